// gik_bounds_init.cu -- bound smoothing and MDS initialisation, one CTA per goal.
//
//   bound_smoothing            utils/dgp.py:192-231
//   generate_initialization    solvers/riemannian_solver.py:67-75
//     gram_from_distance_matrix  utils/dgp.py:28-31
//     factor / MDS               utils/dgp.py:150-171
//     linear_projection          utils/dgp.py:174-183
//
// bound_smoothing.  The reference runs all-pairs Bellman-Ford on a 2N-node digraph
// (two copies of the graph joined by -LOWER arcs).  A shortest u -> v' path uses
// exactly one joining arc, so with Up = min-plus closure of the UPPER matrix
//   upper[u,v] = Up[u,v]
//   lower[u,v] = max(0, max_{a,b} (L[a,b] - Up[u,a] - Up[b,v]))
// i.e. one Floyd-Warshall on an N x N matrix in shared memory plus two max-plus
// products evaluated row by row.  Only the 2 * n_anchor goal edges differ
// between goals; they are patched in from goal_d2.
//
// Initialisation.  Three symmetric eigenproblems per goal (Gram matrix with
// vectors; the reference's rank heuristic -- eigenvalues of the matrix numpy's eigh
// reads from the LOWER triangle of the non-symmetric factor; the K x K scatter matrix
// of the linear projection): CTA-parallel Householder tridiagonalisation in shared memory,
// then implicit QL with the rotation lists applied row-parallel (vectors) or a Sturm count
// (rank heuristic), then the reflectors applied to the eigenvectors that are used.
// Round 1's cyclic Jacobi solver did ~10x the arithmetic and was bound by shared-memory bandwidth.
#include <cstdio>
#include <cstdlib>

#include "gik_common.cuh"

namespace {

struct BiArgs {
    int N, n_goal, n_goal_edges, n_omega_edges, goal_p, goal_q;
    const double *bs_lower, *bs_upper;
    const int32_t *low_ptr, *low_row;       // positive static LOWER entries by column (pairs with p_n / q_n excluded)
    const double *low_val;
    const int32_t *goal_edge_i, *goal_edge_j, *goal_edge_slot;
    const int32_t *omega_ptr, *omega_adj;   // CSR of omega (both directions)
    const double *goal_d2;   // [B][n_goal] (bounds from goals) or null
    const double *lb_in, *ub_in;  // [B][N][N] (init from given bounds) or null
    int B;
    double *lb_out, *ub_out;  // [B][N][N] or null
    double *Y_init;           // [B][N][3] or null
    double *scratch;          // caller's workspace when the matrices do not fit in shared memory
    int use_scratch;
    int do_bounds;            // 1: bounds from the plan tables (+ goal_d2 patches), 0: bounds given
};

// Sum over the CTA; every thread gets the same bits (warp butterflies, then the warp totals in index order).
// One barrier: the totals alternate between two buffers (`par`), so a buffer is rewritten only after a
// later barrier has been passed by every reader.  red: 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *red, int &par)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(GIK_FULL_MASK, v, off);
    if (nw == 1) { __syncwarp(); return v; }
    double *r = red + par * 16;
    par ^= 1;
    if (lane == 0) r[wid] = v;
    __syncthreads();
    double t = r[0];
#pragma unroll 1
    for (int w = 1; w < nw; ++w) t += r[w];
    return t;
}

#ifdef GIK_BI_PROFILE
#define BI_TICK(name) do { __syncthreads(); if (blockIdx.x == 0 && threadIdx.x == 0 && b == 0) { long long t1 = clock64(); printf("bi_profile %-10s %10lld cycles\n", name, t1 - t0); t0 = clock64(); } } while (0)
#else
#define BI_TICK(name)
#endif

// max / min of operands that are never NaN (fmax / fmin carry NaN handling that triples the code of the tile loops;
// the kernel's code size matters: CTAs in different phases share the instruction cache)
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

__device__ __forceinline__ int pow2_at_least(int n, int cap)
{
    int t = 1;
    while (t < n && t < cap) t <<= 1;
    return t;
}

// Householder reduction of the symmetric n x n matrix A (leading dimension ld) to tridiagonal form, in place:
// d[0..n) diagonal, e[i] couples i and i + 1.  Reflector k = I - tau[k] v v^T acts on rows k+1..n-1; v stays in
// column k of A below the diagonal (tau[k] = 0: identity).  v: n doubles, part: 3n doubles, red: 32 doubles of shared
// memory.  Thread maps: matrix-vector product thread = (column, slice of the contraction), partial sums added in
// slice order; rank-2 update thread = (column, row group).
__device__ __noinline__ void tridiagonalize(double *A, int n, int ld, double *d, double *e, double *tau, double *v, double *part,
                               double *red, int &par)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int TJ = pow2_at_least(n, nt), GF = nt / TJ, G = GF < 3 ? GF : 3;
    const int tj = tid & (TJ - 1), tg = tid / TJ;
    for (int k = 0; k + 2 < n; ++k) {
        const int m = n - k - 1;                 // size of the trailing block A22 = A[k+1.., k+1..]
        double *A22 = A + (k + 1) * ld + (k + 1);
        double *col = A + (k + 1) * ld + k;      // x_i = col[i * ld]
        const double x0 = col[0];
        double ps = 0.0;
        for (int i = 1 + tid; i < m; i += nt) { const double xi = col[i * ld]; ps = fma(xi, xi, ps); }
        const double tail = block_sum(ps, red, par);      // ||x[1:]||^2
        if (!(tail > 0.0)) {                      // already tridiagonal in this column (same bits on every thread)
            if (tid == 0) { d[k] = A[k * ld + k]; e[k] = x0; tau[k] = 0.0; }
            continue;
        }
        // v = x - alpha e_0, alpha = -sign(x0) ||x||, beta = 2 / v^T v = 1 / (||x||^2 + |x0| ||x||)
        const double xx = fma(x0, x0, tail);
        const double nrm = xx * rsqrt(xx);
        const double alpha = x0 > 0.0 ? -nrm : nrm;
        const double v0 = x0 - alpha;
        const double beta = gik_rcp(fma(fabs(x0), nrm, xx));
        for (int i = tid; i < m; i += nt) v[i] = i == 0 ? v0 : col[i * ld];
        __syncthreads();                          // every thread has read x0
        if (tid == 0) { col[0] = v0; d[k] = A[k * ld + k]; e[k] = alpha; tau[k] = beta; }
        if (tg < G && tj < m) {                   // m <= TJ: one column per thread
            const double *ap = A22 + tj;          // A22 symmetric: row j, lanes over the column
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j = tg;
            for (; j + 3 * G < m; j += 4 * G) {
                a0 = fma(ap[j * ld], v[j], a0);
                a1 = fma(ap[(j + G) * ld], v[j + G], a1);
                a2 = fma(ap[(j + 2 * G) * ld], v[j + 2 * G], a2);
                a3 = fma(ap[(j + 3 * G) * ld], v[j + 3 * G], a3);
            }
            for (; j < m; j += G) a0 = fma(ap[j * ld], v[j], a0);
            part[tg * n + tj] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        double pi = 0.0, vi = 0.0;
        if (tid < m) {
            pi = part[tid];
            for (int gsel = 1; gsel < G; ++gsel) pi += part[gsel * n + tid];
            pi *= beta;
            vi = v[tid];
        }
        const double pTv = block_sum(pi * vi, red, par);
        const double Kc = 0.5 * beta * pTv;
        if (tid < m) part[tid] = pi - Kc * vi;    // w = p - K v  (every thread has read `part`: barrier in block_sum)
        __syncthreads();
        if (tj < m) {                             // A22 <- A22 - v w^T - w v^T, thread = (column, row group)
            const double wj = part[tj], vj = v[tj];
            double *ap = A22 + tj;
            int i = tg;
            for (; i + 3 * GF < m; i += 4 * GF) {
                const double v0r = v[i], v1r = v[i + GF], v2r = v[i + 2 * GF], v3r = v[i + 3 * GF];
                const double w0r = part[i], w1r = part[i + GF], w2r = part[i + 2 * GF], w3r = part[i + 3 * GF];
                const double b0 = ap[i * ld], b1 = ap[(i + GF) * ld], b2 = ap[(i + 2 * GF) * ld], b3 = ap[(i + 3 * GF) * ld];
                ap[i * ld] = b0 - fma(v0r, wj, w0r * vj);
                ap[(i + GF) * ld] = b1 - fma(v1r, wj, w1r * vj);
                ap[(i + 2 * GF) * ld] = b2 - fma(v2r, wj, w2r * vj);
                ap[(i + 3 * GF) * ld] = b3 - fma(v3r, wj, w3r * vj);
            }
            for (; i < m; i += GF) ap[i * ld] -= fma(v[i], wj, part[i] * vj);
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (n >= 2) { d[n - 2] = A[(n - 2) * ld + (n - 2)]; e[n - 2] = A[(n - 1) * ld + (n - 2)]; }
        d[n - 1] = A[(n - 1) * ld + (n - 1)];
    }
    __syncthreads();
}

// Implicit QL iteration (EISPACK tql2) on the tridiagonal (d, e): on exit d holds the eigenvalues (unordered) and
// W <- W Q, Q = eigenvector matrix of the tridiagonal (W n x n, leading dimension ldw; pass the identity to get Q).
// The scalar recurrence runs on the CTA's last thread; the other threads apply the previous sweep's rotations, one
// thread per row of W, while it computes the next sweep (two rotation lists, one barrier per sweep).
// rot: 2 x 2n doubles, meta: 4 ints of shared memory.
__device__ __noinline__ void ql_implicit(double *d, double *e, int n, double *W, int ldw, double *rot, int *meta)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const bool scalar = tid == nt - 1;
    int l = 0, iter = 0, buf = 0;
    bool done = false, fresh = true;
    double tst1 = 0.0;
    if (scalar) { e[n - 1] = 0.0; meta[2] = 0; meta[3] = 0; }
    __syncthreads();
#ifdef GIK_BI_PROFILE
    long long t_scalar = 0, t_apply = 0, t_search = 0; int n_sweeps = 0, n_rot = 0;
#endif
    for (;;) {
#ifdef GIK_BI_PROFILE
        long long tq0 = clock64();
#endif
        if (scalar && !done) {
            double *r2 = rot + buf * 2 * n;
            int cnt = 0, mtop = 0;
            for (;;) {
                if (l >= n) { done = true; cnt = -1; break; }
                if (fresh) { tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l])); iter = 0; fresh = false; }
#ifdef GIK_BI_PROFILE
                long long ts0 = clock64();
#endif
                // first negligible off-diagonal element at or after l (e[n - 1] = 0 ends the scan).  |e| > small compared on
                // the bit patterns (both sides are non-negative, finite): integer compares cost a fraction of DSETP's latency
                int m = l;
                {
                    const long long small = __double_as_longlong(2.220446049250313e-16 * tst1);
                    const long long *eb = reinterpret_cast<const long long *>(e);
                    const long long absmask = 0x7fffffffffffffffLL;
                    for (;; m += 4) {
                        const long long e0 = eb[min(m, n - 1)] & absmask, e1 = eb[min(m + 1, n - 1)] & absmask;
                        const long long e2 = eb[min(m + 2, n - 1)] & absmask, e3 = eb[min(m + 3, n - 1)] & absmask;
                        const int hit = e0 <= small ? 0 : (e1 <= small ? 1 : (e2 <= small ? 2 : (e3 <= small ? 3 : 4)));
                        if (hit < 4) { m += hit; break; }
                    }
                }
#ifdef GIK_BI_PROFILE
                t_search += clock64() - ts0;
#endif
                if (m == l || ++iter > 60) { ++l; fresh = true; continue; }
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = sqrt(fma(g, g, 1.0));
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? r : -r));
                double s = 1.0, c = 1.0, p = 0.0;
                bool under = false;
                double ei = e[m - 1], di = d[m - 1], di1 = d[m];
                // tql2's recurrence with the second rotation product taken off the critical path:
                // r = (d_i - g1) s' + 2 c' b = rinv (t1 f + 2 b g), so only c', r and g' hang on the reciprocal root
                for (int i = m - 1; i >= l; --i) {
                    const double f = s * ei, bq = c * ei;
                    const double h = fma(f, f, g * g);
                    const double g1 = di1 - p, t1 = di - g1;
                    const double qn = fma(t1, f, (bq + bq) * g);
                    const double rinv = rsqrt(h);
                    const double cn = g * rinv;
                    const double r = qn * rinv;
                    const double gn = fma(cn, r, -bq);
                    if (h == 0.0) { e[i + 1] = 0.0; d[i + 1] = di1 - p; e[m] = 0.0; under = true; break; }
                    c = cn;
                    g = gn;
                    s = f * rinv;
                    p = s * r;
                    e[i + 1] = h * rinv;
                    d[i + 1] = g1 + p;
                    r2[2 * cnt] = c;
                    r2[2 * cnt + 1] = s;
                    ++cnt;
                    di1 = di;
                    if (i > l) { ei = e[i - 1]; di = d[i - 1]; }
                }
                if (!under) { d[l] = di1 - p; e[l] = g; e[m] = 0.0; }
                mtop = m;
                if (cnt > 0) break;
            }
            meta[buf * 2] = cnt;
            meta[buf * 2 + 1] = mtop;
#ifdef GIK_BI_PROFILE
            t_scalar += clock64() - tq0; ++n_sweeps; n_rot += cnt > 0 ? cnt : 0;
#endif
        }
        const int cnt = meta[(buf ^ 1) * 2], mtop = meta[(buf ^ 1) * 2 + 1];
        if (cnt < 0) break;
        if (cnt > 0 && tid < n) {
            const double *r2 = rot + (buf ^ 1) * 2 * n;
            double *row = W + tid * ldw;
            double f = row[mtop];
            for (int q = 0; q < cnt; ++q) {
                const double c = r2[2 * q], s = r2[2 * q + 1];
                const double zi = row[mtop - 1 - q];
                row[mtop - q] = fma(s, zi, c * f);
                f = fma(c, zi, -s * f);
            }
            row[mtop - cnt] = f;
        }
#ifdef GIK_BI_PROFILE
        if (tid == 0) t_apply += clock64() - tq0;
#endif
        __syncthreads();
        buf ^= 1;
    }
#ifdef GIK_BI_PROFILE
    if (blockIdx.x == 0 && scalar) printf("bi_profile   ql n %d sweeps %d rotations %d scalar %lld (search %lld) cycles\n", n, n_sweeps, n_rot, t_scalar, t_search);
    if (blockIdx.x == 0 && tid == 0) printf("bi_profile   ql apply (thread 0) %lld cycles\n", t_apply);
#endif
    __syncthreads();
}

// W[:, c] <- H_0 H_1 ... H_{n-3} W[:, c] for the columns c = cols[0..ncols): eigenvectors of the tridiagonal ->
// eigenvectors of the matrix `tridiagonalize` reduced.  No CTA barriers: one thread per column when there are many
// columns, one warp per column (lanes over the rows, butterfly for the dot product) when there are few.
template <int ROWS>   // rows per lane of the warp-per-column variant: n <= 32 ROWS + 1
__device__ __noinline__ void apply_reflectors(const double *A, int n, int ld, const double *tau, double *W, int ldw, const int *cols,
                                 int ncols)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int rounds = (ncols + nw - 1) / nw;
    if (rounds * 16 < n) {
        for (int q = wid; q < ncols; q += nw) {
            const int c = cols[q];
            for (int k = n - 3; k >= 0; --k) {
                const double beta = tau[k];
                if (beta == 0.0) continue;
                const int m = n - k - 1;
                const double *v = A + (k + 1) * ld + k;
                double *w = W + (k + 1) * ldw + c;
                double vr[ROWS], wr[ROWS], t = 0.0;
#pragma unroll
                for (int u = 0; u < ROWS; ++u) {
                    if (ROWS > 4 && 32 * u >= m) break;
                    const int j = lane + 32 * u;
                    vr[u] = j < m ? v[j * ld] : 0.0;
                    wr[u] = j < m ? w[j * ldw] : 0.0;
                    t = fma(vr[u], wr[u], t);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(GIK_FULL_MASK, t, off);
                t *= -beta;
#pragma unroll
                for (int u = 0; u < ROWS; ++u) {
                    if (ROWS > 4 && 32 * u >= m) break;
                    const int j = lane + 32 * u;
                    if (j < m) w[j * ldw] = fma(t, vr[u], wr[u]);
                }
                __syncwarp();
            }
        }
        return;
    }
    for (int q = tid; q < ncols; q += nt) {
        const int c = cols[q];
        for (int k = n - 3; k >= 0; --k) {
            const double beta = tau[k];
            if (beta == 0.0) continue;
            const int m = n - k - 1;
            const double *v = A + (k + 1) * ld + k;
            double *w = W + (k + 1) * ldw + c;
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            int j = 0;
            for (; j + 3 < m; j += 4) {
                t0 = fma(v[j * ld], w[j * ldw], t0);
                t1 = fma(v[(j + 1) * ld], w[(j + 1) * ldw], t1);
                t2 = fma(v[(j + 2) * ld], w[(j + 2) * ldw], t2);
                t3 = fma(v[(j + 3) * ld], w[(j + 3) * ldw], t3);
            }
#pragma unroll 1
            for (; j < m; ++j) t0 = fma(v[j * ld], w[j * ldw], t0);
            const double t = -beta * ((t0 + t1) + (t2 + t3));
            for (j = 0; j + 3 < m; j += 4) {
                const double w0 = w[j * ldw], w1 = w[(j + 1) * ldw], w2 = w[(j + 2) * ldw], w3 = w[(j + 3) * ldw];
                const double v0 = v[j * ld], v1 = v[(j + 1) * ld], v2 = v[(j + 2) * ld], v3 = v[(j + 3) * ld];
                w[j * ldw] = fma(t, v0, w0);
                w[(j + 1) * ldw] = fma(t, v1, w1);
                w[(j + 2) * ldw] = fma(t, v2, w2);
                w[(j + 3) * ldw] = fma(t, v3, w3);
            }
#pragma unroll 1
            for (; j < m; ++j) w[j * ldw] = fma(t, v[j * ld], w[j * ldw]);
        }
    }
}

// rank[i] = position of d[i] in descending order (ties: lower index first); order[rank[i]] = i
__device__ void rank_desc(const double *d, int n, int *rank, int *order)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double di = d[i];
        int r = 0;
#pragma unroll 2
        for (int j = 0; j < n; ++j) r += (d[j] > di) || (d[j] == di && j < i);
        rank[i] = r;
        order[r] = i;
    }
    __syncthreads();
}

// Number of eigenvalues > sigma of the symmetric tridiagonal (d, e): Sturm count (signs of the pivots of T - sigma I).
__device__ int sturm_count_above(const double *d, const double *e, int n, double sigma)
{
    int below = 0;
    double q = d[0] - sigma;
    if (q < 0.0) ++below;
    for (int i = 1; i < n; ++i) {
        if (q == 0.0) q = 1e-300;
        q = d[i] - sigma - e[i - 1] * e[i - 1] / q;
        if (q < 0.0) ++below;
    }
    return n - below;
}

// TS = 4 x 4 register tiles per thread (TS * blockDim.x >= ceil(N / 4)^2; 0: none, looped fallback for N > 128),
// MAXT / MINB = launch bounds: the kernel is
// latency bound (one thread runs the QL recurrence while the CTA waits), so resident CTAs per SM are what counts.
template <int TS, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_bounds_init(const BiArgs a)
{
    extern __shared__ double smem[];
    const int N = a.N, NN = N * N, tid = threadIdx.x, nt = blockDim.x;
    // small arrays first (sizes: gik_bi_small_doubles in gik_common.cuh)
    double *red = smem;                // 32
    double *lam = red + 32;            // N   eigenvalues / row means
    double *sc = lam + N;              // 7N + 16, by phase:
    double *Lp = sc, *Lq = sc + N;     //   bound smoothing: LOWER of the pairs (., p_n) / (., q_n); then two rows of
    double *rowbuf = sc + 2 * N;       //   4 ceil(N / 4) doubles for Floyd-Warshall
    double *dd = sc, *ee = sc + N, *tau = sc + 2 * N, *hv = sc + 3 * N, *part = sc + 4 * N;   // eigen: d, e, tau, v, 3N
    double *rot = sc + 3 * N;          //   QL rotation lists (2 x 2N) reuse v and the partial sums
    int *order = reinterpret_cast<int *>(sc + 7 * N + 16);   // N ints
    int *rank = order + N;             // N ints
    int *meta = rank + N;              // 4 ints
    double *mats = smem + gik_bi_small_doubles(N);
    double *M1, *M2, *M3;
    if (a.use_scratch == 2) {          // nothing fits: all three matrices in the caller's workspace
        M1 = a.scratch + (size_t)blockIdx.x * 3 * NN;
        M2 = M1 + NN;
        M3 = M2 + NN;
    } else {
        M1 = mats;
        M2 = M1 + NN;
        M3 = a.use_scratch == 1 ? a.scratch + (size_t)blockIdx.x * NN : M2 + NN;
    }
    int par = 0;
    // 4 x 4 register tiles of an N x N matrix: tile t covers rows 4 (t / NT4).., columns 4 (t % NT4)..; a thread owns
    // tiles tid, tid + nt, ... (the launch guarantees TS nt >= NT4^2)
    const int NT4 = (N + 3) >> 2, ntiles = NT4 * NT4, RB = 4 * NT4;
    const int TJ = pow2_at_least(N, nt), G = nt / TJ;   // thread = (column, group)
    const int tj = tid & (TJ - 1), tg = tid / TJ;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
#ifdef GIK_BI_PROFILE
        long long t0 = clock64();
#endif
        double *Up = M1, *D = M2;
        if (a.do_bounds) {
            // ---------------- bound smoothing
            const double *gd = a.goal_d2 ? a.goal_d2 + (size_t)b * a.n_goal : nullptr;
            for (int k = tid; k < NN; k += nt) Up[k] = a.bs_upper[k];
            __syncthreads();
            for (int e = tid; e < a.n_goal_edges; e += nt) {
                const int i = a.goal_edge_i[e], j = a.goal_edge_j[e];
                const double dist = sqrt(gd[a.goal_edge_slot[e]]);
                Up[i * N + j] = dist;
                Up[j * N + i] = dist;
            }
            for (int j = tid; j < 2 * RB; j += nt) rowbuf[j] = INFINITY;
            __syncthreads();
            // min-plus closure (Floyd-Warshall) with the matrix in register tiles.  Step k needs row k only (the matrix
            // stays symmetric bit for bit: u_ik + u_kj and u_jk + u_ki add the same two numbers), which its owners
            // publish after step k - 1; row k is a fixed point of step k.
            if (TS == 0) {
                // graphs too large for register tiles (N > 128): the plain in-place iteration, thread = (column, row group)
                for (int k = 0; k < N; ++k) {
                    for (int i = tg; i < N; i += G) {
                        const double uik = Up[i * N + k];
                        for (int j = tj; j < N; j += TJ) Up[i * N + j] = dmin(Up[i * N + j], uik + Up[k * N + j]);
                    }
                    __syncthreads();
                }
            } else {
                constexpr int TSA = TS > 0 ? TS : 1;
                double U[TSA][4][4];
                int i0[TSA], j0[TSA];
                bool valid[TSA];
#pragma unroll
                for (int ts = 0; ts < TS; ++ts) {
                    const int t = tid + ts * nt;
                    valid[ts] = t < ntiles;
                    i0[ts] = (t / NT4) * 4;
                    j0[ts] = (t % NT4) * 4;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            U[ts][r][c] = (valid[ts] && i0[ts] + r < N && j0[ts] + c < N) ? Up[(i0[ts] + r) * N + j0[ts] + c] : INFINITY;
                }
                for (int j = tid; j < N; j += nt) rowbuf[j] = Up[j];
                __syncthreads();
                for (int k = 0; k < N; ++k) {
                    const double *rb = rowbuf + (k & 1) * RB;
                    double *wb = rowbuf + ((k + 1) & 1) * RB;
#pragma unroll
                    for (int ts = 0; ts < TS; ++ts) {
                        if (!valid[ts]) continue;
                        double ci[4], cj[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) { ci[r] = rb[i0[ts] + r]; cj[r] = rb[j0[ts] + r]; }
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int c = 0; c < 4; ++c) U[ts][r][c] = dmin(U[ts][r][c], ci[r] + cj[c]);
                        const int rn = k + 1 - i0[ts];
                        if (rn >= 0 && rn < 4 && k + 1 < N) {
#pragma unroll
                            for (int r = 0; r < 4; ++r)
                                if (r == rn) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c) wb[j0[ts] + c] = U[ts][r][c];
                                }
                        }
                    }
                    __syncthreads();
                }
#pragma unroll
                for (int ts = 0; ts < TS; ++ts) {
                    if (!valid[ts]) continue;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (i0[ts] + r < N && j0[ts] + c < N) Up[(i0[ts] + r) * N + j0[ts] + c] = U[ts][r][c];
                }
            }
            __syncthreads();
            BI_TICK("floyd");
            // LOWER of this goal: the static table, except the pairs with p_n / q_n (rows Lp, Lq), which hold
            // the goal's exact distances on the goal edges (their static entries are 0)
            const int gp = a.goal_p, gq = a.goal_q;
            for (int i = tid; i < N; i += nt) {
                Lp[i] = gp >= 0 ? a.bs_lower[i * N + gp] : 0.0;
                Lq[i] = gq >= 0 ? a.bs_lower[i * N + gq] : 0.0;
            }
            __syncthreads();
            for (int e = tid; e < a.n_goal_edges; e += nt) {
                const int i = a.goal_edge_i[e], j = a.goal_edge_j[e];
                const double l = sqrt(gd[a.goal_edge_slot[e]]);
                if (j == gp) Lp[i] = l; else if (i == gp) Lp[j] = l;
                if (j == gq) Lq[i] = l; else if (i == gq) Lq[j] = l;
            }
            __syncthreads();
            // lower bounds, two max-plus products (max is exact, so any evaluation order gives the same bits):
            //   Mt[b][u] = max(-Up[u,b], max_{a: L[a,b] > 0} (L[a,b] - Up[u,a]));  lower[u,v] = max(0, max_b (Mt[b][u] - Up[b,v]))
            // first product from the sparse column lists of L (thread = (u, column group), Up read through its symmetry)
            double *Mt = M3;
            for (int bb = tg; bb < N; bb += G) {
                for (int u = tj; u < N; u += TJ) {
                    double m = -Up[bb * N + u];   // joining arc b -> b' of weight 0
                    if (bb == gp) {
#pragma unroll 1
                        for (int aa = 0; aa < N; ++aa) { const double l = Lp[aa]; if (l > 0.0) m = dmax(m, l - Up[aa * N + u]); }
                    } else if (bb == gq) {
#pragma unroll 1
                        for (int aa = 0; aa < N; ++aa) { const double l = Lq[aa]; if (l > 0.0) m = dmax(m, l - Up[aa * N + u]); }
                    } else {
                        const int e0 = a.low_ptr[bb], e1 = a.low_ptr[bb + 1];
#pragma unroll 2
                        for (int e = e0; e < e1; ++e) m = dmax(m, a.low_val[e] - Up[a.low_row[e] * N + u]);
                        if (gp >= 0) { const double l = Lp[bb]; if (l > 0.0) m = dmax(m, l - Up[gp * N + u]); }
                        if (gq >= 0) { const double l = Lq[bb]; if (l > 0.0) m = dmax(m, l - Up[gq * N + u]); }
                    }
                    Mt[bb * N + u] = m;
                }
            }
            __syncthreads();
            // second product in 4 x 4 register tiles (edge tiles compute on clamped indices and drop the duplicates)
            const int tile_rounds = TS > 0 ? TS : (ntiles + nt - 1) / nt;   // (no accumulator outlives a tile here)
#pragma unroll 1
            for (int ts = 0; ts < tile_rounds; ++ts) {
                const int t = tid + ts * nt;
                if (t >= ntiles) break;
                const int u0 = (t / NT4) * 4, v0 = (t % NT4) * 4;
                int uo[4], vo[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) { uo[r] = min(u0 + r, N - 1); vo[r] = min(v0 + r, N - 1); }
                double acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
                for (int bb = 0; bb < N; ++bb) {
                    double mu[4], up[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { mu[r] = Mt[bb * N + uo[r]]; up[r] = Up[bb * N + vo[r]]; }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][c] = dmax(acc[r][c], mu[r] - up[c]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int u = u0 + r, v = v0 + c;
                        if (u < N && v < N) {
                            const double lo = (u == v) ? 0.0 : acc[r][c];
                            const double up = Up[u * N + v];
                            if (a.lb_out) a.lb_out[(size_t)b * NN + u * N + v] = lo;
                            if (a.ub_out) a.ub_out[(size_t)b * NN + u * N + v] = up;
                            const double dr = lo + 0.9 * (up - lo);   // riemannian_solver.py:72
                            D[u * N + v] = dr * dr;
                        }
                    }
            }
            __syncthreads();
        } else {
            const double *lb = a.lb_in + (size_t)b * NN, *ub = a.ub_in + (size_t)b * NN;
            for (int k = tid; k < NN; k += nt) {
                const double dr = lb[k] + 0.9 * (ub[k] - lb[k]);
                D[k] = dr * dr;
            }
            __syncthreads();
        }
        BI_TICK("lower");
        if (!a.Y_init) continue;

        // ---------------- Gram matrix B = -1/2 J D J  (dgp.py:28-31), in place in D
        double *cmean = sc;
        for (int i = tid; i < N; i += nt) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s += D[i * N + j];
            lam[i] = s / N;   // row means
        }
        for (int j = tid; j < N; j += nt) {
            double s = 0.0;
            for (int i = 0; i < N; ++i) s += D[i * N + j];
            cmean[j] = s / N;  // column means
        }
        __syncthreads();
        double tot = 0.0;
        for (int i = tid; i < N; i += nt) tot += lam[i];
        tot = block_sum(tot, red, par) / N;
        double *G = M2, *V = M1;
        for (int k = tid; k < NN; k += nt) {
            const int i = k / N, j = k % N;
            G[k] = -0.5 * (D[k] - lam[i] - cmean[j] + tot);
        }
        __syncthreads();
        // symmetrise against rounding (D is symmetric up to the evaluation order of the max-plus products)
        for (int k = tid; k < NN; k += nt) {
            const int i = k / N, j = k % N;
            if (i < j) {
                const double m = 0.5 * (G[i * N + j] + G[j * N + i]);
                G[i * N + j] = m;
                G[j * N + i] = m;
            }
            V[k] = i == j ? 1.0 : 0.0;
        }
        __syncthreads();
        // ---------------- factor (dgp.py:150-159): X = V sqrt(max(lambda,0)), columns by descending lambda
        BI_TICK("gram");
        tridiagonalize(G, N, N, dd, ee, tau, hv, part, red, par);
        BI_TICK("tridiag");
        ql_implicit(dd, ee, N, V, N, rot, meta);
        BI_TICK("ql");
        // An eigenvalue below 1e-13 of the largest one is rounding noise of an exactly singular Gram matrix (J D J
        // always annihilates the vector of ones; coincident nodes add more).  Its sign is arbitrary -- LAPACK's,
        // Jacobi's and QL's noise differ -- yet a "positive" one would enter the factor as a column of size ~ 3e-8
        // and can lift the rank count below over its 1e-8 threshold.  Such eigenvalues count as zero here.
        double lmax = 0.0;
#pragma unroll 1
        for (int i = 0; i < N; ++i) lmax = dmax(lmax, dd[i]);
        const double lcut = 1e-13 * lmax;
        int npos = 0;                    // the columns that are kept = the first npos of the descending order
#pragma unroll 1
        for (int i = 0; i < N; ++i) npos += dd[i] > lcut;
        rank_desc(dd, N, rank, order);
        for (int i = tid; i < N; i += nt) { lam[i] = dd[i]; if (!(dd[i] > lcut)) rank[i] = N; }
        __syncthreads();
        apply_reflectors<(TS == 0 ? 15 : 4)>(G, N, N, tau, V, N, order, npos);
        __syncthreads();
        BI_TICK("backtr");
        // Eigenvector signs are arbitrary, yet the rank heuristic below is NOT invariant to them
        // (it reads a triangle of the non-symmetric factor).  The reference inherits whatever
        // LAPACK returns; here the sign is fixed canonically: the entry of largest magnitude of
        // every eigenvector is positive (first such entry on ties).
        for (int col = tid; col < N; col += nt) {
            if (rank[col] >= N) continue;
            double best = 0.0, sgn = 1.0;
#pragma unroll 2
            for (int i = 0; i < N; ++i) {
                const double v = V[i * N + col];
                if (fabs(v) > best) { best = fabs(v); sgn = v < 0.0 ? -1.0 : 1.0; }
            }
            if (sgn < 0.0)
                for (int i = 0; i < N; ++i) V[i * N + col] = -V[i * N + col];
        }
        __syncthreads();
        double *X = M2;
        for (int k = tid; k < NN; k += nt) {
            const int i = k / N, col = k % N;
            const int src = order[col];
            X[k] = rank[src] < N ? V[i * N + src] * sqrt(lam[src]) : 0.0;
        }
        __syncthreads();
        // ---------------- MDS rank (dgp.py:163-171): eigh of the lower triangle of X, count > 1e-8
        double *Aw = M1;
        for (int k = tid; k < NN; k += nt) {
            const int i = k / N, j = k % N;
            Aw[k] = i >= j ? X[i * N + j] : X[j * N + i];
        }
        __syncthreads();
        BI_TICK("factor");
        // only the COUNT of eigenvalues above 1e-8 is used: tridiagonal reduction + Sturm count
        tridiagonalize(Aw, N, N, dd, ee, tau, hv, part, red, par);
        if (tid == 0) meta[0] = sturm_count_above(dd, ee, N, 1e-8);
        __syncthreads();
        int K = meta[0];
        __syncthreads();
        BI_TICK("eig_rank");
        if (K > N) K = N;
        // ---------------- linear_projection (dgp.py:174-183): S = sum_{omega} (P_i-P_j)(P_i-P_j)^T, P = X[:, :K]
        // evaluated node-wise: W_i = sum_{j ~ i} (P_i - P_j) (N x K, kept where E will go), S = 2 sum_i P_i^T W_i
        // (the K^2 |omega| edge-wise evaluation cost 5 M cycles per goal at N = 118, K = 40), then symmetrised
        double *S = M1, *E = M3, *Wn = M3;
        {
            // thread = (coordinate r, node group); fixed neighbour order -> deterministic sums
            const int TK = pow2_at_least(K, nt), GK = nt / TK;
            const int tr = tid & (TK - 1), tgk = tid / TK;
            for (int i = tgk; i < N; i += GK) {
                const int e0 = a.omega_ptr[i], e1 = a.omega_ptr[i + 1];
                for (int r = tr; r < K; r += TK) {
                    const double xi = X[i * N + r];
                    double acc = 0.0;
#pragma unroll 2
                    for (int e = e0; e < e1; ++e) acc += xi - X[a.omega_adj[e] * N + r];
                    Wn[i * K + r] = acc;
                }
            }
        }
        __syncthreads();
        for (int k = tid; k < K * K; k += nt) {
            const int r = k / K, cidx = k % K;
            double s = 0.0;
            for (int i = 0; i < N; ++i) s = fma(X[i * N + r], Wn[i * K + cidx], s);
            S[r * K + cidx] = 2.0 * s;   // both (i,j) and (j,i) are nonzeros of omega
        }
        __syncthreads();
        for (int k = tid; k < K * K; k += nt) {
            const int r = k / K, cidx = k % K;
            if (r < cidx) {
                const double m = 0.5 * (S[r * K + cidx] + S[cidx * K + r]);
                S[r * K + cidx] = m;
                S[cidx * K + r] = m;
            }
            E[k] = r == cidx ? 1.0 : 0.0;
        }
        __syncthreads();
        BI_TICK("scatter");
        double *Yo = a.Y_init + (size_t)b * N * 3;
        if (K > 0) {
            tridiagonalize(S, K, K, dd, ee, tau, hv, part, red, par);
            ql_implicit(dd, ee, K, E, K, rot, meta);
            rank_desc(dd, K, rank, order);
            apply_reflectors<(TS == 0 ? 15 : 4)>(S, K, K, tau, E, K, order, K < 3 ? K : 3);   // only the three leading eigenvectors are used
            __syncthreads();
        }
        BI_TICK("eig_proj");
        for (int k = tid; k < N * 3; k += nt) {
            const int i = k / 3, cidx = k % 3;
            double s = 0.0;
            if (cidx < K) {
                const int col = order[cidx];
                for (int r = 0; r < K; ++r) s += X[i * N + r] * E[r * K + col];
            }
            Yo[k] = s;
        }
        __syncthreads();
    }
}

size_t small_bytes(int N) { return (size_t)gik_bi_small_doubles(N) * sizeof(double); }

int launch(const GikPlan *p, BiArgs &a, void *workspace, cudaStream_t st)
{
    const int N = p->N;
    a.N = N;
    a.n_goal = p->n_goal;
    a.n_goal_edges = p->n_goal_edges;
    a.n_omega_edges = p->n_omega_edges;
    a.goal_p = p->n_goal_edges > 0 ? p->goal_p : -1;
    a.goal_q = p->n_goal_edges > 0 ? p->goal_q : -1;
    a.bs_lower = p->bs_lower;
    a.bs_upper = p->bs_upper;
    a.goal_edge_i = p->goal_edge_i;
    a.goal_edge_j = p->goal_edge_j;
    a.goal_edge_slot = p->goal_edge_slot;
    a.low_ptr = p->low_ptr;
    a.low_row = p->low_row;
    a.low_val = p->low_val;
    a.omega_ptr = p->omega_ptr;
    a.omega_adj = p->omega_adj;
    const size_t mat = (size_t)N * N * sizeof(double);
    size_t smem = small_bytes(N) + (p->bi_mode == 0 ? 3 : (p->bi_mode == 1 ? 2 : 0)) * mat;
    const int variant = gik_bi_variant(N);
    const int threads = variant == 0 ? 32 : (variant >= 3 ? 512 : 128);
    int blocks = a.B;
    a.use_scratch = p->bi_mode;
    a.scratch = static_cast<double *>(workspace);
    if (p->bi_mode && !workspace) {
        gik_set_error("gik_bounds/gik_init/gik_bounds_init: a plan with N=%d needs a workspace of gik_workspace_bytes() "
                      "bytes (one per concurrently running call)", N);
        return GIK_EINVAL;
    }
    void (*kern)(const BiArgs) = variant == 0 ? k_bounds_init<1, 32, 24> : (variant == 1 ? k_bounds_init<1, 128, 6> :
                                 (variant == 2 ? k_bounds_init<2, 128, 4> : (variant == 3 ? k_bounds_init<2, 512, 1> :
                                                                              k_bounds_init<0, 512, 1>)));
    // shared-memory opt-in and occupancy depend on (device, N) only: looked up once per plan geometry
    static int cached_dev = -1, cached_N = -1, cached_mode = -1, cached_cap = 0;
    if (cached_dev != p->device || cached_N != N || cached_mode != p->bi_mode) {
        GIK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        cached_cap = p->sm_count * (per_sm < 1 ? 1 : per_sm);
        cached_dev = p->device;
        cached_N = N;
        cached_mode = p->bi_mode;
    }
    const int cap = p->bi_mode == 0 ? cached_cap : (cached_cap < p->bi_blocks ? cached_cap : p->bi_blocks);
    if (blocks > cap) blocks = cap;
    kern<<<blocks, threads, smem, st>>>(a);
    return gik_check_cuda(cudaGetLastError(), "k_bounds_init launch");
}

int check_device(const GikPlan *p, const char *fn)
{
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != p->device) {
        gik_set_error("%s: plan belongs to device %d but device %d is current", fn, p->device, dev);
        return GIK_EINVAL;
    }
    return GIK_OK;
}

}  // namespace

extern "C" int gik_bounds(const GikPlan *p, const double *goal_d2, int32_t B, double *lb, double *ub, void *workspace,
                          void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !lb || !ub || B < 0 || (!goal_d2 && p->n_goal_edges > 0)) { gik_set_error("gik_bounds: bad argument"); return GIK_EINVAL; }
    if (!p->bs_lower || !p->bs_upper) { gik_set_error("gik_bounds: plan was created without bound tables"); return GIK_EINVAL; }
    if (int rc = check_device(p, "gik_bounds")) return rc;
    if (B == 0) return GIK_OK;
    BiArgs a = {};
    a.goal_d2 = goal_d2;
    a.do_bounds = 1;
    a.B = B;
    a.lb_out = lb;
    a.ub_out = ub;
    return launch(p, a, workspace, (cudaStream_t)stream);
}

extern "C" int gik_init(const GikPlan *p, const double *lb, const double *ub, int32_t B, double *Y_init, void *workspace,
                        void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !lb || !ub || !Y_init || B < 0) { gik_set_error("gik_init: bad argument"); return GIK_EINVAL; }
    if (int rc = check_device(p, "gik_init")) return rc;
    if (B == 0) return GIK_OK;
    BiArgs a = {};
    a.lb_in = lb;
    a.ub_in = ub;
    a.B = B;
    a.Y_init = Y_init;
    return launch(p, a, workspace, (cudaStream_t)stream);
}

extern "C" int gik_bounds_init(const GikPlan *p, const double *goal_d2, int32_t B, double *Y_init, void *workspace,
                               void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !goal_d2 || !Y_init || B < 0) { gik_set_error("gik_bounds_init: bad argument"); return GIK_EINVAL; }
    if (!p->bs_lower || !p->bs_upper) { gik_set_error("gik_bounds_init: plan was created without bound tables"); return GIK_EINVAL; }
    if (int rc = check_device(p, "gik_bounds_init")) return rc;
    if (B == 0) return GIK_OK;
    BiArgs a = {};
    a.goal_d2 = goal_d2;
    a.do_bounds = 1;
    a.B = B;
    a.Y_init = Y_init;
    return launch(p, a, workspace, (cudaStream_t)stream);
}

// Bytes of the per-call workspace of gik_bounds / gik_init / gik_bounds_init: 0 while the three N x N matrices of
// a goal stay in shared memory, else the matrices every resident CTA keeps in L2 (see gik_plan.cu: bi_mode).
extern "C" int64_t gik_workspace_bytes(const GikPlan *p)
{
    if (!p || !p->bi_mode) return 0;
    return (int64_t)p->bi_blocks * (p->bi_mode == 1 ? 1 : 3) * p->N * p->N * (int64_t)sizeof(double);
}
