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
// of the linear projection) are solved by a CTA-parallel cyclic Jacobi iteration
// (round-robin pair ordering, N/2 disjoint rotations per step) in shared memory.
#include <cstdio>
#include <cstdlib>

#include "gik_common.cuh"

namespace {

struct BiArgs {
    int N, n_goal, n_goal_edges, n_omega_edges, goal_p, goal_q;
    const double *bs_lower, *bs_upper;
    const int32_t *goal_edge_i, *goal_edge_j, *goal_edge_slot;
    const int32_t *omega_i, *omega_j;
    const int32_t *omega_ptr, *omega_adj;   // CSR of omega (both directions)
    const double *goal_d2;   // [B][n_goal] (bounds from goals) or null
    const double *lb_in, *ub_in;  // [B][N][N] (init from given bounds) or null
    int B;
    double *lb_out, *ub_out;  // [B][N][N] or null
    double *Y_init;           // [B][N][3] or null
    double *scratch;          // global scratch when the matrices do not fit in shared memory
    int use_scratch;
    int do_bounds;            // 1: bounds from the plan tables (+ goal_d2 patches), 0: bounds given
};

__device__ __forceinline__ double block_sum(double v, double *red)
{
    // red: >= 33 doubles of shared memory
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(GIK_FULL_MASK, v, off);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < nw ? red[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(GIK_FULL_MASK, t, off);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

#ifdef GIK_BI_PROFILE
#define BI_TICK(name) do { __syncthreads(); if (blockIdx.x == 0 && threadIdx.x == 0 && b == 0) { long long t1 = clock64(); printf("bi_profile %-10s %10lld cycles\n", name, t1 - t0); t0 = clock64(); } } while (0)
#else
#define BI_TICK(name)
#endif

__device__ __forceinline__ int pow2_at_least(int n, int cap)
{
    int t = 1;
    while (t < n && t < cap) t <<= 1;
    return t;
}

// Cyclic Jacobi for the symmetric n x n matrix A (leading dimension ld), in place:
// on exit diag(A) holds the eigenvalues and, if V != null, the columns of V the
// eigenvectors.  cs: 2 * (n/2 + 1) doubles, pq: n/2 + 1 ints, red: 33 doubles of shared memory.
//
// One step applies the n/2 disjoint rotations of a round-robin pairing.  Threads are mapped in 2-D so
// that a thread's rotation (p, q, c, s) is fixed for the whole phase and no per-element index arithmetic
// remains (the first version spent ~150 instructions per element on k / half, k % half and two
// runtime modulos: 39 k cycles per step at n = 118):
//   column phase  A <- A J, V <- V J : thread = (pair t, row group)   -> elements (i, p), (i, q)
//   row phase     A <- J^T A         : thread = (column j, pair group) -> elements (p, j), (q, j)
__device__ void jacobi_eig(double *A, double *V, int n, int ld, double *cs, int *pq, double *red)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    if (V) {
        for (int k = tid; k < n * n; k += nt) V[(k / n) * ld + (k % n)] = (k / n == k % n) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (n < 2) return;
    const int ne = n + (n & 1);
    const int half = ne / 2;
    const int TT = pow2_at_least(half, nt), RG = nt / TT;     // column phase: pair index x row groups
    const int tt = tid & (TT - 1), ti = tid / TT;
    const int TJ = pow2_at_least(n, nt), PG = nt / TJ;        // row phase: column index x pair groups
    const int tj = tid & (TJ - 1), tg = tid / TJ;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, tot = 0.0;
        for (int i = tg; i < n; i += PG) {
            for (int j = tj; j < n; j += TJ) {
                const double a = A[i * ld + j];
                tot += a * a;
                if (i != j) off += a * a;
            }
        }
        off = block_sum(off, red);
        tot = block_sum(tot, red);
#ifdef GIK_BI_PROFILE
        if (blockIdx.x == 0 && tid == 0) printf("bi_profile   sweep %d n %d off/tot %.3e\n", sweep, n, off / tot);
#endif
        if (off <= 1e-33 * tot || tot == 0.0) break;
        for (int step = 0; step < ne - 1; ++step) {
            // rotation angles of this step's disjoint pairs; the pairs of the previous step are now
            // exactly diagonal (no pair repeats inside a sweep, so these writes touch nobody's reads)
            for (int t = tid; t < half; t += nt) {
                if (step > 0) {
                    const int old = pq[t];
                    if (old >= 0 && cs[2 * t + 1] != 0.0) {
                        const int p0 = old & 0xffff, q0 = old >> 16;
                        A[p0 * ld + q0] = 0.0;
                        A[q0 * ld + p0] = 0.0;
                    }
                }
                int p, q;
                if (t == 0) { p = ne - 1; q = step; }
                else { p = (step + t) % (ne - 1); q = (step - t + (ne - 1)) % (ne - 1); }
                double c = 1.0, s = 0.0;
                const bool real = p < n && q < n;
                if (real) {
                    const double apq = A[p * ld + q];
                    if (apq != 0.0) {
                        const double theta = (A[q * ld + q] - A[p * ld + p]) / (2.0 * apq);
                        const double tt2 = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(tt2 * tt2 + 1.0);
                        s = tt2 * c;
                    }
                }
                cs[2 * t] = c;
                cs[2 * t + 1] = s;
                pq[t] = real ? (p | (q << 16)) : -1;
            }
            __syncthreads();
            // A <- A J (and V <- V J): columns p, q of every row
            for (int t = tt; t < half; t += TT) {
                const int code = pq[t];
                const double c = cs[2 * t], s = cs[2 * t + 1];
                if (code < 0 || s == 0.0) continue;
                const int p = code & 0xffff, q = code >> 16;
                for (int i = ti; i < n; i += RG) {
                    const double aip = A[i * ld + p], aiq = A[i * ld + q];
                    A[i * ld + p] = c * aip - s * aiq;
                    A[i * ld + q] = s * aip + c * aiq;
                    if (V) {
                        const double vip = V[i * ld + p], viq = V[i * ld + q];
                        V[i * ld + p] = c * vip - s * viq;
                        V[i * ld + q] = s * vip + c * viq;
                    }
                }
            }
            __syncthreads();
            // A <- J^T A: rows p, q of every column
            for (int t = tg; t < half; t += PG) {
                const int code = pq[t];
                const double c = cs[2 * t], s = cs[2 * t + 1];
                if (code < 0 || s == 0.0) continue;
                const int p = code & 0xffff, q = code >> 16;
                for (int j = tj; j < n; j += TJ) {
                    const double apj = A[p * ld + j], aqj = A[q * ld + j];
                    A[p * ld + j] = c * apj - s * aqj;
                    A[q * ld + j] = s * apj + c * aqj;
                }
            }
            __syncthreads();
        }
        // the last step's pairs
        for (int t = tid; t < half; t += nt) {
            const int old = pq[t];
            if (old >= 0 && cs[2 * t + 1] != 0.0) {
                const int p0 = old & 0xffff, q0 = old >> 16;
                A[p0 * ld + q0] = 0.0;
                A[q0 * ld + p0] = 0.0;
            }
        }
        __syncthreads();
    }
    __syncthreads();
}

// Number of eigenvalues > sigma of the symmetric n x n matrix A (in place, destroyed): Householder reduction to
// tridiagonal form (d, e) and a Sturm count (signs of the pivots of T - sigma I).  The reference's rank heuristic
// (dgp.py:163-171) only needs this count; a full Jacobi solve cost 18.6 M cycles per goal at n = 118.
// v, pv, d, e: n doubles of shared memory each; red: 33 doubles.  Returns the same value on every thread.
__device__ int count_eigenvalues_above(double *A, int n, int ld, double sigma, double *v, double *pv, double *d,
                                       double *e, double *red)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int TJ = pow2_at_least(n, nt), G = nt / TJ;
    const int tj = tid & (TJ - 1), tg = tid / TJ;
    for (int k = 0; k + 2 < n; ++k) {
        const int m = n - k - 1;                 // size of the trailing block A22 = A[k+1.., k+1..]
        double *A22 = A + (k + 1) * ld + (k + 1);
        // x = A[k+1.., k]
        double part = 0.0;
        for (int i = tid; i < m; i += nt) { const double xi = A[(k + 1 + i) * ld + k]; v[i] = xi; part += xi * xi; }
        const double xx = block_sum(part, red);   // (synchronises)
        const double x0 = v[0];
        const double tail = xx - x0 * x0;         // ||x[1:]||^2
        if (!(tail > 0.0)) {                      // already tridiagonal in this column
            if (tid == 0) { d[k] = A[k * ld + k]; e[k] = x0; }
            __syncthreads();
            continue;
        }
        const double alpha = x0 > 0.0 ? -sqrt(xx) : sqrt(xx);
        const double v0 = x0 - alpha;
        const double vv = tail + v0 * v0;
        const double beta = 2.0 / vv;
        __syncthreads();
        if (tid == 0) { v[0] = v0; d[k] = A[k * ld + k]; e[k] = alpha; }
        __syncthreads();
        // p = beta A22 v: thread = (row, column group); the groups add their partial sums in turn (fixed order)
        for (int i = tid; i < m; i += nt) pv[i] = 0.0;
        __syncthreads();
        for (int gsel = 0; gsel < G; ++gsel) {
            if (tg == gsel) {
                for (int i = tj; i < m; i += TJ) {
                    double acc = 0.0;
                    for (int j = tg; j < m; j += G) acc = fma(A22[j * ld + i], v[j], acc);   // A22 is symmetric: row j, lanes over i
                    pv[i] += beta * acc;
                }
            }
            __syncthreads();
        }
        part = 0.0;
        for (int i = tid; i < m; i += nt) part += pv[i] * v[i];
        const double pTv = block_sum(part, red);
        const double Kc = 0.5 * beta * pTv;
        __syncthreads();
        for (int i = tid; i < m; i += nt) pv[i] -= Kc * v[i];     // w = p - K v
        __syncthreads();
        // A22 <- A22 - v w^T - w v^T
        for (int i = tg; i < m; i += G) {
            const double vi = v[i], wi = pv[i];
            for (int j = tj; j < m; j += TJ) A22[i * ld + j] -= vi * pv[j] + wi * v[j];
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (n >= 2) { d[n - 2] = A[(n - 2) * ld + (n - 2)]; e[n - 2] = A[(n - 1) * ld + (n - 2)]; }
        d[n - 1] = A[(n - 1) * ld + (n - 1)];
        // Sturm count: negative pivots of T - sigma I = eigenvalues below sigma
        int below = 0;
        double q = d[0] - sigma;
        if (q < 0.0) ++below;
        for (int i = 1; i < n; ++i) {
            if (q == 0.0) q = 1e-300;
            q = d[i] - sigma - e[i - 1] * e[i - 1] / q;
            if (q < 0.0) ++below;
        }
        red[0] = (double)(n - below);
    }
    __syncthreads();
    const int count = (int)(red[0] + 0.5);
    __syncthreads();
    return count;
}

// order[k] = index of the k-th largest value of d[0..n) (ties: lower index first); serial, n <= 128
__device__ void sort_desc(const double *d, int n, int *order)
{
    if (threadIdx.x == 0) {
        for (int k = 0; k < n; ++k) order[k] = k;
        for (int a = 1; a < n; ++a) {
            const int idx = order[a];
            const double v = d[idx];
            int bpos = a - 1;
            while (bpos >= 0 && d[order[bpos]] < v) { order[bpos + 1] = order[bpos]; --bpos; }
            order[bpos + 1] = idx;
        }
    }
    __syncthreads();
}

__global__ void k_bounds_init(const BiArgs a)
{
    extern __shared__ double smem[];
    const int N = a.N, NN = N * N, tid = threadIdx.x, nt = blockDim.x;
    // small arrays first (sizes: gik_bi_small_doubles in gik_common.cuh)
    double *cs = smem;                 // 2*(N/2+1) <= N+2
    double *red = cs + (N + 2);        // 33
    double *lam = red + 34;            // N   eigenvalues / row buffer
    double *rowM = lam + N;            // N
    double *Lp = rowM + N;             // N   LOWER of the pairs (., p_n) of this goal
    double *Lq = Lp + N;               // N   LOWER of the pairs (., q_n)
    double *part = Lq + N;             // 2 * 128 partial maxima of the two halves of a max-plus row
    int *order = reinterpret_cast<int *>(part + 256);  // N ints
    int *pq = order + N + (N & 1);     // N/2 + 1 ints: rotation pairs of the current Jacobi step
    double *mats = smem + gik_bi_small_doubles(N);
    double *M1, *M2, *M3;
    if (a.use_scratch == 2) {          // nothing fits: all three matrices in global scratch
        M1 = a.scratch + (size_t)blockIdx.x * 3 * NN;
        M2 = M1 + NN;
        M3 = M2 + NN;
    } else {
        M1 = mats;
        M2 = M1 + NN;
        M3 = a.use_scratch == 1 ? a.scratch + (size_t)blockIdx.x * NN : M2 + NN;
    }

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
            __syncthreads();
            // min-plus closure (Floyd-Warshall); row k / column k are fixed points of step k
            const int TJ = pow2_at_least(N, nt), G = nt / TJ;   // thread = (column, row group)
            const int tj = tid & (TJ - 1), tg = tid / TJ;
            for (int k = 0; k < N; ++k) {
                for (int i = tg; i < N; i += G) {
                    const double uik = Up[i * N + k];
                    for (int j = tj; j < N; j += TJ) {
                        const double via = uik + Up[k * N + j];
                        if (via < Up[i * N + j]) Up[i * N + j] = via;
                    }
                }
                __syncthreads();
            }
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
            // lower bounds row by row, two max-plus products (max is exact, so any split gives the same bits):
            //   rowM[b] = max(-Up[u,b], max_{a: L[a,b] > 0} (L[a,b] - Up[u,a]));  lower[u,v] = max(0, max_b (rowM[b] - Up[b,v]))
            // thread = (column, half of the contraction range)
            const int H = G < 2 ? 1 : 2;
            const int th = tg;               // which half (threads with tg >= H idle in these loops)
            const int chunk = (N + H - 1) / H;
            for (int u = 0; u < N; ++u) {
                if (th < H) {
                    const int a0 = th * chunk, a1 = min(N, a0 + chunk);
                    for (int bb = tj; bb < N; bb += TJ) {
                        double m = -INFINITY;
                        for (int aa = a0; aa < a1; ++aa) {
                            const double l = bb == gp ? Lp[aa] : (bb == gq ? Lq[aa] :
                                             (aa == gp ? Lp[bb] : (aa == gq ? Lq[bb] : a.bs_lower[aa * N + bb])));
                            if (l > 0.0) m = fmax(m, l - Up[u * N + aa]);
                        }
                        part[th * 128 + bb] = m;
                    }
                }
                __syncthreads();
                for (int bb = tid; bb < N; bb += nt) {
                    double m = -Up[u * N + bb];  // joining arc b -> b' of weight 0
                    for (int h = 0; h < H; ++h) m = fmax(m, part[h * 128 + bb]);
                    rowM[bb] = m;
                }
                __syncthreads();
                if (th < H) {
                    const int b0 = th * chunk, b1 = min(N, b0 + chunk);
                    for (int v = tj; v < N; v += TJ) {
                        double m = 0.0;
                        for (int bb = b0; bb < b1; ++bb) m = fmax(m, rowM[bb] - Up[bb * N + v]);
                        part[th * 128 + v] = m;
                    }
                }
                __syncthreads();
                for (int v = tid; v < N; v += nt) {
                    double m = part[v];
                    for (int h = 1; h < H; ++h) m = fmax(m, part[h * 128 + v]);
                    const double lo = (u == v) ? 0.0 : m;
                    const double up = Up[u * N + v];
                    if (a.lb_out) a.lb_out[(size_t)b * NN + u * N + v] = lo;
                    if (a.ub_out) a.ub_out[(size_t)b * NN + u * N + v] = up;
                    const double dr = lo + 0.9 * (up - lo);   // riemannian_solver.py:72
                    D[u * N + v] = dr * dr;
                }
                __syncthreads();
            }
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
        for (int i = tid; i < N; i += nt) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s += D[i * N + j];
            lam[i] = s / N;   // row means
        }
        __syncthreads();
        for (int j = tid; j < N; j += nt) {
            double s = 0.0;
            for (int i = 0; i < N; ++i) s += D[i * N + j];
            rowM[j] = s / N;  // column means
        }
        __syncthreads();
        double tot = 0.0;
        for (int i = tid; i < N; i += nt) tot += lam[i];
        tot = block_sum(tot, red) / N;
        double *G = M2, *V = M1;
        for (int k = tid; k < NN; k += nt) {
            const int i = k / N, j = k % N;
            G[k] = -0.5 * (D[k] - lam[i] - rowM[j] + tot);
        }
        __syncthreads();
        // symmetrise against rounding (D is symmetric up to the row-wise evaluation order)
        for (int k = tid; k < NN; k += nt) {
            const int i = k / N, j = k % N;
            if (i < j) {
                const double m = 0.5 * (G[i * N + j] + G[j * N + i]);
                G[i * N + j] = m;
                G[j * N + i] = m;
            }
        }
        __syncthreads();
        // ---------------- factor (dgp.py:150-159): X = V sqrt(max(lambda,0)), columns by descending lambda
        BI_TICK("gram");
        jacobi_eig(G, V, N, N, cs, pq, red);
        BI_TICK("eig_gram");
        // Eigenvector signs are arbitrary, yet the rank heuristic below is NOT invariant to them
        // (it reads a triangle of the non-symmetric factor).  The reference inherits whatever
        // LAPACK returns; here the sign is fixed canonically: the entry of largest magnitude of
        // every eigenvector is positive (first such entry on ties).
        for (int col = tid; col < N; col += nt) {
            double best = 0.0, sgn = 1.0;
            for (int i = 0; i < N; ++i) {
                const double v = V[i * N + col];
                if (fabs(v) > best) { best = fabs(v); sgn = v < 0.0 ? -1.0 : 1.0; }
            }
            if (sgn < 0.0)
                for (int i = 0; i < N; ++i) V[i * N + col] = -V[i * N + col];
            lam[col] = G[col * N + col];
        }
        __syncthreads();
        sort_desc(lam, N, order);
        double *X = M2;
        for (int k = tid; k < NN; k += nt) {
            const int i = k / N, col = k % N;
            const double ev = lam[order[col]];
            X[k] = ev > 0.0 ? V[i * N + order[col]] * sqrt(ev) : 0.0;
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
        // only the COUNT of eigenvalues above 1e-8 is used: tridiagonal reduction + Sturm count for the larger graphs
        // (65 536 chain20 goals: 371 -> 255 ms), the Jacobi solve for N <= 32 where its few barriers are cheaper
        int K;
        if (N > 32) {
            K = count_eigenvalues_above(Aw, N, N, 1e-8, lam, rowM, Lp, Lq, red);
        } else {
            jacobi_eig(Aw, nullptr, N, N, cs, pq, red);
            double cnt = 0.0;
            for (int i = tid; i < N; i += nt) cnt += Aw[i * N + i] > 1e-8 ? 1.0 : 0.0;
            K = (int)(block_sum(cnt, red) + 0.5);
        }
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
        }
        __syncthreads();
        BI_TICK("scatter");
        jacobi_eig(S, E, K, K, cs, pq, red);
        BI_TICK("eig_proj");
        for (int i = tid; i < K; i += nt) lam[i] = S[i * K + i];
        __syncthreads();
        sort_desc(lam, K, order);
        double *Yo = a.Y_init + (size_t)b * N * 3;
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
    a.omega_i = p->omega_i;
    a.omega_j = p->omega_j;
    a.omega_ptr = p->omega_ptr;
    a.omega_adj = p->omega_adj;
    const size_t mat = (size_t)N * N * sizeof(double);
    size_t smem = small_bytes(N) + (p->bi_mode == 0 ? 3 : (p->bi_mode == 1 ? 2 : 0)) * mat;
    // measured on B200: N = 16: 32 threads 3.9 k goals / ms (64 threads: 3.2 k); N = 44, 65 536 goals: 128 threads
    // 371 ms, 256 threads 479 ms (96 registers: 5 vs 2 CTAs / SM); N = 118 (one CTA / SM): 256 threads 5.8 goals / ms, 512: 7.9
    const int threads = N <= 20 ? 32 : (N <= 64 ? 128 : 512);
    int blocks = a.B;
    a.use_scratch = p->bi_mode;
    a.scratch = static_cast<double *>(workspace);
    if (p->bi_mode && !workspace) {
        gik_set_error("gik_bounds/gik_init/gik_bounds_init: a plan with N=%d needs a workspace of gik_workspace_bytes() "
                      "bytes (one per concurrently running call)", N);
        return GIK_EINVAL;
    }
    // shared-memory opt-in and occupancy depend on (device, N) only: looked up once per plan geometry
    static int cached_dev = -1, cached_N = -1, cached_cap = 0;
    if (cached_dev != p->device || cached_N != N) {
        GIK_CUDA(cudaFuncSetAttribute(k_bounds_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounds_init, threads, smem));
        cached_cap = p->sm_count * (per_sm < 1 ? 1 : per_sm);
        cached_dev = p->device;
        cached_N = N;
    }
    const int cap = p->bi_mode == 0 ? cached_cap : p->bi_blocks;
    if (blocks > cap) blocks = cap;
    k_bounds_init<<<blocks, threads, smem, st>>>(a);
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
// a goal fit in shared memory (N <= 96), else the spilled matrices of every resident CTA.
extern "C" int64_t gik_workspace_bytes(const GikPlan *p)
{
    if (!p || !p->bi_mode) return 0;
    return (int64_t)p->bi_blocks * (p->bi_mode == 1 ? 1 : 3) * p->N * p->N * (int64_t)sizeof(double);
}
