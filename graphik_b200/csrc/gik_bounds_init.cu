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
#include "gik_common.cuh"

namespace {

struct BiArgs {
    int N, n_goal, n_goal_edges, n_omega_edges;
    const double *bs_lower, *bs_upper;
    const int32_t *goal_edge_i, *goal_edge_j, *goal_edge_slot;
    const int32_t *omega_i, *omega_j;
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

// Cyclic Jacobi for the symmetric n x n matrix A (leading dimension ld), in place:
// on exit diag(A) holds the eigenvalues and, if V != null, the columns of V the
// eigenvectors.  cs: 2 * (n/2 + 1) doubles, red: 33 doubles of shared memory.
__device__ void jacobi_eig(double *A, double *V, int n, int ld, double *cs, double *red)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    if (V) {
        for (int k = tid; k < n * n; k += nt) V[(k / n) * ld + (k % n)] = (k / n == k % n) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (n < 2) return;
    const int ne = n + (n & 1);
    const int half = ne / 2;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, tot = 0.0;
        for (int k = tid; k < n * n; k += nt) {
            const int i = k / n, j = k % n;
            const double a = A[i * ld + j];
            tot += a * a;
            if (i != j) off += a * a;
        }
        off = block_sum(off, red);
        tot = block_sum(tot, red);
        if (off <= 1e-33 * tot || tot == 0.0) break;
        for (int step = 0; step < ne - 1; ++step) {
            // rotation angles of this step's disjoint pairs
            for (int t = tid; t < half; t += nt) {
                int p, q;
                if (t == 0) { p = ne - 1; q = step; }
                else { p = (step + t) % (ne - 1); q = (step - t + (ne - 1)) % (ne - 1); }
                double c = 1.0, s = 0.0;
                if (p < n && q < n) {
                    const double apq = A[p * ld + q];
                    if (apq != 0.0) {
                        const double theta = (A[q * ld + q] - A[p * ld + p]) / (2.0 * apq);
                        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(tt * tt + 1.0);
                        s = tt * c;
                    }
                }
                cs[2 * t] = c;
                cs[2 * t + 1] = s;
            }
            __syncthreads();
            // A <- A J (and V <- V J): columns p, q of every row
            for (int k = tid; k < n * half; k += nt) {
                const int i = k / half, t = k % half;
                int p, q;
                if (t == 0) { p = ne - 1; q = step; }
                else { p = (step + t) % (ne - 1); q = (step - t + (ne - 1)) % (ne - 1); }
                if (p >= n || q >= n) continue;
                const double c = cs[2 * t], s = cs[2 * t + 1];
                if (s == 0.0) continue;
                const double aip = A[i * ld + p], aiq = A[i * ld + q];
                A[i * ld + p] = c * aip - s * aiq;
                A[i * ld + q] = s * aip + c * aiq;
                if (V) {
                    const double vip = V[i * ld + p], viq = V[i * ld + q];
                    V[i * ld + p] = c * vip - s * viq;
                    V[i * ld + q] = s * vip + c * viq;
                }
            }
            __syncthreads();
            // A <- J^T A: rows p, q of every column
            for (int k = tid; k < n * half; k += nt) {
                const int j = k % n, t = k / n;
                int p, q;
                if (t == 0) { p = ne - 1; q = step; }
                else { p = (step + t) % (ne - 1); q = (step - t + (ne - 1)) % (ne - 1); }
                if (p >= n || q >= n) continue;
                const double c = cs[2 * t], s = cs[2 * t + 1];
                if (s == 0.0) continue;
                const double apj = A[p * ld + j], aqj = A[q * ld + j];
                A[p * ld + j] = c * apj - s * aqj;
                A[q * ld + j] = s * apj + c * aqj;
            }
            __syncthreads();
            // the rotated pair is now exactly diagonal
            for (int t = tid; t < half; t += nt) {
                int p, q;
                if (t == 0) { p = ne - 1; q = step; }
                else { p = (step + t) % (ne - 1); q = (step - t + (ne - 1)) % (ne - 1); }
                if (p < n && q < n && cs[2 * t + 1] != 0.0) {
                    A[p * ld + q] = 0.0;
                    A[q * ld + p] = 0.0;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
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
    // small arrays first
    double *cs = smem;                 // 2*(N/2+1) <= N+2
    double *red = cs + (N + 2);        // 33
    double *lam = red + 34;            // N   eigenvalues / row buffer
    double *rowM = lam + N;            // N
    int *order = reinterpret_cast<int *>(rowM + N);  // N ints
    double *mats = rowM + N + (N + 1) / 2 + 1;
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
            for (int k = 0; k < N; ++k) {
                for (int e = tid; e < NN; e += nt) {
                    const int i = e / N, j = e % N;
                    const double via = Up[i * N + k] + Up[k * N + j];
                    if (via < Up[e]) Up[e] = via;
                }
                __syncthreads();
            }
            // lower bounds row by row: rowM[b] = max_a (L[a,b] - Up[u,a]); lower[u,v] = max_b (rowM[b] - Up[b,v])
            for (int u = 0; u < N; ++u) {
                for (int bb = tid; bb < N; bb += nt) {
                    double m = -Up[u * N + bb];  // joining arc b -> b' of weight 0
                    for (int aa = 0; aa < N; ++aa) {
                        const double l = a.bs_lower[aa * N + bb];
                        if (l > 0.0) m = fmax(m, l - Up[u * N + aa]);
                    }
                    rowM[bb] = m;
                }
                __syncthreads();
                // goal edges: L[i,j] = L[j,i] = sqrt(goal_d2); only 2 * n_anchor of them -> folded serially
                if (tid == 0) {
                    for (int e = 0; e < a.n_goal_edges; ++e) {
                        const int i = a.goal_edge_i[e], j = a.goal_edge_j[e];
                        const double l = sqrt(gd[a.goal_edge_slot[e]]);
                        rowM[j] = fmax(rowM[j], l - Up[u * N + i]);
                        rowM[i] = fmax(rowM[i], l - Up[u * N + j]);
                    }
                }
                __syncthreads();
                for (int v = tid; v < N; v += nt) {
                    double m = 0.0;
                    for (int bb = 0; bb < N; ++bb) m = fmax(m, rowM[bb] - Up[bb * N + v]);
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
        jacobi_eig(G, V, N, N, cs, red);
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
        jacobi_eig(Aw, nullptr, N, N, cs, red);
        double cnt = 0.0;
        for (int i = tid; i < N; i += nt) cnt += Aw[i * N + i] > 1e-8 ? 1.0 : 0.0;
        int K = (int)(block_sum(cnt, red) + 0.5);
        if (K > N) K = N;
        // ---------------- linear_projection (dgp.py:174-183): S = sum_{omega} (P_i-P_j)(P_i-P_j)^T, P = X[:, :K]
        double *S = M1, *E = M3;
        for (int k = tid; k < K * K; k += nt) {
            const int r = k / K, cidx = k % K;
            double s = 0.0;
            for (int e = 0; e < a.n_omega_edges; ++e) {
                const int i = a.omega_i[e], j = a.omega_j[e];
                s += (X[i * N + r] - X[j * N + r]) * (X[i * N + cidx] - X[j * N + cidx]);
            }
            S[r * K + cidx] = 2.0 * s;   // both (i,j) and (j,i) are nonzeros of omega
        }
        __syncthreads();
        jacobi_eig(S, E, K, K, cs, red);
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

size_t small_bytes(int N) { return (size_t)((N + 2) + 34 + N + N + (N + 1) / 2 + 1) * sizeof(double); }

int launch(const GikPlan *p, BiArgs &a, cudaStream_t st)
{
    const int N = p->N;
    a.N = N;
    a.n_goal = p->n_goal;
    a.n_goal_edges = p->n_goal_edges;
    a.n_omega_edges = p->n_omega_edges;
    a.bs_lower = p->bs_lower;
    a.bs_upper = p->bs_upper;
    a.goal_edge_i = p->goal_edge_i;
    a.goal_edge_j = p->goal_edge_j;
    a.goal_edge_slot = p->goal_edge_slot;
    a.omega_i = p->omega_i;
    a.omega_j = p->omega_j;
    const size_t mat = (size_t)N * N * sizeof(double);
    size_t smem = small_bytes(N) + (p->bi_mode == 0 ? 3 : (p->bi_mode == 1 ? 2 : 0)) * mat;
    int threads = N <= 20 ? 32 : (N <= 48 ? 128 : 256);
    int blocks = a.B;
    a.use_scratch = p->bi_mode;
    a.scratch = p->bi_scratch;
    if (p->bi_mode && !p->bi_scratch) {
        gik_set_error("gik_bounds/gik_init: plan for N=%d was created without bound tables (no scratch)", N);
        return GIK_EINVAL;
    }
    GIK_CUDA(cudaFuncSetAttribute(k_bounds_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int cap = p->bi_blocks;
    if (p->bi_mode == 0) {
        int per_sm = 0;
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounds_init, threads, smem));
        if (per_sm < 1) per_sm = 1;
        cap = p->sm_count * per_sm;
    }
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

extern "C" int gik_bounds(const GikPlan *p, const double *goal_d2, int32_t B, double *lb, double *ub, void *stream)
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
    return launch(p, a, (cudaStream_t)stream);
}

extern "C" int gik_init(const GikPlan *p, const double *lb, const double *ub, int32_t B, double *Y_init, void *stream)
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
    return launch(p, a, (cudaStream_t)stream);
}

extern "C" int gik_bounds_init(const GikPlan *p, const double *goal_d2, int32_t B, double *Y_init, void *stream)
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
    return launch(p, a, (cudaStream_t)stream);
}
