// gik_sdp.cu -- the semidefinite programs of CIDGIK's convex iteration, batched:
// gik_sdp_solve (one program per goal per launch) and gik_cidgik_solve (the whole convex iteration of a goal in one
// launch: programs, Fantope steps and the stopping test of solvers/convex_iteration.py:221-266).
//
// Reference: solve_linear_cost_sdp (solvers/sdp_snl.py:874-967) hands
//     minimise <C, Z>  s.t.  <A_k, Z> = b_k,  Z[-d:, -d:] = I,  Z >= 0
// to MOSEK through cvxpy (sdp_snl.py:952), once per convex iteration (solvers/convex_iteration.py:221-234).  MOSEK is
// closed third-party code that is not in the reference tree and cannot be installed here: this kernel is NOT a
// restatement of it and PARITY WITH IT IS UNPINNED.  It solves the same programs with the same class of method
// (infeasible-start primal-dual interior point, HKM direction, Mehrotra predictor-corrector), to the tolerances the
// reference hands MOSEK (1e-6, solvers/sdp_formulations.py:10) or tighter, and is checked against a numpy statement
// of the same algorithm (the oracle's solve_sdp) and through solver-independent optimality certificates
// (tests/test_gpu_cidgik.py).
//
// Form.  Every constraint of the reference's program is a squared distance between two points, or an entry of the
// identity block: in the coordinates the host prepares (solvers/convex_iteration.py: the face of the cone the
// feasible set lives on) each one reads  w_k^T X w_k = b_k  with a vector w_k, so the solver takes
//     minimise <C, X>  s.t.  w_k^T X w_k = b_k (k < M),  X >= 0          (dual: S = C - sum y_k w_k w_k^T >= 0)
// and the Schur complement of the HKM direction is a Hadamard product,
//     M_kl = <w_k w_k^T, X w_l w_l^T S^-1> = (w_k^T X w_l) (w_l^T S^-1 w_k).
// Inequality rows (distance_range_constraints, sdp_snl.py:356-398: lower / upper bounds on the distance to an
// obstacle) carry a slack in a 1 x 1 block of the cone: w_k^T X w_k + tau_k s_k = b_k, s_k >= 0.
//
// Mapping: one warp per program (32-thread CTAs, 20 per SM), everything in shared memory (N <= 32, M <= 96; a UR10
// program has N = 6, M = 15, 8 KB).  The factorisations (Cholesky of S, of M and of the trial points of the step-length
// search -- no eigenvalue problem anywhere) and the triangular solves run with a lane per row, the small matrix
// products with a lane per entry.  Wider CTAs were measured and dropped: 128 threads 3.7 ms, 64 threads 1.7 ms for the
// first launch of 1024 UR10 programs (the second warp only waits at barriers), one warp the same latency at twice the
// programs in flight (16 384 programs per batch: 147 k -> 197 k solves/s).  The code is kept small on purpose (loops
// not unrolled, helpers not inlined, one body for predictor and corrector): with 20 warps per SM at different places of
// the kernel, instruction fetch was the largest stall of the 7 k-instruction version.  HBM traffic is the problem data
// in and the solution out, once.
#include "gik_fantope.cuh"

namespace {

constexpr int kThreads = 32;     // one warp per program: the matrices have 36-100 entries and half of the work is serial on one warp anyway
constexpr int kMaxN = 32, kMaxM = 96;

struct SdpArgs {
    int N, M, B;
    const double *C, *W, *b, *tau;
    const int32_t *active;
    GikSdpOpts o;
    double *X, *y, *obj, *resid;
    int32_t *iters, *status;
    // fused convex iteration (gik_cidgik_solve)
    const double *Lc;
    int max_convex, d;
    double abs_tol, rel_tol, accept;
    double *Cs, *values, *eig_sums;
    int32_t *n_convex, *feasible;
};

// sums of K per-thread values over the CTA, the same on every thread afterwards (fixed order: deterministic)
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(GIK_FULL_MASK, v[k], o, 32);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) scratch[warp * K + k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
#pragma unroll 1
        for (int w = 0; w < kThreads / 32; ++w) s += scratch[w * K + k];
        v[k] = s;
    }
}

// Out[r x c] = A[r x n] B[n x c], row major
__device__ __noinline__ void matmul(double *Out, const double *A, const double *Bm, int r, int n, int c)
{
#pragma unroll 1
    for (int e = threadIdx.x; e < r * c; e += kThreads) {
        const int i = e / c, j = e % c;
        double s = 0.0;
#pragma unroll 1
        for (int k = 0; k < n; ++k) s = fma(A[i * n + k], Bm[k * c + j], s);
        Out[e] = s;
    }
}

// Lower Cholesky factor of the n x n matrix A (leading dimension n) into L (may alias A) on ONE warp, lane r owning
// the rows r, r + 32, ... (R of them: n <= 32 R); the diagonal of L is stored INVERTED.  False (on every lane) when a
// pivot is not positive and finite.
template <int R>
__device__ __noinline__ bool warp_cholesky(const double *A, double *L, int n, int lane)
{
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int i = lane + 32 * q;
            s[q] = 0.0;
            if (i >= j && i < n) {
                double t = A[i * n + j];
#pragma unroll 1
                for (int k = 0; k < j; ++k) t = fma(-L[i * n + k], L[j * n + k], t);
                s[q] = t;
            }
        }
        double mine = s[0];
#pragma unroll
        for (int q = 1; q < R; ++q) mine = (j >> 5) == q ? s[q] : mine;
        const double dj = __shfl_sync(GIK_FULL_MASK, mine, j & 31, 32);
        if (!(dj > 0.0) || !isfinite(dj)) return false;
        const double rinv = rsqrt(dj);           // one reciprocal square root instead of a square root and a division
        __syncwarp();
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int i = lane + 32 * q;
            if (i == j) L[i * n + j] = rinv;     // the diagonal holds 1 / L_jj: every later use divides by it
            else if (i > j && i < n) L[i * n + j] = s[q] * rinv;
        }
        __syncwarp();
    }
    return true;
}

// x <- (L L^T)^-1 x on one warp (L as warp_cholesky leaves it: inverted diagonal)
__device__ __noinline__ void warp_cholesky_solve(const double *L, double *x, int n, int lane)
{
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        double p = 0.0;
#pragma unroll 1
        for (int k = lane; k < j; k += 32) p = fma(L[j * n + k], x[k], p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(GIK_FULL_MASK, p, o, 32);
        if (lane == 0) x[j] = (x[j] - p) * L[j * n + j];
        __syncwarp();
    }
#pragma unroll 1
    for (int j = n - 1; j >= 0; --j) {
        double p = 0.0;
#pragma unroll 1
        for (int k = j + 1 + lane; k < n; k += 32) p = fma(L[k * n + j], x[k], p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(GIK_FULL_MASK, p, o, 32);
        if (lane == 0) x[j] = (x[j] - p) * L[j * n + j];
        __syncwarp();
    }
}

// Largest alpha in {1} U (0, 1) with X + alpha dX positive definite: geometric backtracking by 0.7, then `bisections`
// halvings of the bracket (four: within 3 %; the predictor only feeds the centring heuristic and takes none), as
// _max_step of the oracle does.  Warp 0 only; T, L: n x n scratch.
__device__ __noinline__ double warp_max_step(const double *X, const double *dX, double *T, double *L, int n, int lane,
                                             int bisections)
{
    auto inside = [&](double alpha) {
#pragma unroll 1
        for (int e = lane; e < n * n; e += 32) T[e] = fma(alpha, dX[e], X[e]);
        __syncwarp();
        const bool ok = warp_cholesky<1>(T, L, n, lane);
        __syncwarp();
        return ok;
    };
    if (inside(1.0)) return 1.0;
    double hi = 1.0, lo = 0.7;
    while (!inside(lo)) {
        hi = lo;
        lo *= 0.7;
        if (lo < 1e-12) return 0.0;
    }
#pragma unroll 1
    for (int r = 0; r < bisections; ++r) {
        const double mid = 0.5 * (lo + hi);
        if (inside(mid)) lo = mid; else hi = mid;
    }
    return lo;
}

// The two step-length searches of an iteration side by side (n <= 16): lanes 0-15 search X + alpha dX, lanes 16-31
// S + alpha dS, each half with its own candidate, both running the same instruction stream (a half that has finished,
// or whose trial factorisation has already failed, is predicated off).  Same candidates and same arithmetic per trial
// as warp_max_step.  T, L: two n x n scratch matrices each.  Returns this lane's half's step (lanes 0-15: primal).
__device__ __noinline__ double warp_dual_max_step(const double *X, const double *dX, const double *S, const double *dS,
                                                  double *Ta, double *Tb, double *La, double *Lb, int n, int lane,
                                                  int bisections)
{
    const int half = lane >> 4, hl = lane & 15;
    const double *Mx = half ? S : X, *Dx = half ? dS : dX;
    double *T = half ? Tb : Ta, *L = half ? Lb : La;
    int phase = 0, count = 0;            // 0: full step, 1: backtracking, 2: bisection, 3: done
    double hi = 1.0, lo = 0.7, result = 0.0;
#pragma unroll 1
    while (__any_sync(GIK_FULL_MASK, phase != 3)) {
        const bool busy = phase != 3;
        const double alpha = phase == 0 ? 1.0 : (phase == 1 ? lo : 0.5 * (lo + hi));
        if (busy) {
#pragma unroll 1
            for (int e = hl; e < n * n; e += 16) T[e] = fma(alpha, Dx[e], Mx[e]);
        }
        __syncwarp();
        // Cholesky of T on this half (row hl), inverted diagonal as warp_cholesky; ok: no pivot failed so far
        bool ok = busy;
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            double t = 0.0;
            if (ok && hl >= j && hl < n) {
                t = T[hl * n + j];
#pragma unroll 1
                for (int k = 0; k < j; ++k) t = fma(-L[hl * n + k], L[j * n + k], t);
            }
            const double dj = __shfl_sync(GIK_FULL_MASK, t, j, 16);
            ok = ok && dj > 0.0 && isfinite(dj);
            if (!__any_sync(GIK_FULL_MASK, ok)) break;
            const double rinv = ok ? rsqrt(dj) : 0.0;
            __syncwarp();
            if (ok && hl >= j && hl < n) L[hl * n + j] = hl == j ? rinv : t * rinv;
            __syncwarp();
        }
        __syncwarp();
        if (phase == 0) {
            if (ok) { result = 1.0; phase = 3; } else phase = 1;
        } else if (phase == 1) {
            if (ok) {
                phase = bisections > 0 ? 2 : 3;
                result = lo;
            } else {
                hi = lo;
                lo *= 0.7;
                if (lo < 1e-12) { result = 0.0; phase = 3; }
            }
        } else if (phase == 2) {
            if (ok) lo = 0.5 * (lo + hi); else hi = 0.5 * (lo + hi);
            if (++count == bisections) { result = lo; phase = 3; }
        }
    }
    return result;
}

// LP: the program has inequality rows (slack blocks); the equality-only instantiation carries none of that code.
// FUSED: the whole convex iteration of a goal (convex_iteration.py:221-266) in the warp that owns it -- program, Fantope
// step in the coordinates of the face (C <- L (I - sum of the d largest u u^T) L^T with V^T V = L L^T, eigenpairs of
// L^T X L), stopping test -- instead of one launch per convex iteration and host glue in between.
template <bool LP, bool FUSED>
__global__ void __launch_bounds__(kThreads, 16) k_sdp(const SdpArgs a)
{
    extern __shared__ double sm[];
    const int N = a.N, M = a.M, NN = N * N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *C = sm, *X = C + NN, *S = X + NN, *Sinv = S + NN, *Rd = Sinv + NN, *dX = Rd + NN, *dS = dX + NN,
           *corr = dS + NN, *T1 = corr + NN, *T2 = T1 + NN, *L = T2 + NN, *Xbest = L + NN;
    double *W = Xbest + NN, *P = W + M * N, *Q = P + M * N, *G = Q + M * N;
    double *bv = G + M * M, *yv = bv + M, *rp = yv + M, *dy = rp + M, *ybest = dy + M, *tau = ybest + M,
           *sv = tau + M, *zv = sv + M, *dsv = zv + M, *dzv = dsv + M, *clp = dzv + M, *red = clp + M;   // red: [2 * 8]
    double *flag = red + 32;                                                           // [4]
    double *Lm = flag + 4, *Csm = Lm + NN, *JA = Csm + NN, *JV = JA + N * (N + 1), *lam = JV + N * (N + 1);   // lam: [32]
    static_assert(!FUSED || kThreads == 32, "the Fantope step runs on the warp that owns the program");
    unsigned char *row_of = reinterpret_cast<unsigned char *>(lam + 32), *col_of = row_of + NN;   // entry -> (i, j)
#pragma unroll 1
    for (int e = tid; e < NN; e += kThreads) {
        row_of[e] = (unsigned char)(e / N);
        col_of[e] = (unsigned char)(e % N);
    }
    __syncthreads();

#pragma unroll 1
    for (int prob = blockIdx.x; prob < a.B; prob += gridDim.x) {
        if (a.active && !a.active[prob]) continue;
        __syncthreads();
#pragma unroll 1
        for (int e = tid; e < NN; e += kThreads) {
            C[e] = a.C[(size_t)prob * NN + e];
            if (FUSED) Lm[e] = a.Lc[(size_t)prob * NN + e];
        }
#pragma unroll 1
        for (int e = tid; e < M * N; e += kThreads) W[e] = a.W[(size_t)prob * M * N + e];
#pragma unroll 1
        for (int k = tid; k < M; k += kThreads) {
            bv[k] = a.b[(size_t)prob * M + k];
            yv[k] = 0.0;
            // inequality k: w_k^T X w_k + tau_k s_k = b_k with a slack s_k >= 0 (tau = +1 upper, -1 lower bound) and its
            // dual z_k = -tau_k y_k >= 0 -- a 1 x 1 block of the cone next to X
            tau[k] = LP ? a.tau[k] : 0.0;
        }
        __syncthreads();
        int n_ineq = 0;
#pragma unroll 1
        for (int k = 0; LP && k < M; ++k) n_ineq += tau[k] != 0.0;
        // state of the convex iteration (FUSED; otherwise the loop body runs once)
        double last_cost = 1e6;
        int n_convex = 0, feasible = 0, sdp_total = 0;
        int status = GIK_SDP_INACCURATE, it = 0;
        double resid = INFINITY, pobj = 0.0;
#pragma unroll 1
        for (int cit = 0; cit < (FUSED ? a.max_convex : 1); ++cit) {
#pragma unroll 1
        for (int e = tid; e < NN; e += kThreads) X[e] = S[e] = row_of[e] == col_of[e] ? a.o.x0 : 0.0;
#pragma unroll 1
        for (int k = tid; k < M; k += kThreads) {
            yv[k] = 0.0;
            sv[k] = zv[k] = (LP && tau[k] != 0.0) ? a.o.x0 : 0.0;
            dsv[k] = dzv[k] = clp[k] = 0.0;
        }
        __syncthreads();
        double nb, nC;
        {
            double v[2] = {0.0, 0.0};
#pragma unroll 1
            for (int k = tid; k < M; k += kThreads) v[0] = fma(bv[k], bv[k], v[0]);
#pragma unroll 1
            for (int e = tid; e < NN; e += kThreads) v[1] = fma(C[e], C[e], v[1]);
            block_sum<2>(v, red);
            nb = 1.0 + sqrt(v[0]);
            nC = 1.0 + sqrt(v[1]);
        }
        status = GIK_SDP_INACCURATE;
        resid = INFINITY;
        double best_resid = INFINITY, best_obj = 0.0;
        int best_it = 0;
#pragma unroll 1
        for (it = 0;; ++it) {
            // residuals
            matmul(P, W, X, M, N, N);
#pragma unroll 1
            for (int e = tid; e < NN; e += kThreads) {
                const int i = row_of[e], j = col_of[e];
                double s = C[e] - S[e];
#pragma unroll 1
                for (int k = 0; k < M; ++k) s = fma(-yv[k] * W[k * N + i], W[k * N + j], s);
                Rd[e] = s;
            }
            __syncthreads();
#pragma unroll 1
            for (int k = tid; k < M; k += kThreads) {
                double s = LP ? bv[k] - tau[k] * sv[k] : bv[k];
#pragma unroll 1
                for (int j = 0; j < N; ++j) s = fma(-P[k * N + j], W[k * N + j], s);
                rp[k] = s;
            }
            __syncthreads();
            double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int e = tid; e < NN; e += kThreads) {
                v[0] = fma(X[e], S[e], v[0]);
                v[1] = fma(C[e], X[e], v[1]);
                v[2] = fma(Rd[e], Rd[e], v[2]);
                const double t = C[e] - Rd[e];
                v[3] = fma(t, t, v[3]);
            }
#pragma unroll 1
            for (int k = tid; k < M; k += kThreads) {
                v[4] = fma(bv[k], yv[k], v[4]);
                v[5] = fma(rp[k], rp[k], v[5]);
                if ((LP && tau[k] != 0.0)) {
                    const double rz = -tau[k] * yv[k] - zv[k];        // dual residual of the slack
                    v[0] = fma(sv[k], zv[k], v[0]);
                    v[2] = fma(rz, rz, v[2]);
                    v[3] = fma(rz, rz, v[3]);
                }
            }
            block_sum<6>(v, red);
            const double mu = v[0] / (N + n_ineq), dobj = v[4];
            pobj = v[1];
            const double pres = sqrt(v[5]) / nb, dres = sqrt(v[2]) / nC;
            const double gap = fabs(pobj - dobj) / (1.0 + fabs(pobj) + fabs(dobj));
            if (!isfinite(pobj + dobj + pres + dres)) { status = GIK_SDP_NUMERIC; break; }
            resid = fmax(pres, fmax(dres, gap));
            if (resid < a.o.tol) { status = GIK_SDP_OPTIMAL; break; }
            // Close to the solution cond(M) ~ 1 / mu^2 and further steps can make the iterate worse: remember the
            // best one and, once below 1e-4, give up after three steps without progress or a tenfold loss
            if (resid < best_resid) {
                best_resid = resid;
                best_obj = pobj;
                best_it = it;
#pragma unroll 1
                for (int e = tid; e < NN; e += kThreads) Xbest[e] = X[e];
#pragma unroll 1
                for (int k = tid; k < M; k += kThreads) ybest[k] = yv[k];
            } else if (best_resid < 1e-4 && (resid > 10.0 * best_resid || it - best_it >= 3)) {
                break;
            }
            // dual improving ray: A^T y + S ~ 0 with b^T y > 0 certifies that the program has no feasible point
            if (dobj > 0.0 && sqrt(v[3]) / dobj < 1e-8) { status = GIK_SDP_INFEASIBLE; break; }
            if (it >= a.o.maxiter) break;

            // S = L L^T, S^-1
            if (warp == 0) {
                const bool ok = warp_cholesky<1>(S, L, N, lane);
                if (lane == 0) flag[0] = ok ? 1.0 : 0.0;
                if (ok && lane < N) {            // column `lane` of L^-1 into T1
                    const int j = lane;
#pragma unroll 1
                    for (int i = 0; i < j; ++i) T1[i * N + j] = 0.0;
#pragma unroll 1
                    for (int i = j; i < N; ++i) {
                        double s = i == j ? 1.0 : 0.0;
#pragma unroll 1
                        for (int k = j; k < i; ++k) s = fma(-L[i * N + k], T1[k * N + j], s);
                        T1[i * N + j] = s * L[i * N + i];
                    }
                }
            }
            __syncthreads();
            if (flag[0] == 0.0) break;
#pragma unroll 1
            for (int e = tid; e < NN; e += kThreads) {
                const int i = row_of[e], j = col_of[e];
                double s = 0.0;
#pragma unroll 1
                for (int k = i > j ? i : j; k < N; ++k) s = fma(T1[k * N + i], T1[k * N + j], s);
                Sinv[e] = s;
            }
            __syncthreads();
            // Schur complement (lower triangle) and its factor
            matmul(Q, W, Sinv, M, N, N);
            __syncthreads();
#pragma unroll 1
            for (int e = tid; e < M * M; e += kThreads) {
                const int k = e / M, l = e % M;
                if (l > k) continue;
                double u = 0.0, w = 0.0;
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    u = fma(P[k * N + j], W[l * N + j], u);
                    w = fma(Q[l * N + j], W[k * N + j], w);
                }
                G[e] = u * w + (k == l && (LP && tau[k] != 0.0) ? sv[k] / zv[k] : 0.0);
            }
            __syncthreads();
            if (warp == 0) {
                const bool ok = M <= 32 ? warp_cholesky<1>(G, G, M, lane) : warp_cholesky<3>(G, G, M, lane);
                if (lane == 0) flag[0] = ok ? 1.0 : 0.0;
            }
            __syncthreads();
            if (flag[0] == 0.0) break;      // cond(M) ~ 1 / mu^2: rounding broke the factorisation; keep the iterate

            // predictor (nu = 0), then corrector with the centring target sigma mu and the second-order term
            // corr = dXa dSa: one body run twice keeps the kernel's code within the instruction cache
            double nu = 0.0, ap = 0.0, ad = 0.0;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const bool use_corr = pass == 1;
#pragma unroll 1
                for (int e = tid; e < NN; e += kThreads) T1[e] = Rd[e] + S[e];
                __syncthreads();
                matmul(T2, X, T1, N, N, N);
                __syncthreads();
#pragma unroll 1
                for (int e = tid; e < NN; e += kThreads) {
                    double t = T2[e];
                    if (row_of[e] == col_of[e]) t -= nu;
                    if (use_corr) t += corr[e];
                    T2[e] = t;
                }
                __syncthreads();
                matmul(T1, T2, Sinv, N, N, N);
                __syncthreads();
#pragma unroll 1
                for (int k = tid; k < M; k += kThreads) {
                    double s = rp[k];
#pragma unroll 1
                    for (int i = 0; i < N; ++i) {
                        double r = 0.0;
#pragma unroll 1
                        for (int j = 0; j < N; ++j) r = fma(T1[i * N + j], W[k * N + j], r);
                        s = fma(W[k * N + i], r, s);
                    }
                    if ((LP && tau[k] != 0.0)) {
                        const double rz = -tau[k] * yv[k] - zv[k];
                        s -= tau[k] * (((nu - clp[k]) / zv[k] - sv[k]) - sv[k] / zv[k] * rz);
                    }
                    dy[k] = s;
                }
                __syncthreads();
                if (warp == 0) warp_cholesky_solve(G, dy, M, lane);
                __syncthreads();
#pragma unroll 1
                for (int k = tid; LP && k < M; k += kThreads)
                    if ((LP && tau[k] != 0.0)) {
                        dzv[k] = -tau[k] * yv[k] - zv[k] - tau[k] * dy[k];
                        dsv[k] = ((nu - clp[k]) / zv[k] - sv[k]) - sv[k] / zv[k] * dzv[k];
                    }
#pragma unroll 1
                for (int e = tid; e < NN; e += kThreads) {
                    const int i = row_of[e], j = col_of[e];
                    double s = Rd[e];
#pragma unroll 1
                    for (int k = 0; k < M; ++k) s = fma(-dy[k] * W[k * N + i], W[k * N + j], s);
                    dS[e] = s;
                }
                __syncthreads();
                matmul(T2, X, dS, N, N, N);
                __syncthreads();
                if (use_corr) {
#pragma unroll 1
                    for (int e = tid; e < NN; e += kThreads) T2[e] += corr[e];
                    __syncthreads();
                }
                matmul(T1, T2, Sinv, N, N, N);
                __syncthreads();
#pragma unroll 1
                for (int e = tid; e < NN; e += kThreads) T2[e] = nu * Sinv[e] - X[e] - T1[e];
                __syncthreads();
#pragma unroll 1
                for (int e = tid; e < NN; e += kThreads) {
                    const int i = row_of[e], j = col_of[e];
                    dX[e] = 0.5 * (T2[i * N + j] + T2[j * N + i]);
                }
                __syncthreads();
                if (warp == 0) {                       // step lengths to the boundary of the cone
                    if (N <= 16) {                     // both searches at once, one per half warp (T2, corr are free here)
                        const double st = warp_dual_max_step(X, dX, S, dS, T1, T2, L, corr, N, lane, pass ? 4 : 0);
                        if (lane == 0) flag[1] = st;
                        if (lane == 16) flag[2] = st;
                    } else {
                        const double sp = warp_max_step(X, dX, T1, L, N, lane, pass ? 4 : 0);
                        const double sd = warp_max_step(S, dS, T1, L, N, lane, pass ? 4 : 0);
                        if (lane == 0) { flag[1] = sp; flag[2] = sd; }
                    }
                }
                __syncthreads();
                ap = flag[1];
                ad = flag[2];
#pragma unroll 1
                for (int k = 0; LP && k < M; ++k)    // ratio test on the slacks (the same on every thread)
                    if ((LP && tau[k] != 0.0)) {
                        if (dsv[k] < 0.0) ap = fmin(ap, -sv[k] / dsv[k]);
                        if (dzv[k] < 0.0) ad = fmin(ad, -zv[k] / dzv[k]);
                    }
                if (pass == 0) {
                    double m[1] = {0.0};
#pragma unroll 1
                    for (int e = tid; e < NN; e += kThreads)
                        m[0] = fma(fma(ap, dX[e], X[e]), fma(ad, dS[e], S[e]), m[0]);
#pragma unroll 1
                    for (int k = tid; LP && k < M; k += kThreads)
                        if ((LP && tau[k] != 0.0)) m[0] = fma(fma(ap, dsv[k], sv[k]), fma(ad, dzv[k], zv[k]), m[0]);
                    block_sum<1>(m, red);
                    const double ratio = fmax(m[0] / (N + n_ineq) / mu, 0.0);
                    nu = fmin(1.0, ratio * ratio * ratio) * mu;
                    matmul(corr, dX, dS, N, N, N);
#pragma unroll 1
                    for (int k = tid; LP && k < M; k += kThreads) clp[k] = dsv[k] * dzv[k];
                    __syncthreads();
                }
            }
            ap = ap == 1.0 ? 1.0 : a.o.tau * ap;         // a full Newton step when it stays inside the cone
            ad = ad == 1.0 ? 1.0 : a.o.tau * ad;
            __syncthreads();
#pragma unroll 1
            for (int e = tid; e < NN; e += kThreads) {
                X[e] = fma(ap, dX[e], X[e]);
                S[e] = fma(ad, dS[e], S[e]);
            }
#pragma unroll 1
            for (int k = tid; k < M; k += kThreads) {
                yv[k] = fma(ad, dy[k], yv[k]);
                if (LP) {
                    sv[k] = fma(ap, dsv[k], sv[k]);
                    zv[k] = fma(ad, dzv[k], zv[k]);
                    clp[k] = 0.0;
                }
            }
            __syncthreads();
        }
        __syncthreads();
        if (status == GIK_SDP_INACCURATE && best_resid < resid) {        // the best iterate is the answer
#pragma unroll 1
            for (int e = tid; e < NN; e += kThreads) X[e] = Xbest[e];
#pragma unroll 1
            for (int k = tid; k < M; k += kThreads) yv[k] = ybest[k];
            pobj = best_obj;
            resid = best_resid;
            __syncthreads();
        }
        if (!FUSED) break;
        // ---- convex_iteration.py:236-266 for this goal
        sdp_total += it;
        const int code = status == GIK_SDP_INACCURATE && resid > a.accept ? GIK_SDP_NUMERIC : status;
        if (code >= GIK_SDP_INFEASIBLE) { feasible = code - 1; break; }      // INFEASIBLE / SOLVER_ERROR (:237-245)
        if (tid == 0) a.values[(size_t)prob * a.max_convex + cit] = pobj;
        // Zs = L^T X L (symmetrised as eigh reads it) into the Jacobi buffer, V = I
        matmul(T1, X, Lm, N, N, N);
        __syncthreads();
        const int ld = N + 1;
#pragma unroll 1
        for (int e = tid; e < NN; e += kThreads) {
            const int i = row_of[e], j = col_of[e];
            double zij = 0.0, zji = 0.0;
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
                zij = fma(Lm[k * N + i], T1[k * N + j], zij);
                zji = fma(Lm[k * N + j], T1[k * N + i], zji);
            }
            JA[i * ld + j] = 0.5 * (zij + zji);
            JV[i * ld + j] = i == j ? 1.0 : 0.0;
        }
        __syncthreads();
        int rank;
        double mine;
        const unsigned top = gik_warp_fantope_eig(JA, JV, lam, N, a.d, lane, &rank, &mine);
        double small[1] = {lane < N && rank < N - a.d ? mine : 0.0};       // sum of the N - d smallest eigenvalues
        block_sum<1>(small, red);
        if (tid == 0) a.eig_sums[(size_t)prob * a.max_convex + cit] = small[0];
        // Cs = I - sum over the d largest of u u^T (into T2), C <- L Cs L^T
#pragma unroll 1
        for (int e = tid; e < NN; e += kThreads) {
            const int i = row_of[e], j = col_of[e];
            double acc = i == j ? 1.0 : 0.0;
#pragma unroll 1
            for (int k = 0; k < N; ++k)
                if (top >> k & 1u) acc = fma(-JV[i * ld + k], JV[j * ld + k], acc);
            Csm[e] = acc;
        }
        __syncthreads();
        matmul(T1, Lm, Csm, N, N, N);
        __syncthreads();
#pragma unroll 1
        for (int e = tid; e < NN; e += kThreads) {
            const int i = row_of[e], j = col_of[e];
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k < N; ++k) acc = fma(T1[i * N + k], Lm[j * N + k], acc);
            C[e] = acc;
        }
        __syncthreads();
        ++n_convex;
        const double change = last_cost - pobj;                              // :262-266
        if (fabs(change) <= a.abs_tol || pobj <= a.abs_tol || fabs(change) / fabs(last_cost) < a.rel_tol) break;
        last_cost = pobj;
        }   // convex iterations
        __syncthreads();
#pragma unroll 1
        for (int e = tid; e < NN; e += kThreads) {
            a.X[(size_t)prob * NN + e] = X[e];
            if (FUSED) a.Cs[(size_t)prob * NN + e] = n_convex ? Csm[e] : (row_of[e] == col_of[e] ? 1.0 : 0.0);
        }
        if (a.y) for (int k = tid; k < M; k += kThreads) a.y[(size_t)prob * M + k] = yv[k];
        if (tid == 0) {
            a.obj[prob] = pobj;
            a.resid[prob] = resid;
            a.iters[prob] = FUSED ? sdp_total : it;
            a.status[prob] = status;
            if (FUSED) {
                a.n_convex[prob] = n_convex;
                a.feasible[prob] = feasible;
            }
        }
    }
}

}  // namespace

extern "C" int gik_sdp_default_opts(GikSdpOpts *o)
{
    if (!o) { gik_set_error("gik_sdp_default_opts: null argument"); return GIK_EINVAL; }
    o->tol = 1e-7;       // the reference asks MOSEK for 1e-6 (sdp_formulations.py:10)
    o->maxiter = 50;
    o->tau = 0.95;       // fraction of the step to the boundary of the cone
    o->x0 = 10.0;        // X = S = x0 I at the start
    return GIK_OK;
}

static int launch_sdp(SdpArgs &a, bool fused, void *stream)
{
    const int N = a.N, M = a.M;
    if (N > kMaxN || M > kMaxM) {
        gik_set_error("gik_sdp_solve: N = %d, M = %d exceed the limits %d, %d", N, M, kMaxN, kMaxM);
        return GIK_ELIMIT;
    }
    if (!(a.o.tol > 0.0) || a.o.maxiter < 1 || !(a.o.tau > 0.0 && a.o.tau < 1.0) || !(a.o.x0 > 0.0)) {
        gik_set_error("gik_sdp_solve: tol > 0, maxiter >= 1, 0 < tau < 1, x0 > 0 required");
        return GIK_EINVAL;
    }
    const size_t doubles = (size_t)14 * N * N + (size_t)2 * N * (N + 1) + (size_t)3 * M * N + (size_t)M * M +
                           (size_t)11 * M + 32 + 4 + 32 + (2 * (size_t)N * N + 7) / 8;
    const size_t smem = doubles * sizeof(double);
    if (smem > 227 * 1024) {
        gik_set_error("gik_sdp_solve: needs %zu bytes of shared memory per CTA", smem);
        return GIK_ELIMIT;
    }
    // shared-memory opt-in, occupancy and SM count are properties of (kernel, smem size, device): looked up once
    static size_t cached_smem[4] = {~(size_t)0, ~(size_t)0, ~(size_t)0, ~(size_t)0};
    static int cached_blocks[4] = {0, 0, 0, 0}, cached_dev[4] = {-1, -1, -1, -1};
    const int v = (a.tau ? 1 : 0) + (fused ? 2 : 0);
    void (*kernel)(SdpArgs) = v == 0 ? k_sdp<false, false> : v == 1 ? k_sdp<true, false>
                            : v == 2 ? k_sdp<false, true> : k_sdp<true, true>;
    int dev = 0;
    GIK_CUDA(cudaGetDevice(&dev));
    if (cached_smem[v] != smem || cached_dev[v] != dev) {
        int sms = 0, per_sm = 0;
        GIK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        GIK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
        cached_blocks[v] = sms * (per_sm < 1 ? 1 : per_sm);
        cached_smem[v] = smem;
        cached_dev[v] = dev;
    }
    int blocks = cached_blocks[v];
    if (blocks > a.B) blocks = a.B;
    kernel<<<blocks, kThreads, smem, (cudaStream_t)stream>>>(a);
    return gik_check_cuda(cudaGetLastError(), "k_sdp launch");
}

extern "C" int gik_sdp_solve(int32_t N, int32_t M, const double *C, const double *W, const double *b,
                             const double *tau, const int32_t *active, int32_t B, const GikSdpOpts *opts, double *X,
                             double *y, double *obj, double *resid, int32_t *iters, int32_t *status, void *stream)
{
    if (B == 0) return GIK_OK;
    if (B < 0 || N < 1 || M < 1 || !C || !W || !b || !X || !obj || !resid || !iters || !status) {
        gik_set_error("gik_sdp_solve: bad argument");
        return GIK_EINVAL;
    }
    SdpArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.M = M; a.B = B; a.C = C; a.W = W; a.b = b; a.tau = tau; a.active = active;
    if (opts) a.o = *opts; else gik_sdp_default_opts(&a.o);
    a.X = X; a.y = y; a.obj = obj; a.resid = resid; a.iters = iters; a.status = status;
    return launch_sdp(a, false, stream);
}

extern "C" int gik_cidgik_solve(int32_t N, int32_t M, int32_t d, const double *G, const double *Lc, const double *W,
                                const double *b, const double *tau, int32_t B, const GikSdpOpts *opts,
                                int32_t max_iters, double abs_eig_sum_tol, double rel_eig_sum_tol, double sdp_accept,
                                double *X, double *y, double *Cs, double *values, double *eig_sums, int32_t *n_iters,
                                int32_t *feasible, double *obj, double *resid, int32_t *sdp_iters, int32_t *status,
                                void *stream)
{
    if (B == 0) return GIK_OK;
    if (B < 0 || N < 1 || M < 1 || d < 0 || d > N || max_iters < 1 || !G || !Lc || !W || !b || !X || !Cs || !values ||
        !eig_sums || !n_iters || !feasible || !obj || !resid || !sdp_iters || !status) {
        gik_set_error("gik_cidgik_solve: bad argument");
        return GIK_EINVAL;
    }
    SdpArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.M = M; a.B = B; a.C = G; a.W = W; a.b = b; a.tau = tau;
    if (opts) a.o = *opts; else gik_sdp_default_opts(&a.o);
    a.X = X; a.y = y; a.obj = obj; a.resid = resid; a.iters = sdp_iters; a.status = status;
    a.Lc = Lc; a.max_convex = max_iters; a.d = d;
    a.abs_tol = abs_eig_sum_tol; a.rel_tol = rel_eig_sum_tol; a.accept = sdp_accept;
    a.Cs = Cs; a.values = values; a.eig_sums = eig_sums; a.n_convex = n_iters; a.feasible = feasible;
    return launch_sdp(a, true, stream);
}
