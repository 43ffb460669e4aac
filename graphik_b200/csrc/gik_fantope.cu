// gik_fantope.cu -- the closed-form Fantope step of CIDGIK's convex iteration, batched.
//
// Reference: solve_fantope_closed_form (solvers/convex_iteration.py:43-53), called once per convex iteration on the
// (n + d) x (n + d) Gram matrix G the SDP returned (convex_iteration.py:236-239):
//     _, Q = eigh(G);  Q = flip(Q, 1);  U = Q[:, d:];  C = U U^T
// i.e. C is the orthogonal projector onto the eigenvectors of the n - d SMALLEST eigenvalues of G, the minimiser of
// <G, Z> over the Fantope {0 <= Z <= I, tr Z = n - d}.  With an orthonormal eigenbasis C = I - sum over the d largest
// eigenvalues of v v^T, which is what the kernel forms (independent of eigenvector signs, hence comparable with
// numpy bit for tolerance).  This is the only part of the CIDGIK path that is arithmetic of
// the reference itself; the semidefinite programs in between are solved by MOSEK through cvxpy and have no
// counterpart here.
//
// One warp per matrix (n <= 32): cyclic Jacobi on A (symmetric, shared memory, stride n + 1), eigenvectors
// accumulated in V; lane k owns row / column k of every rotation update.
#include "gik_common.cuh"

namespace {

constexpr int kWarps = 4;

__global__ void __launch_bounds__(kWarps * 32) k_fantope(int n, int d, const double *__restrict__ G, int B,
                                                         double *__restrict__ C, double *__restrict__ evals)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ld = n + 1;
    double *A = smem + (size_t)warp * (2 * n * ld + 32);
    double *V = A + n * ld;
    double *lam = V + n * ld;   // [32]
    for (int b = blockIdx.x * kWarps + warp; b < B; b += gridDim.x * kWarps) {
        const double *Gb = G + (size_t)b * n * n;
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e % n;
            A[i * ld + j] = 0.5 * (Gb[i * n + j] + Gb[j * n + i]);   // eigh reads one triangle; G is symmetric up to solver noise
            V[i * ld + j] = i == j ? 1.0 : 0.0;
        }
        __syncwarp();
        for (int sweep = 0; sweep < 40; ++sweep) {
            // off-diagonal mass against the diagonal: stop at rounding level
            double off = 0.0, dia = 0.0;
            if (lane < n) {
                for (int j = 0; j < n; ++j) {
                    const double v = A[lane * ld + j];
                    if (j == lane) dia = v * v; else off = fma(v, v, off);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                off += __shfl_xor_sync(GIK_FULL_MASK, off, o, 32);
                dia += __shfl_xor_sync(GIK_FULL_MASK, dia, o, 32);
            }
            if (off <= 1e-32 * dia || off == 0.0) break;
            for (int p = 0; p < n - 1; ++p) {
                for (int q = p + 1; q < n; ++q) {
                    const double apq = A[p * ld + q];
                    if (apq == 0.0) continue;                       // uniform: every lane reads the same entry
                    const double app = A[p * ld + p], aqq = A[q * ld + q];
                    const double theta = (aqq - app) / (2.0 * apq);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                    const double c = 1.0 / sqrt(fma(t, t, 1.0)), s = t * c;
                    __syncwarp();
                    if (lane < n) {                                  // columns p, q of A and V
                        const double akp = A[lane * ld + p], akq = A[lane * ld + q];
                        A[lane * ld + p] = c * akp - s * akq;
                        A[lane * ld + q] = s * akp + c * akq;
                        const double vkp = V[lane * ld + p], vkq = V[lane * ld + q];
                        V[lane * ld + p] = c * vkp - s * vkq;
                        V[lane * ld + q] = s * vkp + c * vkq;
                    }
                    __syncwarp();
                    if (lane < n) {                                  // rows p, q of A
                        const double apk = A[p * ld + lane], aqk = A[q * ld + lane];
                        A[p * ld + lane] = c * apk - s * aqk;
                        A[q * ld + lane] = s * apk + c * aqk;
                    }
                    __syncwarp();
                }
            }
        }
        // rank of every eigenvalue (ascending, ties by index): lane k counts the eigenvalues before its own
        const double mine = lane < n ? A[lane * ld + lane] : 0.0;
        lam[lane] = mine;
        __syncwarp();
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (lam[j] < mine) || (lam[j] == mine && j < lane);
        if (evals && lane < n) evals[(size_t)b * n + rank] = mine;
        const unsigned top = __ballot_sync(GIK_FULL_MASK, lane < n && rank >= n - d);   // the d largest eigenvalues
        // C = I - sum over the d largest of v v^T
        double *Cb = C + (size_t)b * n * n;
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e % n;
            double acc = i == j ? 1.0 : 0.0;
            for (int k = 0; k < n; ++k)
                if (top >> k & 1u) acc = fma(-V[i * ld + k], V[j * ld + k], acc);
            Cb[e] = acc;
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int gik_fantope(int32_t n, int32_t d, const double *G, int32_t B, double *C, double *eigvals, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!G || !C || B < 0 || n < 1 || d < 0 || d > n) { gik_set_error("gik_fantope: bad argument"); return GIK_EINVAL; }
    if (n > 32) { gik_set_error("gik_fantope: n=%d exceeds the compiled limit of 32", n); return GIK_ELIMIT; }
    const size_t smem = (size_t)kWarps * (2 * n * (n + 1) + 32) * sizeof(double);
    int blocks = (B + kWarps - 1) / kWarps;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (smem > 48 * 1024)
        GIK_CUDA(cudaFuncSetAttribute(k_fantope, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fantope<<<blocks, kWarps * 32, smem, (cudaStream_t)stream>>>(n, d, G, B, C, eigvals);
    return gik_check_cuda(cudaGetLastError(), "k_fantope launch");
}
