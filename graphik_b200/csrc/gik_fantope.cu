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
#include "gik_fantope.cuh"

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
        int rank;
        double mine;
        const unsigned top = gik_warp_fantope_eig(A, V, lam, n, d, lane, &rank, &mine);
        if (evals && lane < n) evals[(size_t)b * n + rank] = mine;
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
