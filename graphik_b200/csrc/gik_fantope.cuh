// gik_fantope.cuh -- warp-level pieces of the closed-form Fantope step (solve_fantope_closed_form,
// solvers/convex_iteration.py:43-53), shared by k_fantope (gik_fantope.cu) and the fused convex iteration (gik_sdp.cu).
#pragma once
#include "gik_common.cuh"

// Cyclic Jacobi on the symmetric n x n matrix A (shared memory, leading dimension ld = n + 1, overwritten: its diagonal
// ends as the eigenvalues), eigenvectors accumulated in the columns of V (must hold the identity on entry); lane k owns
// row / column k of every rotation update.  Then the rank of every eigenvalue (ascending, ties by index): returns the mask
// of the d largest; *rank_out / *mine_out: rank and value of this lane's eigenvalue (lane < n).  lam: 32 doubles scratch.
__device__ __forceinline__ unsigned gik_warp_fantope_eig(double *A, double *V, double *lam, int n, int d, int lane,
                                                         int *rank_out, double *mine_out)
{
    const int ld = n + 1;
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
    const unsigned top = __ballot_sync(GIK_FULL_MASK, lane < n && rank >= n - d);   // the d largest eigenvalues
    *rank_out = rank;
    *mine_out = mine;
    return top;
}
