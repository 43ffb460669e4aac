// dmma_ab.cu -- one-off A/B behind DESIGN.md section 3 ("tensor cores only where N is large enough to pay").
//
// The only dense contraction of the IK hot path is the N x 3 . 3 x N Gram-type product of the dense obstacle
// case (KUKA + table, N = 118): s_ij = <x_i - x_j, w_i - w_j> = a_i + a_j - (X W^T)_ij - (W X^T)_ij needs the
// N x N matrix X W^T with contraction length K = 3.  This program produces that matrix for a batch of problems
//   (a) with plain DFMA: 3 FMA per entry, the way k_rtr_cta evaluates its pairs, and
//   (b) with the FP64 tensor-core instruction mma.sync.m8n8k4 (K padded from 3 to 4, one 8 x 8 tile per
//       instruction),
// reduces it with the same mask so that both variants do the same downstream work, checks that the results
// agree, and prints the time per problem-matrix.  Build and run (also under ncu for the pipe utilisation):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/dmma_ab tools/dmma_ab.cu && tools/dmma_ab
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int N = 128;          // nodes padded to a multiple of 8 (KUKA + table: 118)
constexpr int TILES = N / 8;
constexpr int REPS = 64;        // matrices per problem and launch (stands for tCG iterations)

__global__ void __launch_bounds__(128) k_fma(const double *__restrict__ X, const double *__restrict__ W,
                                             const unsigned *__restrict__ mask, double *__restrict__ out, int B)
{
    __shared__ double sx[4][N], sw[4][N];
    const int b = blockIdx.x;
    if (b >= B) return;
    for (int k = threadIdx.x; k < 4 * N; k += blockDim.x) {
        sx[k / N][k % N] = X[(size_t)b * 4 * N + k];
        sw[k / N][k % N] = W[(size_t)b * 4 * N + k];
    }
    __syncthreads();
    const int i = threadIdx.x;      // one row per thread, all columns
    double acc = 0.0;
    for (int rep = 0; rep < REPS; ++rep) {
        const double xi0 = sx[0][i] + rep, xi1 = sx[1][i], xi2 = sx[2][i];
        for (int j = 0; j < N; ++j) {
            const double s = fma(xi0, sw[0][j], fma(xi1, sw[1][j], xi2 * sw[2][j]));
            acc += (mask[i * (N / 32) + (j >> 5)] >> (j & 31) & 1u) ? s : 0.0;
        }
    }
    out[(size_t)b * N + i] = acc;
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// warp w of the CTA owns the row tiles w, w + 4, ...; per row tile it walks the 16 column tiles
__global__ void __launch_bounds__(128) k_dmma(const double *__restrict__ X, const double *__restrict__ W,
                                              const unsigned *__restrict__ mask, double *__restrict__ out, int B)
{
    __shared__ double sx[4][N], sw[4][N], srow[N];
    const int b = blockIdx.x;
    if (b >= B) return;
    for (int k = threadIdx.x; k < 4 * N; k += blockDim.x) {
        sx[k / N][k % N] = X[(size_t)b * 4 * N + k];
        sw[k / N][k % N] = W[(size_t)b * 4 * N + k];
    }
    for (int k = threadIdx.x; k < N; k += blockDim.x) srow[k] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ar = lane >> 2, ak = lane & 3;          // A fragment: row ar, k index ak; B fragment: k ak, column ar
    for (int rep = 0; rep < REPS; ++rep) {
        for (int ti = warp; ti < TILES; ti += 4) {
            const int i = ti * 8 + ar;
            double a = sx[ak][i];
            if (ak == 0) a += rep;
            if (ak == 3) a = 0.0;                      // K = 3 padded to 4
            double acc0 = 0.0, acc1 = 0.0;
            for (int tj = 0; tj < TILES; ++tj) {
                const double bfrag = ak == 3 ? 0.0 : sw[ak][tj * 8 + ar];
                double d0, d1;
                dmma884(d0, d1, a, bfrag, 0.0, 0.0);   // D[row ar][cols 2 ak, 2 ak + 1] of tile (ti, tj)
                const int j0 = tj * 8 + 2 * ak;
                const unsigned m = mask[i * (N / 32) + (j0 >> 5)] >> (j0 & 31);
                acc0 += (m & 1u) ? d0 : 0.0;
                acc1 += (m & 2u) ? d1 : 0.0;
            }
            double acc = acc0 + acc1;                  // row sums: the 4 lanes of a row hold disjoint columns
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (ak == 0) srow[i] += acc;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) out[(size_t)b * N + k] = srow[k];
}

int main()
{
    const int B = 148 * 16;
    std::vector<double> X((size_t)B * 4 * N), W((size_t)B * 4 * N);
    std::vector<unsigned> mask(N * N / 32);
    srand(1);
    for (auto &v : X) v = rand() / (double)RAND_MAX - 0.5;
    for (auto &v : W) v = rand() / (double)RAND_MAX - 0.5;
    for (auto &m : mask) m = (unsigned)rand() * 2654435761u | 0x11111111u;   // ~80 % dense, like omega of KUKA + table
    double *dX, *dW, *dO1, *dO2;
    unsigned *dM;
    cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dW, W.size() * 8);
    cudaMalloc(&dO1, (size_t)B * N * 8); cudaMalloc(&dO2, (size_t)B * N * 8);
    cudaMalloc(&dM, mask.size() * 4);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dM, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[2] = {0, 0};
    for (int variant = 0; variant < 2; ++variant) {
        for (int it = 0; it < 3; ++it) {   // last run is timed
            cudaEventRecord(e0);
            if (variant == 0) k_fma<<<B, 128>>>(dX, dW, dM, dO1, B);
            else k_dmma<<<B, 128>>>(dX, dW, dM, dO2, B);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms[variant], e0, e1);
        }
    }
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    std::vector<double> o1((size_t)B * N), o2((size_t)B * N);
    cudaMemcpy(o1.data(), dO1, o1.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(o2.data(), dO2, o2.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0, ref = 0;
    for (size_t k = 0; k < o1.size(); ++k) { err = fmax(err, fabs(o1[k] - o2[k])); ref = fmax(ref, fabs(o1[k])); }
    const double mats = (double)B * REPS;
    printf("N=%d problems=%d matrices/problem=%d\n", N, B, REPS);
    printf("DFMA  (3 FMA per entry)        : %8.3f ms  %7.1f ns per N x N matrix  %6.2f TFLOP/s useful (K=3)\n", ms[0],
           ms[0] * 1e6 / mats, mats * N * N * 6.0 / (ms[0] * 1e-3) / 1e12);
    printf("DMMA  (m8n8k4, K padded to 4) : %8.3f ms  %7.1f ns per N x N matrix  %6.2f TFLOP/s useful (K=3)\n", ms[1],
           ms[1] * 1e6 / mats, mats * N * N * 6.0 / (ms[1] * 1e-3) / 1e12);
    printf("max |difference| = %.3e (relative %.1e)\n", err, err / ref);
    return 0;
}
