// fp64_peak.cu -- measures the sustained FP64 FMA rate of the device (MEASURED_PEAKS.json has no fp64
// entry).  Each thread runs ILP independent DFMA chains; reports TFLOP/s for several occupancies.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_fma(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int warps_per_sm, int sms, double *out)
{
    const int iters = 200000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int threads = 32 * warps_per_sm;
    k_fma<ILP><<<sms, threads>>>(out, 1000, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k_fma<ILP><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * ILP * (double)iters * threads * sms;
    printf("ILP %2d warps/SM %2d : %7.2f TFLOP/s  (%.3f warp-DFMA per clk per SM at 1.965 GHz)\n", ILP, warps_per_sm,
           flops / ms * 1e-9, flops / 64.0 / (ms * 1e-3) / sms / 1.965e9);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {1, 4, 8, 16, 32}) run<1>(w, p.multiProcessorCount, out);
    for (int w : {4, 8, 16, 32}) run<4>(w, p.multiProcessorCount, out);
    for (int w : {4, 8, 16}) run<8>(w, p.multiProcessorCount, out);
    return 0;
}
