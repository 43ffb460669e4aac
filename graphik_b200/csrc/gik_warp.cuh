// gik_warp.cuh -- warp-level reductions shared by the one-warp-per-problem trust-region kernels.
#pragma once
#include "gik_common.cuh"

template <int LPN, int K>
__device__ __forceinline__ void node_allreduce(double (&v)[K])
{
#pragma unroll
    for (int off = 16; off >= LPN; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(GIK_FULL_MASK, v[k], off, 32);
    }
}

// Transposed butterfly all-reduce of KP (4 or 8) scalars over the nodes of a warp.  At each of
// the first log2(KP) levels a lane keeps half of its values and ships the other half to its
// partner, so the shuffle count is KP + 2 * (remaining levels) instead of KP * levels; the
// totals are then fetched from the lanes that own them.  Every lane ends with identical bits.
// (ncu on the plain butterfly: SHFL issues at ~4 cycles each and was 23 % of all instructions.)
template <int LPN, int KP>
__device__ __forceinline__ void node_allreduce_t(double (&v)[KP], int lane)
{
    static_assert(KP == 4 || KP == 8, "KP must be 4 or 8");
    double cur[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) cur[k] = v[k];
    int cnt = KP;
#pragma unroll
    for (int off = 16; off >= LPN; off >>= 1) {
        if (cnt > 1) {
            const int half = cnt / 2;
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < KP / 2; ++k) {
                if (k < half) {
                    const double keep = up ? cur[half + k] : cur[k];
                    const double send = up ? cur[k] : cur[half + k];
                    cur[k] = keep + __shfl_xor_sync(GIK_FULL_MASK, send, off, 32);
                }
            }
            cnt = half;
        } else {
            cur[0] += __shfl_xor_sync(GIK_FULL_MASK, cur[0], off, 32);
        }
    }
    // owner of scalar k: lane bits (16, 8[, 4]) spell k, most significant first
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int src = (KP == 8) ? (((k >> 2) & 1) * 16 + ((k >> 1) & 1) * 8 + (k & 1) * 4)
                                  : (((k >> 1) & 1) * 16 + (k & 1) * 8);
        v[k] = __shfl_sync(GIK_FULL_MASK, cur[0], src, 32);
    }
}

// sum over the LPN lanes of a node (both lanes end with identical bits)
template <int LPN, int K>
__device__ __forceinline__ void pair_combine(double (&v)[K])
{
#pragma unroll
    for (int off = LPN / 2; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(GIK_FULL_MASK, v[k], off, 32);
    }
}

