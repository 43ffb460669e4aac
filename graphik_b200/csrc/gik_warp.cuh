// gik_warp.cuh -- warp-level reductions shared by the one-warp-per-problem trust-region kernels.
#pragma once
#include "gik_common.cuh"

// Shared-memory accesses by 32-bit shared-window address + immediate byte offset.  With generic pointers the
// compiler re-derives the window base (S2R SR_CgaCtaId, LEA) and the element address (SHL, LOP3, IADD) at every
// use inside the register-capped solver loops; an address register formed once costs one register and nothing else.
__device__ __forceinline__ uint32_t gik_saddr(const void *p)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), out;
    // opaque copy: the compiler must keep the address in a register instead of re-deriving it (it would, to stay
    // under the register cap of the solver kernels, at 3..6 integer instructions per use)
    asm volatile("mov.u32 %0, %1;" : "=r"(out) : "r"(a));
    return out;
}
template <int OFF>
__device__ __forceinline__ double gik_lds(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ double2 gik_lds2(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void gik_sts(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0+%1], %2;" :: "r"(a), "n"(OFF), "d"(v) : "memory");
}

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


// All-reduce of K (4 or 8) scalars over the 32 / LPN node lanes of a warp through shared memory.
// Every node lane deposits its K values (row = scalar, column = node); lane L adds a quarter of row
// (L >> 2) % K with 128-bit reads and two xor shuffles finish the row; the K totals go through a
// second tiny buffer that all lanes read back.  Row strides (18 doubles for 16 columns read in blocks,
// 40 for 32 columns read interleaved) keep the 128-bit reads of a quarter-warp on disjoint banks.
// ~22 instructions for K = 8 on 16 nodes against ~68 for the transposed shuffle butterfly (whose
// keep / send selects and final broadcasts dominate), same latency; all lanes end with identical bits
// and the summation order is fixed.  R: 8 * gik_red_stride(NPW) doubles, T: 8 doubles, 16-byte aligned.
__host__ __device__ constexpr int gik_red_stride(int NPW) { return NPW == 16 ? 18 : 40; }

// Addresses of one lane for node_allreduce_s, formed once per kernel.
struct GikRedAddr {
    uint32_t dep;   // &R[column of this lane]
    uint32_t src;   // first 16 bytes this lane adds up
    uint32_t dst;   // &T[row of this lane]
    uint32_t tot;   // &T[0]
    bool writer;
};
template <int LPN>
__device__ __forceinline__ GikRedAddr gik_red_addr(double *R, double *T, int lane, int K)
{
    constexpr int NPW = 32 / LPN, RS = gik_red_stride(NPW);
    GikRedAddr a;
    const int row = (lane >> 2) & (K - 1), part = lane & 3;
    a.dep = gik_saddr(R + lane / LPN);
    a.src = gik_saddr(R + row * RS + (NPW == 16 ? part * 4 : part * 2));
    a.dst = gik_saddr(T + row);
    a.tot = gik_saddr(T);
    a.writer = part == 0;
    return a;
}

template <int LPN, int K>
__device__ __forceinline__ void node_allreduce_s(double (&v)[K], const GikRedAddr &ra)
{
    static_assert(K == 4 || K == 8, "K must be 4 or 8");
    constexpr int NPW = 32 / LPN, RS = gik_red_stride(NPW);
    // the LPN lanes of a node store identical values
    gik_sts<0 * RS * 8>(ra.dep, v[0]); gik_sts<1 * RS * 8>(ra.dep, v[1]);
    gik_sts<2 * RS * 8>(ra.dep, v[2]); gik_sts<3 * RS * 8>(ra.dep, v[3]);
    if (K == 8) {
        gik_sts<4 * RS * 8>(ra.dep, v[K - 4]); gik_sts<5 * RS * 8>(ra.dep, v[K - 3]);
        gik_sts<6 * RS * 8>(ra.dep, v[K - 2]); gik_sts<7 * RS * 8>(ra.dep, v[K - 1]);
    }
    __syncwarp();
    double s;
    if (NPW == 16) {
        const double2 a = gik_lds2<0>(ra.src), b = gik_lds2<16>(ra.src);
        s = __dadd_rn(__dadd_rn(a.x, a.y), __dadd_rn(b.x, b.y));
    } else {
        const double2 a = gik_lds2<0>(ra.src), b = gik_lds2<64>(ra.src), c = gik_lds2<128>(ra.src),
                      d = gik_lds2<192>(ra.src);
        s = __dadd_rn(__dadd_rn(__dadd_rn(a.x, a.y), __dadd_rn(b.x, b.y)), __dadd_rn(__dadd_rn(c.x, c.y), __dadd_rn(d.x, d.y)));
    }
    s += __shfl_xor_sync(GIK_FULL_MASK, s, 1, 32);
    s += __shfl_xor_sync(GIK_FULL_MASK, s, 2, 32);
    if (ra.writer) gik_sts<0>(ra.dst, s);   // for K = 4 the upper half-warp repeats the lower one's rows (same values)
    __syncwarp();
    const double2 t0 = gik_lds2<0>(ra.tot), t1 = gik_lds2<16>(ra.tot);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
    if (K == 8) {
        const double2 t2 = gik_lds2<32>(ra.tot), t3 = gik_lds2<48>(ra.tot);
        v[K - 4] = t2.x; v[K - 3] = t2.y; v[K - 2] = t3.x; v[K - 1] = t3.y;
    }
}

// The same sums as node_allreduce_s -- the same binary tree over the node columns, hence the same bits -- by xor
// butterflies: 4 (5) dependent shuffle levels instead of two shared-memory round trips plus two levels, i.e. a
// shorter critical path at ~3x the instructions.  Used by the latency variants of the solver kernels (few warps in
// flight: a lone warp is bound by its dependency chain, not by issue slots).
template <int LPN, int K>
__device__ __forceinline__ void node_allreduce_b(double (&v)[K])
{
    if (LPN == 2) {
        // columns = lane / 2; tree: neighbours, quads, eights, halves
#pragma unroll
        for (int off = 2; off <= 16; off <<= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(GIK_FULL_MASK, v[k], off, 32);
        }
    } else {
        // node_allreduce_s reads 32 columns interleaved: (c, c^1), then c^8, c^16, then the parts c^2, c^4
        constexpr int order[5] = {1, 8, 16, 2, 4};
#pragma unroll
        for (int l = 0; l < 5; ++l) {
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(GIK_FULL_MASK, v[k], order[l], 32);
        }
    }
}
