// gik_costs.cu -- batched drop-ins for the reference's `costgrd` leaves and
// PSDFixedRank.proj, plus the goal-dependent distance assembly.
//
// These are the streaming forms of the hot-path operators (state in HBM, one
// pass per call): read Y (and W), write f / g / HW.  The persistent solver
// (gik_rtr.cu) inlines the very same device functions (gik_common.cuh) with
// the state held in registers.
#include "gik_common.cuh"

namespace {

constexpr int kThreads = 128;

struct PlanView {
    const uint32_t *slot_info;
    const double *slot_target;
    const int32_t *deg;
    int N, n_goal;
};

PlanView view_of(const GikPlan *p)
{
    PlanView v;
    v.slot_info = p->slot_info;
    v.slot_target = p->slot_target;
    v.deg = p->deg;
    v.N = p->N;
    v.n_goal = p->n_goal;
    return v;
}

template <int W, int NPL>
__device__ __forceinline__ void load_points(const double *src, int N, int lane, double (&x)[NPL][3])
{
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
        const int i = lane + W * m;
        if (i < N) {
            x[m][0] = src[3 * i];
            x[m][1] = src[3 * i + 1];
            x[m][2] = src[3 * i + 2];
        } else {
            x[m][0] = x[m][1] = x[m][2] = 0.0;
        }
    }
}

template <int W, int NPL>
__device__ __forceinline__ void store_points(double *dst, int N, int lane, const double (&x)[NPL][3])
{
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
        const int i = lane + W * m;
        if (i < N) {
            dst[3 * i] = x[m][0];
            dst[3 * i + 1] = x[m][1];
            dst[3 * i + 2] = x[m][2];
        }
    }
}

// mode 0: cost (+ gradient), mode 1: Hessian-vector product
template <int W, int NPL, int MODE>
__global__ void __launch_bounds__(kThreads)
k_costs(PlanView pv, const double *__restrict__ Y, const double *__restrict__ Wd,
        const double *__restrict__ goal_d2, int B, double *__restrict__ f, double *__restrict__ out)
{
    constexpr int NP = W * NPL;
    constexpr int GPB = kThreads / W;
    extern __shared__ double smem[];
    const int gid = threadIdx.x / W;
    const int lane = threadIdx.x % W;
    const int goal_pad = (pv.n_goal + 1) & ~1;
    double *base = smem + (size_t)gid * (6 * NP + goal_pad);
    GikGroupCtx c;
    c.slot_info = pv.slot_info;
    c.slot_target = pv.slot_target;
    c.deg = pv.deg;
    c.P = base;
    c.V = base + 3 * NP;
    double *goal = base + 6 * NP;
    c.goal = goal;
    c.N = pv.N;
    c.mask = gik_group_mask<W>();
    c.lane = lane;
    const int stride = 3 * pv.N;
    for (int b = blockIdx.x * GPB + gid; b < B; b += gridDim.x * GPB) {
        double x[NPL][3], w[NPL][3], r[NPL][3];
        load_points<W, NPL>(Y + (size_t)b * stride, pv.N, lane, x);
        __syncwarp(c.mask);   // the group has finished reading the previous problem's buffers
        for (int k = lane; k < pv.n_goal; k += W) goal[k] = goal_d2[(size_t)b * pv.n_goal + k];
        gik_publish<W, NPL>(c.P, x, lane);
        if (MODE == 1) {
            load_points<W, NPL>(Wd + (size_t)b * stride, pv.N, lane, w);
            gik_publish<W, NPL>(c.V, w, lane);
        }
        __syncwarp(c.mask);
        if (MODE == 0) {
            double v[1] = {gik_pass_cost_grad<W, NPL>(c, x, r)};
            gik_allreduce<W, 1>(v, c.mask);
            if (f && lane == 0) f[b] = v[0];
            if (out) store_points<W, NPL>(out + (size_t)b * stride, pv.N, lane, r);
        } else {
            gik_pass_hess<W, NPL>(c, x, w, r);
            store_points<W, NPL>(out + (size_t)b * stride, pv.N, lane, r);
        }
        __syncwarp(c.mask);
    }
}

// Same operators for N <= 32 with the lane-centric tables of the solver's latency kernel: one warp per
// problem, LPN lanes per node, the slot description of a lane lives in registers for the whole launch,
// so a problem costs one coalesced read of Y (and W), the edge pass, and one coalesced write.
template <int LPN, int SPL, int MODE>
__global__ void __launch_bounds__(kThreads)
k_costs_fast(int N, int n_goal, const uint32_t *__restrict__ fast_info, const double *__restrict__ fast_target,
             const double *__restrict__ Y, const double *__restrict__ Wd, const double *__restrict__ goal_d2,
             int B, double *__restrict__ f, double *__restrict__ out)
{
    constexpr int NPW = 32 / LPN;
    constexpr int WPB = kThreads / 32;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int node = lane / LPN;
    const bool valid = node < N, owner = valid && (lane % LPN == 0);
    const int goal_pad = (n_goal + 1) & ~1;
    double *P = smem + (size_t)warp * (6 * NPW + goal_pad);
    double *V = P + 3 * NPW;
    double *goal = V + 3 * NPW;
    uint32_t info[SPL];
    double tstat[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
        info[s] = fast_info[s * 32 + lane];
        tstat[s] = fast_target[s * 32 + lane];
    }
    for (int b = blockIdx.x * WPB + warp; b < B; b += gridDim.x * WPB) {
        double x[3] = {0.0, 0.0, 0.0}, w[3] = {0.0, 0.0, 0.0};
        if (valid) {
            const double *src = Y + ((size_t)b * N + node) * 3;
            x[0] = src[0]; x[1] = src[1]; x[2] = src[2];
            if (MODE == 1) {
                const double *sw = Wd + ((size_t)b * N + node) * 3;
                w[0] = sw[0]; w[1] = sw[1]; w[2] = sw[2];
            }
        }
        for (int k = lane; k < n_goal; k += 32) goal[k] = goal_d2[(size_t)b * n_goal + k];
        if (lane % LPN == 0) {
            P[node] = x[0]; P[NPW + node] = x[1]; P[2 * NPW + node] = x[2];
            if (MODE == 1) { V[node] = w[0]; V[NPW + node] = w[1]; V[2 * NPW + node] = w[2]; }
        }
        __syncwarp();
        double acc[4] = {0.0, 0.0, 0.0, 0.0};   // x, y, z of the node sum; cost share
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const int j = GIK_SLOT_NBR(info[s]);
            const uint32_t kind = GIK_SLOT_KIND(info[s]);
            const uint32_t gs = GIK_SLOT_GOAL(info[s]);
            const double T = gs ? goal[gs - 1] : tstat[s];
            const double dx = x[0] - P[j], dy = x[1] - P[NPW + j], dz = x[2] - P[2 * NPW + j];
            const double d = gik_sqdist(dx, dy, dz);
            double r = d - T;
            const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (r < 0.0)) |
                             ((kind == GIK_TERM_UP) & (r > 0.0));
            r = act ? r : 0.0;
            if (MODE == 0) {
                acc[3] = fma(r, r, acc[3]);
                acc[0] = fma(r, dx, acc[0]);
                acc[1] = fma(r, dy, acc[1]);
                acc[2] = fma(r, dz, acc[2]);
            } else {
                const double wx = w[0] - V[j], wy = w[1] - V[NPW + j], wz = w[2] - V[2 * NPW + j];
                const double a2 = act ? 2.0 * (dx * wx + dy * wy + dz * wz) : 0.0;
                acc[0] = fma(a2, dx, fma(r, wx, acc[0]));
                acc[1] = fma(a2, dy, fma(r, wy, acc[1]));
                acc[2] = fma(a2, dz, fma(r, wz, acc[2]));
            }
        }
#pragma unroll
        for (int off = LPN / 2; off > 0; off >>= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += __shfl_xor_sync(GIK_FULL_MASK, acc[k], off, 32);
        }
        if (MODE == 0 && f) {
            double fs = 0.5 * acc[3];          // pairs already combined: count each node once
#pragma unroll
            for (int off = 16; off >= LPN; off >>= 1) fs += __shfl_xor_sync(GIK_FULL_MASK, fs, off, 32);
            if (lane == 0) f[b] = fs;
        }
        if (out && owner) {
            double *dst = out + ((size_t)b * N + node) * 3;
            dst[0] = 2.0 * acc[0]; dst[1] = 2.0 * acc[1]; dst[2] = 2.0 * acc[2];
        }
        __syncwarp();
    }
}

// Same operators for N <= 16 with the node-centric tables of the two-problems-per-warp solver kernel: a half-warp per
// problem (lane = node), the slot description of a lane in registers for the whole launch, the loop over the slots fully
// unrolled, and the NEXT pair of problems loaded into registers before the current one is evaluated.  ncu on the generic
// kernel showed these operators bound by instruction issue (~420 warp instructions per problem: slot decode from
// global tables, run-time loop, address arithmetic), not by HBM; this form needs ~150.
// (launch bounds measured on UR10: cost + gradient 0.40 of the HBM peak without a register cap, 0.43 at 80 registers / six
// CTAs per SM; Hessian-vector 0.52 without a cap, 0.43 with it)
template <int SPL, int MODE>
__global__ void __launch_bounds__(kThreads, MODE == 0 ? 6 : 0)
k_costs_duo(int N, int n_goal, const uint32_t *__restrict__ duo_info, const double *__restrict__ duo_target,
            const double *__restrict__ Y, const double *__restrict__ Wd, const double *__restrict__ goal_d2,
            int B, double *__restrict__ f, double *__restrict__ out)
{
    constexpr int WPB = kThreads / 32;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, node = lane & 15;
    const bool valid = node < N;
    const int goal_pad = (n_goal + 1) & ~1;
    double *P = smem + (size_t)(warp * 2 + half) * (96 + goal_pad);   // [3][16] of this half's problem
    double *V = P + 48;
    double *goal = V + 48;
    int nbr[SPL];
    uint32_t kg[SPL];          // kind | goal slot + 1 << 2
    double tstat[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
        const uint32_t info = duo_info[s * 16 + node];
        nbr[s] = GIK_SLOT_NBR(info);
        kg[s] = GIK_SLOT_KIND(info) | (GIK_SLOT_GOAL(info) << 2);
        tstat[s] = duo_target[s * 16 + node];
    }
    const int step = gridDim.x * WPB;
    int pb = blockIdx.x * WPB + warp;                 // pair of problems 2 pb, 2 pb + 1
    double xn[3] = {0.0, 0.0, 0.0}, wn[3] = {0.0, 0.0, 0.0};
    auto fetch = [&](int pair) {
        const long long b = 2LL * pair + half;
        if (b < B && valid) {
            const double *src = Y + ((size_t)b * N + node) * 3;
            xn[0] = src[0]; xn[1] = src[1]; xn[2] = src[2];
            if (MODE == 1) {
                const double *sw = Wd + ((size_t)b * N + node) * 3;
                wn[0] = sw[0]; wn[1] = sw[1]; wn[2] = sw[2];
            }
        }
    };
    fetch(pb);
    for (; 2LL * pb < B; pb += step) {
        const long long b = 2LL * pb + half;
        const bool active = b < B;
        const double x[3] = {xn[0], xn[1], xn[2]}, w[3] = {wn[0], wn[1], wn[2]};
        __syncwarp();                                  // the previous pair's readers are done
        P[node] = x[0]; P[16 + node] = x[1]; P[32 + node] = x[2];
        if (MODE == 1) { V[node] = w[0]; V[16 + node] = w[1]; V[32 + node] = w[2]; }
        if (active)
            for (int k = node; k < n_goal; k += 16) goal[k] = goal_d2[(size_t)b * n_goal + k];
        fetch(pb + step);                              // in flight while this pair is evaluated
        __syncwarp();
        double acc[4] = {0.0, 0.0, 0.0, 0.0};          // x, y, z of the node sum; cost share
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const int j = nbr[s];
            const uint32_t kind = kg[s] & 3u, gs = kg[s] >> 2;
            const double T = gs ? goal[gs - 1] : tstat[s];
            const double dx = x[0] - P[j], dy = x[1] - P[16 + j], dz = x[2] - P[32 + j];
            const double d = gik_sqdist(dx, dy, dz);
            double r = d - T;
            const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (r < 0.0)) |
                             ((kind == GIK_TERM_UP) & (r > 0.0));
            r = act ? r : 0.0;
            if (MODE == 0) {
                acc[3] = fma(r, r, acc[3]);
                acc[0] = fma(r, dx, acc[0]);
                acc[1] = fma(r, dy, acc[1]);
                acc[2] = fma(r, dz, acc[2]);
            } else {
                const double wx = w[0] - V[j], wy = w[1] - V[16 + j], wz = w[2] - V[32 + j];
                const double a2 = act ? 2.0 * (dx * wx + dy * wy + dz * wz) : 0.0;
                acc[0] = fma(a2, dx, fma(r, wx, acc[0]));
                acc[1] = fma(a2, dy, fma(r, wy, acc[1]));
                acc[2] = fma(a2, dz, fma(r, wz, acc[2]));
            }
        }
        if (MODE == 0 && f) {
            double fs = 0.5 * acc[3];                  // every undirected term is seen from both ends
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) fs += __shfl_xor_sync(GIK_FULL_MASK, fs, off, 32);
            if (node == 0 && active) f[b] = fs;
        }
        if (out && valid && active) {
            double *dst = out + ((size_t)b * N + node) * 3;
            dst[0] = 2.0 * acc[0]; dst[1] = 2.0 * acc[1]; dst[2] = 2.0 * acc[2];
        }
    }
}

// fixed_rank_psd_sym.py:91-113 via the 3x3 form: with X = Y^T Y and c = sum_i Z_i x Y_i,
// (tr(X) I - X) omega = c and proj(Z)_i = Z_i - Y_i x omega.
template <int W, int NPL>
__global__ void __launch_bounds__(kThreads)
k_proj(int N, const double *__restrict__ Y, const double *__restrict__ Z, int B, double *__restrict__ out)
{
    constexpr int GPB = kThreads / W;
    const int gid = threadIdx.x / W;
    const int lane = threadIdx.x % W;
    const unsigned mask = gik_group_mask<W>();
    const int stride = 3 * N;
    for (int b = blockIdx.x * GPB + gid; b < B; b += gridDim.x * GPB) {
        double y[NPL][3], z[NPL][3];
        load_points<W, NPL>(Y + (size_t)b * stride, N, lane, y);
        load_points<W, NPL>(Z + (size_t)b * stride, N, lane, z);
        double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int m = 0; m < NPL; ++m) {
            v[0] = fma(y[m][0], y[m][0], v[0]);
            v[1] = fma(y[m][0], y[m][1], v[1]);
            v[2] = fma(y[m][0], y[m][2], v[2]);
            v[3] = fma(y[m][1], y[m][1], v[3]);
            v[4] = fma(y[m][1], y[m][2], v[4]);
            v[5] = fma(y[m][2], y[m][2], v[5]);
            v[6] += z[m][1] * y[m][2] - z[m][2] * y[m][1];
            v[7] += z[m][2] * y[m][0] - z[m][0] * y[m][2];
            v[8] += z[m][0] * y[m][1] - z[m][1] * y[m][0];
        }
        gik_allreduce<W, 9>(v, mask);
        double Mi[6], om[3];
        gik_sylvester_inverse(v, Mi);
        gik_sym_mul(Mi, v + 6, om);
#pragma unroll
        for (int m = 0; m < NPL; ++m) {
            z[m][0] -= y[m][1] * om[2] - y[m][2] * om[1];
            z[m][1] -= y[m][2] * om[0] - y[m][0] * om[2];
            z[m][2] -= y[m][0] * om[1] - y[m][1] * om[0];
        }
        store_points<W, NPL>(out + (size_t)b * stride, N, lane, z);
    }
}

// graph_revolute.py:243-249 + dgp.py:139 for the goal-dependent entries.
__global__ void k_goal_distances(const double *__restrict__ T_goal, int B, int n_anchor,
                                 const double *__restrict__ anchor_pos, double axis_length,
                                 double *__restrict__ goal_d2)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int G = 2 * n_anchor;
    if (idx >= B * G) return;
    const int b = idx / G, s = idx % G;
    const int a = s % n_anchor;
    const double *T = T_goal + (size_t)b * 16;
    double p[3] = {T[3], T[7], T[11]};
    if (s >= n_anchor) {
        // q_n = R e_z * axis_length + t, evaluated without contraction like the reference
        p[0] = __dadd_rn(__dmul_rn(T[2], axis_length), p[0]);
        p[1] = __dadd_rn(__dmul_rn(T[6], axis_length), p[1]);
        p[2] = __dadd_rn(__dmul_rn(T[10], axis_length), p[2]);
    }
    const double dx = p[0] - anchor_pos[3 * a], dy = p[1] - anchor_pos[3 * a + 1],
                 dz = p[2] - anchor_pos[3 * a + 2];
    // DIST = ||.|| (dgp.py:139), D_goal = DIST ** 2 (dgp.py:50)
    const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    goal_d2[idx] = __dmul_rn(dist, dist);
}

size_t costs_smem(int W, int NPL, int n_goal)
{
    const int NP = W * NPL, GPB = kThreads / W, goal_pad = (n_goal + 1) & ~1;
    return (size_t)GPB * (6 * NP + goal_pad) * sizeof(double);
}

template <int W, int NPL, int MODE>
int launch_costs(const GikPlan *p, const double *Y, const double *Wd, const double *goal_d2, int B,
                 double *f, double *out, cudaStream_t st)
{
    constexpr int GPB = kThreads / W;
    const size_t smem = costs_smem(W, NPL, p->n_goal);
    if (smem > 48 * 1024)
        GIK_CUDA(cudaFuncSetAttribute(k_costs<W, NPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (B + GPB - 1) / GPB;
    const int cap = p->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_costs<W, NPL, MODE><<<blocks, kThreads, smem, st>>>(view_of(p), Y, Wd, goal_d2, B, f, out);
    return gik_check_cuda(cudaGetLastError(), "k_costs launch");
}

template <int LPN, int SPL, int MODE>
int launch_costs_fast(const GikPlan *p, const double *Y, const double *Wd, const double *goal_d2, int B,
                      double *f, double *out, cudaStream_t st)
{
    constexpr int NPW = 32 / LPN, WPB = kThreads / 32;
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t smem = (size_t)WPB * (6 * NPW + goal_pad) * sizeof(double);
    if (smem > 48 * 1024)
        GIK_CUDA(cudaFuncSetAttribute(k_costs_fast<LPN, SPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (B + WPB - 1) / WPB;
    const int cap = p->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_costs_fast<LPN, SPL, MODE><<<blocks, kThreads, smem, st>>>(p->N, p->n_goal, p->fast_info, p->fast_target, Y, Wd,
                                                                goal_d2, B, f, out);
    return gik_check_cuda(cudaGetLastError(), "k_costs_fast launch");
}

template <int SPL, int MODE>
int launch_costs_duo(const GikPlan *p, const double *Y, const double *Wd, const double *goal_d2, int B,
                     double *f, double *out, cudaStream_t st)
{
    constexpr int WPB = kThreads / 32;
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t smem = (size_t)WPB * 2 * (96 + goal_pad) * sizeof(double);
    if (smem > 48 * 1024) return 1;
    const int pairs = (B + 1) / 2;
    int blocks = (pairs + WPB - 1) / WPB;
    const int cap = p->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_costs_duo<SPL, MODE><<<blocks, kThreads, smem, st>>>(p->N, p->n_goal, p->duo_info, p->duo_target, Y, Wd, goal_d2, B, f, out);
    return gik_check_cuda(cudaGetLastError(), "k_costs_duo launch");
}

template <int MODE>
int dispatch_costs(const GikPlan *p, const double *Y, const double *Wd, const double *goal_d2, int B,
                   double *f, double *out, cudaStream_t st)
{
    if (p->duo_info) {      // N <= 16: half a warp per problem, slot tables in registers
        int rc = 1;
        const int d = p->maxdeg;
        if (d <= 6) rc = launch_costs_duo<6, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        else if (d <= 9) rc = launch_costs_duo<9, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        else if (d <= 12) rc = launch_costs_duo<12, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        if (rc <= 0) return rc;
    }
    // 17..32 nodes: warp per problem with register-resident slot tables (KUKA: 1.4x the group kernel);
    // N <= 16 stays on the 16-lane groups, which put two problems in a warp (measured faster there)
    if (p->fast_info && p->fast_LPN == 1) {
        if (p->fast_SPL <= 9) return launch_costs_fast<1, 9, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        if (p->fast_SPL <= 12) return launch_costs_fast<1, 12, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
    }
    if (p->W == 16) return launch_costs<16, 1, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
    switch (p->NPL) {
        case 1: return launch_costs<32, 1, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        case 2: return launch_costs<32, 2, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        case 4: return launch_costs<32, 4, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        case 8: return launch_costs<32, 8, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        default: return launch_costs<32, 15, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
    }
}

template <int W, int NPL>
int launch_proj(int N, const double *Y, const double *Z, int B, double *out, cudaStream_t st)
{
    constexpr int GPB = kThreads / W;
    int blocks = (B + GPB - 1) / GPB;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_proj<W, NPL><<<blocks, kThreads, 0, st>>>(N, Y, Z, B, out);
    return gik_check_cuda(cudaGetLastError(), "k_proj launch");
}

int check_plan_device(const GikPlan *p, const char *fn)
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

extern "C" int gik_goal_distances(const GikPlan *p, const double *T_goal, int32_t B, double *goal_d2, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !T_goal || !goal_d2 || B < 0) { gik_set_error("gik_goal_distances: bad argument"); return GIK_EINVAL; }
    if (p->n_goal != 2 * p->n_anchor) { gik_set_error("gik_goal_distances: plan has no pose-goal layout"); return GIK_EINVAL; }
    if (int rc = check_plan_device(p, "gik_goal_distances")) return rc;
    if (B == 0) return GIK_OK;
    const int total = B * p->n_goal;
    k_goal_distances<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(T_goal, B, p->n_anchor, p->anchor_pos,
                                                                         p->axis_length, goal_d2);
    return gik_check_cuda(cudaGetLastError(), "k_goal_distances launch");
}

extern "C" int gik_cost_grad(const GikPlan *p, const double *Y, const double *goal_d2, int32_t B, double *f,
                             double *g, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !Y || B < 0 || (p->n_goal > 0 && !goal_d2)) { gik_set_error("gik_cost_grad: bad argument"); return GIK_EINVAL; }
    if (int rc = check_plan_device(p, "gik_cost_grad")) return rc;
    if (B == 0 || (!f && !g)) return GIK_OK;
    return dispatch_costs<0>(p, Y, nullptr, goal_d2, B, f, g, (cudaStream_t)stream);
}

extern "C" int gik_hessvec(const GikPlan *p, const double *Y, const double *W, const double *goal_d2, int32_t B,
                           double *HW, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !Y || !W || !HW || B < 0 || (p->n_goal > 0 && !goal_d2)) { gik_set_error("gik_hessvec: bad argument"); return GIK_EINVAL; }
    if (int rc = check_plan_device(p, "gik_hessvec")) return rc;
    if (B == 0) return GIK_OK;
    return dispatch_costs<1>(p, Y, W, goal_d2, B, nullptr, HW, (cudaStream_t)stream);
}

extern "C" int gik_proj(int32_t N, const double *Y, const double *Z, int32_t B, double *out, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!Y || !Z || !out || B < 0 || N < 2) { gik_set_error("gik_proj: bad argument"); return GIK_EINVAL; }
    if (N > 480) { gik_set_error("gik_proj: N=%d exceeds the compiled limit of 480", N); return GIK_ELIMIT; }
    if (B == 0) return GIK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 16) return launch_proj<16, 1>(N, Y, Z, B, out, st);
    if (N <= 32) return launch_proj<32, 1>(N, Y, Z, B, out, st);
    if (N <= 64) return launch_proj<32, 2>(N, Y, Z, B, out, st);
    if (N <= 128) return launch_proj<32, 4>(N, Y, Z, B, out, st);
    if (N <= 256) return launch_proj<32, 8>(N, Y, Z, B, out, st);
    return launch_proj<32, 15>(N, Y, Z, B, out, st);
}
