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

template <int MODE>
int dispatch_costs(const GikPlan *p, const double *Y, const double *Wd, const double *goal_d2, int B,
                   double *f, double *out, cudaStream_t st)
{
    if (p->W == 16) return launch_costs<16, 1, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
    switch (p->NPL) {
        case 1: return launch_costs<32, 1, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        case 2: return launch_costs<32, 2, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
        default: return launch_costs<32, 4, MODE>(p, Y, Wd, goal_d2, B, f, out, st);
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
    if (N > 128) { gik_set_error("gik_proj: N=%d exceeds the compiled limit of 128", N); return GIK_ELIMIT; }
    if (B == 0) return GIK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 16) return launch_proj<16, 1>(N, Y, Z, B, out, st);
    if (N <= 32) return launch_proj<32, 1>(N, Y, Z, B, out, st);
    if (N <= 64) return launch_proj<32, 2>(N, Y, Z, B, out, st);
    return launch_proj<32, 4>(N, Y, Z, B, out, st);
}
