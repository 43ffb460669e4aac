// gik_rtr_duo.cu -- throughput-optimised trust-region solve for graphs with N <= 16 nodes:
// TWO problems per warp advancing in lock-step.
//
// Same algorithm and arithmetic conventions as k_rtr_fast / k_rtr (reference
// trust_region.py:112-599, costs.py:79-207, fixed_rank_psd_sym.py:91-137).  k_rtr_fast spends a
// whole warp on one problem (best latency for the slowest problem of a batch); when many problems
// are queued the limiter is instruction issue (ncu, UR10 B = 65536: 381 warp instructions per inner
// iteration, IPC 0.47 per scheduler).  Here each half-warp owns one problem, lane <-> node, and the
// inner tCG iteration is written as straight-line code that BOTH halves execute together, so one
// instruction stream serves two problems:
//
//   * every shuffle in the inner iteration uses xor offsets <= 8 and the full mask, hence stays
//     inside a half and is executed once per warp;
//   * the three ways a tCG iteration can end (boundary / negative curvature, model increase,
//     target reached) only set the half's phase; state commits are per-lane predicated;
//   * what happens once per outer iteration (proposal cost/gradient, rho test, radius update,
//     accept/reject, start of the next subproblem) and once per problem (fetch from the work
//     queue, initial cost/gradient, final store) runs in short divergent sections guarded by a
//     warp vote, ~2 of every 70 ticks per half;
//   * the x-dependent part of each edge term (2 act D, 2 act r) is cached per accepted iterate in
//     SHARED memory ([slot][lane], conflict-free) instead of registers: with one lane per node a lane
//     has up to 12 slots, and the loads do not depend on delta, so they are off the critical path.
//
// A half whose phase is not INNER executes the inner block on dead state; nothing it computes is
// committed or stored.
#include <cstdlib>

#include "gik_rtr.cuh"

namespace {

constexpr int kThreads = 32;
enum { PH_NEED_PROBLEM = 0, PH_NEED_OUTER = 1, PH_INNER = 2, PH_IDLE = 3 };

// transposed butterfly over the 16 lanes of each half (offsets 8, 4, 2, 1), full-warp mask
template <int KP>
__device__ __forceinline__ void half_allreduce_t(double (&v)[KP], int lane)
{
    static_assert(KP == 4 || KP == 8, "KP must be 4 or 8");
    double cur[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) cur[k] = v[k];
    int cnt = KP;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
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
    const int base = lane & 16;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int src = (KP == 8) ? (((k >> 2) & 1) * 8 + ((k >> 1) & 1) * 4 + (k & 1) * 2)
                                  : (((k >> 1) & 1) * 8 + (k & 1) * 4);
        v[k] = __shfl_sync(GIK_FULL_MASK, cur[0], base + src, 32);
    }
}

// plain butterfly inside one half with the half's own mask (divergent sections)
template <int K>
__device__ __forceinline__ void half_allreduce(double (&v)[K], unsigned gmask)
{
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(gmask, v[k], off, 32);
    }
}

template <int SPL>
__global__ void __launch_bounds__(kThreads, 12) k_rtr_duo(const RtrArgs a, const uint32_t *__restrict__ duo_info,
                                                          const double *__restrict__ duo_target)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int node = lane & 15;
    const int hbase = lane & 16;                       // first lane of this half
    const unsigned gmask = hbase ? 0xffff0000u : 0x0000ffffu;
    const bool valid = node < a.N;
    const int goal_pad = (a.n_goal + 1) & ~1;
    // per-warp shared memory: P[3][32], V[3][32], goal[2][goal_pad], tgt[SPL][32], cache[SPL][4][32]
    double *P = smem;
    double *V = P + 96;
    double *goal = V + 96 + (hbase ? goal_pad : 0);
    double *tgt = V + 96 + 2 * goal_pad + lane;
    double *cache = V + 96 + 2 * goal_pad + SPL * 32 + lane;
    const GikSolveOpts &o = a.o;
    const double eps = 2.220446049250313e-16;  // np.spacing(1), trust_region.py:293

    // static slot description of this lane; neighbour indices are rebased into this half's buffers
    uint32_t info[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) info[s] = duo_info[s * 16 + node];

    double x[3] = {0, 0, 0}, g[3] = {0, 0, 0}, eta[3] = {0, 0, 0}, Heta[3] = {0, 0, 0}, r[3] = {0, 0, 0},
           dl[3] = {0, 0, 0}, Hd[3];
    double fx = 0, gg = 0, norm_grad = 0, Mi[6] = {0, 0, 0, 0, 0, 0}, Delta = 0, Delta2 = 0;
    double e_Pe = 0, r_r = 1, z_r = 1, inv_z_r = 1, d_Pd = 1, e_Pd = 0, model_value = 0, r_target2 = 0, pw = 0;
    int b = 0, phase = PH_NEED_PROBLEM, k_outer = 0, inner_total = 0, status = 0, stop = MAX_INNER_ITER, j = 0,
        numit = 0;

    // cost / gradient at p (published in P) + rebuild of the slot cache; returns this lane's cost share
    auto rebuild = [&](const double (&p)[3], double (&gout)[3]) -> double {
        double fpart = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const int jn = hbase + (int)GIK_SLOT_NBR(info[s]);
            const uint32_t kind = GIK_SLOT_KIND(info[s]);
            const double dx = p[0] - P[jn], dy = p[1] - P[32 + jn], dz = p[2] - P[64 + jn];
            const double d = dx * dx + dy * dy + dz * dz;
            double rr = d - tgt[s * 32];
            const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (rr < 0.0)) |
                             ((kind == GIK_TERM_UP) & (rr > 0.0));
            rr = act ? rr : 0.0;
            fpart = fma(rr, rr, fpart);
            gx = fma(rr, dx, gx);
            gy = fma(rr, dy, gy);
            gz = fma(rr, dz, gz);
            const double two = act ? 2.0 : 0.0;
            cache[(s * 4 + 0) * 32] = two * dx;
            cache[(s * 4 + 1) * 32] = two * dy;
            cache[(s * 4 + 2) * 32] = two * dz;
            cache[(s * 4 + 3) * 32] = 2.0 * rr;
        }
        gout[0] = 2.0 * gx; gout[1] = 2.0 * gy; gout[2] = 2.0 * gz;
        return 0.5 * fpart;
    };

    // start of a trust-region subproblem (trust_region.py:436-490), eta0 = 0, precon = identity
    auto start_tcg = [&]() {
#pragma unroll
        for (int q = 0; q < 3; ++q) { eta[q] = 0.0; Heta[q] = 0.0; r[q] = g[q]; dl[q] = -g[q]; }
        e_Pe = 0.0;
        r_r = gg;
        const double norm_r0 = sqrt(r_r);
        z_r = r_r; d_Pd = r_r; e_Pd = 0.0; model_value = 0.0;
        pw = o.theta == 1.0 ? norm_r0 : pow(norm_r0, o.theta);
        const double r_target = norm_r0 * fmin(pw, o.kappa);
        r_target2 = r_target * r_target;
        Delta2 = Delta * Delta;
        inv_z_r = gik_rcp(z_r);
        stop = MAX_INNER_ITER;
        j = 0;
    };

    auto store_result = [&]() {
        if (valid) {
            double *dst = a.Y_out + ((size_t)b * a.N + node) * 3;
            dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
        }
        if (node == 0) {
            a.f[b] = fx;
            a.gradnorm[b] = norm_grad;
            a.iters[b] = k_outer;
            a.status[b] = status;
            if (a.n_inner) a.n_inner[b] = inner_total;
        }
    };

    for (;;) {
        // ================= A. halves without a problem pull the next one from the queue
        if (__any_sync(GIK_FULL_MASK, phase == PH_NEED_PROBLEM)) {
            if (phase == PH_NEED_PROBLEM) {
                int nb = 0;
                if (node == 0) nb = atomicAdd(a.work_counter, 1);
                b = __shfl_sync(gmask, nb, hbase, 32);
                if (b >= a.B) {
                    phase = PH_IDLE;
                } else {
                    x[0] = x[1] = x[2] = 0.0;
                    if (valid) {
                        const double *src = a.Y_init + ((size_t)b * a.N + node) * 3;
                        x[0] = src[0]; x[1] = src[1]; x[2] = src[2];
                    }
                    for (int k = node; k < a.n_goal; k += 16) goal[k] = a.goal_d2[(size_t)b * a.n_goal + k];
                    P[lane] = x[0]; P[32 + lane] = x[1]; P[64 + lane] = x[2];
                    __syncwarp(gmask);
#pragma unroll
                    for (int s = 0; s < SPL; ++s) {
                        const uint32_t gs = GIK_SLOT_GOAL(info[s]);
                        tgt[s * 32] = gs ? goal[gs - 1] : duo_target[s * 16 + node];
                    }
                    double v[8];
                    v[0] = rebuild(x, g);
                    v[1] = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
                    v[2] = x[0] * x[0]; v[3] = x[0] * x[1]; v[4] = x[0] * x[2];
                    v[5] = x[1] * x[1]; v[6] = x[1] * x[2]; v[7] = x[2] * x[2];
                    half_allreduce<8>(v, gmask);
                    fx = v[0];
                    gg = v[1];
                    gik_sylvester_inverse(v + 2, Mi);
                    norm_grad = sqrt(gg);
                    Delta = o.Delta0;
                    k_outer = 0;
                    inner_total = 0;
                    if (!(isfinite(fx) && isfinite(gg))) {
                        status = GIK_STATUS_NAN;
                        store_result();          // stays PH_NEED_PROBLEM: fetch again next tick
                    } else {
                        status = GIK_STATUS_MAXITER;
                        start_tcg();
                        phase = PH_INNER;
                    }
                }
            }
            __syncwarp();
        }
        if (__all_sync(GIK_FULL_MASK, phase == PH_IDLE)) break;

        // ================= B. halves whose subproblem ended: proposal, rho test, accept / reject
        if (__any_sync(GIK_FULL_MASK, phase == PH_NEED_OUTER)) {
            if (phase == PH_NEED_OUTER) {
                double xp[3], gp[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) xp[q] = x[q] + eta[q];
                P[lane] = xp[0]; P[32 + lane] = xp[1]; P[64 + lane] = xp[2];
                __syncwarp(gmask);
                double v[10];
                v[0] = rebuild(xp, gp);
                v[1] = g[0] * eta[0] + g[1] * eta[1] + g[2] * eta[2];
                v[2] = eta[0] * Heta[0] + eta[1] * Heta[1] + eta[2] * Heta[2];
                v[3] = gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2];
                v[4] = xp[0] * xp[0]; v[5] = xp[0] * xp[1]; v[6] = xp[0] * xp[2];
                v[7] = xp[1] * xp[1]; v[8] = xp[1] * xp[2]; v[9] = xp[2] * xp[2];
                half_allreduce<10>(v, gmask);
                const double fx_prop = v[0];
                double rhonum = fx - fx_prop;
                double rhoden = -v[1] - 0.5 * v[2];
                const double rho_reg = fmax(1.0, fabs(fx)) * eps * o.rho_regularization;
                rhonum += rho_reg;
                rhoden += rho_reg;
                const bool model_decreased = rhoden >= 0.0;
                const double rho = rhonum / rhoden;
                const double Delta_used = Delta;
                if (rho < 0.25 || !model_decreased || isnan(rho)) {
                    Delta = Delta / 4.0;
                } else if (rho > 0.75 && (stop == NEGATIVE_CURVATURE || stop == EXCEEDED_TR)) {
                    Delta = fmin(2.0 * Delta, o.Delta_bar);
                }
                const bool accept = model_decreased && rho > o.rho_prime;
                if (accept) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { x[q] = xp[q]; g[q] = gp[q]; }
                    fx = fx_prop;
                    gg = v[3];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 4, Mi);
                } else {
                    __syncwarp(gmask);
                    P[lane] = x[0]; P[32 + lane] = x[1]; P[64 + lane] = x[2];
                    __syncwarp(gmask);
                    double gtmp[3];
                    rebuild(x, gtmp);
                }
                if (a.trace && k_outer < a.trace_rows && node == 0) {
                    double *row = a.trace + ((size_t)b * a.trace_rows + k_outer) * 6;
                    row[0] = Delta_used;
                    row[1] = (double)numit;
                    row[2] = (double)stop;
                    row[3] = fx_prop;
                    row[4] = accept ? 1.0 : 0.0;
                    row[5] = accept ? norm_grad : nan("");
                }
                ++k_outer;
                // pymanopt Solver._check_stopping_criterion: maxiter before mingradnorm
                if (k_outer >= o.maxiter) {
                    status = GIK_STATUS_MAXITER;
                    store_result();
                    phase = PH_NEED_PROBLEM;
                } else if (norm_grad < o.mingradnorm) {
                    status = GIK_STATUS_CONVERGED;
                    store_result();
                    phase = PH_NEED_PROBLEM;
                } else {
                    start_tcg();
                    phase = PH_INNER;
                }
            }
            __syncwarp();
        }

        // ================= C. one tCG iteration (trust_region.py:495-597), both halves together
        V[lane] = dl[0]; V[32 + lane] = dl[1]; V[64 + lane] = dl[2];
        __syncwarp();
        double z[3] = {0.0, 0.0, 0.0}, zb[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const int jn = hbase + (int)GIK_SLOT_NBR(info[s]);
            const double cx = cache[(s * 4 + 0) * 32], cy = cache[(s * 4 + 1) * 32], cz = cache[(s * 4 + 2) * 32];
            const double c2 = cache[(s * 4 + 3) * 32];
            const double wx = dl[0] - V[jn], wy = dl[1] - V[32 + jn], wz = dl[2] - V[64 + jn];
            const double t = fma(cx, wx, fma(cy, wy, cz * wz));
            z[0] = fma(c2, wx, z[0]);
            z[1] = fma(c2, wy, z[1]);
            z[2] = fma(c2, wz, z[2]);
            zb[0] = fma(t, cx, zb[0]);
            zb[1] = fma(t, cy, zb[1]);
            zb[2] = fma(t, cz, zb[2]);
        }
        z[0] += zb[0]; z[1] += zb[1]; z[2] += zb[2];
        double v[8];
        v[7] = 0.0;
        v[0] = dl[0] * z[0] + dl[1] * z[1] + dl[2] * z[2];
        v[1] = z[1] * x[2] - z[2] * x[1];      // c = sum Z_i x Y_i
        v[2] = z[2] * x[0] - z[0] * x[2];
        v[3] = z[0] * x[1] - z[1] * x[0];
        v[4] = dl[1] * x[2] - dl[2] * x[1];    // u = sum delta_i x Y_i
        v[5] = dl[2] * x[0] - dl[0] * x[2];
        v[6] = dl[0] * x[1] - dl[1] * x[0];
        half_allreduce_t<8>(v, lane);
        double om[3];
        gik_sym_mul(Mi, v + 1, om);
        Hd[0] = z[0] - (x[1] * om[2] - x[2] * om[1]);
        Hd[1] = z[1] - (x[2] * om[0] - x[0] * om[2]);
        Hd[2] = z[2] - (x[0] * om[1] - x[1] * om[0]);
        const double d_Hd = v[0] - (om[0] * v[4] + om[1] * v[5] + om[2] * v[6]);
        inner_total += phase == PH_INNER;
        const double alpha = gik_div(z_r, d_Hd, gik_rcp(d_Hd));
        const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
        // boundary / negative curvature (!(d_Hd > 0) also catches NaN): rare, handled divergently
        const bool exit1 = phase == PH_INNER && (!(d_Hd > 0.0) || e_Pe_new >= Delta2);
        if (__any_sync(GIK_FULL_MASK, exit1)) {
            if (exit1) {
                const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    eta[q] = fma(tau, dl[q], eta[q]);
                    Heta[q] = fma(tau, Hd[q], Heta[q]);
                }
                stop = d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                numit = j;
                phase = PH_NEED_OUTER;
            }
            __syncwarp();
        }
        double ne[3], nh[3], nr[3], sdot[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            ne[q] = fma(alpha, dl[q], eta[q]);
            nh[q] = fma(alpha, Hd[q], Heta[q]);
            nr[q] = fma(alpha, Hd[q], r[q]);
            sdot[0] = fma(ne[q], g[q], sdot[0]);
            sdot[1] = fma(ne[q], nh[q], sdot[1]);
            sdot[2] = fma(nr[q], nr[q], sdot[2]);
        }
        half_allreduce_t<4>(sdot, lane);
        const double new_model_value = sdot[0] + 0.5 * sdot[1];
        const bool cont = phase == PH_INNER;
        const bool model_inc = cont && new_model_value >= model_value;
        const bool commit = cont && !model_inc;
        if (model_inc) { stop = MODEL_INCREASED; numit = j; phase = PH_NEED_OUTER; }
        if (commit) {
            e_Pe = e_Pe_new;
#pragma unroll
            for (int q = 0; q < 3; ++q) { eta[q] = ne[q]; Heta[q] = nh[q]; r[q] = nr[q]; }
            model_value = new_model_value;
            r_r = sdot[2];
        }
        const bool reached = commit && j >= o.mininner && r_r <= r_target2;
        if (reached) {
            stop = o.kappa < pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
            numit = j;
            phase = PH_NEED_OUTER;
        }
        const double beta = gik_div(r_r, z_r, inv_z_r);
        if (commit && !reached) {
            z_r = r_r;
            inv_z_r = gik_rcp(z_r);
#pragma unroll
            for (int q = 0; q < 3; ++q) dl[q] = fma(beta, dl[q], -r[q]);
            e_Pd = beta * (e_Pd + alpha * d_Pd);
            d_Pd = z_r + beta * beta * d_Pd;
            ++j;
            if (j >= o.maxinner) { stop = MAX_INNER_ITER; numit = o.maxinner - 1; phase = PH_NEED_OUTER; }
        }
    }
}

template <int SPL>
int launch(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t smem = (size_t)(192 + 2 * goal_pad + SPL * 32 + SPL * 4 * 32) * sizeof(double);
    if (smem > 48 * 1024)
        GIK_CUDA(cudaFuncSetAttribute(k_rtr_duo<SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rtr_duo<SPL>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = p->sm_count * per_sm;
    const int need = (a.B + 1) / 2;
    if (blocks > need) blocks = need;
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    k_rtr_duo<SPL><<<blocks, kThreads, smem, st>>>(a, p->duo_info, p->duo_target);
    return gik_check_cuda(cudaGetLastError(), "k_rtr_duo launch");
}

}  // namespace

int gik_launch_rtr_duo(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    if (!p->duo_info) return 1;
    const int d = p->maxdeg;
    if (d <= 6) return launch<6>(p, a, st);
    if (d <= 8) return launch<8>(p, a, st);
    if (d <= 9) return launch<9>(p, a, st);
    if (d <= 10) return launch<10>(p, a, st);
    if (d <= 12) return launch<12>(p, a, st);
    return 1;
}
