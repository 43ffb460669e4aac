// gik_rtr_fast.cu -- latency-optimised trust-region solve for graphs with N <= 32 nodes.
//
// Same algorithm and same arithmetic conventions as k_rtr (gik_rtr.cu; reference
// trust_region.py:112-599, costs.py:79-207, fixed_rank_psd_sym.py:91-137), laid out for
// the critical path of ONE problem, because a batch's wall time is set by its slowest
// problems (outer iteration counts range from ~50 to the 3000 cap):
//
//   * one full warp per problem -- two problems sharing a warp diverge almost always
//     (ncu: 16.5 of 32 threads active per instruction in k_rtr<16,1>) and serialise;
//   * LPN lanes per node (LPN = 2 for N <= 16): the lanes of a node split its slot list
//     and combine the partial sums with one xor-shuffle, halving the serial slot loop;
//   * everything about an edge term that only depends on the current iterate x --
//     D_ij = x_i - x_j, the activity of the hinge and the residual -- is cached in
//     REGISTERS once per accepted outer iteration (SPL slots per lane, compile-time),
//     so an inner tCG iteration loads only the neighbour's direction delta_j:
//       Z_i = sum_slots <D', w> D' + c w,   w = delta_i - delta_j,  D' = 2 act D,  c = 2 act r
//     (= costs.py lhess: 2 * sum act [2 <D,w> D + (d - T) w]);
//   * the cache is rebuilt by the cost/gradient pass of the proposal x + eta, which the
//     outer iteration needs anyway (trust_region.py:248-251); a rejected step rebuilds
//     it at x.
#include <cstdlib>

#include "gik_rtr.cuh"
#include "gik_warp.cuh"

namespace {

constexpr int kThreads = 32;            // one warp = one problem per CTA: a straggler pins only its own warp's
                                        // registers, so the next batch's kernel can move in beside it
constexpr int kWarps = kThreads / 32;

struct SlotCache {
    double dx, dy, dz;  // 2 * act * (x_i - x_j): zero for an inactive hinge, so (D.w) D = 4 act <d,w> d
    double c2;          // 2 * act * (d_ij - T_ij)
};

// SMC: keep the slot cache in shared memory ([slot][4][lane], conflict-free) instead of registers --
// used for the one-lane-per-node layouts (17..32 nodes, up to 12 slots per lane), where the register
// cache costs 238+ registers (8 warps / SM); the loads do not depend on delta, so they stay off the
// critical path
template <int LPN, int SPL, bool SMC>
__device__ __forceinline__ void rtr_fast_body(const RtrArgs &a, const uint32_t *__restrict__ fast_info,
                                              const double *__restrict__ fast_target)
{
    constexpr int NPW = 32 / LPN;  // node slots per warp
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int node = lane / LPN;
    const bool valid = node < a.N;
    const int goal_pad = (a.n_goal + 1) & ~1;
    double *P = smem + (size_t)warp * (6 * NPW + goal_pad + SPL * 32 + (SMC ? SPL * 128 : 0));
    double *V = P + 3 * NPW;
    double *goal = V + 3 * NPW;
    double *tgt = goal + goal_pad + lane;   // [SPL][32] per-problem targets of this warp's slots
    // slot cache when SMC: [SPL][2][32 lanes] double2 {dx, dy}, {dz, c2} -- two conflict-free LDS.128 per slot
    double2 *scm = reinterpret_cast<double2 *>(goal + goal_pad + SPL * 32) + lane;
    const GikSolveOpts &o = a.o;
    const double eps = 2.220446049250313e-16;  // np.spacing(1), trust_region.py:293

    // static slot description of this lane (identical for every problem)
    uint32_t info[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) info[s] = fast_info[s * 32 + lane];

    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(a.work_counter, 1);
        b = __shfl_sync(GIK_FULL_MASK, b, 0, 32);
        if (b >= a.B) break;

        double x[3] = {0.0, 0.0, 0.0}, g[3], eta[3], Heta[3], r[3], dl[3], Hd[3];
        SlotCache sc[SMC ? 1 : SPL];
        if (valid) {
            const double *src = a.Y_init + ((size_t)b * a.N + node) * 3;
            x[0] = src[0]; x[1] = src[1]; x[2] = src[2];
        }
        __syncwarp();
        for (int k = lane; k < a.n_goal; k += 32) goal[k] = a.goal_d2[(size_t)b * a.n_goal + k];
        if (lane % LPN == 0) { P[node] = x[0]; P[NPW + node] = x[1]; P[2 * NPW + node] = x[2]; }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const uint32_t gs = GIK_SLOT_GOAL(info[s]);
            tgt[s * 32] = gs ? goal[gs - 1] : fast_target[s * 32 + lane];
        }

        // cost / gradient at point p (published in P) + rebuild of the slot cache.
        // Returns this lane's cost share; gout = full half-gradient of the node.
        auto rebuild = [&](const double (&p)[3], double (&gout)[3]) -> double {
            double fpart = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const int j = GIK_SLOT_NBR(info[s]);
                const uint32_t kind = GIK_SLOT_KIND(info[s]);
                const double dx = p[0] - P[j], dy = p[1] - P[NPW + j], dz = p[2] - P[2 * NPW + j];
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
                if (SMC) {
                    scm[(s * 2 + 0) * 32] = make_double2(two * dx, two * dy);
                    scm[(s * 2 + 1) * 32] = make_double2(two * dz, 2.0 * rr);
                } else {
                    sc[s].dx = two * dx; sc[s].dy = two * dy; sc[s].dz = two * dz;
                    sc[s].c2 = 2.0 * rr;
                }
            }
            double v[3] = {2.0 * gx, 2.0 * gy, 2.0 * gz};
            pair_combine<LPN, 3>(v);
            gout[0] = v[0]; gout[1] = v[1]; gout[2] = v[2];
            return 0.5 * fpart;
        };

        double fx, gg, Mi[6];
        {
            double f1[1] = {rebuild(x, g)};
            pair_combine<LPN, 1>(f1);
            double v[8] = {f1[0], g[0] * g[0] + g[1] * g[1] + g[2] * g[2],
                           x[0] * x[0], x[0] * x[1], x[0] * x[2], x[1] * x[1], x[1] * x[2], x[2] * x[2]};
            node_allreduce<LPN, 8>(v);
            fx = v[0];
            gg = v[1];
            gik_sylvester_inverse(v + 2, Mi);
        }
        double norm_grad = sqrt(gg);
        double Delta = o.Delta0;
        int k_outer = 0, inner_total = 0, status = GIK_STATUS_MAXITER;
        if (!(isfinite(fx) && isfinite(gg))) {
            status = GIK_STATUS_NAN;
        } else {
            for (;;) {
                // ================= tCG (trust_region.py:436-599), eta0 = 0, precon = identity
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    eta[q] = 0.0; Heta[q] = 0.0; r[q] = g[q]; dl[q] = -g[q];
                }
                double e_Pe = 0.0, r_r = gg;
                const double norm_r0 = sqrt(r_r);
                double z_r = r_r, d_Pd = r_r, e_Pd = 0.0, model_value = 0.0;
                const double pw = o.theta == 1.0 ? norm_r0 : pow(norm_r0, o.theta);
                const double r_target = norm_r0 * fmin(pw, o.kappa);
                // sqrt is monotone: ||r|| <= target  <=>  <r,r> <= target^2 (saves a sqrt per inner iteration)
                const double r_target2 = r_target * r_target;
                const double Delta2 = Delta * Delta;
                double inv_z_r = gik_rcp(z_r);  // for beta = z_r_new / z_r, formed one iteration ahead
                int stop = MAX_INNER_ITER;
                int j = 0;
                for (j = 0; j < o.maxinner; ++j) {
                    // ---- Hdelta = proj(x, lhess(x, delta))
                    if (lane % LPN == 0) { V[node] = dl[0]; V[NPW + node] = dl[1]; V[2 * NPW + node] = dl[2]; }
                    __syncwarp();
                    // two independent accumulator chains (c w and <D', w> D') halve the dependent FMA depth
                    double z[3] = {0.0, 0.0, 0.0}, zb[3] = {0.0, 0.0, 0.0};
#pragma unroll
                    for (int s = 0; s < SPL; ++s) {
                        const int jn = GIK_SLOT_NBR(info[s]);
                        const double wx = dl[0] - V[jn], wy = dl[1] - V[NPW + jn], wz = dl[2] - V[2 * NPW + jn];
                        double cx, cy, cz, c2;
                        if (SMC) {
                            const double2 a0 = scm[(s * 2 + 0) * 32], a1 = scm[(s * 2 + 1) * 32];
                            cx = a0.x; cy = a0.y; cz = a1.x; c2 = a1.y;
                        } else {
                            cx = sc[s].dx; cy = sc[s].dy; cz = sc[s].dz; c2 = sc[s].c2;
                        }
                        const double t = fma(cx, wx, fma(cy, wy, cz * wz));
                        z[0] = fma(c2, wx, z[0]);
                        z[1] = fma(c2, wy, z[1]);
                        z[2] = fma(c2, wz, z[2]);
                        zb[0] = fma(t, cx, zb[0]);
                        zb[1] = fma(t, cy, zb[1]);
                        zb[2] = fma(t, cz, zb[2]);
                    }
                    z[0] += zb[0]; z[1] += zb[1]; z[2] += zb[2];
                    pair_combine<LPN, 3>(z);
                    double v[8];
                    v[7] = 0.0;
                    v[0] = dl[0] * z[0] + dl[1] * z[1] + dl[2] * z[2];
                    v[1] = z[1] * x[2] - z[2] * x[1];      // c = sum Z_i x Y_i
                    v[2] = z[2] * x[0] - z[0] * x[2];
                    v[3] = z[0] * x[1] - z[1] * x[0];
                    v[4] = dl[1] * x[2] - dl[2] * x[1];    // u = sum delta_i x Y_i
                    v[5] = dl[2] * x[0] - dl[0] * x[2];
                    v[6] = dl[0] * x[1] - dl[1] * x[0];
                    node_allreduce_t<LPN, 8>(v, lane);
                    double om[3];
                    gik_sym_mul(Mi, v + 1, om);
                    Hd[0] = z[0] - (x[1] * om[2] - x[2] * om[1]);
                    Hd[1] = z[1] - (x[2] * om[0] - x[0] * om[2]);
                    Hd[2] = z[2] - (x[0] * om[1] - x[1] * om[0]);
                    const double d_Hd = v[0] - (om[0] * v[4] + om[1] * v[5] + om[2] * v[6]);
                    ++inner_total;
                    const double alpha = gik_div(z_r, d_Hd, gik_rcp(d_Hd));
                    const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
                    // !(d_Hd > 0) also catches NaN (the reference would spin to maxinner on it)
                    if (!(d_Hd > 0.0) || e_Pe_new >= Delta2) {
                        const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            eta[q] = fma(tau, dl[q], eta[q]);
                            Heta[q] = fma(tau, Hd[q], Heta[q]);
                        }
                        stop = d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                        break;
                    }
                    e_Pe = e_Pe_new;
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
                    node_allreduce_t<LPN, 4>(sdot, lane);
                    const double new_model_value = sdot[0] + 0.5 * sdot[1];
                    if (new_model_value >= model_value) {
                        stop = MODEL_INCREASED;
                        break;
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) { eta[q] = ne[q]; Heta[q] = nh[q]; r[q] = nr[q]; }
                    model_value = new_model_value;
                    r_r = sdot[2];
                    if (j >= o.mininner && r_r <= r_target2) {
                        stop = o.kappa < pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
                        break;
                    }
                    const double beta = gik_div(r_r, z_r, inv_z_r);
                    z_r = r_r;
                    inv_z_r = gik_rcp(z_r);
#pragma unroll
                    for (int q = 0; q < 3; ++q) dl[q] = fma(beta, dl[q], -r[q]);
                    e_Pd = beta * (e_Pd + alpha * d_Pd);
                    d_Pd = z_r + beta * beta * d_Pd;
                }
                const int numit = j < o.maxinner ? j : o.maxinner - 1;

                // ================= proposal x + eta; dl <- x_prop, Hd <- grad(x_prop), cache <- x_prop
#pragma unroll
                for (int q = 0; q < 3; ++q) dl[q] = x[q] + eta[q];
                if (lane % LPN == 0) { P[node] = dl[0]; P[NPW + node] = dl[1]; P[2 * NPW + node] = dl[2]; }
                __syncwarp();
                double f1[1] = {rebuild(dl, Hd)};
                pair_combine<LPN, 1>(f1);
                double v[10] = {f1[0],
                                g[0] * eta[0] + g[1] * eta[1] + g[2] * eta[2],
                                eta[0] * Heta[0] + eta[1] * Heta[1] + eta[2] * Heta[2],
                                Hd[0] * Hd[0] + Hd[1] * Hd[1] + Hd[2] * Hd[2],
                                dl[0] * dl[0], dl[0] * dl[1], dl[0] * dl[2], dl[1] * dl[1], dl[1] * dl[2],
                                dl[2] * dl[2]};
                node_allreduce<LPN, 10>(v);
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
                    for (int q = 0; q < 3; ++q) { x[q] = dl[q]; g[q] = Hd[q]; }
                    fx = fx_prop;
                    gg = v[3];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 4, Mi);
                } else {
                    // rejected: bring the exchange buffer and the slot cache back to x
                    __syncwarp();
                    if (lane % LPN == 0) { P[node] = x[0]; P[NPW + node] = x[1]; P[2 * NPW + node] = x[2]; }
                    __syncwarp();
                    double gtmp[3];
                    rebuild(x, gtmp);
                }
                if (a.trace && k_outer < a.trace_rows && lane == 0) {
                    double *row = a.trace + ((size_t)b * a.trace_rows + k_outer) * 6;
                    row[0] = Delta_used;
                    row[1] = (double)numit;
                    row[2] = (double)stop;
                    row[3] = fx_prop;
                    row[4] = accept ? 1.0 : 0.0;
                    row[5] = accept ? norm_grad : nan("");
                }
                ++k_outer;
                if (k_outer >= o.maxiter) { status = GIK_STATUS_MAXITER; break; }
                if (norm_grad < o.mingradnorm) { status = GIK_STATUS_CONVERGED; break; }
            }
        }
        if (valid && lane % LPN == 0) {
            double *dst = a.Y_out + ((size_t)b * a.N + node) * 3;
            dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
        }
        if (lane == 0) {
            a.f[b] = fx;
            a.gradnorm[b] = norm_grad;
            a.iters[b] = k_outer;
            a.status[b] = status;
            if (a.n_inner) a.n_inner[b] = inner_total;
        }
        __syncwarp();
    }
}

// Where the slot cache lives and how many warps fit: the paired layout (N <= 16, <= 6 slots per lane) keeps it in
// registers at 168 registers = 12 warps / SM; one lane per node with <= 9 slots (7-DOF arms: KUKA, LWA4D, Panda)
// keeps it in registers too, at 238 registers = 8 warps / SM -- measured on KUKA IIWA against the shared-memory cache
// at 12 warps / SM: same throughput at 65 536 goals (1439 ms), 9 % lower latency at 16 384 (437 vs 480 ms), because the
// kernel is bound by FP64 / MIO issue and dependent latency, not by resident warps; longer slot lists use shared memory.
__host__ __device__ constexpr bool fast_cache_in_smem(int LPN, int SPL) { return LPN == 1 ? SPL > 9 : SPL > 6; }
__host__ __device__ constexpr int fast_min_blocks(int LPN, int SPL) { return (LPN == 1 && SPL <= 9) ? 8 : 12; }

template <int LPN, int SPL>
__global__ void __launch_bounds__(kThreads, fast_min_blocks(LPN, SPL))
k_rtr_fast(const RtrArgs a, const uint32_t *__restrict__ fast_info, const double *__restrict__ fast_target)
{
    rtr_fast_body<LPN, SPL, fast_cache_in_smem(LPN, SPL)>(a, fast_info, fast_target);
}

template <int LPN, int SPL>
int launch(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    auto kern = k_rtr_fast<LPN, SPL>;
    constexpr int NPW = 32 / LPN;
    constexpr bool SMC = fast_cache_in_smem(LPN, SPL);
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t smem = (size_t)kWarps * (6 * NPW + goal_pad + SPL * 32 + (SMC ? SPL * 128 : 0)) * sizeof(double);
    if (smem > 48 * 1024)
        GIK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = p->sm_count * per_sm;
    const int need = (a.B + kWarps - 1) / kWarps;
    if (blocks > need) blocks = need;
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    kern<<<blocks, kThreads, smem, st>>>(a, p->fast_info, p->fast_target);
    return gik_check_cuda(cudaGetLastError(), "k_rtr_fast launch");
}


// ---------------------------------------------------------------------------------------------------
// 33 .. 64 nodes, sparse (e.g. the 20-DOF chain of BASELINE configs[3]: N = 44, 115 terms, degree 5..7):
// one warp per problem, TWO nodes per lane.  The plan orders the nodes by degree: the 32 highest-degree
// nodes are every lane's first node (S0 slots), the rest the second node of lanes 0 .. N-33 (S1 slots),
// so a lane walks S0 + S1 slots instead of the 2 * maxdeg of a fixed l / l+32 assignment.  Slot cache in
// shared memory ([slot][4][lane], conflict-free; loads independent of delta), node state in registers.
template <int S0, int S1>
__device__ __forceinline__ void rtr_fast2_body(const RtrArgs &a, const uint32_t *__restrict__ info_tbl,
                                               const double *__restrict__ target_tbl,
                                               const int32_t *__restrict__ node_tbl)
{
    constexpr int ST = S0 + S1;
    constexpr int NPW = 64;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int goal_pad = (a.n_goal + 1) & ~1;
    double *P = smem;
    double *V = P + 3 * NPW;
    double *goal = V + 3 * NPW;
    double *tgt = goal + goal_pad + lane;   // [ST][32] per-problem targets of this lane's slots
    double2 *scm = reinterpret_cast<double2 *>(goal + goal_pad + ST * 32) + lane;   // [ST][2][32] double2 {dx, dy}, {dz, c2}
    const GikSolveOpts &o = a.o;
    const double eps = 2.220446049250313e-16;  // np.spacing(1), trust_region.py:293

    // this lane's nodes; a lane without a second node parks it on the unused exchange slot 63
    int nd[2];
    bool valid[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int t = node_tbl[m * 32 + lane];
        valid[m] = t >= 0;
        nd[m] = valid[m] ? t : NPW - 1;
    }
    uint32_t info[ST];
#pragma unroll
    for (int s = 0; s < ST; ++s) info[s] = info_tbl[s * 32 + lane];

    auto publish = [&](double *buf, const double (&v)[2][3]) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            buf[nd[m]] = v[m][0]; buf[NPW + nd[m]] = v[m][1]; buf[2 * NPW + nd[m]] = v[m][2];
        }
    };

    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(a.work_counter, 1);
        b = __shfl_sync(GIK_FULL_MASK, b, 0, 32);
        if (b >= a.B) break;

        double x[2][3], g[2][3], eta[2][3], Heta[2][3], r[2][3], dl[2][3], Hd[2][3];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            x[m][0] = 0.0; x[m][1] = 0.0; x[m][2] = 0.0;
            if (valid[m]) {
                const double *src = a.Y_init + ((size_t)b * a.N + nd[m]) * 3;
                x[m][0] = src[0]; x[m][1] = src[1]; x[m][2] = src[2];
            }
        }
        __syncwarp();
        for (int k = lane; k < a.n_goal; k += 32) goal[k] = a.goal_d2[(size_t)b * a.n_goal + k];
        publish(P, x);
        __syncwarp();
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            const uint32_t gs = GIK_SLOT_GOAL(info[s]);
            tgt[s * 32] = gs ? goal[gs - 1] : target_tbl[s * 32 + lane];
        }

        // cost / gradient at point p (published in P) + rebuild of the slot cache
        auto rebuild = [&](const double (&p)[2][3], double (&gout)[2][3]) -> double {
            double fpart = 0.0;
            double ga[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
            for (int s = 0; s < ST; ++s) {
                const int m = s < S0 ? 0 : 1;
                const int j = GIK_SLOT_NBR(info[s]);
                const uint32_t kind = GIK_SLOT_KIND(info[s]);
                const double dx = p[m][0] - P[j], dy = p[m][1] - P[NPW + j], dz = p[m][2] - P[2 * NPW + j];
                const double d = dx * dx + dy * dy + dz * dz;
                double rr = d - tgt[s * 32];
                const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (rr < 0.0)) |
                                 ((kind == GIK_TERM_UP) & (rr > 0.0));
                rr = act ? rr : 0.0;
                fpart = fma(rr, rr, fpart);
                ga[m][0] = fma(rr, dx, ga[m][0]);
                ga[m][1] = fma(rr, dy, ga[m][1]);
                ga[m][2] = fma(rr, dz, ga[m][2]);
                const double two = act ? 2.0 : 0.0;
                scm[(s * 2 + 0) * 32] = make_double2(two * dx, two * dy);
                scm[(s * 2 + 1) * 32] = make_double2(two * dz, 2.0 * rr);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                gout[m][0] = 2.0 * ga[m][0]; gout[m][1] = 2.0 * ga[m][1]; gout[m][2] = 2.0 * ga[m][2];
            }
            return 0.5 * fpart;
        };

        double fx, gg, Mi[6];
        {
            double v[8];
            v[0] = rebuild(x, g);
            v[1] = 0.0;
#pragma unroll
            for (int k = 2; k < 8; ++k) v[k] = 0.0;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                v[1] = fma(g[m][0], g[m][0], fma(g[m][1], g[m][1], fma(g[m][2], g[m][2], v[1])));
                v[2] = fma(x[m][0], x[m][0], v[2]); v[3] = fma(x[m][0], x[m][1], v[3]);
                v[4] = fma(x[m][0], x[m][2], v[4]); v[5] = fma(x[m][1], x[m][1], v[5]);
                v[6] = fma(x[m][1], x[m][2], v[6]); v[7] = fma(x[m][2], x[m][2], v[7]);
            }
            node_allreduce<1, 8>(v);
            fx = v[0];
            gg = v[1];
            gik_sylvester_inverse(v + 2, Mi);
        }
        double norm_grad = sqrt(gg);
        double Delta = o.Delta0;
        int k_outer = 0, inner_total = 0, status = GIK_STATUS_MAXITER;
        if (!(isfinite(fx) && isfinite(gg))) {
            status = GIK_STATUS_NAN;
        } else {
            for (;;) {
                // ================= tCG (trust_region.py:436-599), eta0 = 0, precon = identity
#pragma unroll
                for (int m = 0; m < 2; ++m) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        eta[m][q] = 0.0; Heta[m][q] = 0.0; r[m][q] = g[m][q]; dl[m][q] = -g[m][q];
                    }
                }
                double e_Pe = 0.0, r_r = gg;
                const double norm_r0 = sqrt(r_r);
                double z_r = r_r, d_Pd = r_r, e_Pd = 0.0, model_value = 0.0;
                const double pw = o.theta == 1.0 ? norm_r0 : pow(norm_r0, o.theta);
                const double r_target = norm_r0 * fmin(pw, o.kappa);
                const double r_target2 = r_target * r_target;   // <r,r> <= target^2 (sqrt is monotone)
                const double Delta2 = Delta * Delta;
                double inv_z_r = gik_rcp(z_r);
                int stop = MAX_INNER_ITER;
                int j = 0;
                for (j = 0; j < o.maxinner; ++j) {
                    // ---- Hdelta = proj(x, lhess(x, delta))
                    publish(V, dl);
                    __syncwarp();
                    double z[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}, zb[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
                    for (int s = 0; s < ST; ++s) {
                        const int m = s < S0 ? 0 : 1;
                        const int jn = GIK_SLOT_NBR(info[s]);
                        const double wx = dl[m][0] - V[jn], wy = dl[m][1] - V[NPW + jn], wz = dl[m][2] - V[2 * NPW + jn];
                        const double2 a0 = scm[(s * 2 + 0) * 32], a1 = scm[(s * 2 + 1) * 32];
                        const double cx = a0.x, cy = a0.y, cz = a1.x, c2 = a1.y;
                        const double t = fma(cx, wx, fma(cy, wy, cz * wz));
                        z[m][0] = fma(c2, wx, z[m][0]);
                        z[m][1] = fma(c2, wy, z[m][1]);
                        z[m][2] = fma(c2, wz, z[m][2]);
                        zb[m][0] = fma(t, cx, zb[m][0]);
                        zb[m][1] = fma(t, cy, zb[m][1]);
                        zb[m][2] = fma(t, cz, zb[m][2]);
                    }
                    double v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = 0.0;
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        z[m][0] += zb[m][0]; z[m][1] += zb[m][1]; z[m][2] += zb[m][2];
                        v[0] += dl[m][0] * z[m][0] + dl[m][1] * z[m][1] + dl[m][2] * z[m][2];
                        v[1] += z[m][1] * x[m][2] - z[m][2] * x[m][1];      // c = sum Z_i x Y_i
                        v[2] += z[m][2] * x[m][0] - z[m][0] * x[m][2];
                        v[3] += z[m][0] * x[m][1] - z[m][1] * x[m][0];
                        v[4] += dl[m][1] * x[m][2] - dl[m][2] * x[m][1];    // u = sum delta_i x Y_i
                        v[5] += dl[m][2] * x[m][0] - dl[m][0] * x[m][2];
                        v[6] += dl[m][0] * x[m][1] - dl[m][1] * x[m][0];
                    }
                    node_allreduce_t<1, 8>(v, lane);
                    double om[3];
                    gik_sym_mul(Mi, v + 1, om);
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        Hd[m][0] = z[m][0] - (x[m][1] * om[2] - x[m][2] * om[1]);
                        Hd[m][1] = z[m][1] - (x[m][2] * om[0] - x[m][0] * om[2]);
                        Hd[m][2] = z[m][2] - (x[m][0] * om[1] - x[m][1] * om[0]);
                    }
                    const double d_Hd = v[0] - (om[0] * v[4] + om[1] * v[5] + om[2] * v[6]);
                    ++inner_total;
                    const double alpha = gik_div(z_r, d_Hd, gik_rcp(d_Hd));
                    const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
                    if (!(d_Hd > 0.0) || e_Pe_new >= Delta2) {   // also catches NaN
                        const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                eta[m][q] = fma(tau, dl[m][q], eta[m][q]);
                                Heta[m][q] = fma(tau, Hd[m][q], Heta[m][q]);
                            }
                        }
                        stop = d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                        break;
                    }
                    e_Pe = e_Pe_new;
                    double ne[2][3], nh[2][3], nr[2][3], sdot[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            ne[m][q] = fma(alpha, dl[m][q], eta[m][q]);
                            nh[m][q] = fma(alpha, Hd[m][q], Heta[m][q]);
                            nr[m][q] = fma(alpha, Hd[m][q], r[m][q]);
                            sdot[0] = fma(ne[m][q], g[m][q], sdot[0]);
                            sdot[1] = fma(ne[m][q], nh[m][q], sdot[1]);
                            sdot[2] = fma(nr[m][q], nr[m][q], sdot[2]);
                        }
                    }
                    node_allreduce_t<1, 4>(sdot, lane);
                    const double new_model_value = sdot[0] + 0.5 * sdot[1];
                    if (new_model_value >= model_value) {
                        stop = MODEL_INCREASED;
                        break;
                    }
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) { eta[m][q] = ne[m][q]; Heta[m][q] = nh[m][q]; r[m][q] = nr[m][q]; }
                    }
                    model_value = new_model_value;
                    r_r = sdot[2];
                    if (j >= o.mininner && r_r <= r_target2) {
                        stop = o.kappa < pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
                        break;
                    }
                    const double beta = gik_div(r_r, z_r, inv_z_r);
                    z_r = r_r;
                    inv_z_r = gik_rcp(z_r);
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) dl[m][q] = fma(beta, dl[m][q], -r[m][q]);
                    }
                    e_Pd = beta * (e_Pd + alpha * d_Pd);
                    d_Pd = z_r + beta * beta * d_Pd;
                }
                const int numit = j < o.maxinner ? j : o.maxinner - 1;

                // ================= proposal x + eta; dl <- x_prop, Hd <- grad(x_prop), cache <- x_prop
#pragma unroll
                for (int m = 0; m < 2; ++m) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) dl[m][q] = x[m][q] + eta[m][q];
                }
                __syncwarp();
                publish(P, dl);
                __syncwarp();
                double v[10];
                v[0] = rebuild(dl, Hd);
#pragma unroll
                for (int k = 1; k < 10; ++k) v[k] = 0.0;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    v[1] += g[m][0] * eta[m][0] + g[m][1] * eta[m][1] + g[m][2] * eta[m][2];
                    v[2] += eta[m][0] * Heta[m][0] + eta[m][1] * Heta[m][1] + eta[m][2] * Heta[m][2];
                    v[3] += Hd[m][0] * Hd[m][0] + Hd[m][1] * Hd[m][1] + Hd[m][2] * Hd[m][2];
                    v[4] = fma(dl[m][0], dl[m][0], v[4]); v[5] = fma(dl[m][0], dl[m][1], v[5]);
                    v[6] = fma(dl[m][0], dl[m][2], v[6]); v[7] = fma(dl[m][1], dl[m][1], v[7]);
                    v[8] = fma(dl[m][1], dl[m][2], v[8]); v[9] = fma(dl[m][2], dl[m][2], v[9]);
                }
                node_allreduce<1, 10>(v);
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
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) { x[m][q] = dl[m][q]; g[m][q] = Hd[m][q]; }
                    }
                    fx = fx_prop;
                    gg = v[3];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 4, Mi);
                } else {
                    // rejected: bring the exchange buffer and the slot cache back to x
                    __syncwarp();
                    publish(P, x);
                    __syncwarp();
                    double gtmp[2][3];
                    rebuild(x, gtmp);
                }
                if (a.trace && k_outer < a.trace_rows && lane == 0) {
                    double *row = a.trace + ((size_t)b * a.trace_rows + k_outer) * 6;
                    row[0] = Delta_used;
                    row[1] = (double)numit;
                    row[2] = (double)stop;
                    row[3] = fx_prop;
                    row[4] = accept ? 1.0 : 0.0;
                    row[5] = accept ? norm_grad : nan("");
                }
                ++k_outer;
                if (k_outer >= o.maxiter) { status = GIK_STATUS_MAXITER; break; }
                if (norm_grad < o.mingradnorm) { status = GIK_STATUS_CONVERGED; break; }
                __syncwarp();
            }
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (valid[m]) {
                double *dst = a.Y_out + ((size_t)b * a.N + nd[m]) * 3;
                dst[0] = x[m][0]; dst[1] = x[m][1]; dst[2] = x[m][2];
            }
        }
        if (lane == 0) {
            a.f[b] = fx;
            a.gradnorm[b] = norm_grad;
            a.iters[b] = k_outer;
            a.status[b] = status;
            if (a.n_inner) a.n_inner[b] = inner_total;
        }
        __syncwarp();
    }
}

// 254 registers without spills (8 warps / SM) against 168 with ~20 spilled doubles (11 warps / SM, shared-memory bound)
#ifndef GIK_FAST2_MINB
#define GIK_FAST2_MINB 8
#endif
template <int S0, int S1>
__global__ void __launch_bounds__(kThreads, GIK_FAST2_MINB) k_rtr_fast2(const RtrArgs a, const uint32_t *__restrict__ info_tbl,
                                                            const double *__restrict__ target_tbl,
                                                            const int32_t *__restrict__ node_tbl)
{
    rtr_fast2_body<S0, S1>(a, info_tbl, target_tbl, node_tbl);
}

template <int S0, int S1>
int launch2(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    auto kern = k_rtr_fast2<S0, S1>;
    constexpr int ST = S0 + S1;
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t smem = (size_t)(6 * 64 + goal_pad + ST * 32 + ST * 128) * sizeof(double);
    if (smem > 227 * 1024) return 1;
    if (smem > 48 * 1024)
        GIK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = p->sm_count * per_sm;
    if (blocks > a.B) blocks = a.B;
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    kern<<<blocks, kThreads, smem, st>>>(a, p->fast2_info, p->fast2_target, p->fast2_node);
    return gik_check_cuda(cudaGetLastError(), "k_rtr_fast2 launch");
}

}  // namespace

int gik_launch_rtr_fast(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    if (!p->fast_info) return 1;
    if (p->fast_LPN == 2) {
        switch (p->fast_SPL) {
            case 1: case 2: case 3: case 4: return launch<2, 4>(p, a, st);
            case 5: return launch<2, 5>(p, a, st);
            case 6: return launch<2, 6>(p, a, st);
            case 7: case 8: return launch<2, 8>(p, a, st);
            default: return 1;
        }
    }
    switch (p->fast_SPL) {
        case 1: case 2: case 3: case 4: case 5: case 6: return launch<1, 6>(p, a, st);
        case 7: case 8: return launch<1, 8>(p, a, st);
        case 9: return launch<1, 9>(p, a, st);
        case 10: return launch<1, 10>(p, a, st);
        case 11: case 12: return launch<1, 12>(p, a, st);
        default: return 1;
    }
}

// 33 .. 64 nodes, two nodes per lane.  The tables are laid out for the instantiated (S0, S1) the plan chose.
int gik_launch_rtr_fast2(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    if (!p->fast2_info) return 1;
    switch (p->fast2_S0 * 16 + p->fast2_S1) {
        case 6 * 16 + 5: return launch2<6, 5>(p, a, st);
        case 7 * 16 + 5: return launch2<7, 5>(p, a, st);
        case 8 * 16 + 5: return launch2<8, 5>(p, a, st);
        case 8 * 16 + 8: return launch2<8, 8>(p, a, st);
        case 12 * 16 + 8: return launch2<12, 8>(p, a, st);
        case 12 * 16 + 12: return launch2<12, 12>(p, a, st);
        default: return 1;
    }
}
