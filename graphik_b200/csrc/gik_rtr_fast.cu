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
#include "gik_tr_math.cuh"
#include "gik_warp.cuh"

namespace {

constexpr int kThreads = 32;            // one warp = one problem per CTA: a straggler pins only its own warp's
                                        // registers, so the next batch's kernel can move in beside it

struct SlotCache {
    double dx, dy, dz;  // 2 * act * (x_i - x_j): zero for an inactive hinge, so (D.w) D = 4 act <d,w> d
    double c2;          // 2 * act * (d_ij - T_ij)
};

// per-warp shared memory (doubles): P, V [3][NPW], goal row, per-problem slot targets [SPL][32], slot cache when
// SMC, reduction rows [8][gik_red_stride] + totals [8], start time of the problem (+ 1 spare word)
__host__ __device__ constexpr int fast_smem_doubles(int LPN, int SPL, bool SMC, int goal_pad)
{
    return 6 * (32 / LPN) + goal_pad + SPL * 32 + (SMC ? SPL * 128 : 0) + 8 * gik_red_stride(32 / LPN) + 8 + 2;
}

template <int LPN, int K, bool LAT>
__device__ __forceinline__ void fast_allreduce(double (&v)[K], const GikRedAddr &ra, int lane)
{

    // measured on UR10 (one 4096-goal batch, i.e. the slowest goal alone on its SM): the shared-memory reduction
    // beats xor butterflies in the latency variant as well (140.6 vs 148.7 ms) -- shuffles and shared-memory accesses
    // issue at one per ~4.4 cycles per scheduler, and the butterfly needs 64 of them against 26
    node_allreduce_s<LPN, K>(v, ra);
}

// SMC: keep the slot cache in shared memory ([slot][2][lane] double2, conflict-free) instead of registers --
// used for the one-lane-per-node layouts with long slot lists (17..32 nodes, up to 12 slots per lane), where the
// register cache costs 238+ registers (8 warps / SM); the loads do not depend on delta, so they stay off the
// critical path
template <int LPN, int SPL, bool SMC, bool LAT>
__device__ __forceinline__ void rtr_fast_body(const RtrArgs &a, const uint32_t *__restrict__ fast_info,
                                              const double *__restrict__ fast_target)
{
    constexpr int NPW = 32 / LPN;  // node slots per warp
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int node = lane / LPN;
    const bool valid = node < a.N;
    const int goal_pad = (a.n_goal + 1) & ~1;
    double *P = smem;
    double *V = P + 3 * NPW;
    double *goal = V + 3 * NPW;
    double *tgt = goal + goal_pad + lane;   // [SPL][32] per-problem targets of this warp's slots
    // slot cache when SMC: [SPL][2][32 lanes] double2 {dx, dy}, {dz, c2} -- two conflict-free LDS.128 per slot
    double2 *scm = reinterpret_cast<double2 *>(goal + goal_pad + SPL * 32) + lane;
    double *R = goal + goal_pad + SPL * 32 + (SMC ? SPL * 128 : 0);   // reduction rows
    double *Tt = R + 8 * gik_red_stride(NPW);                                  // reduction totals
    unsigned long long *t_start = reinterpret_cast<unsigned long long *>(Tt + 8);
    const GikSolveOpts &o = a.o;

    // static slot description of this lane (identical for every problem): where the neighbour's coordinates sit in
    // the exchange buffers (kept as addresses so that the inner loop does no index arithmetic) and the term kinds
    uint32_t vj[SPL];        // shared address of V[neighbour]; the same node of P is 3 * NPW doubles below
    uint32_t kinds = 0;      // 2 bits per slot
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
        const uint32_t info = fast_info[s * 32 + lane];
        vj[s] = gik_saddr(V + GIK_SLOT_NBR(info));
        kinds |= GIK_SLOT_KIND(info) << (2 * s);
    }
    const uint32_t vown = gik_saddr(V + node);   // where this lane's node publishes
    const bool publisher = lane % LPN == 0;
    const GikRedAddr ra = gik_red_addr<LPN>(R, Tt, lane, 8);   // rows of a K = 4 reduction are the first four
    constexpr int PO = -3 * NPW * 8, CS = NPW * 8;              // byte offsets: V -> P, coordinate stride

    // parked problems of the incoming queue are resumed before any new problem starts
    int n_res = 0;
    if (a.carry_in) {
        const GikCarryHdr *h = reinterpret_cast<const GikCarryHdr *>(a.carry_in);
        n_res = min(h->count, h->capacity);
    }

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(GIK_FULL_MASK, w, 0, 32);
        if (w >= n_res + a.B) break;
        const bool resumed = w < n_res;
        const int b = w - n_res;
        const double *ent = resumed ? gik_carry_entry(a.carry_in, w) : nullptr;
        const unsigned long long *entp = reinterpret_cast<const unsigned long long *>(ent);
        const double *goal_row = resumed ? reinterpret_cast<const double *>(entp[CW_GOAL])
                                         : a.goal_d2 + (size_t)b * a.n_goal;

        double x[3] = {0.0, 0.0, 0.0}, g[3] = {0.0, 0.0, 0.0}, eta[3], Heta[3], r[3], dl[3], Hd[3];
        SlotCache sc[SMC ? 1 : SPL];
        if (valid) {
            const double *src = resumed ? ent + CW_X + 3 * node : a.Y_init + ((size_t)b * a.N + node) * 3;
            x[0] = src[0]; x[1] = src[1]; x[2] = src[2];
            if (resumed) {
                src = ent + CW_X + 3 * (a.N + node);
                g[0] = src[0]; g[1] = src[1]; g[2] = src[2];
            }
        }
        __syncwarp();
        for (int k = lane; k < a.n_goal; k += 32) goal[k] = goal_row[k];
        if (publisher) { gik_sts<PO>(vown, x[0]); gik_sts<PO + CS>(vown, x[1]); gik_sts<PO + 2 * CS>(vown, x[2]); }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const uint32_t gs = GIK_SLOT_GOAL(fast_info[s * 32 + lane]);
            tgt[s * 32] = gs ? goal[gs - 1] : fast_target[s * 32 + lane];
        }

        // cost / gradient at point p (published in P) + rebuild of the slot cache.
        // Returns the node's cost share (both lanes of a pair); gout = full half-gradient of the node.
        auto rebuild = [&](const double (&p)[3], double (&gout)[3]) -> double {
            double fpart = 0.0, gacc[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const uint32_t kind = (kinds >> (2 * s)) & 3u;
                const double dx = trm::sub(p[0], gik_lds<PO>(vj[s])), dy = trm::sub(p[1], gik_lds<PO + CS>(vj[s])),
                             dz = trm::sub(p[2], gik_lds<PO + 2 * CS>(vj[s]));
                const trm::SlotEval e = trm::slot_cost(dx, dy, dz, tgt[s * 32], kind, fpart, gacc);
                const double cx = trm::mul(e.two, dx), cy = trm::mul(e.two, dy), cz = trm::mul(e.two, dz),
                             c2 = trm::mul(2.0, e.rr);
                if (SMC) {
                    scm[(s * 2 + 0) * 32] = make_double2(cx, cy);
                    scm[(s * 2 + 1) * 32] = make_double2(cz, c2);
                } else {
                    sc[s].dx = cx; sc[s].dy = cy; sc[s].dz = cz; sc[s].c2 = c2;
                }
            }
            double v[4] = {trm::mul(2.0, gacc[0]), trm::mul(2.0, gacc[1]), trm::mul(2.0, gacc[2]), trm::mul(0.5, fpart)};
            pair_combine<LPN, 4>(v);
            gout[0] = v[0]; gout[1] = v[1]; gout[2] = v[2];
            return v[3];
        };
        // two K = 8 all-reduces of v[0..15]
        auto reduce16 = [&](double (&v)[16]) {
            double lo[8], hi[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { lo[k] = v[k]; hi[k] = v[8 + k]; }
            node_allreduce_s<LPN, 8>(lo, ra);
            node_allreduce_s<LPN, 8>(hi, ra);
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = lo[k]; v[8 + k] = hi[k]; }
        };

        double fx, gg, Mi[6], sg[3], Delta;
        int k_outer, inner_total;
        if (resumed) {
            fx = ent[CW_FX]; gg = ent[CW_GG]; Delta = ent[CW_DELTA];
#pragma unroll
            for (int k = 0; k < 6; ++k) Mi[k] = ent[CW_MI + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) sg[k] = ent[CW_SG + k];
            const unsigned long long cnt = entp[CW_COUNTS];
            k_outer = (int)(cnt & 0xffffffffu);
            inner_total = (int)(cnt >> 32);
            if (lane == 0) *t_start = entp[CW_T0];
            double gtmp[3];
            rebuild(x, gtmp);   // slot cache at x, as after a rejected step
        } else {
            double v[16];
            v[0] = rebuild(x, g);
            trm::point_scalars(x, g, v + 1);
#pragma unroll
            for (int k = 11; k < 16; ++k) v[k] = 0.0;
            reduce16(v);
            fx = v[0];
            gg = v[1];
            gik_sylvester_inverse(v + 2, Mi);
            sg[0] = v[8]; sg[1] = v[9]; sg[2] = v[10];
            Delta = o.Delta0;
            k_outer = 0;
            inner_total = 0;
            if (lane == 0) *t_start = a.maxtime_ns ? gik_globaltimer() : 0ull;
        }
        const int inner_entry = inner_total;
        bool may_park = a.carry_out != nullptr;
        int park_slot = -1;
        double norm_grad = sqrt(gg);
        int status = GIK_STATUS_MAXITER;
        if (!(isfinite(fx) && isfinite(gg))) {
            status = GIK_STATUS_NAN;
        } else {
            for (;;) {
                // ================= tCG (trust_region.py:436-599), eta0 = 0, precon = identity
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    eta[q] = 0.0; Heta[q] = 0.0; r[q] = g[q]; dl[q] = -g[q];
                }
                const trm::TcgStart ts = trm::tcg_start(gg, Delta, o);
                double e_Pe = 0.0, z_r = gg, d_Pd = gg, e_Pd = 0.0, model_value = 0.0, inv_z_r = ts.inv_z_r;
                // u = sum delta_i x Y_i.  H delta is horizontal (sum (H delta)_i x Y_i = 0 by construction of omega), so
                // sum r_i x Y_i keeps its starting value sum g_i x Y_i =: sg during the run and delta' = -r' + beta delta
                // gives u' = -sg + beta u.  (sg itself is rounding noise: the cost is rotation invariant.)
                double u[3] = {-sg[0], -sg[1], -sg[2]};
                int stop = MAX_INNER_ITER;
                int j = 0;
                for (j = 0; j < o.maxinner; ++j) {
                    // ---- Hdelta = proj(x, lhess(x, delta))
                    if (publisher) { gik_sts<0>(vown, dl[0]); gik_sts<CS>(vown, dl[1]); gik_sts<2 * CS>(vown, dl[2]); }
                    __syncwarp();
                    double z[3] = {0.0, 0.0, 0.0}, zb[3] = {0.0, 0.0, 0.0};
#pragma unroll
                    for (int s = 0; s < SPL; ++s) {
                        const double wx = trm::sub(dl[0], gik_lds<0>(vj[s])), wy = trm::sub(dl[1], gik_lds<CS>(vj[s])),
                                     wz = trm::sub(dl[2], gik_lds<2 * CS>(vj[s]));
                        double cx, cy, cz, c2;
                        if (SMC) {
                            const double2 a0 = scm[(s * 2 + 0) * 32], a1 = scm[(s * 2 + 1) * 32];
                            cx = a0.x; cy = a0.y; cz = a1.x; c2 = a1.y;
                        } else {
                            cx = sc[s].dx; cy = sc[s].dy; cz = sc[s].dz; c2 = sc[s].c2;
                        }
                        trm::slot_hess(cx, cy, cz, c2, wx, wy, wz, z, zb);
                    }
                    z[0] = trm::add(z[0], zb[0]); z[1] = trm::add(z[1], zb[1]); z[2] = trm::add(z[2], zb[2]);
                    pair_combine<LPN, 3>(z);
                    double v[4];
                    trm::hess_scalars(dl, z, x, v);
                    fast_allreduce<LPN, 4, LAT>(v, ra, lane);
                    const trm::InnerScalars is = trm::inner_scalars(Mi, v, u, z_r, e_Pe, e_Pd, d_Pd, ts.Delta2);
                    trm::project(z, x, is.om, Hd);
                    ++inner_total;
                    // is.leave also catches a NaN curvature (the reference would spin to maxinner on it)
                    if (is.leave) {
                        const double tau = trm::boundary_tau(e_Pe, e_Pd, d_Pd, ts.Delta2);
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            eta[q] = fma(tau, dl[q], eta[q]);
                            Heta[q] = fma(tau, Hd[q], Heta[q]);
                        }
                        stop = is.d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                        break;
                    }
                    e_Pe = is.e_Pe_new;
                    double ne[3], nh[3], nr[3], sdot[4];
                    trm::inner_step(is.alpha, dl, Hd, eta, Heta, r, g, ne, nh, nr, sdot);
                    fast_allreduce<LPN, 4, LAT>(sdot, ra, lane);
                    const double new_model_value = trm::model_value(sdot);
                    if (new_model_value >= model_value) {
                        stop = MODEL_INCREASED;
                        break;
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) { eta[q] = ne[q]; Heta[q] = nh[q]; r[q] = nr[q]; }
                    model_value = new_model_value;
                    const double r_r = sdot[2];
                    if (j >= o.mininner && r_r <= ts.r_target2) {
                        stop = o.kappa < ts.pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
                        break;
                    }
                    const trm::NextDir nd = trm::next_direction(r_r, z_r, inv_z_r, is.alpha, e_Pd, d_Pd);
                    z_r = r_r;
                    inv_z_r = gik_rcp(z_r);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        dl[q] = fma(nd.beta, dl[q], -r[q]);
                        u[q] = fma(nd.beta, u[q], -sg[q]);
                    }
                    e_Pd = nd.e_Pd;
                    d_Pd = nd.d_Pd;
                }
                const int numit = j < o.maxinner ? j : o.maxinner - 1;

                // ================= proposal x + eta; dl <- x_prop, Hd <- grad(x_prop), cache <- x_prop
                double v[16];
                v[11] = trm::dot3(g, eta);
                v[12] = trm::dot3(eta, Heta);
#pragma unroll
                for (int q = 0; q < 3; ++q) dl[q] = trm::add(x[q], eta[q]);
                if (publisher) { gik_sts<PO>(vown, dl[0]); gik_sts<PO + CS>(vown, dl[1]); gik_sts<PO + 2 * CS>(vown, dl[2]); }
                __syncwarp();
                v[0] = rebuild(dl, Hd);
                trm::point_scalars(dl, Hd, v + 1);
                v[13] = 0.0; v[14] = 0.0; v[15] = 0.0;
                reduce16(v);
                const double fx_prop = v[0];
                const double Delta_used = Delta;
                const trm::OuterDecision od = trm::outer_decision(fx, fx_prop, v[11], v[12], Delta, stop, o);
                Delta = od.Delta;
                const bool accept = od.accept;
                if (accept) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { x[q] = dl[q]; g[q] = Hd[q]; }
                    fx = fx_prop;
                    gg = v[1];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 2, Mi);
                    sg[0] = v[8]; sg[1] = v[9]; sg[2] = v[10];
                } else {
                    // rejected: bring the exchange buffer and the slot cache back to x
                    __syncwarp();
                    if (publisher) { gik_sts<PO>(vown, x[0]); gik_sts<PO + CS>(vown, x[1]); gik_sts<PO + 2 * CS>(vown, x[2]); }
                    __syncwarp();
                    double gtmp[3];
                    rebuild(x, gtmp);
                }
                if (a.trace && !resumed && k_outer < a.trace_rows && lane == 0) {
                    double *row = a.trace + ((size_t)b * a.trace_rows + k_outer) * 6;
                    row[0] = Delta_used;
                    row[1] = (double)numit;
                    row[2] = (double)stop;
                    row[3] = fx_prop;
                    row[4] = accept ? 1.0 : 0.0;
                    row[5] = accept ? norm_grad : nan("");
                }
                ++k_outer;
                // pymanopt Solver._check_stopping_criterion: maxtime, then maxiter, then mingradnorm
                if (a.maxtime_ns) {
                    unsigned long long now = 0;
                    if (lane == 0) now = gik_globaltimer() - *t_start;
                    now = __shfl_sync(GIK_FULL_MASK, now, 0, 32);
                    if (now >= a.maxtime_ns) { status = GIK_STATUS_MAXTIME; break; }
                }
                if (k_outer >= o.maxiter) { status = GIK_STATUS_MAXITER; break; }
                if (norm_grad < o.mingradnorm) { status = GIK_STATUS_CONVERGED; break; }
                if (may_park && inner_total - inner_entry >= a.inner_budget) {
                    if (lane == 0) park_slot = gik_try_park(a, n_res + a.B);
                    park_slot = __shfl_sync(GIK_FULL_MASK, park_slot, 0, 32);
                    if (park_slot >= 0) { status = GIK_STATUS_PENDING; break; }
                    if (park_slot == -1) may_park = false;   // queue full: run this problem to its end
                }
            }
        }

        // ---- final values (or, for a parked problem, its current ones) go where the problem came from
        double *Yrow = resumed ? reinterpret_cast<double *>(entp[CW_Y]) : a.Y_out + (size_t)b * a.N * 3;
        double *cx = status == GIK_STATUS_PENDING ? gik_carry_slot(a.carry_out, park_slot) : nullptr;
        if (valid && publisher) {
            double *dst = Yrow + node * 3;
            dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
            if (cx) {
                dst = cx + CW_X + 3 * node;
                dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
                dst += 3 * a.N;
                dst[0] = g[0]; dst[1] = g[1]; dst[2] = g[2];
            }
        }
        if (lane == 0)
            gik_finish_problem(a, resumed, b, entp, goal_row, cx, *t_start, status, k_outer, inner_total, fx, gg,
                               norm_grad, Delta, Mi, sg, Yrow);
        __syncwarp();
    }
}

// Where the slot cache lives and how many warps fit: the paired layout (N <= 16, <= 6 slots per lane) keeps it in
// registers at 168 registers = 12 warps / SM; one lane per node with <= 9 slots (7-DOF arms: KUKA, LWA4D, Panda)
// keeps it in registers too, at 238 registers = 8 warps / SM -- measured on KUKA IIWA against the shared-memory cache
// at 12 warps / SM: same throughput at 65 536 goals (1439 ms), 9 % lower latency at 16 384 (437 vs 480 ms), because the
// kernel is bound by FP64 / MIO issue and dependent latency, not by resident warps; longer slot lists use shared memory.
#ifndef GIK_FAST_SMC2   // A/B builds only: slot cache of the paired layouts in shared memory as well
#define GIK_FAST_SMC2 0
#endif
__host__ __device__ constexpr bool fast_cache_in_smem(int LPN, int SPL) { return LPN == 1 ? SPL > 9 : (GIK_FAST_SMC2 || SPL > 6); }
#ifndef GIK_FAST_MINB   // A/B builds only: resident warps per SM of the paired / shared-memory-cache layouts
#define GIK_FAST_MINB 12
#endif
__host__ __device__ constexpr int fast_min_blocks(int LPN, int SPL) { return (LPN == 1 && SPL <= 9) ? 8 : GIK_FAST_MINB; }

// LAT = false: throughput variant (as many warps per SM as the registers allow, shared-memory reductions).
// LAT = true: latency variant for launches with few problems -- small batches and the draining launches of
// gik_rtr_solve_sliced, where the wall time is the dependency chain of the slowest problems: 8 warps / SM (255
// registers: no rematerialised addresses, more loads in flight).  Both variants compile the same explicitly rounded
// arithmetic (gik_tr_math.cuh) and give bit-identical results.
template <int LPN, int SPL, bool LAT>
__global__ void __launch_bounds__(kThreads, LAT ? 8 : fast_min_blocks(LPN, SPL))
k_rtr_fast(const RtrArgs a, const uint32_t *__restrict__ fast_info, const double *__restrict__ fast_target)
{
    rtr_fast_body<LPN, SPL, fast_cache_in_smem(LPN, SPL), LAT>(a, fast_info, fast_target);
}

template <int LPN, int SPL, bool LAT>
int launch_variant(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    auto kern = k_rtr_fast<LPN, SPL, LAT>;
    constexpr bool SMC = fast_cache_in_smem(LPN, SPL);
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t smem = (size_t)fast_smem_doubles(LPN, SPL, SMC, goal_pad) * sizeof(double);
    // shared-memory opt-in and occupancy are properties of (kernel, smem size, device): looked up once
    static size_t cached_smem = ~(size_t)0;
    static int cached_per_sm = 0, cached_dev = -1;
    if (cached_smem != smem || cached_dev != p->device) {
        if (smem > 48 * 1024)
            GIK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
        cached_per_sm = per_sm < 1 ? 1 : per_sm;
        cached_smem = smem;
        cached_dev = p->device;
    }
    int blocks = p->sm_count * cached_per_sm;
    if (!a.carry_in && blocks > a.B) blocks = a.B;   // the number of parked problems is only known on the device
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    kern<<<blocks, kThreads, smem, st>>>(a, p->fast_info, p->fast_target);
    return gik_check_cuda(cudaGetLastError(), "k_rtr_fast launch");
}

// Which variant: a launch that only drains a queue, or a one-piece solve of a batch small enough that its wall
// time is its slowest problem (measured on UR10: 144 vs 178 ms at 4096 goals, break-even near 40 k goals).
template <int LPN, int SPL>
int launch(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    const bool few = a.carry_in ? a.B == 0 : (a.carry_out == nullptr && a.B <= 32768);
    return few ? launch_variant<LPN, SPL, true>(p, a, st) : launch_variant<LPN, SPL, false>(p, a, st);
}


// ---------------------------------------------------------------------------------------------------
// 33 .. 64 nodes, sparse (e.g. the 20-DOF chain of BASELINE configs[3]: N = 44, 115 terms, degree 5..7):
// one warp per problem, TWO nodes per lane.  The plan orders the nodes by degree: the 32 highest-degree
// nodes are every lane's first node (S0 slots), the rest the second node of lanes 0 .. N-33 (S1 slots),
// so a lane walks S0 + S1 slots instead of the 2 * maxdeg of a fixed l / l+32 assignment.  Slot cache in
// shared memory ([slot][4][lane], conflict-free; loads independent of delta), node state in registers.
__host__ __device__ constexpr int fast2_smem_doubles(int ST, int goal_pad)
{
    return 6 * 64 + goal_pad + ST * 32 + ST * 128 + 8 * gik_red_stride(32) + 8 + 2;
}

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
    double *R = goal + goal_pad + ST * 32 + ST * 128;   // reduction rows
    double *Tt = R + 8 * gik_red_stride(32);             // reduction totals
    unsigned long long *t_start = reinterpret_cast<unsigned long long *>(Tt + 8);
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
    uint32_t vj[ST];         // shared address of V[neighbour]; the same node of P is 3 * NPW doubles below
    uint32_t kinds = 0;      // 2 bits per slot (ST <= 16 in every instantiation with more than 16 slots the kinds
    uint32_t kinds_hi = 0;   // of slots 16.. go to the second word)
#pragma unroll
    for (int s = 0; s < ST; ++s) {
        const uint32_t info = info_tbl[s * 32 + lane];
        vj[s] = gik_saddr(V + GIK_SLOT_NBR(info));
        if (s < 16) kinds |= GIK_SLOT_KIND(info) << (2 * s);
        else kinds_hi |= GIK_SLOT_KIND(info) << (2 * (s - 16));
    }
    const uint32_t vown[2] = {gik_saddr(V + nd[0]), gik_saddr(V + nd[1])};
    const GikRedAddr ra = gik_red_addr<1>(R, Tt, lane, 8);
    constexpr int PO = -3 * NPW * 8, CS = NPW * 8;   // byte offsets: V -> P, coordinate stride

    auto publish = [&](int off, const double (&v)[2][3]) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (off == 0) { gik_sts<0>(vown[m], v[m][0]); gik_sts<CS>(vown[m], v[m][1]); gik_sts<2 * CS>(vown[m], v[m][2]); }
            else { gik_sts<PO>(vown[m], v[m][0]); gik_sts<PO + CS>(vown[m], v[m][1]); gik_sts<PO + 2 * CS>(vown[m], v[m][2]); }
        }
    };

    int n_res = 0;
    if (a.carry_in) {
        const GikCarryHdr *h = reinterpret_cast<const GikCarryHdr *>(a.carry_in);
        n_res = min(h->count, h->capacity);
    }

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(GIK_FULL_MASK, w, 0, 32);
        if (w >= n_res + a.B) break;
        const bool resumed = w < n_res;
        const int b = w - n_res;
        const double *ent = resumed ? gik_carry_entry(a.carry_in, w) : nullptr;
        const unsigned long long *entp = reinterpret_cast<const unsigned long long *>(ent);
        const double *goal_row = resumed ? reinterpret_cast<const double *>(entp[CW_GOAL])
                                         : a.goal_d2 + (size_t)b * a.n_goal;

        double x[2][3], g[2][3], eta[2][3], Heta[2][3], r[2][3], dl[2][3], Hd[2][3];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            x[m][0] = 0.0; x[m][1] = 0.0; x[m][2] = 0.0;
            g[m][0] = 0.0; g[m][1] = 0.0; g[m][2] = 0.0;
            if (valid[m]) {
                const double *src = resumed ? ent + CW_X + 3 * nd[m] : a.Y_init + ((size_t)b * a.N + nd[m]) * 3;
                x[m][0] = src[0]; x[m][1] = src[1]; x[m][2] = src[2];
                if (resumed) {
                    src = ent + CW_X + 3 * (a.N + nd[m]);
                    g[m][0] = src[0]; g[m][1] = src[1]; g[m][2] = src[2];
                }
            }
        }
        __syncwarp();
        for (int k = lane; k < a.n_goal; k += 32) goal[k] = goal_row[k];
        publish(1, x);
        __syncwarp();
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            const uint32_t gs = GIK_SLOT_GOAL(info_tbl[s * 32 + lane]);
            tgt[s * 32] = gs ? goal[gs - 1] : target_tbl[s * 32 + lane];
        }

        // cost / gradient at point p (published in P) + rebuild of the slot cache
        auto rebuild = [&](const double (&p)[2][3], double (&gout)[2][3]) -> double {
            double fpart = 0.0;
            double ga[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
            for (int s = 0; s < ST; ++s) {
                const int m = s < S0 ? 0 : 1;
                const uint32_t kind = s < 16 ? (kinds >> (2 * s)) & 3u : (kinds_hi >> (2 * (s - 16))) & 3u;
                const double dx = trm::sub(p[m][0], gik_lds<PO>(vj[s])), dy = trm::sub(p[m][1], gik_lds<PO + CS>(vj[s])),
                             dz = trm::sub(p[m][2], gik_lds<PO + 2 * CS>(vj[s]));
                const trm::SlotEval e = trm::slot_cost(dx, dy, dz, tgt[s * 32], kind, fpart, ga[m]);
                scm[(s * 2 + 0) * 32] = make_double2(trm::mul(e.two, dx), trm::mul(e.two, dy));
                scm[(s * 2 + 1) * 32] = make_double2(trm::mul(e.two, dz), trm::mul(2.0, e.rr));
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                gout[m][0] = trm::mul(2.0, ga[m][0]); gout[m][1] = trm::mul(2.0, ga[m][1]);
                gout[m][2] = trm::mul(2.0, ga[m][2]);
            }
            return trm::mul(0.5, fpart);
        };
        // this lane's share (two nodes) of the scalars of an accepted iterate: <h,h>, y^T y (6), sum h_i x y_i (3)
        auto point_scalars = [&](const double (&y)[2][3], const double (&h)[2][3], double *v) {
            double v0[10], v1[10];
            trm::point_scalars(y[0], h[0], v0);
            trm::point_scalars(y[1], h[1], v1);
#pragma unroll
            for (int k = 0; k < 10; ++k) v[k] = trm::add(v0[k], v1[k]);
        };
        auto reduce16 = [&](double (&v)[16]) {
            double lo[8], hi[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { lo[k] = v[k]; hi[k] = v[8 + k]; }
            node_allreduce_s<1, 8>(lo, ra);
            node_allreduce_s<1, 8>(hi, ra);
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = lo[k]; v[8 + k] = hi[k]; }
        };

        double fx, gg, Mi[6], sg[3], Delta;
        int k_outer, inner_total;
        if (resumed) {
            fx = ent[CW_FX]; gg = ent[CW_GG]; Delta = ent[CW_DELTA];
#pragma unroll
            for (int k = 0; k < 6; ++k) Mi[k] = ent[CW_MI + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) sg[k] = ent[CW_SG + k];
            const unsigned long long cnt = entp[CW_COUNTS];
            k_outer = (int)(cnt & 0xffffffffu);
            inner_total = (int)(cnt >> 32);
            if (lane == 0) *t_start = entp[CW_T0];
            double gtmp[2][3];
            rebuild(x, gtmp);   // slot cache at x, as after a rejected step
        } else {
            double v[16];
            v[0] = rebuild(x, g);
            point_scalars(x, g, v + 1);
#pragma unroll
            for (int k = 11; k < 16; ++k) v[k] = 0.0;
            reduce16(v);
            fx = v[0];
            gg = v[1];
            gik_sylvester_inverse(v + 2, Mi);
            sg[0] = v[8]; sg[1] = v[9]; sg[2] = v[10];
            Delta = o.Delta0;
            k_outer = 0;
            inner_total = 0;
            if (lane == 0) *t_start = a.maxtime_ns ? gik_globaltimer() : 0ull;
        }
        const int inner_entry = inner_total;
        bool may_park = a.carry_out != nullptr;
        int park_slot = -1;
        double norm_grad = sqrt(gg);
        int status = GIK_STATUS_MAXITER;
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
                const trm::TcgStart ts = trm::tcg_start(gg, Delta, o);
                double e_Pe = 0.0, z_r = gg, d_Pd = gg, e_Pd = 0.0, model_value = 0.0, inv_z_r = ts.inv_z_r;
                double u[3] = {-sg[0], -sg[1], -sg[2]};   // sum delta_i x Y_i by recurrence, see k_rtr_fast
                int stop = MAX_INNER_ITER;
                int j = 0;
                for (j = 0; j < o.maxinner; ++j) {
                    // ---- Hdelta = proj(x, lhess(x, delta))
                    publish(0, dl);
                    __syncwarp();
                    double z[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}, zb[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
                    for (int s = 0; s < ST; ++s) {
                        const int m = s < S0 ? 0 : 1;
                        const double wx = trm::sub(dl[m][0], gik_lds<0>(vj[s])), wy = trm::sub(dl[m][1], gik_lds<CS>(vj[s])),
                                     wz = trm::sub(dl[m][2], gik_lds<2 * CS>(vj[s]));
                        const double2 a0 = scm[(s * 2 + 0) * 32], a1 = scm[(s * 2 + 1) * 32];
                        trm::slot_hess(a0.x, a0.y, a1.x, a1.y, wx, wy, wz, z[m], zb[m]);
                    }
                    double v[4], v1[4];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        z[m][0] = trm::add(z[m][0], zb[m][0]); z[m][1] = trm::add(z[m][1], zb[m][1]);
                        z[m][2] = trm::add(z[m][2], zb[m][2]);
                    }
                    trm::hess_scalars(dl[0], z[0], x[0], v);
                    trm::hess_scalars(dl[1], z[1], x[1], v1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = trm::add(v[k], v1[k]);
                    fast_allreduce<1, 4, false>(v, ra, lane);
                    const trm::InnerScalars is = trm::inner_scalars(Mi, v, u, z_r, e_Pe, e_Pd, d_Pd, ts.Delta2);
                    trm::project(z[0], x[0], is.om, Hd[0]);
                    trm::project(z[1], x[1], is.om, Hd[1]);
                    ++inner_total;
                    if (is.leave) {   // negative curvature, boundary, or NaN
                        const double tau = trm::boundary_tau(e_Pe, e_Pd, d_Pd, ts.Delta2);
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                eta[m][q] = fma(tau, dl[m][q], eta[m][q]);
                                Heta[m][q] = fma(tau, Hd[m][q], Heta[m][q]);
                            }
                        }
                        stop = is.d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                        break;
                    }
                    e_Pe = is.e_Pe_new;
                    double ne[2][3], nh[2][3], nr[2][3], sdot[4], sdot1[4];
                    trm::inner_step(is.alpha, dl[0], Hd[0], eta[0], Heta[0], r[0], g[0], ne[0], nh[0], nr[0], sdot);
                    trm::inner_step(is.alpha, dl[1], Hd[1], eta[1], Heta[1], r[1], g[1], ne[1], nh[1], nr[1], sdot1);
#pragma unroll
                    for (int k = 0; k < 3; ++k) sdot[k] = trm::add(sdot[k], sdot1[k]);
                    fast_allreduce<1, 4, false>(sdot, ra, lane);
                    const double new_model_value = trm::model_value(sdot);
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
                    const double r_r = sdot[2];
                    if (j >= o.mininner && r_r <= ts.r_target2) {
                        stop = o.kappa < ts.pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
                        break;
                    }
                    const trm::NextDir nd = trm::next_direction(r_r, z_r, inv_z_r, is.alpha, e_Pd, d_Pd);
                    z_r = r_r;
                    inv_z_r = gik_rcp(z_r);
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) dl[m][q] = fma(nd.beta, dl[m][q], -r[m][q]);
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) u[q] = fma(nd.beta, u[q], -sg[q]);
                    e_Pd = nd.e_Pd;
                    d_Pd = nd.d_Pd;
                }
                const int numit = j < o.maxinner ? j : o.maxinner - 1;

                // ================= proposal x + eta; dl <- x_prop, Hd <- grad(x_prop), cache <- x_prop
                double v[16];
                v[11] = trm::add(trm::dot3(g[0], eta[0]), trm::dot3(g[1], eta[1]));
                v[12] = trm::add(trm::dot3(eta[0], Heta[0]), trm::dot3(eta[1], Heta[1]));
#pragma unroll
                for (int m = 0; m < 2; ++m) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) dl[m][q] = trm::add(x[m][q], eta[m][q]);
                }
                __syncwarp();
                publish(1, dl);
                __syncwarp();
                v[0] = rebuild(dl, Hd);
                point_scalars(dl, Hd, v + 1);
                v[13] = 0.0; v[14] = 0.0; v[15] = 0.0;
                reduce16(v);
                const double fx_prop = v[0];
                const double Delta_used = Delta;
                const trm::OuterDecision od = trm::outer_decision(fx, fx_prop, v[11], v[12], Delta, stop, o);
                Delta = od.Delta;
                const bool accept = od.accept;
                if (accept) {
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) { x[m][q] = dl[m][q]; g[m][q] = Hd[m][q]; }
                    }
                    fx = fx_prop;
                    gg = v[1];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 2, Mi);
                    sg[0] = v[8]; sg[1] = v[9]; sg[2] = v[10];
                } else {
                    // rejected: bring the exchange buffer and the slot cache back to x
                    __syncwarp();
                    publish(1, x);
                    __syncwarp();
                    double gtmp[2][3];
                    rebuild(x, gtmp);
                }
                if (a.trace && !resumed && k_outer < a.trace_rows && lane == 0) {
                    double *row = a.trace + ((size_t)b * a.trace_rows + k_outer) * 6;
                    row[0] = Delta_used;
                    row[1] = (double)numit;
                    row[2] = (double)stop;
                    row[3] = fx_prop;
                    row[4] = accept ? 1.0 : 0.0;
                    row[5] = accept ? norm_grad : nan("");
                }
                ++k_outer;
                // pymanopt Solver._check_stopping_criterion: maxtime, then maxiter, then mingradnorm
                if (a.maxtime_ns) {
                    unsigned long long now = 0;
                    if (lane == 0) now = gik_globaltimer() - *t_start;
                    now = __shfl_sync(GIK_FULL_MASK, now, 0, 32);
                    if (now >= a.maxtime_ns) { status = GIK_STATUS_MAXTIME; break; }
                }
                if (k_outer >= o.maxiter) { status = GIK_STATUS_MAXITER; break; }
                if (norm_grad < o.mingradnorm) { status = GIK_STATUS_CONVERGED; break; }
                if (may_park && inner_total - inner_entry >= a.inner_budget) {
                    if (lane == 0) park_slot = gik_try_park(a, n_res + a.B);
                    park_slot = __shfl_sync(GIK_FULL_MASK, park_slot, 0, 32);
                    if (park_slot >= 0) { status = GIK_STATUS_PENDING; break; }
                    if (park_slot == -1) may_park = false;   // queue full: run this problem to its end
                }
                __syncwarp();
            }
        }

        // ---- final values (or, for a parked problem, its current ones) go where the problem came from
        double *Yrow = resumed ? reinterpret_cast<double *>(entp[CW_Y]) : a.Y_out + (size_t)b * a.N * 3;
        double *cx = status == GIK_STATUS_PENDING ? gik_carry_slot(a.carry_out, park_slot) : nullptr;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (valid[m]) {
                double *dst = Yrow + nd[m] * 3;
                dst[0] = x[m][0]; dst[1] = x[m][1]; dst[2] = x[m][2];
                if (cx) {
                    dst = cx + CW_X + 3 * nd[m];
                    dst[0] = x[m][0]; dst[1] = x[m][1]; dst[2] = x[m][2];
                    dst += 3 * a.N;
                    dst[0] = g[m][0]; dst[1] = g[m][1]; dst[2] = g[m][2];
                }
            }
        }
        if (lane == 0) {
            gik_finish_problem(a, resumed, b, entp, goal_row, cx, *t_start, status, k_outer, inner_total, fx, gg,
                               norm_grad, Delta, Mi, sg, Yrow);
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
    const size_t smem = (size_t)fast2_smem_doubles(ST, goal_pad) * sizeof(double);
    if (smem > 227 * 1024) return 1;
    static size_t cached_smem = ~(size_t)0;
    static int cached_per_sm = 0, cached_dev = -1;
    if (cached_smem != smem || cached_dev != p->device) {
        if (smem > 48 * 1024)
            GIK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
        cached_per_sm = per_sm < 1 ? 1 : per_sm;
        cached_smem = smem;
        cached_dev = p->device;
    }
    int blocks = p->sm_count * cached_per_sm;
    if (!a.carry_in && blocks > a.B) blocks = a.B;
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
