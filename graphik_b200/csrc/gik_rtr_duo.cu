// gik_rtr_duo.cu -- throughput kernel for graphs with N <= 16 nodes: TWO problems per warp.
//
// Same algorithm as k_rtr_fast (reference trust_region.py:112-599, costs.py:79-207,
// fixed_rank_psd_sym.py:91-137) and -- through gik_tr_math.cuh -- the same floating-point operations in
// the same order: a problem solved here, by k_rtr_fast<2, S> or partly by one and partly by the other
// (gik_rtr_solve_sliced parks and resumes problems across launches) ends with the same bits.
//
// k_rtr_fast spends a whole warp on one problem: best for the critical path of a single problem, but
// half of its instructions compute values every lane of the warp already has (the scalar recurrences
// of tCG, the 3 x 3 projection, the reductions).  Here each HALF-warp owns one problem, lane <-> node,
// and one instruction stream serves both:
//
//   * the inner tCG iteration is straight-line code executed by both halves together; its two
//     reductions go through shared memory once for both problems (gik_warp.cuh: node_allreduce_s
//     layout, 4 lanes per (half, scalar)); the three ways an iteration can end only set the half's
//     phase, state commits are per-lane predicated;
//   * what happens once per outer iteration (proposal cost / gradient, rho test, radius update,
//     accept / reject, parking) and once per problem (fetch from the queue or from the carry queue,
//     initial cost / gradient, final store) runs in divergent sections guarded by a warp vote;
//   * a lane evaluates ALL slots of its node (k_rtr_fast<2, S> splits them over two lanes) but keeps
//     the two lanes' accumulator chains apart (slot k goes to chain k & 1) and adds them at the end --
//     exactly the sum k_rtr_fast forms with its pair shuffle, and twice the instruction-level
//     parallelism in the slot loop;
//   * the x-dependent part of each term (2 act D, 2 act r) is cached per accepted iterate in shared
//     memory ([slot][2][lane] double2, conflict-free); the loads do not depend on delta.
//
// A half whose phase is not INNER executes the inner block on dead state; nothing it computes is
// committed or stored.  A stalled problem (short tCG runs) makes its partner wait during its outer
// sections -- that costs latency, not throughput: the instructions issued per step are the same as for
// two separate warps.  The kernel is therefore used where latency is hidden: large one-piece batches
// and the bulk launches of gik_rtr_solve_sliced; small batches and draining launches use k_rtr_fast.
#include <cstdlib>

#include "gik_rtr.cuh"
#include "gik_tr_math.cuh"
#include "gik_warp.cuh"

namespace {

constexpr int kThreads = 32;
constexpr int kRS = 34;   // row stride of the reduction buffer: 128-bit reads of a quarter-warp hit disjoint banks
enum { PH_NEED_PROBLEM = 0, PH_NEED_OUTER = 1, PH_INNER = 2, PH_IDLE = 3 };

// plain butterfly inside one half (lane = node: neighbours, quads, eights, all sixteen -- the tree of
// node_allreduce_s / node_allreduce_b over 16 node columns), used by the divergent sections
template <int K>
__device__ __forceinline__ void half_allreduce(double (&v)[K], unsigned gmask)
{
#pragma unroll
    for (int off = 1; off <= 8; off <<= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(gmask, v[k], off, 32);
    }
}

__host__ __device__ constexpr int duo_smem_doubles(int SPL, int goal_pad)
{
    // P, V [3][32]; goal rows of the two halves; per-problem slot targets [SPL][32]; slot cache [SPL][2][32] double2;
    // reduction rows [4][kRS] + totals [8]; start times [2]
    return 192 + 2 * goal_pad + SPL * 32 + SPL * 128 + 4 * kRS + 8 + 2;
}

// Where the slot cache lives.  Measured on UR10 (SPL = 9, 65536 goals, maxiter 300): 12 warps / SM with the cache in
// shared memory and 8 warps / SM (248 registers) run at the same 2.11 G tCG iterations / s -- the shared-memory pipe is
// the common limit (168 wavefronts per trip, 43 % of them cache reads) -- so up to 9 slots the cache moves into the
// registers that 8 warps / SM leave free.
#ifndef GIK_DUO_MINB
#define GIK_DUO_MINB 8     // resident warps per SM with the slot cache in registers (A/B: tools/, -DGIK_DUO_MINB=...)
#endif
#ifndef GIK_DUO_REGCACHE
#define GIK_DUO_REGCACHE 1
#endif
__host__ __device__ constexpr bool duo_cache_in_regs(int SPL) { return GIK_DUO_REGCACHE && SPL <= 9; }

template <int SPL>
__global__ void __launch_bounds__(kThreads, duo_cache_in_regs(SPL) ? GIK_DUO_MINB : 12) k_rtr_duo(const RtrArgs a, const uint32_t *__restrict__ duo_info,
                                                          const double *__restrict__ duo_target)
{
    constexpr bool RC = duo_cache_in_regs(SPL);
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int node = lane & 15;
    const int hbase = lane & 16;                       // first lane of this half
    const int half = lane >> 4;
    const unsigned gmask = hbase ? 0xffff0000u : 0x0000ffffu;
    const bool valid = node < a.N;
    const int goal_pad = (a.n_goal + 1) & ~1;
    double *P = smem;
    double *V = P + 96;
    double *goal = V + 96 + half * goal_pad;
    double *tgt = V + 96 + 2 * goal_pad + lane;                                               // [SPL][32]
    double2 *scm = reinterpret_cast<double2 *>(V + 96 + 2 * goal_pad + SPL * 32) + lane;      // [SPL][2][32]
    double *R = V + 96 + 2 * goal_pad + SPL * 32 + SPL * 128;
    double *Tt = R + 4 * kRS;
    unsigned long long *t_start = reinterpret_cast<unsigned long long *>(Tt + 8) + half;
    const GikSolveOpts &o = a.o;

    // static slot description of this lane; neighbour addresses are rebased into this half's exchange columns
    uint32_t vj[SPL];
    uint32_t kinds = 0;
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
        const uint32_t info = duo_info[s * 16 + node];
        vj[s] = gik_saddr(V + hbase + GIK_SLOT_NBR(info));
        kinds |= GIK_SLOT_KIND(info) << (2 * s);
    }
    const uint32_t vown = gik_saddr(V + lane);
    constexpr int PO = -96 * 8, CS = 32 * 8;   // byte offsets: V -> P, coordinate stride
    // reduction addresses: every lane deposits into column `lane`; lane L sums 4 nodes of row (L >> 2) & 3 of its half
    const uint32_t r_dep = gik_saddr(R + lane);
    const uint32_t r_src = gik_saddr(R + ((lane >> 2) & 3) * kRS + hbase + (lane & 3) * 4);
    const uint32_t r_dst = gik_saddr(Tt + half * 4 + ((lane >> 2) & 3));
    const uint32_t r_tot = gik_saddr(Tt + half * 4);
    const bool r_writer = (lane & 3) == 0;
    // all-reduce of 4 scalars over the 16 nodes of EACH half at once; same tree as node_allreduce_s<2, 4>
    auto reduce4 = [&](double (&v)[4]) {
        gik_sts<0 * kRS * 8>(r_dep, v[0]); gik_sts<1 * kRS * 8>(r_dep, v[1]);
        gik_sts<2 * kRS * 8>(r_dep, v[2]); gik_sts<3 * kRS * 8>(r_dep, v[3]);
        __syncwarp();
        const double2 p0 = gik_lds2<0>(r_src), p1 = gik_lds2<16>(r_src);
        double s = __dadd_rn(__dadd_rn(p0.x, p0.y), __dadd_rn(p1.x, p1.y));
        s += __shfl_xor_sync(GIK_FULL_MASK, s, 1, 32);
        s += __shfl_xor_sync(GIK_FULL_MASK, s, 2, 32);
        if (r_writer) gik_sts<0>(r_dst, s);
        __syncwarp();
        const double2 t0 = gik_lds2<0>(r_tot), t1 = gik_lds2<16>(r_tot);
        v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
    };

    double x[3] = {0, 0, 0}, g[3] = {0, 0, 0}, eta[3] = {0, 0, 0}, Heta[3] = {0, 0, 0}, r[3] = {0, 0, 0},
           dl[3] = {0, 0, 0}, Hd[3];
    double sc[RC ? SPL : 1][4];   // slot cache in registers: 2 act (x_i - x_j), 2 act (d_ij - T_ij)
    if (RC) {
#pragma unroll
        for (int s = 0; s < (RC ? SPL : 1); ++s) { sc[s][0] = 0.0; sc[s][1] = 0.0; sc[s][2] = 0.0; sc[s][3] = 0.0; }
    }
    double fx = 0, gg = 0, norm_grad = 0, Mi[6] = {0, 0, 0, 0, 0, 0}, sg[3] = {0, 0, 0}, u[3] = {0, 0, 0}, Delta = 0;
    double e_Pe = 0, z_r = 1, inv_z_r = 1, d_Pd = 1, e_Pd = 0, model_value = 0;
    trm::TcgStart ts = {0, 0, 0, 1};
    int w = 0, phase = PH_NEED_PROBLEM, k_outer = 0, inner_total = 0, inner_entry = 0, status = 0,
        stop = MAX_INNER_ITER, j = 0, numit = 0;
    bool may_park = false;

    int n_res = 0;
    if (a.carry_in) {
        const GikCarryHdr *h = reinterpret_cast<const GikCarryHdr *>(a.carry_in);
        n_res = min(h->count, h->capacity);
    }

    // cost / gradient at p (published in P) + rebuild of the slot cache; returns the node's cost share.
    // Slot k accumulates into chain k & 1: the two lanes of k_rtr_fast<2, S>.
    auto rebuild = [&](const double (&p)[3], double (&gout)[3]) -> double {
        double fpart[2] = {0.0, 0.0}, gacc[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const uint32_t kind = (kinds >> (2 * s)) & 3u;
            const double dx = trm::sub(p[0], gik_lds<PO>(vj[s])), dy = trm::sub(p[1], gik_lds<PO + CS>(vj[s])),
                         dz = trm::sub(p[2], gik_lds<PO + 2 * CS>(vj[s]));
            const trm::SlotEval e = trm::slot_cost(dx, dy, dz, tgt[s * 32], kind, fpart[s & 1], gacc[s & 1]);
            if (RC) {
                sc[RC ? s : 0][0] = trm::mul(e.two, dx); sc[RC ? s : 0][1] = trm::mul(e.two, dy);
                sc[RC ? s : 0][2] = trm::mul(e.two, dz); sc[RC ? s : 0][3] = trm::mul(2.0, e.rr);
            } else {
                scm[(s * 2 + 0) * 32] = make_double2(trm::mul(e.two, dx), trm::mul(e.two, dy));
                scm[(s * 2 + 1) * 32] = make_double2(trm::mul(e.two, dz), trm::mul(2.0, e.rr));
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) gout[q] = trm::add(trm::mul(2.0, gacc[0][q]), trm::mul(2.0, gacc[1][q]));
        return trm::add(trm::mul(0.5, fpart[0]), trm::mul(0.5, fpart[1]));
    };

    // start of a trust-region subproblem (trust_region.py:436-490), eta0 = 0, precon = identity
    auto start_tcg = [&]() {
#pragma unroll
        for (int q = 0; q < 3; ++q) { eta[q] = 0.0; Heta[q] = 0.0; r[q] = g[q]; dl[q] = -g[q]; u[q] = -sg[q]; }
        ts = trm::tcg_start(gg, Delta, o);
        e_Pe = 0.0; z_r = gg; d_Pd = gg; e_Pd = 0.0; model_value = 0.0; inv_z_r = ts.inv_z_r;
        stop = MAX_INNER_ITER;
        j = 0;
    };

    // the problem of this half leaves the warp: final values, or its state into the outgoing queue (park_slot >= 0)
    auto leave_problem = [&](int park_slot) {
        const bool resumed = w < n_res;
        const int b = w - n_res;
        const double *ent = resumed ? gik_carry_entry(a.carry_in, w) : nullptr;
        const unsigned long long *entp = reinterpret_cast<const unsigned long long *>(ent);
        double *Yrow = resumed ? reinterpret_cast<double *>(entp[CW_Y]) : a.Y_out + (size_t)b * a.N * 3;
        double *cx = park_slot >= 0 ? gik_carry_slot(a.carry_out, park_slot) : nullptr;
        if (valid) {
            double *dst = Yrow + node * 3;
            dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
            if (cx) {
                dst = cx + CW_X + 3 * node;
                dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
                dst += 3 * a.N;
                dst[0] = g[0]; dst[1] = g[1]; dst[2] = g[2];
            }
        }
        if (node == 0) {
            const double *goal_row = resumed ? reinterpret_cast<const double *>(entp[CW_GOAL])
                                             : a.goal_d2 + (size_t)b * a.n_goal;
            gik_finish_problem(a, resumed, b, entp, goal_row, cx, *t_start, status, k_outer, inner_total, fx, gg,
                               norm_grad, Delta, Mi, sg, Yrow);
        }
    };

    for (;;) {
        // ================= A. halves without a problem pull the next one: parked problems first, then new ones
        if (__any_sync(GIK_FULL_MASK, phase == PH_NEED_PROBLEM)) {
            if (phase == PH_NEED_PROBLEM) {
                int nw = 0;
                if (node == 0) nw = atomicAdd(a.work_counter, 1);
                w = __shfl_sync(gmask, nw, hbase, 32);
                if (w >= n_res + a.B) {
                    phase = PH_IDLE;
                } else {
                    const bool resumed = w < n_res;
                    const int b = w - n_res;
                    const double *ent = resumed ? gik_carry_entry(a.carry_in, w) : nullptr;
                    const unsigned long long *entp = reinterpret_cast<const unsigned long long *>(ent);
                    const double *goal_row = resumed ? reinterpret_cast<const double *>(entp[CW_GOAL])
                                                     : a.goal_d2 + (size_t)b * a.n_goal;
                    x[0] = x[1] = x[2] = 0.0;
                    g[0] = g[1] = g[2] = 0.0;
                    if (valid) {
                        const double *src = resumed ? ent + CW_X + 3 * node : a.Y_init + ((size_t)b * a.N + node) * 3;
                        x[0] = src[0]; x[1] = src[1]; x[2] = src[2];
                        if (resumed) {
                            src = ent + CW_X + 3 * (a.N + node);
                            g[0] = src[0]; g[1] = src[1]; g[2] = src[2];
                        }
                    }
                    for (int k = node; k < a.n_goal; k += 16) goal[k] = goal_row[k];
                    gik_sts<PO>(vown, x[0]); gik_sts<PO + CS>(vown, x[1]); gik_sts<PO + 2 * CS>(vown, x[2]);
                    __syncwarp(gmask);
#pragma unroll
                    for (int s = 0; s < SPL; ++s) {
                        const uint32_t gs = GIK_SLOT_GOAL(duo_info[s * 16 + node]);
                        tgt[s * 32] = gs ? goal[gs - 1] : duo_target[s * 16 + node];
                    }
                    if (resumed) {
                        fx = ent[CW_FX]; gg = ent[CW_GG]; Delta = ent[CW_DELTA];
#pragma unroll
                        for (int k = 0; k < 6; ++k) Mi[k] = ent[CW_MI + k];
#pragma unroll
                        for (int k = 0; k < 3; ++k) sg[k] = ent[CW_SG + k];
                        const unsigned long long cnt = entp[CW_COUNTS];
                        k_outer = (int)(cnt & 0xffffffffu);
                        inner_total = (int)(cnt >> 32);
                        if (node == 0) *t_start = entp[CW_T0];
                        double gtmp[3];
                        rebuild(x, gtmp);   // slot cache at x, as after a rejected step
                    } else {
                        double v[16];
                        v[0] = rebuild(x, g);
                        trm::point_scalars(x, g, v + 1);
#pragma unroll
                        for (int k = 11; k < 16; ++k) v[k] = 0.0;
                        half_allreduce<11>(reinterpret_cast<double (&)[11]>(v), gmask);
                        fx = v[0];
                        gg = v[1];
                        gik_sylvester_inverse(v + 2, Mi);
                        sg[0] = v[8]; sg[1] = v[9]; sg[2] = v[10];
                        Delta = o.Delta0;
                        k_outer = 0;
                        inner_total = 0;
                        if (node == 0) *t_start = a.maxtime_ns ? gik_globaltimer() : 0ull;
                    }
                    inner_entry = inner_total;
                    may_park = a.carry_out != nullptr;
                    norm_grad = sqrt(gg);
                    if (!(isfinite(fx) && isfinite(gg))) {
                        status = GIK_STATUS_NAN;
                        leave_problem(-1);       // stays PH_NEED_PROBLEM: fetch again next tick
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
                double xp[3], gp[3], v[16];
                v[11] = trm::dot3(g, eta);
                v[12] = trm::dot3(eta, Heta);
#pragma unroll
                for (int q = 0; q < 3; ++q) xp[q] = trm::add(x[q], eta[q]);
                gik_sts<PO>(vown, xp[0]); gik_sts<PO + CS>(vown, xp[1]); gik_sts<PO + 2 * CS>(vown, xp[2]);
                __syncwarp(gmask);
                v[0] = rebuild(xp, gp);
                trm::point_scalars(xp, gp, v + 1);
                half_allreduce<13>(reinterpret_cast<double (&)[13]>(v), gmask);
                const double fx_prop = v[0];
                const double Delta_used = Delta;
                const trm::OuterDecision od = trm::outer_decision(fx, fx_prop, v[11], v[12], Delta, stop, o);
                Delta = od.Delta;
                const bool accept = od.accept;
                if (accept) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { x[q] = xp[q]; g[q] = gp[q]; }
                    fx = fx_prop;
                    gg = v[1];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 2, Mi);
                    sg[0] = v[8]; sg[1] = v[9]; sg[2] = v[10];
                } else {
                    __syncwarp(gmask);
                    gik_sts<PO>(vown, x[0]); gik_sts<PO + CS>(vown, x[1]); gik_sts<PO + 2 * CS>(vown, x[2]);
                    __syncwarp(gmask);
                    double gtmp[3];
                    rebuild(x, gtmp);
                }
                if (a.trace && w >= n_res && k_outer < a.trace_rows && node == 0) {
                    double *row = a.trace + ((size_t)(w - n_res) * a.trace_rows + k_outer) * 6;
                    row[0] = Delta_used;
                    row[1] = (double)numit;
                    row[2] = (double)stop;
                    row[3] = fx_prop;
                    row[4] = accept ? 1.0 : 0.0;
                    row[5] = accept ? norm_grad : nan("");
                }
                ++k_outer;
                // pymanopt Solver._check_stopping_criterion: maxtime, then maxiter, then mingradnorm
                bool timed_out = false;
                if (a.maxtime_ns) {
                    unsigned long long now = 0;
                    if (node == 0) now = gik_globaltimer() - *t_start;
                    now = __shfl_sync(gmask, now, hbase, 32);
                    timed_out = now >= a.maxtime_ns;
                }
                if (timed_out) {
                    status = GIK_STATUS_MAXTIME;
                    leave_problem(-1);
                    phase = PH_NEED_PROBLEM;
                } else if (k_outer >= o.maxiter) {
                    status = GIK_STATUS_MAXITER;
                    leave_problem(-1);
                    phase = PH_NEED_PROBLEM;
                } else if (norm_grad < o.mingradnorm) {
                    status = GIK_STATUS_CONVERGED;
                    leave_problem(-1);
                    phase = PH_NEED_PROBLEM;
                } else {
                    int park_slot = -1;
                    if (may_park && inner_total - inner_entry >= a.inner_budget) {
                        if (node == 0) park_slot = gik_try_park(a, n_res + a.B);
                        park_slot = __shfl_sync(gmask, park_slot, hbase, 32);
                        if (park_slot == -1) may_park = false;   // queue full: run this problem to its end
                    }
                    if (park_slot >= 0) {
                        status = GIK_STATUS_PENDING;
                        leave_problem(park_slot);
                        phase = PH_NEED_PROBLEM;
                    } else {
                        start_tcg();
                        phase = PH_INNER;
                    }
                }
            }
            __syncwarp();
        }

        // ================= C. one tCG iteration (trust_region.py:495-597), both halves together
        gik_sts<0>(vown, dl[0]); gik_sts<CS>(vown, dl[1]); gik_sts<2 * CS>(vown, dl[2]);
        __syncwarp();
        double z[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}, zb[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const double wx = trm::sub(dl[0], gik_lds<0>(vj[s])), wy = trm::sub(dl[1], gik_lds<CS>(vj[s])),
                         wz = trm::sub(dl[2], gik_lds<2 * CS>(vj[s]));
            if (RC) {
                trm::slot_hess(sc[RC ? s : 0][0], sc[RC ? s : 0][1], sc[RC ? s : 0][2], sc[RC ? s : 0][3], wx, wy, wz,
                               z[s & 1], zb[s & 1]);
            } else {
                const double2 a0 = scm[(s * 2 + 0) * 32], a1 = scm[(s * 2 + 1) * 32];
                trm::slot_hess(a0.x, a0.y, a1.x, a1.y, wx, wy, wz, z[s & 1], zb[s & 1]);
            }
        }
        double Z[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) Z[q] = trm::add(trm::add(z[0][q], zb[0][q]), trm::add(z[1][q], zb[1][q]));
        double v[4];
        trm::hess_scalars(dl, Z, x, v);
        reduce4(v);
        const trm::InnerScalars is = trm::inner_scalars(Mi, v, u, z_r, e_Pe, e_Pd, d_Pd, ts.Delta2);
        trm::project(Z, x, is.om, Hd);
        inner_total += phase == PH_INNER;
        // boundary / negative curvature (also a NaN curvature): rare, handled divergently
        const bool exit1 = phase == PH_INNER && is.leave;
        if (__any_sync(GIK_FULL_MASK, exit1)) {
            if (exit1) {
                const double tau = trm::boundary_tau(e_Pe, e_Pd, d_Pd, ts.Delta2);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    eta[q] = fma(tau, dl[q], eta[q]);
                    Heta[q] = fma(tau, Hd[q], Heta[q]);
                }
                stop = is.d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                numit = j;
                phase = PH_NEED_OUTER;
            }
            __syncwarp();
        }
        double ne[3], nh[3], nr[3], sdot[4];
        trm::inner_step(is.alpha, dl, Hd, eta, Heta, r, g, ne, nh, nr, sdot);
        reduce4(sdot);
        const double new_model_value = trm::model_value(sdot);
        const bool cont = phase == PH_INNER;
        const bool model_inc = cont && new_model_value >= model_value;
        const bool commit = cont && !model_inc;
        if (model_inc) { stop = MODEL_INCREASED; numit = j; phase = PH_NEED_OUTER; }
        const double r_r = sdot[2];
        if (commit) {
            e_Pe = is.e_Pe_new;
#pragma unroll
            for (int q = 0; q < 3; ++q) { eta[q] = ne[q]; Heta[q] = nh[q]; r[q] = nr[q]; }
            model_value = new_model_value;
        }
        const bool reached = commit && j >= o.mininner && r_r <= ts.r_target2;
        if (reached) {
            stop = o.kappa < ts.pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
            numit = j;
            phase = PH_NEED_OUTER;
        }
        const trm::NextDir nd = trm::next_direction(r_r, z_r, inv_z_r, is.alpha, e_Pd, d_Pd);
        if (commit && !reached) {
            z_r = r_r;
            inv_z_r = gik_rcp(z_r);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                dl[q] = fma(nd.beta, dl[q], -r[q]);
                u[q] = fma(nd.beta, u[q], -sg[q]);
            }
            e_Pd = nd.e_Pd;
            d_Pd = nd.d_Pd;
            ++j;
            if (j >= o.maxinner) { stop = MAX_INNER_ITER; numit = o.maxinner - 1; phase = PH_NEED_OUTER; }
        }
    }
}

template <int SPL>
int launch(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    auto kern = k_rtr_duo<SPL>;
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t smem = (size_t)duo_smem_doubles(SPL, goal_pad) * sizeof(double);
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
    const int need = (a.B + 1) / 2;
    if (!a.carry_in && blocks > need) blocks = need;   // the number of parked problems is only known on the device
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    kern<<<blocks, kThreads, smem, st>>>(a, p->duo_info, p->duo_target);
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
