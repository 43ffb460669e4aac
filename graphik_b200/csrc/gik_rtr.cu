// gik_rtr.cu -- persistent batched Riemannian trust-region solve.
//
// Replaces, for B problems per launch, the reference call stack
//   TrustRegions.solve                  (solvers/trust_region.py:112-434)
//   -> _truncated_conjugate_gradient    (solvers/trust_region.py:436-599)
//      -> problem.hess = proj(x, lhess) (utils/manifolds/fixed_rank_psd_sym.py:91-127,
//                                        solvers/costs.py:171-207)
//   -> problem.cost / problem.grad      (solvers/costs.py:79-169)
// on the manifold PSDFixedRank(N, 3) with metric <U,V> = sum U.*V and retraction
// Y + U (fixed_rank_psd_sym.py:75-79,137).
//
// One group of W lanes (gik_common.cuh) owns one problem from start to finish;
// x, g, eta, H eta, r, delta, H delta stay in registers, the only shared-memory
// traffic is the neighbour-coordinate exchange of the edge passes, and the only
// HBM traffic is Y_init in / Y_out out.  Problems are handed out through a
// global atomic counter, so a group that converges early immediately starts
// the next problem (outer iteration counts range from ~50 to 3000).
//
// Algebra that differs from the reference's literal evaluation order (results
// agree to rounding):
//   * proj uses the 3x3 system (tr(X) I - X) omega = sum_i Z_i x Y_i instead of
//     the 9x9 Sylvester system; X = Y^T Y is formed once per accepted iterate.
//   * <delta, H delta> = <delta, Z> - omega . sum_i (delta_i x Y_i) shares the
//     reduction that produces omega, so one inner iteration needs two
//     reductions (7 and 3 scalars) instead of six.
//   * pymanopt's redundant egrad(x) inside every Hessian call is dropped.
//   * cost and gradient of the proposal are evaluated in one pass.
#include <cstdlib>

#include "gik_common.cuh"
#include <cstdint>

#include "gik_rtr.cuh"

namespace {

constexpr int kThreads = 128;

template <int NPL>
__device__ __forceinline__ double dot3(const double (&a)[NPL][3], const double (&b)[NPL][3])
{
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < NPL; ++m)
        s = fma(a[m][0], b[m][0], fma(a[m][1], b[m][1], fma(a[m][2], b[m][2], s)));
    return s;
}

// X = Y^T Y packed (xx, xy, xz, yy, yz, zz), this lane's share
template <int NPL>
__device__ __forceinline__ void gram_partial(const double (&y)[NPL][3], double *X)
{
#pragma unroll
    for (int k = 0; k < 6; ++k) X[k] = 0.0;
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
        X[0] = fma(y[m][0], y[m][0], X[0]);
        X[1] = fma(y[m][0], y[m][1], X[1]);
        X[2] = fma(y[m][0], y[m][2], X[2]);
        X[3] = fma(y[m][1], y[m][1], X[3]);
        X[4] = fma(y[m][1], y[m][2], X[4]);
        X[5] = fma(y[m][2], y[m][2], X[5]);
    }
}

template <int W, int NPL>
__global__ void __launch_bounds__(kThreads) k_rtr(const RtrArgs a)
{
    constexpr int NP = W * NPL;
    extern __shared__ double smem[];

    // ---- stage the per-graph tables once per CTA
    const size_t tbl = (size_t)a.maxdeg * a.N;
    double *s_target = smem;
    uint32_t *s_info = reinterpret_cast<uint32_t *>(s_target + (a.tables_in_smem ? tbl : 0));
    int32_t *s_deg = reinterpret_cast<int32_t *>(s_info + (a.tables_in_smem ? tbl : 0));
    double *groups = reinterpret_cast<double *>(s_deg + NP + (((a.tables_in_smem ? tbl : 0) + NP) & 1));
    for (int k = threadIdx.x; k < NP; k += kThreads) s_deg[k] = a.deg[k];
    if (a.tables_in_smem) {
        for (size_t k = threadIdx.x; k < tbl; k += kThreads) {
            s_target[k] = a.slot_target[k];
            s_info[k] = a.slot_info[k];
        }
    }
    __syncthreads();

    const int gid = threadIdx.x / W;
    const int lane = threadIdx.x % W;
    const int goal_pad = (a.n_goal + 1) & ~1;
    double *base = groups + (size_t)gid * (6 * NP + goal_pad);
    GikGroupCtx c;
    c.slot_info = a.tables_in_smem ? s_info : a.slot_info;
    c.slot_target = a.tables_in_smem ? s_target : a.slot_target;
    c.deg = s_deg;
    c.P = base;
    c.V = base + 3 * NP;
    double *goal = base + 6 * NP;
    c.goal = goal;
    c.N = a.N;
    c.mask = gik_group_mask<W>();
    c.lane = lane;
    const unsigned mask = c.mask;
    const int leader = (threadIdx.x & 31) & ~(W - 1);
    const GikSolveOpts &o = a.o;
    const double eps = 2.220446049250313e-16;  // np.spacing(1), trust_region.py:293

    // parked problems of the incoming queue are resumed before any new problem starts (gik_rtr.cuh)
    int n_res = 0;
    if (a.carry_in) {
        const GikCarryHdr *h = reinterpret_cast<const GikCarryHdr *>(a.carry_in);
        n_res = min(h->count, h->capacity);
    }

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(mask, w, leader, 32);
        if (w >= n_res + a.B) break;
        const bool resumed = w < n_res;
        const int b = w - n_res;
        const double *ent = resumed ? gik_carry_entry(a.carry_in, w) : nullptr;
        const unsigned long long *entp = reinterpret_cast<const unsigned long long *>(ent);
        const double *goal_row = resumed ? reinterpret_cast<const double *>(entp[CW_GOAL])
                                         : a.goal_d2 + (size_t)b * a.n_goal;

        double x[NPL][3], g[NPL][3], eta[NPL][3], Heta[NPL][3], r[NPL][3], dl[NPL][3], Hd[NPL][3];
        // ---- load the problem
        {
            const double *src = resumed ? ent + CW_X : a.Y_init + (size_t)b * 3 * a.N;
#pragma unroll
            for (int m = 0; m < NPL; ++m) {
                const int i = lane + W * m;
                g[m][0] = g[m][1] = g[m][2] = 0.0;
                if (i < a.N) {
                    x[m][0] = src[3 * i]; x[m][1] = src[3 * i + 1]; x[m][2] = src[3 * i + 2];
                    if (resumed) {
                        const double *gs = src + 3 * a.N;
                        g[m][0] = gs[3 * i]; g[m][1] = gs[3 * i + 1]; g[m][2] = gs[3 * i + 2];
                    }
                } else {
                    x[m][0] = x[m][1] = x[m][2] = 0.0;
                }
            }
            __syncwarp(mask);
            for (int k = lane; k < a.n_goal; k += W) goal[k] = goal_row[k];
            gik_publish<W, NPL>(c.P, x, lane);
            __syncwarp(mask);
        }
        // fx = cost(x); fgradx = grad(x); norm_grad (trust_region.py:158-160)
        double fx, gg, Mi[6], Delta;
        int k_outer, inner_total;
        unsigned long long t0;
        if (resumed) {
            fx = ent[CW_FX]; gg = ent[CW_GG]; Delta = ent[CW_DELTA];
#pragma unroll
            for (int k = 0; k < 6; ++k) Mi[k] = ent[CW_MI + k];
            const unsigned long long cnt = entp[CW_COUNTS];
            k_outer = (int)(cnt & 0xffffffffu);
            inner_total = (int)(cnt >> 32);
            t0 = entp[CW_T0];
        } else {
            double v[8];
            v[0] = gik_pass_cost_grad<W, NPL>(c, x, g);
            v[1] = dot3<NPL>(g, g);
            gram_partial<NPL>(x, v + 2);
            gik_allreduce<W, 8>(v, mask);
            fx = v[0];
            gg = v[1];
            gik_sylvester_inverse(v + 2, Mi);
            Delta = o.Delta0;
            k_outer = 0;
            inner_total = 0;
            t0 = a.maxtime_ns ? gik_globaltimer() : 0ull;
            t0 = __shfl_sync(mask, t0, leader, 32);
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
                for (int m = 0; m < NPL; ++m)
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        eta[m][q] = 0.0;
                        Heta[m][q] = 0.0;
                        r[m][q] = g[m][q];
                        dl[m][q] = -g[m][q];
                    }
                double e_Pe = 0.0, r_r = gg;
                const double norm_r0 = sqrt(r_r);
                double z_r = r_r, d_Pd = r_r, e_Pd = 0.0, model_value = 0.0;
                const double pw = pow(norm_r0, o.theta);
                const double r_target = norm_r0 * fmin(pw, o.kappa);
                const double Delta2 = Delta * Delta;
                int stop = MAX_INNER_ITER;
                int j = 0;
                for (j = 0; j < o.maxinner; ++j) {
                    // Hdelta = proj(x, lhess(x, delta))
                    __syncwarp(mask);   // every lane has read the previous direction out of V
                    gik_publish<W, NPL>(c.V, dl, lane);
                    __syncwarp(mask);
                    gik_pass_hess<W, NPL>(c, x, dl, Hd);
                    double v[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
                    for (int m = 0; m < NPL; ++m) {
                        v[0] = fma(dl[m][0], Hd[m][0], fma(dl[m][1], Hd[m][1], fma(dl[m][2], Hd[m][2], v[0])));
                        // c = sum Z_i x Y_i
                        v[1] += Hd[m][1] * x[m][2] - Hd[m][2] * x[m][1];
                        v[2] += Hd[m][2] * x[m][0] - Hd[m][0] * x[m][2];
                        v[3] += Hd[m][0] * x[m][1] - Hd[m][1] * x[m][0];
                        // u = sum delta_i x Y_i
                        v[4] += dl[m][1] * x[m][2] - dl[m][2] * x[m][1];
                        v[5] += dl[m][2] * x[m][0] - dl[m][0] * x[m][2];
                        v[6] += dl[m][0] * x[m][1] - dl[m][1] * x[m][0];
                    }
                    gik_allreduce<W, 7>(v, mask);
                    double om[3];
                    gik_sym_mul(Mi, v + 1, om);
#pragma unroll
                    for (int m = 0; m < NPL; ++m) {
                        Hd[m][0] -= x[m][1] * om[2] - x[m][2] * om[1];
                        Hd[m][1] -= x[m][2] * om[0] - x[m][0] * om[2];
                        Hd[m][2] -= x[m][0] * om[1] - x[m][1] * om[0];
                    }
                    const double d_Hd = v[0] - (om[0] * v[4] + om[1] * v[5] + om[2] * v[6]);
                    ++inner_total;
                    if (!isfinite(d_Hd)) break;  // guard: the reference would spin to maxinner on NaN
                    const double alpha = z_r / d_Hd;
                    const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
                    if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {
                        const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
#pragma unroll
                        for (int m = 0; m < NPL; ++m)
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                eta[m][q] = fma(tau, dl[m][q], eta[m][q]);
                                Heta[m][q] = fma(tau, Hd[m][q], Heta[m][q]);
                            }
                        stop = d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                        break;
                    }
                    e_Pe = e_Pe_new;
                    // new_eta, new_Heta, r + alpha Hdelta evaluated on the fly (committed below)
                    double s[3] = {0, 0, 0};
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const double ne = fma(alpha, dl[m][q], eta[m][q]);
                            const double nh = fma(alpha, Hd[m][q], Heta[m][q]);
                            const double nr = fma(alpha, Hd[m][q], r[m][q]);
                            s[0] = fma(ne, g[m][q], s[0]);
                            s[1] = fma(ne, nh, s[1]);
                            s[2] = fma(nr, nr, s[2]);
                        }
                    gik_allreduce<W, 3>(s, mask);
                    const double new_model_value = s[0] + 0.5 * s[1];
                    if (new_model_value >= model_value) {
                        stop = MODEL_INCREASED;
                        break;
                    }
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            eta[m][q] = fma(alpha, dl[m][q], eta[m][q]);
                            Heta[m][q] = fma(alpha, Hd[m][q], Heta[m][q]);
                            r[m][q] = fma(alpha, Hd[m][q], r[m][q]);
                        }
                    model_value = new_model_value;
                    r_r = s[2];
                    const double norm_r = sqrt(r_r);
                    if (j >= o.mininner && norm_r <= r_target) {
                        stop = o.kappa < pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
                        break;
                    }
                    const double zold_rold = z_r;
                    z_r = r_r;
                    const double beta = z_r / zold_rold;
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) dl[m][q] = fma(beta, dl[m][q], -r[m][q]);
                    e_Pd = beta * (e_Pd + alpha * d_Pd);
                    d_Pd = z_r + beta * beta * d_Pd;
                }
                const int numit = j < o.maxinner ? j : o.maxinner - 1;

                // ================= proposal x + eta (trust_region.py:248-251); dl <- x_prop, Hd <- grad(x_prop)
#pragma unroll
                for (int m = 0; m < NPL; ++m)
#pragma unroll
                    for (int q = 0; q < 3; ++q) dl[m][q] = x[m][q] + eta[m][q];
                __syncwarp(mask);       // every lane has finished the edge passes that read P
                gik_publish<W, NPL>(c.P, dl, lane);
                __syncwarp(mask);
                double v[10];
                v[0] = gik_pass_cost_grad<W, NPL>(c, dl, Hd);
                v[1] = dot3<NPL>(g, eta);
                v[2] = dot3<NPL>(eta, Heta);
                v[3] = dot3<NPL>(Hd, Hd);
                gram_partial<NPL>(dl, v + 4);
                gik_allreduce<W, 10>(v, mask);
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
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            x[m][q] = dl[m][q];
                            g[m][q] = Hd[m][q];
                        }
                    fx = fx_prop;
                    gg = v[3];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 4, Mi);
                } else {
                    // restore the exchange buffer to x for the next subproblem
                    __syncwarp(mask);
                    gik_publish<W, NPL>(c.P, x, lane);
                    __syncwarp(mask);
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
                    unsigned long long now = gik_globaltimer() - t0;
                    now = __shfl_sync(mask, now, leader, 32);
                    if (now >= a.maxtime_ns) { status = GIK_STATUS_MAXTIME; break; }
                }
                if (k_outer >= o.maxiter) { status = GIK_STATUS_MAXITER; break; }
                if (norm_grad < o.mingradnorm) { status = GIK_STATUS_CONVERGED; break; }
                if (may_park && inner_total - inner_entry >= a.inner_budget) {
                    if (lane == 0) park_slot = gik_try_park(a, n_res + a.B);
                    park_slot = __shfl_sync(mask, park_slot, leader, 32);
                    if (park_slot >= 0) { status = GIK_STATUS_PENDING; break; }
                    if (park_slot == -1) may_park = false;   // queue full: run this problem to its end
                }
            }
        }
        // ---- store optlog final_values (or, for a parked problem, its current ones and its queue entry)
        {
            double *dst = resumed ? reinterpret_cast<double *>(entp[CW_Y]) : a.Y_out + (size_t)b * 3 * a.N;
            double *cx = status == GIK_STATUS_PENDING ? gik_carry_slot(a.carry_out, park_slot) : nullptr;
#pragma unroll
            for (int m = 0; m < NPL; ++m) {
                const int i = lane + W * m;
                if (i < a.N) {
                    dst[3 * i] = x[m][0]; dst[3 * i + 1] = x[m][1]; dst[3 * i + 2] = x[m][2];
                    if (cx) {
                        double *cd = cx + CW_X + 3 * i;
                        cd[0] = x[m][0]; cd[1] = x[m][1]; cd[2] = x[m][2];
                        cd += 3 * a.N;
                        cd[0] = g[m][0]; cd[1] = g[m][1]; cd[2] = g[m][2];
                    }
                }
            }
            if (lane == 0) {
                const double sg[3] = {0.0, 0.0, 0.0};   // only k_rtr_fast / k_rtr_fast2 carry sum g_i x Y_i
                gik_finish_problem(a, resumed, b, entp, goal_row, cx, t0, status, k_outer, inner_total, fx, gg,
                                   norm_grad, Delta, Mi, sg, dst);
            }
        }
        __syncwarp(mask);
    }
}

template <int W, int NPL>
int launch_rtr(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    constexpr int NP = W * NPL, GPB = kThreads / W;
    const size_t tbl = (size_t)p->maxdeg * p->N;
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t group_bytes = (size_t)GPB * (6 * NP + goal_pad) * sizeof(double);
    const size_t tbl_bytes = tbl * (sizeof(double) + sizeof(uint32_t));
    a.tables_in_smem = tbl_bytes <= 64 * 1024;
    size_t smem = group_bytes + (NP + 2) * sizeof(int32_t) + (a.tables_in_smem ? tbl_bytes + 8 : 0);
    smem = (smem + 15) & ~(size_t)15;
    if (smem > 227 * 1024) {
        gik_set_error("gik_rtr_solve: needs %zu bytes of shared memory per CTA", smem);
        return GIK_ELIMIT;
    }
    GIK_CUDA(cudaFuncSetAttribute(k_rtr<W, NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rtr<W, NPL>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = p->sm_count * per_sm;
    const int need = (a.B + GPB - 1) / GPB;
    if (!a.carry_in && blocks > need) blocks = need;   // the number of parked problems is only known on the device
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    k_rtr<W, NPL><<<blocks, kThreads, smem, st>>>(a);
    return gik_check_cuda(cudaGetLastError(), "k_rtr launch");
}

}  // namespace

// Kernel selection shared by gik_rtr_solve and gik_rtr_solve_sliced.  k_rtr_fast and k_rtr_duo return the same bits
// (gik_tr_math.cuh), so for N <= 16 the choice between them is free per launch: two problems per warp where the
// device is kept full (large one-piece batches, the bulk launches of a sliced stream), one problem per warp where
// the wall time is the slowest problem (small batches, draining launches).
static int rtr_dispatch(const GikPlan *p, RtrArgs &a, bool sliced, cudaStream_t st)
{
    int kernel = a.o.kernel;
    if (kernel == GIK_KERNEL_AUTO) {
        if (p->N <= 32) {
            const bool bulk = sliced ? (a.B > 0 && a.carry_out != nullptr) : a.B > 32768;
            kernel = (bulk && p->duo_info) ? GIK_KERNEL_THROUGHPUT : GIK_KERNEL_LATENCY;
        } else {
            kernel = (p->dense_target && 8 * (long)p->n_terms >= (long)p->N * p->N) ? GIK_KERNEL_DENSE
                     : (p->fast2_info ? GIK_KERNEL_LATENCY : GIK_KERNEL_GENERIC);
        }
    }
    if (kernel == GIK_KERNEL_DENSE) {
        const int rc = gik_launch_rtr_cta(p, a, st);
        if (rc <= 0) return rc;
        kernel = GIK_KERNEL_GENERIC;
    }
    if (kernel == GIK_KERNEL_THROUGHPUT) {
        const int rc = gik_launch_rtr_duo(p, a, st);
        if (rc <= 0) return rc;
        kernel = GIK_KERNEL_LATENCY;   // no lock-step specialisation for this plan
    }
    if (kernel == GIK_KERNEL_LATENCY) {
        int rc = gik_launch_rtr_fast(p, a, st);
        if (rc <= 0) return rc;
        rc = gik_launch_rtr_fast2(p, a, st);   // 33 .. 64 nodes: two nodes per lane
        if (rc <= 0) return rc;
    }
    if (p->W == 16) return launch_rtr<16, 1>(p, a, st);
    switch (p->NPL) {
        case 1: return launch_rtr<32, 1>(p, a, st);
        case 2: return launch_rtr<32, 2>(p, a, st);
        case 4: return launch_rtr<32, 4>(p, a, st);
        case 8: return launch_rtr<32, 8>(p, a, st);    // 129 .. 256 nodes: the state no longer fits the register file; the
                                                         // compiler keeps part of it in (L1-cached) local memory
        default: return launch_rtr<32, 15>(p, a, st);  // .. 480 nodes
    }
}

static int rtr_fill_args(const char *fn, const GikPlan *p, const double *goal_d2, const double *Y_init, int32_t B,
                         const GikSolveOpts *opts, double *Y_out, double *f, double *gradnorm, int32_t *iters,
                         int32_t *status, int32_t *n_inner, int32_t *work_counter, RtrArgs &a)
{
    if (!p || !work_counter || B < 0 ||
        (B > 0 && (!Y_init || !Y_out || !f || !gradnorm || !iters || !status || (p->n_goal > 0 && !goal_d2)))) {
        gik_set_error("%s: bad argument", fn);
        return GIK_EINVAL;
    }
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != p->device) {
        gik_set_error("%s: plan belongs to device %d but device %d is current", fn, p->device, dev);
        return GIK_EINVAL;
    }
    a.slot_info = p->slot_info;
    a.slot_target = p->slot_target;
    a.deg = p->deg;
    a.N = p->N;
    a.n_goal = p->n_goal;
    a.maxdeg = p->maxdeg;
    a.tables_in_smem = 0;
    a.goal_d2 = goal_d2;
    a.Y_init = Y_init;
    a.B = B;
    if (opts) a.o = *opts; else gik_default_opts(&a.o);
    if (a.o.maxiter < 1 || a.o.maxinner < 1) { gik_set_error("%s: maxiter/maxinner must be >= 1", fn); return GIK_EINVAL; }
    a.Y_out = Y_out; a.f = f; a.gradnorm = gradnorm; a.iters = iters; a.status = status; a.n_inner = n_inner;
    a.trace = nullptr; a.trace_rows = 0; a.work_counter = work_counter;
    a.inner_budget = 0; a.carry_in = nullptr; a.carry_out = nullptr; a.pending = nullptr;
    a.maxtime_ns = a.o.maxtime > 0.0 ? (unsigned long long)(a.o.maxtime * 1e9) : 0ull;
    return GIK_OK;
}

extern "C" int gik_rtr_solve(const GikPlan *p, const double *goal_d2, const double *Y_init, int32_t B,
                             const GikSolveOpts *opts, double *Y_out, double *f, double *gradnorm,
                             int32_t *iters, int32_t *status, int32_t *n_inner, double *trace,
                             int32_t trace_rows, int32_t *work_counter, void *stream)
{
    if (B == 0) return GIK_OK;
    RtrArgs a;
    if (int rc = rtr_fill_args("gik_rtr_solve", p, goal_d2, Y_init, B, opts, Y_out, f, gradnorm, iters, status,
                               n_inner, work_counter, a))
        return rc;
    a.trace = trace;
    a.trace_rows = trace ? trace_rows : 0;
    return rtr_dispatch(p, a, false, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// deferred stragglers

namespace {
__global__ void k_carry_init(GikCarryHdr *h, int capacity, int stride)
{
    h->count = 0; h->capacity = capacity; h->stride = stride; h->full = 0;
}
}  // namespace

extern "C" int64_t gik_carry_bytes(const GikPlan *p, int32_t capacity)
{
    if (!p || capacity < 0) return 0;
    return (int64_t)sizeof(GikCarryHdr) + (int64_t)capacity * gik_carry_stride(p->N) * (int64_t)sizeof(double);
}

extern "C" int gik_carry_init(const GikPlan *p, void *carry, int32_t capacity, void *stream)
{
    if (!p || !carry || capacity < 0) { gik_set_error("gik_carry_init: bad argument"); return GIK_EINVAL; }
    if ((uintptr_t)carry % 16) { gik_set_error("gik_carry_init: the buffer must be 16-byte aligned"); return GIK_EINVAL; }
    k_carry_init<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<GikCarryHdr *>(carry), capacity, gik_carry_stride(p->N));
    return gik_check_cuda(cudaGetLastError(), "k_carry_init launch");
}

extern "C" int gik_rtr_solve_sliced(const GikPlan *p, const double *goal_d2, const double *Y_init, int32_t B,
                                    const GikSolveOpts *opts, double *Y_out, double *f, double *gradnorm,
                                    int32_t *iters, int32_t *status, int32_t *n_inner, int32_t inner_budget,
                                    const void *carry_in, void *carry_out, int32_t *pending,
                                    int32_t *work_counter, void *stream)
{
    if (B == 0 && !carry_in) return GIK_OK;
    RtrArgs a;
    if (int rc = rtr_fill_args("gik_rtr_solve_sliced", p, goal_d2, Y_init, B, opts, Y_out, f, gradnorm, iters,
                               status, n_inner, work_counter, a))
        return rc;
    if (carry_in && carry_in == carry_out) {
        gik_set_error("gik_rtr_solve_sliced: carry_in and carry_out must be different buffers");
        return GIK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    a.carry_in = static_cast<const char *>(carry_in);
    if (carry_out && inner_budget > 0) {
        a.carry_out = static_cast<char *>(carry_out);
        a.inner_budget = inner_budget;
        a.pending = pending;
    }
    if (carry_out)   // empty the outgoing queue (capacity, stride and the `full` statistic stay)
        GIK_CUDA(cudaMemsetAsync(carry_out, 0, sizeof(int32_t), st));
    return rtr_dispatch(p, a, true, st);
}
