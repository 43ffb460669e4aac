// gik_cg.cu -- RiemannianSolver(params={"solver": "ConjugateGradient"}) (reference riemannian_solver.py:52-60).
//
// The reference builds pymanopt.solvers.ConjugateGradient(mingradnorm=1e-9, maxiter=1e5, minstepsize=1e-10,
// orth_value=1e11, beta_type=BetaTypes[3] = HagerZhang) with pymanopt's default LineSearchAdaptive.  pymanopt 0.2.5
// is a third-party dependency that is NOT in the reference tree (setup.py:20 pins it), so this kernel restates the
// PUBLISHED algorithm of that version (pymanopt/solvers/conjugate_gradient.py and linesearch.py, themselves ports
// of Manopt's conjugategradient.m / linesearch_adaptive.m) on the reference's manifold
// (fixed_rank_psd_sym.py: retr(Y, U) = Y + U, transp(Y, Z, U) = proj(Z, U), egrad2rgrad = identity, Frobenius inner
// product, no preconditioner).  PARITY UNPINNED: neither the reference's tests nor its tree hold a vector for this
// branch; the kernel is checked against a numpy restatement of the same published algorithm (the oracle's
// Problem.solve_cg, tests/test_gpu_cg.py), and against the property that it reaches the cost the trust-region solver reaches.
//
// One W-lane group per problem (the layout of k_rtr, gik_rtr.cu); per iteration: the line search (cost + gradient
// passes at the trial points: the gradient of the accepted one is reused, pymanopt evaluates the cost there a second
// time), one reduction for X = newx^T newx and the two transports, one for the scalars of the Hager-Zhang rule.
#include "gik_rtr.cuh"

namespace {

constexpr int kThreads = 128;

struct CgArgs {
    RtrArgs r;          // tables, batch, outputs, trace, work counter (opts inside are unused)
    GikCgOpts o;
};

template <int NPL>
__device__ __forceinline__ double dot3(const double (&a)[NPL][3], const double (&b)[NPL][3])
{
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < NPL; ++m) s = fma(a[m][2], b[m][2], fma(a[m][1], b[m][1], fma(a[m][0], b[m][0], s)));
    return s;
}

// c += sum_i z_i x y_i (this lane's nodes)
template <int NPL>
__device__ __forceinline__ void cross_partial(const double (&z)[NPL][3], const double (&y)[NPL][3], double *c)
{
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
        c[0] += z[m][1] * y[m][2] - z[m][2] * y[m][1];
        c[1] += z[m][2] * y[m][0] - z[m][0] * y[m][2];
        c[2] += z[m][0] * y[m][1] - z[m][1] * y[m][0];
    }
}

// z_i -= y_i x omega: with omega = (tr(X) I - X)^-1 sum z_i x y_i this is PSDFixedRank.proj(Y, Z)
// (fixed_rank_psd_sym.py:91-113 in its 3 x 3 form, as in the trust-region kernels)
template <int NPL>
__device__ __forceinline__ void remove_vertical(double (&z)[NPL][3], const double (&y)[NPL][3], const double (&om)[3])
{
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
        z[m][0] -= y[m][1] * om[2] - y[m][2] * om[1];
        z[m][1] -= y[m][2] * om[0] - y[m][0] * om[2];
        z[m][2] -= y[m][0] * om[1] - y[m][1] * om[0];
    }
}

template <int W, int NPL>
__global__ void __launch_bounds__(kThreads) k_cg(const CgArgs ca)
{
    constexpr int NP = W * NPL;
    const RtrArgs &a = ca.r;
    const GikCgOpts &o = ca.o;
    extern __shared__ double smem[];
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

    const int gid = threadIdx.x / W, lane = threadIdx.x % W;
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
    const unsigned long long maxtime_ns = o.maxtime > 0.0 ? (unsigned long long)(o.maxtime * 1e9) : 0ull;

    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(a.work_counter, 1);
        b = __shfl_sync(mask, b, leader, 32);
        if (b >= a.B) break;

        double x[NPL][3], grad[NPL][3], desc[NPL][3], nx[NPL][3], ngrad[NPL][3];
        {
            const double *src = a.Y_init + (size_t)b * 3 * a.N;
#pragma unroll
            for (int m = 0; m < NPL; ++m) {
                const int i = lane + W * m;
                x[m][0] = x[m][1] = x[m][2] = 0.0;
                if (i < a.N) { x[m][0] = src[3 * i]; x[m][1] = src[3 * i + 1]; x[m][2] = src[3 * i + 2]; }
            }
            __syncwarp(mask);
            for (int k = lane; k < a.n_goal; k += W) goal[k] = a.goal_d2[(size_t)b * a.n_goal + k];
            gik_publish<W, NPL>(c.P, x, lane);
            __syncwarp(mask);
        }
        // cost = objective(x); grad = gradient(x); gradnorm; Pgrad = grad; gradPgrad = <grad, grad>
        double cost, gradPgrad;
        {
            double v[2];
            v[0] = gik_pass_cost_grad<W, NPL>(c, x, grad);
            v[1] = dot3<NPL>(grad, grad);
            gik_allreduce<W, 2>(v, mask);
            cost = v[0];
            gradPgrad = v[1];
        }
        double gradnorm = sqrt(gradPgrad);
#pragma unroll
        for (int m = 0; m < NPL; ++m)
#pragma unroll
            for (int q = 0; q < 3; ++q) desc[m][q] = -grad[m][q];      // initial descent direction
        int iter = 0, costevals = 1, status = GIK_STATUS_MAXITER;
        double stepsize = nan("");
        bool have_oldalpha = false;
        double oldalpha = 0.0;
        unsigned long long t0 = maxtime_ns ? gik_globaltimer() : 0ull;
        t0 = __shfl_sync(mask, t0, leader, 32);
        if (!(isfinite(cost) && isfinite(gradPgrad))) {
            status = GIK_STATUS_NAN;
        } else {
            for (;;) {
                // Solver._check_stopping_criterion(time0, gradnorm=gradnorm, iter=iter + 1, stepsize=stepsize)
                if (maxtime_ns) {
                    unsigned long long now = gik_globaltimer() - t0;
                    now = __shfl_sync(mask, now, leader, 32);
                    if (now >= maxtime_ns) { status = GIK_STATUS_MAXTIME; break; }
                }
                if (iter + 1 >= o.maxiter) { status = GIK_STATUS_MAXITER; break; }
                if (gradnorm < o.mingradnorm) { status = GIK_STATUS_CONVERGED; break; }
                if (stepsize < o.minstepsize) { status = GIK_STATUS_MINSTEP; break; }   // (NaN before the first step)

                // df0 = <grad, desc_dir>; no descent direction: restart from the negative gradient
                double v[2] = {dot3<NPL>(grad, desc), dot3<NPL>(desc, desc)};
                gik_allreduce<W, 2>(v, mask);
                double df0 = v[0], dd = v[1];
                bool restarted = false;
                if (df0 >= 0.0) {
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) desc[m][q] = -grad[m][q];
                    df0 = -gradPgrad;
                    dd = gradPgrad;
                    restarted = true;
                }
                // ---- LineSearchAdaptive.search(objective, man, x, desc_dir, cost, df0)
                const double norm_d = sqrt(dd);
                double alpha = have_oldalpha ? oldalpha : o.ls_initial_stepsize / norm_d;
                double newf = 0.0;
                int evals = 0;
                for (;;) {
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) nx[m][q] = fma(alpha, desc[m][q], x[m][q]);     // retr(x, alpha d)
                    __syncwarp(mask);
                    gik_publish<W, NPL>(c.P, nx, lane);
                    __syncwarp(mask);
                    double fv[1] = {gik_pass_cost_grad<W, NPL>(c, nx, ngrad)};
                    gik_allreduce<W, 1>(fv, mask);
                    newf = fv[0];
                    ++evals;
                    if (!(newf > cost + o.ls_suff_decr * alpha * df0 && evals <= o.ls_maxiter)) break;
                    alpha *= o.ls_contraction;
                }
                costevals += evals;
                if (newf > cost) {               // no decrease found: stay where we are
                    alpha = 0.0;
                    newf = cost;
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) { nx[m][q] = x[m][q]; ngrad[m][q] = grad[m][q]; }
                    __syncwarp(mask);
                    gik_publish<W, NPL>(c.P, nx, lane);
                    __syncwarp(mask);
                }
                stepsize = alpha * norm_d;
                oldalpha = evals == 2 ? alpha : 2.0 * alpha;     // keep pace after one backtrack, else speed up
                have_oldalpha = true;

                // ---- new cost-related quantities, transports to newx
                double s[13];
                s[0] = dot3<NPL>(ngrad, ngrad);
                s[1] = s[2] = s[3] = s[4] = s[5] = s[6] = 0.0;
#pragma unroll
                for (int m = 0; m < NPL; ++m) {
                    s[1] = fma(nx[m][0], nx[m][0], s[1]); s[2] = fma(nx[m][0], nx[m][1], s[2]);
                    s[3] = fma(nx[m][0], nx[m][2], s[3]); s[4] = fma(nx[m][1], nx[m][1], s[4]);
                    s[5] = fma(nx[m][1], nx[m][2], s[5]); s[6] = fma(nx[m][2], nx[m][2], s[6]);
                }
                s[7] = s[8] = s[9] = s[10] = s[11] = s[12] = 0.0;
                cross_partial<NPL>(grad, nx, s + 7);
                cross_partial<NPL>(desc, nx, s + 10);
                gik_allreduce<W, 13>(s, mask);
                const double newgradPnewgrad = s[0];
                const double newgradnorm = sqrt(newgradPnewgrad);
                double Mi[6], om_g[3], om_d[3];
                gik_sylvester_inverse(s + 1, Mi);
                gik_sym_mul(Mi, s + 7, om_g);
                gik_sym_mul(Mi, s + 10, om_d);
                // oldgrad = transp(x, newx, grad) (kept in `grad`), desc_dir = transp(x, newx, desc_dir)
                remove_vertical<NPL>(grad, nx, om_g);
                remove_vertical<NPL>(desc, nx, om_d);
                // scalars of Powell's restart test and of the Hager-Zhang rule (Pdiff = diff: no preconditioner)
                double h[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int m = 0; m < NPL; ++m)
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const double df = ngrad[m][q] - grad[m][q];           // diff = newgrad - oldgrad
                        h[0] = fma(grad[m][q], ngrad[m][q], h[0]);            // <oldgrad, Pnewgrad>
                        h[1] = fma(df, desc[m][q], h[1]);                     // deno
                        h[2] = fma(df, ngrad[m][q], h[2]);                    // numo
                        h[3] = fma(df, df, h[3]);                             // <diff, Pdiff>
                        h[4] = fma(desc[m][q], ngrad[m][q], h[4]);            // <desc_dir, newgrad>
                        h[5] = fma(desc[m][q], desc[m][q], h[5]);             // |desc_dir|^2
                    }
                gik_allreduce<W, 6>(h, mask);
                const double orth_grads = h[0] / newgradPnewgrad;
                double beta = 0.0;
                if (fabs(orth_grads) >= o.orth_value) {
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) desc[m][q] = -ngrad[m][q];
                } else {
                    if (o.beta_type == 0) {                       // FletcherReeves
                        beta = newgradPnewgrad / gradPgrad;
                    } else if (o.beta_type == 1) {                // PolakRibiere
                        beta = fmax(0.0, h[2] / gradPgrad);
                    } else if (o.beta_type == 2) {                // HestenesStiefel (0 / 0 -> 1 as pymanopt's except branch)
                        beta = h[1] == 0.0 ? 1.0 : fmax(0.0, h[2] / h[1]);
                    } else {                                      // HagerZhang
                        const double numo = h[2] - 2.0 * h[3] * h[4] / h[1];
                        beta = numo / h[1];
                        const double eta_HZ = -1.0 / (sqrt(h[5]) * fmin(0.01, newgradnorm));
                        beta = fmax(beta, eta_HZ);
                    }
#pragma unroll
                    for (int m = 0; m < NPL; ++m)
#pragma unroll
                        for (int q = 0; q < 3; ++q) desc[m][q] = fma(beta, desc[m][q], -ngrad[m][q]);
                }
                if (a.trace && iter < a.trace_rows && lane == 0) {
                    double *row = a.trace + ((size_t)b * a.trace_rows + iter) * 6;
                    row[0] = stepsize;
                    row[1] = (double)evals;
                    row[2] = beta;
                    row[3] = newf;
                    row[4] = restarted ? 1.0 : 0.0;
                    row[5] = newgradnorm;
                }
                // x = newx, cost = newcost, grad = newgrad, ...
#pragma unroll
                for (int m = 0; m < NPL; ++m)
#pragma unroll
                    for (int q = 0; q < 3; ++q) { x[m][q] = nx[m][q]; grad[m][q] = ngrad[m][q]; }
                cost = newf;
                gradnorm = newgradnorm;
                gradPgrad = newgradPnewgrad;
                ++iter;
                if (!isfinite(cost)) { status = GIK_STATUS_NAN; break; }
            }
        }
        {
            double *dst = a.Y_out + (size_t)b * 3 * a.N;
#pragma unroll
            for (int m = 0; m < NPL; ++m) {
                const int i = lane + W * m;
                if (i < a.N) { dst[3 * i] = x[m][0]; dst[3 * i + 1] = x[m][1]; dst[3 * i + 2] = x[m][2]; }
            }
            if (lane == 0) {
                a.f[b] = cost;
                a.gradnorm[b] = gradnorm;
                a.iters[b] = iter;
                a.status[b] = status;
                if (a.n_inner) a.n_inner[b] = costevals;
            }
        }
        __syncwarp(mask);
    }
}

template <int W, int NPL>
int launch_cg(const GikPlan *p, CgArgs &ca, cudaStream_t st)
{
    constexpr int NP = W * NPL, GPB = kThreads / W;
    RtrArgs &a = ca.r;
    const size_t tbl = (size_t)p->maxdeg * p->N;
    const int goal_pad = (p->n_goal + 1) & ~1;
    const size_t group_bytes = (size_t)GPB * (6 * NP + goal_pad) * sizeof(double);
    const size_t tbl_bytes = tbl * (sizeof(double) + sizeof(uint32_t));
    a.tables_in_smem = tbl_bytes <= 64 * 1024;
    size_t smem = group_bytes + (NP + 2) * sizeof(int32_t) + (a.tables_in_smem ? tbl_bytes + 8 : 0);
    smem = (smem + 15) & ~(size_t)15;
    if (smem > 227 * 1024) {
        gik_set_error("gik_cg_solve: needs %zu bytes of shared memory per CTA", smem);
        return GIK_ELIMIT;
    }
    // shared-memory opt-in and occupancy are properties of (kernel, smem size, device): looked up once
    static size_t cached_smem = ~(size_t)0;
    static int cached_per_sm = 0, cached_dev = -1;
    if (cached_smem != smem || cached_dev != p->device) {
        GIK_CUDA(cudaFuncSetAttribute(k_cg<W, NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int q = 0;
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_cg<W, NPL>, kThreads, smem));
        cached_per_sm = q < 1 ? 1 : q;
        cached_smem = smem;
        cached_dev = p->device;
    }
    const int per_sm = cached_per_sm;
    int blocks = p->sm_count * per_sm;
    const int need = (a.B + GPB - 1) / GPB;
    if (blocks > need) blocks = need;
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    k_cg<W, NPL><<<blocks, kThreads, smem, st>>>(ca);
    return gik_check_cuda(cudaGetLastError(), "k_cg launch");
}

}  // namespace

extern "C" int gik_cg_default_opts(GikCgOpts *o)
{
    if (!o) { gik_set_error("gik_cg_default_opts: null argument"); return GIK_EINVAL; }
    o->mingradnorm = 1e-9;          // riemannian_solver.py:54
    o->maxiter = 100000;            // :56 (10e4)
    o->minstepsize = 1e-10;         // :57
    o->orth_value = 1e11;           // :58 (10e10)
    o->beta_type = 3;               // :59 BetaTypes[3] = HagerZhang
    o->maxtime = 1000.0;            // pymanopt Solver default
    o->ls_contraction = 0.5;        // pymanopt LineSearchAdaptive defaults
    o->ls_suff_decr = 0.5;
    o->ls_maxiter = 10;
    o->ls_initial_stepsize = 1.0;
    return GIK_OK;
}

extern "C" int gik_cg_solve(const GikPlan *p, const double *goal_d2, const double *Y_init, int32_t B,
                            const GikCgOpts *opts, double *Y_out, double *f, double *gradnorm, int32_t *iters,
                            int32_t *status, int32_t *n_costevals, double *trace, int32_t trace_rows,
                            int32_t *work_counter, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !work_counter || B < 0 || !Y_init || !Y_out || !f || !gradnorm || !iters || !status ||
        (p->n_goal > 0 && !goal_d2)) {
        gik_set_error("gik_cg_solve: bad argument");
        return GIK_EINVAL;
    }
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != p->device) {
        gik_set_error("gik_cg_solve: plan belongs to device %d but device %d is current", p->device, dev);
        return GIK_EINVAL;
    }
    CgArgs ca;
    memset(&ca, 0, sizeof(ca));
    if (opts) ca.o = *opts; else gik_cg_default_opts(&ca.o);
    if (ca.o.maxiter < 1 || ca.o.ls_maxiter < 1 || ca.o.beta_type < 0 || ca.o.beta_type > 3) {
        gik_set_error("gik_cg_solve: maxiter / ls_maxiter must be >= 1, beta_type in 0..3");
        return GIK_EINVAL;
    }
    RtrArgs &a = ca.r;
    a.slot_info = p->slot_info;
    a.slot_target = p->slot_target;
    a.deg = p->deg;
    a.N = p->N;
    a.n_goal = p->n_goal;
    a.maxdeg = p->maxdeg;
    a.goal_d2 = goal_d2;
    a.Y_init = Y_init;
    a.B = B;
    a.Y_out = Y_out; a.f = f; a.gradnorm = gradnorm; a.iters = iters; a.status = status; a.n_inner = n_costevals;
    a.trace = trace;
    a.trace_rows = trace ? trace_rows : 0;
    a.work_counter = work_counter;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->W == 16) return launch_cg<16, 1>(p, ca, st);
    switch (p->NPL) {
        case 1: return launch_cg<32, 1>(p, ca, st);
        case 2: return launch_cg<32, 2>(p, ca, st);
        case 4: return launch_cg<32, 4>(p, ca, st);
        case 8: return launch_cg<32, 8>(p, ca, st);
        default: return launch_cg<32, 15>(p, ca, st);
    }
}
