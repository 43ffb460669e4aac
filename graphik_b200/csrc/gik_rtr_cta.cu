// gik_rtr_cta.cu -- trust-region solve for LARGE, DENSE graphs (32 < N <= 128): one CTA per problem.
//
// Same algorithm and arithmetic conventions as the warp kernels (reference trust_region.py:112-599,
// costs.py:79-207, fixed_rank_psd_sym.py:91-137).  With spherical obstacles the reference's graph has
// N = 118 nodes of which 104 are mutually fixed anchors, i.e. 5609 of 6903 node pairs carry an equality
// term (BASELINE configs[2]); a slot list of ~109 entries per node walked by one lane (k_rtr<32,4>: 436
// slots per lane and iteration, tables in global memory) leaves the machine idle.  Here
//
//   * per accepted iterate x the pair quantity c2[j][i] = 2 act (d_ij - T_ij) is cached as a DENSE matrix in
//     shared memory (N x 128 doubles + a byte "hinge active" per pair), rebuilt by the cost/gradient pass of
//     the proposal from the static targets in global memory (L2) -- the 2 * n_anchor goal-dependent targets
//     all involve p_n or q_n and live in two per-problem rows Tp, Tq;
//   * 256 threads: thread t owns node i = t % 128 and one half of the neighbour range j; in the edge pass
//     every lane of a warp looks at the SAME j, so the neighbour's coordinates are shared-memory
//     broadcasts and c2[j][i] is a conflict-free row read; no index loads at all;
//   * the two halves of a node exchange their partial sums through shared memory (fixed order, so both
//     hold identical bits); inner products are warp butterflies + an 8-entry shared-memory stage, summed
//     by every thread in the same order -> all scalars are block-uniform and every branch of tCG / RTR is
//     taken by the whole CTA.
//
// A pair carrying more than one term kind is not representable densely; gik_launch_rtr_cta then returns 1
// and the caller falls back to k_rtr.
#include "gik_rtr.cuh"

namespace {

#ifndef GIK_CTA_THREADS
#define GIK_CTA_THREADS 256
#endif
constexpr int kThreads = GIK_CTA_THREADS;   // 2 (or 4) threads per node, each owning a slice of the neighbour range
constexpr int kWarps = kThreads / 32;
constexpr int NPAD = 128;
constexpr int kParts = kThreads / NPAD;

template <int K>
__device__ __forceinline__ void block_allreduce(double (&v)[K], double *red, int warp, int lane)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(GIK_FULL_MASK, v[k], off, 32);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[warp * K + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = red[k];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) s += red[w * K + k];
        v[k] = s;
    }
    __syncthreads();
}

struct CtaTables {
    const double *target;       // [N][N] squared targets (static part)
    const unsigned char *kind;  // [N][N] GIK_TERM_* or 3 = no term
    const int32_t *goal_i, *goal_j, *goal_slot;
    int n_goal_edges;
    int gp, gq;                 // node indices of p_n, q_n (-1: no goal-dependent targets)
};

__global__ void __launch_bounds__(kThreads, 1) k_rtr_cta(const RtrArgs a, const CtaTables tb)
{
    extern __shared__ double smem[];
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int node = tid % NPAD, part = tid / NPAD;
    const bool valid = node < N;
    const bool owner = valid && part == 0;
    const int jchunk = (N + kParts - 1) / kParts;
    const int jlo = part * jchunk, jhi = min(N, jlo + jchunk);
    double *C2 = smem;                                  // [N][NPAD] 2 act (d - T) at the current iterate
    double *P = C2 + (size_t)N * NPAD;                  // [3][NPAD]
    double *V = P + 3 * NPAD;                           // [3][NPAD]
    double *Zx = V + 3 * NPAD;                          // [kParts][3][NPAD] partial sums of the slices
    double *red = Zx + kParts * 3 * NPAD;               // [kWarps][10]
    double *Tp = red + kWarps * 10;                     // [NPAD] targets of the pairs (., p_n) for this problem
    double *Tq = Tp + NPAD;                             // [NPAD] targets of the pairs (., q_n)
    double *goal = Tq + NPAD;                           // [n_goal]
    int *s_b = reinterpret_cast<int *>(goal + ((a.n_goal + 1) & ~1));
    unsigned char *Kd = reinterpret_cast<unsigned char *>(s_b + 2);  // [N][NPAD] static term kind
    unsigned char *Ak = Kd + (size_t)N * NPAD;                       // [N][NPAD] term active at the current iterate
    const int gp = tb.gp, gq = tb.gq;
    const GikSolveOpts &o = a.o;
    const double eps = 2.220446049250313e-16;

    // ---- static tables once per CTA (transposed so that a warp reads a row of consecutive i)
    for (int e = tid; e < N * NPAD; e += kThreads) {
        const int j = e / NPAD, i = e % NPAD;
        Kd[e] = i < N ? tb.kind[(size_t)i * N + j] : 3;
    }
    __syncthreads();

    // combine the partial node sums of the two halves in a fixed order
    auto combine3 = [&](double (&z)[3]) {
        Zx[(part * 3 + 0) * NPAD + node] = z[0];
        Zx[(part * 3 + 1) * NPAD + node] = z[1];
        Zx[(part * 3 + 2) * NPAD + node] = z[2];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double acc = Zx[q * NPAD + node];
#pragma unroll
            for (int pp = 1; pp < kParts; ++pp) acc += Zx[(pp * 3 + q) * NPAD + node];
            z[q] = acc;
        }
        __syncthreads();
    };

    // costs.py:125-169 at point p (published in P): this thread's cost share, full half-gradient of the node
    auto cost_grad = [&](const double (&p)[3], double (&gout)[3]) -> double {
        double fpart = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
        if (valid) {
#pragma unroll 4
            for (int j = jlo; j < jhi; ++j) {
                const unsigned kind = Kd[j * NPAD + node];
                const double dx = p[0] - P[j], dy = p[1] - P[NPAD + j], dz = p[2] - P[2 * NPAD + j];
                const double d = dx * dx + dy * dy + dz * dz;
                const double tgt = j == gp ? Tp[node] : (j == gq ? Tq[node] :
                                   (node == gp ? Tp[j] : (node == gq ? Tq[j] : tb.target[(size_t)j * N + node])));
                double rr = d - tgt;
                const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (rr < 0.0)) |
                                 ((kind == GIK_TERM_UP) & (rr > 0.0));
                rr = act ? rr : 0.0;
                fpart = fma(rr, rr, fpart);
                gx = fma(rr, dx, gx);
                gy = fma(rr, dy, gy);
                gz = fma(rr, dz, gz);
                C2[j * NPAD + node] = 2.0 * rr;
                Ak[j * NPAD + node] = act;
            }
        }
        gout[0] = 2.0 * gx; gout[1] = 2.0 * gy; gout[2] = 2.0 * gz;
        combine3(gout);
        return 0.5 * fpart;
    };

    // costs.py:171-207 at x (in P) along w (in V)
    auto hess = [&](const double (&xx)[3], const double (&w)[3], double (&z)[3]) {
        double zx = 0.0, zy = 0.0, zz = 0.0;
        if (valid) {
#pragma unroll 4
            for (int j = jlo; j < jhi; ++j) {
                const double c2 = C2[j * NPAD + node];
                const bool act = Ak[j * NPAD + node] != 0;
                const double dx = xx[0] - P[j], dy = xx[1] - P[NPAD + j], dz = xx[2] - P[2 * NPAD + j];
                const double wx = w[0] - V[j], wy = w[1] - V[NPAD + j], wz = w[2] - V[2 * NPAD + j];
                const double s = dx * wx + dy * wy + dz * wz;
                const double aa = act ? 4.0 * s : 0.0;
                zx = fma(aa, dx, fma(c2, wx, zx));
                zy = fma(aa, dy, fma(c2, wy, zy));
                zz = fma(aa, dz, fma(c2, wz, zz));
            }
        }
        z[0] = zx; z[1] = zy; z[2] = zz;
        combine3(z);
    };

    auto publish = [&](double *buf, const double (&v)[3]) {
        if (part == 0) { buf[node] = v[0]; buf[NPAD + node] = v[1]; buf[2 * NPAD + node] = v[2]; }
    };

    for (;;) {
        if (tid == 0) s_b[0] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int b = s_b[0];
        __syncthreads();
        if (b >= a.B) break;

        double x[3] = {0.0, 0.0, 0.0}, g[3], eta[3], Heta[3], r[3], dl[3], Hd[3];
        if (valid) {
            const double *src = a.Y_init + ((size_t)b * N + node) * 3;
            x[0] = src[0]; x[1] = src[1]; x[2] = src[2];
        }
        for (int k = tid; k < a.n_goal; k += kThreads) goal[k] = a.goal_d2[(size_t)b * a.n_goal + k];
        publish(P, x);
        __syncthreads();
        if (gp >= 0) {
            for (int i = tid; i < N; i += kThreads) {
                Tp[i] = tb.target[(size_t)gp * N + i];
                Tq[i] = tb.target[(size_t)gq * N + i];
            }
            __syncthreads();
            for (int e = tid; e < tb.n_goal_edges; e += kThreads) {
                const int i = tb.goal_i[e], j = tb.goal_j[e];
                const double t = goal[tb.goal_slot[e]];
                if (j == gp) Tp[i] = t; else if (i == gp) Tp[j] = t;
                if (j == gq) Tq[i] = t; else if (i == gq) Tq[j] = t;
            }
        }
        __syncthreads();

        double fx, gg, Mi[6];
        {
            double v[8];
            v[0] = cost_grad(x, g);
            const double w8 = owner ? 1.0 : 0.0;   // node quantities are counted once (by the first half)
            v[1] = w8 * (g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
            v[2] = w8 * x[0] * x[0]; v[3] = w8 * x[0] * x[1]; v[4] = w8 * x[0] * x[2];
            v[5] = w8 * x[1] * x[1]; v[6] = w8 * x[1] * x[2]; v[7] = w8 * x[2] * x[2];
            block_allreduce<8>(v, red, warp, lane);
            fx = v[0];
            gg = v[1];
            gik_sylvester_inverse(v + 2, Mi);
        }
        const double w1 = owner ? 1.0 : 0.0;
        double norm_grad = sqrt(gg);
        double Delta = o.Delta0;
        int k_outer = 0, inner_total = 0, status = GIK_STATUS_MAXITER;
        if (!(isfinite(fx) && isfinite(gg))) {
            status = GIK_STATUS_NAN;
        } else {
            for (;;) {
                // ================= tCG (trust_region.py:436-599), eta0 = 0, precon = identity
#pragma unroll
                for (int q = 0; q < 3; ++q) { eta[q] = 0.0; Heta[q] = 0.0; r[q] = g[q]; dl[q] = -g[q]; }
                double e_Pe = 0.0, r_r = gg;
                const double norm_r0 = sqrt(r_r);
                double z_r = r_r, d_Pd = r_r, e_Pd = 0.0, model_value = 0.0;
                const double pw = o.theta == 1.0 ? norm_r0 : pow(norm_r0, o.theta);
                const double r_target = norm_r0 * fmin(pw, o.kappa);
                const double r_target2 = r_target * r_target;
                const double Delta2 = Delta * Delta;
                int stop = MAX_INNER_ITER;
                int j = 0;
                for (j = 0; j < o.maxinner; ++j) {
                    publish(V, dl);
                    __syncthreads();
                    hess(x, dl, Hd);   // raw Z; projected below
                    double v[7];
                    v[0] = w1 * (dl[0] * Hd[0] + dl[1] * Hd[1] + dl[2] * Hd[2]);
                    v[1] = w1 * (Hd[1] * x[2] - Hd[2] * x[1]);      // c = sum Z_i x Y_i
                    v[2] = w1 * (Hd[2] * x[0] - Hd[0] * x[2]);
                    v[3] = w1 * (Hd[0] * x[1] - Hd[1] * x[0]);
                    v[4] = w1 * (dl[1] * x[2] - dl[2] * x[1]);      // u = sum delta_i x Y_i
                    v[5] = w1 * (dl[2] * x[0] - dl[0] * x[2]);
                    v[6] = w1 * (dl[0] * x[1] - dl[1] * x[0]);
                    block_allreduce<7>(v, red, warp, lane);
                    double om[3];
                    gik_sym_mul(Mi, v + 1, om);
                    Hd[0] -= x[1] * om[2] - x[2] * om[1];
                    Hd[1] -= x[2] * om[0] - x[0] * om[2];
                    Hd[2] -= x[0] * om[1] - x[1] * om[0];
                    const double d_Hd = v[0] - (om[0] * v[4] + om[1] * v[5] + om[2] * v[6]);
                    ++inner_total;
                    const double alpha = gik_div(z_r, d_Hd, gik_rcp(d_Hd));
                    const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
                    if (!(d_Hd > 0.0) || e_Pe_new >= Delta2) {   // also catches NaN
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
                    double ne[3], nh[3], nr[3], sdot[3] = {0.0, 0.0, 0.0};
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        ne[q] = fma(alpha, dl[q], eta[q]);
                        nh[q] = fma(alpha, Hd[q], Heta[q]);
                        nr[q] = fma(alpha, Hd[q], r[q]);
                        sdot[0] = fma(ne[q], g[q], sdot[0]);
                        sdot[1] = fma(ne[q], nh[q], sdot[1]);
                        sdot[2] = fma(nr[q], nr[q], sdot[2]);
                    }
                    sdot[0] *= w1; sdot[1] *= w1; sdot[2] *= w1;
                    block_allreduce<3>(sdot, red, warp, lane);
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
                    const double beta = r_r / z_r;
                    z_r = r_r;
#pragma unroll
                    for (int q = 0; q < 3; ++q) dl[q] = fma(beta, dl[q], -r[q]);
                    e_Pd = beta * (e_Pd + alpha * d_Pd);
                    d_Pd = z_r + beta * beta * d_Pd;
                }
                const int numit = j < o.maxinner ? j : o.maxinner - 1;

                // ================= proposal x + eta (trust_region.py:248-251)
#pragma unroll
                for (int q = 0; q < 3; ++q) dl[q] = x[q] + eta[q];
                __syncthreads();
                publish(P, dl);
                __syncthreads();
                double v[10];
                v[0] = cost_grad(dl, Hd);
                v[1] = w1 * (g[0] * eta[0] + g[1] * eta[1] + g[2] * eta[2]);
                v[2] = w1 * (eta[0] * Heta[0] + eta[1] * Heta[1] + eta[2] * Heta[2]);
                v[3] = w1 * (Hd[0] * Hd[0] + Hd[1] * Hd[1] + Hd[2] * Hd[2]);
                v[4] = w1 * dl[0] * dl[0]; v[5] = w1 * dl[0] * dl[1]; v[6] = w1 * dl[0] * dl[2];
                v[7] = w1 * dl[1] * dl[1]; v[8] = w1 * dl[1] * dl[2]; v[9] = w1 * dl[2] * dl[2];
                block_allreduce<10>(v, red, warp, lane);
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
                    // rejected: bring the exchange buffer and the pair cache back to x
                    __syncthreads();
                    publish(P, x);
                    __syncthreads();
                    double gtmp[3];
                    cost_grad(x, gtmp);
                }
                if (a.trace && k_outer < a.trace_rows && tid == 0) {
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
        if (owner) {
            double *dst = a.Y_out + ((size_t)b * N + node) * 3;
            dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
        }
        if (tid == 0) {
            a.f[b] = fx;
            a.gradnorm[b] = norm_grad;
            a.iters[b] = k_outer;
            a.status[b] = status;
            if (a.n_inner) a.n_inner[b] = inner_total;
        }
        __syncthreads();
    }
}

}  // namespace

int gik_launch_rtr_cta(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    if (!p->dense_target || p->N > NPAD) return 1;
    const int N = p->N;
    const int goal_pad = (p->n_goal + 1) & ~1;
    size_t smem = ((size_t)N * NPAD + (8 + 3 * kParts) * NPAD + kWarps * 10 + goal_pad) * sizeof(double) + 2 * sizeof(int) +
                  2 * (size_t)N * NPAD;
    smem = (smem + 15) & ~(size_t)15;
    if (smem > 227 * 1024) return 1;
    GIK_CUDA(cudaFuncSetAttribute(k_rtr_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rtr_cta, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = p->sm_count * per_sm;
    if (blocks > a.B) blocks = a.B;
    CtaTables tb;
    tb.target = p->dense_target;
    tb.kind = p->dense_kind;
    tb.goal_i = p->dense_goal_i;
    tb.goal_j = p->dense_goal_j;
    tb.goal_slot = p->dense_goal_slot;
    tb.n_goal_edges = p->n_dense_goal;
    tb.gp = p->n_dense_goal > 0 ? p->goal_p : -1;
    tb.gq = p->n_dense_goal > 0 ? p->goal_q : -1;
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    k_rtr_cta<<<blocks, kThreads, smem, st>>>(a, tb);
    return gik_check_cuda(cudaGetLastError(), "k_rtr_cta launch");
}
