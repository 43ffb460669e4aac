// gik_tr_math.cuh -- the arithmetic of one tCG / trust-region iteration, written ONCE.
//
// k_rtr_fast (one problem per warp, throughput and latency variants) and k_rtr_duo (two problems per warp) must
// return bit-identical results for the same problem: a goal parked by one of them may be resumed by another
// (gik_rtr_solve_sliced), and a result must not depend on which kernel the batch size selected.  Every
// floating-point operation of the iteration therefore lives here with its rounding spelled out -- explicit
// fma / __dmul_rn / __dadd_rn, which the compiler neither contracts nor re-associates -- and the kernels only
// decide which lane evaluates what.  Formulas: trust_region.py:436-599 (tCG), :248-391 (outer step),
// costs.py:79-207 (cost, half gradient, Hessian-vector product), fixed_rank_psd_sym.py:91-113 (projection).
#pragma once
#include "gik_common.cuh"

namespace trm {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ double dot3(const double (&a)[3], const double (&b)[3])
{
    return fma(a[2], b[2], fma(a[1], b[1], mul(a[0], b[0])));
}

// out = a x b
__device__ __forceinline__ void cross(const double (&a)[3], const double (&b)[3], double *out)
{
    out[0] = fma(a[1], b[2], -mul(a[2], b[1]));
    out[1] = fma(a[2], b[0], -mul(a[0], b[2]));
    out[2] = fma(a[0], b[1], -mul(a[1], b[0]));
}

// ---- one slot (a directed term i -> j) of the cost / half-gradient pass at the point p (costs.py:125-169).
// d = p_i - p_j; T the squared target; returns the masked residual r = act (|d|^2 - T) and accumulates
// f += r^2, g += r d.  The caller caches (2 act d, 2 r) for the Hessian products at this point.
struct SlotEval {
    double rr;     // act * (|d|^2 - T)
    double two;    // 2 * act
};
__device__ __forceinline__ SlotEval slot_cost(double dx, double dy, double dz, double T, uint32_t kind,
                                              double &fpart, double (&gacc)[3])
{
    const double d = gik_sqdist(dx, dy, dz);
    double rr = sub(d, T);
    const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (rr < 0.0)) | ((kind == GIK_TERM_UP) & (rr > 0.0));
    rr = act ? rr : 0.0;
    fpart = fma(rr, rr, fpart);
    gacc[0] = fma(rr, dx, gacc[0]);
    gacc[1] = fma(rr, dy, gacc[1]);
    gacc[2] = fma(rr, dz, gacc[2]);
    SlotEval e;
    e.rr = rr;
    e.two = act ? 2.0 : 0.0;
    return e;
}

// ---- one slot of the Hessian-vector product (costs.py:171-207) from the cached c = 2 act d, c2 = 2 act r and
// w = delta_i - delta_j: z += c2 w, zb += <c, w> c  (two independent accumulator chains)
__device__ __forceinline__ void slot_hess(double cx, double cy, double cz, double c2, double wx, double wy, double wz,
                                          double (&z)[3], double (&zb)[3])
{
    const double t = fma(cx, wx, fma(cy, wy, mul(cz, wz)));
    z[0] = fma(c2, wx, z[0]);
    z[1] = fma(c2, wy, z[1]);
    z[2] = fma(c2, wz, z[2]);
    zb[0] = fma(t, cx, zb[0]);
    zb[1] = fma(t, cy, zb[1]);
    zb[2] = fma(t, cz, zb[2]);
}

// ---- per-node contributions to the inner products of the first reduction: <delta, Z>, c = sum Z_i x Y_i
__device__ __forceinline__ void hess_scalars(const double (&dl)[3], const double (&z)[3], const double (&x)[3],
                                             double (&v)[4])
{
    v[0] = dot3(dl, z);
    cross(z, x, v + 1);
}

// ---- everything between the two reductions of an inner iteration that is the same for every lane of a problem
struct InnerScalars {
    double om[3];     // omega = (tr(X) I - X)^-1 c
    double d_Hd;      // <delta, H delta>
    double alpha;     // z_r / d_Hd
    double e_Pe_new;  // |eta + alpha delta|^2 by its recurrence
    bool leave;       // negative curvature or trust-region boundary crossed (NaN-safe)
};
__device__ __forceinline__ InnerScalars inner_scalars(const double (&Mi)[6], const double (&v)[4], const double (&u)[3],
                                                      double z_r, double e_Pe, double e_Pd, double d_Pd, double Delta2)
{
    InnerScalars s;
    // <delta, H delta> = <delta, Z> - omega . u differs from <delta, Z> by rounding noise (u ~ eps): the reciprocal is
    // formed from <delta, Z> while omega is still being computed, the correction steps use the true denominator
    const double rcp = gik_rcp(v[0]);
    s.om[0] = fma(Mi[2], v[3], fma(Mi[1], v[2], mul(Mi[0], v[1])));
    s.om[1] = fma(Mi[4], v[3], fma(Mi[3], v[2], mul(Mi[1], v[1])));
    s.om[2] = fma(Mi[5], v[3], fma(Mi[4], v[2], mul(Mi[2], v[1])));
    s.d_Hd = sub(v[0], fma(s.om[2], u[2], fma(s.om[1], u[1], mul(s.om[0], u[0]))));
    s.alpha = gik_div_near(z_r, s.d_Hd, rcp);
    s.e_Pe_new = fma(mul(s.alpha, s.alpha), d_Pd, fma(mul(2.0, s.alpha), e_Pd, e_Pe));
    s.leave = !(s.d_Hd > 0.0) || !(s.e_Pe_new < Delta2);
    return s;
}

// H delta = Z - Y x omega (projection onto the horizontal space)
__device__ __forceinline__ void project(const double (&z)[3], const double (&x)[3], const double (&om)[3], double (&Hd)[3])
{
    double yxo[3];
    cross(x, om, yxo);
    Hd[0] = sub(z[0], yxo[0]);
    Hd[1] = sub(z[1], yxo[1]);
    Hd[2] = sub(z[2], yxo[2]);
}

// step to the trust-region boundary along delta (trust_region.py:515-523)
__device__ __forceinline__ double boundary_tau(double e_Pe, double e_Pd, double d_Pd, double Delta2)
{
    const double disc = fma(d_Pd, sub(Delta2, e_Pe), mul(e_Pd, e_Pd));
    return (sqrt(disc) - e_Pd) / d_Pd;
}

// candidate eta' = eta + alpha delta, H eta' = H eta + alpha H delta, r' = r + alpha H delta and this node's share of
// <eta', g>, <eta', H eta'>, <r', r'>
__device__ __forceinline__ void inner_step(double alpha, const double (&dl)[3], const double (&Hd)[3],
                                           const double (&eta)[3], const double (&Heta)[3], const double (&r)[3],
                                           const double (&g)[3], double (&ne)[3], double (&nh)[3], double (&nr)[3],
                                           double (&sdot)[4])
{
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        ne[q] = fma(alpha, dl[q], eta[q]);
        nh[q] = fma(alpha, Hd[q], Heta[q]);
        nr[q] = fma(alpha, Hd[q], r[q]);
    }
    sdot[0] = dot3(ne, g);
    sdot[1] = dot3(ne, nh);
    sdot[2] = dot3(nr, nr);
    sdot[3] = 0.0;
}

__device__ __forceinline__ double model_value(const double (&sd)[4]) { return fma(0.5, sd[1], sd[0]); }

// new search direction and its recurrences (trust_region.py:585-597); u = sum delta_i x Y_i follows
// u' = -sg + beta u because sum r_i x Y_i stays sum g_i x Y_i = sg while H delta is horizontal
struct NextDir {
    double beta, e_Pd, d_Pd;
};
__device__ __forceinline__ NextDir next_direction(double r_r_new, double z_r, double inv_z_r, double alpha, double e_Pd,
                                                  double d_Pd)
{
    NextDir n;
    n.beta = gik_div(r_r_new, z_r, inv_z_r);
    n.e_Pd = mul(n.beta, fma(alpha, d_Pd, e_Pd));
    n.d_Pd = fma(mul(n.beta, n.beta), d_Pd, r_r_new);
    return n;
}

// ---- per-node scalars an accepted iterate y with half gradient h needs: <h,h>, y^T y (6), sum h_i x y_i (3)
__device__ __forceinline__ void point_scalars(const double (&y)[3], const double (&h)[3], double *v)
{
    v[0] = dot3(h, h);
    v[1] = mul(y[0], y[0]); v[2] = mul(y[0], y[1]); v[3] = mul(y[0], y[2]);
    v[4] = mul(y[1], y[1]); v[5] = mul(y[1], y[2]); v[6] = mul(y[2], y[2]);
    cross(h, y, v + 7);
}

// ---- accept / reject of the outer iteration (trust_region.py:255-391)
struct OuterDecision {
    double Delta;       // radius for the next subproblem
    bool accept;
};
__device__ __forceinline__ OuterDecision outer_decision(double fx, double fx_prop, double g_eta, double eta_Heta,
                                                        double Delta, int stop, const GikSolveOpts &o)
{
    const double eps = 2.220446049250313e-16;  // np.spacing(1), trust_region.py:293
    const double rho_reg = mul(mul(fmax(1.0, fabs(fx)), eps), o.rho_regularization);
    const double rhonum = add(sub(fx, fx_prop), rho_reg);
    const double rhoden = add(sub(-g_eta, mul(0.5, eta_Heta)), rho_reg);
    const bool model_decreased = rhoden >= 0.0;
    const double rho = rhonum / rhoden;
    OuterDecision d;
    d.Delta = Delta;
    if (rho < 0.25 || !model_decreased || isnan(rho)) {
        d.Delta = Delta / 4.0;
    } else if (rho > 0.75 && (stop == 0 /* NEGATIVE_CURVATURE */ || stop == 1 /* EXCEEDED_TR */)) {
        d.Delta = fmin(mul(2.0, Delta), o.Delta_bar);
    }
    d.accept = model_decreased && rho > o.rho_prime;
    return d;
}

// ---- start of a subproblem (trust_region.py:436-490 with eta0 = 0 and the identity preconditioner)
struct TcgStart {
    double pw, r_target2, Delta2, inv_z_r;
};
__device__ __forceinline__ TcgStart tcg_start(double gg, double Delta, const GikSolveOpts &o)
{
    TcgStart t;
    const double norm_r0 = sqrt(gg);
    t.pw = o.theta == 1.0 ? norm_r0 : pow(norm_r0, o.theta);
    const double r_target = mul(norm_r0, fmin(t.pw, o.kappa));
    t.r_target2 = mul(r_target, r_target);   // ||r|| <= target  <=>  <r,r> <= target^2 (sqrt is monotone)
    t.Delta2 = mul(Delta, Delta);
    t.inv_z_r = gik_rcp(gg);                 // for beta = z_r_new / z_r, formed one iteration ahead
    return t;
}

}  // namespace trm
