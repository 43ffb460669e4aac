/*
 * gik_oracle.c -- CPU restatement of GraphIK's Riemannian IK hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call it.  Nothing under graphik_b200/
 * imports or links it, and the product fails loudly without its CUDA library.
 *
 * Parity status: the reference's own tests hold NO golden vector for the cost /
 * gradient / Hessian kernels, proj, the trust-region loop or the
 * initialisation, so this restatement is pinned against
 * OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD CONTAINER
 * (oracle/gen_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py),
 * and bound_smoothing additionally against the containment property of
 * reference tests/test_bound_smoothing.py:99-117.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/graphik).  Plain C99, double precision, no FMA contraction
 * (build with -ffp-contract=off) so that edge loops reproduce the numba
 * build of costs.py operation for operation.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DIM 3

typedef struct {
    int N;               /* nodes */
    int E;               /* entries of inds */
    const int64_t *ii;   /* inds[0]: row indices (i < j), row-major order of np.nonzero */
    const int64_t *jj;   /* inds[1] */
    const double *D;     /* D_goal  [N*N], squared distances */
    const double *omega; /* [N*N] 0/1 */
    const double *psiL;  /* [N*N] squared lower limits (0 = none) */
    const double *psiU;  /* [N*N] squared upper limits (0 = none) */
} orc_problem;

/* ---------------------------------------------------------------- costs.py */

/* solvers/costs.py:79-93 (lcost) */
double orc_lcost(const orc_problem *p, const double *Y)
{
    double cost = 0;
    for (int e = 0; e < p->E; ++e) {
        int64_t i = p->ii[e], j = p->jj[e], ij = i * p->N + j;
        double nrm = 0;
        for (int k = 0; k < DIM; ++k) {
            double t = Y[i * DIM + k] - Y[j * DIM + k];
            nrm += t * t;
        }
        if (p->omega[ij] > 0) {
            double r = p->D[ij] - nrm;
            cost += r * r;
        }
        if (p->psiL[ij] > 0) {
            double r = fmax(p->psiL[ij] - nrm, 0);
            cost += r * r;
        }
        if (p->psiU[ij] > 0) {
            double r = fmax(-p->psiU[ij] + nrm, 0);
            cost += r * r;
        }
    }
    return cost;
}

/* solvers/costs.py:95-123 (lgrad): returns 2*sum, i.e. HALF the true gradient of lcost */
void orc_lgrad(const orc_problem *p, const double *Y, double *grad)
{
    memset(grad, 0, sizeof(double) * p->N * DIM);
    for (int e = 0; e < p->E; ++e) {
        int64_t i = p->ii[e], j = p->jj[e], ij = i * p->N + j;
        double nrm = 0;
        for (int k = 0; k < DIM; ++k) {
            double t = Y[i * DIM + k] - Y[j * DIM + k];
            nrm += t * t;
        }
        if (p->omega[ij] != 0) {
            for (int k = 0; k < DIM; ++k) {
                double a = (nrm - p->D[ij]) * (Y[i * DIM + k] - Y[j * DIM + k]);
                grad[i * DIM + k] += a;
                grad[j * DIM + k] += -a;
            }
        }
        if (p->psiL[ij] != 0 && fmax(p->psiL[ij] - nrm, 0) > 0) {
            for (int k = 0; k < DIM; ++k) {
                double a = (nrm - p->psiL[ij]) * (Y[i * DIM + k] - Y[j * DIM + k]);
                grad[i * DIM + k] += a;
                grad[j * DIM + k] += -a;
            }
        }
        if (p->psiU[ij] != 0 && fmax(-p->psiU[ij] + nrm, 0) > 0) {
            for (int k = 0; k < DIM; ++k) {
                double a = (nrm - p->psiU[ij]) * (Y[i * DIM + k] - Y[j * DIM + k]);
                grad[i * DIM + k] += a;
                grad[j * DIM + k] += -a;
            }
        }
    }
    for (int k = 0; k < p->N * DIM; ++k) grad[k] = 2 * grad[k];
}

/* solvers/costs.py:171-207 (lhess) */
void orc_lhess(const orc_problem *p, const double *Y, const double *w, double *hess)
{
    memset(hess, 0, sizeof(double) * p->N * DIM);
    for (int e = 0; e < p->E; ++e) {
        int64_t i = p->ii[e], j = p->jj[e], ij = i * p->N + j;
        double nrm = 0, sc = 0;
        for (int k = 0; k < DIM; ++k) {
            double t = Y[i * DIM + k] - Y[j * DIM + k];
            nrm += t * t;
            sc += t * (w[i * DIM + k] - w[j * DIM + k]);
        }
        const double tgt[3] = {p->D[ij], p->psiL[ij], p->psiU[ij]};
        const int on[3] = {
            p->omega[ij] != 0,
            p->psiL[ij] != 0 && fmax(p->psiL[ij] - nrm, 0) > 0,
            p->psiU[ij] != 0 && fmax(-p->psiU[ij] + nrm, 0) > 0,
        };
        for (int t = 0; t < 3; ++t) {
            if (!on[t]) continue;
            for (int k = 0; k < DIM; ++k) {
                double a = 2 * sc * (Y[i * DIM + k] - Y[j * DIM + k]);
                double b = (nrm - tgt[t]) * (w[i * DIM + k] - w[j * DIM + k]);
                double c = a + b;
                hess[i * DIM + k] += c;
                hess[j * DIM + k] += -c;
            }
        }
    }
    for (int k = 0; k < p->N * DIM; ++k) hess[k] = 2 * hess[k];
}

/* solvers/costs.py:7-16 (jcost): equality edges only, 0.5 * sum 2 r^2 */
double orc_jcost(const orc_problem *p, const double *Y)
{
    double cost = 0;
    for (int e = 0; e < p->E; ++e) {
        int64_t i = p->ii[e], j = p->jj[e], ij = i * p->N + j;
        double nrm = 0;
        for (int k = 0; k < DIM; ++k) {
            double t = Y[i * DIM + k] - Y[j * DIM + k];
            nrm += t * t;
        }
        double r = p->D[ij] - nrm;
        cost += 2 * (r * r);
    }
    return 0.5 * cost;
}

/* solvers/costs.py:19-35 (jgrad) */
void orc_jgrad(const orc_problem *p, const double *Y, double *grad)
{
    memset(grad, 0, sizeof(double) * p->N * DIM);
    for (int e = 0; e < p->E; ++e) {
        int64_t i = p->ii[e], j = p->jj[e], ij = i * p->N + j, ji = j * p->N + i;
        double nrm = 0;
        for (int k = 0; k < DIM; ++k) {
            double t = Y[i * DIM + k] - Y[j * DIM + k];
            nrm += t * t;
        }
        for (int k = 0; k < DIM; ++k) {
            grad[i * DIM + k] += -4 * (p->D[ij] - nrm) * (Y[i * DIM + k] - Y[j * DIM + k]);
            grad[j * DIM + k] += -4 * (p->D[ji] - nrm) * (Y[j * DIM + k] - Y[i * DIM + k]);
        }
    }
    for (int k = 0; k < p->N * DIM; ++k) grad[k] = 0.5 * grad[k];
}

/* solvers/costs.py:38-58 (jhess) */
void orc_jhess(const orc_problem *p, const double *Y, const double *w, double *hess)
{
    memset(hess, 0, sizeof(double) * p->N * DIM);
    for (int e = 0; e < p->E; ++e) {
        int64_t i = p->ii[e], j = p->jj[e], ij = i * p->N + j, ji = j * p->N + i;
        double nrm = 0, sc = 0;
        for (int k = 0; k < DIM; ++k) {
            double t = Y[i * DIM + k] - Y[j * DIM + k];
            sc += t * (w[i * DIM + k] - w[j * DIM + k]);
            nrm += t * t;
        }
        for (int k = 0; k < DIM; ++k) {
            double yi = Y[i * DIM + k], yj = Y[j * DIM + k];
            double wi = w[i * DIM + k], wj = w[j * DIM + k];
            hess[i * DIM + k] += 4 * (2 * sc * (yi - yj) + (nrm - p->D[ij]) * (wi - wj));
            hess[j * DIM + k] += 4 * (2 * sc * (yj - yi) + (nrm - p->D[ji]) * (wj - wi));
        }
    }
    for (int k = 0; k < p->N * DIM; ++k) hess[k] = 0.5 * hess[k];
}

/* ------------------------------------------------- fixed_rank_psd_sym.py */

/* Gaussian elimination with partial pivoting, n <= 9 (stands for np.linalg.solve) */
static int solve_dense(int n, double *A, double *b)
{
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
        if (A[piv * n + c] == 0.0) return -1;
        if (piv != c) {
            for (int k = 0; k < n; ++k) {
                double t = A[c * n + k];
                A[c * n + k] = A[piv * n + k];
                A[piv * n + k] = t;
            }
            double t = b[c];
            b[c] = b[piv];
            b[piv] = t;
        }
        for (int r = c + 1; r < n; ++r) {
            double f = A[r * n + c] / A[c * n + c];
            if (f == 0.0) continue;
            for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
            b[r] -= f * b[c];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        double s = b[r];
        for (int k = r + 1; k < n; ++k) s -= A[r * n + k] * b[k];
        b[r] = s / A[r * n + r];
    }
    return 0;
}

/* utils/manifolds/fixed_rank_psd_sym.py:91-113 (proj, dim == 3):
 * solve Omega X + X Omega = Y^T Z - Z^T Y through the explicit 9x9 system,
 * return Z - Y Omega. */
int orc_proj(int N, const double *Y, const double *Z, double *out)
{
    double X[3][3] = {{0}}, C[9];
    double YtZ[3][3] = {{0}};
    for (int n = 0; n < N; ++n)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                X[a][b] += Y[n * 3 + a] * Y[n * 3 + b];
                YtZ[a][b] += Y[n * 3 + a] * Z[n * 3 + b];
            }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) C[a * 3 + b] = YtZ[a][b] - YtZ[b][a];
    /* rows exactly as written at fixed_rank_psd_sym.py:97-105 */
    double A[81] = {
        X[0][0] + X[0][0], X[0][1], X[0][2], X[1][0], 0, 0, X[2][0], 0, 0,
        X[1][0], X[1][1] + X[0][0], X[1][2], 0, X[1][0], 0, 0, X[2][0], 0,
        X[2][0], X[2][1], X[2][2] + X[0][0], 0, 0, X[1][0], 0, 0, X[2][0],
        X[0][1], 0, 0, X[0][0] + X[1][1], X[0][1], X[0][2], X[2][1], 0, 0,
        0, X[0][1], 0, X[1][0], X[1][1] + X[1][1], X[1][2], 0, X[2][1], 0,
        0, 0, X[0][1], X[2][0], X[2][1], X[2][2] + X[1][1], 0, 0, X[2][1],
        X[0][2], 0, 0, X[1][2], 0, 0, X[0][0] + X[2][2], X[0][1], X[0][2],
        0, X[0][2], 0, 0, X[1][2], 0, X[1][0], X[1][1] + X[2][2], X[1][2],
        0, 0, X[0][2], 0, 0, X[1][2], X[2][0], X[2][1], X[2][2] + X[2][2]};
    int rc = solve_dense(9, A, C);
    for (int n = 0; n < N; ++n)
        for (int b = 0; b < 3; ++b) {
            double s = 0;
            for (int a = 0; a < 3; ++a) s += Y[n * 3 + a] * C[a * 3 + b];
            out[n * 3 + b] = Z[n * 3 + b] - s;
        }
    return rc;
}

/* fixed_rank_psd_sym.py:75-79 (inner: Euclidean dot of the flattened matrices) */
static double inner(int n, const double *a, const double *b)
{
    double s = 0;
    for (int k = 0; k < n; ++k) s += a[k] * b[k];
    return s;
}

/* ------------------------------------------------------- trust_region.py */

enum { NEGATIVE_CURVATURE = 0, EXCEEDED_TR, REACHED_TARGET_LINEAR, REACHED_TARGET_SUPERLINEAR,
       MAX_INNER_ITER, MODEL_INCREASED };

typedef struct {
    double mingradnorm;        /* riemannian_solver.py:45  (5e-10) */
    int maxiter;               /* :47  (3000) */
    double theta, kappa;       /* :48-49 (1.0, 0.1) */
    double rho_prime;          /* trust_region.py:90 (0.1) */
    double rho_regularization; /* :92 (1e3) */
    int mininner, maxinner;    /* :116-118 (1, 10000) */
    double Delta_bar, Delta0;  /* :132-138 (typicaldist = 10 + k = 13; Delta_bar / 8) */
    int use_limits;            /* 1: lcost/lgrad/lhess, 0: jcost/jgrad/jhess */
} orc_params;

void orc_default_params(orc_params *q)
{
    q->mingradnorm = 0.5 * 1e-9;
    q->maxiter = 3000;
    q->theta = 1.0;
    q->kappa = 0.1;
    q->rho_prime = 0.1;
    q->rho_regularization = 1e3;
    q->mininner = 1;
    q->maxinner = 10000;
    q->Delta_bar = 13.0;
    q->Delta0 = 13.0 / 8;
    q->use_limits = 1;
}

typedef struct {
    const orc_problem *p;
    const orc_params *q;
    double *tmp;      /* N*3 scratch for the Euclidean Hessian */
    long n_hess;      /* Hessian-vector products so far */
} orc_ctx;

static double cost_fn(orc_ctx *c, const double *Y)
{
    return c->q->use_limits ? orc_lcost(c->p, Y) : orc_jcost(c->p, Y);
}
/* pymanopt Problem.grad + PSDFixedRank.egrad2rgrad (fixed_rank_psd_sym.py:123): identity */
static void grad_fn(orc_ctx *c, const double *Y, double *g)
{
    if (c->q->use_limits) orc_lgrad(c->p, Y, g); else orc_jgrad(c->p, Y, g);
}
/* pymanopt Problem.hess + ehess2rhess (fixed_rank_psd_sym.py:126): proj(Y, ehess).
 * (pymanopt also evaluates egrad(x) here and discards it; omitted -- no effect.) */
static void hess_fn(orc_ctx *c, const double *Y, const double *v, double *Hv)
{
    if (c->q->use_limits) orc_lhess(c->p, Y, v, c->tmp); else orc_jhess(c->p, Y, v, c->tmp);
    orc_proj(c->p->N, Y, c->tmp, Hv);
    c->n_hess++;
}

/* trust_region.py:436-599 (_truncated_conjugate_gradient), use_rand = False */
static int tcg(orc_ctx *c, const double *x, const double *fgradx, double Delta,
               double *eta, double *Heta, int *numit, double *work)
{
    const int n = c->p->N * DIM;
    const orc_params *q = c->q;
    double *r = work, *delta = work + n, *Hdelta = work + 2 * n, *new_eta = work + 3 * n,
           *new_Heta = work + 4 * n;
    memset(eta, 0, sizeof(double) * n);
    memset(Heta, 0, sizeof(double) * n);
    memcpy(r, fgradx, sizeof(double) * n);
    double e_Pe = 0;
    double r_r = inner(n, r, r);
    double norm_r = sqrt(r_r);
    double norm_r0 = norm_r;
    double z_r = r_r; /* z = precon(r) = r */
    double d_Pd = z_r;
    for (int k = 0; k < n; ++k) delta[k] = -r[k];
    double e_Pd = 0;
    double model_value = 0;
    int stop = MAX_INNER_ITER;
    int j = 0;
    for (j = 0; j < q->maxinner; ++j) {
        hess_fn(c, x, delta, Hdelta);
        double d_Hd = inner(n, delta, Hdelta);
        double alpha = z_r / d_Hd;
        double e_Pe_new = e_Pe + 2 * alpha * e_Pd + alpha * alpha * d_Pd;
        if (d_Hd <= 0 || e_Pe_new >= Delta * Delta) {
            double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
            for (int k = 0; k < n; ++k) {
                eta[k] = eta[k] + tau * delta[k];
                Heta[k] = Heta[k] + tau * Hdelta[k];
            }
            stop = d_Hd <= 0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
            break;
        }
        e_Pe = e_Pe_new;
        for (int k = 0; k < n; ++k) {
            new_eta[k] = eta[k] + alpha * delta[k];
            new_Heta[k] = Heta[k] + alpha * Hdelta[k];
        }
        double new_model_value = inner(n, new_eta, fgradx) + 0.5 * inner(n, new_eta, new_Heta);
        if (new_model_value >= model_value) {
            stop = MODEL_INCREASED;
            break;
        }
        memcpy(eta, new_eta, sizeof(double) * n);
        memcpy(Heta, new_Heta, sizeof(double) * n);
        model_value = new_model_value;
        for (int k = 0; k < n; ++k) r[k] = r[k] + alpha * Hdelta[k];
        r_r = inner(n, r, r);
        norm_r = sqrt(r_r);
        if (j >= q->mininner && norm_r <= norm_r0 * fmin(pow(norm_r0, q->theta), q->kappa)) {
            stop = q->kappa < pow(norm_r0, q->theta) ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
            break;
        }
        double zold_rold = z_r;
        z_r = r_r;
        double beta = z_r / zold_rold;
        for (int k = 0; k < n; ++k) delta[k] = -r[k] + beta * delta[k];
        e_Pd = beta * (e_Pd + alpha * d_Pd);
        d_Pd = z_r + beta * beta * d_Pd;
    }
    if (j == q->maxinner) j = q->maxinner - 1; /* python loop variable after exhaustion */
    *numit = j;
    return stop;
}

/* per-outer-iteration trace row: Delta, numit, stop, fx_prop, accepted, gradnorm(after) */
#define TRACE_COLS 6

/* trust_region.py:112-434 (TrustRegions.solve) driven as riemannian_solver.py:178-218 does.
 * Returns the number of outer iterations; status: 0 gradnorm reached, 1 maxiter. */
int orc_rtr_solve(const orc_problem *p, const orc_params *q, const double *Y_init, double *x,
                  double *f_out, double *gradnorm_out, int *status, long *n_hess,
                  double *trace, int trace_rows)
{
    const int n = p->N * DIM;
    double *buf = (double *)malloc(sizeof(double) * n * 10);
    double *fgradx = buf, *eta = buf + n, *Heta = buf + 2 * n, *x_prop = buf + 3 * n,
           *work = buf + 4 * n; /* 5n */
    orc_ctx c = {p, q, buf + 9 * n, 0};
    memcpy(x, Y_init, sizeof(double) * n);
    int k = 0;
    double fx = cost_fn(&c, x);
    grad_fn(&c, x, fgradx);
    double norm_grad = sqrt(inner(n, fgradx, fgradx));
    double Delta = q->Delta0;
    const double eps = 2.220446049250313e-16; /* np.spacing(1) */
    *status = 1;
    for (;;) {
        int numit;
        int stop_inner = tcg(&c, x, fgradx, Delta, eta, Heta, &numit, work);
        for (int t = 0; t < n; ++t) x_prop[t] = x[t] + eta[t]; /* retr, fixed_rank_psd_sym.py:137 */
        double fx_prop = cost_fn(&c, x_prop);
        double rhonum = fx - fx_prop;
        double rhoden = -inner(n, fgradx, eta) - 0.5 * inner(n, eta, Heta);
        double rho_reg = fmax(1, fabs(fx)) * eps * q->rho_regularization;
        rhonum = rhonum + rho_reg;
        rhoden = rhoden + rho_reg;
        int model_decreased = rhoden >= 0;
        double rho = rhonum / rhoden;
        double Delta_used = Delta;
        if (rho < 1.0 / 4 || !model_decreased || isnan(rho)) {
            Delta = Delta / 4;
        } else if (rho > 3.0 / 4 && (stop_inner == NEGATIVE_CURVATURE || stop_inner == EXCEEDED_TR)) {
            Delta = fmin(2 * Delta, q->Delta_bar);
        }
        int accepted = 0;
        if (model_decreased && rho > q->rho_prime) {
            accepted = 1;
            memcpy(x, x_prop, sizeof(double) * n);
            fx = fx_prop;
            grad_fn(&c, x, fgradx);
            norm_grad = sqrt(inner(n, fgradx, fgradx));
        }
        if (trace && k < trace_rows) {
            double *row = trace + (size_t)k * TRACE_COLS;
            row[0] = Delta_used; row[1] = numit; row[2] = stop_inner; row[3] = fx_prop;
            row[4] = accepted; row[5] = accepted ? norm_grad : NAN;
        }
        k = k + 1;
        /* pymanopt Solver._check_stopping_criterion order: (time,) iter, gradnorm */
        if (k >= q->maxiter) { *status = 1; break; }
        if (norm_grad < q->mingradnorm) { *status = 0; break; }
    }
    *f_out = fx;
    *gradnorm_out = norm_grad;
    if (n_hess) *n_hess = c.n_hess;
    free(buf);
    return k;
}

/* Batch driver used by the CPU baseline: B independent problems that share omega / psi
 * (goal-independent) and differ in D_goal and Y_init.  OpenMP over problems. */
void orc_rtr_solve_batch(int B, int N, int E, const int64_t *ii, const int64_t *jj,
                         const double *D /*[B,N,N]*/, const double *omega, const double *psiL,
                         const double *psiU, const orc_params *q, const double *Y_init /*[B,N,3]*/,
                         double *Y_out, double *f, double *gradnorm, int *iters, int *status,
                         long *n_hess, int threads)
{
    /* the thread count is set here, not through OMP_NUM_THREADS: libgomp reads the environment once when it is
     * loaded, and torchrun exports OMP_NUM_THREADS=1 to every rank */
    if (threads < 1) threads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int b = 0; b < B; ++b) {
        orc_problem p = {N, E, ii, jj, D + (size_t)b * N * N, omega, psiL, psiU};
        iters[b] = orc_rtr_solve(&p, q, Y_init + (size_t)b * N * DIM, Y_out + (size_t)b * N * DIM,
                                 f + b, gradnorm + b, status + b, n_hess ? n_hess + b : 0, 0, 0);
    }
}

/* ------------------------------------------------------------------ dgp.py */

/* utils/dgp.py:192-231 (bound_smoothing) on the literal 2N-node digraph H:
 * nodes u (0..N-1) and u' (N..2N-1); per edge {u,v}: u->u' 0, v->v' 0,
 * u->v' -LOWER, v->u' -LOWER, u<->v UPPER, u'<->v' UPPER; all-pairs shortest
 * paths (the reference calls networkx Bellman-Ford; Floyd-Warshall here);
 * upper[u,v] = dist(u,v); lower[u,v] = max(0, -dist(u,v')).
 * edge[N*N] symmetric 0/1; lower/upper[N*N] edge attributes (unsquared). */
void orc_bound_smoothing(int N, const unsigned char *edge, const double *lower, const double *upper,
                         double *lb, double *ub)
{
    const int M = 2 * N;
    double *d = (double *)malloc(sizeof(double) * M * M);
    for (int a = 0; a < M * M; ++a) d[a] = INFINITY;
    for (int a = 0; a < M; ++a) d[a * M + a] = 0;
    for (int u = 0; u < N; ++u)
        for (int v = 0; v < N; ++v) {
            if (!edge[u * N + v] || u == v) continue;
            double lo = lower[u * N + v], up = upper[u * N + v];
            d[u * M + (u + N)] = fmin(d[u * M + (u + N)], 0);
            d[v * M + (v + N)] = fmin(d[v * M + (v + N)], 0);
            d[u * M + (v + N)] = fmin(d[u * M + (v + N)], -lo);
            d[u * M + v] = fmin(d[u * M + v], up);
            d[(u + N) * M + (v + N)] = fmin(d[(u + N) * M + (v + N)], up);
        }
    for (int k = 0; k < M; ++k)
        for (int a = 0; a < M; ++a) {
            double dak = d[a * M + k];
            if (dak == INFINITY) continue;
            for (int b = 0; b < M; ++b) {
                double t = dak + d[k * M + b];
                if (t < d[a * M + b]) d[a * M + b] = t;
            }
        }
    for (int u = 0; u < N; ++u)
        for (int v = 0; v < N; ++v) {
            double s = d[u * M + (v + N)];
            lb[u * N + v] = s < 0 ? -s : 0;
            ub[u * N + v] = d[u * M + v];
        }
    free(d);
}
