/*
 * graphik_b200.h -- C ABI of libgraphik_b200.so
 *
 * Batched distance-geometric inverse kinematics on NVIDIA B200 (sm_100a).
 * This is the drop-in boundary for the Riemannian IK hot path of
 * utiasSTARS/GraphIK.  The reference is pure Python; its only native plug-in
 * for this path is the numba-AOT extension module `costgrd`
 * (graphik/solvers/costs.py:3-5,208-209, imported at
 * graphik/solvers/riemannian_solver.py:18-21).  Each entry point below names
 * the reference routine it replaces, evaluated for B goal poses per call.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative GIK_E* code on failure;
 *     gik_last_error() returns a thread-local message.  Nothing throws.
 *   - All array arguments except GikPlanDesc's are DEVICE pointers owned by the
 *     caller (e.g. torch tensors); the library allocates only inside
 *     gik_plan_create.  All floating-point data is IEEE double, row-major.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls
 *     are asynchronous on that stream.  A plan is immutable after creation, owns
 *     no scratch memory and may be shared between streams and threads (per-call
 *     scratch -- work_counter, workspace, carry queues -- is the caller's, one per
 *     concurrently running call); it belongs to the device that was current when
 *     it was created.
 *   - Point sets are Y[B][N][3]; node order is the ProblemGraph's node order
 *     (p0, x, y, q0, p1, q1, ..., pn, qn, obstacles...).
 */
#ifndef GRAPHIK_B200_H
#define GRAPHIK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIK_OK 0
#define GIK_EINVAL (-1)   /* bad argument */
#define GIK_ECUDA (-2)    /* CUDA runtime error (message has the details) */
#define GIK_ELIMIT (-3)   /* problem exceeds a compiled-in limit */

/* cost-term kinds (riemannian_solver.py:121-138 / costs.py:79-93) */
#define GIK_TERM_EQ 0     /* omega_ij > 0 : (D_ij - d_ij)^2            */
#define GIK_TERM_LO 1     /* psi_L,ij > 0 : max(psi_L,ij - d_ij, 0)^2  */
#define GIK_TERM_UP 2     /* psi_U,ij > 0 : max(d_ij - psi_U,ij, 0)^2  */

/* per-problem status written by gik_rtr_solve */
#define GIK_STATUS_CONVERGED 0   /* gradnorm < mingradnorm            */
#define GIK_STATUS_MAXITER 1     /* outer iterations reached maxiter  */
#define GIK_STATUS_NAN 2         /* non-finite cost at the start      */
#define GIK_STATUS_LIMITS 3      /* solved, but a distance limit is broken after joint recovery (set by
                                    gik_check_limits when asked to; riemannian_solver.py:230-232)     */
#define GIK_STATUS_PENDING 4     /* parked by gik_rtr_solve_sliced: final values arrive with a later launch */
#define GIK_STATUS_MAXTIME 5     /* wall time since the problem started reached maxtime               */
#define GIK_STATUS_MINSTEP 6     /* gik_cg_solve: the line search's step fell below minstepsize       */

typedef struct GikPlan GikPlan;

/*
 * Host-side description of one robot + environment (goal independent).  It is
 * what graph.from_pose / distance_matrix_from_graph / adjacency_matrix_from_graph
 * / distance_bound_matrices (graph_base.py:171-180,262-279; dgp.py:42-65) produce
 * per goal in the reference, factored into a static part and the few entries
 * that depend on the goal pose.
 */
typedef struct {
    int32_t n_nodes;            /* N */
    /* cost terms, one per (i<j, kind); order = the reference's `inds` order */
    int32_t n_terms;
    const int32_t *term_i;
    const int32_t *term_j;
    const int32_t *term_kind;   /* GIK_TERM_* */
    const double *term_target;  /* squared distance / squared limit; ignored if term_goal >= 0 */
    const int32_t *term_goal;   /* >= 0: target is goal_d2[b][term_goal]; -1: static */
    int32_t n_goal;             /* row length of goal_d2 (= 2 * n_anchor for pose goals) */
    /* goal assembly (_pose_goal, graph_revolute.py:243-249) */
    int32_t n_anchor;           /* nodes with a fixed known position */
    const int32_t *anchor_node; /* [n_anchor] */
    const double *anchor_pos;   /* [n_anchor][3] */
    int32_t goal_p, goal_q;     /* node indices of p_n and q_n */
    double axis_length;
    /* bound smoothing (dgp.py:192-231): static edge bounds, unsquared;
     * no edge: lower = 0, upper = +inf.  Goal edges (anchor x {p_n,q_n} without a
     * static distance) are listed separately and take sqrt(goal_d2) per goal. */
    const double *bs_lower;     /* [N][N] */
    const double *bs_upper;     /* [N][N] */
    int32_t n_goal_edges;
    const int32_t *goal_edge_i; /* [n_goal_edges] */
    const int32_t *goal_edge_j;
    const int32_t *goal_edge_slot; /* index into goal_d2 row */
    /* initialisation (riemannian_solver.py:67-75): adjacency incl. goal edges */
    const uint8_t *omega;       /* [N][N] 0/1, may be NULL if gik_init is not used */
    /* joint recovery (graph_revolute.py:251-318) */
    int32_t n_joints;           /* 0: tables absent */
    const double *T0;           /* [(n_joints+1)][4][4] zero-configuration frames */
    /* check_distance_limits (graph_base.py:219-260): the edges carrying a BELOW / ABOVE limit with
     * their unsquared bounds, independent of the goal (an edge that is also a goal edge keeps its
     * limit here although bound smoothing uses the goal's exact distance for it) */
    int32_t n_limits;
    const int32_t *limit_i;     /* [n_limits] */
    const int32_t *limit_j;
    const double *limit_lower;  /* [n_limits]; no lower limit: 0    */
    const double *limit_upper;  /* [n_limits]; no upper limit: +inf */
} GikPlanDesc;

/* Options of the trust-region solve; defaults = riemannian_solver.py:44-50 over
 * trust_region.py:85-122 and fixed_rank_psd_sym.py:72. */
typedef struct {
    double mingradnorm;         /* 5e-10 */
    int32_t maxiter;            /* 3000  */
    double theta;               /* 1.0   */
    double kappa;               /* 0.1   */
    double rho_prime;           /* 0.1   */
    double rho_regularization;  /* 1e3   */
    int32_t mininner;           /* 1     */
    int32_t maxinner;           /* 10000 */
    double Delta_bar;           /* 13 = typicaldist = 10 + k */
    double Delta0;              /* Delta_bar / 8 */
    int32_t kernel;             /* GIK_KERNEL_*: which implementation gik_rtr_solve launches (same results) */
    double maxtime;             /* 1000 s (pymanopt Solver default, tested first at the end of every outer
                                   iteration, then maxiter, then mingradnorm); <= 0: no time limit */
} GikSolveOpts;

/* Options of the conjugate-gradient solve (params["solver"] = "ConjugateGradient", riemannian_solver.py:52-60:
 * pymanopt 0.2.5's ConjugateGradient with its default LineSearchAdaptive). */
typedef struct {
    double mingradnorm;         /* 1e-9   */
    int32_t maxiter;            /* 100000 (the loop stops when iter + 1 >= maxiter, as pymanopt's does) */
    double minstepsize;         /* 1e-10  */
    double orth_value;          /* 1e11: Powell restart when |<oldgrad, newgrad>| / <newgrad, newgrad> reaches it */
    int32_t beta_type;          /* 0 FletcherReeves, 1 PolakRibiere, 2 HestenesStiefel, 3 HagerZhang (reference default) */
    double maxtime;             /* 1000 s; <= 0: no limit */
    double ls_contraction;      /* 0.5  LineSearchAdaptive(contraction_factor) */
    double ls_suff_decr;        /* 0.5  (suff_decr) */
    int32_t ls_maxiter;         /* 10   (maxiter: cost evaluations per search, + 1) */
    double ls_initial_stepsize; /* 1.0  */
} GikCgOpts;

/* gik_rtr_solve implementations (identical algorithm, different mapping to the SM) */
#define GIK_KERNEL_AUTO 0        /* N <= 16: two problems per warp for batches > 32768 and for the bulk launches of
                                    gik_rtr_solve_sliced, one warp per problem otherwise (same bits either way);
                                    N <= 32: one warp per problem;
                                    N > 32: dense kernel when >= 1/8 of the node pairs carry a term (and no pair
                                    carries two), else the two-nodes-per-lane warp kernel (N <= 64), else generic */
#define GIK_KERNEL_LATENCY 1     /* one warp per problem: register slot cache (N <= 32), two nodes per lane with a
                                    shared-memory slot cache (32 < N <= 64) */
#define GIK_KERNEL_THROUGHPUT 2  /* two problems per warp (N <= 16); bit-identical to GIK_KERNEL_LATENCY */
#define GIK_KERNEL_GENERIC 3     /* W-lane groups, any N <= 480 (beyond 128 nodes part of the per-lane state lives in
                                    local memory: a fallback, several times slower per node pair) */
#define GIK_KERNEL_DENSE 4       /* one CTA per problem, two CTAs per SM, packed symmetric pair cache in shared
                                    memory (32 < N <= 128) */

const char *gik_last_error(void);
int gik_version(void);

/* Fills *opts with the reference defaults. */
int gik_default_opts(GikSolveOpts *opts);

int gik_plan_create(const GikPlanDesc *desc, GikPlan **plan_out);
int gik_plan_destroy(GikPlan *plan);
/* what[0]=N, [1]=n_terms, [2]=n_goal, [3]=max node degree (slots), [4]=n_joints,
 * [5]=lanes per problem chosen for gik_rtr_solve, [6]=nodes per lane */
int gik_plan_info(const GikPlan *plan, int32_t what[8]);

/* _pose_goal + graph_complete_edges for the goal-dependent entries only
 * (graph_revolute.py:243-249, dgp.py:124-147):
 * T_goal[B][4][4] -> goal_d2[B][n_goal] squared distances
 * ([a] = |p_n - anchor_a|^2, [n_anchor + a] = |q_n - anchor_a|^2). */
int gik_goal_distances(const GikPlan *plan, const double *T_goal, int32_t B, double *goal_d2,
                       void *stream);

/* costgrd.lcost / lgrad / lcost_and_grad (costs.py:79-169); with a plan holding
 * only EQ terms: jcost / jgrad / jcost_and_grad (costs.py:7-77).
 * f[B] and/or g[B][N][3] may be NULL.  g is the reference's HALF gradient. */
int gik_cost_grad(const GikPlan *plan, const double *Y, const double *goal_d2, int32_t B,
                  double *f, double *g, void *stream);

/* costgrd.lhess / jhess (costs.py:171-207, 38-58): HW[B][N][3] = ehess(Y)[W]. */
int gik_hessvec(const GikPlan *plan, const double *Y, const double *W, const double *goal_d2,
                int32_t B, double *HW, void *stream);

/* PSDFixedRank.proj (fixed_rank_psd_sym.py:91-113), dim = 3:
 * out[B][N][3] = Z - Y * Omega,  Omega X + X Omega = Y^T Z - Z^T Y. */
int gik_proj(int32_t N, const double *Y, const double *Z, int32_t B, double *out, void *stream);

/* Per-call device workspace of gik_bounds / gik_init / gik_bounds_init in bytes: 0 when the three N x N matrices
 * of a goal stay in shared memory (small graphs; pass NULL), else the N x N matrices the resident CTAs keep in L2
 * (N > 20: the third matrix, which buys more CTAs per SM; N > 118: all three).  The plan owns no scratch memory:
 * calls that may run concurrently (different streams) need one workspace each. */
int64_t gik_workspace_bytes(const GikPlan *plan);

/* bound_smoothing (dgp.py:192-231) for B goals: lb, ub [B][N][N] (unsquared). */
int gik_bounds(const GikPlan *plan, const double *goal_d2, int32_t B, double *lb, double *ub,
               void *workspace, void *stream);

/* RiemannianSolver.generate_initialization (riemannian_solver.py:67-75; dgp.py:28-31,
 * 150-183) from bounds: lb, ub [B][N][N] -> Y_init[B][N][3]. */
int gik_init(const GikPlan *plan, const double *lb, const double *ub, int32_t B, double *Y_init,
             void *workspace, void *stream);

/* bound_smoothing + generate_initialization fused (bounds never leave the SM):
 * goal_d2[B][n_goal] -> Y_init[B][N][3]. */
int gik_bounds_init(const GikPlan *plan, const double *goal_d2, int32_t B, double *Y_init,
                    void *workspace, void *stream);

/* TrustRegions.solve + _truncated_conjugate_gradient on PSDFixedRank(N, 3)
 * (trust_region.py:112-599; fixed_rank_psd_sym.py; riemannian_solver.py:178-218)
 * for B problems in one persistent launch.
 *   Y_init[B][N][3]        starting points
 *   Y_out[B][N][3], f[B], gradnorm[B], iters[B], status[B]  = optlog final_values
 *   n_inner[B]  (may be NULL) Hessian-vector products spent per problem
 *   trace (may be NULL): [B][trace_rows][6] per outer iteration
 *       (Delta, numit, stop_reason, fx_prop, accepted, gradnorm)
 *   work_counter: device int32 scratch, zeroed by the call (persistent work queue). */
int gik_rtr_solve(const GikPlan *plan, const double *goal_d2, const double *Y_init, int32_t B,
                  const GikSolveOpts *opts, double *Y_out, double *f, double *gradnorm,
                  int32_t *iters, int32_t *status, int32_t *n_inner, double *trace,
                  int32_t trace_rows, int32_t *work_counter, void *stream);

/* RiemannianSolver(params={"solver": "ConjugateGradient"}).solver.solve (riemannian_solver.py:52-60, 206-209):
 * pymanopt 0.2.5 ConjugateGradient + LineSearchAdaptive restated from the published algorithm (third-party source,
 * not in the reference tree: parity unpinned, see gik_cg.cu).  Same arguments as gik_rtr_solve; n_costevals[B]
 * (may be NULL) = cost evaluations; trace[B][trace_rows][6] (may be NULL) per iteration: step size, cost evaluations
 * of the line search, beta, new cost, 1 if the direction was reset to the negative gradient, new gradient norm.
 * status: CONVERGED, MAXITER, NAN, MAXTIME or MINSTEP. */
int gik_cg_default_opts(GikCgOpts *opts);
int gik_cg_solve(const GikPlan *plan, const double *goal_d2, const double *Y_init, int32_t B,
                 const GikCgOpts *opts, double *Y, double *f, double *gradnorm, int32_t *iters,
                 int32_t *status, int32_t *n_costevals, double *trace, int32_t trace_rows,
                 int32_t *work_counter, void *stream);

/* ---- Deferred stragglers -------------------------------------------------------------------------
 * The number of tCG iterations a goal needs spans two orders of magnitude (UR10: median 5 k, 0.4 % of the
 * goals run into maxiter = 3000 with ~240 k), and gik_rtr_solve returns with its slowest problem.
 * gik_rtr_solve_sliced bounds a launch instead: once the launch has handed out its last new problem, every
 * problem that has spent at least `inner_budget` tCG iterations in it PARKS at its next outer-iteration
 * boundary of trust_region.py:179-422 -- its state goes to an entry of `carry_out` together with the
 * addresses of its outputs, its status reads GIK_STATUS_PENDING -- and the next call that receives that
 * buffer as `carry_in` resumes it before it starts any new problem, writing the final values to the
 * ORIGINAL output locations (which the caller must keep alive).  So a launch lasts as long as its work,
 * long problems run without interruption while there is new work around them, and a parked problem
 * follows bit for bit the trajectory of an unparked one.
 *   inner_budget <= 0 or carry_out == NULL: nothing parks (a call with carry_in only drains the queue).
 *   pending (may be NULL): device int32, += 1 for each of this call's problems that parks; the call that
 *       finishes a parked problem does -= 1 on the address that problem was parked with -- one counter
 *       per batch tells when the batch is complete.
 *   B == 0 with carry_in: only resumes.  The two buffers must differ; use them alternately.
 *   One queue must always be used with the same plan; GIK_KERNEL_LATENCY and GIK_KERNEL_THROUGHPUT may
 *   alternate on it (identical arithmetic), the other kernels may not be mixed with them.
 * gik_carry_bytes: size of a queue for `capacity` parked problems; gik_carry_init: format it once
 * (16-byte aligned device memory).  When the outgoing queue is full, problems simply run on. */
int64_t gik_carry_bytes(const GikPlan *plan, int32_t capacity);
int gik_carry_init(const GikPlan *plan, void *carry, int32_t capacity, void *stream);
int gik_rtr_solve_sliced(const GikPlan *plan, const double *goal_d2, const double *Y_init, int32_t B,
                         const GikSolveOpts *opts, double *Y_out, double *f, double *gradnorm,
                         int32_t *iters, int32_t *status, int32_t *n_inner, int32_t inner_budget,
                         const void *carry_in, void *carry_out, int32_t *pending,
                         int32_t *work_counter, void *stream);

/* ProblemGraphRevolute.joint_variables (graph_revolute.py:251-318):
 * Y[B][N][3] (+ T_goal[B][4][4] or NULL) -> q[B][n_joints]. */
int gik_joints(const GikPlan *plan, const double *Y, const double *T_goal, int32_t B, double *q,
               void *stream);

/* RobotRevolute.pose for every joint + graph.realization points
 * (robot_revolute.py:85-103, graph_base.py:112-121):
 * q[B][n_joints] -> T_ee[B][4][4] (may be NULL), Y[B][N][3] (may be NULL). */
int gik_fk(const GikPlan *plan, const double *q, int32_t B, double *T_ee, double *Y, void *stream);

/* ProblemGraph.check_distance_limits (graph_base.py:219-260) with the INTENDED semantics: for every
 * edge carrying a BELOW/ABOVE limit (GikPlanDesc.limit_*) count the realised distances outside
 * [LOWER - tol, UPPER + tol].  (As shipped, the reference's own node-type test never fires on revolute
 * graphs, so it reports no violation at all; callers wanting that behaviour simply skip this call.)
 * Y[B][N][3] -> n_broken[B]; status (may be NULL): status[b] = GIK_STATUS_LIMITS where n_broken[b] > 0 and
 * the solve itself had ended normally (converged or maxiter) -- the (None, None) of riemannian_solver.py:230-232. */
int gik_check_limits(const GikPlan *plan, const double *Y, double tol, int32_t B, int32_t *n_broken,
                     int32_t *status, void *stream);

/* CIDGIK, the closed-form Fantope step: solve_fantope_closed_form (solvers/convex_iteration.py:43-53),
 * C[b] = U U^T with U the eigenvectors of the n - d smallest eigenvalues of the symmetric G[b][n][n] (n <= 32);
 * eigvals[B][n] ascending (may be NULL). */
int gik_fantope(int32_t n, int32_t d, const double *G, int32_t B, double *C, double *eigvals, void *stream);

/* CIDGIK, the semidefinite programs of the convex iteration: solve_linear_cost_sdp (solvers/sdp_snl.py:874-967,
 * cvxpy -> MOSEK at :952).  MOSEK is closed third-party code outside the reference tree: PARITY UNPINNED; this is a
 * primal-dual interior-point method (HKM direction, Mehrotra predictor-corrector) for B programs of the form
 *     minimise <C, X>  s.t.  w_k^T X w_k = b_k  (tau_k = 0),  <= b_k  (tau_k = +1),  >= b_k  (tau_k = -1),  X >= 0 (N x N)
 * which is the form every constraint of the reference's program takes (a squared distance between two points, an
 * entry of the identity block) once the host has written the points in the coordinates of the face the feasible set
 * lives on (graphik_b200/solvers/convex_iteration.py); the inequalities are distance_range_constraints' bounds
 * (sdp_snl.py:356-398, 586-618), each with a slack in a 1 x 1 block of the cone.  C[B][N][N], W[B][M][N], b[B][M];
 * tau[M] shared by all programs (may be NULL: all equalities); active[B] (may be NULL):
 * programs with active[b] == 0 are skipped and their outputs left untouched.  X[B][N][N], y[B][M] (dual, may be
 * NULL), obj[B] = <C, X>, resid[B] = max(relative primal residual, relative dual residual, relative gap) of the
 * returned iterate, iters[B], status[B].  N <= 32, M <= 96 (GIK_ELIMIT beyond). */
#define GIK_SDP_OPTIMAL 0        /* resid < tol                                                             */
#define GIK_SDP_INACCURATE 1     /* stopped by maxiter, by a numerical breakdown of the Schur complement
                                    factorisation (cond ~ 1 / mu^2) or by three iterations without progress:
                                    the best iterate is returned, see resid                                 */
#define GIK_SDP_INFEASIBLE 2     /* dual improving ray found: the program has no feasible point
                                    (INFEASIBLE of convex_iteration.py:237-240)                             */
#define GIK_SDP_NUMERIC 3        /* non-finite data or iterate (SOLVER_ERROR of convex_iteration.py:241-244) */
typedef struct {
    double tol;                 /* 1e-7 (the reference asks MOSEK for 1e-6, sdp_formulations.py:10) */
    int32_t maxiter;            /* 50   */
    double tau;                 /* 0.95 fraction of the step to the boundary of the cone */
    double x0;                  /* 10   X = S = x0 I at the start */
} GikSdpOpts;
int gik_sdp_default_opts(GikSdpOpts *opts);
int gik_sdp_solve(int32_t N, int32_t M, const double *C, const double *W, const double *b,
                  const double *tau, const int32_t *active, int32_t B, const GikSdpOpts *opts, double *X, double *y,
                  double *obj, double *resid, int32_t *iters, int32_t *status, void *stream);

/* CIDGIK, the whole convex iteration of B goals in one launch: convex_iterate_sdp_snl_graph (solvers/
 * convex_iteration.py:160-276; dense, closed-form Fantope step) with the programs in the form gik_sdp_solve takes.
 * Per goal, in the warp that owns it: C = I, solve, Fantope step, stopping test on the change of the optimum
 * (:262-266), at most max_iters times.  The Fantope step is taken in the coordinates of the face: with Z = V X V^T and
 * V^T V = L L^T the non-zero eigenpairs of Z are those of L^T X L, so the caller passes G[B][N][N] = V^T V (= the cost
 * matrix of the first program, C = I) and Lc[B][N][N] = L (lower), and the cost of the next program is
 * L (I - sum over the d largest of u u^T) L^T.  W, b, tau, opts, X, y, obj, resid, status: as gik_sdp_solve (values of
 * the last program solved for each goal); Cs[B][N][N]: the last I - sum u u^T; values / eig_sums[B][max_iters]: optimum
 * and sum of the smallest eigenvalues per convex iteration (the caller pre-fills them, e.g. with NaN); n_iters[B];
 * feasible[B]: 0, 1 (INFEASIBLE, :237-240) or 2 (SOLVER_ERROR, :241-244: non-finite, or stopped with resid > sdp_accept);
 * sdp_iters[B]: interior-point iterations, summed. */
int gik_cidgik_solve(int32_t N, int32_t M, int32_t d, const double *G, const double *Lc, const double *W,
                     const double *b, const double *tau, int32_t B, const GikSdpOpts *opts, int32_t max_iters,
                     double abs_eig_sum_tol, double rel_eig_sum_tol, double sdp_accept, double *X, double *y,
                     double *Cs, double *values, double *eig_sums, int32_t *n_iters, int32_t *feasible, double *obj,
                     double *resid, int32_t *sdp_iters, int32_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHIK_B200_H */
