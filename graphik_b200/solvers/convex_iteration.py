"""CIDGIK: rank-constrained semidefinite relaxation by convex iteration (reference solvers/convex_iteration.py,
solvers/sdp_snl.py), batched on the GPU.

The reference alternates (convex_iteration.py:160-276), starting from C = I,
    1. minimise <C, Z> over the SDP relaxation of the distance-geometry problem (cvxpy -> MOSEK, sdp_snl.py:874-967),
    2. the closed-form Fantope step C = U U^T (convex_iteration.py:43-53),
until the optimum of step 1 stops changing (:262-266), then reads the points off Z (sdp_snl.py:763-780) and recovers
the joint angles (graph.joint_variables).

What is a restatement of the reference and what is not
------------------------------------------------------
* The PROGRAM (which constraints, on which variables, with which right-hand sides) is the reference's: dense
  relaxation, anchors p0, q0, p_n, q_n, one equality per DIST edge, identity block, no inequalities
  (`distance_range_constraints` only looks at obstacle pairs, sdp_snl.py:383-385, and the reference graph never has a
  robot-obstacle edge).  It is pinned against the reference's own matrices
  (tests/golden/cidgik_constraints.npz, tests/test_cidgik_cpu.py).
* The loop, the Fantope step (`gik_fantope`), the stopping test and the extraction are the reference's.
* The SDP SOLVER is not: MOSEK is closed third-party code outside the reference tree.  `gik_sdp_solve`
  (csrc/gik_sdp.cu) is a primal-dual interior-point method of its own; PARITY WITH MOSEK IS UNPINNED.  What the tests
  pin instead: kernel == numpy statement of the same method, optimality certificates of every solve, and the
  solver-independent end result (the recovered joint angles reach the goal pose).

Coordinates
-----------
Nodes joined pairwise by DIST edges form a rigid body in any dimension, so affine dependencies inside such a body
(points of one joint axis, coincident points) hold for every feasible Z: the feasible set has no interior, which an
interior-point method without MOSEK's self-dual embedding cannot work with.  `CidgikPlan` finds those dependencies once
per robot (maximal cliques of the DIST graph, null space of their homogeneous coordinates), eliminates the dependent
nodes and writes every node as a row vector r_u over (kept nodes, homogeneous coordinates); Z = V Zr V^T with the rows
of V the r_u.  This is the same program on the face it lives on -- same feasible Z, same optimum.  Every distance
constraint becomes (r_u - r_v)^T Zr (r_u - r_v) = d^2, the form `gik_sdp_solve` takes.
"""
import ctypes
import time

import numpy as np

from graphik_b200 import _lib
from graphik_b200.utils.se3 import as_matrix4

FEASIBLE, INFEASIBLE, SOLVER_ERROR = "feasible", "infeasible", "solver_error"


def solve_fantope_closed_form_batch(G, d):
    """C[b] = projector onto the eigenvectors of the n - d smallest eigenvalues of G[b] (convex_iteration.py:43-53),
    for G[B, n, n] (array or CUDA tensor).  Returns (C, eigvals) as CUDA tensors; eigvals ascending like numpy.eigh."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.GikError("graphik_b200 needs a CUDA device (B200); there is no CPU fallback")
    Gt = torch.as_tensor(np.ascontiguousarray(G, dtype=np.float64)) if not isinstance(G, torch.Tensor) else G
    Gt = Gt.to(device="cuda", dtype=torch.float64).contiguous()
    if Gt.dim() == 2:
        Gt = Gt[None]
    B, n = Gt.shape[0], Gt.shape[-1]
    C = torch.empty_like(Gt)
    ev = torch.empty((B, n), dtype=torch.float64, device=Gt.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(Gt.device).cuda_stream)
    with torch.cuda.device(Gt.device):
        _lib.check(_lib.load().gik_fantope(n, int(d), ctypes.c_void_p(Gt.data_ptr()), B, ctypes.c_void_p(C.data_ptr()),
                                           ctypes.c_void_p(ev.data_ptr()), stream), "gik_fantope")
    return C, ev


def solve_fantope_closed_form(G, d):
    """Reference signature (convex_iteration.py:43-53): one matrix in, (U U^T, seconds) out."""
    t0 = time.perf_counter()
    C, _ = solve_fantope_closed_form_batch(np.asarray(G, dtype=float)[None], d)
    return C[0].cpu().numpy(), time.perf_counter() - t0


def make_sdp_opts(params=None):
    o = _lib.SdpOpts()
    _lib.check(_lib.load().gik_sdp_default_opts(ctypes.byref(o)), "gik_sdp_default_opts")
    for key, val in (params or {}).items():
        if key not in ("tol", "maxiter", "tau", "x0"):
            raise ValueError("unknown SDP option %r" % key)
        setattr(o, key, int(val) if key == "maxiter" else float(val))
    return o


def sdp_solve_batch(C, W, b, active=None, opts=None, tau=None):
    """`gik_sdp_solve` on CUDA tensors C[B,N,N], W[B,M,N], b[B,M] (float64, contiguous), tau[M] (0 equality, +1 / -1
    upper / lower bound; None: all equalities); returns a dict of tensors."""
    import torch
    B, M, N = W.shape
    dev = W.device
    out = {"X": torch.zeros((B, N, N), dtype=torch.float64, device=dev),
           "y": torch.zeros((B, M), dtype=torch.float64, device=dev),
           "obj": torch.zeros(B, dtype=torch.float64, device=dev),
           "resid": torch.full((B,), float("inf"), dtype=torch.float64, device=dev),
           "iters": torch.zeros(B, dtype=torch.int32, device=dev),
           "status": torch.full((B,), 3, dtype=torch.int32, device=dev)}
    _sdp_launch(C, W, b, active, opts, out, tau)
    return out


def _sdp_launch(C, W, b, active, opts, out, tau=None):
    import torch
    B, M, N = W.shape
    for t in (C, W, b):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError("gik_sdp_solve takes contiguous float64 CUDA tensors")
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    o = opts or make_sdp_opts()
    stream = ctypes.c_void_p(torch.cuda.current_stream(W.device).cuda_stream)
    with torch.cuda.device(W.device):
        _lib.check(_lib.load().gik_sdp_solve(N, M, p(C), p(W), p(b), p(tau), p(active), B, ctypes.byref(o), p(out["X"]),
                                             p(out["y"]), p(out["obj"]), p(out["resid"]), p(out["iters"]),
                                             p(out["status"]), stream), "gik_sdp_solve")


# ------------------------------------------------------------------------------------------------------------------
# per-robot structure (host, once)
# ------------------------------------------------------------------------------------------------------------------
def _maximal_cliques(adj):
    """Bron-Kerbosch with pivoting on a boolean adjacency matrix (a few dozen nodes)."""
    nbrs = [set(np.nonzero(row)[0].tolist()) for row in adj]
    cliques, stack = [], [(set(), set(range(len(nbrs))), set())]
    while stack:
        R, P, X = stack.pop()
        if not P and not X:
            cliques.append(sorted(R))
            continue
        pivot = max(P | X, key=lambda u: len(nbrs[u] & P))
        for v in sorted(P - nbrs[pivot]):
            stack.append((R | {v}, P & nbrs[v], X & nbrs[v]))
            P = P - {v}
            X = X | {v}
    return cliques


class CidgikPlan:
    """Static data of the relaxation for one ProblemGraphRevolute (anchors p0, q0, p_n, q_n as solve_with_cidgik sets
    them, convex_iteration.py:283-289)."""

    def __init__(self, graph, tol=1e-9):
        self.graph = graph
        robot = graph.robot
        n, d = robot.n, 3
        if graph.dim != 3:
            raise NotImplementedError("graphik_b200 supports dim = 3")
        ids = graph.node_ids
        # convex_iteration.py:179-180 removes x and y; obstacles (nodes with a position: anchors, :186-189) are joined to
        # the other anchors only and to robot nodes by BOUNDED edges only, so they never enter an equality
        self.obstacle_names = [o["name"] for o in graph.obstacles]
        self.names = [u for u in ids if u not in ("x", "y") and u not in self.obstacle_names]
        self.anchor_names = ["p0", "q0", "p%d" % n, "q%d" % n]
        self.free = [u for u in self.names if u not in self.anchor_names]  # canonical_point_order (:193)
        sel = [graph.idx(u) for u in self.names]
        dist = graph.dist[np.ix_(sel, sel)]
        nn = len(self.names)
        is_anchor = np.array([u in self.anchor_names for u in self.names])

        # -- affine dependencies inside rigid bodies, from one generic configuration
        rng = np.random.RandomState(20231)
        q_gen = rng.uniform(-2.0, 2.0, size=n)
        P = np.asarray(graph.realization_points(q_gen), dtype=float)[sel]
        adj = ~np.isnan(dist)
        np.fill_diagonal(adj, False)
        deps = []
        for K in _maximal_cliques(adj):
            if len(K) < 2:
                continue
            H = np.vstack([P[K].T, np.ones(len(K))])
            _, s, Vt = np.linalg.svd(H)
            rank = int(np.sum(s > tol * s[0]))
            for c in Vt[rank:]:
                row = np.zeros(nn)
                row[K] = c
                deps.append(row)
        R = np.array(deps).reshape(-1, nn)
        # reduced row echelon form with pivots on free nodes: pivot node = combination of the others
        pivots, row = [], 0
        cols = [j for j in range(nn) if not is_anchor[j]]
        while row < R.shape[0] and cols:
            sub = np.abs(R[row:][:, cols])
            i, jj = np.unravel_index(int(np.argmax(sub)), sub.shape)
            if sub[i, jj] < 1e-7:
                break
            j = cols.pop(jj)
            R[[row, row + i]] = R[[row + i, row]]
            R[row] /= R[row, j]
            for k in range(R.shape[0]):
                if k != row:
                    R[k] -= R[k, j] * R[row]
            pivots.append(j)
            row += 1
        R = R[:row]
        self.eliminated = [self.names[j] for j in pivots]
        self.kept = [u for u in self.free if u not in self.eliminated]
        nk = len(self.kept)
        kidx = {u: k for k, u in enumerate(self.kept)}
        aidx = {u: k for k, u in enumerate(self.anchor_names)}
        # every node as a row over (kept nodes | anchors): x_u = sum coefK[u, k] x_k + sum coefA[u, a] x_a
        coefK, coefA = np.zeros((nn, nk)), np.zeros((nn, 4))
        for j, u in enumerate(self.names):
            if u in kidx:
                coefK[j, kidx[u]] = 1.0
            elif u in aidx:
                coefA[j, aidx[u]] = 1.0
        for r, j in zip(R, pivots):
            for k in np.nonzero(np.abs(r) > 1e-13)[0]:
                if k == j:
                    continue
                w = self.names[k]
                if w in aidx:
                    coefA[j, aidx[w]] -= r[k]
                else:
                    coefK[j, kidx[w]] -= r[k]
        self.n_free, self.N = len(self.free), len(self.free) + d
        self.Nr = nk + d

        # -- constraints: one per DIST edge not between two anchors (sdp_snl.py:159-198) + the identity block
        WK, WA, WH, b, pairs = [], [], [], [], []
        for i in range(nn):
            for j in range(i + 1, nn):
                if np.isnan(dist[i, j]) or (is_anchor[i] and is_anchor[j]):
                    continue
                WK.append(coefK[i] - coefK[j])
                WA.append(coefA[i] - coefA[j])
                WH.append(np.zeros(d))
                b.append(dist[i, j] ** 2)
                pairs.append((self.names[i], self.names[j]))
        self.n_distance_constraints = len(b)
        self.pairs = pairs
        for p in range(d):                       # Z[-d:, -d:] = I as w^T Z w = b with w = e_p, e_p + e_q
            for q in range(p, d):
                w = np.zeros(d)
                w[p] += 1.0
                if q != p:
                    w[q] += 1.0
                WK.append(np.zeros(nk))
                WA.append(np.zeros(4))
                WH.append(w)
                b.append(1.0 if p == q else 2.0)
        WK, WA, WH, b = np.array(WK), np.array(WA), np.array(WH), np.array(b)
        # -- inequalities (distance_range_constraints, sdp_snl.py:356-398: BOUNDED edges with an obstacle at one end;
        # anchor_inequality_constraint :586-618).  The reference's own graphs never carry such an edge (SURVEY App. C.1:
        # obstacle_semantics="reference" -> none); obstacle_semantics="intended" has p_i -- obstacle with LOWER = radius.
        IK, IA, IH, ib, itau = [], [], [], [], []
        for oname in self.obstacle_names:
            o = graph.idx(oname)
            for j, u in enumerate(self.names):
                if is_anchor[j]:
                    continue                                  # both ends anchored: ignored (:364-366)
                iu = graph.idx(u)
                for flag, bound, sense in ((graph.below, graph.lower, -1.0), (graph.above, graph.upper, 1.0)):
                    if flag[iu, o]:
                        IK.append(coefK[j])
                        IA.append(coefA[j])
                        IH.append(-graph.pos[o])
                        ib.append(bound[iu, o] ** 2)
                        itau.append(sense)
        self.n_inequalities = len(ib)
        # an independent subset (the eliminated nodes make many of them repeat each other), chosen on one sample goal
        A0 = self._anchors_numpy(np.asarray(robot.fk_all(q_gen[None]))[0, n][None])[0]
        Wfull = np.hstack([WK, WA.dot(A0) + WH])
        rows = np.einsum("ki,kj->kij", Wfull, Wfull).reshape(len(b), -1)
        import scipy.linalg as sla
        _, Rq, piv = sla.qr(rows.T, pivoting=True, mode="economic")
        dg = np.abs(np.diag(Rq))
        keep = np.sort(piv[:int(np.sum(dg > tol * dg[0]))])
        self.keep = keep
        self.WK, self.WA, self.WH, self.b = WK[keep], WA[keep], WH[keep], b[keep]
        self.tau = np.zeros(len(keep))
        if self.n_inequalities:
            self.WK = np.vstack([self.WK, np.array(IK)])
            self.WA = np.vstack([self.WA, np.array(IA)])
            self.WH = np.vstack([self.WH, np.array(IH)])
            self.b = np.concatenate([self.b, np.array(ib)])
            self.tau = np.concatenate([self.tau, np.array(itau)])
        self.M = len(self.b)
        # -- V: rows of the free nodes (reference order: graph order), then the homogeneous coordinates
        fsel = [self.names.index(u) for u in self.free]
        self.VK = np.vstack([coefK[fsel], np.zeros((d, nk))])
        self.VA = np.vstack([coefA[fsel], np.zeros((d, 4))])
        self.VH = np.vstack([np.zeros((len(fsel), d)), np.eye(d)])
        self._dev = {}

    # anchors[B, 4, 3] = p0, q0, p_n, q_n (convex_iteration.py:284-289)
    def _anchors_numpy(self, T):
        g, n = self.graph, self.graph.robot.n
        T = np.asarray(T, dtype=float).reshape(-1, 4, 4)
        A = np.empty((T.shape[0], 4, 3))
        A[:, 0] = g.pos[g.idx("p0")]
        A[:, 1] = g.pos[g.idx("q0")]
        A[:, 2] = T[:, :3, 3]
        A[:, 3] = T[:, :3, 3] + T[:, :3, 2] * g.axis_length
        return A

    def _tensors(self, device):
        import torch
        key = str(device)
        if key not in self._dev:
            f = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)
            g = self.graph
            self._dev[key] = {k: f(getattr(self, k)) for k in ("WK", "WA", "WH", "b", "VK", "VA", "VH", "tau")}
            self._dev[key]["base"] = f(np.stack([g.pos[g.idx("p0")], g.pos[g.idx("q0")]]))
        return self._dev[key]

    def assemble(self, T_goals, device="cuda"):
        """anchors[B,4,3], W[B,M,Nr], b[B,M], V[B,N,Nr] as CUDA tensors for goals T_goals[B,4,4]."""
        import torch
        s = self._tensors(device)
        T = torch.as_tensor(T_goals, dtype=torch.float64, device=device).reshape(-1, 4, 4)
        B = T.shape[0]
        pn = T[:, :3, 3]
        anchors = torch.stack([s["base"][0].expand(B, 3), s["base"][1].expand(B, 3), pn,
                               pn + T[:, :3, 2] * self.graph.axis_length], dim=1)
        W = torch.cat([s["WK"].expand(B, -1, -1), torch.matmul(s["WA"], anchors) + s["WH"]], dim=2).contiguous()
        V = torch.cat([s["VK"].expand(B, -1, -1), torch.matmul(s["VA"], anchors) + s["VH"]], dim=2).contiguous()
        b = s["b"].expand(B, -1).contiguous()
        return anchors, W, b, V


def _plan_for(graph):
    sig = hash((graph.number_of_nodes(), graph.dist.tobytes(), graph.lower.tobytes(), graph.upper.tobytes(),
                graph.below.tobytes(), graph.above.tobytes()))
    cached = getattr(graph, "_gik_cidgik_cache", None)
    if cached is None or cached[0] != sig:
        cached = (sig, CidgikPlan(graph))
        graph._gik_cidgik_cache = cached
    return cached[1]


# ------------------------------------------------------------------------------------------------------------------
# the convex iteration, batched
# ------------------------------------------------------------------------------------------------------------------
def convex_iterate_batch(graph, T_goals, max_iters=10, abs_eig_sum_tol=1e-6, rel_eig_sum_tol=1e-3, W_init=None,
                         sdp_params=None, sdp_accept=1e-3, device="cuda", fused=True):
    """convex_iterate_sdp_snl_graph (convex_iteration.py:160-276; ranges=True, sparse=False, closed_form=True) for
    every goal of T_goals[B,4,4] at once.  Returns a dict of CUDA tensors: Z[B,N,N] (the last SDP solution of each
    goal), C, values[B,max_iters] (SDP optimum per convex iteration, NaN beyond the last), eig_sums (sum of the N - d
    smallest eigenvalues of Z, the reference's eig_value_sum_vs_iterations), n_iters[B], feasible[B] (0 feasible,
    1 infeasible, 2 solver error), resid[B], sdp_iters[B] (interior-point iterations, summed), anchors, plan.
    A program the solver leaves with a residual above `sdp_accept` counts as a solver error (the reference's
    SOLVER_ERROR branch, :241-244); below it an inaccurate answer is used as cvxpy's 'optimal_inaccurate' is.
    fused=True (default, C = I start only): the whole loop of a goal runs inside one kernel launch (`gik_cidgik_solve`);
    fused=False: one `gik_sdp_solve` + one `gik_fantope` launch per convex iteration with the bookkeeping in torch --
    the same arithmetic up to the rounding of the small matrix products."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.GikError("graphik_b200 needs a CUDA device (B200); there is no CPU fallback")
    plan = _plan_for(graph)
    d, N = 3, plan.N
    anchors, W, b, V = plan.assemble(T_goals, device)
    B = W.shape[0]
    dev = W.device
    tau = plan._tensors(device)["tau"] if plan.n_inequalities else None
    opts = make_sdp_opts(sdp_params)
    f64 = dict(dtype=torch.float64, device=dev)
    C = torch.eye(N, **f64).expand(B, N, N).contiguous() if W_init is None else \
        torch.as_tensor(W_init, **f64).expand(B, N, N).contiguous()
    sdp = {"X": torch.zeros((B, plan.Nr, plan.Nr), **f64), "y": torch.zeros((B, plan.M), **f64),
           "obj": torch.zeros(B, **f64), "resid": torch.full((B,), float("inf"), **f64),
           "iters": torch.zeros(B, dtype=torch.int32, device=dev),
           "status": torch.zeros(B, dtype=torch.int32, device=dev)}
    active = torch.ones(B, dtype=torch.int32, device=dev)
    feasible = torch.zeros(B, dtype=torch.int32, device=dev)
    last_cost = torch.full((B,), 1e6, **f64)
    values = torch.full((B, max_iters), float("nan"), **f64)
    eig_sums = torch.full((B, max_iters), float("nan"), **f64)
    n_iters = torch.zeros(B, dtype=torch.int32, device=dev)
    sdp_iters = torch.zeros(B, dtype=torch.int32, device=dev)
    # Z = V Zr V^T has its range in span(V): with V^T V = L L^T its non-zero eigenpairs are those of the Nr x Nr matrix
    # L^T Zr L (eigenvectors V L^-T u), so the Fantope step C = I - sum over the d largest of v v^T (:43-53) is taken
    # there and V^T C V = L (I - sum u u^T) L^T -- the N x N matrices are only formed once, at the end
    Vt = V.transpose(1, 2).contiguous()
    G = torch.matmul(Vt, V)
    Lc = torch.linalg.cholesky(G)
    Lct = Lc.transpose(1, 2)
    Nr = plan.Nr
    Cs = torch.eye(Nr, **f64).expand(B, Nr, Nr).contiguous()            # C = I
    Cr = G.contiguous() if W_init is None else torch.matmul(torch.matmul(Vt, C), V).contiguous()
    launches = 0
    sdp_iters_max = torch.zeros(max_iters, dtype=torch.int32, device=dev)   # per launch: its slowest program
    n_active = torch.zeros(max_iters, dtype=torch.int32, device=dev)
    fused = fused and W_init is None
    if fused:
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        Lcc, Cs = Lc.contiguous(), torch.empty((B, Nr, Nr), **f64)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().gik_cidgik_solve(
                Nr, plan.M, d, p(Cr), p(Lcc), p(W), p(b), p(tau), B, ctypes.byref(opts), int(max_iters),
                float(abs_eig_sum_tol), float(rel_eig_sum_tol), float(sdp_accept), p(sdp["X"]), p(sdp["y"]), p(Cs),
                p(values), p(eig_sums), p(n_iters), p(feasible), p(sdp["obj"]), p(sdp["resid"]), p(sdp_iters),
                p(sdp["status"]), stream), "gik_cidgik_solve")
        launches, it = 1, 1
    for it in range(0 if fused else max_iters):
        _sdp_launch(Cr, W, b, active, opts, sdp, tau)
        launches += 1
        on = active.bool()
        code = torch.where((sdp["status"] == 1) & (sdp["resid"] > sdp_accept), torch.full_like(sdp["status"], 3),
                           sdp["status"])
        bad = on & (code >= 2)                                # convex_iteration.py:237-245
        feasible = torch.where(bad, code - 1, feasible)
        on = on & ~bad
        Zs = torch.matmul(torch.matmul(Lct, sdp["X"]), Lc)
        Cs_new, ev = solve_fantope_closed_form_batch(Zs, d)   # :249-251
        launches += 1
        Cs = torch.where(on[:, None, None], Cs_new, Cs)
        Cr = torch.where(on[:, None, None], torch.matmul(torch.matmul(Lc, Cs_new), Lct), Cr).contiguous()
        values[:, it] = torch.where(on, sdp["obj"], values[:, it])
        eig_sums[:, it] = torch.where(on, ev[:, :Nr - d].sum(dim=1), eig_sums[:, it])   # the other N - Nr are zero
        n_iters += on.int()
        spent = torch.where(active.bool(), sdp["iters"], torch.zeros_like(sdp["iters"]))
        sdp_iters += spent
        sdp_iters_max[it] = spent.max()
        n_active[it] = active.sum()
        change = last_cost - sdp["obj"]                       # :262-266
        done = (change.abs() <= abs_eig_sum_tol) | (sdp["obj"] <= abs_eig_sum_tol) | \
               (change.abs() / last_cost.abs() < rel_eig_sum_tol)
        last_cost = torch.where(on & ~done, sdp["obj"], last_cost)
        active = (on & ~done).int()
        if int(active.sum()) == 0:
            break
    Z = torch.matmul(torch.matmul(V, sdp["X"]), Vt)           # the last program solved for each goal
    if W_init is None or it > 0:
        Qb = torch.linalg.solve_triangular(Lc, Vt, upper=False).transpose(1, 2)      # V L^-T, orthonormal columns
        eye = torch.eye(N, **f64)
        C = eye - torch.matmul(torch.matmul(Qb, torch.eye(Nr, **f64) - Cs), Qb.transpose(1, 2))
    return {"Z": Z, "C": C, "values": values, "eig_sums": eig_sums, "n_iters": n_iters, "feasible": feasible,
            "resid": sdp["resid"], "sdp_iters": sdp_iters, "anchors": anchors, "plan": plan, "launches": launches,
            "y": sdp["y"], "V": V, "W": W, "b": b, "tau": tau, "sdp_iters_max": sdp_iters_max, "n_active": n_active}


def solve_batch_with_cidgik(graph, T_goals, as_numpy=False, **kw):
    """solve_with_cidgik for T_goals[B,4,4]: dict with q[B,n] (joint angles, graph.joint_variables), x[B,Nnodes,3]
    (all node positions, graph order), feasible[B] (0 feasible, 1 infeasible, 2 solver error: the reference returns
    (None, None) for those), n_iters[B], values, eig_sums, resid."""
    import torch
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver
    out = convex_iterate_batch(graph, T_goals, **kw)
    plan, Z, anchors = out["plan"], out["Z"], out["anchors"]
    B, d = Z.shape[0], 3
    dev = Z.device
    X = Z[:, -d:, :plan.n_free]                                   # extract_solution (sdp_snl.py:763-780)
    Y = torch.zeros((B, graph.number_of_nodes(), 3), dtype=torch.float64, device=dev)
    for name in ["p0", "x", "y", "q0"] + plan.obstacle_names:     # convex_iteration.py:312-315 (+ the other anchors)
        Y[:, graph.idx(name)] = torch.as_tensor(graph.pos[graph.idx(name)], dtype=torch.float64, device=dev)
    Y[:, graph.idx(plan.anchor_names[2])] = anchors[:, 2]         # :308-310
    Y[:, graph.idx(plan.anchor_names[3])] = anchors[:, 3]
    for k, u in enumerate(plan.free):
        Y[:, graph.idx(u)] = X[:, :, k]
    T = torch.as_tensor(T_goals, dtype=torch.float64, device=dev).reshape(B, 4, 4).contiguous()
    q = RiemannianSolver(graph).engine.joints(Y.contiguous(), T)   # graph.joint_variables(G_sol, {p_n: T_goal}) (:317)
    res = {"q": q, "x": Y, "feasible": out["feasible"], "n_iters": out["n_iters"], "values": out["values"],
           "eig_sums": out["eig_sums"], "resid": out["resid"], "sdp_iters": out["sdp_iters"], "Z": Z,
           "launches": out["launches"] + 1, "sdp_iters_max": out["sdp_iters_max"], "n_active": out["n_active"]}
    if as_numpy:
        res = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items()}
    return res


def solve_with_cidgik(graph, T_goal):
    """convex_iteration.py:279-319: (q_sol dict, solution dict name -> position) or (None, None) when infeasible."""
    T = as_matrix4(T_goal)
    out = solve_batch_with_cidgik(graph, T[None], as_numpy=True)
    if int(out["feasible"][0]) != 0:
        return None, None
    q_sol = graph.robot.q_dict(out["q"][0])
    solution = {u: out["x"][0, graph.idx(u)] for u in graph.node_ids}
    return q_sol, solution
