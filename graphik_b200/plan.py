"""Plan compiler: ProblemGraph (or raw matrices) -> GikPlanDesc -> device plan.

In the reference every goal pose rebuilds, through networkx, the four N x N
matrices the solver consumes (`solve_with_riemannian`, riemannian_solver.py:
220-226: from_pose -> distance_matrix_from_graph / adjacency_matrix_from_graph /
distance_bound_matrices) and then the edge index list `inds`
(riemannian_solver.py:79,123-125).  omega, psi_L, psi_U and `inds` do not
depend on the goal; only the 2 * n_anchor squared distances between the anchors
(p0, x, y, q0, obstacles) and the end-effector points (p_n, q_n) do.  The plan
holds the static part once, as a term list in the reference's `inds` order, and
marks the goal-dependent targets by their slot in a per-goal row `goal_d2`.
"""
import ctypes

import numpy as np

from graphik_b200 import _lib

TERM_EQ, TERM_LO, TERM_UP = 0, 1, 2


def _terms_from_matrices(omega, psi_L, psi_U, D, use_limits, goal_slot=None, inds=None):
    """Term list in the order of the reference's `inds` (riemannian_solver.py:79,123-125)
    with the per-edge tests of costs.py:79-93 (omega > 0, psi_L > 0, psi_U > 0).  With
    use_limits=False every listed pair is an equality term (costs.py:7-16 has no test)."""
    if inds is None:
        if use_limits:
            diff = psi_L != psi_U
            inds = np.nonzero(np.triu(omega) + np.triu(diff * (psi_L > 0)) + np.triu(diff * (psi_U > 0)))
        else:
            inds = np.nonzero(np.triu(omega))
    ti, tj, tk, tt, tg = [], [], [], [], []
    for i, j in zip(*inds):
        if i == j:
            continue
        gs = -1 if goal_slot is None else int(goal_slot[i, j])
        if omega[i, j] > 0 or not use_limits:
            ti.append(i); tj.append(j); tk.append(TERM_EQ); tt.append(D[i, j]); tg.append(gs)
        if use_limits and psi_L[i, j] > 0:
            ti.append(i); tj.append(j); tk.append(TERM_LO); tt.append(psi_L[i, j]); tg.append(-1)
        if use_limits and psi_U[i, j] > 0:
            ti.append(i); tj.append(j); tk.append(TERM_UP); tt.append(psi_U[i, j]); tg.append(-1)
    return (np.asarray(ti, np.int32), np.asarray(tj, np.int32), np.asarray(tk, np.int32),
            np.asarray(tt, np.float64), np.asarray(tg, np.int32))


class Plan:
    """Owns a device-side GikPlan.  Create through `from_graph` or `from_matrices`."""

    def __init__(self, arrays):
        self._a = arrays  # keep host arrays alive while the descriptor points at them
        L = _lib.load()
        d = _lib.PlanDesc()

        def ptr(key):
            arr = arrays.get(key)
            if arr is None or arr.size == 0:
                return None
            return arr.ctypes.data_as(ctypes.c_void_p)

        d.n_nodes = int(arrays["n_nodes"])
        d.n_terms = len(arrays["term_i"])
        for key in ("term_i", "term_j", "term_kind", "term_target", "term_goal", "anchor_node", "anchor_pos",
                    "bs_lower", "bs_upper", "goal_edge_i", "goal_edge_j", "goal_edge_slot", "omega", "T0",
                    "limit_i", "limit_j", "limit_lower", "limit_upper"):
            setattr(d, key, ptr(key))
        d.n_limits = len(arrays["limit_i"]) if arrays.get("limit_i") is not None else 0
        d.n_goal = int(arrays.get("n_goal", 0))
        d.n_anchor = len(arrays["anchor_node"]) if arrays.get("anchor_node") is not None else 0
        d.goal_p = int(arrays.get("goal_p", -1))
        d.goal_q = int(arrays.get("goal_q", -1))
        d.axis_length = float(arrays.get("axis_length", 1.0))
        d.n_goal_edges = len(arrays["goal_edge_i"]) if arrays.get("goal_edge_i") is not None else 0
        d.n_joints = int(arrays.get("n_joints", 0))
        handle = ctypes.c_void_p()
        _lib.check(L.gik_plan_create(ctypes.byref(d), ctypes.byref(handle)), "gik_plan_create")
        self.handle = handle
        info = (ctypes.c_int32 * 8)()
        _lib.check(L.gik_plan_info(handle, ctypes.byref(info)), "gik_plan_info")
        (self.N, self.n_terms, self.n_goal, self.maxdeg, self.n_joints, self.lanes, self.nodes_per_lane,
         self.sm_count) = [int(v) for v in info]
        self.n_anchor = d.n_anchor

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                _lib.load().gik_plan_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ------------------------------------------------------------------
    @staticmethod
    def arrays_from_graph(graph, use_limits=True):
        """Host arrays of the descriptor for a ProblemGraphRevolute (pose goals on p_n, q_n)."""
        N = graph.number_of_nodes()
        n = graph.robot.n
        gp, gq = graph.idx("p%d" % n), graph.idx("q%d" % n)
        anchors = np.asarray([k for k in graph.anchor_indices if k not in (gp, gq)], dtype=np.int32)
        A = len(anchors)
        has_dist = ~np.isnan(graph.dist)
        omega = has_dist.astype(np.float64)
        D = np.where(has_dist, np.nan_to_num(graph.dist) ** 2, 0.0)
        goal_slot = np.full((N, N), -1, dtype=np.int64)
        ge_i, ge_j, ge_s = [], [], []
        # graph_complete_edges (dgp.py:124-147): anchors x {p_n, q_n} lacking a DIST get the exact distance
        for base, g in ((0, gp), (A, gq)):
            for a, node in enumerate(anchors):
                if not has_dist[node, g]:
                    omega[node, g] = omega[g, node] = 1.0
                    goal_slot[node, g] = goal_slot[g, node] = base + a
                    ge_i.append(min(node, g)); ge_j.append(max(node, g)); ge_s.append(base + a)
        psi_L, psi_U = graph.distance_bound_matrices()
        ti, tj, tk, tt, tg = _terms_from_matrices(omega, psi_L, psi_U, D, use_limits, goal_slot)
        bs_lower = np.where(np.isnan(graph.lower), 0.0, graph.lower)
        bs_upper = np.where(np.isnan(graph.upper), np.inf, graph.upper)
        for i, j in zip(ge_i, ge_j):
            bs_lower[i, j] = bs_lower[j, i] = 0.0
            bs_upper[i, j] = bs_upper[j, i] = np.inf
        np.fill_diagonal(bs_lower, 0.0)
        np.fill_diagonal(bs_upper, 0.0)
        # check_distance_limits (graph_base.py:219-260, intended semantics): every BELOW / ABOVE edge with the graph's
        # own bounds -- NOT the bound-smoothing tables above, where goal edges lose their static limits
        li, lj = np.nonzero(np.triu(graph.below | graph.above, 1))
        lim_lo = np.where(np.isnan(graph.lower[li, lj]), 0.0, graph.lower[li, lj])
        lim_up = np.where(np.isnan(graph.upper[li, lj]), np.inf, graph.upper[li, lj])
        return {
            "limit_i": li.astype(np.int32), "limit_j": lj.astype(np.int32),
            "limit_lower": np.ascontiguousarray(lim_lo, dtype=np.float64),
            "limit_upper": np.ascontiguousarray(lim_up, dtype=np.float64),
            "n_nodes": N, "term_i": ti, "term_j": tj, "term_kind": tk, "term_target": tt, "term_goal": tg,
            "n_goal": 2 * A, "anchor_node": anchors,
            "anchor_pos": np.ascontiguousarray(graph.pos[anchors], dtype=np.float64),
            "goal_p": gp, "goal_q": gq, "axis_length": float(graph.axis_length),
            "bs_lower": np.ascontiguousarray(bs_lower), "bs_upper": np.ascontiguousarray(bs_upper),
            "goal_edge_i": np.asarray(ge_i, np.int32), "goal_edge_j": np.asarray(ge_j, np.int32),
            "goal_edge_slot": np.asarray(ge_s, np.int32),
            "omega": np.ascontiguousarray(omega != 0, dtype=np.uint8),
            "n_joints": n, "T0": np.ascontiguousarray(graph.robot.T0, dtype=np.float64),
            # kept for host-side users (not part of the descriptor)
            "omega_f": omega, "psi_L": psi_L, "psi_U": psi_U, "D_static": D, "goal_slot": goal_slot,
        }

    @classmethod
    def from_graph(cls, graph, use_limits=True):
        return cls(cls.arrays_from_graph(graph, use_limits))

    @classmethod
    def from_matrices(cls, D_goal, omega, psi_L=None, psi_U=None, use_limits=True, inds=None):
        """Static plan for one explicit problem, as RiemannianSolver.solve receives it
        (riemannian_solver.py:178-195): no goal-dependent slots, no bound/joint tables.
        `inds` overrides the edge list (the costgrd functions take it as an argument)."""
        D_goal = np.asarray(D_goal, dtype=np.float64)
        N = D_goal.shape[0]
        omega = np.zeros((N, N)) if omega is None else np.asarray(omega, dtype=np.float64)
        psi_L = np.zeros((N, N)) if psi_L is None else np.asarray(psi_L, dtype=np.float64)
        psi_U = np.zeros((N, N)) if psi_U is None else np.asarray(psi_U, dtype=np.float64)
        if inds is not None:
            inds = (np.asarray(inds[0], dtype=np.int64), np.asarray(inds[1], dtype=np.int64))
        ti, tj, tk, tt, tg = _terms_from_matrices(omega, psi_L, psi_U, D_goal, use_limits, inds=inds)
        om = omega != 0
        if inds is not None and not use_limits:
            om = np.zeros((N, N), dtype=bool)
            om[inds[0], inds[1]] = True
        return cls({
            "n_nodes": N, "term_i": ti, "term_j": tj, "term_kind": tk, "term_target": tt, "term_goal": tg,
            "n_goal": 0, "anchor_node": None, "omega": np.ascontiguousarray(om | om.T, dtype=np.uint8),
        })

    # ------------------------------------------------------------------
    def goal_row_from_matrix(self, D_goal):
        """goal_d2 row of one problem from its full D_goal matrix (host helper for tests)."""
        gs = self._a["goal_slot"]
        row = np.zeros(self.n_goal)
        ii, jj = np.nonzero(np.triu(gs >= 0))
        for i, j in zip(ii, jj):
            row[gs[i, j]] = D_goal[i, j]
        return row
