"""Distance-geometric problem graph of a serial revolute manipulator (host side).

Mirrors `ProblemGraph` / `ProblemGraphRevolute` of the reference
(graphs/graph_base.py:19-279, graphs/graph_revolute.py:15-318) for the methods
the Riemannian IK path calls: construction of the base tetrahedron, structure
and joint-limit edges, `from_pose`, `distance_bound_matrices`,
`add_spherical_obstacle`, `realization`, `joint_variables`,
`check_distance_limits`.

The reference stores everything as networkx attribute dictionaries and deep
copies the graph for every goal pose.  Here the graph is a handful of dense
N x N arrays built once per robot (it is the input of the device-side plan,
graphik_b200/plan.py); a goal only changes 2*(4+n_obstacles) entries, so
`from_pose` returns a light `GoalGraph` view and the batched path never
materialises per-goal matrices on the host at all.

Node order (defines every matrix row): p0, x, y, q0, p1, q1, ..., pn, qn,
then obstacles in insertion order (graph_revolute.py:32-51,59-81;
graph_base.py:192).
"""
from typing import Dict, List, Optional

import numpy as np

from graphik_b200.robots.robot_revolute import RobotRevolute
from graphik_b200.utils.se3 import SE3, as_matrix4, hat

BELOW, ABOVE = "below", "above"
ROBOT, BASE, OBSTACLE = "robot", "base", "obstacle"


def _max_min_distance_revolute(r, P, C, N):
    """geometry.py:45-58: extreme distances from P to a circle (centre C, normal N, radius r)."""
    delta = P - C
    along = N.dot(delta) ** 2
    radial = np.linalg.norm(np.cross(N, delta))
    d_min_s = along + (radial - r) ** 2
    d_max_s = along + (radial + r) ** 2
    d_min = np.sqrt(d_min_s) if d_min_s > 0 else 0
    d_max = np.sqrt(d_max_s) if d_max_s > 0 else 0
    return d_max, d_min


class GoalGraph:
    """Per-goal matrices of the IK problem: what `graph.from_pose(T)` stands for.

    dist/lower/upper: N x N, NaN where the attribute is absent; `edge` marks
    node pairs joined by an edge (symmetric)."""

    def __init__(self, node_ids, edge, dist, lower, upper, pos):
        self.node_ids = node_ids
        self.edge, self.dist, self.lower, self.upper, self.pos = edge, dist, lower, upper, pos

    def number_of_nodes(self):
        return len(self.node_ids)


class ProblemGraphRevolute:
    def __init__(self, robot: RobotRevolute, params: Dict = {}):
        self.robot = robot
        self.dim = robot.dim
        self.axis_length = params.get("axis_length", 1)
        self.obstacle_semantics = params.get("obstacle_semantics", "reference")
        n = robot.n
        self.node_ids: List[str] = ["p0", "x", "y", "q0"]
        for i in range(1, n + 1):
            self.node_ids += ["p%d" % i, "q%d" % i]
        self._index = {u: k for k, u in enumerate(self.node_ids)}
        self.types: List = [[ROBOT, BASE], [BASE], [BASE], [ROBOT, BASE]] + [[ROBOT]] * (2 * n)
        N = len(self.node_ids)
        self.edge = np.zeros((N, N), dtype=bool)
        self.dist = np.full((N, N), np.nan)
        self.lower = np.full((N, N), np.nan)
        self.upper = np.full((N, N), np.nan)
        self.below = np.zeros((N, N), dtype=bool)
        self.above = np.zeros((N, N), dtype=bool)
        self.pos = np.full((N, 3), np.nan)  # POS attribute (anchors only)
        self.obstacles: List[dict] = []
        self.limited_joints: List[str] = []

        self._base_and_structure()
        self._set_limits()
        self._root_angle_limits()

    # ------------------------------------------------------------------ build
    def idx(self, name: str) -> int:
        return self._index[name]

    def number_of_nodes(self) -> int:
        return len(self.node_ids)

    def _set_edge(self, u, v, dist=None, lower=None, upper=None, bounded=None):
        i, j = self._index[u], self._index[v]
        for a, b in ((i, j), (j, i)):
            self.edge[a, b] = True
            if dist is not None:
                self.dist[a, b] = dist
            if lower is not None:
                self.lower[a, b] = lower
            if upper is not None:
                self.upper[a, b] = upper
            if bounded is not None:
                self.below[a, b] = bounded == BELOW
                self.above[a, b] = bounded == ABOVE

    def _aux(self, T):
        """Position of the auxiliary point: frame origin moved axis_length along its z."""
        return T[:3, :3] @ np.array([0.0, 0.0, self.axis_length]) + T[:3, 3]

    def _base_and_structure(self):
        L, T0, n = self.axis_length, self.robot.T0, self.robot.n
        # base tetrahedron (graph_revolute.py:32-57)
        base_pos = {"p0": np.array([0, 0, 0]), "x": np.array([L, 0, 0]),
                    "y": np.array([0, -L, 0]), "q0": np.array([0, 0, L])}
        for u, p in base_pos.items():
            self.pos[self._index[u]] = p
        for u, v in (("p0", "x"), ("p0", "y"), ("p0", "q0"), ("x", "y"), ("y", "q0"), ("q0", "x")):
            d = np.linalg.norm(base_pos[u] - base_pos[v])
            self._set_edge(u, v, d, d, d)
        # structure: p_i, q_i rigidly tied to p_{i-1}, q_{i-1} (graph_revolute.py:59-95)
        pts = {}
        for i in range(n + 1):
            pts["p%d" % i] = T0[i][:3, 3]
            pts["q%d" % i] = self._aux(T0[i])
            d = np.linalg.norm(pts["p%d" % i] - pts["q%d" % i])
            self._set_edge("p%d" % i, "q%d" % i, d, d, d)
            if i:
                for u in ("p%d" % (i - 1), "q%d" % (i - 1)):
                    for v in ("p%d" % i, "q%d" % i):
                        d = np.linalg.norm(pts[u] - pts[v])
                        self._set_edge(u, v, d, d, d)

    def _limit_edge(self, u, v, P, T1, T2, ub_joint):
        """Distance interval swept by point T2 (rigid w.r.t. joint frame T1's successor)
        seen from fixed point P while the joint about T1's z turns
        (graph_revolute.py:189-239 and :126-165 share this body)."""
        t1, t2 = T1[:3, 3], T2[:3, 3]
        Nz = T1[:3, 2]
        C = t1 + (Nz.dot(t2 - t1)) * Nz
        r = np.linalg.norm(t2 - C)
        d_max, d_min = _max_min_distance_revolute(r, P, C, Nz)
        d = np.linalg.norm(t2 - P)
        if d_max == d_min:
            limit = False
        elif d == d_max:
            limit = BELOW
        elif d == d_min:
            limit = ABOVE
        else:
            limit = None
        if limit:
            T_rel = np.linalg.inv(T1) @ T2
            Rz = np.eye(4)
            c, s = np.cos(ub_joint), np.sin(ub_joint)
            Rz[:2, :2] = [[c, -s], [s, c]]
            d_limit = np.linalg.norm(((T1 @ Rz) @ T_rel)[:3, 3] - P)
            if limit == ABOVE:
                d_max = d_limit
            else:
                d_min = d_limit
        self._set_edge(u, v, d_max if d_max == d_min else None, d_min, d_max, limit if limit else "")
        return bool(limit)

    def _with_aux(self, T):
        Tq = T.copy()
        Tq[:3, 3] = self._aux(T)
        return Tq

    def _set_limits(self):
        T0, n, ub = self.robot.T0, self.robot.n, self.robot.ub
        for i in range(2, n + 1):
            cur, mid, prev = i, i - 1, i - 2
            for a in ("p", "q"):
                for b in ("p", "q"):
                    Ta = T0[prev] if a == "p" else self._with_aux(T0[prev])
                    Tb = T0[cur] if b == "p" else self._with_aux(T0[cur])
                    if self._limit_edge("%s%d" % (a, prev), "%s%d" % (b, cur), Ta[:3, 3], T0[mid], Tb,
                                        ub["p%d" % cur]):
                        self.limited_joints.append("p%d" % cur)

    def _root_angle_limits(self):
        T0, ub = self.robot.T0, self.robot.ub
        if self.robot.n < 1:
            return
        for base_node in ("x", "y"):
            for node in ("p1", "q1"):
                P = self.pos[self._index[base_node]].copy()
                T2 = T0[1] if node == "p1" else self._with_aux(T0[1])
                if self._limit_edge(base_node, node, P, T0[0], T2, ub["p1"]):
                    self.limited_joints.append("p1")

    # ------------------------------------------------------------- obstacles
    def add_anchor_node(self, name: str, data: Dict):
        """graph_base.py:182-199: a point with known position, tied by exact
        distances to every other known-position node."""
        if "pos" not in data:
            raise KeyError("Node needs to gave a position to be added.")
        p = np.asarray(data["pos"], dtype=float)
        N = len(self.node_ids)
        for name_arr in ("edge", "below", "above"):
            a = getattr(self, name_arr)
            g = np.zeros((N + 1, N + 1), dtype=bool)
            g[:N, :N] = a
            setattr(self, name_arr, g)
        for name_arr in ("dist", "lower", "upper"):
            a = getattr(self, name_arr)
            g = np.full((N + 1, N + 1), np.nan)
            g[:N, :N] = a
            setattr(self, name_arr, g)
        self.pos = np.vstack([self.pos, p[None, :]])
        self.node_ids.append(name)
        self._index[name] = N
        self.types.append(data.get("type", OBSTACLE))
        for k in range(N):
            if not np.isnan(self.pos[k, 0]):
                d = np.linalg.norm(self.pos[k] - p)
                self._set_edge(self.node_ids[k], name, d, d, d)

    def add_spherical_obstacle(self, name: str, position, radius: float):
        """graph_base.py:201-211.  In the reference the robot-node loop never
        fires on revolute graphs (`node_type == ROBOT` compares a list with a
        string), so an obstacle only adds an anchor -- obstacle_semantics
        "reference" reproduces exactly that.  "intended" adds the lower-bound
        edges p_i -- obstacle (LOWER=radius, UPPER=100, BELOW) for i = 1..n."""
        self.add_anchor_node(name, {"pos": np.asarray(position, dtype=float), "type": OBSTACLE})
        self.obstacles.append({"name": name, "pos": np.asarray(position, float), "radius": float(radius)})
        if self.obstacle_semantics == "intended":
            for i in range(1, self.robot.n + 1):
                self._set_edge("p%d" % i, name, None, radius, 100, BELOW)

    def clear_obstacles(self):
        keep = [k for k, t in enumerate(self.types) if t != OBSTACLE]
        ix = np.ix_(keep, keep)
        for name_arr in ("edge", "below", "above", "dist", "lower", "upper"):
            setattr(self, name_arr, getattr(self, name_arr)[ix])
        self.pos = self.pos[keep]
        self.node_ids = [self.node_ids[k] for k in keep]
        self.types = [self.types[k] for k in keep]
        self._index = {u: k for k, u in enumerate(self.node_ids)}
        self.obstacles = []

    # ------------------------------------------------------------- per goal
    @property
    def anchor_indices(self) -> np.ndarray:
        """Indices of nodes with a goal-independent known position."""
        return np.nonzero(~np.isnan(self.pos[:, 0]))[0]

    def goal_points(self, T_goal) -> np.ndarray:
        """_pose_goal (graph_revolute.py:243-249): p_n = t, q_n = t + R e_z * axis_length."""
        T = as_matrix4(T_goal)
        return np.stack([T[:3, 3], self._aux(T)])

    def from_pos(self, P: Dict[str, np.ndarray]) -> GoalGraph:
        """graph_base.py:146-165 + dgp.py:124-147: pin the given nodes and join every
        pair of known-position nodes that has no DIST yet by an exact-distance edge."""
        pos = self.pos.copy()
        for name, p in P.items():
            if name in self._index:
                pos[self._index[name]] = np.asarray(p, dtype=float)
        edge, dist = self.edge.copy(), self.dist.copy()
        lower, upper = self.lower.copy(), self.upper.copy()
        known = np.nonzero(~np.isnan(pos[:, 0]))[0]
        for a, i in enumerate(known):
            for j in known[a + 1:]:
                if np.isnan(dist[i, j]):
                    d = np.linalg.norm(pos[i] - pos[j])
                    for u, v in ((i, j), (j, i)):
                        edge[u, v] = True
                        dist[u, v] = lower[u, v] = upper[u, v] = d
        return GoalGraph(self.node_ids, edge, dist, lower, upper, pos)

    def from_pose(self, T_goal) -> GoalGraph:
        """graph_base.py:171-180."""
        if isinstance(T_goal, dict):
            P = {}
            for u, T in T_goal.items():
                pq = self.goal_points(T)
                P[u], P["q" + u[1:]] = pq[0], pq[1]
            return self.from_pos(P)
        n = self.robot.n
        pq = self.goal_points(T_goal)
        return self.from_pos({"p%d" % n: pq[0], "q%d" % n: pq[1]})

    def distance_bound_matrices(self):
        """graph_base.py:262-279: squared lower / upper limits of BELOW / ABOVE edges."""
        L = np.where(self.below, self.lower, 0.0) ** 2
        U = np.where(self.above, self.upper, 0.0) ** 2
        return L, U

    # ---------------------------------------------------------- realization
    def points_from_frames(self, T) -> np.ndarray:
        """Node positions for joint frames T[..., n+1, 4, 4] -> [..., N, 3]
        (base nodes and obstacles at their fixed positions)."""
        T = np.asarray(T, dtype=float)
        lead = T.shape[:-3]
        N = len(self.node_ids)
        Y = np.broadcast_to(self.pos, lead + (N, 3)).copy()
        n = self.robot.n
        L = self.axis_length
        p = T[..., :, :3, 3]
        q = p + L * T[..., :, :3, 2]
        Y[..., 0, :] = p[..., 0, :]
        Y[..., 3, :] = q[..., 0, :]
        Y[..., 4:4 + 2 * n:2, :] = p[..., 1:, :]
        Y[..., 5:5 + 2 * n:2, :] = q[..., 1:, :]
        return Y

    def realization_points(self, joint_angles) -> np.ndarray:
        """Positions of all nodes for a configuration (or a batch Q[B,n])."""
        Q = self.robot.q_array(joint_angles)
        T = self.robot.fk_all(np.atleast_2d(Q))
        Y = self.points_from_frames(T)
        return Y[0] if Q.ndim == 1 else Y

    def realization(self, joint_angles) -> GoalGraph:
        """graph_base.py:112-121: complete graph over the realised points."""
        Y = self.realization_points(joint_angles)
        return self.from_pos({u: Y[k] for k, u in enumerate(self.node_ids)})

    def distance_matrix_from_joints(self, joint_angles) -> np.ndarray:
        Y = self.realization_points(joint_angles)
        diff = Y[:, None, :] - Y[None, :, :]
        return np.sum(diff * diff, axis=-1)

    # ------------------------------------------------- joint angle recovery
    def joint_variables(self, G, T_final: Optional[Dict] = None) -> Dict[str, float]:
        """graph_revolute.py:251-318 for one realization.  `G` may be a GoalGraph
        (uses its positions) or an [N,3] array of points in node order."""
        Y = G.pos if isinstance(G, GoalGraph) else np.asarray(G, dtype=float)
        T_goal = None
        if T_final is not None:
            T_goal = as_matrix4(next(iter(T_final.values())) if isinstance(T_final, dict) else T_final)[None]
        q = self.joint_variables_batch(Y[None], T_goal)[0]
        return self.robot.q_dict(q)

    def joint_variables_batch(self, Y, T_goal=None) -> np.ndarray:
        """Vectorised restatement of joint_variables: Y[B,N,3] (+ T_goal[B,4,4]) -> q[B,n]."""
        Y = np.asarray(Y, dtype=float)
        Bsz, n, T0 = Y.shape[0], self.robot.n, self.robot.T0

        def unit(v):
            nv = np.linalg.norm(v, axis=-1, keepdims=True)
            return np.where(nv == 0, v, v / np.where(nv == 0, 1.0, nv))

        p0 = Y[:, 0]
        x, y, z = unit(Y[:, 1] - p0), unit(Y[:, 2] - p0), unit(Y[:, 3] - p0)
        R = np.stack([x, -y, z], axis=-1)            # columns x, -y, z
        # B^{-1} p = R^T (p - p0)
        def to_base(p):
            return np.einsum("bji,bj->bi", R, p - p0)

        Tprev = np.broadcast_to(T0[0], (Bsz, 4, 4)).copy()
        ez = hat(np.array([0.0, 0.0, 1.0]))
        q = np.zeros((Bsz, n))
        tol = 1e-10
        T_rel = None
        for i in range(1, n + 1):
            T_rel = self.robot.T_rel[i - 1]
            Tq0 = self._with_aux(T0[i])
            qs_0 = (np.linalg.inv(T0[i - 1]) @ Tq0)[:3, 3]
            pc, qc = Y[:, 2 + 2 * i], Y[:, 3 + 2 * i]
            qn = pc + unit(qc - pc)
            qb = to_base(qn)
            Rp, tp = Tprev[:, :3, :3], Tprev[:, :3, 3]
            qs = np.einsum("bji,bj->bi", Rp, qb - tp)
            num = -np.einsum("i,bi->b", qs_0 @ ez, qs)
            den = np.einsum("i,bi->b", qs_0 @ (ez @ ez.T), qs)
            th = np.arctan2(num, den)
            q[:, i - 1] = th
            c, s = np.cos(th), np.sin(th)
            Rz = np.zeros((Bsz, 4, 4))
            Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1] = c, -s, s, c
            Rz[:, 2, 2] = Rz[:, 3, 3] = 1.0
            Tprev = (Tprev @ Rz) @ T_rel
        if T_goal is not None and np.linalg.norm(np.cross(T_rel[:3, 3], np.array([0.0, 0.0, 1.0]))) < tol:
            T_th = np.linalg.inv(Tprev) @ np.asarray(T_goal, dtype=float)
            extra = np.arctan2(T_th[:, 1, 0], T_th[:, 0, 0])
            q[:, n - 1] = np.mod(q[:, n - 1] + extra + np.pi, 2 * np.pi) - np.pi
        return q

    def get_pose(self, joint_angles, query_node: str) -> SE3:
        T = self.robot.pose(joint_angles, "p" + query_node[1:])
        if query_node[0] == "q":
            T = SE3.from_matrix(self._with_aux(T.as_matrix()))
        return T

    # ---------------------------------------------------------- limit check
    def check_distance_limits(self, G, tol=1e-10, semantics: Optional[str] = None) -> List[Dict]:
        """graph_base.py:219-260.

        semantics="reference": the reference's node-type comparisons
        (`typ[u] == ROBOT`) compare a list with a string on revolute graphs and are
        never true, so the reference ALWAYS returns [] here; reproduced as is.
        semantics="intended": report every BELOW/ABOVE edge whose realised distance
        leaves [LOWER - tol, UPPER + tol]."""
        semantics = semantics or "reference"
        if semantics == "reference":
            return []
        Y = G.pos if isinstance(G, GoalGraph) else np.asarray(G, dtype=float)
        out = []
        iu, ju = np.nonzero(np.triu(self.below | self.above))
        for i, j in zip(iu, ju):
            d = np.linalg.norm(Y[i] - Y[j])
            kind = OBSTACLE if OBSTACLE in (self.types[i], self.types[j]) else "joint"
            if d < self.lower[i, j] - tol:
                out.append({"edge": (self.node_ids[i], self.node_ids[j]), "value": d - self.lower[i, j],
                            "type": kind, "side": "lower_limit"})
            if d > self.upper[i, j] + tol:
                out.append({"edge": (self.node_ids[i], self.node_ids[j]), "value": d - self.upper[i, j],
                            "type": kind, "side": "upper_limit"})
        return out
