"""`RiemannianSolver` / `solve_with_riemannian` with the reference's signatures
(graphik/solvers/riemannian_solver.py:33-234), executed by the CUDA engine.

Drop-in use (the reference's README / experiments/riemannian_example.py):

    from graphik_b200.utils.roboturdf import load_ur10
    from graphik_b200.solvers.riemannian_solver import solve_with_riemannian
    robot, graph = load_ur10()
    T_goal = robot.pose(robot.random_configuration(), f"p{robot.n}")
    q_sol, points = solve_with_riemannian(graph, T_goal)

New, batched entry points: `RiemannianSolver.solve_batch(T_goals)` and
`solve_batch_with_riemannian(graph, T_goals)`.

`solver="ConjugateGradient"` (riemannian_solver.py:52-60) runs `gik_cg_solve`: pymanopt 0.2.5's published
ConjugateGradient + LineSearchAdaptive (third-party source outside the reference tree: parity unpinned).
"""
import time

import numpy as np

from graphik_b200.engine import BatchIK, make_opts
from graphik_b200.plan import Plan
from graphik_b200.utils.se3 import as_matrix4


class RiemannianSolver:
    def __init__(self, graph, params={}):
        self.params = params
        self.graph = graph
        self.dim = graph.dim
        self.N = graph.number_of_nodes()
        solver_type = params.get("solver", "TrustRegions")
        if solver_type not in ("TrustRegions", "ConjugateGradient"):
            raise ValueError("params[\"solver\"] must be one of 'ConjugateGradient', 'TrustRegions'")
        # TrustRegions: mingradnorm 5e-10, maxiter 3000, theta 1, kappa 0.1 (:44-50); ConjugateGradient: mingradnorm 1e-9,
        # maxiter 1e5, minstepsize 1e-10, orth_value 1e11, HagerZhang (:52-60; pymanopt 0.2.5's published algorithm
        # restated in csrc/gik_cg.cu -- the third-party source is not in the reference tree, parity unpinned)
        self.opts = make_opts(params)
        self._engine = None
        self._static_cache = {}

    # -- engine bound to the graph (pose-goal pipeline) -------------------------
    @property
    def engine(self) -> BatchIK:
        """Device plan + buffers for this graph.  The reference builds a RiemannianSolver per call
        (riemannian_solver.py:222); to keep that usage cheap the engine is cached on the graph object and
        rebuilt only when the graph's edge data changed (e.g. an obstacle was added)."""
        g = self.graph
        import torch
        dev = torch.cuda.current_device() if torch.cuda.is_available() else -1   # a plan belongs to one device
        sig = hash((dev, tuple(sorted((k, repr(v)) for k, v in self.params.items())),
                    g.number_of_nodes(), g.dist.tobytes(), g.lower.tobytes(), g.upper.tobytes(),
                    g.below.tobytes(), g.above.tobytes()))
        cached = getattr(g, "_gik_engine_cache", None)
        if self._engine is None and cached is not None and cached[0] == sig:
            self._engine = cached[1]
        if self._engine is None or getattr(self, "_engine_sig", sig) != sig:
            self._engine = BatchIK(g, self.params)
        self._engine_sig = sig
        g._gik_engine_cache = (sig, self._engine)
        self.N = g.number_of_nodes()
        return self._engine

    # -- reference API -------------------------------------------------------------
    def solve(self, D_goal, omega, use_limits=False, bounds=None, Y_init=None, jit=True, output_log=True):
        """riemannian_solver.py:178-218 for one explicit problem.  `jit` is accepted for
        compatibility; both values run the CUDA path.  Returns optlog["final_values"]
        ({x, f(x), time, gradnorm, iterations}) or, with output_log=False, only x."""
        D_goal = np.asarray(D_goal, dtype=float)
        omega = np.asarray(omega, dtype=float)
        if use_limits:
            psi_L, psi_U = self.graph.distance_bound_matrices()
        else:
            psi_L, psi_U = 0 * omega, 0 * omega
        key = (use_limits, D_goal.tobytes(), omega.tobytes(), psi_L.tobytes(), psi_U.tobytes())
        eng = self._static_cache.get(key)
        if eng is None:
            self._static_cache.clear()
            eng = BatchIK(plan=Plan.from_matrices(D_goal, omega, psi_L, psi_U, use_limits=use_limits))
            eng.opts = self.opts
            self._static_cache[key] = eng
        t0 = time.time()
        if bounds is not None:
            Y0 = eng.init_from_bounds(np.asarray(bounds[0], float)[None], np.asarray(bounds[1], float)[None])
        elif Y_init is None:
            raise Exception("If not using bounds, provide an initialization!")
        else:
            Y0 = np.asarray(Y_init, dtype=float)[None]
        out = eng.solve_points(None, Y0)
        x = out["x"][0].cpu().numpy()
        if not output_log:
            return x
        return {"x": x, "f(x)": float(out["f(x)"][0]), "time": time.time() - t0,
                "gradnorm": float(out["gradnorm"][0]), "iterations": int(out["iterations"][0])}

    @staticmethod
    def generate_initialization(bounds, dim, omega, psi_L=None, psi_U=None):
        """riemannian_solver.py:67-75 on the device (B = 1)."""
        if dim != 3:
            raise NotImplementedError("graphik_b200 supports dim = 3")
        omega = np.asarray(omega, dtype=float)
        N = omega.shape[0]
        eng = BatchIK(plan=Plan.from_matrices(np.zeros((N, N)), omega, use_limits=False))
        return eng.init_from_bounds(np.asarray(bounds[0], float)[None], np.asarray(bounds[1], float)[None])[0] \
            .cpu().numpy()

    # -- batched API ---------------------------------------------------------------
    def solve_batch(self, T_goals, Y_init=None, check=True, as_numpy=False):
        """Full pipeline for T_goals[B,4,4] (array or CUDA tensor): device tensors
        q, x, f(x), gradnorm, iterations, status, n_inner (+ pos_err, rot_err)."""
        eng = self.engine
        eng.opts = self.opts
        out = eng.solve(T_goals, Y_init=Y_init, check=check)
        if as_numpy:
            out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items() if v is not None}
        return out


    def stream(self, **kw):
        """Throughput mode: an `IKStream` (graphik_b200/pipeline.py) over this solver's engine -- batches are
        submitted asynchronously, their small kernels overlap other batches' trust-region launches and goals
        that need many more iterations than the rest finish in the shadow of later batches.
        Keywords: slots, inner_budget, carry_capacity, to_host."""
        from graphik_b200.pipeline import IKStream
        eng = self.engine
        eng.opts = self.opts
        return IKStream(eng, **kw)

    def solve_batches(self, batches, **kw):
        """solve_batch for a sequence of T_goals arrays / tensors through one IKStream; returns the list of
        result dicts (device tensors), complete."""
        st = self.stream(**kw)
        tickets = [st.submit(T) for T in batches]
        st.drain()
        return [st.result(t) for t in tickets]


def solve_with_riemannian(graph, T_goal, use_jit=True, jit=None, limit_semantics="reference"):
    """riemannian_solver.py:220-234: (q_sol dict, points[N,3]) or (None, None).

    limit_semantics="reference" reproduces the reference as shipped, whose check_distance_limits never
    reports a violation on revolute graphs (graph_base.py:226-258 compares a list with a string), so a
    solution is always returned; "intended" runs the limit check (on the device) and returns (None, None)
    when a joint-limit or obstacle distance bound is broken by more than 1e-6."""
    solver = RiemannianSolver(graph)
    T = as_matrix4(T_goal)
    out = solver.solve_batch(T[None], check=(limit_semantics == "intended"))
    Y = out["x"][0].cpu().numpy()
    q = out["q"][0].cpu().numpy()
    q_sol = graph.robot.q_dict(q)
    # "reference": riemannian_solver.py:230 calls check_distance_limits, which (as shipped) returns []
    # for every revolute graph -- nothing to evaluate
    n_broken = int(out["n_broken"][0]) if limit_semantics == "intended" else 0
    if n_broken > 0:
        return None, None
    return q_sol, Y


def solve_batch_with_riemannian(graph, T_goals, params=None, as_numpy=True):
    """Batched form of solve_with_riemannian: every goal of T_goals[B,4,4] in one launch."""
    return RiemannianSolver(graph, params or {}).solve_batch(T_goals, as_numpy=as_numpy)
