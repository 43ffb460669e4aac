"""Generate golden fixtures by RUNNING THE UNMODIFIED REFERENCE in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (read-only) and the stand-ins
in oracle/shims (see ref_runner.py); cannot run on the GPU box.  Outputs:

  graphik_b200/robots/models/<robot>.json   zero-configuration joint frames
        extracted by the reference's URDF loader (roboturdf.py:122-153,226-264)
        -- numeric kinematic parameters only, so that the product can build
        the same robots without the reference tree or a URDF parser.
  tests/golden/<robot>_graph.npz            static ProblemGraphRevolute data:
        node order, DIST/LOWER/UPPER/BOUNDED per edge (graph_revolute.py:15-241)
  tests/golden/<robot>_goals.npz            per seeded goal: q_goal, T_goal,
        D_goal, omega, psi_L, psi_U (riemannian_solver.py:220-226), lb, ub
        (dgp.py:192-231), Y_init (riemannian_solver.py:67-75), the solver's
        final_values (x, f, gradnorm, iterations), q_sol
        (graph_revolute.py:251-318) and a per-outer-iteration trace
        (Delta, numit, stop_reason, fx_prop, accepted, gradnorm) recorded by
        proxying the pymanopt Problem handed to TrustRegions.solve.
  tests/golden/costgrd_vectors.npz          random (Y, w) -> the reference's own
        numba-AOT costgrd outputs (costs.py) and PSDFixedRank.proj outputs
        (fixed_rank_psd_sym.py:91-113) on the UR10 / KUKA problem matrices.

Usage:  python oracle/gen_golden.py [--robots ur10 kuka ...] [--goals 6]
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from ref_runner import load_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODELS = os.path.join(ROOT, "graphik_b200", "robots", "models")


def _loaders():
    from graphik.utils import roboturdf as ru
    return {
        "ur10": ru.load_ur10,
        "kuka": ru.load_kuka,
        "lwa4d": ru.load_schunk_lwa4d,
        "lwa4p": ru.load_schunk_lwa4p,
        "panda": ru.load_panda,
    }


def random_dh_chain(n, seed):
    """tests/test_joint_variables.py:80-102 with fixed n and seed (BASELINE config 4)."""
    from graphik.graphs import ProblemGraphRevolute
    from graphik.robots import RobotRevolute
    rng_state = np.random.get_state()
    np.random.seed(seed)
    a = np.random.rand(n)
    d = np.random.rand(n)
    al = np.random.rand(n) * np.pi / 2 - 2 * np.random.rand(n) * np.pi / 2
    th = 0 * np.ones(n)
    np.random.set_state(rng_state)
    params = {"a": a, "alpha": al, "d": d, "theta": th, "modified_dh": False, "num_joints": n}
    robot = RobotRevolute(params)
    graph = ProblemGraphRevolute(robot)
    return robot, graph, {"a": a.tolist(), "d": d.tolist(), "alpha": al.tolist(), "theta": th.tolist()}


def dump_model(name, robot, extra=None):
    os.makedirs(MODELS, exist_ok=True)
    T0 = [robot.nodes["p%d" % i]["T0"].as_matrix().tolist() for i in range(robot.n + 1)]
    model = {"name": name, "num_joints": int(robot.n), "T_zero": T0,
             "source": "extracted by oracle/gen_golden.py from the reference's loader"}
    if extra:
        model["dh"] = extra
    with open(os.path.join(MODELS, name + ".json"), "w") as f:
        json.dump(model, f, indent=1)


def graph_static(graph):
    from graphik.utils.constants import ABOVE, BELOW, BOUNDED, DIST, LOWER, POS, UPPER
    ids = graph.node_ids
    N = len(ids)
    ix = {u: i for i, u in enumerate(ids)}
    out = {k: np.full((N, N), np.nan) for k in ("dist", "lower", "upper")}
    below = np.zeros((N, N), bool)
    above = np.zeros((N, N), bool)
    edge = np.zeros((N, N), bool)
    for u, v, d in graph.edges(data=True):
        i, j = ix[u], ix[v]
        edge[i, j] = True
        for key, lab in (("dist", DIST), ("lower", LOWER), ("upper", UPPER)):
            if lab in d:
                out[key][i, j] = d[lab]
        if BOUNDED in d:
            below[i, j] = BELOW in d[BOUNDED]
            above[i, j] = ABOVE in d[BOUNDED]
    pos = np.full((N, 3), np.nan)
    for u, d in graph.nodes(data=True):
        if POS in d:
            pos[ix[u]] = d[POS]
    return dict(node_ids=np.array(ids), edge=edge, below=below, above=above, pos=pos, **out)


class _ProblemProxy:
    """Wraps the pymanopt Problem given to TrustRegions.solve and records calls."""

    def __init__(self, problem, log):
        self._p = problem
        self.manifold = problem.manifold
        self.verbosity = problem.verbosity
        self.precon = problem.precon
        self.hess = problem.hess
        self._log = log
        p_cost, p_grad = problem.cost, problem.grad

        def cost(x):
            f = p_cost(x)
            log.append(("cost", float(f)))
            return f

        def grad(x):
            g = p_grad(x)
            log.append(("grad", float(np.linalg.norm(g))))
            return g

        self.cost, self.grad = cost, grad


def trace_from_log(log):
    """Reassemble per-outer-iteration records from the call log."""
    it = iter(log)
    kind, f0 = next(it)
    assert kind == "cost"
    kind, g0 = next(it)
    assert kind == "grad"
    rows = []
    cur = None
    for kind, val in it:
        if kind == "tcg":
            if cur is not None:
                rows.append(cur)
            cur = {"Delta": val[0], "numit": val[1], "stop": val[2], "fx_prop": np.nan,
                   "accepted": 0, "gradnorm": np.nan}
        elif kind == "cost":
            cur["fx_prop"] = val
        elif kind == "grad":
            cur["accepted"] = 1
            cur["gradnorm"] = val
    if cur is not None:
        rows.append(cur)
    arr = np.array([[r["Delta"], r["numit"], r["stop"], r["fx_prop"], r["accepted"], r["gradnorm"]]
                    for r in rows], dtype=float).reshape(-1, 6)
    return f0, g0, arr


def solve_traced(graph, T_goal):
    """solve_with_riemannian (riemannian_solver.py:220-234) with every
    intermediate kept; the calls are the reference's own, in its order."""
    from graphik.solvers.riemannian_solver import RiemannianSolver
    from graphik.solvers.trust_region import TrustRegions
    from graphik.utils.dgp import (adjacency_matrix_from_graph, bound_smoothing,
                                   distance_matrix_from_graph, graph_from_pos)
    log = []
    orig_solve = TrustRegions.solve
    orig_tcg = TrustRegions._truncated_conjugate_gradient

    def solve(self, problem, x=None, **kw):
        return orig_solve(self, _ProblemProxy(problem, log), x=x, **kw)

    def tcg(self, problem, x, fgradx, eta, Delta, theta, kappa, mininner, maxinner):
        out = orig_tcg(self, problem, x, fgradx, eta, Delta, theta, kappa, mininner, maxinner)
        log.append(("tcg", (float(Delta), int(out[2]), int(out[3]))))
        return out

    TrustRegions.solve = solve
    TrustRegions._truncated_conjugate_gradient = tcg
    try:
        G = graph.from_pose(T_goal)
        solver = RiemannianSolver(graph)
        D_goal = distance_matrix_from_graph(G)
        omega = adjacency_matrix_from_graph(G)
        lb, ub = bound_smoothing(G)
        psi_L, psi_U = graph.distance_bound_matrices()
        Y_init = RiemannianSolver.generate_initialization((lb, ub), graph.dim, omega, psi_L, psi_U)
        sol = solver.solve(D_goal, omega, use_limits=True, bounds=(lb, ub), jit=True)
    finally:
        TrustRegions.solve = orig_solve
        TrustRegions._truncated_conjugate_gradient = orig_tcg
    G_sol = graph_from_pos(sol["x"], graph.node_ids)
    q_sol = graph.joint_variables(G_sol, {"p%d" % graph.robot.n: T_goal})
    broken = graph.check_distance_limits(graph.realization(q_sol), tol=1e-6)
    f0, g0, trace = trace_from_log(log)
    n = graph.robot.n
    return dict(D_goal=D_goal, omega=omega, psi_L=psi_L, psi_U=psi_U, lb=lb, ub=ub, Y_init=Y_init,
                Y_sol=sol["x"], f=float(sol["f(x)"]), gradnorm=float(sol["gradnorm"]),
                iterations=int(sol["iterations"]), f0=f0, g0=g0, trace=trace,
                q_sol=np.array([q_sol["p%d" % i] for i in range(1, n + 1)]),
                n_broken=len(broken))


def dump_goals(name, robot, graph, n_goals, seed):
    n = robot.n
    np.random.seed(seed)
    recs = []
    for g in range(n_goals):
        q = robot.random_configuration()
        T_goal = robot.pose(q, "p%d" % n)
        r = solve_traced(graph, T_goal)
        r["q_goal"] = np.array([q["p%d" % i] for i in range(1, n + 1)])
        r["T_goal"] = T_goal.as_matrix()
        T_sol = robot.pose({"p%d" % (i + 1): r["q_sol"][i] for i in range(n)}, "p%d" % n).as_matrix()
        r["pose_err"] = float(np.linalg.norm(T_sol[:3, 3] - r["T_goal"][:3, 3]))
        recs.append(r)
        print("  %s goal %d: iters %d f %.3e |g| %.3e pos_err %.2e" %
              (name, g, r["iterations"], r["f"], r["gradnorm"], r["pose_err"]), flush=True)
    out = {}
    for key in recs[0]:
        if key == "trace":
            L = max(len(r["trace"]) for r in recs)
            tr = np.full((len(recs), L, 6), np.nan)
            for i, r in enumerate(recs):
                tr[i, :len(r["trace"])] = r["trace"]
            out["trace"] = tr
        else:
            out[key] = np.array([r[key] for r in recs])
    np.savez_compressed(os.path.join(GOLDEN, name + "_goals.npz"), **out)
    return out


def dump_costgrd_vectors(problems):
    """The reference's own costgrd / PSDFixedRank.proj on random inputs."""
    from graphik.solvers import costgrd
    from graphik.utils.manifolds.fixed_rank_psd_sym import PSDFixedRank
    rng = np.random.default_rng(1234)
    out = {}
    for name, p in problems.items():
        D, om, pL, pU = (np.ascontiguousarray(p[k][0]) for k in ("D_goal", "omega", "psi_L", "psi_U"))
        N = D.shape[0]
        diff = pL != pU
        inds = np.nonzero(np.triu(om) + np.triu(diff * (pL > 0)) + np.triu(diff * (pU > 0)))
        jinds = np.nonzero(np.triu(om))
        K = 8
        Y = rng.normal(size=(K, N, 3))
        # half the samples near a true realization so that hinge terms switch on/off
        Y[K // 2:] = p["Y_sol"][0] + 0.05 * rng.normal(size=(K - K // 2, N, 3))
        W = rng.normal(size=(K, N, 3))
        res = {k: [] for k in ("lcost", "lgrad", "lhess", "jcost", "jgrad", "jhess", "proj")}
        for k in range(K):
            y, w = np.ascontiguousarray(Y[k]), np.ascontiguousarray(W[k])
            res["lcost"].append(costgrd.lcost(y, D, om, pL, pU, inds))
            res["lgrad"].append(costgrd.lgrad(y, D, om, pL, pU, inds))
            res["lhess"].append(costgrd.lhess(y, w, D, om, pL, pU, inds))
            res["jcost"].append(costgrd.jcost(y, D, jinds))
            res["jgrad"].append(costgrd.jgrad(y, D, jinds))
            res["jhess"].append(costgrd.jhess(y, w, D, jinds))
            res["proj"].append(PSDFixedRank.proj(y, w))
        out[name + "_Y"] = Y
        out[name + "_W"] = W
        for k, v in res.items():
            out[name + "_" + k] = np.array(v)
        for k, v in (("D_goal", D), ("omega", om), ("psi_L", pL), ("psi_U", pU)):
            out[name + "_" + k] = v
    np.savez_compressed(os.path.join(GOLDEN, "costgrd_vectors.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", nargs="*", default=["ur10", "kuka", "lwa4d", "lwa4p", "panda", "chain20"])
    ap.add_argument("--goals", type=int, default=6)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    problems = {}
    for name in args.robots:
        if name == "kuka_table":
            # BASELINE config 3 as the reference literally runs it: KUKA + table_environment()
            # (utils.py:179-191, experiments/riemannian_example.py:13-17); few goals -- ~20 s each
            from graphik.utils.utils import table_environment
            robot, graph = _loaders()["kuka"]()
            for idx, obs in enumerate(table_environment()):
                graph.add_spherical_obstacle(f"o{idx}", obs[0], obs[1])
            np.savez_compressed(os.path.join(GOLDEN, name + "_graph.npz"), **graph_static(graph))
            problems[name] = dump_goals(name, robot, graph, min(args.goals, 2), args.seed)
            continue
        if name.startswith("chain"):
            robot, graph, dh = random_dh_chain(int(name[5:]), args.seed)
            dump_model(name, robot, dh)
        else:
            robot, graph = _loaders()[name]()
            dump_model(name, robot)
        np.savez_compressed(os.path.join(GOLDEN, name + "_graph.npz"), **graph_static(graph))
        problems[name] = dump_goals(name, robot, graph, args.goals, args.seed)
    if all(k in problems for k in ("ur10", "kuka", "chain20")):
        dump_costgrd_vectors({k: v for k, v in problems.items() if k in ("ur10", "kuka", "chain20")})


if __name__ == "__main__":
    main()
