"""Golden fixture for CIDGIK's problem construction, made by RUNNING THE UNMODIFIED REFERENCE in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference.  graphik/solvers/sdp_snl.py imports cvxpy (and, through
sdp_formulations.py, mosek) at module level; neither can be installed here, so empty stand-in modules are registered
first -- the functions called below (distance_constraints_graph :270-314, distance_range_constraints :356-398) are
numpy / networkx only and never touch them.  No SDP is solved: the reference's solver is MOSEK.

Output: tests/golden/cidgik_constraints.npz -- per robot and seeded goal: T_goal, the anchors solve_with_cidgik sets
(convex_iteration.py:284-289), the variable order of the reference's clique, and its constraint matrices A[m,N,N],
right-hand sides b[m] and node pairs; the number of inequality constraints (always 0).

Usage:  python oracle/gen_golden_cidgik.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from ref_runner import load_reference  # noqa: E402


def _stand_ins():
    class _Any:
        def __getattr__(self, k):
            return _Any()

    for name in ("cvxpy", "mosek", "progress", "progress.bar"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    cp = sys.modules["cvxpy"]
    cp.MOSEK, cp.Problem, cp.Variable, cp.Parameter = "MOSEK", object, object, object
    cp.error = types.SimpleNamespace(SolverError=Exception)
    sys.modules["progress.bar"].ShadyBar = object
    ms = sys.modules["mosek"]
    ms.iparam, ms.dparam, ms.scalingtype, ms.solveform, ms.onoffkey = _Any(), _Any(), _Any(), _Any(), _Any()


def main():
    load_reference(with_costgrd=False)
    _stand_ins()
    import networkx as nx
    from graphik.solvers import sdp_snl
    from graphik.utils import roboturdf as ru
    from graphik.utils.constants import POS

    out = {}
    for name, loader in (("ur10", ru.load_ur10), ("kuka", ru.load_kuka), ("lwa4d", ru.load_schunk_lwa4d)):
        robot, graph = loader()
        n = robot.n
        np.random.seed(0)
        for g in range(2):
            q = robot.random_configuration()
            T = robot.pose(q, "p%d" % n)
            anchors = {"p0": graph.nodes["p0"][POS], "q0": graph.nodes["q0"][POS], "p%d" % n: T.trans,
                       "q%d" % n: T.trans + T.rot.as_matrix()[:, 2]}              # convex_iteration.py:284-289
            G = nx.DiGraph(graph)                                                 # :176-180
            G.remove_node("x")
            G.remove_node("y")
            ccd = sdp_snl.distance_constraints_graph(G, anchors, False, ee_cost=False, angle_limits=True)
            assert len(ccd) == 1
            (clique, (A, b, mapping, augmented)), = ccd.items()
            assert augmented
            ineq = sdp_snl.distance_range_constraints(G, ccd, anchors)
            names = sorted((k for k in mapping if isinstance(k, str)), key=lambda k: mapping[k])
            pairs = sorted((k for k in mapping if not isinstance(k, str)), key=lambda k: mapping[k])
            key = "%s_%d_" % (name, g)
            out[key + "T_goal"] = T.as_matrix()
            out[key + "anchors"] = np.stack([anchors[u] for u in ("p0", "q0", "p%d" % n, "q%d" % n)])
            out[key + "order"] = np.array(names)
            out[key + "pairs"] = np.array([sorted(p) for p in pairs])
            out[key + "A"] = np.array(A)
            out[key + "b"] = np.array(b)
            out[key + "n_inequalities"] = np.array(sum(len(v) for v in ineq.values()))
    # The inequality form.  The reference's own graphs never carry a robot -- obstacle edge, so
    # distance_range_constraints returns nothing for them; the function that WOULD build the row,
    # anchor_inequality_constraint (sdp_snl.py:586-618), is called here directly for every free joint point of the
    # UR10 against one sphere (lower bound = radius), on a graph that holds the sphere as the reference adds it.
    robot, graph = ru.load_ur10()
    n = robot.n
    centre, radius = np.array([0.3, 0.3, 0.2]), 0.3
    graph.add_spherical_obstacle("o0", centre, radius)
    np.random.seed(1)
    T = robot.pose(robot.random_configuration(), "p%d" % n)
    anchors = {"p0": graph.nodes["p0"][POS], "q0": graph.nodes["q0"][POS], "p%d" % n: T.trans,
               "q%d" % n: T.trans + T.rot.as_matrix()[:, 2], "o0": graph.nodes["o0"][POS]}   # :186-189: POS -> anchor
    G = nx.DiGraph(graph)
    G.remove_node("x")
    G.remove_node("y")
    ccd = sdp_snl.distance_constraints_graph(G, anchors, False, ee_cost=False, angle_limits=True)
    (clique, (A, b, mapping, augmented)), = ccd.items()
    assert len(sdp_snl.distance_range_constraints(G, ccd, anchors)) == 0       # as shipped: nothing
    names = sorted((k for k in mapping if isinstance(k, str)), key=lambda k: mapping[k])
    rows, rhs, nodes = [], [], []
    for i in range(1, n):
        _, (Ai, bi) = sdp_snl.anchor_inequality_constraint(ccd, frozenset(("p%d" % i, "o0")), radius, anchors, False)
        rows.append(Ai)
        rhs.append(bi)
        nodes.append("p%d" % i)
    out["ineq_T_goal"] = T.as_matrix()
    out["ineq_order"] = np.array(names)
    out["ineq_nodes"] = np.array(nodes)
    out["ineq_A"] = np.array(rows)                 # <A, Z> <= b
    out["ineq_b"] = np.array(rhs)
    out["ineq_centre"], out["ineq_radius"] = centre, np.array(radius)
    out["ineq_n_equalities"] = np.array(len(A))    # the sphere adds no equality (anchor -- anchor pairs are skipped)
    path = os.path.join(ROOT, "tests", "golden", "cidgik_constraints.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_A")})


if __name__ == "__main__":
    main()
