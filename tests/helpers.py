"""Shared test helpers: golden loading, oracle problems, engines."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

ROBOTS = ["ur10", "kuka", "lwa4d", "lwa4p", "panda", "chain20"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def load_robot(name, **kw):
    from graphik_b200.utils.roboturdf import load_model
    return load_model(name, **kw)


def matrices_for_goal(graph, T_goal):
    """D_goal, omega, psi_L, psi_U exactly as solve_with_riemannian builds them
    (riemannian_solver.py:221-224, 189) through the host graph model."""
    G = graph.from_pose(T_goal)
    D = np.where(np.isnan(G.dist), np.where(G.edge, 1.0, 0.0), G.dist) ** 2
    omega = (~np.isnan(G.dist)).astype(float)
    psi_L, psi_U = graph.distance_bound_matrices()
    return G, D, omega, psi_L, psi_U


def random_goals(robot, B, seed):
    """Reachable goals: q ~ U(lb, ub), T = FK(q) (README usage / SURVEY 8d)."""
    rng = np.random.RandomState(seed)
    n = robot.n
    lb = np.array([robot.lb["p%d" % i] for i in range(1, n + 1)])
    ub = np.array([robot.ub["p%d" % i] for i in range(1, n + 1)])
    Q = lb + (ub - lb) * rng.rand(B, n)
    T = robot.fk_all(Q)[:, n]
    return Q, T


def align_columns(Y, Yref):
    """Flip column signs of Y to match Yref (eigenvector sign freedom)."""
    s = np.sign(np.sum(Y * Yref, axis=-2, keepdims=True))
    s[s == 0] = 1
    return Y * s


def load_kuka_table(**kw):
    """BASELINE config 3 as the reference runs it: KUKA IIWA + table_environment() obstacles
    (reference semantics: obstacles are anchors only)."""
    from graphik_b200.utils.utils import table_environment
    robot, graph = load_robot("kuka", **kw)
    for k, (c, r) in enumerate(table_environment()):
        graph.add_spherical_obstacle("o%d" % k, c, r)
    return robot, graph
