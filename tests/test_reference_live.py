"""Runs only where the reference tree is mounted (the build container): the oracle against the
LIVE unmodified reference on fresh random inputs, beyond the committed goldens."""
import os

import numpy as np
import pytest

from helpers import ROOT

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/graphik"),
                                reason="reference tree not mounted (GPU box)")


def test_oracle_matches_live_reference():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ref_runner import load_reference
    load_reference()
    from graphik.solvers import costgrd
    from graphik.utils.dgp import adjacency_matrix_from_graph, bound_smoothing, distance_matrix_from_graph
    from graphik.utils.manifolds.fixed_rank_psd_sym import PSDFixedRank
    from graphik.utils.roboturdf import load_kuka
    from oracle import oracle as orc
    from graphik_b200.utils.roboturdf import load_kuka as my_load_kuka
    robot, graph = load_kuka()
    my_robot, my_graph = my_load_kuka()
    np.random.seed(5)
    for _ in range(3):
        q = robot.random_configuration()
        T = robot.pose(q, "p7")
        G = graph.from_pose(T)
        D, om = distance_matrix_from_graph(G), adjacency_matrix_from_graph(G)
        pL, pU = graph.distance_bound_matrices()
        lb, ub = bound_smoothing(G)
        myG = my_graph.from_pose(T.as_matrix())
        lb2, ub2 = orc.bound_smoothing(myG.edge, myG.lower, myG.upper)
        assert np.max(np.abs(lb - lb2)) <= 1e-13 and np.max(np.abs(ub - ub2)) <= 1e-13
        P = orc.Problem(D, om, pL, pU)
        inds = orc.limit_inds(om, pL, pU)
        Y, W = np.random.randn(18, 3), np.random.randn(18, 3)
        assert costgrd.lcost(Y, D, om, pL, pU, inds) == P.cost(Y)
        assert np.array_equal(costgrd.lgrad(Y, D, om, pL, pU, inds), P.grad(Y))
        assert np.array_equal(costgrd.lhess(Y, W, D, om, pL, pU, inds), P.hess(Y, W))
        assert np.max(np.abs(PSDFixedRank.proj(Y, W) - orc.proj(Y, W))) <= 1e-13
