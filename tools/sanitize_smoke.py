#!/usr/bin/env python
"""Tiny invocations of every kernel of libgraphik_b200.so for compute-sanitizer (racecheck / memcheck):

  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py

Every trust-region kernel variant (k_rtr_fast throughput / latency, k_rtr_duo, k_rtr_fast2, k_rtr_cta, k_rtr), the
sliced solve with parking and resuming, k_bounds_init at N = 16 / 44 / 118 (shared memory only / with workspace),
the streaming cost kernels, goal distances, joints, FK, the limit check, the conjugate-gradient solve and the CIDGIK
kernels (SDP interior point, Fantope step) -- on a handful of goals with the
iteration counts cut down (a sanitizer slows a kernel by two orders of magnitude).
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from bench import goals_for, load_workload
    from graphik_b200 import _lib
    from graphik_b200.engine import BatchIK, _p
    from graphik_b200.pipeline import IKStream
    params = {"maxiter": 3, "maxinner": 12}
    done = []
    for name, kernels, B in (("ur10", ("latency", "throughput", "generic"), 6), ("kuka", ("latency",), 4),
                             ("chain20", ("latency", "generic"), 3), ("kuka_table", ("dense",), 2)):
        robot, graph = load_workload(name)
        _, T = goals_for(robot, B, seed=1)
        for kern in kernels:
            eng = BatchIK(graph, params=dict(params, kernel=kern))
            out = eng.solve(T, check=True)               # goal distances, bounds + init, solve, joints, fk, limits
            torch.cuda.synchronize()
            assert int((out["iterations"] == 3).sum()) == B
            # sliced: everything parks after one outer iteration, then drains (throughput variant + latency variant)
            st = IKStream(eng, slots=1, inner_budget=1, carry_capacity=16)
            tk = [st.submit(T), st.submit(T)]
            st.drain()
            for t in tk:
                r = st.result(t)
                for k in ("x", "f(x)", "iterations"):
                    assert (r[k] == out[k]).all(), (name, kern, k)
            done.append("%s/%s" % (name, kern))
        eng = BatchIK(graph)
        g2 = eng.goal_distances(T)
        lb, ub = eng.bounds(g2)
        Y0 = eng.init_from_bounds(lb, ub)
        f, g = eng.cost_grad(Y0, g2)
        eng.hessvec(Y0, g, g2)
        eng.proj(Y0, g)
        torch.cuda.synchronize()
    # a graph above 128 nodes: group kernel with eight nodes per lane, bound smoothing without register tiles
    from helpers import load_robot
    from graphik_b200.utils.utils import table_environment
    robot, graph = load_robot("kuka")
    for k, (c, r) in enumerate(table_environment(n_height=12, n_width=10)):
        graph.add_spherical_obstacle("o%d" % k, c, r)
    _, T = goals_for(robot, 2, seed=1)
    eng = BatchIK(graph, params=params)
    out = eng.solve(T, check=True)
    torch.cuda.synchronize()
    assert int((out["iterations"] == 3).sum()) == 2
    done.append("kuka+148 obstacles (N = %d)/generic" % graph.number_of_nodes())
    # conjugate-gradient solve (k_cg) and CIDGIK (k_sdp with a skipped program, k_fantope), cut down
    from graphik_b200.engine import make_opts
    from graphik_b200.solvers.convex_iteration import convex_iterate_batch
    for name in ("ur10", "chain20"):
        robot, graph = load_workload(name)
        _, T = goals_for(robot, 3, seed=2)
        eng = BatchIK(graph)
        g2 = eng.goal_distances(T)
        out = eng.solve_points(g2, eng.initialization(g2), opts=make_opts({"solver": "ConjugateGradient", "maxiter": 6}))
        torch.cuda.synchronize()
        assert int((out["iterations"] == 5).sum()) == 3
        done.append("%s/cg" % name)
    for name in ("ur10", "kuka"):
        robot, graph = load_workload(name)
        _, T = goals_for(robot, 3, seed=2)
        for fused in (True, False):          # gik_cidgik_solve / gik_sdp_solve + gik_fantope per convex iteration
            out = convex_iterate_batch(graph, T, max_iters=2, sdp_params={"maxiter": 4}, sdp_accept=float("inf"),
                                       fused=fused)
            torch.cuda.synchronize()
            assert int(out["n_iters"].max()) == 2
        done.append("%s/cidgik fused + per-iteration" % name)
    from helpers import load_robot as _lr
    robot, graph = _lr("ur10", graph_params={"obstacle_semantics": "intended"})
    graph.add_spherical_obstacle("o0", np.array([0.3, 0.3, 0.2]), 0.3)
    _, T = goals_for(robot, 3, seed=2)
    for fused in (True, False):
        convex_iterate_batch(graph, T, max_iters=2, sdp_params={"maxiter": 4}, sdp_accept=float("inf"), fused=fused)
        torch.cuda.synchronize()
    done.append("ur10+sphere/cidgik with inequalities fused + per-iteration")
    print("sanitize smoke ok:", " ".join(done))


if __name__ == "__main__":
    main()
