"""Larger reference sample for end-state parity (joint angles / EDM residual): runs the UNMODIFIED
reference's solve_with_riemannian pipeline on N seeded goals per robot and stores only what the
end-state comparison needs.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

  tests/golden/<robot>_stats.npz : T_goal, Y_init, Y_sol, q_sol, f, gradnorm, iterations, pose_err
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from ref_runner import load_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", nargs="*", default=["ur10", "kuka"])
    ap.add_argument("--goals", type=int, default=64)
    ap.add_argument("--seed", type=int, default=123)
    args = ap.parse_args()
    load_reference()
    from gen_golden import _loaders, random_dh_chain, solve_traced
    for name in args.robots:
        if name in ("kuka_table", "kuka_table_intended"):      # BASELINE config 3 (~25 s per goal)
            from graphik.utils.utils import table_environment
            robot, graph = _loaders()["kuka"]()
            for idx, obs in enumerate(table_environment()):
                graph.add_spherical_obstacle(f"o{idx}", obs[0], obs[1])
                if name == "kuka_table_intended":
                    # what add_spherical_obstacle (graph_base.py:201-211) means to do and never does on revolute graphs
                    # (SURVEY App. C.1): the reference's own statements, applied by hand to the joint points p_1..p_n
                    for i in range(1, robot.n + 1):
                        graph.add_edge(f"p{i}", f"o{idx}")
                        graph[f"p{i}"][f"o{idx}"]["bounded"] = ["below"]
                        graph[f"p{i}"][f"o{idx}"]["lower_limit"] = obs[1]
                        graph[f"p{i}"][f"o{idx}"]["upper_limit"] = 100
        elif name.startswith("chain"):  # BASELINE config 4's robot: the same random-DH chain as tests/golden/chain20_*
            robot, graph, _ = random_dh_chain(int(name[5:]), 0)
        else:
            robot, graph = _loaders()[name]()
        n = robot.n
        np.random.seed(args.seed)
        keys = ("T_goal", "Y_init", "Y_sol", "q_sol", "f", "gradnorm", "iterations", "pose_err")
        rec = {k: [] for k in keys}
        for k in range(args.goals):
            q = robot.random_configuration()
            T = robot.pose(q, "p%d" % n)
            r = solve_traced(graph, T)
            T_sol = robot.pose({"p%d" % (i + 1): r["q_sol"][i] for i in range(n)}, "p%d" % n).as_matrix()
            rec["T_goal"].append(T.as_matrix())
            rec["pose_err"].append(np.linalg.norm(T_sol[:3, 3] - T.as_matrix()[:3, 3]))
            for key in ("Y_init", "Y_sol", "q_sol", "f", "gradnorm", "iterations"):
                rec[key].append(r[key])
            print(name, k, r["iterations"], "%.2e" % r["f"], "%.2e" % rec["pose_err"][-1], flush=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + "_stats.npz"),
                            **{k: np.array(v) for k, v in rec.items()})


if __name__ == "__main__":
    main()
