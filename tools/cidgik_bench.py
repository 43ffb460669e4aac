"""BASELINE configs[4]: CIDGIK on UR10, 1024 goals per batch -- timing of solve_batch_with_cidgik (CUDA events),
success statistics, one JSON line.  `--reps 1` under ncu gives the launch list behind profiles/r2ae_*."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="ur10")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    import torch
    from graphik_b200.solvers.convex_iteration import solve_batch_with_cidgik
    from graphik_b200.utils.roboturdf import load_model
    robot, graph = load_model(a.robot)
    n = robot.n
    rng = np.random.RandomState(0)
    lb = np.array([robot.lb["p%d" % i] for i in range(1, n + 1)])
    ub = np.array([robot.ub["p%d" % i] for i in range(1, n + 1)])
    Q = lb + (ub - lb) * rng.rand(a.batch, n)
    T = robot.fk_all(Q)[:, n]
    Td = torch.as_tensor(T, dtype=torch.float64, device="cuda")
    for _ in range(a.warmup):
        out = solve_batch_with_cidgik(graph, Td)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = solve_batch_with_cidgik(graph, Td)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    q = out["q"].cpu().numpy()
    Tq = robot.fk_all(q)[:, n]
    pos = np.linalg.norm(Tq[:, :3, 3] - T[:, :3, 3], axis=1)
    rot = np.abs(Tq[:, :3, :3] - T[:, :3, :3]).max(axis=(1, 2))
    ms = float(np.median(times))
    print(json.dumps({"workload": "configs[4] CIDGIK %s, %d goals per batch" % (a.robot, a.batch),
                      "solves_per_s": a.batch / ms * 1e3, "ms_per_batch": ms, "ms_all": times,
                      "pose_reached_frac": float(np.mean((pos < 1e-2) & (rot < 1e-2))),
                      "median_pos_err": float(np.median(pos)),
                      "convex_iters_hist": np.bincount(out["n_iters"].cpu().numpy(), minlength=11).tolist(),
                      "sdp_iters_mean": float(out["sdp_iters"].float().mean()),
                      "feasible_hist": np.bincount(out["feasible"].cpu().numpy(), minlength=3).tolist(),
                      "launches": int(out["launches"]),
                      "programs_per_sdp_launch": out["n_active"].cpu().numpy().tolist(),
                      "slowest_program_iters_per_launch": out["sdp_iters_max"].cpu().numpy().tolist()}))


if __name__ == "__main__":
    main()
