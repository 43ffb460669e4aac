#!/usr/bin/env python
"""Per-workload numbers of the trust-region kernel: solves/s, iteration counts, cost of one tCG
inner iteration (SM-time), streaming-roofline fraction (72N/outer + 240N/inner bytes)
and FP64 rate (38 flop per term + 60N + 60 per inner iteration).

    python tools/profile_workloads.py ur10:4096:latency ur10:65536:throughput chain20:8192 kuka_table:296

Each spec is robot:batch[:kernel[:repeats[:maxiter]]]; with a small maxiter every problem does about the
same work, so solve time / max inner iterations is the latency of one inner iteration of one problem.  One JSON line per spec (not a bench value: no e2e, one launch).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    for a in sys.argv[1:]:
        if a.startswith("--lib="):       # an A/B variant built by hand into graphik_b200/lib/
            from graphik_b200 import _lib
            _lib.LIBPATH = os.path.join(_lib.LIBDIR, a[6:])
            _lib.needs_build = lambda: False
    import torch
    from bench import goals_for, load_workload, measured_peak_hbm
    from graphik_b200.engine import BatchIK

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak, _ = measured_peak_hbm()
    for spec in sys.argv[1:]:
        if spec.startswith("--"):
            continue
        parts = spec.split(":")
        robot_name, B = parts[0], int(parts[1])
        kernel = parts[2] if len(parts) > 2 and parts[2] else "auto"
        reps = int(parts[3]) if len(parts) > 3 and parts[3] else 1
        maxiter = int(parts[4]) if len(parts) > 4 else None
        robot, graph = load_workload(robot_name)
        params = {"kernel": kernel}
        if maxiter:
            params["maxiter"] = maxiter
        eng = BatchIK(graph, params=params, device=dev)
        N = graph.number_of_nodes()
        _, T = goals_for(robot, B, seed=1000)
        T = torch.as_tensor(T, device=dev)
        g2 = eng.goal_distances(T)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        Y0 = eng.initialization(g2[: min(B, 64)])                   # warm (module load)
        torch.cuda.synchronize()
        e0.record()
        Y0 = eng.initialization(g2)
        e1.record()
        torch.cuda.synchronize()
        init_ms = e0.elapsed_time(e1)
        eng.solve_points(g2[: min(B, 64)], Y0[: min(B, 64)])       # warm (module load, smem attribute)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            e1.record()
            out = eng.solve_points(g2, Y0)
            e2.record()
            torch.cuda.synchronize()
            ms.append(e1.elapsed_time(e2))
        t = float(np.min(ms)) * 1e-3
        it = out["iterations"].double()
        inn = out["n_inner"].double()
        f = out["f(x)"]
        n_terms = eng.plan.n_terms
        alg_bytes = float((72.0 * N * it + 240.0 * N * inn).sum())
        flops = float(inn.sum()) * (38.0 * n_terms + 60.0 * N + 60.0)
        line = {
            "robot": robot_name, "B": B, "kernel": kernel, "N": N, "terms": n_terms,
            "init_ms": init_ms, "solve_ms": [round(m, 3) for m in ms],
            "solves_per_s": B / t,
            "mean_outer": float(it.mean()), "mean_inner": float(inn.mean()), "max_inner": float(inn.max()),
            "converged_frac": float((out["status"] == 0).double().mean()),
            "median_f": float(f.median()),
            "inner_iters_per_s": float(inn.sum()) / t,
            "sm_ns_per_inner_iter": t * sms / float(inn.sum()) * 1e9,
            "maxiter": maxiter, "us_per_inner_iter_of_slowest_problem": t / float(inn.max()) * 1e6,
            "stream_GBs": alg_bytes / t / 1e9, "stream_frac": alg_bytes / t / 1e9 / peak,
            "fp64_TFLOPs": flops / t / 1e12,
        }
        print(json.dumps(line), flush=True)
        del eng


if __name__ == "__main__":
    main()
