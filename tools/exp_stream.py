#!/usr/bin/env python
"""A/B harness for the trust-region kernels and the pipelined stream (one process per library variant).

  python tools/exp_stream.py [--lib path.so] [--robot ur10] [--what steady,single,stream]

steady : B = 65536 goals, maxiter = 300 (stragglers bounded): tCG iterations per second in the throughput regime
single : one 4096-goal batch through BatchIK.solve (latency of a synchronous call)
stream : K batches of 4096 through IKStream for a grid of (slots, inner_budget)
Prints one JSON line per measurement.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--robot", default="ur10")
    ap.add_argument("--what", default="steady,single,stream")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--slots", default="1,2,3,4")
    ap.add_argument("--budgets", default="4096,8192,12288,16384,32768")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    from graphik_b200 import _lib
    if args.lib:
        _lib.LIBPATH = os.path.abspath(args.lib)
        _lib.needs_build = lambda: False
    import torch
    from bench import goals_for, load_workload
    from graphik_b200.engine import BatchIK
    from graphik_b200.pipeline import IKStream
    robot, graph = load_workload(args.robot)
    dev = torch.device("cuda", 0)
    tag = args.tag or (os.path.basename(args.lib) if args.lib else "default")

    def emit(**kw):
        kw.update(lib=tag, robot=args.robot)
        print(json.dumps(kw), flush=True)

    what = args.what.split(",")
    for kern in [k for k in ("latency", "throughput") if "steady" in what or "steady_" + k in what]:
        eng = BatchIK(graph, params={"maxiter": 300, "kernel": kern}, device=dev)
        T = torch.as_tensor(goals_for(robot, 65536, seed=5)[1], device=dev)
        g2 = eng.goal_distances(T)
        Y0 = eng.initialization(g2)
        torch.cuda.synchronize()
        best = None
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.solve_points(g2, Y0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        inner = float(out["n_inner"].sum())
        emit(what="steady", kernel=kern, ms=best, inner=inner, giter_per_s=inner / best / 1e6,
             solves_per_s=65536 / best * 1e3)
    if "single" in what:
        eng = BatchIK(graph, device=dev)
        Ts = [torch.as_tensor(goals_for(robot, args.batch, seed=1000 + s)[1], device=dev) for s in range(4)]
        eng.solve(Ts[0], check=False)
        torch.cuda.synchronize()
        ms = []
        for T in Ts[1:]:
            t0 = time.perf_counter()
            out = eng.solve(T, check=False)
            torch.cuda.synchronize()
            ms.append(1e3 * (time.perf_counter() - t0))
        emit(what="single", ms=float(np.mean(ms)), solves_per_s=args.batch / np.mean(ms) * 1e3,
             inner_mean=float(out["n_inner"].float().mean()), inner_max=int(out["n_inner"].max()))
    if "stream" in what:
        eng = BatchIK(graph, device=dev)
        K = args.steps
        Ts = [torch.as_tensor(goals_for(robot, args.batch, seed=1000 + s)[1], device=dev) for s in range(K + 4)]
        for slots in [int(v) for v in args.slots.split(",")]:
            for budget in [int(v) for v in args.budgets.split(",")]:
                st = IKStream(eng, slots=slots, inner_budget=budget)
                for s in range(4):
                    st.submit(Ts[s])
                st.drain()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tk = [st.submit(Ts[4 + s]) for s in range(K)]
                t_sub = time.perf_counter() - t0
                st.drain()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                conv = float(np.mean([float((st.result(t)["status"] == 0).float().mean()) for t in tk]))
                info = st.stats()
                info.pop("slots"), info.pop("inner_budget")
                emit(what="stream", slots=slots, budget=budget, steps=K, ms_total=1e3 * dt, ms_submit=1e3 * t_sub,
                     solves_per_s=K * args.batch / dt, converged=conv, **info)
                del st


if __name__ == "__main__":
    main()
