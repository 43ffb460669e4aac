#!/usr/bin/env python
"""Streaming forms of the costgrd drop-ins (gik_cost_grad / gik_hessvec / gik_proj) against the HBM
roofline: B problems, state in HBM, one pass.  Algorithmic bytes per problem: cost+grad 24N + 8 n_goal in,
24N + 8 out; Hess-vec 48N + 8 n_goal in, 24N out; proj 48N in, 24N out."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphik_b200 import _lib  # noqa: E402
for _a in sys.argv[1:]:
    if _a.startswith("--lib="):       # an A/B variant built by hand into graphik_b200/lib/
        _lib.LIBPATH = os.path.join(_lib.LIBDIR, _a[6:])
        _lib.needs_build = lambda: False
from graphik_b200.engine import BatchIK  # noqa: E402
from graphik_b200.utils.roboturdf import load_model  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peak = 6454.0
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    robots = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name, B in (("ur10", 1 << 21), ("kuka", 1 << 21), ("chain20", 1 << 20)):
        if robots and name not in robots:
            continue
        robot, graph = load_model(name)
        eng = BatchIK(graph)
        N, G = eng.plan.N, eng.plan.n_goal
        Y = torch.randn(B, N, 3, dtype=torch.float64, device="cuda")   # 24 N B bytes >> 126 MB L2
        W = torch.randn(B, N, 3, dtype=torch.float64, device="cuda")
        gd = torch.rand(B, G, dtype=torch.float64, device="cuda") + 0.5
        rows = {}
        ms = timed(lambda: eng.cost_grad(Y, gd))
        rows["cost_grad"] = (B * (48 * N + 8 * G + 8)) / (ms * 1e-3) / 1e9
        ms = timed(lambda: eng.hessvec(Y, W, gd))
        rows["hessvec"] = (B * (72 * N + 8 * G)) / (ms * 1e-3) / 1e9
        ms = timed(lambda: eng.proj(Y, W))
        rows["proj"] = (B * 72 * N) / (ms * 1e-3) / 1e9
        print(json.dumps({"robot": name, "B": B, "N": N, "peak_GBs": peak,
                          **{k: {"GBs": round(v, 1), "frac": round(v / peak, 3)} for k, v in rows.items()}}))


if __name__ == "__main__":
    main()
