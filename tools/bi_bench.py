#!/usr/bin/env python
"""k_bounds_init: time per batch (CUDA events) for the BASELINE workloads and, with --profile, the per-phase cycle counts
of one goal printed by a -DGIK_BI_PROFILE variant of the library (tools/bi_bench.py --build-profile builds it in-tree so
that it travels to the GPU box).

    python tools/bi_bench.py --build-profile                 # here (no GPU needed)
    python tools/bi_bench.py ur10:65536 chain20:65536 kuka_table:2048 [--profile]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphik_b200 import _lib

ALT = os.path.join(_lib.LIBDIR, "alt_bi_profile.so")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if "--build-profile" in sys.argv:
    print(_lib.build(force=True, defines=("GIK_BI_PROFILE",), out=ALT))
    sys.exit(0)
if "--profile" in sys.argv:
    _lib.LIBPATH = ALT
    _lib.needs_build = lambda: False
for a in sys.argv[1:]:
    if a.startswith("--lib="):       # an A/B variant built by hand into graphik_b200/lib/
        _lib.LIBPATH = os.path.join(_lib.LIBDIR, a[6:])
        _lib.needs_build = lambda: False
import json
import torch
from bench import goals_for, load_workload
from graphik_b200.engine import BatchIK

for spec in args:
    name, B = spec.split(":")
    B = int(B)
    robot, graph = load_workload(name)
    eng = BatchIK(graph)
    _, T = goals_for(robot, min(B, 4096), seed=1000)
    T = torch.as_tensor(T, device="cuda")
    if B > T.shape[0]:
        T = T.repeat((B + T.shape[0] - 1) // T.shape[0], 1, 1)[:B]
    g2 = eng.goal_distances(T)
    if "--profile" in sys.argv:
        print("==", name, flush=True)
        eng.initialization(g2[:1])
        torch.cuda.synchronize()
        continue
    for _ in range(2):
        eng.initialization(g2)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 3
    ev[0].record()
    for _ in range(reps):
        eng.initialization(g2)
    ev[1].record()
    torch.cuda.synchronize()
    print(json.dumps({"workload": name, "goals": B, "nodes": graph.n_nodes if hasattr(graph, "n_nodes") else None,
                      "bounds_init_ms": ev[0].elapsed_time(ev[1]) / reps}), flush=True)
