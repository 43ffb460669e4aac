#!/usr/bin/env python
"""Per-phase cycle counts of k_bounds_init (Floyd-Warshall, lower bounds, Gram, the three eigenproblems, scatter
matrix) for one goal, printed by the kernel itself.  Needs a library built with -DGIK_BI_PROFILE:

    cd graphik_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --shared \
        -cudart static -DGIK_BI_PROFILE -o ../lib/alt_bi_profile.so *.cu
    GRAPHIK_B200_LIB=$PWD/graphik_b200/lib/alt_bi_profile.so python tools/bi_profile.py ur10 chain20 kuka_table
"""
import sys, torch
sys.path.insert(0, '.')
from bench import goals_for, load_workload
from graphik_b200.engine import BatchIK
for name in sys.argv[1:]:
    robot, graph = load_workload(name)
    eng = BatchIK(graph)
    _, T = goals_for(robot, 4, seed=1000)
    g2 = eng.goal_distances(torch.as_tensor(T, device='cuda'))
    print("==", name, flush=True)
    Y0 = eng.initialization(g2[:1])
    torch.cuda.synchronize()
