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
