import sys, os, torch
sys.path.insert(0, '/root/repo')
from graphik_b200.engine import BatchIK
from graphik_b200.utils.roboturdf import load_model
robot, graph = load_model("ur10")
eng = BatchIK(graph)
B = 1 << 20
N, G = eng.plan.N, eng.plan.n_goal
Y = torch.randn(B, N, 3, dtype=torch.float64, device="cuda")
W = torch.randn(B, N, 3, dtype=torch.float64, device="cuda")
gd = torch.rand(B, G, dtype=torch.float64, device="cuda") + 0.5
for _ in range(3):
    eng.hessvec(Y, W, gd)
torch.cuda.synchronize()
