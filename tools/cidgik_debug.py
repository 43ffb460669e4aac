import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import load_robot, random_goals
from oracle import cidgik as cg
from graphik_b200.solvers.convex_iteration import CidgikPlan, sdp_solve_batch, make_sdp_opts, solve_batch_with_cidgik
name = sys.argv[1] if len(sys.argv) > 1 else "ur10"
robot, graph = load_robot(name)
plan = CidgikPlan(graph)
Q, T = random_goals(robot, 8, 11)
anchors, W, b, V = plan.assemble(T, device="cuda")
C = torch.matmul(V.transpose(1, 2), V).contiguous()
for mi in (50,):
    out = sdp_solve_batch(C, W, b, opts=make_sdp_opts({"maxiter": mi}))
    torch.cuda.synchronize()
    for k in range(8):
        A = np.einsum("ki,kj->kij", W[k].cpu().numpy(), W[k].cpu().numpy())
        tr = []
        ref = cg.solve_sdp(C[k].cpu().numpy(), A, b[k].cpu().numpy(), maxiter=mi, trace=tr)
        print("maxiter", mi, "k", k, "gpu: st %d it %d obj %.10e resid %.3e | ref: st %d it %d obj %.10e resid %.3e | dX %.2e" % (
            int(out["status"][k]), int(out["iters"][k]), float(out["obj"][k]), float(out["resid"][k]),
            ref["status"], ref["iters"], ref["obj"], ref["resid"], np.abs(out["X"][k].cpu().numpy() - ref["X"]).max()))
n = robot.n
Q, T = random_goals(robot, 1024, 21)
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    out = solve_batch_with_cidgik(graph, T, as_numpy=True)
    torch.cuda.synchronize(); dt = time.time() - t
Tq = robot.fk_all(out["q"])[:, n]
pos = np.linalg.norm(Tq[:, :3, 3] - T[:, :3, 3], axis=1)
print("1024 goals: %.3f s (%.0f solves/s)" % (dt, 1024 / dt), "feasible", np.bincount(out["feasible"]), "n_iters", np.bincount(out["n_iters"]),
      "pos<1e-2: %.3f" % np.mean(pos < 1e-2), "median pos %.1e" % np.median(pos), "sdp iters mean %.1f" % out["sdp_iters"].mean())
