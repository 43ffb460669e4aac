"""BASELINE.json's full size (UR10, 4096 goals in one batch) checked through size-independent properties:
the oracle cannot solve 4096 problems inside a unit test, invariants can be checked on all of them."""
import numpy as np
import pytest

from helpers import load_robot, random_goals

pytestmark = pytest.mark.gpu
B = 4096


@pytest.fixture(scope="module")
def solved():
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot("ur10")
    eng = BatchIK(graph)
    Q, T = random_goals(robot, B, seed=2024)
    out = eng.solve(T, check=True)
    return robot, graph, eng, T, out


def test_status_cost_and_success(solved):
    robot, graph, eng, T, out = solved
    st = out["status"].cpu().numpy()
    f = out["f(x)"].cpu().numpy()
    gn = out["gradnorm"].cpu().numpy()
    it = out["iterations"].cpu().numpy()
    assert set(np.unique(st)) <= {0, 1}
    assert np.mean(st == 0) >= 0.99                      # same stopping rule as the reference
    assert np.all(gn[st == 0] < 5e-10) and np.all(it[st == 1] == 3000) and np.all(it[st == 0] < 3000)
    # reported cost == the cost kernel (== lcost, costs.py:79-93) on the returned points
    f2, g2 = eng.cost_grad(out["x"], out["goal_d2"])
    f2 = f2.cpu().numpy()
    assert np.max(np.abs(f2 - f) / np.maximum(1e-300 + np.abs(f), 1e-18)) < 1e-6 or np.allclose(f2, f, rtol=1e-9, atol=1e-24)
    assert np.allclose(np.linalg.norm(g2.cpu().numpy().reshape(B, -1), axis=1), gn, rtol=1e-6, atol=1e-16)
    # pose success rate in the reference's range (0.95-1.00 below 1e-2 m on 12-20 goals; 0.92-0.94 on
    # larger samples, tests/golden/ur10_stats.npz)
    pos = out["pos_err"].cpu().numpy()
    assert 0.88 <= np.mean(pos < 1e-2) <= 1.0
    assert np.median(f) < 1e-14


def test_equality_edges_are_met(solved):
    """EDM residual: for converged problems every equality edge holds to sqrt(f)."""
    robot, graph, eng, T, out = solved
    a = eng.plan._a
    x = out["x"].cpu().numpy()
    f = out["f(x)"].cpu().numpy()
    gd = out["goal_d2"].cpu().numpy()
    eq = a["term_kind"] == 0
    i, j, tgt, gs = a["term_i"][eq], a["term_j"][eq], a["term_target"][eq], a["term_goal"][eq]
    d2 = np.sum((x[:, i] - x[:, j]) ** 2, axis=-1)
    target = np.where(gs[None, :] >= 0, gd[:, np.maximum(gs, 0)], tgt[None, :])
    worst = np.max(np.abs(d2 - target), axis=1)
    assert np.all(worst <= np.sqrt(f) + 1e-15)


def test_rigid_motion_invariance(solved):
    """The cost only sees distances: f(Y Q + t) = f(Y), g(Y Q + t) = g(Y) Q for any orthogonal Q."""
    robot, graph, eng, T, out = solved
    import torch
    rng = np.random.default_rng(1)
    Qm = torch.as_tensor(np.linalg.qr(rng.normal(size=(3, 3)))[0], device=out["x"].device)
    Y = torch.randn(B, 16, 3, dtype=torch.float64, device=out["x"].device)
    f1, g1 = eng.cost_grad(Y, out["goal_d2"])
    f2, g2 = eng.cost_grad((Y @ Qm + 0.37).contiguous(), out["goal_d2"])
    assert torch.allclose(f1, f2, rtol=1e-11, atol=0)
    assert torch.allclose(g1 @ Qm, g2, rtol=1e-9, atol=1e-9)


def test_resolving_from_the_solution_is_idempotent(solved):
    robot, graph, eng, T, out = solved
    conv = (out["status"] == 0)
    again = eng.solve_points(out["goal_d2"], out["x"])
    it = again["iterations"][conv].cpu().numpy()
    # already below mingradnorm: the loop body runs once (as in the reference, the test is at its end),
    # takes at most a tiny Newton step and leaves; the cost never increases (monotone trust region)
    assert np.all(it <= 2)
    # (the rho regularisation of trust_region.py:293 tolerates an increase of ~2e-13)
    assert bool(((again["f(x)"] <= out["f(x)"] + 1e-12)[conv]).all())
    assert bool((again["status"][conv] == 0).all())
    assert float((again["x"] - out["x"])[conv].abs().max()) < 1e-2   # one small Newton step at most


def test_any_subset_gives_identical_bits(solved):
    robot, graph, eng, T, out = solved
    sub = eng.solve(T[1000:1128], check=False)
    for key in ("x", "f(x)", "iterations", "q"):
        assert np.array_equal(sub[key].cpu().numpy(), out[key][1000:1128].cpu().numpy()), key


# ------------------------------------------------------------------ BASELINE configs[3] and configs[2] at size

def _edge_residuals(eng, out):
    a = eng.plan._a
    x = out["x"].cpu().numpy()
    gd = out["goal_d2"].cpu().numpy()
    eq = a["term_kind"] == 0
    i, j, tgt, gs = a["term_i"][eq], a["term_j"][eq], a["term_target"][eq], a["term_goal"][eq]
    d2 = np.sum((x[:, i] - x[:, j]) ** 2, axis=-1)
    target = np.where(gs[None, :] >= 0, gd[:, np.maximum(gs, 0)], tgt[None, :])
    return np.max(np.abs(d2 - target), axis=1)


@pytest.fixture(scope="module")
def solved_chain20():
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot("chain20")
    eng = BatchIK(graph)
    Q, T = random_goals(robot, 8192, seed=77)
    return robot, graph, eng, T, eng.solve(T, check=True)


def test_chain20_8192_goals_properties(solved_chain20):
    """The 20-DOF chain of BASELINE configs[3] (N = 44, two nodes per lane) on 8192 goals in one batch: stopping rule,
    reported cost = cost kernel on the returned points, equality edges met to sqrt(f), any subset solved alone returns
    the same bits."""
    robot, graph, eng, T, out = solved_chain20
    st, f = out["status"].cpu().numpy(), out["f(x)"].cpu().numpy()
    gn, it = out["gradnorm"].cpu().numpy(), out["iterations"].cpu().numpy()
    assert set(np.unique(st)) <= {0, 1}
    assert np.mean(st == 0) >= 0.97
    assert np.all(gn[st == 0] < 5e-10) and np.all(it[st == 1] == 3000) and np.all(it[st == 0] < 3000)
    f2, g2 = eng.cost_grad(out["x"], out["goal_d2"])
    assert np.allclose(f2.cpu().numpy(), f, rtol=1e-9, atol=1e-24)
    assert np.allclose(np.linalg.norm(g2.cpu().numpy().reshape(len(f), -1), axis=1), gn, rtol=1e-6, atol=1e-16)
    assert np.all(_edge_residuals(eng, out)[st == 0] <= np.sqrt(f[st == 0]) + 1e-15)
    assert np.median(f) < 1e-13
    # (no pose assertion: on this random-DH chain the reference's own joint_variables returns angles whose pose is
    # 0.4-4 m off although its EDM residual is 1e-21 -- tests/golden/chain20_stats.npz -- a reference quirk that the
    # joint-recovery kernel reproduces; the angle recovery itself is pinned in test_gpu_kernels.py)
    sub = eng.solve(T[4000:4096], check=False)
    for key in ("x", "f(x)", "iterations", "q"):
        assert np.array_equal(sub[key].cpu().numpy(), out[key][4000:4096].cpu().numpy()), key


def test_kuka_table_512_goals_properties():
    """KUKA IIWA + table_environment() of BASELINE configs[2] (N = 118, dense kernel, bound smoothing with its third
    matrix in the caller's workspace) on 512 goals: stopping rule, cost consistency, equality edges, bound smoothing
    contains the realised distances of the solved configuration, sub-batch bit-identity."""
    from helpers import load_kuka_table
    from graphik_b200.engine import BatchIK
    robot, graph = load_kuka_table()
    eng = BatchIK(graph)
    Q, T = random_goals(robot, 512, seed=5)
    out = eng.solve(T, check=True)
    st, f = out["status"].cpu().numpy(), out["f(x)"].cpu().numpy()
    gn, it = out["gradnorm"].cpu().numpy(), out["iterations"].cpu().numpy()
    assert set(np.unique(st)) <= {0, 1}
    assert np.mean(st == 0) >= 0.8          # the reference leaves ~10 % of these goals at maxiter too
    assert np.all(gn[st == 0] < 5e-10) and np.all(it[st == 1] == 3000) and np.all(it[st == 0] < 3000)
    f2, g2 = eng.cost_grad(out["x"], out["goal_d2"])
    assert np.allclose(f2.cpu().numpy(), f, rtol=1e-9, atol=1e-22)
    assert np.allclose(np.linalg.norm(g2.cpu().numpy().reshape(len(f), -1), axis=1), gn, rtol=1e-4, atol=1e-12)   # 354 components of rounding noise ~1e-14 each
    assert np.all(_edge_residuals(eng, out)[st == 0] <= np.sqrt(f[st == 0]) + 1e-14)
    assert np.median(f) < 1e-13
    # bound smoothing at size: the distances of the configuration that generated the goal lie inside the bounds
    # (the property of reference tests/test_bound_smoothing.py:99-117), 64 goals
    lb, ub = eng.bounds(out["goal_d2"][:64])
    lb, ub = lb.cpu().numpy(), ub.cpu().numpy()
    for k in range(64):
        D = np.sqrt(graph.distance_matrix_from_joints(Q[k]))
        assert np.all(D < ub[k] + 1e-6) and np.all(lb[k] - 1e-6 < D)
    sub = eng.solve(T[100:116], check=False)
    for key in ("x", "f(x)", "iterations", "q"):
        assert np.array_equal(sub[key].cpu().numpy(), out[key][100:116].cpu().numpy()), key
