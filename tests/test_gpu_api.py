"""The reference-facing Python API on the GPU: same names, arguments and return shapes as
graphik.solvers.riemannian_solver / graphik.solvers.costgrd / graphik.utils.dgp."""
import numpy as np
import pytest

from helpers import golden, load_robot, matrices_for_goal

pytestmark = pytest.mark.gpu


def test_solve_with_riemannian_readme_usage():
    """README.md:30-46 / experiments/riemannian_example.py, with the loader and solver swapped."""
    from graphik_b200.solvers.riemannian_solver import solve_with_riemannian
    from graphik_b200.utils.roboturdf import load_ur10
    robot, graph = load_ur10()
    np.random.seed(0)
    q_goal = robot.random_configuration()
    T_goal = robot.pose(q_goal, f"p{robot.n}")
    q_sol, points = solve_with_riemannian(graph, T_goal, use_jit=False)
    assert sorted(q_sol) == ["p%d" % i for i in range(1, 7)] and points.shape == (16, 3)
    T_sol = robot.pose(q_sol, "p6").as_matrix()
    assert np.linalg.norm(T_sol[:3, 3] - T_goal.as_matrix()[:3, 3]) < 1e-2
    # the misspelt keyword of the reference README is accepted too
    q2, _ = solve_with_riemannian(graph, T_goal, jit=True)
    assert q2 == q_sol   # deterministic


def test_riemannian_solver_solve_signature_and_log():
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver
    from graphik_b200.utils.dgp import (adjacency_matrix_from_graph, bound_smoothing,
                                         distance_matrix_from_graph)
    robot, graph = load_robot("ur10")
    g = golden("ur10_goals")
    G = graph.from_pose(g["T_goal"][0])
    D_goal, omega = distance_matrix_from_graph(G), adjacency_matrix_from_graph(G)
    assert np.array_equal(D_goal, g["D_goal"][0]) and np.array_equal(omega, g["omega"][0])
    lb, ub = bound_smoothing(G)
    assert np.max(np.abs(lb - g["lb"][0])) <= 1e-12 and np.max(np.abs(ub - g["ub"][0])) <= 1e-12
    solver = RiemannianSolver(graph)
    sol = solver.solve(D_goal, omega, use_limits=True, bounds=(lb, ub), jit=True)
    assert sorted(sol) == ["f(x)", "gradnorm", "iterations", "time", "x"]
    assert sol["f(x)"] < 1e-12 and sol["gradnorm"] < 5e-10 and sol["x"].shape == (16, 3)
    # injected initial point, no limits, output_log=False returns the points only
    Y = solver.solve(D_goal, omega, use_limits=False, Y_init=g["Y_init"][0], output_log=False)
    assert isinstance(Y, np.ndarray) and Y.shape == (16, 3)
    with pytest.raises(Exception):
        solver.solve(D_goal, omega)                      # neither bounds nor Y_init (:199-200)
    with pytest.raises(ValueError):                      # riemannian_solver.py:61-64
        RiemannianSolver(graph, {"solver": "SteepestDescent"})
    sol3 = RiemannianSolver(graph, {"maxiter": 5}).solve(D_goal, omega, use_limits=True, Y_init=g["Y_init"][0])
    assert sol3["iterations"] == 5


def test_costgrd_facade_matches_reference_vectors():
    from graphik_b200.solvers import costgrd
    from oracle import oracle as orc
    cv = golden("costgrd_vectors")
    name = "kuka"
    D, om, pL, pU = (cv[name + "_" + k] for k in ("D_goal", "omega", "psi_L", "psi_U"))
    inds, jinds = orc.limit_inds(om, pL, pU), orc.equality_inds(om)
    Y, W = cv[name + "_Y"][5], cv[name + "_W"][5]

    def close(a, b):
        return np.max(np.abs(np.asarray(a) - b)) <= 1e-12 * max(1.0, np.max(np.abs(b)))

    assert close(costgrd.lcost(Y, D, om, pL, pU, inds), cv[name + "_lcost"][5])
    assert close(costgrd.lgrad(Y, D, om, pL, pU, inds), cv[name + "_lgrad"][5])
    assert close(costgrd.lhess(Y, W, D, om, pL, pU, inds), cv[name + "_lhess"][5])
    f, g = costgrd.lcost_and_grad(Y, D, om, pL, pU, inds)
    assert close(f, cv[name + "_lcost"][5]) and close(g, cv[name + "_lgrad"][5])
    assert close(costgrd.jcost(Y, D, jinds), cv[name + "_jcost"][5])
    assert close(costgrd.jgrad(Y, D, jinds), cv[name + "_jgrad"][5])
    assert close(costgrd.jhess(Y, W, D, jinds), cv[name + "_jhess"][5])
    f, g = costgrd.jcost_and_grad(Y, D, jinds)
    assert close(f, cv[name + "_jcost"][5]) and close(g, cv[name + "_jgrad"][5])


def test_solve_batch_outputs_and_obstacle_anchors():
    """Batched entry point; KUKA with a few spherical obstacles (reference semantics: anchors only)."""
    from graphik_b200.solvers.riemannian_solver import solve_batch_with_riemannian
    from helpers import random_goals
    robot, graph = load_robot("kuka")
    for k, c in enumerate(([0.6, 0.1, 0.4], [-0.3, 0.5, 0.8], [0.2, -0.6, 0.3])):
        graph.add_spherical_obstacle("o%d" % k, np.array(c), 0.1)
    Q, T = random_goals(robot, 32, seed=9)
    out = solve_batch_with_riemannian(graph, T)
    assert out["q"].shape == (32, 7) and out["x"].shape == (32, 21, 3)
    assert np.mean(out["pos_err"] < 1e-2) >= 0.8
    assert np.median(out["f(x)"]) < 1e-12


def test_multi_gpu_sharding_is_exact():
    """Sharding the batch over two devices reproduces the single-device results bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from graphik_b200.engine import BatchIK
    from helpers import random_goals
    robot, graph = load_robot("ur10")
    Q, T = random_goals(robot, 64, seed=4)
    full = BatchIK(graph, device=0).solve(T, check=False)
    with torch.cuda.device(1):
        part = BatchIK(graph, device=1).solve(T[32:], check=False)
    assert np.array_equal(full["x"][32:].cpu().numpy(), part["x"].cpu().numpy())


def test_c_abi_error_codes_and_empty_batches():
    """Every entry point: B = 0 is a no-op, bad arguments / limits come back as negative codes with a
    message, nothing throws across the ABI."""
    import ctypes
    import torch
    from graphik_b200 import _lib
    from graphik_b200.engine import BatchIK
    from graphik_b200.plan import Plan
    robot, graph = load_robot("ur10")
    eng = BatchIK(graph)
    L, h = eng.lib, eng.plan.handle
    null = ctypes.c_void_p(0)
    assert L.gik_goal_distances(h, null, 0, null, null) == 0
    assert L.gik_cost_grad(h, null, null, 0, null, null, null) == 0
    assert L.gik_hessvec(h, null, null, null, 0, null, null) == 0
    assert L.gik_proj(16, null, null, 0, null, null) == 0
    assert L.gik_bounds(h, null, 0, null, null, null, null) == 0
    assert L.gik_init(h, null, null, 0, null, null, null) == 0
    assert L.gik_bounds_init(h, null, 0, null, null, null) == 0
    assert L.gik_joints(h, null, null, 0, null, null) == 0
    assert L.gik_fk(h, null, 0, null, null, null) == 0
    assert L.gik_check_limits(h, null, 1e-6, 0, null, null, null) == 0
    assert L.gik_rtr_solve_sliced(h, null, null, 0, None, null, null, null, null, null, null, 0, null, null, null,
                                  null, null) == 0
    assert L.gik_workspace_bytes(h) == 0 and L.gik_carry_bytes(h, 10) == 16 + 10 * 8 * (22 + 6 * 16)
    assert L.gik_rtr_solve(h, null, null, 0, None, null, null, null, null, null, null, null, 0, null, null) == 0
    # null pointers with B > 0 -> GIK_EINVAL (-1) and a message
    assert L.gik_cost_grad(h, null, null, 4, null, null, null) == -1
    assert b"gik_cost_grad" in L.gik_last_error()
    assert L.gik_rtr_solve(h, null, null, 4, None, null, null, null, null, null, null, null, 0, null, null) == -1
    assert L.gik_proj(1, null, null, 4, null, null) == -1
    # more nodes than any kernel is compiled for -> GIK_ELIMIT (-3) at plan creation
    N = 500
    with pytest.raises(_lib.GikError, match="exceeds the compiled limit"):
        Plan.from_matrices(np.ones((N, N)), np.triu(np.ones((N, N)), 1))
    # a plan without joint tables refuses joint recovery instead of reading garbage
    static = BatchIK(plan=Plan.from_matrices(np.ones((16, 16)), np.triu(np.ones((16, 16)), 1)))
    with pytest.raises(_lib.GikError, match="without joint tables"):
        static.joints(np.zeros((1, 16, 3)))
    # empty batch through the Python engine
    out = eng.solve(np.zeros((0, 4, 4)), check=False)
    assert out["q"].shape == (0, 6) and out["x"].shape == (0, 16, 3)


def test_engine_is_cached_on_the_graph_and_invalidated_by_edits():
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver, solve_with_riemannian
    robot, graph = load_robot("ur10")
    g = golden("ur10_goals")
    solve_with_riemannian(graph, g["T_goal"][0])
    eng1 = graph._gik_engine_cache[1]
    solve_with_riemannian(graph, g["T_goal"][1])
    assert graph._gik_engine_cache[1] is eng1                  # same device plan reused
    graph.add_spherical_obstacle("o0", np.array([0.5, 0.5, 0.5]), 0.1)
    q, Y = solve_with_riemannian(graph, g["T_goal"][1])
    assert graph._gik_engine_cache[1] is not eng1 and Y.shape == (17, 3)
    assert RiemannianSolver(graph).engine is graph._gik_engine_cache[1]


def test_fantope_closed_form_vs_numpy():
    """CIDGIK's closed-form Fantope step (reference convex_iteration.py:43-53) on a batch of Gram matrices of the size
    the UR10 relaxation has (n + d = 13) and at the kernel's limit (32): against numpy.eigh, which is what the
    reference calls.  The projector is independent of eigenvector signs, so the comparison is direct."""
    import torch
    from graphik_b200.solvers.convex_iteration import (solve_fantope_closed_form, solve_fantope_closed_form_batch,
                                                       solve_with_cidgik)
    rng = np.random.default_rng(3)
    for n, d, B in ((13, 3, 257), (32, 3, 33), (5, 2, 4)):
        P = rng.normal(size=(B, n, d + 2))
        G = P @ P.transpose(0, 2, 1) + 1e-3 * np.eye(n)          # near low rank, like an SDP iterate
        G[0] = np.diag(np.arange(n, dtype=float))                  # already diagonal
        C, ev = solve_fantope_closed_form_batch(G, d)
        C, ev = C.cpu().numpy(), ev.cpu().numpy()
        for b in range(B):
            w, Q = np.linalg.eigh(G[b])
            U = np.flip(Q, 1)[:, d:]
            ref = U @ U.T
            assert np.max(np.abs(C[b] - ref)) <= 1e-10, (n, b, np.max(np.abs(C[b] - ref)))
            assert np.max(np.abs(ev[b] - w)) <= 1e-12 * max(1.0, np.max(np.abs(w)))
        assert np.allclose(np.trace(C, axis1=1, axis2=2), n - d, atol=1e-10)
        assert np.max(np.abs(C @ C - C)) <= 1e-10                  # a projector
    C1, seconds = solve_fantope_closed_form(G[1], d)               # the reference's signature
    assert np.max(np.abs(C1 - C[1])) == 0.0 and seconds >= 0.0
