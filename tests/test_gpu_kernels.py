"""Kernel-level parity (tier T1): every CUDA leaf against the CPU oracle / goldens.

The oracle's cost/grad/hess are bit-identical to the reference's own numba `costgrd`
(tests/test_oracle_golden.py), so agreement with the oracle is agreement with the
reference.  Tolerance: 1e-12 relative (fp64, different summation order).
"""
import numpy as np
import pytest

from helpers import ROBOTS, align_columns, golden, load_robot, matrices_for_goal, random_goals

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _engine(name, **kw):
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot(name)
    return robot, graph, BatchIK(graph, **kw)


@pytest.mark.parametrize("name", ROBOTS)
def test_goal_distances_match_reference_D_goal(name):
    robot, graph, eng = _engine(name)
    g = golden(name + "_goals")
    gd = eng.goal_distances(g["T_goal"]).cpu().numpy()
    for k in range(len(g["f"])):
        row = eng.plan.goal_row_from_matrix(g["D_goal"][k])
        assert np.max(np.abs(gd[k] - row)) <= 1e-15 * np.max(row), \
            (name, k, np.max(np.abs(gd[k] - row)))


@pytest.mark.parametrize("name", ["ur10", "kuka", "lwa4d", "chain20"])
@pytest.mark.parametrize("use_limits", [True, False])
def test_cost_grad_hess_vs_oracle(name, use_limits):
    from oracle import oracle as orc
    robot, graph, eng = _engine(name, use_limits=use_limits)
    g = golden(name + "_goals")
    rng = np.random.default_rng(7)
    K = len(g["f"])
    N = graph.number_of_nodes()
    # half random, half near a solution so that hinge terms switch on and off
    Y = np.concatenate([rng.normal(size=(K, N, 3)), g["Y_sol"] + 0.05 * rng.normal(size=(K, N, 3))])
    W = rng.normal(size=(2 * K, N, 3))
    rows = np.stack([eng.plan.goal_row_from_matrix(g["D_goal"][k % K]) for k in range(2 * K)])
    f, gr = eng.cost_grad(Y, rows)
    hv = eng.hessvec(Y, W, rows)
    f, gr, hv = f.cpu().numpy(), gr.cpu().numpy(), hv.cpu().numpy()
    for k in range(2 * K):
        P = orc.Problem(g["D_goal"][k % K], g["omega"][k % K], g["psi_L"][k % K], g["psi_U"][k % K],
                        use_limits=use_limits)
        fo, go, ho = P.cost(Y[k]), P.grad(Y[k]), P.hess(Y[k], W[k])
        assert abs(f[k] - fo) <= RTOL * max(1.0, abs(fo)), (name, k, f[k], fo)
        assert np.max(np.abs(gr[k] - go)) <= RTOL * max(1.0, np.max(np.abs(go))), (name, k)
        assert np.max(np.abs(hv[k] - ho)) <= RTOL * max(1.0, np.max(np.abs(ho))), (name, k)


def test_cost_matches_reference_costgrd_vectors():
    """Directly against the reference's numba-AOT outputs stored in the golden file."""
    from graphik_b200.engine import BatchIK
    from graphik_b200.plan import Plan
    cv = golden("costgrd_vectors")
    for name in ("ur10", "kuka", "chain20"):
        for lim in (True, False):
            plan = Plan.from_matrices(cv[name + "_D_goal"], cv[name + "_omega"], cv[name + "_psi_L"],
                                      cv[name + "_psi_U"], use_limits=lim)
            eng = BatchIK(plan=plan)
            Y, W = cv[name + "_Y"], cv[name + "_W"]
            f, gr = eng.cost_grad(Y)
            hv = eng.hessvec(Y, W)
            p = "l" if lim else "j"
            for got, ref in ((f, cv[name + "_%scost" % p]), (gr, cv[name + "_%sgrad" % p]),
                             (hv, cv[name + "_%shess" % p])):
                got = got.cpu().numpy()
                assert np.max(np.abs(got - ref)) <= RTOL * max(1.0, np.max(np.abs(ref))), (name, lim)
            pr = eng.proj(Y, W).cpu().numpy()
            assert np.max(np.abs(pr - cv[name + "_proj"])) <= 1e-11 * np.max(np.abs(W)), name


def test_cost_ragged_and_empty_batches():
    robot, graph, eng = _engine("ur10")
    g = golden("ur10_goals")
    row = eng.plan.goal_row_from_matrix(g["D_goal"][0])
    for B in (0, 1, 7, 33, 1000):
        Y = np.tile(g["Y_init"][0], (B, 1, 1))
        f, gr = eng.cost_grad(Y, np.tile(row, (B, 1)))
        assert f.shape == (B,) and gr.shape == (B, 16, 3)
        if B:
            fc = f.cpu().numpy()
            assert np.all(fc == fc[0])  # deterministic, identical problems give identical bits


@pytest.mark.parametrize("name", ROBOTS)
def test_bound_smoothing_vs_reference(name):
    robot, graph, eng = _engine(name)
    g = golden(name + "_goals")
    gd = eng.goal_distances(g["T_goal"])
    lb, ub = eng.bounds(gd)
    lb, ub = lb.cpu().numpy(), ub.cpu().numpy()
    assert np.max(np.abs(ub - g["ub"])) <= 1e-12 * np.max(g["ub"]), name
    assert np.max(np.abs(lb - g["lb"])) <= 1e-12 * np.max(g["ub"]), name


def test_bound_smoothing_contains_truth():
    """reference tests/test_bound_smoothing.py:99-117 (UR10 containment), B = 100 configurations."""
    robot, graph, eng = _engine("ur10")
    Q, T = random_goals(robot, 100, seed=22)
    lb, ub = eng.bounds(eng.goal_distances(T))
    lb, ub = lb.cpu().numpy(), ub.cpu().numpy()
    TOL = 1e-6
    for k in range(100):
        D = graph.distance_matrix_from_joints(Q[k])
        assert np.all(D < ub[k] ** 2 + TOL)
        assert np.all(lb[k] ** 2 - TOL < D)


@pytest.mark.parametrize("name", ROBOTS)
def test_initialisation_vs_oracle(name):
    """generate_initialization (riemannian_solver.py:67-75).  The reference's rank heuristic
    (dgp.py:163-171) depends on the arbitrary SIGNS LAPACK gives the Gram eigenvectors, so the
    kernel fixes them canonically; the oracle restates both conventions: "lapack" is pinned
    to the reference's golden Y_init (tests/test_oracle_golden.py), "canonical" is what the
    kernel must reproduce (to 1e-8: Jacobi vs LAPACK eigenvectors, column signs aligned)."""
    from oracle import oracle as orc
    robot, graph, eng = _engine(name)
    g = golden(name + "_goals")
    Y1 = eng.init_from_bounds(g["lb"], g["ub"]).cpu().numpy()      # reference bounds in
    Y2 = eng.initialization(eng.goal_distances(g["T_goal"])).cpu().numpy()  # fused path
    n_same_as_reference = 0
    for k in range(len(g["f"])):
        ref = orc.generate_initialization(g["lb"][k], g["ub"][k], g["omega"][k], signs="canonical")
        scale = np.max(np.abs(ref))
        for Y in (Y1[k], Y2[k]):
            err = np.max(np.abs(align_columns(Y, ref) - ref))
            assert err <= 1e-8 * scale, (name, k, err)
        n_same_as_reference += np.max(np.abs(align_columns(Y1[k], g["Y_init"][k]) - g["Y_init"][k])) <= 1e-8 * scale
    print(name, "goals whose Y_init also equals the reference's LAPACK-signed one:", n_same_as_reference)


@pytest.mark.parametrize("name", ROBOTS)
def test_joint_variables_and_fk(name):
    robot, graph, eng = _engine(name)
    g = golden(name + "_goals")
    q = eng.joints(g["Y_sol"], g["T_goal"]).cpu().numpy()
    d = np.abs(np.mod(q - g["q_sol"] + np.pi, 2 * np.pi) - np.pi)
    assert np.max(d) <= 1e-9, (name, np.max(d))
    # round trip reference tests/test_joint_variables.py:55-78: joint_variables(realization(q)) == q
    Q, T = random_goals(robot, 64, seed=3)
    T_dev, Y_dev = eng.fk(Q)
    assert np.max(np.abs(T_dev.cpu().numpy() - T)) <= 1e-12
    assert np.max(np.abs(Y_dev.cpu().numpy() - graph.realization_points(Q))) <= 1e-12
    q_rec = eng.joints(Y_dev, T_dev).cpu().numpy()
    np.testing.assert_allclose(q_rec, Q, rtol=1e-5, atol=1e-9)


def test_check_limits_device_vs_host_intended():
    """gik_check_limits against the host restatement of check_distance_limits with the intended
    semantics, on realisations that respect / violate a tightened joint limit."""
    import numpy as np
    from graphik_b200.engine import BatchIK
    lim = 0.6 * np.pi * np.ones(6)
    robot, graph = load_robot("ur10", limits=(-lim, lim))
    eng = BatchIK(graph)
    rng = np.random.RandomState(1)
    Q = -np.pi + 2 * np.pi * rng.rand(200, 6)        # many of these break the +-0.6 pi limits
    T, Y = eng.fk(Q)
    dev = eng.check_limits(Y, tol=1e-6).cpu().numpy()
    host = np.array([len(graph.check_distance_limits(Y[k].cpu().numpy(), tol=1e-6, semantics="intended"))
                     for k in range(200)])
    assert np.array_equal(dev, host)
    assert dev.max() > 0 and (dev == 0).any()
    assert all(len(graph.check_distance_limits(Y[k].cpu().numpy(), tol=1e-6)) == 0 for k in range(5))  # as shipped


def test_check_limits_obstacle_on_goal_edge_and_status_code():
    """Intended obstacle semantics: the pair (p_n, obstacle) is a GOAL edge for bound smoothing (exact distance per
    goal) and still carries the obstacle's lower limit for check_distance_limits.  A realisation whose end-effector
    point lies inside an obstacle sphere must be reported by the device check exactly as by the host one, and
    `status` becomes GIK_STATUS_LIMITS (3) for those goals only."""
    import numpy as np
    import torch
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot("kuka", graph_params={"obstacle_semantics": "intended"})
    rng = np.random.RandomState(4)
    Q = -np.pi + 2 * np.pi * rng.rand(64, robot.n)
    P = robot.fk_all(Q)[:, robot.n, :3, 3]
    # spheres centred on the end-effector points of the first goals: those goals end inside an obstacle
    for k in range(3):
        graph.add_spherical_obstacle("o%d" % k, P[k] + np.array([0.0, 0.0, 0.01]), 0.25)
    eng = BatchIK(graph)
    T, Y = eng.fk(Q)
    status = torch.zeros(64, dtype=torch.int32, device=Y.device)
    status[10] = 2                                    # a NaN start keeps its own code
    dev = eng.check_limits(Y, tol=1e-6, status=status).cpu().numpy()
    host = np.array([len(graph.check_distance_limits(Y[k].cpu().numpy(), tol=1e-6, semantics="intended"))
                     for k in range(64)])
    assert np.array_equal(dev, host)
    assert (dev[:3] > 0).all()
    st = status.cpu().numpy()
    assert np.array_equal(st == 3, (dev > 0) & (np.arange(64) != 10)) and st[10] == 2
