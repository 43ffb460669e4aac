"""Graph sizes around the kernels' tiling boundaries (the BASELINE robots only hit N = 16, 18, 44 and 118):

* one-warp kernels: N = 32 (one lane per node, every lane used), 34 (two second nodes), 48, 64 (every lane two nodes);
* dense CTA kernel: N = 38 (two 32-node blocks, ragged), 68 (three blocks, ragged), 100 (four blocks, 4 rows in the last),
  100 with intended obstacle semantics (second terms around the hub node p_n).

Each case runs the whole device pipeline (goal distances, bound smoothing + initialisation, solve) and checks the selected
kernel against the group kernel (identical leading decisions, different bits -- i.e. no silent fallback) and against the
oracle (cost / gradient at the initial point, leading trust-region decisions, reported cost == lcost of the result)."""
import numpy as np
import pytest

from helpers import load_robot

pytestmark = pytest.mark.gpu


def _random_dh_chain(n, seed):
    """RobotRevolute from random DH parameters, as the reference's tests build them (tests/test_joint_variables.py:80-102)."""
    from graphik_b200.graphs.graph_revolute import ProblemGraphRevolute
    from graphik_b200.robots.robot_revolute import RobotRevolute
    rng = np.random.RandomState(seed)
    a, d = rng.rand(n), rng.rand(n)
    al = rng.rand(n) * np.pi / 2 - 2 * rng.rand(n) * np.pi / 2
    params = {"a": a, "alpha": al, "d": d, "theta": np.zeros(n), "modified_dh": False, "num_joints": n,
              "joint_limits_upper": np.pi * np.ones(n), "joint_limits_lower": -np.pi * np.ones(n)}
    robot = RobotRevolute(params)
    return robot, ProblemGraphRevolute(robot)


def _check_against_group_kernel_and_oracle(robot, graph, kernel, B=4, maxiter=40, seed=3, has_other_kernel=True):
    from graphik_b200.engine import BatchIK, make_opts
    from oracle import oracle as orc
    eng = BatchIK(graph)
    a = eng.plan._a
    rng = np.random.RandomState(seed)
    Q = -np.pi + 2 * np.pi * rng.rand(B, robot.n)
    T = robot.fk_all(Q)[:, robot.n]
    gd = eng.goal_distances(T)
    Y0 = eng.initialization(gd)
    assert bool(np.isfinite(Y0.cpu().numpy()).all())
    f0, g0 = eng.cost_grad(Y0, gd)
    out = eng.solve_points(gd, Y0, trace_rows=8, opts=make_opts({"maxiter": maxiter, "kernel": kernel}))
    gen = eng.solve_points(gd, Y0, trace_rows=8, opts=make_opts({"maxiter": 8, "kernel": "generic"}))
    tr, tg = out["trace"].cpu().numpy(), gen["trace"].cpu().numpy()
    assert np.array_equal(tr[:, :4][:, :, [1, 2, 4]], tg[:, :4][:, :, [1, 2, 4]]), (tr[:, :4], tg[:, :4])
    np.testing.assert_allclose(tr[:, :4, 3], tg[:, :4, 3], rtol=1e-6)
    if has_other_kernel:
        assert not np.array_equal(tr[:, :8, 3], tg[:, :8, 3]), "the requested kernel fell back to the group kernel"
    gdh, Y0h = gd.cpu().numpy(), Y0.cpu().numpy()
    x, fx = out["x"].cpu().numpy(), out["f(x)"].cpu().numpy()
    gs = a["goal_slot"]
    ii, jj = np.nonzero(gs >= 0)
    for k in range(B):
        D = a["D_static"].copy()
        D[ii, jj] = gdh[k, gs[ii, jj]]
        P = orc.Problem(D, a["omega_f"], a["psi_L"], a["psi_U"])
        fo, go = P.cost(Y0h[k]), P.grad(Y0h[k])
        assert abs(float(f0[k]) - fo) <= 1e-12 * fo
        assert np.max(np.abs(g0[k].cpu().numpy() - go)) <= 1e-12 * np.max(np.abs(go))
        ref = P.solve(Y0h[k], params={"maxiter": 6}, trace_rows=6)["trace"]
        m = min(len(ref), 4)
        assert np.array_equal(tr[k, :m][:, [1, 2, 4]], ref[:m][:, [1, 2, 4]]), (k, tr[k, :m], ref[:m])
        np.testing.assert_allclose(tr[k, :m, 3], ref[:m, 3], rtol=1e-6)
        assert abs(P.cost(x[k]) - fx[k]) <= 1e-11 * max(1.0, fx[k])
    assert np.all(fx < f0.cpu().numpy())
    return eng


@pytest.mark.parametrize("n", [14, 15, 22, 30])
def test_warp_kernels_at_lane_boundaries(n):
    robot, graph = _random_dh_chain(n, seed=n)
    N = graph.number_of_nodes()
    assert N == 2 * n + 4
    eng = _check_against_group_kernel_and_oracle(robot, graph, "latency")
    assert eng.plan.N == N


@pytest.mark.parametrize("n_obstacles,semantics", [(20, "reference"), (50, "reference"), (82, "reference"), (82, "intended")])
def test_dense_kernel_at_block_boundaries(n_obstacles, semantics):
    from graphik_b200.utils.utils import table_environment
    robot, graph = load_robot("kuka", graph_params={"obstacle_semantics": semantics})
    for k, (c, r) in enumerate(table_environment()[:n_obstacles]):
        graph.add_spherical_obstacle("o%d" % k, c, r)
    assert graph.number_of_nodes() == 18 + n_obstacles
    _check_against_group_kernel_and_oracle(robot, graph, "dense", B=3, maxiter=30)


@pytest.mark.parametrize("n_height,n_width", [(12, 10), (14, 12), (18, 16)])
def test_graphs_above_128_nodes(n_height, n_width):
    """Beyond the dense kernel's 128 nodes (table_environment() with a finer grid: N = 166 / 218 / 346, reference
    utils.py:179-191): the group kernels with eight / fifteen nodes per lane, bound smoothing without register tiles and all three
    matrices in the caller's workspace.  Whole device pipeline against the oracle, as for the other shapes, plus the
    bounds against the oracle's bound_smoothing on one goal."""
    from oracle import oracle as orc
    from graphik_b200.utils.utils import table_environment
    robot, graph = load_robot("kuka")
    for k, (c, r) in enumerate(table_environment(n_height=n_height, n_width=n_width)):
        graph.add_spherical_obstacle("o%d" % k, c, r)
    N = graph.number_of_nodes()
    assert 128 < N <= 480
    eng = _check_against_group_kernel_and_oracle(robot, graph, "auto", B=2, maxiter=12, has_other_kernel=False)
    assert eng.plan.N == N
    rng = np.random.RandomState(5)
    T = robot.fk_all(-np.pi + 2 * np.pi * rng.rand(1, robot.n))[:, robot.n]
    lb, ub = eng.bounds(eng.goal_distances(T))
    G = graph.from_pose(T[0])
    lo, up = orc.bound_smoothing(G.edge, G.lower, G.upper)
    assert np.max(np.abs(ub[0].cpu().numpy() - up)) <= 1e-12 * np.max(up)
    assert np.max(np.abs(lb[0].cpu().numpy() - lo)) <= 1e-12 * np.max(up)
