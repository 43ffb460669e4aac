"""CIDGIK on the CPU (no GPU needed): the reference's problem construction, pinned; the product's host-side plan against
the oracle.  The SDP solver itself has no reference to pin against (MOSEK, closed source): see oracle/cidgik.py."""
import numpy as np
import pytest

from helpers import golden, load_robot, random_goals

ROBOTS = ["ur10", "kuka", "lwa4d"]


def _anchors(graph, T):
    n = graph.robot.n
    pq = graph.goal_points(T)
    return {"p0": graph.pos[graph.idx("p0")], "q0": graph.pos[graph.idx("q0")], "p%d" % n: pq[0], "q%d" % n: pq[1]}


@pytest.mark.parametrize("name", ROBOTS)
def test_constraint_construction_matches_the_reference(name):
    """oracle.sdp_constraints == the matrices distance_constraints_graph (sdp_snl.py:270-314) built for the anchors
    solve_with_cidgik sets (convex_iteration.py:284-289), up to the order of variables and constraints."""
    from oracle import cidgik as cg
    gold = golden("cidgik_constraints")
    robot, graph = load_robot(name)
    for g in range(2):
        key = "%s_%d_" % (name, g)
        T = gold[key + "T_goal"]
        anchors = _anchors(graph, T)
        np.testing.assert_allclose(np.stack([anchors[u] for u in anchors]), gold[key + "anchors"], atol=1e-12)
        free, A, b, pairs = cg.sdp_constraints(graph.node_ids, graph.dist, anchors)
        order = [str(u) for u in gold[key + "order"]]
        assert sorted(order) == sorted(free)
        perm = [order.index(u) for u in free] + [len(free), len(free) + 1, len(free) + 2]   # mine -> reference index
        ref = {tuple(sorted(map(str, p))): k for k, p in enumerate(gold[key + "pairs"])}
        assert len(ref) == len(pairs) == len(gold[key + "b"])
        for Ak, bk, (u, v) in zip(A, b, pairs):
            k = ref[tuple(sorted((u, v)))]
            np.testing.assert_allclose(Ak, gold[key + "A"][k][np.ix_(perm, perm)], rtol=0, atol=1e-13)
            np.testing.assert_allclose(bk, gold[key + "b"][k], rtol=1e-12, atol=1e-13)
        assert int(gold[key + "n_inequalities"]) == 0       # distance_range_constraints: obstacle pairs only, none exist


@pytest.mark.parametrize("name", ROBOTS)
def test_true_configuration_satisfies_the_program(name):
    """tests/test_sdp_snl.py of the reference: the linear maps evaluate to zero at the true points."""
    from oracle import cidgik as cg
    robot, graph = load_robot(name)
    Q, T = random_goals(robot, 3, 5)
    for k in range(3):
        anchors = _anchors(graph, T[k])
        free, A, b, _ = cg.sdp_constraints(graph.node_ids, graph.dist, anchors)
        P = np.asarray(graph.realization_points(Q[k]))
        X = np.hstack([np.stack([P[graph.idx(u)] for u in free]).T, np.eye(3)])
        Z = X.T.dot(X)
        r = [np.sum(Ak * Z) - bk for Ak, bk in zip(A, b)]
        assert np.max(np.abs(r)) < 1e-12


@pytest.mark.parametrize("name", ROBOTS)
def test_plan_coordinates_describe_the_same_program(name):
    """The product's coordinates (eliminated nodes, rank-one constraints) against the oracle's (orthonormal face basis,
    the reference's matrices): same face, the true configuration lies on it, the reduced constraints imply all of the
    reference's, and the two programs have the same optimum."""
    from oracle import cidgik as cg
    from graphik_b200.solvers.convex_iteration import CidgikPlan
    robot, graph = load_robot(name)
    n = robot.n
    plan = CidgikPlan(graph)
    assert plan.Nr < plan.N and plan.M < plan.n_distance_constraints + 6
    Q, T = random_goals(robot, 2, 3)
    anchors, W, b, V = [t.numpy() for t in plan.assemble(T, device="cpu")]
    P_generic = np.asarray(graph.realization_points(np.random.RandomState(7).uniform(-2, 2, n)))
    for k in range(2):
        an = {u: anchors[k, i] for i, u in enumerate(plan.anchor_names)}
        np.testing.assert_allclose(anchors[k], np.stack(list(_anchors(graph, T[k]).values())), atol=1e-14)
        Vo = cg.face_basis(graph.node_ids, graph.dist, an, P_generic)
        assert Vo.shape == V[k].shape
        assert np.abs(Vo.dot(Vo.T).dot(V[k]) - V[k]).max() < 1e-10                 # same column space
        free, A, bb, _ = cg.sdp_constraints(graph.node_ids, graph.dist, an)
        assert free == plan.free
        Ai, bi = cg.identity_block_constraints(plan.N)
        A, bb = np.array(A + Ai), np.array(bb + bi)
        # the true configuration: on the face, and feasible in both coordinates
        Pk = np.asarray(graph.realization_points(Q[k]))
        X = np.hstack([np.stack([Pk[graph.idx(u)] for u in free]).T, np.eye(3)])
        Z = X.T.dot(X)
        Zr = np.linalg.pinv(V[k]).dot(Z).dot(np.linalg.pinv(V[k]).T)
        assert np.abs(V[k].dot(Zr).dot(V[k].T) - Z).max() < 1e-10
        assert np.abs(np.einsum("ki,ij,kj->k", W[k], Zr, W[k]) - b[k]).max() < 1e-10
        # optimum of the first program of the convex iteration (C = I) in the two coordinate systems
        sol = cg.solve_sdp(V[k].T.dot(V[k]), np.einsum("ki,kj->kij", W[k], W[k]), b[k])
        out = cg.convex_iterate(graph.node_ids, graph.dist, an, P_generic, max_iters=1)
        assert sol["resid"] < 1e-5 and out["sdp"]["resid"] < 1e-4
        assert abs(sol["obj"] - out["values"][0]) < 2e-4 * (1 + abs(sol["obj"]))
        Zp = V[k].dot(sol["X"]).dot(V[k].T)
        assert np.abs(A.reshape(len(bb), -1).dot(Zp.ravel()) - bb).max() < 1e-4     # every constraint of the reference
        cert = cg.certificate(V[k].T.dot(V[k]), np.einsum("ki,kj->kij", W[k], W[k]), b[k], sol["X"], sol["y"])
        assert cert["pres"] < 1e-5 and cert["gap"] < 1e-5 and cert["min_eig_X"] > -1e-9 and cert["min_eig_S"] > -1e-7


def test_oracle_convex_iteration_reaches_the_goal():
    """The whole CPU statement on three UR10 goals: the convex iteration stops within its 10 iterations
    (convex_iteration.py:164; supplement II-B: 'typically fewer than 10') and the recovered joint angles reach the goal."""
    from oracle import cidgik as cg
    robot, graph = load_robot("ur10")
    n = robot.n
    Q, T = random_goals(robot, 3, 0)
    P_generic = np.asarray(graph.realization_points(np.random.RandomState(7).uniform(-2, 2, n)))
    for k in range(3):
        anchors = _anchors(graph, T[k])
        out = cg.convex_iterate(graph.node_ids, graph.dist, anchors, P_generic)
        assert out["feasible"] and len(out["values"]) <= 10 and out["values"][-1] < 1e-6
        pts = cg.extract_points(out["free"], out["Z"])
        P = {u: graph.pos[graph.idx(u)] for u in ("p0", "x", "y", "q0")}
        P.update(pts)
        P.update({u: anchors[u] for u in ("p%d" % n, "q%d" % n)})
        Y = np.stack([P[u] for u in graph.node_ids])
        q = graph.joint_variables_batch(Y[None], T[k][None])[0]
        Tq = robot.fk_all(q[None])[0, n]
        assert np.abs(Tq - T[k]).max() < 1e-3


def test_infeasible_program_is_reported():
    """A goal out of reach: the program has no feasible point and the interior-point method finds the dual ray."""
    from oracle import cidgik as cg
    robot, graph = load_robot("ur10")
    T = np.eye(4)
    T[:3, 3] = [5.0, 0.0, 0.0]
    out = cg.convex_iterate(graph.node_ids, graph.dist, _anchors(graph, T),
                            np.asarray(graph.realization_points(np.random.RandomState(7).uniform(-2, 2, robot.n))))
    assert not out["feasible"] and out["status"] == cg.STATUS_INFEASIBLE


def test_inequality_rows_match_the_reference_form():
    """The plan's inequality rows (obstacle_semantics="intended") against anchor_inequality_constraint
    (sdp_snl.py:586-618), called by the golden script for every free joint point of the UR10 and one sphere: both
    expressions take the same value on matrices of the feasible face (lifted true configurations and a mixture of two
    of them, which has rank 6), with the same sense.  As shipped, the reference's distance_range_constraints returns
    nothing even with the sphere in its graph (the golden script asserts it)."""
    from graphik_b200.solvers.convex_iteration import CidgikPlan
    gold = golden("cidgik_constraints")
    robot, graph = load_robot("ur10", graph_params={"obstacle_semantics": "intended"})
    centre, radius = gold["ineq_centre"], float(gold["ineq_radius"])
    graph.add_spherical_obstacle("o0", centre, radius)
    plan = CidgikPlan(graph)
    n = robot.n
    assert plan.n_inequalities == n - 1 and plan.M == int(np.sum(plan.tau == 0)) + n - 1
    assert int(gold["ineq_n_equalities"]) == plan.n_distance_constraints        # the sphere adds no equality
    T = gold["ineq_T_goal"]
    anchors, W, b, V = [t.numpy()[0] for t in plan.assemble(T[None], device="cpu")]
    order = [str(u) for u in gold["ineq_order"]]
    perm = [order.index(u) for u in plan.free] + [len(order), len(order) + 1, len(order) + 2]
    rows = W[plan.tau != 0]
    rhs = b[plan.tau != 0]
    assert np.all(plan.tau[plan.tau != 0] == -1.0)                              # lower bounds: w^T Zr w >= r^2
    # two configurations with the same end-effector pose do not come for free: use matrices of the face built from the
    # product's own coordinates instead -- Zr = R R^T with the hom block forced to I
    rng = np.random.RandomState(3)
    for trial in range(3):
        R = rng.randn(plan.Nr, 3 + trial)                                      # rank 3, 4, 5
        R[-3:, :3], R[-3:, 3:] = np.eye(3), 0.0
        Zr = R.dot(R.T)
        Z = V.dot(Zr).dot(V.T)
        assert np.abs(Z[-3:, -3:] - np.eye(3)).max() < 1e-12
        for k, u in enumerate(str(v) for v in gold["ineq_nodes"]):
            A_ref = gold["ineq_A"][k][np.ix_(perm, perm)]
            ref = np.sum(A_ref * Z) - gold["ineq_b"][k]                         # reference: <A, Z> - b <= 0
            mine = -(rows[k].dot(Zr).dot(rows[k]) - rhs[k])                     # product: -(w^T Zr w - r^2) <= 0
            assert abs(ref - mine) < 1e-10 * (1 + abs(ref)), (u, ref, mine)


def test_obstacles_with_reference_semantics_change_nothing():
    """The reference's obstacles are anchors joined to the other anchors only: the program of KUKA +
    table_environment() is the program of the bare KUKA; with the intended semantics every free joint point gets one
    lower bound per sphere (600 rows -- beyond the kernel's 96, reported by gik_sdp_solve as GIK_ELIMIT)."""
    from helpers import load_kuka_table
    from graphik_b200.solvers.convex_iteration import CidgikPlan
    bare = CidgikPlan(load_robot("kuka")[1])
    table = CidgikPlan(load_kuka_table()[1])
    assert (table.N, table.Nr, table.M, table.n_inequalities) == (bare.N, bare.Nr, bare.M, 0)
    for k in ("WK", "WA", "WH", "b", "VK", "VA", "VH"):
        np.testing.assert_array_equal(getattr(table, k), getattr(bare, k))
    intended = CidgikPlan(load_kuka_table(graph_params={"obstacle_semantics": "intended"})[1])
    assert intended.n_inequalities == 100 * 6 and intended.M == bare.M + 600
