"""Host-side robot / graph model against the reference (goldens) and the reference's own
property tests (tests/test_joint_variables.py, tests/test_distance_matrix.py)."""
import numpy as np
import pytest

from helpers import ROBOTS, golden, load_robot, matrices_for_goal


@pytest.mark.parametrize("name", ROBOTS)
def test_static_graph_identical_to_reference(name):
    robot, graph = load_robot(name)
    g = golden(name + "_graph")
    assert list(g["node_ids"]) == graph.node_ids

    def sym(a):
        return np.where(np.isnan(a), a.T, a)

    for key in ("dist", "lower", "upper"):
        ref, mine = sym(g[key]), getattr(graph, key)
        assert np.array_equal(np.isnan(ref), np.isnan(mine)), key
        assert np.array_equal(np.nan_to_num(ref), np.nan_to_num(mine)), key
    assert np.array_equal(g["below"] | g["below"].T, graph.below)
    assert np.array_equal(g["above"] | g["above"].T, graph.above)


@pytest.mark.parametrize("name", ROBOTS)
def test_goal_matrices_identical_to_reference(name):
    robot, graph = load_robot(name)
    g = golden(name + "_goals")
    for k in range(len(g["f"])):
        T = robot.pose(g["q_goal"][k], "p%d" % robot.n).as_matrix()
        assert np.max(np.abs(T - g["T_goal"][k])) <= 1e-13
        G, D, omega, psi_L, psi_U = matrices_for_goal(graph, g["T_goal"][k])
        assert np.array_equal(D, g["D_goal"][k]) and np.array_equal(omega, g["omega"][k])
        assert np.array_equal(psi_L, g["psi_L"][k]) and np.array_equal(psi_U, g["psi_U"][k])
        q = graph.joint_variables(g["Y_sol"][k], {"p%d" % robot.n: g["T_goal"][k]})
        d = np.array([q["p%d" % (i + 1)] for i in range(robot.n)]) - g["q_sol"][k]
        assert np.max(np.abs(np.mod(d + np.pi, 2 * np.pi) - np.pi)) <= 1e-10


def test_plan_terms_follow_reference_inds_order():
    from graphik_b200.plan import Plan
    from oracle import oracle as orc
    robot, graph = load_robot("ur10")
    g = golden("ur10_goals")
    a = Plan.arrays_from_graph(graph)
    ii, jj = orc.limit_inds(g["omega"][0], g["psi_L"][0], g["psi_U"][0])
    pairs = list(dict.fromkeys(zip(a["term_i"].tolist(), a["term_j"].tolist())))
    assert pairs == list(zip(ii.tolist(), jj.tolist()))
    assert a["n_goal"] == 8 and len(a["goal_edge_i"]) == 8
    assert np.array_equal(a["omega_f"], g["omega"][0])


def test_joint_variables_random_dh_chains():
    """reference tests/test_joint_variables.py:80-111."""
    from graphik_b200.graphs import ProblemGraphRevolute
    from graphik_b200.robots import RobotRevolute
    rng = np.random.RandomState(1)
    for _ in range(25):
        n = rng.randint(3, 20)
        params = {"a": rng.rand(n), "d": rng.rand(n), "alpha": rng.rand(n) * np.pi / 2 - 2 * rng.rand(n) * np.pi / 2,
                  "theta": np.zeros(n), "modified_dh": False, "num_joints": n}
        robot = RobotRevolute(params)
        graph = ProblemGraphRevolute(robot)
        q = robot.random_configuration()
        T_goal = {"p%d" % n: robot.pose(q, "p%d" % n)}
        q_rec = graph.joint_variables(graph.realization(q), T_goal)
        np.testing.assert_allclose(list(q.values()), list(q_rec.values()), rtol=1e-5)


@pytest.mark.parametrize("name", ["ur10", "kuka", "panda", "lwa4d"])
def test_joint_variables_urdf_robots_and_rigid_invariance(name):
    """reference tests/test_joint_variables.py:30-78 (incl. invariance to a rigid motion)."""
    robot, graph = load_robot(name)
    rng = np.random.RandomState(2)
    n = robot.n
    for _ in range(10):
        q = robot.random_configuration()
        T = robot.pose(q, "p%d" % n)
        Y = graph.realization_points(q)
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        Yr = (Y - Y.mean(0)) @ A                       # rotation / reflection + translation
        for P in (Y, Yr):
            q_rec = graph.joint_variables(P, {"p%d" % n: T})
            np.testing.assert_allclose(list(q.values()), list(q_rec.values()), rtol=1e-5, atol=1e-9)


def test_distance_matrix_from_joints_matches_edges():
    """reference tests/test_distance_matrix.py: realised distances reproduce the graph's DIST edges."""
    robot, graph = load_robot("ur10")
    q = robot.random_configuration()
    D = graph.distance_matrix_from_joints(q)
    has = ~np.isnan(graph.dist)
    assert np.max(np.abs(D[has] - graph.dist[has] ** 2)) <= 1e-12


def test_obstacles_reference_and_intended_semantics():
    """in the reference add_spherical_obstacle only adds an anchor."""
    from graphik_b200.utils.utils import table_environment
    robot, graph = load_robot("kuka")
    for k, (c, r) in enumerate(table_environment()):
        graph.add_spherical_obstacle("o%d" % k, c, r)
    assert graph.number_of_nodes() == 118
    L, U = graph.distance_bound_matrices()
    assert int(np.sum(np.triu(L) > 0)) == 9 and int(np.sum(np.triu(U) > 0)) == 6
    G = graph.from_pose(robot.pose(robot.random_configuration(), "p7"))
    assert int(np.sum(np.triu(~np.isnan(G.dist)))) == 5609
    robot2, graph2 = load_robot("kuka", graph_params={"obstacle_semantics": "intended"})
    for k, (c, r) in enumerate(table_environment()):
        graph2.add_spherical_obstacle("o%d" % k, c, r)
    L2, _ = graph2.distance_bound_matrices()
    assert int(np.sum(np.triu(L2) > 0)) == 9 + 7 * 100
    graph2.clear_obstacles()
    assert graph2.number_of_nodes() == 18
