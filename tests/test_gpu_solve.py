"""Solver parity on the GPU (tiers T2 and T3 of DESIGN.md) through the C ABI.

T2  with the reference's own Y_init injected, the per-outer-iteration decisions
    (tCG iterations, stop reason, accept/reject) agree with the oracle for the leading
    iterations and fx_prop to 1e-6 relative; RTR trajectories are chaotic w.r.t. rounding
    (the C oracle itself leaves the reference's trajectory after 9-18 iterations), so
    later iterations are compared through the end state only.
T3  end-to-end statistics on a few hundred goals against the oracle run on the same
    goals: success rate, final cost, iteration counts.
FP tolerances are written next to each assertion.
"""
import numpy as np
import pytest

from helpers import golden, load_robot, random_goals

pytestmark = pytest.mark.gpu


def _engine(name, **kw):
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot(name)
    return robot, graph, BatchIK(graph, **kw)


@pytest.mark.parametrize("name", ["ur10", "kuka", "lwa4d", "lwa4p", "chain20"])
def test_trace_agrees_with_oracle_on_leading_iterations(name):
    from oracle import oracle as orc
    robot, graph, eng = _engine(name)
    g = golden(name + "_goals")
    K = len(g["f"])
    rows = np.stack([eng.plan.goal_row_from_matrix(g["D_goal"][k]) for k in range(K)])
    out = eng.solve_points(rows, g["Y_init"], trace_rows=64)
    tr = out["trace"].cpu().numpy()
    iters = out["iterations"].cpu().numpy()
    f = out["f(x)"].cpu().numpy()
    gn = out["gradnorm"].cpu().numpy()
    status = out["status"].cpu().numpy()
    lead = []
    for k in range(K):
        P = orc.Problem(g["D_goal"][k], g["omega"][k], g["psi_L"][k], g["psi_U"][k])
        ref = P.solve(g["Y_init"][k], trace_rows=64)["trace"]
        m = min(len(ref), iters[k], 64)
        same = np.all(tr[k, :m][:, [1, 2, 4]] == ref[:m][:, [1, 2, 4]], axis=1)
        n_same = m if same.all() else int(np.argmin(same))
        lead.append(n_same)
        # the first outer iterations must be decision-identical and numerically tight (observed on the 30 golden
        # goals: 7 .. 46 identical leading iterations before rounding differences flip a decision)
        assert n_same >= min(7, m), (name, k, n_same, tr[k, :8], ref[:8])
        np.testing.assert_allclose(tr[k, :min(4, m), 3], ref[:min(4, m), 3], rtol=1e-6)
        np.testing.assert_allclose(tr[k, :min(4, m), 0], ref[:min(4, m), 0], rtol=0)  # Delta: exact
    print(name, "decision-identical leading outer iterations per goal:", lead)
    # end state: same stopping rule as the reference (gradnorm < 5e-10 unless maxiter)
    assert np.all((status == 0) | (status == 1))
    assert np.all(gn[status == 0] < 5e-10)
    assert np.all(iters[status == 1] == 3000)
    # final EDM residual in the range the reference reaches: ~1e-15 when converged, up to ~1e-10 on the
    # slowly converging goals that run into maxiter (which goals those are is trajectory dependent)
    ok_ref = g["f"] < 1e-12
    assert np.all(f[ok_ref] < 1e-9), (name, f, g["f"])
    assert np.median(f[ok_ref]) < 1e-13


def test_reported_cost_is_lcost_of_returned_points():
    """f(x) returned by the solver == lcost (costs.py:79-93) recomputed by the oracle on x."""
    from oracle import oracle as orc
    robot, graph, eng = _engine("ur10")
    g = golden("ur10_goals")
    K = len(g["f"])
    rows = np.stack([eng.plan.goal_row_from_matrix(g["D_goal"][k]) for k in range(K)])
    out = eng.solve_points(rows, g["Y_init"])
    x, f = out["x"].cpu().numpy(), out["f(x)"].cpu().numpy()
    for k in range(K):
        P = orc.Problem(g["D_goal"][k], g["omega"][k], g["psi_L"][k], g["psi_U"][k])
        assert abs(P.cost(x[k]) - f[k]) <= 1e-12 * max(1.0, f[k]) + 1e-25


@pytest.mark.parametrize("name,B", [("ur10", 256), ("kuka", 128), ("chain20", 96)])
def test_batch_statistics_match_oracle(name, B):
    from oracle import oracle as orc
    robot, graph, eng = _engine(name)
    Q, T = random_goals(robot, B, seed=0)
    out = eng.solve(T, check=True)
    f = out["f(x)"].cpu().numpy()
    it = out["iterations"].cpu().numpy()
    pos = out["pos_err"].cpu().numpy()
    n_inner = out["n_inner"].cpu().numpy()
    # oracle on the same goals, same initial points
    gd = out["goal_d2"].cpu().numpy()
    Y0 = eng.initialization(out["goal_d2"]).cpu().numpy()
    a = eng.plan._a
    D = np.repeat(a["D_static"][None], B, 0)
    gs = a["goal_slot"]
    ii, jj = np.nonzero(gs >= 0)
    D[:, ii, jj] = gd[:, gs[ii, jj]]
    ref = orc.solve_batch(D, a["omega_f"], a["psi_L"], a["psi_U"], Y0)
    q_ref = graph.joint_variables_batch(ref["x"], T)
    pos_ref = np.linalg.norm(robot.fk_all(q_ref)[:, robot.n, :3, 3] - T[:, :3, 3], axis=1)
    succ, succ_ref = np.mean(pos < 1e-2), np.mean(pos_ref < 1e-2)
    print(name, "success gpu/oracle", succ, succ_ref, "median iters", np.median(it), np.median(ref["iterations"]),
          "median f", np.median(f), np.median(ref["f(x)"]), "mean inner", n_inner.mean(), ref["n_hess"].mean())
    assert abs(succ - succ_ref) <= 0.05                      # success rate within 5 points
    if name != "chain20":
        # (the EDM of the random-DH 20-DOF chain is completed to 1e-25 by both, but `joint_variables` on the completed
        # points does not reproduce the pose for either -- mirror-image realisations; a property of the reference path)
        assert succ >= 0.9
    assert np.median(f) < 1e-13 and np.median(ref["f(x)"]) < 1e-13
    r = np.median(it) / np.median(ref["iterations"])
    assert 0.8 < r < 1.25, r                                   # same iteration-count distribution (observed 0.98 - 1.03)
    r = n_inner.mean() / ref["n_hess"].mean()
    assert 0.7 < r < 1.4, r                                    # (a mean over a heavy tail: observed 0.99 - 1.05)


def test_results_do_not_depend_on_batch_composition():
    """No cross-problem arithmetic: any sharding of the batch gives bit-identical per-problem results
    (this is what makes the multi-GPU split exact)."""
    robot, graph, eng = _engine("ur10")
    Q, T = random_goals(robot, 96, seed=5)
    full = eng.solve(T, check=False)
    parts = [eng.solve(T[s], check=False) for s in (slice(0, 1), slice(1, 40), slice(40, 96))]
    for key in ("x", "f(x)", "gradnorm", "iterations", "q"):
        a = full[key].cpu().numpy()
        b = np.concatenate([p[key].cpu().numpy() for p in parts])
        assert np.array_equal(a, b), key


def test_maxiter_and_nan_status():
    from graphik_b200.engine import make_opts
    robot, graph, eng = _engine("ur10")
    g = golden("ur10_goals")
    rows = np.stack([eng.plan.goal_row_from_matrix(g["D_goal"][k]) for k in range(2)])
    out = eng.solve_points(rows, g["Y_init"][:2], opts=make_opts({"maxiter": 3}))
    assert np.all(out["iterations"].cpu().numpy() == 3) and np.all(out["status"].cpu().numpy() == 1)
    Y0 = g["Y_init"][:2].copy()
    Y0[1, 0, 0] = np.nan
    out = eng.solve_points(rows, Y0)
    st = out["status"].cpu().numpy()
    assert st[0] == 0 and st[1] == 2   # a NaN problem is flagged and does not poison its neighbour


def test_kernel_variants_agree():
    """gik_rtr_solve has three implementations of the same algorithm -- latency (one warp per problem,
    register slot cache), throughput (two problems per warp in lock-step) and generic (W-lane groups, any
    N) -- that differ only in summation order: identical leading decisions, same end quality."""
    from graphik_b200.engine import make_opts
    robot, graph, eng = _engine("ur10")
    g = golden("ur10_goals")
    K = len(g["f"])
    rows = np.stack([eng.plan.goal_row_from_matrix(g["D_goal"][k]) for k in range(K)])
    outs = {name: eng.solve_points(rows, g["Y_init"], trace_rows=32, opts=make_opts({"kernel": name}))
            for name in ("latency", "throughput", "generic")}
    tf = outs["latency"]["trace"].cpu().numpy()
    for name in ("throughput", "generic"):
        tg = outs[name]["trace"].cpu().numpy()
        for k in range(K):
            assert np.array_equal(tf[k, :6][:, [1, 2, 4]], tg[k, :6][:, [1, 2, 4]]), (name, k, tf[k, :6], tg[k, :6])
            np.testing.assert_allclose(tf[k, :6, 3], tg[k, :6, 3], rtol=1e-6)
    for name, out in outs.items():
        assert np.all(out["f(x)"].cpu().numpy() < 1e-11), name
        assert np.all(out["gradnorm"].cpu().numpy() < 5e-10), name
        assert np.all(out["status"].cpu().numpy() == 0), name


def test_two_nodes_per_lane_kernel_agrees_with_generic():
    """33 .. 64 nodes (20-DOF chain, N = 44): the one-warp-per-problem kernel with two nodes per lane (degree-ordered
    node assignment, slot cache in shared memory) against the generic group kernel: identical leading decisions,
    same end quality, and results that do not depend on which problems share the launch."""
    from graphik_b200.engine import make_opts
    robot, graph, eng = _engine("chain20")
    g = golden("chain20_goals")
    K = len(g["f"])
    rows = np.stack([eng.plan.goal_row_from_matrix(g["D_goal"][k]) for k in range(K)])
    outs = {name: eng.solve_points(rows, g["Y_init"], trace_rows=32, opts=make_opts({"kernel": name}))
            for name in ("latency", "generic")}
    tf, tg = outs["latency"]["trace"].cpu().numpy(), outs["generic"]["trace"].cpu().numpy()
    for k in range(K):
        assert np.array_equal(tf[k, :6][:, [1, 2, 4]], tg[k, :6][:, [1, 2, 4]]), (k, tf[k, :6], tg[k, :6])
        np.testing.assert_allclose(tf[k, :6, 3], tg[k, :6, 3], rtol=1e-6)
    for name, out in outs.items():
        assert np.all(out["status"].cpu().numpy() == 0), name
        assert np.all(out["gradnorm"].cpu().numpy() < 5e-10), name
        assert np.all(out["f(x)"].cpu().numpy() < 1e-11), name
    # the two kernels must really be different code paths: summation order differs, so bits differ somewhere
    assert not np.array_equal(outs["latency"]["x"].cpu().numpy(), outs["generic"]["x"].cpu().numpy())
    # sharding exactness: reversed order, one problem at a time
    rev = eng.solve_points(rows[::-1].copy(), g["Y_init"][::-1].copy(), opts=make_opts({"kernel": "latency"}))
    assert np.array_equal(rev["x"].cpu().numpy()[::-1], outs["latency"]["x"].cpu().numpy())
    one = eng.solve_points(rows[2:3], g["Y_init"][2:3], opts=make_opts({"kernel": "latency"}))
    assert np.array_equal(one["x"].cpu().numpy()[0], outs["latency"]["x"].cpu().numpy()[2])


def test_throughput_kernel_statistics_and_independence():
    """Lock-step kernel on a ragged batch (odd size, so one half-warp idles at the end): same statistics as
    the latency kernel, bit-identical results whatever problem shares the warp."""
    from graphik_b200.engine import make_opts
    robot, graph, eng = _engine("ur10")
    Q, T = random_goals(robot, 301, seed=11)
    gd = eng.goal_distances(T)
    Y0 = eng.initialization(gd)
    thr = eng.solve_points(gd, Y0, opts=make_opts({"kernel": "throughput"}))
    lat = eng.solve_points(gd, Y0, opts=make_opts({"kernel": "latency"}))
    perm = np.random.RandomState(0).permutation(301)
    import torch
    pt = torch.as_tensor(perm, device=gd.device)
    thr2 = eng.solve_points(gd[pt].contiguous(), Y0[pt].contiguous(), opts=make_opts({"kernel": "throughput"}))
    for key in ("x", "f(x)", "gradnorm", "iterations", "n_inner"):
        assert np.array_equal(thr[key].cpu().numpy()[perm], thr2[key].cpu().numpy()), key
    it_t, it_l = thr["iterations"].cpu().numpy(), lat["iterations"].cpu().numpy()
    assert 0.8 < np.median(it_t) / np.median(it_l) < 1.25
    assert np.median(thr["f(x)"].cpu().numpy()) < 1e-13
    assert abs(np.mean(thr["status"].cpu().numpy() == 0) - np.mean(lat["status"].cpu().numpy() == 0)) < 0.03
    # NaN input does not disturb the problem sharing its warp
    Yn = Y0[:2].clone()
    Yn[0, 3, 1] = float("nan")
    o = eng.solve_points(gd[:2], Yn, opts=make_opts({"kernel": "throughput"}))
    assert o["status"].cpu().numpy().tolist() == [2, 0]
    assert np.array_equal(o["x"][1].cpu().numpy(), thr["x"][1].cpu().numpy())


@pytest.mark.parametrize("name", ["ur10", "kuka", "lwa4d", "lwa4p", "panda"])
def test_end_state_vs_reference_sample(name):
    """256 (UR10, KUKA) / 48 (LWA4D, LWA4P, Panda) goals solved by the UNMODIFIED reference (tests/golden/<robot>_stats.npz, made by
    oracle/gen_golden_stats.py) against the GPU started from the reference's own Y_init:
    recovered joint angles (same IK branch, 1e-3 rad -- both solvers stop at |g| < 5e-10 but follow
    rounding-perturbed trajectories, so a minority of goals ends in another of the <= 16 IK branches),
    EDM residual (<= 1e-9 wherever the reference reaches 1e-12) and end-effector error (<= 1e-2 m
    wherever the reference achieves it)."""
    import os
    from helpers import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, name + "_stats.npz")):
        pytest.skip("reference sample not generated")
    robot, graph, eng = _engine(name)
    g = golden(name + "_stats")
    T = g["T_goal"]
    gd = eng.goal_distances(T)
    out = eng.solve_points(gd, g["Y_init"])
    q = eng.joints(out["x"], T).cpu().numpy()
    f = out["f(x)"].cpu().numpy()
    T_sol, _ = eng.fk(q, want_points=False)
    pos = np.linalg.norm(T_sol.cpu().numpy()[:, :3, 3] - T[:, :3, 3], axis=1)
    dq = np.max(np.abs(np.mod(q - g["q_sol"] + np.pi, 2 * np.pi) - np.pi), axis=1)
    # 6-DOF: isolated IK branches; 7-DOF: a one-parameter family of solutions per pose, so
    # rounding-perturbed trajectories stop at nearby points of the same family
    same = dq < (1e-3 if robot.n == 6 else 5e-2)
    ref_ok, ref_conv = g["pose_err"] < 1e-2, g["f"] < 1e-12
    print(name, "same IK branch as the reference: %d/%d; success gpu %.3f ref %.3f; median f gpu %.2e ref %.2e; "
          "median iters gpu %d ref %d" % (same.sum(), len(same), np.mean(pos < 1e-2), np.mean(ref_ok),
                                          np.median(f), np.median(g["f"]),
                                          np.median(out["iterations"].cpu().numpy()), np.median(g["iterations"])))
    # observed with the round-2 kernels: UR10 222/256, KUKA 249/256, LWA4D 47/48, LWA4P 45/48, Panda 48/48
    assert np.mean(same) >= {"ur10": 0.85, "kuka": 0.95, "lwa4d": 0.9, "lwa4p": 0.85, "panda": 0.95}[name]
    # a goal may end in a local minimum in one run and not in the other (rounding-perturbed trajectories):
    # compare rates, not goal by goal
    assert np.mean(f[ref_conv] < 1e-9) >= 0.9
    assert abs(np.mean(f < 1e-9) - np.mean(g["f"] < 1e-9)) <= 0.08
    assert np.mean(pos[ref_ok] < 1e-2) >= 0.9
    assert abs(np.mean(pos < 1e-2) - np.mean(ref_ok)) <= 0.06
    # where both land in the same branch the points agree up to the rigid motion fixed by the base nodes
    if same.any():
        k = int(np.argmax(same))
        Dg = np.linalg.norm(out["x"][k].cpu().numpy()[:, None] - out["x"][k].cpu().numpy()[None], axis=-1)
        Dr = np.linalg.norm(g["Y_sol"][k][:, None] - g["Y_sol"][k][None], axis=-1)
        assert np.max(np.abs(Dg - Dr)) < (1e-3 if robot.n == 6 else 5e-2)


@pytest.mark.parametrize("name", ["ur10", "kuka", "lwa4d", "lwa4p", "panda", "chain20"])
def test_product_pipeline_with_its_own_initialisation_vs_reference_sample(name):
    """The same reference samples, but through the PRODUCT pipeline end to end (BatchIK.solve: goal distances, bound
    smoothing, the kernel's own initialisation, solve, joint recovery) -- nothing of the reference is injected.  The
    kernel's Y_init equals the reference's only up to LAPACK's eigenvector signs, i.e. it is another valid
    draw of the same initialisation rule, so what must agree are the RATES: converged fraction, pose success, and the
    iteration statistics."""
    import os
    from helpers import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, name + "_stats.npz")):
        pytest.skip("reference sample not generated")
    robot, graph, eng = _engine(name)
    g = golden(name + "_stats")
    T = g["T_goal"]
    out = eng.solve(T, check=True)
    f, it = out["f(x)"].cpu().numpy(), out["iterations"].cpu().numpy()
    pos = out["pos_err"].cpu().numpy()
    n = len(f)
    print(name, "own init: converged gpu %.3f ref %.3f; pose success gpu %.3f ref %.3f; median iters gpu %d ref %d"
          % (np.mean(f < 1e-9), np.mean(g["f"] < 1e-9), np.mean(pos < 1e-2), np.mean(g["pose_err"] < 1e-2),
             np.median(it), np.median(g["iterations"])))
    tol = 0.02 + 0.8 / np.sqrt(n)     # two binomial samples of n goals (2 sigma of their difference ~ 0.11 at n = 64)
    assert abs(np.mean(f < 1e-9) - np.mean(g["f"] < 1e-9)) <= tol
    assert abs(np.mean(pos < 1e-2) - np.mean(g["pose_err"] < 1e-2)) <= tol
    r = np.median(it) / np.median(g["iterations"])
    assert 0.8 < r < 1.25, r                  # observed 0.96 - 1.04
    # the joint angles returned reproduce the goal wherever the solve converged to a realisation
    T_sol, _ = eng.fk(out["q"], want_points=False)
    err = np.linalg.norm(T_sol.cpu().numpy()[:, :3, 3] - T[:, :3, 3], axis=1)
    assert np.allclose(err, pos, atol=1e-9)


@pytest.mark.parametrize("name", ["chain20", "kuka_table", "kuka_table_intended"])
def test_end_state_vs_reference_sample_redundant_and_dense(name):
    """BASELINE configs 3 and 4 against the UNMODIFIED reference (tests/golden/<name>_stats.npz from
    oracle/gen_golden_stats.py: 32 goals of the 20-DOF chain, 6 of KUKA + table), GPU started from the reference's own
    Y_init.  Redundant arms have a continuum of solutions per pose and rounding-perturbed trajectories end at different
    points of it, so what is compared is what the reference's stopping rule pins: the EDM residual, convergence, and the
    outer-iteration counts (the CPU oracle differs from the reference by the same amounts).
    "kuka_table_intended": the reference run with the obstacle edges add_spherical_obstacle means to add applied to its
    graph by hand (oracle/gen_golden_stats.py) against obstacle_semantics="intended" here -- one of the
    six goals ends with an active obstacle hinge (f = 9.26e-5), which the GPU must reproduce."""
    import os
    from helpers import GOLDEN, load_kuka_table
    from graphik_b200.engine import BatchIK
    if not os.path.exists(os.path.join(GOLDEN, name + "_stats.npz")):
        pytest.skip("reference sample not generated")
    if name.startswith("kuka_table"):
        robot, graph = load_kuka_table(graph_params={"obstacle_semantics": "intended" if name.endswith("intended")
                                                     else "reference"})
        eng = BatchIK(graph)
    else:
        robot, graph, eng = _engine(name)
    g = golden(name + "_stats")
    T = g["T_goal"]
    out = eng.solve_points(eng.goal_distances(T), g["Y_init"])
    f, it, st = out["f(x)"].cpu().numpy(), out["iterations"].cpu().numpy(), out["status"].cpu().numpy()
    ref_it = g["iterations"]
    print(name, "iterations gpu", it.tolist(), "reference", ref_it.tolist(), "max f gpu %.2e ref %.2e" % (f.max(), g["f"].max()))
    # same end cost as the reference: ~0 where it converged to a realisation, the same local minimum where a hinge stays
    # active (a perturbed trajectory may stall where the reference did not, hence rates)
    assert np.mean((st == 0) & (np.abs(f - g["f"]) <= np.maximum(1e-9, 1e-3 * g["f"]))) >= 0.8
    assert np.median(f) < 1e-13
    r = np.median(it) / np.median(ref_it)
    assert 0.6 < r < 1.6, r
    assert np.median(np.abs(it - ref_it) / ref_it) < 0.3
    if name.startswith("kuka_table"):
        q = eng.joints(out["x"], T).cpu().numpy()
        dq = np.max(np.abs(np.mod(q - g["q_sol"] + np.pi, 2 * np.pi) - np.pi), axis=1)
        print("joint-angle distance to the reference's solution:", np.round(dq, 4).tolist())
        assert np.mean(dq < 0.1) >= 0.5                # 7-DOF: nearby points of the same solution family
