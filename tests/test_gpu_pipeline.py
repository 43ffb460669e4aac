"""Deferred stragglers (gik_rtr_solve_sliced) and the pipelined IKStream: a goal that parks and resumes must
follow BIT FOR BIT the trajectory of a goal solved in one piece (trust_region.py:179-422 carries only x, the
gradient, f, Delta and the counters from one outer iteration to the next), for every kernel that can park."""
import ctypes

import numpy as np
import pytest

from helpers import load_kuka_table, load_robot, random_goals

pytestmark = pytest.mark.gpu

KEYS = ("x", "f(x)", "gradnorm", "iterations", "status", "n_inner")


def _sliced_to_completion(eng, g2, Y0, budget, capacity, max_launches=4000):
    """Drive gik_rtr_solve_sliced by hand through the raw ABI: first launch with the batch, then launches that
    only resume, alternating the two queues, until the batch's pending counter is back to zero."""
    import torch
    from graphik_b200 import _lib
    from graphik_b200.engine import _p
    lib, dev = eng.lib, eng.device
    B, N = Y0.shape[0], eng.plan.N
    nbytes = int(lib.gik_carry_bytes(eng.plan.handle, capacity))
    carry = [torch.zeros((nbytes + 15) // 16 * 2, dtype=torch.int64, device=dev) for _ in range(2)]
    for c in carry:
        _lib.check(lib.gik_carry_init(eng.plan.handle, _p(c), capacity, eng._stream()), "gik_carry_init")
    out = {"x": torch.empty(B, N, 3, dtype=torch.float64, device=dev),
           "f(x)": torch.empty(B, dtype=torch.float64, device=dev),
           "gradnorm": torch.empty(B, dtype=torch.float64, device=dev),
           "iterations": torch.empty(B, dtype=torch.int32, device=dev),
           "status": torch.empty(B, dtype=torch.int32, device=dev),
           "n_inner": torch.empty(B, dtype=torch.int32, device=dev)}
    pending = torch.zeros(1, dtype=torch.int32, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    flip, launches, parked_first = 0, 0, None
    while True:
        first = launches == 0
        _lib.check(lib.gik_rtr_solve_sliced(
            eng.plan.handle, _p(g2) if first else None, _p(Y0) if first else None, B if first else 0,
            ctypes.byref(eng.opts), _p(out["x"]), _p(out["f(x)"]), _p(out["gradnorm"]), _p(out["iterations"]),
            _p(out["status"]), _p(out["n_inner"]), budget, _p(carry[flip]), _p(carry[1 - flip]), _p(pending),
            _p(counter), eng._stream()), "gik_rtr_solve_sliced")
        flip ^= 1
        launches += 1
        left = int(pending.item())
        if first:
            parked_first = left
            assert int((out["status"] == 4).sum()) == left     # parked goals read GIK_STATUS_PENDING meanwhile
        if left == 0:
            break
        assert launches < max_launches
    full = sum(int(c.view(torch.int32)[3]) for c in carry)
    return out, launches, parked_first, full


def _assert_identical(a, b):
    for k in KEYS:
        assert (a[k] == b[k]).all(), "%s differs between a sliced and a one-piece solve" % k


@pytest.mark.parametrize("robot_name,kernel,B,budget", [
    ("ur10", "latency", 768, 1500),      # k_rtr_fast<2,5>: throughput variant for the bulk launch, latency variant to drain
    ("ur10", "throughput", 768, 1500),   # k_rtr_duo<9>, two problems per warp
    ("ur10", "auto", 768, 1500),         # k_rtr_duo for the bulk launch, k_rtr_fast (latency variant) to drain
    ("lwa4p", "auto", 256, 1500),
    ("kuka", "latency", 512, 3000),      # k_rtr_fast<1,9>, register slot cache
    ("ur10", "generic", 256, 1500),      # k_rtr<16,1>
    ("chain20", "latency", 192, 800),    # k_rtr_fast2<7,5>
])
def test_sliced_solve_is_bit_identical(robot_name, kernel, B, budget):
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot(robot_name)
    eng = BatchIK(graph, params={"kernel": kernel})
    _, T = random_goals(robot, B, seed=77)
    g2 = eng.goal_distances(T)
    Y0 = eng.initialization(g2)
    ref = eng.solve_points(g2, Y0)
    out, launches, parked, full = _sliced_to_completion(eng, g2, Y0, budget, capacity=B)
    assert parked > 0 and launches > 2 and full == 0             # the test does exercise parking
    _assert_identical(out, ref)
    assert int((out["status"] == 4).sum()) == 0


def test_kernels_for_small_graphs_agree_bit_for_bit():
    """k_rtr_fast (throughput and latency variants) and k_rtr_duo share their arithmetic (gik_tr_math.cuh, explicitly
    rounded operations): the same problems give the same bits whichever of them runs, so results do not depend on
    the batch size that selected the kernel."""
    import torch
    from graphik_b200.engine import BatchIK
    for name in ("ur10", "lwa4p"):
        robot, graph = load_robot(name)
        _, T = random_goals(robot, 640, seed=31)
        outs = []
        for kernel in ("latency", "throughput"):
            eng = BatchIK(graph, params={"kernel": kernel})
            outs.append(eng.solve(T, check=False))
        _assert_identical(outs[0], outs[1])
        assert (outs[0]["q"] == outs[1]["q"]).all()
        # the latency VARIANT (small batch) against the throughput variant (same goals inside a batch > 32768)
        eng = BatchIK(graph, params={"kernel": "latency", "maxiter": 40})
        g2 = eng.goal_distances(T)
        Y0 = eng.initialization(g2)
        small = eng.solve_points(g2, Y0)
        reps = 33280 // 640
        big = eng.solve_points(g2.repeat(reps, 1), Y0.repeat(reps, 1, 1))
        for k in KEYS:
            assert (big[k][:640] == small[k]).all() and (big[k][-640:] == small[k]).all(), k


def test_sliced_solve_full_queue_runs_on():
    """A goal that finds the outgoing queue full keeps iterating; nothing is lost, results unchanged."""
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot("ur10")
    eng = BatchIK(graph)
    _, T = random_goals(robot, 512, seed=5)
    g2 = eng.goal_distances(T)
    Y0 = eng.initialization(g2)
    ref = eng.solve_points(g2, Y0)
    out, launches, parked, full = _sliced_to_completion(eng, g2, Y0, 1000, capacity=8)
    assert parked <= 8 and full > 0
    _assert_identical(out, ref)


def test_sliced_dense_kernel_is_bit_identical():
    """k_rtr_cta (KUKA + table obstacles, N = 118): CTA-wide park / resume."""
    from graphik_b200.engine import BatchIK
    robot, graph = load_kuka_table()
    eng = BatchIK(graph, params={"maxiter": 60})
    _, T = random_goals(robot, 24, seed=3)
    g2 = eng.goal_distances(T)
    Y0 = eng.initialization(g2)
    ref = eng.solve_points(g2, Y0)
    out, launches, parked, full = _sliced_to_completion(eng, g2, Y0, 400, capacity=64)
    assert parked > 0 and launches > 2
    _assert_identical(out, ref)


def test_ikstream_matches_batch_solve_bitwise():
    """IKStream over several batches with a budget small enough that many goals finish in later launches:
    every result equals BatchIK.solve of the same batch, including the joint angles recovered afterwards."""
    import torch
    from graphik_b200.engine import BatchIK
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver
    robot, graph = load_robot("ur10")
    solver = RiemannianSolver(graph)
    batches = [random_goals(robot, 384, seed=100 + k)[1] for k in range(7)]
    stream = solver.stream(slots=2, inner_budget=2000, to_host=True)
    tickets = [stream.submit(torch.as_tensor(T).pin_memory()) for T in batches]
    results = [stream.result(t) for t in tickets]
    hosts = [stream.result(t, host=True) for t in tickets]
    eng = BatchIK(graph)
    for T, res, h in zip(batches, results, hosts):
        ref = eng.solve(T, check=False)
        _assert_identical(res, ref)
        assert (res["q"] == ref["q"]).all()
        assert int((res["status"] == 4).sum()) == 0
        assert np.array_equal(h["q"].numpy(), ref["q"].cpu().numpy())
        assert np.array_equal(h["status"].numpy(), ref["status"].cpu().numpy())
    st = stream.stats()
    assert st["queue_full_events"] == 0 and st["launches"] >= 3 * len(batches)
    # a drained stream can be used again
    t = stream.submit(batches[0])
    stream.drain()
    _assert_identical(stream.result(t), eng.solve(batches[0], check=False))


def test_solve_batches_convenience_and_kuka_table_workspace():
    """solve_batches on a plan whose bound-smoothing kernel needs a workspace (N = 118): each slot owns one."""
    from graphik_b200.engine import BatchIK
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver
    robot, graph = load_kuka_table()
    solver = RiemannianSolver(graph, {"maxiter": 30})
    batches = [random_goals(robot, 12, seed=40 + k)[1] for k in range(4)]
    results = solver.solve_batches(batches, slots=2, inner_budget=300)
    eng = BatchIK(graph, params={"maxiter": 30})
    for T, res in zip(batches, results):
        ref = eng.solve(T, check=False)
        assert (res["Y_init"] == eng.initialization(ref["goal_d2"])).all()
        _assert_identical(res, ref)


def test_shared_plan_on_two_streams_is_race_free():
    """Two engines sharing ONE plan run bound smoothing + initialisation + solve concurrently on two streams
    (N = 118 needs the per-call workspace that used to live in the plan): results equal the serial ones."""
    import torch
    from graphik_b200.engine import BatchIK
    robot, graph = load_kuka_table()
    e0 = BatchIK(graph, params={"maxiter": 20})
    e1 = BatchIK(plan=e0.plan, params={"maxiter": 20})
    Ts = [random_goals(robot, 160, seed=9 + k)[1] for k in range(2)]
    serial = [e0.solve(T, check=False) for T in Ts]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(2)]
    conc = []
    for rep in range(3):
        conc = []
        for e, T, s in zip((e0, e1), Ts, streams):
            with torch.cuda.stream(s):
                conc.append(e.solve(T, check=False))
        torch.cuda.synchronize()
        for a, b in zip(conc, serial):
            assert (a["Y_init"] == b["Y_init"]).all() if "Y_init" in a else True
            _assert_identical(a, b)


def test_one_engine_two_streams():
    """One engine used from two streams at once: the work-queue counter is per stream."""
    import torch
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot("ur10")
    eng = BatchIK(graph)
    Ts = [random_goals(robot, 2048, seed=60 + k)[1] for k in range(2)]
    serial = [eng.solve(T, check=False) for T in Ts]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(2)]
    conc = []
    for T, s in zip(Ts, streams):
        with torch.cuda.stream(s):
            conc.append(eng.solve(T, check=False))
    torch.cuda.synchronize()
    for a, b in zip(conc, serial):
        _assert_identical(a, b)


def test_maxtime_stops_first_and_status_codes():
    """pymanopt's Solver checks maxtime before maxiter and mingradnorm: with a limit that has
    already passed after the first outer iteration every goal stops there with its own status code."""
    from graphik_b200.engine import BatchIK
    robot, graph = load_robot("ur10")
    _, T = random_goals(robot, 64, seed=8)
    for kernel in ("latency", "generic"):
        eng = BatchIK(graph, params={"maxtime": 1e-9, "kernel": kernel})
        out = eng.solve(T, check=False)
        assert (out["status"] == 5).all() and (out["iterations"] == 1).all()
    eng = BatchIK(graph, params={"maxtime": 0.0})          # <= 0: no time limit
    assert int((eng.solve(T, check=False)["status"] == 5).sum()) == 0
    assert BatchIK(graph).opts.maxtime == 1000.0             # the reference's (pymanopt's) default
