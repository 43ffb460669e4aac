#!/usr/bin/env python
"""Headline benchmark: IK solves/sec on batches of random reachable goal poses.

Workload (BASELINE.json configs[1]): UR10 ProblemGraphRevolute, batch = 4096 random goal poses per GPU,
no obstacles.  One "step" = one pass of the hot path over one batch:
T_goal[B,4,4] resident in HBM -> q[B,n], status[B], f[B] resident in HBM
(goal distances, bound smoothing + initialisation, trust-region solve, joint recovery).

  python bench.py --gpus N --steps K --warmup W      this repo's CUDA path, through the public API only:
        RiemannianSolver.stream() (graphik_b200/pipeline.py) -- K submit() calls, one drain()
  python bench.py --impl reference ...               CPU arm on the box's host cores: the UNMODIFIED reference
        (oracle/_ref/reference, one pinned process per core; kind = "reference") or, where no copy of the
        reference exists, the oracle's C restatement (kind = "port")

Prints ONE JSON line (rank 0).  DESIGN.md section "Measurement" explains every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ik_solves_per_sec"
UNIT = "solves/s"
FP64_PEAK_TFLOPS = 34.2     # FP64 FMA peak measured on this pool's B200 with tools/fp64_peak.cu (profiles/r1d_fp64_peak_b200.txt)


def goals_for(robot, B, seed):
    """Reachable goals: q ~ U(-pi, pi)^n, T = FK(q) (reference README usage)."""
    rng = np.random.RandomState(seed)
    n = robot.n
    lb = np.array([robot.lb["p%d" % i] for i in range(1, n + 1)])
    ub = np.array([robot.ub["p%d" % i] for i in range(1, n + 1)])
    Q = lb + (ub - lb) * rng.rand(B, n)
    return Q, robot.fk_all(Q)[:, n]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING a timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        self.rows = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload)
    return None


def load_workload(name):
    """Robot + graph of a named workload; "kuka_table" = KUKA IIWA + table_environment() obstacles
    with the reference's (anchor-only) obstacle semantics (BASELINE configs[2])."""
    from graphik_b200.utils.roboturdf import load_model
    if name == "kuka_table":
        from graphik_b200.utils.utils import table_environment
        robot, graph = load_model("kuka")
        for k, (c, r) in enumerate(table_environment()):
            graph.add_spherical_obstacle("o%d" % k, c, r)
        return robot, graph
    return load_model(name)


def baseline_label(robot, batch, world):
    """Which BASELINE.json config a (robot, per-GPU batch, GPUs) combination is."""
    if robot == "ur10" and batch == 4096:
        return ", no obstacles (BASELINE configs[1])"
    if robot == "kuka_table" and batch * world == 16384:
        return ", table_environment() obstacles with the reference's semantics (BASELINE configs[2])"
    if robot == "chain20" and batch * world == 65536:
        return ", no obstacles (BASELINE configs[3]: 65536 goals sharded over %d GPUs)" % world
    return ", no obstacles" if robot != "kuka_table" else ", table_environment() obstacles"


# ---------------------------------------------------------------------------------------------------
# CPU arms (oracle/ is measurement infrastructure: only this leg of bench.py executes it)

class CpuPort:
    """The reference algorithm restated on the CPU (oracle/): per goal from_pose distances, bound smoothing (C),
    initialisation (numpy eigh) -- both spread over worker processes --, trust-region solves (C, OpenMP over goals),
    joint recovery.  All host cores."""

    kind = "port"

    def __init__(self, robot, graph):
        from graphik_b200.plan import Plan
        from oracle import ref_arm
        self.robot, self.graph = robot, graph
        self.a = Plan.arrays_from_graph(graph)
        self.cores = len(ref_arm.usable_cores())
        self.cpu = ref_arm.cpu_model()
        self.pool = ref_arm.PortPool(self.cores)
        a, g = self.a, graph
        self.lower = np.where(np.isnan(g.lower), 0.0, g.lower)
        self.upper = np.where(np.isnan(g.upper), np.inf, g.upper)
        gs = a["goal_slot"]
        self.ii, self.jj = np.nonzero(gs >= 0)
        self.slots = gs[self.ii, self.jj]
        self.edge = g.edge.copy()
        self.edge[self.ii, self.jj] = True

    def solve(self, T):
        from oracle import oracle as orc
        a, graph = self.a, self.graph
        B = T.shape[0]
        t0 = time.perf_counter()
        pq = np.stack([T[:, :3, 3], T[:, :3, 3] + graph.axis_length * T[:, :3, 2]], 1)       # [B,2,3]
        d = np.linalg.norm(pq[:, :, None, :] - a["anchor_pos"][None, None], axis=-1)          # [B,2,A]
        goal_d2 = (d ** 2).reshape(B, -1)
        D = np.repeat(a["D_static"][None], B, 0)
        D[:, self.ii, self.jj] = goal_d2[:, self.slots]
        Y0 = self.pool.init_points(self.edge, self.lower, self.upper, self.ii, self.jj, self.slots, goal_d2,
                                   a["omega_f"])
        res = orc.solve_batch(D, a["omega_f"], a["psi_L"], a["psi_U"], Y0, threads=self.cores)
        res["q"] = graph.joint_variables_batch(res["x"], T)
        return time.perf_counter() - t0, res

    def describe(self, sample, steps):
        return ("%d goals/step x %d steps; oracle C port of TrustRegions + tCG (OpenMP over goals, %d threads set from "
                "C), bound smoothing (C) + numpy-eigh initialisation on %d worker processes, joint recovery; %s"
                % (sample, steps, self.cores, self.cores, self.cpu))

    def close(self):
        self.pool.close()


class CpuReference:
    """The unmodified reference through its public API, solve_with_riemannian(graph, T_goal, use_jit=True)
    (riemannian_solver.py:220-234), one pinned process per host core."""

    kind = "reference"

    def __init__(self, workload):
        from oracle import ref_arm
        self.pool = ref_arm.ReferencePool(workload)
        self.cores = len(self.pool.cores)
        self.cpu = ref_arm.cpu_model()

    def solve(self, T):
        dt, errs, ok = self.pool.solve(T)
        return dt, {"pos_err": np.asarray(errs), "ok": np.asarray(ok)}

    def describe(self, sample, steps):
        return ("%d goals/step x %d steps; UNMODIFIED reference solve_with_riemannian(use_jit=True) incl. from_pose, "
                "bound_smoothing, joint_variables; one pinned process per core (%d), numba costgrd AOT-built on this "
                "host, first solve of every process and step discarded; third-party pymanopt/liegroups/urdfpy through "
                "the stand-ins of oracle/shims; %s" % (sample, steps, self.cores, self.cpu))

    def close(self):
        self.pool.close()


def make_cpu_arm(kind, workload, robot, graph):
    from oracle import ref_arm
    if kind in ("auto", "reference") and ref_arm.reference_root() is not None:
        try:
            return CpuReference(workload)
        except Exception as e:   # e.g. numba missing on this host
            if kind == "reference":
                raise
            print("note: reference arm unavailable (%s); using the port" % e, file=sys.stderr)
    return CpuPort(robot, graph)


def run_reference(args):
    """--impl reference: CPU arm.  Rank 0 only; every step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    robot, graph = load_workload(args.robot)
    arm = make_cpu_arm(args.cpu_kind, args.robot, robot, graph)
    sample = args.cpu_sample or (6 * arm.cores if arm.kind == "reference" else 1024)
    times, n_done, errs = [], 0, []
    for s in range(args.warmup + args.steps):
        _, T = goals_for(robot, sample, seed=1000 + s)
        dt, res = arm.solve(T)
        if s >= args.warmup:
            times.append(dt)
            n_done += sample
    arm.close()
    total = float(np.sum(times))
    value = n_done / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s ProblemGraphRevolute, %d random reachable goal poses per step (bounded "
                               "sample of the batch=%d workload)%s" % (args.robot, sample, args.batch,
                                                                       baseline_label(args.robot, args.batch, 1)),
                   "robot": args.robot, "batch": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                         "sample": arm.describe(sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------

def make_stream(solver, n_batches, B, slots, budget, to_host, record_events=False):
    """One IKStream of the public API, its allocations done (stream objects, carry queues, pinned snapshots and the
    caching allocator primed for `n_batches` outstanding batches): set-up, not part of any timed region."""
    st = solver.stream(slots=slots, inner_budget=budget, to_host=to_host, record_events=record_events)
    st.reserve(n_batches, B)    # no cudaMalloc (= device synchronisation) once batches flow
    return st


def run_stream(st, T_list, to_host):
    """K batches through an existing stream: submit all, drain, wall time on the host between two device
    synchronisations.  Returns (seconds, tickets)."""
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tickets = [st.submit(T) for T in T_list]
    st.drain()
    if to_host:
        for tk in tickets:
            tk.ready.synchronize()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, tickets


def measure_stream(solver, T_list, slots, budget, to_host, record_events=False):
    """Set-up + run for the side measurements (configs block): returns (seconds of the run, tickets, stream)."""
    st = make_stream(solver, len(T_list), T_list[0].shape[0], slots, budget, to_host, record_events)
    sec, tickets = run_stream(st, T_list, to_host)
    return sec, tickets, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--robot", default="ur10")
    ap.add_argument("--batch", type=int, default=4096, help="goal poses per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: goal poses per step over ALL GPUs (overrides --batch)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="goals per step of the CPU arm (0: chosen per kind)")
    ap.add_argument("--cpu-kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--slots", type=int, default=1, help="IKStream slots (CUDA streams with a carry queue each)")
    ap.add_argument("--inner-budget", type=int, default=0, help="tCG iterations per goal and launch (0: default)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "latency", "throughput", "generic", "dense"],
                    help="gik_rtr_solve implementation (same results)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs / batch-size sweep")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from graphik_b200.distributed import gather_stats, summary_stats
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    strong = args.global_batch > 0
    if strong:
        args.batch = args.global_batch // world

    robot, graph = load_workload(args.robot)
    solver = RiemannianSolver(graph, {"kernel": args.kernel})
    eng = solver.engine
    budget = args.inner_budget or None
    B, N, n, K = args.batch, graph.number_of_nodes(), robot.n, args.steps
    n_warm = max(args.warmup, 3)
    total_steps = n_warm + K
    # a different goal set per step and per rank; all resident in HBM before the timed region
    T_host = [goals_for(robot, B, seed=1000 + s + 7919 * rank)[1] for s in range(total_steps)]
    T_dev = [torch.as_tensor(T, device=dev).contiguous() for T in T_host]
    T_pinned = [torch.as_tensor(T).pin_memory() for T in T_host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------ set-up + warm-up (untimed): both variants of the stream, the same
    # stream objects the timed regions use (a service keeps its stream; creating one allocates and pins memory)
    st = make_stream(solver, K, B, args.slots, budget, False, record_events=True)
    e2e_st = make_stream(solver, K, B, args.slots, budget, True)
    run_stream(st, T_dev[:n_warm], False)
    run_stream(e2e_st, T_pinned[:n_warm], True)
    barrier()

    # ------------------------------------------------ timed region: K steps, inputs resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    events0, stream_launches0 = len(st.launch_events), st.launches
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    _, tickets = run_stream(st, T_dev[n_warm:], False)
    t_end.record()
    barrier()
    launches = eng.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    local_ms = float(t_begin.elapsed_time(t_end))
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * K / (total_ms * 1e-3)
    outs = [st.result(tk) for tk in tickets]
    rtr_ms = [a.elapsed_time(b) for a, b in st.launch_events[events0:]]
    stream_stats = st.stats()
    stream_stats["launches"] -= stream_launches0

    # summary statistics: the path's single collective (all-gather of a fixed-size vector)
    agg = {k: torch.cat([o[k] for o in outs]) for k in ("iterations", "status", "f(x)", "n_inner")}
    per_rank, stats = gather_stats(summary_stats(agg, local_ms))

    # sharding exactness: every rank re-solves the head of rank 0's first timed batch; one more all-gather
    shard_ok = None
    if world > 1:
        head = min(256, B)
        T0 = torch.as_tensor(goals_for(robot, B, seed=1000 + n_warm)[1][:head], device=dev)
        mine = solver.solve_batch(T0, check=False)["x"].contiguous()
        everyone = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
        shard_ok = bool(all(torch.equal(everyone[0], e) for e in everyone))
        if rank == 0:
            shard_ok = shard_ok and bool(torch.equal(outs[0]["x"][:head], everyone[0]))

    # roofline of the dominant kernel (k_rtr*): algorithmic bytes of the streaming formulation
    it_sum = float(agg["iterations"].sum())
    in_sum = float(agg["n_inner"].sum())
    alg_bytes_total = 72.0 * N * it_sum + 240.0 * N * in_sum
    n_launch = max(len(rtr_ms), 1)
    alg_bytes_per_launch = alg_bytes_total / n_launch
    rtr_avg_ms = float(np.mean(rtr_ms)) if rtr_ms else float("nan")
    achieved = alg_bytes_per_launch / (rtr_avg_ms * 1e-3) / 1e9
    aggregate = alg_bytes_total / (local_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_hbm()
    workload = "%s_b%d" % (args.robot, B)
    alg_flops_total = in_sum * (38.0 * eng.plan.n_terms + 60.0 * N + 60.0)

    # ------------------------------------------------ end to end through the same public API with HOST buffers
    barrier()
    e2e_s, e2e_tickets = run_stream(e2e_st, T_pinned[n_warm:], True)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / float(t.item())
    h2d = B * 16 * 8
    d2h = B * n * 8 + B * 8 + B * 4
    # (the pinned host copies live in a ring of IKStream.HOST_RING buffers per slot: compare the latest batches)
    e2e_same = bool(all(np.array_equal(e2e_st.result(a, host=True)[k].numpy(), o[k].cpu().numpy())
                        for a, o in zip(e2e_tickets[-2:], outs[-2:]) for k in ("status", "q", "f(x)")))

    # ------------------------------------------------ one batch at a time through solve_batch (synchronous call)
    serial_ms = []
    for s in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.solve_batch(T_dev[s], check=False)
        torch.cuda.synchronize()
        serial_ms.append(1e3 * (time.perf_counter() - t0))
    # ... and one batch at a time through a ONE-slot stream (the default is one slot: then this is `value` itself)
    if args.slots == 1:
        one_value = value
    else:
        one_s, _, _ = measure_stream(solver, T_dev[n_warm:n_warm + min(K, 16)], 1, budget, False)
        one_value = world * B * min(K, 16) / one_s
    barrier()

    # ------------------------------------------------ other BASELINE configs + batch-size sweep (N = 1 only)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        configs = measure_configs(args, dev)

    # ------------------------------------------------ CPU baselines (rank 0, N = 1 only)
    cpu, cpu_ref = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, cpu_ref = cpu_baselines(args, robot, graph, n_warm)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "warmup_actual": n_warm, "ms_per_step": total_ms / K,
            "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s ProblemGraphRevolute, batch=%d random reachable goal poses per GPU%s"
                                   % (args.robot, B, baseline_label(args.robot, B, world)),
                       "robot": args.robot, "batch_per_gpu": B, "nodes": N, "cost_terms": eng.plan.n_terms,
                       "api": "RiemannianSolver.stream(): K submit() calls + drain()",
                       "stream_slots": args.slots, "inner_budget": stream_stats["inner_budget"],
                       "rtr_kernel": args.kernel,
                       "l2": "not flushed: every step reads its own goal set (%d steps x %.1f MB of inputs, outputs and "
                             "parked states) and batches on different slots evict each other" %
                             (K, B * (16 + 2 * 3 * N + n + 4) * 8 / 1e6),
                       "parallelism": "goals sharded, dp%d" % world},
            "serial": {"value": world * B / (float(np.mean(serial_ms)) * 1e-3), "unit": UNIT,
                       "ms_per_batch": float(np.mean(serial_ms)),
                       "note": "one synchronous RiemannianSolver.solve_batch call at a time: its latency is the "
                               "slowest goal of the batch (maxiter = 3000: ~230 k tCG iterations at the ~0.6 us a lone "
                               "warp needs per iteration), which no scheduling can shorten without changing results"},
            "serial_deferred": {"value": one_value, "unit": UNIT, "slots": 1,
                                "note": "one batch at a time through a ONE-slot stream (one CUDA stream, launches strictly "
                                        "one after the other): goals that are still iterating when a launch runs out of "
                                        "new goals park and finish in the shadow of the following batches; one drain at "
                                        "the end, inside the timed region"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "same_results_as_device_run": e2e_same},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_rtr_fast / k_rtr_fast2 / k_rtr_cta (persistent trust-region solve)",
                         # launches of different slots overlap on the device, so the launch duration that matters
                         # for the roofline is the timed region divided by the launches it retired; the raw
                         # CUDA-event duration of one launch (which includes time-sharing the SMs) is per_launch_*
                         "achieved": aggregate, "peak": peak, "unit": "GB/s", "frac": aggregate / peak,
                         "per_launch_achieved": achieved, "per_launch_frac": achieved / peak,
                         "peak_source": peak_src, "traffic": ncu_traffic(workload),
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launches": n_launch,
                         "kernel_ms": rtr_avg_ms, "effective_kernel_ms": local_ms / n_launch,
                         "kernel_share_of_region": float(np.sum(rtr_ms)) / (local_ms * max(args.slots, 1)),
                         "fp64": {"achieved": alg_flops_total / (local_ms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS,
                                  "unit": "TFLOP/s",
                                  "frac": alg_flops_total / (local_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                                  "note": "algorithmic flops (38 per term + 60 N + 60 per tCG iteration) / region "
                                          "time vs the measured FP64 FMA peak"},
                         "tcg_iterations_per_s": in_sum / (local_ms * 1e-3),
                         "note": "algorithmic bytes = sum over problems of 72N*outer + 240N*inner: the state a "
                                 "kernel-per-iteration formulation streams through HBM (SURVEY 8d), from the iteration "
                                 "counts actually executed; the persistent kernel keeps that state in registers (traffic "
                                 "= real DRAM bytes per launch from ncu), so the binding resources are FP64 issue and "
                                 "dependent-instruction latency (profiles/)"},
            "stream": stream_stats,
            "shard_bit_identical": shard_ok,
            "cpu_baseline": cpu,
            "cpu_baseline_reference": cpu_ref,
            "clocks": clocks,
            "stats": {"converged_frac": stats["converged"] / max(stats["count"], 1),
                      "mean_outer_iters": stats["sum_outer"] / max(stats["count"], 1),
                      "mean_inner_iters": stats["sum_inner"] / max(stats["count"], 1),
                      "max_outer_iters": stats["max_outer"], "mean_f": stats["sum_f"] / max(stats["count"], 1)},
            "configs": configs,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_configs(args, dev):
    """The other BASELINE.json configs that fit one GPU and the 1k..64k batch-size sweep of north_star, each measured
    like the headline (stream of batches, inputs resident in HBM, host wall time between device synchronisations) and
    each with its own clock record."""
    import torch
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver
    plan = [  # (label, robot, batch, steps)
        ("BASELINE configs[2]: KUKA-IIWA + table_environment(), 16384 goals, obstacle semantics = reference "
         "(obstacles are anchors only, SURVEY App. C.1)", "kuka_table", 16384, 1),
        ("BASELINE configs[3] on one GPU: 20-DOF chain, 65536 goals", "chain20", 65536, 2),
        ("sweep ur10 1k", "ur10", 1024, 16), ("sweep ur10 16k", "ur10", 16384, 4), ("sweep ur10 64k", "ur10", 65536, 2),
        ("sweep kuka 1k", "kuka", 1024, 16), ("sweep kuka 4k", "kuka", 4096, 8), ("sweep kuka 16k", "kuka", 16384, 4),
        ("sweep kuka 64k", "kuka", 65536, 2),
    ]
    out = []
    for label, name, B, K in plan:
        robot, graph = load_workload(name)
        solver = RiemannianSolver(graph, {"kernel": args.kernel})
        eng = solver.engine
        N = graph.number_of_nodes()
        Ts = [torch.as_tensor(goals_for(robot, B, seed=5000 + s)[1], device=dev) for s in range(K)]
        warm = torch.as_tensor(goals_for(robot, min(B, 512), seed=4999)[1], device=dev)
        measure_stream(solver, [warm], args.slots, None, False)
        sampler = ClockSampler(dev.index).start()
        dt, tickets, st = measure_stream(solver, Ts, args.slots, None, False)
        clocks = sampler.stop()
        res = [st.result(t) for t in tickets]
        inner = float(sum(r["n_inner"].sum() for r in res))
        outer = float(sum(r["iterations"].sum() for r in res))
        conv = float(sum((r["status"] == 0).sum() for r in res)) / (B * K)
        alg = 72.0 * N * outer + 240.0 * N * inner
        peak, _ = measured_peak_hbm()
        flops = inner * (38.0 * eng.plan.n_terms + 60.0 * N + 60.0)
        out.append({"workload": label, "robot": name, "batch": B, "steps": K, "nodes": N, "value": B * K / dt,
                    "unit": UNIT, "seconds": dt, "converged_frac": conv, "streaming_frac": alg / dt / 1e9 / peak,
                    "fp64_frac": flops / dt / 1e12 / FP64_PEAK_TFLOPS, "tcg_iterations_per_s": inner / dt,
                    "clocks": clocks})
        del solver, eng, Ts, res, tickets, st
        torch.cuda.empty_cache()
    out.extend(measure_cidgik(dev))
    return out


def measure_cidgik(dev, B=1024, K=8):
    """BASELINE configs[4]: CIDGIK (solve_with_cidgik, convex_iteration.py:279-319) on UR10, 1024 goals per batch,
    `ranges=True` as the reference calls it.  Two entries: (1) the graph as the reference builds it, for which
    `ranges=True` yields no inequality (sdp_snl.py:383-385 looks at obstacle pairs only and the reference's obstacles are
    anchors without robot edges); (2) obstacle_semantics="intended" with one sphere in the workspace:
    every free joint point carries a lower-bound inequality.  The SDP solver is this repo's interior-point kernel, not
    MOSEK: parity unpinned; `pose_reached_frac` / `clear_of_obstacle_frac` are the solver-independent
    checks.  `cpu_port` times the numpy statement of the same algorithm on one host core."""
    import torch
    from graphik_b200.solvers.convex_iteration import solve_batch_with_cidgik
    from graphik_b200.utils.roboturdf import load_model
    entries = []
    centre, radius = np.array([0.3, 0.3, 0.2]), 0.3
    for with_sphere in (False, True):
        if with_sphere:
            robot, graph = load_model("ur10", graph_params={"obstacle_semantics": "intended"})
            graph.add_spherical_obstacle("o0", centre, radius)
        else:
            robot, graph = load_workload("ur10")
        n = robot.n
        Tn = [goals_for(robot, B, seed=7000 + s)[1] for s in range(K)]
        Ts = [torch.as_tensor(T, device=dev) for T in Tn]
        solve_batch_with_cidgik(graph, Ts[0])
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(dev.index).start()
        # this path is ~170 short launches per batch: let nvidia-smi finish starting up (it holds driver locks that
        # delay launches while it initialises) before the timed region begins
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.01)
        t0 = time.perf_counter()
        res = [solve_batch_with_cidgik(graph, T) for T in Ts]
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        ok, iters, launches, feas, clear, rank3 = [], [], 0, [], [], []
        pid = [graph.idx("p%d" % i) for i in range(1, n)]
        for T, r in zip(Tn, res):
            Tq = robot.fk_all(r["q"].cpu().numpy())[:, n]
            ok.append((np.linalg.norm(Tq[:, :3, 3] - T[:, :3, 3], axis=1) < 1e-2) &
                      (np.abs(Tq[:, :3, :3] - T[:, :3, :3]).max(axis=(1, 2)) < 1e-2))
            iters.append(r["n_iters"].cpu().numpy())
            feas.append(r["feasible"].cpu().numpy())
            clear.append(np.linalg.norm(r["x"].cpu().numpy()[:, pid] - centre, axis=-1).min(axis=1) > radius - 1e-3)
            vals = r["values"].cpu().numpy()
            rank3.append(np.array([v[~np.isnan(v)][-1] < 1e-6 if np.any(~np.isnan(v)) else False for v in vals]))
            launches += int(r["launches"])
        ok, feas, clear, rank3 = np.concatenate(ok), np.concatenate(feas), np.concatenate(clear), np.concatenate(rank3)
        e = {"workload": "BASELINE configs[4]: CIDGIK (SDP relaxation + convex iteration), UR10, 1024 goals per batch, "
                         "ranges=True; own interior-point SDP kernel, parity with MOSEK unpinned; " +
                         ("obstacle_semantics=intended, one sphere r = 0.3 m: 5 lower-bound inequalities per program"
                          if with_sphere else "graph as the reference builds it: no inequality arises (SURVEY App. C.1)"),
             "robot": "ur10", "batch": B, "steps": K, "value": B * K / dt, "unit": UNIT, "seconds": dt,
             "feasible_frac": float(np.mean(feas == 0)),
             "pose_reached_frac": float(np.mean(ok[feas == 0])),
             "rank3_frac": float(np.mean(rank3[feas == 0])),     # convex iteration ended with excess rank < 1e-6
             "convex_iterations_mean": float(np.mean(np.concatenate(iters))), "gpu_launches": launches,
             "clocks": clocks}
        if with_sphere:
            e["clear_of_obstacle_frac_where_rank3"] = float(np.mean(clear[(feas == 0) & rank3]))
        entries.append(e)
    robot, graph = load_workload("ur10")
    Tn = [goals_for(robot, B, seed=7000)[1]]
    # CPU: the numpy statement of the same algorithm (oracle/cidgik.py), 4 goals, one core
    from oracle import cidgik as cg
    from graphik_b200.solvers.convex_iteration import CidgikPlan
    plan = CidgikPlan(graph)
    an, W, b, V = [t.numpy() for t in plan.assemble(Tn[0][:4], device="cpu")]
    t0 = time.perf_counter()
    for k in range(4):
        cg.convex_iterate(graph.node_ids, graph.dist, {u: an[k, i] for i, u in enumerate(plan.anchor_names)},
                          coordinates=(W[k], b[k], V[k]))
    cpu = 4 / (time.perf_counter() - t0)
    entries[0]["cpu_port"] = {"value": cpu, "unit": UNIT, "cores": 1, "kind": "port",
                              "sample": "4 goals, numpy statement of the same algorithm"}
    return entries


def cpu_baselines(args, robot, graph, n_warm):
    """cpu_baseline (port) and cpu_baseline_reference (the unmodified reference) on bounded samples of the first timed
    batch, timed on this box's host cores."""
    from oracle import ref_arm
    cpu, cpu_ref = None, None
    port = CpuPort(robot, graph)
    sample = args.cpu_sample or 1024
    _, Tc = goals_for(robot, sample, seed=1000 + n_warm)
    port.solve(Tc[:64])                                # warm the OpenMP pool and the workers
    dt, res = port.solve(Tc)
    cpu = {"value": sample / dt, "unit": UNIT, "cores": port.cores, "kind": "port",
           "sample": "first " + port.describe(sample, 1) + "; median outer iters %d" % int(np.median(res["iterations"]))}
    port.close()
    if args.cpu_kind != "port" and ref_arm.reference_root() is not None:
        try:
            ref = CpuReference(args.robot)
            sample = 4 * ref.cores
            _, Tr = goals_for(robot, sample, seed=1000 + n_warm)
            dt, res = ref.solve(Tr)
            cpu_ref = {"value": sample / dt, "unit": UNIT, "cores": ref.cores, "kind": "reference",
                       "sample": "first " + ref.describe(sample, 1) + "; pose error < 1e-2 m on %.2f of the goals"
                                 % float(np.mean(res["pos_err"] < 1e-2))}
            ref.close()
        except Exception as e:
            cpu_ref = {"unavailable": str(e)[:200]}
    return cpu, cpu_ref


if __name__ == "__main__":
    main()
