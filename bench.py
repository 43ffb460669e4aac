#!/usr/bin/env python
"""Headline benchmark: IK solves/sec on batches of random reachable goal poses.

Workload (BASELINE.json configs[1]): UR10 ProblemGraphRevolute, batch = 4096 random goal
poses per GPU, no obstacles.  One "step" = one pass of the hot path over one batch:
T_goal[B,4,4] resident in HBM -> q[B,n], status[B], f[B] resident in HBM
(goal distances, bound smoothing + initialisation, trust-region solve, joint recovery).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference ...                     CPU arm: the oracle port of the
        reference's algorithm on all host cores (the reference itself is Python and cannot
        travel to the GPU box; kind = "port")

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
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


def goals_for(robot, B, seed):
    """Reachable goals: q ~ U(-pi, pi)^n, T = FK(q) (reference README usage)."""
    rng = np.random.RandomState(seed)
    n = robot.n
    lb = np.array([robot.lb["p%d" % i] for i in range(1, n + 1)])
    ub = np.array([robot.ub["p%d" % i] for i in range(1, n + 1)])
    Q = lb + (ub - lb) * rng.rand(B, n)
    return Q, robot.fk_all(Q)[:, n]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

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


def oracle_inputs(eng_arrays, goal_d2, Y0):
    """Per-goal D_goal matrices for the CPU port from the plan's static arrays + goal rows."""
    a = eng_arrays
    B = goal_d2.shape[0]
    D = np.repeat(a["D_static"][None], B, 0)
    gs = a["goal_slot"]
    ii, jj = np.nonzero(gs >= 0)
    D[:, ii, jj] = goal_d2[:, gs[ii, jj]]
    return D


def cpu_port_solve(robot, graph, T, threads):
    """The reference algorithm restated on the CPU (oracle/): per goal from_pose distances,
    bound smoothing, initialisation (numpy eigh), trust-region solve (C, OpenMP over goals),
    joint recovery.  Returns (elapsed seconds, results)."""
    from graphik_b200.plan import Plan
    from oracle import oracle as orc
    a = Plan.arrays_from_graph(graph)
    B = T.shape[0]
    anchors = a["anchor_pos"]
    t0 = time.perf_counter()
    pq = np.stack([T[:, :3, 3], T[:, :3, 3] + graph.axis_length * T[:, :3, 2]], 1)       # [B,2,3]
    d = np.linalg.norm(pq[:, :, None, :] - anchors[None, None], axis=-1)                  # [B,2,A]
    goal_d2 = (d ** 2).reshape(B, -1)
    D = oracle_inputs(a, goal_d2, None)
    lower = np.where(np.isnan(graph.lower), 0.0, graph.lower)
    upper = np.where(np.isnan(graph.upper), np.inf, graph.upper)
    edge = graph.edge.copy()
    gs = a["goal_slot"]
    ii, jj = np.nonzero(gs >= 0)
    edge[ii, jj] = True
    Y0 = np.empty((B, graph.number_of_nodes(), 3))
    for b in range(B):
        lo, up = lower.copy(), upper.copy()
        lo[ii, jj] = up[ii, jj] = np.sqrt(goal_d2[b, gs[ii, jj]])
        lb, ub = orc.bound_smoothing(edge, lo, up)
        Y0[b] = orc.generate_initialization(lb, ub, a["omega_f"])
    res = orc.solve_batch(D, a["omega_f"], a["psi_L"], a["psi_U"], Y0, threads=threads)
    res["q"] = graph.joint_variables_batch(res["x"], T)
    return time.perf_counter() - t0, res


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


def run_reference(args):
    """--impl reference: CPU arm.  Rank 0 only; bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    robot, graph = load_workload(args.robot)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    times, n_done = [], 0
    for s in range(args.warmup + args.steps):
        _, T = goals_for(robot, sample, seed=1000 + s)
        dt, res = cpu_port_solve(robot, graph, T, cores)
        if s >= args.warmup:
            times.append(dt)
            n_done += sample
    total = float(np.sum(times))
    value = n_done / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s ProblemGraphRevolute, %d random reachable goal poses per step (bounded "
                               "sample of the batch=%d workload), no obstacles" % (args.robot, sample, args.batch),
                   "robot": args.robot, "batch": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d goals/step x %d steps; oracle C port of TrustRegions+tCG with OpenMP over "
                                   "goals, numpy eigh initialisation, C bound smoothing" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--robot", default="ur10")
    ap.add_argument("--batch", type=int, default=4096, help="goal poses per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="goals per step of the CPU arm / cpu_baseline")
    ap.add_argument("--concurrent", type=int, default=8,
                    help="batches in flight (CUDA streams); 1 = strictly one batch at a time")
    ap.add_argument("--kernel", default="auto", choices=["auto", "latency", "throughput", "generic", "dense"],
                    help="gik_rtr_solve implementation (same results)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from graphik_b200.distributed import gather_stats, summary_stats
    from graphik_b200.engine import BatchIK
    from graphik_b200.utils.roboturdf import load_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    robot, graph = load_workload(args.robot)
    C = max(1, args.concurrent)              # batches in flight (one CUDA stream + work counter each)
    eng = BatchIK(graph, params={"kernel": args.kernel}, device=dev)
    engs = [eng] + [BatchIK(plan=eng.plan, params={"kernel": args.kernel}, device=dev) for _ in range(C - 1)]   # one work counter per slot
    streams = [torch.cuda.Stream(device=dev) for _ in range(C)]
    B, N, n = args.batch, graph.number_of_nodes(), robot.n
    n_warm = max(args.warmup, C)             # every slot/stream is warmed at least once (untimed)
    total_steps = n_warm + args.steps
    # a different goal set per step and per rank; all resident in HBM before the timed region
    T_host = [goals_for(robot, B, seed=1000 + s + 7919 * rank)[1] for s in range(total_steps)]
    T_dev = [torch.as_tensor(T, device=dev).contiguous() for T in T_host]
    T_pinned = [torch.as_tensor(T).pin_memory() for T in T_host]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(e, T):
        """One pass of the hot path with inputs resident in HBM (4 kernels + 1 memset)."""
        g2 = e.goal_distances(T)
        Y0 = e.initialization(g2)
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        out = e.solve_points(g2, Y0)
        ev1.record()
        out["q"] = e.joints(out["x"], T)
        return out, ev0, ev1

    # ------------------------------------------------ warm-up: every slot / stream once (untimed)
    for s in range(n_warm):
        c = s % C
        with torch.cuda.stream(streams[c]):
            device_step(engs[c], T_dev[s])
        streams[c].synchronize()
    barrier()

    # ------------------------------------------------ timed region: K steps, C batches in flight
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(e.launches for e in engs)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    outs, ev_rtr, ev_step = [], [], []
    t_begin.record()
    for st in streams:
        st.wait_event(t_begin)
    for s in range(args.steps):
        c = s % C
        with torch.cuda.stream(streams[c]):
            flush.zero_() if C == 1 else None      # L2 flush between serial steps; with C > 1 the
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()                              # concurrent batches evict each other's lines
            out, k0, k1 = device_step(engs[c], T_dev[n_warm + s])
            e1.record()
        outs.append(out)
        ev_rtr.append((k0, k1))
        ev_step.append((e0, e1))
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    t_end.record()
    barrier()
    launches = sum(e.launches for e in engs) - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev_step]
    rtr_ms = [a.elapsed_time(b) for a, b in ev_rtr]
    local_ms = float(t_begin.elapsed_time(t_end))
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # summary statistics: the path's single collective (all-gather of a fixed-size vector)
    agg = {k: torch.cat([o[k] for o in outs]) for k in ("iterations", "status", "f(x)", "n_inner")}
    per_rank, stats = gather_stats(summary_stats(agg, local_ms))

    # roofline of the dominant kernel (k_rtr*): algorithmic bytes of the streaming formulation
    it_sum = float(agg["iterations"].sum())
    in_sum = float(agg["n_inner"].sum())
    alg_bytes_total = 72.0 * N * it_sum + 240.0 * N * in_sum
    alg_bytes_per_launch = alg_bytes_total / args.steps
    rtr_avg_ms = float(np.mean(rtr_ms))
    achieved = alg_bytes_per_launch / (rtr_avg_ms * 1e-3) / 1e9
    aggregate = alg_bytes_total / (local_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_hbm()
    workload = "%s_b%d" % (args.robot, B)
    # secondary bound: algorithmic fp64 flops per tCG iteration = 38 per cost term + 60 N + 60,
    # against the FP64 FMA peak measured on this pool's B200 with tools/fp64_peak.cu (profiles/r1d_fp64_peak_b200.txt)
    alg_flops_total = in_sum * (38.0 * eng.plan.n_terms + 60.0 * N + 60.0)
    fp64_peak_tflops = 34.2

    # ------------------------------------------------ end to end through the public API (host buffers)
    from graphik_b200.solvers.riemannian_solver import RiemannianSolver
    solvers = []
    for e in engs:
        sv = RiemannianSolver(graph, {"kernel": args.kernel})
        sv._engine = e
        solvers.append(sv)
    h_out = [(torch.empty((B, n), dtype=torch.float64).pin_memory(), torch.empty((B,), dtype=torch.float64).pin_memory(),
              torch.empty((B,), dtype=torch.int32).pin_memory()) for _ in range(C)]

    def e2e_step(s, c):
        streams[c].synchronize()                     # the consumer has taken slot c's previous result
        with torch.cuda.stream(streams[c]):
            Tg = T_pinned[s].to(dev, non_blocking=True)
            o = solvers[c].solve_batch(Tg, check=False)
            h_out[c][0].copy_(o["q"], non_blocking=True)
            h_out[c][1].copy_(o["f(x)"], non_blocking=True)
            h_out[c][2].copy_(o["status"], non_blocking=True)

    for s in range(C):
        e2e_step(s, s % C)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        e2e_step(n_warm + s, s % C)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    h2d = B * 16 * 8
    d2h = B * n * 8 + B * 8 + B * 4

    # ------------------------------------------------ serial view: one batch at a time, everything warm.  Reported as
    # `serial` (latency of a single solve_batch call) and as the trust-region kernel's share of a step, which is
    # comparable with ncu's serialised launch list
    serial_ms, serial_rtr_ms = [], []
    for s in range(min(3, n_warm)):
        with torch.cuda.stream(streams[0]):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            _, w0, w1 = device_step(engs[0], T_dev[s])
            a1.record()
        a1.synchronize()
        serial_ms.append(a0.elapsed_time(a1))
        serial_rtr_ms.append(w0.elapsed_time(w1))
    barrier()

    # ------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        _, Tc = goals_for(robot, args.cpu_sample, seed=1000 + n_warm)
        cpu_port_solve(robot, graph, Tc[:16], cores)       # warm the OpenMP pool / page in numpy
        dt, res = cpu_port_solve(robot, graph, Tc, cores)
        cpu = {"value": args.cpu_sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d goals of the first timed batch; oracle C port (OpenMP over goals) incl. "
                         "bound smoothing + numpy-eigh initialisation + joint recovery; median outer iters %d"
                         % (args.cpu_sample, int(np.median(res["iterations"])))}

    serial_warm = serial_ms
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "warmup_actual": n_warm, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s ProblemGraphRevolute, batch=%d random reachable goal poses per GPU%s"
                                   % (args.robot, B, baseline_label(args.robot, B, world)),
                       "robot": args.robot, "batch_per_gpu": B, "nodes": N, "cost_terms": eng.plan.n_terms,
                       "concurrent_batches": C, "rtr_kernel": args.kernel,
                       "l2": ("flushed between timed steps (256 MiB write)" if C == 1 else
                              "not flushed: %d batches in flight on separate streams evict each other" % C),
                       "parallelism": "goals sharded, dp%d" % world},
            "serial": {"value": world * B / (float(np.mean(serial_warm)) * 1e-3) if serial_warm else None, "unit": UNIT,
                       "ms_per_batch": float(np.mean(serial_warm)) if serial_warm else None,
                       "note": "one batch at a time (3 steps after the timed regions): latency of a single solve_batch "
                               "call, set by the batch's slowest goal (maxiter = 3000)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_rtr_fast / k_rtr_duo / k_rtr (persistent trust-region solve)",
                         # C launches overlap on the device, so the launch duration that matters for the
                         # roofline is the timed region divided by the launches it retired; the raw
                         # CUDA-event duration of one launch (which includes time-sharing the SMs with
                         # C-1 others) is kept as per_launch_*
                         "achieved": aggregate, "peak": peak, "unit": "GB/s", "frac": aggregate / peak,
                         "per_launch_achieved": achieved, "per_launch_frac": achieved / peak,
                         "peak_source": peak_src, "traffic": ncu_traffic(workload),
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "kernel_ms": rtr_avg_ms, "effective_kernel_ms": local_ms / args.steps,
                         # share of the trust-region kernel in a step: from the serial (one batch at a time) steps,
                         # comparable with ncu's serialised launch list (profiles/r1e_launches_bench_steps2.csv: 99.3 %);
                         # with C batches in flight a step's event interval also contains the time its small kernels
                         # wait for SM slots held by the other batches' persistent launches
                         "kernel_share_of_step": float(np.sum(serial_rtr_ms) / np.sum(serial_ms)) if serial_ms else None,
                         "kernel_share_of_step_concurrent": float(np.sum(rtr_ms) / np.sum(step_ms)),
                         "fp64": {"achieved": alg_flops_total / (local_ms * 1e-3) / 1e12, "peak": fp64_peak_tflops,
                                  "unit": "TFLOP/s", "frac": alg_flops_total / (local_ms * 1e-3) / 1e12 / fp64_peak_tflops,
                                  "note": "algorithmic flops (38 per term + 60 N + 60 per tCG iteration) / region "
                                          "time vs the measured FP64 FMA peak; the kernel issues ~2.6x that many "
                                          "FP64 instructions (terms seen from both end nodes, warp-uniform scalars)"},
                         "note": "algorithmic bytes = sum over problems of 72N*outer + 240N*inner: the state a "
                                 "kernel-per-iteration formulation streams through HBM (SURVEY 8d), from the iteration "
                                 "counts actually executed; the persistent kernel keeps that state in registers (traffic "
                                 "= real DRAM bytes per launch from ncu), so the binding resources are FP64 issue and "
                                 "shuffle/FMA latency (profiles/)"},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "stats": {"converged_frac": stats["converged"] / max(stats["count"], 1),
                      "mean_outer_iters": stats["sum_outer"] / max(stats["count"], 1),
                      "mean_inner_iters": stats["sum_inner"] / max(stats["count"], 1),
                      "max_outer_iters": stats["max_outer"], "mean_f": stats["sum_f"] / max(stats["count"], 1)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
