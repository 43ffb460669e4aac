"""CPU timing arms of bench.py (TEST / MEASUREMENT INFRASTRUCTURE -- never imported by graphik_b200).

kind = "reference": the UNMODIFIED reference (`graphik.solvers.riemannian_solver.solve_with_riemannian`,
    riemannian_solver.py:220-234) driven exactly as BASELINE.md section 3 prescribes: one OS process per core,
    pinned, NUMBA / OMP / MKL threads = 1, `costgrd` AOT-built by numba.pycc on the machine that runs it,
    every process solving a contiguous slice of the same goal list, the first solve of every process
    discarded.  The reference is pure Python; `install()` (called by `__graft_entry__.build()` in the build
    container, where /root/reference exists) copies its package -- *.py and *.urdf only, no meshes -- into
    the git-ignored oracle/_ref/reference/, which travels to the GPU box with the repository snapshot.
kind = "port": the C restatement (oracle/gik_oracle.c) with the per-goal bound smoothing + numpy-eigh
    initialisation spread over worker processes and the trust-region solves over OpenMP threads.
"""
import os
import pickle
import shutil
import struct
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_COPY = os.path.join(HERE, "_ref", "reference")


def install(src="/root/reference"):
    """Copy the reference package (python sources + URDFs) into oracle/_ref/reference/.  Build container only."""
    pkg = os.path.join(src, "graphik")
    if not os.path.isdir(pkg):
        return False
    dst = os.path.join(REF_COPY, "graphik")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for root, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "meshes", "old_urdfs")]
        for f in files:
            if f.endswith((".py", ".urdf")):
                out = os.path.join(dst, os.path.relpath(root, pkg))
                os.makedirs(out, exist_ok=True)
                shutil.copy(os.path.join(root, f), os.path.join(out, f))
    return True


def reference_root():
    """Where an importable copy of the reference lives on this machine, or None."""
    for cand in (os.environ.get("GRAPHIK_REFERENCE"), "/root/reference", REF_COPY):
        if cand and os.path.isdir(os.path.join(cand, "graphik")):
            return cand
    return None


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def usable_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except AttributeError:
        return list(range(os.cpu_count() or 1))


# ---------------------------------------------------------------------------------------------------
# kind = "reference"

_W = {}


def _ref_worker_init(workload, core, root):
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ[k] = "1"
    try:
        os.sched_setaffinity(0, {core})
    except (AttributeError, OSError):
        pass
    os.environ["GRAPHIK_REFERENCE"] = root
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_runner
    ref_runner.REFERENCE = root
    ref_runner.load_reference()
    from graphik.solvers.riemannian_solver import solve_with_riemannian
    if workload.startswith("chain"):
        from gen_golden import random_dh_chain
        robot, graph, _ = random_dh_chain(int(workload[5:]), 0)
    else:
        from gen_golden import _loaders
        base = "kuka" if workload.startswith("kuka_table") else workload
        robot, graph = _loaders()[base]()
        if workload.startswith("kuka_table"):
            from graphik.utils.utils import table_environment
            for idx, obs in enumerate(table_environment()):
                graph.add_spherical_obstacle("o%d" % idx, obs[0], obs[1])
    _W.update(robot=robot, graph=graph, solve=solve_with_riemannian)


def _ref_worker_solve(T_slice):
    """solve_with_riemannian on every goal of the slice; the first one is solved once more up front and discarded."""
    from liegroups import SE3
    robot, graph, solve = _W["robot"], _W["graph"], _W["solve"]
    n = robot.n
    goals = [SE3.from_matrix(T) for T in T_slice]
    if len(goals) == 0:
        return 0.0, [], []
    solve(graph, goals[0], use_jit=True)          # discarded: numba dispatch caches, first-touch
    t0 = time.perf_counter()
    errs, ok = [], []
    for T in goals:
        q, Y = solve(graph, T, use_jit=True)
        ok.append(q is not None)
        if q is not None:
            Ts = robot.pose(q, "p%d" % n).as_matrix()
            errs.append(float(np.linalg.norm(Ts[:3, 3] - T.as_matrix()[:3, 3])))
    return time.perf_counter() - t0, errs, ok


class _Worker:
    """A child interpreter (`python -m oracle.ref_arm --worker ...`) talking length-prefixed pickles over pipes.
    Plain subprocesses instead of multiprocessing: nothing of the parent (bench.py, torch, CUDA) is re-imported."""

    def __init__(self, kind, workload, core, root):
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1",
                   NUMBA_NUM_THREADS="1", GRAPHIK_REFERENCE=root or "", PYTHONDONTWRITEBYTECODE="1")
        self.log = open(os.path.join(HERE, "_ref", "worker_%s.log" % core), "w")
        self.p = subprocess.Popen([sys.executable, "-m", "oracle.ref_arm", "--worker", kind, workload, str(core)],
                                  cwd=ROOT, env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=self.log)

    def send(self, obj):
        data = pickle.dumps(obj, protocol=4)
        self.p.stdin.write(struct.pack("<q", len(data)) + data)
        self.p.stdin.flush()

    def recv(self):
        head = self.p.stdout.read(8)
        if len(head) < 8:
            raise RuntimeError("CPU-arm worker died (see oracle/_ref/worker_*.log)")
        (n,) = struct.unpack("<q", head)
        return pickle.loads(self.p.stdout.read(n))

    def close(self):
        try:
            self.p.stdin.close()
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.log.close()


def _worker_main(kind, workload, core):
    out = os.fdopen(os.dup(1), "wb")      # replies go to the real stdout; stray prints of the libraries to stderr
    os.dup2(2, 1)
    inp = sys.stdin.buffer
    if kind == "reference":
        _ref_worker_init(workload, core, os.environ["GRAPHIK_REFERENCE"])
        fn = _ref_worker_solve
    else:
        _port_worker_init()
        fn = _port_worker_init_points

    def reply(obj):
        data = pickle.dumps(obj, protocol=4)
        out.write(struct.pack("<q", len(data)) + data)
        out.flush()

    reply("ready")
    while True:
        head = inp.read(8)
        if len(head) < 8:
            return
        (n,) = struct.unpack("<q", head)
        reply(fn(pickle.loads(inp.read(n))))


class ReferencePool:
    """Persistent worker processes (one per core) holding the imported reference + its robot graph."""

    def __init__(self, workload, cores=None):
        root = reference_root()
        if root is None:
            raise RuntimeError("no copy of the reference on this machine (oracle/_ref/reference is made by build())")
        self.cores = cores or usable_cores()
        # numba.pycc build of the reference's costs.py on THIS machine, once, before the workers import it
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        from oracle import ref_runner
        ref_runner.REFERENCE = root
        t0 = time.perf_counter()
        ref_runner._build_costgrd()
        self.costgrd_build_s = time.perf_counter() - t0
        self.workers = [_Worker("reference", workload, c, root) for c in self.cores]
        for w in self.workers:                    # wait until every worker has imported and built its graph
            assert w.recv() == "ready"

    def solve(self, T):
        """T[B,4,4] split into contiguous slices, one per core.  Returns (seconds = slowest worker, pose errors, ok)."""
        B, P = len(T), len(self.workers)
        bounds = [(k * B) // P for k in range(P + 1)]
        for k, w in enumerate(self.workers):
            w.send(np.asarray(T[bounds[k]:bounds[k + 1]]))
        out = [w.recv() for w in self.workers]
        errs = [e for o in out for e in o[1]]
        ok = [e for o in out for e in o[2]]
        return max(o[0] for o in out), errs, ok

    def close(self):
        for w in self.workers:
            w.close()


# ---------------------------------------------------------------------------------------------------
# kind = "port"

def _port_worker_init():
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)


def _port_worker_init_points(args):
    """Bound smoothing (C) + generate_initialization (numpy eigh) for a slice of goals."""
    from oracle import oracle as orc
    edge, lower, upper, ii, jj, slots, goal_d2, omega = args
    N = edge.shape[0]
    Y0 = np.empty((goal_d2.shape[0], N, 3))
    for b in range(goal_d2.shape[0]):
        lo, up = lower.copy(), upper.copy()
        lo[ii, jj] = up[ii, jj] = np.sqrt(goal_d2[b, slots])
        lb, ub = orc.bound_smoothing(edge, lo, up)
        Y0[b] = orc.generate_initialization(lb, ub, omega)
    return Y0


class PortPool:
    """Worker processes for the Python part of the port arm (the per-goal initialisation)."""

    def __init__(self, procs=None):
        cores = usable_cores()
        self.procs = procs or len(cores)
        self.workers = [_Worker("port", "-", cores[k % len(cores)], None) for k in range(self.procs)]
        for w in self.workers:
            assert w.recv() == "ready"

    def init_points(self, edge, lower, upper, ii, jj, slots, goal_d2, omega):
        B = goal_d2.shape[0]
        P = min(self.procs, max(1, B))
        bounds = [(k * B) // P for k in range(P + 1)]
        for k in range(P):
            self.workers[k].send((edge, lower, upper, ii, jj, slots, goal_d2[bounds[k]:bounds[k + 1]], omega))
        return np.concatenate([self.workers[k].recv() for k in range(P)], 0)

    def close(self):
        for w in self.workers:
            w.close()


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "--worker":
        _worker_main(sys.argv[2], sys.argv[3], int(sys.argv[4]))
