"""Import the UNMODIFIED reference (GraphIK at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

The reference is pure Python and needs three third-party packages that are not
vendored and not installable offline (pymanopt==0.2.5, liegroups, urdfpy;
reference setup.py:19-24).  `oracle/shims/` holds small stand-ins (SURVEY.md
Appendix B).  This module:

  1. puts the shims and /root/reference on sys.path,
  2. applies two compat patches for numpy 2 / networkx 3 WITHOUT editing the
     reference (nx.shortest_path -> dict, geometry.skew ravel),
  3. AOT-builds the reference's own `costgrd` extension from
     graphik/solvers/costs.py with numba.pycc into oracle/_ref/ (git-ignored)
     and registers it as `graphik.solvers.costgrd`.

/root/reference does not exist on the GPU box; anything that imports this
module must only run here (golden generation, local validation tests that
skip when the reference is absent).
"""
import importlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(HERE, "_ref")


def _default_reference():
    """The read-only tree in the build container; on the GPU box the copy `oracle.ref_arm.install()` made."""
    for cand in (os.environ.get("GRAPHIK_REFERENCE"), "/root/reference", os.path.join(REF_BUILD, "reference")):
        if cand and os.path.isdir(os.path.join(cand, "graphik")):
            return cand
    return "/root/reference"


REFERENCE = _default_reference()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE, "graphik"))


def costgrd_dir():
    """numba.pycc emits code for the CPU it runs on: one build per CPU model (the build container's and the
    GPU box's hosts differ, and oracle/_ref/ travels between them)."""
    import hashlib
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("model name", "flags")):
                    model += line
                    if line.startswith("flags"):
                        break
    except OSError:
        pass
    return os.path.join(REF_BUILD, "costgrd_" + hashlib.sha1(model.encode()).hexdigest()[:10])


def _build_costgrd():
    """numba.pycc AOT build of the reference's costs.py (costs.py:208-209)."""
    out_dir = costgrd_dir()
    os.makedirs(out_dir, exist_ok=True)
    for f in os.listdir(out_dir):
        if f.startswith("costgrd") and f.endswith(".so"):
            return
    scratch = os.path.join(out_dir, "build")
    os.makedirs(scratch, exist_ok=True)
    # numba.pycc compiles the module it is handed; it must be run from a
    # writable directory, and the reference tree is read-only.  The copy is a
    # build input placed in the git-ignored scratch dir, never in the repo.
    shutil.copy(os.path.join(REFERENCE, "graphik", "solvers", "costs.py"),
                os.path.join(scratch, "costs.py"))
    env = dict(os.environ, NUMBA_CACHE_DIR=os.path.join(REF_BUILD, "numba_cache"))
    subprocess.run([sys.executable, "-W", "ignore", "costs.py"], cwd=scratch,
                   check=True, env=env, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    for f in os.listdir(scratch):
        if f.startswith("costgrd") and f.endswith(".so"):
            shutil.move(os.path.join(scratch, f), os.path.join(out_dir, f))
    shutil.rmtree(scratch, ignore_errors=True)


_loaded = {}


def load_reference(with_costgrd=True):
    """Returns the imported `graphik` package of the reference."""
    if "graphik" in _loaded:
        return _loaded["graphik"]
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE)
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(REF_BUILD, "numba_cache"))
    sys.dont_write_bytecode = True
    for p in (os.path.join(HERE, "shims"), REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)

    import networkx as nx
    import numpy as np

    # compat 1: networkx>=3 returns a generator from shortest_path(G)
    if not getattr(nx.shortest_path, "_gik_patched", False):
        _orig_sp = nx.shortest_path

        def _shortest_path(G, source=None, target=None, *a, **k):
            out = _orig_sp(G, source, target, *a, **k)
            if source is None and target is None and not isinstance(out, dict):
                out = dict(out)
            return out

        _shortest_path._gik_patched = True
        nx.shortest_path = _shortest_path

    if with_costgrd:
        _build_costgrd()
        if costgrd_dir() not in sys.path:
            sys.path.insert(0, costgrd_dir())
        costgrd = importlib.import_module("costgrd")
        sys.modules["graphik.solvers.costgrd"] = costgrd

    import graphik  # noqa: E402
    import graphik.utils.geometry as geom

    # compat 2: geometry.skew builds an array of arrays under numpy 2 when
    # handed a (3,1) column (roboturdf.py:287)
    if not getattr(geom.skew, "_gik_patched", False):
        _orig_skew = geom.skew

        def _skew(x):
            return _orig_skew(np.asarray(x, dtype=float).ravel())

        _skew._gik_patched = True
        for modname in ("graphik.utils.geometry", "graphik.utils", "graphik.utils.roboturdf",
                        "graphik.graphs.graph_revolute"):
            mod = sys.modules.get(modname) or importlib.import_module(modname)
            if hasattr(mod, "skew"):
                setattr(mod, "skew", _skew)

    import graphik.solvers.riemannian_solver  # noqa: F401,E402
    if with_costgrd:
        rs = sys.modules["graphik.solvers.riemannian_solver"]
        for name in ("jcost", "jgrad", "jhess", "lcost", "lgrad", "lhess"):
            setattr(rs, name, getattr(sys.modules["graphik.solvers.costgrd"], name))
    _loaded["graphik"] = graphik
    return graphik


if __name__ == "__main__":
    import time

    import numpy as np

    g = load_reference()
    from graphik.solvers.riemannian_solver import solve_with_riemannian
    from graphik.utils.roboturdf import load_ur10

    robot, graph = load_ur10()
    np.random.seed(0)
    q = robot.random_configuration()
    T = robot.pose(q, "p%d" % robot.n)
    t0 = time.time()
    q_sol, Y = solve_with_riemannian(graph, T, use_jit=True)
    print("solve s:", time.time() - t0)
    print("q_goal", q)
    print("q_sol ", q_sol)
    if q_sol:
        print("pose err", np.linalg.norm(robot.pose(q_sol, "p6").as_matrix() - T.as_matrix()))
