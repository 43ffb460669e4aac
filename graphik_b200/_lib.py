"""ctypes binding of libgraphik_b200.so (the C ABI in include/graphik_b200.h).

The library is built in-tree with nvcc for sm_100a (`build()`); there is no CPU
fallback: if it cannot be loaded every solver entry point raises.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libgraphik_b200.so")
SOURCES = ["gik_plan.cu", "gik_costs.cu", "gik_rtr.cu", "gik_rtr_fast.cu", "gik_rtr_duo.cu", "gik_rtr_cta.cu", "gik_bounds_init.cu", "gik_joints.cu", "gik_fantope.cu", "gik_cg.cu", "gik_sdp.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static"]


class GikError(RuntimeError):
    pass


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(_HERE), "include", "graphik_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile every CUDA source for sm_100a into graphik_b200/lib/libgraphik_b200.so (one nvcc process
    per source file, then one link).  `defines` / `out` build a variant library for A/B experiments
    (tools/); the product always loads LIBPATH."""
    out = out or LIBPATH
    if not force and out == LIBPATH and not needs_build():
        return LIBPATH
    from concurrent.futures import ThreadPoolExecutor
    tag = "" if out == LIBPATH else "_" + os.path.splitext(os.path.basename(out))[0]
    objdir = os.path.join(LIBDIR, "obj" + tag)
    os.makedirs(objdir, exist_ok=True)
    dflags = ["-D" + d for d in defines]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + dflags + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise GikError("nvcc failed on %s:\n%s\n%s" % (src, res.stdout, res.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [_nvcc()] + LINK_FLAGS + ["-o", out] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise GikError("nvcc link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return out


class PlanDesc(ctypes.Structure):
    _fields_ = [
        ("n_nodes", ctypes.c_int32), ("n_terms", ctypes.c_int32),
        ("term_i", ctypes.c_void_p), ("term_j", ctypes.c_void_p), ("term_kind", ctypes.c_void_p),
        ("term_target", ctypes.c_void_p), ("term_goal", ctypes.c_void_p),
        ("n_goal", ctypes.c_int32), ("n_anchor", ctypes.c_int32),
        ("anchor_node", ctypes.c_void_p), ("anchor_pos", ctypes.c_void_p),
        ("goal_p", ctypes.c_int32), ("goal_q", ctypes.c_int32), ("axis_length", ctypes.c_double),
        ("bs_lower", ctypes.c_void_p), ("bs_upper", ctypes.c_void_p),
        ("n_goal_edges", ctypes.c_int32),
        ("goal_edge_i", ctypes.c_void_p), ("goal_edge_j", ctypes.c_void_p), ("goal_edge_slot", ctypes.c_void_p),
        ("omega", ctypes.c_void_p),
        ("n_joints", ctypes.c_int32), ("T0", ctypes.c_void_p),
        ("n_limits", ctypes.c_int32), ("limit_i", ctypes.c_void_p), ("limit_j", ctypes.c_void_p),
        ("limit_lower", ctypes.c_void_p), ("limit_upper", ctypes.c_void_p),
    ]


class SolveOpts(ctypes.Structure):
    _fields_ = [
        ("mingradnorm", ctypes.c_double), ("maxiter", ctypes.c_int32),
        ("theta", ctypes.c_double), ("kappa", ctypes.c_double),
        ("rho_prime", ctypes.c_double), ("rho_regularization", ctypes.c_double),
        ("mininner", ctypes.c_int32), ("maxinner", ctypes.c_int32),
        ("Delta_bar", ctypes.c_double), ("Delta0", ctypes.c_double),
        ("kernel", ctypes.c_int32), ("maxtime", ctypes.c_double),
    ]


class CgOpts(ctypes.Structure):
    _fields_ = [
        ("mingradnorm", ctypes.c_double), ("maxiter", ctypes.c_int32), ("minstepsize", ctypes.c_double),
        ("orth_value", ctypes.c_double), ("beta_type", ctypes.c_int32), ("maxtime", ctypes.c_double),
        ("ls_contraction", ctypes.c_double), ("ls_suff_decr", ctypes.c_double), ("ls_maxiter", ctypes.c_int32),
        ("ls_initial_stepsize", ctypes.c_double),
    ]


class SdpOpts(ctypes.Structure):
    _fields_ = [("tol", ctypes.c_double), ("maxiter", ctypes.c_int32), ("tau", ctypes.c_double),
                ("x0", ctypes.c_double)]


# every symbol include/graphik_b200.h declares
EXPORTS = [
    "gik_last_error", "gik_version", "gik_default_opts", "gik_plan_create", "gik_plan_destroy",
    "gik_plan_info", "gik_goal_distances", "gik_cost_grad", "gik_hessvec", "gik_proj", "gik_bounds",
    "gik_init", "gik_bounds_init", "gik_rtr_solve", "gik_joints", "gik_fk", "gik_check_limits",
    "gik_carry_bytes", "gik_carry_init", "gik_rtr_solve_sliced", "gik_workspace_bytes", "gik_fantope",
    "gik_cg_default_opts", "gik_cg_solve", "gik_sdp_default_opts", "gik_sdp_solve",
    "gik_cidgik_solve",
]

_lib = None


def load():
    """Load the CUDA library (building it first if nvcc is present and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        try:
            build()
        except (GikError, OSError) as e:
            if not os.path.exists(LIBPATH):
                raise GikError(
                    "libgraphik_b200.so is missing and could not be built (%s). graphik_b200 has no "
                    "CPU fallback: run `python -c 'import __graft_entry__ as g; g.build()'`." % e)
    try:
        L = ctypes.CDLL(LIBPATH)
    except OSError as e:
        raise GikError("cannot load %s: %s (no CPU fallback exists)" % (LIBPATH, e))
    vp, i32, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_double
    L.gik_last_error.restype = ctypes.c_char_p
    L.gik_last_error.argtypes = []
    L.gik_version.restype = ctypes.c_int
    sig = {
        "gik_default_opts": [ctypes.POINTER(SolveOpts)],
        "gik_plan_create": [ctypes.POINTER(PlanDesc), ctypes.POINTER(vp)],
        "gik_plan_destroy": [vp],
        "gik_plan_info": [vp, ctypes.POINTER(i32 * 8)],
        "gik_goal_distances": [vp, vp, i32, vp, vp],
        "gik_cost_grad": [vp, vp, vp, i32, vp, vp, vp],
        "gik_hessvec": [vp, vp, vp, vp, i32, vp, vp],
        "gik_proj": [i32, vp, vp, i32, vp, vp],
        "gik_bounds": [vp, vp, i32, vp, vp, vp, vp],
        "gik_init": [vp, vp, vp, i32, vp, vp, vp],
        "gik_bounds_init": [vp, vp, i32, vp, vp, vp],
        "gik_rtr_solve": [vp, vp, vp, i32, ctypes.POINTER(SolveOpts), vp, vp, vp, vp, vp, vp, vp, i32, vp, vp],
        "gik_carry_init": [vp, vp, i32, vp],
        "gik_rtr_solve_sliced": [vp, vp, vp, i32, ctypes.POINTER(SolveOpts), vp, vp, vp, vp, vp, vp, i32, vp, vp, vp,
                                 vp, vp],
        "gik_fantope": [i32, i32, vp, i32, vp, vp, vp],
        "gik_sdp_default_opts": [ctypes.POINTER(SdpOpts)],
        "gik_sdp_solve": [i32, i32, vp, vp, vp, vp, vp, i32, ctypes.POINTER(SdpOpts), vp, vp, vp, vp, vp, vp, vp],
        "gik_cidgik_solve": [i32, i32, i32, vp, vp, vp, vp, vp, i32, ctypes.POINTER(SdpOpts), i32, dbl, dbl, dbl,
                             vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "gik_cg_default_opts": [ctypes.POINTER(CgOpts)],
        "gik_cg_solve": [vp, vp, vp, i32, ctypes.POINTER(CgOpts), vp, vp, vp, vp, vp, vp, vp, i32, vp, vp],
        "gik_joints": [vp, vp, vp, i32, vp, vp],
        "gik_fk": [vp, vp, i32, vp, vp, vp],
        "gik_check_limits": [vp, vp, dbl, i32, vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.restype = ctypes.c_int
        fn.argtypes = args
    L.gik_carry_bytes.restype = ctypes.c_int64
    L.gik_carry_bytes.argtypes = [vp, i32]
    L.gik_workspace_bytes.restype = ctypes.c_int64
    L.gik_workspace_bytes.argtypes = [vp]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().gik_last_error().decode("utf-8", "replace")
        raise GikError("%s failed (%d): %s" % (what or "libgraphik_b200 call", rc, msg))
