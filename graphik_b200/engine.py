"""Batched IK engine: the device pipeline behind `solve_with_riemannian`.

For B goal poses at once it runs, on the current CUDA stream, what the reference
does per pose in Python (riemannian_solver.py:220-234):

    T_goal --gik_goal_distances--> goal_d2            (from_pose / D_goal entries)
           --gik_bounds_init----> Y_init              (bound_smoothing + generate_initialization)
           --gik_rtr_solve------> Y, f, |g|, iters    (TrustRegions.solve on PSDFixedRank)
           --gik_joints---------> q                   (joint_variables)
           --gik_fk-------------> realised points     (realization, for the limit check)

torch is used only as the device-memory container (tensor.data_ptr()) and for
stream handles; all arithmetic is in libgraphik_b200.so.
"""
import ctypes

import numpy as np

from graphik_b200 import _lib
from graphik_b200.plan import Plan


def _torch():
    import torch
    return torch


def make_cg_opts(params=None):
    """GikCgOpts from the reference's params dict (riemannian_solver.py:52-60)."""
    o = _lib.CgOpts()
    _lib.check(_lib.load().gik_cg_default_opts(ctypes.byref(o)), "gik_cg_default_opts")
    params = params or {}
    for key in ("mingradnorm", "minstepsize", "orth_value", "maxtime"):
        if key in params:
            setattr(o, key, float(params[key]))
    if "maxiter" in params:
        o.maxiter = int(min(float(params["maxiter"]), 2 ** 31 - 1))
    if "beta_type" in params:      # BetaTypes value 0..3 or its name
        bt = params["beta_type"]
        names = ["FletcherReeves", "PolakRibiere", "HestenesStiefel", "HagerZhang"]
        o.beta_type = names.index(bt) if isinstance(bt, str) else int(bt)
    return o


def make_opts(params=None):
    """GikSolveOpts from the reference's params dict (riemannian_solver.py:41-50); GikCgOpts when
    params["solver"] == "ConjugateGradient"."""
    params = params or {}
    if params.get("solver", "TrustRegions") == "ConjugateGradient":
        return make_cg_opts(params)
    o = _lib.SolveOpts()
    _lib.check(_lib.load().gik_default_opts(ctypes.byref(o)), "gik_default_opts")
    for key in ("mingradnorm", "theta", "kappa", "rho_prime", "rho_regularization", "Delta_bar", "Delta0", "maxtime"):
        if key in params:
            setattr(o, key, float(params[key]))
    for key in ("maxiter", "mininner", "maxinner"):
        if key in params:
            setattr(o, key, int(params[key]))
    if "kernel" in params:   # "auto" | "latency" | "throughput" | "generic" | "dense" (same results, different SM mapping)
        o.kernel = {"auto": 0, "latency": 1, "throughput": 2, "generic": 3, "dense": 4}[params["kernel"]]
    if "Delta_bar" in params and "Delta0" not in params:
        o.Delta0 = o.Delta_bar / 8  # trust_region.py:137-138
    return o


def _p(t):
    """Raw device pointer of a contiguous CUDA tensor (or NULL)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


class BatchIK:
    def __init__(self, graph=None, params=None, device=None, use_limits=True, plan=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.GikError("graphik_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        self.lib = _lib.load()
        self.graph = graph
        with torch.cuda.device(self.device):
            self.plan = plan if plan is not None else Plan.from_graph(graph, use_limits=use_limits)
        self.opts = make_opts(params)
        self._counters = {}   # work-queue counter of gik_rtr_solve per CUDA stream (calls on one stream serialise)
        self._workspaces = {}
        self._ws_bytes = int(self.lib.gik_workspace_bytes(self.plan.handle))
        self.launches = 0  # kernels launched through this engine (bench.py's gpu_launches)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _f64(self, x, shape=None):
        torch = self.torch
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
        x = x.to(device=self.device, dtype=torch.float64).contiguous()
        if shape is not None:
            x = x.reshape(shape)
        return x

    def _empty(self, *shape, dtype=None):
        return self.torch.empty(shape, dtype=dtype or self.torch.float64, device=self.device)

    def _counter(self):
        """Per-stream scratch for the persistent work queue: two solves issued on different streams from one
        engine must not share it."""
        key = self.torch.cuda.current_stream(self.device).cuda_stream
        c = self._counters.get(key)
        if c is None:
            c = self._counters[key] = self.torch.zeros(1, dtype=self.torch.int32, device=self.device)
        return c

    def _workspace(self):
        """Per-stream workspace of the bound-smoothing / initialisation kernel (None while everything fits in
        shared memory: small graphs only, gik_workspace_bytes() says)."""
        if not self._ws_bytes:
            return None
        key = self.torch.cuda.current_stream(self.device).cuda_stream
        w = self._workspaces.get(key)
        if w is None:
            w = self._workspaces[key] = self.torch.empty(self._ws_bytes // 8, dtype=self.torch.float64,
                                                         device=self.device)
        return w

    # ------------------------------------------------------------------ stages
    def goal_distances(self, T_goal):
        T = self._f64(T_goal).reshape(-1, 4, 4)
        B = T.shape[0]
        out = self._empty(B, self.plan.n_goal)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_goal_distances(self.plan.handle, _p(T), B, _p(out), self._stream()),
                       "gik_goal_distances")
        self.launches += 1
        return out

    def bounds(self, goal_d2, B=None):
        N = self.plan.N
        g = self._f64(goal_d2).reshape(-1, self.plan.n_goal) if goal_d2 is not None else None
        B = g.shape[0] if g is not None else int(B)
        lb, ub = self._empty(B, N, N), self._empty(B, N, N)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_bounds(self.plan.handle, _p(g), B, _p(lb), _p(ub), _p(self._workspace()),
                                           self._stream()), "gik_bounds")
        self.launches += 1
        return lb, ub

    def init_from_bounds(self, lb, ub):
        N = self.plan.N
        lb, ub = self._f64(lb).reshape(-1, N, N), self._f64(ub).reshape(-1, N, N)
        B = lb.shape[0]
        Y = self._empty(B, N, 3)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_init(self.plan.handle, _p(lb), _p(ub), B, _p(Y), _p(self._workspace()),
                                         self._stream()), "gik_init")
        self.launches += 1
        return Y

    def initialization(self, goal_d2):
        g = self._f64(goal_d2).reshape(-1, self.plan.n_goal)
        B, N = g.shape[0], self.plan.N
        Y = self._empty(B, N, 3)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_bounds_init(self.plan.handle, _p(g), B, _p(Y), _p(self._workspace()),
                                                self._stream()), "gik_bounds_init")
        self.launches += 1
        return Y

    def cost_grad(self, Y, goal_d2=None, want_grad=True):
        N = self.plan.N
        Y = self._f64(Y).reshape(-1, N, 3)
        B = Y.shape[0]
        g2 = self._f64(goal_d2).reshape(B, self.plan.n_goal) if self.plan.n_goal else None
        f = self._empty(B)
        g = self._empty(B, N, 3) if want_grad else None
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_cost_grad(self.plan.handle, _p(Y), _p(g2), B, _p(f), _p(g), self._stream()),
                       "gik_cost_grad")
        self.launches += 1
        return f, g

    def hessvec(self, Y, W, goal_d2=None):
        N = self.plan.N
        Y, W = self._f64(Y).reshape(-1, N, 3), self._f64(W).reshape(-1, N, 3)
        B = Y.shape[0]
        g2 = self._f64(goal_d2).reshape(B, self.plan.n_goal) if self.plan.n_goal else None
        out = self._empty(B, N, 3)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_hessvec(self.plan.handle, _p(Y), _p(W), _p(g2), B, _p(out), self._stream()),
                       "gik_hessvec")
        self.launches += 1
        return out

    def proj(self, Y, Z):
        Y, Z = self._f64(Y), self._f64(Z)
        N = Y.shape[-2]
        Y, Z = Y.reshape(-1, N, 3), Z.reshape(-1, N, 3)
        out = self._empty(Y.shape[0], N, 3)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_proj(N, _p(Y), _p(Z), Y.shape[0], _p(out), self._stream()), "gik_proj")
        self.launches += 1
        return out

    def solve_points(self, goal_d2, Y_init, trace_rows=0, opts=None):
        """TrustRegions.solve for B problems.  Returns device tensors."""
        torch = self.torch
        N = self.plan.N
        Y0 = self._f64(Y_init).reshape(-1, N, 3)
        B = Y0.shape[0]
        g2 = self._f64(goal_d2).reshape(B, self.plan.n_goal) if self.plan.n_goal else None
        Y = self._empty(B, N, 3)
        f, gn = self._empty(B), self._empty(B)
        iters = self._empty(B, dtype=torch.int32)
        status = self._empty(B, dtype=torch.int32)
        n_inner = self._empty(B, dtype=torch.int32)
        trace = torch.full((B, trace_rows, 6), float("nan"), dtype=torch.float64, device=self.device) \
            if trace_rows else None
        o = opts or self.opts
        if isinstance(o, _lib.CgOpts):      # params["solver"] == "ConjugateGradient"; n_inner = cost evaluations
            with torch.cuda.device(self.device):
                _lib.check(self.lib.gik_cg_solve(self.plan.handle, _p(g2), _p(Y0), B, ctypes.byref(o), _p(Y), _p(f),
                                                 _p(gn), _p(iters), _p(status), _p(n_inner), _p(trace),
                                                 int(trace_rows), _p(self._counter()), self._stream()), "gik_cg_solve")
            self.launches += 2
            return {"x": Y, "f(x)": f, "gradnorm": gn, "iterations": iters, "status": status,
                    "n_inner": n_inner, "trace": trace}
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gik_rtr_solve(self.plan.handle, _p(g2), _p(Y0), B, ctypes.byref(o), _p(Y), _p(f),
                                              _p(gn), _p(iters), _p(status), _p(n_inner), _p(trace),
                                              int(trace_rows), _p(self._counter()), self._stream()), "gik_rtr_solve")
        self.launches += 2  # memset of the work counter + the persistent kernel
        return {"x": Y, "f(x)": f, "gradnorm": gn, "iterations": iters, "status": status,
                "n_inner": n_inner, "trace": trace}

    def joints(self, Y, T_goal=None):
        N, n = self.plan.N, self.plan.n_joints
        Y = self._f64(Y).reshape(-1, N, 3)
        B = Y.shape[0]
        T = self._f64(T_goal).reshape(B, 4, 4) if T_goal is not None else None
        q = self._empty(B, n)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_joints(self.plan.handle, _p(Y), _p(T), B, _p(q), self._stream()), "gik_joints")
        self.launches += 1
        return q

    def fk(self, q, want_points=True):
        n, N = self.plan.n_joints, self.plan.N
        q = self._f64(q).reshape(-1, n)
        B = q.shape[0]
        T = self._empty(B, 4, 4)
        Y = self._empty(B, N, 3) if want_points else None
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_fk(self.plan.handle, _p(q), B, _p(T), _p(Y), self._stream()), "gik_fk")
        self.launches += 1
        return T, Y

    def check_limits(self, Y, tol=1e-6, status=None):
        """Intended-semantics check_distance_limits on the device: violations per problem.  `status` (int32[B],
        optional) gets GIK_STATUS_LIMITS (3) where a limit is broken."""
        N = self.plan.N
        Y = self._f64(Y).reshape(-1, N, 3)
        B = Y.shape[0]
        out = self._empty(B, dtype=self.torch.int32)
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.gik_check_limits(self.plan.handle, _p(Y), float(tol), B, _p(out), _p(status),
                                                 self._stream()), "gik_check_limits")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ pipeline
    def solve(self, T_goal, Y_init=None, trace_rows=0, check=True):
        """solve_with_riemannian for a batch of goal poses T_goal[B,4,4] (device or host).

        Returns device tensors: q[B,n], x[B,N,3], f(x), gradnorm, iterations, status,
        n_inner, and -- when `check` -- the realised end-effector pose error of q."""
        T = self._f64(T_goal).reshape(-1, 4, 4)
        g2 = self.goal_distances(T)
        if Y_init is None:
            Y_init = self.initialization(g2)
        out = self.solve_points(g2, Y_init, trace_rows=trace_rows)
        out["goal_d2"], out["T_goal"], out["Y_init"] = g2, T, self._f64(Y_init).reshape(-1, self.plan.N, 3)
        out["q"] = self.joints(out["x"], T)
        if check:
            T_sol, Y_real = self.fk(out["q"], want_points=True)
            out["n_broken"] = self.check_limits(Y_real, tol=1e-6)   # riemannian_solver.py:230 (intended semantics)
            out["pos_err"] = (T_sol[:, :3, 3] - T[:, :3, 3]).norm(dim=1)
            R = T_sol[:, :3, :3].transpose(1, 2) @ T[:, :3, :3]
            cosang = ((R.diagonal(dim1=1, dim2=2).sum(1) - 1.0) * 0.5).clamp(-1.0, 1.0)
            out["rot_err"] = self.torch.arccos(cosang)
        return out
