"""`costgrd`-compatible facade over the batched C ABI (B = 1).

The reference's only native plug-in for this path is the numba-AOT module
`graphik.solvers.costgrd` (costs.py:3-5,208-209) with eight functions taking numpy
arrays.  This module exposes the same names, argument order and return types so
that `from graphik_b200.solvers.costgrd import lcost, lgrad, lhess, ...` can stand
where `from graphik.solvers.costgrd import ...` stood (riemannian_solver.py:18-21).
It exists for API parity and tests; production code should call the batched engine.
"""
import numpy as np

from graphik_b200.engine import BatchIK
from graphik_b200.plan import Plan

_cache = {}


def _engine(D_goal, omega, psi_L, psi_U, inds, use_limits):
    D_goal = np.ascontiguousarray(D_goal, dtype=float)
    key = (use_limits, D_goal.tobytes(), np.asarray(omega).tobytes() if omega is not None else b"",
           np.asarray(psi_L).tobytes() if psi_L is not None else b"",
           np.asarray(psi_U).tobytes() if psi_U is not None else b"",
           np.asarray(inds[0]).tobytes(), np.asarray(inds[1]).tobytes())
    eng = _cache.get(key)
    if eng is None:
        if len(_cache) > 8:
            _cache.clear()
        # the edge set is whatever `inds` lists (costs.py loops over zip(*inds))
        plan = Plan.from_matrices(D_goal, omega, psi_L, psi_U, use_limits=use_limits, inds=inds)
        eng = BatchIK(plan=plan)
        _cache[key] = eng
    return eng


def _one(Y):
    return np.ascontiguousarray(Y, dtype=float)[None]


def jcost(Y, D_goal, inds):
    f, _ = _engine(D_goal, None, None, None, inds, False).cost_grad(_one(Y), want_grad=False)
    return float(f[0])


def jgrad(Y, D_goal, inds):
    _, g = _engine(D_goal, None, None, None, inds, False).cost_grad(_one(Y))
    return g[0].cpu().numpy()


def jhess(Y, w, D_goal, inds):
    return _engine(D_goal, None, None, None, inds, False).hessvec(_one(Y), _one(w))[0].cpu().numpy()


def jcost_and_grad(Y, D_goal, inds):
    f, g = _engine(D_goal, None, None, None, inds, False).cost_grad(_one(Y))
    return float(f[0]), g[0].cpu().numpy()


def lcost(Y, D_goal, omega, psi_L, psi_U, inds):
    f, _ = _engine(D_goal, omega, psi_L, psi_U, inds, True).cost_grad(_one(Y), want_grad=False)
    return float(f[0])


def lgrad(Y, D_goal, omega, psi_L, psi_U, inds):
    _, g = _engine(D_goal, omega, psi_L, psi_U, inds, True).cost_grad(_one(Y))
    return g[0].cpu().numpy()


def lcost_and_grad(Y, D_goal, omega, psi_L, psi_U, inds):
    f, g = _engine(D_goal, omega, psi_L, psi_U, inds, True).cost_grad(_one(Y))
    return float(f[0]), g[0].cpu().numpy()


def lhess(Y, w, D_goal, omega, psi_L, psi_U, inds):
    return _engine(D_goal, omega, psi_L, psi_U, inds, True).hessvec(_one(Y), _one(w))[0].cpu().numpy()
