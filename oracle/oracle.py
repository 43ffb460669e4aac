"""Python face of the CPU oracle (ctypes over gik_oracle.c + numpy for the
eigen-decomposition based initialisation).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by graphik_b200/.

Parity status: pinned against outputs of the reference itself run in the build
container (tests/golden/*.npz, made by oracle/gen_golden.py); the reference's
own test-suite has no fixture for this path.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libgik_oracle.so")

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_long_p = ctypes.POINTER(ctypes.c_long)


class _Problem(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("E", ctypes.c_int), ("ii", _c_i64_p), ("jj", _c_i64_p),
                ("D", _c_double_p), ("omega", _c_double_p), ("psiL", _c_double_p), ("psiU", _c_double_p)]


class Params(ctypes.Structure):
    _fields_ = [("mingradnorm", ctypes.c_double), ("maxiter", ctypes.c_int),
                ("theta", ctypes.c_double), ("kappa", ctypes.c_double),
                ("rho_prime", ctypes.c_double), ("rho_regularization", ctypes.c_double),
                ("mininner", ctypes.c_int), ("maxinner", ctypes.c_int),
                ("Delta_bar", ctypes.c_double), ("Delta0", ctypes.c_double),
                ("use_limits", ctypes.c_int)]


def build(force=False):
    src = os.path.join(HERE, "gik_oracle.c")
    if (not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
                    "-o", LIB, src, "-lm"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.orc_lcost.restype = ctypes.c_double
        L.orc_jcost.restype = ctypes.c_double
        L.orc_rtr_solve.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def limit_inds(omega, psi_L, psi_U):
    """riemannian_solver.py:123-125."""
    diff = psi_L != psi_U
    return np.nonzero(np.triu(omega) + np.triu(diff * (psi_L > 0)) + np.triu(diff * (psi_U > 0)))


def equality_inds(omega):
    """riemannian_solver.py:79."""
    return np.nonzero(np.triu(omega))


class Problem:
    """One EDM-completion problem (D_goal, omega, psi_L, psi_U, inds)."""

    def __init__(self, D_goal, omega, psi_L=None, psi_U=None, use_limits=True):
        self.D, self.omega = _f64(D_goal), _f64(omega)
        self.psiL = _f64(psi_L if psi_L is not None else 0 * self.omega)
        self.psiU = _f64(psi_U if psi_U is not None else 0 * self.omega)
        self.use_limits = bool(use_limits)
        inds = limit_inds(self.omega, self.psiL, self.psiU) if use_limits else equality_inds(self.omega)
        self.ii = np.ascontiguousarray(inds[0], dtype=np.int64)
        self.jj = np.ascontiguousarray(inds[1], dtype=np.int64)
        self.N = self.D.shape[0]
        self.c = _Problem(self.N, len(self.ii), self.ii.ctypes.data_as(_c_i64_p),
                          self.jj.ctypes.data_as(_c_i64_p), _dp(self.D), _dp(self.omega),
                          _dp(self.psiL), _dp(self.psiU))

    # costs.py leaves ------------------------------------------------------
    def cost(self, Y):
        Y = _f64(Y)
        f = lib().orc_lcost if self.use_limits else lib().orc_jcost
        return f(ctypes.byref(self.c), _dp(Y))

    def grad(self, Y):
        Y = _f64(Y)
        out = np.empty_like(Y)
        f = lib().orc_lgrad if self.use_limits else lib().orc_jgrad
        f(ctypes.byref(self.c), _dp(Y), _dp(out))
        return out

    def hess(self, Y, w):
        Y, w = _f64(Y), _f64(w)
        out = np.empty_like(Y)
        f = lib().orc_lhess if self.use_limits else lib().orc_jhess
        f(ctypes.byref(self.c), _dp(Y), _dp(w), _dp(out))
        return out

    # trust-region solve ------------------------------------------------------
    def solve(self, Y_init, params=None, trace_rows=0):
        q = default_params(params)
        q.use_limits = int(self.use_limits)
        Y_init = _f64(Y_init)
        x = np.empty_like(Y_init)
        f, gn = ctypes.c_double(), ctypes.c_double()
        status, nh = ctypes.c_int(), ctypes.c_long()
        trace = np.full((max(trace_rows, 1), 6), np.nan)
        iters = lib().orc_rtr_solve(ctypes.byref(self.c), ctypes.byref(q), _dp(Y_init), _dp(x),
                                    ctypes.byref(f), ctypes.byref(gn), ctypes.byref(status),
                                    ctypes.byref(nh), _dp(trace) if trace_rows else None,
                                    ctypes.c_int(trace_rows))
        return {"x": x, "f(x)": f.value, "gradnorm": gn.value, "iterations": iters,
                "status": status.value, "n_hess": nh.value, "trace": trace[:min(iters, trace_rows)]}


    # conjugate-gradient solve ------------------------------------------------
    def solve_cg(self, Y_init, params=None, trace_rows=0):
        """RiemannianSolver(params={"solver": "ConjugateGradient"}) (riemannian_solver.py:52-60): pymanopt 0.2.5's
        ConjugateGradient + LineSearchAdaptive on the reference's manifold (retr(Y, U) = Y + U, transp(Y, Z, U) =
        proj(Z, U), egrad2rgrad = identity, Frobenius inner product, no preconditioner).

        PARITY UNPINNED: pymanopt is a third-party dependency that is not in the reference tree (setup.py:20 pins
        0.2.5) and the reference holds no test or vector for this branch.  This is a restatement of the published
        algorithm of that version (pymanopt/solvers/conjugate_gradient.py, linesearch.py = Manopt's
        conjugategradient.m, linesearch_adaptive.m), statement by statement; it is what the CUDA kernel is
        checked against (gik_cg.cu)."""
        q = {"mingradnorm": 1e-9, "maxiter": 10e4, "minstepsize": 1e-10, "orth_value": 10e10, "beta_type": 3,
             "contraction_factor": .5, "suff_decr": .5, "ls_maxiter": 10, "initial_stepsize": 1}
        q.update(params or {})
        inner = lambda a, b: float(np.tensordot(a, b))
        x = _f64(Y_init).copy()
        it, stepsize, oldalpha = 0, np.nan, None
        cost, grad = self.cost(x), self.grad(x)
        gradnorm = np.sqrt(inner(grad, grad))
        gradPgrad = inner(grad, grad)
        desc = -grad
        trace, status, costevals = [], 1, 1
        while True:
            # Solver._check_stopping_criterion(time0, gradnorm=gradnorm, iter=iter + 1, stepsize=stepsize)
            if it + 1 >= q["maxiter"]:
                status = 1
                break
            if gradnorm < q["mingradnorm"]:
                status = 0
                break
            if stepsize < q["minstepsize"]:
                status = 6
                break
            df0 = inner(grad, desc)
            restarted = False
            if df0 >= 0:
                desc = -grad
                df0 = -gradPgrad
                restarted = True
            # LineSearchAdaptive.search
            norm_d = np.sqrt(inner(desc, desc))
            alpha = float(oldalpha) if oldalpha is not None else q["initial_stepsize"] / norm_d
            newx = x + alpha * desc
            newf = self.cost(newx)
            evals = 1
            while newf > cost + q["suff_decr"] * alpha * df0 and evals <= q["ls_maxiter"]:
                alpha *= q["contraction_factor"]
                newx = x + alpha * desc
                newf = self.cost(newx)
                evals += 1
            if newf > cost:
                alpha = 0
                newx = x
            stepsize = alpha * norm_d
            oldalpha = alpha if evals == 2 else 2 * alpha
            costevals += evals
            # new cost-related quantities
            newcost, newgrad = self.cost(newx), self.grad(newx)
            newgradPnewgrad = inner(newgrad, newgrad)
            newgradnorm = np.sqrt(newgradPnewgrad)
            oldgrad = proj(newx, grad)
            orth_grads = inner(oldgrad, newgrad) / newgradPnewgrad
            if abs(orth_grads) >= q["orth_value"]:
                beta = 0
                desc = -newgrad
            else:
                desc = proj(newx, desc)
                bt = q["beta_type"]
                if bt == 0:
                    beta = newgradPnewgrad / gradPgrad
                elif bt == 1:
                    beta = max(0, inner(newgrad, newgrad - oldgrad) / gradPgrad)
                elif bt == 2:
                    diff = newgrad - oldgrad
                    den = inner(diff, desc)
                    beta = 1 if den == 0 else max(0, inner(newgrad, diff) / den)
                else:
                    diff = newgrad - oldgrad
                    Pdiff = newgrad - oldgrad          # Poldgrad = transp(x, newx, Pgrad) = oldgrad (no preconditioner)
                    deno = inner(diff, desc)
                    numo = inner(diff, newgrad)
                    numo -= 2 * inner(diff, Pdiff) * inner(desc, newgrad) / deno
                    beta = numo / deno
                    eta_HZ = -1 / (np.sqrt(inner(desc, desc)) * min(0.01, newgradnorm))
                    beta = max(beta, eta_HZ)
                desc = -newgrad + beta * desc
            if len(trace) < trace_rows:
                trace.append([stepsize, evals, beta, newcost, 1.0 if restarted else 0.0, newgradnorm])
            x, cost, grad, gradnorm, gradPgrad = newx, newcost, newgrad, newgradnorm, newgradPnewgrad
            it += 1
        return {"x": x, "f(x)": cost, "gradnorm": gradnorm, "iterations": it, "status": status,
                "stepsize": stepsize, "costevals": costevals, "trace": np.array(trace).reshape(-1, 6)}


def default_params(params=None):
    q = Params()
    lib().orc_default_params(ctypes.byref(q))
    for k, v in (params or {}).items():
        if hasattr(q, k):
            setattr(q, k, v)
    return q


def proj(Y, Z):
    Y, Z = _f64(Y), _f64(Z)
    out = np.empty_like(Y)
    lib().orc_proj(ctypes.c_int(Y.shape[0]), _dp(Y), _dp(Z), _dp(out))
    return out


def solve_batch(D, omega, psi_L, psi_U, Y_init, params=None, threads=None):
    """B problems sharing omega/psi (goal-independent), OpenMP over problems (`threads`: 0 / None = all cores)."""
    D, Y_init = _f64(D), _f64(Y_init)
    B, N = D.shape[0], D.shape[1]
    proto = Problem(D[0], omega, psi_L, psi_U, True)
    q = default_params(params)
    q.use_limits = 1
    Y = np.empty_like(Y_init)
    f, gn = np.empty(B), np.empty(B)
    iters, status = np.empty(B, np.int32), np.empty(B, np.int32)
    nh = np.empty(B, np.int64)
    lib().orc_rtr_solve_batch(
        ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(len(proto.ii)),
        proto.ii.ctypes.data_as(_c_i64_p), proto.jj.ctypes.data_as(_c_i64_p), _dp(D),
        _dp(proto.omega), _dp(proto.psiL), _dp(proto.psiU), ctypes.byref(q), _dp(Y_init), _dp(Y),
        _dp(f), _dp(gn), iters.ctypes.data_as(_c_int_p), status.ctypes.data_as(_c_int_p),
        nh.ctypes.data_as(_c_long_p), ctypes.c_int(int(threads or 0)))
    return {"x": Y, "f(x)": f, "gradnorm": gn, "iterations": iters, "status": status, "n_hess": nh}


# ------------------------------------------------------------------ dgp.py

def bound_smoothing(edge, lower, upper):
    """dgp.py:192-231 on dense edge-attribute matrices (unsquared)."""
    edge = np.ascontiguousarray(edge, dtype=np.uint8)
    lo = _f64(np.nan_to_num(lower, nan=0.0))
    up = _f64(np.nan_to_num(upper, nan=np.inf))
    N = edge.shape[0]
    lb, ub = np.empty((N, N)), np.empty((N, N))
    lib().orc_bound_smoothing(ctypes.c_int(N), edge.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                              _dp(lo), _dp(up), _dp(lb), _dp(ub))
    return lb, ub


def gram_from_distance_matrix(D):
    """dgp.py:28-31."""
    n = D.shape[0]
    J = np.identity(n) - (1 / n) * np.ones(D.shape)
    return -0.5 * J @ D @ J


def canonical_signs(V):
    """Make the entry of largest magnitude of every column positive (first on ties)."""
    V = V.copy()
    for c in range(V.shape[1]):
        k = int(np.argmax(np.abs(V[:, c])))
        if V[k, c] < 0:
            V[:, c] = -V[:, c]
    return V


def factor(A, signs="lapack"):
    """dgp.py:150-159.  `signs`: eigenvector sign convention -- "lapack" keeps what
    numpy/LAPACK returns (what the reference inherits), "canonical" applies
    canonical_signs (what the CUDA kernel does, see gik_bounds_init.cu)."""
    evals, evecs = np.linalg.eigh(A)
    if signs == "canonical":
        evecs = canonical_signs(evecs)
        # second half of the canonical convention: an eigenvalue below 1e-13 of the largest is rounding noise of the
        # exactly singular Gram matrix (J D J annihilates the vector of ones); its sign differs between LAPACK, Jacobi
        # and QL, and a "positive" one would add a ~3e-8 column that can tip the 1e-8 rank count of MDS.  It counts
        # as zero.  (No effect on any of the 38 golden goals, where "lapack" stays pinned to the reference.)
        evals = np.where(evals > 1e-13 * max(np.max(evals), 0.0), evals, 0.0)
    evals[evals < 0] = 0
    X = evecs.dot(np.diag(np.sqrt(evals)))
    return np.fliplr(X)


def MDS(B, eps=1e-5, signs="lapack"):
    """dgp.py:163-171 -- including the eigh() of the NON-symmetric factor (numpy reads
    the lower triangle) that sets the kept rank K.  NB: K is not invariant to the signs
    of the eigenvectors inside `factor`, which LAPACK leaves arbitrary."""
    n = B.shape[0]
    x = factor(B, signs)
    evals, _ = np.linalg.eigh(x)
    K = int(np.sum(evals > eps))
    if K < n:
        x = x[:, 0:K]
    return x


def linear_projection(P, F, dim):
    """dgp.py:174-183."""
    I = np.nonzero(F)
    d = P[I[0]] - P[I[1]]
    S = d.T @ d
    _, eigvec = np.linalg.eigh(S)
    return P @ np.fliplr(eigvec)[:, :dim]


def generate_initialization(lb, ub, omega, dim=3, signs="lapack"):
    """riemannian_solver.py:67-75."""
    D_rand = (lb + 0.9 * (ub - lb)) ** 2
    X_rand = MDS(gram_from_distance_matrix(D_rand), eps=1e-8, signs=signs)
    return linear_projection(X_rand, omega, dim)
