"""CIDGIK (reference solvers/convex_iteration.py) -- the part of it that is arithmetic of the reference itself.

The reference's convex iteration alternates (convex_iteration.py:160-276)
    1. a semidefinite program with linear cost <C, Z> (cvxpy -> MOSEK, sdp_snl.py:874-967), and
    2. the closed-form Fantope step C = U U^T (convex_iteration.py:43-53).
Step 2 is implemented here on the GPU for a batch of Gram matrices (`gik_fantope`, csrc/gik_fantope.cu) and is
checked against numpy.  Step 1 lives inside MOSEK: there is no source to restate, neither cvxpy nor MOSEK can be
installed next to this repository to produce golden outputs, and the reference's own tests pin only how the
constraints are built (tests/test_sdp_snl.py), never a solve.  `solve_with_cidgik` therefore raises: a result that
cannot be compared with the reference's would not be a drop-in.
"""
import ctypes
import time

import numpy as np

from graphik_b200 import _lib


def solve_fantope_closed_form_batch(G, d):
    """C[b] = projector onto the eigenvectors of the n - d smallest eigenvalues of G[b] (convex_iteration.py:43-53),
    for G[B, n, n] (array or CUDA tensor).  Returns (C, eigvals) as CUDA tensors; eigvals ascending like numpy.eigh."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.GikError("graphik_b200 needs a CUDA device (B200); there is no CPU fallback")
    Gt = torch.as_tensor(np.ascontiguousarray(G, dtype=np.float64)) if not isinstance(G, torch.Tensor) else G
    Gt = Gt.to(device="cuda", dtype=torch.float64).contiguous()
    if Gt.dim() == 2:
        Gt = Gt[None]
    B, n = Gt.shape[0], Gt.shape[-1]
    C = torch.empty_like(Gt)
    ev = torch.empty((B, n), dtype=torch.float64, device=Gt.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(Gt.device).cuda_stream)
    with torch.cuda.device(Gt.device):
        _lib.check(_lib.load().gik_fantope(n, int(d), ctypes.c_void_p(Gt.data_ptr()), B, ctypes.c_void_p(C.data_ptr()),
                                           ctypes.c_void_p(ev.data_ptr()), stream), "gik_fantope")
    return C, ev


def solve_fantope_closed_form(G, d):
    """Reference signature (convex_iteration.py:43-53): one matrix in, (U U^T, seconds) out."""
    t0 = time.perf_counter()
    C, _ = solve_fantope_closed_form_batch(np.asarray(G, dtype=float)[None], d)
    return C[0].cpu().numpy(), time.perf_counter() - t0


def solve_with_cidgik(graph, T_goal):
    raise NotImplementedError(
        "CIDGIK's semidefinite programs are solved by MOSEK through cvxpy in the reference (sdp_snl.py:874-967); neither is "
        "available to restate or to pin results against, so graphik_b200 ships only the closed-form Fantope step "
        "(solve_fantope_closed_form).  Use solve_with_riemannian.")
