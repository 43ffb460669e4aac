"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol the
header declares; the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


def test_library_exports_every_declared_symbol():
    from graphik_b200 import _lib
    path = _lib.build()
    L = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "graphik_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gik_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(L, sym), "missing export " + sym
    assert sorted(_lib.EXPORTS) == declared
    assert L.gik_version() >= 100


def test_default_opts_are_the_reference_defaults():
    from graphik_b200 import _lib
    from graphik_b200.engine import make_opts
    o = make_opts()
    assert (o.mingradnorm, o.maxiter, o.theta, o.kappa) == (0.5e-9, 3000, 1.0, 0.1)   # riemannian_solver.py:44-50
    assert (o.rho_prime, o.rho_regularization, o.mininner, o.maxinner) == (0.1, 1e3, 1, 10000)
    assert (o.Delta_bar, o.Delta0) == (13.0, 13.0 / 8)
    o = make_opts({"maxiter": 10, "Delta_bar": 4.0})
    assert o.maxiter == 10 and o.Delta0 == 0.5


def test_null_arguments_are_rejected_without_touching_the_gpu():
    from graphik_b200 import _lib
    L = _lib.load()
    assert L.gik_plan_create(None, None) == -1
    assert b"null" in L.gik_last_error()
    assert L.gik_default_opts(None) == -1
    assert L.gik_plan_destroy(None) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from graphik_b200._lib import GikError
    from graphik_b200.engine import BatchIK
    from graphik_b200.utils.roboturdf import load_ur10
    robot, graph = load_ur10()
    with pytest.raises(GikError):
        BatchIK(graph)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graphik_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle/gen_golden.py", ""), \
                    os.path.join(dirpath, f)


def test_kernel_selector_names_match_the_header():
    """engine.make_opts({"kernel": name}) must send the GIK_KERNEL_* value include/graphik_b200.h defines."""
    from graphik_b200.engine import make_opts
    header = open(os.path.join(ROOT, "include", "graphik_b200.h")).read()
    defines = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define GIK_KERNEL_([A-Z]+)\s+(\d+)", header)}
    assert set(defines) == {"auto", "latency", "throughput", "generic", "dense"}
    for name, value in defines.items():
        assert make_opts({"kernel": name}).kernel == value, name
    with pytest.raises(KeyError):
        make_opts({"kernel": "no-such-kernel"})


def test_every_cuda_source_is_in_the_build_list():
    """A .cu file that is not in _lib.SOURCES would silently stay out of libgraphik_b200.so."""
    from graphik_b200 import _lib
    on_disk = sorted(f for f in os.listdir(_lib.CSRC) if f.endswith(".cu"))
    assert on_disk == sorted(_lib.SOURCES)
