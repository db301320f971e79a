"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/saa_fem.h declares,
and refuses to compute without a CUDA device (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest

import saa_b200  # noqa: F401
from saa_b200 import plan as splan
from util import ROOT, load_golden


def header_symbols():
    src = open(os.path.join(ROOT, "include", "saa_fem.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(saa_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    splan.build()
    L = ctypes.CDLL(splan.library_path())
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/saa_fem.h but not exported"
    # the Python binding declares exactly the header's functions
    assert sorted(splan.ABI) == syms
    assert splan.lib().saa_version() >= 100


def test_compute_fails_loudly_without_gpu():
    if splan.device_count() > 0:
        pytest.skip("a CUDA device is present")
    g = load_golden("beam_coarse_P1")
    import scipy.sparse as sp
    r = g["ranks"][0]
    n = r["F"].size
    K = sp.csr_matrix((r["K_data"], r["K_indices"], r["K_indptr"]), shape=(n, n))
    with pytest.raises(splan.SaaError, match="no CPU fallback"):
        splan.StepPlan(K, r["F"], r["lM"], r["dirichlet"], g["dt"], 0.5)


def test_product_package_never_touches_the_oracle():
    """The product must not import / link / execute anything under oracle/."""
    pkg = os.path.join(ROOT, "synchronization-avoiding-algorithms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "fem_oracle" not in txt and "ref_harness" not in txt, os.path.join(dirpath, f)


def test_step_scalars_follow_python_expressions():
    dt = np.float64(0.00024784067462642383)
    s = splan.step_scalars(dt, 0.5)
    assert s == (float(dt), float(dt ** 2), float(dt / 2), 0.25, 0.5)


def test_argument_errors_are_reported_before_any_device_work():
    """Shape errors come back as status + message, whatever the machine (checked before the device is touched)."""
    L = splan.lib()
    h = ctypes.c_void_p()
    ip = np.array([0, 1, 2, 3, 4], dtype=np.int32)
    ix = np.zeros(4, dtype=np.int32)
    dv = np.ones(4)
    v = np.ones(4)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = L.saa_plan_create(ctypes.byref(h), 0, 4, p(ip), p(ix), p(dv), p(v), p(v), None, 0, 1e-3, 1e-6, 5e-4, 0.25, 0.5)
    assert rc != 0 and b"multiple of 3" in L.saa_last_error()
    rc = L.saa_plan_create(ctypes.byref(h), 0, 0, p(ip), p(ix), p(dv), p(v), p(v), None, 0, 1e-3, 1e-6, 5e-4, 0.25, 0.5)
    assert rc != 0 and b"null or empty" in L.saa_last_error()
    assert L.saa_plan_step(None, 1, 0, 0) != 0 and b"null plan" in L.saa_last_error()
    assert L.saa_plan_destroy(None) == 0
