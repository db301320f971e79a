"""ctypes front end of oracle/fem_oracle.c (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libfem_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "fem_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, i64, f64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
        L.oracle_problem_create.restype = vp
        L.oracle_problem_create.argtypes = [i32, i64]
        L.oracle_problem_destroy.argtypes = [vp]
        L.oracle_problem_set_rank.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, i64, vp, vp]
        L.oracle_problem_set_state.argtypes = [vp, i32, vp, vp, f64]
        L.oracle_problem_get_state.argtypes = [vp, i32, vp, vp, vp]
        L.oracle_problem_run.argtypes = [vp, i64, f64, f64, f64, f64, f64, i32]
        L.oracle_csr_matvec.argtypes = [i64, vp, vp, vp, vp, vp]
        L.oracle_cd_update.argtypes = [i64, f64, f64, f64, f64, f64, f64, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def step_scalars(dt, alpha):
    """The scalar sub-expressions of Dynamic_solver.py:17 evaluated with the reference's own Python
    expressions on the reference's own types (dt is np.float64, alpha a Python float)."""
    dt = np.float64(dt)
    return float(dt), float(dt ** 2), float(dt / 2), float(0.5 * alpha), float(alpha)


class OracleProblem:
    """All ranks of one partitioned run, stepped by oracle_problem_run (Data_prepare.py:223-235)."""

    def __init__(self, n_global_nodes, ranks, dt, alpha):
        """ranks: list of dicts with K_indptr, K_indices, K_data, F, lM, dirichlet, nodes."""
        self.L = lib()
        self.size = len(ranks)
        self.h = self.L.oracle_problem_create(self.size, int(n_global_nodes))
        self._keep = []
        self.n_dof = []
        for r, q in enumerate(ranks):
            arrs = dict(indptr=np.ascontiguousarray(q["K_indptr"], dtype=np.int32),
                        indices=np.ascontiguousarray(q["K_indices"], dtype=np.int32),
                        data=np.ascontiguousarray(q["K_data"], dtype=np.float64),
                        F=np.ascontiguousarray(q["F"], dtype=np.float64).reshape(-1),
                        M=np.ascontiguousarray(q["lM"], dtype=np.float64).reshape(-1),
                        dir=np.ascontiguousarray(q["dirichlet"], dtype=np.int64),
                        nodes=np.ascontiguousarray(q["nodes"], dtype=np.int64))
            self._keep.append(arrs)
            n = arrs["F"].size
            self.n_dof.append(n)
            self.L.oracle_problem_set_rank(self.h, r, n, _p(arrs["indptr"]), _p(arrs["indices"]), _p(arrs["data"]),
                                           _p(arrs["F"]), _p(arrs["M"]), arrs["dir"].size, _p(arrs["dir"]),
                                           _p(arrs["nodes"]))
        self.dt, self.dt2, self.dt_half, self.half_alpha, self.alpha = step_scalars(dt, alpha)

    def run(self, n_steps, model=False):
        self.L.oracle_problem_run(self.h, int(n_steps), self.dt, self.dt2, self.dt_half, self.half_alpha,
                                  self.alpha, 1 if model else 0)

    def set_state(self, r, d0, dn, tn):
        d0 = np.ascontiguousarray(d0, dtype=np.float64).reshape(-1)
        dn = np.ascontiguousarray(dn, dtype=np.float64).reshape(-1)
        self.L.oracle_problem_set_state(self.h, r, _p(d0), _p(dn), float(tn))

    def state(self, r):
        d0 = np.empty(self.n_dof[r])
        dn = np.empty(self.n_dof[r])
        tn = ctypes.c_double(0)
        self.L.oracle_problem_get_state(self.h, r, _p(d0), _p(dn), ctypes.byref(tn))
        return d0, dn, tn.value

    def d0(self, r):
        return self.state(r)[0]

    def close(self):
        if self.h:
            self.L.oracle_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def csr_matvec(indptr, indices, data, x):
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    y = np.empty(indptr.size - 1)
    lib().oracle_csr_matvec(indptr.size - 1, _p(indptr), _p(indices), _p(data), _p(x), _p(y))
    return y
