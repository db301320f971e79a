"""Element -> rank partition vectors (the `epart` of /root/reference/Data_prepare.py:94-101).

The reference calls ParMETIS through `mgmetis.parmetis.part_mesh_kway(size, eptr, eind)`
(Data_prepare.py:94), which is not installable offline.  CUDA 12.9 ships a complete serial METIS
(`libmetis_static.a`, 64-bit idx_t, 32-bit real_t); `build_metis()` links it into
`csrc/libsaa_metis.so` and `metis_part_mesh` calls `METIS_PartMeshDual` (same dual-graph k-way
objective ParMETIS_V3_PartMeshKway optimises; ncommon = 3: tets sharing a face).  The exact
ParMETIS output of the reference is recorded nowhere, so the partition is an input of every
parity comparison.  `slab_partition` is the fallback for very large structured
beams: balanced slabs along x by element centroid.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_METIS_SO = os.path.join(_HERE, "csrc", "libsaa_metis.so")
_METIS_A = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"
_lib = None


def build_metis(force=False):
    """Link CUDA's static METIS into a shared object next to the kernels (no sources involved)."""
    if os.path.isfile(_METIS_SO) and not force:
        return _METIS_SO
    if not os.path.isfile(_METIS_A):
        raise RuntimeError(f"{_METIS_A} not found; cannot build METIS")
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-o", _METIS_SO, "-Wl,--whole-archive", _METIS_A,
                           "-Wl,--no-whole-archive", "-lm"])
    return _METIS_SO


def _metis():
    global _lib
    if _lib is None:
        if not os.path.isfile(_METIS_SO):
            build_metis()
        _lib = ctypes.CDLL(_METIS_SO)
    return _lib


def metis_part_mesh(cells, n_nodes, nparts, ncommon=3):
    """METIS_PartMeshDual on a tet mesh -> epart (nE,) int64.  nparts == 1 returns zeros."""
    cells = np.ascontiguousarray(cells, dtype=np.int64)
    ne = cells.shape[0]
    if nparts == 1:
        return np.zeros(ne, dtype=np.int64)
    lib = _metis()
    i64 = ctypes.c_int64
    eptr = np.arange(0, 4 * ne + 1, 4, dtype=np.int64)
    eind = cells.reshape(-1).copy()
    epart = np.zeros(ne, dtype=np.int64)
    npart = np.zeros(n_nodes, dtype=np.int64)
    objval = i64(0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.METIS_PartMeshDual(ctypes.byref(i64(ne)), ctypes.byref(i64(n_nodes)), p(eptr), p(eind), None, None,
                                ctypes.byref(i64(ncommon)), ctypes.byref(i64(nparts)), None, None,
                                ctypes.byref(objval), p(epart), p(npart))
    if rc != 1:  # METIS_OK
        raise RuntimeError(f"METIS_PartMeshDual failed with code {rc}")
    return epart


def slab_partition(points, cells, nparts, axis=0):
    """Balanced slabs along `axis` by element centroid (ties broken by element id)."""
    cells = np.asarray(cells)
    c = points[cells[:, :4], axis].sum(axis=1)
    order = np.argsort(c, kind="stable")
    epart = np.empty(cells.shape[0], dtype=np.int64)
    epart[order] = (np.arange(cells.shape[0], dtype=np.int64) * nparts) // cells.shape[0]
    return epart
