"""The explicit time step behind the reference's signature (/root/reference/Tools/Dynamic_solver.py:9-34)."""
import weakref

import numpy as np

from saa_b200 import maps as _maps
from saa_b200 import plan as _plan
from Tools.Mat_construction import *   # noqa: F401,F403  (same re-export chain as the reference, :2-5)
from Tools.commons import *            # noqa: F401,F403
from Tools.Distributed_tools import *  # noqa: F401,F403
import Tools.Distributed_tools as _dt

_plans = {}


def _fingerprint(a):
    """Cheap identity + content check of an input the plan uploaded once: address, shape and the sum of <= 1024
    evenly spaced entries.  A caller that replaces or rescales F_rankwise / l_M / Local_Dirichlet between two calls
    (the reference re-reads them on every call, Dynamic_solver.py:13-32) gets a fresh plan instead of stale device copies."""
    if isinstance(a, np.ndarray):
        v = a.reshape(-1)
        return (a.__array_interface__["data"][0], a.shape, float(v[::max(1, v.size // 1024)].sum()) if v.size else 0.0)
    v = np.asarray(a, dtype=np.int64).reshape(-1)      # Local_Dirichlet is a Python list in the reference drivers
    return (len(v), int(v[::max(1, v.size // 1024)].sum()) if v.size else 0)


def _plan_for(LocalK, F_rankwise, Local_nodes, Local_Dirichlet, T, l_M, alpha, size, rank):
    """One device plan per LocalK object: built (and, for size > 1, given its interface description — a
    collective) at the first call, reused while the matrix object, the scalars and the fingerprints of the load,
    mass and Dirichlet inputs are unchanged."""
    key = id(LocalK)
    sig = (float(T.dt), float(alpha), size, rank, LocalK.data.__array_interface__["data"][0], LocalK.nnz,
           _fingerprint(F_rankwise), _fingerprint(l_M), _fingerprint(Local_Dirichlet))
    hit = _plans.get(key)
    if hit is not None and hit[0]() is LocalK and hit[2] == sig:
        return hit[1]
    if hit is not None:                                                # same matrix object, different inputs: rebuild
        _plans.pop(key)[1].close()
    for k in [k for k, v in _plans.items() if v[0]() is None]:      # matrices that no longer exist: release their plans
        _plans.pop(k)[1].close()
    halo = None
    if size != 1:
        nodes = np.asarray(Local_nodes, dtype=np.int64)
        c = _dt.comm
        lists = c.allgather(nodes) if hasattr(c, "allgather") else c.bcast(c.gather(nodes, root=0), root=0)
        halo = _maps.halo_plan(rank, size, lists)
    p = _plan.StepPlan(LocalK, F_rankwise, l_M, Local_Dirichlet, T.dt, alpha, halo=halo, rank=rank, size=size)
    _plans[key] = (weakref.ref(LocalK), p, sig)
    return p


def parallel_explicit_solver_dis_pre(LocalK, F_rankwise, Points, Local_nodes, Local_Dirichlet,
                                     T, Elas, l_M, alpha, size, rank, MODEL=False):
    """d_{n+1} from T = (tn, dt, d0 = d_n, dn = d_{n-1}); returns a fresh writable (3n,1) float64 array.

    MODEL == False and size != 1: collective — the partial internal forces of shared nodes are summed over
    their holders in ascending rank order (syn_cpus semantics) before the update.  MODEL == True: local.
    T.d0 / T.dn are not modified.
    """
    p = _plan_for(LocalK, F_rankwise, Local_nodes, Local_Dirichlet, T, l_M, alpha, size, rank)
    if MODEL or size == 1:
        d1 = p.step_host(T.d0, T.dn, T.tn, _plan.MODE_LOCAL)
    else:
        p.set_state(T.d0, T.dn, T.tn)
        p.step_exchange(_dt.comm.exchange)
        d1 = p.d0()
    return d1.reshape(-1, 1)
