"""Partition maps and the shared-node force synchronisation — names of /root/reference/Tools/Distributed_tools.py.

The maps are the vectorised, sequence-exact restatements of `saa_b200.maps`; `syn_cpus` runs the pack /
rank-ordered-sum kernels of the device plan and moves only the shared-node partial forces between
neighbours (the reference gathers whole vectors to rank 0 and broadcasts the global vector, :77-92 — the
numbers returned are identical, see tests/test_gpu_tools.py).
"""
import numpy as np

from saa_b200 import comm as _comm
from saa_b200 import maps as _maps
from Tools.commons import *  # noqa: F401,F403  (the reference re-exports commons through this module)

comm = _comm.world()
rank = comm.Get_rank()


def rankwise_dist(rank, recvbuf, Points, Cells):
    """(element ids of `rank` ascending, their nodes in first-appearance order) — :14-24."""
    return _maps.rankwise_dist(rank, recvbuf, Cells)


def find_shared_nodes(rank, size, rank_nodal_num, rank_nodal_list):
    """Nodes also held by other ranks, other-rank-major order — :29-40."""
    return _maps.find_shared_nodes(rank, size, rank_nodal_list)


def sort_shared(G_shared_nodes):
    """Sorted union of the ranks' shared lists — :44-51."""
    return _maps.sort_shared(G_shared_nodes)


def Dirichlet_rank_dist(D_node, Local_N_list):
    """Local clamped DOF ids — :55-62."""
    return _maps.Dirichlet_rank_dist(D_node, Local_N_list)


def local_mat_node(G_ID, L_N):
    """Local positions of global node ids — :66-73."""
    return _maps.local_mat_node(G_ID, L_N)


_halo_only_plans = {}


def _halo_plan_for(size, rank, Local_nodes):
    """Device plan that only carries the interface description of this rank (no matrix).  Cached under a digest of
    the whole node list; whether to (re)build is decided COLLECTIVELY, so that the allgather below is entered by
    every rank or by none (a rank-local cache miss would deadlock the others)."""
    import hashlib
    from scipy.sparse import csr_matrix
    from saa_b200 import plan as _plan
    nodes = np.ascontiguousarray(Local_nodes, dtype=np.int64)
    key = (size, rank, nodes.size, hashlib.blake2b(nodes.tobytes(), digest_size=16).digest())
    p = _halo_only_plans.get(key)
    gather = (lambda v: comm.allgather(v)) if hasattr(comm, "allgather") else (lambda v: comm.bcast(comm.gather(v, root=0), root=0))
    if any(gather(p is None)):
        lists = gather(nodes)
        if p is None:
            n = 3 * nodes.size
            p = _plan.StepPlan(csr_matrix((n, n)), np.zeros(n), np.ones(n), np.zeros(0, dtype=np.int64), 1.0, 0.0,
                               halo=_maps.halo_plan(rank, size, lists), rank=rank, size=size)
            _halo_only_plans[key] = p
    return p


def syn_cpus(size, rank, f, L_g, Local_nodes):
    """f_global[dofs_local] with f_global = sum over ranks (ascending) of the scattered partial forces — :77-92.
    Collective over `comm`; returns a fresh (3n,1) array."""
    p = _halo_plan_for(size, rank, Local_nodes)
    return p.sync_forces(f, comm.exchange).reshape(-1, 1)
