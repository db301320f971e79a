"""oracle/numpy_step.py — the reference's CPU path as numpy/scipy statements (TEST INFRASTRUCTURE ONLY).

Only tests/ and the cpu_baseline / --impl reference legs of bench.py may import this module; the product never
does.  /root/reference does not exist on the GPU box, so the reference's per-step statement sequence is restated
here, operator for operator, and pinned against golden histories produced by the unmodified reference
(tests/test_oracle.py::test_numpy_step_*):

  step()        /root/reference/Tools/Dynamic_solver.py:12-32 — scipy `LocalK.dot(d0)`, the numpy update
                expression evaluated with F_int local (:17), Dirichlet rows zeroed (:20) and, for size != 1 and
                MODEL == False, syn_cpus (:26) followed by the SAME expression again (:29) and the zeroing (:32).
  syn_cpus()    /root/reference/Tools/Distributed_tools.py:77-92 — every rank sends its whole force vector and its
                node list (a Python list of ints, pickled, like mpi4py's lowercase gather) to rank 0; rank 0 builds
                the DOF lists with the list comprehension of commons.py:66-71 (`node_to_dof`) and scatter-adds in
                ascending rank order; the global vector is broadcast (pickled) and indexed with the local DOF list.
  run_ranks()   P single-threaded processes (what `mpirun -np P` gives; pipes carry the pickles), the loop of
                /root/reference/Data_prepare.py:223-235.

This is the baseline BASELINE.md describes ("the reference's mpirun CPU path on the box's own host cores").
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def linear_ramp(t):                                   # commons.py:7-11
    if t <= 1:
        return t
    return 1.0


def node_to_dof(dim, var, nodes):                     # commons.py:66-71 (a Python loop over the node list)
    dofs = []
    for n in nodes:
        for i in var:
            dofs.append(dim * n + i)
    return dofs


class _Comm:
    """gather / bcast of Python objects between P processes over pipes (pickle), root = 0."""

    def __init__(self, rank, size, to_root, from_root):
        self.rank, self.size, self.to_root, self.from_root = rank, size, to_root, from_root

    def gather(self, obj):
        if self.rank == 0:
            return [obj] + [c.recv() for c in self.to_root]
        self.to_root.send(obj)
        return None

    def bcast(self, obj):
        if self.rank == 0:
            for c in self.from_root:
                c.send(obj)
            return obj
        return self.from_root.recv()


def syn_cpus(comm, size, rank, f, L_g, Local_nodes):
    """Distributed_tools.py:77-92"""
    f_all = comm.gather(f)
    n_all = comm.gather(Local_nodes)
    f_global = None
    if rank == 0:
        f_global = np.zeros((3 * L_g, 1))
        for i in range(size):                         # :85-86 ascending rank
            f_global[node_to_dof(3, [0, 1, 2], n_all[i])] += f_all[i]
    f_global = comm.bcast(f_global)
    return f_global[node_to_dof(3, [0, 1, 2], Local_nodes)]


def step(comm, LocalK, F_rankwise, L_g, Local_nodes, Local_Dirichlet, tn, dt, d0, dn, l_M, alpha, size, rank, MODEL=False):
    """Dynamic_solver.py:12-32; all vectors are (3n,1) columns like the reference's."""
    F_int = LocalK.dot(d0)                                                                       # :12
    F_ext = F_rankwise * linear_ramp(tn)                                                         # :13
    d1 = (dt ** 2 * (F_ext - F_int) + 2 * l_M * d0 - l_M * dn + dt / 2 * l_M * alpha * dn) / (l_M + 0.5 * alpha * l_M * dt)   # :17
    d1[Local_Dirichlet] = 0                                                                      # :20
    if MODEL == False:                                                                           # noqa: E712  (:22)
        if size != 1:                                                                            # :25
            F_int = syn_cpus(comm, size, rank, F_int, L_g, Local_nodes)                          # :26
            d1 = (dt ** 2 * (F_ext - F_int) + 2 * l_M * d0 - l_M * dn + dt / 2 * l_M * alpha * dn) / (l_M + 0.5 * alpha * l_M * dt)   # :29
            d1[Local_Dirichlet] = 0                                                              # :32
    return d1


def _worker(rank, size, q, n_global_nodes, dt, alpha, n_steps, warmup, model, to_root, from_root, out):
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    from scipy.sparse import csr_matrix
    comm = _Comm(rank, size, to_root, from_root)
    n = q["F"].size
    K = csr_matrix((q["K_data"], q["K_indices"], q["K_indptr"]), shape=(n, n))
    F = np.asarray(q["F"], dtype=np.float64).reshape(-1, 1)
    lM = np.asarray(q["lM"], dtype=np.float64).reshape(-1, 1)
    nodes = [int(v) for v in q["nodes"]]              # Local_nodal_list is a Python list in the reference
    dirichlet = [int(v) for v in q["dirichlet"]]
    dt = np.float64(dt)
    d_0 = np.zeros((n, 1))
    d_n = np.zeros((n, 1))
    tn = 0
    t0 = None
    for i in range(warmup + n_steps):                 # Data_prepare.py:223-235
        if i == warmup:
            comm.bcast(comm.gather(0) and 0)          # line the ranks up before the timed part
            t0 = time.perf_counter()
        d1 = step(comm, K, F, n_global_nodes, nodes, dirichlet, tn, dt, d_0, d_n, lM, alpha, size, rank, MODEL=model)
        d_n = d_0
        d_0 = d1
        tn = tn + dt
    comm.bcast(comm.gather(0) and 0)
    secs = time.perf_counter() - t0
    out.put((rank, secs, d_0.reshape(-1) if q.get("want_state") else None))


def run_ranks(ranks, n_global_nodes, dt, alpha, n_steps, warmup=0, model=False, want_state=False):
    """Run the loop on len(ranks) single-threaded processes.  ranks: dicts with K_indptr / K_indices / K_data / F / lM /
    dirichlet / nodes.  Returns (seconds of the timed part = max over ranks, [final d0 per rank] or None)."""
    size = len(ranks)
    ctx = mp.get_context("fork")
    up = [ctx.Pipe(duplex=False) for _ in range(size - 1)]      # rank r+1 -> root
    down = [ctx.Pipe(duplex=False) for _ in range(size - 1)]    # root -> rank r+1
    out = ctx.Queue()
    procs = []
    for r in range(size):
        q = dict(ranks[r])
        q["want_state"] = want_state
        to_root = [u[0] for u in up] if r == 0 else up[r - 1][1]
        from_root = [d[1] for d in down] if r == 0 else down[r - 1][0]
        p = ctx.Process(target=_worker, args=(r, size, q, n_global_nodes, dt, alpha, n_steps, warmup, model, to_root, from_root, out))
        p.start()
        procs.append(p)
    res = [out.get() for _ in range(size)]
    for p in procs:
        p.join()
    res.sort(key=lambda t: t[0])
    return max(t[1] for t in res), ([t[2] for t in res] if want_state else None)
