"""Process-group facade with the mpi4py surface the reference drivers use.

The reference talks to `MPI.COMM_WORLD` (module global `comm`, /root/reference/Tools/Distributed_tools.py:9-11;
Data_prepare.py:13-15): lower-case pickle collectives `bcast` / `gather` and buffer collectives `Gather` /
`Gatherv`.  `world()` returns the real mpi4py communicator when mpi4py is importable, otherwise an object with
the same methods on top of `torch.distributed` (launched by torchrun; gloo for host objects), otherwise a
single-process stand-in.  On top of that, `exchange` moves the per-neighbour halo messages of
`StepPlan.step_exchange` through host memory with whatever communicator is in use.
"""
from __future__ import annotations

import os

import numpy as np


class SerialComm:
    """size == 1: every collective is the identity."""

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, obj, root=0):
        return obj

    def gather(self, obj, root=0):
        return [obj]

    def allgather(self, obj):
        return [obj]

    def Gather(self, sendbuf, recvbuf, root=0):
        if recvbuf is not None:
            np.asarray(recvbuf).reshape(-1)[:] = np.asarray(sendbuf).reshape(-1)

    Gatherv = Gather

    def Barrier(self):
        pass

    def exchange(self, send, nb, off):
        return np.zeros(0)


class TorchDistComm:
    """mpi4py-shaped view of the default torch.distributed process group."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        if not dist.is_initialized():
            backend = os.environ.get("SAA_DIST_BACKEND", "gloo")
            dist.init_process_group(backend)
        # host objects and host halo messages travel over gloo whatever the default backend is
        self.host_group = None
        if dist.get_backend() != "gloo":
            self.host_group = dist.new_group(backend="gloo")

    def Get_rank(self):
        return self.dist.get_rank()

    def Get_size(self):
        return self.dist.get_world_size()

    def bcast(self, obj, root=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=root, group=self.host_group)
        return box[0]

    def gather(self, obj, root=0):
        out = [None] * self.Get_size() if self.Get_rank() == root else None
        self.dist.gather_object(obj, out, dst=root, group=self.host_group)
        return out

    def allgather(self, obj):
        out = [None] * self.Get_size()
        self.dist.all_gather_object(out, obj, group=self.host_group)
        return out

    def Gatherv(self, sendbuf, recvbuf, root=0):
        """Concatenate the ranks' (differently sized) arrays into recvbuf on root (Data_prepare.py:100)."""
        parts = self.gather(np.asarray(sendbuf).reshape(-1), root)
        if self.Get_rank() == root:
            np.asarray(recvbuf).reshape(-1)[:] = np.concatenate(parts)

    Gather = Gatherv

    def Barrier(self):
        self.dist.barrier(group=self.host_group)

    def exchange(self, send, nb, off):
        """Neighbour-wise halo exchange: send[off[k]:off[k+1]] -> rank nb[k]; returns what they sent here."""
        import torch
        recv = np.empty_like(send)
        ops = []
        ts, tr = torch.from_numpy(send), torch.from_numpy(recv)
        for k, r in enumerate(nb):
            a, b = int(off[k]), int(off[k + 1])
            if b > a:
                ops.append(self.dist.P2POp(self.dist.isend, ts[a:b], int(r), group=self.host_group))
                ops.append(self.dist.P2POp(self.dist.irecv, tr[a:b], int(r), group=self.host_group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        return recv


class Mpi4pyComm:
    """Thin wrapper that adds `exchange` / `allgather` to a real mpi4py communicator."""

    def __init__(self, comm):
        self._c = comm

    def __getattr__(self, name):
        return getattr(self._c, name)

    def exchange(self, send, nb, off):
        from mpi4py import MPI
        recv = np.empty_like(send)
        reqs = []
        for k, r in enumerate(nb):
            a, b = int(off[k]), int(off[k + 1])
            if b > a:
                reqs.append(self._c.Irecv([recv[a:b], MPI.DOUBLE], source=int(r), tag=77))
                reqs.append(self._c.Isend([send[a:b], MPI.DOUBLE], dest=int(r), tag=77))
        MPI.Request.Waitall(reqs)
        return recv


_world = None


def world():
    """The communicator of this process: mpi4py if present, torch.distributed under torchrun, else serial."""
    global _world
    if _world is None:
        try:
            from mpi4py import MPI  # noqa: WPS433
            if getattr(MPI, "_saa_stub", False):
                raise ImportError
            _world = Mpi4pyComm(MPI.COMM_WORLD)
        except Exception:
            if int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("SAA_FORCE_TORCH_DIST"):
                _world = TorchDistComm()
            else:
                _world = SerialComm()
    return _world


def reset_world():
    global _world
    _world = None
