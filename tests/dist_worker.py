"""Worker launched under torch.distributed.run by tests/test_tools_shim.py (world_size 2..4, gloo).

mode "host": host-side logic of the N>1 path without any GPU — communicator facade, distributed partition
call, collectively built maps, halo plans of all ranks consistent with each other, and the neighbour exchange
(`comm.exchange`) + ascending-rank sum reproducing the literal syn_cpus (Distributed_tools.py:77-92).
mode "gpu": the same exchange driven through the device plan (`Tools.Distributed_tools.syn_cpus`).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "synchronization-avoiding-algorithms_b200")
sys.path[:0] = [PKG, os.path.join(ROOT, "tests")]
sys.path.append(os.path.join(PKG, "compat"))

mode, golden = sys.argv[1], sys.argv[2]
from mpi4py import MPI                                   # real mpi4py if installed, else the stand-in
from mgmetis.parmetis import part_mesh_kway
from Tools.commons import node_to_dof
import Tools.Distributed_tools as DT
from saa_b200 import maps
from util import bits_equal, load_golden

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
g = load_golden(golden)
assert g["P"] == size
Cells, Points = g["cells"], g["points"]

# collectives of the facade
assert comm.bcast("x" if rank == 0 else None, root=0) == "x"
got = comm.gather(rank * 10, root=0)
assert got == [10 * r for r in range(size)] if rank == 0 else got is None
buf = np.empty(size, dtype=float) if rank == 0 else None
comm.Gather(np.float64(rank + 0.5), buf, root=0)
if rank == 0:
    assert np.array_equal(buf, np.arange(size) + 0.5)

# distributed partition call on contiguous chunks (Data_prepare.py:66-101)
bounds = np.linspace(0, len(Cells), size + 1).astype(int)
mine = Cells[bounds[rank]:bounds[rank + 1]]
_, ep = part_mesh_kway(size, 4 * np.arange(len(mine) + 1), mine.reshape(-1))
recv = np.empty(len(Cells), dtype=int) if rank == 0 else None
comm.Gatherv(ep, recv, root=0)
epart = comm.bcast(recv, root=0)
if "beam_coarse" in golden:
    assert np.array_equal(epart, g["epart"])             # same METIS call as the fixture's partition
epart = g["epart"]

# maps, built the way the driver builds them
ele, nodes = DT.rankwise_dist(rank, epart, Points, Cells)
lists = comm.bcast(comm.gather(nodes, root=0), root=0)
shared = DT.find_shared_nodes(rank, size, [len(x) for x in lists], lists)
assert np.array_equal(nodes, g["ranks"][rank]["nodes"]) and np.array_equal(shared, g["ranks"][rank]["shared"])

# literal syn_cpus on seeded partial forces
rng = np.random.default_rng(100 + rank)
f = rng.standard_normal(3 * len(nodes)) * 10.0 ** rng.integers(-3, 3, 3 * len(nodes))
allf = comm.bcast(comm.gather(f, root=0), root=0)
fg = np.zeros(3 * len(Points))
for r in range(size):
    fg[node_to_dof(3, [0, 1, 2], lists[r])] += allf[r]
want = fg[node_to_dof(3, [0, 1, 2], nodes)]

if mode == "host":
    hp = maps.halo_plan(rank, size, lists)
    nb = np.asarray(hp["neighbours"], dtype=np.int32)
    off = np.zeros(len(nb) + 1, dtype=np.int64)
    for k, r in enumerate(nb):
        off[k + 1] = off[k] + 3 * hp["send_idx"][int(r)].size
    own = f.reshape(-1, 3)
    send = np.concatenate([own[hp["shared_pos"][hp["send_idx"][int(r)]]].reshape(-1) for r in nb]) if len(nb) else np.zeros(0)
    got = DT.comm.exchange(send, nb, off)
    out = (0.0 + f).reshape(-1, 3)
    for j, pos in enumerate(hp["shared_pos"]):
        acc = np.zeros(3)
        for k in range(hp["holders_ptr"][j], hp["holders_ptr"][j + 1]):
            hr, slot = int(hp["holders_rank"][k]), int(hp["holders_slot"][k])
            if slot < 0:
                acc = acc + own[pos]
            else:
                kk = int(np.nonzero(nb == hr)[0][0])
                acc = acc + got[off[kk] + 3 * slot: off[kk] + 3 * slot + 3]
        out[pos] = acc
    assert bits_equal(out, want)
    # the set-up exchanges of the device path (device_setup.gather_node_lists / dist_exchange / holders_sum) over a gloo
    # group pass through host memory: rank-ordered sums of per-node partials (lumped mass, load) at the shared nodes
    import torch
    from saa_b200 import device_setup as ds
    lists2 = ds.gather_node_lists(torch.as_tensor(np.asarray(nodes, dtype=np.int64)), size)
    assert len(lists2) == size and all(np.array_equal(a, b) for a, b in zip(lists2, lists))
    part = rng.standard_normal((len(nodes), 2)) * 10.0 ** rng.integers(-3, 3, (len(nodes), 2))
    own_t = torch.as_tensor(part)[torch.as_tensor(np.asarray(hp["shared_pos"], dtype=np.int64))]
    tot = ds.holders_sum(hp, rank, own_t, ds.dist_exchange(ds.holders_send(hp, own_t)))
    allp = comm.bcast(comm.gather(part, root=0), root=0)
    glob = np.zeros((len(Points), 2))
    for r in range(size):
        glob[np.asarray(lists[r], dtype=np.int64)] += allp[r]
    assert bits_equal(tot.numpy(), glob[np.asarray(nodes, dtype=np.int64)[np.asarray(hp["shared_pos"], dtype=np.int64)]])
else:
    got = DT.syn_cpus(size, rank, f.reshape(-1, 1), len(Points), nodes)
    assert got.shape == (f.size, 1) and bits_equal(got, want)
comm.Barrier()
print(f"rank {rank}/{size} ok ({mode})")
