"""Worker for the multi-GPU parity tests (one process per GPU, torchrun): a P-way partition stepped with the
peer-memory (NVLink) or NCCL transport must reproduce golden histories bit for bit.

    dist_gpu_worker.py <peer|nccl> <fixture>

fixture = beam_coarse_P{N}: histories of the unmodified reference (110 nodes, one boundary slice per rank);
          mid_np{N}:     CPU-oracle histories of the 29 025-DOF METIS case (oracle/gen_golden_mid.py): several boundary
                            slices and shared-row units per rank, nodes held by >= 3 ranks, device set-up.
When the box has fewer GPUs than ranks the ranks share the GPUs (gloo process group, peer transport only: the
receive areas are mapped through CUDA IPC all the same, the kernels of the ranks time-slice on the device).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import scipy.sparse as sp
import torch
import torch.distributed as dist

import saa_b200  # noqa: F401
from saa_b200 import device_setup, maps, mesh, multi, plan as splan
from util import GOLDEN, bits_equal, load_golden

transport, golden = sys.argv[1], sys.argv[2]
world = int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
local = int(os.environ["LOCAL_RANK"]) % ndev
torch.cuda.set_device(local)
shared_gpu = ndev < world
if shared_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, size = dist.get_rank(), dist.get_world_size()

if golden.startswith("mid_"):
    z = np.load(os.path.join(GOLDEN, golden + ".npz"))
    assert int(z["size"]) == size
    pts, cells, fac = mesh.structured_beam(int(z["m"]), length=int(z["length"]))
    pl, info = device_setup.build_mesh_rank(pts, cells, fac, z["epart"].astype(np.int64), rank, size, device_index=local)
    steps, hist = [int(x) for x in z["steps"]], (lambda s: z[f"hist_{s}_r{rank}"])
    if ndev < int(os.environ["WORLD_SIZE"]):
        steps = [s for s in steps if s <= 60]           # time-sliced ranks: a context switch per step
else:
    g = load_golden(golden)
    assert g["P"] == size
    r = g["ranks"][rank]
    n = r["F"].size
    K = sp.csr_matrix((r["K_data"], r["K_indices"], r["K_indptr"]), shape=(n, n))
    lists = [q["nodes"] for q in g["ranks"]]
    pl = splan.StepPlan(K, r["F"], r["lM"], r["dirichlet"], g["dt"], float(g["alpha"]), device=local,
                        halo=maps.halo_plan(rank, size, lists), rank=rank, size=size)
    steps, hist = [int(x) for x in g["steps"]], (lambda s: g[f"hist_{s}_r{rank}"])
    if shared_gpu:
        steps = [s for s in steps if s <= 100]          # time-sliced ranks: a context switch per step

assert multi.attach_transport(pl, transport) == transport
done = 0
for s in steps:
    pl.step(s - done, splan.MODE_SYNC)
    pl.synchronize()
    done = s
    assert bits_equal(pl.d0(), hist(s)), (golden, transport, s, rank)
# the same steps again from the same start with the other execution forms of a synchronised step: all bit-identical
if transport == "peer":
    d0, dn, tn = pl.get_state()
    ref = None
    for fused in (1, 0):
        pl.set_option(splan.OPT_PEER_FUSED, fused)
        for launch in (splan.LAUNCH_GRAPH, splan.LAUNCH_PER_STEP):
            pl.set_state(d0, dn, tn)
            pl.step(9, splan.MODE_SYNC, launch)
            pl.synchronize()
            got = pl.d0()
            assert ref is None or bits_equal(got, ref), (golden, fused, launch, rank)
            ref = got
    pl.set_option(splan.OPT_PEER_FUSED, 1)
    if not shared_gpu:                                   # cooperative persistent synchronised loop: one launch for all steps
        for n_steps in (9, 4):
            pl.set_state(d0, dn, tn)
            pl.step(n_steps, splan.MODE_SYNC, splan.LAUNCH_PERSISTENT)
            if n_steps == 4:
                pl.step(5, splan.MODE_SYNC)              # continues seamlessly with the per-step form (odd / even buffer parity)
            pl.synchronize()
            assert bits_equal(pl.d0(), ref), (golden, "persistent", n_steps, rank)
# the reference-facing host call in synchronised mode, pipelined (interior chunks as their uploads arrive, boundary
# slices as soon as theirs have, shared rows last) == plain sequence == device-resident steps
dt_ = float(info["dt"]) if golden.startswith("mid_") else float(g["dt"])
d0, dn, tn = pl.get_state()


def host_loop(k):
    a0, an, t = d0.copy(), dn.copy(), tn
    for _ in range(k):
        d1 = pl.step_host(a0, an, t, splan.MODE_SYNC)
        an, a0, t = a0, d1, t + dt_
    return a0


os.environ["SAA_STEP_HOST_PIPELINE"] = "5"
piped = host_loop(6)
assert pl.host_pipe_info(splan.MODE_SYNC)[0] == min(5, pl.n_dof // 3), pl.host_pipe_info(splan.MODE_SYNC)
os.environ["SAA_STEP_HOST_PIPELINE"] = "0"
plain = host_loop(6)
del os.environ["SAA_STEP_HOST_PIPELINE"]
pl.set_state(d0, dn, tn)
pl.step(6, splan.MODE_SYNC)
pl.synchronize()
assert bits_equal(piped, plain) and bits_equal(piped, pl.d0()), (golden, transport, "host call", rank)
# interleave un-synchronised steps (MODEL=True) and a per-step launch; all ranks stay in lockstep
pl.step(5, splan.MODE_LOCAL)
pl.step(3, splan.MODE_SYNC, splan.LAUNCH_PER_STEP)
pl.step(4, splan.MODE_SYNC)
pl.synchronize()
dist.barrier()
print(f"rank {rank}/{size} ok ({transport})")
dist.destroy_process_group()
