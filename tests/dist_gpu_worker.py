"""Worker for the multi-GPU parity test (one process per GPU, torchrun): a P-way partition stepped with the
peer-memory (NVLink) or NCCL transport must reproduce the golden histories of the reference bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import scipy.sparse as sp
import torch
import torch.distributed as dist

import saa_b200  # noqa: F401
from saa_b200 import maps, multi, plan as splan
from util import bits_equal, load_golden

transport, golden = sys.argv[1], sys.argv[2]
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, size = dist.get_rank(), dist.get_world_size()
g = load_golden(golden)
assert g["P"] == size
r = g["ranks"][rank]
n = r["F"].size
K = sp.csr_matrix((r["K_data"], r["K_indices"], r["K_indptr"]), shape=(n, n))
lists = [q["nodes"] for q in g["ranks"]]
pl = splan.StepPlan(K, r["F"], r["lM"], r["dirichlet"], g["dt"], float(g["alpha"]), device=local,
                    halo=maps.halo_plan(rank, size, lists), rank=rank, size=size)
multi.attach_transport(pl, transport)
done = 0
for s in [int(x) for x in g["steps"]]:
    pl.step(s - done, splan.MODE_SYNC)
    pl.synchronize()
    done = s
    assert bits_equal(pl.d0(), g[f"hist_{s}_r{rank}"]), (golden, transport, s, rank)
# interleave un-synchronised steps (MODEL=True) and a per-step launch; all ranks stay in lockstep
pl.step(5, splan.MODE_LOCAL)
pl.step(3, splan.MODE_SYNC, splan.LAUNCH_PER_STEP)
pl.step(4, splan.MODE_SYNC)
pl.synchronize()
dist.barrier()
print(f"rank {rank}/{size} ok ({transport})")
dist.destroy_process_group()
