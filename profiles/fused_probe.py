"""Single-GPU timing probe of the fused synchronised step (profiles only, not part of the product or the tests).

Builds rank 0 of a 2-slab partition of the m=65 beam on ONE GPU, attaches the peer transport to the plan itself
(SAA_DEBUG_PEER_SELF=1: the 'neighbour' receive area is the plan's own), and times MODE_SYNC (saa_k_step_fused)
against MODE_LOCAL (saa_k_step) on identical data.  Results are not physically meaningful; the timings are.
"""
import os, sys, json
os.environ["SAA_DEBUG_PEER_SELF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import saa_b200
from saa_b200 import device_setup as ds, plan as splan

m = int(sys.argv[1]) if len(sys.argv) > 1 else 65
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
loc = ds.structured_rank_local(m, 0, 2)
import numpy as np
from saa_b200 import maps
# the neighbour's node list: only its interface layer matters -> build it from rank 1's local data
loc1 = ds.structured_rank_local(m, 1, 2)
lists = [loc["local_nodes"].cpu().numpy(), loc1["local_nodes"].cpu().numpy()]
loc1["K"].free(); del loc1
halo = maps.halo_plan(0, 2, lists)
pl, info = ds.structured_rank_plan(loc, halo, {1: torch.zeros((len(halo["send_idx"][1]), 4), dtype=torch.float64, device="cuda")}, loc["dt_loc"])
exp = pl.peer_export()
fake = dict(rank=1, handle=exp["handle"], nb=[0], off=exp["off"])
pl.peer_attach([exp, fake])
st = torch.cuda.ExternalStream(pl.stream)
def t(mode, n):
    pl.step(50, mode); pl.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); pl.step(n, mode); e1.record(st); pl.synchronize()
    return e0.elapsed_time(e1) / n
out = {}
for rep in range(2):
    out[f"local_{rep}"] = t(splan.MODE_LOCAL, steps)
    out[f"sync_{rep}"] = t(splan.MODE_SYNC, steps)
out["dbg"] = os.environ.get("SAA_DEBUG_PEER", "0")
print(json.dumps(out))
