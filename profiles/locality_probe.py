"""Gather-locality probe (profiles only): the same mesh with (a) the structured element order, (b) a random element order
(what an incoherent mesh file would give: Local_nodal_list is first-appearance order), each with and without the Morton
node-order hint.  One GPU, ms per step."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import saa_b200
from saa_b200 import device_setup as ds, mesh, plan as splan

m = int(sys.argv[1]) if len(sys.argv) > 1 else 40
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nx, ny, nz = mesh.structured_beam_dims(m)
out = {"m": m}
for elem_order in ("structured", "random"):
    for reorder in (None, "morton"):
        cells = ds.structured_slab_cells(m, 0, 1)
        if elem_order == "random":
            torch.manual_seed(0)
            cells = cells[torch.randperm(cells.shape[0], device=cells.device)]
        loc = ds.rank_local(cells, lambda ids: ds.structured_points(m, ids), lambda ids: ids < (ny + 1) * (nz + 1), 0, 1, reorder=reorder)
        del cells
        pl, info = ds.structured_rank_plan(loc, None, None, loc["dt_loc"])
        st = torch.cuda.ExternalStream(pl.stream)
        pl.step(30, splan.MODE_LOCAL); pl.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); pl.step(steps, splan.MODE_LOCAL); e1.record(st); pl.synchronize()
        out[f"{elem_order}/{reorder}"] = {"ms_per_step": e0.elapsed_time(e1) / steps, "n_dof": pl.n_dof,
                                          "G_dof_steps_per_s": pl.n_dof * steps / e0.elapsed_time(e1) / 1e6}
        pl.close(); del pl, loc
        torch.cuda.empty_cache()
print(json.dumps(out))
