#!/usr/bin/env python
"""Synchronization-avoiding run — the flow of the reference's Online_predictor.py (/root/reference/Online_predictor.py:
set-up :66-235, loop :251-318, output :321-324) with the time loop resident on the GPU.

    PKG=synchronization-avoiding-algorithms_b200
    PYTHONPATH=$PKG:$PKG/compat python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        examples/online_predictor_driver.py --mesh beam.vtk --steps 400 --n-past 4 --n-future 3 --filter-size 5 [--model-dir DIR]

Every rank: warm-up of n_past*filter_size synchronised steps, then refill blocks in which the rank's LSTM
encoder-decoder predicts its shared-DOF displacements on the device and the FE steps run without any exchange.
--model-dir DIR names the directory the reference's Model_training.py writes into (`Distributed_save`): the weights are
read from DIR/Rank-<r>/nB-<nB>-nH-<hidden>-Lr-<lr>-filter=<filter>/model.pth (Online_predictor.py:139-141) and the scaling
constants are recomputed the way the reference does (:130-136) from Results/sol_on_shared/rank=<r>-shared_dof.hdf5 under
--out (training windows of the first --cut-off fraction, joint extrema).  Without --model-dir: seeded random weights.  --resync K adds a true exchange every K steps.  Ranks may share a GPU: the warm-up
exchange then travels through host memory with the process group (gloo / MPI).
Writes Results/Dynamics/Modeled_Local-rank-<r>.hdf5 (dataset 'Displacement', (3n, n_saved)).
"""
import argparse
import os

import numpy as np
import torch
import meshio
import h5py
from mpi4py import MPI
from mgmetis.parmetis import part_mesh_kway

from Tools.commons import *
from Tools.Distributed_tools import *
from Tools.DNN_prediction import *
from saa_b200 import maps, plan as splan, problem, sync_avoiding

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", required=True)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--save-every", type=int, default=1)
ap.add_argument("--out", default=".")
ap.add_argument("--n-past", type=int, default=20)
ap.add_argument("--n-future", type=int, default=20)
ap.add_argument("--filter-size", type=int, default=150)
ap.add_argument("--hidden", type=int, default=50)
ap.add_argument("--model-dir", default="")
ap.add_argument("--nB", type=int, default=10)
ap.add_argument("--learning-rate", type=float, default=5e-4)
ap.add_argument("--cut-off", type=float, default=0.5)
ap.add_argument("--resync", type=int, default=0)
args = ap.parse_args()

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
out_dir = os.path.join(args.out, "Results", "Dynamics")
os.makedirs(out_dir, exist_ok=True)

# mesh + partition (every rank reads the mesh; the partition call is the driver's, Online_predictor.py:94-114)
Mesh = meshio.read(args.mesh)
Cells, Facets, Points = Mesh.cells_dict['tetra'], Mesh.cells_dict['triangle'], Mesh.points
# element chunks handed to ParMETIS: the reference's elmdist (Online_predictor.py:83-86)
nEach = len(Cells) // size
nLeft = len(Cells) - nEach * size
bounds = np.append((nEach + 1) * np.arange(nLeft + 1), ((nEach + 1) * nLeft) + nEach * np.arange(1, size - nLeft + 1)).astype(np.int64)
mine = np.asarray(Cells[bounds[rank]:bounds[rank + 1]], dtype=np.int64)
_, epart = part_mesh_kway(size, 4 * np.arange(len(mine) + 1, dtype=np.int64), mine.reshape(-1))
recvbuf = np.empty(len(Cells), dtype='int') if rank == 0 else None
comm.Gatherv(epart, recvbuf, root=0)
epart = comm.bcast(recvbuf, root=0)

pb = problem.build_problem(Points, Cells, Facets, epart, size, ranks=[rank])        # maps, dt, K, lumped mass, load
q = pb["ranks"][rank]
Damp = 0.5
pl = splan.StepPlan(q["K"], q["F"], q["lM"], q["dirichlet"], pb["dt"], Damp, halo=q["halo"], rank=rank, size=size)
loc_dof_shared = q["loc_dof_shared"]                                                  # Online_predictor.py:129
input_size = loc_dof_shared.size


class HostExchangeStepper:
    """synchronised steps with the process group carrying the halo messages through host memory"""

    def step(self, n, mode):
        if mode == splan.MODE_SYNC and size > 1:
            for _ in range(n):
                pl.step_exchange(comm.exchange)
        else:
            pl.step(n, splan.MODE_LOCAL if mode == splan.MODE_SYNC else mode)


# surrogate (Online_predictor.py:139-141) and scaling constants (:130-136)
if args.model_dir:
    data_path = os.path.join(args.out, "Results", "sol_on_shared", f"rank={rank}-shared_dof.hdf5")
    X, Y = Dis_data_filtered_subset_coronary("cpu", input_size, args.filter_size, args.n_past, args.n_future, data_path, args.cut_off)
    _, _, scale_max, scale_min = Scale_to_zero_one(X, Y)                              # :130-136
    scale_max, scale_min = float(scale_max), float(scale_min)
    model_path = os.path.join(args.model_dir, f"Rank-{rank}", f"nB-{args.nB}-nH-{args.hidden}-Lr-{args.learning_rate}-filter={args.filter_size}",
                              "model.pth")                                            # :139-140
    model = call_model("cuda", args.filter_size, input_size, args.hidden, model_path)
else:
    torch.manual_seed(100 + rank)
    model = LSTM_encoder_decoder(max(input_size, 1), args.hidden, 2, True, 0.0, 0.0)
    scale_max, scale_min = 1e-3, -1e-2

run = sync_avoiding.SyncAvoidingRun([pl], HostExchangeStepper(), [loc_dof_shared], [model], [(scale_max, scale_min)], args.n_past,
                                    args.n_future, args.filter_size, resync_every=(args.resync or None))
n_saved = int(args.steps / args.save_every)
d1_save = np.zeros((pl.n_dof, n_saved))
i = 0
while i < args.steps:                                                                 # save_every-strided snapshots (:244, 267-270)
    nxt = min(args.steps, (i // args.save_every + 1) * args.save_every) if i % args.save_every else i + 1
    run.run(nxt)
    if (nxt - 1) % args.save_every == 0 and (nxt - 1) // args.save_every < n_saved:
        d1_save[:, (nxt - 1) // args.save_every] = pl.d0()
    i = nxt

hf = h5py.File(os.path.join(out_dir, f'Modeled_Local-rank-{rank}.hdf5'), 'w')
hf.create_dataset('Displacement', data=d1_save, compression='gzip')
hf.close()
comm.Barrier()
if rank == 0:
    print(f"done: {args.steps} steps ({args.n_past * args.filter_size} synchronised), size={size}")
