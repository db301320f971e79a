#!/usr/bin/env python
"""Distributed explicit-dynamics run through the drop-in `Tools` package — the flow of the reference's
Data_prepare.py (/root/reference/Data_prepare.py:56-246) with the mesh, step count and output directory as
arguments.  Every solver call below is a name the reference driver imports from `Tools.*`.

    PKG=synchronization-avoiding-algorithms_b200
    PYTHONPATH=$PKG:$PKG/compat python examples/data_prepare_driver.py --mesh beam.vtk --steps 2000
    PYTHONPATH=$PKG:$PKG/compat python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        examples/data_prepare_driver.py --mesh beam.vtk --steps 2000

Writes Results/Rankwised_Data, Results/Shared_Data, Results/Rankwised_Element (CSV maps) and
Results/Dynamics/Local-rank-<r>.hdf5 (dataset 'Displacement', (3n, n_saved)) under --out.
"""
import argparse
import os
from math import floor

from Tools.commons import *
from Tools.Distributed_tools import *
from Tools.Steady_solvers import *
from Tools.Dynamic_solver import *
from mgmetis.parmetis import part_mesh_kway
import numpy as np
import meshio
from mpi4py import MPI
import h5py

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", required=True)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--save-every", type=int, default=1)
ap.add_argument("--out", default=".")
ap.add_argument("--steady", action="store_true", help="also solve and write the steady solution")
args = ap.parse_args()

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
dirs = {k: os.path.join(args.out, "Results", k) for k in
        ("Rankwised_Data", "Shared_Data", "Static", "Dynamics", "Rankwised_Element")}
for d in dirs.values():
    os.makedirs(d, exist_ok=True)

E, nu, rho, fz = 1e6, 0.3, 1, 0.5                       # material / load of the reference example
Damp, Ramp, p, n_basis, facet_node = 0.5, True, 1, 4, 3
gamma = .9
elas = elasticity(E*nu/((1+nu)*(1-2*nu)), E/(2*(1+nu)), rho, fz, Ramp)

# mesh on rank 0, broadcast; contiguous element chunks for the partitioner
Cells = Facets = Points = elmdist = None
if rank == 0:
    Mesh = meshio.read(args.mesh)
    Cells, Facets, Points = Mesh.cells_dict['tetra'], Mesh.cells_dict['triangle'], Mesh.points
    nELE = len(Cells)
    nEach = floor(nELE / size)
    nLeft = nELE - nEach * size
    elmdist = np.append((nEach + 1) * np.arange(nLeft + 1),
                        (nEach + 1) * nLeft + nEach * np.arange(1, size - nLeft + 1)).astype(np.int64)
Cells, Facets, Points, elmdist = (comm.bcast(x, root=0) for x in (Cells, Facets, Points, elmdist))
mine = Cells[elmdist[rank]:elmdist[rank + 1]]
eptr = 4 * np.arange(len(mine) + 1, dtype=np.int64)
eind = np.asarray(mine, dtype=np.int64).reshape(-1)
_, epart = part_mesh_kway(size, eptr, eind)
recvbuf = np.empty(len(Cells), dtype='int') if rank == 0 else None
comm.Gatherv(epart, recvbuf, root=0)
recvbuf = comm.bcast(recvbuf, root=0)

# maps
Local_ele_list, Local_nodal_list = rankwise_dist(rank, recvbuf, Points, Cells)
rank_nodal_num = comm.bcast(comm.gather(len(Local_nodal_list), root=0), root=0)
rank_nodal_list = comm.bcast(comm.gather(Local_nodal_list, root=0), root=0)
shared_nodes = find_shared_nodes(rank, size, rank_nodal_num, rank_nodal_list)
np.savetxt(os.path.join(dirs["Shared_Data"], f'Rank={rank}_shared.csv'), shared_nodes, delimiter=',', fmt='%d')
np.savetxt(os.path.join(dirs["Rankwised_Data"], f'Rank={rank}_local_nodes.csv'), rank_nodal_list[rank], delimiter=',', fmt='%d')
np.savetxt(os.path.join(dirs["Rankwised_Element"], f'Rank={rank}_elements.csv'), Local_ele_list, delimiter=',', fmt='%d')
G_shared_nodes = comm.gather(shared_nodes, root=0)
if rank == 0:
    np.savetxt(os.path.join(dirs["Shared_Data"], 'Global_shared.csv'), sort_shared(G_shared_nodes), delimiter=',', fmt='%d')

# clamp x = 0
Dirichlet_node = None
if rank == 0:
    from saa_b200 import mesh as _mesh
    Dirichlet_node = _mesh.dirichlet_nodes(Points, Facets)
    Dirichlet_global_dof = node_to_dof(3, [0, 1, 2], Dirichlet_node)
Dirichlet_node = comm.bcast(Dirichlet_node, root=0)
Local_Dirichlet = Dirichlet_rank_dist(Dirichlet_node, Local_nodal_list)

# time step: CFL on the local elements, minimum over ranks
dt = gamma * Meshsize(Cells[Local_ele_list, :], Points) / np.sqrt(E/rho/(1-nu**2))
recvbuf2 = np.empty(size, dtype='float') if rank == 0 else None
comm.Gather(dt, recvbuf2, root=0)
dt = min(comm.bcast(recvbuf2, root=0))

# rank 0: lumped mass, load vector, initial data (ramped load => ghost step is exactly zero)
lumped_M = d0 = dn = F_pre = None
if rank == 0:
    elas_steady = elasticity(E*nu/((1+nu)*(1-2*nu)), E/(2*(1+nu)), rho, fz, False)
    if args.steady:
        d_steady = Steady_Elasticity_solver(p, Cells, Points, Dirichlet_global_dof, elas_steady)
        meshio.write_points_cells(os.path.join(dirs["Static"], 'steady_distributed.vtk'), Points, Mesh.cells,
                                  {'displacement-x': d_steady[0::3], 'displacement-y': d_steady[1::3],
                                   'displacement-z': d_steady[2::3]})
    d0 = np.zeros((len(Points)*3, 1))
    M_0, _, F_pre = Global_Assembly_no_bc(p, Cells, Points, elas_steady, 0)
    lumped_M = lumping_to_vec(M_0)
    dn = np.zeros((len(Points)*3, 1))
lumped_M, d0, dn, F_pre = (comm.bcast(x, root=0) for x in (lumped_M, d0, dn, F_pre))

local_dof = node_to_dof(3, [0, 1, 2], Local_nodal_list)
F_rankwise, l_M, d_0, d_n = F_pre[local_dof], lumped_M[local_dof], d0[local_dof], dn[local_dof]
LocalK = Local_assembly_for_stiffness(Local_nodal_list, Cells[Local_ele_list, :], Points, p, n_basis, elas, rank)

# time integration
tn = 0
d1_save = np.zeros((len(Local_nodal_list)*3, int(args.steps/args.save_every)))
counter = 0
for i in range(args.steps):
    Time = Time_integration_displacement(tn, dt, d_0, d_n)
    d1 = parallel_explicit_solver_dis_pre(LocalK, F_rankwise, Points, Local_nodal_list, Local_Dirichlet,
                                          Time, elas, l_M, Damp, size, rank, MODEL=False)
    d_n, d_0, tn = d_0, d1, tn + dt
    if i % args.save_every == 0 and counter < d1_save.shape[1]:
        d1_save[:, counter] = d1.reshape(len(d1))
        counter += 1

hf = h5py.File(os.path.join(dirs["Dynamics"], f'Local-rank-{rank}.hdf5'), 'w')
hf.create_dataset('Displacement', data=d1_save, compression='gzip')
hf.close()
if rank == 0:
    print(f"done: {args.steps} steps, dt={dt!r}, size={size}")
