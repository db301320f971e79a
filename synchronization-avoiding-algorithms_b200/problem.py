"""Set-up of a partitioned explicit-dynamics problem without dense matrices.

Restates the set-up section of /root/reference/Data_prepare.py:104-209 (maps, Dirichlet DOFs, dt, lumped
mass, load vector, per-rank stiffness) on top of the vectorised, sequence-exact helpers of `maps`,
`mesh` and `assembly`, and turns each rank's data into a device `StepPlan`.
"""
from __future__ import annotations

import numpy as np

from . import assembly, maps, mesh
from .plan import PlanGroup, StepPlan

# constants of the reference example (Data_prepare.py:35-48)
E_DEFAULT, NU_DEFAULT, RHO_DEFAULT, FZ_DEFAULT = 1e6, 0.3, 1, 0.5
DAMP_DEFAULT, GAMMA_DEFAULT = 0.5, .9


def lame(E, nu):
    """(lambda, mu) exactly as Data_prepare.py:47 spells them."""
    return E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


def build_problem(points, cells, facets, epart, size, E=E_DEFAULT, nu=NU_DEFAULT, rho=RHO_DEFAULT, fz=FZ_DEFAULT,
                  gamma=GAMMA_DEFAULT, ranks=None, exact_rowsum=None):
    """Host-side problem description for a `size`-way element partition `epart`.

    ranks: which ranks to assemble (default all) — a one-process-per-GPU driver passes [its rank].
    Returns dict(dt, size, Dirichlet_node, Global_shared, node_lists, ranks={r: {...}}).
    """
    cells = np.asarray(cells, dtype=np.int64)
    lmd, mu = lame(E, nu)
    D = mesh.dirichlet_nodes(points, facets)                                       # Data_prepare.py:127-135
    per, gshared = maps.partition_maps(epart, cells, size, D)                      # :104-144
    node_lists = [p["Local_nodal_list"] for p in per]
    dts = [mesh.stable_dt(cells[p["Local_ele_list"]], points, E, nu, rho, gamma) for p in per]   # :147
    dt = min(dts)                                                                  # :151-154
    lM, F = assembly.lumped_mass_and_load(points, cells, rho, fz, exact_rowsum=exact_rowsum)     # :175-176
    out = {}
    for r in (range(size) if ranks is None else ranks):
        p = per[r]
        dof = maps.node_to_dof(3, [0, 1, 2], p["Local_nodal_list"])                # :200
        q = dict(rank=r, nodes=p["Local_nodal_list"], ele=p["Local_ele_list"], shared=p["shared_nodes"],
                 loc_dof_shared=p["loc_dof_shared"], dirichlet=p["Local_Dirichlet"],
                 F=F[dof], lM=lM[dof],                                             # :201-202
                 K=assembly.local_stiffness_csr(p["Local_nodal_list"], cells[p["Local_ele_list"]], points, lmd, mu))  # :207-209
        q["halo"] = maps.halo_plan(r, size, node_lists) if size > 1 else None
        out[r] = q
    return dict(dt=dt, size=size, Dirichlet_node=D, Global_shared=gshared, node_lists=node_lists, ranks=out,
                n_global_dof=3 * len(points))


def make_plan(rank_data, dt, alpha, size, device=0):
    """StepPlan of one rank from the dict build_problem (or a golden fixture) holds for it."""
    q = rank_data
    K = q["K"]
    return StepPlan(K, q["F"], q["lM"], q["dirichlet"], dt, alpha, device=device, halo=q.get("halo"),
                    rank=q["rank"], size=size)


def make_group(problem, alpha=DAMP_DEFAULT, device=0):
    """All ranks of `problem` on one GPU: (plans, PlanGroup or None for size 1)."""
    plans = [make_plan(problem["ranks"][r], problem["dt"], alpha, problem["size"], device) for r in range(problem["size"])]
    return plans, (PlanGroup(plans) if problem["size"] > 1 else None)
