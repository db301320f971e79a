"""oracle/ref_harness.py — run the UNMODIFIED reference functions in this container.

TEST INFRASTRUCTURE ONLY (never imported by the product package).  It exists
only where `/root/reference` exists (the authoring container); the GPU box
consumes the golden vectors this harness produced (`tests/golden/*.npz`, made
by `oracle/gen_golden.py`).

What it does
------------
* puts `oracle/refshim/` stubs on `sys.path` for third-party modules the
  reference imports at module top but that are not installed here (mpi4py,
  meshio, h5py, matplotlib, mgmetis) — only when the real module is missing;
* imports the reference's `Tools.*` from `/root/reference` (read-only);
* restates the ~30 set-up lines of `/root/reference/Data_prepare.py` by
  CALLING the reference's own functions, for an explicit element->rank vector
  `epart` (the ParMETIS output of Data_prepare.py:94 is recorded nowhere in the
  reference, so the partition is an input here);
* runs the step loop of Data_prepare.py:223-240 with the reference's real
  `parallel_explicit_solver_dis_pre` (Tools/Dynamic_solver.py:9-34) and the
  reference's real `syn_cpus` (Tools/Distributed_tools.py:77-92); for P>1 the
  P ranks are executed in-process, one after the other, with a stand-in `comm`
  whose `gather`/`bcast` hand the real `syn_cpus` body the objects the MPI
  calls would have delivered.
"""
from __future__ import annotations

import importlib
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

REFERENCE_ROOT = os.environ.get("SAA_REFERENCE_ROOT", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
_ref = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Tools", "Dynamic_solver.py"))


def _install_shims():
    for name in ("mpi4py", "meshio", "h5py", "matplotlib", "mgmetis"):
        try:
            importlib.import_module(name)
        except Exception:
            if _SHIM_DIR not in sys.path:
                sys.path.append(_SHIM_DIR)  # appended: a real install always wins
            importlib.import_module(name)


class _Ref:
    """Namespace holding the imported reference modules."""


def load_reference():
    """Import /root/reference/Tools/* once and return a namespace of modules."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    r = _Ref()
    r.commons = importlib.import_module("Tools.commons")
    r.dist = importlib.import_module("Tools.Distributed_tools")
    r.mat = importlib.import_module("Tools.Mat_construction")
    r.dyn = importlib.import_module("Tools.Dynamic_solver")
    r.dnn_tools = importlib.import_module("Tools.DNN_tools")
    r.dnn_pred = importlib.import_module("Tools.DNN_prediction")
    assert r.dyn.__file__.startswith(REFERENCE_ROOT), r.dyn.__file__
    _ref = r
    return r


# --------------------------------------------------------------------------------------
# constants of the reference example (Data_prepare.py:35-50)
E, NU, RHO, FZ = 1e6, 0.3, 1, 0.5
DAMP, RAMP, P_ORDER, N_BASIS, FACET_NODE = 0.5, True, 1, 4, 3
GAMMA = .9


def read_mesh(path):
    """Data_prepare.py:57-61 (through the meshio stub when meshio is absent)."""
    load_reference()
    import meshio
    m = meshio.read(path)
    return m.points, m.cells_dict["tetra"], m.cells_dict["triangle"]


def ref_setup(points, cells, facets, epart, size, quiet=True):
    """Data_prepare.py:104-209 for every rank of a `size`-way partition `epart`.

    Returns a dict with global quantities and a list `ranks` of per-rank dicts, every
    array produced by the reference's own functions.
    """
    r = load_reference()
    c, d, mat = r.commons, r.dist, r.mat
    Points, Cells, Facets = points, cells, facets
    recvbuf = np.asarray(epart)
    elas = c.elasticity(E * NU / ((1 + NU) * (1 - 2 * NU)), E / (2 * (1 + NU)), RHO, FZ, RAMP)  # :47

    per = []
    for rank in range(size):
        ele, nod = d.rankwise_dist(rank, recvbuf, Points, Cells)                     # :104
        per.append(dict(rank=rank, Local_ele_list=ele, Local_nodal_list=nod))
    rank_nodal_num = [len(p["Local_nodal_list"]) for p in per]                       # :107
    rank_nodal_list = [p["Local_nodal_list"] for p in per]                           # :108
    for p in per:
        p["shared_nodes"] = d.find_shared_nodes(p["rank"], size, rank_nodal_num, rank_nodal_list)  # :112
    Global_shared = d.sort_shared([p["shared_nodes"] for p in per])                  # :123

    Dirichlet_node = []                                                              # :127-135
    for i in range(len(Facets)):
        if all(abs(Points[Facets[i][k]][0]) < 1e-9 for k in range(FACET_NODE)):
            for j in range(FACET_NODE):
                if Facets[i][j] not in Dirichlet_node:
                    Dirichlet_node.append(Facets[i][j])
    Dirichlet_global_dof = c.node_to_dof(3, [0, 1, 2], Dirichlet_node)              # :136

    dts = []
    for p in per:
        p["Local_Dirichlet"] = d.Dirichlet_rank_dist(Dirichlet_node, p["Local_nodal_list"])   # :144
        dts.append(GAMMA * c.Meshsize(Cells[p["Local_ele_list"], :], Points) / np.sqrt(E / RHO / (1 - NU ** 2)))  # :147
    dt = min(np.array(dts, dtype="float"))                                           # :151-154

    elas_steady = c.elasticity(E * NU / ((1 + NU) * (1 - 2 * NU)), E / (2 * (1 + NU)), RHO, FZ, False)  # :161
    d0 = np.zeros((len(Points) * 3, 1))                                              # :171
    sink = io.StringIO()
    with redirect_stdout(sink if quiet else sys.stdout):
        M_0, _, F_pre = mat.Global_Assembly_no_bc(P_ORDER, Cells, Points, elas_steady, 0)   # :175
    lumped_M = c.lumping_to_vec(M_0)                                                 # :176
    # ghost step (:179-189): with Ramp=True the load at t=0 is exactly zero, so a0 = 0 and
    # dn = d0 - dt*v0 + dt**2/2*a0 = 0 exactly; the dense solve is skipped here.
    dn = np.zeros((len(Points) * 3, 1))

    for p in per:
        local_dof = c.node_to_dof(3, [0, 1, 2], p["Local_nodal_list"])               # :200
        p["F_rankwise"] = F_pre[local_dof]                                           # :201
        p["l_M"] = lumped_M[local_dof]                                               # :202
        p["d_0"] = d0[local_dof]                                                     # :203
        p["d_n"] = dn[local_dof]                                                     # :204
        Local_cell = Cells[p["Local_ele_list"], :]                                   # :207
        with redirect_stdout(sink if quiet else sys.stdout):
            p["LocalK"] = mat.Local_assembly_for_stiffness(p["Local_nodal_list"], Local_cell, Points,
                                                           P_ORDER, N_BASIS, elas, p["rank"])  # :208
        p["loc_dof_shared"] = c.node_to_dof(3, [0, 1, 2], d.local_mat_node(p["shared_nodes"], p["Local_nodal_list"]))  # Online_predictor.py:129
    return dict(size=size, dt=dt, elas=elas, Points=Points, Cells=Cells, Facets=Facets, epart=recvbuf,
                Dirichlet_node=Dirichlet_node, Dirichlet_global_dof=Dirichlet_global_dof,
                Global_shared=Global_shared, lumped_M=lumped_M, F_pre=F_pre, ranks=per)


class _InProcessComm:
    """Hands the real syn_cpus (Distributed_tools.py:77-92) what MPI would have delivered.

    Within one time step the harness first evaluates every rank's `LocalK.dot(d0)` (the same
    scipy call Dynamic_solver.py:12 makes), then calls the real
    `parallel_explicit_solver_dis_pre` for rank 0, 1, ... in turn.  Inside it `syn_cpus`
    calls `comm.gather(f)`, `comm.gather(Local_nodes)` and `comm.bcast(f_global)`.
    """

    def __init__(self):
        self.forces = None
        self.node_lists = None
        self._ncall = 0
        self._f_global = None

    def new_step(self, forces, node_lists):
        self.forces, self.node_lists = forces, node_lists
        self._ncall = 0
        self._f_global = None

    def gather(self, obj, root=0):
        k = self._ncall % 2
        self._ncall += 1
        return self.forces if k == 0 else self.node_lists

    def bcast(self, obj, root=0):
        if obj is not None:            # rank 0 built f_global
            self._f_global = obj
        return self._f_global


def ref_run(setup, nsteps, save_steps=(), mode_model=False):
    """Data_prepare.py:215-240 with the reference's real step function, all ranks in-process.

    `save_steps`: 1-based step counts n after which the displacement d_n (the `d1` returned by
    the n-th call) is recorded.  Returns {n: [d1 of rank 0, d1 of rank 1, ...]} and final state.
    """
    r = load_reference()
    c, dyn, dist = r.commons, r.dyn, r.dist
    size = setup["size"]
    per = setup["ranks"]
    dt = setup["dt"]
    Points = setup["Points"]
    elas = setup["elas"]
    d_0 = [p["d_0"] for p in per]
    d_n = [p["d_n"] for p in per]
    tn = 0                                                                            # :215
    save_steps = set(int(s) for s in save_steps)
    out = {}
    comm = _InProcessComm()
    saved_comm = dist.comm
    dist.comm = comm
    try:
        for i in range(nsteps):                                                       # :223
            if size != 1 and not mode_model:
                forces = [per[q]["LocalK"].dot(d_0[q]) for q in range(size)]
                comm.new_step(forces, [per[q]["Local_nodal_list"] for q in range(size)])
            d1s = []
            for q in range(size):
                Time = c.Time_integration_displacement(tn, dt, d_0[q], d_n[q])        # :224
                d1 = dyn.parallel_explicit_solver_dis_pre(
                    per[q]["LocalK"], per[q]["F_rankwise"], Points, per[q]["Local_nodal_list"],
                    per[q]["Local_Dirichlet"], Time, elas, per[q]["l_M"], DAMP, size, q, MODEL=mode_model)  # :227
                d1s.append(d1)
            d_n = d_0                                                                 # :233
            d_0 = d1s                                                                 # :234
            tn = tn + dt                                                              # :235
            if (i + 1) in save_steps:
                out[i + 1] = [a.reshape(-1).copy() for a in d1s]
    finally:
        dist.comm = saved_comm
    return out, dict(d_0=d_0, d_n=d_n, tn=tn)
