"""oracle/gen_golden.py — produce tests/golden/*.npz from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs /root/reference):

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py beam_coarse_P2

Every array in a fixture is the output of the reference's own functions driven by
oracle/ref_harness.py (set-up lines of Data_prepare.py:104-209, step loop :223-240 with the real
parallel_explicit_solver_dis_pre / syn_cpus).  Inputs that the reference obtains from tools not
available offline are recorded in the fixture as inputs: the mesh arrays (meshio / gmsh) and the
element->rank vector `epart` (ParMETIS).  The structured meshes and the METIS/slab partitions are
produced with the product's mesh/partition helpers — they are inputs, not results.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as H  # noqa: E402
import saa_b200  # noqa: E402,F401
from saa_b200 import mesh, partition  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (mesh, P, partitioner, sync steps to record, no-sync (MODEL=True) steps to record)
    "beam_coarse_P1": ("beam_coarse", 1, "metis", [1, 2, 10, 100, 1000, 2000, 10000], []),
    "beam_coarse_P2": ("beam_coarse", 2, "metis", [1, 2, 10, 100, 1000, 2000], [1, 2, 10, 100]),
    "beam_coarse_P3": ("beam_coarse", 3, "metis", [1, 10, 100, 1000], []),
    "beam_coarse_P4": ("beam_coarse", 4, "metis", [1, 10, 100, 1000], [10]),
    "beam_coarse_P8": ("beam_coarse", 8, "metis", [1, 10, 100, 500], []),
    "struct_m2_P1": ("struct2", 1, "metis", [1, 100, 1000], []),
    "struct_m2_P2": ("struct2", 2, "metis", [1, 100, 1000], [100]),
    "struct_m2_P4": ("struct2", 4, "slab", [1, 100, 1000], []),
    "struct_m3_P1": ("struct3", 1, "metis", [1, 100, 500], []),
    "struct_m3_P3": ("struct3", 3, "metis", [1, 100, 500], []),
    # the larger oracle case of SURVEY.md section 8c (2525 nodes, 7575 DOF; the reference's dense set-up takes ~2 minutes)
    "struct_m4_P2": ("struct4", 2, "slab", [1, 100, 1000, 3000], [50]),
    # 6517 nodes, 19 551 DOF, 32 400 tets — near the practical ceiling of the reference's dense (3N)^2 set-up
    "struct_m6_P3": ("struct6", 3, "metis", [1, 100, 1000], []),
}


def get_mesh(name):
    if name == "beam_coarse":
        return H.read_mesh(os.path.join(H.REFERENCE_ROOT, "Mesh_info", "beam_coarse.vtk"))
    if name.startswith("struct"):
        return mesh.structured_beam(int(name[len("struct"):]))
    raise KeyError(name)


def make_case(name):
    mesh_name, P, part, steps, nosync_steps = CASES[name]
    pts, cells, fac = get_mesh(mesh_name)
    if part == "metis":
        epart = partition.metis_part_mesh(cells, len(pts), P)
    else:
        epart = partition.slab_partition(pts, cells, P)
    t0 = time.time()
    s = H.ref_setup(pts, cells, fac, epart, P)
    t1 = time.time()
    hist, _ = H.ref_run(s, max(steps), steps)
    out = dict(points=pts, cells=np.asarray(cells, dtype=np.int64), facets=np.asarray(fac, dtype=np.int64),
               epart=np.asarray(epart, dtype=np.int64), size=np.int64(P), dt=np.float64(s["dt"]),
               alpha=np.float64(H.DAMP), E=np.float64(H.E), nu=np.float64(H.NU), rho=np.float64(H.RHO),
               fz=np.float64(H.FZ), gamma=np.float64(H.GAMMA),
               Dirichlet_node=np.asarray(s["Dirichlet_node"], dtype=np.int64),
               Global_shared=np.asarray(s["Global_shared"], dtype=np.int64),
               lumped_M=s["lumped_M"], F_pre=s["F_pre"],
               steps=np.asarray(steps, dtype=np.int64), nosync_steps=np.asarray(nosync_steps, dtype=np.int64))
    for q, p in enumerate(s["ranks"]):
        K = p["LocalK"]
        assert K.has_sorted_indices
        out[f"r{q}_ele"] = np.asarray(p["Local_ele_list"], dtype=np.int64)
        out[f"r{q}_nodes"] = np.asarray(p["Local_nodal_list"], dtype=np.int64)
        out[f"r{q}_shared"] = np.asarray(p["shared_nodes"], dtype=np.int64)
        out[f"r{q}_dirichlet"] = np.asarray(p["Local_Dirichlet"], dtype=np.int64)
        out[f"r{q}_loc_dof_shared"] = np.asarray(p["loc_dof_shared"], dtype=np.int64)
        out[f"r{q}_K_indptr"] = K.indptr
        out[f"r{q}_K_indices"] = K.indices
        out[f"r{q}_K_data"] = K.data
        out[f"r{q}_lM"] = p["l_M"]
        out[f"r{q}_F"] = p["F_rankwise"]
        for n in steps:
            out[f"hist_{n}_r{q}"] = hist[n][q]
    if nosync_steps:
        # MODEL=True branch of Dynamic_solver.py:22 (no syn_cpus at all), plain un-synchronised steps
        h2, _ = H.ref_run(s, max(nosync_steps), nosync_steps, mode_model=True)
        for n in nosync_steps:
            for q in range(P):
                out[f"nosync_{n}_r{q}"] = h2[n][q]
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: setup {t1 - t0:.1f}s, total {time.time() - t0:.1f}s, {os.path.getsize(path) / 1024:.0f} KiB, "
          f"dt={s['dt']!r}, nodes/rank={[len(p['Local_nodal_list']) for p in s['ranks']]}, "
          f"shared/rank={[len(p['shared_nodes']) for p in s['ranks']]}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        make_case(nm)
