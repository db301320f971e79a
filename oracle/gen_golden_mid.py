"""oracle/gen_golden_mid.py — mid-size multi-rank golden histories (TEST INFRASTRUCTURE; needs a GPU: run under gpurun).

    python oracle/gen_golden_mid.py [out_dir]      ->  out_dir/mid_np{2,4,8}.npz   (default out_dir: gpurun_out/golden)

Case: a stubby 3 x 1 x 1 structured beam, m = 14 (9 675 nodes, 29 025 DOF, 49 392 tets), METIS_PartMeshDual partitions into
2 / 4 / 8 parts — irregular interfaces, several boundary slices and several 256-row shared-row units per rank, nodes held
by three and more ranks (a k-way cut of the 25:1:1 cantilever degenerates into slabs with two-rank interfaces only).  The reference's own dense set-up cannot run at this size, so the matrices come from the product's device
assembly (deterministic: row-owned accumulation in ascending element order) and the HISTORIES COME FROM THE CPU ORACLE
(oracle/fem_oracle.c, pinned bit for bit to the unmodified reference) stepping exactly those matrices with the syn_cpus
semantics of Distributed_tools.py:77-92.  The fixtures let bench.py's N > 1 parity pre-check (which must not touch oracle/)
and the multi-GPU tests compare the fused peer-memory step with the oracle on a mesh that has more than one boundary slice.
tests/test_gpu_parity.py::test_mid_fixture_is_reproducible regenerates them on the test box and compares bitwise.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

M, LENGTH, STEPS = 14, 3, (1, 60, 200)


def generate(P, device_index=0):
    import saa_b200  # noqa: F401
    from saa_b200 import device_setup, mesh, partition
    import fem_oracle
    pts, cells, fac = mesh.structured_beam(M, length=LENGTH)
    epart = partition.metis_part_mesh(cells, len(pts), P)
    plans, infos = device_setup.build_mesh_in_process(pts, cells, fac, epart, P, device_index=device_index, keep_csr=True)
    ranks = []
    for q in range(P):
        K = infos[q]["K"].to_scipy()
        ranks.append(dict(K_indptr=K.indptr, K_indices=K.indices, K_data=K.data, F=infos[q]["F"].cpu().numpy(), lM=infos[q]["lM"].cpu().numpy(),
                          dirichlet=infos[q]["dirichlet"], nodes=infos[q]["local_nodes"].cpu().numpy()))
    o = fem_oracle.OracleProblem(len(pts), ranks, infos[0]["dt"], 0.5)
    out = dict(m=M, length=LENGTH, size=P, epart=epart.astype(np.int8), steps=np.asarray(STEPS), dt=infos[0]["dt"], alpha=0.5)
    done = 0
    for s in STEPS:
        o.run(s - done)
        done = s
        for q in range(P):
            out[f"hist_{s}_r{q}"] = o.d0(q)
    halos = [i["halo"] for i in infos]
    multi = np.bincount(np.concatenate([np.asarray(r["nodes"]) for r in ranks]))
    out["stats"] = np.asarray([len(pts), int((multi >= 2).sum()), int((multi >= 3).sum()), max(len(h["neighbours"]) for h in halos),
                               max(len(h["shared_pos"]) for h in halos)])
    for p in plans:
        p.close()
    return out


if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for P in (2, 4, 8):
        g = generate(P)
        np.savez_compressed(os.path.join(out_dir, f"mid_np{P}.npz"), **g)
        print(f"P={P}: nodes, shared(>=2), shared(>=3), max neighbours, max shared per rank = {g['stats'].tolist()}")
