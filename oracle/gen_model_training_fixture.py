"""oracle/gen_model_training_fixture.py — how tests/golden/model_training_ref_rank{0,1}.pth were made (authoring
container only: needs /root/reference; TEST INFRASTRUCTURE).

    python oracle/gen_model_training_fixture.py /tmp/mt            # writes the inputs, prints the two commands to run

1. Inputs of the reference's Model_training.py for beam_coarse, 2 partitions (the partition of tests/golden/
   beam_coarse_P2.npz): Results/Shared_Data/Rank=r_shared.csv, Results/Rankwised_Data/Rank=r_local_nodes.csv and
   Results/sol_on_shared/rank=r-shared_dof.hdf5 (30 000 fully synchronised steps of the shared DOFs, produced by the CPU
   oracle, which is pinned bit for bit to the reference; written through the compat h5py stand-in as .hdf5.npz).
2. The reference's UNMODIFIED script, two ranks:
       cd /tmp/mt && python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
           $REPO/synchronization-avoiding-algorithms_b200/run_driver.py /root/reference/Model_training.py
   (3450 epochs with its own hyper-parameters, ~9 min on 8 CPU cores; log: profiles/r2/model_training_unchanged_2ranks_cpu.log)
3. cp Distributed_save/Rank-r/nB-10-nH-50-Lr-0.0005-filter=150/model.pth  ->  tests/golden/model_training_ref_rank{r}.pth
The weights are not reproducible bit for bit (Model_training.py:101 shuffles unseeded); the fixtures are the outcome of one run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

if __name__ == "__main__":
    from util import load_golden, make_oracle
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/mt"
    g = load_golden("beam_coarse_P2")
    o = make_oracle(g)
    T = 30000
    dofs = [r["loc_dof_shared"] for r in g["ranks"]]
    H = [np.zeros((d.size, T)) for d in dofs]
    for i in range(T):
        o.run(1)
        for q in range(2):
            H[q][:, i] = o.d0(q)[dofs[q]]
    for d in ("Results/Shared_Data", "Results/Rankwised_Data", "Results/sol_on_shared"):
        os.makedirs(os.path.join(out, d), exist_ok=True)
    for q, r in enumerate(g["ranks"]):
        np.savetxt(os.path.join(out, f"Results/Shared_Data/Rank={q}_shared.csv"), r["shared"], delimiter=",", fmt="%d")
        np.savetxt(os.path.join(out, f"Results/Rankwised_Data/Rank={q}_local_nodes.csv"), r["nodes"], delimiter=",", fmt="%d")
        np.savez_compressed(os.path.join(out, f"Results/sol_on_shared/rank={q}-shared_dof.hdf5.npz"), Displacement=H[q])
    print(__doc__.split("2. The reference")[1].split("3. cp")[0])
