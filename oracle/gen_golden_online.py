"""oracle/gen_golden_online.py — golden trajectory of the synchronization-avoiding loop from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY; run in the authoring container:  python oracle/gen_golden_online.py
The loop of Online_predictor.py:251-318 is restated by CALLING the reference's own functions
(parallel_explicit_solver_dis_pre with MODEL=False / True, syn_cpus through the in-process communicator of
ref_harness, encoder_decoder_predictor, LSTM_encoder_decoder) for beam_coarse with the fixture's 2-way partition;
the surrogates have seeded random weights (the reference ships no trained model).  -> tests/golden/online_beam_coarse_np2.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

r = H.load_reference()
c, dyn, dist = r.commons, r.dyn, r.dist
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

g = np.load(os.path.join(GOLDEN, "beam_coarse_P2.npz"))
pts, cells, fac, epart = g["points"], g["cells"], g["facets"], g["epart"]
P = 2
s = H.ref_setup(pts, cells, fac, epart, P)
per, dt, elas, Points = s["ranks"], s["dt"], s["elas"], s["Points"]
n_p, n_f, n_s, hidden, test_num = 4, 3, 5, 8, 64
i_cri = n_p * n_s - 1                                                           # Online_predictor.py:63
scales = [(1e-3, -1e-2), (1e-3, -1e-2)]
models = []
for q in range(P):
    torch.manual_seed(100 + q)
    models.append(r.dnn_tools.LSTM_encoder_decoder(len(per[q]["loc_dof_shared"]), hidden, 2, True, 0.0, 0.0))

d_0 = [p["d_0"] for p in per]
d_n = [p["d_n"] for p in per]
d_sol = [np.zeros((test_num, len(p["loc_dof_shared"]))) for p in per]             # :244
tn = 0
comm = H._InProcessComm()
saved = dist.comm
dist.comm = comm
try:
    i = 0
    counter2 = 0
    while i < test_num:                                                         # :251
        if i <= i_cri:                                                          # :253-275
            forces = [per[q]["LocalK"].dot(d_0[q]) for q in range(P)]
            comm.new_step(forces, [per[q]["Local_nodal_list"] for q in range(P)])
            d1s = []
            for q in range(P):
                T = c.Time_integration_displacement(tn, dt, d_0[q], d_n[q])
                d1 = dyn.parallel_explicit_solver_dis_pre(per[q]["LocalK"], per[q]["F_rankwise"], Points, per[q]["Local_nodal_list"],
                                                          per[q]["Local_Dirichlet"], T, elas, per[q]["l_M"], H.DAMP, P, q, MODEL=False)
                d_sol[q][i, :] = d1[per[q]["loc_dof_shared"], 0]                # :260
                d1s.append(d1)
            d_n, d_0, tn, i = d_0, d1s, tn + dt, i + 1
        else:
            d_shared = [r.dnn_pred.encoder_decoder_predictor("cpu", i, models[q], n_p, n_f, n_s, len(per[q]["loc_dof_shared"]),
                                                             d_sol[q], scales[q][0], scales[q][1]) for q in range(P)]   # :280
            for k in range(i, i + n_f * n_s):                                   # :284
                if k >= test_num:
                    break
                d1s = []
                for q in range(P):
                    T = c.Time_integration_displacement(tn, dt, d_0[q], d_n[q])
                    d1 = dyn.parallel_explicit_solver_dis_pre(per[q]["LocalK"], per[q]["F_rankwise"], Points, per[q]["Local_nodal_list"],
                                                              per[q]["Local_Dirichlet"], T, elas, per[q]["l_M"], H.DAMP, P, q, MODEL=True)
                    n_in = len(per[q]["loc_dof_shared"])
                    d1[per[q]["loc_dof_shared"]] = d_shared[q][k - i_cri - 1 - n_f * n_s * counter2, :].reshape((n_in, 1))   # :298
                    d_sol[q][i, :] = d1[per[q]["loc_dof_shared"], 0]            # :301
                    d1s.append(d1)
                d_n, d_0, tn, i = d_0, d1s, tn + dt, i + 1
            counter2 += 1                                                       # :318
finally:
    dist.comm = saved

out = dict(n_p=n_p, n_f=n_f, n_s=n_s, hidden=hidden, test_num=test_num, scale_max=scales[0][0], scale_min=scales[0][1])
for q in range(P):
    out[f"final_r{q}"] = d_0[q].reshape(-1)
    out[f"d_sol_r{q}"] = d_sol[q]
    for k, v in models[q].state_dict().items():
        out[f"sd{q}__" + k] = v.numpy()
np.savez_compressed(os.path.join(GOLDEN, "online_beam_coarse_np2.npz"), **out)
print("online golden:", test_num, "steps,", [float(np.abs(d_0[q]).max()) for q in range(P)], os.path.getsize(os.path.join(GOLDEN, "online_beam_coarse_np2.npz")) // 1024, "KiB")
