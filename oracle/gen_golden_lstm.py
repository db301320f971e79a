"""oracle/gen_golden_lstm.py — golden vectors of the LSTM refill predictor from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY; run in the authoring container:  python oracle/gen_golden_lstm.py
Builds the reference's LSTM_encoder_decoder (Tools/DNN_tools.py:85-98) with seeded random weights, a seeded
synthetic shared-DOF history, and records the output of the reference's own encoder_decoder_predictor
(Tools/DNN_prediction.py:38-55; CPU, batch-1 recursion) -> tests/golden/lstm_*.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

r = H.load_reference()
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

for name, inp, hid, n_p, n_f, n_s, n in (("lstm_small", 24, 16, 20, 20, 7, 300), ("lstm_wide", 60, 50, 20, 20, 3, 100),
                                        ("lstm_short", 9, 8, 5, 4, 2, 40)):
    torch.manual_seed(1234)
    rng = np.random.default_rng(99)
    model = r.dnn_tools.LSTM_encoder_decoder(inp, hid, 2, True, 0.0, 0.0)
    t = np.arange(n + n_f * n_s + 5)[:, None] * 0.01
    d_sol = 1e-3 * np.sin(3.0 * t + rng.uniform(0, 6.28, (1, inp))) * rng.uniform(0.2, 1.0, (1, inp)) - 2e-3 * t
    scale_max, scale_min = float(d_sol[:n].max() * 1.1), float(d_sol[:n].min() * 1.1)
    NF = r.dnn_pred.encoder_decoder_predictor("cpu", n, model, n_p, n_f, n_s, inp, d_sol, scale_max, scale_min)
    out = dict(input_size=inp, hidden_size=hid, n_p=n_p, n_f=n_f, n_s=n_s, n=n, d_sol=d_sol, scale_max=scale_max,
               scale_min=scale_min, NF=NF)
    for k, v in model.state_dict().items():
        out["sd__" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, NF.shape, float(np.abs(NF).max()), os.path.getsize(os.path.join(GOLDEN, name + ".npz")) // 1024, "KiB")
