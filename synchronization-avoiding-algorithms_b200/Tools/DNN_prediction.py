"""Online prediction of the shared-node displacements (/root/reference/Tools/DNN_prediction.py)."""
import numpy as np
import torch

from Tools.DNN_tools import *  # noqa: F401,F403
from Tools.DNN_tools import LSTM_encoder_decoder, model_predict_batch, scale_forward, scale_it_back


def call_model(device, filter_size, input_size, hidden_size, model_path):
    """Fixed architecture of the reference (2 encoder layers, bidirectional, no dropout) with learned weights
    loaded from `model_path` (DNN_prediction.py:18-34)."""
    model = LSTM_encoder_decoder(input_size, hidden_size, 2, True, 0.0, 0.0)
    model.load_state_dict(torch.load(model_path, map_location=device))
    return model.to(device)


def comb_indices(n, n_p, n_f, n_s):
    """History / future row indices of the n_s interleaved combs of one refill block, with the reference's own
    arange expressions (DNN_prediction.py:44-45): (n_s, len) int arrays."""
    past = np.stack([np.arange(i + n - n_p * n_s, i + n - 1, n_s) for i in range(n_s)])
    fut = np.stack([np.arange(i + n, n + i + n_f * n_s - 1, n_s) for i in range(n_s)])
    return past, fut


def predict_block(model, hist_rows, n_p, n_f, n_s, scale_max, scale_min):
    """hist_rows: float64 tensor (n_p*n_s, input) = rows [n - n_p*n_s, n) of the shared-DOF history, on the
    model's device.  Returns the float64 (n_s*n_f, input) refill table, row t = prediction for step n + t.
    All n_s combs are one batch; scaling as DNN_tools.py:272-279; model I/O in float32 (DNN_prediction.py:49)."""
    inp = hist_rows.shape[1]
    X = scale_forward(hist_rows.view(n_p, n_s, inp).permute(1, 0, 2), scale_max, scale_min).float()   # comb i: rows i + j*n_s
    Y = model_predict_batch(model, X.contiguous(), n_f)                                                  # (n_s, n_f, input) fp32
    Y = scale_it_back(Y, scale_max, scale_min)
    return Y.permute(1, 0, 2).reshape(n_f * n_s, inp).double().contiguous()                              # row j*n_s + i


def encoder_decoder_predictor(device, n, model, n_p, n_f, n_s, input_size, d_sol, scale_max, scale_min):
    """Reference signature (DNN_prediction.py:38-55): d_sol is the (steps, input) float64 history array; returns
    the (n_s*n_f, input) float64 table as a numpy array."""
    past, fut = comb_indices(n, n_p, n_f, n_s)
    NF = np.zeros((n_s * n_f, input_size))
    X = scale_forward(np.asarray(d_sol)[past, :], scale_max, scale_min)                # (n_s, n_p', input)
    X = torch.from_numpy(X).float().to(device)
    Y = scale_it_back(model_predict_batch(model, X, n_f), scale_max, scale_min)        # (n_s, n_f, input)
    Y = Y.cpu().numpy()
    for i in range(n_s):
        NF[fut[i] - n, :] = Y[i, :fut.shape[1], :]
    return NF
