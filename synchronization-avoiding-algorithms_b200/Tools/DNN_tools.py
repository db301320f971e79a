"""LSTM encoder-decoder surrogate of the shared-node displacements — names and state_dict layout of
/root/reference/Tools/DNN_tools.py, so that a `model.pth` trained by the reference's Model_training.py loads
unchanged (`encoder.lstm_encoder.*`, `decoder.lstm_decoder.*`, `decoder.fc.*`).

Per the north star the surrogate stays plain PyTorch (cuDNN LSTM on the GPU).  What is added for the B200 path
is `model_predict_batch`: the reference predicts the n_s interleaved "combs" of a refill block one by one with
batch size 1 (DNN_prediction.py:43-54); they are independent, so they run here as ONE batch on the device.
"""
import random

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import Dataset

try:                                    # the reference's scripts reach `plt` through `from Tools.DNN_tools import *`
    from matplotlib import pyplot as plt
except Exception:                       # plotting is optional
    plt = None


class LSTM_Encoder(nn.Module):
    """Stacked (optionally bidirectional) LSTM; forward returns the LAST layer's final (h, c), the two directions
    concatenated, shaped (1, N, D*hidden) ready to seed the decoder (DNN_tools.py:16-59)."""

    def __init__(self, input_size, hidden_size, num_layers, Bi_dir, dp):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers, self.Bi_dir, self.dp = input_size, hidden_size, num_layers, Bi_dir, dp
        self.D = 2 if Bi_dir else 1
        self.lstm_encoder = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, batch_first=True,
                                    dropout=dp, bidirectional=bool(Bi_dir))

    def forward(self, x):
        _, (hn, cn) = self.lstm_encoder(x)                      # (D*layers, N, H)
        N = x.shape[0]
        hn = hn.view(self.num_layers, self.D, N, self.hidden_size)[-1]
        cn = cn.view(self.num_layers, self.D, N, self.hidden_size)[-1]
        if self.D == 1:
            return hn, cn
        return torch.cat((hn[0], hn[1]), 1).unsqueeze(0), torch.cat((cn[0], cn[1]), 1).unsqueeze(0)


class LSTM_Decoder(nn.Module):
    """One LSTM step from (x, h, c) followed by a dense layer back to the input width (DNN_tools.py:63-80)."""

    def __init__(self, input_size, hidden_size, Bi_dir, dp):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size * 2 if Bi_dir else hidden_size
        self.dp = dp
        self.lstm_decoder = nn.LSTM(input_size=input_size, hidden_size=self.hidden_size, num_layers=1, batch_first=True,
                                    bidirectional=False)
        self.fc = nn.Linear(self.hidden_size, input_size)
        self.dropout = nn.Dropout(dp)

    def forward(self, x, encoded_hn, encoded_cn):
        out, (h, c) = self.lstm_decoder(x.unsqueeze(1), (encoded_hn, encoded_cn))
        return self.fc(self.dropout(out.squeeze(1))), h, c


class LSTM_encoder_decoder(nn.Module):
    """Container with the sub-module names the reference's checkpoints use (DNN_tools.py:85-98)."""

    def __init__(self, input_size, hidden_size, num_layers_encoder, Bi_dir_encoder, dp_encoder, dp_decoder):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.num_layers_encoder, self.Bi_dir_encoder = num_layers_encoder, Bi_dir_encoder
        self.dp_encoder, self.dp_decoder = dp_encoder, dp_decoder
        self.encoder = LSTM_Encoder(input_size, hidden_size, num_layers_encoder, Bi_dir_encoder, dp_encoder)
        self.decoder = LSTM_Decoder(input_size, hidden_size, Bi_dir_encoder, dp_decoder)


def model_predict_batch(model, X, n_future):
    """X: (B, n_past, input) -> (B, n_future, input): encode once, decode recursively feeding the output back,
    starting from the last input row (the recursion of DNN_tools.py:212-234 for B sequences at once)."""
    model.eval()
    with torch.no_grad():
        h, c = model.encoder(X)
        out = torch.empty((X.shape[0], n_future, X.shape[2]), device=X.device, dtype=X.dtype)
        y = X[:, -1, :]
        for i in range(n_future):
            y, h, c = model.decoder(y, h, c)
            out[:, i, :] = y
    return out


def model_predict(device, model, X, n_future):
    """Reference signature (DNN_tools.py:212-234): X (n_past, input) -> (n_future, input) tensor on `device`."""
    return model_predict_batch(model, X.unsqueeze(0).to(device), n_future)[0]


def _decode(model, X, n_future, truth=None, ratio=0.0):
    """encoder once, n_future recursive decoder steps; with `truth` and ratio > 0: mixed teacher forcing"""
    h, c = model.encoder(X)
    y = X[:, -1, :]
    outs = []
    for i in range(n_future):
        y, h, c = model.decoder(y, h, c)
        outs.append(y)
        if truth is not None and random.random() < ratio:
            y = truth[:, i, :]
    return torch.stack(outs, dim=1)


def _scores(criterion, out, truth):
    """(mse, R2-type accuracy, relative accuracy) as the reference reports them (DNN_tools.py:146-155)"""
    loss = criterion(out, truth)
    r2 = 1.0 - loss / criterion(truth, torch.mean(truth) + torch.zeros_like(truth))
    rel = 1.0 - loss / criterion(truth, torch.zeros_like(truth))
    return loss, r2, rel


def model_train(device, model, trainloader, criterion, optimizer, n_future, training_method='recursive', ratio=0.5):
    """One epoch over `trainloader` (DNN_tools.py:103-165): returns (sum loss, sum R2, sum rel, model).
    training_method 'recursive' feeds predictions back; 'mtf' mixes in the truth with probability `ratio`,
    lowered by 0.005 per batch."""
    tot = [0.0, 0.0, 0.0]
    model.train()
    for xb, yb in trainloader:
        optimizer.zero_grad()
        out = _decode(model, xb, n_future, yb if training_method == 'mtf' else None, ratio)
        loss, r2, rel = _scores(criterion, out, yb)
        tot[0] += loss.item(); tot[1] += r2.item(); tot[2] += rel.item()
        loss.backward()
        optimizer.step()
        if ratio > 0.005:
            ratio -= 0.005
    return tot[0], tot[1], tot[2], model


def model_test(device, model, testloader, criterion, n_future):
    """Validation pass without gradients (DNN_tools.py:170-207): (sum loss, sum R2, sum rel)."""
    tot = [0.0, 0.0, 0.0]
    model.eval()
    with torch.no_grad():
        for xb, yb in testloader:
            loss, r2, rel = _scores(criterion, _decode(model, xb, n_future), yb)
            tot[0] += loss.item(); tot[1] += r2.item(); tot[2] += rel.item()
    return tot[0], tot[1], tot[2]


class MyDataset(Dataset):
    """(x[i], y[i]) pairs, first dimension = sample (DNN_tools.py:239-253)."""

    def __init__(self, x, y):
        super().__init__()
        assert x.shape[0] == y.shape[0]
        self.x, self.y = x, y

    def __len__(self):
        return self.y.shape[0]

    def __getitem__(self, index):
        return self.x[index], self.y[index]


def Scale_to_zero_one(X, Y):
    """Map both tensors with the joint extrema to [-1, 0]; returns (X, Y, scale_max, scale_min) (DNN_tools.py:259-269)."""
    scale_min, scale_max = min(X.min(), Y.min()), max(X.max(), Y.max())
    span = -scale_min + scale_max
    return (X - scale_max) / span, (Y - scale_max) / span, scale_max, scale_min


def scale_forward(X, scale_max, scale_min):
    """(X - max) / (max - min) with the training extrema (DNN_tools.py:272-274)."""
    return (X - scale_max) / (-scale_min + scale_max)


def scale_it_back(X, scale_max, scale_min):
    """Inverse of scale_forward (DNN_tools.py:277-279)."""
    return X * (scale_max - scale_min) + scale_max


def windows_from_history(history, input_size, filter_size, n_past, n_future, cut_off, device="cpu"):
    """Training windows from a (steps, input) displacement history: keep the first `cut_off` fraction, every
    `filter_size`-th row, then all (n_past, n_future) sliding windows (DNN_tools.py:284-313)."""
    H = np.asarray(history)
    H = H[0:int(cut_off * len(H)), :][0::filter_size, :]
    T = torch.from_numpy(np.ascontiguousarray(H)).float().to(device)
    groups = T.shape[0] - n_future - n_past + 1
    if groups <= 0:
        z = torch.zeros((0, n_past, input_size), device=device)
        return z, torch.zeros((0, n_future, input_size), device=device)
    W = T.unfold(0, n_past + n_future, 1).permute(0, 2, 1)          # (groups, n_past + n_future, input)
    return W[:, :n_past, :].contiguous(), W[:, n_past:, :].contiguous()


def Dis_data_filtered_subset_coronary(device, input_size, filter_size, n_past, n_future, Path, cut_off):
    """Same from the 'Displacement' dataset (input, steps) of an HDF5 file (DNN_tools.py:284-313)."""
    import h5py
    with h5py.File(Path, 'r') as f:
        data = np.array(f['Displacement'])
    return windows_from_history(data.transpose(), input_size, filter_size, n_past, n_future, cut_off, device)
