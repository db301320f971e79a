"""Synchronization-avoiding time loop on the device (the loop of /root/reference/Online_predictor.py:251-318).

Phase 1 (steps 0 .. i_cri = n_past*filter_size - 1): synchronised steps; the shared-DOF rows of every d1 are
recorded into the plan's device history ring (`d_sol_shared`, Online_predictor.py:260).
Phase 2: per refill block of n_future*filter_size steps — read the last n_past*filter_size history rows on the
device, run the LSTM encoder-decoder for all filter_size interleaved combs as one batch
(`Tools.DNN_prediction.predict_block`), hand the float64 table to the plan (`saa_plan_set_prediction`) and run
the block with SAA_MODE_PREDICT: un-synchronised step, shared DOFs overwritten by the table row
(Online_predictor.py:294-298), the overwritten values appended to the history (:301).  Nothing crosses PCIe or
NVLink inside phase 2.

`resync_every = k` is the extension BASELINE config 5 asks for: during phase 2 every step whose index is a
multiple of k is a true synchronised step (the exchange runs every k steps); k = None is the reference's
behaviour (never synchronise again).
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import plan as _plan

_PKG = os.path.dirname(os.path.abspath(__file__))


def _dnn_prediction():
    if _PKG not in sys.path:
        sys.path.insert(0, _PKG)
    import Tools.DNN_prediction as P
    return P


class SyncAvoidingRun:
    """plans: the StepPlan(s) this process owns (one per GPU process, or all P of a PlanGroup on one GPU).
    stepper: the object whose .step(n, mode) advances all of them together — the plan itself for one process per
    GPU, the PlanGroup otherwise.  models / loc_dof_shared / scales: one entry per plan."""

    def __init__(self, plans, stepper, loc_dof_shared, models, scales, n_past, n_future, filter_size, device="cuda",
                 resync_every=None, keep_tables=False):
        import torch
        self.torch = torch
        self.plans, self.stepper = list(plans), stepper
        self.dofs = [np.asarray(d, dtype=np.int64) for d in loc_dof_shared]
        self.models, self.scales = list(models), list(scales)
        self.n_p, self.n_f, self.n_s = int(n_past), int(n_future), int(filter_size)
        self.device = torch.device(device)
        self.k = resync_every
        self.i = 0
        self.i_cri = self.n_p * self.n_s - 1                                   # Online_predictor.py:63
        self.block = self.n_f * self.n_s
        self.tables = [] if keep_tables else None
        self.t_predict = 0.0                                                   # wall seconds spent in LSTM inference
        self.n_predict = 0                                                     # refill inferences so far
        self._live = []                                                        # tables of the refill block in progress (referenced by the plans)
        self._blk_start = 0                                                    # step index that took row 0 of those tables
        cap = self.n_p * self.n_s + self.block
        for pl, d in zip(self.plans, self.dofs):
            pl.set_history(d, capacity=cap, save_every=1)
        for m in self.models:
            m.to(self.device).eval()

    def _history_block(self, q):
        """rows [i - n_p*n_s, i) of d_sol_shared of plan q as a device tensor"""
        torch = self.torch
        rows = self.n_p * self.n_s
        buf = torch.empty((rows, self.dofs[q].size), dtype=torch.float64, device=self.device)
        self.plans[q].read_history_dev(self.i - rows, rows, buf.data_ptr())
        self.plans[q].synchronize()
        return buf

    def _predict_tables(self):
        P = _dnn_prediction()
        out = []
        for q, pl in enumerate(self.plans):
            smax, smin = self.scales[q]
            t = P.predict_block(self.models[q], self._history_block(q), self.n_p, self.n_f, self.n_s, smax, smin)
            out.append(t)
        self.torch.cuda.synchronize(self.device)
        return out

    def run(self, test_num):
        """advance to step index `test_num` (exclusive), like `while i < test_num` of Online_predictor.py:251.

        May be called repeatedly with growing `test_num` (e.g. once per saved step): the refill block in progress —
        its prediction tables, its first step (the reference's `i` at :280) and the rows already consumed (the
        reference's counter2, :284-316) — is kept between calls, so a chunked run performs exactly the same
        inferences and uses exactly the same table rows as a single `run(T)`."""
        import time as _time
        while self.i < test_num:
            if self.i <= self.i_cri:                                           # phase 1: synchronised
                n = min(test_num, self.i_cri + 1) - self.i
                self.stepper.step(n, _plan.MODE_SYNC)
                self.i += n
                continue
            done = self.i - self._blk_start if self._live else self.block     # rows of the current block already used
            if done >= self.block:                                             # block boundary: new inference (:280)
                for pl in self.plans:
                    pl.synchronize()
                _t0 = _time.perf_counter()
                tables = self._predict_tables()
                self.t_predict += _time.perf_counter() - _t0
                self.n_predict += 1
                if self.tables is not None:
                    self.tables.append([t.cpu().numpy() for t in tables])
                self._live, self._blk_start, done = tables, self.i, 0
            tables = self._live
            self._point_plans_at(done)
            n_blk = min(self.block, done + (test_num - self.i))
            while done < n_blk:                                                # :284-316
                if self.k and (self.i % self.k == 0):
                    # true exchange on this step; its table row is skipped
                    self.stepper.step(1, _plan.MODE_SYNC)
                    self._point_plans_at(done + 1)
                    seg = 1
                else:
                    nxt = n_blk - done
                    if self.k:
                        nxt = min(nxt, self.k - (self.i % self.k))
                    self.stepper.step(nxt, _plan.MODE_PREDICT)
                    seg = nxt
                done += seg
                self.i += seg
            for pl in self.plans:
                pl.synchronize()
        return self.i

    def _point_plans_at(self, row):
        """the next SAA_MODE_PREDICT step of every plan takes row `row` of the current block's table"""
        for pl, d, t in zip(self.plans, self.dofs, self._live):
            rest = t[row:]
            # the DOF list is handed over once; afterwards only the table changes (keeps the plan's step graphs valid)
            pl.set_prediction(None if getattr(pl, "_sa_dofs_set", False) else d, rest.data_ptr() if rest.shape[0] else t.data_ptr(),
                              rest.shape[0])
            pl._sa_dofs_set = True
