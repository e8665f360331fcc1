"""Evaluation metrics of the reference's test loop, computed and accumulated on the device.

Mirrors the objects `test.py:198-201` builds and the way `test.py:241-245,291-301` uses them:

    si_snr_criterion = torchmetrics.audio.ScaleInvariantSignalNoiseRatio().cuda()
    sdr_criterion    = torchmetrics.audio.SignalDistortionRatio().cuda()
    criterion.update(enhanced_sp.float(), gt_sp.cuda())      # per batch
    val = criterion.compute()                                 # running mean over every item

The per-item values come from `cse_tm_si_snr` / `cse_sdr`; the running (sum, count) state lives in two doubles of
device memory updated by `cse_metric_update`, so an evaluation epoch never synchronises with the host until
`compute()`.  `EvalMeter` bundles the whole per-batch bookkeeping of test.py (enhanced and unprocessed metrics,
improvements, selection accuracy).  No CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .runtime import current_stream


def _prep(t, name):
    if not t.is_cuda:
        raise _lib.CseError(f"{name} is on {t.device}: the CUDA path has no CPU fallback")
    return t.detach().float().contiguous()


def _st(t):
    return C.c_void_p(current_stream(t.device))


def signal_distortion_ratio(preds, target, use_cg_iter=None, filter_length=512, zero_mean=False, load_diag=None):
    """torchmetrics.functional.audio.signal_distortion_ratio: [..., T] -> SDR in dB per item (float32)."""
    if use_cg_iter is not None:
        raise NotImplementedError("use_cg_iter: the reference uses the default direct solver (test.py:200)")
    if preds.shape != target.shape:
        raise RuntimeError("Predictions and targets are expected to have the same shape")
    shape, T = preds.shape[:-1], preds.shape[-1]
    p = _prep(preds.reshape(-1, T), "preds")
    t = _prep(target.reshape(-1, T), "target")
    B = p.shape[0]
    out = torch.empty(B, dtype=torch.float32, device=p.device)
    nbytes = _lib.load().cse_sdr_workspace_bytes(B, T, filter_length)
    ws = torch.empty(nbytes // 8, dtype=torch.float64, device=p.device)
    _lib.call("cse_sdr", _lib.ptr(p), _lib.ptr(t), B, T, int(filter_length), int(bool(zero_mean)),
              int(load_diag is not None), float(load_diag or 0.0), _lib.ptr(out), _lib.ptr(ws), nbytes, _st(p))
    return out.reshape(shape)


class _RunningMean(nn.Module):
    """(sum, count) of a torchmetrics metric object as two device doubles."""

    def __init__(self):
        super().__init__()
        self._acc = None

    def reset(self):
        self._acc = None

    def _accumulate(self, values):
        v = values.reshape(-1)
        if self._acc is None or self._acc.device != v.device:
            self._acc = torch.zeros(2, dtype=torch.float64, device=v.device)
        _lib.call("cse_metric_update", _lib.ptr(v), v.numel(), _lib.ptr(self._acc), _st(v))

    def compute(self):
        if self._acc is None:
            raise RuntimeError("compute() called before update()")
        return (self._acc[0] / self._acc[1]).float()

    def _values(self, preds, target):
        raise NotImplementedError

    def update(self, preds, target):
        self._accumulate(self._values(preds, target))

    def forward(self, preds, target):
        v = self._values(preds, target)
        self._accumulate(v)
        return v.mean()


class SignalDistortionRatio(_RunningMean):
    """torchmetrics.audio.SignalDistortionRatio (test.py:200-201)."""

    def __init__(self, use_cg_iter=None, filter_length=512, zero_mean=False, load_diag=None):
        super().__init__()
        self.kw = dict(use_cg_iter=use_cg_iter, filter_length=filter_length, zero_mean=zero_mean, load_diag=load_diag)

    def _values(self, preds, target):
        return signal_distortion_ratio(preds, target, **self.kw)


class StreamingSiSnr(_RunningMean):
    """torchmetrics.audio.ScaleInvariantSignalNoiseRatio used as an evaluation accumulator (test.py:198-199,241,244):
    same per-item values as losses.ScaleInvariantSignalNoiseRatio, state kept on the device."""

    def _values(self, preds, target):
        T = preds.shape[-1]
        p = _prep(preds.reshape(-1, T), "preds")
        t = _prep(target.reshape(-1, T), "target")
        out = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
        _lib.call("cse_tm_si_snr", _lib.ptr(p), _lib.ptr(t), p.shape[0], T, _lib.ptr(out), _st(p))
        return out


class EvalMeter:
    """The per-batch metric bookkeeping of test.py:241-255,291-301 without host round trips.

        meter = EvalMeter()
        meter.update(enhanced [B,T], mixed [B,T], gt [B,T], interferers=[ns1 [B,T], ...])
        meter.compute() -> dict(si_snr, sdr, si_snr_i, sdr_i, acc)       # the only synchronisation
    """

    def __init__(self):
        self.si_snr, self.si_snr_prev = StreamingSiSnr(), StreamingSiSnr()
        self.sdr, self.sdr_prev = SignalDistortionRatio(), SignalDistortionRatio()
        self._acc = _RunningMean()

    def update(self, enhanced, mixed, gt, interferers=()):
        from . import selection
        self.si_snr.update(enhanced, gt)
        self.sdr.update(enhanced, gt)
        self.si_snr_prev.update(mixed, gt)
        self.sdr_prev.update(mixed, gt)
        if len(interferers):
            acc, _ = selection.selection_accuracy(enhanced, torch.stack([gt] + list(interferers), -1))
            self._acc._accumulate(acc.float())

    def compute(self):
        out = dict(si_snr=self.si_snr.compute(), sdr=self.sdr.compute())
        out["si_snr_i"] = out["si_snr"] - self.si_snr_prev.compute()
        out["sdr_i"] = out["sdr"] - self.sdr_prev.compute()
        if self._acc._acc is not None:
            out["acc"] = self._acc.compute()
        return {k: float(v) for k, v in out.items()}
