"""SI-SNR losses with the reference's call signatures, computed by the CUDA reduction kernels.

Mirrors `speechbrain.nnet.losses.{cal_si_snr, get_si_snr_with_pitwrapper}`
(train_ContSep.py:346,352,386,391-393) and `torchmetrics.audio.ScaleInvariantSignalNoiseRatio`
(train_ContExt.py:339,367).  Unlike speechbrain's cal_si_snr, inputs are never mutated.
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


def cal_si_snr(source, estimate):
    """speechbrain call shape: source, estimate [T,B,C] -> NEGATIVE SI-SNR [1,B,C]."""
    assert source.size() == estimate.size()
    T, B, Cn = source.shape
    s = _prep(source.permute(1, 0, 2), "source")
    e = _prep(estimate.permute(1, 0, 2), "estimate")
    out = torch.empty(B, Cn, dtype=torch.float32, device=s.device)
    _lib.call("cse_si_snr", _lib.ptr(s), _lib.ptr(e), B, T, Cn, _lib.ptr(out), C.c_void_p(current_stream(s.device)))
    return out.unsqueeze(0)


def get_si_snr_with_pitwrapper(source, estimate_source, return_perms=False):
    """[B,T,C] x2 -> permutation-invariant loss [B] (min over permutations of the mean pairwise
    negative SI-SNR)."""
    B, T, Cn = source.shape
    s = _prep(source, "source")
    e = _prep(estimate_source, "estimate_source")
    loss = torch.empty(B, dtype=torch.float32, device=s.device)
    perm = torch.empty(B, Cn, dtype=torch.int32, device=s.device)
    _lib.call("cse_pit_si_snr", _lib.ptr(s), _lib.ptr(e), B, T, Cn, _lib.ptr(loss), _lib.ptr(perm),
              C.c_void_p(current_stream(s.device)))
    if return_perms:
        return loss, perm
    return loss


def scale_invariant_signal_noise_ratio(preds, target):
    """torchmetrics functional: [..., T] -> SI-SNR in dB per item."""
    shape = preds.shape[:-1]
    T = preds.shape[-1]
    p = _prep(preds.reshape(-1, T), "preds")
    t = _prep(target.reshape(-1, T), "target")
    out = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
    _lib.call("cse_tm_si_snr", _lib.ptr(p), _lib.ptr(t), p.shape[0], T, _lib.ptr(out),
              C.c_void_p(current_stream(p.device)))
    return out.reshape(shape)


class ScaleInvariantSignalNoiseRatio(nn.Module):
    """torchmetrics.audio.ScaleInvariantSignalNoiseRatio: forward() returns the batch mean and
    accumulates the running mean returned by compute()."""

    def __init__(self):
        super().__init__()
        self.reset()

    def reset(self):
        self._sum, self._n = 0.0, 0

    def update(self, preds, target):
        v = scale_invariant_signal_noise_ratio(preds, target)
        self._sum = self._sum + v.sum()
        self._n += v.numel()

    def compute(self):
        return self._sum / self._n

    def forward(self, preds, target):
        v = scale_invariant_signal_noise_ratio(preds, target)
        self._sum = self._sum + v.sum()
        self._n += v.numel()
        return v.mean()
