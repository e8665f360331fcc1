"""SI-SNR losses with the reference's call signatures, computed by the CUDA reduction kernels.

Mirrors `speechbrain.nnet.losses.{cal_si_snr, get_si_snr_with_pitwrapper}`
(train_ContSep.py:346,352,386,391-393) and `torchmetrics.audio.ScaleInvariantSignalNoiseRatio`
(train_ContExt.py:339,367).  Unlike speechbrain's cal_si_snr, inputs are never mutated.

All three are differentiable: each is an `autograd.Function` whose backward is the matching
`cse_*_si_snr_bwd` kernel (the training losses `loss.backward()` starts from,
train_ContSep.py:391-393,402; train_ContExt.py:367,372).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .runtime import current_stream


def _prep(t, name):
    """fp32, contiguous, on the GPU — through differentiable torch plumbing (no detach)."""
    if not t.is_cuda:
        raise _lib.CseError(f"{name} is on {t.device}: the CUDA path has no CPU fallback")
    return t.float().contiguous()


def _st(t):
    return C.c_void_p(current_stream(t.device))


class _CalSiSnr(torch.autograd.Function):
    """cse_si_snr / cse_si_snr_bwd on batch-major [B,T,C] tensors."""

    @staticmethod
    def forward(ctx, s, e):
        B, T, Cn = s.shape
        out = torch.empty(B, Cn, dtype=torch.float32, device=s.device)
        _lib.call("cse_si_snr", _lib.ptr(s), _lib.ptr(e), B, T, Cn, _lib.ptr(out), _st(s))
        ctx.save_for_backward(s, e)
        return out

    @staticmethod
    def backward(ctx, g):
        s, e = ctx.saved_tensors
        B, T, Cn = s.shape
        g = g.float().contiguous()
        ds = torch.empty_like(s) if ctx.needs_input_grad[0] else None
        de = torch.empty_like(e) if ctx.needs_input_grad[1] else None
        _lib.call("cse_si_snr_bwd", _lib.ptr(s), _lib.ptr(e), _lib.ptr(g), B, T, Cn, _lib.ptr(ds), _lib.ptr(de),
                  _st(s))
        return ds, de


class _PitSiSnr(torch.autograd.Function):
    """cse_pit_si_snr / cse_pit_si_snr_bwd; the chosen permutation is a constant of the backward pass."""

    @staticmethod
    def forward(ctx, s, e):
        B, T, Cn = s.shape
        loss = torch.empty(B, dtype=torch.float32, device=s.device)
        perm = torch.empty(B, Cn, dtype=torch.int32, device=s.device)
        _lib.call("cse_pit_si_snr", _lib.ptr(s), _lib.ptr(e), B, T, Cn, _lib.ptr(loss), _lib.ptr(perm), _st(s))
        ctx.save_for_backward(s, e, perm)
        ctx.mark_non_differentiable(perm)
        return loss, perm

    @staticmethod
    def backward(ctx, g, _g_perm):
        s, e, perm = ctx.saved_tensors
        B, T, Cn = s.shape
        g = g.float().contiguous()
        ds = torch.empty_like(s) if ctx.needs_input_grad[0] else None
        de = torch.empty_like(e) if ctx.needs_input_grad[1] else None
        _lib.call("cse_pit_si_snr_bwd", _lib.ptr(s), _lib.ptr(e), _lib.ptr(g), _lib.ptr(perm), B, T, Cn,
                  _lib.ptr(ds), _lib.ptr(de), _st(s))
        return ds, de


class _TmSiSnr(torch.autograd.Function):
    """cse_tm_si_snr / cse_tm_si_snr_bwd on [B,T]."""

    @staticmethod
    def forward(ctx, p, t):
        out = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
        _lib.call("cse_tm_si_snr", _lib.ptr(p), _lib.ptr(t), p.shape[0], p.shape[1], _lib.ptr(out), _st(p))
        ctx.save_for_backward(p, t)
        return out

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        g = g.float().contiguous()
        dp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        dt = torch.empty_like(t) if ctx.needs_input_grad[1] else None
        _lib.call("cse_tm_si_snr_bwd", _lib.ptr(p), _lib.ptr(t), _lib.ptr(g), p.shape[0], p.shape[1],
                  _lib.ptr(dp), _lib.ptr(dt), _st(p))
        return dp, dt


def cal_si_snr(source, estimate):
    """speechbrain call shape: source, estimate [T,B,C] -> NEGATIVE SI-SNR [1,B,C]."""
    assert source.size() == estimate.size()
    s = _prep(source.permute(1, 0, 2), "source")
    e = _prep(estimate.permute(1, 0, 2), "estimate")
    return _CalSiSnr.apply(s, e).unsqueeze(0)


def get_si_snr_with_pitwrapper(source, estimate_source, return_perms=False):
    """[B,T,C] x2 -> permutation-invariant loss [B] (min over permutations of the mean pairwise
    negative SI-SNR)."""
    s = _prep(source, "source")
    e = _prep(estimate_source, "estimate_source")
    loss, perm = _PitSiSnr.apply(s, e)
    if return_perms:
        return loss, perm
    return loss


def scale_invariant_signal_noise_ratio(preds, target):
    """torchmetrics functional: [..., T] -> SI-SNR in dB per item."""
    shape = preds.shape[:-1]
    T = preds.shape[-1]
    p = _prep(preds.reshape(-1, T), "preds")
    t = _prep(target.reshape(-1, T), "target")
    return _TmSiSnr.apply(p, t).reshape(shape)


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
        self._sum = self._sum + v.detach().sum()
        self._n += v.numel()

    def compute(self):
        return self._sum / self._n

    def forward(self, preds, target):
        v = scale_invariant_signal_noise_ratio(preds, target)
        self._sum = self._sum + v.detach().sum()
        self._n += v.numel()
        return v.mean()
