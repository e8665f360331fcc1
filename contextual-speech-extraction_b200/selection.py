"""ContSep selection tail on the device: the consumer of `Sepformer.forward`'s
`(est_source, context_pred)` in the reference's train / eval loops, without their host round trips.

  selection_loss     train_ContSep.py:386-388   per-stream SI-SNR vs the ground truth -> argmax label ->
                                                CrossEntropy / BCEWithLogits of the selector logits
  select_stream      test.py:234-239            stream chosen by the selector (`ctx_pred.cpu()` in the reference)
  selection_accuracy test.py:248-255            SI-SNR against the target >= SI-SNR against every interferer

Each is one C-ABI call (`cse_selection_loss`, `cse_select_stream`, `cse_selection_accuracy`); no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from .runtime import current_stream


def _prep(t, name):
    if not t.is_cuda:
        raise _lib.CseError(f"{name} is on {t.device}: the CUDA path has no CPU fallback")
    return t.float().contiguous()


def _st(t):
    return C.c_void_p(current_stream(t.device))


class _SelectionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, est, gt, ce):
        B, T, n = est.shape
        dev = est.device
        sisnr = torch.empty(B, n, dtype=torch.float32, device=dev)
        label = torch.empty(B, dtype=torch.int64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dlogits = torch.empty_like(logits)
        item = torch.empty(B, dtype=torch.float32, device=dev)
        _lib.call("cse_selection_loss", _lib.ptr(gt), _lib.ptr(est), _lib.ptr(logits), B, T, n, int(ce),
                  _lib.ptr(sisnr), _lib.ptr(label), _lib.ptr(loss), _lib.ptr(dlogits), _lib.ptr(item), _st(est))
        ctx.save_for_backward(dlogits)
        ctx.mark_non_differentiable(label, sisnr)
        return loss.squeeze(0), label, sisnr

    @staticmethod
    def backward(ctx, g_loss, g_label, g_sisnr):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g_loss, None, None, None


def selection_loss(ctx_pred, est, gt, ce=True):
    """train_ContSep.py:386-388.  ctx_pred [B,spk] (ce) or [B,1] (2-spk BCE head), est [B,T,spk], gt [B,T] ->
    (ctx_loss scalar — differentiable w.r.t. ctx_pred only, the estimate is detached as in the reference —,
     context_index [B] int64, sisnrs [B,spk])."""
    est = _prep(est.detach(), "est")
    gt = _prep(gt, "gt")
    B, T, n = est.shape
    if gt.shape != (B, T):
        raise RuntimeError(f"gt must be [B,T] = {(B, T)}, got {tuple(gt.shape)}")
    logits = _prep(ctx_pred, "ctx_pred")
    if ce:
        if logits.shape != (B, n):
            raise RuntimeError(f"ctx_pred must be [B,{n}] for the cross-entropy head, got {tuple(logits.shape)}")
    else:
        if n != 2 or logits.numel() != B:
            raise RuntimeError("the BCE head needs 2 streams and one logit per item (ContSep.py:46-51)")
        logits = logits.reshape(B)
    loss, label, sisnr = _SelectionLoss.apply(logits, est, gt, bool(ce))
    return loss, label, sisnr


def select_stream(est, ctx_pred, ce=True):
    """test.py:234-239: (enhanced [B,T], pick [B] int64) — the stream the selector chose, no host sync."""
    est = _prep(est, "est")
    B, T, n = est.shape
    logits = _prep(ctx_pred, "ctx_pred").reshape(B, -1)
    if (ce and logits.shape[1] != n) or (not ce and (n != 2 or logits.shape[1] != 1)):
        raise RuntimeError(f"ctx_pred {tuple(ctx_pred.shape)} does not match {n} streams (ce={ce})")
    out = torch.empty(B, T, dtype=torch.float32, device=est.device)
    pick = torch.empty(B, dtype=torch.int64, device=est.device)
    _lib.call("cse_select_stream", _lib.ptr(est), _lib.ptr(logits), B, T, n, int(ce), _lib.ptr(out), _lib.ptr(pick),
              _st(est))
    return out, pick


def selection_accuracy(enhanced, sources):
    """test.py:248-255: enhanced [B,T]; sources [B,T,C] with the target speaker in column 0 and the interferers after it ->
    (acc [B] int32, sisnrs [B,C])."""
    enhanced = _prep(enhanced, "enhanced")
    sources = _prep(sources, "sources")
    B, T = enhanced.shape
    if sources.dim() != 3 or sources.shape[:2] != (B, T):
        raise RuntimeError(f"sources must be [B,T,C] with B,T = {(B, T)}, got {tuple(sources.shape)}")
    n = sources.shape[2]
    sisnr = torch.empty(B, n, dtype=torch.float32, device=enhanced.device)
    acc = torch.empty(B, dtype=torch.int32, device=enhanced.device)
    _lib.call("cse_selection_accuracy", _lib.ptr(enhanced), _lib.ptr(sources), B, T, n, _lib.ptr(sisnr), _lib.ptr(acc),
              _st(enhanced))
    return acc, sisnr
