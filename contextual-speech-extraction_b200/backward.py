"""Host side of the training path: hand-written CUDA gradients behind torch.autograd.

The reference trains through autograd (`loss.backward()`, train_ContSep.py:402-419,
train_ContExt.py:372-389).  Here each differentiable unit of the hot path is an
`autograd.Function` whose forward AND backward are C-ABI calls (`cse_layer_fwd` / `cse_layer_bwd`,
`cse_*_si_snr` / `cse_*_si_snr_bwd`): PyTorch only carries the graph, the parameter ownership and
the `.grad` buffers (so stock DDP gradient all-reduce applies unchanged).  No CPU / eager fallback.

This module: one pre-norm transformer layer (TransformerEncoderLayer.forward, CSE_transformer.py:385-416)
as an autograd node with activation checkpointing at layer granularity (fp32 parity mode).  The other
stages and the composition of the whole forward live in training.py, the losses in losses.py.
"""
import ctypes as C
import os

import torch

from . import _lib
from .runtime import current_stream

LAYER_KEYS = (
    ("in_proj_w", "self_att.att.in_proj_weight"), ("in_proj_b", "self_att.att.in_proj_bias"),
    ("out_proj_w", "self_att.att.out_proj.weight"), ("out_proj_b", "self_att.att.out_proj.bias"),
    ("ffn1_w", "pos_ffn.ffn.0.weight"), ("ffn1_b", "pos_ffn.ffn.0.bias"),
    ("ffn2_w", "pos_ffn.ffn.3.weight"), ("ffn2_b", "pos_ffn.ffn.3.bias"),
    ("ln1_g", "norm1.norm.weight"), ("ln1_b", "norm1.norm.bias"),
    ("ln2_g", "norm2.norm.weight"), ("ln2_b", "norm2.norm.bias"),
)


# Performance-mode training layers (bf16 tensor-core forward through cse_layer_fwd, backward through
# cse_layer_bwd_bf16): used when the training step runs under torch.autocast (the reference's --fp16 / --bf16
# switch, train_ContSep.py:383).  CSE_TRAIN_BF16=0 forces the fp32 parity kernels even under autocast,
# CSE_TRAIN_BF16=1 forces the tensor-core layers even without autocast (A/B aid).
_ENV_BF16 = os.environ.get("CSE_TRAIN_BF16")


def use_bf16_layers(autocast_bf16):
    if _ENV_BF16 == "1":
        return True
    if _ENV_BF16 == "0":
        return False
    return bool(autocast_bf16)


def _bf16_packs(params, lp):
    packs = {}
    for field, key in LAYER_KEYS[:8:2]:                          # the four weight matrices
        packs[field] = params[key].detach().to(torch.bfloat16).contiguous()
        setattr(lp, field + "_bf16", C.c_void_p(packs[field].data_ptr()))
    return packs


def _check(t, name):
    if not t.is_cuda:
        raise _lib.CseError(f"{name} is on {t.device}: the CUDA path has no CPU fallback")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise _lib.CseError(f"{name} must be a contiguous float32 tensor (got {t.dtype})")
    return t


def _layer_struct(params, cls):
    """params: {reference key relative to the layer prefix -> tensor}."""
    s = cls()
    for field, key in LAYER_KEYS:
        setattr(s, field, C.c_void_p(_check(params[key], key).data_ptr()))
    return s


def _workspace(nseq, n, device):
    nbytes = _lib.load().cse_layer_workspace_bytes(nseq, n)
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _zero_grads(params):
    """Gradient buffers of one layer as views of ONE zero-filled allocation (one fill launch instead of twelve)."""
    sizes = [params[key].numel() for _, key in LAYER_KEYS]
    offs, total = [], 0
    for sz in sizes:                      # every view 16-byte aligned (float4 / TMA accesses in the kernels)
        offs.append(total)
        total += (sz + 3) // 4 * 4
    ref = params[LAYER_KEYS[0][1]]
    flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
    return {key: flat[o:o + sz].view(params[key].shape) for (_, key), o, sz in zip(LAYER_KEYS, offs, sizes)}


def layer_forward(params, R, nseq, n, experimental_bf16=False, packs_out=None):
    """TransformerEncoderLayer.forward on the fp32 residual stream R [nseq*n, 256]; returns a new tensor.
    fp32 arithmetic unless experimental_bf16 (the bench path's tensor-core launch sequence, one layer).
    packs_out: list that receives the bf16 weight copies, for the backward pass of the same step to reuse."""
    _check(R, "R")
    if R.shape != (nseq * n, _lib.N):
        raise _lib.CseError(f"R has shape {tuple(R.shape)}, expected {(nseq * n, _lib.N)}")
    out = R.clone()
    ws, nbytes = _workspace(nseq, n, R.device)
    lp = _layer_struct(params, _lib.LayerParams)
    packs = _bf16_packs(params, lp) if experimental_bf16 else None
    _lib.call("cse_layer_fwd", C.byref(lp), _lib.ptr(out), nseq, n, _lib.BF16 if experimental_bf16 else _lib.FP32,
              _lib.ptr(ws), nbytes, C.c_void_p(current_stream(R.device)))
    if packs_out is not None and packs is not None:
        packs_out.append(packs)
    del packs
    return out


def layer_backward(params, R_in, dR_out, nseq, n, grads=None, experimental_bf16=False, packs=None):
    """Gradient of layer_forward: returns (dR_in, grads) with grads keyed like `params`.
    `grads` may carry existing buffers to accumulate into (autograd .grad semantics).

    experimental_bf16: route to cse_layer_bwd_bf16 (tensor-core recompute / dgrad / wgrad / attention backward:
    the autocast training path)."""
    _check(R_in, "R_in")
    _check(dR_out, "dR_out")
    if grads is None:
        grads = _zero_grads(params)
    dR = dR_out.clone()
    lp = _layer_struct(params, _lib.LayerParams)
    lg = _layer_struct(grads, _lib.LayerGrads)
    st = C.c_void_p(current_stream(R_in.device))
    if experimental_bf16:
        if packs is None:                # (kept alive until the call below is enqueued)
            packs = _bf16_packs(params, lp)
        else:                            # the forward's bf16 weight copies: the weights do not change within a step
            for field, _ in LAYER_KEYS[:8:2]:
                setattr(lp, field + "_bf16", C.c_void_p(packs[field].data_ptr()))
        nbytes = _lib.load().cse_layer_bwd_bf16_workspace_bytes(nseq, n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=R_in.device)
        _lib.call("cse_layer_bwd_bf16", C.byref(lp), C.byref(lg), _lib.ptr(R_in), _lib.ptr(dR), nseq, n,
                  _lib.ptr(ws), nbytes, st)
        return dR, grads
    ws, nbytes = _workspace(nseq, n, R_in.device)
    _lib.call("cse_layer_bwd", C.byref(lp), C.byref(lg), _lib.ptr(R_in), _lib.ptr(dR), nseq, n, _lib.ptr(ws),
              nbytes, st)
    return dR, grads


class _LayerFn(torch.autograd.Function):
    """One transformer layer as an autograd node: saves only its input (checkpointing)."""

    @staticmethod
    def forward(ctx, R, nseq, n, bf16, *weights):
        params = {key: w for (_, key), w in zip(LAYER_KEYS, weights)}
        ctx.save_for_backward(R, *weights)
        ctx.shape = (nseq, n, bool(bf16))
        ctx.packs = []
        ctx.versions = [w._version for w in weights]
        return layer_forward(params, R.contiguous(), nseq, n, experimental_bf16=bool(bf16), packs_out=ctx.packs)

    @staticmethod
    def backward(ctx, dR_out):
        R, *weights = ctx.saved_tensors
        nseq, n, bf16 = ctx.shape
        params = {key: w for (_, key), w in zip(LAYER_KEYS, weights)}
        same = all(w._version == v for w, v in zip(weights, ctx.versions))   # untouched since the forward pass
        packs = ctx.packs[0] if (ctx.packs and same) else None
        dR, grads = layer_backward(params, R.contiguous(), dR_out.contiguous(), nseq, n, experimental_bf16=bf16,
                                   packs=packs)
        return (dR, None, None, None) + tuple(grads[key] for _, key in LAYER_KEYS)


def transformer_layer(params, R, nseq, n, bf16=False):
    """Differentiable transformer layer: R [nseq*n,256] fp32 -> same; fp32 parity kernels (cse_layer_bwd) or, with
    bf16=True, the tensor-core forward / recompute / dgrad / wgrad (cse_layer_bwd_bf16)."""
    return _LayerFn.apply(R, nseq, n, use_bf16_layers(bf16), *[params[key] for _, key in LAYER_KEYS])
