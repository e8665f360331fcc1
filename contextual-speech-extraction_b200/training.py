"""Differentiable Sepformer forward for the training scripts (fp32 parity mode).

`Sepformer.forward` of the reference under autograd (train_ContSep.py:384-419,
train_ContExt.py:366-389): every stage is an `autograd.Function` whose forward and backward are
C-ABI calls into libcse_b200.so (include/cse_b200.h, "backward" / "training path" sections).
PyTorch carries the graph, the `nn.Parameter`s and their `.grad` — nothing is computed by eager
ops except views, the division by K in pred_head and the bias-scale of one [256*spk] vector — so
stock DDP (gradient all-reduce over NCCL, train_ContExt.py:269-273) works unchanged on top.

Layout is channels-last throughout: [B,L,256], chunks [B,S,K,256], residual streams
[(b,s), c+k, 256] / [(b,k), c+s, 256].  Transformer layers checkpoint their input only.
"""
import ctypes as C

import torch
from torch.autograd import Function

from . import _lib
from .backward import LAYER_KEYS, transformer_layer
from .runtime import current_stream

N, K, LAYERS, BLOCKS, CTX = _lib.N, _lib.K_CHUNK, _lib.LAYERS, _lib.BLOCKS, _lib.CTX
FP32 = _lib.FP32


def _st(t):
    return C.c_void_p(current_stream(t.device))


def _f32(t, name="tensor"):
    if not t.is_cuda:
        raise _lib.CseError(f"{name} is on {t.device}: the CUDA path has no CPU fallback")
    return t.contiguous().float()


_ZEROS = {}


def _zeros(device, rows):
    """Shared all-zero [rows,256] table (stands in for `pe` / `ctok` in adjoint relayouts)."""
    key = (device, rows)
    z = _ZEROS.get(key)
    if z is None:
        z = torch.zeros(rows, N, dtype=torch.float32, device=device)
        _ZEROS[key] = z
    return z


class EncoderFn(Function):
    """speechbrain Encoder (ContSep.py:10,69): mix [B,T] -> relu(conv1d) channels-last [B,L,256]."""

    @staticmethod
    def forward(ctx, mix, w):
        B, T = mix.shape
        L = (T - 16) // 8 + 1
        out = torch.empty(B, L, N, dtype=torch.float32, device=mix.device)
        part = torch.empty(B, (L + 63) // 64 + 1, 2, dtype=torch.float32, device=mix.device)
        n_parts = C.c_int(0)
        _lib.call("cse_encoder_fwd", _lib.ptr(mix), _lib.ptr(w), B, T, FP32, _lib.ptr(out), _lib.ptr(part),
                  C.byref(n_parts), _st(mix))
        ctx.save_for_backward(mix, w, out)
        return out

    @staticmethod
    def backward(ctx, dE):
        mix, w, E = ctx.saved_tensors
        dw = torch.zeros_like(w)
        _lib.call("cse_encoder_bwd", _lib.ptr(mix), _lib.ptr(E), _lib.ptr(dE.contiguous()), mix.shape[0],
                  mix.shape[1], _lib.ptr(dw), _st(mix))
        return None, dw


class GroupNormFn(Function):
    """nn.GroupNorm(1,256,eps=1e-8) per sample over x [B,rows,256] (+ skip) (ContSep.py:226,498-502,527-531)."""

    @staticmethod
    def forward(ctx, x, g, b, skip):
        B, rows = x.shape[0], x.shape[1]
        out = torch.empty_like(x)
        stat = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        part = torch.empty(B * 64 * 2, dtype=torch.float32, device=x.device)
        _lib.call("cse_groupnorm_fwd", _lib.ptr(x), _lib.ptr(g), _lib.ptr(b), _lib.ptr(skip), B, rows, 1e-8,
                  _lib.ptr(out), _lib.ptr(stat), _lib.ptr(part), _st(x))
        ctx.save_for_backward(x, g, stat)
        ctx.has_skip = skip is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, g, stat = ctx.saved_tensors
        B, rows = x.shape[0], x.shape[1]
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg, db = torch.zeros_like(g), torch.zeros_like(g)
        part = torch.empty(B * 64 * 2, dtype=torch.float32, device=x.device)
        _lib.call("cse_groupnorm_bwd", _lib.ptr(x), _lib.ptr(stat), _lib.ptr(g), _lib.ptr(dy), B, rows,
                  _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(part), _st(x))
        return dx, dg, db, (dy if ctx.has_skip else None)


class LinearFn(Function):
    """nn.Linear / 1x1 conv in fp32: y [M,Nout] = a [M,Kin] W^T + bias_scale * bias."""

    @staticmethod
    def forward(ctx, a, W, bias, bias_scale):
        M, Kin = a.shape
        Nout = W.shape[0]
        y = torch.empty(M, Nout, dtype=torch.float32, device=a.device)
        _lib.call("cse_linear", _lib.ptr(a), Kin, _lib.ptr(W), _lib.ptr(bias), float(bias_scale), None,
                  _lib.ptr(y), Nout, M, Nout, Kin, 0, 1, FP32, _st(a))
        ctx.save_for_backward(a, W)
        ctx.bias_scale = float(bias_scale) if bias is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        a, W = ctx.saved_tensors
        M, Kin = a.shape
        Nout = W.shape[0]
        dy = dy.contiguous()
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        dW = torch.zeros_like(W)
        db = torch.zeros(Nout, dtype=torch.float32, device=a.device) if ctx.bias_scale is not None else None
        wt = torch.empty(Nout * Kin, dtype=torch.float32, device=a.device) if da is not None else None
        _lib.call("cse_linear_bwd", _lib.ptr(a), Kin, _lib.ptr(W), _lib.ptr(dy), Nout, M, Nout, Kin,
                  _lib.ptr(da), Kin, _lib.ptr(dW), _lib.ptr(db), _lib.ptr(wt), _st(a))
        if db is not None and ctx.bias_scale != 1.0:
            db = db * ctx.bias_scale
        return da, dW, db, None


class ContextMapFn(Function):
    """{intra,inter}_context_mapper = nn.Linear(4096,256) on the B*c prompt rows (ContSep.py:480,511)."""

    @staticmethod
    def forward(ctx, x, W, b):
        rows = x.shape[0]
        out = torch.empty(rows, N, dtype=torch.float32, device=x.device)
        _lib.call("cse_context_map", _lib.ptr(x), _lib.ptr(W), _lib.ptr(b), rows, x.shape[1], _lib.ptr(out), _st(x))
        ctx.save_for_backward(x, W)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        dW, db = torch.zeros_like(W), torch.zeros(N, dtype=torch.float32, device=x.device)
        da, wt = None, None
        if ctx.needs_input_grad[0]:      # H-ContExt: the speaker embedding reaches ctx through se_embedding
            da = torch.empty_like(x)
            wt = torch.empty(W.numel(), dtype=torch.float32, device=x.device)
        _lib.call("cse_linear_bwd", _lib.ptr(x), x.shape[1], _lib.ptr(W), _lib.ptr(dy), N, x.shape[0], N,
                  x.shape[1], _lib.ptr(da), x.shape[1], _lib.ptr(dW), _lib.ptr(db), _lib.ptr(wt), _st(x))
        return da, dW, db


class SegmentFn(Function):
    """_padding + _Segmentation (ContSep.py:270-335): x0 [B,L,256] -> X [B,S,K,256]; the adjoint is
    the overlap-add kernel with a unit PReLU slope."""

    @staticmethod
    def forward(ctx, x0, S):
        B, L, _ = x0.shape
        X = torch.empty(B, S, K, N, dtype=torch.float32, device=x0.device)
        _lib.call("cse_segment", _lib.ptr(x0), B, L, S, _lib.ptr(X), _st(x0))
        ctx.dims = (B, L, S)
        return X

    @staticmethod
    def backward(ctx, dX):
        B, L, S = ctx.dims
        dX = dX.contiguous()
        one = torch.ones(1, dtype=torch.float32, device=dX.device)
        dx0 = torch.empty(B, L, N, dtype=torch.float32, device=dX.device)
        _lib.call("cse_prelu_overlap_add", _lib.ptr(dX), _lib.ptr(one), B, S, L, FP32, _lib.ptr(dx0), _st(dX))
        return dx0, None


class BuildSequencesFn(Function):
    """Chunk tensor -> residual stream of a stack: prompt token(s) prepended, sinusoid table added
    (ContSep.py:474-482 / :506-513, CSE_transformer.py:102-104)."""

    @staticmethod
    def forward(ctx, X, ctok, pe, inter):
        B, S = X.shape[0], X.shape[1]
        c = 0 if ctok is None else ctok.shape[1]
        nseq, n = (B * K, S + c) if inter else (B * S, K + c)
        R = torch.empty(nseq * n, N, dtype=torch.float32, device=X.device)
        _lib.call("cse_build_sequences", _lib.ptr(X), _lib.ptr(ctok), _lib.ptr(pe), B, S, c, int(inter),
                  _lib.ptr(R), _st(X))
        ctx.dims = (B, S, c, int(inter))
        return R

    @staticmethod
    def backward(ctx, dR):
        B, S, c, inter = ctx.dims
        dR = dR.contiguous()
        dX = torch.empty(B, S, K, N, dtype=torch.float32, device=dR.device)
        dtok = torch.empty(B, c, N, dtype=torch.float32, device=dR.device) if c else None
        _lib.call("cse_sequences_to_chunks", _lib.ptr(dR), B, S, c, inter, _lib.ptr(dX), _lib.ptr(dtok), _st(dR))
        return dX, dtok, None, None


class SequencesToChunksFn(Function):
    """Residual stream -> chunk tensor without the prompt rows (ContSep.py:487-489 / :518-521), plus
    the per-sample sum of the prompt rows (pred_head is that sum / K, ContSep.py:516-517)."""

    @staticmethod
    def forward(ctx, R, B, S, c, inter):
        X = torch.empty(B, S, K, N, dtype=torch.float32, device=R.device)
        tok_sum = torch.empty(B, c, N, dtype=torch.float32, device=R.device) if c else None
        _lib.call("cse_sequences_to_chunks", _lib.ptr(R), B, S, c, int(inter), _lib.ptr(X), _lib.ptr(tok_sum),
                  _st(R))
        ctx.dims = (B, S, c, int(inter))
        if tok_sum is None:
            tok_sum = torch.zeros(B, 0, N, dtype=torch.float32, device=R.device)
        return X, tok_sum

    @staticmethod
    def backward(ctx, dX, dtok_sum):
        B, S, c, inter = ctx.dims
        dX = dX.contiguous()
        nseq, n = (B * K, S + c) if inter else (B * S, K + c)
        dR = torch.empty(nseq * n, N, dtype=torch.float32, device=dX.device)
        dtok = dtok_sum.contiguous() if c else None
        _lib.call("cse_build_sequences", _lib.ptr(dX), _lib.ptr(dtok), _lib.ptr(_zeros(dX.device, 2500)), B, S,
                  c, inter, _lib.ptr(dR), _st(dX))
        return dR, None, None, None, None


class LayerNormFn(Function):
    """Final LayerNorm of a stack (CSE_transformer.py:197,248; eps 1e-6) on rows [M,256]."""

    @staticmethod
    def forward(ctx, x, g, b):
        out = torch.empty_like(x)
        _lib.call("cse_layernorm_fwd", _lib.ptr(x), _lib.ptr(g), _lib.ptr(b), x.shape[0], 1e-6, FP32,
                  _lib.ptr(out), _st(x))
        ctx.save_for_backward(x, g)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, g = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg, db = torch.zeros_like(g), torch.zeros_like(g)
        _lib.call("cse_layernorm_bwd", _lib.ptr(x), _lib.ptr(g), _lib.ptr(dy), x.shape[0], 1e-6, 0, _lib.ptr(dx),
                  _lib.ptr(dg), _lib.ptr(db), _st(x))
        return dx, dg, db


class PreluOverlapAddFn(Function):
    """PReLU then _over_add (ContSep.py:244,337-370), commuted in front of conv2d."""

    @staticmethod
    def forward(ctx, X, a, L):
        B, S = X.shape[0], X.shape[1]
        U = torch.empty(B, L, N, dtype=torch.float32, device=X.device)
        _lib.call("cse_prelu_overlap_add", _lib.ptr(X), _lib.ptr(a), B, S, L, FP32, _lib.ptr(U), _st(X))
        ctx.save_for_backward(X, a)
        ctx.L = L
        return U

    @staticmethod
    def backward(ctx, dU):
        X, a = ctx.saved_tensors
        B, S = X.shape[0], X.shape[1]
        dU = dU.contiguous()
        dX = torch.empty_like(X)
        da = torch.zeros_like(a)
        _lib.call("cse_prelu_overlap_add_bwd", _lib.ptr(X), _lib.ptr(a), _lib.ptr(dU), B, S, ctx.L, _lib.ptr(dX),
                  _lib.ptr(da), _st(X))
        return dX, da, None


class GateFn(Function):
    """tanh(output) * sigmoid(output_gate) (ContSep.py:255)."""

    @staticmethod
    def forward(ctx, o, g):
        y = torch.empty_like(o)
        _lib.call("cse_gate", _lib.ptr(o), _lib.ptr(g), o.numel(), FP32, _lib.ptr(y), _st(o))
        ctx.save_for_backward(o, g)
        return y

    @staticmethod
    def backward(ctx, dy):
        o, g = ctx.saved_tensors
        dy = dy.contiguous()
        do, dg = torch.empty_like(o), torch.empty_like(g)
        _lib.call("cse_gate_bwd", _lib.ptr(o), _lib.ptr(g), _lib.ptr(dy), o.numel(), _lib.ptr(do), _lib.ptr(dg),
                  _st(o))
        return do, dg


class MaskDecodeFn(Function):
    """relu(mask) * mix_w -> ConvTranspose1d(256,1,16,stride 8) -> pad / trim (ContSep.py:263,79-95)."""

    @staticmethod
    def forward(ctx, mask_pre, E, dec_w, T, n_masks):
        B, L, _ = E.shape
        frames = torch.empty(B * L * n_masks, 16, dtype=torch.float32, device=E.device)
        est = torch.empty(B, T, n_masks, dtype=torch.float32, device=E.device)
        _lib.call("cse_mask_decode", _lib.ptr(mask_pre), _lib.ptr(E), _lib.ptr(dec_w), B, L, T, n_masks, FP32,
                  _lib.ptr(frames), _lib.ptr(est), _st(E))
        ctx.save_for_backward(mask_pre, E, dec_w)
        ctx.dims = (B, L, T, n_masks)
        return est

    @staticmethod
    def backward(ctx, d_est):
        mask_pre, E, dec_w = ctx.saved_tensors
        B, L, T, n_masks = ctx.dims
        d_est = d_est.contiguous()
        dm, dE, dw = torch.empty_like(mask_pre), torch.empty_like(E), torch.zeros_like(dec_w)
        _lib.call("cse_mask_decode_bwd", _lib.ptr(mask_pre), _lib.ptr(E), _lib.ptr(dec_w), _lib.ptr(d_est), B, L,
                  T, n_masks, _lib.ptr(dm), _lib.ptr(dE), _lib.ptr(dw), _st(E))
        return dm, dE, dw, None, None


# ----------------------------------------------------------------------------------------------
# composition
# ----------------------------------------------------------------------------------------------
def _stack(sd, prefix, X, tok, B, S, c, inter, bf16=False):
    """SBTransformerBlock_CSE on the chunk tensor: returns (chunks [B,S,K,256], prompt-row sums)."""
    R = BuildSequencesFn.apply(X, tok, sd[prefix + "pos_enc.pe"], inter)
    nseq, n = (B * K, S + c) if inter else (B * S, K + c)
    for l in range(LAYERS):
        q = f"{prefix}mdl.layers.{l}."
        R = transformer_layer({key: sd[q + key] for _, key in LAYER_KEYS}, R, nseq, n, bf16)
    R = LayerNormFn.apply(R, sd[prefix + "mdl.norm.norm.weight"], sd[prefix + "mdl.norm.norm.bias"])
    return SequencesToChunksFn.apply(R, B, S, c, inter)


def forward_train(sd, mix, ctx, n_masks, want_pred_head=False, bf16=False):
    """Differentiable encoder -> masknet(ctx) -> mask * mix_w -> decoder -> pad/trim.

    sd: {reference state_dict key -> tensor} (`_SepformerBase._tensors()`); mix [B,T]; ctx [B,c,4096] or
    None.  Returns (est [B,T,n_masks], pred_head [B,256] | None), same as `_SepformerBase._run`.
    bf16 (the step runs under torch.autocast): the 32 transformer layers — 97 % of the FLOPs — use the bf16
    tensor-core kernels forward and backward; every other stage, the residual stream, the norms, the losses and all
    gradients stay fp32 (autocast keeps them fp32 too)."""
    mix = _f32(mix, "mix")
    B, T = mix.shape
    c = 0
    if ctx is not None:
        ctx = _f32(ctx, "ctx")
        c = ctx.shape[1]
    sh = _lib.path_shape(B, T, c, n_masks)
    L, S = sh.L, sh.S
    E = EncoderFn.apply(mix, sd["encoder.conv1d.weight"])
    x = GroupNormFn.apply(E, sd["masknet.norm.weight"], sd["masknet.norm.bias"], None)
    x0 = LinearFn.apply(x.view(B * L, N), sd["masknet.conv1d.weight"], None, 0.0)
    X = SegmentFn.apply(x0.view(B, L, N), S)
    pred_head = None
    for i in range(BLOCKS):
        p = f"masknet.dual_mdl.{i}."
        tok_i = tok_e = None
        if c:
            flat = ctx.reshape(B * c, CTX)
            tok_i = ContextMapFn.apply(flat, sd[p + "intra_context_mapper.weight"],
                                       sd[p + "intra_context_mapper.bias"]).view(B, c, N)
            tok_e = ContextMapFn.apply(flat, sd[p + "inter_context_mapper.weight"],
                                       sd[p + "inter_context_mapper.bias"]).view(B, c, N)
        Y, _ = _stack(sd, p + "intra_mdl.", X, tok_i, B, S, c, False, bf16)
        X1 = GroupNormFn.apply(Y.view(B, S * K, N), sd[p + "intra_norm.weight"], sd[p + "intra_norm.bias"],
                               X.view(B, S * K, N)).view(B, S, K, N)
        Y, tok_sum = _stack(sd, p + "inter_mdl.", X1, tok_e, B, S, c, True, bf16)
        if want_pred_head and c:
            pred_head = tok_sum[:, 0, :] / K                               # ContSep.py:516-517
        X = GroupNormFn.apply(Y.view(B, S * K, N), sd[p + "inter_norm.weight"], sd[p + "inter_norm.bias"],
                              X1.view(B, S * K, N)).view(B, S, K, N)
    U = PreluOverlapAddFn.apply(X, sd["masknet.prelu.weight"], L)
    w2d, b2d = sd["masknet.conv2d.weight"], sd["masknet.conv2d.bias"]
    if w2d.shape[0] != n_masks * N:                                        # ContExt decodes mask 0 only
        w2d, b2d = w2d[: n_masks * N], b2d[: n_masks * N]
    V = LinearFn.apply(U.view(B * L, N), w2d.reshape(n_masks * N, N), b2d, 2.0).view(B * L * n_masks, N)
    O = LinearFn.apply(V, sd["masknet.output.0.weight"], sd["masknet.output.0.bias"], 1.0)
    G = LinearFn.apply(V, sd["masknet.output_gate.0.weight"], sd["masknet.output_gate.0.bias"], 1.0)
    MP = LinearFn.apply(GateFn.apply(O, G), sd["masknet.end_conv1x1.weight"], None, 0.0)
    est = MaskDecodeFn.apply(MP, E, sd["decoder.weight"], T, n_masks)
    return est, pred_head
