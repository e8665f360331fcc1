"""Stock-PyTorch (eager) GPU yardstick for the Sepformer hot path — NOT a product path.

SURVEY.md §2b / §7 step 0: the reference owns no kernel, so "the reference on a B200" is whatever stock
PyTorch dispatches for its modules — cuDNN/cuBLAS GEMMs, the bundled flash-SDPA under autocast, ATen
norm / elementwise kernels, eager launches.  This file runs that: the same op sequence as
`Sepformer.forward` (ContSep.py:53-100, ContExt.py:54-129), `Dual_Path_Model_CSE.forward`
(ContSep.py:205-268), `Dual_Computation_Block_CSE.forward` (ContSep.py:453-533) and
`SBTransformerBlock_CSE` / `TransformerEncoderLayer` / `MultiheadAttention` (CSE_transformer.py:90-106,
385-416, 535-557), composed from the stock `torch.nn` modules our module mirror already owns as parameter
containers (`nn.Conv1d`, `nn.GroupNorm`, `nn.MultiheadAttention`, `nn.LayerNorm`, `nn.Linear`, `nn.PReLU`,
`nn.Conv2d`, `nn.ConvTranspose1d`).  None of our kernels, C ABI or engine is on this path.

    python tools/eager_yardstick.py [--batch 16] [--seconds 4] [--steps 10] [--dtype bf16|fp32]
        -> one JSON line: audio-s/s of the eager path, ours beside it (same weights, same inputs), ours / eager.

Also imported by tests/test_baseline_shapes_gpu.py as the CUDA-autocast drift yardstick (the reference's real
`--bf16` arithmetic, train_ContSep.py:383), after that test has checked this file's fp32 output against the oracle.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K = 250
P = K // 2


def _mha(att_wrapper, x):
    """CSE_transformer.py:535-557: three distinct (T,B,E) tensors -> nn.MultiheadAttention slow path."""
    q = x.permute(1, 0, 2)
    k = x.permute(1, 0, 2)
    v = x.permute(1, 0, 2)
    out, _ = att_wrapper.att(q, k, v, attn_mask=None, key_padding_mask=None, need_weights=False)
    return out.permute(1, 0, 2)


def _ffn(pos_ffn, x):
    """speechbrain PositionalwiseFeedForward: permute, Linear-ReLU-Dropout-Linear, permute back."""
    return pos_ffn.ffn(x.permute(1, 0, 2)).permute(1, 0, 2)


def _encoder_layer(layer, src):
    """Pre-norm TransformerEncoderLayer.forward (CSE_transformer.py:385-416), dropout 0."""
    src1 = layer.norm1.norm(src)
    src = src + _mha(layer.self_att, src1)
    src1 = layer.norm2.norm(src)
    return src + _ffn(layer.pos_ffn, src1)


def _block(blk, x):
    """SBTransformerBlock_CSE.forward (CSE_transformer.py:90-106): x + pe, 8 layers, final LayerNorm."""
    out = x + blk.pos_enc.pe[:, : x.size(1)].clone().detach()
    for layer in blk.mdl.layers:
        out = _encoder_layer(layer, out)
    return blk.mdl.norm.norm(out)


def _dual_block(dual, x, ctx):
    """Dual_Computation_Block_CSE.forward (ContSep.py:453-533)."""
    B, N, Kc, S = x.shape
    c = 0 if ctx is None else ctx.size(1)
    intra = x.permute(0, 3, 2, 1).contiguous().view(B * S, Kc, N)
    if c:
        tok = dual.intra_context_mapper(ctx)                                   # [B,c,N]
        intra = torch.cat([tok.unsqueeze(1).repeat(1, S, 1, 1).view(B * S, c, N), intra], 1)
    intra = _block(dual.intra_mdl, intra)
    if c:
        intra = intra[:, c:, :]
    intra = intra.reshape(B, S, Kc, N).permute(0, 3, 2, 1).contiguous()
    intra = dual.intra_norm(intra) + x                                         # skip around intra
    inter = intra.permute(0, 2, 3, 1).contiguous().view(B * Kc, S, N)
    if c:
        tok = dual.inter_context_mapper(ctx)
        inter = torch.cat([tok.unsqueeze(1).repeat(1, Kc, 1, 1).view(B * Kc, c, N), inter], 1)
    inter = _block(dual.inter_mdl, inter)
    pred_head = None
    if c:
        pred_head = inter[:, 0, :].view(B, Kc, -1).mean(1)
        inter = inter[:, c:, :]
    inter = inter.reshape(B, Kc, S, N).permute(0, 3, 1, 2).contiguous()
    out = dual.inter_norm(inter) + intra
    return out, pred_head


def _segment(x):
    """_padding + _Segmentation (ContSep.py:270-335)."""
    B, N, L = x.shape
    gap = K - (P + L % K) % K
    x = F.pad(x, (0, gap))
    pad = x.new_zeros(B, N, P)
    x = torch.cat([pad, x, pad], 2)
    a = x[:, :, :-P].contiguous().view(B, N, -1, K)
    b = x[:, :, P:].contiguous().view(B, N, -1, K)
    return torch.cat([a, b], 3).view(B, N, -1, K).transpose(2, 3).contiguous(), gap


def _over_add(x, gap):
    """_over_add (ContSep.py:337-370)."""
    B, N, Kc, S = x.shape
    x = x.transpose(2, 3).contiguous().view(B, N, -1, K * 2)
    a = x[:, :, :, :K].contiguous().view(B, N, -1)[:, :, P:]
    b = x[:, :, :, K:].contiguous().view(B, N, -1)[:, :, :-P]
    out = a + b
    return out[:, :, :-gap] if gap > 0 else out


def masknet(mn, x, ctx):
    """Dual_Path_Model_CSE.forward (ContSep.py:205-268): x [B,N,L] -> (mask [spk,B,N,L], pred_head)."""
    x = mn.conv1d(mn.norm(x))
    x, gap = _segment(x)
    pred = None
    for dual in mn.dual_mdl:
        x, pred = _dual_block(dual, x, ctx)
    x = mn.conv2d(mn.prelu(x))
    B, _, Kc, S = x.shape
    x = x.view(B * mn.num_spks, -1, Kc, S)
    x = _over_add(x, gap)
    x = mn.output(x) * mn.output_gate(x)
    x = mn.end_conv1x1(x)
    _, N, L = x.shape
    x = mn.activation(x.view(B, mn.num_spks, N, L))
    return x.transpose(0, 1), pred


def eager_forward(model, mix, ctx=None, se=None, cue="joint"):
    """The whole path with stock torch ops on `model`'s parameters.  `model` is one of our Sepformer mirrors
    (only its stock nn.Module parameter containers are used).  Returns est [B,T,spk|1] (, context_pred)."""
    extraction = hasattr(model, "add_se")                 # ContExt / H-ContExt flavour
    if extraction and model.add_se and ctx is not None:   # cue assembly, eval branch (ContExt.py:105-111)
        se = model.se_embedding(se)
        if cue == "joint":
            ctx = torch.cat([ctx, se], 1)
        elif cue == "history":
            ctx = torch.cat([ctx, torch.zeros_like(ctx)], 1)
        else:
            ctx = torch.cat([torch.zeros_like(se), se], 1)
    mix_w = F.relu(model.encoder.conv1d(mix.unsqueeze(1)))
    mask, pred_head = masknet(model.masknet, mix_w, ctx)
    dec = model.decoder
    if extraction and ctx is not None:
        est = nn.ConvTranspose1d.forward(dec, mix_w * mask[0]).squeeze(1).unsqueeze(-1)
    else:
        sep_h = torch.stack([mix_w] * model.num_spks) * mask
        est = torch.cat([nn.ConvTranspose1d.forward(dec, sep_h[i]).squeeze(1).unsqueeze(-1)
                         for i in range(model.num_spks)], -1)
    T, T_est = mix.size(1), est.size(1)
    est = F.pad(est, (0, 0, 0, T - T_est)) if T > T_est else est[:, :T, :]
    if getattr(model, "context_selector", None) is not None and pred_head is not None:
        return est, model.context_selector(pred_head)
    return est


def run_eager(model, mix, ctx, dtype, se=None, cue="joint"):
    """fp32 (TF32 off: true fp32 like the CPU reference) or the reference's autocast + flash-SDPA context
    (train_ContSep.py:383)."""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    with torch.no_grad():
        if dtype == "fp32":
            old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            try:
                return eager_forward(model, mix, ctx, se, cue)
            finally:
                torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        adt = torch.bfloat16 if dtype == "bf16" else torch.float16
        with torch.autocast("cuda", dtype=adt), sdpa_kernel(SDPBackend.FLASH_ATTENTION):
            return eager_forward(model, mix, ctx, se, cue)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seconds", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    args = ap.parse_args()
    import cse_b200  # noqa: F401
    from cse_b200 import synth
    from cse_b200.models.ContSep import Sepformer

    dev = torch.device("cuda", 0)
    T = args.seconds * 8000
    model = Sepformer(2, add_mt=True)
    model.add_mt_pipeline()
    model.load_state_dict(synth.make_state_dict("contsep", 2, seed=0))
    model = model.to(dev).eval()
    mix, _ = synth.make_mixture(args.batch, T, 2, seed=1234)
    ctx = synth.make_context(args.batch, 1, seed=1234)
    mix, ctx = mix.to(dev), ctx.to(dev)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, (time.perf_counter() - t0) / args.steps * 1e3

    eager_ms, eager_wall = timed(lambda: run_eager(model, mix, ctx, args.dtype))
    model.precision = "fp32" if args.dtype == "fp32" else "bf16"

    def ours():
        with torch.no_grad():
            return model(mix, ctx)

    ours_ms, _ = timed(ours)
    model.use_cuda_graph = True
    graph_ms, _ = timed(ours)
    est_e, _ = run_eager(model, mix, ctx, args.dtype)
    est_o, _ = ours()
    rel = ((est_o.double() - est_e.double()).norm() / est_e.double().norm()).item()
    audio = args.batch * args.seconds
    print(json.dumps({
        "yardstick": "stock PyTorch eager (cuBLAS/cuDNN + flash-SDPA under autocast + ATen), reference op sequence",
        "config": {"batch": args.batch, "seconds": args.seconds, "dtype": args.dtype, "model": "ContSep 2-spk c=1"},
        "eager_ms": eager_ms, "eager_wall_ms": eager_wall, "eager_audio_s_per_s": audio / (eager_ms / 1e3),
        "ours_eager_launch_ms": ours_ms, "ours_graph_ms": graph_ms, "ours_audio_s_per_s": audio / (graph_ms / 1e3),
        "ours_over_eager": eager_ms / graph_ms, "rel_l2_ours_vs_eager": rel,
        "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}), flush=True)


if __name__ == "__main__":
    main()
