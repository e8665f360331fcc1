"""The reference's op sequence on STOCK torch modules (nn.MultiheadAttention, nn.LayerNorm, nn.GroupNorm, nn.Conv1d ...).
TEST / BASELINE INFRASTRUCTURE ONLY — never imported by the product package.

Unlike `sepformer_oracle.py` (elementary einsum arithmetic, an independent statement of the maths), this file calls
the same torch modules the reference's speechbrain path instantiates, in the same order, with the same copies and
permutes: `Sepformer.forward` (ContSep.py:53-100, ContExt.py:54-129), `Dual_Path_Model_CSE.forward`
(ContSep.py:205-268), `Dual_Computation_Block_CSE.forward` (ContSep.py:453-533), `SBTransformerBlock_CSE` /
`TransformerEncoderLayer` / `MultiheadAttention` (CSE_transformer.py:90-106, 385-416, 535-557).  Its fp32 CPU output is
bit-identical to the reference modules' (`tests/test_oracle.py::test_eager_reference_is_bit_identical_to_the_reference_fixtures`).
Uses:
  * bench.py `--impl reference` / `cpu_baseline`: the reference's own CPU speed (stock torch kernels, all host threads);
  * tools/eager_yardstick.py and tests/test_baseline_shapes_gpu.py: the reference on the B200 itself — cuBLAS/cuDNN,
    flash-SDPA under autocast (train_ContSep.py:383), ATen norms, eager launches — as speed and bf16-drift yardstick.
The parameters come from our module mirror used purely as a container of stock nn.Modules (state_dict-compatible
with the reference); none of our kernels or the C ABI is on this path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

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
