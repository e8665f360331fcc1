"""CPU restatement of the reference's Sepformer hot path.  TEST INFRASTRUCTURE ONLY.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
pin is the reference's own module code, imported unchanged from /root/reference over the
speechbrain shim (`oracle/run_reference.py`): `tests/golden/make_golden.py` wrote the fixtures
in `tests/golden/*.npz` from it, `tests/test_oracle.py` checks this file against those fixtures
everywhere and against the live reference wherever /root/reference exists.
The speechbrain/torchmetrics leaves themselves are un-vendored, un-pinned third-party code
(README.md:15,21); their semantics are restated from the public 1.0.x sources in
`oracle/sb_shim` — that boundary cannot be verified offline and is flagged in DESIGN.md.

Every function works on plain tensors and a flat `state_dict` (keys = SURVEY.md §8b) in the
dtype of the weights (float32 like the reference, or float64 as an error yardstick), using
only elementary torch CPU ops so that it is an independent statement of the arithmetic
rather than a re-wrapping of nn.MultiheadAttention / nn.GroupNorm.
"""
import math
from itertools import permutations

import torch

K_CHUNK = 250
N_HEAD = 8
N_LAYER = 8
N_BLOCK = 2


# --------------------------------------------------------------------------------------
# leaves
# --------------------------------------------------------------------------------------
def encoder(sd, mix):
    """speechbrain Encoder (ContSep.py:10,69): relu(conv1d(mix.unsqueeze(1))), k=16, s=8, no bias.
    mix [B,T] -> [B,256,L]."""
    w = sd["encoder.conv1d.weight"][:, 0, :]                    # [N,16]
    frames = mix.unfold(1, w.shape[1], w.shape[1] // 2)         # [B,L,16]
    return torch.relu(torch.einsum("blk,nk->bnl", frames, w))


def group_norm1(x, weight, bias, eps=1e-8):
    """select_norm('ln') = nn.GroupNorm(1, N, eps=1e-8) (ContSep.py:164,423-424):
    one mean / biased variance per sample over every non-batch element, per-channel affine."""
    B = x.shape[0]
    flat = x.reshape(B, -1)
    mean = flat.mean(1)
    var = flat.var(1, unbiased=False)
    shape = [B] + [1] * (x.dim() - 1)
    xn = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps)
    cshape = [1, -1] + [1] * (x.dim() - 2)
    return xn * weight.view(cshape) + bias.view(cshape)


def layer_norm(x, weight, bias, eps=1e-6):
    """sb LayerNorm -> nn.LayerNorm(256, eps=1e-6) (CSE_transformer.py:197,358-359)."""
    mean = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def pad_and_segment(x, K=K_CHUNK):
    """_padding + _Segmentation (ContSep.py:270-335): out[b,n,k,s] = padded[b,n,s*P+k]."""
    B, N, L = x.shape
    P = K // 2
    gap = K - (P + L % K) % K
    padded = torch.cat([x.new_zeros(B, N, P), x, x.new_zeros(B, N, gap + P)], dim=2)
    S = (padded.shape[2] - P) // P - 1 + 1  # windows of K at hop P
    S = (padded.shape[2] - K) // P + 1
    seg = padded.unfold(2, K, P)            # [B,N,S,K]
    assert seg.shape[2] == S
    return seg.permute(0, 1, 3, 2).contiguous(), gap


def overlap_add(x, gap):
    """_over_add (ContSep.py:337-370): every frame is the sum of its two chunk copies."""
    B, N, K, S = x.shape
    P = K // 2
    total = (S + 1) * P
    out = x.new_zeros(B, N, total)
    for s in range(S):
        out[:, :, s * P:s * P + K] += x[:, :, :, s]
    out = out[:, :, P:total - P]
    if gap > 0:
        out = out[:, :, :-gap]
    return out


def multihead_attention(sd, prefix, x):
    """MultiheadAttention.forward -> nn.MultiheadAttention (CSE_transformer.py:468-477,535-557):
    packed in_proj split into q/k/v, 8 heads of 32, softmax(q k^T / sqrt(32)) v, out_proj.
    No mask, no dropout.  x [B',n,256]."""
    w_in, b_in = sd[prefix + "att.in_proj_weight"], sd[prefix + "att.in_proj_bias"]
    w_out, b_out = sd[prefix + "att.out_proj.weight"], sd[prefix + "att.out_proj.bias"]
    Bp, n, E = x.shape
    d = E // N_HEAD
    qkv = x @ w_in.t() + b_in
    q, k, v = qkv.split(E, dim=-1)
    q = q.view(Bp, n, N_HEAD, d).transpose(1, 2)
    k = k.view(Bp, n, N_HEAD, d).transpose(1, 2)
    v = v.view(Bp, n, N_HEAD, d).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(Bp, n, E)
    return o @ w_out.t() + b_out


def encoder_layer(sd, prefix, x):
    """Pre-norm TransformerEncoderLayer.forward (CSE_transformer.py:385-416), dropout 0."""
    h = layer_norm(x, sd[prefix + "norm1.norm.weight"], sd[prefix + "norm1.norm.bias"])
    x = x + multihead_attention(sd, prefix + "self_att.", h)
    h = layer_norm(x, sd[prefix + "norm2.norm.weight"], sd[prefix + "norm2.norm.bias"])
    h = torch.relu(h @ sd[prefix + "pos_ffn.ffn.0.weight"].t() + sd[prefix + "pos_ffn.ffn.0.bias"])
    return x + (h @ sd[prefix + "pos_ffn.ffn.3.weight"].t() + sd[prefix + "pos_ffn.ffn.3.bias"])


def transformer_block(sd, prefix, x):
    """SBTransformerBlock_CSE.forward (CSE_transformer.py:90-106) + TransformerEncoder.forward
    (:201-250): x + pe[:, :n], 8 layers, final LayerNorm."""
    x = x + sd[prefix + "pos_enc.pe"][:, : x.shape[1]].to(x.dtype)
    for l in range(N_LAYER):
        x = encoder_layer(sd, f"{prefix}mdl.layers.{l}.", x)
    return layer_norm(x, sd[prefix + "mdl.norm.norm.weight"], sd[prefix + "mdl.norm.norm.bias"])


def dual_block(sd, i, x, ctx):
    """Dual_Computation_Block_CSE.forward (ContSep.py:453-533; ContExt.py:479-557).
    x [B,N,K,S], ctx [B,c,4096] or None -> (out [B,N,K,S], pred_head [B,N] or None)."""
    p = f"masknet.dual_mdl.{i}."
    B, N, K, S = x.shape
    c = 0 if ctx is None else ctx.shape[1]
    intra = x.permute(0, 3, 2, 1).reshape(B * S, K, N)
    if c:
        tok = ctx @ sd[p + "intra_context_mapper.weight"].t() + sd[p + "intra_context_mapper.bias"]
        tok = tok.unsqueeze(1).expand(B, S, c, N).reshape(B * S, c, N)
        intra = torch.cat([tok, intra], dim=1)
    intra = transformer_block(sd, p + "intra_mdl.", intra)[:, c:]
    intra = intra.reshape(B, S, K, N).permute(0, 3, 2, 1)
    intra = group_norm1(intra, sd[p + "intra_norm.weight"], sd[p + "intra_norm.bias"]) + x
    inter = intra.permute(0, 2, 3, 1).reshape(B * K, S, N)
    if c:
        tok = ctx @ sd[p + "inter_context_mapper.weight"].t() + sd[p + "inter_context_mapper.bias"]
        tok = tok.unsqueeze(1).expand(B, K, c, N).reshape(B * K, c, N)
        inter = torch.cat([tok, inter], dim=1)
    inter = transformer_block(sd, p + "inter_mdl.", inter)
    pred_head = inter[:, 0].reshape(B, K, N).mean(1)             # ContSep.py:516-517
    inter = inter[:, c:].reshape(B, K, S, N).permute(0, 3, 1, 2)
    inter = group_norm1(inter, sd[p + "inter_norm.weight"], sd[p + "inter_norm.bias"])
    return inter + intra, pred_head


def masknet(sd, mix_w, ctx, num_spks):
    """Dual_Path_Model_CSE.forward (ContSep.py:205-268). -> (mask [spk,B,N,L], pred_head)."""
    x = group_norm1(mix_w, sd["masknet.norm.weight"], sd["masknet.norm.bias"])
    x = torch.einsum("oi,bil->bol", sd["masknet.conv1d.weight"][:, :, 0], x)
    x, gap = pad_and_segment(x)
    pred_head = None
    for i in range(N_BLOCK):
        x, pred_head = dual_block(sd, i, x, ctx)
    a = sd["masknet.prelu.weight"]
    x = torch.where(x >= 0, x, a * x)
    x = torch.einsum("oi,biks->boks", sd["masknet.conv2d.weight"][:, :, 0, 0], x) \
        + sd["masknet.conv2d.bias"].view(1, -1, 1, 1)
    B, _, K, S = x.shape
    x = x.reshape(B * num_spks, -1, K, S)
    x = overlap_add(x, gap)
    o = torch.tanh(torch.einsum("oi,bil->bol", sd["masknet.output.0.weight"][:, :, 0], x)
                   + sd["masknet.output.0.bias"].view(1, -1, 1))
    g = torch.sigmoid(torch.einsum("oi,bil->bol", sd["masknet.output_gate.0.weight"][:, :, 0], x)
                      + sd["masknet.output_gate.0.bias"].view(1, -1, 1))
    x = torch.einsum("oi,bil->bol", sd["masknet.end_conv1x1.weight"][:, :, 0], o * g)
    _, N, L = x.shape
    x = torch.relu(x.reshape(B, num_spks, N, L))
    return x.transpose(0, 1), pred_head


def decoder(sd, x):
    """speechbrain Decoder = ConvTranspose1d(256,1,16,stride=8,bias=False) (ContSep.py:40,84).
    x [B,N,L] -> [B, 8(L-1)+16]."""
    w = sd["decoder.weight"][:, 0, :]                            # [N,16]
    B, N, L = x.shape
    k = w.shape[1]
    hop = k // 2
    frames = torch.einsum("bnl,nk->blk", x, w)                   # [B,L,16]
    out = x.new_zeros(B, hop * (L - 1) + k)
    for j in range(k // hop):                                    # 2 half-frames
        seg = frames[:, :, j * hop:(j + 1) * hop].reshape(B, L * hop)
        out[:, j * hop: j * hop + L * hop] += seg
    return out


def fix_length(est, T):
    """ContSep.py:90-95."""
    T_est = est.shape[1]
    if T > T_est:
        return torch.nn.functional.pad(est, (0, 0, 0, T - T_est))
    return est[:, :T, :]


def assemble_context(sd, ctx, se, cue="joint"):
    """ContExt.Sepformer.forward eval-mode cue assembly (ContExt.py:96-111)."""
    se = se @ sd["se_embedding.weight"].t() + sd["se_embedding.bias"]
    if cue == "joint":
        return torch.cat([ctx, se], 1)
    if cue == "history":
        return torch.cat([ctx, torch.zeros_like(ctx)], 1)
    if cue == "voice":
        return torch.cat([torch.zeros_like(se), se], 1)
    return ctx                                                    # unknown cue: ctx unchanged


def sepformer_forward(sd, mix, ctx=None, variant="contsep", num_spks=2, se=None, cue="joint"):
    """Top-level forward of the four model flavours.

    sepformer : sepformer.py:42-81                     -> est [B,T,spk]
    contsep   : ContSep.py:53-100                      -> (est [B,T,spk], context_pred)
    context   : ContExt.py:54-129 (add_ctx)            -> est [B,T,1] (mask 0 only)
    hcontext  : same with se/cue assembly (add_se)     -> est [B,T,1]
    """
    dt = sd["encoder.conv1d.weight"].dtype
    mix = mix.to(dt)
    if ctx is not None:
        ctx = ctx.to(dt)
    T = mix.shape[1]
    mix_w = encoder(sd, mix)
    if variant == "sepformer":
        ctx = None
    if variant == "hcontext":
        ctx = assemble_context(sd, ctx, se.to(dt), cue)
    mask, pred_head = masknet(sd, mix_w, ctx, num_spks)
    if variant in ("context", "hcontext"):
        est = decoder(sd, mix_w * mask[0]).unsqueeze(-1)
        return fix_length(est, T)
    est = torch.stack([decoder(sd, mix_w * mask[i]) for i in range(num_spks)], dim=-1)
    est = fix_length(est, T)
    if variant == "sepformer":
        return est
    pred = pred_head @ sd["context_selector.weight"].t() + sd["context_selector.bias"]
    return est, pred


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def cal_si_snr(source, estimate, eps=1e-8):
    """speechbrain cal_si_snr (train_ContSep.py:352,386): inputs [T,B,C]; returns the NEGATIVE
    SI-SNR [1,B,C] with `source` as the projection target.  (Does not mutate its inputs.)"""
    s = source - source.mean(0, keepdim=True)
    e = estimate - estimate.mean(0, keepdim=True)
    dot = (e * s).sum(0, keepdim=True)
    energy = (s * s).sum(0, keepdim=True) + eps
    proj = dot * s / energy
    noise = e - proj
    ratio = (proj * proj).sum(0) / ((noise * noise).sum(0) + eps)
    return -(10 * torch.log10(ratio + eps)).unsqueeze(0)


def pit_si_snr(source, estimate_source):
    """get_si_snr_with_pitwrapper(source [B,T,C], estimate_source [B,T,C]) -> (loss [B], perms).
    Per item: loss_mat[i,j] = cal_si_snr(source=source[:,j], estimate=estimate_source[:,i]);
    loss = min over permutations p of mean_i loss_mat[i, p[i]] (first minimum wins)."""
    losses, perms = [], []
    C = source.shape[-1]
    for pred, target in zip(source, estimate_source):
        mat = torch.empty(C, C, dtype=pred.dtype)
        for i in range(C):
            for j in range(C):
                mat[i, j] = cal_si_snr(pred[:, j].view(-1, 1, 1), target[:, i].view(-1, 1, 1)).squeeze()
        best, best_p = None, None
        for p in permutations(range(C)):
            v = mat[list(range(C)), list(p)].mean()
            if best is None or best > v:
                best, best_p = v, p
        losses.append(best)
        perms.append(best_p)
    return torch.stack(losses), perms


def tm_si_snr(preds, target):
    """torchmetrics ScaleInvariantSignalNoiseRatio (train_ContExt.py:339,367): [B,T] -> [B] dB."""
    eps = torch.finfo(preds.dtype).eps
    t = target - target.mean(-1, keepdim=True)
    p = preds - preds.mean(-1, keepdim=True)
    alpha = ((p * t).sum(-1, keepdim=True) + eps) / ((t * t).sum(-1, keepdim=True) + eps)
    ts = alpha * t
    noise = ts - p
    return 10 * torch.log10(((ts * ts).sum(-1) + eps) / ((noise * noise).sum(-1) + eps))
