"""CPU restatement of the GRADIENTS of the hot path.  TEST INFRASTRUCTURE ONLY.

The reference has no backward code of its own: training calls `loss.backward()`
(train_ContSep.py:402-419, train_ContExt.py:372-389) and autograd differentiates the forward.
So the pin for every gradient is autograd over the forward — the live reference modules where
/root/reference exists, `oracle/sepformer_oracle.py` (itself pinned against the reference) elsewhere.
`autograd_*` below return those reference gradients; `manual_*` restate, in closed form, exactly
the formulas the CUDA backward kernels implement (csrc/backward.cu, csrc/loss.cu), so that the
derivations are checked on CPU (tests/test_backward_oracle.py) before any GPU time is spent.
Parity status: PINNED through autograd of the pinned forward.
"""
import math

import torch

from . import sepformer_oracle as O

LAYER_KEYS = ("self_att.att.in_proj_weight", "self_att.att.in_proj_bias", "self_att.att.out_proj.weight",
              "self_att.att.out_proj.bias", "pos_ffn.ffn.0.weight", "pos_ffn.ffn.0.bias",
              "pos_ffn.ffn.3.weight", "pos_ffn.ffn.3.bias", "norm1.norm.weight", "norm1.norm.bias",
              "norm2.norm.weight", "norm2.norm.bias")


# --------------------------------------------------------------------------------------
# reference gradients: autograd over the pinned forward
# --------------------------------------------------------------------------------------
def autograd_layer(params, x, dy):
    """params: {key relative to the layer prefix -> tensor}; x, dy [B',n,256].
    Returns (y, dx, {key -> grad}) of encoder_layer (CSE_transformer.py:385-416)."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    xi = x.detach().clone().requires_grad_(True)
    y = O.encoder_layer(p, "", xi)
    y.backward(dy)
    return y.detach(), xi.grad, {k: v.grad for k, v in p.items()}


def autograd_loss(kind, a, b):
    """Gradients of sum(loss) for the three SI-SNR flavours with the argument order of
    cse_b200.losses: cal_si_snr(source [T,B,C], estimate), pit(source [B,T,C], estimate_source),
    tm(preds [B,T], target).  Returns (value, d/da, d/db)."""
    a = a.detach().clone().requires_grad_(True)
    b = b.detach().clone().requires_grad_(True)
    if kind == "cal_si_snr":
        v = O.cal_si_snr(a, b)
    elif kind == "pit":
        v, _ = O.pit_si_snr(a, b)
    else:
        v = O.tm_si_snr(a, b)
    v.sum().backward()
    return v.detach(), a.grad, b.grad


# --------------------------------------------------------------------------------------
# closed forms the kernels implement
# --------------------------------------------------------------------------------------
def manual_layernorm_bwd(x, g, dy, eps=1e-6):
    """layernorm_bwd_kernel: returns (dx, dg, db)."""
    mean = x.mean(-1, keepdim=True)
    xc = x - mean
    rstd = torch.rsqrt((xc * xc).mean(-1, keepdim=True) + eps)
    xh = xc * rstd
    dh = dy * g
    dx = rstd * (dh - dh.mean(-1, keepdim=True) - xh * (dh * xh).mean(-1, keepdim=True))
    red = tuple(range(x.dim() - 1))
    return dx, (dy * xh).sum(red), dy.sum(red)


def manual_attention_bwd(qkv, d_out, heads=8):
    """attention_bwd_kernel: qkv [B',n,768] (q|k|v column blocks), d_out [B',n,256] -> (out, d_qkv)."""
    Bp, n, E3 = qkv.shape
    E = E3 // 3
    d = E // heads
    scale = 1.0 / math.sqrt(d)
    q, k, v = (t.view(Bp, n, heads, d).transpose(1, 2) for t in qkv.split(E, dim=-1))
    go = d_out.view(Bp, n, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * scale
    lse = torch.logsumexp(s, dim=-1, keepdim=True)
    p = torch.exp(s - lse)
    o = p @ v
    D = (go * o).sum(-1, keepdim=True)
    dv = p.transpose(-1, -2) @ go
    ds = p * (go @ v.transpose(-1, -2) - D)
    dq = (ds @ k) * scale
    dk = (ds.transpose(-1, -2) @ q) * scale
    back = lambda t: t.transpose(1, 2).reshape(Bp, n, E)
    return back(o), torch.cat([back(dq), back(dk), back(dv)], dim=-1)


def manual_linear_bwd(a, w, dc):
    """cse_linear_bwd for c = a w^T + b: (da, dw, db)."""
    a2, dc2 = a.reshape(-1, a.shape[-1]), dc.reshape(-1, dc.shape[-1])
    return dc @ w, dc2.t() @ a2, dc2.sum(0)


def manual_layer_bwd(params, x, dy, rnd=None):
    """The launch sequence of cse_layer_bwd (csrc/backward_abi.cu): recompute, then walk back.
    Returns (dx, {key -> grad}).

    rnd: optional rounding applied wherever the performance-mode composition (cse_layer_bwd_bf16,
    csrc/backward_tc.cu) holds a bf16 value — stored activations (norm outputs, qkv, attention output, FFN
    hidden), GEMM weight operands and the incoming gradient of every dgrad / wgrad; residual stream, LayerNorm and
    attention gradients stay in full precision.  None = the fp32 parity path."""
    P = params
    r = rnd if rnd is not None else (lambda t: t)
    g = {}

    def lin_bwd(a, w, dc):                      # operands as the tensor-core GEMMs see them
        return manual_linear_bwd(r(a), r(w), r(dc))

    h1 = r(O.layer_norm(x, P["norm1.norm.weight"], P["norm1.norm.bias"]))
    qkv = r(h1 @ r(P["self_att.att.in_proj_weight"]).t() + P["self_att.att.in_proj_bias"])
    ao, _ = manual_attention_bwd(qkv, torch.zeros_like(x))
    ao = r(ao)
    rmid = x + ao @ r(P["self_att.att.out_proj.weight"]).t() + P["self_att.att.out_proj.bias"]
    h2 = r(O.layer_norm(rmid, P["norm2.norm.weight"], P["norm2.norm.bias"]))
    f1 = r(torch.relu(h2 @ r(P["pos_ffn.ffn.0.weight"]).t() + P["pos_ffn.ffn.0.bias"]))
    dR = dy.clone()
    df1, g["pos_ffn.ffn.3.weight"], g["pos_ffn.ffn.3.bias"] = lin_bwd(f1, P["pos_ffn.ffn.3.weight"], dR)
    g["pos_ffn.ffn.3.bias"] = dR.reshape(-1, dR.shape[-1]).sum(0)           # bias gradients: fp32 column sums
    df1 = df1 * (f1 > 0)
    dh2, g["pos_ffn.ffn.0.weight"], _ = lin_bwd(h2, P["pos_ffn.ffn.0.weight"], df1)
    g["pos_ffn.ffn.0.bias"] = df1.reshape(-1, df1.shape[-1]).sum(0)
    dx2, g["norm2.norm.weight"], g["norm2.norm.bias"] = manual_layernorm_bwd(rmid, P["norm2.norm.weight"], dh2)
    dR = dR + dx2
    dao, g["self_att.att.out_proj.weight"], _ = lin_bwd(ao, P["self_att.att.out_proj.weight"], dR)
    g["self_att.att.out_proj.bias"] = dR.reshape(-1, dR.shape[-1]).sum(0)
    _, dqkv = manual_attention_bwd(qkv, dao)
    dh1, g["self_att.att.in_proj_weight"], _ = lin_bwd(h1, P["self_att.att.in_proj_weight"], dqkv)
    g["self_att.att.in_proj_bias"] = dqkv.reshape(-1, dqkv.shape[-1]).sum(0)
    dx1, g["norm1.norm.weight"], g["norm1.norm.bias"] = manual_layernorm_bwd(x, P["norm1.norm.weight"], dh1)
    return dR + dx1, g


def _centred_stats(s, e):
    sc = s - s.mean()
    ec = e - e.mean()
    return sc, ec, (sc * sc).sum(), (ec * ec).sum(), (sc * ec).sum()


def manual_sb_pair_grad(s, e, eps=1e-8):
    """sb_neg_si_snr_grad (csrc/loss.cu): gradient of cal_si_snr(source=s, estimate=e) for 1-D
    signals, returned as (d/ds, d/de)."""
    sc, ec, ss, ee, dot = _centred_stats(s.double(), e.double())
    energy = ss + eps
    proj2 = dot * dot * ss / energy ** 2
    noise2 = (ee - 2 * dot * dot / energy + proj2).clamp_min(0)
    ratio = proj2 / (noise2 + eps)
    dl_dr = -(10 / math.log(10)) / (ratio + eps)
    dr_dp = 1 / (noise2 + eps)
    dr_dn = -proj2 / (noise2 + eps) ** 2
    dp_dd = 2 * dot * ss / energy ** 2
    dp_dss = dot * dot * (energy - 2 * ss) / energy ** 3
    dn_dd = -4 * dot / energy + dp_dd
    dn_dss = 2 * dot * dot / energy ** 2 + dp_dss
    g_d = dl_dr * (dr_dp * dp_dd + dr_dn * dn_dd)
    g_ss = dl_dr * (dr_dp * dp_dss + dr_dn * dn_dss)
    g_ee = dl_dr * dr_dn
    return 2 * g_ss * sc + g_d * ec, 2 * g_ee * ec + g_d * sc


def manual_pit_grad(source, estimate_source, perms):
    """si_snr_bwd_kernel mode 1: gradient of sum_b pit loss, given the forward's permutations
    (estimate_source column i is paired with source column perms[b][i], weight 1/C)."""
    ds, de = torch.zeros_like(source, dtype=torch.float64), torch.zeros_like(source, dtype=torch.float64)
    C = source.shape[-1]
    for b, perm in enumerate(perms):
        for i, j in enumerate(perm):
            gs, ge = manual_sb_pair_grad(source[b, :, j], estimate_source[b, :, i])
            ds[b, :, j] += gs / C
            de[b, :, i] += ge / C
    return ds, de


def manual_tm_grad(preds, target):
    """tm_si_snr_bwd_kernel: gradient of sum_b torchmetrics SI-SNR(preds[b], target[b])."""
    eps = torch.finfo(torch.float32).eps
    dps, dts = [], []
    for p, t in zip(preds.double(), target.double()):
        tc, pc, tt, pp, pt = _centred_stats(t, p)
        alpha = (pt + eps) / (tt + eps)
        ts2 = alpha * alpha * tt
        noise2 = (ts2 - 2 * alpha * pt + pp).clamp_min(0)
        k10 = 10 / math.log(10)
        g_ts2, g_n = k10 / (ts2 + eps), -k10 / (noise2 + eps)
        g_alpha = g_ts2 * 2 * alpha * tt + g_n * (2 * alpha * tt - 2 * pt)
        g_pt = g_alpha / (tt + eps) - 2 * alpha * g_n
        g_tt = -g_alpha * alpha / (tt + eps) + (g_ts2 + g_n) * alpha * alpha
        dps.append(g_pt * tc + 2 * g_n * pc)
        dts.append(g_pt * pc + 2 * g_tt * tc)
    return torch.stack(dps), torch.stack(dts)
