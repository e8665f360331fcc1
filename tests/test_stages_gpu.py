"""GPU: every per-stage C-ABI entry point against the CPU oracle on seeded inputs (fp32 mode
bit-for-tolerance, bf16 mode within bf16 rounding)."""
import ctypes as C
import math

import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib, synth
from cse_b200._lib import BF16, FP32
from helpers import rel_l2
from oracle import sepformer_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_KEEP = []


def _p(t):
    """Pointer of a tensor that is kept alive until the test ends (a temporary `.to(DEV)` would be
    freed — and its memory reused by the next temporary — before the kernel runs)."""
    _KEEP.append(t)
    return _lib.ptr(t)


@pytest.fixture(autouse=True)
def _release_kept_tensors():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("B,T", [(1, 16), (2, 4003), (3, 32000)])
@pytest.mark.parametrize("prec", [FP32, BF16])
def test_encoder_and_groupnorm_stats(B, T, prec):
    w = _rand(256, 1, 16, seed=1) / 4
    mix = _rand(B, T, seed=2) * 0.3
    ref = O.encoder({"encoder.conv1d.weight": w}, mix).transpose(1, 2).contiguous()     # [B,L,256]
    L = ref.shape[1]
    adt = torch.bfloat16 if prec == BF16 else torch.float32
    out = torch.empty(B, L, 256, dtype=adt, device=DEV)
    part = torch.zeros(B, (L + 63) // 64, 2, device=DEV)
    n_parts = C.c_int(0)
    _lib.call("cse_encoder_fwd", _p(mix.to(DEV)), _p(w.to(DEV)), B, T, prec, _p(out), _p(part),
              C.byref(n_parts), _st())
    assert n_parts.value == part.shape[1]
    tol = 1e-6 if prec == FP32 else 4e-3
    assert rel_l2(out.float().cpu(), ref) < tol
    stat = torch.empty(B, 2, device=DEV)
    _lib.call("cse_gn_finalize", _p(part), B, n_parts.value, float(L * 256), 1e-8, _p(stat), _st())
    x = out.float().cpu().reshape(B, -1)
    assert torch.allclose(stat[:, 0].cpu(), x.mean(1), rtol=1e-4, atol=1e-6)
    assert torch.allclose(stat[:, 1].cpu(), 1 / torch.sqrt(x.var(1, unbiased=False) + 1e-8), rtol=1e-4)
    # apply
    g, b = 1 + 0.1 * _rand(256, seed=3), 0.1 * _rand(256, seed=4)
    y = torch.empty_like(out)
    _lib.call("cse_gn_apply", _p(out), _p(stat), _p(g.to(DEV)), _p(b.to(DEV)), B, L, prec, _p(y), _st())
    ref_y = O.group_norm1(out.float().cpu().transpose(1, 2), g, b).transpose(1, 2)
    assert rel_l2(y.float().cpu(), ref_y) < (1e-5 if prec == FP32 else 4e-3)


@pytest.mark.parametrize("M", [1, 77, 1000])
@pytest.mark.parametrize("prec", [FP32, BF16])
def test_layernorm(M, prec):
    x = _rand(M, 256, seed=5) * 3 + 0.5
    g, b = 1 + 0.1 * _rand(256, seed=6), 0.1 * _rand(256, seed=7)
    adt = torch.bfloat16 if prec == BF16 else torch.float32
    out = torch.empty(M, 256, dtype=adt, device=DEV)
    _lib.call("cse_layernorm_fwd", _p(x.to(DEV)), _p(g.to(DEV)), _p(b.to(DEV)), M, 1e-6, prec, _p(out), _st())
    ref = O.layer_norm(x, g, b)
    assert rel_l2(out.float().cpu(), ref) < (2e-6 if prec == FP32 else 4e-3)


@pytest.mark.parametrize("M,N,K", [(1, 128, 16), (130, 256, 256), (1000, 768, 256), (517, 256, 1024), (300, 1024, 256)])
def test_linear_fp32(M, N, K):
    A, W, bias = _rand(M, K, seed=8), _rand(N, K, seed=9) / math.sqrt(K), _rand(N, seed=10)
    res = _rand(M, N, seed=11)
    Ad, Wd, bd = A.to(DEV), W.to(DEV), bias.to(DEV)
    out = torch.empty(M, N, device=DEV)
    _lib.call("cse_linear", _p(Ad), K, _p(Wd), _p(bd), 1.0, None, _p(out), N, M, N, K, 0, 1, FP32, _st())
    assert rel_l2(out.cpu(), A.double() @ W.double().t() + bias.double()) < 2e-6
    _lib.call("cse_linear", _p(Ad), K, _p(Wd), _p(bd), 2.0, None, _p(out), N, M, N, K, 1, 1, FP32, _st())
    assert rel_l2(out.cpu(), torch.relu(A.double() @ W.double().t() + 2 * bias.double())) < 2e-6
    r = res.to(DEV).clone()                                    # in-place residual (C aliases residual)
    _lib.call("cse_linear", _p(Ad), K, _p(Wd), None, 0.0, _p(r), _p(r), N, M, N, K, 0, 1, FP32, _st())
    assert rel_l2(r.cpu(), A.double() @ W.double().t() + res.double()) < 2e-6


def _attention_ref(qkv, nseq, n):
    q, k, v = qkv.double().view(nseq, n, 3, 8, 32).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(32), -1)
    return (att @ v).permute(0, 2, 1, 3).reshape(nseq * n, 256)


@pytest.mark.parametrize("nseq,n", [(3, 251), (5, 35), (2, 132), (4, 1), (2, 17), (1, 260), (2, 64),
                                    (7, 35), (300, 35), (40, 252), (1, 128), (3, 129), (2, 256), (5, 100),
                                    (544, 251), (4000, 35), (130, 252), (250, 132), (1001, 67)])
@pytest.mark.parametrize("prec", [FP32, BF16, "bf16-mma", "bf16-tc", "bf16-tc4"])
def test_attention(nseq, n, prec):
    """fp32 SIMT kernel, the automatic bf16 choice, and each bf16 kernel forced: tcgen05 (n <= 256;
    packed tiles for n <= 128, split tiles above) and mma.sync online softmax (any n), all against
    the float64 softmax(q k^T / sqrt(d)) v."""
    # tc = v1 kernel (the default); tc4 = v4 (per-buffer issuers, 16 softmax warps, P in TMEM; opt-in, CSE_ATTN_VER=4)
    mode = {"bf16-mma": 1, "bf16-tc": 2, "bf16-tc4": 4}.get(prec, 0)
    if mode >= 2 and n > 256:
        pytest.skip("tcgen05 attention handles n <= 256")
    prec = BF16 if mode else prec
    qkv = _rand(nseq * n, 768, seed=12) * 1.5
    adt = torch.bfloat16 if prec == BF16 else torch.float32
    q = qkv.to(adt)
    out = torch.zeros(nseq * n, 256, dtype=adt, device=DEV)
    _lib.load().cse_debug_force_mma_attention(mode)
    try:
        _lib.call("cse_attention_fwd", _p(q.to(DEV)), nseq, n, prec, _p(out), _st())
        torch.cuda.synchronize()
    finally:
        _lib.load().cse_debug_force_mma_attention(0)
    ref = _attention_ref(q.float(), nseq, n)
    assert rel_l2(out.float().cpu(), ref) < (3e-6 if prec == FP32 else 1e-2)


@pytest.mark.parametrize("B,L", [(2, 124), (1, 375), (2, 499), (1, 1)])
def test_segment_build_and_overlap_add(B, L):
    sh = _lib.path_shape(B, 8 * (L - 1) + 16, 1, 2)
    S = sh.S
    x0 = _rand(B, L, 256, seed=13)
    X = torch.empty(B, S, 250, 256, device=DEV)
    _lib.call("cse_segment", _p(x0.to(DEV)), B, L, S, _p(X), _st())
    seg, gap = O.pad_and_segment(x0.transpose(1, 2))                     # [B,N,K,S]
    assert gap == sh.gap and seg.shape[3] == S
    assert torch.equal(X.cpu(), seg.permute(0, 3, 2, 1).contiguous())
    # residual-stream builder, both layouts, c = 2
    c = 2
    pe = synth.positional_table()[0]
    ctok = _rand(B, c, 256, seed=14)
    for inter in (0, 1):
        n = (S if inter else 250) + c
        nseq = B * (250 if inter else S)
        R = torch.empty(nseq, n, 256, device=DEV)
        _lib.call("cse_build_sequences", _p(X), _p(ctok.to(DEV)), _p(pe.to(DEV)), B, S, c, inter, _p(R), _st())
        Xc = X.cpu()
        body = Xc.permute(0, 2, 1, 3).reshape(nseq, S, 256) if inter else Xc.reshape(nseq, 250, 256)
        per_b = 250 if inter else S
        tok = ctok.unsqueeze(1).expand(B, per_b, c, 256).reshape(nseq, c, 256)
        ref = torch.cat([tok, body], 1) + pe[:n]
        assert torch.allclose(R.cpu(), ref, atol=1e-6)
    # PReLU + overlap-add (commuted form) against prelu -> _over_add
    a = torch.tensor([0.2])
    Xr = _rand(B, S, 250, 256, seed=15)
    for prec in (FP32, BF16):
        adt = torch.bfloat16 if prec == BF16 else torch.float32
        U = torch.empty(B, L, 256, dtype=adt, device=DEV)
        _lib.call("cse_prelu_overlap_add", _p(Xr.to(DEV)), _p(a.to(DEV)), B, S, L, prec, _p(U), _st())
        y = Xr.permute(0, 3, 2, 1)
        y = torch.where(y >= 0, y, a * y)
        ref = O.overlap_add(y, gap).transpose(1, 2)
        assert rel_l2(U.float().cpu(), ref) < (1e-6 if prec == FP32 else 4e-3)


def test_context_map():
    ctx = _rand(6, 4096, seed=16)
    w, b = _rand(256, 4096, seed=17) / 64, _rand(256, seed=18)
    out = torch.empty(6, 256, device=DEV)
    _lib.call("cse_context_map", _p(ctx.to(DEV)), _p(w.to(DEV)), _p(b.to(DEV)), 6, 4096, _p(out), _st())
    assert rel_l2(out.cpu(), ctx.double() @ w.double().t() + b.double()) < 2e-6


@pytest.mark.parametrize("inter", [0, 1])
@pytest.mark.parametrize("c", [0, 2])
def test_stack_finish_and_pred_head(inter, c):
    B, S = 2, 6
    n = (S if inter else 250) + c
    nseq = B * (250 if inter else S)
    R = _rand(nseq, n, 256, seed=19) * 2 + 0.3
    lg, lb = 1 + 0.1 * _rand(256, seed=20), 0.1 * _rand(256, seed=21)
    gg, gb = 1 + 0.1 * _rand(256, seed=22), 0.1 * _rand(256, seed=23)
    skip = _rand(B, S, 250, 256, seed=24)
    out = torch.empty(B, S, 250, 256, device=DEV)
    part = torch.empty(B, 64, 2, device=DEV)
    stat = torch.empty(B, 2, device=DEV)
    _lib.call("cse_stack_finish", _p(R.to(DEV)), _p(lg.to(DEV)), _p(lb.to(DEV)), _p(gg.to(DEV)), _p(gb.to(DEV)),
              _p(skip.to(DEV)), B, S, c, inter, _p(out), _p(part), _p(stat), _st())
    y = O.layer_norm(R.double(), lg.double(), lb.double())[:, c:]
    if inter:
        y = y.reshape(B, 250, S, 256).permute(0, 3, 1, 2)          # [B,N,K,S]
    else:
        y = y.reshape(B, S, 250, 256).permute(0, 3, 2, 1)
    ref = O.group_norm1(y, gg.double(), gb.double()) + skip.double().permute(0, 3, 2, 1)
    assert rel_l2(out.cpu(), ref.permute(0, 3, 2, 1)) < 5e-6
    if inter:
        ph = torch.empty(B, 256, device=DEV)
        _lib.call("cse_pred_head", _p(R.to(DEV)), _p(lg.to(DEV)), _p(lb.to(DEV)), B, S, c, _p(ph), _st())
        ref_ph = O.layer_norm(R.double(), lg.double(), lb.double())[:, 0].reshape(B, 250, 256).mean(1)
        assert rel_l2(ph.cpu(), ref_ph) < 5e-6


@pytest.mark.parametrize("prec", [FP32, BF16])
def test_gate_and_mask_decode(prec):
    B, L, n_masks = 2, 300, 2
    T = 8 * (L - 1) + 16 + 5                                   # exercises the zero-pad tail
    adt = torch.bfloat16 if prec == BF16 else torch.float32
    o, g = _rand(B * L * n_masks, 256, seed=25).to(adt), _rand(B * L * n_masks, 256, seed=26).to(adt)
    out = torch.empty(B * L * n_masks, 256, dtype=adt, device=DEV)
    _lib.call("cse_gate", _p(o.to(DEV)), _p(g.to(DEV)), o.numel(), prec, _p(out), _st())
    ref = torch.tanh(o.double()) * torch.sigmoid(g.double())
    assert rel_l2(out.float().cpu(), ref) < (2e-6 if prec == FP32 else 4e-3)
    mp = _rand(B * L * n_masks, 256, seed=27).to(adt)
    E = torch.relu(_rand(B, L, 256, seed=28)).to(adt)
    wd = _rand(256, 1, 16, seed=29) / 16
    frames = torch.empty(B * L * n_masks, 16, device=DEV)
    est = torch.empty(B, T, n_masks, device=DEV)
    _lib.call("cse_mask_decode", _p(mp.to(DEV)), _p(E.to(DEV)), _p(wd.to(DEV)), B, L, T, n_masks, prec,
              _p(frames), _p(est), _st())
    mask = torch.relu(mp.double()).view(B, L, n_masks, 256)
    sd = {"decoder.weight": wd.double()}
    cols = [O.decoder(sd, (E.double() * mask[:, :, s]).transpose(1, 2)) for s in range(n_masks)]
    ref = O.fix_length(torch.stack(cols, -1), T)
    # bf16 mode: the decoder contraction runs on the tensor cores with mask * mix_w and the filter rounded to bf16,
    # as the reference's ConvTranspose1d does under autocast (2^-9 per operand over 256 channels)
    tol = 5e-6 if prec == FP32 else 4e-3
    assert rel_l2(est.cpu(), ref) < tol
    assert torch.all(est[:, 8 * (L - 1) + 16:].cpu() == 0)
    # trim branch
    T2 = 8 * (L - 1) + 16 - 3
    est2 = torch.empty(B, T2, n_masks, device=DEV)
    _lib.call("cse_mask_decode", _p(mp.to(DEV)), _p(E.to(DEV)), _p(wd.to(DEV)), B, L, T2, n_masks, prec,
              _p(frames), _p(est2), _st())
    assert rel_l2(est2.cpu(), ref[:, :T2]) < tol
