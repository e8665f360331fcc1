"""GPU: the hand-written backward kernels (csrc/backward.cu, loss.cu) through the C ABI against autograd
over the pinned CPU oracle (float64), fp32 bar 1e-4 relative L2 (BASELINE north_star)."""
import ctypes as C

import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib, backward, losses, synth
from helpers import rel_l2
from oracle import backward_oracle as BO
from oracle import sepformer_oracle as O
from test_backward_oracle import _layer_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _signals(B, T, Cn, seed):
    _, a = synth.make_mixture(B, T, max(Cn, 2), seed=seed)
    _, b = synth.make_mixture(B, T, max(Cn, 2), seed=seed + 1)
    a, b = a[:, :, :Cn].contiguous(), b[:, :, :Cn].contiguous()
    return (0.6 * a + 0.4 * b.flip(-1)).contiguous(), a


# ------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,Cn", [(2, 3000, 2), (1, 32000, 3), (3, 777, 1)])
def test_cal_si_snr_backward(B, T, Cn):
    est, src = _signals(B, T, Cn, 41)
    w = _rand(1, B, Cn, seed=5)                                   # upstream gradient
    s = src.transpose(0, 1).to(DEV).requires_grad_(True)          # [T,B,C] like the reference call
    e = est.transpose(0, 1).to(DEV).requires_grad_(True)
    v = losses.cal_si_snr(s, e)
    (v * w.to(DEV)).sum().backward()
    sr = src.transpose(0, 1).double().requires_grad_(True)
    er = est.transpose(0, 1).double().requires_grad_(True)
    vr = O.cal_si_snr(sr, er)
    (vr * w.double()).sum().backward()
    assert torch.allclose(v.cpu().double(), vr.detach(), atol=2e-4)
    assert rel_l2(s.grad.cpu(), sr.grad) < TOL
    assert rel_l2(e.grad.cpu(), er.grad) < TOL


@pytest.mark.parametrize("B,T,Cn", [(2, 3000, 2), (2, 16000, 3)])
def test_pit_backward_training_call_order(B, T, Cn):
    """train_ContSep.py:391-393 passes (estimate, targets): the model output is the FIRST argument."""
    est, tgt = _signals(B, T, Cn, 43)
    e = est.to(DEV).requires_grad_(True)
    t = tgt.to(DEV).requires_grad_(True)
    loss = losses.get_si_snr_with_pitwrapper(e, t)
    loss.mean().backward()
    er, tr = est.double().requires_grad_(True), tgt.double().requires_grad_(True)
    lr, _ = O.pit_si_snr(er, tr)
    lr.mean().backward()
    assert torch.allclose(loss.detach().cpu().double(), lr.detach(), atol=2e-4)
    assert rel_l2(e.grad.cpu(), er.grad) < TOL
    assert rel_l2(t.grad.cpu(), tr.grad) < TOL
    # only the estimate needs a gradient in training: the target side may be skipped
    e2 = est.to(DEV).requires_grad_(True)
    losses.get_si_snr_with_pitwrapper(e2, tgt.to(DEV)).mean().backward()
    assert torch.equal(e2.grad, e.grad)


@pytest.mark.parametrize("B,T", [(2, 4000), (1, 128000)])
def test_tm_si_snr_backward(B, T):
    """train_ContExt.py:366-367: loss = -sisnr(est[:, :, 0], gt)."""
    est, tgt = _signals(B, T, 1, 47)
    p = est[:, :, 0].contiguous().to(DEV).requires_grad_(True)
    t = tgt[:, :, 0].contiguous().to(DEV)
    loss = -losses.ScaleInvariantSignalNoiseRatio()(p, t)
    loss.backward()
    pr = est[:, :, 0].float().requires_grad_(True)
    lr = -O.tm_si_snr(pr, tgt[:, :, 0].float()).mean()
    lr.backward()
    assert abs(loss.item() - lr.item()) < 2e-4
    assert rel_l2(p.grad.cpu(), pr.grad) < TOL


# ------------------------------------------------------------------------------------------
# layer pieces
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [1, 77, 2000])
def test_layernorm_backward(M):
    x = _rand(M, 256, seed=5) * 3 + 0.5
    g, dy = 1 + 0.1 * _rand(256, seed=6), _rand(M, 256, seed=7)
    dx_ref, dg_ref, db_ref = BO.manual_layernorm_bwd(x.double(), g.double(), dy.double())
    xd, gd, dyd = x.to(DEV), g.to(DEV), dy.to(DEV)
    base = _rand(M, 256, seed=8)
    dx = base.to(DEV)
    dg, db = torch.zeros(256, device=DEV), torch.ones(256, device=DEV)      # db starts at 1: accumulation
    _lib.call("cse_layernorm_bwd", _lib.ptr(xd), _lib.ptr(gd), _lib.ptr(dyd), M, 1e-6, 1, _lib.ptr(dx),
              _lib.ptr(dg), _lib.ptr(db), _st())
    assert rel_l2(dx.cpu().double() - base.double(), dx_ref) < TOL
    assert rel_l2(dg.cpu(), dg_ref) < TOL
    assert rel_l2(db.cpu() - 1, db_ref) < TOL
    dx2 = torch.full((M, 256), 7.0, device=DEV)
    _lib.call("cse_layernorm_bwd", _lib.ptr(xd), _lib.ptr(gd), _lib.ptr(dyd), M, 1e-6, 0, _lib.ptr(dx2),
              None, None, _st())
    assert rel_l2(dx2.cpu(), dx_ref) < TOL


@pytest.mark.parametrize("M,N,K", [(300, 768, 256), (77, 256, 1024), (5000, 1024, 256), (1, 256, 256)])
def test_linear_backward(M, N, K):
    a, w, dc = _rand(M, K, seed=1), _rand(N, K, seed=2) / K ** 0.5, _rand(M, N, seed=3)
    da_ref, dw_ref, db_ref = BO.manual_linear_bwd(a.double(), w.double(), dc.double())
    ad, wd, dcd = a.to(DEV), w.to(DEV), dc.to(DEV)
    da = torch.empty(M, K, device=DEV)
    dw, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    wt = torch.empty(N * K, device=DEV)
    args = (_lib.ptr(ad), K, _lib.ptr(wd), _lib.ptr(dcd), N, M, N, K, _lib.ptr(da), K, _lib.ptr(dw), _lib.ptr(db),
            _lib.ptr(wt), _st())
    _lib.call("cse_linear_bwd", *args)
    assert rel_l2(da.cpu(), da_ref) < TOL
    assert rel_l2(dw.cpu(), dw_ref) < TOL
    assert rel_l2(db.cpu(), db_ref) < TOL
    _lib.call("cse_linear_bwd", *args)                                      # parameter gradients accumulate
    assert rel_l2(dw.cpu(), 2 * dw_ref) < TOL
    assert rel_l2(db.cpu(), 2 * db_ref) < TOL
    with pytest.raises(_lib.CseError):
        _lib.call("cse_linear_bwd", _lib.ptr(ad), K, _lib.ptr(wd), _lib.ptr(dcd), N, M, N + 1, K, None, K,
                  _lib.ptr(dw), None, None, _st())


@pytest.mark.parametrize("nseq,n", [(3, 35), (2, 251), (1, 300), (5, 1)])
def test_attention_backward(nseq, n):
    qkv = _rand(nseq, n, 768, seed=9)
    do = _rand(nseq, n, 256, seed=10)
    o_ref, dqkv_ref = BO.manual_attention_bwd(qkv.double(), do.double())
    qd, dod = qkv.to(DEV), do.to(DEV)
    out = torch.empty(nseq * n, 256, device=DEV)
    _lib.call("cse_attention_fwd", _lib.ptr(qd), nseq, n, _lib.FP32, _lib.ptr(out), _st())
    assert rel_l2(out.cpu().view(nseq, n, 256), o_ref) < 1e-5
    dqkv = torch.empty(nseq, n, 768, device=DEV)
    _lib.call("cse_attention_bwd", _lib.ptr(qd), _lib.ptr(out), _lib.ptr(dod), nseq, n, _lib.ptr(dqkv), _st())
    for name, sl in (("dq", slice(0, 256)), ("dk", slice(256, 512)), ("dv", slice(512, 768))):
        got, ref = dqkv.cpu()[..., sl].double(), dqkv_ref[..., sl]
        # n == 1: softmax of a single score is constant, dq = dk = 0 exactly (the reference holds rounding noise)
        assert (got - ref).norm() <= TOL * max(ref.norm().item(), 1e-3), name
    with pytest.raises(_lib.CseError):
        _lib.call("cse_attention_bwd", _lib.ptr(qd), _lib.ptr(out), _lib.ptr(dod), 1, 1000, _lib.ptr(dqkv), _st())


# ------------------------------------------------------------------------------------------
# one whole transformer layer, forward + backward, as an autograd node
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nseq,n", [(4, 36), (2, 251), (1, 1)])
def test_layer_forward_backward(nseq, n):
    p64 = _layer_params(51)
    x = _rand(nseq, n, 256, seed=52)
    dy = _rand(nseq, n, 256, seed=53)
    y_ref, dx_ref, g_ref = BO.autograd_layer(p64, x.double(), dy.double())
    params = {k: v.float().to(DEV).requires_grad_(True) for k, v in p64.items()}
    R = x.reshape(nseq * n, 256).to(DEV).requires_grad_(True)
    y = backward.transformer_layer(params, R, nseq, n)
    assert rel_l2(y.detach().cpu().view(nseq, n, 256), y_ref) < 1e-5
    y.backward(dy.reshape(nseq * n, 256).to(DEV))
    assert rel_l2(R.grad.cpu().view(nseq, n, 256), dx_ref) < TOL
    for k in BO.LAYER_KEYS:
        assert rel_l2(params[k].grad.cpu(), g_ref[k]) < TOL, k
    # the layer input is not modified by either pass
    assert torch.equal(R.detach().cpu(), x.reshape(nseq * n, 256))


def test_layer_forward_bf16_entry_matches_fp32_entry():
    """cse_layer_fwd in performance mode = the launch sequence of the bench path, one layer."""
    nseq, n = 3, 251
    p64 = _layer_params(61)
    params = {k: v.float().to(DEV) for k, v in p64.items()}
    x = _rand(nseq * n, 256, seed=62).to(DEV)
    y32 = backward.layer_forward(params, x, nseq, n)
    lp = backward._layer_struct(params, _lib.LayerParams)
    packs = {}
    for field in ("in_proj_w", "out_proj_w", "ffn1_w", "ffn2_w"):
        key = dict(backward.LAYER_KEYS)[field]
        packs[field] = params[key].to(torch.bfloat16).contiguous()
        setattr(lp, field + "_bf16", C.c_void_p(packs[field].data_ptr()))
    y16 = x.clone()
    nbytes = _lib.load().cse_layer_workspace_bytes(nseq, n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.call("cse_layer_fwd", C.byref(lp), _lib.ptr(y16), nseq, n, _lib.BF16, _lib.ptr(ws), nbytes, _st())
    assert rel_l2(y16.cpu(), y32.cpu()) < 1e-2
    with pytest.raises(_lib.CseError):
        _lib.call("cse_layer_fwd", C.byref(lp), _lib.ptr(y16), nseq, n, _lib.BF16, _lib.ptr(ws), 16, _st())
