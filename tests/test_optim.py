"""Fused optimiser step: clip_grad_norm_ + AdamW(amsgrad) + GradScaler bookkeeping.
CPU: the oracle restatement against stock torch (the reference's own calls, train_ContSep.py:233,402-419) and the
host-side chunk table.  GPU: `cse_optim_step` through `cse_b200.optim.AdamW` against both."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib
from oracle.optim_oracle import AdamWOracle

SHAPES = [(256, 4096), (768, 256), (768,), (256, 1, 16), (1,), (3, 5), (40000,)]
HYPER = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=True)


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) * 0.3 for s in SHAPES]


def _grads(step, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(1000 * seed + step)
    mag = 10.0 if step % 2 == 0 else 0.01  # alternate clipped / unclipped steps
    return [torch.randn(s, generator=g) * mag * scale for s in SHAPES]


def test_oracle_matches_stock_torch_adamw_and_clip_grad_norm():
    ps = [torch.nn.Parameter(p.clone()) for p in _params()]
    opt = torch.optim.AdamW(ps, **HYPER)
    orc = AdamWOracle([p.detach().numpy() for p in ps], **HYPER)
    for step in range(6):
        gs = _grads(step)
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_(ps, max_norm=5.0)
        opt.step()
        onorm = orc.step([g.numpy() for g in gs], max_norm=5.0)
        assert abs(onorm - norm.item()) <= 2e-5 * norm.item()  # torch accumulates the norm in fp32
        for p, q in zip(ps, orc.p):
            np.testing.assert_allclose(q, p.detach().numpy(), rtol=2e-6, atol=1e-7)
    for p, m, v, x in zip(ps, orc.m, orc.v, orc.vmax):
        st = opt.state[p]
        # (the moments are linear / quadratic in the clip coefficient, whose norm torch accumulates in fp32: ~1e-5)
        for a, b in ((m, st["exp_avg"]), (v, st["exp_avg_sq"]), (x, st["max_exp_avg_sq"])):
            b = b.numpy().astype(np.float64)
            assert np.linalg.norm(a - b) <= 6e-5 * np.linalg.norm(b)


def test_oracle_scaler_rule_matches_the_documented_gradscaler_update():
    orc = AdamWOracle([np.zeros(4, np.float32)], init_scale=1024.0, growth_interval=3, **HYPER)
    g = [np.ones(4, np.float32)]
    for _ in range(3):
        orc.step(g)
    assert orc.scale == 2048.0 and orc.tracker == 0 and orc.step_count == 3
    before = orc.p[0].copy()
    orc.step([np.array([1.0, np.inf, 0.0, 0.0], np.float32)])
    assert orc.found_inf and orc.scale == 1024.0 and orc.step_count == 3
    np.testing.assert_array_equal(orc.p[0], before)


def test_chunk_table_is_filled_on_the_host():
    lib = _lib.load()
    numel = (C.c_longlong * 3)(5, 16384, 40000)
    assert lib.cse_optim_chunk_count(3, numel) == 1 + 1 + 3
    assert lib.cse_optim_chunk_count(1, (C.c_longlong * 1)(-1)) == -1
    base = [0x10000000 * (k + 1) for k in range(5)]
    arrs = [(C.c_void_p * 3)(b, b + 0x100000, b + 0x200000) for b in base]
    buf = (C.c_uint8 * (5 * 48))()
    _lib.call("cse_optim_table_fill", 3, numel, *arrs, C.cast(buf, C.c_void_p), 5 * 48)
    rec = np.frombuffer(buf, dtype=np.dtype([("ptr", "<u8", 5), ("n", "<i4"), ("pad", "<i4")]))
    assert rec["n"].tolist() == [5, 16384, 16384, 16384, 40000 - 2 * 16384]
    assert rec["ptr"][3].tolist() == [b + 0x200000 + 4 * 16384 for b in base]
    with pytest.raises(_lib.CseError):
        _lib.call("cse_optim_table_fill", 3, numel, *arrs, C.cast(buf, C.c_void_p), 4 * 48)


def test_cpu_tensors_are_rejected():
    from cse_b200.optim import AdamW
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(_lib.CseError):
        AdamW([p], **HYPER).step()


@pytest.mark.gpu
@pytest.mark.parametrize("amsgrad", [True, False])
def test_fused_step_matches_oracle_and_stock_torch(amsgrad):
    from cse_b200.optim import AdamW
    hyper = dict(HYPER, amsgrad=amsgrad)
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in _params()]
    ref_ps = [torch.nn.Parameter(p.clone().cuda()) for p in _params()]
    opt, ref = AdamW(ps, **hyper), torch.optim.AdamW(ref_ps, **hyper)
    orc = AdamWOracle([p.numpy() for p in _params()], **hyper)
    for step in range(6):
        gs = _grads(step)
        for p, q, g in zip(ps, ref_ps, gs):
            p.grad, q.grad = g.clone().cuda(), g.clone().cuda()
        norm = opt.step(max_norm=5.0, write_back_grads=(step == 1))
        rnorm = torch.nn.utils.clip_grad_norm_(ref_ps, max_norm=5.0)
        ref.step()
        onorm = orc.step([g.numpy() for g in gs], max_norm=5.0)
        assert abs(norm.item() - onorm) <= 2e-6 * onorm and abs(norm.item() - rnorm.item()) <= 2e-5 * onorm
        if step == 1:  # clip_grad_norm_ leaves the clipped gradients in .grad
            for p, q in zip(ps, ref_ps):
                torch.testing.assert_close(p.grad, q.grad, rtol=2e-6, atol=1e-9)
        for p, q, o in zip(ps, ref_ps, orc.p):
            np.testing.assert_allclose(p.detach().cpu().numpy(), o, rtol=3e-6, atol=2e-7)
            torch.testing.assert_close(p.detach(), q.detach(), rtol=3e-6, atol=2e-7)
    assert opt.steps_applied() == 6 and not opt.found_inf
    # checkpoint layout interchangeable with torch.optim.AdamW
    sd = opt.state_dict()
    rsd = ref.state_dict()
    assert sd["state"].keys() == rsd["state"].keys()
    for k in sd["state"]:
        assert set(sd["state"][k]) == set(rsd["state"][k])
        assert float(sd["state"][k]["step"]) == float(rsd["state"][k]["step"]) == 6.0
        a, b = sd["state"][k]["exp_avg_sq"].double(), rsd["state"][k]["exp_avg_sq"].double()
        assert (a - b).norm() <= 6e-5 * b.norm()
    fresh = AdamW(ps, **hyper)
    fresh.load_state_dict(rsd)
    assert fresh.steps_applied() == 6


@pytest.mark.gpu
def test_fused_step_gradscaler_semantics():
    """--fp16 path (train_ContSep.py:397,405-410): scaled loss, unscale, skip + back-off on inf, growth."""
    from cse_b200.optim import AdamW
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in _params()]
    opt = AdamW(ps, init_scale=1024.0, growth_interval=3, **HYPER)
    orc = AdamWOracle([p.numpy() for p in _params()], init_scale=1024.0, growth_interval=3, **HYPER)
    assert opt.scale(torch.ones((), device="cuda")).item() == 1024.0
    for step in range(7):
        gs = _grads(step, scale=orc.scale)
        if step == 4:
            gs[2][5] = float("inf")
        for p, g in zip(ps, gs):
            p.grad = g.clone().cuda()
        norm = opt.step(max_norm=5.0)
        onorm = orc.step([g.numpy() for g in gs], max_norm=5.0)
        assert opt.found_inf == orc.found_inf == (step == 4)
        if step != 4:
            assert abs(norm.item() - onorm) <= 3e-6 * onorm
        assert opt.get_scale() == orc.scale
        for p, o in zip(ps, orc.p):
            np.testing.assert_allclose(p.detach().cpu().numpy(), o, rtol=3e-6, atol=2e-7)
    assert opt.steps_applied() == orc.step_count == 6


@pytest.mark.gpu
def test_fused_step_on_the_real_model_in_three_launches():
    from cse_b200 import synth
    from cse_b200.models.ContExt import Sepformer
    from cse_b200.optim import AdamW
    m = Sepformer(2, add_ctx=True)
    m.add_ctx_pipeline()
    m.load_state_dict(synth.make_state_dict("context", 2, seed=1))
    m = m.cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    for p in m.parameters():
        p.grad = torch.randn(p.shape, generator=g, device="cuda") * 0.05
    ref = [p.detach().clone() for p in m.parameters()]
    ref_ps = [torch.nn.Parameter(r.clone()) for r in ref]
    for q, p in zip(ref_ps, m.parameters()):
        q.grad = p.grad.clone()
    opt = AdamW(m.parameters(), lr=1e-4, weight_decay=1e-6, amsgrad=True)
    n0 = _lib.load().cse_launch_count()
    norm = opt.step(max_norm=5.0)
    assert _lib.load().cse_launch_count() - n0 == 3
    stock = torch.optim.AdamW(ref_ps, lr=1e-4, weight_decay=1e-6, amsgrad=True)
    rnorm = torch.nn.utils.clip_grad_norm_(ref_ps, max_norm=5.0)
    stock.step()
    assert abs(norm.item() - rnorm.item()) <= 2e-5 * rnorm.item()
    for p, q in zip(m.parameters(), ref_ps):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=3e-6, atol=1e-7)
