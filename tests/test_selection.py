"""ContSep selection tail: oracle vs the reference-expression fixtures on CPU, CUDA kernels vs both
on the GPU (train_ContSep.py:386-388; test.py:234-239,248-255)."""
import pytest
import torch

import cse_b200  # noqa: F401
from cases import SELECTION_CASES, selection_inputs
from helpers import load_golden
from oracle import selection_oracle as SO

DEV = "cuda:0"


@pytest.mark.parametrize("name", list(SELECTION_CASES))
def test_selection_oracle_matches_reference_expressions(name):
    case = SELECTION_CASES[name]
    gt, ns, est, ctx_pred = selection_inputs(case)
    gold = load_golden(name)
    logits = ctx_pred.clone().requires_grad_(True)
    loss, index, sisnrs = SO.selection_loss(logits, est, gt, case["ce"])
    loss.backward()
    assert torch.equal(index, gold["context_index"])
    assert torch.allclose(sisnrs, gold["sisnrs"], atol=1e-4)
    assert torch.allclose(loss.detach(), gold["ctx_loss"], atol=1e-6)
    assert torch.allclose(logits.grad, gold["dlogits"], atol=1e-6)
    picked, pick = SO.select_stream(est, ctx_pred, case["ce"])
    assert torch.equal(pick, gold["pick"]) and torch.equal(picked[:, :64], gold["picked_head"])
    acc, vals = SO.selection_accuracy(picked, torch.cat([gt.unsqueeze(-1), ns], -1))
    assert torch.equal(acc, gold["acc"]) and torch.allclose(vals, gold["acc_sisnrs"], atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SELECTION_CASES))
def test_selection_kernels_match_reference_expressions(name):
    from cse_b200 import selection
    case = SELECTION_CASES[name]
    gt, ns, est, ctx_pred = selection_inputs(case)
    gold = load_golden(name)
    logits = ctx_pred.to(DEV).requires_grad_(True)
    loss, index, sisnrs = selection.selection_loss(logits, est.to(DEV), gt.to(DEV), case["ce"])
    (3.0 * loss).backward()
    assert loss.dim() == 0 and index.dtype == torch.int64
    assert torch.equal(index.cpu(), gold["context_index"])
    assert torch.allclose(sisnrs.cpu(), gold["sisnrs"], atol=1e-4)
    assert torch.allclose(loss.detach().cpu(), gold["ctx_loss"], atol=2e-6)
    assert torch.allclose(logits.grad.cpu(), 3.0 * gold["dlogits"], atol=2e-6)
    picked, pick = selection.select_stream(est.to(DEV), ctx_pred.to(DEV), case["ce"])
    assert torch.equal(pick.cpu(), gold["pick"])
    ref_picked, _ = SO.select_stream(est, ctx_pred, case["ce"])
    assert torch.equal(picked.cpu(), ref_picked)                       # a gather: bit-exact
    acc, vals = selection.selection_accuracy(picked, torch.cat([gt.unsqueeze(-1), ns], -1).to(DEV))
    assert torch.equal(acc.cpu(), gold["acc"]) and torch.allclose(vals.cpu(), gold["acc_sisnrs"], atol=1e-4)


@pytest.mark.gpu
def test_selection_at_the_bench_batch_shape_and_edge_cases():
    """B = 16 x 4 s (BASELINE configs[1]) against the oracle live; ties and NaNs follow torch.argmax; CPU tensors raise."""
    from cse_b200 import selection, synth
    _, src = synth.make_mixture(16, 32000, 2, seed=61)
    g = torch.Generator().manual_seed(61)
    est = (0.6 * src + 0.4 * src.flip(-1) + 0.1 * torch.randn(src.shape, generator=g)).contiguous()
    est[::2] = est[::2].flip(-1)
    logits = torch.randn(16, 2, generator=g)
    gt = src[:, :, 0].contiguous()
    loss, index, sisnrs = selection.selection_loss(logits.to(DEV), est.to(DEV), gt.to(DEV), True)
    rl, ri, rs = SO.selection_loss(logits, est, gt, True)
    assert torch.equal(index.cpu(), ri) and torch.allclose(sisnrs.cpu(), rs, atol=1e-4)
    assert abs(loss.item() - rl.item()) < 1e-5
    same = est.clone()
    same[:, :, 1] = same[:, :, 0]                                       # tie -> first index
    _, index, _ = selection.selection_loss(logits.to(DEV), same.to(DEV), gt.to(DEV), True)
    assert index.cpu().tolist() == [0] * 16
    bad = est.clone()
    bad[3, 100, 1] = float("nan")                                       # NaN counts as the maximum (torch.argmax)
    _, index, s = selection.selection_loss(logits.to(DEV), bad.to(DEV), gt.to(DEV), True)
    assert index[3].item() == 1 and torch.isnan(s[3, 1]).item()
    with pytest.raises(RuntimeError):
        selection.selection_loss(logits, est, gt, True)
