"""GPU: SI-SNR / PIT reduction kernels against the reference-generated golden values and the
CPU oracle."""
import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import losses
from helpers import LOSS_CASES, load_golden, loss_case
from oracle import sepformer_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_loss_matches_golden(name):
    kind, est, tgt = loss_case(name)
    gold = load_golden(name)["value"]
    e, t = est.to(DEV), tgt.to(DEV)
    if kind == "cal_si_snr":
        v = losses.cal_si_snr(t.transpose(0, 1), e.transpose(0, 1))
    elif kind == "pit":
        v = losses.get_si_snr_with_pitwrapper(e, t)
    else:
        v = losses.ScaleInvariantSignalNoiseRatio()(e[:, :, 0], t[:, :, 0])
    assert v.shape == gold.shape
    assert torch.allclose(v.cpu(), gold, rtol=1e-4, atol=2e-4)


def test_pit_permutation_and_invariance():
    _, est, tgt = loss_case("pit_b2_t3000_c3")
    loss, perm = losses.get_si_snr_with_pitwrapper(est.to(DEV), tgt.to(DEV), return_perms=True)
    ref_loss, ref_perm = O.pit_si_snr(est, tgt)
    assert torch.allclose(loss.cpu(), ref_loss, atol=2e-4)
    assert perm.cpu().tolist() == [list(p) for p in ref_perm]
    loss2 = losses.get_si_snr_with_pitwrapper(est.flip(-1).contiguous().to(DEV), tgt.to(DEV))
    assert torch.allclose(loss, loss2, atol=2e-4)
    # inputs are not mutated (speechbrain's version multiplies `estimate` in place)
    e = est.to(DEV)
    before = e.clone()
    losses.cal_si_snr(tgt.to(DEV).transpose(0, 1), e.transpose(0, 1))
    assert torch.equal(e, before)


def test_si_snr_full_length_and_perfect_estimate():
    g = torch.Generator().manual_seed(5)
    t = torch.randn(2, 128000, 2, generator=g)
    e = t + 0.01 * torch.randn(2, 128000, 2, generator=g)
    v = losses.cal_si_snr(t.to(DEV).transpose(0, 1), e.to(DEV).transpose(0, 1))
    ref = O.cal_si_snr(t.double().transpose(0, 1), e.double().transpose(0, 1))
    assert torch.allclose(v.cpu().double(), ref, atol=1e-3)
    assert (v < -35).all()                                # ~40 dB SI-SNR, negated
