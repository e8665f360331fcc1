"""Evaluation metrics on the device (test.py:198-201,241-245,291-301): SDR (torchmetrics
signal_distortion_ratio), streaming SI-SNR / SDR accumulators, the whole EvalMeter bookkeeping."""
import numpy as np
import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib, synth
from oracle import metrics_oracle as MO

DEV = "cuda:0"


def _case(B, T, seed, noise=0.05):
    mix, src = synth.make_mixture(B, T, 2, seed=seed)
    g = torch.Generator().manual_seed(seed)
    est = 0.7 * src[:, :, 0] + 0.3 * src[:, :, 1] + noise * torch.randn(B, T, generator=g)
    return mix, src, est


def test_sdr_oracle_agrees_with_time_domain_definition():
    _, src, est = _case(2, 3000, 5)
    for b in range(2):
        a = MO.signal_distortion_ratio(est[b].numpy(), src[b, :, 0].numpy(), filter_length=64)
        d, _ = MO.signal_distortion_ratio_time_domain(est[b].numpy(), src[b, :, 0].numpy(), filter_length=64)
        assert abs(a - d) < 1e-9


def test_sdr_oracle_properties():
    _, src, est = _case(1, 4000, 6)
    t, p = src[0, :, 0].numpy().astype(np.float64), est[0].numpy().astype(np.float64)
    a = MO.signal_distortion_ratio(p, t)
    assert abs(MO.signal_distortion_ratio(3.7 * p, 0.2 * t) - a) < 1e-9          # scale invariance of both arguments
    assert MO.signal_distortion_ratio(t + 1e-4 * p, t) > 60                      # near-perfect estimate
    delayed = np.concatenate([np.zeros(5), t[:-5]])
    assert MO.signal_distortion_ratio(delayed, t) > 20                           # a 5-sample delay is inside the 512-tap filter
    assert MO.signal_distortion_ratio(np.concatenate([t[5:], np.zeros(5)]), t) < 0   # an advance is not (causal filter)


def test_metric_entry_points_reject_bad_arguments():
    lib = _lib.load()
    assert lib.cse_sdr_workspace_bytes(0, 100, 512) == 0
    assert lib.cse_sdr_workspace_bytes(2, 4000, 512) == (2 * 4 + 2 * 2 * 2 * 512) * 8
    with pytest.raises(_lib.CseError):
        _lib.call("cse_sdr", None, None, 1, 10, 512, 0, 0, 0.0, None, None, 0, None)
    from cse_b200 import metrics
    with pytest.raises(_lib.CseError):
        metrics.signal_distortion_ratio(torch.zeros(1, 100), torch.zeros(1, 100))     # CPU tensors: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,L", [(3, 4000, 512), (2, 32000, 512), (1, 700, 512), (2, 5000, 64), (1, 2500, 1024)])
def test_sdr_matches_oracle(B, T, L):
    from cse_b200 import metrics
    mix, src, est = _case(B, T, 20 + B)
    for preds in (est, mix):
        got = metrics.signal_distortion_ratio(preds.to(DEV), src[:, :, 0].contiguous().to(DEV), filter_length=L)
        ref = MO.signal_distortion_ratio(preds.numpy(), src[:, :, 0].numpy(), filter_length=L)
        assert got.dtype == torch.float32 and got.shape == (B,)
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=2e-5)   # dB; float32 output rounding


@pytest.mark.gpu
def test_sdr_options_zero_mean_and_load_diag():
    from cse_b200 import metrics
    _, src, est = _case(2, 6000, 31)
    est = est + 0.05
    t = src[:, :, 0].contiguous()
    for kw in (dict(zero_mean=True), dict(load_diag=1e-3), dict(zero_mean=True, load_diag=1e-2)):
        got = metrics.signal_distortion_ratio(est.to(DEV), t.to(DEV), **kw)
        ref = MO.signal_distortion_ratio(est.numpy(), t.numpy(), **kw)
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=2e-5)
    with pytest.raises(NotImplementedError):
        metrics.signal_distortion_ratio(est.to(DEV), t.to(DEV), use_cg_iter=10)


@pytest.mark.gpu
def test_streaming_metric_objects_accumulate_like_torchmetrics():
    from cse_b200 import metrics
    sdr, sis = metrics.SignalDistortionRatio(), metrics.StreamingSiSnr()
    o_sdr, o_sis = MO.RunningMean(), MO.RunningMean()
    for i, (B, T) in enumerate([(3, 4000), (2, 4000), (1, 5000)]):      # ragged batches, as the eval loader yields
        _, src, est = _case(B, T, 40 + i)
        t = src[:, :, 0].contiguous()
        sdr.update(est.to(DEV), t.to(DEV))
        last = sis(est.to(DEV), t.to(DEV))                              # forward(): batch mean + accumulate
        o_sdr.update(MO.signal_distortion_ratio(est.numpy(), t.numpy()))
        v = MO.scale_invariant_signal_noise_ratio(est.numpy(), t.numpy())
        o_sis.update(v)
        assert abs(last.item() - v.mean()) < 1e-4
    assert abs(sdr.compute().item() - o_sdr.compute()) < 2e-5
    assert abs(sis.compute().item() - o_sis.compute()) < 1e-4
    sdr.reset()
    with pytest.raises(RuntimeError):
        sdr.compute()


@pytest.mark.gpu
def test_eval_meter_matches_the_reference_loop_bookkeeping():
    """test.py:241-255,291-301 — SI-SNR, SDR, their improvements over the unprocessed mixture, selection accuracy."""
    from cse_b200 import metrics
    meter = metrics.EvalMeter()
    o = {k: MO.RunningMean() for k in ("si", "sdr", "si0", "sdr0", "acc")}
    for i in range(2):
        mix, src, est = _case(3, 4000, 50 + i, noise=0.3)
        if i == 1:
            est[0] = 0.2 * src[0, :, 0] + 0.8 * src[0, :, 1]            # one item closer to the interferer
        gt, ns = src[:, :, 0].contiguous(), src[:, :, 1].contiguous()
        meter.update(est.to(DEV), mix.to(DEV), gt.to(DEV), interferers=[ns.to(DEV)])
        o["si"].update(MO.scale_invariant_signal_noise_ratio(est.numpy(), gt.numpy()))
        o["sdr"].update(MO.signal_distortion_ratio(est.numpy(), gt.numpy()))
        o["si0"].update(MO.scale_invariant_signal_noise_ratio(mix.numpy(), gt.numpy()))
        o["sdr0"].update(MO.signal_distortion_ratio(mix.numpy(), gt.numpy()))
        from oracle import selection_oracle as SO
        o["acc"].update(SO.selection_accuracy(est, torch.stack([gt, ns], -1))[0].numpy())
    got = meter.compute()
    assert abs(got["si_snr"] - o["si"].compute()) < 1e-4
    assert abs(got["sdr"] - o["sdr"].compute()) < 2e-5
    assert abs(got["si_snr_i"] - (o["si"].compute() - o["si0"].compute())) < 2e-4
    assert abs(got["sdr_i"] - (o["sdr"].compute() - o["sdr0"].compute())) < 4e-5
    assert abs(got["acc"] - o["acc"].compute()) < 1e-7 and 0 < got["acc"] < 1
