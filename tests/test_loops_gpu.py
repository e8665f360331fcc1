"""The reference's training and evaluation loop bodies (train_ContSep.py:375-419, test.py:228-301) with every step on
the device: mixture synthesis + collate (§8f-2) -> Sepformer forward -> selection loss + PIT SI-SNR (§8f-1, a16/a17)
-> backward -> fused clip + AdamW (§8f-5); and model -> stream pick -> streaming SI-SNR / SDR / accuracy (§8f-1/3).
Small shapes; parity of each piece is tested in its own file — here the chain runs, stays finite, trains, and
matches the same chain evaluated with the CPU oracles."""
import numpy as np
import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import losses, metrics, mixture, selection, synth
from cse_b200.models.ContSep import Sepformer
from cse_b200.optim import AdamW

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


def _clips(seed, lens):
    rng = np.random.default_rng(seed)
    out = []
    for n in lens:
        x = np.convolve(rng.standard_normal(n + 15), np.ones(16) / 16.0, mode="valid").astype(np.float32)
        out.append(x)
    return out


def _batch(seed):
    """Two ragged 16 kHz items -> 8 kHz collated batch, as dataset_train_CSE.py builds them."""
    sig = _clips(seed, (6000, 4800))
    noi = _clips(seed + 1, (5000, 6400))
    s, s_len = mixture.peak_normalize(sig, device=DEV)
    n_, n_len = mixture.peak_normalize(noi, device=DEV)
    mixed, gt, ns, sp_len = mixture.mix_batch([s[b, :l] for b, l in enumerate(s_len)],
                                              [n_[b, :l] for b, l in enumerate(n_len)],
                                              [np.float64(1.0), np.float64(-2.5)], pad=True)
    m8, len8 = mixture.decimate(mixed, sp_len)
    g8, _ = mixture.decimate(gt, sp_len)
    n8, _ = mixture.decimate(ns, sp_len)
    return m8, g8, n8, len8


def _model(train):
    m = Sepformer(2, add_mt=True)
    m.add_mt_pipeline()
    m.load_state_dict(synth.make_state_dict("contsep", 2, seed=3))
    m = m.to(DEV)
    return m.train() if train else m.eval()


def test_training_loop_body_runs_on_the_device_and_learns():
    model = _model(train=True)
    opt = AdamW(model.parameters(), lr=2e-4, weight_decay=1e-6, amsgrad=True)       # train_ContSep.py:233
    mixed, gt, ns, _ = _batch(101)
    ctx = synth.make_context(mixed.shape[0], 1, seed=5).to(DEV)
    history = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):                          # train_ContSep.py:383 (--bf16)
            est, ctx_pred = model(mixed, ctx)
            ctx_loss, context_index, sisnrs = selection.selection_loss(ctx_pred, est, gt, ce=True)   # :386-388
            snr_loss = losses.get_si_snr_with_pitwrapper(est, torch.stack([gt, ns], -1)).mean()      # :391-393
            loss = 0.1 * ctx_loss + snr_loss
        loss.backward()
        grad_norm = opt.step(max_norm=5.0)                                          # :411-416
        history.append(float(loss.detach()))
        assert torch.isfinite(grad_norm) and not opt.found_inf
    assert opt.steps_applied() == 4
    assert history[-1] < history[0], history                                       # the step direction is a descent direction
    assert context_index.shape == (2,) and sisnrs.shape == (2, 2)


def test_evaluation_loop_body_matches_the_cpu_oracles():
    from oracle import metrics_oracle as MO
    from oracle import selection_oracle as SO
    model = _model(train=False)
    meter = metrics.EvalMeter()
    ref = {k: MO.RunningMean() for k in ("si", "sdr", "si0", "sdr0", "acc")}
    for seed in (201, 202):
        mixed, gt, ns, _ = _batch(seed)
        ctx = synth.make_context(mixed.shape[0], 1, seed=seed).to(DEV)
        with torch.no_grad():
            model.precision = "fp32"
            est, ctx_pred = model(mixed, ctx)                                       # test.py:234
            enhanced, pick = selection.select_stream(est, ctx_pred, ce=True)         # test.py:235-239, no .cpu()
        meter.update(enhanced, mixed, gt, interferers=[ns])                          # test.py:241-255
        e, m, g, n = (t.cpu() for t in (enhanced, mixed, gt, ns))
        o_enh, o_pick = SO.select_stream(est.cpu(), ctx_pred.cpu(), ce=True)
        assert torch.equal(pick.cpu(), o_pick) and torch.equal(e, o_enh)
        ref["si"].update(MO.scale_invariant_signal_noise_ratio(e.numpy(), g.numpy()))
        ref["sdr"].update(MO.signal_distortion_ratio(e.numpy(), g.numpy()))
        ref["si0"].update(MO.scale_invariant_signal_noise_ratio(m.numpy(), g.numpy()))
        ref["sdr0"].update(MO.signal_distortion_ratio(m.numpy(), g.numpy()))
        ref["acc"].update(SO.selection_accuracy(e, torch.stack([g, n], -1))[0].numpy())
    got = meter.compute()                                                           # the only host synchronisation
    assert abs(got["si_snr"] - ref["si"].compute()) < 2e-4
    assert abs(got["sdr"] - ref["sdr"].compute()) < 1e-4
    assert abs(got["si_snr_i"] - (ref["si"].compute() - ref["si0"].compute())) < 4e-4
    assert abs(got["sdr_i"] - (ref["sdr"].compute() - ref["sdr0"].compute())) < 2e-4
    assert abs(got["acc"] - ref["acc"].compute()) < 1e-7
