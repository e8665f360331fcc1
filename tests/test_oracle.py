"""CPU: the oracle restatement against (i) the committed golden vectors produced by the
reference's own modules and (ii) the live reference wherever /root/reference exists."""
import pytest
import torch

from helpers import LOSS_CASES, MODEL_CASES, load_golden, loss_case, model_case, rel_l2
from oracle import run_reference as R
from oracle import sepformer_oracle as O

FAST_CASES = [n for n in MODEL_CASES if MODEL_CASES[n][5] <= 4100]


def _run_oracle(sd, mix, ctx, se, meta):
    with torch.no_grad():
        out = O.sepformer_forward(sd, mix, ctx, meta["variant"], meta["spk"], se=se,
                                  cue=meta["cue"] or "joint")
    return out if isinstance(out, tuple) else (out, None)


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_matches_golden(name):
    sd, mix, src, ctx, se, meta = model_case(name)
    gold = load_golden(name)
    est, pred = _run_oracle(sd, mix, ctx, se, meta)
    assert est.shape == gold["est"].shape
    assert rel_l2(est, gold["est"]) < 2e-5            # fp32 vs fp32, different op order
    if "context_pred" in gold:
        assert pred.shape == gold["context_pred"].shape
        assert rel_l2(pred, gold["context_pred"]) < 2e-5


def test_oracle_fp64_brackets_reference():
    """The reference's fp32 output is itself ~1e-6 from exact arithmetic; the fp64 oracle is
    the yardstick used to budget the 1e-4 tolerance of the CUDA fp32 mode."""
    name = "contsep_2spk_bce_b1_t2024"
    sd, mix, src, ctx, se, meta = model_case(name)
    sd64 = {k: v.double() for k, v in sd.items()}
    est64, _ = _run_oracle(sd64, mix, ctx, se, meta)
    assert rel_l2(load_golden(name)["est"], est64) < 1e-5


@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_loss_oracle_matches_golden(name):
    kind, est, tgt = loss_case(name)
    gold = load_golden(name)["value"]
    if kind == "cal_si_snr":
        v = O.cal_si_snr(tgt.transpose(0, 1), est.transpose(0, 1))
    elif kind == "pit":
        v, _ = O.pit_si_snr(est, tgt)
    else:
        v = O.tm_si_snr(est[:, :, 0], tgt[:, :, 0]).mean()
    assert v.shape == gold.shape
    assert torch.allclose(v, gold, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("L", [124, 125, 250, 375, 1000, 1999, 3999])
def test_segmentation_overlap_add_identity(L):
    """_over_add(_Segmentation(x)) == 2 x exactly and the strided-window form."""
    x = torch.randn(2, 3, L)
    seg, gap = O.pad_and_segment(x)
    assert 1 <= gap <= 250
    assert seg.shape[2] == 250 and seg.shape[3] == 2 * (L + gap + 125) // 250
    assert torch.equal(O.overlap_add(seg, gap), 2 * x)


def test_si_snr_invariances():
    _, est, tgt = loss_case("cal_si_snr_b3_t4000_c2")
    e, t = est.transpose(0, 1), tgt.transpose(0, 1)
    base = O.cal_si_snr(t, e)
    assert torch.allclose(O.cal_si_snr(t, 3.7 * e + 0.2), base, atol=1e-3)     # scale + shift
    assert torch.allclose(O.cal_si_snr(2.0 * t, e), base, atol=1e-3)
    loss, perms = O.pit_si_snr(est, tgt)
    loss2, perms2 = O.pit_si_snr(est.flip(-1), tgt)                             # permutation
    assert torch.allclose(loss, loss2, atol=1e-4)


@pytest.mark.skipif(not R.reference_available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("name", ["contsep_2spk_bce_b1_t2024", "hcontext_2spk_voice_b1_t1999",
                                  "sepformer_3spk_b2_t1000"])
def test_oracle_matches_live_reference(name):
    sd, mix, src, ctx, se, meta = model_case(name)
    ref = R.build_reference_model(meta["variant"], meta["spk"], ce=meta["ce"]).eval()
    ref.load_state_dict(sd, strict=True)               # key/shape contract of SURVEY.md §8b
    with torch.no_grad():
        if meta["variant"] == "sepformer":
            r = ref(mix)
        elif meta["variant"] == "hcontext":
            r = ref(mix, ctx, se, cue=meta["cue"])
        else:
            r = ref(mix, ctx)
    r_est = r[0] if isinstance(r, tuple) else r
    est, pred = _run_oracle(sd, mix, ctx, se, meta)
    assert rel_l2(est, r_est) < 2e-5
    assert rel_l2(r_est, load_golden(name)["est"]) < 1e-6     # fixtures are reproducible


@pytest.mark.parametrize("name", ["baseline_cfg2_mix0", "baseline_cfg2_mix15", "baseline_cfg4_hcontext_3spk_16s",
                                  "baseline_32s_contsep_2spk"])
def test_oracle_matches_reference_at_baseline_shapes(name):
    """The pin at the BASELINE.json shapes themselves (cfg 2 items of bench.py's batch, cfg 4 = 16 s 3-spk c = 2,
    32 s): fixtures written by tests/golden/make_golden_baseline.py from the reference's own modules."""
    from cases import BASELINE_CASES, baseline_inputs
    from cse_b200 import synth
    case = BASELINE_CASES[name]
    mix, src, ctx, se = baseline_inputs(case)
    sd = synth.make_state_dict(case["variant"], case["spk"], seed=case["wseed"])
    with torch.no_grad():
        out = O.sepformer_forward(sd, mix, ctx, case["variant"], case["spk"], se=se, cue=case["cue"] or "joint")
    est, pred = out if isinstance(out, tuple) else (out, None)
    fix = load_golden(name)
    if "win" in fix:
        w = int(fix["win"])
        assert rel_l2(est[:, :w], fix["est_head"]) < 2e-5 and rel_l2(est[:, -w:], fix["est_tail"]) < 2e-5
        assert abs(est.double().norm().item() / float(fix["est_norm"]) - 1) < 1e-5
    else:
        assert est.shape == fix["est"].shape and rel_l2(est, fix["est"]) < 2e-5
    if pred is not None:
        assert rel_l2(pred, fix["context_pred"]) < 2e-5


@pytest.mark.parametrize("name", ["contsep_2spk_b2_t4000", "hcontext_3spk_joint_b2_t2500", "sepformer_2spk_b1_t4003",
                                  "contsep_2spk_bce_b1_t2024", "hcontext_2spk_voice_b1_t1999", "context_2spk_c3_b1_t2000"])
def test_eager_reference_is_bit_identical_to_the_reference_fixtures(name):
    """oracle/eager_reference.py (the reference's op sequence on stock torch modules; bench.py's CPU baseline and the
    GPU-eager yardstick) reproduces the reference modules' own fp32 outputs exactly on CPU."""
    from oracle import eager_reference as EY
    from test_forward_gpu import build_model
    sd, mix, src, ctx, se, meta = model_case(name)
    m = build_model(meta)
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        out = EY.eager_forward(m, mix, ctx, se, meta["cue"] or "joint")
    gold = load_golden(name)
    est, pred = out if isinstance(out, tuple) else (out, None)
    assert torch.equal(est, gold["est"])
    if pred is not None:
        assert torch.equal(pred, gold["context_pred"])
