"""Golden-vector case table shared by make_golden.py and the tests."""

# name: (variant, num_spks, ce, c, B, T, cue, weight seed, input seed)
MODEL_CASES = {
    "contsep_2spk_b2_t4000": ("contsep", 2, True, 1, 2, 4000, None, 1, 11),
    "contsep_2spk_bce_b1_t2024": ("contsep", 2, False, 1, 1, 2024, None, 2, 12),
    "contsep_3spk_b1_t2100": ("contsep", 3, True, 1, 1, 2100, None, 3, 13),
    "contsep_2spk_b1_t16000": ("contsep", 2, True, 1, 1, 16000, None, 4, 14),
    "sepformer_2spk_b1_t4003": ("sepformer", 2, True, 0, 1, 4003, None, 5, 15),
    "sepformer_3spk_b2_t1000": ("sepformer", 3, True, 0, 2, 1000, None, 6, 16),
    "context_2spk_b2_t3000": ("context", 2, True, 1, 2, 3000, None, 7, 17),
    "context_2spk_c3_b1_t2000": ("context", 2, True, 3, 1, 2000, None, 8, 18),
    "hcontext_3spk_joint_b2_t2500": ("hcontext", 3, True, 1, 2, 2500, "joint", 9, 19),
    "hcontext_2spk_history_b1_t1999": ("hcontext", 2, True, 1, 1, 1999, "history", 10, 20),
    "hcontext_2spk_voice_b1_t1999": ("hcontext", 2, True, 1, 1, 1999, "voice", 10, 20),
    "contsep_2spk_b1_t16_minimal": ("contsep", 2, True, 1, 1, 16, None, 1, 21),
}

# name: (kind, B, T, C, seed)
LOSS_CASES = {
    "cal_si_snr_b3_t4000_c2": ("cal_si_snr", 3, 4000, 2, 31),
    "pit_b3_t4000_c2": ("pit", 3, 4000, 2, 32),
    "pit_b2_t3000_c3": ("pit", 2, 3000, 3, 33),
    "tm_si_snr_b4_t5000": ("tm_si_snr", 4, 5000, 1, 34),
}

# Gradient fixtures (make_golden_grads.py): name -> (model case, loss kind).  The REFERENCE model in
# train() mode, its own loss objects, `loss.backward()`; stored per parameter: gradient norm + the first
# 64 flattened gradient values.
#   tm_neg_sisnr : -ScaleInvariantSignalNoiseRatio()(est[:, :, 0], gt)            train_ContExt.py:366-367
#   pit_plus_sel : get_si_snr_with_pitwrapper(est, gt).mean() + 0.1 * logsumexp(context_pred).mean()
#                  (PIT SI-SNR of train_ContSep.py:391-393 plus a term that reaches pred_head / context_selector)
GRAD_CASES = {
    "grad_context_2spk_b2_t3000": ("context_2spk_b2_t3000", "tm_neg_sisnr"),
    "grad_contsep_2spk_bce_b1_t2024": ("contsep_2spk_bce_b1_t2024", "pit_plus_sel"),
}
GRAD_HEAD = 64

# BASELINE.json shapes (make_golden_baseline.py; tests/test_baseline_shapes_gpu.py).  `batch`/`index`: the case is
# item `index` of a seeded batch of `batch` mixtures (the bench.py batch for cfg 2), run alone at B=1.
BASELINE_CASES = {
    # configs[1]: ContSep 2-spk, 16 x 4 s, c = 1 — bench.py's own batch (weights seed 0, inputs seed 1234), items 0 and 15
    "baseline_cfg2_mix0": dict(variant="contsep", spk=2, c=1, T=32000, cue=None, wseed=0, iseed=1234, batch=16, index=0),
    "baseline_cfg2_mix15": dict(variant="contsep", spk=2, c=1, T=32000, cue=None, wseed=0, iseed=1234, batch=16, index=15),
    # configs[3]: H-ContExt 3-spk, 16 s (max_sp_len), ctx + speaker token (c = 2): intra n = 252, inter n = 132
    "baseline_cfg4_hcontext_3spk_16s": dict(variant="hcontext", spk=3, c=1, T=128000, cue="joint", wseed=40, iseed=41,
                                           batch=1, index=0),
    # 32 s: inter n = 259 > 256 — the attention path beyond the tcgen05 kernel's whole-row-in-TMEM limit
    "baseline_32s_contsep_2spk": dict(variant="contsep", spk=2, c=1, T=256000, cue=None, wseed=42, iseed=43,
                                     batch=1, index=0, window=16000),
}


def baseline_inputs(case):
    """(mix [1,T], src [1,T,spk], ctx [1,c,4096], se [1,1,192] | None) of a BASELINE_CASES entry."""
    from cse_b200 import synth
    B, i = case["batch"], case["index"]
    mix, src = synth.make_mixture(B, case["T"], max(case["spk"], 2), seed=case["iseed"])
    ctx = synth.make_context(B, case["c"], seed=case["iseed"])
    se = synth.make_speaker_embedding(B, seed=case["iseed"]) if case["variant"] == "hcontext" else None
    sl = slice(i, i + 1)
    return mix[sl].contiguous(), src[sl].contiguous(), ctx[sl].contiguous(), (None if se is None else se[sl].contiguous())

# ContSep selection tail (make_golden_selection.py).
SELECTION_CASES = {
    "selection_ce_2spk_b4_t4000": dict(spk=2, ce=True, B=4, T=4000, seed=51),
    "selection_bce_2spk_b3_t3001": dict(spk=2, ce=False, B=3, T=3001, seed=52),
    "selection_ce_3spk_b5_t2500": dict(spk=3, ce=True, B=5, T=2500, seed=53),
}


def selection_inputs(case):
    """(gt [B,T], interferers [B,T,spk-1], est [B,T,spk], ctx_pred [B,spk | 1]): estimates are imperfect mixtures of
    the sources in a per-item random stream order, logits are seeded noise."""
    import torch
    from cse_b200 import synth
    B, T, n = case["B"], case["T"], case["spk"]
    _, src = synth.make_mixture(B, T, max(n, 2), seed=case["seed"])
    g = torch.Generator().manual_seed(case["seed"])
    est = torch.empty(B, T, n)
    for b in range(B):
        perm = torch.randperm(n, generator=g)
        for s in range(n):
            other = src[b, :, perm[(s + 1) % n]]
            est[b, :, s] = 0.7 * src[b, :, perm[s]] + 0.3 * other + 0.05 * torch.randn(T, generator=g)
    ctx_pred = torch.randn(B, n if case["ce"] else 1, generator=g) * 2.0
    return src[:, :, 0].contiguous(), src[:, :, 1:n].contiguous(), est.contiguous(), ctx_pred


# Loader-side mixture synthesis (make_golden_mixture.py): name -> (clip lengths, seed, pad).
# 2 lengths = mix_audio(signal, noise), 3 = mix_audio_3spk(signal, noise1, noise2).
MIXTURE_CASES = {
    "2spk_pad_signal_longer": ((2500, 1700), 61, True),
    "2spk_pad_signal_shorter": ((1600, 2300), 62, True),
    "2spk_loop_signal_longer": ((2400, 1000), 63, False),
    "2spk_loop_signal_shorter": ((1500, 2000), 64, False),
    "3spk_pad": ((1800, 2600, 1200), 65, True),
    "3spk_loop": ((2000, 900, 2700), 66, False),
    "3spk_pad_signal_longest": ((2600, 2000, 2100), 67, True),
}


def mixture_case(name):
    """(clips: list of float32 numpy arrays, peak-normalised to 0.9 like dataset_train_CSE.py:237; snrs: numpy
    float64 scalars drawn as the dataset draws them, np.clip(N(0, 4), -5, 5); pad)."""
    import numpy as np
    lens, seed, pad = MIXTURE_CASES[name]
    rng = np.random.default_rng(seed)
    clips = []
    for n in lens:
        w = rng.standard_normal(n + 64)
        x = np.convolve(w, np.ones(8) / 8.0, mode="valid")[:n] * (1.0 + 0.5 * np.sin(np.arange(n) / 97.0))
        x = x.astype(np.float32)
        clips.append(x / np.max(np.abs(x)) * 0.9)
    snrs = [np.clip(rng.normal(0, 4), -5, 5) for _ in range(len(lens) - 1)]
    assert all(isinstance(s, np.float64) for s in snrs)
    return clips, snrs, pad
