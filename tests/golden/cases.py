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
