"""Generate the BASELINE-shape fixtures `tests/golden/baseline_*.npz` by running the REFERENCE's own
modules (imported unchanged from /root/reference over oracle/sb_shim) at the shapes BASELINE.json names.

Run in the authoring container only:  python tests/golden/make_golden_baseline.py [case ...]
The reference cannot travel to the GPU box; these fixtures can.  `tests/test_baseline_shapes_gpu.py`
compares the CUDA path with them AND with the CPU oracle run live on the box.

What is stored per case (mixtures are independent — GroupNorm(1,.) is per sample, attention never crosses
samples — so a batch item run alone at B=1 equals its slice of the batched output):
  est, context_pred        the reference's fp32 forward
  est_bf16_ref, ...        the reference's own reduced-precision path (its --bf16 switch,
                           train_ContSep.py:383) under CPU autocast: the yardstick for bf16 drift
  for the 32 s case only windows of the waveforms are kept (`win`), plus the full-length drift scalars.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cse_b200  # noqa: E402,F401
from cse_b200 import synth  # noqa: E402
from oracle import run_reference as R  # noqa: E402
from cases import BASELINE_CASES, baseline_inputs  # noqa: E402


def si_snr_db(est, src):
    e = est.double() - est.double().mean(1, keepdim=True)
    s = src.double() - src.double().mean(1, keepdim=True)
    dot = torch.einsum("bte,bts->bes", e, s)
    s_en = (s * s).sum(1).unsqueeze(1) + 1e-12
    proj = dot * dot / s_en
    noise = (e * e).sum(1).unsqueeze(2) - proj
    return 10 * torch.log10(proj / noise.clamp_min(1e-30) + 1e-30)


def run(model, variant, mix, ctx, se, cue):
    if variant == "sepformer":
        return model(mix)
    if variant == "hcontext":
        return model(mix, ctx, se, cue=cue)
    return model(mix, ctx)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    want = sys.argv[1:] or list(BASELINE_CASES)
    for name in want:
        case = BASELINE_CASES[name]
        variant, spk, c, T, cue, wseed = case["variant"], case["spk"], case["c"], case["T"], case["cue"], case["wseed"]
        sd = synth.make_state_dict(variant, spk, seed=wseed)
        model = R.build_reference_model(variant, spk).eval()
        model.load_state_dict(sd, strict=True)
        mix, src, ctx, se = baseline_inputs(case)
        t0 = time.time()
        with torch.no_grad():
            out = run(model, variant, mix, ctx, se, cue)
        t1 = time.time()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            out16 = run(model, variant, mix, ctx, se, cue)
        t2 = time.time()
        est, pred = (out if isinstance(out, tuple) else (out, None))
        est16, pred16 = (out16 if isinstance(out16, tuple) else (out16, None))
        est16 = est16.float()
        arrays = {}
        rel = ((est16.double() - est.double()).norm() / est.double().norm()).item()
        nsrc = min(src.shape[2], 3)
        raw = (si_snr_db(est16, src[:, :, :nsrc]) - si_snr_db(est, src[:, :, :nsrc])).abs().max().item()
        arrays["bf16_ref_rel_l2"] = np.float64(rel)
        arrays["bf16_ref_raw_dsisnr_db"] = np.float64(raw)
        if case.get("window"):
            w = case["window"]
            arrays["win"] = np.int64(w)
            arrays["est_head"], arrays["est_tail"] = est[:, :w].numpy(), est[:, -w:].numpy()
            arrays["est_norm"] = np.float64(est.double().norm().item())
        else:
            arrays["est"] = est.numpy()
            arrays["est_bf16_ref"] = est16.numpy()
        if pred is not None:
            arrays["context_pred"] = pred.numpy()
            arrays["context_pred_bf16_ref"] = pred16.float().numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(f"{name}: fp32 {t1 - t0:.1f} s, bf16 {t2 - t1:.1f} s; reference bf16 drift rel-L2 {rel:.3e}, "
              f"raw-source |dSI-SNR| {raw:.3f} dB; est {tuple(est.shape)}", flush=True)


if __name__ == "__main__":
    main()
