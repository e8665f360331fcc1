"""Generate tests/golden/*.npz by running the REFERENCE's own modules (imported unchanged from
/root/reference over oracle/sb_shim) on seeded synthetic weights and inputs.

Run in the authoring container only:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box; these fixtures can.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cse_b200  # noqa: E402,F401
from cse_b200 import synth  # noqa: E402
from oracle import run_reference as R  # noqa: E402
from cases import LOSS_CASES, MODEL_CASES  # noqa: E402


def model_inputs(case):
    variant, spk, ce, c, B, T, cue, wseed, iseed = case
    mix, src = synth.make_mixture(B, T, max(spk, 2), seed=iseed)
    ctx = synth.make_context(B, c, seed=iseed) if variant != "sepformer" else None
    se = synth.make_speaker_embedding(B, seed=iseed) if variant == "hcontext" else None
    return mix, src, ctx, se


def loss_inputs(case):
    kind, B, T, C, seed = case
    _, a = synth.make_mixture(B, T, max(C, 2), seed=seed)
    _, b = synth.make_mixture(B, T, max(C, 2), seed=seed + 1000)
    a, b = a[:, :, :C], b[:, :, :C]
    est = 0.6 * a + 0.4 * b.flip(-1)            # correlated with the target, imperfect
    return est.contiguous(), a.contiguous()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    for name, case in MODEL_CASES.items():
        variant, spk, ce, c, B, T, cue, wseed, iseed = case
        sd = synth.make_state_dict(variant, spk, seed=wseed, ce=ce)
        model = R.build_reference_model(variant, spk, ce=ce).eval()
        model.load_state_dict(sd, strict=True)
        mix, src, ctx, se = model_inputs(case)
        with torch.no_grad():
            if variant == "sepformer":
                out = model(mix)
            elif variant == "hcontext":
                out = model(mix, ctx, se, cue=cue)
            else:
                out = model(mix, ctx)
        arrays = {}
        if variant == "contsep":
            arrays["est"], arrays["context_pred"] = out[0].numpy(), out[1].numpy()
        else:
            arrays["est"] = out.numpy()
        # The reference's own reduced-precision path (its --bf16 switch, train_ContSep.py:383), run
        # here under CPU autocast because this container has no GPU: the yardstick for how far a
        # bf16 implementation of this network may drift from the fp32 output.
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            if variant == "sepformer":
                out16 = model(mix)
            elif variant == "hcontext":
                out16 = model(mix, ctx, se, cue=cue)
            else:
                out16 = model(mix, ctx)
        out16 = out16[0] if isinstance(out16, tuple) else out16
        arrays["est_bf16_ref"] = out16.float().numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()}, float(np.abs(arrays["est"]).mean()))

    cal, pit, TM = R.reference_losses()
    for name, case in LOSS_CASES.items():
        kind, B, T, C, seed = case
        est, tgt = loss_inputs(case)
        if kind == "cal_si_snr":      # call shape of train_ContSep.py:386 -> [T,B,C]
            v = cal(tgt.transpose(0, 1).clone(), est.transpose(0, 1).clone())
        elif kind == "pit":           # train_ContSep.py:391-393: (estimate, targets)
            v = pit(est.clone(), tgt.clone())
        else:                         # train_ContExt.py:367: criterion(est[:, :, 0], gt)
            v = TM()(est[:, :, 0], tgt[:, :, 0])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), value=v.numpy())
        print(name, v.flatten()[:4])


if __name__ == "__main__":
    main()
