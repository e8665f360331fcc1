"""Generate tests/golden/grad_*.npz: gradients of the reference's training losses computed by the
REFERENCE's own modules and loss objects (imported unchanged from /root/reference over
oracle/sb_shim) under autograd, on the seeded weights / inputs of MODEL_CASES.

Run in the authoring container only:  python tests/golden/make_golden_grads.py
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
from cases import GRAD_CASES, GRAD_HEAD, MODEL_CASES  # noqa: E402
from make_golden import model_inputs  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    _, pit, TM = R.reference_losses()
    for name, (model_name, kind) in GRAD_CASES.items():
        case = MODEL_CASES[model_name]
        variant, spk, ce, c, B, T, cue, wseed, iseed = case
        model = R.build_reference_model(variant, spk, ce=ce).train()
        model.load_state_dict(synth.make_state_dict(variant, spk, seed=wseed, ce=ce), strict=True)
        mix, src, ctx, se = model_inputs(case)
        out = model(mix, ctx)
        if kind == "tm_neg_sisnr":
            loss = -TM()(out[:, :, 0], src[:, :, 0])
        else:
            est, pred = out
            loss = pit(est, src[:, :, :spk].clone()).mean() + 0.1 * torch.logsumexp(pred, -1).mean()
        loss.backward()
        names, norms, heads = [], [], []
        for k, p in model.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            names.append(k)
            norms.append(g.double().norm().item())
            h = torch.zeros(GRAD_HEAD)
            flat = g.flatten()[:GRAD_HEAD]
            h[: flat.numel()] = flat
            heads.append(h.numpy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), loss=np.float64(loss.item()),
                            names=np.array(names), norms=np.array(norms), heads=np.stack(heads))
        print(name, loss.item(), len(names), float(np.sum(np.square(norms))) ** 0.5)


if __name__ == "__main__":
    main()
