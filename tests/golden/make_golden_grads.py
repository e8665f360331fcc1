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
        # The reference's own reduced-precision training step (its --bf16 switch, train_ContSep.py:383-386: forward
        # and loss under autocast, backward outside), run under CPU autocast because this container has no GPU:
        # the yardstick for how far a bf16 tensor-core backward may drift from the fp32 gradients.
        g32 = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        model.zero_grad(set_to_none=True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out = model(mix, ctx)
            if kind == "tm_neg_sisnr":
                loss16 = -TM()(out[:, :, 0].float(), src[:, :, 0])
            else:
                est, pred = out
                loss16 = pit(est.float(), src[:, :, :spk].clone()).mean() + 0.1 * torch.logsumexp(pred.float(), -1).mean()
        loss16.backward()
        num = sum(((p.grad.double() - g32[k].double()) ** 2).sum().item() for k, p in model.named_parameters() if k in g32)
        den = sum((g.double() ** 2).sum().item() for g in g32.values())
        drift = (num / den) ** 0.5
        per_param = [((p.grad.double() - g32[k].double()).norm() / g32[k].double().norm().clamp_min(1e-30)).item()
                     if k in g32 else 0.0 for k, p in model.named_parameters()]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), loss=np.float64(loss.item()),
                            names=np.array(names), norms=np.array(norms), heads=np.stack(heads),
                            bf16_ref_loss=np.float64(loss16.item()), bf16_ref_global_rel_l2=np.float64(drift),
                            bf16_ref_param_rel_l2=np.array(per_param))
        print(name, loss.item(), len(names), float(np.sum(np.square(norms))) ** 0.5,
              "| reference bf16 autocast: loss", loss16.item(), "global grad rel-L2", drift,
              "median/max per-param", float(np.median(per_param)), float(np.max(per_param)))


if __name__ == "__main__":
    main()
