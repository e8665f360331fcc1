"""Golden vectors of the ContSep selection tail: the reference has no function for it, so this script evaluates the
very expressions of train_ContSep.py:386-388 and test.py:234-239,248-255 over the speechbrain shim's own `cal_si_snr`
(the same object the reference scripts call).  Run in the authoring container:  python tests/golden/make_golden_selection.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cse_b200  # noqa: E402,F401
from oracle import run_reference as R  # noqa: E402
from cases import SELECTION_CASES, selection_inputs  # noqa: E402


def main():
    sp_si_snr_criterion, _, _ = R.reference_losses()
    for name, case in SELECTION_CASES.items():
        gt_sp, ns, enhanced_sp, ctx_pred = selection_inputs(case)
        ce, n = case["ce"], case["spk"]
        # --- train_ContSep.py:386-388 ---
        sisnrs = -1. * sp_si_snr_criterion(gt_sp.unsqueeze(-1).repeat(1, 1, n).transpose(0, 1),
                                           enhanced_sp.transpose(0, 1).float().clone()).squeeze(0).detach()
        context_index = sisnrs.argmax(-1)
        selection_criterion = nn.CrossEntropyLoss() if ce else nn.BCEWithLogitsLoss()
        logits = ctx_pred.clone().requires_grad_(True)
        ctx_loss = selection_criterion(logits.squeeze(1), context_index.float() if n == 2 and not ce else context_index)
        ctx_loss.backward()
        # --- test.py:234-239 ---
        if ce:
            pick = nn.functional.softmax(ctx_pred.squeeze(-1), dim=-1).argmax(-1)
        else:
            pick = (nn.functional.sigmoid(ctx_pred.squeeze(-1)) > 0.5).int()
        picked = enhanced_sp[torch.arange(pick.size(0)), :, pick.long()]
        # --- test.py:248-255 ---
        w_gt = -1. * sp_si_snr_criterion(gt_sp.unsqueeze(-1).transpose(0, 1), picked.unsqueeze(-1).transpose(0, 1).float().clone()).squeeze(0).squeeze(-1)
        acc = torch.ones_like(w_gt).int()
        w_ns = []
        for j in range(ns.shape[2]):
            v = -1. * sp_si_snr_criterion(ns[:, :, j].unsqueeze(-1).transpose(0, 1), picked.unsqueeze(-1).transpose(0, 1).float().clone()).squeeze(0).squeeze(-1)
            acc = acc * (w_gt >= v).int()
            w_ns.append(v)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sisnrs=sisnrs.numpy(), context_index=context_index.numpy(),
                            ctx_loss=ctx_loss.detach().numpy(), dlogits=logits.grad.numpy(), pick=pick.long().numpy(),
                            picked_head=picked[:, :64].numpy(), acc=acc.numpy(),
                            acc_sisnrs=torch.stack([w_gt] + w_ns, -1).numpy())
        print(name, "loss", float(ctx_loss), "index", context_index.tolist(), "pick", pick.tolist(), "acc", acc.tolist())


if __name__ == "__main__":
    main()
