"""CPU restatement of the ContSep selection tail.  TEST INFRASTRUCTURE ONLY.

The reference has no function for it: the arithmetic is inline in its scripts —
  train_ContSep.py:386-388   sisnrs / context_index / ctx_loss
  test.py:234-239            stream pick from ctx_pred
  test.py:248-255            acc_pred
restated here line for line on top of `sepformer_oracle.cal_si_snr` (the speechbrain restatement, pinned by the loss
fixtures).  Pinned itself by `tests/golden/selection_*.npz` (make_golden_selection.py runs the very expressions of
those script lines over the speechbrain shim's own `cal_si_snr`).
"""
import torch
import torch.nn.functional as F

from . import sepformer_oracle as O


def selection_loss(ctx_pred, est, gt, ce=True):
    """ctx_pred [B,spk] or [B,1]; est [B,T,spk]; gt [B,T] -> (ctx_loss, context_index [B], sisnrs [B,spk])."""
    n = est.shape[2]
    source = gt.unsqueeze(-1).repeat(1, 1, n).transpose(0, 1)                     # [T,B,spk]
    sisnrs = (-1.0 * O.cal_si_snr(source, est.transpose(0, 1).float())).squeeze(0).detach()
    context_index = sisnrs.argmax(-1)
    if ce:
        loss = F.cross_entropy(ctx_pred.squeeze(1), context_index)
    else:
        loss = F.binary_cross_entropy_with_logits(ctx_pred.squeeze(1), context_index.float())
    return loss, context_index, sisnrs


def select_stream(est, ctx_pred, ce=True):
    if ce:
        pick = F.softmax(ctx_pred.squeeze(-1), dim=-1).argmax(-1)
    else:
        pick = (torch.sigmoid(ctx_pred.squeeze(-1)) > 0.5).long()
    return est[torch.arange(pick.size(0)), :, pick], pick


def selection_accuracy(enhanced, sources):
    """enhanced [B,T]; sources [B,T,C] (column 0 = gt, then ns_1, ns_2)."""
    e = enhanced.unsqueeze(-1).transpose(0, 1).float()                            # [T,B,1]
    vals = [(-1.0 * O.cal_si_snr(sources[:, :, j].unsqueeze(-1).transpose(0, 1), e)).squeeze(0).squeeze(-1)
            for j in range(sources.shape[2])]
    acc = torch.ones_like(vals[0], dtype=torch.int32)
    for v in vals[1:]:
        acc = acc * (vals[0] >= v).int()
    return acc, torch.stack(vals, -1)
