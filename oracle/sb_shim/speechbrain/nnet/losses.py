"""speechbrain.nnet.losses: cal_si_snr / PitWrapper / get_si_snr_with_pitwrapper
(used at train_ContSep.py:346,352,386,391-393; test.py:248-252).  Test infrastructure."""
from itertools import permutations

import torch
import torch.nn as nn


def get_mask(source, source_lengths):
    """[T,B,1]-broadcastable validity mask (all ones when every length == T)."""
    mask = source.new_ones(source.size()[:-1]).unsqueeze(-1).transpose(1, -2)
    B = source.size(-2)
    for i in range(B):
        mask[source_lengths[i]:, i] = 0
    return mask.transpose(-2, 1)


def cal_si_snr(source, estimate):
    """NEGATIVE SI-SNR, shape [1,B,C]; inputs [T,B,C]; mutates `estimate` in place (x ones)."""
    EPS = 1e-8
    assert source.size() == estimate.size()
    device = estimate.device.type
    source_lengths = torch.tensor([estimate.shape[0]] * estimate.shape[-2], device=device)
    mask = get_mask(source, source_lengths)
    estimate *= mask
    num_samples = source_lengths.contiguous().reshape(1, -1, 1).float()
    mean_target = torch.sum(source, dim=0, keepdim=True) / num_samples
    mean_estimate = torch.sum(estimate, dim=0, keepdim=True) / num_samples
    zero_mean_target = source - mean_target
    zero_mean_estimate = estimate - mean_estimate
    zero_mean_target *= mask
    zero_mean_estimate *= mask
    s_target = zero_mean_target
    s_estimate = zero_mean_estimate
    dot = torch.sum(s_estimate * s_target, dim=0, keepdim=True)
    s_target_energy = torch.sum(s_target ** 2, dim=0, keepdim=True) + EPS
    proj = dot * s_target / s_target_energy
    e_noise = s_estimate - proj
    si_snr_beforelog = torch.sum(proj ** 2, dim=0) / (torch.sum(e_noise ** 2, dim=0) + EPS)
    si_snr = 10 * torch.log10(si_snr_beforelog + EPS)
    return -si_snr.unsqueeze(0)


class PitWrapper(nn.Module):
    def __init__(self, base_loss):
        super().__init__()
        self.base_loss = base_loss

    def _fast_pit(self, loss_mat):
        loss, assigned = None, None
        n = loss_mat.shape[0]
        for p in permutations(range(n)):
            c = loss_mat[range(n), p].mean()
            if loss is None or loss > c:
                loss, assigned = c, p
        return loss, assigned

    def _opt_perm_loss(self, pred, target):
        n = pred.size(-1)
        pred = pred.unsqueeze(-2).repeat(*[1 for _ in range(len(pred.shape) - 1)], n, 1)
        target = target.unsqueeze(-1).repeat(1, *[1 for _ in range(len(target.shape) - 1)], n)
        loss_mat = self.base_loss(pred, target)
        assert len(loss_mat.shape) >= 2 and loss_mat.shape[-2:] == target.shape[-2:]
        lead = list(range(len(loss_mat.shape)))[:-2]
        loss_mat = loss_mat.mean(dim=lead)
        return self._fast_pit(loss_mat)

    def forward(self, preds, targets):
        losses, perms = [], []
        for pred, label in zip(preds, targets):
            loss, p = self._opt_perm_loss(pred, label)
            perms.append(p)
            losses.append(loss)
        return torch.stack(losses), perms


def get_si_snr_with_pitwrapper(source, estimate_source):
    loss, _ = PitWrapper(cal_si_snr)(source, estimate_source)
    return loss
