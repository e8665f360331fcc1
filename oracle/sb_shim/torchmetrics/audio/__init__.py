import torch
import torch.nn as nn


def scale_invariant_signal_noise_ratio(preds, target):
    """torchmetrics functional SI-SNR = SI-SDR with zero_mean=True, eps = finfo(dtype).eps."""
    eps = torch.finfo(preds.dtype).eps
    target = target - torch.mean(target, dim=-1, keepdim=True)
    preds = preds - torch.mean(preds, dim=-1, keepdim=True)
    alpha = (torch.sum(preds * target, dim=-1, keepdim=True) + eps) / (
        torch.sum(target ** 2, dim=-1, keepdim=True) + eps
    )
    target_scaled = alpha * target
    noise = target_scaled - preds
    val = (torch.sum(target_scaled ** 2, dim=-1) + eps) / (torch.sum(noise ** 2, dim=-1) + eps)
    return 10 * torch.log10(val)


class ScaleInvariantSignalNoiseRatio(nn.Module):
    """Running mean over all items seen; forward() returns the batch value (as Metric.forward)."""

    def __init__(self):
        super().__init__()
        self.reset()

    def reset(self):
        self._sum, self._n = 0.0, 0

    def update(self, preds, target):
        v = scale_invariant_signal_noise_ratio(preds, target)
        self._sum = self._sum + v.sum()
        self._n += v.numel()

    def compute(self):
        return self._sum / self._n

    def forward(self, preds, target):
        v = scale_invariant_signal_noise_ratio(preds, target)
        self._sum = self._sum + v.detach().sum()
        self._n += v.numel()
        return v.mean()
