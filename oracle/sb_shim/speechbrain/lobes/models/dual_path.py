"""speechbrain.lobes.models.dual_path stand-in (test infrastructure; see sb_shim/README.md).

The reference does `from speechbrain.lobes.models.dual_path import *` (ContSep.py:4,
ContExt.py:4), which must supply: copy, select_norm, Encoder, Decoder, SBRNNBlock, Linear,
PositionalEncoding; sepformer.py:4 additionally needs Dual_Path_Model.
"""
import copy  # noqa: F401  (re-exported through `import *`)
import math  # noqa: F401

import torch
import torch.nn as nn
import torch.nn.functional as F

from speechbrain.lobes.models.transformer.Transformer import PositionalEncoding  # noqa: F401
from speechbrain.nnet.linear import Linear  # noqa: F401

EPS = 1e-8


def select_norm(norm, dim, shape, eps=1e-8):
    """'ln' -> GroupNorm(1, dim): one mean/var per sample over all (channel, position)."""
    if norm == "ln":
        return nn.GroupNorm(1, dim, eps=eps)
    if norm in ("gln", "cln"):
        raise NotImplementedError(f"norm={norm!r} is not used by the reference models")
    return nn.BatchNorm1d(dim)


class Encoder(nn.Module):
    """Conv1d(in, out, k, stride=k//2, bias=False) on x.unsqueeze(1), then ReLU."""

    def __init__(self, kernel_size=2, out_channels=64, in_channels=1):
        super().__init__()
        self.conv1d = nn.Conv1d(
            in_channels=in_channels,
            out_channels=out_channels,
            kernel_size=kernel_size,
            stride=kernel_size // 2,
            groups=1,
            bias=False,
        )
        self.in_channels = in_channels

    def forward(self, x):
        if self.in_channels == 1:
            x = torch.unsqueeze(x, dim=1)
        return F.relu(self.conv1d(x))


class Decoder(nn.ConvTranspose1d):
    """ConvTranspose1d whose forward accepts [B,N,L] (or [N,L]) and squeezes the channel dim."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def forward(self, x):
        if x.dim() not in [2, 3]:
            raise RuntimeError("{} accept 3/4D tensor as input".format(type(self).__name__))
        x = super().forward(x if x.dim() == 3 else torch.unsqueeze(x, 1))
        if torch.squeeze(x).dim() == 1:
            x = torch.squeeze(x, dim=1)
        else:
            x = torch.squeeze(x)
        return x


class SBRNNBlock(nn.Module):  # only used in isinstance() checks (ContSep.py:428,437)
    def __init__(self, *a, **k):
        raise NotImplementedError("SBRNNBlock is outside the hot path")


class SBTransformerBlock(nn.Module):  # imported by CSE_transformer.py:9, never instantiated
    def __init__(self, *a, **k):
        raise NotImplementedError("use SBTransformerBlock_CSE")


class Dual_Computation_Block(nn.Module):
    """Stock dual-path block = the reference's _CSE fork without the ##ADD context lines."""

    def __init__(self, intra_mdl, inter_mdl, out_channels, norm="ln",
                 skip_around_intra=True, linear_layer_after_inter_intra=True):
        super().__init__()
        self.intra_mdl = intra_mdl
        self.inter_mdl = inter_mdl
        self.skip_around_intra = skip_around_intra
        self.linear_layer_after_inter_intra = linear_layer_after_inter_intra
        self.norm = norm
        if norm is not None:
            self.intra_norm = select_norm(norm, out_channels, 4)
            self.inter_norm = select_norm(norm, out_channels, 4)
        if linear_layer_after_inter_intra:
            self.intra_linear = Linear(out_channels, input_size=out_channels)
            self.inter_linear = Linear(out_channels, input_size=out_channels)

    def forward(self, x):
        B, N, K, S = x.shape
        y = x.permute(0, 3, 2, 1).contiguous().view(B * S, K, N)
        y = self.intra_mdl(y)
        if self.linear_layer_after_inter_intra:
            y = self.intra_linear(y)
        y = y.view(B, S, K, N).permute(0, 3, 2, 1).contiguous()
        if self.norm is not None:
            y = self.intra_norm(y)
        if self.skip_around_intra:
            y = y + x
        z = y.permute(0, 2, 3, 1).contiguous().view(B * K, S, N)
        z = self.inter_mdl(z)
        if self.linear_layer_after_inter_intra:
            z = self.inter_linear(z)
        z = z.view(B, K, S, N).permute(0, 3, 1, 2).contiguous()
        if self.norm is not None:
            z = self.inter_norm(z)
        return z + y


class Dual_Path_Model(nn.Module):
    """Stock dual-path mask estimator (sepformer.py:11): norm -> 1x1 -> chunk -> blocks ->
    PReLU -> 1x1 (x spk) -> overlap-add -> tanh*sigmoid gate -> 1x1 -> ReLU."""

    def __init__(self, in_channels, out_channels, intra_model, inter_model, num_layers=1,
                 norm="ln", K=200, num_spks=2, skip_around_intra=True,
                 linear_layer_after_inter_intra=True, use_global_pos_enc=False, max_length=20000):
        super().__init__()
        self.K = K
        self.num_spks = num_spks
        self.num_layers = num_layers
        self.norm = select_norm(norm, in_channels, 3)
        self.conv1d = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.use_global_pos_enc = use_global_pos_enc
        if use_global_pos_enc:
            self.pos_enc = PositionalEncoding(max_length)
        self.dual_mdl = nn.ModuleList(
            [
                copy.deepcopy(
                    Dual_Computation_Block(
                        intra_model, inter_model, out_channels, norm,
                        skip_around_intra=skip_around_intra,
                        linear_layer_after_inter_intra=linear_layer_after_inter_intra,
                    )
                )
                for _ in range(num_layers)
            ]
        )
        self.conv2d = nn.Conv2d(out_channels, out_channels * num_spks, kernel_size=1)
        self.end_conv1x1 = nn.Conv1d(out_channels, in_channels, 1, bias=False)
        self.prelu = nn.PReLU()
        self.activation = nn.ReLU()
        self.output = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Tanh())
        self.output_gate = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Sigmoid())

    def forward(self, x):
        x = self.conv1d(self.norm(x))
        if self.use_global_pos_enc:
            x = self.pos_enc(x.transpose(1, -1)).transpose(1, -1) + x * (x.size(1) ** 0.5)
        x, gap = self._Segmentation(x, self.K)
        for blk in self.dual_mdl:
            x = blk(x)
        x = self.conv2d(self.prelu(x))
        B, _, K, S = x.shape
        x = x.view(B * self.num_spks, -1, K, S)
        x = self._over_add(x, gap)
        x = self.output(x) * self.output_gate(x)
        x = self.end_conv1x1(x)
        _, N, L = x.shape
        x = self.activation(x.view(B, self.num_spks, N, L))
        return x.transpose(0, 1)

    def _padding(self, x, K):
        B, N, L = x.shape
        P = K // 2
        gap = K - (P + L % K) % K
        if gap > 0:
            x = torch.cat([x, x.new_zeros(B, N, gap)], dim=2)
        edge = x.new_zeros(B, N, P)
        return torch.cat([edge, x, edge], dim=2), gap

    def _Segmentation(self, x, K):
        B, N, L = x.shape
        P = K // 2
        x, gap = self._padding(x, K)
        a = x[:, :, :-P].contiguous().view(B, N, -1, K)
        b = x[:, :, P:].contiguous().view(B, N, -1, K)
        x = torch.cat([a, b], dim=3).view(B, N, -1, K).transpose(2, 3)
        return x.contiguous(), gap

    def _over_add(self, x, gap):
        B, N, K, S = x.shape
        P = K // 2
        x = x.transpose(2, 3).contiguous().view(B, N, -1, K * 2)
        a = x[:, :, :, :K].contiguous().view(B, N, -1)[:, :, P:]
        b = x[:, :, :, K:].contiguous().view(B, N, -1)[:, :, :-P]
        x = a + b
        if gap > 0:
            x = x[:, :, :-gap]
        return x
