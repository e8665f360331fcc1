"""PositionalEncoding as used by CSE_transformer.py:5,88,103 (sinusoidal table)."""
import math
import torch
import torch.nn as nn


class PositionalEncoding(nn.Module):
    """pe[:, 0::2] = sin(pos * w_i), pe[:, 1::2] = cos(pos * w_i), w_i = exp(-2i ln(1e4)/d).

    forward() ignores the values of x and returns the first x.size(1) rows.
    """

    def __init__(self, input_size, max_len=2500):
        super().__init__()
        if input_size % 2 != 0:
            raise ValueError(f"Cannot use sin/cos positional encoding with odd channels (got channels={input_size})")
        self.max_len = max_len
        table = torch.zeros(max_len, input_size, requires_grad=False)
        pos = torch.arange(0, max_len).unsqueeze(1).float()
        freq = torch.exp(torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size))
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, : x.size(1)].clone().detach()
