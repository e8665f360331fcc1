"""speechbrain.nnet.attention pieces touched by CSE_transformer.py:322,335."""
import torch.nn as nn


class PositionalwiseFeedForward(nn.Module):
    """Linear -> act -> Dropout -> Linear, applied between two (1,0,2) permutes.

    Parameter keys: ffn.0.{weight,bias}, ffn.3.{weight,bias}.
    """

    def __init__(self, d_ffn, input_shape=None, input_size=None, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        if input_shape is None and input_size is None:
            raise ValueError("Expected one of input_shape or input_size")
        if input_size is None:
            input_size = input_shape[-1]
        self.ffn = nn.Sequential(
            nn.Linear(input_size, d_ffn),
            activation(),
            nn.Dropout(dropout),
            nn.Linear(d_ffn, input_size),
        )

    def forward(self, x):
        x = x.permute(1, 0, 2)
        x = self.ffn(x)
        return x.permute(1, 0, 2)


class RelPosMHAXL(nn.Module):  # name only; branch unused (CSE_transformer.py:321-324)
    def __init__(self, *a, **k):
        raise NotImplementedError("RelPosMHAXL is outside the hot path")
