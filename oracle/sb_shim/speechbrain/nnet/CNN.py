"""Name-only stub: CSE_transformer.py:6 imports Conv1d for the unused '1dcnn' FFN branch."""
import torch.nn as nn


class Conv1d(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("speechbrain.nnet.CNN.Conv1d is outside the hot path")
