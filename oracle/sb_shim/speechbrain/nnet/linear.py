"""speechbrain.nnet.linear.Linear (only reachable with linear_layer_after_inter_intra=True)."""
import torch.nn as nn


class Linear(nn.Module):
    def __init__(self, n_neurons, input_shape=None, input_size=None, bias=True, combine_dims=False):
        super().__init__()
        if input_size is None:
            input_size = input_shape[-1]
        self.w = nn.Linear(input_size, n_neurons, bias=bias)

    def forward(self, x):
        return self.w(x)
