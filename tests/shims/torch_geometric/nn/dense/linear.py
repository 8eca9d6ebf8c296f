import math

import torch

from ..inits import glorot


class Linear(torch.nn.Linear):
    """torch_geometric.nn.dense.linear.Linear: weight_initializer 'glorot' or kaiming-uniform default, zero bias."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        self.weight_initializer = weight_initializer
        super().__init__(in_channels, out_channels, bias=bias)

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            glorot(self.weight)
        else:
            torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)
