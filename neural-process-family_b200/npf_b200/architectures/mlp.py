"""MLP whose whole Linear->ReLU chain runs as one autograd node over the CUDA GEMM kernels.

Interface, parameter names (``to_hidden``, ``linears.i``, ``out``) and hidden-size clamping follow upstream
npf/architectures/mlp.py:12-115 so upstream ``state_dict``s load key-for-key.
"""
import warnings

import torch.nn as nn

from .. import ops
from ..utils.initialization import linear_init

__all__ = ["MLP"]


class MLP(nn.Module):
    """``out(relu(linears[-1](... relu(to_hidden(x)))))``.

    Only the configuration the hot path uses is implemented on the GPU: ReLU activation, no dropout, no residual
    connections (upstream defaults).  Anything else raises ``NotImplementedError`` at construction.
    """

    def __init__(self, input_size, output_size, hidden_size=32, n_hidden_layers=1, activation=nn.ReLU(), is_bias=True,
                 dropout=0, is_force_hid_smaller=False, is_res=False):
        super().__init__()
        if not isinstance(activation, nn.ReLU):
            raise NotImplementedError("npf_b200.MLP: only nn.ReLU activations are implemented in the fused kernels")
        if dropout > 0 or is_res:
            raise NotImplementedError("npf_b200.MLP: dropout / residual MLPs are outside the B200 hot path")
        self.input_size, self.output_size = input_size, output_size
        self.n_hidden_layers, self.is_res = n_hidden_layers, is_res
        lo, hi = min(input_size, output_size), max(input_size, output_size)
        if is_force_hid_smaller and hidden_size > hi:  # upstream mlp.py:64-79
            warnings.warn(f"hidden_size={hidden_size} larger than output={output_size} and input={input_size}. Setting it to {hi}.")
            hidden_size = hi
        elif hidden_size < lo:
            warnings.warn(f"hidden_size={hidden_size} smaller than output={output_size} and input={input_size}. Setting it to {lo}.")
            hidden_size = lo
        self.hidden_size = hidden_size
        self.activation = activation
        self.dropout = nn.Identity()
        self.to_hidden = nn.Linear(input_size, hidden_size, bias=is_bias)
        self.linears = nn.ModuleList(nn.Linear(hidden_size, hidden_size, bias=is_bias) for _ in range(n_hidden_layers - 1))
        self.out = nn.Linear(hidden_size, output_size, bias=is_bias)
        self.precision = None  # None: global npf_b200 precision; 'fp32' pins this MLP to the FFMA kernels
        self.reset_parameters()

    def _layers(self):
        return [self.to_hidden, *self.linears, self.out]

    def forward(self, x):
        layers = self._layers()
        return ops.mlp_chain(x, [l.weight for l in layers], [l.bias for l in layers], precision=self.precision)

    def reset_parameters(self):
        for lin in self._layers()[:-1]:
            linear_init(lin, activation=self.activation)
        linear_init(self.out)
