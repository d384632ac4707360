"""SetConv (RBF set convolution) on the CUDA kernels.  Interface and parameter names
(``radial_basis_func.length_scale_param``, ``resizer``) follow upstream npf/architectures/setcnn.py:86-268."""
import math

import torch
import torch.nn as nn

from .. import ops

__all__ = ["SetConv", "ExpRBF"]


class ExpRBF(nn.Module):
    """Holder of the length-scale parameter theta (sigma = 1e-5 + softplus(theta)).  Initialised so that a query at
    ``max_dist`` from a key gets weight ``max_dist_weight`` (upstream setcnn.py:105-124); only p=2 is implemented."""

    def __init__(self, x_dim, max_dist=1 / 256, max_dist_weight=0.9, p=2, **kwargs):
        super().__init__()
        if p != 2:
            raise NotImplementedError("npf_b200.ExpRBF: only the Gaussian (p=2) radial basis is implemented")
        self.max_dist, self.max_dist_weight, self.p = max_dist, max_dist_weight, p
        self.length_scale_param = nn.Parameter(torch.tensor([0.0]))
        self.reset_parameters()

    def reset_parameters(self):
        sigma0 = self.max_dist / math.sqrt(-math.log(self.max_dist_weight))
        with torch.no_grad():
            self.length_scale_param.fill_(math.log(math.expm1(sigma0)))  # inverse softplus

    def forward(self, diff):
        raise RuntimeError("npf_b200.ExpRBF is evaluated inside the fused SetConv kernel, not stand-alone")


class SetConv(nn.Module):
    """{(x_k, v_k)}, {x_q} -> Linear([sum_k softmax_k(-(d/sigma)^2) v_k ; sum_k exp(-(d/sigma)^2)]).

    ``keys`` / ``queries`` are [B, n, 1] tensors, or 1-D tensors shared by the whole batch (the induced grid:
    pass ``keys_regular=True`` to enable the exact sigma-window over an increasing uniform grid)."""

    def __init__(self, x_dim, in_channels, out_channels, RadialBasisFunc=ExpRBF, **kwargs):
        super().__init__()
        assert x_dim == 1, "Currently only supports single spatial dimension `x_dim==1`"
        if RadialBasisFunc is not ExpRBF:
            raise NotImplementedError("npf_b200.SetConv: only ExpRBF is implemented in the CUDA kernels")
        self.radial_basis_func = RadialBasisFunc(x_dim, **kwargs)
        self.resizer = nn.Linear(in_channels + 1, out_channels)

    def reset_parameters(self):
        pass

    def forward(self, keys, queries, values, keys_regular=False):
        return ops.setconv(keys, queries, values, self.radial_basis_func.length_scale_param, self.resizer.weight,
                           self.resizer.bias, keys_regular=keys_regular)
