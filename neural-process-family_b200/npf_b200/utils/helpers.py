"""Host-side helpers mirroring the pieces of upstream npf/utils/helpers.py that the hot path touches."""
import operator
from functools import reduce

import torch
import torch.nn as nn
from torch.distributions import Normal
from torch.distributions.independent import Independent

__all__ = [
    "MultivariateNormalDiag", "isin_range", "channels_to_2nd_dim", "channels_to_last_dim", "prod",
    "sum_from_nth_dim", "make_abs_conv", "make_depth_sep_conv", "make_padded_conv", "CircularPad2d",
]


def MultivariateNormalDiag(loc, scale_diag):
    """Independent(Normal(loc, scale), 1) -- upstream helpers.py:125-129.  Callers rely on ``.base_dist.loc``,
    ``.base_dist.scale``, ``.log_prob`` and ``.batch_shape`` of this object."""
    if loc.dim() < 1:
        raise ValueError("loc must be at least one-dimensional.")
    return Independent(Normal(loc, scale_diag, validate_args=False), 1)


def isin_range(x, valid_range):
    """upstream helpers.py:55-57 (host-synchronising; the models use the device-side flag instead)."""
    return ((x >= valid_range[0]) & (x <= valid_range[1])).all()


def channels_to_2nd_dim(X):
    return X.permute(0, X.dim() - 1, *range(1, X.dim() - 1))


def channels_to_last_dim(X):
    return X.permute(0, *range(2, X.dim()), 1)


def prod(iterable):
    return reduce(operator.mul, iterable, 1)


def sum_from_nth_dim(t, dim):
    return t.reshape(*t.shape[:dim], -1).sum(-1)


def make_abs_conv(Conv):
    """Convolution whose effective weights are |w| (upstream helpers.py:316-331).  The class only *holds* the
    parameters; GridConvCNP reads ``weight`` and applies abs() inside its CUDA kernel."""

    class AbsConv(Conv):
        _npf_abs = True

    return AbsConv


def make_depth_sep_conv(Conv):
    """Depthwise (groups=in) followed by 1x1 pointwise convolution (upstream helpers.py:354-403).  Parameter
    holder with the upstream sub-module names ``depthwise`` / ``pointwise``; executed by ResConvBlock."""

    class DepthSepConv(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, confidence=False, bias=True, **kwargs):
            super().__init__()
            self.depthwise = Conv(in_channels, in_channels, kernel_size, groups=in_channels, bias=bias, **kwargs)
            self.pointwise = Conv(in_channels, out_channels, 1, bias=bias)

        def reset_parameters(self):
            pass

    return DepthSepConv


def make_padded_conv(Conv, Padder):
    raise NotImplementedError(
        "npf_b200: custom padders (e.g. CircularPad2d for `model_2d_extrap`) are outside the B200 hot path; "
        "the CUDA depthwise kernels implement zero padding only.")


class CircularPad2d(nn.Module):
    def __init__(self, padding):
        super().__init__()
        raise NotImplementedError("npf_b200: circular padding is not implemented by the CUDA depthwise kernels.")
