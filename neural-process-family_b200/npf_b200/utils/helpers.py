"""Host-side helpers mirroring the pieces of upstream npf/utils/helpers.py that the hot path touches."""
import operator
from functools import reduce

import torch
import torch.nn as nn
from torch.distributions import Normal
from torch.distributions.independent import Independent

__all__ = [
    "MultivariateNormalDiag", "isin_range", "channels_to_2nd_dim", "channels_to_last_dim", "prod",
    "sum_from_nth_dim", "make_abs_conv", "make_depth_sep_conv", "make_padded_conv", "CircularPad2d", "conv_padding",
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
    """Convolution preceded by an arbitrary padder (upstream helpers.py:334-351): ``padding`` goes to ``Padder(padding)``
    and the convolution itself is built unpadded; ``Padder=None`` keeps the plain zero padding.  Parameter holder like
    the other conv classes here: ResConvBlock / GridConvCNP read ``padder`` and run pad -> CUDA kernel -> crop."""

    class PaddedConv(Conv):
        _npf_padded = True

        def __init__(self, *args, Padder=Padder, padding=0, **kwargs):
            old_padding = 0
            if Padder is None:
                Padder = nn.Identity
                old_padding = padding
            super().__init__(*args, padding=old_padding, **kwargs)
            self.padder = Padder(padding)

    return PaddedConv


class CircularPad2d(nn.Module):
    """Wrap-around padding of both grid axes (upstream helpers.py:406-414, ``F.pad(mode="circular")``).  ``forward``
    takes the CHANNEL-LAST signal the kernels work on, [B, H, W, C], and returns [B, H+2p, W+2p, C]."""

    def __init__(self, padding):
        super().__init__()
        self.padding = int(padding)

    def forward(self, x):
        p = self.padding
        if p == 0:
            return x
        if x.dim() != 4 or p > x.shape[1] or p > x.shape[2]:
            raise ValueError(f"CircularPad2d({p}): expected a channel-last [B,H,W,C] signal with H, W >= {p}, got {tuple(x.shape)}")
        x = torch.cat([x[:, -p:], x, x[:, :p]], dim=1)
        return torch.cat([x[:, :, -p:], x, x[:, :, :p]], dim=2)


def conv_padding(conv):
    """(padder, p): how a conv held by the models is padded.  Plain zero-padded conv -> (None, k // 2) after checking
    that it is 'same' padding; ``make_padded_conv`` conv -> (its CircularPad2d, p) or (None, ...) for ``Padder=None``."""
    k = conv.kernel_size[0]
    padder = getattr(conv, "padder", None)
    if padder is None or isinstance(padder, nn.Identity):
        if any(q != k // 2 for q in conv.padding):
            raise NotImplementedError("npf_b200: convolutions must use 'same' padding (kernel_size // 2)")
        return None, k // 2
    if not isinstance(padder, CircularPad2d) or any(q != 0 for q in conv.padding):
        raise NotImplementedError("npf_b200: the only custom padder implemented is CircularPad2d")
    if padder.padding != k // 2:
        raise NotImplementedError("npf_b200: CircularPad2d must pad by kernel_size // 2 ('same' output size)")
    return (padder if padder.padding > 0 else None), padder.padding
