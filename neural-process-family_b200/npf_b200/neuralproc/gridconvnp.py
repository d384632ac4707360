"""GridConvCNP / GridConvLNP (on-grid, images): masked abs-conv encoding -> depthwise-separable CNN -> per-pixel MLP.
Constructor / argument contract of upstream npf/neuralproc/gridconvnp.py (``GridConvCNP`` 28-178, ``GridConvLNP``
181-289): ``X_cntxt`` is the context MASK [B,*grid,1], ``Y_cntxt`` the full image [B,*grid,y], ``X_trgt`` the target
mask; every pixel is predicted."""
import logging
from functools import partial

import torch.nn as nn

from .. import ops
from ..architectures import CNN, ResConvBlock
from ..utils.helpers import conv_padding, make_abs_conv
from .base import LatentNeuralProcessFamily, NeuralProcessFamily
from .convnp import ConvCNP, ConvLNP
from .helpers import collapse_z_samples_batch

__all__ = ["GridConvCNP", "GridConvLNP"]
logger = logging.getLogger(__name__)


def _default_abs_conv(y_dim):
    return make_abs_conv(nn.Conv2d)(y_dim, y_dim, groups=y_dim, kernel_size=11, padding=11 // 2, bias=False)


class GridConvCNP(NeuralProcessFamily):
    _valid_paths = ["deterministic"]

    def __init__(self, x_dim, y_dim, Conv=_default_abs_conv,
                 CNN=partial(CNN, ConvBlock=ResConvBlock, Conv=nn.Conv2d, n_blocks=3, Normalization=nn.Identity,
                             is_chan_last=True, kernel_size=11),
                 **kwargs):
        assert x_dim == 1 or x_dim == y_dim, "Ensure that featrue masks can be multiplied with Y"
        if "Decoder" in kwargs and kwargs["Decoder"] != nn.Identity:
            logger.warning("`Decoder` was given to `GridConvCNP`. To be translation equivariant it should disregard its "
                           "first argument, e.g. `discard_ith_arg(Decoder, i=0)` (the default).")
        kwargs["encoded_path"] = kwargs.get("encoded_path", "deterministic")
        super().__init__(x_dim, y_dim, x_transf_dim=None, XEncoder=nn.Identity, **kwargs)
        self.CNN = CNN
        self.conv = Conv(y_dim)
        c = self.conv
        ok = (isinstance(c, nn.Conv2d) and getattr(c, "_npf_abs", False) and c.groups == y_dim and c.bias is None
              and c.kernel_size[0] == c.kernel_size[1] and c.padding_mode == "zeros")
        if not ok:
            raise NotImplementedError("npf_b200.GridConvCNP: `Conv` must be a bias-free depthwise `make_abs_conv(nn.Conv2d)` "
                                      "(the upstream default), zero-padded or wrapped by make_padded_conv(.., CircularPad2d)")
        conv_padding(c)
        self.resizer = nn.Linear(self.y_dim * 2, self.r_dim)  # signal + confidence channels
        self.induced_to_induced = CNN(self.r_dim)

    dflt_Modules = ConvCNP.dflt_Modules

    def cntxt_to_induced(self, mask_cntxt, X):
        """[signal / max(density, 1e-5) ; density] of the masked image under the |w| filter, resized to r_dim."""
        padder, p = conv_padding(self.conv)
        if padder is None:
            feat = ops.gridconv_in(X, mask_cntxt, self.conv.weight)
        else:   # wrap-around first layer of `model_2d_extrap`: extend image and mask, same kernel, crop
            feat = ops.gridconv_in(padder(X).contiguous(), padder(mask_cntxt).contiguous(), self.conv.weight)[:, p:-p, p:-p, :].contiguous()
        return ops.linear(feat, self.resizer.weight, self.resizer.bias)

    def encode_globally(self, mask_cntxt, X):
        return self.induced_to_induced(self.cntxt_to_induced(mask_cntxt, X))

    def trgt_dependent_representation(self, _, __, R_induced, ___):
        return R_induced.unsqueeze(0)

    def set_extrapolation(self, min_max):
        raise NotImplementedError("GridConvCNP cannot be used for extrapolation.")


class GridConvLNP(LatentNeuralProcessFamily, GridConvCNP):
    _valid_paths = ["latent", "both"]

    def __init__(self, x_dim, y_dim, CNNPostZ=None, encoded_path="latent", is_global=False, **kwargs):
        super().__init__(x_dim, y_dim, encoded_path=encoded_path, **kwargs)
        self.is_global = is_global
        if CNNPostZ is None:
            CNNPostZ = self.CNN
        self.induced_to_induced_post_sampling = CNNPostZ(self.r_dim)

    dflt_Modules = ConvLNP.dflt_Modules
    add_global_latent = ConvLNP.add_global_latent
    rep_to_lat_input = ConvLNP.rep_to_lat_input

    def trgt_dependent_representation(self, X_cntxt, z_samples, R_induced, X_trgt):
        B, *grid_shape, _ = X_trgt.shape
        n_z = z_samples.size(0)
        if self.encoded_path == "latent":
            z = collapse_z_samples_batch(z_samples)  # [n_z*B, *grid, z]
            if self.is_global:  # on the grid the pooling comes BEFORE the CNN (upstream gridconvnp.py:255-257)
                z = self.add_global_latent(z)
            if self.z_dim != self.r_dim:
                z = ops.linear(z, self.reshaper_z.weight, self.reshaper_z.bias)
            R_trgt = self.induced_to_induced_post_sampling(z)
        else:
            z = z_samples.view(n_z, B, *([1] * len(grid_shape)), self.r_dim).expand(n_z, B, *grid_shape, self.r_dim)
            R_trgt = self.induced_to_induced_post_sampling(collapse_z_samples_batch(self.merge_r_z(R_induced, z)))
        return R_trgt.view(n_z, B, *grid_shape, self.r_dim)
