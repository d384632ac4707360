"""ConvCNP / ConvLNP (off-grid, 1-D): SetConv -> depthwise-separable CNN -> SetConv -> MLP on the B200 kernels.
Constructor / attribute contract of upstream npf/neuralproc/convnp.py (``ConvCNP`` 26-181, ``ConvLNP`` 184-335)."""
import logging
from functools import partial

import torch
import torch.nn as nn

from .. import ops
from ..architectures import CNN, ResConvBlock, SetConv, discard_ith_arg
from .base import LatentNeuralProcessFamily, NeuralProcessFamily
from .helpers import collapse_z_samples_batch

logger = logging.getLogger(__name__)

__all__ = ["ConvCNP", "ConvLNP"]


class ConvCNP(NeuralProcessFamily):
    _valid_paths = ["deterministic"]

    def __init__(self, x_dim, y_dim, density_induced=128, Interpolator=SetConv,
                 CNN=partial(CNN, ConvBlock=ResConvBlock, Conv=nn.Conv1d, n_blocks=3, Normalization=nn.Identity,
                             is_chan_last=True, kernel_size=11),
                 **kwargs):
        if "Decoder" in kwargs and kwargs["Decoder"] != nn.Identity:
            logger.warning("`Decoder` was given to `ConvCNP`. To be translation equivariant you should disregard the "
                           "first argument, e.g. `discard_ith_arg(Decoder, i=0)` (the default).")
        kwargs["encoded_path"] = kwargs.get("encoded_path", "deterministic")
        super().__init__(x_dim, y_dim, x_transf_dim=None, XEncoder=nn.Identity, **kwargs)
        self.density_induced = density_induced
        # [-1, 1] plus half a unit of margin on each side against boundary effects (upstream convnp.py:102-104)
        self.X_induced = torch.linspace(-1.5, 1.5, int(self.density_induced * 3))
        self.CNN = CNN
        self.cntxt_to_induced = Interpolator(self.x_dim, self.y_dim, self.r_dim)
        self.induced_to_induced = CNN(self.r_dim)
        self.induced_to_trgt = Interpolator(self.x_dim, self.r_dim, self.r_dim)

    @property
    def n_induced(self):
        return len(self.X_induced)

    @property
    def dflt_Modules(self):
        d = NeuralProcessFamily.dflt_Modules.__get__(self)
        d["Decoder"] = discard_ith_arg(d["SubDecoder"], i=0)  # the decoder must not see x (equivariance)
        return d

    def _grid(self, device):
        """The induced grid as a 1-D device tensor (moved once, like upstream's lazy ``.to``)."""
        if self.X_induced.device != device:
            self.X_induced = self.X_induced.to(device)
        return self.X_induced

    def _get_X_induced(self, X):
        return self._grid(X.device).view(1, -1, 1).expand(X.shape[0], self.n_induced, self.x_dim)

    def _interp(self, module, keys, queries, values, grid_is_keys):
        """Run an Interpolator; our SetConv takes the shared grid as a 1-D tensor (no per-task copies) and, when the
        grid is on the key side, enables the exact sigma-window."""
        if isinstance(module, SetConv):
            return module(keys, queries, values, keys_regular=grid_is_keys)
        B = values.shape[0]
        expand = lambda g: g.view(1, -1, 1).expand(B, g.numel(), self.x_dim)
        if grid_is_keys:
            return module(expand(keys), queries, values)
        return module(keys, expand(queries), values)

    def encode_globally(self, X_cntxt, Y_cntxt):
        B, n_cntxt, _ = X_cntxt.shape
        grid = self._grid(X_cntxt.device)
        if n_cntxt == 0:
            # no context: zero functional representation (density channel included), then the CNN (upstream 146-151)
            R_induced = torch.zeros(B, self.n_induced, self.r_dim, device=X_cntxt.device)
        else:
            R_induced = self._interp(self.cntxt_to_induced, X_cntxt, grid, Y_cntxt, grid_is_keys=False)
        return self.induced_to_induced(R_induced)

    def trgt_dependent_representation(self, X_cntxt, z_samples, R_induced, X_trgt):
        grid = self._grid(X_trgt.device)
        R_trgt = self._interp(self.induced_to_trgt, grid, X_trgt, R_induced, grid_is_keys=True)
        return R_trgt.unsqueeze(0)

    def set_extrapolation(self, min_max):
        """Re-grid the induced points over [min-0.5, max+0.5] at the training density (upstream convnp.py:170-181)."""
        lo, hi = min_max[0] - 0.5, min_max[1] + 0.5
        self.X_induced = torch.linspace(lo, hi, int(self.density_induced * (hi - lo)))


class ConvLNP(LatentNeuralProcessFamily, ConvCNP):
    _valid_paths = ["latent", "both"]

    def __init__(self, x_dim, y_dim, CNNPostZ=None, encoded_path="latent", is_global=False, **kwargs):
        super().__init__(x_dim, y_dim, encoded_path=encoded_path, **kwargs)
        self.is_global = is_global
        if CNNPostZ is None:
            CNNPostZ = self.CNN
        self.induced_to_induced_post_sampling = CNNPostZ(self.r_dim)

    @property
    def dflt_Modules(self):
        d = ConvCNP.dflt_Modules.__get__(self)
        d.update(LatentNeuralProcessFamily.dflt_Modules.__get__(self))
        d["Decoder"] = discard_ith_arg(nn.Linear, i=0)  # small decoder: the post-sampling CNN did the work
        return d

    def rep_to_lat_input(self, R):
        if self.encoded_path == "latent":
            return R
        return ops.mean_pool(R.reshape(R.shape[0], -1, self.r_dim))

    def add_global_latent(self, z_samples):
        return ops.global_latent(z_samples)

    def trgt_dependent_representation(self, X_cntxt, z_samples, R_induced, X_trgt):
        B, n_trgt, _ = X_trgt.shape
        n_z = z_samples.shape[0]
        grid = self._grid(X_trgt.device)
        if self.encoded_path == "latent":
            z = collapse_z_samples_batch(z_samples)  # [n_z*B, I, z]
            if self.z_dim != self.r_dim:
                z = ops.linear(z, self.reshaper_z.weight, self.reshaper_z.bias)
            R = self.induced_to_induced_post_sampling(z)
            if self.is_global:  # upstream applies it AFTER the CNN in the off-grid model (convnp.py:290-293)
                R = self.add_global_latent(R)
        else:
            z = z_samples.expand(n_z, B, self.n_induced, self.z_dim)
            R = self.induced_to_induced_post_sampling(collapse_z_samples_batch(self.merge_r_z(R_induced, z)))
        X_t = X_trgt.unsqueeze(0).expand(n_z, B, n_trgt, self.x_dim).reshape(n_z * B, n_trgt, self.x_dim)
        R_trgt = self._interp(self.induced_to_trgt, grid, X_t, R, grid_is_keys=True)
        return R_trgt.view(n_z, B, n_trgt, self.r_dim)
