"""CNP (DeepSet encoder + mean aggregation + MLP decoder) and LNP on the B200 kernels.
Constructor / attribute contract of upstream npf/neuralproc/np.py (``CNP`` 19-110, ``LNP`` 113-163)."""
from functools import partial

import torch

from .. import ops
from ..architectures import MLP, merge_flat_input
from .base import LatentNeuralProcessFamily, NeuralProcessFamily

__all__ = ["CNP", "LNP"]


class CNP(NeuralProcessFamily):
    _valid_paths = ["deterministic"]

    def __init__(self, x_dim, y_dim, XYEncoder=None, **kwargs):
        kwargs["encoded_path"] = kwargs.get("encoded_path", "deterministic")
        super().__init__(x_dim, y_dim, **kwargs)
        if XYEncoder is None:
            XYEncoder = self.dflt_Modules["XYEncoder"]
        self.xy_encoder = XYEncoder(self.x_transf_dim, self.y_dim, self.r_dim)

    @property
    def dflt_Modules(self):
        d = NeuralProcessFamily.dflt_Modules.__get__(self)
        sub = partial(MLP, n_hidden_layers=2, is_force_hid_smaller=True, hidden_size=self.r_dim)
        d["XYEncoder"] = merge_flat_input(sub, is_sum_merge=True)
        return d

    def encode_globally(self, X_cntxt, Y_cntxt):
        """R = mean_c xy_encoder(x_c, y_c), one vector per task; zeros without context (upstream np.py:86-101)."""
        B, n_cntxt, _ = X_cntxt.shape
        if n_cntxt == 0:
            return torch.zeros(B, 1, self.r_dim, device=X_cntxt.device)
        return ops.mean_pool(self.xy_encoder(X_cntxt, Y_cntxt))

    def trgt_dependent_representation(self, _, __, R, X_trgt):
        B, n_trgt, _ = X_trgt.shape
        return R.expand(B, n_trgt, self.r_dim).unsqueeze(0)  # a broadcast view: [1, B, T, r]


class LNP(LatentNeuralProcessFamily, CNP):
    def __init__(self, x_dim, y_dim, encoded_path="latent", **kwargs):
        super().__init__(x_dim, y_dim, encoded_path=encoded_path, **kwargs)

    def trgt_dependent_representation(self, _, z_samples, R, X_trgt):
        B, n_trgt, _ = X_trgt.shape
        n_z = z_samples.size(0)
        if self.encoded_path == "both":
            R_trgt = self.merge_r_z(R, z_samples)
        else:
            R_trgt = z_samples
            if self.z_dim != self.r_dim:
                R_trgt = ops.linear(R_trgt, self.reshaper_z.weight, self.reshaper_z.bias)
        return R_trgt.expand(n_z, B, n_trgt, self.r_dim)
