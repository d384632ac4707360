"""AttnCNP / AttnLNP: per-context representations attended by the targets with the fused cross-attention kernel.
Constructor contract of upstream npf/neuralproc/attnnp.py (``AttnCNP`` 27-131, ``AttnLNP`` 134-202)."""
import torch

from .. import ops
from ..architectures import SelfAttention, get_attender, merge_flat_input
from .base import LatentNeuralProcessFamily, NeuralProcessFamily
from .np import CNP

__all__ = ["AttnCNP", "AttnLNP"]


class AttnCNP(NeuralProcessFamily):
    _valid_paths = ["deterministic"]

    def __init__(self, x_dim, y_dim, XYEncoder=None, attention="scaledot", attention_kwargs={}, self_attention_kwargs={},
                 is_self_attn=False, **kwargs):
        kwargs["encoded_path"] = kwargs.get("encoded_path", "deterministic")
        super().__init__(x_dim, y_dim, **kwargs)
        self.is_self_attn = is_self_attn
        if self.is_self_attn:  # upstream attnnp.py:88-91: the XYEncoder argument is ignored
            XYEncoder = merge_flat_input(SelfAttention, is_sum_merge=True, **self_attention_kwargs)
        elif XYEncoder is None:
            XYEncoder = self.dflt_Modules["XYEncoder"]
        self.xy_encoder = XYEncoder(self.x_transf_dim, self.y_dim, self.r_dim)
        self.attender = get_attender(attention, self.x_transf_dim, self.r_dim, self.r_dim, **attention_kwargs)
        if hasattr(self.x_encoder, "precision"):
            self.x_encoder.precision = "fp32"  # keys / queries of the attention: see MultiheadAttender._attend

    dflt_Modules = CNP.dflt_Modules

    def encode_globally(self, X_cntxt, Y_cntxt):
        B, n_cntxt, _ = X_cntxt.shape
        if n_cntxt == 0:
            return torch.zeros(B, 0, self.r_dim, device=X_cntxt.device)
        return self.xy_encoder(X_cntxt, Y_cntxt)  # [B, C, r]: one representation per context point

    def trgt_dependent_representation(self, X_cntxt, _, R, X_trgt):
        B, n_cntxt, _ = X_cntxt.shape
        if n_cntxt == 0:
            R_trgt = torch.zeros(B, X_trgt.size(1), self.r_dim, device=R.device)
        else:
            R_trgt = self.attender(X_cntxt, X_trgt, R)  # keys, queries, values
        return R_trgt.unsqueeze(0)


class AttnLNP(LatentNeuralProcessFamily, AttnCNP):
    """Deterministic attention path + a global latent inferred from the mean context representation
    (upstream attnnp.py:134-202)."""

    def __init__(self, x_dim, y_dim, encoded_path="both", **kwargs):
        super().__init__(x_dim, y_dim, encoded_path=encoded_path, **kwargs)

    @property
    def dflt_Modules(self):
        d = AttnCNP.dflt_Modules.__get__(self)
        d.update(LatentNeuralProcessFamily.dflt_Modules.__get__(self))
        return d

    def rep_to_lat_input(self, R):
        """One latent per task from the per-context representations: mean over the context set, zeros when the set
        is empty (upstream attnnp.py:174-183)."""
        if R.shape[1] == 0:
            return torch.zeros(R.shape[0], 1, self.r_dim, device=R.device)
        return ops.mean_pool(R)

    def trgt_dependent_representation(self, X_cntxt, z_samples, R, X_trgt):
        B, n_trgt, _ = X_trgt.shape
        n_z = z_samples.size(0)
        if self.encoded_path == "both":
            R_attn = AttnCNP.trgt_dependent_representation(self, X_cntxt, None, R, X_trgt).squeeze(0)  # [B,T,r]
            R_trgt = self.merge_r_z(R_attn, z_samples.expand(n_z, B, n_trgt, self.z_dim))
        else:
            R_trgt = z_samples.expand(n_z, B, n_trgt, self.r_dim)
        return R_trgt
