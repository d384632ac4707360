"""Self-attention context encoder: a stack of attenders applied with keys = queries = values.

Interface and sub-module names (``attn_layers``, ``resize``) of upstream npf/architectures/selfattn.py:10-100, which
the 2-D AttnCNP / AttnLNP notebooks reach through ``is_self_attn=True`` (attnnp.py:88-91).  Every layer runs the same
fused attention kernel as the cross-attention path (``npf_xattn_{fwd,bwd}``) with the context set on both sides.
Positional encodings (``positional="absolute"/"relative"``) are not used by any upstream config and raise here.
"""
import torch.nn as nn

from .. import ops
from .attention import get_attender

__all__ = ["SelfAttention"]


class SelfAttention(nn.Module):
    def __init__(self, x_dim, out_dim=None, n_attn_layers=2, attention="transformer", positional=None,
                 position_dim=None, max_len=2000, **kwargs):
        super().__init__()
        if positional in ("absolute", "relative"):
            raise NotImplementedError("npf_b200.SelfAttention: positional encodings are not implemented "
                                      "(no upstream config uses them)")
        if positional is not None:
            raise ValueError("Unknown positional={}.".format(positional))
        self.positional = None
        self.attn_layers = nn.ModuleList(
            [get_attender(attention, x_dim, x_dim, x_dim, **kwargs) for _ in range(n_attn_layers)])
        self.is_resize = out_dim is not None
        if self.is_resize:
            self.resize = nn.Linear(x_dim, out_dim)

    def reset_parameters(self):
        pass

    def forward(self, X, positions=None):
        out = X
        for attn_layer in self.attn_layers:
            out = attn_layer(out, out, out)  # keys, queries, values (upstream selfattn.py:94-95)
        if self.is_resize:
            out = ops.linear(out, self.resize.weight, self.resize.bias)
        return out
