"""Cross-attention attenders on the fused (flash-style) CUDA attention kernel.

Implements the three mechanisms the upstream notebooks / defaults use -- ``"scaledot"``, ``"multihead"`` and
``"transformer"`` -- with the interface and parameter names of upstream npf/architectures/attention.py
(``get_attender`` 16-86, ``DotAttender`` 180-220, ``MultiheadAttender`` 375-527, ``TransformerAttender`` 530-588).
The other six scoring functions of that file are never constructed by any upstream config and raise here.
"""
import math

import torch.nn as nn

from .. import ops
from .mlp import MLP

__all__ = ["get_attender", "DotAttender", "MultiheadAttender", "TransformerAttender"]


def get_attender(attention, kq_size, value_size, out_size, **kwargs):
    if not isinstance(attention, str):
        return attention(kq_size, value_size, out_size, **kwargs)
    name = attention.lower()
    if name == "scaledot":
        return DotAttender(kq_size, value_size, out_size, is_scale=True, **kwargs)
    if name == "multihead":
        return MultiheadAttender(kq_size, value_size, out_size, **kwargs)
    if name == "transformer":
        return TransformerAttender(kq_size, value_size, out_size, **kwargs)
    if name in ("multiplicative", "additive", "cosine", "manhattan", "euclidean", "weighted_dist"):
        raise NotImplementedError(f"npf_b200: attention='{name}' is not on the B200 hot path (no upstream config uses it)")
    raise ValueError("Unknown attention method {}".format(attention))


class DotAttender(nn.Module):
    """softmax_k(q.k / sqrt(d)) @ values, single head (upstream BaseAttender.forward + DotAttender.score)."""

    def __init__(self, kq_size, value_size, out_size, is_scale=True, is_normalize=True, dropout=0):
        super().__init__()
        if not is_normalize or dropout > 0:
            raise NotImplementedError("npf_b200.DotAttender: un-normalised attention / dropout are not implemented")
        self.kq_size, self.value_size, self.out_size, self.is_scale = kq_size, value_size, out_size, is_scale
        self.is_resize = value_size != out_size
        self.dropout = nn.Identity()
        if self.is_resize:
            self.resizer = nn.Linear(value_size, out_size)

    def reset_parameters(self):
        pass

    def forward(self, keys, queries, values, n_heads=1):
        d = queries.shape[-1] // n_heads
        ctx = ops.xattn(queries, keys, values, n_heads, 1.0 / math.sqrt(d) if self.is_scale else 1.0)
        if self.is_resize:
            ctx = ops.linear(ctx, self.resizer.weight, self.resizer.bias)
        return ctx


class MultiheadAttender(nn.Module):
    """K = Wk k, Q = Wq q + bq, V = Wv v; per-head scaled dot-product (head h = channels [h*d, (h+1)*d), i.e. the
    same split as upstream ``_make_multiheaded`` without materialising the permuted copies); optional output Linear."""

    def __init__(self, kq_size, value_size, out_size, n_heads=8, is_post_process=True, dropout=0, is_relative_pos=False):
        super().__init__()
        if is_relative_pos or dropout > 0:
            raise NotImplementedError("npf_b200.MultiheadAttender: relative positions / dropout are not implemented")
        assert kq_size % n_heads == 0, "{} % {} != 0".format(kq_size, n_heads)
        assert value_size % n_heads == 0, "{} % {} != 0".format(value_size, n_heads)
        self.is_relative_pos = False
        self.key_transform = nn.Linear(kq_size, kq_size, bias=False)
        self.query_transform = nn.Linear(kq_size, kq_size, bias=True)
        self.value_transform = nn.Linear(value_size, value_size, bias=False)
        self.dot = DotAttender(kq_size, value_size, out_size, is_scale=True)
        self.n_heads = n_heads
        self.kq_head_size = kq_size // n_heads
        self.value_head_size = kq_size // n_heads  # sic: upstream attention.py:432
        self.kq_size, self.value_size, self.out_size = kq_size, value_size, out_size
        self.post_processor = nn.Linear(value_size, out_size) if (is_post_process or value_size != out_size) else None
        self.reset_parameters()

    def reset_parameters(self):
        # upstream attention.py:446-455: the effective fan-out of a head is head_size, not kq_size
        std = math.sqrt(2.0 / (self.kq_size + self.kq_head_size))
        nn.init.normal_(self.key_transform.weight, mean=0, std=std)
        nn.init.normal_(self.query_transform.weight, mean=0, std=std)
        std = math.sqrt(2.0 / (self.value_size + self.value_head_size))
        nn.init.normal_(self.value_transform.weight, mean=0, std=std)

    def _attend(self, keys, queries, values):
        # everything upstream of the softmax logits stays in fp32: an absolute logit error is a relative error of the
        # attention weight, and trained attention has |logit| ~ 1e2 (the 16-bit products of 'bf16x3' are not enough)
        k = ops.linear(keys, self.key_transform.weight, precision="fp32")
        q = ops.linear(queries, self.query_transform.weight, self.query_transform.bias, precision="fp32")
        v = ops.linear(values, self.value_transform.weight)
        return ops.xattn(q, k, v, self.n_heads, 1.0 / math.sqrt(self.kq_head_size))

    def forward(self, keys, queries, values, rel_pos_enc=None, **kwargs):
        ctx = self._attend(keys, queries, values)
        if self.post_processor is not None:
            ctx = ops.linear(ctx, self.post_processor.weight, self.post_processor.bias)
        return ctx


class TransformerAttender(MultiheadAttender):
    """Multi-head attention + residual/LayerNorm + position-wise MLP + residual/LayerNorm (upstream 569-588)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, is_post_process=False, **kwargs)
        assert self.kq_size == self.out_size
        self.layer_norm1 = nn.LayerNorm(self.out_size)
        self.layer_norm2 = nn.LayerNorm(self.out_size)
        self.mlp = MLP(self.out_size, self.out_size, hidden_size=self.out_size, activation=nn.ReLU())

    def forward(self, keys, queries, values, **kwargs):
        ctx = super().forward(keys, queries, values)
        ctx = ops.add_layernorm(ctx, queries, self.layer_norm1.weight, self.layer_norm1.bias)
        return ops.add_layernorm(ctx, self.mlp(ctx), self.layer_norm2.weight, self.layer_norm2.bias)
