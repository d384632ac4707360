"""Two-input wrappers of the factory protocol: ``merge_flat_input`` (sum-merge) and ``discard_ith_arg``.

Mirrors the interface of upstream npf/architectures/encoders.py:103-213 (same sub-module names ``resizer``,
``flat_module``, ``destination``); the positional-encoding classes of that file are out of scope.
"""
import torch.nn as nn

from .. import ops
from .mlp import MLP

__all__ = ["merge_flat_input", "discard_ith_arg", "MergeFlatInputs", "DiscardIthArg"]


class DiscardIthArg(nn.Module):
    """Drops the i-th positional argument of constructor and forward before delegating to ``To``."""

    def __init__(self, *args, i=0, To=nn.Identity, **kwargs):
        super().__init__()
        self.i = i
        self.destination = To(*self._keep(args), **kwargs)

    def _keep(self, args):
        return [a for j, a in enumerate(args) if j != self.i]

    def forward(self, *args, **kwargs):
        dest = self.destination
        kept = self._keep(args)
        if isinstance(dest, nn.Linear):  # bare Linear decoder of the ConvLNP family
            return ops.linear(kept[0], dest.weight, dest.bias)
        return dest(*kept, **kwargs)


def discard_ith_arg(module, i, **kwargs):
    def discarded_arg(*args, **kwargs2):
        return DiscardIthArg(*args, i=i, To=module, **kwargs, **kwargs2)

    return discarded_arg


class MergeFlatInputs(nn.Module):
    """``flat_module(relu(x1 + resizer(x2)))`` (is_sum_merge=True, upstream encoders.py:175-183).

    ``x2`` may carry fewer broadcastable rows than ``x1`` (the CNP decoder passes one representation per task):
    the resizer MLP then runs once per task instead of once per target, and the broadcast happens inside the
    fused add+relu kernel.  Concatenation merge (is_sum_merge=False) is not used by any upstream config.
    """

    def __init__(self, FlatModule, x1_dim, x2_dim, n_out, is_sum_merge=False, **kwargs):
        super().__init__()
        if not is_sum_merge:
            raise NotImplementedError("npf_b200.MergeFlatInputs: only is_sum_merge=True is implemented")
        self.is_sum_merge = True
        self.resizer = MLP(x2_dim, x1_dim)
        self.flat_module = FlatModule(x1_dim, n_out, **kwargs)

    def forward(self, x1, x2):
        if x2.dim() == 4 and x2.shape[2] > 1 and x2.stride(2) == 0:
            x2 = x2[:, :, :1]  # a broadcast view over the targets: resize once per task, broadcast in the kernel
        merged = ops.merge_relu(x1, self.resizer(x2))
        fm = self.flat_module
        if isinstance(fm, nn.Linear):
            return ops.linear(merged, fm.weight, fm.bias)
        return fm(merged)

    def reset_parameters(self):
        pass


def merge_flat_input(module, is_sum_merge=False, **kwargs):
    def merged_flat_input(x_shape, flat_dim, n_out, **kwargs2):
        assert isinstance(x_shape, int)
        return MergeFlatInputs(module, x_shape, flat_dim, n_out, is_sum_merge=is_sum_merge, **kwargs2, **kwargs)

    return merged_flat_input
