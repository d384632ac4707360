"""Context / target split of a meta-batch ON THE DEVICE (the collate step in front of the hot path).

Same classes, constructor arguments and call contract as upstream npf/utils/datasplit.py (``GetRandomIndcs`` 60-145,
``CntxtTrgtGetter`` 148-255, ``RandomMasker`` 259-278, ``no_masker`` 329-333, ``GridCntxtTrgtGetter`` 336-452), but
the data never leaves HBM: the per-row random subsets come from ``npf_random_subset`` / ``npf_random_mask``
(Philox-driven partial Fisher-Yates, one CTA per row), the gathers from ``npf_select_points`` / ``npf_grid_select``.

What stays on the host, exactly as upstream: HOW MANY points (``random.randint(a, b)`` on python's ``random``, one draw
per batch).  The key of the device generator is drawn from numpy's global RNG -- the generator upstream spends on its
per-row shuffles -- so python's ``random`` stream is consumed exactly as upstream consumes it: after upstream's
``set_seed`` a run sees the SAME sequence of context sizes as upstream, and is reproducible.  The index *values* follow
the device generator, not numpy's Mersenne Twister: same distribution as upstream's split, not the same stream.

CUDA only: indices and masks are created on the device the getter was given; there is no CPU path.
"""
import random

import numpy as np

import torch

from .. import _cabi
from .helpers import channels_to_last_dim, prod

__all__ = ["get_all_indcs", "GetRangeIndcs", "GetRandomIndcs", "CntxtTrgtGetter", "RandomMasker", "no_masker", "half_masker",
           "GridCntxtTrgtGetter"]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device(device):
    device = torch.device("cuda" if device is None else device)
    if device.type != "cuda":
        raise RuntimeError("npf_b200.utils.datasplit runs on CUDA devices only (there is no CPU fallback)")
    return device


def _draw_seed():
    """63-bit key for the device generator, from numpy's global RNG (see the module docstring)."""
    return int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))


def ratio_to_int(percentage, max_val):
    """A ratio in [0, 1) becomes a count out of ``max_val``; counts pass through (upstream helpers.py:99-108)."""
    if 1 <= percentage <= max_val:
        out = percentage
    elif 0 <= percentage < 1:
        out = percentage * max_val
    else:
        raise ValueError("percentage={} outside of [0,{}].".format(percentage, max_val))
    return int(out)


class _AllIndcs:
    """Marker returned by ``get_all_indcs``: every point, in order -- ``select`` then returns the inputs themselves
    instead of gathering a copy."""

    def __init__(self, batch_size, n_possible_points):
        self.shape = (batch_size, n_possible_points)

    def materialize(self, device):
        return torch.arange(self.shape[1], dtype=torch.int32, device=device).expand(*self.shape).contiguous()


def get_all_indcs(batch_size, n_possible_points, device=None):
    return _AllIndcs(batch_size, n_possible_points)


class GetRangeIndcs:
    """All indices in ``arange`` (upstream 37-45), shared by the batch."""

    def __init__(self, arange):
        self.arange = arange

    def __call__(self, batch_size, n_possible_points, device=None):
        indcs = torch.arange(*self.arange, dtype=torch.int32, device=_device(device))
        return indcs.expand(batch_size, len(indcs)).contiguous()


class GetRandomIndcs:
    """Random subset of indices: the count is drawn on the host like upstream, the subsets on the device."""

    def __init__(self, a=0.1, b=0.5, is_batch_share=False, range_indcs=None, is_ensure_one=False, is_beta_binomial=False,
                 proba_uniform=0):
        self.a, self.b = a, b
        self.is_batch_share = is_batch_share
        self.range_indcs = range_indcs
        self.is_ensure_one = is_ensure_one
        self.is_beta_binomial = is_beta_binomial
        self.proba_uniform = proba_uniform

    def n_indcs(self, n_possible_points):
        """The number of points of this batch (upstream 112-128; same host RNG calls in the same order)."""
        if np.random.uniform(size=1) < self.proba_uniform:
            n = random.randint(0, n_possible_points)
        elif self.is_beta_binomial:
            from scipy.stats import betabinom
            n = int(betabinom(n_possible_points, self.a, self.b).rvs())
        else:
            n = random.randint(ratio_to_int(self.a, n_possible_points), ratio_to_int(self.b, n_possible_points))
        if self.is_ensure_one and n < 1:
            n = 1
        return n

    def __call__(self, batch_size, n_possible_points, device=None):
        device = _device(device)
        if self.range_indcs is not None:
            n_possible_points = self.range_indcs[1] - self.range_indcs[0]
        n = self.n_indcs(n_possible_points)
        seed = _draw_seed()
        rows = 1 if self.is_batch_share else batch_size
        indcs = torch.empty(rows, n, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            _cabi.call("npf_random_subset", indcs.data_ptr(), rows, n_possible_points, n, seed, _stream())
        if self.is_batch_share:
            indcs = indcs.expand(batch_size, n).contiguous()
        if self.range_indcs is not None:
            indcs += self.range_indcs[0]
        return indcs


class CntxtTrgtGetter:
    """Split ``X [B, N, x_dim]``, ``y [B, N, y_dim]`` (device tensors) into context and target sets."""

    def __init__(self, contexts_getter=GetRandomIndcs(), targets_getter=get_all_indcs, is_add_cntxts_to_trgts=False):
        self.contexts_getter = contexts_getter
        self.targets_getter = targets_getter
        self.is_add_cntxts_to_trgts = is_add_cntxts_to_trgts

    def __call__(self, X, y=None, context_indcs=None, target_indcs=None, is_return_indcs=False):
        batch_size, num_points = self.getter_inputs(X)
        if context_indcs is None:
            context_indcs = self.contexts_getter(batch_size, num_points, device=X.device)
        if target_indcs is None:
            target_indcs = self.targets_getter(batch_size, num_points, device=X.device)
        if self.is_add_cntxts_to_trgts:
            target_indcs = self.add_cntxts_to_trgts(num_points, target_indcs, context_indcs)
        X_pre_cntxt = self.preprocess_context(X)
        if is_return_indcs:
            return context_indcs, X_pre_cntxt, target_indcs, X
        X_cntxt, Y_cntxt = self.select(X_pre_cntxt, y, context_indcs)
        X_trgt, Y_trgt = self.select(X, y, target_indcs)
        return X_cntxt, Y_cntxt, X_trgt, Y_trgt

    def preprocess_context(self, X):
        return X

    def add_cntxts_to_trgts(self, num_points, target_indcs, context_indcs):
        if isinstance(target_indcs, _AllIndcs):
            target_indcs = target_indcs.materialize(context_indcs.device)
        return torch.cat([target_indcs, context_indcs.to(target_indcs.dtype)], dim=-1)[:, :num_points].contiguous()

    def getter_inputs(self, X):
        batch_size, num_points, _ = X.shape
        return batch_size, num_points

    def select(self, X, y, indcs):
        if isinstance(indcs, _AllIndcs):
            return X.contiguous(), y.contiguous()
        if not X.is_cuda:
            raise RuntimeError("npf_b200.utils.datasplit runs on CUDA tensors only (there is no CPU fallback)")
        B, N, xd = X.shape
        yd = y.size(-1)
        indcs = indcs.to(device=X.device, dtype=torch.int32).contiguous()
        n = indcs.shape[1]
        X, y = X.contiguous().float(), y.contiguous().float()
        Xo = torch.empty(B, n, xd, dtype=torch.float32, device=X.device)
        Yo = torch.empty(B, n, yd, dtype=torch.float32, device=X.device)
        with torch.cuda.device(X.device):
            _cabi.call("npf_select_points", X.data_ptr(), y.data_ptr(), indcs.data_ptr(), Xo.data_ptr(), Yo.data_ptr(), B, N, n, xd,
                       yd, _stream())
        return Xo, Yo


class RandomMasker(GetRandomIndcs):
    """Random boolean mask ``[B, *mask_shape, 1]`` with the same number of ones in every row (upstream 259-278)."""

    def __call__(self, batch_size, mask_shape, device=None):
        device = _device(device)
        P = prod(mask_shape)
        n = self.n_indcs(P)
        seed = _draw_seed()
        rows = 1 if self.is_batch_share else batch_size
        mask = torch.empty(rows, P, dtype=torch.bool, device=device)
        with torch.cuda.device(device):
            _cabi.call("npf_random_mask", mask.data_ptr(), rows, P, n, seed, _stream())
        if self.is_batch_share:
            mask = mask.expand(batch_size, P)
        return mask.view(batch_size, *mask_shape, 1).contiguous()


def half_masker(batch_size, mask_shape, dim=0, device=None):
    """Mask of the first half of ``dim`` (upstream 319-326)."""
    mask = torch.zeros(mask_shape, dtype=torch.bool, device=_device(device))
    slcs = [slice(None)] * len(mask_shape)
    slcs[dim] = slice(0, mask_shape[dim] // 2)
    mask[tuple(slcs)] = True
    return mask.unsqueeze(-1).expand(batch_size, *mask_shape, 1)


def no_masker(batch_size, mask_shape, device=None):
    """All-ones mask as a broadcast view (upstream 329-333)."""
    return torch.ones(1, dtype=torch.bool, device=_device(device)).expand(batch_size, *mask_shape, 1)


class GridCntxtTrgtGetter(CntxtTrgtGetter):
    """Split grids of values ``X [B, y_dim, *grid]`` (e.g. images) into context / target points (upstream 336-452).
    With ``is_return_masks=True`` (what the on-grid GridConv* models consume) nothing is gathered at all."""

    def __init__(self, context_masker=RandomMasker(), target_masker=no_masker, upscale_factor=1, **kwargs):
        self.upscale_factor = upscale_factor
        super().__init__(contexts_getter=context_masker, targets_getter=target_masker, **kwargs)

    def __call__(self, X, y=None, context_mask=None, target_mask=None, is_return_masks=False, **kwargs):
        return super().__call__(channels_to_last_dim(X), context_indcs=context_mask, target_indcs=target_mask,
                                is_return_indcs=is_return_masks, **kwargs)

    def add_cntxts_to_trgts(self, grid_shape, target_mask, context_mask):
        return target_mask | context_mask

    def getter_inputs(self, X):
        batch_size, *grid_shape, _ = X.shape
        return batch_size, grid_shape

    def select(self, X, y, mask, extrapolation=1):
        """Masked grid points in row-major order: coordinates normalised to [-1, 1] (times ``upscale_factor``) and values.
        Every row must mask the same number of points (upstream assumes it silently; here a mismatch raises)."""
        if not X.is_cuda:
            raise RuntimeError("npf_b200.utils.datasplit runs on CUDA tensors only (there is no CPU fallback)")
        B, *grid, yd = X.shape
        if len(grid) not in (1, 2):
            raise NotImplementedError("npf_b200.GridCntxtTrgtGetter.select: 1-D and 2-D grids only")
        H, W = (1, grid[0]) if len(grid) == 1 else grid
        mask = mask.to(X.device).expand(B, *grid, 1).contiguous()
        n = int(mask[0].sum())  # upstream reads the count of the first row the same way (one host sync)
        img = X.contiguous().float()
        Xo = torch.empty(B, n, len(grid), dtype=torch.float32, device=X.device)
        Yo = torch.empty(B, n, yd, dtype=torch.float32, device=X.device)
        counts = torch.empty(B, dtype=torch.int32, device=X.device)
        with torch.cuda.device(X.device):
            _cabi.call("npf_grid_select", mask.data_ptr(), img.data_ptr(), Xo.data_ptr(), Yo.data_ptr(), counts.data_ptr(), B, H, W,
                       len(grid), yd, n, float(self.upscale_factor), _stream())
        if not bool((counts == n).all()):
            raise ValueError("GridCntxtTrgtGetter.select: rows mask different numbers of points")
        return Xo, Yo
