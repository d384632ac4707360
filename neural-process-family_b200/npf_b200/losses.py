"""Losses of the neural-process family with the per-point Gaussian log-likelihood and its reduction over targets
fused into one kernel.  Same classes / call contract as upstream npf/losses.py: ``Loss()(pred_outputs, Y_trgt)`` with
``pred_outputs`` the 4-tuple returned by the models; ``reduction`` in {None, "mean", "sum"} over the batch."""
import abc
import math

import torch
import torch.nn as nn
from torch.distributions import Normal
from torch.distributions.independent import Independent
from torch.distributions.kl import kl_divergence

from . import ops
from .utils.helpers import sum_from_nth_dim

__all__ = ["CNPFLoss", "ELBOLossLNPF", "NLLLossLNPF", "sum_log_prob"]


def sum_log_prob(prob, sample):
    """sum_t sum_y log p(sample) -> [n_z, B] (upstream losses.py:18-24).  Diagonal Gaussians on the GPU go through
    the fused kernel; any other distribution through its own ``log_prob``."""
    if isinstance(prob, Independent) and isinstance(prob.base_dist, Normal) and prob.base_dist.loc.is_cuda \
            and prob.base_dist.loc.dim() >= 3:
        loc, scale = prob.base_dist.loc, prob.base_dist.scale
        if sample.dim() == loc.dim() - 1:  # targets [B, *, y] against [n_z, B, *, y] predictions
            return ops.gauss_sum_log_prob(loc, scale, sample)
        if sample.shape == loc.shape:  # latent samples [n_z, B, *, z] against q [B, *, z]: handled below
            pass
    return sum_from_nth_dim(prob.log_prob(sample), 2)


class BaseLossNPF(nn.Module, abc.ABC):
    def __init__(self, reduction="mean", is_force_mle_eval=True):
        super().__init__()
        self.reduction = reduction
        self.is_force_mle_eval = is_force_mle_eval

    def forward(self, pred_outputs, Y_trgt):
        p_yCc, z_samples, q_zCc, q_zCct = pred_outputs
        if self.training:
            loss = self.get_loss(p_yCc, z_samples, q_zCc, q_zCct, Y_trgt)
        else:  # evaluation always reports the (approximate) log marginal likelihood (upstream losses.py:63-69)
            if self.is_force_mle_eval:
                q_zCct = None
            loss = NLLLossLNPF.get_loss(self, p_yCc, z_samples, q_zCc, q_zCct, Y_trgt)
        if self.reduction is None:
            return loss
        if self.reduction == "mean":
            return loss.mean(0)
        if self.reduction == "sum":
            return loss.sum(0)
        raise ValueError(f"Unknown {self.reduction}")

    @abc.abstractmethod
    def get_loss(self, p_yCc, z_samples, q_zCc, q_zCct, Y_trgt):
        pass


class CNPFLoss(BaseLossNPF):
    """-sum_t log p(y_t | context) per task (conditional NPs; upstream losses.py:112-123)."""

    def get_loss(self, p_yCc, _, q_zCc, ___, Y_trgt):
        assert q_zCc is None
        return -sum_log_prob(p_yCc, Y_trgt).squeeze(0)


class ELBOLossLNPF(BaseLossNPF):
    """-(E_z sum_t log p(y_t|z) - sum_l KL[q(z_l|C,T) || q(z_l|C)]) (upstream losses.py:126-150)."""

    def get_loss(self, p_yCc, _, q_zCc, q_zCct, Y_trgt):
        e_ll = sum_log_prob(p_yCc, Y_trgt).mean(0)
        kl = sum_from_nth_dim(kl_divergence(q_zCct, q_zCc), 1)
        return -(e_ll - kl)


class NLLLossLNPF(BaseLossNPF):
    """-(logsumexp_z sum_t log p(y_t|z) - log n_z), with importance weights q(z|C)/q(z|C,T) when the samples came
    from q(z|C,T) (upstream losses.py:153-203)."""

    def get_loss(self, p_yCc, z_samples, q_zCc, q_zCct, Y_trgt):
        n_z = p_yCc.batch_shape[0]
        w = sum_log_prob(p_yCc, Y_trgt)
        if q_zCct is not None:
            w = w + sum_log_prob(q_zCc, z_samples) - sum_log_prob(q_zCct, z_samples)
        return -(torch.logsumexp(w, 0) - math.log(n_z))
