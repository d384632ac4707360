"""Base classes of the (latent) neural-process family on the B200 kernels.

Same constructor keywords, attributes, ``forward(X_cntxt, Y_cntxt, X_trgt, Y_trgt=None)`` signature and returned
4-tuple ``(p_yCc, z_samples, q_zCc, q_zCct)`` as upstream npf/neuralproc/base.py (``NeuralProcessFamily`` 23-371,
``LatentNeuralProcessFamily`` 374-575).  Differences, all internal:
  * the input range check (base.py:241-247) runs on the device and raises one step late instead of forcing a
    device->host sync in every training forward (``strict_validation`` restores the synchronous behaviour);
  * the predictive head (split + ``0.01 + 0.99 softplus``) is one fused kernel;
  * ``rsample`` is ``q_loc + q_scale * eps`` with ``eps ~ N(0,1)`` drawn by torch on the device, fused with the
    ``0.1 + 0.9 sigmoid`` scale transform.
"""
import abc
from functools import partial

import torch
import torch.nn as nn

from .. import ops
from ..architectures import MLP, merge_flat_input
from ..utils.helpers import MultivariateNormalDiag
from .helpers import pool_and_replicate_middle

__all__ = ["NeuralProcessFamily", "LatentNeuralProcessFamily", "MinSoftplus", "BoundedSigmoid"]


class MinSoftplus:
    """y -> lo + (1 - lo) softplus(y): the default predictive-scale transform (upstream base.py:116, lo = 0.01).
    A recognisable callable (instead of upstream's lambda) so that ``decode`` can use the fused head kernel."""

    def __init__(self, lo=0.01):
        self.lo = lo

    def __call__(self, y):
        return self.lo + (1 - self.lo) * nn.functional.softplus(y)


class BoundedSigmoid:
    """z -> 0.1 + 0.9 sigmoid(z): the default latent-scale transform (upstream base.py:432)."""

    def __call__(self, z):
        return 0.1 + 0.9 * torch.sigmoid(z)


class NeuralProcessFamily(nn.Module, abc.ABC):
    _valid_paths = ["deterministic", "latent", "both"]
    strict_validation = False  # True: synchronous range check exactly like upstream (one D2H sync per step)

    def __init__(self, x_dim, y_dim, encoded_path, r_dim=128, x_transf_dim=-1, is_heteroskedastic=True, XEncoder=None,
                 Decoder=None, PredictiveDistribution=MultivariateNormalDiag, p_y_loc_transformer=nn.Identity(),
                 p_y_scale_transformer=None):
        super().__init__()
        self.x_dim, self.y_dim, self.r_dim = x_dim, y_dim, r_dim
        self.is_heteroskedastic = is_heteroskedastic
        if x_transf_dim is None:
            self.x_transf_dim = x_dim
        elif x_transf_dim == -1:
            self.x_transf_dim = r_dim
        else:
            self.x_transf_dim = x_transf_dim
        self.encoded_path = encoded_path.lower()
        if self.encoded_path not in self._valid_paths:
            raise ValueError(f"Unknown encoded_path={self.encoded_path}.")
        XEncoder = XEncoder if XEncoder is not None else self.dflt_Modules["XEncoder"]
        Decoder = Decoder if Decoder is not None else self.dflt_Modules["Decoder"]
        self.x_encoder = XEncoder(self.x_dim, self.x_transf_dim)
        self.decoder = Decoder(self.x_transf_dim, self.r_dim, self.y_dim * 2)  # loc and scale
        self.PredictiveDistribution = PredictiveDistribution
        self.p_y_loc_transformer = p_y_loc_transformer
        self.p_y_scale_transformer = p_y_scale_transformer if p_y_scale_transformer is not None else MinSoftplus(0.01)
        self._range_state = None
        self.reset_parameters()

    def reset_parameters(self):
        pass

    @property
    def dflt_Modules(self):
        d = dict()
        d["XEncoder"] = partial(MLP, n_hidden_layers=1, hidden_size=self.r_dim)
        d["SubDecoder"] = partial(MLP, n_hidden_layers=4, hidden_size=self.r_dim)
        d["Decoder"] = merge_flat_input(d["SubDecoder"], is_sum_merge=True)
        return d

    # -------------------------------------------------------------------------------------------- forward template
    def forward(self, X_cntxt, Y_cntxt, X_trgt, Y_trgt=None):
        """Returns ``(p_yCc, z_samples, q_zCc, q_zCct)``; ``p_yCc`` has batch shape [n_z, B, *n_trgt] and event
        shape [y_dim] (upstream base.py:177-239)."""
        self._validate_inputs(X_cntxt, Y_cntxt, X_trgt, Y_trgt)
        X_cntxt = self.x_encoder(X_cntxt)
        X_trgt = self.x_encoder(X_trgt)
        R = self.encode_globally(X_cntxt, Y_cntxt)
        if self.encoded_path in ("latent", "both"):
            z_samples, q_zCc, q_zCct = self.latent_path(X_cntxt, R, X_trgt, Y_trgt)
        else:
            z_samples, q_zCc, q_zCct = None, None, None
        if self.encoded_path == "latent":
            R = None
        R_trgt = self.trgt_dependent_representation(X_cntxt, z_samples, R, X_trgt)
        p_yCc = self.decode(X_trgt, R_trgt)
        return p_yCc, z_samples, q_zCc, q_zCct

    # -------------------------------------------------------------------------------------------- validation
    def _validate_inputs(self, X_cntxt, Y_cntxt, X_trgt, Y_trgt):
        """Training-time check that features lie in [-1, 1] (upstream base.py:241-247), on the device."""
        if not self.training or not torch.is_floating_point(X_cntxt):
            return
        if not X_cntxt.is_cuda:
            raise RuntimeError("npf_b200 models run on CUDA tensors only (there is no CPU fallback)")
        st = self._range_state
        if st is None or st["flag"].device != X_cntxt.device:
            st = dict(flag=torch.zeros(1, dtype=torch.int32, device=X_cntxt.device),
                      host=torch.zeros(1, dtype=torch.int32).pin_memory(), event=None)
            self._range_state = st
        if torch.cuda.is_current_stream_capturing():     # CUDA-graph capture (graph.GraphedStep): record the check kernel only;
            ops.range_flag(st["flag"], X_cntxt, X_trgt, lo=-1.0, hi=1.0)   # the flag is read back after each replay
            return
        self._raise_if_flagged(wait=False)
        ops.range_flag(st["flag"], X_cntxt, X_trgt, lo=-1.0, hi=1.0)
        self._post_range_check()

    def _post_range_check(self):
        st = self._range_state
        st["host"].copy_(st["flag"], non_blocking=True)
        st["event"] = torch.cuda.Event()
        st["event"].record()
        if self.strict_validation:
            self._raise_if_flagged(wait=True)

    def _after_graph_replay(self):
        """Host side of the range validation for a replayed step: raise for the previous replay, queue this one's read."""
        if self._range_state is None or not self.training:
            return
        self._raise_if_flagged(wait=False)
        self._post_range_check()

    def _raise_if_flagged(self, wait):
        st = self._range_state
        if st is None or st["event"] is None:
            return
        if wait:
            st["event"].synchronize()
        elif not st["event"].query():
            return
        if int(st["host"][0]) != 0:
            st["flag"].zero_()
            st["host"].zero_()
            st["event"] = None
            raise ValueError("Features during training should be in [-1,1].")

    def validate_now(self):
        """Block until the last asynchronous range check has finished and raise ``ValueError`` if it failed."""
        self._raise_if_flagged(wait=True)

    # -------------------------------------------------------------------------------------------- abstract parts
    @abc.abstractmethod
    def encode_globally(self, X_cntxt, Y_cntxt):
        pass

    @abc.abstractmethod
    def trgt_dependent_representation(self, X_cntxt, z_samples, R, X_trgt):
        pass

    def latent_path(self, X_cntxt, R, X_trgt, Y_trgt):
        raise NotImplementedError(f"`latent_path` not implemented. Cannot use encoded_path={self.encoded_path} in such case.")

    # -------------------------------------------------------------------------------------------- decoding
    def decode(self, X_trgt, R_trgt):
        """decoder -> split -> transforms -> predictive distribution (upstream base.py:327-367)."""
        suff = self.decoder(X_trgt, R_trgt)
        fused = isinstance(self.p_y_scale_transformer, MinSoftplus) and isinstance(self.p_y_loc_transformer, nn.Identity)
        if fused:
            loc, scale = ops.gauss_head(suff, self.p_y_scale_transformer.lo)
        else:  # user-supplied transforms: applied as given
            loc, scale = suff.split(self.y_dim, dim=-1)
            loc = self.p_y_loc_transformer(loc)
            scale = self.p_y_scale_transformer(scale)
        if not self.is_heteroskedastic:
            n_z, B, *n_trgt, y = scale.shape
            scale = pool_and_replicate_middle(scale.reshape(n_z * B, *n_trgt, y)).reshape(n_z, B, *n_trgt, y)
        return self.PredictiveDistribution(loc, scale)

    def set_extrapolation(self, min_max):
        pass


class LatentNeuralProcessFamily(NeuralProcessFamily):
    _valid_paths = ["latent", "both"]

    def __init__(self, *args, is_q_zCct=False, n_z_samples_train=32, n_z_samples_test=32, LatentEncoder=None,
                 LatentDistribution=MultivariateNormalDiag, q_z_loc_transformer=nn.Identity(), q_z_scale_transformer=None,
                 z_dim=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.is_q_zCct = is_q_zCct
        self.n_z_samples_train, self.n_z_samples_test = n_z_samples_train, n_z_samples_test
        self.z_dim = self.r_dim if z_dim is None else z_dim
        LatentEncoder = LatentEncoder if LatentEncoder is not None else self.dflt_Modules["LatentEncoder"]
        self.latent_encoder = LatentEncoder(self.r_dim, self.z_dim * 2)
        if self.encoded_path == "both":
            self.r_z_merger = nn.Linear(self.r_dim + self.z_dim, self.r_dim)
        self.LatentDistribution = LatentDistribution
        self.q_z_loc_transformer = q_z_loc_transformer
        self.q_z_scale_transformer = q_z_scale_transformer if q_z_scale_transformer is not None else BoundedSigmoid()
        if not (isinstance(self.q_z_scale_transformer, BoundedSigmoid) and isinstance(q_z_loc_transformer, nn.Identity)):
            raise NotImplementedError("npf_b200: custom latent transforms are not implemented in the fused sampler")
        if self.z_dim != self.r_dim and self.encoded_path == "latent":
            self.reshaper_z = nn.Linear(self.z_dim, self.r_dim)
        self._eps_override = None  # tests: feed a fixed eps once instead of drawing it

    @property
    def dflt_Modules(self):
        d = NeuralProcessFamily.dflt_Modules.__get__(self)
        d["LatentEncoder"] = partial(MLP, n_hidden_layers=1, hidden_size=self.r_dim)
        return d

    def forward(self, *args, **kwargs):
        n = self.n_z_samples_train if self.training else self.n_z_samples_test
        self.n_z_samples = n.rvs() if hasattr(n, "rvs") else n  # scipy random variable allowed (upstream 478-488)
        return super().forward(*args, **kwargs)

    def latent_path(self, X_cntxt, R, X_trgt, Y_trgt):
        """q(z|C) (and q(z|C,T) when ``is_q_zCct`` and targets are known), then n_z reparameterised samples."""
        suff_c = self._latent_suffstat(X_cntxt, R)
        if self.is_q_zCct and Y_trgt is not None:
            suff_ct = self._latent_suffstat(X_trgt, self.encode_globally(X_trgt, Y_trgt))
            eps = self._draw_eps(suff_ct)
            q_loc, q_scale, z_samples = ops.latent_sample(suff_ct, eps)
            q_zCct = self.LatentDistribution(q_loc, q_scale)
            c_loc, c_scale, _ = ops.latent_sample(suff_c, eps[:0])  # parameters only (no samples drawn)
            q_zCc = self.LatentDistribution(c_loc, c_scale)
        else:
            eps = self._draw_eps(suff_c)
            q_loc, q_scale, z_samples = ops.latent_sample(suff_c, eps)
            q_zCc, q_zCct = self.LatentDistribution(q_loc, q_scale), None
        return z_samples, q_zCc, q_zCct

    def _draw_eps(self, suff):
        shape = (self.n_z_samples, *suff.shape[:-1], self.z_dim)
        if self._eps_override is not None:
            eps, self._eps_override = self._eps_override, None
            assert tuple(eps.shape) == tuple(shape), f"eps override has shape {tuple(eps.shape)}, expected {shape}"
            return eps.to(suff.device)
        return torch.randn(shape, device=suff.device, dtype=torch.float32)

    def _latent_suffstat(self, X, R):
        return self.latent_encoder(self.rep_to_lat_input(R))

    def infer_latent_dist(self, X, R):
        """q(z | .) as a distribution object (upstream base.py:516-547)."""
        suff = self._latent_suffstat(X, R)
        q_loc, q_scale, _ = ops.latent_sample(suff, suff.new_zeros((0, *suff.shape[:-1], self.z_dim)))
        return self.LatentDistribution(q_loc, q_scale)

    def rep_to_lat_input(self, R):
        return R

    def merge_r_z(self, R, z_samples):
        """relu(Linear([R ; z])) with R broadcast over the z-sample axis (upstream base.py:554-575)."""
        if R.shape != z_samples.shape:
            R = R.unsqueeze(0).expand(*z_samples.shape[:-1], self.r_dim)
        return ops.linear(torch.cat((R, z_samples), dim=-1), self.r_z_merger.weight, self.r_z_merger.bias, relu=True)
