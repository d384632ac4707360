"""Channel-last residual depthwise-separable CNN on the CUDA kernels.

Interface and parameter names follow upstream npf/architectures/cnn.py (``ResConvBlock`` 126-215, ``CNN`` 307-380):
``conv_blocks.i.{norm1, conv1.depthwise, conv1.pointwise, norm2, conv2_depthwise, conv2_pointwise}``.  The ``Conv``
and ``Normalization`` classes given by the caller (``nn.Conv1d`` / ``nn.Conv2d``, ``nn.Identity`` /
``nn.BatchNorm{1,2}d``) are instantiated only to *hold* parameters (same shapes, same default init as upstream); the
arithmetic is done by ``ops.dwconv`` (depthwise k-tap + pre-activation + residual) and ``ops.linear`` (1x1 pointwise).
The signal stays channel-last end to end: the permutes of upstream ``CNN.forward`` (cnn.py:364-370) do not exist.
"""
import torch
import torch.nn as nn

from .. import ops
from ..utils.helpers import conv_padding, make_depth_sep_conv

__all__ = ["ResConvBlock", "CNN"]


def _check_conv(conv, what):
    if not isinstance(conv, (nn.Conv1d, nn.Conv2d)) or conv.padding_mode != "zeros" or any(s != 1 for s in conv.stride) \
            or any(d != 1 for d in conv.dilation):
        raise NotImplementedError(f"npf_b200: {what} must be a stride-1 nn.Conv1d / nn.Conv2d, zero-padded or wrapped by "
                                  "make_padded_conv(Conv, CircularPad2d)")
    conv_padding(conv)   # 'same' padding, zero or circular


def _depthwise(x, conv, res, sc, sh):
    """relu(sc * x + sh) -> depthwise conv (+ res).  With a circular padder (upstream helpers.py:334-351, 406-414) the
    signal is extended by wrap-around, run through the same zero-padded kernel and cropped: on the original grid every tap
    then reads a wrapped sample, never the kernel's own zero padding."""
    padder, p = conv_padding(conv)
    if padder is None:
        return ops.dwconv(x, conv.weight, conv.bias, res, True, sc, sh)
    y = ops.dwconv(padder(x).contiguous(), conv.weight, conv.bias, None, True, sc, sh)[:, p:-p, p:-p, :]
    return y + res if res is not None else y.contiguous()


class _PreNorm:
    """Folds ``Normalization(in_chan)`` into a per-channel affine (scale, shift) consumed by the depthwise kernel's
    pre-activation.  Identity -> (None, None).  BatchNorm: eval mode uses running statistics; train mode uses the
    batch statistics (two-pass, differentiable through ``ops.channel_moments``) and updates the running ones."""

    @staticmethod
    def affine(norm, x):
        if isinstance(norm, nn.Identity):
            return None, None
        if not isinstance(norm, (nn.BatchNorm1d, nn.BatchNorm2d)):
            raise NotImplementedError("npf_b200: Normalization must be nn.Identity or nn.BatchNorm{1,2}d")
        if norm.training or not norm.track_running_stats:
            mean, var = ops.channel_moments(x)
            n = x.numel() // x.shape[-1]
            sync = getattr(norm, "_npf_sync_group", None)
            if sync is not None and norm.training:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size(sync[0]) > 1:
                    from ..parallel import sync_moments
                    mean, var, n = sync_moments(mean, var, n, sync[0])   # statistics of the global meta-batch
                    n = n.detach()                                       # total count stays a device scalar: no host sync
            if norm.track_running_stats:
                with torch.no_grad():
                    norm.num_batches_tracked += 1
                    unbias = n / (n - 1).clamp(min=1) if torch.is_tensor(n) else n / max(n - 1, 1)
                    if norm.momentum is None:      # torch semantics: cumulative moving average, factor 1 / num_batches_tracked
                        mom = 1.0 / norm.num_batches_tracked.to(torch.float32)        # device scalar: no host sync
                        norm.running_mean.lerp_(mean.detach(), mom)
                        norm.running_var.lerp_(var.detach() * unbias, mom)
                    else:
                        mom = norm.momentum
                        norm.running_mean.mul_(1 - mom).add_(mean.detach(), alpha=mom)
                        norm.running_var.mul_(1 - mom).add_(var.detach() * unbias, alpha=mom)
        else:
            mean, var = norm.running_mean, norm.running_var
        rstd = torch.rsqrt(var + norm.eps)
        scale = norm.weight * rstd if norm.affine else rstd
        shift = (norm.bias if norm.affine else 0) - mean * scale
        return scale.contiguous(), shift.contiguous()


class ResConvBlock(nn.Module):
    """Pre-activation residual block: ``pw2(dw2(relu(norm2(h))) + X)`` with ``h = pw1(dw1(relu(norm1(X))))`` when
    ``n_conv_layers == 2`` and ``h = X`` otherwise (upstream cnn.py:204-215: the residual is the block *input* and is
    added before the last pointwise)."""

    def __init__(self, in_chan, out_chan, Conv, kernel_size=5, activation=nn.ReLU(), Normalization=nn.Identity,
                 is_bias=True, n_conv_layers=1):
        super().__init__()
        if not isinstance(activation, nn.ReLU):
            raise NotImplementedError("npf_b200.ResConvBlock: only nn.ReLU is implemented")
        assert n_conv_layers in (1, 2)
        if kernel_size % 2 == 0:
            raise ValueError("`kernel_size={}`, but should be odd.".format(kernel_size))
        self.activation, self.n_conv_layers = activation, n_conv_layers
        pad = kernel_size // 2
        if n_conv_layers == 2:
            self.norm1 = Normalization(in_chan)
            self.conv1 = make_depth_sep_conv(Conv)(in_chan, in_chan, kernel_size, padding=pad, bias=is_bias)
            _check_conv(self.conv1.depthwise, "Conv")
        self.norm2 = Normalization(in_chan)
        self.conv2_depthwise = Conv(in_chan, in_chan, kernel_size, padding=pad, groups=in_chan, bias=is_bias)
        self.conv2_pointwise = Conv(in_chan, out_chan, 1, bias=is_bias)
        _check_conv(self.conv2_depthwise, "Conv")
        _check_conv(self.conv2_pointwise, "Conv")

    def reset_parameters(self):
        pass

    @staticmethod
    def _pointwise(x, conv):
        return ops.linear(x, conv.weight, conv.bias)   # 1x1 conv weight [out, in, 1(, 1)] read as [out, in]

    def forward(self, X):
        """X channel-last: [B, L, C] or [B, H, W, C]."""
        h = X
        if self.n_conv_layers == 1 and isinstance(self.norm2, nn.Identity) and self.conv2_depthwise.bias is not None \
                and getattr(self.conv2_depthwise, "padder", None) is None \
                and self.conv2_pointwise.bias is not None and ops.resblock1d_supported(X, self.conv2_depthwise.weight, self.conv2_pointwise.weight) \
                and self.conv2_depthwise.weight.is_contiguous() and self.conv2_pointwise.weight.is_contiguous():
            # the ConvCNP default block: one kernel (depthwise out of a TMA-staged raw tile, pointwise on the tensor core)
            return ops.resblock1d(X, self.conv2_depthwise.weight, self.conv2_depthwise.bias, self.conv2_pointwise.weight, self.conv2_pointwise.bias)
        if self.n_conv_layers == 2:
            sc, sh = _PreNorm.affine(self.norm1, X)
            h = _depthwise(X, self.conv1.depthwise, None, sc, sh)
            h = self._pointwise(h, self.conv1.pointwise)
        sc, sh = _PreNorm.affine(self.norm2, h)
        h = _depthwise(h, self.conv2_depthwise, X, sc, sh)
        return self._pointwise(h, self.conv2_pointwise)


class CNN(nn.Module):
    """Stack of ``ConvBlock``s over a channel-last signal (upstream cnn.py:307-380).  ``is_chan_last`` must be True:
    every upstream model builds its CNN that way and the kernels never leave the channel-last layout."""

    def __init__(self, n_channels, ConvBlock, n_blocks=3, is_chan_last=False, **kwargs):
        super().__init__()
        if not is_chan_last:
            raise NotImplementedError("npf_b200.CNN: only is_chan_last=True is implemented")
        if ConvBlock is not ResConvBlock and not (isinstance(ConvBlock, type) and issubclass(ConvBlock, ResConvBlock)):
            raise NotImplementedError("npf_b200.CNN: only ResConvBlock is implemented")
        self.n_blocks, self.is_chan_last = n_blocks, is_chan_last
        chans = [n_channels] * (n_blocks + 1) if isinstance(n_channels, int) else list(n_channels)
        assert len(chans) == n_blocks + 1, "{} != {}".format(len(chans), n_blocks + 1)
        self.in_out_channels = list(zip(chans, chans[1:]))
        self.conv_blocks = nn.ModuleList(ConvBlock(i, o, **kwargs) for i, o in self.in_out_channels)
        self.is_return_rep = False

    def reset_parameters(self):
        pass

    def forward(self, X):
        for block in self.conv_blocks:
            X = block(X)
        return X
