"""torch.autograd.Functions over the C ABI of libnpf_b200.so.

PyTorch is used for device memory (torch.empty / zeros from the caching allocator), the current CUDA stream
and autograd bookkeeping; every arithmetic op on the hot path is a kernel of the library.  All tensors are
fp32, contiguous, on a CUDA device.  No CPU fallback: calling an op with CPU tensors raises.
"""
import ctypes

import torch

from . import _cabi
from ._cabi import ACCUM, ADD_DY, MASK_X, RELU_IN, RELU_OUT, call

__all__ = [
    "set_precision", "get_precision", "linear", "mlp_chain", "setconv", "dwconv", "resblock1d", "resblock1d_supported", "channel_moments",
    "merge_relu", "mean_pool", "add_layernorm", "xattn", "gauss_head", "gauss_sum_log_prob", "latent_sample",
    "global_latent", "gridconv_in", "range_flag", "launch_count", "set_direct_grad_accumulation",
]

_PRECISION = {"fp32": _cabi.PREC_FP32, "bf16": _cabi.PREC_BF16, "bf16x3": _cabi.PREC_BF16X3}
_precision = _cabi.PREC_FP32


def set_precision(name):
    """'fp32' (FFMA, 1e-4 parity), 'bf16' (tensor cores, 1e-2 parity) or 'bf16x3' (split-bf16 tensor cores)."""
    global _precision
    _precision = _PRECISION[name]


def get_precision():
    return {v: k for k, v in _PRECISION.items()}[_precision]


def launch_count():
    return _cabi.launch_count()


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("npf_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if t.dtype != torch.float32 and t.dtype != torch.uint8 and t.dtype != torch.bool and t.dtype != torch.int32:
            raise TypeError(f"npf_b200 kernels take fp32 tensors, got {t.dtype}")


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _shape(cond, msg):
    """The kernels take raw pointers + sizes: every Function checks its operands' shapes on the host first (a mismatch
    that upstream would surface as a broadcast error must not become an out-of-bounds device read)."""
    if not cond:
        raise ValueError("npf_b200: " + msg)


# ------------------------------------------------------------------------------------------------------
# Direct gradient accumulation: every weight-gradient kernel ADDS into its output, so when a parameter already has a
# contiguous fp32 ``.grad`` (e.g. a view into ``parallel.FlatGradients``' bucket) the kernel can write there directly
# and the Function returns ``None`` for it -- no zeros_like fill, no autograd accumulate kernel per parameter.
# Scope: only parameters tagged by ``parallel.FlatGradients`` (``p._npf_direct_grad``) take this path -- for every other
# parameter the Functions return the gradient to autograd as usual, so hooks (``register_hook``, DDP reducers,
# post-accumulate hooks) and ``torch.autograd.grad`` keep working.  ``set_direct_grad_accumulation(True)`` forces it for
# every parameter that has a ``.grad`` (process-wide; off by default).
# ------------------------------------------------------------------------------------------------------
_direct_grads = False


def set_direct_grad_accumulation(on):
    global _direct_grads
    _direct_grads = bool(on)


def _gbuf(p):
    """(buffer to accumulate the gradient of ``p`` into, value to return to autograd for it)."""
    if p is None:
        return None, None
    if (_direct_grads or getattr(p, "_npf_direct_grad", False)) and p.is_leaf and p.requires_grad and p.grad is not None \
            and p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.shape == p.shape and p.is_contiguous():
        return p.grad, None
    g = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return g, g


_CHAIN_BWD_MAX_ROWS = 148 * 256   # npf_mlp_chain_bwd: one resident row group per SM
_CHAIN_MAX_ROWS = 1 << 30       # npf_mlp_chain_fwd: one 256-row block per CTA up to 37 888 rows, persistent CTAs over blocks beyond


# ======================================================================================================
# Linear / MLP chain
# ======================================================================================================
def _lin_fwd(x2, W, b, N, K, flags=0, ldw=None, u=None, w2_ptr=None, ldw2=0, prec=None):
    M = x2.shape[0]
    y = torch.empty(M, N, device=x2.device, dtype=torch.float32)
    call("npf_linear_fwd", _p(x2), x2.stride(0) if M else K, _p(W) if not isinstance(W, int) else W, ldw or K, _p(b),
         _p(y), N, M, K, N, flags, _p(u), w2_ptr, ldw2, _precision if prec is None else prec, _stream())
    return y


def _lin_bwd_data(dz, W_ptr, ldw, M, K, N, mask=None, out=None, accumulate=False, prec=None):
    dx = out if out is not None else torch.empty(M, K, device=dz.device, dtype=torch.float32)
    call("npf_linear_bwd_data", _p(dz), N, W_ptr, ldw, _p(dx), K, M, K, N, _p(mask), K if mask is not None else 0,
         ACCUM if accumulate else 0, _precision if prec is None else prec, _stream())
    return dx


def _lin_bwd(dz, x2, W_ptr, ldw, dW_ptr, lddw, db, M, K, N, mask, prec=None):
    """dX, dW +=, db += of one Linear in a single call (one pass over dz and x2 for 128 -> 128 layers)."""
    dx = torch.empty(M, K, device=dz.device, dtype=torch.float32)
    call("npf_linear_bwd", _p(dz), N, _p(x2), K, W_ptr, ldw, _p(dx), K, dW_ptr, lddw, _p(db), M, K, N,
         MASK_X if mask else 0, _precision if prec is None else prec, _stream())
    return dx


def _lin_bwd_weight(dz, x2, dW_ptr, lddw, db, M, K, N, flags=0, u=None, dw2_ptr=None, ldw2=0, prec=None):
    call("npf_linear_bwd_weight", _p(dz), N, _p(x2), K, dW_ptr, lddw, _p(db), M, K, N, flags, _p(u), dw2_ptr, ldw2,
         _precision if prec is None else prec, _stream())


class _MLPChain(torch.autograd.Function):
    """y = L_n(relu(L_{n-1}(... relu(L_1(x))))) with every L_i an nn.Linear (bias optional); optionally a final
    relu.  Saves the post-relu activations; the backward passes the pre-activation gradient from layer to layer
    with the relu mask fused into the data-gradient GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, final_relu, has_bias, prec, *params):
        _chk(x, *params)
        ctx.prec = prec
        n_layers = len(params) // 2 if has_bias else len(params)
        Ws = params[:n_layers]
        bs = params[n_layers:] if has_bias else (None,) * n_layers
        lead = x.shape[:-1]
        h = _c(x).reshape(-1, x.shape[-1])
        acts = [h]
        p_eff = _precision if prec is None else prec
        i = 0
        while i < n_layers:
            W, b = Ws[i], bs[i]
            K = W.numel() // W.shape[0]
            # run of consecutive square 128-wide layers: one kernel keeps the row block on chip between layers
            j = i
            if p_eff != _PRECISION["fp32"] and K == 128 and h.shape[0] > 0:
                while j < n_layers and Ws[j].shape[0] == 128 and Ws[j].numel() == 128 * 128 and Ws[j].is_contiguous():
                    j += 1
            if j - i >= 2 and h.shape[0] <= _CHAIN_MAX_ROWS:
                L, M = j - i, h.shape[0]
                ys = [torch.empty(M, 128, device=h.device, dtype=torch.float32) for _ in range(L)]
                mask = 0
                for l in range(L):
                    if (i + l) != n_layers - 1 or final_relu:
                        mask |= 1 << l
                Wp = (ctypes.c_void_p * L)(*[Ws[i + l].data_ptr() for l in range(L)])
                bp = (ctypes.c_void_p * L)(*[(bs[i + l].data_ptr() if bs[i + l] is not None else None) for l in range(L)])
                Yp = (ctypes.c_void_p * L)(*[y.data_ptr() for y in ys])
                call("npf_mlp_chain_fwd", _p(h), h.stride(0), Wp, bp, Yp, L, M, 128, 0, mask, p_eff, _stream())
                acts.extend(ys)
                h = ys[-1]
                i = j
                continue
            last = i == n_layers - 1
            flags = RELU_OUT if (not last or final_relu) else 0
            # W is [out, in] or a 1x1 conv weight [out, in, 1(, 1)]: same memory, K = numel / out
            h = _lin_fwd(h, _c(W), None if b is None else _c(b), W.shape[0], K, flags, prec=prec)
            acts.append(h)
            i += 1
        ctx.save_for_backward(*acts, *Ws, *(bs if has_bias else ()))
        ctx.n_layers, ctx.final_relu, ctx.has_bias = n_layers, final_relu, has_bias
        ctx.x_shape = x.shape
        return h.reshape(*lead, Ws[-1].shape[0])

    @staticmethod
    def backward(ctx, dy):
        n = ctx.n_layers
        saved = ctx.saved_tensors
        acts, Ws = saved[: n + 1], saved[n + 1: 2 * n + 1]
        bs = saved[2 * n + 1:] if ctx.has_bias else (None,) * n
        M = acts[0].shape[0]
        dz = _c(dy).reshape(M, -1)
        if ctx.final_relu:
            dzm = torch.empty_like(dz)
            call("npf_relu_bwd", _p(dz), _p(acts[n]), _p(dzm), dz.numel(), _stream())
            dz = dzm
        dWs, dbs = [None] * n, [None] * n
        dx = None
        p_eff = _precision if ctx.prec is None else ctx.prec
        i = n - 1
        while i >= 0:
            # run of consecutive square 128-wide layers i0 .. i: ONE kernel keeps the gradient on chip between the layers
            # (npf_mlp_chain_bwd: dY and every saved input read once, only the run's dX written)
            i0 = i
            # (one row group of <= 256 rows per CTA: beyond 148 x 256 rows the kernel would re-stage the weights and flush dW with
            # atomics once per group and layer -- measured slower than one npf_linear_bwd per layer at M = 131 072)
            if p_eff != _PRECISION["fp32"] and 64 <= M <= _CHAIN_BWD_MAX_ROWS and dz.shape[1] == 128 and dz.is_contiguous():
                while i0 >= 0 and Ws[i0].shape[0] == 128 and Ws[i0].numel() == 128 * 128 and Ws[i0].is_contiguous() \
                        and (bs[i0] is None or bs[i0].is_contiguous()):
                    i0 -= 1
                i0 += 1
            if i - i0 + 1 >= 2:
                Lr = i - i0 + 1
                need_dx = i0 > 0 or ctx.needs_input_grad[0]
                bufs = [(_gbuf(Ws[l]), _gbuf(bs[l])) for l in range(i0, i + 1)]
                for l, ((dW, rW), (db, rb)) in zip(range(i0, i + 1), bufs):
                    dWs[l], dbs[l] = rW, rb
                dxr = torch.empty(M, 128, device=dz.device, dtype=torch.float32) if need_dx else None
                Xp = (ctypes.c_void_p * Lr)(*[acts[l].data_ptr() for l in range(i0, i + 1)])
                Wp = (ctypes.c_void_p * Lr)(*[Ws[l].data_ptr() for l in range(i0, i + 1)])
                dWp = (ctypes.c_void_p * Lr)(*[b_[0][0].data_ptr() for b_ in bufs])
                dbp = (ctypes.c_void_p * Lr)(*[(b_[1][0].data_ptr() if b_[1][0] is not None else None) for b_ in bufs])
                call("npf_mlp_chain_bwd", _p(dz), 128, Xp, Wp, _p(dxr), 128, dWp, dbp, Lr, M, 128, MASK_X if i0 > 0 else 0, p_eff, _stream())
                dz = dxr
                if i0 == 0 and need_dx:
                    dx = dz.reshape(ctx.x_shape)
                i = i0 - 1
                continue
            i = _MLPChain._one_layer_bwd(ctx, i, dz, acts, Ws, bs, dWs, dbs, M)
            dz, dx_i = i[1], i[2]
            if dx_i is not None:
                dx = dx_i
            i = i[0]
        grads = tuple(dWs) + (tuple(dbs) if ctx.has_bias else ())
        return (dx, None, None, None) + grads

    @staticmethod
    def _one_layer_bwd(ctx, i, dz, acts, Ws, bs, dWs, dbs, M):
        """backward of layer i alone; returns (next layer index, gradient for it, dx or None)"""
        dx = None
        if True:
            W = _c(Ws[i])
            N = W.shape[0]
            K = W.numel() // N
            dW, dWs[i] = _gbuf(Ws[i])
            db, dbs[i] = _gbuf(bs[i])
            need_dx = i > 0 or ctx.needs_input_grad[0]
            if M > 0 and need_dx:        # one pass over dz and acts[i]: data + weight + bias gradient
                dz = _lin_bwd(dz, acts[i], _p(W), K, _p(dW), K, db, M, K, N, mask=i > 0, prec=ctx.prec)
            elif M > 0:
                _lin_bwd_weight(dz, acts[i], _p(dW), K, db, M, K, N, prec=ctx.prec)
            elif need_dx:
                dz = torch.empty(0, K, device=dz.device, dtype=torch.float32)
            if i == 0 and need_dx:
                dx = dz.reshape(ctx.x_shape)
        return i - 1, dz, dx


def mlp_chain(x, weights, biases, final_relu=False, precision=None):
    """weights: list of [out,in] tensors; biases: list of [out] tensors or None (all or none).
    precision: None (global setting) or 'fp32' / 'bf16' / 'bf16x3' for this chain only."""
    has_bias = biases is not None and biases[0] is not None
    params = tuple(weights) + (tuple(biases) if has_bias else ())
    return _MLPChain.apply(x, final_relu, has_bias, None if precision is None else _PRECISION[precision], *params)


def linear(x, weight, bias=None, relu=False, precision=None):
    return mlp_chain(x, [weight], None if bias is None else [bias], final_relu=relu, precision=precision)


# ======================================================================================================
# SetConv
# ======================================================================================================
class _SetConv(torch.autograd.Function):
    """out = Linear([ sum_k softmax_k(a) v_k ; sum_k exp(a) ]) with the exp-quadratic RBF.  The [.., C+1] concatenation
    never exists as a separate pass: with many channels the density column of the Linear is a rank-1 epilogue term of the
    GEMM; with few channels (context -> induced, C = y_dim) the SetConv kernel writes [feat | dens] interleaved and the
    resizer and its whole backward are single streaming passes (npf_linear_fwd / npf_linear_bwd, K = C + 1)."""

    @staticmethod
    def forward(ctx, keys, queries, values, theta, W, b, keys_regular):
        _chk(keys, queries, values, theta, W, b)
        _shape(values.dim() == 3, f"setconv: values must be [B,K,C], got {tuple(values.shape)}")
        B, K, C = values.shape
        Q = queries.shape[-1]  # queries: [B,Q] per task or [Q] shared; keys likewise
        _shape(keys.dim() in (1, 2) and keys.shape[-1] == K and (keys.dim() == 1 or keys.shape[0] == B),
               f"setconv: keys {tuple(keys.shape)} do not match values {tuple(values.shape)}")
        _shape(queries.dim() in (1, 2) and (queries.dim() == 1 or queries.shape[0] == B),
               f"setconv: queries {tuple(queries.shape)} do not match batch {B}")
        _shape(W.dim() == 2 and W.shape[1] == C + 1 and theta.numel() == 1 and (b is None or b.numel() == W.shape[0]),
               f"setconv: resizer weight {tuple(W.shape)} must be [N, {C + 1}]")
        keys, queries, values = _c(keys), _c(queries), _c(values)
        key_bs = 0 if keys.dim() == 1 else K
        qry_bs = 0 if queries.dim() == 1 else Q
        dev = values.device
        assert W.is_contiguous(), "SetConv resizer weight must be contiguous"
        N = W.shape[0]
        M = B * Q
        ilv = C <= 4 and not keys_regular and K * (1 + C) * 4 <= 40 * 1024 and not ctx.needs_input_grad[2]
        mstat = torch.empty(B, Q, 2, device=dev, dtype=torch.float32)
        out = torch.empty(B, Q, N, device=dev, dtype=torch.float32)
        if ilv:
            feat = torch.empty(B, Q, C + 1, device=dev, dtype=torch.float32)       # [feat | dens]
            dens = feat
            call("npf_setconv_fwd", _p(keys), key_bs, _p(queries), qry_bs, _p(values), _p(theta), _p(feat),
                 feat.data_ptr() + 4 * C, _p(mstat), B, K, Q, C, 0, C + 1, C + 1, _stream())
            if M > 0:
                call("npf_linear_fwd", _p(feat), C + 1, _p(W), C + 1, _p(b), _p(out), N, M, C + 1, N, 0, None, None, 0,
                     _precision, _stream())
        else:
            feat = torch.empty(B, Q, C, device=dev, dtype=torch.float32)
            dens = torch.empty(B, Q, device=dev, dtype=torch.float32)
            call("npf_setconv_fwd", _p(keys), key_bs, _p(queries), qry_bs, _p(values), _p(theta), _p(feat), _p(dens),
                 _p(mstat), B, K, Q, C, int(keys_regular), C, 1, _stream())
            if M > 0:
                call("npf_linear_fwd", _p(feat), C, _p(W), C + 1, _p(b), _p(out), N, M, C, N, 0, _p(dens),
                     W.data_ptr() + 4 * C, C + 1, _precision, _stream())
        ctx.save_for_backward(keys, queries, values, theta, W, feat, dens, mstat)
        ctx.bias_ref = b
        ctx.dims = (B, K, Q, C, N, key_bs, qry_bs, int(keys_regular), ilv)
        return out

    @staticmethod
    def backward(ctx, dout):
        keys, queries, values, theta, W, feat, dens, mstat = ctx.saved_tensors
        B, K, Q, C, N, key_bs, qry_bs, regular, ilv = ctx.dims
        dout = _c(dout).reshape(B * Q, N)
        M = B * Q
        dW, rW = _gbuf(W)
        db, rb = _gbuf(ctx.bias_ref)
        dtheta, rtheta = _gbuf(theta)
        dvalues = None
        if M > 0 and ilv:
            dxe = torch.empty(M, C + 1, device=dout.device, dtype=torch.float32)   # [dfeat | ddens]
            call("npf_linear_bwd", _p(dout), N, _p(feat), C + 1, _p(W), C + 1, _p(dxe), C + 1, _p(dW), C + 1, _p(db), M,
                 C + 1, N, 0, _precision, _stream())
            call("npf_setconv_bwd", _p(keys), key_bs, _p(queries), qry_bs, _p(values), _p(theta), _p(feat),
                 feat.data_ptr() + 4 * C, _p(mstat), _p(dxe), dxe.data_ptr() + 4 * C, None, _p(dtheta), B, K, Q, C, 0,
                 C + 1, C + 1, _stream())
        elif M > 0:
            # feature block of the resizer: dfeat, dW[:, :C], db in one pass; density column: ddens and dW[:, C] in a thin one
            dfeat = _lin_bwd(dout, feat, _p(W), C + 1, _p(dW), C + 1, db, M, C, N, mask=False)
            ddens = _lin_bwd(dout, dens, W.data_ptr() + 4 * C, C + 1, dW.data_ptr() + 4 * C, C + 1, None, M, 1, N, mask=False)
            if ctx.needs_input_grad[2]:
                dvalues = torch.empty_like(values)
            call("npf_setconv_bwd", _p(keys), key_bs, _p(queries), qry_bs, _p(values), _p(theta), _p(feat), _p(dens),
                 _p(mstat), _p(dfeat), _p(ddens), _p(dvalues), _p(dtheta), B, K, Q, C, regular, C, 1, _stream())
        elif ctx.needs_input_grad[2]:
            dvalues = torch.zeros_like(values)
        return None, None, dvalues, rtheta, rW, rb, None


def setconv(keys, queries, values, theta, weight, bias, keys_regular=False):
    """keys [B,K,1] or shared [K]; queries [B,Q,1] or shared [Q]; values [B,K,C]; weight [N, C+1]; -> [B,Q,N]."""
    if keys.dim() == 3:
        keys = keys.squeeze(-1)
    if queries.dim() == 3:
        queries = queries.squeeze(-1)
    return _SetConv.apply(keys, queries, values, theta, weight, bias, keys_regular)


# ======================================================================================================
# Depthwise conv (+ folded pre-activation affine) and batch statistics
# ======================================================================================================
class _DWConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wt, bias, res, relu_in, scale, shift):
        _chk(x, Wt, bias, res, scale, shift)
        x = _c(x)
        _shape(x.dim() in (3, 4), f"dwconv: x must be [B,L,C] or [B,H,W,C], got {tuple(x.shape)}")
        B, C = x.shape[0], x.shape[-1]
        _shape(Wt.dim() == x.dim() and Wt.shape[0] == C and Wt.shape[1] == 1, f"dwconv: weight {tuple(Wt.shape)} does not match {C} channels")
        _shape(res is None or res.shape == x.shape, "dwconv: residual must have the shape of x")
        _shape(bias is None or bias.numel() == C, "dwconv: bias must have one entry per channel")
        _shape((scale is None) == (shift is None) and (scale is None or (scale.numel() == C and shift.numel() == C)),
               "dwconv: scale/shift must both be given with one entry per channel")
        if x.dim() == 3:
            H, Wd = 1, x.shape[1]
            kh, kw = 1, Wt.shape[-1]
        else:
            H, Wd = x.shape[1], x.shape[2]
            kh, kw = Wt.shape[-2], Wt.shape[-1]
        assert Wt.is_contiguous(), "depthwise weight must be contiguous"
        y = torch.empty_like(x)
        res_c = None if res is None else _c(res)
        call("npf_dwconv_fwd", _p(x), _p(Wt), _p(bias), _p(res_c), _p(y), B, H, Wd, C, kh, kw,
             RELU_IN if relu_in else 0, _p(scale), _p(shift), _stream())
        ctx.save_for_backward(x, Wt, scale, shift, bias)
        # residual == conv input (ResConvBlock with one conv layer): its gradient is folded into the dX kernel
        res_is_x = res is not None and res_c.data_ptr() == x.data_ptr() and res_c.shape == x.shape
        ctx.cfg = (B, H, Wd, C, kh, kw, relu_in, bias is not None, res is not None, res_is_x)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wt, scale, shift, bias = ctx.saved_tensors
        B, H, Wd, C, kh, kw, relu_in, has_bias, has_res, res_is_x = ctx.cfg
        dy = _c(dy)
        fuse_res = res_is_x and ctx.needs_input_grad[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW, rW = _gbuf(Wt)
        db, rb = _gbuf(bias)
        need_aff = scale is not None and (ctx.needs_input_grad[5] or ctx.needs_input_grad[6])
        dscale = torch.zeros_like(scale) if need_aff else None
        dshift = torch.zeros_like(shift) if need_aff else None
        if need_aff and dx is None:
            dx = torch.empty_like(x)
        call("npf_dwconv_bwd", _p(dy), _p(x), _p(Wt), _p(dx), _p(dW), _p(db), B, H, Wd, C, kh, kw,
             (RELU_IN if relu_in else 0) | (ADD_DY if fuse_res else 0), _p(scale), _p(shift), _p(dscale), _p(dshift),
             _stream())
        dres = None if (fuse_res or not has_res) else dy
        return (dx if ctx.needs_input_grad[0] else None, rW, rb, dres, None, dscale, dshift)


def dwconv(x, weight, bias=None, res=None, relu_in=False, scale=None, shift=None):
    """Channel-last depthwise conv.  x [B,L,C] (1-D) or [B,H,W,C] (2-D); weight [C,1,k] / [C,1,k,k]."""
    return _DWConv.apply(x, weight, bias, res, relu_in, scale, shift)


class _ResBlock1d(torch.autograd.Function):
    """y = pw(dw_k(relu(x)) + b_dw + x) + b_pw for [B,L,128] signals: ONE kernel per direction (npf_resblock1d_fwd / _bwd: raw rows
    by TMA, depthwise in registers out of shared memory, pointwise on tcgen05).  Only x is saved: the backward recomputes the
    intermediate O, so neither O nor its gradient ever exists in HBM."""

    @staticmethod
    def forward(ctx, x, wdw, bdw, wpw, bpw):
        _chk(x, wdw, bdw, wpw, bpw)
        x = _c(x)
        B, L, C = x.shape
        k = wdw.shape[-1]
        y = torch.empty_like(x)
        call("npf_resblock1d_fwd", _p(x), _p(wdw), _p(bdw), _p(wpw), _p(bpw), None, _p(y), B, L, C, k, _precision, _stream())
        ctx.save_for_backward(x, wdw, wpw, bdw)
        ctx.refs = (bdw, bpw)
        ctx.cfg = (B, L, C, k, _precision)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wdw, wpw, bdw_saved = ctx.saved_tensors
        bdw, bpw = ctx.refs
        B, L, C, k, prec = ctx.cfg
        dy = _c(dy)
        dWp, rWp = _gbuf(wpw)
        dbp, rbp = _gbuf(bpw)
        dWd, rWd = _gbuf(wdw)
        dbd, rbd = _gbuf(bdw)
        dx = torch.empty_like(x)
        call("npf_resblock1d_bwd", _p(dy), _p(x), _p(wdw), _p(bdw_saved), _p(wpw), _p(dx), _p(dWd), _p(dbd), _p(dWp), _p(dbp), B, L, C, k, prec,
             _stream())
        return dx, rWd, rbd, rWp, rbp


def resblock1d_supported(x, dw_weight, pw_weight):
    """The fused block kernel covers the ConvCNP default: [B,L,128] signals, kernel size 11, the bf16x3 arithmetic mode."""
    return (x.dim() == 3 and x.shape[-1] == 128 and dw_weight.shape[-1] == 11 and dw_weight.shape[0] == 128 and pw_weight.shape[0] == 128
            and pw_weight.numel() == 128 * 128 and _precision == _PRECISION["bf16x3"] and x.is_cuda and x.shape[0] > 0)


def resblock1d(x, dw_weight, dw_bias, pw_weight, pw_bias):
    """Fused pre-activation residual block (depthwise k + residual + pointwise); x [B,L,128] channel-last."""
    return _ResBlock1d.apply(x, dw_weight, dw_bias, pw_weight, pw_bias)


class _ChannelMoments(torch.autograd.Function):
    """mean[c], biased var[c] of a channel-last tensor (two passes, like ATen's batch_norm statistics)."""

    @staticmethod
    def forward(ctx, x):
        _chk(x)
        x = _c(x)
        C = x.shape[-1]
        M = x.numel() // C
        s = torch.zeros(C, device=x.device, dtype=torch.float32)
        call("npf_channel_stats", _p(x), None, _p(s), None, M, C, _stream())
        mean = s / M
        q = torch.zeros(C, device=x.device, dtype=torch.float32)
        call("npf_channel_stats", _p(x), _p(mean), None, _p(q), M, C, _stream())
        var = q / M
        ctx.save_for_backward(x, mean)
        ctx.M = M
        return mean, var

    @staticmethod
    def backward(ctx, dmean, dvar):
        x, mean = ctx.saved_tensors
        M = ctx.M
        a = (2.0 / M) * dvar
        b = dmean / M - a * mean
        dx = torch.empty_like(x)
        call("npf_channel_affine", _p(x), _p(_c(a)), _p(_c(b)), _p(dx), M, x.shape[-1], 0, _stream())
        return dx


def channel_moments(x):
    return _ChannelMoments.apply(x)


# ======================================================================================================
# sum-merge / pooling / layernorm
# ======================================================================================================
class _MergeRelu(torch.autograd.Function):
    """out[z,b,t,:] = relu(x1[b,t,:] + x2[z,b,(t|0),:])"""

    @staticmethod
    def forward(ctx, x1, x2, x2_has_t):
        _chk(x1, x2)
        x1, x2 = _c(x1), _c(x2)
        _shape(x1.dim() == 3 and x2.dim() == 4, f"merge_relu: x1 must be [B,T,C] and x2 [Z,B,T|1,C], got {tuple(x1.shape)}, {tuple(x2.shape)}")
        B, T, C = x1.shape
        Z = x2.shape[0]
        _shape(x2.shape[1] == B and x2.shape[3] == C and x2.shape[2] == (T if x2_has_t else 1),
               f"merge_relu: x2 {tuple(x2.shape)} does not broadcast against x1 {tuple(x1.shape)}")
        out = torch.empty(Z, B, T, C, device=x1.device, dtype=torch.float32)
        call("npf_merge_relu_fwd", _p(x1), _p(x2), _p(out), Z, B, T, C, int(x2_has_t), _stream())
        ctx.save_for_backward(out)
        ctx.cfg = (Z, B, T, C, int(x2_has_t), x2.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        Z, B, T, C, has_t, x2_shape = ctx.cfg
        dout = _c(dout)
        dx1 = torch.empty(B, T, C, device=out.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        dx2 = torch.empty(x2_shape, device=out.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        call("npf_merge_relu_bwd", _p(dout), _p(out), _p(dx1), _p(dx2), Z, B, T, C, has_t, _stream())
        return dx1, dx2, None


def merge_relu(x1, x2):
    """x1 [B,T,C]; x2 [B,T,C] | [Z,B,T,C] | [Z,B,1,C]  ->  [B,T,C] (if x2 is 3-D) else [Z,B,T,C]."""
    squeeze = x2.dim() == 3
    if squeeze:
        x2 = x2.unsqueeze(0)
    _shape(x1.dim() == 3 and x2.dim() == 4 and x2.shape[2] in (1, x1.shape[1]),
           f"merge_relu: x2 {tuple(x2.shape)} must have the {x1.shape[1]} targets of x1 or a singleton target axis")
    has_t = x2.shape[2] == x1.shape[1]
    out = _MergeRelu.apply(x1, x2, has_t)
    return out.squeeze(0) if squeeze else out


class _MeanPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _chk(x)
        x = _c(x)
        _shape(x.dim() == 3, f"mean_pool: x must be [B,N,C], got {tuple(x.shape)}")
        B, N, C = x.shape
        r = torch.empty(B, 1, C, device=x.device, dtype=torch.float32)
        call("npf_mean_pool_fwd", _p(x), _p(r), B, N, C, _stream())
        ctx.cfg = (B, N, C)
        return r

    @staticmethod
    def backward(ctx, dr):
        B, N, C = ctx.cfg
        dx = torch.empty(B, N, C, device=dr.device, dtype=torch.float32)
        call("npf_mean_pool_bwd", _p(_c(dr)), _p(dx), B, N, C, _stream())
        return dx


def mean_pool(x):
    """[B,N,C] -> [B,1,C] mean over the set axis."""
    return _MeanPool.apply(x)


class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, gamma, beta):
        _chk(a, b, gamma, beta)
        a, b = _c(a), _c(b)
        C = a.shape[-1]
        _shape(a.shape == b.shape and gamma.numel() == C and beta.numel() == C,
               f"add_layernorm: a {tuple(a.shape)}, b {tuple(b.shape)}, gamma/beta [{gamma.numel()}]/[{beta.numel()}] do not agree")
        M = a.numel() // C
        y = torch.empty_like(a)
        rstat = torch.empty(M, 2, device=a.device, dtype=torch.float32)
        call("npf_add_layernorm_fwd", _p(a), _p(b), _p(gamma), _p(beta), _p(y), _p(rstat), M, C, _stream())
        ctx.save_for_backward(a, b, gamma, rstat)
        ctx.beta_ref = beta
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b, gamma, rstat = ctx.saved_tensors
        C = a.shape[-1]
        M = a.numel() // C
        ds = torch.empty_like(a)
        dg, rg = _gbuf(gamma)
        dbeta, rbeta = _gbuf(ctx.beta_ref)
        call("npf_add_layernorm_bwd", _p(_c(dy)), _p(a), _p(b), _p(gamma), _p(rstat), _p(ds), _p(dg), _p(dbeta), M, C,
             _stream())
        return ds, ds, rg, rbeta


def add_layernorm(a, b, gamma, beta):
    """LayerNorm(a + b) over the last dim (eps 1e-5, affine)."""
    return _AddLayerNorm.apply(a, b, gamma, beta)


# ======================================================================================================
# attention
# ======================================================================================================
class _XAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, n_heads, scale):
        _chk(q, k, v)
        q, k, v = _c(q), _c(k), _c(v)
        _shape(q.dim() == 3 and k.dim() == 3 and v.dim() == 3, "xattn: q, k, v must be [B,T,E]")
        _shape(k.shape[0] == q.shape[0] and v.shape[0] == q.shape[0] and k.shape[1] == v.shape[1] and k.shape[2] == q.shape[2],
               f"xattn: q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)} do not agree (k/v share B and Tk; q/k share E)")
        _shape(q.shape[2] % n_heads == 0 and v.shape[2] % n_heads == 0, f"xattn: {n_heads} heads do not divide E={q.shape[2]} / Ev={v.shape[2]}")
        B, Tq, E = q.shape
        Tk = k.shape[1]
        Ev = v.shape[2]
        D, Dv = E // n_heads, Ev // n_heads
        o = torch.empty(B, Tq, Ev, device=q.device, dtype=torch.float32)
        lse = torch.empty(B, n_heads, Tq, device=q.device, dtype=torch.float32)
        call("npf_xattn_fwd", _p(q), _p(k), _p(v), _p(o), _p(lse), B, Tq, Tk, n_heads, D, Dv, float(scale), _precision,
             _stream())
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.cfg = (B, Tq, Tk, n_heads, D, Dv, float(scale))
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        B, Tq, Tk, H, D, Dv, scale = ctx.cfg
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        call("npf_xattn_bwd", _p(q), _p(k), _p(v), _p(o), _p(lse), _p(_c(do)), _p(dq), _p(dk), _p(dv), B, Tq, Tk, H, D, Dv,
             scale, _precision, _stream())
        return dq, dk, dv, None, None


def xattn(q, k, v, n_heads, scale):
    """softmax(q k^T * scale) v per head; head h = channels [h*D, (h+1)*D).  q [B,Tq,E], k [B,Tk,E], v [B,Tk,Ev]."""
    return _XAttn.apply(q, k, v, n_heads, scale)


# ======================================================================================================
# predictive head / log-likelihood
# ======================================================================================================
class _GaussHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, suff, min_scale):
        _chk(suff)
        suff = _c(suff)
        _shape(suff.shape[-1] % 2 == 0 and suff.shape[-1] > 0, f"gauss_head: last dim {suff.shape[-1]} must be 2*y_dim")
        y = suff.shape[-1] // 2
        M = suff.numel() // (2 * y)
        loc = torch.empty(*suff.shape[:-1], y, device=suff.device, dtype=torch.float32)
        scale = torch.empty_like(loc)
        call("npf_gauss_head_fwd", _p(suff), _p(loc), _p(scale), M, y, float(min_scale), _stream())
        ctx.save_for_backward(suff)
        ctx.cfg = (M, y, float(min_scale))
        return loc, scale

    @staticmethod
    def backward(ctx, dloc, dscale):
        (suff,) = ctx.saved_tensors
        M, y, min_scale = ctx.cfg
        dsuff = torch.empty_like(suff)
        call("npf_gauss_head_bwd", _p(suff), _p(None if dloc is None else _c(dloc)), _p(None if dscale is None else _c(dscale)),
             _p(dsuff), M, y, min_scale, _stream())
        return dsuff, None


def gauss_head(suff, min_scale=0.01):
    """suff [..., 2y] -> loc [..., y], scale = min_scale + (1 - min_scale) softplus(.) [..., y]."""
    return _GaussHead.apply(suff, min_scale)


class _GaussSLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, scale, Y):
        _chk(loc, scale, Y)
        loc, scale, Y = _c(loc), _c(scale), _c(Y)
        _shape(loc.dim() >= 3 and loc.shape == scale.shape, f"gauss_sum_log_prob: loc {tuple(loc.shape)} / scale {tuple(scale.shape)} must be equal [Z,B,*,y]")
        _shape(tuple(Y.shape) == tuple(loc.shape[1:]), f"gauss_sum_log_prob: Y {tuple(Y.shape)} must be {tuple(loc.shape[1:])} (targets of loc/scale)")
        Z, B = loc.shape[0], loc.shape[1]
        n = loc.numel() // max(Z * B, 1) if Z * B > 0 else 0
        slp = torch.empty(Z, B, device=loc.device, dtype=torch.float32)
        call("npf_gauss_nll_fwd", _p(loc), _p(scale), _p(Y), _p(slp), Z, B, n, _stream())
        ctx.save_for_backward(loc, scale, Y)
        ctx.cfg = (Z, B, n)
        return slp

    @staticmethod
    def backward(ctx, g):
        loc, scale, Y = ctx.saved_tensors
        Z, B, n = ctx.cfg
        dloc, dscale = torch.empty_like(loc), torch.empty_like(scale)
        call("npf_gauss_nll_bwd", _p(loc), _p(scale), _p(Y), _p(_c(g)), _p(dloc), _p(dscale), Z, B, n, _stream())
        return dloc, dscale, None


def gauss_sum_log_prob(loc, scale, Y):
    """sum over targets and y of log N(Y; loc, scale):  loc/scale [Z,B,*,y], Y [B,*,y] -> [Z,B]."""
    return _GaussSLP.apply(loc, scale, Y)


# ======================================================================================================
# latent path
# ======================================================================================================
class _LatentSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, suff, eps):
        _chk(suff, eps)
        suff, eps = _c(suff), _c(eps)
        _shape(suff.shape[-1] % 2 == 0 and eps.dim() == suff.dim() + 1 and tuple(eps.shape[1:-1]) == tuple(suff.shape[:-1])
               and eps.shape[-1] == suff.shape[-1] // 2,
               f"latent_sample: eps {tuple(eps.shape)} must be [S, *{tuple(suff.shape[:-1])}, {suff.shape[-1] // 2}]")
        zd = suff.shape[-1] // 2
        M = suff.numel() // (2 * zd)
        S = eps.shape[0]
        q_loc = torch.empty(*suff.shape[:-1], zd, device=suff.device, dtype=torch.float32)
        q_scale = torch.empty_like(q_loc)
        z = torch.empty_like(eps)
        call("npf_latent_sample_fwd", _p(suff), _p(eps), _p(q_loc), _p(q_scale), _p(z), S, M, zd, _stream())
        ctx.save_for_backward(suff, eps)
        ctx.cfg = (S, M, zd)
        return q_loc, q_scale, z

    @staticmethod
    def backward(ctx, dq_loc, dq_scale, dz):
        suff, eps = ctx.saved_tensors
        S, M, zd = ctx.cfg
        dsuff = torch.empty_like(suff)
        call("npf_latent_sample_bwd", _p(suff), _p(eps), _p(None if dz is None else _c(dz)),
             _p(None if dq_loc is None else _c(dq_loc)), _p(None if dq_scale is None else _c(dq_scale)), _p(dsuff), S, M,
             zd, _stream())
        return dsuff, None


def latent_sample(suff, eps):
    """suff [*lat, 2z], eps [S, *lat, z] -> (q_loc, q_scale [*lat, z], z [S, *lat, z]); q_scale = 0.1 + 0.9 sigmoid."""
    return _LatentSample.apply(suff, eps)


class _GlobalLatent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        _chk(z)
        z = _c(z)
        _shape(z.dim() >= 3 and z.shape[-1] % 2 == 0, f"global_latent: z must be [N, *spatial, C] with even C, got {tuple(z.shape)}")
        N, C = z.shape[0], z.shape[-1]
        P = z.numel() // (N * C)
        out = torch.empty_like(z)
        call("npf_global_latent_fwd", _p(z), _p(out), N, P, C, _stream())
        ctx.cfg = (N, P, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, P, C = ctx.cfg
        dout = _c(dout)
        dz = torch.empty_like(dout)
        call("npf_global_latent_bwd", _p(dout), _p(dz), N, P, C, _stream())
        return dz


def global_latent(z):
    """[N, *spatial, C]: second half of the channels replaced by its mean over all spatial positions."""
    return _GlobalLatent.apply(z)


# ======================================================================================================
# on-grid context encoding
# ======================================================================================================
class _GridConvIn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, mask_u8, Wt):
        _chk(img, mask_u8, Wt)
        img, mask_u8 = _c(img), _c(mask_u8)
        assert Wt.is_contiguous()
        _shape(img.dim() == 4 and mask_u8.dim() == 4 and mask_u8.shape[:3] == img.shape[:3] and mask_u8.shape[3] in (1, img.shape[3]),
               f"gridconv_in: image {tuple(img.shape)} (channel-last [B,H,W,y]) and mask {tuple(mask_u8.shape)} do not agree")
        _shape(Wt.dim() == 4 and Wt.shape[0] == img.shape[3] and Wt.shape[1] == 1 and Wt.shape[2] == Wt.shape[3] and Wt.shape[2] % 2 == 1,
               f"gridconv_in: weight {tuple(Wt.shape)} must be [y,1,k,k] with odd k")
        B, H, Wd, y = img.shape
        k = Wt.shape[-1]
        feat = torch.empty(B, H, Wd, 2 * y, device=img.device, dtype=torch.float32)
        call("npf_gridconv_in_fwd", _p(img), _p(mask_u8), mask_u8.shape[-1], _p(Wt), _p(feat), B, H, Wd, y, k, _stream())
        ctx.save_for_backward(img, mask_u8, Wt, feat)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        img, mask_u8, Wt, feat = ctx.saved_tensors
        B, H, Wd, y = img.shape
        k = Wt.shape[-1]
        dW, rW = _gbuf(Wt)
        call("npf_gridconv_in_bwd", _p(img), _p(mask_u8), mask_u8.shape[-1], _p(Wt), _p(feat), _p(_c(dfeat)), _p(dW), B, H,
             Wd, y, k, _stream())
        return None, None, rW


def gridconv_in(img, mask, weight):
    """img [B,H,W,y] fp32, mask [B,H,W,1|y] bool/uint8, weight [y,1,k,k] (abs applied inside) -> [B,H,W,2y]."""
    if mask.dtype != torch.uint8:
        mask = mask.to(torch.uint8)
    return _GridConvIn.apply(img, mask, weight)


# ======================================================================================================
# input validation
# ======================================================================================================
def range_flag(flag, *tensors, lo=-1.0, hi=1.0):
    """OR into the device int32 ``flag`` whether any element of the tensors lies outside [lo, hi] (no sync)."""
    for t in tensors:
        if t.numel() == 0:
            continue
        _chk(t)
        t = _c(t)
        call("npf_range_check", _p(t), t.numel(), float(lo), float(hi), _p(flag), _stream())
