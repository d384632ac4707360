"""CPU oracle for the Neural-Process hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this file.  Nothing under ``neural-process-family_b200/`` imports it; the product
path is CUDA-only and raises if its extension is missing.

What it is: a *functional* restatement (plain functions over a ``state_dict``; no ``nn.Module``) of
the algorithm the reference implements in ``npf/`` for the per-task forward / loss of CNP, AttnCNP,
ConvCNP, GridConvCNP, ConvLNP, GridConvLNP (+ LNP, AttnLNP, self-attention encoders).  Arithmetic is torch CPU (fp32 or fp64), which is
the reference's own third-party arithmetic (every FLOP of the reference is an
ATen op); gradients come from torch autograd over this restatement.  Each function cites the
reference ``file:line`` it follows (paths relative to the upstream repo root).

Pinned: ``tests/test_oracle_golden.py`` checks every function here against the fixtures in
``tests/golden/*.pt`` which were produced by the *real* reference imported in the build container
(``oracle/gen_golden.py``, committed), including the upstream pretrained checkpoints.

The op order deliberately mirrors the reference (e.g. the ``[B, n_q, n_k, C]`` broadcast multiply of
SetConv is materialised exactly as upstream does) so that timing this file on host cores is a
faithful "reference CPU path" baseline.
"""
import math

import torch
import torch.nn.functional as F

__all__ = [
    "mlp", "merge_flat_sum", "setconv", "res_conv_cnn", "dot_attention", "multihead_attention",
    "transformer_attention", "self_attention", "cnp_forward", "attncnp_forward", "attnlnp_forward", "convcnp_forward",
    "gridconvcnp_forward", "lnp_forward", "convlnp_forward", "gridconvlnp_forward",
    "gauss_sum_log_prob", "cnpf_loss", "nll_lnpf_loss", "elbo_lnpf_loss", "induced_grid",
    "p_y_scale", "q_z_scale",
]


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    """Sub-dictionary of ``sd`` whose keys start with ``prefix`` (prefix stripped)."""
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _lin(sd, prefix, x):
    """nn.Linear: y = x W^T + b (bias optional)."""
    return F.linear(x, sd[prefix + "weight"], sd.get(prefix + "bias"))


def p_y_scale(s):
    """Predictive std transform, npf/neuralproc/base.py:116."""
    return 0.01 + 0.99 * F.softplus(s)


def q_z_scale(s):
    """Latent std transform, npf/neuralproc/base.py:432."""
    return 0.1 + 0.9 * torch.sigmoid(s)


def induced_grid(density_induced=128, lo=-1.5, hi=1.5, dtype=torch.float32):
    """Induced (pseudo) inputs of ConvCNP, npf/neuralproc/convnp.py:104 (and 170-181 for
    ``set_extrapolation``: ``lo = min - 0.5``, ``hi = max + 0.5``, ``int(density*(hi-lo))`` points).
    The reference always builds the grid in fp32 (torch.linspace default)."""
    n = int(density_induced * (hi - lo))
    return torch.linspace(lo, hi, n).to(dtype)


# --------------------------------------------------------------------------------------------
# building blocks (npf/architectures)
# --------------------------------------------------------------------------------------------
def mlp(sd, prefix, x):
    """MLP.forward, npf/architectures/mlp.py:95-109 (activation ReLU, no dropout, is_res=False):
    out(relu(linears[n-2](... relu(to_hidden(x)))))."""
    h = torch.relu(_lin(sd, prefix + "to_hidden.", x))
    i = 0
    while (prefix + f"linears.{i}.weight") in sd:
        h = torch.relu(_lin(sd, prefix + f"linears.{i}.", h))
        i += 1
    return _lin(sd, prefix + "out.", h)


def merge_flat_sum(sd, prefix, x1, x2):
    """MergeFlatInputs.forward with is_sum_merge=True, npf/architectures/encoders.py:175-183:
    flat_module(relu(x1 + resizer(x2))).  ``flat_module`` is an MLP or a bare Linear."""
    h = torch.relu(x1 + mlp(sd, prefix + "resizer.", x2))
    return _flat(sd, prefix + "flat_module.", h)


def _flat(sd, prefix, h):
    if (prefix + "to_hidden.weight") in sd:
        return mlp(sd, prefix, h)
    return _lin(sd, prefix, h)  # nn.Linear decoder (ConvLNP default, convnp.py:249)


def setconv(sd, prefix, keys, queries, values):
    """SetConv.forward + ExpRBF.forward, npf/architectures/setcnn.py:234-268 and 126-142.

    keys [B,K,1], queries [B,Q,1], values [B,K,Cin] -> [B,Q,Cout].
    sigma = 1e-5 + softplus(theta); a = -(|xk - xq| / sigma)^2; weights = softmax_k(a);
    density = sum_k exp(a); out = Linear([sum_k w v ; density])."""
    k = keys.unsqueeze(1)
    q = queries.unsqueeze(2)
    v = values.unsqueeze(1)
    dist = torch.norm(k - q, p=2, dim=-1, keepdim=True)  # [B,Q,K,1]
    sigma = 1e-5 + F.softplus(sd[prefix + "radial_basis_func.length_scale_param"])
    inp = -(dist / sigma).pow(2)
    weight = torch.softmax(inp, dim=-2)
    density = torch.exp(inp).sum(dim=-2)  # [B,Q,1]
    targets = (weight * v).sum(dim=2)  # the [B,Q,K,Cin] broadcast of setcnn.py:263
    targets = torch.cat([targets, density], dim=-1)
    return _lin(sd, prefix + "resizer.", targets)


def _norm(sd, prefix, X, training, bn_eps=1e-5):
    """Normalization(in_chan): nn.Identity (no keys) or BatchNorm{1,2}d (affine + running stats),
    channels on dim 1.  Train mode uses batch statistics (biased variance), like nn.BatchNorm."""
    if (prefix + "weight") not in sd:
        return X
    return F.batch_norm(
        X, sd[prefix + "running_mean"].clone(), sd[prefix + "running_var"].clone(),
        sd[prefix + "weight"], sd[prefix + "bias"], training=training, momentum=0.1, eps=bn_eps,
    )


def _conv(sd, prefix, X, groups, circular=False):
    w = sd[prefix + "weight"]
    b = sd.get(prefix + "bias")
    pad = w.shape[-1] // 2
    if circular and pad > 0:  # make_padded_conv(Conv, CircularPad2d), npf/utils/helpers.py:334-351, 406-414
        return F.conv2d(F.pad(X, (pad,) * 4, mode="circular"), w, b, padding=0, groups=groups)
    if w.dim() == 3:
        return F.conv1d(X, w, b, padding=pad, groups=groups)
    return F.conv2d(X, w, b, padding=pad, groups=groups)


def res_conv_cnn(sd, prefix, X, training=True, circular=False, bn_eps=1e-5):
    """CNN.forward (is_chan_last=True) over ResConvBlocks, npf/architectures/cnn.py:363-380, 204-215;
    depth-separable conv of npf/utils/helpers.py:354-403.

    X channel-last [B, *spatial, C].  Per block:
        h   = pw1(dw1(relu(norm1(X))))            only if n_conv_layers == 2
        out = pw2(dw2(relu(norm2(h))) + X)
    """
    nd = X.dim()
    X = X.permute(0, nd - 1, *range(1, nd - 1))  # channels_to_2nd_dim, helpers.py:60-65
    C = X.shape[1]
    i = 0
    while (prefix + f"conv_blocks.{i}.conv2_depthwise.weight") in sd:
        p = prefix + f"conv_blocks.{i}."
        if (p + "conv1.depthwise.weight") in sd:
            h = _conv(sd, p + "conv1.depthwise.", torch.relu(_norm(sd, p + "norm1.", X, training, bn_eps)), C, circular)
            h = _conv(sd, p + "conv1.pointwise.", h, 1)
        else:
            h = X
        h = _conv(sd, p + "conv2_depthwise.", torch.relu(_norm(sd, p + "norm2.", h, training, bn_eps)), C, circular)
        h = h + X
        X = _conv(sd, p + "conv2_pointwise.", h.contiguous(), 1)
        i += 1
    return X.permute(0, *range(2, nd), 1)  # channels_to_last_dim, helpers.py:68-73


def dot_attention(keys, queries, values, is_scale=True):
    """BaseAttender.forward + DotAttender.score, npf/architectures/attention.py:129-156, 204-220:
    softmax_k(q.k / sqrt(d)) @ values."""
    logits = torch.einsum("bkd,bqd->bqk", keys, queries)
    if is_scale:
        logits = logits / math.sqrt(queries.size(-1))
    attn = logits.softmax(dim=-1)
    return torch.bmm(attn, values)


def multihead_attention(sd, prefix, keys, queries, values, n_heads=8):
    """MultiheadAttender.forward, npf/architectures/attention.py:457-527: K = Wk k (no bias),
    Q = Wq q + bq, V = Wv v (no bias); heads are contiguous channel slices stacked head-major on the
    batch axis; per-head scaled-dot; concat; optional post_processor Linear."""
    K = _lin(sd, prefix + "key_transform.", keys)
    Q = _lin(sd, prefix + "query_transform.", queries)
    V = _lin(sd, prefix + "value_transform.", values)

    def split(t):  # _make_multiheaded 507-516
        b, n, d = t.shape
        return t.view(b, n, n_heads, d // n_heads).permute(2, 0, 1, 3).contiguous().view(b * n_heads, n, d // n_heads)

    ctx = dot_attention(split(K), split(Q), split(V))
    b = keys.shape[0]
    hd = ctx.shape[-1]
    ctx = ctx.view(n_heads, b, -1, hd).permute(1, 2, 0, 3).contiguous().view(b, -1, n_heads * hd)  # 518-527
    if (prefix + "post_processor.weight") in sd:
        ctx = _lin(sd, prefix + "post_processor.", ctx)
    return ctx


def transformer_attention(sd, prefix, keys, queries, values, n_heads=8):
    """TransformerAttender.forward, npf/architectures/attention.py:569-588:
    c = LN1(MHA + queries); c = LN2(c + MLP(c))."""
    ctx = multihead_attention(sd, prefix, keys, queries, values, n_heads)
    d = ctx.shape[-1]
    ctx = F.layer_norm(ctx + queries, (d,), sd[prefix + "layer_norm1.weight"], sd[prefix + "layer_norm1.bias"], 1e-5)
    ctx = F.layer_norm(ctx + mlp(sd, prefix + "mlp.", ctx), (d,), sd[prefix + "layer_norm2.weight"], sd[prefix + "layer_norm2.bias"], 1e-5)
    return ctx


def self_attention(sd, prefix, X, attention="transformer", n_attn_layers=2, n_heads=8):
    """SelfAttention.forward, npf/architectures/selfattn.py:82-100 with positional=None: every layer
    attends the set to itself (keys = queries = values = previous output); Linear ``resize`` at the end
    when an output size was given."""
    out = X
    for i in range(n_attn_layers):
        pre = prefix + f"attn_layers.{i}."
        if attention == "scaledot":
            out = dot_attention(out, out, out)
        elif attention == "multihead":
            out = multihead_attention(sd, pre, out, out, out, n_heads)
        elif attention == "transformer":
            out = transformer_attention(sd, pre, out, out, out, n_heads)
        else:
            raise ValueError(f"Unknown attention method {attention}")
    if (prefix + "resize.weight") in sd:
        out = _lin(sd, prefix + "resize.", out)
    return out


def _attn_xy_encode(sd, Xe, Y, is_self_attn, self_attention_type="transformer"):
    """The per-point context encoder of AttnCNP / AttnLNP: MergeFlatInputs around an MLP, or around
    SelfAttention when ``is_self_attn`` (npf/neuralproc/attnnp.py:88-96).  The self-attention layers
    take their type from ``self_attention_kwargs`` (default "transformer", selfattn.py:50), NOT from the
    ``attention`` argument, which only selects the cross-attention."""
    if not is_self_attn:
        return merge_flat_sum(sd, "xy_encoder.", Xe, Y)
    h = torch.relu(Xe + mlp(sd, "xy_encoder.resizer.", Y))
    return self_attention(sd, "xy_encoder.flat_module.", h, self_attention_type)


def _attend(sd, Xe_c, Xe_t, R_c, attention, n_heads):
    if attention == "scaledot":
        return dot_attention(Xe_c, Xe_t, R_c)
    if attention == "multihead":
        return multihead_attention(sd, "attender.", Xe_c, Xe_t, R_c, n_heads)
    if attention == "transformer":
        return transformer_attention(sd, "attender.", Xe_c, Xe_t, R_c, n_heads)
    raise ValueError(f"Unknown attention method {attention}")


# --------------------------------------------------------------------------------------------
# models (npf/neuralproc): each returns (loc, scale) of shape [n_z, B, *n_trgt, y_dim] (+ latents)
# --------------------------------------------------------------------------------------------
def _decode(sd, Xe_t, R_trgt, y_dim):
    """NeuralProcessFamily.decode, npf/neuralproc/base.py:327-367 (heteroskedastic).
    Decoder is MergeFlatInputs (CNP/AttnCNP, keys ``decoder.resizer``/``decoder.flat_module``) or
    DiscardIthArg (Conv*, key ``decoder.destination``, encoders.py:105-120)."""
    if "decoder.resizer.to_hidden.weight" in sd:
        suff = merge_flat_sum(sd, "decoder.", Xe_t, R_trgt)
    else:
        suff = _flat(sd, "decoder.destination.", R_trgt)
    loc, s = suff.split(y_dim, dim=-1)
    return loc, p_y_scale(s)


def cnp_forward(sd, X_cntxt, Y_cntxt, X_trgt):
    """CNP, npf/neuralproc/base.py:177-239 + np.py:86-110.  R = mean_c xy_encoder(x_enc(Xc), Yc);
    zeros when n_cntxt == 0 (np.py:97-99)."""
    y_dim = Y_cntxt.shape[-1]
    Xe_c = mlp(sd, "x_encoder.", X_cntxt)
    Xe_t = mlp(sd, "x_encoder.", X_trgt)
    B, C, _ = Xe_c.shape
    r_dim = sd["xy_encoder.flat_module.out.weight"].shape[0]
    if C == 0:
        R = torch.zeros(B, 1, r_dim, dtype=Xe_t.dtype)
    else:
        R = merge_flat_sum(sd, "xy_encoder.", Xe_c, Y_cntxt).mean(dim=1, keepdim=True)
    R_trgt = R.expand(B, X_trgt.shape[1], r_dim).unsqueeze(0)
    return _decode(sd, Xe_t, R_trgt, y_dim)


def attncnp_forward(sd, X_cntxt, Y_cntxt, X_trgt, attention="transformer", n_heads=8, is_self_attn=False):
    """AttnCNP, npf/neuralproc/attnnp.py:105-131.  attention in {"scaledot","multihead","transformer"}."""
    y_dim = Y_cntxt.shape[-1]
    Xe_c = mlp(sd, "x_encoder.", X_cntxt)
    Xe_t = mlp(sd, "x_encoder.", X_trgt)
    B, C, _ = Xe_c.shape
    r_dim = sd["decoder.resizer.to_hidden.weight"].shape[1]  # MergeFlatInputs resizes R (r_dim) onto x
    if C == 0:
        R_trgt = torch.zeros(B, X_trgt.shape[1], r_dim, dtype=Xe_t.dtype)
    else:
        R_c = _attn_xy_encode(sd, Xe_c, Y_cntxt, is_self_attn)
        R_trgt = _attend(sd, Xe_c, Xe_t, R_c, attention, n_heads)
    return _decode(sd, Xe_t, R_trgt.unsqueeze(0), y_dim)


def _convcnp_encode(sd, X_cntxt, Y_cntxt, X_induced, training):
    """ConvCNP.encode_globally, npf/neuralproc/convnp.py:137-156."""
    B, C, _ = X_cntxt.shape
    Xi = X_induced.view(1, -1, 1).expand(B, -1, 1)
    R = setconv(sd, "cntxt_to_induced.", X_cntxt, Xi, Y_cntxt)
    if C == 0:
        R = torch.zeros_like(R)
    return res_conv_cnn(sd, "induced_to_induced.", R, training), Xi


def convcnp_forward(sd, X_cntxt, Y_cntxt, X_trgt, X_induced=None, training=True):
    """ConvCNP (off-grid, 1-D), npf/neuralproc/convnp.py:128-168: SetConv -> CNN -> SetConv -> MLP."""
    y_dim = Y_cntxt.shape[-1]
    if X_induced is None:
        X_induced = induced_grid(128).to(X_cntxt.dtype)
    R_ind, Xi = _convcnp_encode(sd, X_cntxt, Y_cntxt, X_induced, training)
    R_trgt = setconv(sd, "induced_to_trgt.", Xi, X_trgt, R_ind)
    return _decode(sd, X_trgt, R_trgt.unsqueeze(0), y_dim)


def _grid_cntxt_to_induced(sd, mask_cntxt, Y, circular=False):
    """GridConvCNP.cntxt_to_induced, npf/neuralproc/gridconvnp.py:136-162; abs-weight depthwise conv
    of npf/utils/helpers.py:316-331 (no bias, one filter per y channel)."""
    X = Y.permute(0, 3, 1, 2)
    m = mask_cntxt.permute(0, 3, 1, 2).to(Y.dtype)
    w = sd["conv.weight"].abs()
    y_dim = X.shape[1]
    pad = w.shape[-1] // 2
    if circular:  # the padded first layer of `model_2d_extrap` (helpers.py:334-351): wrap-around, then an unpadded conv
        wrap = lambda t: F.pad(t, (pad,) * 4, mode="circular")
        signal = F.conv2d(wrap(X * m), w, None, padding=0, groups=y_dim)
        density = F.conv2d(wrap(m.expand_as(X)), w, None, padding=0, groups=y_dim)
    else:
        signal = F.conv2d(X * m, w, None, padding=pad, groups=y_dim)
        density = F.conv2d(m.expand_as(X), w, None, padding=pad, groups=y_dim)
    out = signal / torch.clamp(density, min=1e-5)
    out = torch.cat([out, density], dim=1).permute(0, 2, 3, 1)
    return _lin(sd, "resizer.", out)


def gridconvcnp_forward(sd, mask_cntxt, Y, mask_trgt=None, training=True, circular=False, bn_eps=1e-5):
    """GridConvCNP, npf/neuralproc/gridconvnp.py:136-175.  mask_cntxt bool [B,H,W,1], Y [B,H,W,y].  ``circular``: every
    convolution wraps around (`model_2d_extrap` of the notebooks)."""
    y_dim = Y.shape[-1]
    R = res_conv_cnn(sd, "induced_to_induced.", _grid_cntxt_to_induced(sd, mask_cntxt, Y, circular), training, circular, bn_eps)
    return _decode(sd, None, R.unsqueeze(0), y_dim)


def _latent_dist(sd, R, z_dim):
    """LatentNeuralProcessFamily.infer_latent_dist, npf/neuralproc/base.py:516-547."""
    suff = mlp(sd, "latent_encoder.", R)
    loc, s = suff.split(z_dim, dim=-1)
    return loc, q_z_scale(s)


def _add_global_latent(z):
    """ConvLNP.add_global_latent, npf/neuralproc/convnp.py:322-335 + helpers.py:20-32: second half
    of the channels is mean-pooled over all middle (spatial) dims and broadcast back."""
    half = z.shape[-1] // 2
    loc_z, glob_z = z.split(half, dim=-1)
    first, *middle, last = glob_z.shape
    g = glob_z.reshape(first, -1, last).mean(1, keepdim=True)
    g = g.view(first, *([1] * len(middle)), last).expand(first, *middle, last)
    return torch.cat([loc_z, g], dim=-1)


def lnp_forward(sd, X_cntxt, Y_cntxt, X_trgt, eps, encoded_path="latent"):
    """LNP, npf/neuralproc/np.py:113-163 with base.py:495-514, 554-575.  eps [n_z,B,1,z] ~ N(0,1);
    z = loc + scale*eps is what Normal.rsample does (q(z|cntxt) sampling, is_q_zCct=False)."""
    y_dim = Y_cntxt.shape[-1]
    Xe_c = mlp(sd, "x_encoder.", X_cntxt)
    Xe_t = mlp(sd, "x_encoder.", X_trgt)
    B, C, _ = Xe_c.shape
    r_dim = sd["xy_encoder.flat_module.out.weight"].shape[0]
    if C == 0:
        R = torch.zeros(B, 1, r_dim, dtype=Xe_t.dtype)
    else:
        R = merge_flat_sum(sd, "xy_encoder.", Xe_c, Y_cntxt).mean(dim=1, keepdim=True)
    z_dim = sd["latent_encoder.out.weight"].shape[0] // 2
    q_loc, q_scale = _latent_dist(sd, R, z_dim)
    z = q_loc.unsqueeze(0) + q_scale.unsqueeze(0) * eps
    if encoded_path == "both":
        Rz = R.unsqueeze(0).expand(*z.shape[:-1], r_dim)
        R_trgt = torch.relu(_lin(sd, "r_z_merger.", torch.cat((Rz, z), dim=-1)))
    else:
        R_trgt = z
        if "reshaper_z.weight" in sd:
            R_trgt = _lin(sd, "reshaper_z.", R_trgt)
    R_trgt = R_trgt.expand(z.shape[0], B, X_trgt.shape[1], r_dim)
    loc, scale = _decode(sd, Xe_t, R_trgt, y_dim)
    return loc, scale, z, q_loc, q_scale


def attnlnp_forward(sd, X_cntxt, Y_cntxt, X_trgt, eps, attention="transformer", n_heads=8, is_self_attn=False,
                    Y_trgt=None):
    """AttnLNP (encoded_path="both"), npf/neuralproc/attnnp.py:134-202 with base.py:495-514, 554-575:
    deterministic path = AttnCNP's cross-attention; latent path = one global z per task inferred from
    the MEAN of the per-context representations (zeros without context).  With ``Y_trgt`` given, z is
    sampled from q(z | targets) (is_q_zCct=True in training, base.py:501-506): the target set goes
    through the same xy-encoder.  Returns loc, scale, z, q_c (loc, scale), q_ct (loc, scale) or None."""
    y_dim = Y_cntxt.shape[-1]
    Xe_c = mlp(sd, "x_encoder.", X_cntxt)
    Xe_t = mlp(sd, "x_encoder.", X_trgt)
    B, C, _ = Xe_c.shape
    r_dim = sd["decoder.resizer.to_hidden.weight"].shape[1]  # MergeFlatInputs resizes R (r_dim) onto x
    z_dim = sd["latent_encoder.out.weight"].shape[0] // 2

    def pooled(R):
        if R.shape[1] == 0:
            return torch.zeros(R.shape[0], 1, r_dim, dtype=Xe_t.dtype)
        return R.mean(dim=1, keepdim=True)

    R_c = (_attn_xy_encode(sd, Xe_c, Y_cntxt, is_self_attn) if C > 0
           else torch.zeros(B, 0, r_dim, dtype=Xe_t.dtype))
    q_loc, q_scale = _latent_dist(sd, pooled(R_c), z_dim)
    q_ct = None
    if Y_trgt is not None:
        q_ct = _latent_dist(sd, pooled(_attn_xy_encode(sd, Xe_t, Y_trgt, is_self_attn)), z_dim)
        z = q_ct[0].unsqueeze(0) + q_ct[1].unsqueeze(0) * eps
    else:
        z = q_loc.unsqueeze(0) + q_scale.unsqueeze(0) * eps
    T = X_trgt.shape[1]
    if C == 0:
        R_det = torch.zeros(B, T, r_dim, dtype=Xe_t.dtype)
    else:
        R_det = _attend(sd, Xe_c, Xe_t, R_c, attention, n_heads)
    n_z = z.shape[0]
    Rz = R_det.unsqueeze(0).expand(n_z, B, T, r_dim)
    zz = z.expand(n_z, B, T, z_dim)
    R_trgt = torch.relu(_lin(sd, "r_z_merger.", torch.cat((Rz, zz), dim=-1)))
    loc, scale = _decode(sd, Xe_t, R_trgt, y_dim)
    return loc, scale, z, (q_loc, q_scale), q_ct


def convlnp_forward(sd, X_cntxt, Y_cntxt, X_trgt, eps, X_induced=None, is_global=False, training=True):
    """ConvLNP (encoded_path="latent"), npf/neuralproc/convnp.py:253-320: latent per induced point,
    CNN after sampling on the (n_z*B) collapsed batch, (global pooling AFTER the CNN), SetConv to targets."""
    y_dim = Y_cntxt.shape[-1]
    if X_induced is None:
        X_induced = induced_grid(128).to(X_cntxt.dtype)
    R_ind, Xi = _convcnp_encode(sd, X_cntxt, Y_cntxt, X_induced, training)
    z_dim = sd["latent_encoder.out.weight"].shape[0] // 2
    q_loc, q_scale = _latent_dist(sd, R_ind, z_dim)
    z = q_loc.unsqueeze(0) + q_scale.unsqueeze(0) * eps  # [n_z,B,I,z]
    n_z, B = z.shape[:2]
    zc = z.reshape(n_z * B, *z.shape[2:])
    if "reshaper_z.weight" in sd:
        zc = _lin(sd, "reshaper_z.", zc)
    zc = res_conv_cnn(sd, "induced_to_induced_post_sampling.", zc, training)
    if is_global:
        zc = _add_global_latent(zc)
    Xi_r = Xi.unsqueeze(0).expand(n_z, *Xi.shape).reshape(n_z * B, *Xi.shape[1:])
    Xt_r = X_trgt.unsqueeze(0).expand(n_z, *X_trgt.shape).reshape(n_z * B, *X_trgt.shape[1:])
    R_trgt = setconv(sd, "induced_to_trgt.", Xi_r, Xt_r, zc).view(n_z, B, X_trgt.shape[1], -1)
    loc, scale = _decode(sd, None, R_trgt, y_dim)
    return loc, scale, z, q_loc, q_scale


def gridconvlnp_forward(sd, mask_cntxt, Y, eps, is_global=False, training=True, circular=False, bn_eps=1e-5):
    """GridConvLNP (encoded_path="latent"), npf/neuralproc/gridconvnp.py:244-289: latent per pixel,
    (global pooling BEFORE the post-sampling CNN), CNN on the (n_z*B) collapsed batch."""
    y_dim = Y.shape[-1]
    R = res_conv_cnn(sd, "induced_to_induced.", _grid_cntxt_to_induced(sd, mask_cntxt, Y, circular), training, circular, bn_eps)
    z_dim = sd["latent_encoder.out.weight"].shape[0] // 2
    q_loc, q_scale = _latent_dist(sd, R, z_dim)
    z = q_loc.unsqueeze(0) + q_scale.unsqueeze(0) * eps  # [n_z,B,H,W,z]
    n_z, B = z.shape[:2]
    zc = z.reshape(n_z * B, *z.shape[2:])
    if is_global:
        zc = _add_global_latent(zc)
    if "reshaper_z.weight" in sd:
        zc = _lin(sd, "reshaper_z.", zc)
    R_trgt = res_conv_cnn(sd, "induced_to_induced_post_sampling.", zc, training, circular, bn_eps)
    R_trgt = R_trgt.view(n_z, B, *R_trgt.shape[1:])
    loc, scale = _decode(sd, None, R_trgt, y_dim)
    return loc, scale, z, q_loc, q_scale


# --------------------------------------------------------------------------------------------
# losses (npf/losses.py)
# --------------------------------------------------------------------------------------------
def gauss_sum_log_prob(loc, scale, Y):
    """sum_log_prob, npf/losses.py:18-24 with Independent(Normal): sum over targets and y of
    -log(scale) - 0.5 log(2 pi) - 0.5 ((Y - loc)/scale)^2  ->  [n_z, B]."""
    lp = -((Y - loc) ** 2) / (2 * scale ** 2) - scale.log() - math.log(math.sqrt(2 * math.pi))
    return lp.reshape(*lp.shape[:2], -1).sum(-1)


def cnpf_loss(loc, scale, Y, reduction="mean"):
    """CNPFLoss, npf/losses.py:112-123 (+ reduction 71-79)."""
    nll = -gauss_sum_log_prob(loc, scale, Y).squeeze(0)
    return _reduce(nll, reduction)


def nll_lnpf_loss(loc, scale, Y, reduction="mean"):
    """NLLLossLNPF without importance weights, npf/losses.py:169-203: -(logsumexp_z - log n_z)."""
    s = gauss_sum_log_prob(loc, scale, Y)
    nll = -(torch.logsumexp(s, 0) - math.log(s.shape[0]))
    return _reduce(nll, reduction)


def elbo_lnpf_loss(loc, scale, Y, q_ct_loc, q_ct_scale, q_c_loc, q_c_scale, reduction="mean"):
    """ELBOLossLNPF, npf/losses.py:135-150: -(mean_z sum log p - sum KL(q_ct || q_c)); KL between
    diagonal Gaussians summed over latent points and z dims."""
    s = gauss_sum_log_prob(loc, scale, Y).mean(0)
    var_ratio = (q_ct_scale / q_c_scale) ** 2
    t1 = ((q_ct_loc - q_c_loc) / q_c_scale) ** 2
    kl = 0.5 * (var_ratio + t1 - 1 - var_ratio.log())
    kl = kl.reshape(kl.shape[0], -1).sum(-1)
    return _reduce(-(s - kl), reduction)


def _reduce(loss, reduction):
    if reduction is None:
        return loss
    if reduction == "mean":
        return loss.mean(0)
    if reduction == "sum":
        return loss.sum(0)
    raise ValueError(f"Unknown {reduction}")
