"""Generate tests/golden/*.pt from the REAL reference (imported read-only from /root/reference).

Run in the build container only:   python oracle/gen_golden.py
The reference cannot travel to the GPU box, so the vectors it produces are committed as fixtures;
this script is committed so they can be regenerated / audited.  Nothing at test / bench / smoke time
reads /root/reference.

Each fixture ``<name>.pt`` is a dict:
    cfg         : dict describing how the model was built (consumed by tests/_cfg.py::build_model)
    state_dict  : the reference module's state_dict (random init under a seed, or upstream pretrained)
    cases       : list of dicts {name, training, inputs{...}, eps (latent only), loc, scale, loss_per_task,
                                 loss, q_loc, q_scale, grad_proj{param_name: [9] fp64 = 8 random projections + L2 norm}
                                 (train cases only; z = q_loc + q_scale * eps is not stored)}
All tensors fp32 (masks bool).
"""
import os
import random
import sys
from functools import partial

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
sys.path.insert(0, REF)
import npf  # noqa: E402
from npf import (  # noqa: E402
    CNP, LNP, AttnCNP, AttnLNP, ConvCNP, ConvLNP, GridConvCNP, GridConvLNP, CNPFLoss, ELBOLossLNPF, NLLLossLNPF,
)
from npf.architectures import CNN, MLP, ResConvBlock, SetConv, discard_ith_arg, merge_flat_input  # noqa: E402
import torch.distributions.normal as _tdn  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
PRE = os.path.join(REF, "results", "pretrained")

torch.set_num_threads(8)


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


# ---- record the eps drawn by Normal.rsample so the latent models can be replayed exactly ----------
_EPS_LOG = []
_orig_std_normal = _tdn._standard_normal


_EPS_SEED = [None]   # when set: eps = randn(shape, Generator(seed)) so that a large eps need not be stored in the fixture


def _recording_standard_normal(shape, dtype, device):
    if _EPS_SEED[0] is not None:
        e = torch.randn(tuple(shape), generator=torch.Generator().manual_seed(_EPS_SEED[0]), dtype=dtype)
    else:
        e = _orig_std_normal(shape, dtype, device)
    _EPS_LOG.append(e.detach().clone())
    return e


_tdn._standard_normal = _recording_standard_normal

R_DIM = 128


# ---- the configurations (cfg dict is the portable description; build() turns it into the reference) ----
def build(cfg):
    fam = cfg["family"]
    kw = {}
    if cfg.get("notebook"):
        if fam in ("CNP", "AttnCNP", "LNP"):
            kw["XEncoder"] = partial(MLP, n_hidden_layers=1, hidden_size=R_DIM)
            kw["Decoder"] = merge_flat_input(partial(MLP, n_hidden_layers=4, hidden_size=R_DIM), is_sum_merge=True)
        if fam in ("CNP", "AttnCNP", "LNP", "AttnLNP"):
            kw["r_dim"] = R_DIM
            if not cfg.get("is_self_attn"):  # the 2-D notebooks use the self-attention encoder instead
                kw["XYEncoder"] = merge_flat_input(
                    partial(MLP, n_hidden_layers=2, hidden_size=cfg["xy_hidden"]), is_sum_merge=True)
        elif fam in ("ConvCNP", "GridConvCNP"):
            kw["r_dim"] = R_DIM
            kw["Decoder"] = discard_ith_arg(partial(MLP, n_hidden_layers=4, hidden_size=R_DIM), i=0)
        elif fam in ("ConvLNP", "GridConvLNP"):
            kw["r_dim"] = R_DIM
            kw["Decoder"] = discard_ith_arg(torch.nn.Linear, i=0)
            kw["is_q_zCct"] = False
    if "cnn" in cfg:
        c = cfg["cnn"]
        Conv = nn.Conv1d if c["dim"] == 1 else nn.Conv2d
        Norm = {None: nn.Identity, "bn": nn.BatchNorm1d if c["dim"] == 1 else nn.BatchNorm2d}[c.get("norm")]
        if "bn_eps" in c:
            Norm = partial(Norm, eps=c["bn_eps"])
        if cfg.get("circular"):  # `model_2d_extrap` of ConvCNP.ipynb / ConvLNP.ipynb: wrap-around padding everywhere
            from npf.utils.helpers import CircularPad2d, make_abs_conv, make_padded_conv
            Conv = make_padded_conv(Conv, CircularPad2d)
            kw["Conv"] = lambda y_dim: make_padded_conv(make_abs_conv(nn.Conv2d), CircularPad2d)(
                y_dim, y_dim, groups=y_dim, kernel_size=11, padding=11 // 2, bias=False)
        kw["CNN"] = partial(CNN, ConvBlock=ResConvBlock, Conv=Conv, Normalization=Norm, n_blocks=c["n_blocks"],
                            kernel_size=c["kernel_size"], is_chan_last=True, n_conv_layers=c["n_conv_layers"])
    for k in ("density_induced", "attention", "n_z_samples_train", "n_z_samples_test", "is_global", "encoded_path",
              "is_q_zCct", "is_self_attn"):
        if k in cfg and not (k == "encoded_path" and fam == "AttnLNP"):
            kw[k] = cfg[k]
    if fam in ("ConvCNP", "ConvLNP") and cfg.get("notebook"):
        kw["Interpolator"] = SetConv
    cls = dict(CNP=CNP, LNP=LNP, AttnCNP=AttnCNP, AttnLNP=AttnLNP, ConvCNP=ConvCNP, ConvLNP=ConvLNP, GridConvCNP=GridConvCNP,
               GridConvLNP=GridConvLNP)[fam]
    seed_all(cfg.get("init_seed", 0))
    model = cls(cfg["x_dim"], cfg["y_dim"], **kw)
    if cfg.get("pretrained"):
        sd = torch.load(os.path.join(PRE, cfg["pretrained"], "run_0", "params.pt"), map_location="cpu")
        print("   load pretrained:", model.load_state_dict(sd))
    return model


def gp_like(B, N, y_dim, seed):
    """Smooth-ish synthetic functions on sorted x in [-1,1] (values do not matter for parity)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, 1, generator=g) * 2 - 1
    ph = torch.rand(B, 1, y_dim, generator=g) * 6.28
    fr = 2 + 5 * torch.rand(B, 1, y_dim, generator=g)
    y = torch.sin(fr * x + ph) + 0.1 * torch.randn(B, N, y_dim, generator=g)
    return x, y


def offgrid_inputs(B, C, T, y_dim, seed, x_scale=1.0, dup=False):
    x, y = gp_like(B, C + T, y_dim, seed)
    x = x * x_scale
    Xc, Yc, Xt, Yt = x[:, :C], y[:, :C], x[:, C:], y[:, C:]
    if dup and C >= 2:
        Xc = Xc.clone(); Xc[:, 1] = Xc[:, 0]
        Xt = Xt.clone(); Xt[:, 0] = Xc[:, 0]
    return dict(X_cntxt=Xc.contiguous(), Y_cntxt=Yc.contiguous(), X_trgt=Xt.contiguous(), Y_trgt=Yt.contiguous())


def grid_inputs(B, H, W, y_dim, frac, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, H, W, y_dim, generator=g)
    n = int(frac * H * W)
    mask = torch.zeros(B, H * W, dtype=torch.bool)
    for b in range(B):
        mask[b, torch.randperm(H * W, generator=g)[:n]] = True
    mask = mask.view(B, H, W, 1)
    return dict(X_cntxt=mask, Y_cntxt=img, X_trgt=torch.ones(B, H, W, 1, dtype=torch.bool), Y_trgt=img.clone())


def blob_inputs(B, H, W, frac, seed):
    """Digit-like single-channel images (a few soft strokes on a black background, values in [0, 1]) for the checkpoints trained on
    MNIST-type data: uniform noise drives those models to sigma = 0.01 everywhere and NLLs of 1e5-1e6, whose gradients are sums of huge
    cancelling terms (nothing a fp32 implementation can be pinned on to 1e-3)."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    img = torch.zeros(B, H, W)
    for b in range(B):
        for _ in range(4):
            cy, cx = torch.rand(2, generator=g) * torch.tensor([H - 1.0, W - 1.0])
            s = 1.2 + 1.8 * torch.rand(1, generator=g)
            img[b] += torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    img = img.clamp(0, 1).unsqueeze(-1)
    out = grid_inputs(B, H, W, 1, frac, seed + 1)
    out["Y_cntxt"], out["Y_trgt"] = img, img.clone()
    return out


N_PROJ = 8


def grad_projection(g):
    """Compact pin of a gradient tensor: 8 fixed random projections <g, v_i> + its L2 norm (fp64 accumulate).
    v_i = randn(shape) under Generator seed 1000+i -- reproduced by tests/_util.py::grad_projection."""
    out = []
    g64 = g.double().reshape(-1)
    for i in range(N_PROJ):
        v = torch.randn(g64.numel(), generator=torch.Generator().manual_seed(1000 + i), dtype=torch.float64)
        out.append(torch.dot(g64, v))
    out.append(g64.norm())
    return torch.stack(out)


def run_case(model, cfg, name, training, inputs, loss_name, with_grads, extrap=None, seed=0, eps_seed=None):
    _EPS_SEED[0] = eps_seed
    model.train(training)
    if extrap is not None:
        model.set_extrapolation(extrap)
    crit = dict(cnpf=CNPFLoss, nll=NLLLossLNPF, elbo=ELBOLossLNPF)[loss_name](reduction=None)
    crit.train(training)
    model.zero_grad()
    _EPS_LOG.clear()
    seed_all(seed)
    bn_before = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k or "num_batches" in k}
    out = model(inputs["X_cntxt"], inputs["Y_cntxt"], inputs["X_trgt"], inputs["Y_trgt"])
    p, z, q_c, q_ct = out
    per_task = crit(out, inputs["Y_trgt"])
    loss = per_task.mean(0)
    case = dict(name=name, training=training, inputs=inputs, loss_name=loss_name,
                loc=p.base_dist.loc.detach().clone(), scale=p.base_dist.scale.detach().clone(),
                loss_per_task=per_task.detach().clone(), loss=loss.detach().clone())
    if extrap is not None:
        case["extrap"] = list(extrap)
        case["X_induced"] = model.X_induced.detach().clone()
    if z is not None:
        case["q_loc"] = q_c.base_dist.loc.detach().clone()
        case["q_scale"] = q_c.base_dist.scale.detach().clone()
        assert len(_EPS_LOG) == 1
        if eps_seed is None:
            case["eps"] = _EPS_LOG[0]
        else:   # tests/_util.py::load_fixture regenerates it and verifies the checksum
            e = _EPS_LOG[0]
            case["eps_seed"], case["eps_shape"] = eps_seed, tuple(e.shape)
            case["eps_check"] = torch.stack([e.double().sum(), e.double().abs().sum(), e.reshape(-1)[:: max(1, e.numel() // 7)].double().sum()])
        if q_ct is not None:
            case["q_ct_loc"] = q_ct.base_dist.loc.detach().clone()
            case["q_ct_scale"] = q_ct.base_dist.scale.detach().clone()
    if with_grads:
        loss.backward()
        case["grad_proj"] = {k: grad_projection(p_.grad.detach()) for k, p_ in model.named_parameters()
                             if p_.grad is not None}
    if training and bn_before:
        # restore BN running stats so every case starts from the fixture's state_dict; keep the updated ones
        case["bn_after"] = {k: v.detach().clone() for k, v in model.state_dict().items() if k in bn_before}
        model.load_state_dict({**model.state_dict(), **bn_before})
    if extrap is not None:
        model.set_extrapolation((-1, 1))
    print(f"   case {name:28s} train={training} loss={loss.item():.6f}")
    return case


def dump(name, cfg, model, cases):
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    n_params = sum(p.numel() for p in model.parameters())
    torch.save(dict(cfg=cfg, state_dict=sd, cases=cases, n_params=n_params), os.path.join(OUT, name + ".pt"))
    print(f"== {name}: {n_params} params, {len(cases)} cases")


def main():
    os.makedirs(OUT, exist_ok=True)

    # ---------------- CNP ----------------
    cfg = dict(family="CNP", x_dim=1, y_dim=1)
    m = build(cfg)
    dump("cnp_default", cfg, m, [
        run_case(m, cfg, "train_b16_c32_t64", True, offgrid_inputs(16, 32, 64, 1, 1), "cnpf", True),
        run_case(m, cfg, "eval_c0", False, offgrid_inputs(3, 0, 9, 1, 2), "cnpf", False),
        run_case(m, cfg, "train_c1_t1", True, offgrid_inputs(2, 1, 1, 1, 3), "cnpf", True),
    ])
    cfg = dict(family="CNP", x_dim=1, y_dim=1, notebook=True, xy_hidden=256, pretrained="RBF_Kernel/CNP")
    m = build(cfg)
    dump("cnp_notebook_pretrained", cfg, m, [
        run_case(m, cfg, "train_b4_c20_t30", True, offgrid_inputs(4, 20, 30, 1, 4), "cnpf", True),
        run_case(m, cfg, "eval_b4", False, offgrid_inputs(4, 10, 40, 1, 5), "cnpf", False),
    ])
    cfg = dict(family="CNP", x_dim=2, y_dim=3, init_seed=3)
    m = build(cfg)
    dump("cnp_x2_y3", cfg, m, [
        run_case(m, cfg, "train_b3_c12_t17", True, _xy2(3, 12, 17, 3, 6), "cnpf", True),
    ])

    # ---------------- AttnCNP ----------------
    cfg = dict(family="AttnCNP", x_dim=1, y_dim=1)  # scaledot default
    m = build(cfg)
    dump("attncnp_scaledot", cfg, m, [
        run_case(m, cfg, "train_b4_c19_t23", True, offgrid_inputs(4, 19, 23, 1, 7), "cnpf", True),
        run_case(m, cfg, "eval_c0", False, offgrid_inputs(2, 0, 5, 1, 8), "cnpf", False),
    ])
    cfg = dict(family="AttnCNP", x_dim=1, y_dim=1, notebook=True, xy_hidden=128, attention="transformer",
               pretrained="RBF_Kernel/AttnCNP")
    m = build(cfg)
    dump("attncnp_transformer_pretrained", cfg, m, [
        run_case(m, cfg, "train_b4_c33_t70", True, offgrid_inputs(4, 33, 70, 1, 9), "cnpf", True),
        run_case(m, cfg, "eval_b2_c130_t150", False, offgrid_inputs(2, 130, 150, 1, 10), "cnpf", False),
        run_case(m, cfg, "train_c1_t1", True, offgrid_inputs(2, 1, 1, 1, 11), "cnpf", True),
    ])
    cfg = dict(family="AttnCNP", x_dim=1, y_dim=2, attention="multihead", init_seed=5)
    m = build(cfg)
    dump("attncnp_multihead_y2", cfg, m, [
        run_case(m, cfg, "train_b3_c16_t40", True, offgrid_inputs(3, 16, 40, 2, 12), "cnpf", True),
    ])

    # ---------------- ConvCNP (off-grid 1-D) ----------------
    cfg = dict(family="ConvCNP", x_dim=1, y_dim=1)
    m = build(cfg)
    dump("convcnp_default", cfg, m, [
        run_case(m, cfg, "train_b3_c9_t13", True, offgrid_inputs(3, 9, 13, 1, 13), "cnpf", True),
        run_case(m, cfg, "train_b2_c128_t128", True, offgrid_inputs(2, 128, 128, 1, 14), "cnpf", True),
        run_case(m, cfg, "eval_c0", False, offgrid_inputs(2, 0, 7, 1, 15), "cnpf", False),
        run_case(m, cfg, "train_c1_t1", True, offgrid_inputs(2, 1, 1, 1, 16), "cnpf", True),
        run_case(m, cfg, "train_dup_x", True, offgrid_inputs(2, 6, 5, 1, 17, dup=True), "cnpf", True),
        run_case(m, cfg, "eval_extrap", False, offgrid_inputs(2, 10, 12, 1, 18, x_scale=2.0), "cnpf", False,
                 extrap=(-2, 2)),
    ])
    cfg = dict(family="ConvCNP", x_dim=1, y_dim=2, init_seed=7)
    m = build(cfg)
    dump("convcnp_default_y2", cfg, m, [
        run_case(m, cfg, "train_b2_c11_t9", True, offgrid_inputs(2, 11, 9, 2, 19), "cnpf", True),
    ])
    cfg = dict(family="ConvCNP", x_dim=1, y_dim=1, notebook=True, density_induced=64, pretrained="RBF_Kernel/ConvCNP",
               cnn=dict(dim=1, norm="bn", n_blocks=5, kernel_size=19, n_conv_layers=2))
    m = build(cfg)
    dump("convcnp_notebook_pretrained", cfg, m, [
        run_case(m, cfg, "eval_b4_c30_t50", False, offgrid_inputs(4, 30, 50, 1, 20), "cnpf", False),
        run_case(m, cfg, "train_b4_c25_t40", True, offgrid_inputs(4, 25, 40, 1, 21), "cnpf", True),
    ])

    # ---------------- GridConvCNP (on-grid 2-D) ----------------
    for y in (1, 3):
        cfg = dict(family="GridConvCNP", x_dim=1, y_dim=y, init_seed=y)
        m = build(cfg)
        dump(f"gridconvcnp_default_y{y}", cfg, m, [
            run_case(m, cfg, "train_b2_32x32", True, grid_inputs(2, 32, 32, y, 0.3, 22 + y), "cnpf", True),
            run_case(m, cfg, "eval_b1_20x28_sparse", False, grid_inputs(1, 20, 28, y, 0.02, 30 + y), "cnpf", False),
        ])
    cfg = dict(family="GridConvCNP", x_dim=1, y_dim=3, notebook=True, pretrained="celeba32/ConvCNP",
               cnn=dict(dim=2, norm="bn", n_blocks=5, kernel_size=9, n_conv_layers=2))
    m = build(cfg)
    dump("gridconvcnp_notebook_pretrained", cfg, m, [
        run_case(m, cfg, "eval_b2_32x32", False, grid_inputs(2, 32, 32, 3, 0.2, 40), "cnpf", False),
        run_case(m, cfg, "train_b2_32x32", True, grid_inputs(2, 32, 32, 3, 0.3, 41), "cnpf", True),
    ])

    # ---------------- latent models ----------------
    cfg = dict(family="GridConvLNP", x_dim=1, y_dim=3, n_z_samples_train=4, n_z_samples_test=3, init_seed=9)
    m = build(cfg)
    dump("gridconvlnp_default_y3", cfg, m, [
        run_case(m, cfg, "train_b1_16x16_nz4", True, grid_inputs(1, 16, 16, 3, 0.3, 42), "nll", True),
        run_case(m, cfg, "eval_b1_12x20_nz3", False, grid_inputs(1, 12, 20, 3, 0.3, 43), "nll", False),
    ])
    cfg = dict(family="GridConvLNP", x_dim=1, y_dim=3, notebook=True, n_z_samples_train=16, n_z_samples_test=32,
               is_global=True, pretrained="celeba32/ConvLNP",
               cnn=dict(dim=2, norm="bn", n_blocks=4, kernel_size=9, n_conv_layers=2))
    m = build(cfg)
    m.n_z_samples_test = 3
    cfg["n_z_samples_test"] = 3
    dump("gridconvlnp_notebook_pretrained", cfg, m, [
        run_case(m, cfg, "eval_b1_24x24_nz3", False, grid_inputs(1, 24, 24, 3, 0.25, 44), "nll", False),
    ])
    cfg = dict(family="ConvLNP", x_dim=1, y_dim=1, n_z_samples_train=3, n_z_samples_test=2, init_seed=11)
    m = build(cfg)
    dump("convlnp_default", cfg, m, [
        run_case(m, cfg, "train_b2_c10_t14_nz3", True, offgrid_inputs(2, 10, 14, 1, 45), "nll", True),
        run_case(m, cfg, "eval_b1_nz2", False, offgrid_inputs(1, 6, 9, 1, 46), "nll", False),
    ])
    cfg = dict(family="ConvLNP", x_dim=1, y_dim=1, notebook=True, density_induced=64, n_z_samples_train=16,
               n_z_samples_test=32, is_global=True, pretrained="RBF_Kernel/ConvLNP",
               cnn=dict(dim=1, norm="bn", n_blocks=4, kernel_size=19, n_conv_layers=2))
    m = build(cfg)
    m.n_z_samples_test = 4
    cfg["n_z_samples_test"] = 4
    dump("convlnp_notebook_pretrained", cfg, m, [
        run_case(m, cfg, "eval_b2_c20_t30_nz4", False, offgrid_inputs(2, 20, 30, 1, 47), "nll", False),
    ])
    cfg = dict(family="LNP", x_dim=1, y_dim=1, notebook=True, xy_hidden=128, n_z_samples_train=5, n_z_samples_test=4,
               init_seed=13)  # NB: bare LNP(1,1) raises KeyError upstream (no default XYEncoder), so pass one
    m = build(cfg)
    dump("lnp_default", cfg, m, [
        run_case(m, cfg, "train_b3_c8_t12_nz5", True, offgrid_inputs(3, 8, 12, 1, 48), "nll", True),
    ])
    cfg = dict(family="LNP", x_dim=1, y_dim=1, notebook=True, xy_hidden=128, n_z_samples_train=3, n_z_samples_test=4, init_seed=14,
               is_q_zCct=True, encoded_path="both")
    m = build(cfg)
    dump("lnp_both_elbo", cfg, m, [
        run_case(m, cfg, "train_b3_c8_t12_nz3_elbo", True, offgrid_inputs(3, 8, 12, 1, 49), "elbo", True),
    ])


def main_attn_family():
    """AttnLNP and the self-attention (2-D notebook) encoders -- separate entry so the older fixtures
    need not be regenerated:  python oracle/gen_golden.py attn"""
    os.makedirs(OUT, exist_ok=True)
    # AttnCNP.ipynb model_2d: x = pixel coordinates, SelfAttention xy-encoder, upstream celeba32 weights
    cfg = dict(family="AttnCNP", x_dim=2, y_dim=3, notebook=True, attention="transformer", is_self_attn=True,
               pretrained="celeba32/AttnCNP")
    m = build(cfg)
    dump("attncnp_selfattn_pretrained", cfg, m, [
        run_case(m, cfg, "train_b3_c40_t70", True, _xy2(3, 40, 70, 3, 50, unit=True), "cnpf", True),
        # seed 63: the reference's own fp32 result sits 1.7e-5 (sigma) from its fp64 re-run on these inputs; for e.g.
        # seed 51 it is 7e-5 (pixel-trained sharp attention on synthetic inputs), which would leave no room under 1e-4
        run_case(m, cfg, "eval_b2_c150_t200", False, _xy2(2, 150, 200, 3, 63, unit=True), "cnpf", False),
        run_case(m, cfg, "eval_c0", False, _xy2(2, 0, 6, 3, 52, unit=True), "cnpf", False),
        run_case(m, cfg, "train_c1_t1", True, _xy2(2, 1, 1, 3, 53, unit=True), "cnpf", True),
    ])
    cfg = dict(family="AttnCNP", x_dim=1, y_dim=1, attention="multihead", is_self_attn=True, init_seed=15)
    m = build(cfg)
    dump("attncnp_selfattn_multihead", cfg, m, [
        run_case(m, cfg, "train_b3_c17_t21", True, offgrid_inputs(3, 17, 21, 1, 54), "cnpf", True),
    ])
    # AttnLNP.ipynb model_1d: NPVI (ELBO, z ~ q(z | targets)), n_z_samples_train=1, upstream RBF weights
    cfg = dict(family="AttnLNP", x_dim=1, y_dim=1, notebook=True, xy_hidden=128, attention="transformer",
               is_q_zCct=True, n_z_samples_train=1, n_z_samples_test=8, encoded_path="both",
               pretrained="RBF_Kernel/AttnLNP")
    m = build(cfg)
    m.n_z_samples_test = 3
    cfg["n_z_samples_test"] = 3
    dump("attnlnp_pretrained", cfg, m, [
        run_case(m, cfg, "train_b4_c25_t45_elbo", True, offgrid_inputs(4, 25, 45, 1, 55), "elbo", True),
        run_case(m, cfg, "eval_b2_c60_t80_nz3", False, offgrid_inputs(2, 60, 80, 1, 56), "nll", False),
        run_case(m, cfg, "eval_c0_nz3", False, offgrid_inputs(2, 0, 5, 1, 57), "nll", False),
    ])
    # AttnLNP.ipynb model_2d: self-attention encoder, NLL with several samples, upstream mnist weights
    cfg = dict(family="AttnLNP", x_dim=2, y_dim=1, notebook=True, attention="transformer", is_self_attn=True,
               is_q_zCct=False, n_z_samples_train=2, n_z_samples_test=8, encoded_path="both",
               pretrained="mnist/AttnLNP")
    m = build(cfg)
    dump("attnlnp_selfattn_pretrained", cfg, m, [
        run_case(m, cfg, "train_b2_c30_t50_nz2", True, _xy2(2, 30, 50, 1, 58, unit=True), "nll", True),
    ])


def main_baseline_shapes():
    """Fixtures at the shapes of BASELINE.json's configs (the sizes bench.py measures), from the real reference:
    python oracle/gen_golden.py baseline"""
    os.makedirs(OUT, exist_ok=True)
    # configs[1]: ConvCNP default ctor, C = T = 128 (8 of the 256 tasks; tests/test_gpu_baseline_shapes.py also runs B = 256)
    cfg = dict(family="ConvCNP", x_dim=1, y_dim=1)
    m = build(cfg)
    dump("baseline_convcnp_b8_c128_t128", cfg, m, [
        run_case(m, cfg, "train_b8_c128_t128", True, bench_offgrid(8, 128, 128, 101), "cnpf", True),
    ])
    # configs[2]: AttnCNP transformer attention (notebook cfg, 252 738 params), C = T = 512
    cfg = dict(family="AttnCNP", x_dim=1, y_dim=1, notebook=True, xy_hidden=128, attention="transformer", init_seed=21)
    m = build(cfg)
    dump("baseline_attncnp_b2_c512_t512", cfg, m, [
        run_case(m, cfg, "train_b2_c512_t512", True, bench_offgrid(2, 512, 512, 102), "cnpf", True),
        run_case(m, cfg, "eval_b2_c512_t512", False, offgrid_inputs(2, 512, 512, 1, 103), "cnpf", False),
    ])
    # configs[3]: GridConvCNP(1, 3) default ctor, 32x32, 30 % context
    cfg = dict(family="GridConvCNP", x_dim=1, y_dim=3, init_seed=22)
    m = build(cfg)
    dump("baseline_gridconvcnp_b4_32x32", cfg, m, [
        run_case(m, cfg, "train_b4_32x32", True, grid_inputs(4, 32, 32, 3, 0.3, 104), "cnpf", True),
    ])
    # configs[4]: GridConvLNP(1, 3, n_z_samples_train=16), 32x32 (eps regenerated from its seed: 16.8 MB otherwise)
    cfg = dict(family="GridConvLNP", x_dim=1, y_dim=3, n_z_samples_train=16, n_z_samples_test=4, is_q_zCct=False, init_seed=23)
    m = build(cfg)
    dump("baseline_gridconvlnp_b2_32x32_nz16", cfg, m, [
        run_case(m, cfg, "train_b2_32x32_nz16", True, grid_inputs(2, 32, 32, 3, 0.3, 105), "nll", True, eps_seed=9001),
    ])


def bench_offgrid(B, C, T, seed):
    """bench.py's synthetic inputs: X ~ U(-1, 1) unsorted, Y ~ N(0, 1)."""
    g = torch.Generator().manual_seed(seed)
    return dict(X_cntxt=torch.rand(B, C, 1, generator=g) * 2 - 1, Y_cntxt=torch.randn(B, C, 1, generator=g),
                X_trgt=torch.rand(B, T, 1, generator=g) * 2 - 1, Y_trgt=torch.randn(B, T, 1, generator=g))


def _xy2(B, C, T, y_dim, seed, unit=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, C + T, 2, generator=g) * 2 - 1
    y = torch.sin(3 * x.sum(-1, keepdim=True)) + 0.1 * torch.randn(B, C + T, y_dim, generator=g)
    if unit:  # image-like targets in [0, 1] for the checkpoints trained on pixels
        y = (0.5 + 0.4 * y).clamp(0, 1)
    return dict(X_cntxt=x[:, :C].contiguous(), Y_cntxt=y[:, :C].contiguous(), X_trgt=x[:, C:].contiguous(),
                Y_trgt=y[:, C:].contiguous())


def main_circular():
    """`model_2d_extrap` (wrap-around padding, upstream helpers.py:334-351, 406-414) with the upstream zsmms checkpoints, and the
    12-block `model_2d_XL` celeba128 checkpoint:  python oracle/gen_golden.py circular"""
    os.makedirs(OUT, exist_ok=True)
    cnn = dict(dim=2, norm="bn", bn_eps=1e-2, n_blocks=5, kernel_size=9, n_conv_layers=2)
    cfg = dict(family="GridConvCNP", x_dim=1, y_dim=1, notebook=True, circular=True, pretrained="zsmms/ConvCNP", cnn=cnn)
    m = build(cfg)
    dump("gridconvcnp_extrap_pretrained", cfg, m, [
        run_case(m, cfg, "eval_b2_40x40", False, blob_inputs(2, 40, 40, 0.2, 70), "cnpf", False),
        run_case(m, cfg, "eval_b1_28x20", False, blob_inputs(1, 28, 20, 0.3, 71), "cnpf", False),
    ])
    # (no train case here: train-mode batch statistics of these MNIST-trained weights on a 2-image batch make the reference's own fp32
    # gradients differ from its fp64 ones by 3e-3 -- nothing to pin a 1e-3 bar on; gradients through the wrap-around path are pinned by
    # the GridConvLNP fixture below, whose fp32 / fp64 gradients agree to 3e-5)
    cfg = dict(family="GridConvLNP", x_dim=1, y_dim=1, notebook=True, circular=True, is_global=False, pretrained="zsmms/ConvLNP",
               n_z_samples_train=3, n_z_samples_test=2, cnn=dict(cnn, n_blocks=4))
    m = build(cfg)
    dump("gridconvlnp_extrap_pretrained", cfg, m, [
        run_case(m, cfg, "train_b2_24x24_nz3", True, blob_inputs(2, 24, 24, 0.3, 72), "nll", True, eps_seed=720),
        run_case(m, cfg, "eval_b1_32x32_nz2", False, blob_inputs(1, 32, 32, 0.1, 73), "nll", False, eps_seed=730),
    ])
    cfg = dict(family="GridConvCNP", x_dim=1, y_dim=3, notebook=True, pretrained="celeba128/ConvCNPXL",
               cnn=dict(dim=2, norm="bn", n_blocks=12, kernel_size=9, n_conv_layers=2))
    m = build(cfg)
    dump("gridconvcnp_xl_pretrained", cfg, m, [
        run_case(m, cfg, "eval_b1_48x48", False, grid_inputs(1, 48, 48, 3, 0.15, 74), "cnpf", False),
        run_case(m, cfg, "train_b2_24x24", True, grid_inputs(2, 24, 24, 3, 0.3, 75), "cnpf", True),
    ])


if __name__ == "__main__":
    if sys.argv[1:] == ["attn"]:
        main_attn_family()
    elif sys.argv[1:] == ["circular"]:
        main_circular()
    elif sys.argv[1:] == ["baseline"]:
        main_baseline_shapes()
    else:
        main()
        main_attn_family()
        main_baseline_shapes()
        main_circular()
    os.system(f"du -sh {OUT}")
