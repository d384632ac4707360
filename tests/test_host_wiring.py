"""CPU test of the HOST-SIDE wiring of every model family (CNP, LNP, AttnCNP, AttnLNP incl. the self-attention encoders,
ConvCNP, ConvLNP, GridConvCNP, GridConvLNP incl. BatchNorm folding, extrapolation grids and global latents): every ``npf_b200.ops`` entry these models call is replaced -- in this test only -- by a one-line torch
expression of the operator's documented contract, and the models are then run on the golden fixtures of the real
reference.  What is exercised is the module graph, argument order, broadcasting and the latent / attention plumbing of
``npf_b200/neuralproc`` and ``npf_b200/architectures``; the CUDA kernels behind the same entry points are checked
against the same fixtures by the ``-m gpu`` tests.  (The product has no CPU path: without this monkeypatch the models
raise on CPU tensors, tests/test_host_logic.py::test_no_cpu_fallback.)"""
import math

import pytest
import torch
import torch.nn.functional as F

from _cfg import build_model, loss_for
from _util import fixture_names, grad_projection, load_fixture, rel_err



def _mlp_chain(x, weights, biases, final_relu=False, precision=None):
    h = x
    for i, W in enumerate(weights):
        # a 1x1 convolution's weight [out, in, 1(, 1)] is read as [out, in] by the kernel
        h = F.linear(h, W.reshape(W.shape[0], -1), None if biases is None or biases[i] is None else biases[i])
        if i < len(weights) - 1 or final_relu:
            h = torch.relu(h)
    return h


def _linear(x, weight, bias=None, relu=False, precision=None):
    return _mlp_chain(x, [weight], None if bias is None else [bias], final_relu=relu)


def _merge_relu(x1, x2):
    return torch.relu(x1 + x2) if x2.dim() == 3 else torch.relu(x1.unsqueeze(0) + x2)


def _xattn(q, k, v, n_heads, scale):
    B, Tq, E = q.shape
    Tk, Ev = k.shape[1], v.shape[-1]
    qh = q.view(B, Tq, n_heads, E // n_heads).transpose(1, 2)
    kh = k.view(B, Tk, n_heads, E // n_heads).transpose(1, 2)
    vh = v.view(B, Tk, n_heads, Ev // n_heads).transpose(1, 2)
    w = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (w @ vh).transpose(1, 2).reshape(B, Tq, Ev)


def _gauss_head(suff, min_scale=0.01):
    loc, s = suff.split(suff.shape[-1] // 2, dim=-1)
    return loc, min_scale + (1 - min_scale) * F.softplus(s)


def _latent_sample(suff, eps):
    loc, s = suff.split(suff.shape[-1] // 2, dim=-1)
    scale = 0.1 + 0.9 * torch.sigmoid(s)
    return loc, scale, loc.unsqueeze(0) + scale.unsqueeze(0) * eps


def _setconv(keys, queries, values, theta, weight, bias, keys_regular=False):
    B = values.shape[0]
    k = keys.squeeze(-1) if keys.dim() == 3 else keys.expand(B, -1)
    q = queries.squeeze(-1) if queries.dim() == 3 else queries.expand(B, -1)
    sigma = 1e-5 + F.softplus(theta)
    inp = -((k[:, None, :] - q[:, :, None]).abs() / sigma) ** 2          # [B, Q, K]
    feat = torch.softmax(inp, dim=-1) @ values
    dens = torch.exp(inp).sum(-1, keepdim=True)
    return F.linear(torch.cat([feat, dens], dim=-1), weight, bias)


def _dwconv(x, weight, bias=None, res=None, relu_in=False, scale=None, shift=None):
    h = x if scale is None else x * scale + shift
    if relu_in:
        h = torch.relu(h)
    C = x.shape[-1]
    k = weight.shape[-1]
    if x.dim() == 3:
        o = F.conv1d(h.transpose(1, 2), weight, bias, padding=k // 2, groups=C).transpose(1, 2)
    else:
        o = F.conv2d(h.permute(0, 3, 1, 2), weight, bias, padding=k // 2, groups=C).permute(0, 2, 3, 1)
    return o if res is None else o + res


def _channel_moments(x):
    x2 = x.reshape(-1, x.shape[-1])
    return x2.mean(0), x2.var(0, unbiased=False)


def _gridconv_in(img, mask, weight):
    y, k = img.shape[-1], weight.shape[-1]
    m = mask.to(img.dtype).expand(*img.shape[:-1], y).permute(0, 3, 1, 2)
    w = weight.abs()
    sig = F.conv2d(img.permute(0, 3, 1, 2) * m, w, None, padding=k // 2, groups=y)
    den = F.conv2d(m, w, None, padding=k // 2, groups=y)
    return torch.cat([sig / den.clamp(min=1e-5), den], dim=1).permute(0, 2, 3, 1)


def _global_latent(z):
    half = z.shape[-1] // 2
    loc, glob = z.split(half, dim=-1)
    g = glob.reshape(z.shape[0], -1, half).mean(1).view(z.shape[0], *([1] * (z.dim() - 2)), half).expand_as(glob)
    return torch.cat([loc, g], dim=-1)


@pytest.fixture
def torch_ops(monkeypatch):
    from npf_b200 import ops
    from npf_b200.neuralproc.base import NeuralProcessFamily
    monkeypatch.setattr(ops, "mlp_chain", _mlp_chain)
    monkeypatch.setattr(ops, "linear", _linear)
    monkeypatch.setattr(ops, "merge_relu", _merge_relu)
    monkeypatch.setattr(ops, "mean_pool", lambda x: x.mean(dim=1, keepdim=True))
    monkeypatch.setattr(ops, "add_layernorm", lambda a, b, g, be: F.layer_norm(a + b, (a.shape[-1],), g, be, 1e-5))
    monkeypatch.setattr(ops, "xattn", _xattn)
    monkeypatch.setattr(ops, "gauss_head", _gauss_head)
    monkeypatch.setattr(ops, "latent_sample", _latent_sample)
    monkeypatch.setattr(ops, "setconv", _setconv)
    monkeypatch.setattr(ops, "dwconv", _dwconv)
    monkeypatch.setattr(ops, "channel_moments", _channel_moments)
    monkeypatch.setattr(ops, "gridconv_in", _gridconv_in)
    monkeypatch.setattr(ops, "global_latent", _global_latent)
    monkeypatch.setattr(NeuralProcessFamily, "_validate_inputs", lambda self, *a: None)


@pytest.mark.parametrize("name", fixture_names())
def test_model_wiring_matches_reference_golden(torch_ops, name):
    torch.set_num_threads(4)
    fx = load_fixture(name)
    model = build_model(fx["cfg"])
    model.load_state_dict(fx["state_dict"])
    for case in fx["cases"]:
        tag = f"{name}/{case['name']}"
        model.load_state_dict(fx["state_dict"])          # BatchNorm running statistics back to the fixture's
        model.train(case["training"])
        if "extrap" in case:
            model.set_extrapolation(tuple(case["extrap"]))
        model.zero_grad(set_to_none=True)
        if "eps" in case:
            model._eps_override = case["eps"]
        inp = case["inputs"]
        crit = loss_for(case["loss_name"])
        crit.train(case["training"])
        out = model(inp["X_cntxt"], inp["Y_cntxt"], inp["X_trgt"], inp["Y_trgt"])
        per_task = crit(out, inp["Y_trgt"])
        p, z, q_c, q_ct = out
        assert tuple(p.base_dist.loc.shape) == tuple(case["loc"].shape), tag
        # loc against max(|loc|, 1 % of the predictive std): without context the ConvCNP mean is ~2e-4 beside sigma = 0.7
        e_loc = (p.base_dist.loc.detach().double() - case["loc"].double()).abs().max().item() / max(
            case["loc"].abs().max().item(), 1e-2 * case["scale"].abs().max().item())
        assert e_loc < 2e-5, f"{tag} loc {e_loc}"
        assert rel_err(p.base_dist.scale, case["scale"]) < 2e-5, tag
        assert rel_err(per_task, case["loss_per_task"]) < 2e-5, tag
        if "q_loc" in case:
            assert rel_err(q_c.base_dist.loc, case["q_loc"]) < 2e-5 and rel_err(q_c.base_dist.scale, case["q_scale"]) < 2e-5, tag
        if "q_ct_loc" in case:
            assert rel_err(q_ct.base_dist.loc, case["q_ct_loc"]) < 2e-5, tag
        if "bn_after" in case:
            for k, v in case["bn_after"].items():
                assert ("num_batches" in k and int(model.state_dict()[k]) == int(v)) or rel_err(model.state_dict()[k], v) < 2e-5, f"{tag} {k}"
        if "grad_proj" in case:
            per_task.mean(0).backward()
            got = {k: v.grad for k, v in model.named_parameters() if v.grad is not None}
            assert set(got) == set(case["grad_proj"]), f"{tag}: {set(got) ^ set(case['grad_proj'])}"
            G = max(v[-1].item() for v in case["grad_proj"].values())
            for k, g in got.items():
                ref = case["grad_proj"][k]
                denom = max(ref[-1].abs().item(), 1e-4 * G)
                assert ((grad_projection(g) - ref).abs().max() / denom).item() < 5e-4, f"{tag}/{k}"
        if "extrap" in case:
            model.set_extrapolation((-1, 1))
