"""CPU test of the HOST-SIDE wiring of the set-based model families (CNP, LNP, AttnCNP, AttnLNP incl. the self-attention
encoders): every ``npf_b200.ops`` entry these models call is replaced -- in this test only -- by a one-line torch
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

FAMILIES = ("CNP", "LNP", "AttnCNP", "AttnLNP")


def _mlp_chain(x, weights, biases, final_relu=False, precision=None):
    h = x
    for i, W in enumerate(weights):
        h = F.linear(h, W, None if biases is None or biases[i] is None else biases[i])
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
    monkeypatch.setattr(NeuralProcessFamily, "_validate_inputs", lambda self, *a: None)


@pytest.mark.parametrize("name", [n for n in fixture_names() if load_fixture(n)["cfg"]["family"] in FAMILIES])
def test_set_models_wiring_matches_reference_golden(torch_ops, name):
    torch.set_num_threads(4)
    fx = load_fixture(name)
    model = build_model(fx["cfg"])
    model.load_state_dict(fx["state_dict"])
    for case in fx["cases"]:
        tag = f"{name}/{case['name']}"
        model.train(case["training"])
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
        assert rel_err(p.base_dist.loc, case["loc"]) < 2e-5, tag
        assert rel_err(p.base_dist.scale, case["scale"]) < 2e-5, tag
        assert rel_err(per_task, case["loss_per_task"]) < 2e-5, tag
        if "q_loc" in case:
            assert rel_err(q_c.base_dist.loc, case["q_loc"]) < 2e-5 and rel_err(q_c.base_dist.scale, case["q_scale"]) < 2e-5, tag
        if "q_ct_loc" in case:
            assert rel_err(q_ct.base_dist.loc, case["q_ct_loc"]) < 2e-5, tag
        if "grad_proj" in case:
            per_task.mean(0).backward()
            got = {k: v.grad for k, v in model.named_parameters() if v.grad is not None}
            assert set(got) == set(case["grad_proj"]), f"{tag}: {set(got) ^ set(case['grad_proj'])}"
            G = max(v[-1].item() for v in case["grad_proj"].values())
            for k, g in got.items():
                ref = case["grad_proj"][k]
                denom = max(ref[-1].abs().item(), 1e-4 * G)
                assert ((grad_projection(g) - ref).abs().max() / denom).item() < 5e-4, f"{tag}/{k}"
