"""-m gpu: end-to-end parity of the npf_b200 models (CUDA, through the C ABI) with
  (a) the golden vectors produced by the real reference (tests/golden, incl. upstream pretrained checkpoints), and
  (b) the CPU oracle run here on the same inputs (full gradients, not only projections).
Bars (north_star): predictive mu, sigma and the NLL/ELBO within 1e-4 relative (fp32 path); gradients within 1e-3 of
the largest gradient entry of the parameter (fp32 accumulation order differs: atomics / split reductions)."""
import pytest
import torch

from _cfg import build_model, loss_for
from _util import fixture_names, grad_projection, load_fixture, oracle_run, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
GRAD_TOL = 1e-3


def _to(inp, dev):
    return {k: v.to(dev) for k, v in inp.items()}


def _run_cuda(model, cfg, case):
    model.train(case["training"])
    if "extrap" in case:
        model.set_extrapolation(tuple(case["extrap"]))
    if "eps" in case:
        model._eps_override = case["eps"].cuda()
    model.zero_grad(set_to_none=True)
    inp = _to(case["inputs"], "cuda")
    crit = loss_for(case["loss_name"])
    crit.train(case["training"])
    out = model(inp["X_cntxt"], inp["Y_cntxt"], inp["X_trgt"], inp["Y_trgt"])
    per_task = crit(out, inp["Y_trgt"])
    return out, per_task


@pytest.mark.parametrize("name", fixture_names())
def test_model_parity(name):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import npf_b200
    npf_b200.set_precision("fp32")
    fx = load_fixture(name)
    cfg, sd = fx["cfg"], fx["state_dict"]
    model = build_model(cfg)
    model.load_state_dict(sd)
    model.cuda()
    type(model).strict_validation = True
    for case in fx["cases"]:
        tag = f"{name}/{case['name']}"
        model.load_state_dict(sd)  # BatchNorm running stats back to the fixture's
        (p, z, q_c, q_ct), per_task = _run_cuda(model, cfg, case)
        loc, scale = p.base_dist.loc, p.base_dist.scale
        assert tuple(loc.shape) == tuple(case["loc"].shape), tag
        assert tuple(p.batch_shape) == tuple(case["loc"].shape[:-1]), tag
        # (a) golden vectors of the real reference
        assert rel_err(loc, case["loc"]) < TOL, f"{tag} loc {rel_err(loc, case['loc'])}"
        assert rel_err(scale, case["scale"]) < TOL, f"{tag} scale {rel_err(scale, case['scale'])}"
        assert rel_err(per_task, case["loss_per_task"]) < TOL, f"{tag} loss {rel_err(per_task, case['loss_per_task'])}"
        if "q_loc" in case:
            assert rel_err(q_c.base_dist.loc, case["q_loc"]) < TOL, tag
            assert rel_err(q_c.base_dist.scale, case["q_scale"]) < TOL, tag
        if "bn_after" in case:
            for k, v in case["bn_after"].items():
                if "num_batches" in k:
                    assert int(model.state_dict()[k]) == int(v), f"{tag} {k}"
                else:
                    assert rel_err(model.state_dict()[k], v) < TOL, f"{tag} {k}"
        if "grad_proj" not in case:
            continue
        # (b) full gradients against the CPU oracle, projections against the golden pins
        per_task.mean(0).backward()
        ora = oracle_run(cfg, sd, case, torch.float32, with_grads=True)
        got = {k: v.grad for k, v in model.named_parameters() if v.grad is not None}
        assert set(got) == set(ora["grads"]), f"{tag}: {set(got) ^ set(ora['grads'])}"
        G = max(g.abs().max().item() for g in ora["grads"].values())
        Gn = max(v[-1].item() for v in case["grad_proj"].values())
        for k, g_ref in ora["grads"].items():
            g = got[k].detach().cpu()
            denom = max(g_ref.abs().max().item(), 1e-4 * G)
            err = (g.double() - g_ref.double()).abs().max().item() / denom
            assert err < GRAD_TOL, f"{tag} grad {k}: {err}"
            ref_p = case["grad_proj"][k]
            pd = max(ref_p[-1].abs().item(), 1e-4 * Gn)
            assert ((grad_projection(g) - ref_p).abs().max() / pd).item() < 10 * GRAD_TOL, f"{tag} grad-proj {k}"
    type(model).strict_validation = False


def test_out_of_range_raises_value_error():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from npf_b200 import ConvCNP
    m = ConvCNP(1, 1).cuda().train()
    Xc, Yc, Xt = torch.rand(2, 5, 1).cuda(), torch.rand(2, 5, 1).cuda(), torch.rand(2, 7, 1).cuda() * 3
    m(Xc, Yc, Xt)  # asynchronous check: must surface at the latest on validate_now()
    with pytest.raises(ValueError):
        m.validate_now()
    m.eval()
    m(Xc, Yc, Xt)  # no check in eval mode (extrapolation plots)


def test_full_size_properties_convcnp():
    """Config-2 size (B=256, C=T=128): size-independent properties instead of an oracle run --
    permutation invariance in the context set, equivariance to target order, batch independence."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from npf_b200 import ConvCNP
    torch.manual_seed(0)
    m = ConvCNP(1, 1).cuda().eval()
    B, C, T = 256, 128, 128
    Xc, Yc, Xt = torch.rand(B, C, 1).cuda() * 2 - 1, torch.randn(B, C, 1).cuda(), torch.rand(B, T, 1).cuda() * 2 - 1
    with torch.no_grad():
        p0 = m(Xc, Yc, Xt)[0].base_dist
        perm_c, perm_t = torch.randperm(C).cuda(), torch.randperm(T).cuda()
        p1 = m(Xc[:, perm_c], Yc[:, perm_c], Xt[:, perm_t])[0].base_dist
        p2 = m(Xc[17:18], Yc[17:18], Xt[17:18])[0].base_dist
    assert rel_err(p1.loc, p0.loc[:, :, perm_t]) < TOL and rel_err(p1.scale, p0.scale[:, :, perm_t]) < TOL
    assert rel_err(p2.loc, p0.loc[:, 17:18]) < TOL and rel_err(p2.scale, p0.scale[:, 17:18]) < TOL
