"""-m gpu: parity at the SHAPES of BASELINE.json's configs (the sizes bench.py measures), in both arithmetic modes
(`fp32` FFMA and `bf16x3`, the benched mode), whole-model gradients included.

  * test_baseline_fixture_parity -- the `baseline_*` golden vectors produced by the REAL reference
    (oracle/gen_golden.py::main_baseline_shapes): ConvCNP default B=8 C=T=128, AttnCNP transformer B=2 C=T=512,
    GridConvCNP(1,3) B=4 32x32 30 % context, GridConvLNP(1,3,n_z=16) B=2 32x32.  mu, sigma, per-task loss <= 1e-4 in
    BOTH modes; every parameter gradient (full tensor against the CPU oracle's autograd, max-abs error relative to the
    largest entry) <= 1e-3 in fp32 (measured <= 3e-5) and <= 1e-2 in bf16x3 (max-norm and L2).  The bf16x3 gradient bar is
    the north star's bf16 tier, not 1e-3, because of ReLU mask flips, not arithmetic: a unit whose pre-activation lies
    within the mode's 1.5e-5 forward error of zero switches one sample's gradient on or off, which moves a weight gradient
    by O(1/sqrt(#samples)) of a typical entry.  oracle/check_relu_flip_sensitivity.py reproduces the measured deviations
    (8.0e-3, 4.1e-3) to the digit with the fp64 oracle and 1e-5 noise on the linear outputs; trials without a flip sit at
    1e-4.
  * test_full_batch_slice_parity -- the bench-sized launch itself (ConvCNP B=256, AttnCNP B=64, GridConvCNP B=128,
    GridConvLNP B=64 x 16 z) with the loss restricted to a few tasks: predictions, loss and the gradients of those tasks
    must equal the oracle run on just those tasks.  This drives the persistent multi-task loops / 98 304-row tiles with
    real values (the tasks are independent in every benched configuration: no BatchNorm).
"""
import pytest
import torch

from _cfg import build_model, loss_for
from _util import load_fixture, oracle_run, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4         # mu, sigma, loss (north_star: 1e-4 rel fp32; bf16x3 is held to the same bar here)
GRAD_TOL = {"fp32": 1e-3, "bf16x3": 1e-2}    # whole-model gradients (bf16x3: ReLU-flip bound, see the module docstring)

BASELINE_FIXTURES = ["baseline_convcnp_b8_c128_t128", "baseline_attncnp_b2_c512_t512", "baseline_gridconvcnp_b4_32x32",
                     "baseline_gridconvlnp_b2_32x32_nz16"]


@pytest.fixture
def npf():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import npf_b200
    yield npf_b200
    npf_b200.set_precision("fp32")


def _grad_errors(model, ora_grads):
    """per parameter: max(max-abs error / largest entry, L2 error / L2 norm), both floored at 1e-4 of the largest gradient"""
    got = {k: v.grad for k, v in model.named_parameters() if v.grad is not None}
    assert set(got) == set(ora_grads), set(got) ^ set(ora_grads)
    G = max(g.abs().max().item() for g in ora_grads.values())
    Gn = max(g.double().norm().item() for g in ora_grads.values())
    errs = {}
    for k, g_ref in ora_grads.items():
        d = got[k].detach().double().cpu() - g_ref.double()
        e_max = d.abs().max().item() / max(g_ref.abs().max().item(), 1e-4 * G)
        e_l2 = d.norm().item() / max(g_ref.double().norm().item(), 1e-4 * Gn)
        errs[k] = max(e_max, e_l2)
    return errs


def _forward(model, case, inp, dev="cuda"):
    model.train(case["training"])
    if case.get("eps") is not None:
        model._eps_override = case["eps"].to(dev)
    model.zero_grad(set_to_none=True)
    crit = loss_for(case["loss_name"])
    crit.train(case["training"])
    out = model(inp["X_cntxt"], inp["Y_cntxt"], inp["X_trgt"], inp["Y_trgt"])
    return out, crit(out, inp["Y_trgt"])


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", BASELINE_FIXTURES)
def test_baseline_fixture_parity(npf, name, prec):
    npf.set_precision(prec)
    fx = load_fixture(name)
    cfg, sd = fx["cfg"], fx["state_dict"]
    model = build_model(cfg)
    model.load_state_dict(sd)
    model.cuda()
    for case in fx["cases"]:
        tag = f"{name}/{case['name']}/{prec}"
        inp = {k: v.cuda() for k, v in case["inputs"].items()}
        (p, z, q_c, q_ct), per_task = _forward(model, case, inp)
        e = dict(loc=rel_err(p.base_dist.loc, case["loc"]), scale=rel_err(p.base_dist.scale, case["scale"]),
                 loss=rel_err(per_task, case["loss_per_task"]))
        if "q_loc" in case:
            e["q_loc"], e["q_scale"] = rel_err(q_c.base_dist.loc, case["q_loc"]), rel_err(q_c.base_dist.scale, case["q_scale"])
        print(tag, {k: f"{v:.2e}" for k, v in e.items()})
        assert max(e.values()) < TOL, f"{tag}: {e}"
        if "grad_proj" not in case:
            continue
        per_task.mean(0).backward()
        ora = oracle_run(cfg, sd, case, torch.float32, with_grads=True)
        errs = _grad_errors(model, ora["grads"])
        worst = max(errs, key=errs.get)
        print(tag, "worst gradient", worst, f"{errs[worst]:.2e}")
        assert errs[worst] < GRAD_TOL[prec], f"{tag}: gradient of {worst}: {errs[worst]}"


# (fixture that carries cfg + weights, bench-sized B, task slices to check, seed)
FULL = {
    "convcnp_b256": ("baseline_convcnp_b8_c128_t128", 256, [(0, 8), (124, 132), (248, 256)]),
    "attncnp_b64": ("baseline_attncnp_b2_c512_t512", 64, [(0, 2), (62, 64)]),
    "gridconvcnp_b128": ("baseline_gridconvcnp_b4_32x32", 128, [(0, 2), (126, 128)]),
    "gridconvlnp_b64_nz16": ("baseline_gridconvlnp_b2_32x32_nz16", 64, [(31, 32), (63, 64)]),
}


def _bench_inputs(cfg, B, case, seed):
    g = torch.Generator().manual_seed(seed)
    if cfg["family"].startswith("Grid"):
        y = cfg["y_dim"]
        img = torch.rand(B, 32, 32, y, generator=g)
        mask = torch.zeros(B, 1024, dtype=torch.bool)
        for b in range(B):
            mask[b, torch.randperm(1024, generator=g)[:307]] = True
        return dict(X_cntxt=mask.view(B, 32, 32, 1), Y_cntxt=img, X_trgt=torch.ones(B, 32, 32, 1, dtype=torch.bool), Y_trgt=img.clone())
    C, T = case["inputs"]["X_cntxt"].shape[1], case["inputs"]["X_trgt"].shape[1]
    return dict(X_cntxt=torch.rand(B, C, 1, generator=g) * 2 - 1, Y_cntxt=torch.randn(B, C, 1, generator=g),
                X_trgt=torch.rand(B, T, 1, generator=g) * 2 - 1, Y_trgt=torch.randn(B, T, 1, generator=g))


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("which", list(FULL))
def test_full_batch_slice_parity(npf, which, prec):
    name, B, slices = FULL[which]
    npf.set_precision(prec)
    fx = load_fixture(name)
    cfg, sd, case0 = fx["cfg"], fx["state_dict"], fx["cases"][0]
    model = build_model(cfg)
    model.load_state_dict(sd)
    model.cuda()
    inp = _bench_inputs(cfg, B, case0, seed=777)
    dinp = {k: v.cuda() for k, v in inp.items()}
    eps = None
    if "eps" in case0:
        nz = case0["eps"].shape[0]
        eps = torch.randn(nz, B, *case0["eps"].shape[2:], generator=torch.Generator().manual_seed(778))
    for (a, b) in slices:
        tag = f"{which}[{a}:{b}]/{prec}"
        case = dict(training=True, loss_name=case0["loss_name"], inputs={k: v[a:b].contiguous() for k, v in inp.items()},
                    eps=None if eps is None else eps[:, a:b].contiguous())
        full_case = dict(training=True, loss_name=case0["loss_name"], eps=eps)
        (p, z, q_c, q_ct), per_task = _forward(model, full_case, dinp)
        per_task[a:b].mean(0).backward()                     # loss of the slice only, out of the full-size launch
        torch.cuda.synchronize()
        ora = oracle_run(cfg, sd, case, torch.float32, with_grads=True)
        e = dict(loc=rel_err(p.base_dist.loc[:, a:b], ora["loc"]), scale=rel_err(p.base_dist.scale[:, a:b], ora["scale"]),
                 loss=rel_err(per_task[a:b], ora["loss_per_task"]))
        errs = _grad_errors(model, ora["grads"])
        worst = max(errs, key=errs.get)
        print(tag, {k: f"{v:.2e}" for k, v in e.items()}, "worst gradient", worst, f"{errs[worst]:.2e}")
        assert max(e.values()) < TOL, f"{tag}: {e}"
        assert errs[worst] < GRAD_TOL[prec], f"{tag}: gradient of {worst}: {errs[worst]}"
