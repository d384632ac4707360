"""-m gpu: the fused flat-buffer Adam step (npf_adam_step / parallel.FlatAdam) tracks torch.optim.Adam step for step, including
weight decay, a changing learning rate (ExponentialLR) and gradient scaling, on a real model trained through GraphedStep."""
import copy

import pytest
import torch

from _cfg import build_model, loss_for
from _util import load_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(wd):
    import npf_b200
    from npf_b200.parallel import FlatAdam, FlatGradients
    npf_b200.set_precision("fp32")
    fx = load_fixture("cnp_default")
    c = [c for c in fx["cases"] if c["training"]][0]
    inp = [c["inputs"][k].float().cuda() for k in ("X_cntxt", "Y_cntxt", "X_trgt", "Y_trgt")]
    m1 = build_model(fx["cfg"]).cuda().train()
    m1.load_state_dict(fx["state_dict"])
    m2 = copy.deepcopy(m1)
    crit = loss_for(c["loss_name"], reduction="mean").train()
    ref = torch.optim.Adam(m1.parameters(), lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    flat = FlatGradients(m2)
    opt = FlatAdam(flat, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    step = npf_b200.GraphedStep(m2, crit, flat=flat)
    for it in range(6):
        ref.zero_grad(set_to_none=True)
        l1 = crit(m1(*inp), inp[3])
        l1.backward()
        ref.step()
        l2 = step(*inp)
        opt.step()
        assert abs(l1.item() - l2.item()) <= 2e-4 * max(1.0, abs(l1.item())), (it, l1.item(), l2.item())
        if it == 2:                      # ExponentialLR-style decay
            for g in ref.param_groups:
                g["lr"] *= 0.5
            opt.lr *= 0.5
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        err = ((p1 - p2).norm() / p1.norm().clamp_min(1e-12)).item()
        # the two models see gradients that differ in the last bits (atomic accumulation order, eager vs graph replay) and
        # Adam's g / sqrt(v) turns that into O(lr) differences on near-zero gradients: 2e-4 was observed, a wrong update
        # rule is off by >= 1e-2 (the exact-gradient comparison is test_flat_adam_grad_scale_and_state below)
        assert err < 1e-3, (n, err)
    # the module's parameters live in the flat buffer
    assert all(p.data_ptr() >= opt.param.data_ptr() and p.data_ptr() < opt.param.data_ptr() + 4 * opt.param.numel() for p in m2.parameters())


def test_flat_adam_grad_scale_and_state():
    import npf_b200
    from npf_b200.parallel import FlatAdam, FlatGradients
    lin = torch.nn.Linear(8, 4).cuda()
    ref = copy.deepcopy(lin)
    flat = FlatGradients(lin)
    opt = FlatAdam(flat, lr=1e-2)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    x = torch.randn(16, 8, device="cuda")
    for _ in range(3):
        flat.zero_()
        lin(x).square().mean().backward()          # torch ops: grads accumulate into the flat views
        ropt.zero_grad(set_to_none=True)
        ref(x).square().mean().backward()
        scale = min(1.0, 0.05 / float(opt.global_grad_norm()))
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        opt.step(grad_scale=scale)
        ropt.step()
    assert torch.allclose(lin.weight, ref.weight, rtol=1e-4, atol=1e-6) and torch.allclose(lin.bias, ref.bias, rtol=1e-4, atol=1e-6)
    sd = opt.state_dict()
    opt2 = FlatAdam(FlatGradients(copy.deepcopy(lin)), lr=5.0)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 3 and opt2.lr == 1e-2 and torch.equal(opt2.exp_avg, opt.exp_avg)


def test_flat_adam_device_side_clipping():
    """max_grad_norm: the global norm is reduced and applied on the device (npf_sqnorm + npf_adam_step_clipped) and must
    track torch.nn.utils.clip_grad_norm_ + torch.optim.Adam, both when the clip is active and when it is not."""
    from npf_b200.parallel import FlatAdam, FlatGradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 33), torch.nn.ReLU(), torch.nn.Linear(33, 4)).cuda()
    ref = copy.deepcopy(net)
    flat = FlatGradients(net)
    opt = FlatAdam(flat, lr=1e-2, weight_decay=1e-3)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-3)
    x = torch.randn(64, 8, device="cuda")
    for it, max_norm in enumerate([0.05, 0.05, 1e3, 0.2, 1e3]):
        flat.zero_()
        (net(x).square().mean() * 3).backward()
        ropt.zero_grad(set_to_none=True)
        (ref(x).square().mean() * 3).backward()
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        opt.step(max_grad_norm=max_norm)
        ropt.step()
        assert abs(opt.last_grad_norm().item() - total.item()) <= 1e-5 * total.item(), it
        opt.lr *= 0.9                                   # ExponentialLR(gamma=0.9)
        for g in ropt.param_groups:
            g["lr"] *= 0.9
    for p1, p2 in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p1, p2, rtol=1e-4, atol=1e-6), (p1 - p2).abs().max().item()
