"""-m gpu: npf_b200.GraphedStep (CUDA-graph replay of forward + loss + backward) gives the eager step's loss and
gradients, leaves the model's buffers as one eager step would, re-captures per shape signature, and still raises the
training-time [-1, 1] range error."""
import copy

import pytest
import torch

from _cfg import build_model, loss_for
from _util import load_fixture

pytestmark = pytest.mark.gpu


def _inputs(fx):
    c = [c for c in fx["cases"] if c["training"]][0]
    return [c["inputs"][k].float().cuda() for k in ("X_cntxt", "Y_cntxt", "X_trgt", "Y_trgt")], c["loss_name"]


@pytest.mark.parametrize("name", ["convcnp_default", "cnp_default", "attncnp_multihead_y2", "gridconvcnp_default_y1", "lnp_default", "convlnp_default"])
def test_graphed_step_matches_eager(name):
    import npf_b200
    from npf_b200.parallel import FlatGradients
    npf_b200.set_precision("fp32")
    fx = load_fixture(name)
    inp, loss_name = _inputs(fx)
    m1 = build_model(fx["cfg"]).cuda().train()
    m1.load_state_dict(fx["state_dict"])
    m2 = copy.deepcopy(m1)
    crit = loss_for(loss_name, reduction="mean").train()
    latent = hasattr(m1, "n_z_samples_train")
    f1 = FlatGradients(m1)
    step = npf_b200.GraphedStep(m2, crit)
    for it in range(3):                       # replay more than once: buffers (BatchNorm statistics) must track the eager model
        f1.zero_()
        if latent:
            torch.manual_seed(1234 + it)
        l1 = crit(m1(*inp), inp[3])
        l1.backward()
        if latent:
            torch.manual_seed(1234 + it)
        l2 = step(*inp)
        if latent:                            # graph replays draw their own eps: compare statistics only loosely
            assert torch.isfinite(l2) and abs(l2.item() - l1.item()) < 0.5 * abs(l1.item()) + 1.0
            continue
        assert abs(l1.item() - l2.item()) <= 1e-5 * max(1.0, abs(l1.item())), (it, l1.item(), l2.item())
        g1, g2 = f1.flat, step.flat.flat
        assert ((g1 - g2).norm() / g1.norm().clamp_min(1e-20)).item() < 1e-4     # atomics reorder sums
        for (n, b1), (_, b2) in zip(m1.named_buffers(), m2.named_buffers()):
            assert torch.allclose(b1.float(), b2.float(), rtol=1e-5, atol=1e-6), (it, n)
    assert step.last_launches > 0


def test_graphed_step_new_shapes_and_range_error():
    import npf_b200
    npf_b200.set_precision("fp32")
    fx = load_fixture("cnp_default")
    m = build_model(fx["cfg"]).cuda().train()
    crit = loss_for("cnpf", reduction="mean").train()
    step = npf_b200.GraphedStep(m, crit, max_graphs=2)
    g = torch.Generator().manual_seed(0)
    for nc, nt in ((5, 7), (9, 7), (5, 7), (3, 11)):
        xc, xt = torch.rand(4, nc, 1, generator=g) * 2 - 1, torch.rand(4, nt, 1, generator=g) * 2 - 1
        loss = step(xc.cuda(), torch.randn(4, nc, 1, generator=g).cuda(), xt.cuda(), torch.randn(4, nt, 1, generator=g).cuda())
        assert torch.isfinite(loss)
    assert len(step._graphs) == 2
    bad = torch.full((4, 3, 1), 1.5).cuda()
    step(bad, torch.zeros(4, 3, 1).cuda(), torch.zeros(4, 11, 1).cuda(), torch.zeros(4, 11, 1).cuda())
    with pytest.raises(ValueError):
        m.validate_now()


def test_pipelined_step_matches_sequential():
    """npf_b200.PipelinedStep: batches from pinned host memory, the copy of step i+1 and the loss read of step i-1 overlapping step i;
    the losses it returns (one step late) must be the ones a plain sequence of GraphedStep calls on the same batches gives."""
    import npf_b200
    npf_b200.set_precision("fp32")
    fx = load_fixture("convcnp_default")
    case = fx["cases"][1]
    crit = loss_for("cnpf", reduction="mean").train()
    batches = []
    for i in range(5):
        g = torch.Generator().manual_seed(i)
        b = {k: v.clone() for k, v in case["inputs"].items()}
        b["Y_cntxt"] = b["Y_cntxt"] + 0.1 * torch.randn(b["Y_cntxt"].shape, generator=g)
        b["Y_trgt"] = b["Y_trgt"] + 0.1 * torch.randn(b["Y_trgt"].shape, generator=g)
        batches.append({k: v.float().pin_memory() for k, v in b.items()})
    m1 = build_model(fx["cfg"]).cuda().train(); m1.load_state_dict(fx["state_dict"])
    m2 = copy.deepcopy(m1)
    s1 = npf_b200.GraphedStep(m1, crit)
    ref = [float(s1(*[b[k].cuda() for k in ("X_cntxt", "Y_cntxt", "X_trgt", "Y_trgt")]).item()) for b in batches]
    pipe = npf_b200.PipelinedStep(npf_b200.GraphedStep(m2, crit))
    got = [pipe.submit(b) for b in batches]
    assert got[0] is None
    got = got[1:] + [pipe.drain()]
    for a, b in zip(ref, got):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (ref, got)
