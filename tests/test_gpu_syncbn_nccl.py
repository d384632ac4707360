"""-m gpu, needs >= 2 GPUs (skipped otherwise): synchronised BatchNorm over NCCL with the real CUDA kernels -- upstream's
pretrained BatchNorm ConvCNP, tasks sharded over 2 ranks, against a single-GPU run on the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "neural-process-family_b200"), os.path.join(root, "tests"), root]
    from _cfg import build_model, loss_for
    from _util import load_fixture
    import npf_b200
    from npf_b200.parallel import FlatGradients, shard_tasks, sync_batchnorm_
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    npf_b200.set_precision("fp32")
    fx = load_fixture("convcnp_notebook_pretrained")
    case = [c for c in fx["cases"] if c["training"]][0]
    inputs = {k: v.cuda() for k, v in case["inputs"].items()}
    crit = loss_for("cnpf", reduction="mean").train()

    def run(model, batch):
        model.cuda().train()
        flat = FlatGradients(model)
        flat.zero_()
        out = model(batch["X_cntxt"], batch["Y_cntxt"], batch["X_trgt"], batch["Y_trgt"])
        crit(out, batch["Y_trgt"]).backward()
        return out[0].base_dist.loc.detach(), flat

    ref = build_model(fx["cfg"]); ref.load_state_dict(fx["state_dict"])
    loc_ref, flat_ref = run(ref, inputs)
    m = sync_batchnorm_(build_model(fx["cfg"])); m.load_state_dict(fx["state_dict"])
    loc, flat = run(m, shard_tasks(inputs, rank, world))
    flat.all_reduce_mean()
    torch.cuda.synchronize()
    s = inputs["X_cntxt"].shape[0] // world
    e_loc = ((loc - loc_ref[:, rank * s:(rank + 1) * s]).abs().max() / loc_ref.abs().max()).item()
    e_grad = ((flat.flat - flat_ref.flat).norm() / flat_ref.flat.norm()).item()
    sd, sd_ref = m.state_dict(), ref.state_dict()
    e_bn = max(((sd[k].float() - sd_ref[k].float()).abs().max() / sd_ref[k].float().abs().max().clamp_min(1e-12)).item()
               for k in sd if "running_" in k)
    # the plain (no BatchNorm) model through GraphedStep on 2 ranks: replay + flat-gradient all-reduce == eager gradient of the whole batch
    fx2 = load_fixture("convcnp_default")
    case2 = fx2["cases"][1]
    inp2 = {k: v.cuda() for k, v in case2["inputs"].items()}
    r2 = build_model(fx2["cfg"]); r2.load_state_dict(fx2["state_dict"])
    _, flat_ref2 = run(r2, inp2)
    m2 = build_model(fx2["cfg"]); m2.load_state_dict(fx2["state_dict"]); m2.cuda().train()
    flat2 = FlatGradients(m2, process_group=dist.group.WORLD)
    gstep = npf_b200.GraphedStep(m2, crit, flat=flat2)
    shard = shard_tasks(inp2, rank, world)
    for _ in range(2):
        gstep(shard["X_cntxt"], shard["Y_cntxt"], shard["X_trgt"], shard["Y_trgt"])
    torch.cuda.synchronize()
    e_graph = ((flat2.flat - flat_ref2.flat).norm() / flat_ref2.flat.norm()).item()
    ret[rank] = (e_loc, e_grad, e_bn, e_graph)
    dist.destroy_process_group()


def test_sync_batchnorm_nccl_world2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in range(2):
        e_loc, e_grad, e_bn, e_graph = ret[r]
        assert e_loc < 1e-4 and e_grad < 1e-3 and e_bn < 1e-4 and e_graph < 1e-3, ret[r]
