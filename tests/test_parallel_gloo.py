"""World-size-2 gloo test (CPU) of the multi-GPU host logic: task sharding + the single flat-gradient all-reduce
reproduce the single-process gradient of the global batch mean."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "neural-process-family_b200"))
    from npf_b200.parallel import FlatGradients, shard_tasks
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # a stand-in differentiable module (plain torch on CPU): the flat-bucket logic is backend- and model-agnostic
    net = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))
    batch = dict(x=torch.randn(8, 5, 3), y=torch.randn(8, 5, 1))
    flat = FlatGradients(net)
    flat.zero_()
    mine = shard_tasks(batch, rank, world)
    per_task = ((net(mine["x"]) - mine["y"]) ** 2).sum((1, 2))
    per_task.mean().backward()  # local mean over the rank's tasks
    flat.all_reduce_mean()
    # single-process reference on the whole batch
    ref = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))
    ref.load_state_dict(net.state_dict())
    ((ref(batch["x"]) - batch["y"]) ** 2).sum((1, 2)).mean().backward()
    err = max((p.grad - q.grad).abs().max().item() for p, q in zip(net.parameters(), ref.parameters()))
    views_ok = all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in net.parameters())
    ret[rank] = (err, views_ok, flat.flat.numel())
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        err, views_ok, n = ret[r]
        assert err < 1e-6 and views_ok and n == 32 * 4  # four parameters, each padded to a 128-byte slot


def test_shard_tasks_rejects_ragged():
    from npf_b200.parallel import shard_tasks
    with pytest.raises(ValueError):
        shard_tasks(dict(x=torch.zeros(7, 2)), 0, 2)


def _syncbn_worker(rank, world, port, ret):
    """ConvCNP with BatchNorm CNN blocks (the notebooks' configuration), tasks sharded over 2 ranks with
    parallel.sync_batchnorm_: predictions of the local shard, the averaged flat gradient and the BatchNorm running statistics
    must equal a single-process run on the whole batch.  The CUDA entry points are replaced by their torch contracts
    (tests/test_host_wiring.py) -- what is under test is the collective plumbing, which is backend-agnostic."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "neural-process-family_b200"), os.path.join(root, "tests"), root]
    import test_host_wiring as hw
    from _cfg import build_model, loss_for
    from _util import load_fixture
    from npf_b200 import ops
    from npf_b200.neuralproc.base import NeuralProcessFamily
    from npf_b200.parallel import FlatGradients, shard_tasks, sync_batchnorm_
    import torch.nn.functional as F
    for name, fn in dict(mlp_chain=hw._mlp_chain, linear=hw._linear, merge_relu=hw._merge_relu, gauss_head=hw._gauss_head,
                         setconv=hw._setconv, dwconv=hw._dwconv, channel_moments=hw._channel_moments,
                         mean_pool=lambda x: x.mean(dim=1, keepdim=True)).items():
        setattr(ops, name, fn)
    NeuralProcessFamily._validate_inputs = lambda self, *a: None
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    fx = load_fixture("convcnp_notebook_pretrained")
    case = [c for c in fx["cases"] if c["training"]][0]            # B = 4 tasks
    crit = loss_for("cnpf", reduction="mean").train()

    def run(model, batch):
        model.train()
        flat = FlatGradients(model)
        flat.zero_()
        out = model(batch["X_cntxt"], batch["Y_cntxt"], batch["X_trgt"], batch["Y_trgt"])
        crit(out, batch["Y_trgt"]).backward()
        return out[0].base_dist.loc.detach(), flat

    ref = build_model(fx["cfg"]); ref.load_state_dict(fx["state_dict"])
    loc_ref, flat_ref = run(ref, case["inputs"])                   # single process, whole batch (local BN == global BN)
    m = sync_batchnorm_(build_model(fx["cfg"])); m.load_state_dict(fx["state_dict"])
    mine = shard_tasks(case["inputs"], rank, world)
    loc, flat = run(m, mine)
    flat.all_reduce_mean()
    s = case["inputs"]["X_cntxt"].shape[0] // world
    e_loc = (loc - loc_ref[:, rank * s:(rank + 1) * s]).abs().max().item() / loc_ref.abs().max().item()
    e_grad = ((flat.flat - flat_ref.flat).norm() / flat_ref.flat.norm()).item()
    sd, sd_ref = m.state_dict(), ref.state_dict()
    e_bn = max(((sd[k].float() - sd_ref[k].float()).abs().max() / sd_ref[k].float().abs().max().clamp_min(1e-12)).item()
               for k in sd if "running_" in k)
    # control: WITHOUT synchronisation the shard statistics differ and so do the predictions
    m2 = build_model(fx["cfg"]); m2.load_state_dict(fx["state_dict"])
    loc2, _ = run(m2, mine)
    e_local = (loc2 - loc_ref[:, rank * s:(rank + 1) * s]).abs().max().item() / loc_ref.abs().max().item()
    ret[rank] = (e_loc, e_grad, e_bn, e_local)
    dist.destroy_process_group()


def test_sync_batchnorm_world2_equals_single_process():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_syncbn_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        e_loc, e_grad, e_bn, e_local = ret[r]
        assert e_loc < 2e-5 and e_grad < 1e-4 and e_bn < 2e-5, ret[r]
        assert e_local > 1e-3, ret[r]          # local BN is a different model: the stated caveat of DESIGN.md section 6
