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
