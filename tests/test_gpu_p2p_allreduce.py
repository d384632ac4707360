"""-m gpu, needs >= 2 GPUs (skipped otherwise): the one-kernel mean all-reduce over NVLink peer memory (csrc/p2p.cu,
parallel.P2PAllReduce) against NCCL's ReduceOp.AVG, over several epochs (the kernel advances its own epoch counter), for a bucket
of the ConvCNP size and a tiny one, and through FlatGradients."""
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
    from npf_b200.parallel import FlatGradients, P2PAllReduce
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    worst = 0.0
    for numel, two_shot in ((137476 + 60, False), (8, False), (137476 + 60, True), (12, True)):
        ar = P2PAllReduce(numel, dev)
        ar.two_shot = two_shot
        for epoch in range(6):
            g = torch.Generator(device="cpu").manual_seed(100 * epoch + rank)
            x = torch.randn(ar.n, generator=g).to(dev) * (epoch + 1)
            ref = x.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.AVG)
            ar.bucket.copy_(x)
            ar.reduce_()
            torch.cuda.synchronize()
            worst = max(worst, ((ar.bucket - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item())
    # through FlatGradients (what GraphedStep / bench.py use): p.grad views live in the shared bucket
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 4)).to(dev)
    flat = FlatGradients(net, process_group=dist.group.WORLD)
    used_p2p = flat.p2p is not None
    flat.zero_()
    for p in net.parameters():
        p.grad.fill_(float(rank + 1))
    flat.all_reduce_mean()
    torch.cuda.synchronize()
    mean = sum(range(1, world + 1)) / world
    e_flat = max((p.grad - mean).abs().max().item() for p in net.parameters())
    ret[rank] = (worst, e_flat, used_p2p)
    dist.destroy_process_group()


def test_p2p_allreduce_matches_nccl_world2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in range(2):
        worst, e_flat, used_p2p = ret[r]
        assert used_p2p, "peer access between the two GPUs of the box was expected"
        assert worst < 1e-6 and e_flat < 1e-6, ret[r]
