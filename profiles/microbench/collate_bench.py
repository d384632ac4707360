"""Context/target split of one meta-batch: on the device (npf_b200.utils.datasplit, data resident in HBM) against the
host-side collate the reference performs (per-row np.random.shuffle + torch.gather on the CPU, datasplit.py:108-145,
246-255, then the host->device copy of the four tensors).  Prints one JSON line per configuration.
    python profiles/microbench/collate_bench.py"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "neural-process-family_b200"))
from npf_b200.utils import datasplit as ds  # noqa: E402


def host_collate(X, Y, a, b):
    B, N, _ = X.shape
    n = random.randint(a, b)
    idx = np.arange(N).reshape(1, N).repeat(B, axis=0)
    for r in range(B):
        np.random.shuffle(idx[r])
    idx = torch.from_numpy(idx[:, :n])
    Xc = torch.gather(X, 1, idx.unsqueeze(-1).expand(B, -1, X.shape[-1])).contiguous()
    Yc = torch.gather(Y, 1, idx.unsqueeze(-1).expand(B, -1, Y.shape[-1])).contiguous()
    return [t.pin_memory().cuda(non_blocking=True) for t in (Xc, Yc, X, Y)]


def main():
    for (B, N, xd, yd, a, b) in [(256, 128, 1, 1, 0, 50), (256, 128, 1, 1, 128, 128), (1024, 1024, 2, 3, 0, 307)]:
        X, Y = torch.rand(B, N, xd) * 2 - 1, torch.randn(B, N, yd)
        Xd, Yd = X.cuda(), Y.cuda()
        getter = ds.CntxtTrgtGetter(contexts_getter=ds.GetRandomIndcs(a=a, b=b), targets_getter=ds.get_all_indcs)
        random.seed(0); np.random.seed(0)
        for _ in range(10):
            getter(Xd, Yd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 200
        t0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            getter(Xd, Yd)
        e1.record()
        torch.cuda.synchronize()
        wall_dev = (time.perf_counter() - t0) / iters
        dev_us = e0.elapsed_time(e1) * 1e3 / iters
        random.seed(0); np.random.seed(0)
        host_collate(X, Y, a, b)
        torch.cuda.synchronize()
        it_h = 20
        t0 = time.perf_counter()
        for _ in range(it_h):
            host_collate(X, Y, a, b)
        torch.cuda.synchronize()
        host_us = (time.perf_counter() - t0) / it_h * 1e6
        print(json.dumps(dict(workload=f"split_b{B}_n{N}_x{xd}_y{yd}_c{a}-{b}", device_us_per_batch=round(dev_us, 1),
                              device_wall_us_per_batch=round(wall_dev * 1e6, 1), host_collate_us_per_batch=round(host_us, 1),
                              ratio=round(host_us / max(wall_dev * 1e6, 1e-9), 1))), flush=True)


if __name__ == "__main__":
    main()
