#!/usr/bin/env python
"""Per-role timeline of CTA 0 of npf_mlp_chain_bwd (diagnostics hook npf_debug_set_trace): where does a block iteration go?
   python profiles/microbench/trace_timeline.py [M] [L]
roles: 0 producer thread 0 (1 = before the X-stage waits, 2 = after them, 3 = X block staged), 1 MMA thread (1 = before xfull
wait, 2 = after, 3 = all MMAs of the block issued + committed), 2 epilogue warp 0 lane 0 (1 = before tfull wait, 2 = after,
3 = TMEM drained, 4 = wgrad done + mask released, 5 = next-layer image written)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "neural-process-family_b200"))
from npf_b200 import _cabi  # noqa: E402


def main(M=32768, L=4):
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)
    Xs = [torch.relu(torch.randn(M, 128, generator=g)).to(dev) for _ in range(L)]
    Ws = [(torch.randn(128, 128, generator=g) * 128 ** -0.5).to(dev) for _ in range(L)]
    dWs = [torch.zeros(128, 128, device=dev) for _ in range(L)]
    dbs = [torch.zeros(128, device=dev) for _ in range(L)]
    dY = torch.randn(M, 128, generator=g).to(dev)
    dX = torch.empty(M, 128, device=dev)
    arr = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: _cabi.call("npf_mlp_chain_bwd", dY.data_ptr(), 128, arr(Xs), arr(Ws), dX.data_ptr(), 128, arr(dWs), arr(dbs), L, M, 128, 0, 2, st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); run(); ev[1].record(); torch.cuda.synchronize()
    print(f"kernel time (no trace, warm L2): {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
    flush = torch.empty(64 * 1024 * 1024, device=dev); flush.fill_(1.0)
    ev[0].record(); run(); ev[1].record(); torch.cuda.synchronize()
    print(f"kernel time (no trace, L2 flushed): {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    _cabi.call("npf_debug_set_trace", buf.data_ptr())
    flush.fill_(2.0)
    run()
    torch.cuda.synchronize()
    _cabi.call("npf_debug_set_trace", None)
    h = buf.cpu().tolist()
    recs = []
    for role in range(3):
        n = h[role * 256]
        for i in range(n):
            v = h[role * 256 + 1 + i] & 0xFFFFFFFFFFFFFFFF
            recs.append((v & 0x00FFFFFFFFFFFFFF, role, v >> 56))
    recs.sort()
    t0 = recs[0][0]
    print("cycles since first record | role event")
    last = {0: t0, 1: t0, 2: t0}
    for t, role, evn in recs[:220]:
        print(f"{t - t0:9d}  (+{t - last[role]:6d})  " + "    " * role * 4 + f"r{role} e{evn}")
        last[role] = t
    print("span:", recs[-1][0] - t0, "cycles;", len(recs), "records")


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
