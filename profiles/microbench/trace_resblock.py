#!/usr/bin/env python
"""Per-role timeline of CTA 0 of npf_resblock1d_fwd (diagnostics hook npf_debug_set_trace).
roles: 0 producer thread 0 (1 before raw wait, 2 raw landed + sync, 3 depthwise done, 4 next TMA issued, 5 image free, 6 image stored),
1 MMA thread (1 before afull wait, 2 after, 3 issued + committed), 2 epilogue warp 0 lane 0 (1 before tfull wait, 2 after)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "neural-process-family_b200"))
from npf_b200 import _cabi  # noqa: E402


def main(B=256, L=384, bwd=0):
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, L, 128, generator=g).to(dev)
    wd, bd = (torch.randn(128, 11, generator=g) * 0.3).to(dev), torch.randn(128, generator=g).to(dev)
    wp, bp = (torch.randn(128, 128, generator=g) * 128 ** -0.5).to(dev), torch.randn(128, generator=g).to(dev)
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: _cabi.call("npf_resblock1d_fwd", x.data_ptr(), wd.data_ptr(), bd.data_ptr(), wp.data_ptr(), bp.data_ptr(), None, y.data_ptr(), B, L, 128, 11, 2, st)
    if bwd:      # backward roles: 0 producer (1 before X wait, 2 X landed, 3 O computed, 4 images free, 5 images stored), 1 MMA (1 before afull,
        #          2 after, 3 TMEM free, 4 issued), 2 epilogue warp 0 (1 before tfull, 2 after, 3 tile done)
        dy = torch.randn(B, L, 128, generator=g).to(dev)
        dx = torch.empty_like(x)
        gw = [torch.zeros_like(t) for t in (wd, bd, wp, bp)]
        run = lambda: _cabi.call("npf_resblock1d_bwd", dy.data_ptr(), x.data_ptr(), wd.data_ptr(), bd.data_ptr(), wp.data_ptr(), dx.data_ptr(), gw[0].data_ptr(),
                                 gw[1].data_ptr(), gw[2].data_ptr(), gw[3].data_ptr(), B, L, 128, 11, 2, st)
    for _ in range(3):
        run()
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    flush.fill_(1.0); ev[0].record(); run(); ev[1].record(); torch.cuda.synchronize()
    print(f"kernel time (L2 flushed): {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    _cabi.call("npf_debug_set_trace", buf.data_ptr())
    flush.fill_(2.0); run(); torch.cuda.synchronize()
    _cabi.call("npf_debug_set_trace", None)
    h = buf.cpu().tolist()
    recs = []
    for role in range(3):
        for i in range(h[role * 256]):
            v = h[role * 256 + 1 + i] & 0xFFFFFFFFFFFFFFFF
            recs.append((v & 0x00FFFFFFFFFFFFFF, role, v >> 56))
    recs.sort()
    t0 = recs[0][0]
    last = {0: t0, 1: t0, 2: t0}
    for t, role, evn in recs[:150]:
        print(f"{t - t0:9d}  (+{t - last[role]:6d})  " + "    " * role * 4 + f"r{role} e{evn}")
        last[role] = t
    print("span:", recs[-1][0] - t0, "cycles;", len(recs), "records")


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
