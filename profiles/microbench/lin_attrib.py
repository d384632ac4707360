"""micro-benchmark of npf_linear_fwd / bwd_data on the hot shape (attribution experiments; not a bench value)"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "neural-process-family_b200"))
from npf_b200 import _cabi
M = int(os.environ.get("M", 75776)); K = N = 128
dev = "cuda"
X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / 11; b = torch.randn(N, device=dev)
Y = torch.empty(M, N, device=dev); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
prec = int(os.environ.get("PREC", 2))
def fwd(): _cabi.call("npf_linear_fwd", X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, M, K, N, 1, 0, 0, 0, prec, st)
def bwd(): _cabi.call("npf_linear_bwd_data", X.data_ptr(), N, W.data_ptr(), K, Y.data_ptr(), K, M, K, N, X.data_ptr(), K, 0, prec, st)
for name, f in (("fwd", fwd), ("bwd_data", bwd)):
    for _ in range(3): f()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"dbg={os.environ.get('NPF_WS_DBG','0'):>2} {name:9s} M={M} median {ts[len(ts)//2]:7.1f} us  min {ts[0]:7.1f} us")
