"""micro-benchmark of the induced -> target SetConv (regular keys, 128 channels): per-entry-point CUDA-event times"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "neural-process-family_b200"))
import npf_b200
from npf_b200 import _cabi, ops
npf_b200.set_precision("bf16x3")
B, K, Q, C, N = 256, 296, 128, 128, 128
torch.manual_seed(0)
keys = torch.linspace(-1.15, 1.15, K, device="cuda")
qs = (torch.rand(B, Q, device="cuda") * 2 - 1)
V = torch.randn(B, K, C, device="cuda", requires_grad=True)
theta = torch.tensor([float(os.environ.get("THETA", -4.4))], device="cuda", requires_grad=True)   # sigma ~ 0.012
W = (torch.randn(N, C + 1, device="cuda") / 11).requires_grad_(True); b = torch.zeros(N, device="cuda", requires_grad=True)
go = torch.randn(B, Q, N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
_cabi.enable_timing(True)
for it in range(8):
    flush.zero_()
    out = ops.setconv(keys, qs, V, theta, W, b, keys_regular=True)
    flush.zero_()
    out.backward(go)
    if it == 2:
        torch.cuda.synchronize(); _cabi.collect_timing()
torch.cuda.synchronize()
kt = _cabi.collect_timing()
for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {1e3 * v[0] / 5:8.1f} us/iter  {v[1] // 5} calls")
