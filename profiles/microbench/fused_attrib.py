import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "neural-process-family_b200"))
from npf_b200 import _cabi
M = int(os.environ.get("M", 75776)); K = N = 128
dY = torch.randn(M, N, device="cuda"); X = torch.relu(torch.randn(M, K, device="cuda")); W = torch.randn(N, K, device="cuda") / 11
dW = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda"); dX = torch.empty(M, K, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def f(): _cabi.call("npf_linear_bwd", dY.data_ptr(), N, X.data_ptr(), K, W.data_ptr(), K, dX.data_ptr(), K, dW.data_ptr(), K, db.data_ptr(), M, K, N, 16, 2, st)
for _ in range(2): f()
ts = []
for _ in range(4):
    flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("fused bwd M=%d: %s us" % (M, ["%.1f" % t for t in ts]))
