import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "neural-process-family_b200"))
from npf_b200 import _cabi
K = N = 128
st = torch.cuda.current_stream().cuda_stream
for pr in (2, 1):
    for M in (128, 300, 1000):
        for mask in (0, 16):
            torch.manual_seed(0)
            dY = torch.randn(M, N, device="cuda"); X = torch.relu(torch.randn(M, K, device="cuda")); W = torch.randn(N, K, device="cuda") / 11
            dW = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda"); dX = torch.empty(M, K, device="cuda")
            _cabi.call("npf_linear_bwd", dY.data_ptr(), N, X.data_ptr(), K, W.data_ptr(), K, dX.data_ptr(), K, dW.data_ptr(), K, db.data_ptr(), M, K, N, mask, pr, st)
            torch.cuda.synchronize()
            ref = dY.double().sum(0)
            bad = ~torch.isfinite(db)
            print(f"prec={pr} M={M} mask={mask}: db nan count {int(bad.sum())}, err {((db.double()-ref).norm()/ref.norm()).item():.3e}; "
                  f"dW err {((dW.double()-dY.double().t()@X.double()).norm()/(dY.double().t()@X.double()).norm()).item():.3e} first db {db[:4].tolist()} ref {ref[:4].tolist()}")
