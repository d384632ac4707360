"""Bisect which part of the GraphedStep body makes compute-sanitizer report cudaErrorStreamCaptureIsolation."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "neural-process-family_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import npf_b200
from npf_b200.parallel import FlatGradients
variant = sys.argv[1]
torch.manual_seed(0)
m = npf_b200.CNP(1, 1).cuda().train()
crit = npf_b200.CNPFLoss(reduction="mean").train()
flat = FlatGradients(m)
xc, yc = torch.rand(4, 5, 1, device="cuda") * 2 - 1, torch.randn(4, 5, 1, device="cuda")
xt, yt = torch.rand(4, 7, 1, device="cuda") * 2 - 1, torch.randn(4, 7, 1, device="cuda")
if variant == "eval":
    m.eval()
def body():
    if variant != "nozero":
        flat.zero_()
    out = m(xc, yc, xt, yt)
    loss = crit(out, yt)
    if variant not in ("fwdonly",):
        loss.backward()
    return loss.detach()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        body()
torch.cuda.current_stream().wait_stream(s)
if hasattr(m, "validate_now") and variant != "novalidate":
    m.validate_now()
g = torch.cuda.CUDAGraph()
SAME = os.environ.get("SAME_STREAM", "1") == "1"
try:
    with (torch.cuda.graph(g, stream=s) if SAME else torch.cuda.graph(g)):
        loss = body()
    g.replay(); torch.cuda.synchronize()
    print(variant, "capture ok", float(loss))
except Exception as e:
    print(variant, "capture FAILED", type(e).__name__, str(e)[:80])
