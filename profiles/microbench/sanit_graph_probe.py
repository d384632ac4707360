"""Does whole-network CUDA-graph capture of a pure-torch fwd+bwd survive compute-sanitizer?  (control experiment for
the cudaErrorStreamCaptureIsolation that memcheck reports on GraphedStep; it reports the same on this torch-only graph)"""
import torch
lin = torch.nn.Linear(64, 64).cuda()
x = torch.randn(32, 64, device="cuda")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        lin.zero_grad(set_to_none=True); lin(x).square().mean().backward()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
lin.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    loss = lin(x).square().mean(); loss.backward()
g.replay(); torch.cuda.synchronize()
print("pure-torch capture ok", float(loss))
