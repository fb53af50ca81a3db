import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
dev = torch.device("cuda:0")
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
R = 16384
dgh = torch.randn(R, 15, 1536, device=dev); W = torch.randn(1536, 512, device=dev); dh = torch.zeros(R, 512, device=dev)
h = torch.randn(R, 15, 512, device=dev); gh = torch.empty(R, 1536, device=dev)
print("note fwd  NT strided A (30720B)      us", t(lambda: ops.gemm_nt(h[:, 3], W, gh)))
print("note bwd  NN strided A (92160B) acc  us", t(lambda: ops.gemm_nn(dgh[:, 3], W, dh, accumulate=True)))
print("note bwd  NN strided A (92160B) noacc us", t(lambda: ops.gemm_nn(dgh[:, 3], W, dh, accumulate=False)))
dghc = dgh[:, 3].contiguous()
print("note bwd  NN contiguous A acc        us", t(lambda: ops.gemm_nn(dghc, W, dh, accumulate=True)))
dw = torch.zeros(1536, 512, device=dev)
print("note dW   TN strided both acc        us", t(lambda: ops.gemm_tn(dgh[:, 3], h[:, 2], dw, accumulate=True)))
B = 512
dg = torch.randn(B, 32, 3072, device=dev); hp = torch.randn(B, 32, 1024, device=dev); dW = torch.zeros(3072, 1024, device=dev)
print("time dW   TN strided (393KB) acc     us", t(lambda: ops.gemm_tn(dg[:, 5], hp[:, 4], dW, accumulate=True)))
dgc, hpc = dg[:, 5].contiguous(), hp[:, 4].contiguous()
print("time dW   TN contiguous acc          us", t(lambda: ops.gemm_tn(dgc, hpc, dW, accumulate=True)))
Wt = torch.randn(3072, 1024, device=dev); dhh = torch.zeros(B, 1024, device=dev)
print("time bwd  NN strided A acc           us", t(lambda: ops.gemm_nn(dg[:, 5], Wt, dhh, accumulate=True)))
print("time fwd  NT strided A               us", t(lambda: ops.gemm_nt(hp[:, 4], Wt, torch.empty(B, 3072, device=dev))))
