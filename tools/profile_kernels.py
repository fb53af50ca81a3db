"""Launch the training step's dominant kernels once each at B=512 shapes (for `ncu --set full`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
R = 16384
h = torch.randn(R, 15, 512, device=dev); W = torch.randn(1536, 512, device=dev) * 0.05; b = torch.randn(1536, device=dev)
gh = torch.empty(R, 1536, device=dev)
dgh = torch.randn(R, 15, 1536, device=dev); dh = torch.empty(R, 512, device=dev); dW = torch.empty(1536, 512, device=dev)
Q = 245760
h0 = torch.randn(Q, 64, device=dev)
par = [torch.randn(192, 5, device=dev) * 0.3, torch.randn(192, device=dev) * 0.1, torch.randn(192, 64, device=dev) * 0.2,
       torch.randn(192, device=dev) * 0.1, torch.rand(5, device=dev), torch.randn(2, 64, device=dev) * 0.3, torch.randn(2, device=dev) * 0.1]
lg = torch.empty(Q, 5, 2, device=dev); S = torch.empty(Q, 6, 72, device=dev); GX = torch.empty(Q, 6, 264, device=dev); dh0 = torch.empty(Q, 64, device=dev)
gi = torch.randn(R, 15, 1536, device=dev); gi2 = torch.randn(R, 1536, device=dev); rzn = torch.empty(R, 15, 1536, device=dev); hn = torch.empty(R, 15, 512, device=dev)
hout = torch.empty(R, 15, 512, device=dev)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    ops.gemm_nt(h[:, 3], W, gh, b)                                   # note-GRU recurrent GEMM (fwd)
    ops.gemm_nn(dgh[:, 3], W, dh)                                    # note-GRU dh GEMM (bwd, MN-major B)
    ops.gemm_tn(dgh.view(R * 15, 1536), h.view(R * 15, 512), dW)     # note-GRU dW_hh (split-K, MN-major A and B)
    ops._gates_fwd(gi[:, 3], gi2, gh, h[:, 2], hout[:, 3], rzn[:, 3], hn[:, 3], None, 3)
    ops._call("pd_dur_decode_fwd", h0.data_ptr(), 64, Q, *[p.data_ptr() for p in par], lg.data_ptr(), S.data_ptr(), st)
    ops._call("pd_dur_decode_bwd", S.data_ptr(), lg.data_ptr(), Q, *[p.data_ptr() for p in par], GX.data_ptr(), dh0.data_ptr(), 64, st)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
