import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import _lib
dev = torch.device("cuda:0")
st = lambda: torch.cuda.current_stream().cuda_stream
def run(M, N, K, pad=8, bias=True):
    torch.manual_seed(0)
    A = torch.randn(M, K + pad, device=dev); B = torch.randn(N, K + 4, device=dev)
    C = torch.zeros(M, N + 4, device=dev); b = torch.randn(N, device=dev) if bias else None
    _lib.call("pd_gemm_tf32", A.data_ptr(), K + pad, 1, B.data_ptr(), 1, K + 4, C.data_ptr(), N + 4, None if b is None else b.data_ptr(), M, N, K, 0, st())
    torch.cuda.synchronize()
    ref = A[:, :K].double() @ B[:, :K].double().t() + (b.double() if bias else 0)
    err = (C[:, :N].double() - ref).abs()
    bad = (err > 0.05).nonzero()
    print(f"M={M} N={N} K={K} pad={pad} bias={bias}: max err {err.max().item():.3e}; bad count {bad.shape[0]}; first bad {bad[:4].tolist()}; rows bad {sorted(set(bad[:,0].tolist()))[:8]} cols bad range {(bad[:,1].min().item(), bad[:,1].max().item()) if bad.numel() else None}")
for M in (1, 3, 4, 6, 8, 128):
    run(M, 3072, 36)
run(4, 3072, 36, bias=False)
run(4, 256, 36)
run(4, 3072, 64)
run(4, 3072, 36, pad=0)
