import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import _lib
dev = torch.device("cuda:0")
st = lambda: torch.cuda.current_stream().cuda_stream
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def run(M, N, K, cfgs):
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.zeros(M, N, device=dev)
    for cfg in cfgs:
        r = []
        for dbg in (0, 1, 2, 3):
            r.append(t(lambda: _lib.call("pd_gemm_tf32_cfg", A.data_ptr(), K, 1, B.data_ptr(), 1, K, C.data_ptr(), N, None, M, N, K, 0, cfg + 1000000 * dbg, st())))
        print(f"M={M} N={N} K={K} cfg={cfg}: full {r[0]:.1f}us | no-epilogue {r[1]:.1f} | no-mma {r[2]:.1f} | tma-only {r[3]:.1f}", flush=True)
run(16384, 1536, 512, [25622, 25641, 12823, 125631])
run(16384, 1536, 2048, [25622, 25641])
run(16384, 512, 1536, [25622])
