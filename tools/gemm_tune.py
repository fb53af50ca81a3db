"""Time pd_gemm_tf32_cfg tile configurations on the training step's dominant GEMM shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import _lib
dev = torch.device("cuda:0")
st = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def t(fn, n=10):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3

def shape(name, M, N, K, layout, cfgs, acc=0):
    if layout == "tn":
        A = torch.randn(K, M, device=dev); sam, sak = 1, M
    else:
        A = torch.randn(M, K, device=dev); sam, sak = K, 1
    if layout == "nt":
        B = torch.randn(N, K, device=dev); sbk, sbn = 1, K
    else:
        B = torch.randn(K, N, device=dev); sbk, sbn = N, 1
    C = torch.zeros(M, N, device=dev)
    out = []
    for cfg in cfgs:
        try:
            us = t(lambda: _lib.call("pd_gemm_tf32_cfg", A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), N, None, M, N, K, acc, cfg, st()))
            out.append(f"{cfg}:{us:7.1f}us/{2.0*M*N*K/us/1e6:5.0f}TF")
        except RuntimeError as e:
            out.append(f"{cfg}:ERR")
    print(f"{name:34s} {layout} M={M:6d} N={N:5d} K={K:6d} | " + "  ".join(out), flush=True)

if __name__ == "__main__":
    big = [25622, 12823, 925641, 912861]
    shape("note fwd step", 16384, 1536, 512, "nt", big)
    shape("gi_tok (output-bound)", 262144, 1536, 128, "nt", big)
    shape("emb gi (output-bound)", 262144, 384, 128, "nt", big)
    shape("time fwd step", 512, 3072, 1024, "nt", [12823, 6441, 6433, 912861])
    shape("note bwd dh", 16384, 512, 1536, "nn", big)
    shape("time bwd dh", 512, 1024, 3072, "nn", [12823, 6441, 6433])
    shape("note dW (split-K)", 1536, 512, 245760, "tn", big)
    shape("pitch head", 245760, 136, 512, "nt", [25622, 12823, 925641, 912861])
    shape("pitch dX", 245760, 512, 136, "nn", big)
    shape("dur_hid (N=64)", 245760, 64, 512, "nt", [6441, 6433])
    shape("note dX tok", 262144, 128, 1536, "nn", [12823, 25622, 912861])
