"""pd_gemm_bf16 (tcgen05 kind::f16, bf16 operands) vs fp64 for the three layouts + timing vs TF32."""
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
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3

def run(M, N, K, layout, acc=0, bias=True, pad=8, time_it=False):
    torch.manual_seed(0)
    if layout == "tn":
        A32 = torch.randn(K, M + pad, device=dev); sam, sak = 1, M + pad; Am = A32[:, :M].t()
    else:
        A32 = torch.randn(M, K + pad, device=dev); sam, sak = K + pad, 1; Am = A32[:, :K]
    if layout == "nt":
        B32 = torch.randn(N, K + pad, device=dev); sbk, sbn = 1, K + pad; Bm = B32[:, :K].t()
    else:
        B32 = torch.randn(K, N + pad, device=dev); sbk, sbn = N + pad, 1; Bm = B32[:, :N]
    A, B = A32.to(torch.bfloat16), B32.to(torch.bfloat16)
    Amr = (A[:, :M].t() if layout == "tn" else A[:, :K]).double()
    Bmr = (B[:, :K].t() if layout == "nt" else B[:, :N]).double()
    ldc = N + 4
    C = torch.randn(M, ldc, device=dev); C0 = C.clone()
    b = torch.randn(N, device=dev) if bias else None
    ref = Amr @ Bmr + (b.double() if bias else 0) + (C0[:, :N].double() if acc else 0)
    try:
        _lib.call("pd_gemm_bf16", A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc,
                  None if b is None else b.data_ptr(), M, N, K, acc, st())
        torch.cuda.synchronize()
    except RuntimeError as e:
        print(f"bf16 {layout} M={M} N={N} K={K}: ERROR {e}"); return
    err = (C[:, :N].double() - ref).abs().max().item(); scale = ref.abs().max().item()
    msg = f"bf16 {layout} M={M:6d} N={N:5d} K={K:6d} acc={acc} rel_err={err/scale:.2e} pad_ok={torch.equal(C[:, N:], C0[:, N:])}"
    if time_it:
        us = t(lambda: _lib.call("pd_gemm_bf16", A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc, None, M, N, K, 0, st()))
        us32 = t(lambda: _lib.call("pd_gemm_tf32", A32.data_ptr(), sam, sak, B32.data_ptr(), sbk, sbn, C.data_ptr(), ldc, None, M, N, K, 0, st()))
        msg += f"  bf16 {us:7.1f}us ({2.0*M*N*K/us/1e6:5.0f} TF)  tf32 {us32:7.1f}us"
    print(msg, flush=True)

for lay in sys.argv[1].split(","):
    run(128, 128, 64, lay, bias=False, pad=0)
    run(256, 512, 128, lay)
    run(300, 136, 200, lay, acc=1)
    run(1000, 64, 40, lay)
    run(512, 3072, 1024, lay, time_it=True)
    run(16384, 1536, 512, lay, time_it=True)
    run(1536, 512, 61440, lay, time_it=True)
