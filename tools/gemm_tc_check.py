"""Check pd_gemm_tf32 (tcgen05) against an fp64 matmul for the three layouts; optional timing."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import _lib

dev = torch.device("cuda:0")
CFG = int(os.environ.get("PD_GEMM_CFG", "0"))
st = lambda: torch.cuda.current_stream().cuda_stream


def run(name, M, N, K, layout, acc=0, bias=True, pad=4, time_it=False):
    torch.manual_seed(0)
    if layout == "tn":
        A = torch.randn(K, M + pad, device=dev); sam, sak = 1, M + pad; Am = A[:, :M].t()
    else:
        A = torch.randn(M, K + pad, device=dev); sam, sak = K + pad, 1; Am = A[:, :K]
    if layout == "nt":
        B = torch.randn(N, K + pad, device=dev); sbk, sbn = 1, K + pad; Bm = B[:, :K].t()
    else:
        B = torch.randn(K, N + pad, device=dev); sbk, sbn = N + pad, 1; Bm = B[:, :N]
    ldc = N + pad
    C = torch.randn(M, ldc, device=dev)
    C0 = C.clone()
    b = torch.randn(N, device=dev) if bias else None
    ref = Am.double() @ Bm.double()
    if bias:
        ref = ref + b.double()
    if acc:
        ref = ref + C0[:, :N].double()
    try:
        if CFG and name == "pd_gemm_tf32":
            _lib.call("pd_gemm_tf32_cfg", A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc,
                      None if b is None else b.data_ptr(), M, N, K, acc, CFG, st())
        else:
            _lib.call(name, A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc,
                      None if b is None else b.data_ptr(), M, N, K, acc, st())
        torch.cuda.synchronize()
    except RuntimeError as e:
        print(f"{name} {layout} M={M} N={N} K={K}: ERROR {e}")
        return
    err = (C[:, :N].double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    untouched = torch.equal(C[:, N:], C0[:, N:])
    msg = f"{name} {layout} M={M:6d} N={N:5d} K={K:6d} acc={acc} max_abs_err={err:.3e} (ref max {scale:.2f}) rel={err/scale:.2e} pad_ok={untouched}"
    if time_it:
        for _ in range(3):
            _lib.call(name, A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc, None, M, N, K, 0, st())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.call(name, A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc, None, M, N, K, 0, st())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        msg += f"  {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s"
    print(msg, flush=True)


layouts = sys.argv[1].split(",") if len(sys.argv) > 1 else ["nt"]
for lay in layouts:
    run("pd_gemm_tf32", 128, 128, 32, lay, bias=False, pad=0)
    run("pd_gemm_tf32", 128, 256, 64, lay, bias=False, pad=0)
    run("pd_gemm_tf32", 256, 512, 128, lay)
    run("pd_gemm_tf32", 300, 130, 200, lay, acc=1)
    run("pd_gemm_tf32", 1000, 64, 36, lay)
    run("pd_gemm_tf32", 512, 3072, 1024, lay, time_it=True)
    run("pd_gemm_tf32", 16384, 1536, 512, lay, time_it=True)
    run("pd_gemm_tf32", 1536, 512, 245760 // 4, lay, time_it=True)
    run("pd_gemm_f32", 16384, 1536, 512, lay, time_it=True)
