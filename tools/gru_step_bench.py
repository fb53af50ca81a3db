import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def t(fn, n=10):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
for (B, T, H, bc) in [(16384, 15, 512, True), (512, 32, 1024, True), (16384, 16, 128, False), (512, 8, 512, True)]:
    gi = torch.randn(B, T, 3 * H, device=dev); gi2 = torch.randn(B, 3 * H, device=dev) if bc else None
    h_all = torch.randn(B, T, H, device=dev); w = torch.randn(3 * H, H, device=dev) * 0.03; b = torch.randn(3 * H, device=dev)
    rzn = torch.empty(B, T, 3 * H, device=dev); hn = torch.empty(B, T, H, device=dev); gh = torch.empty(B, 3 * H, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def fused():
        ops._call("pd_gru_step_tf32", h_all[:, 2].data_ptr(), T * H, w.data_ptr(), H, b.data_ptr(), gi[:, 3].data_ptr(), T * 3 * H,
                  None if gi2 is None else gi2.data_ptr(), 3 * H, h_all[:, 3].data_ptr(), T * H, rzn[:, 3].data_ptr(), T * 3 * H,
                  hn[:, 3].data_ptr(), T * H, None, 3, B, H, st)
    def fused_tma():     # experimental persistent variant with TMA epilogue I/O (no length mask)
        ops._call("pd_gru_step_tma", h_all[:, 2].data_ptr(), T * H, w.data_ptr(), H, b.data_ptr(), gi[:, 3].data_ptr(), T * 3 * H,
                  None if gi2 is None else gi2.data_ptr(), 3 * H, h_all[:, 3].data_ptr(), T * H, rzn[:, 3].data_ptr(), T * 3 * H,
                  hn[:, 3].data_ptr(), T * H, B, H, st)
    def split():
        ops.gemm_nt(h_all[:, 2], w, gh, b)
        ops._gates_fwd(gi[:, 3], gi2, gh, h_all[:, 2], h_all[:, 3], rzn[:, 3], hn[:, 3], None, 3)
    line = f"B={B} H={H}: fused {t(fused):7.1f} us   gemm+gates {t(split):7.1f} us"
    split(); ref = h_all[:, 3].clone(); ref_rzn = rzn[:, 3].clone()
    for variant in (1, 2, 0):
        ops._lib.lib.pd_gru_step_tma_variant(variant)
        h_all[:, 3].zero_(); fused_tma(); torch.cuda.synchronize()
        err = float((h_all[:, 3] - ref).abs().max()); err_s = float((rzn[:, 3] - ref_rzn).abs().max())
        line += f"   fused-TMA v{variant} {t(fused_tma):7.1f} us (max |dh| {err:.2e}, |d rzn| {err_s:.2e})"
    if bc and H % 64 == 0:      # x-projection folded into the step (second K segment, K2 = 128): no gi read
        x = torch.randn(B, T + 1, 128, device=dev); wx = torch.randn(3 * H, 128, device=dev) * 0.05
        def fused_x():
            ops._call("pd_gru_step_tmax", h_all[:, 2].data_ptr(), T * H, w.data_ptr(), H, x[:, 3].data_ptr(), (T + 1) * 128,
                      wx.data_ptr(), 128, 128, b.data_ptr(), gi2.data_ptr(), 3 * H, h_all[:, 3].data_ptr(), T * H,
                      rzn[:, 3].data_ptr(), T * 3 * H, hn[:, 3].data_ptr(), T * H, B, H, st)
        gi[:, 3] = x[:, 3] @ wx.t()
        split(); ref = h_all[:, 3].clone()
        h_all[:, 3].zero_(); fused_x(); torch.cuda.synchronize()
        line += f"   fused-X {t(fused_x):7.1f} us (max |dh| {float((h_all[:, 3] - ref).abs().max()):.2e})"
    print(line, flush=True)
