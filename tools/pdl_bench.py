"""Programmatic dependent launch on the recurrent chain: a CUDA graph of 32 dependent fused steps (512 x 1024, bf16
operands) and of 32 backward steps (gate gradients + split-K dgh.W_hh GEMM), replayed with ordinary and with PDL launches
(pd_set_pdl).  Only the fused step kernel opts in today (the backward pair was measured slower with PDL and reverted);
the backward line is the per-step latency reference of that chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops, _lib
dev = torch.device("cuda:0")
B, H, T = 512, 1024, 32
torch.manual_seed(0)
w = torch.randn(3 * H, H, device=dev) * 0.03
b = torch.randn(3 * H, device=dev) * 0.1
gi = torch.randn(B, T + 1, 3 * H, device=dev)
h = torch.zeros(B, T + 1, H, device=dev)
h[:, 0] = torch.randn(B, H, device=dev) * 0.3
rzn = torch.empty(B, T, 3 * H, device=dev); hn = torch.empty(B, T, H, device=dev)
wb = ops.to_bf16(w)
hb = torch.empty(2, B, H, device=dev, dtype=torch.bfloat16)
dout = torch.randn(B, T, H, device=dev) * 0.1
dgi = torch.empty(B, T, 3 * H, device=dev); dgh = torch.empty(B, T, 3 * H, device=dev)
bufs = torch.zeros(4, B, H, device=dev)
dghb = torch.empty(B, 3 * H, device=dev, dtype=torch.bfloat16)

def fwd():
    st = torch.cuda.current_stream().cuda_stream
    ops._call("pd_f32_to_bf16", h[:, 0].data_ptr(), h.stride(0), B, H, hb[0].data_ptr(), H, st)
    for t in range(T):
        ops._call("pd_gru_step_tma_bf16", hb[t & 1].data_ptr(), H, wb.data_ptr(), H, b.data_ptr(), gi[:, t].data_ptr(), gi.stride(0),
                  None, 0, h[:, t].data_ptr(), h.stride(0), h[:, t + 1].data_ptr(), h.stride(0), hb[(t & 1) ^ 1].data_ptr(), H,
                  rzn[:, t].data_ptr(), rzn.stride(0), hn[:, t].data_ptr(), hn.stride(0), B, H, st)

def bwd():
    st = torch.cuda.current_stream().cuda_stream
    dz = dm = None
    for t in range(T - 1, -1, -1):
        nz = bufs[1] if dz is bufs[0] else bufs[0]
        nm = bufs[3] if dm is bufs[2] else bufs[2]
        P = lambda x: None if x is None else x.data_ptr()
        ops._call("pd_gru_gates_bwd_zb", P(dz), 0 if dz is None else H, dout[:, t].data_ptr(), dout.stride(0), P(dm), 0 if dm is None else H,
                  rzn[:, t].data_ptr(), rzn.stride(0), hn[:, t].data_ptr(), hn.stride(0), h[:, t].data_ptr(), h.stride(0),
                  dgi[:, t].data_ptr(), dgi.stride(0), dgh[:, t].data_ptr(), dgh.stride(0), nz.data_ptr(), H, None, t, B, H,
                  nm.data_ptr(), H, dghb.data_ptr(), 3 * H, st)
        ops._call("pd_gemm_bf16", dghb.data_ptr(), 3 * H, 1, wb.data_ptr(), H, 1, nm.data_ptr(), H, None, B, H, 3 * H, 1, st)
        dz, dm = nz, nm

res = {}
for pdl in (0, 1, 0, 1):
    _lib.lib.pd_set_pdl(pdl)
    for name, fn in (("forward 32 fused steps", fwd), ("backward 32 steps (gates + GEMM)", bwd)):
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        out = h[:, T].clone() if fn is fwd else (bufs[0] + bufs[2]).clone()
        key = name
        if pdl == 0 and key not in res:
            res[key] = out
        diff = float((out - res[key]).abs().max())
        print(f"PDL={pdl} {name:36s} {e0.elapsed_time(e1) / 20 * 1e3 / T:7.2f} us/step   max |diff vs ordinary| {diff:.2e}", flush=True)
        del g
_lib.lib.pd_set_pdl(0)
