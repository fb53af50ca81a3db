"""Time the weight-resident GRU128 kernels against the per-step GEMM+gates path (note-summary bi-GRU shape)."""
import sys, torch
sys.path.insert(0, ".")
import polydis_b200  # noqa
from polydis_b200 import ops

R, T, H = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 16, 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
gi = torch.randn(R, T, 3 * H, device=dev)
w = torch.randn(3 * H, H, device=dev) / H ** 0.5
b = torch.randn(3 * H, device=dev) * 0.1
lengths = torch.randint(1, T + 1, (R,), device=dev, dtype=torch.int32)
dout = torch.randn(R, T, H, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for res in (True, False):
    ops.RESIDENT_GRU128 = res
    for rev in (False, True):
        g = gi.clone().requires_grad_(True)
        ww, bb = w.clone().requires_grad_(True), b.clone().requires_grad_(True)

        def fwd():
            return ops.gru_sequence(g, None, None, ww, bb, lengths, rev, T)
        out = fwd()

        def fb():
            o = fwd()
            o.backward(dout)
        with torch.no_grad():
            tf = timeit(lambda: ops.gru_sequence_nograd(g, None, None, ww, bb, lengths, rev, None, T))
        tfb = timeit(fb)
        print(f"resident={res} reverse={rev}: fwd {tf:.0f} us, fwd+bwd {tfb:.0f} us")
