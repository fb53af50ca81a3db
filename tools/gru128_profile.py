"""One launch each of the resident GRU128 forward / backward kernels and the duration-decoder kernels, for ncu."""
import sys, torch
sys.path.insert(0, ".")
import polydis_b200  # noqa
from polydis_b200 import ops

R, T, H = 16384, 16, 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
gi = torch.randn(R, T, 3 * H, device=dev, requires_grad=True)
w = (torch.randn(3 * H, H, device=dev) / H ** 0.5).requires_grad_(True)
b = (torch.randn(3 * H, device=dev) * 0.1).requires_grad_(True)
lengths = torch.randint(1, T + 1, (R,), device=dev, dtype=torch.int32)
ops.RESIDENT_GRU128_MAX_ROWS_TF32 = 1 << 30
for it in range(2):
    if it == 1:
        torch.cuda.cudart().cudaProfilerStart()
    out = ops.gru_sequence(gi, None, None, w, b, lengths, False, T)
    out.backward(torch.ones_like(out))
    Q = 245760
    h0 = torch.randn(Q, 64, device=dev, requires_grad=True)
    par = [torch.randn(192, 5, device=dev) * .3, torch.randn(192, device=dev) * .1, torch.randn(192, 64, device=dev) * .12,
           torch.randn(192, device=dev) * .1, torch.randn(5, device=dev), torch.randn(2, 64, device=dev) * .2, torch.randn(2, device=dev) * .1]
    par = [p.requires_grad_(True) for p in par]
    lg = ops.dur_decode(h0, *par)
    lg.backward(torch.ones_like(lg))
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
