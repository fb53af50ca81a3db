"""FusedClipAdam (clip + Adam + MinExponentialLR in two kernels per bucket) against
torch.nn.utils.clip_grad_norm_ + torch.optim.Adam + the reference's LR rule, on the numpy ABI emulation."""
import torch

from tests import cpu_backend


def test_fused_clip_adam_matches_torch(monkeypatch):
    cpu_backend.install(monkeypatch)
    from polydis_b200.optim import FusedClipAdam
    torch.manual_seed(0)
    shapes = [(7, 5), (33,), (4, 3, 2), (130, 17)]
    a = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    gamma, lr_min, lr0 = 0.9, 5e-4, 1e-3
    ref = torch.optim.Adam(b, lr=lr0)
    opt = FusedClipAdam(a, lr=lr0, clip=1.0, lr_gamma=gamma, lr_min=lr_min, bucket_mb=0.002)
    assert len(opt.reducer.buckets) >= 2
    for it in range(6):
        gs = [torch.randn(s) * (3.0 if it % 2 == 0 else 0.01) for s in shapes]     # clipped and unclipped steps
        opt.zero_grad()
        ref.zero_grad()
        for p, q, g in zip(a, b, gs):
            (p * g).sum().backward()
            (q * g).sum().backward()
        norm = torch.nn.utils.clip_grad_norm_(b, 1.0)
        for grp in ref.param_groups:
            grp["lr"] = max(lr0 * gamma ** it, lr_min)          # MinExponentialLR after `it` scheduler steps
        ref.step()
        opt.step()
        assert abs(float(opt.grad_norm()) - float(norm)) < 1e-4 * float(norm)
        for p, q in zip(a, b):
            assert torch.allclose(p, q, atol=2e-7, rtol=1e-5), it
