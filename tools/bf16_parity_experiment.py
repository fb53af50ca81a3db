"""Would bf16 GEMM operands (fp32 accumulate) meet the gradient tolerance?  Emulate by rounding every GEMM
operand to bf16 before the TF32 kernel (bf16 values are exact in TF32) and compare all 81 gradients with the
CPU oracle."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
from polydis_b200.model import DisentangleVAE
from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict
from oracle import polydis_oracle as O
dev = torch.device("cuda:0")
B = 6
xs, cs, prs = (torch.from_numpy(a) for a in synth_batch(B, 21))
sd = {k: v.requires_grad_(True) for k, v in make_state_dict(9).items()}
torch.manual_seed(3); e1, e2 = torch.randn(B, 256), torch.randn(B, 256)
random.seed(11); plan = O.draw_plan(1., 1., 1.)
ref = O.loss(sd, xs, cs, prs, plan, e1, e2); ref[0].backward()
orig = ops._gemm
def rounded(a, sam, sak, b, sbk, sbn, out, bias, M, N, K, acc):
    if K == 0:
        return orig(a, sam, sak, b, sbk, sbn, out, bias, M, N, K, acc)
    a2 = a.to(torch.bfloat16).to(torch.float32).contiguous()
    b2 = b.to(torch.bfloat16).to(torch.float32).contiguous()
    if sak == 1: sam2, sak2 = a2.shape[1], 1
    else: sam2, sak2 = 1, a2.shape[1]
    if sbk == 1: sbk2, sbn2 = 1, b2.shape[1]
    else: sbk2, sbn2 = b2.shape[1], 1
    return orig(a2, sam2, sak2, b2, sbk2, sbn2, out, bias, M, N, K, acc)
for mode in ("tf32", "bf16-emulated"):
    ops._gemm = rounded if mode != "tf32" else orig
    m = DisentangleVAE.init_model(device=dev); m.load_state_dict(make_state_dict(9)); m.to(dev).train()
    random.seed(11)
    got = m.loss(xs.to(dev), cs.to(dev), prs.to(dev), 1., 1., 1., eps=(e1.to(dev), e2.to(dev)))
    got[0].backward()
    lerr = max(abs(float(a) - float(b)) / (abs(float(b)) + 1e-9) for a, b in zip(got, ref))
    errs = sorted(((float((p.grad.cpu().double() - sd[n].grad.double()).norm() / (sd[n].grad.double().norm() + 1e-20)), n)
                   for n, p in m.named_parameters()), reverse=True)
    print(f"{mode}: max loss rel err {lerr:.2e}; worst grads: " + ", ".join(f"{n.split('.',1)[1][:28]} {e:.2e}" for e, n in errs[:5]), flush=True)
