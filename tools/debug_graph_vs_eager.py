import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
from polydis_b200.ddp import BucketedGradAllReduce
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "4"))
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 60))
torch.manual_seed(123)
e = (torch.randn(B, 256, device=dev), torch.randn(B, 256, device=dev))

def fresh():
    m = DisentangleVAE.init_model(device=dev)
    m.load_state_dict(make_state_dict(3))
    return m.to(dev).train()

def eager(use_red):
    m = fresh()
    red = BucketedGradAllReduce(list(m.parameters()), bucket_mb=8) if use_red else None
    for _ in range(2):
        if red: red.reset()
        else:
            for p in m.parameters(): p.grad = None
        random.seed(0)
        m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5), eps=e)[0].backward()
        if red: red.finish()
    torch.cuda.synchronize()
    g = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    if red: red.remove()
    return g

def graphed(use_red):
    m = fresh()
    params = list(m.parameters())
    red = BucketedGradAllReduce(params, bucket_mb=8) if use_red else None
    opt = torch.optim.Adam(params, lr=0.0, fused=True, capturable=True)
    random.seed(0)
    g = GraphedTrainStep(m, opt, B, reducer=red, warmup=3, inject_eps=True, clip=0).capture(x, c, pr)
    g.eps[0].copy_(e[0]); g.eps[1].copy_(e[1])
    g(x, c, pr)
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in m.named_parameters()}

def cmp(a, b, tag):
    worst = sorted(((float((a[n] - b[n]).norm() / (b[n].norm() + 1e-20)), n) for n in a), reverse=True)[:3]
    print(f"{tag:44s}", " | ".join(f"{n} {v:.2e}" for v, n in worst), flush=True)

for spec in sys.argv[1:] or [""]:
    saved = {}
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        saved[k] = getattr(ops, k); setattr(ops, k, bool(int(v)))
    e1 = eager(False); e2 = eager(False)
    cmp(e1, e2, f"[{spec}] eager vs eager")
    cmp(graphed(False), e1, f"[{spec}] graph vs eager")
    cmp(eager(True), e1, f"[{spec}] eager+reducer vs eager")
    cmp(graphed(True), e1, f"[{spec}] graph+reducer vs eager")
    for k, v in saved.items(): setattr(ops, k, v)
