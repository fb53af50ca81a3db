import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict
dev = torch.device("cuda:0")
B = 512
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 4243))
for spec in sys.argv[1:] or [""]:
    saved = {}
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("="); saved[k] = getattr(ops, k); setattr(ops, k, bool(int(v)))
    m = DisentangleVAE.init_model(device=dev); m.load_state_dict(make_state_dict(31)); m.to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=0.0, fused=True, capturable=True)
    vals = []
    for i in range(6):
        opt.zero_grad(set_to_none=True)
        l = m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))
        l[0].backward()
        torch.cuda.synchronize()
        vals.append((float(l[5]), float(l[6]), float(l[7])))
    print(f"[{spec}] eager  kl_chd/kl_rhy/chord:", " | ".join(f"{a:.6e} {b:.6e} {c_:.5f}" for a, b, c_ in vals), flush=True)
    g = GraphedTrainStep(m, opt, B, warmup=1).capture(x, c, pr)
    vals = []
    for i in range(6):
        l = g(x, c, pr).clone(); torch.cuda.synchronize()
        vals.append((float(l[5]), float(l[6]), float(l[7])))
    print(f"[{spec}] graph  kl_chd/kl_rhy/chord:", " | ".join(f"{a:.6e} {b:.6e} {c_:.5f}" for a, b, c_ in vals), flush=True)
    for k, v in saved.items(): setattr(ops, k, v)
