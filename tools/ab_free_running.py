"""A/B of the graph-replayed FREE-RUNNING training step (tfr = 0) under ``ops`` switches (as tools/step_ab.py)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch

dev = torch.device("cuda:0")
B = 512
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))
for spec in (sys.argv[1:] or [""]):
    saved = {}
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        saved[k] = getattr(ops, k)
        setattr(ops, k, type(saved[k])(float(v)) if not isinstance(saved[k], bool) else bool(int(v)))
    torch.manual_seed(0); random.seed(0)
    m = DisentangleVAE.init_model(device=dev).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
    g = GraphedTrainStep(m, opt, B, tfr=(0., 0., 0.), warmup=2, inject_eps=True).capture(x, c, pr)
    l0 = float(g(x, c, pr)[0])
    for _ in range(2):
        g(x, c, pr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        g(x, c, pr)
    e1.record(); torch.cuda.synchronize()
    print(f"{spec or 'default':44s} {e0.elapsed_time(e1) / 10:7.2f} ms/step  first loss {l0:.5f}", flush=True)
    for k, v in saved.items():
        setattr(ops, k, v)
    del g, m, opt
