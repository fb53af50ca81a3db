import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200.model import DisentangleVAE
from polydis_b200.synth import synth_batch
from polydis_b200 import _lib
dev = torch.device("cuda:0")
m = DisentangleVAE.init_model(device=dev).to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))
for tfr in [(0., 0., 0.), (0.5, 0.5, 0.5)]:
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); c0 = _lib.call_count
        opt.zero_grad(set_to_none=True)
        l = m('train', x, c, pr, tfr1=tfr[0], tfr2=tfr[1], tfr3=tfr[2], beta=0.1, weights=(1, 0.5))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        l[0].backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0); opt.step()
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"B={B} tfr={tfr}: fwd {1e3*(t1-t0):.0f} ms, bwd+opt {1e3*(t2-t1):.0f} ms, {B/(t2-t0):.0f} samples/s, lib calls {_lib.call_count-c0}", flush=True)

del l          # the last eager graph keeps AccumulateGrad nodes bound to the legacy stream alive
from polydis_b200.graphs import GraphedTrainStep
opt2 = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
g = GraphedTrainStep(m, opt2, B, tfr=(0., 0., 0.), warmup=1).capture(x, c, pr)
for _ in range(2): g(x, c, pr)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): g(x, c, pr)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"B={B} tfr=(0,0,0) CUDA graph: {1e3*dt:.1f} ms/step, {B/dt:.0f} samples/s, loss {float(g.losses[0]):.4f}", flush=True)
