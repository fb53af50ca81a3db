"""Kernel summary (CUPTI) of ONE graph-replayed FREE-RUNNING training step (tfr = 0: no-grad greedy pass + batched phases).

    python tools/trace_free_running.py
"""
import collections, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch

dev = torch.device("cuda:0")
B = 512
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))
torch.manual_seed(0); random.seed(0)
m = DisentangleVAE.init_model(device=dev).to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
g = GraphedTrainStep(m, opt, B, tfr=(0., 0., 0.), warmup=2).capture(x, c, pr)
for _ in range(3):
    g(x, c, pr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g(x, c, pr)
e1.record(); torch.cuda.synchronize()
print(f"free-running step: {e0.elapsed_time(e1) / 5:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(x, c, pr)
    torch.cuda.synchronize()
rows = [(e.name, e.time_range.start, e.time_range.end - e.time_range.start) for e in prof.events()
        if e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: r[1])
t0 = rows[0][1]
end = max(s + d for _, s, d in rows) - t0
agg = collections.defaultdict(lambda: [0, 0.0])
for n, s, d in rows:
    k = n.replace("void ", "").replace("(anonymous namespace)::", "").split("(")[0][:70]
    agg[k][0] += 1; agg[k][1] += d
print(f"{len(rows)} kernels, span {end / 1e3:.2f} ms, sum of durations {sum(v[1] for v in agg.values()) / 1e3:.2f} ms")
for k, (cnt, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{d / 1e3:8.3f} ms {cnt:5d}  {d / cnt:7.1f} us each  {k}")
# idle gaps: time with no kernel running
iv = sorted((s - t0, s - t0 + d) for _, s, d in rows)
cov, cs, ce = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > ce:
        cov += ce - cs; cs, ce = s, e
    else:
        ce = max(ce, e)
cov += ce - cs
print(f">= 1 kernel running: {cov / 1e3:.2f} ms of {end / 1e3:.2f} ms")
