"""Kernel timeline of ONE graph-replayed training step (torch.profiler / CUPTI activity records: kernels are NOT
serialised as under ncu, so overlap between the step's streams is visible).

    python tools/trace_step.py gpurun_out/trace.json [NAME=value ...]      # ops switches as in tools/step_ab.py

Writes a compact JSON list of [name, stream, start_us, dur_us] and prints busy / idle statistics.
"""
import json, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from polydis_b200 import ops
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch

out = sys.argv[1]
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    cur = getattr(ops, k)
    setattr(ops, k, bool(int(v)) if isinstance(cur, bool) else type(cur)(float(v)))
dev = torch.device("cuda:0")
B = 512
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))
torch.manual_seed(0); random.seed(0)
m = DisentangleVAE.init_model(device=dev).to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
g = GraphedTrainStep(m, opt, B, warmup=2).capture(x, c, pr)
for _ in range(3):
    g(x, c, pr)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(x, c, pr)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = []
for e in ev:
    rows.append([e.name[:80], getattr(e, "stream", 0) if hasattr(e, "stream") else 0, e.time_range.start, e.time_range.end - e.time_range.start])
rows.sort(key=lambda r: r[2])
t0 = rows[0][2]
for r in rows:
    r[2] -= t0
json.dump(rows, open(out, "w"))
end = max(r[2] + r[3] for r in rows)
busy = sum(r[3] for r in rows)
# union of intervals = time with at least one kernel running
iv = sorted((r[2], r[2] + r[3]) for r in rows)
cov, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > cur_e:
        cov += cur_e - cur_s; cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
cov += cur_e - cur_s
print(f"{len(rows)} kernels, span {end / 1e3:.3f} ms, sum of durations {busy / 1e3:.3f} ms, >=1 kernel running {cov / 1e3:.3f} ms")
