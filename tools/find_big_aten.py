"""Which torch (aten) ops of the EAGER training step launch the largest kernels, with shapes and the python frame that
issued them (torch.profiler with_stack): finds stray layout copies / fills the hand-written kernels should absorb.

    python tools/find_big_aten.py [min_us]
"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from polydis_b200.model import DisentangleVAE
from polydis_b200.synth import synth_batch

min_us = float(sys.argv[1]) if len(sys.argv) > 1 else 15.0
dev = torch.device("cuda:0")
B = 512
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))
torch.manual_seed(0); random.seed(0)
m = DisentangleVAE.init_model(device=dev).to(dev)


def step():
    for p in m.parameters():
        p.grad = None
    m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))[0].backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::") and e.device_time_total >= min_us:
        # leaf-ish aten ops only: skip wrappers whose children carry the same time
        kids = [k for k in e.cpu_children if k.name.startswith("aten::") and k.device_time_total >= 0.9 * e.device_time_total]
        if kids:
            continue
        stack = [s for s in (e.stack or []) if "polyphonic" in s or "polydis" in s]
        rows.append((e.device_time_total, e.name, str(e.input_shapes)[:120], stack[:3]))
rows.sort(key=lambda r: -r[0])
tot = sum(r[0] for r in rows)
print(f"aten ops with >= {min_us} us of device time in one eager step: {len(rows)}, {tot / 1e3:.3f} ms in total")
for t, n, sh, st in rows[:40]:
    print(f"{t:8.1f} us  {n:28s} {sh}")
    for s in st:
        print(f"              {s[-150:]}")
