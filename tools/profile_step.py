"""One warm-up + one profiled training step (and optionally one decode) between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 512] [--decode 0]
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200.model import DisentangleVAE
from polydis_b200.synth import synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--decode", type=int, default=0)
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
random.seed(0)
m = DisentangleVAE.init_model(device=dev).to(dev)
params = list(m.parameters())
opt = torch.optim.Adam(params, lr=1e-3, fused=True)
x, c, pr = (torch.from_numpy(t).to(dev) for t in synth_batch(a.batch, 0))


def step():
    opt.zero_grad(set_to_none=True)
    l = m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))
    l[0].backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)
    opt.step()


for _ in range(a.warm):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
if a.decode:
    xd, cd, pd = (torch.from_numpy(t).to(dev) for t in synth_batch(a.decode, 1))
    m.swap(pd, pd, cd, cd, True, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
