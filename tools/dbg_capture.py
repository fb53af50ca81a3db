import os, sys, random, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200.model import DisentangleVAE
from polydis_b200.synth import synth_batch
from polydis_b200 import _lib, ops
dev = torch.device("cuda:0")
m = DisentangleVAE.init_model(device=dev).to(dev)
B = 8
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))
tfr = (0., 0., 0.)
def fwd():
    return m('train', x, c, pr, tfr1=tfr[0], tfr2=tfr[1], tfr3=tfr[2], beta=0.1, weights=(1, 0.5))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    l = fwd(); l[0].backward()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
for p in m.parameters(): p.grad = None
orig = _lib.call
bad = []
def chk(name, *a):
    if not torch.cuda.is_current_stream_capturing():
        cs = torch.cuda.current_stream()
        bad.append((name, cs.cuda_stream))
        if len(bad) < 4:
            print("NOT CAPTURING:", name, "stream", cs.cuda_stream, flush=True)
            traceback.print_stack(limit=8)
        return   # skip the launch so capture survives
    orig(name, *a)
ops._call = chk
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        l = fwd()
        print("forward captured; bad so far", len(bad), flush=True)
        l[0].backward()
    print("backward captured; bad", len(bad), flush=True)
except Exception as e:
    print("EXC", repr(e)[:300])
print("bad calls:", len(bad), bad[:5])
