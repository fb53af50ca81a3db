"""One warm-up + one profiled greedy decode (model.swap, default token-parity precision) between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_decode.csv python tools/profile_decode.py [--batch 16384]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200.model import DisentangleVAE
from polydis_b200.synth import synth_batch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = DisentangleVAE.init_model(device=dev).to(dev).eval()
z1, z2 = torch.randn(a.batch, 256, device=dev), torch.randn(a.batch, 256, device=dev)
with torch.no_grad():
    m.decode_tokens(z1[:1024], z2[:1024])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m.decode_tokens(z1, z2)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
