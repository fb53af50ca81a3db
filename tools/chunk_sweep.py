"""Graph-replayed training step time vs the row-chunking parameters of the big recurrences (ops._over_row_chunks)."""
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


def timed(name):
    torch.manual_seed(0); random.seed(0)
    m = DisentangleVAE.init_model(device=dev).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
    g = GraphedTrainStep(m, opt, B, warmup=2).capture(x, c, pr)
    for _ in range(3): g(x, c, pr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): g(x, c, pr)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1)/10:7.2f} ms/step", flush=True)
    del g, m, opt


if __name__ == "__main__":
    for mb, lanes in [(0, 1), (12, 3), (25, 2), (25, 3), (50, 2), (6, 4), (12, 2), (25, 1), (50, 1)]:
        ops.ROW_CHUNK_BYTES = mb << 20
        ops.ROW_CHUNK_LANES = lanes
        timed(f"chunk {mb} MB x {lanes} lanes")
