"""A/B timing of the graph-replayed training step (batch 512, teacher-forced) under ``ops`` switches.

    python tools/step_ab.py DEFER_WGRAD=0 DEFER_WGRAD=1 "DEFER_WGRAD=1,DEFER_MIN_ROWS=100000"

Every argument is one variant: comma-separated NAME=value assignments applied to ``polydis_b200.ops`` before the step
is captured.  Prints ms/step (CUDA events over 20 replays) and the first loss of each variant (same seed: equal).
"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch

dev = torch.device("cuda:0")
B = int(os.environ.get("PD_AB_BATCH", "512"))
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))


def run(spec):
    from polydis_b200 import _lib
    pdl = "PDL=1" in spec
    _lib.lib.pd_set_pdl(1 if pdl else 0)
    from polydis_b200 import graphs
    graphs.FOLD_CLIP_INTO_ADAM = "FOLD=0" not in spec          # FOLD=0: separate clip_grad_norm_ pass
    spec = ",".join(kv for kv in spec.split(",") if not kv.startswith("PDL=") and not kv.startswith("FOLD="))
    saved = {}
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        saved[k] = getattr(ops, k)
        setattr(ops, k, type(saved[k])(float(v)) if not isinstance(saved[k], bool) else bool(int(v)))
    torch.manual_seed(0); random.seed(0)
    m = DisentangleVAE.init_model(device=dev).to(dev)
    if os.environ.get("PD_AB_FUSED_OPT"):       # the repo's clip + Adam + LR decay on the reducer's flat buckets
        from polydis_b200.optim import FusedClipAdam
        opt = FusedClipAdam(list(m.parameters()), lr=1e-3, clip=1.0, lr_gamma=0.9999, lr_min=1e-5)
    else:
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
    g = GraphedTrainStep(m, opt, B, warmup=2, inject_eps=True).capture(x, c, pr)
    l0 = float(g(x, c, pr)[0])
    for _ in range(3):
        g(x, c, pr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        g(x, c, pr)
    e1.record(); torch.cuda.synchronize()
    print(f"{(spec or 'default') + (' PDL' if pdl else ''):50s} {e0.elapsed_time(e1) / 20:7.3f} ms/step   first loss {l0:.6f}", flush=True)
    for k, v in saved.items():
        setattr(ops, k, v)
    del g, m, opt


for spec in (sys.argv[1:] or [""]):
    run(spec)
