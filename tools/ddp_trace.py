"""Gradient-exchange cost of the data-parallel step: A/B timings of the graph-replayed step under reducer settings, and a
CUPTI kernel timeline of rank 0 (where the exchange kernels sit relative to backward's tail, the clip and Adam).

    torchrun --nproc-per-node 2 tools/ddp_trace.py [--trace out.json] "bucket_mb=8" "bucket_mb=8,ready=1" "bucket_mb=32,prio=-1"
    python tools/ddp_trace.py "reducer=0" "bucket_mb=8"            # world 1: cost of the reducer's gather path alone

Variant keys: reducer (0: plain p.grad, world 1 only), bucket_mb, ready (bucket order = observed gradient-ready order),
prio (communication stream priority), impl ("nccl" | "p2p": the repo's peer-memory all-reduce kernel), streams / blocks
(p2p: exchange streams, CTAs per exchange).
"""
import json, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from polydis_b200 import ops
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch
from polydis_b200 import ddp

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = int(os.environ.get("PD_AB_BATCH", "512"))
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 100 + rank))
args = sys.argv[1:]
trace_out = None
if args and args[0] == "--trace":
    trace_out = args[1]; args = args[2:]


def build(spec):
    kv = dict(a.split("=") for a in filter(None, spec.split(",")))
    torch.manual_seed(0); random.seed(0)
    m = DisentangleVAE.init_model(device=dev).to(dev)
    params = list(m.parameters())
    if world > 1:
        for p in params:
            dist.broadcast(p.data, 0)
    reducer = None
    if int(kv.get("reducer", "1")):
        order = None
        if int(kv.get("ready", "0")):
            def fb():
                m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))[0].backward()
                torch.cuda.synchronize()
            order = ddp.observe_ready_order(params, fb)
        kw = {}
        if "impl" in kv:
            kw["impl"] = kv["impl"]
        if "streams" in kv:
            os.environ["PD_AR_STREAMS"] = kv["streams"]
        if "blocks" in kv:
            kw["ar_blocks"] = int(kv["blocks"])
        reducer = ddp.BucketedGradAllReduce(params, bucket_mb=float(kv.get("bucket_mb", "8")), ready_order=order,
                                            comm_priority=int(kv.get("prio", "0")), **kw)
    opt = torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True)
    g = GraphedTrainStep(m, opt, B, reducer=reducer, warmup=3 if world == 1 else 11).capture(x, c, pr)
    return m, opt, g, reducer


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(g, n=20):
    for _ in range(3):
        g(x, c, pr)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g(x, c, pr)
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def trace(g, out):
    from torch.profiler import profile, ProfilerActivity
    barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g(x, c, pr)
        torch.cuda.synchronize()
    barrier()
    if rank != 0:
        return
    rows = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            rows.append([e.name[:70], e.time_range.start, e.time_range.end - e.time_range.start])
    rows.sort(key=lambda r: r[1])
    t0 = rows[0][1]
    for r in rows:
        r[1] -= t0
    json.dump(rows, open(out, "w"))
    end = max(r[1] + r[2] for r in rows)
    is_x = lambda n: ("nccl" in n.lower()) or ("allreduce_p2p" in n)
    xs = [r for r in rows if is_x(r[0])]
    other = [r for r in rows if not is_x(r[0])]
    print(f"trace: {len(rows)} kernels, span {end / 1e3:.3f} ms; exchange kernels {len(xs)}, "
          f"sum {sum(r[2] for r in xs) / 1e3:.3f} ms", flush=True)
    for r in xs:
        print(f"   exchange  start {r[1] / 1e3:7.3f} ms  dur {r[2]:8.1f} us  {r[0][:50]}")
    # the tail: everything that starts after the first exchange kernel of the last quarter
    t_tail = end - 1500.0
    print("tail (last 1.5 ms):")
    for r in rows:
        if r[1] + r[2] >= t_tail:
            print(f"   {r[1] / 1e3:7.3f} ms +{r[2]:7.1f} us  {r[0][:60]}")


for spec in (args or [""]):
    m, opt, g, red = build(spec)
    t = timed(g)
    if rank == 0:
        nb = len(red.buckets) if red is not None else 0
        print(f"world {world}  {spec or 'default':40s} {t:7.3f} ms/step  ({nb} buckets)", flush=True)
    if trace_out is not None:
        trace(g, trace_out.replace(".json", f"_{(spec or 'default').replace(',', '_').replace('=', '')}.json"))
    del g, m, opt, red
if world > 1:
    dist.destroy_process_group()
