"""Micro-benchmark of the peer-memory gradient-exchange kernel (csrc/allreduce_p2p.cu) against ncclAllReduce.

    PD_AR_VARIANT=0 torchrun --nproc-per-node 2 tools/ar_bench.py

Per bucket size and grid: 20 exchanges captured in one CUDA graph (as the training step issues them), time per exchange
= max over ranks, with and without the in-kernel gather.  (profiles/r02_ar_bench_2gpu.txt also holds the PD_AR_VARIANT runs
of the first kernel version: sys-scope strong vs weak accesses, unroll.)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PD_AR_STREAMS", "1")
import torch
import torch.distributed as dist
from polydis_b200.ddp import BucketedGradAllReduce

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sizes_mb = [1, 4, 8, 32]
params = [torch.nn.Parameter(torch.zeros(int(mb * (1 << 20)) // 4, device=dev)) for mb in sizes_mb]
N = 20


def time_graph(fn):
    fn(); torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                fn()
    torch.cuda.synchronize(); dist.barrier()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / N * 1e3)
    t = torch.tensor([best], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


rows = []
for blocks in (16, 32, 64):
    red = BucketedGradAllReduce(params, bucket_mb=0.001, impl="p2p", ar_blocks=blocks)   # one bucket per parameter
    for b in red.buckets:
        mb = b["n"] * 4 / (1 << 20)
        src = [torch.randn(b["n"], device=dev)]
        for gather in (True, False):
            if gather:
                def fn():       # the exchange as the training step issues it: gather from the producer's buffer, then average
                    red._exchange_p2p(b["index"], b, src)
            else:
                def fn():       # exchange only: the bucket already in place
                    from polydis_b200 import _lib
                    _lib.call("pd_allreduce_p2p", red._peer_ptrs, red.rank, red.world, red.flag_bytes, b["off"], b["n"],
                              1.0 / red.world, red._epoch.data_ptr(), red._err.data_ptr(), b["index"], 1, red.ar_blocks,
                              None, None, None, 0, torch.cuda.current_stream().cuda_stream)
            us = time_graph(fn)
            rows.append((f"p2p blocks={blocks} gather={int(gather)}", mb, us))
    assert not red.peer_error()
    red.remove()
for mb in sizes_mb:
    buf = torch.randn(int(mb * (1 << 20)) // 4, device=dev)
    us = time_graph(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG))
    rows.append(("nccl all_reduce(AVG)", float(mb), us))
if rank == 0:
    print(f"world={world}")
    for name, mb, us in rows:
        # bus: bytes one rank RECEIVES as read responses (= what it also receives as peer stores) per second
        print(f"  {name:34s} {mb:6.1f} MB  {us:8.1f} us   bus {mb * 1.048576 * (world - 1) / world / us * 1e3:7.1f} GB/s/dir")
dist.destroy_process_group()
