"""world_size-2 gloo test (CPU) of the bucketed gradient all-reduce used for multi-GPU training:
after backward + finish(), every rank holds the MEAN over ranks of the per-shard gradients
(the reference's DataParallel loss-averaging semantics, amc_dl/torch_plus/module.py:152-157)."""
import os
import random
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from polydis_b200 import _lib, ops
    from polydis_b200.ddp import BucketedGradAllReduce
    from polydis_b200.model import DisentangleVAE
    from polydis_b200.synth import synth_batch
    from polydis_b200.weights import make_state_dict
    from tests.cpu_backend import CpuBackend
    be = CpuBackend()                                   # numpy emulation of the C-ABI (no GPU here)
    _lib.call = ops._call = be.call
    ops._stream = lambda: None
    ops._chk = lambda t, name="tensor": t

    m = DisentangleVAE.init_model(device=torch.device("cpu"))
    m.load_state_dict(make_state_dict(3))
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(2, 40 + rank))       # a different shard per rank
    torch.manual_seed(5 + rank)
    eps = (torch.randn(2, 256), torch.randn(2, 256))

    def backward():
        random.seed(0)
        m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5), eps=eps)[0].backward()

    backward()                                          # plain local gradients
    local = [p.grad.clone() for p in m.parameters()]
    expect = []
    for g in local:
        g = g.clone()
        dist.all_reduce(g)
        expect.append(g / world)
    for p in m.parameters():
        p.grad = None
    red = BucketedGradAllReduce(list(m.parameters()), bucket_mb=8)
    assert len(red.buckets) >= 4
    for _ in range(2):                                  # twice: reset() must re-arm the hooks
        red.reset()
        backward()
        red.finish()
    err = max(float((p.grad - e).abs().max() / (e.abs().max() + 1e-12)) for p, e in zip(m.parameters(), expect))
    differs = max(float((l - e).abs().max()) for l, e in zip(local, expect))
    # the reducer's own clip == torch's clip_grad_norm_ on the same (averaged) gradients
    tn = torch.sqrt(sum((e.double() ** 2).sum() for e in expect))
    got = red.clip_grad_norm_(0.5)
    coef = min(1.0, 0.5 / (float(tn) + 1e-6))
    e_norm = abs(float(got) - float(tn)) / float(tn)
    e_clip = max(float((p.grad - e * coef).abs().max() / (e.abs().max() + 1e-12)) for p, e in zip(m.parameters(), expect))
    # (torch's float32 CPU norm over 2 M-element buckets is only good to ~1e-4; the CUDA reductions are tree-shaped)
    assert e_norm < 5e-4 and e_clip < 5e-4, (e_norm, e_clip)
    red.remove()
    # buckets in observed gradient-ready order: a permutation of the parameters, same averaged gradients
    from polydis_b200.ddp import observe_ready_order
    order = observe_ready_order(list(m.parameters()), backward)
    assert sorted(map(id, order)) == sorted(map(id, m.parameters())) and all(p.grad is None for p in m.parameters())
    red2 = BucketedGradAllReduce(list(m.parameters()), bucket_mb=8, ready_order=order)
    red2.reset()
    backward()
    red2.finish()
    err = max(err, max(float((p.grad - e).abs().max() / (e.abs().max() + 1e-12)) for p, e in zip(m.parameters(), expect)))
    q.put((rank, err, differs))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=400) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, differs in res:
        assert err < 1e-5, (rank, err)
        assert differs > 1e-6            # the shards really had different gradients
