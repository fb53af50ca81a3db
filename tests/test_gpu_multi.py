"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with ``gpurun --gpus 2``): the data-parallel exchange on hardware.

Two NCCL ranks, a different shard each.  Checked on every rank:
  1. eager step through ``BucketedGradAllReduce``: every gradient == mean over ranks of the CPU ORACLE's per-shard
     gradients (the reference's DataParallel semantics, amc_dl/torch_plus/module.py:152-157), <= 1e-2 relative;
  2. the same exchange captured in the training step's CUDA graph (side-stream backward, event-ordered buckets):
     gradients equal the eager ones, post-step parameters are identical on both ranks and equal a single-process
     step on the rank-averaged gradients.
"""
import os
import random
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_kernel(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from polydis_b200.ddp import BucketedGradAllReduce
        # the exchange kernel alone: odd-sized parameters, sources that are only 4-byte aligned, several rounds on the
        # same flags (epochs), exact expected values; norm partials vs torch
        sizes = (5, 1024, 3, 77777, 1, 4096 * 33 + 2, 130, 2000003)
        ps = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in sizes]
        redk = BucketedGradAllReduce(ps, bucket_mb=0.3, impl="p2p")
        assert len(redk.buckets) >= 3
        for rnd in range(3):
            redk.reset()
            keep = []
            for i, p_ in enumerate(ps):
                base = torch.arange(p_.numel() + 1, device=dev, dtype=torch.float32) + (2 * rank + rnd + i)
                keep.append(base)
                p_.grad = base[1:] if i % 2 else base[1:].clone()      # odd ones: data_ptr % 16 == 4
            torch.cuda.synchronize()            # (the hooks' events order this in training; here the gradients are set by hand)
            for b in redk.buckets:
                b["pending"] = 0
                redk._launch(b)
            redk.finish()
            total = redk.clip_grad_norm_(1e30)
            torch.cuda.synchronize()
            assert not redk.peer_error()
            ref_sq = 0.0
            for i, p_ in enumerate(ps):
                # mean over ranks of (k + 2 rank + rnd + i): every partial sum is an integer below 2^24 and world is a
                # power of two, so the fp32 result is exact
                want = (torch.arange(p_.numel() + 1, device=dev, dtype=torch.float64)[1:] + (world - 1) + (rnd + i)).float()
                assert torch.equal(p_.grad, want), (rnd, i, float((p_.grad - want).abs().max()))
                ref_sq += float(want.double().pow(2).sum())
            assert abs(float(total) - ref_sq ** 0.5) <= 1e-5 * ref_sq ** 0.5, (float(total), ref_sq ** 0.5)
        redk.remove()

        q.put((rank, "ok"))
        torch.cuda.synchronize()
        dist.barrier()
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, "error: " + repr(ex) + "\n" + traceback.format_exc()))
    finally:
        os._exit(0)


def test_p2p_exchange_kernel_exact():
    """csrc/allreduce_p2p.cu alone on all GPUs of the box (2, 4 or 8 ranks): exact averages, in-kernel gather from unaligned sources, epochs over several
    rounds, several buckets in flight on different streams, norm partials."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = torch.cuda.device_count()
    world = 8 if n >= 8 else 4 if n >= 4 else 2            # (powers of two: the expected averages are exact)
    procs = [ctx.Process(target=_worker_kernel, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status in res:
        assert status == "ok", status


def _worker(rank, world, port, q, impl="nccl"):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from oracle import polydis_oracle as O
        from polydis_b200.ddp import BucketedGradAllReduce
        from polydis_b200.graphs import GraphedTrainStep
        from polydis_b200.model import DisentangleVAE
        from polydis_b200.synth import synth_batch
        from polydis_b200.weights import make_state_dict

        B = 4
        shards = [[torch.from_numpy(a) for a in synth_batch(B, 60 + r)] for r in range(world)]
        torch.manual_seed(123)
        eps = [(torch.randn(B, 256), torch.randn(B, 256)) for _ in range(world)]
        # oracle: per-shard gradients on the CPU, averaged over the shards
        sd = {k: v.requires_grad_(True) for k, v in make_state_dict(3).items()}
        for r in range(world):
            random.seed(0)
            O.loss(sd, *shards[r], O.draw_plan(1., 1., 1.), *eps[r])[0].backward()
        expect = {k: v.grad / world for k, v in sd.items()}

        def fresh():
            m = DisentangleVAE.init_model(device=dev)
            m.load_state_dict(make_state_dict(3))
            return m.to(dev).train()

        x, c, pr = (t.to(dev) for t in shards[rank])
        e = tuple(t.to(dev) for t in eps[rank])
        # 1. eager, bucketed all-reduce
        m = fresh()
        red = BucketedGradAllReduce(list(m.parameters()), bucket_mb=8, impl=impl)
        for _ in range(2):                               # twice: reset() must re-arm everything
            red.reset()
            random.seed(0)
            m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5), eps=e)[0].backward()
            red.finish()
        torch.cuda.synchronize()
        worst = 0.0
        eager_grads = {}
        for name, p in m.named_parameters():
            g, rf = p.grad.detach().cpu().double(), expect[name].double()
            worst = max(worst, float((g - rf).norm() / (rf.norm() + 1e-20)))
            eager_grads[name] = p.grad.detach().clone()
        red.remove()
        # 2. the same exchange inside the step's CUDA graph + Adam; reference = single-process Adam on the averaged grads
        m2 = fresh()
        params2 = list(m2.parameters())
        red2 = BucketedGradAllReduce(params2, bucket_mb=8, impl=impl)
        opt2 = torch.optim.Adam(params2, lr=1e-3, fused=True, capturable=True)
        random.seed(0)
        g = GraphedTrainStep(m2, opt2, B, reducer=red2, warmup=11, inject_eps=True, clip=1.0).capture(x, c, pr)
        g.eps[0].copy_(e[0]); g.eps[1].copy_(e[1])
        g(x, c, pr)
        torch.cuda.synchronize()
        # reference for the captured step: the eager rank-averaged gradients, clipped to norm 1, then one Adam step
        m3 = fresh()
        params3 = list(m3.parameters())
        for (n, p) in m3.named_parameters():
            p.grad = eager_grads[n].clone()
        torch.nn.utils.clip_grad_norm_(params3, 1.0, foreach=True)
        # after the replay p.grad holds the all-reduced gradients AFTER the in-graph clip_grad_norm_(1)
        graph_grad_err = max(float((p2.grad - p3.grad).norm() / (p3.grad.norm() + 1e-20))
                             for p2, p3 in zip(params2, params3))
        # the update itself: one Adam step on the gradients the replay left behind (all-reduced, clipped) from fresh weights
        # must land exactly on the replayed parameters.  (Stepping on the EAGER gradients instead would test noise: two
        # runs of the same backward differ by ~5e-4 relative in the encoder gradients -- atomics / split-K order through
        # the 32-step recurrences -- and Adam turns a sign flip of a near-zero element into a 2 x lr difference.)
        for p2, p3 in zip(params2, params3):
            p3.grad = p2.grad.detach().clone()
        torch.optim.Adam(params3, lr=1e-3, fused=True, capturable=True).step()
        torch.cuda.synchronize()
        step_err = max(float((a - b).abs().max()) for a, b in zip(params2, params3))
        # identical parameters on every rank
        flat = torch.cat([p.detach().reshape(-1) for p in params2])
        other = flat.clone()
        dist.broadcast(other, 0)
        rank_diff = float((flat - other).abs().max())
        q.put((rank, "ok", worst, graph_grad_err, step_err, rank_diff))
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, "error: " + repr(ex) + "\n" + traceback.format_exc(), 0, 0, 0, 0))
    finally:
        # process groups whose collectives were captured in CUDA graphs can block in destroy_process_group()
        os._exit(0)


@pytest.mark.parametrize("impl", ["nccl", "p2p"])
def test_two_rank_nccl_gradients_match_per_shard_oracle_mean(impl):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, impl)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    print(f"2-rank {impl} (rank, status, oracle err, graph-vs-eager err, step err, rank diff):", res)
    for rank, status, worst, graph_err, step_err, rank_diff in res:
        assert status == "ok", status
        assert worst <= 1e-2, (rank, worst)                  # N-rank gradients == mean of per-shard oracle gradients
        assert graph_err <= 2e-3, (rank, graph_err)          # captured exchange == eager exchange up to run-to-run noise
        assert step_err <= 1e-6, (rank, step_err)            # Adam on the exchanged, clipped gradients
        assert rank_diff == 0.0, (rank, rank_diff)           # replicas stay bit-identical
