#!/usr/bin/env python
"""PolyDis hot-path benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

Headline metric (BASELINE.json): PolyDis **training samples/s** -- one step = zero_grad + forward +
loss + backward + clip_grad_norm_(1) + Adam on a synthetic teacher-forced batch of 512 2-bar segments
per GPU (config "PolyDisVAE training, batch=512 on 1 B200, teacher-forced PianoTree decoder"), with
greedy-decode segments/s reported beside it in the same JSON line (`decode`).  `value` is measured with
the batch resident in HBM; `e2e` runs the same step through the public model API from pinned HOST
buffers (H2D of x / c / pr_mat and D2H of the loss inside the timed region).  Under torchrun each rank
runs the same per-GPU batch (weak scaling) with bucketed NCCL gradient all-reduce overlapped with
backward (polydis_b200.ddp, captured inside the step's CUDA graph); timing is CUDA events, max over ranks.

`--impl reference` times the reference's CPU path: the oracle port (oracle/polydis_oracle.py, same op
granularity as the reference, pinned to it by golden vectors) on all host cores, bounded batches.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

EXCHANGE_NOTE = {
    "p2p": "polydis_b200 allreduce_p2p_kernel: one kernel per bucket gathers the gradients, two-shot all-reduce with peer "
           "loads / stores over NVLink (CUDA IPC symmetric region), clip-norm partials; captured in the step graph",
    "nccl": "multi-tensor gather + ncclAllReduce(AVG) per bucket on a side stream, captured in the step graph",
}

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_GFLOP_PER_SAMPLE = 5.45      # algorithmic minimum fwd+bwd, SURVEY.md 8(d)
DECODE_GFLOP_PER_SEGMENT = 1.92
METRIC = "polydis_train_samples_per_sec"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")   # ncu --set full dram bytes per launch


def _measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of ``kernel`` from the committed ncu capture of the
    variant this bench times (profiles/r02_kernel_traffic.json, written from the .ncu-rep by tools/ncu_traffic.py)."""
    try:
        return json.load(open(TRAFFIC_FILE))[kernel]["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1389.3), d.get("hbm_gbs", 6454.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v == "Active":
                        reasons.add(name)
        mx = float(self.rows[0][2]) if self.rows and len(self.rows[0]) > 2 else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_run(batch, steps, warmup, decode_batch):
    """Oracle port on host cores: train step (fwd+bwd+clip+Adam, tfr=1) and greedy decode."""
    import torch
    from oracle import polydis_oracle as O
    from polydis_b200.synth import synth_batch
    from polydis_b200.weights import make_state_dict
    torch.set_num_threads(os.cpu_count())
    sd = {k: v.requires_grad_(True) for k, v in make_state_dict(0).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-3)
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(batch, 0))
    plan = O.draw_plan(1., 1., 1.)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        e1, e2 = torch.randn(batch, 256), torch.randn(batch, 256)
        loss = O.loss(sd, x, c, pr, plan, e1, e2)[0]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), 1.0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    train_sps = batch * len(times) / sum(times)
    sdd = {k: v.detach() for k, v in sd.items()}
    xd, cd, prd = (torch.from_numpy(a) for a in synth_batch(decode_batch, 1))
    t0 = time.perf_counter()
    O.inference(sdd, prd, cd)
    dec = decode_batch / (time.perf_counter() - t0)
    return train_sps, dec, sum(times) / len(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import torch
    batch = 128
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    sps, dec, ms = cpu_reference_run(batch, steps, warmup, 64)
    cores = torch.get_num_threads()
    out = {"metric": METRIC, "value": sps, "unit": "samples/s", "impl": "reference", "n_gpus": args.gpus,
           "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "PolyDisVAE train step (zero_grad+fwd+loss+bwd+clip+Adam), teacher-forced, "
                                  "CPU oracle port of the reference path", "batch_per_step": batch},
           "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                            "sample": f"{steps} steps of batch {batch} after {warmup} warm-up; os.cpu_count()={os.cpu_count()}"},
           "decode": {"value": dec, "unit": "segments/s", "sample": "one greedy inference call, batch 64"},
           "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from polydis_b200 import _lib
    from polydis_b200.model import DisentangleVAE
    from polydis_b200.synth import synth_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    verbose = bool(os.environ.get("PD_BENCH_VERBOSE"))

    def mark(msg):
        if verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(1234)
    random.seed(1234 + rank)
    model = DisentangleVAE.init_model(device=dev).to(dev)
    net = model
    params = [p for p in model.parameters()]
    reducer = None
    if world > 1:
        from polydis_b200.ddp import BucketedGradAllReduce
        for p in params:                                   # identical initial weights on every rank
            dist.broadcast(p.data, 0)
        exchange = args.exchange
        # buckets in the order the gradients become ready (one observed backward; rank 0's order for everybody)
        from polydis_b200.ddp import observe_ready_order
        xo, co, po = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 7))     # the real batch size: which ops defer depends on it

        def _one_backward():
            model('train', xo, co, po, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))[0].backward()
            torch.cuda.synchronize()
        seen = observe_ready_order(params, _one_backward)
        idx = [[next(i for i, q in enumerate(params) if q is p) for p in seen]]
        dist.broadcast_object_list(idx, 0)
        order = [params[i] for i in idx[0]]
        if exchange == "p2p":           # the repo's peer-memory kernel; NCCL only if the node offers no CUDA IPC / P2P
            try:
                reducer = BucketedGradAllReduce(params, bucket_mb=args.bucket_mb, impl="p2p", ready_order=order)
            except RuntimeError as ex:
                if rank == 0:
                    print(f"bench: p2p exchange unavailable, using NCCL: {ex}", file=sys.stderr, flush=True)
                exchange = "nccl"
        if exchange == "nccl":
            reducer = BucketedGradAllReduce(params, bucket_mb=args.bucket_mb, ready_order=order)
    fused_opt = args.fused_optim
    if fused_opt:
        from polydis_b200.optim import FusedClipAdam
    if fused_opt:       # clip_grad_norm_(1) + Adam(lr 1e-3) + MinExponentialLR(0.9999, 1e-5): train.py:18-26,50-51
        opt = FusedClipAdam(params, lr=1e-3, clip=1.0, lr_gamma=0.9999, lr_min=1e-5, reducer=reducer)
        reducer = opt.reducer
    else:
        opt = torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True)
    xh, ch, ph = (torch.from_numpy(a).pin_memory() for a in synth_batch(B, 100 + rank))
    x, c, pr = xh.to(dev), ch.to(dev), ph.to(dev)

    def step(x, c, pr):
        if reducer is not None:
            reducer.reset()
        else:
            opt.zero_grad(set_to_none=True)
        losses = net('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))
        losses[0].backward()
        if reducer is not None:
            reducer.finish()
        if not fused_opt:
            if getattr(reducer, "impl", None) == "p2p":
                reducer.clip_grad_norm_(1.0)
            else:
                torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)
        opt.step()
        return losses[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    def timed_graph(fn, n):
        """ms per replay of a CUDA graph of fn() -- how the step itself issues its kernels (eager launches of the TMA kernels
        are host-bound: every launch encodes its tensor maps on the CPU)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        gr.replay()
        ms = timed(gr.replay, n)
        del gr
        return ms

    mark("model + reducer ready")
    graphed = None
    if not args.eager:
        from polydis_b200.graphs import GraphedTrainStep
        c0 = _lib.call_count
        graphed = GraphedTrainStep(model, opt, B, reducer=reducer, warmup=3 if world == 1 else 11).capture(x, c, pr)
        calls_per_step = (_lib.call_count - c0) // (graphed._warm + 1)

        def step(x, c, pr):  # noqa: F811
            return graphed(x, c, pr)[0]
    mark("graph captured" if graphed is not None else "eager mode")
    for _ in range(args.warmup):
        step(x, c, pr)
    mark("warm-up done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    calls0 = _lib.call_count
    ms_step = timed(lambda: step(x, c, pr), args.steps)
    launches = (_lib.call_count - calls0) if graphed is None else calls_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None

    mark(f"timed region done: {ms_step:.2f} ms/step")
    # BASELINE configs[3]: data-parallel training at GLOBAL batch 4096 (strong-scaling point: 4096 / N per GPU)
    strong = None
    if world > 1 and graphed is not None and not args.no_strong and 4096 % world == 0 and 4096 // world != B:
        from polydis_b200.graphs import GraphedTrainStep
        Bs = 4096 // world
        xs_, cs_, ps_ = (torch.from_numpy(a).to(dev) for a in synth_batch(Bs, 300 + rank))
        gs = GraphedTrainStep(model, opt, Bs, reducer=reducer, warmup=11).capture(xs_, cs_, ps_)
        for _ in range(3):
            gs(xs_, cs_, ps_)
        ms_s = timed(lambda: gs(xs_, cs_, ps_), max(3, args.steps // 2))
        strong = {"global_batch": 4096, "batch_per_gpu": Bs, "ms_per_step": ms_s, "value": 4096 / (ms_s * 1e-3),
                  "unit": "samples/s", "what": "BASELINE configs[3]: global batch 4096 over N GPUs"}
        del gs, xs_, cs_, ps_
        mark(f"strong-scaling point: {ms_s:.2f} ms/step at {Bs}/GPU")
    # free-running regime (tfr = 0,0,0: what train.py's schedule yields after its first step, scheduler.py:48-49):
    # 32 x 15 sequential note steps with greedy feedback, same kernels step-wise, whole step in one CUDA graph
    ms_tfr0 = None
    if world == 1 and graphed is not None and not args.no_tfr0:
        from polydis_b200.graphs import GraphedTrainStep
        from polydis_b200.ptvae import PtvaeDecoder
        PtvaeDecoder.batched_sampling = not args.stepwise_sampling     # default: greedy pass + batched phases
        g0 = GraphedTrainStep(model, opt, B, tfr=(0., 0., 0.), warmup=1).capture(x, c, pr)
        g0(x, c, pr)
        ms_tfr0 = timed(lambda: g0(x, c, pr), 3)
        del g0
        PtvaeDecoder.batched_sampling = True
        mark(f"tfr=0 step: {ms_tfr0:.1f} ms")
    # end-to-end: pinned host buffers -> device inside the timed region, loss read back
    # (graph mode: double-buffered -- every step copies one full batch from pinned host memory, the one the NEXT
    # step trains on, behind the current replay, as a training loop's prefetcher does; eager mode: copy, then step)
    if graphed is not None:
        graphed.prefetch(xh, ch, ph)

        def e2e_step():
            return float(graphed.step_prefetched(next_batch=(xh, ch, ph))[0])
    else:
        def e2e_step():
            xd, cd, pd = xh.to(dev, non_blocking=True), ch.to(dev, non_blocking=True), ph.to(dev, non_blocking=True)
            return float(step(xd, cd, pd))
    e2e_step()
    ms_e2e = timed(e2e_step, max(2, args.steps // 2))
    h2d = xh.numel() * 8 + ch.numel() * 4 + ph.numel() * 4

    mark("e2e done")
    # greedy decode (encode chord+texture -> means -> PianoTree decode -> int tokens on device), replayed
    # from a CUDA graph; fp32-faithful GEMMs (token parity with the fp32 reference) and TF32 tensor cores
    from polydis_b200.graphs import GraphedDecode
    Bd = args.decode_batch
    xd_, cd_, pd_ = (torch.from_numpy(a).to(dev) for a in synth_batch(Bd, 500 + rank))
    dec_ms = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        model.decode_precision = prec
        gd = GraphedDecode(model, Bd).capture(pd_, cd_)
        gd(pd_, cd_)
        dec_ms[prec] = timed(lambda: gd(pd_, cd_), 3)
        del gd
    model.decode_precision = "tf32x3"
    ms_dec = dec_ms["tf32x3"]
    # decode end to end (BASELINE configs[2], 65,536 segments on one GPU): 4 replays of 16,384 segments, every replay
    # fed from its own pinned host batch (pr_mat + c, H2D inside the timed region) and its tokens brought back to the
    # host in the compact 2-byte format (ops.pack_tokens; the reference copies int64 tokens AND the logits,
    # ptvae.py:537-544)
    dec_e2e = None
    if not args.no_decode_e2e:
        n_chunks = max(1, 65536 // Bd)
        host = []
        for k in range(n_chunks):
            _, c_k, p_k = synth_batch(Bd, 7000 + 10 * rank + k)
            host.append((torch.from_numpy(p_k).pin_memory(), torch.from_numpy(c_k).pin_memory()))
        gd = GraphedDecode(model, Bd, pack=True).capture(pd_, cd_)
        out_host = torch.empty((n_chunks,) + tuple(gd.packed.shape), dtype=torch.uint8).pin_memory()

        def decode_all():
            for k, (p_k, c_k) in enumerate(host):
                gd(p_k, c_k)                                   # H2D into the static buffers + replay (+ pack)
                out_host[k].copy_(gd.packed, non_blocking=True)
        decode_all()
        ms_all = timed(decode_all, 2)
        torch.cuda.synchronize()
        h2d_dec = sum(p_k.numel() * 4 + c_k.numel() * 4 for p_k, c_k in host)
        dec_e2e = {"value": world * n_chunks * Bd / (ms_all * 1e-3), "unit": "segments/s", "segments_per_gpu": n_chunks * Bd,
                   "replays": n_chunks, "ms_total": ms_all, "h2d_bytes": h2d_dec, "d2h_bytes": out_host.numel(),
                   "how": "pinned host pr_mat + c per replay -> device, greedy decode graph, packed uint8 tokens -> pinned host"}
        del gd, host
    # latency-bound corner (BASELINE configs[4]: a 256-bar arrangement = 128 segments over 8 GPUs = 16 per GPU)
    xs_, cs_, ps_ = (torch.from_numpy(a).to(dev) for a in synth_batch(16, 900 + rank))
    gd16 = GraphedDecode(model, 16).capture(ps_, cs_)
    gd16(ps_, cs_)
    ms_dec16 = timed(lambda: gd16(ps_, cs_), 3)
    del gd16
    model.train()

    mark("decode done")
    # dominant kernel timed alone: the note-GRU recurrent GEMM [32B x 512] . [512 x 1536]
    from polydis_b200 import ops
    R = 32 * B
    hA = torch.randn(R, 512, device=dev)
    wB = torch.randn(1536, 512, device=dev)
    oC = torch.empty(R, 1536, device=dev)
    for _ in range(3):
        ops.gemm_nt(hA, wB, oC)
    ms_gemm = timed(lambda: ops.gemm_nt(hA, wB, oC), 20)
    gemm_tflops = 2.0 * R * 512 * 1536 / (ms_gemm * 1e-3) / 1e12
    del hA, oC
    # dominant kernel of the step, timed alone: the fused note-GRU step (recurrent GEMM on tcgen05 + gate math in the
    # epilogue, the step's x-projection as a second K segment, all operands / results as TMA boxes).  HBM-bound: per
    # launch it must read h_prev, gi2 and the 128-wide embedding rows and write h, r|z|n, W_hn h -- 15 launches over
    # distinct (b,t) slices, so every launch streams ~300 MB of fresh data (> 126 MB L2).
    T_, H_, K2_ = 15, 512, 128
    x_ = torch.randn(R, T_ + 1, K2_, device=dev)
    gi2_ = torch.randn(R, 3 * H_, device=dev)
    h_ = torch.randn(R, T_ + 1, H_, device=dev) * 0.3
    rzn_ = torch.empty(R, T_, 3 * H_, device=dev)
    hn_ = torch.empty(R, T_, H_, device=dev)
    wB.mul_(0.03)
    wX = torch.randn(3 * H_, K2_, device=dev) * 0.05
    bB = torch.randn(3 * H_, device=dev) * 0.1

    def step_seq():
        # exactly the call ops.gru_seq issues for the teacher-forced note GRU (x-projection folded in: no gi tensor)
        for t_ in range(T_):
            ops._call("pd_gru_step_tmax", h_[:, t_].data_ptr(), h_.stride(0), wB.data_ptr(), H_, x_[:, t_].data_ptr(),
                      x_.stride(0), wX.data_ptr(), K2_, K2_, bB.data_ptr(), gi2_.data_ptr(), 3 * H_,
                      h_[:, t_ + 1].data_ptr(), h_.stride(0), rzn_[:, t_].data_ptr(), rzn_.stride(0), hn_[:, t_].data_ptr(),
                      hn_.stride(0), R, H_, torch.cuda.current_stream().cuda_stream)
    step_seq()
    ms_fstep = timed_graph(step_seq, 5) / T_
    # h_prev + gi2(3) in, h + rzn(3) + hn out = 9 H floats per row-step, + the step's 128 embedding floats, + W_hh | W_x
    step_bytes = R * (9 * H_ + K2_) * 4 + 3 * H_ * (H_ + K2_) * 4
    step_gbs = step_bytes / (ms_fstep * 1e-3) / 1e9
    del x_, gi2_, h_, rzn_, hn_
    # the kernel family with the largest share of the packed step (profiles/r02_v7_train_step_launches.txt): the
    # dgh . W_hh GEMM of the batch-sized recurrences' backward steps (time GRU, both encoder bi-GRUs, chord decoder):
    # [B x 3H] . [3H x H], H = 1024, split over K with a red.add epilogue into an accumulator the gate kernel cleared.
    # 32 launches over distinct gate-gradient slices (as in the 32-step time GRU), W_hh stays L2-resident as in the step.
    Hh = 1024
    dgh_ = torch.randn(32, B, 3 * Hh, device=dev)
    w_hh_ = torch.randn(3 * Hh, Hh, device=dev) * 0.03
    dm_ = torch.zeros(2, B, Hh, device=dev)

    dghb_ = dgh_.to(torch.bfloat16)
    wbb_ = ops.to_bf16(w_hh_)

    def dh_seq():
        # as ops._gru_steps_bwd issues it in training: bf16 copies of dgh (written by the gate kernel) and W_hh
        for t_ in range(32):
            ops._call("pd_gemm_bf16", dghb_[t_].data_ptr(), 3 * Hh, 1, wbb_.data_ptr(), Hh, 1, dm_[t_ & 1].data_ptr(), Hh, None,
                      B, Hh, 3 * Hh, 1, torch.cuda.current_stream().cuda_stream)
    dh_seq()
    ms_dh = timed_graph(dh_seq, 10) / 32
    dh_flop = 2.0 * B * 3 * Hh * Hh
    dh_tflops = dh_flop / (ms_dh * 1e-3) / 1e12
    del dgh_, dm_
    # ... and of their forward steps: the fused step kernel with 32-unit tiles (recurrent GEMM [B x H] . [H x 3H] on tcgen05 +
    # gate math in the epilogue, operands / results as TMA boxes) -- the single kernel with the largest share of the step
    gi_b = torch.randn(B, 33, 3 * Hh, device=dev)
    h_b = torch.randn(B, 33, Hh, device=dev) * 0.3
    rzn_b = torch.empty(B, 32, 3 * Hh, device=dev)
    hn_b = torch.empty(B, 32, Hh, device=dev)
    b_b = torch.randn(3 * Hh, device=dev) * 0.1

    wb_b = ops.to_bf16(w_hh_)
    hb_b = torch.empty(2, B, Hh, device=dev, dtype=torch.bfloat16)
    hb_b[0].copy_(h_b[:, 0])

    def fwd_seq():
        # the call ops._gru_steps_fwd issues for these recurrences in training: bf16 copies of h / W_hh as tcgen05 operands
        for t_ in range(32):
            ops._call("pd_gru_step_tma_bf16", hb_b[t_ & 1].data_ptr(), Hh, wb_b.data_ptr(), Hh, b_b.data_ptr(),
                      gi_b[:, t_].data_ptr(), gi_b.stride(0), None, 0, h_b[:, t_].data_ptr(), h_b.stride(0),
                      h_b[:, t_ + 1].data_ptr(), h_b.stride(0), hb_b[(t_ & 1) ^ 1].data_ptr(), Hh,
                      rzn_b[:, t_].data_ptr(), rzn_b.stride(0), hn_b[:, t_].data_ptr(), hn_b.stride(0), B, Hh,
                      torch.cuda.current_stream().cuda_stream)
    fwd_seq()
    ms_bstep = timed_graph(fwd_seq, 10) / 32
    bstep_tflops = dh_flop / (ms_bstep * 1e-3) / 1e12
    del gi_b, h_b, rzn_b, hn_b
    # live share of the note level in THIS batch (device table of the packed path): executed work, not the dense count
    tok_, len_, _, _ = ops.grid_prepare(x)
    tab_ = ops.Packed(tok_, len_).table.cpu()
    live_frac = float(tab_[18:33].sum()) / (15.0 * R)
    del tok_, len_

    def leave():
        # a process group whose collectives were captured in CUDA graphs can block in
        # destroy_process_group(); all results are out, so synchronise, rendezvous and exit hard
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        leave()
        return
    peak_tf, peak_hbm, peak_src = _peaks()
    sps = world * B / (ms_step * 1e-3)
    # executed algorithmic work: 66.5 % of the dense minimum (603.7 of 908.2 M MAC: note GRU, heads, duration decoder) is
    # note-level and only its live share is computed in loss mode (packed note level)
    gflop_exec = TRAIN_GFLOP_PER_SAMPLE * (1.0 - 0.6647 * (1.0 - live_frac))
    achieved = gflop_exec * B / (ms_step * 1e-3) / 1e3                   # TFLOP/s per GPU
    if getattr(reducer, "impl", None) == "p2p" and reducer.peer_error():
        raise RuntimeError("bench: a gradient-exchange kernel gave up waiting for a peer; the timed steps are invalid")
    out = {"metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "tf32+bf16", "data": "synthetic",
           "dtype_note": "fp32 storage, accumulation, gate math and recurrent state; GEMM operands rounded to TF32 on the tensor "
                         "cores, bf16 operand copies (W_hh, h, dgh) for the recurrent GEMMs of the batch-sized recurrences",
           "config": {"workload": "PolyDisVAE training step (zero_grad+fwd+loss+bwd+clip_grad_norm+Adam), "
                                  "teacher-forced PianoTree decoder, batch 512 per GPU (BASELINE configs[1])",
                      "batch_per_gpu": B, "global_batch": world * B, "tfr": [1, 1, 1],
                      "l2_policy": "working set per step (>5 GB of activations) exceeds the 126 MB L2; no explicit flush",
                      "note_level": "packed: rows sorted by token count, note slots whose target is PAD are not computed in "
                                    "loss mode (losses and gradients unchanged; tests/test_gpu_model.py)",
                      "parallelism": f"dp{world}", "cuda_graph": graphed is not None,
                      "optimizer": "FusedClipAdam (clip 1.0, lr 1e-3, gamma 0.9999, floor 1e-5)" if fused_opt else ("global norm from the exchange kernels' partials, clip folded into torch Adam(fused) via grad_scale"
                                                                  if getattr(reducer, "impl", None) == "p2p" else
                                                                  "torch _foreach_norm, clip folded into torch Adam(fused) via grad_scale"
                                                                  if graphed is not None else "torch clip_grad_norm_ + Adam(fused)"),
                      **({"gradient_exchange": EXCHANGE_NOTE[reducer.impl], "bucket_mb": args.bucket_mb} if reducer is not None and world > 1 else {})},
           "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 4,
                   "how": ("GraphedTrainStep.prefetch/step_prefetched: each step copies one full batch from pinned host "
                           "memory (the next step's, overlapped with the replay) and reads the loss back"
                           if graphed is not None else "copy, step, read the loss back")},
           "gpu_launches": launches, "configs3_global_4096": strong,
           "train_free_running": None if ms_tfr0 is None else
           {"value": B / (ms_tfr0 * 1e-3), "unit": "samples/s", "ms_per_step": ms_tfr0, "tfr": [0, 0, 0], "cuda_graph": True,
            "path": "step-wise" if args.stepwise_sampling else "greedy pass + batched phases"},
           "decode": {"value": world * Bd / (ms_dec * 1e-3), "unit": "segments/s", "batch_per_gpu": Bd,
                      "ms_per_batch": ms_dec,
                      "precision": "tf32x3 (error-compensated tensor-core GEMMs; token parity with the fp32 reference)",
                      "fp32_ffma_value": world * Bd / (dec_ms["fp32"] * 1e-3),
                      "tf32_value": world * Bd / (dec_ms["tf32"] * 1e-3), "cuda_graph": True,
                      "latency_16_segments_ms": ms_dec16, "e2e": dec_e2e},
           "roofline": {"bound": "tensor", "achieved": bstep_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": bstep_tflops / peak_tf, "traffic": _measured_traffic("gru_step_tma_kernel<4, 1, 1, 0, 0, 32, 2>"),
                        "peak_source": peak_src,
                        "kernel": "gru_step_tma_kernel<32-unit tiles, bf16 operands> via pd_gru_step_tma_bf16: one forward step of a "
                                  "batch-sized recurrence (time GRU / encoders / chord decoder), [B x 1024] . [1024 x 3072] on "
                                  "tcgen05 (kind::f16, fp32 accumulate) + gate math in the epilogue",
                        "ms_per_launch": ms_bstep, "algorithmic_flop_per_launch": dh_flop,
                        "what": "the kernel family with the largest share of the packed step (profiles/r02_v7_train_step_launches.txt: "
                                "32 launches of this 32-unit form for the time GRU + 36 of the 64-unit form the short concurrent "
                                "recurrences use, 11.9 % together), timed alone with CUDA events as a captured graph of 32 dependent launches "
                                "over the slices of a (B,33,.) sequence (10 replays): 2 M N K / launch time vs the measured "
                                "sustained bf16 tensor peak. "
                                "At batch 512 the kernel is ONE wave of 128 CTAs with 16 dependent k-blocks: bound by launch + TMA "
                                "pipeline latency, not by the tensor pipe; "
                                "traffic = ncu dram bytes per launch (profiles/r02_kernel_traffic.json)",
                        "recurrent_dh_gemm": {"achieved": dh_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": dh_tflops / peak_tf,
                                              "ms_per_launch": ms_dh,
                                              "traffic": _measured_traffic("gemm_tf32_kernel<2, 1, 64, 3, 3, 0, 1>"),
                                              "kernel": "gemm_tf32_kernel<bf16 operands, 64-wide, 3 stages, NN>: dgh . W_hh of the same "
                                                        "recurrences' backward steps, [B x 3072] . [3072 x 1024], split-K red.add "
                                                        "epilogue (82 launches)"},
                        "fused_note_step_hbm": {"achieved": step_gbs, "peak": peak_hbm, "unit": "GB/s",
                                                "frac": step_gbs / peak_hbm,
                                                "traffic": _measured_traffic("gru_step_tma_kernel<3, 1, 1, 0, 1, 64, 4>"),
                                                "kernel": "gru_step_tma_kernel<SEG2> (fused note-GRU step, all 32B rows live)",
                                                "ms_per_launch": ms_fstep, "algorithmic_bytes_per_launch": step_bytes},
                        "whole_step_tensor": {"achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                                              "frac": achieved / peak_tf, "gflop_per_sample_executed": gflop_exec,
                                              "gflop_per_sample_dense": TRAIN_GFLOP_PER_SAMPLE,
                                              "note_level_live_fraction": live_frac,
                                              "what": "EXECUTED algorithmic GFLOP/sample x batch / step time vs sustained bf16 "
                                                      "peak; dense = the reference's work incl. the PAD-target note slots the "
                                                      "packed path does not compute"},
                        "note_gemm_alone": {"name": "note-GRU recurrent GEMM [32B x 512].[512 x 1536] (unfused route)",
                                            "ms": ms_gemm, "achieved": gemm_tflops, "frac": gemm_tflops / peak_tf,
                                            "traffic": _measured_traffic("gemm_tf32_persistent")}},
           "clocks": clocks}
    if world == 1 and not args.no_cpu:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2",
                                "--warmup", "1"], capture_output=True, text=True, timeout=900,
                               env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
            ref = json.loads(r.stdout.strip().splitlines()[-1])
            out["cpu_baseline"] = ref["cpu_baseline"]
            out["cpu_baseline"]["decode_segments_per_sec"] = ref["decode"]["value"]
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"failed: {e!r}"}
    print(json.dumps(out))
    leave()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--decode-batch", type=int, default=16384)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="gradient exchange under --gpus N>1: the repo's peer-memory all-reduce kernel (default) or NCCL")
    ap.add_argument("--bucket-mb", type=float, default=16.0)
    ap.add_argument("--no-tfr0", action="store_true", help="skip the free-running (tfr=0) training measurement")
    ap.add_argument("--no-decode-e2e", action="store_true", help="skip the 65,536-segment end-to-end decode measurement")
    ap.add_argument("--no-strong", action="store_true", help="skip the configs[3] strong-scaling point under --gpus N>1")
    ap.add_argument("--stepwise-sampling", action="store_true",
                    help="free-running measurement through the step-wise path instead of the batched scheduled-sampling path")
    ap.add_argument("--fused-optim", action="store_true",
                    help="polydis_b200.optim.FusedClipAdam (flat buckets) instead of torch clip_grad_norm_ + fused Adam; "
                         "measured 0.9 ms/step slower at 1 GPU because backward then accumulates into the flat buckets")
    ap.add_argument("--eager", action="store_true", help="issue the training step eagerly (no CUDA graph)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
