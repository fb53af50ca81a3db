"""Data-parallel gradient exchange for PolyDis training: one process per GPU, NCCL all-reduce of the
27.3 M-parameter gradient over NVLink, bucketed and overlapped with backward.

Replaces the reference's single-process ``nn.DataParallel`` wrap (amc_dl/torch_plus/module.py:67-68).
Semantics match its loss averaging (module.py:152-157): every rank computes the per-shard mean losses,
gradients are averaged over ranks (= gradient of the mean of per-shard losses).

Design (round 2):

* parameters are packed, in reverse registration order (decoder / chord-decoder gradients are produced first
  by backward, the encoders' last), into flat fp32 buckets of at most ``bucket_mb``;
* ``p.grad`` is ``None`` while backward runs, so autograd's AccumulateGrad *adopts* the gradient tensor the
  producing kernel wrote (no zero pass over the buckets, no read-modify-write add per parameter -- 81 extra
  kernels on the step's critical path in round 1);
* a post-accumulate-grad hook records a CUDA event on the stream the gradient was produced on (the model forks
  encoders / decoder / chord decoder onto side streams, ``ops.fork_join``, and autograd replays their backward
  on those streams) and counts the bucket down; when the bucket is complete the communication stream waits for
  the events of ALL its parameters, gathers the gradients into the flat buffer with one multi-tensor copy, and
  issues the all-reduce (``ReduceOp.AVG``).  ``p.grad`` is then re-pointed at the bucket views, so the clip and
  the optimizer read the averaged gradients in place;
* every operation is a stream operation, so the whole exchange is captured inside the training step's CUDA graph;
* ``finish()`` joins the communication stream before the global-norm clip and the optimizer step (which therefore
  see identical gradients on every rank).

``impl="p2p"`` (CUDA, world 2..8 on one NVLink / NVSwitch node) replaces the per-bucket multi-tensor copy + NCCL call by
ONE kernel of the repo's library (``csrc/allreduce_p2p.cu``): the buckets live in a CUDA-IPC "symmetric" region every
peer has mapped; the kernel gathers the bucket from the producers' buffers, reduces this rank's slice with peer loads
over NVLink, stores the average into every rank's copy and leaves the squared-norm partials of the clip
(``clip_grad_norm_``).  Measured reason (2 GPUs, ``tools/ddp_trace.py``): with the deferred weight gradients most
buckets complete only when backward ends, so the exchange is an exposed tail of copy + ``ncclAllReduce`` (RING_LL,
15-70 us per 8 MB bucket) launches -- 0.75 ms of an 8.57 ms step.
"""
import ctypes
import os

import torch
import torch.distributed as dist


class _DeviceRegion:
    """A raw device allocation exposed through ``__cuda_array_interface__`` (``torch.as_tensor`` wraps it, no copy)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def observe_ready_order(params, run_backward):
    """Run ``run_backward()`` (one forward + backward) and return the parameters in the order autograd finished
    accumulating their gradients -- the bucket order that lets each all-reduce start as early as possible (the deferred
    weight-gradient jobs of ``ops.defer`` make this differ from reverse registration order).  Gradients are dropped."""
    params = [p for p in params if p.requires_grad]
    seen, handles = [], []
    for p in params:
        handles.append(p.register_post_accumulate_grad_hook(lambda q, seen=seen: seen.append(q)))
    for p in params:
        p.grad = None
    run_backward()
    for h in handles:
        h.remove()
    for p in params:
        p.grad = None
    out, ids = [], set()
    for p in seen:
        if id(p) not in ids:
            ids.add(id(p))
            out.append(p)
    return out + [p for p in reversed(params) if id(p) not in ids]


class BucketedGradAllReduce:
    """Bucketed, backward-overlapped gradient averaging over the ranks of a process group (see the module docstring).
    Per step: ``reset()`` .. forward .. backward .. ``finish()`` .. (``clip_grad_norm_`` | ``grad_norm``) .. optimizer step;
    ``p.grad`` of every parameter is then a view of its bucket holding the averaged gradient."""

    def __init__(self, params, bucket_mb=8, group=None, ready_order=None, comm_priority=0, impl="nccl", ar_blocks=None):
        """``ready_order``: the parameters in the order their gradients become ready (``observe_ready_order``); default
        reverse registration order.  ``comm_priority``: CUDA priority of the communication stream (negative = higher).
        ``impl``: "nccl" (gather + ``ncclAllReduce`` per bucket) | "p2p" (the library's peer-memory exchange kernel; raises
        ``RuntimeError`` on every rank if CUDA IPC / peer access is unavailable -- the caller decides about a fall-back).
        ``ar_blocks``: CTAs per exchange kernel (default ``PD_AR_BLOCKS`` or 32; ``PD_AR_STREAMS``, default 4, exchange
        streams)."""
        self.group = group
        self.impl = impl
        assert impl in ("nccl", "p2p")
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.cuda = self.params[0].is_cuda
        cap = int(bucket_mb * (1 << 20) // 4)
        self.buckets = []                  # dicts: flat, params, views, pending, events, adopted
        cur, cur_n = [], 0
        if ready_order is not None:
            order = [p for p in ready_order if p.requires_grad]
            assert len(order) == len(self.params) and {id(p) for p in order} == {id(p) for p in self.params}
        else:
            order = list(reversed(self.params))
        for p in order:
            if cur and cur_n + p.numel() > cap:
                self._close(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._close(cur)
        if impl == "p2p":
            self._setup_p2p(ar_blocks)
        self.comm = torch.cuda.Stream(priority=comm_priority) if self.cuda else None
        # p2p: the exchanges of different buckets are independent (a flag slot each), so they are dealt onto several
        # streams -- when many buckets become ready together (backward's end) their barrier latencies overlap and the tail
        # is bound by NVLink bytes, not by ~25 us per launch.  n_streams x ar_blocks <= SM count keeps every in-flight
        # exchange block resident on every rank (two ranks that scheduled different buckets first still both progress).
        self.comms = [self.comm]
        if impl == "p2p":
            n_comm = int(os.environ.get("PD_AR_STREAMS", "4"))
            n_sm = torch.cuda.get_device_properties(self.params[0].device).multi_processor_count
            assert n_comm * self.ar_blocks <= n_sm, "PD_AR_STREAMS x PD_AR_BLOCKS must not exceed the SM count"
            self.comms += [torch.cuda.Stream(priority=comm_priority) for _ in range(n_comm - 1)]
        self._handles = []
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b["params"]):
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi, pi)))
        self.reset()

    def _close(self, plist):
        n = sum(-(-p.numel() // 4) * 4 for p in plist)             # 16-byte aligned slots
        flat = None if self.impl == "p2p" else torch.zeros(n, device=plist[0].device, dtype=torch.float32)
        events = [torch.cuda.Event() for _ in plist] if plist[0].is_cuda else None
        b = {"flat": flat, "n": n, "params": plist, "views": None, "pending": len(plist), "events": events, "adopted": None,
             "index": len(self.buckets)}
        if flat is not None:
            self._bind(b, flat)
        self.buckets.append(b)

    @staticmethod
    def _bind(b, flat):
        views, offs, off = [], [], 0
        for p in b["params"]:
            views.append(flat[off:off + p.numel()].view_as(p))
            offs.append(off)
            off += -(-p.numel() // 4) * 4
        b["flat"], b["views"], b["offs"] = flat, views, offs

    def _setup_p2p(self, ar_blocks):
        """Allocate this rank's symmetric region, exchange the IPC handles, map the peers' regions."""
        from ._lib import lib
        if not (self.cuda and dist.is_initialized() and 2 <= self.world <= lib.pd_ar_limit(0)):
            raise RuntimeError("impl='p2p' needs CUDA parameters and a process group of 2..8 ranks on one node")
        self._lib = lib
        self.rank = dist.get_rank(self.group)
        self.ar_blocks = int(ar_blocks or os.environ.get("PD_AR_BLOCKS", "32"))
        assert 1 <= self.ar_blocks <= lib.pd_ar_limit(1)
        self.max_src = lib.pd_ar_limit(2)
        dev = self.params[0].device
        self.flag_bytes = lib.pd_ar_flag_bytes(len(self.buckets))
        total = sum(b["n"] for b in self.buckets)
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        with torch.cuda.device(dev):
            rc = lib.pd_ipc_alloc(self.flag_bytes + total * 4, ctypes.byref(ptr), handle)
            handles = [None] * self.world
            dist.all_gather_object(handles, (rc, bytes(handle.raw)), group=self.group)
            self._region = ptr.value
            self._peer_ptrs = (ctypes.c_void_p * self.world)()
            self._opened = []
            why = "" if rc == 0 else f"pd_ipc_alloc -> {rc}"
            if all(h[0] == 0 for h in handles):
                for r, (_, h) in enumerate(handles):
                    if r == self.rank:
                        self._peer_ptrs[r] = ptr.value
                        continue
                    q = ctypes.c_void_p()
                    rc = lib.pd_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(q))
                    if rc != 0:
                        why = f"pd_ipc_open(rank {r}) -> cudaError {rc}"
                        break
                    self._peer_ptrs[r] = q.value
                    self._opened.append(q.value)
            data = None
            if not why:
                try:
                    self._mem = _DeviceRegion(ptr.value + self.flag_bytes, total)
                    data = torch.as_tensor(self._mem, device=dev)
                    assert data.data_ptr() == ptr.value + self.flag_bytes and data.dtype == torch.float32
                except Exception as e:                  # noqa: BLE001 -- reported to every rank below
                    why = f"wrapping the region as a tensor failed: {e!r}"
            # every rank learns whether ALL ranks are set up (a rank raising alone would leave the others in a collective)
            whys = [None] * self.world
            dist.all_gather_object(whys, why, group=self.group)
            if any(whys):
                raise RuntimeError("impl='p2p' unavailable: " + "; ".join(f"rank {r}: {w}" for r, w in enumerate(whys) if w))
        self.flat_all = data
        off = 0
        for b in self.buckets:
            b["off"] = off
            self._bind(b, data[off:off + b["n"]])
            off += b["n"]
        self._epoch = torch.zeros(len(self.buckets) * lib.pd_ar_limit(1), device=dev, dtype=torch.int32)
        self._err = torch.zeros(1, device=dev, dtype=torch.int32)
        self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        dist.barrier(group=self.group)          # every rank has mapped every region before the first kernel touches one

    def _exchange_p2p(self, bi, b, grads):
        """One kernel: gather ``grads`` into the bucket, all-reduce (average) it over the peers, leave the norm partials."""
        lib = self._lib
        ok = len(grads) <= self.max_src and all(g.is_contiguous() and g.dtype == torch.float32 and g.data_ptr() % 4 == 0
                                                and g.numel() == p.numel() for g, p in zip(grads, b["params"]))
        if ok:
            n_src = len(grads)
            src = (ctypes.c_void_p * n_src)(*[g.data_ptr() for g in grads])
            soff = (ctypes.c_long * n_src)(*b["offs"])
            sn = (ctypes.c_long * n_src)(*[g.numel() for g in grads])
        else:
            torch._foreach_copy_(b["views"], grads)
            n_src, src, soff, sn = 0, None, None, None
        from . import _lib
        _lib.call("pd_allreduce_p2p", self._peer_ptrs, self.rank, self.world, self.flag_bytes, b["off"], b["n"],
                  1.0 / self.world, self._epoch.data_ptr(), self._err.data_ptr(), bi, 1, self.ar_blocks, src, soff, sn, n_src,
                  torch.cuda.current_stream().cuda_stream)

    def _make_hook(self, bi, pi):
        def hook(_p):
            b = self.buckets[bi]
            if b["events"] is not None:
                # the gradient is complete on the stream this AccumulateGrad node runs on, which need not be the
                # stream of the hook that fires last for the bucket
                b["events"][pi].record(torch.cuda.current_stream())
            b["pending"] -= 1
            if b["pending"] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        grads = [p.grad for p in b["params"]]
        if any(g is None for g in grads):
            raise RuntimeError("BucketedGradAllReduce: a parameter's hook fired without a gradient")
        if self.cuda:
            comm = self.comms[b["index"] % len(self.comms)]
            for ev in b["events"]:
                comm.wait_event(ev)
            with torch.cuda.stream(comm):
                if self.impl == "p2p":
                    self._exchange_p2p(b["index"], b, grads)
                else:
                    torch._foreach_copy_(b["views"], grads)
                    if self.world > 1:
                        dist.all_reduce(b["flat"], op=dist.ReduceOp.AVG, group=self.group)
            b["adopted"] = grads           # keep the producers' buffers alive until finish() has joined ``comm``
        else:                                                       # gloo (CPU tests): no streams, no AVG
            torch._foreach_copy_(b["views"], grads)
            if self.world > 1:
                dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group)
                b["flat"].div_(self.world)
        for p, v in zip(b["params"], b["views"]):
            p.grad = v

    def reset(self):
        """Drop the gradients and re-arm the hooks (call instead of optimizer.zero_grad)."""
        for b in self.buckets:
            for p in b["params"]:
                p.grad = None
            b["pending"] = len(b["params"])
            b["adopted"] = None

    def finish(self):
        """Make the compute stream wait for every bucket's gather + all-reduce."""
        for b in self.buckets:
            if b["pending"] != 0:
                raise RuntimeError("a parameter received no gradient; its bucket was never reduced")
        if self.cuda:
            for comm in self.comms:
                torch.cuda.current_stream().wait_stream(comm)
        for b in self.buckets:
            b["adopted"] = None
        self.check_grad_views()

    def grad_norm(self):
        """Global L2 norm of the averaged gradients as a device scalar (call after ``finish()``).  With ``impl="p2p"`` it
        comes from the partials the exchange kernels left: no pass over the gradients, same bits on every rank."""
        if self.impl == "p2p":
            from . import _lib
            _lib.call("pd_ar_norm_total", self._region, len(self.buckets), self.world, self.ar_blocks,
                      self._sumsq.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return self._sumsq.sqrt().squeeze(0)
        return torch.linalg.vector_norm(torch.stack(torch._foreach_norm([b["flat"] for b in self.buckets])))

    def clip_grad_norm_(self, max_norm):
        """Global-norm clip of the averaged gradients (call after ``finish()``): torch's ``clip_grad_norm_`` semantics
        (coef = max_norm / (norm + 1e-6), capped at 1, always applied).  Returns the norm as a device scalar."""
        norm = self.grad_norm()
        coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        torch._foreach_mul_([b["flat"] for b in self.buckets], coef)
        return norm

    def peer_error(self):
        """True if an exchange kernel gave up waiting for a peer (synchronises; not for use inside a capture)."""
        return self.impl == "p2p" and bool(self._err.item())

    def check_grad_views(self):
        """Every ``p.grad`` must be its bucket view: a stray ``zero_grad(set_to_none=True)`` / foreign optimizer would
        otherwise leave the optimizer stepping on stale bucket contents without any error."""
        for b in self.buckets:
            for p, v in zip(b["params"], b["views"]):
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    raise RuntimeError("BucketedGradAllReduce: p.grad is not the bucket view (gradients were reset or "
                                       "replaced outside reset()/finish())")

    def remove(self):
        for h in self._handles:
            h.remove()
        if self.impl == "p2p" and getattr(self, "_region", None):
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)      # no peer is still inside an exchange kernel on this region
            for q in self._opened:
                self._lib.pd_ipc_close(q)
            # the region itself stays allocated while tensors view it (``flat_all``); it is freed with the process
