"""Data-parallel gradient exchange for PolyDis training: one process per GPU, NCCL all-reduce of the
27.3 M-parameter gradient over NVLink, bucketed and overlapped with backward.

Replaces the reference's single-process ``nn.DataParallel`` wrap (amc_dl/torch_plus/module.py:67-68).
Semantics match its loss averaging (module.py:152-157): every rank computes the per-shard mean losses,
gradients are averaged over ranks (= gradient of the mean of per-shard losses).

Design: parameters are packed, in reverse registration order (decoder / chord-decoder gradients are
produced first by backward, the encoders' last), into a few flat fp32 buckets; ``p.grad`` are views into
them, so backward accumulates straight into the communication buffers (no copy).  A
post-accumulate-grad hook counts a bucket's parameters down and, when the bucket is complete, issues
its all-reduce (``ReduceOp.AVG``) on a side stream ordered after the producing kernels -- the exchange
overlaps the rest of backward, and because every launch is a stream operation the whole thing is
capturable in the training step's CUDA graph.  ``finish()`` joins the side stream before the global-norm
clip and the optimizer step (which therefore see identical gradients on every rank).
"""
import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, params, bucket_mb=32, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.cuda = self.params[0].is_cuda
        cap = int(bucket_mb * (1 << 20) // 4)
        self.buckets = []                  # dicts: flat, params, pending
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._close(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._close(cur)
        self.comm = torch.cuda.Stream() if self.cuda else None
        self._handles = []
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
        self.reset()

    def _close(self, plist):
        n = sum(-(-p.numel() // 4) * 4 for p in plist)             # 16-byte aligned slots
        flat = torch.zeros(n, device=plist[0].device, dtype=torch.float32)
        off = 0
        for p in plist:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += -(-p.numel() // 4) * 4
        self.buckets.append({"flat": flat, "params": plist, "pending": len(plist)})

    def _make_hook(self, bi):
        def hook(_p):
            b = self.buckets[bi]
            b["pending"] -= 1
            if b["pending"] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if self.world == 1:
            return
        if self.cuda:
            self.comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                dist.all_reduce(b["flat"], op=dist.ReduceOp.AVG, group=self.group)
        else:                                                       # gloo (CPU tests): no AVG
            dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group)
            b["flat"].div_(self.world)

    def reset(self):
        """Zero the gradient buckets and re-arm the hooks (call instead of optimizer.zero_grad)."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])

    def finish(self):
        """Make the compute stream wait for every bucket's all-reduce."""
        if self.cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm)
        for b in self.buckets:
            if b["pending"] != 0:
                raise RuntimeError("a parameter received no gradient; its bucket was never reduced")

    def remove(self):
        for h in self._handles:
            h.remove()
