"""Fused optimizer tail on libpolydis_b200: global-norm clip + Adam + exponential LR decay with a floor.

What the reference's training loop does after ``loss.backward()`` (amc_dl/torch_plus/module.py:142-143
``clip_grad_norm_``; scheduler.py:69-74 ``Adam.step`` + per-batch LR step; example.py:4-12
``MinExponentialLR``), as two kernels per flat bucket instead of ~10 multi-tensor launches.  Parameters are
re-pointed into flat fp32 buffers that mirror the gradient buckets of ``ddp.BucketedGradAllReduce`` (used
with world size 1 when not distributed), so gradients are already contiguous and -- in the data-parallel
case -- already averaged when ``step()`` runs.  Step count and gradient norm never leave the device: the
whole tail is CUDA-graph capturable.
"""
import torch

from . import _lib, ops
from .ddp import BucketedGradAllReduce


class FusedClipAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clip=1.0, lr_gamma=0.0, lr_min=0.0,
                 reducer=None, bucket_mb=32):
        params = [p for p in params if p.requires_grad]
        self.reducer = reducer if reducer is not None else BucketedGradAllReduce(params, bucket_mb=bucket_mb)
        self.lr, self.betas, self.eps, self.clip = float(lr), betas, float(eps), float(clip or 0.0)
        self.lr_gamma, self.lr_min = float(lr_gamma or 0.0), float(lr_min or 0.0)
        dev = params[0].device
        self.flat_p, self.m, self.v = [], [], []
        for b in self.reducer.buckets:
            fp = torch.empty_like(b["flat"])
            off = 0
            for p in b["params"]:
                n = p.numel()
                fp[off:off + n].copy_(p.data.reshape(-1))
                p.data = fp[off:off + n].view_as(p)            # parameters now live in the flat buffer
                off += -(-n // 4) * 4
            self.flat_p.append(fp)
            self.m.append(torch.zeros_like(fp))
            self.v.append(torch.zeros_like(fp))
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float32)

    def zero_grad(self, set_to_none=False):
        self.reducer.reset()

    def state_dict(self):
        """Moments, step count (which also fixes the position on the LR decay curve) and hyper-parameters; tensors are
        copies, so the dict survives later steps (checkpointing)."""
        return {"m": [t.detach().clone() for t in self.m], "v": [t.detach().clone() for t in self.v],
                "step_count": self.step_count.detach().clone(),
                "hyper": {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "clip": self.clip,
                          "lr_gamma": self.lr_gamma, "lr_min": self.lr_min}}

    def load_state_dict(self, sd):
        """In place (the buffers' addresses may be baked into a captured CUDA graph)."""
        if len(sd["m"]) != len(self.m) or any(a.numel() != b.numel() for a, b in zip(sd["m"], self.m)):
            raise ValueError("FusedClipAdam.load_state_dict: bucket layout differs (bucket_mb / parameter set changed)")
        with torch.no_grad():
            for dst, src in zip(self.m + self.v, list(sd["m"]) + list(sd["v"])):
                dst.copy_(src)
            self.step_count.copy_(sd["step_count"])
        h = sd.get("hyper", {})
        self.lr, self.eps, self.clip = h.get("lr", self.lr), h.get("eps", self.eps), h.get("clip", self.clip)
        self.betas = tuple(h.get("betas", self.betas))
        self.lr_gamma, self.lr_min = h.get("lr_gamma", self.lr_gamma), h.get("lr_min", self.lr_min)

    def grad_norm(self):
        """Global gradient norm of the last step() (device tensor)."""
        return self.sumsq.sqrt()

    def step(self):
        self.reducer.check_grad_views()      # a stray zero_grad(set_to_none=True) would silently step on zero gradients
        st = ops._stream()
        self.sumsq.zero_()
        ops._call("pd_counter_inc", ops._ptr(self.step_count), st)
        for b in self.reducer.buckets:
            ops._call("pd_sumsq_f32", ops._ptr(b["flat"]), b["flat"].numel(), ops._ptr(self.sumsq), st)
        for b, fp, m, v in zip(self.reducer.buckets, self.flat_p, self.m, self.v):
            ops._call("pd_adam_clip_step", ops._ptr(fp), ops._ptr(b["flat"]), ops._ptr(m), ops._ptr(v), fp.numel(),
                      ops._ptr(self.sumsq), ops._ptr(self.step_count), self.lr, self.lr_gamma, self.lr_min,
                      self.betas[0], self.betas[1], self.eps, self.clip, st)
