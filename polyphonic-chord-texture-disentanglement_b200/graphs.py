"""Whole-step CUDA-graph capture of PolyDis training (forward + loss + backward + clip + Adam).

One teacher-forced training step is ~1,400 kernel launches issued from Python (GEMM + gate kernel per
recurrent step, forwards and backwards); eager issue is CPU-bound.  ``GraphedTrainStep`` captures the
whole step once (static input buffers, capturable Adam) and replays it with a single launch per
step -- "CUDA graphs instead of a tracing compiler".  Valid while the host-side control flow is fixed:
the teacher-forcing decisions drawn from python's ``random`` are baked in at capture, so every ratio must
be exactly 0 or 1 (decisions independent of the draw): tfr = (1,1,1) -- the batched teacher-forced path -- or
tfr = (0,0,0), the free-running regime train.py's schedule settles into after its first step
(scheduler.py:48-49); intermediate ratios run eagerly -- or, with ``device_plan=True``, from ONE graph whose
487 teacher-forcing decisions are device data: they are drawn from python ``random`` in the reference's order before
every replay and uploaded, and the step-wise decoder picks ground-truth or predicted rows with a select kernel
(``ops.select_rows``).  (device_plan was written after round 1's GPU budget was spent: its host logic is pinned to the
reference goldens on the CPU emulation, the captured path has not run on hardware yet.)
"""
import torch


#: fold the global-norm clip into torch's fused Adam through ``optimizer.grad_scale`` (see ``_clip_and_step``)
FOLD_CLIP_INTO_ADAM = True


class GraphedTrainStep:
    def __init__(self, model, optimizer, batch, tfr=(1., 1., 1.), beta=0.1, weights=(1, 0.5), clip=1.0,
                 warmup=3, reducer=None, device_plan=False, inject_eps=False, restore_after_capture=True):
        assert device_plan or all(t in (0., 1.) for t in tfr), \
            "graph capture bakes the teacher-forcing plan: every ratio must be 0 or 1 (deterministic decisions); " \
            "use device_plan=True (decisions as device data) or eager steps for 0 < tfr < 1"
        dev = next(model.parameters()).device
        self.device_plan = device_plan
        if device_plan:
            self.plan = torch.zeros(model.N_PLAN, device=dev, dtype=torch.int32)
            self._plan_host = torch.zeros(model.N_PLAN, dtype=torch.int32).pin_memory()
            self._plan_copied = torch.cuda.Event()
            self._plan_copied.record()
        self.model, self.opt = model, optimizer
        self.params = [p for p in model.parameters()]
        self.x = torch.zeros(batch, 32, 16, 6, device=dev, dtype=torch.int64)
        self.c = torch.zeros(batch, 8, 36, device=dev, dtype=torch.float32)
        self.pr = torch.zeros(batch, 32, 128, device=dev, dtype=torch.float32)
        self.tfr, self.beta, self.weights, self.clip = tfr, beta, weights, clip
        #: inject_eps: the reparameterisation noise is read from these static buffers (fill them before a replay)
        #: instead of being drawn inside the graph -- deterministic replays for tests / reproducible runs
        self.eps = (tuple(torch.zeros(batch, 256, device=dev, dtype=torch.float32) for _ in range(2))
                    if inject_eps else None)
        #: restore_after_capture: warm-up and capture run REAL optimizer steps on the first batch; with this on
        #: (default) parameters and optimizer state are put back afterwards, so training starts from the caller's
        #: state and the first batch is trained once (the reference takes one step per batch, module.py:134-149)
        self.restore_after_capture = restore_after_capture
        from .optim import FusedClipAdam
        self.fused = isinstance(optimizer, FusedClipAdam)     # clip + Adam + LR decay inside optimizer.step()
        if self.fused:
            reducer = optimizer.reducer
        self.reducer = reducer          # ddp.BucketedGradAllReduce: its all-reduces are captured in the graph
        self.losses = None
        self.graph = None
        self._warm = warmup

    def _step(self):
        if self.reducer is not None:
            self.reducer.reset()        # grads live in the reducer's flat buckets
        else:
            self.opt.zero_grad(set_to_none=True)
        losses = self.model('train', self.x, self.c, self.pr, tfr1=self.tfr[0], tfr2=self.tfr[1],
                            tfr3=self.tfr[2], beta=self.beta, weights=self.weights,
                            **({"plan_dev": self.plan} if self.device_plan else {}),
                            **({"eps": self.eps} if self.eps is not None else {}))
        losses[0].backward()
        if self.reducer is not None:
            self.reducer.finish()
        self._clip_and_step()
        return torch.stack([l.detach() for l in losses])

    def _clip_and_step(self):
        """Global-norm clip + optimizer step.  With torch's fused Adam the clip is FOLDED into the update: the kernel divides
        every gradient by ``optimizer.grad_scale`` (the GradScaler hook; it also writes the scaled gradient back), so
        1 / coef goes there and the separate ``grads *= coef`` pass over all 27 M gradients disappears (~50 us of a 7.8 ms
        step).  Same semantics as ``clip_grad_norm_``: coef = min(1, clip / (norm + 1e-6))."""
        opt = self.opt
        if not self.clip or self.fused:
            opt.step()
            return
        p2p = getattr(self.reducer, "impl", None) == "p2p"
        fold = (FOLD_CLIP_INTO_ADAM and isinstance(opt, torch.optim.Adam)
                and all(g.get("fused") for g in opt.param_groups) and self.params[0].is_cuda)
        if not fold:
            if p2p:
                self.reducer.clip_grad_norm_(self.clip)     # norm from the exchange kernels' partials
            else:
                torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
            opt.step()
            return
        if p2p:
            norm = self.reducer.grad_norm()
        else:
            grads = [p.grad for p in self.params if p.grad is not None]
            norm = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads)))
        if getattr(self, "_gscale", None) is None:
            self._gscale = torch.ones(1, device=self.params[0].device, dtype=torch.float32)
        self._gscale.copy_(torch.clamp((norm + 1e-6) / self.clip, min=1.0).reshape(1))      # = 1 / coef
        opt.grad_scale = self._gscale
        try:
            opt.step()
        finally:
            del opt.grad_scale

    def set_tfr(self, tfr1, tfr2, tfr3):
        """New teacher-forcing ratios for the following steps (device_plan graphs only: a schedule without re-capture)."""
        assert self.device_plan
        self.tfr = (tfr1, tfr2, tfr3)

    def _upload_plan(self):
        """Draw this step's 487 decisions (python ``random``, reference order) and copy them to the device."""
        self._plan_copied.synchronize()                       # the previous upload has left the pinned buffer
        self._plan_host.copy_(torch.tensor(self.model.draw_plan(*self.tfr), dtype=torch.int32))
        self.plan.copy_(self._plan_host, non_blocking=True)
        self._plan_copied.record()

    def _snapshot(self):
        """Parameters + optimizer state (torch optimizers: state_dict tensors; FusedClipAdam: its flat buffers)."""
        snap = {"p": [p.detach().clone() for p in self.params]}
        if self.fused:
            snap["opt"] = self.opt.state_dict()
        else:
            snap["opt"] = [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.opt.state[p].items()}
                           for p in self.params]
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            for p, q in zip(self.params, snap["p"]):
                p.copy_(q)
            if self.fused:
                self.opt.load_state_dict(snap["opt"])
                return
            for p, st in zip(self.params, snap["opt"]):
                cur = self.opt.state[p]
                if not st:
                    # the optimizer had no state yet: zero what warm-up created, IN PLACE (the captured graph holds
                    # these tensors' addresses)
                    for v in cur.values():
                        if torch.is_tensor(v):
                            v.zero_()
                    continue
                for k, v in st.items():
                    if torch.is_tensor(v):
                        cur[k].copy_(v)
                    else:
                        cur[k] = v

    def capture(self, x, c, pr_mat):
        self.x.copy_(x); self.c.copy_(c); self.pr.copy_(pr_mat)
        if self.device_plan:
            self._upload_plan()
        snap = self._snapshot() if self.restore_after_capture else None
        s = torch.cuda.Stream(priority=-1)      # the step's chain outranks its deferred weight-gradient jobs (ops.defer)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self._warm):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.check_exchange()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = self._step()
        if snap is not None:
            self._restore(snap)
        return self

    def check_exchange(self):
        """Raise if a peer-memory gradient exchange gave up waiting for a peer (the step's gradients would be wrong).
        Synchronises; call it at checkpoints, not every step."""
        if self.reducer is not None and getattr(self.reducer, "impl", None) == "p2p" and self.reducer.peer_error():
            raise RuntimeError("gradient exchange: a peer did not reach its exchange kernel within the time limit")

    def __call__(self, x, c, pr_mat):
        """Copy one batch into the static buffers (H2D if the sources are pinned host tensors), replay,
        return the 11 losses as one device tensor (no host sync).  The returned tensor is a STATIC buffer that the
        next replay overwrites: clone it to keep a step's losses."""
        if self.graph is None:
            self.capture(x, c, pr_mat)
        self.x.copy_(x, non_blocking=True)
        self.c.copy_(c, non_blocking=True)
        self.pr.copy_(pr_mat, non_blocking=True)
        if self.device_plan:
            self._upload_plan()
        self.graph.replay()
        return self.losses

    # -- pipelined input feeding: the host->device copy of the NEXT batch overlaps the replay of the current one --
    def prefetch(self, x, c, pr_mat):
        """Start the asynchronous copy of a (pinned host) batch into the staging buffers on a side stream."""
        if not hasattr(self, "_stage"):
            self._stage = [torch.empty_like(t) for t in (self.x, self.c, self.pr)]
            self._copy_stream = torch.cuda.Stream()
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)          # the previous staged batch has been taken over
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._stage, (x, c, pr_mat)):
                dst.copy_(src, non_blocking=True)
            self._staged.record()

    def step_prefetched(self, next_batch=None):
        """Train on the batch staged by ``prefetch`` and (optionally) start fetching ``next_batch`` behind the
        replay.  Returns the 11 losses as one device tensor."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for dst, src in zip((self.x, self.c, self.pr), self._stage):
            dst.copy_(src, non_blocking=True)                 # device-to-device, microseconds
        self._consumed.record()
        if next_batch is not None:
            self.prefetch(*next_batch)
        if self.device_plan:
            self._upload_plan()
        self.graph.replay()
        return self.losses


class GraphedDecode:
    """CUDA-graph replay of greedy inference: encode chord + texture -> posterior means -> PianoTree greedy
    decode -> int tokens on device (``model.swap`` / ``inference(sample=False)`` semantics, model.py:133-149).
    One segment batch = ~5,600 kernel launches captured once; replays take one launch each."""

    def __init__(self, model, batch, warmup=1, pack=False):
        """``pack``: the graph also packs the tokens to the compact 2-byte device->host format (``self.packed``,
        uint8 (B,32,15,2), ``ops.pack_tokens``)."""
        dev = next(model.parameters()).device
        self.model, self.pack, self.packed = model, pack, None
        self.c = torch.zeros(batch, 8, 36, device=dev, dtype=torch.float32)
        self.pr = torch.zeros(batch, 32, 128, device=dev, dtype=torch.float32)
        self.tokens, self.graph, self._warm = None, None, warmup
        self.eager, self.capture_error = False, None

    def _run(self):
        from . import ops
        m = self.model
        m.eval()
        with torch.no_grad(), ops.precision(m.decode_precision):
            dc, dr = ops.fork_join([lambda: m.chd_encoder(self.c), lambda: m.rhy_encoder(self.pr)])
            tok = m.decoder.greedy_tokens(torch.cat([dc.mean, dr.mean], -1))
            if self.pack:
                self.packed = ops.pack_tokens(tok)
            return tok

    def capture(self, pr_mat, c):
        self.pr.copy_(pr_mat); self.c.copy_(c)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self._warm):
                self._run()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import ops
        ops._lo_cache.clear()           # weight low parts ("tf32x3") must be (re)computed INSIDE the graph
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.tokens = self._run()
        except RuntimeError as e:       # e.g. a launch type the driver cannot capture: replay eagerly instead
            self.graph, self.eager, self.capture_error = None, True, repr(e)
            torch.cuda.synchronize()
            self.tokens = self._run()
        ops._lo_cache.clear()
        return self

    def __call__(self, pr_mat, c):
        """-> (B,32,15,6) int32 tokens on device (static buffer, overwritten by the next call)."""
        if self.graph is None and not self.eager:
            self.capture(pr_mat, c)
        self.pr.copy_(pr_mat, non_blocking=True)
        self.c.copy_(c, non_blocking=True)
        if self.eager:
            self.tokens = self._run()
        else:
            self.graph.replay()
        return self.tokens
