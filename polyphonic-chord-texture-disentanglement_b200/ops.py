"""Host-side operators: thin wrappers that hand device pointers to libpolydis_b200 and the autograd
``Function``s built on them.  PyTorch here is plumbing (allocation, streams, autograd bookkeeping);
every FLOP below runs in the library's kernels.

Conventions: fp32, batch-first; 2-D operands may have an arbitrary row stride but unit inner stride.
"""
import torch

from . import _lib

_call = _lib.call


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _chk(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"polydis_b200: {name} must be a CUDA tensor (there is no CPU path)")
    return t


def _rows(t):
    """View ``t`` as (rows, cols) with unit inner stride; returns (tensor2d, row_stride).  Tensors whose
    leading dimensions collapse to one uniform row stride (e.g. a (...,130) view of a 136-padded buffer)
    are re-strided in place, never copied."""
    if t.dim() != 2:
        if t.dim() > 2 and t.stride(-1) == 1 and not t.is_contiguous():
            sh, st = t.shape, t.stride()
            if all(st[i] == st[i + 1] * sh[i + 1] for i in range(t.dim() - 2)):
                rows = 1
                for d in sh[:-1]:
                    rows *= d
                t = t.as_strided((rows, sh[-1]), (st[-2], 1))
        if t.dim() != 2:
            t = t.reshape(-1, t.shape[-1])
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else t.shape[1])


def _pad4(n):
    return (n + 3) // 4 * 4


def _empty_rows(rows, cols, dev):
    """(rows, cols) fp32 buffer whose row stride is a multiple of 4 floats (TMA-addressable as a GEMM
    operand) -- a view into a padded allocation when cols % 4 != 0."""
    if cols % 4 == 0:
        return torch.empty(rows, cols, device=dev, dtype=torch.float32)
    return torch.empty(rows, _pad4(cols), device=dev, dtype=torch.float32)[:, :cols]


# ------------------------------------------------------------------------------------------------
# Fork-join over CUDA streams.  The two encoders, the two directions of every bi-GRU and the chord decoder
# are independent chains of small kernels (48-192 CTAs each on a 148-SM part); issuing them on side streams
# lets them overlap -- also inside a captured CUDA graph, where the streams become parallel branches, and in
# backward, because autograd replays each op on the stream its forward ran on.
FORK_STREAMS = True
_stream_pool = {}          # parent stream handle -> its side streams


def fork_join(fns, urgent=None):
    """Run the callables in list order (python side effects keep their order); all but the last go to side
    streams forked from the current stream, the last runs on the current stream, then everything is joined.
    ``urgent``: index of the callable on the longest dependency chain -- it gets a side stream of its own with a HIGHER
    priority than the others' (its small kernels then take free SM slots first instead of queueing behind the other
    branches' full waves); ``URGENT_PRIORITY`` = None switches this off.

    Every side stream belongs to ONE parent stream (the pool is keyed by the forking stream, so nested fork-joins under
    different parents never share a side stream).  That is what makes the caching allocator safe here without
    ``record_stream``: a tensor allocated on a side stream is consumed by its parent (or an ancestor) after the join; when
    python frees it the block returns to the side stream's pool, and the side stream's next use starts with a wait on
    that same parent -- i.e. after the consumer.  With a shared side stream a second parent could re-use the block while
    the first parent's consumer was still queued (seen as a wrong first graph replay once buffer sizes lined up)."""
    if not FORK_STREAMS or len(fns) < 2 or not torch.cuda.is_available():
        return [f() for f in fns]
    cur = torch.cuda.current_stream()
    if urgent is not None and URGENT_PRIORITY is None:
        urgent = None
    n_side = len(fns) - 1
    side = _stream_pool.setdefault(cur.cuda_stream, [])
    while len(side) < n_side:
        side.append(torch.cuda.Stream(priority=-1))     # chain streams outrank the weight-gradient stream
    hi = None
    if urgent is not None:
        pool = _stream_pool.setdefault(("urgent", cur.cuda_stream), [])
        if not pool:
            pool.append(torch.cuda.Stream(priority=URGENT_PRIORITY))
        hi = pool[0]
    outs, used, k = [], [], 0
    for i, f in enumerate(fns):
        if i == urgent:
            st = hi
        elif i == len(fns) - 1 or (urgent is not None and k >= n_side):
            st = None
        else:
            st = side[k]
            k += 1
        if st is None:
            outs.append(f())
            continue
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            outs.append(f())
        used.append(st)
    for st in used:
        cur.wait_stream(st)
    return outs


#: issue the decoder's z-independent prologue before the encoders (model.run).  Measured on B200 (GPU call 38): the prologue then
#: starts 0.25 ms earlier (graph nodes start in creation order) but its weight-resident summary GRU (1 CTA per SM for 0.35 ms)
#: holds the encoders back by as much: 7.87 vs 7.78 ms/step.  Off.
PROLOGUE_FIRST = False

#: CUDA priority of the ``urgent`` branch of a fork_join (the other chain streams have -1; None: no special stream)
URGENT_PRIORITY = None       # measured on B200 (GPU call 37): -3 .. 0 all within the run-to-run noise of the step (7.71-7.79 ms)


# ------------------------------------------------------------------------------------------------
# Deferred weight gradients.  Backward's critical path is the chain of INPUT gradients (gate gradients -> dgh.W_hh ->
# next step ...); the weight / bias gradients (split-K GEMMs over all rows of a sequence, column sums, the embedding and
# conv scatter kernels: ~4 ms of a 15 ms step at batch 512) feed nothing but the optimizer.  Issued in line they sit
# between the latency-bound 512-row recurrences of the time GRU and the encoders, which leave most SMs idle.  Here they
# leave the chain: every op that owns such work tags its parameters with ``defer`` right before use; the tag is an
# identity autograd node created under the weight-gradient stream, so autograd replays it (and everything between it
# and the leaf: slices, concatenations, the duration-head fold, AccumulateGrad) on that stream.  The op's backward only
# allocates the gradient buffers and hands the tag a job; the tag's backward launches the job on the weight-gradient
# stream (the engine has made that stream wait for the op's backward).  In a captured step the jobs become parallel
# branches of the graph that join before the clip.  Results are the in-line ones (same kernels, same operands).
DEFER_WGRAD = True
SMALL_GEMM_CFG = 0
BG_GEMM_CFG = 0        # launch configuration of GEMMs issued on the weight-gradient stream (0: the usual heuristic)
DBG_WS_OFF = DBG_PIN_OFF = DBG_INLINE = False
DBG_WS_SKIP = set()
_wg_stream = None
_leaf_pins = []
_warn_off = False
_ws_active = []


#: number of weight-gradient streams the deferred jobs are dealt onto (round robin per tagged op).  The jobs are
#: independent of each other; on ONE in-order stream the last dozen small ones (encoder-side gradients that only exist
#: when backward ends) serialise into a ~1 ms tail after the chain has finished.
DEFER_STREAMS = 4
_wg_extra = []
_wg_next = 0


def wgrad_stream(i=0):
    """The low-priority side streams of the deferred weight-gradient jobs (None without CUDA).  Stream 0 also hosts the
    weight-space nodes (``weight_space``) and the parameters' gradient accumulation (``pin_leaf_streams``)."""
    global _wg_stream
    if not torch.cuda.is_available():
        return None
    if _wg_stream is None:
        _wg_stream = torch.cuda.Stream()
    if i == 0:
        return _wg_stream
    while len(_wg_extra) < i:
        _wg_extra.append(torch.cuda.Stream())
    return _wg_extra[i - 1]


def _on_wgrad_stream():
    if _wg_stream is None:
        return False
    cur = torch.cuda.current_stream()
    return cur == _wg_stream or any(cur == s for s in _wg_extra)


class weight_space:
    """``with ops.weight_space(): w = ...``: build weight-space tensors (column slices, concatenated heads, folded
    products of parameters) on the weight-gradient stream, so their autograd nodes -- which in backward consume deferred
    weight gradients -- run there too instead of making the chain's stream wait.  Joins both ways, so the block is an
    ordinary fork-join for the forward pass.  Mark the results with ``wmark`` to make them deferrable."""

    def __init__(self, site=""):
        self.site = site

    def __enter__(self):
        self.on = (DEFER_WGRAD and not DBG_WS_OFF and self.site not in DBG_WS_SKIP and torch.cuda.is_available() and torch.is_grad_enabled() and not _on_wgrad_stream())
        if self.on:
            self.cur = torch.cuda.current_stream()
            s = wgrad_stream()
            s.wait_stream(self.cur)
            self.ctx = torch.cuda.stream(s)
            self.ctx.__enter__()
            _ws_active.append(self)
        return self

    def __exit__(self, *exc):
        if self.on:
            _ws_active.pop()
            self.ctx.__exit__(*exc)
            self.cur.wait_stream(wgrad_stream())
        return False


def wmark(*ts):
    """Declare tensors built inside ``weight_space`` deferrable (their producers live on the weight-gradient stream)."""
    for t in ts:
        if torch.is_tensor(t):
            t._pd_wspace = True
            if _ws_active and t.is_cuda:
                # allocated on the weight-gradient stream, consumed by the caller's: without this the allocator may hand
                # the block to the next weight-space tensor as soon as python drops the reference (e.g. a merged bias
                # that no backward saves), while the consuming GEMM is still queued on the other stream
                t.record_stream(_ws_active[-1].cur)
    return ts[0] if len(ts) == 1 else ts


def pin_leaf_streams(params):
    """Create the parameters' AccumulateGrad nodes under the weight-gradient stream (a node runs on the stream that was
    current when it was created; they are created on a parameter's first use and die with the graph).  Without this
    the engine makes the stream of a parameter's first forward use -- the chain's -- wait for the deferred job."""
    global _leaf_pins, _wg_next
    _leaf_pins = []
    _wg_next = 0                # the same op -> stream assignment in every forward pass
    s = wgrad_stream()
    if not (DEFER_WGRAD and s is not None and torch.is_grad_enabled()) or _on_wgrad_stream() or DBG_PIN_OFF:
        return
    global _warn_off
    if not _warn_off:       # the stream mismatch between a gradient's producer and AccumulateGrad is the point here
        _warn_off = True
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    with torch.cuda.stream(s):
        _leaf_pins = [p.view_as(p) for p in params if p.requires_grad and p.is_cuda]


class _WGradJob:
    __slots__ = ("job", "stream")

    def __init__(self, stream=None):
        self.job = None
        self.stream = stream

    def set(self, job, keep=()):
        """``job()`` launches the weight-gradient kernels (it runs under the weight-gradient stream); ``keep``: device
        buffers it reads that the calling backward may release before the job has executed."""
        s = self.stream
        if s is not None:
            for t in keep:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(s)
        self.job = job
        if DBG_INLINE:
            self.job = None
            job()


class _DeferTag(torch.autograd.Function):
    """Identity on parameters; its backward runs the consumer's weight-gradient job (see ``defer``)."""

    @staticmethod
    def forward(ctx, wg, *ws):
        ctx.wg = wg
        return tuple(w.view_as(w) for w in ws)

    @staticmethod
    def backward(ctx, *gs):
        job, ctx.wg.job = ctx.wg.job, None
        if job is not None:
            job()
        return (None,) + tuple(gs)


def _deferrable(w):
    return w.is_leaf or getattr(w, "_pd_wspace", False)


def defer(*ws):
    """Tag the parameters of ONE op call for a deferred weight-gradient job.  Returns ``(*tagged, wg)``; ``wg`` is None
    (and ``ws`` come back untouched) when deferral does not apply -- then the op computes its weight gradients in line.
    With a job handle the op's backward must return freshly allocated, UNWRITTEN gradient buffers for these parameters
    and register the job that fills them (``wg.set``)."""
    if not (DEFER_WGRAD and torch.is_grad_enabled()) or _on_wgrad_stream():
        return ws + (None,)
    idx = [i for i, w in enumerate(ws) if torch.is_tensor(w) and w.requires_grad]
    if not idx or not all(_deferrable(ws[i]) for i in idx):
        return ws + (None,)
    global _wg_next
    s = None
    if ws[idx[0]].is_cuda:
        s = wgrad_stream(_wg_next % max(1, DEFER_STREAMS))
        _wg_next += 1
    wg = _WGradJob(s)
    if s is not None:
        with torch.cuda.stream(s):
            tagged = _DeferTag.apply(wg, *[ws[i] for i in idx])
    else:
        tagged = _DeferTag.apply(wg, *[ws[i] for i in idx])
    out = list(ws)
    for i, t in zip(idx, tagged):
        out[i] = t
    return tuple(out) + (wg,)


# ------------------------------------------------------------------------------------------------
# GEMM routing.  "tf32": tcgen05 tensor-core kernel (TF32 multiplies, fp32 accumulate) whenever TMA can
# address the operands, else the fp32 FFMA kernel.  "fp32": always the FFMA kernel -- the fp32-faithful
# arithmetic greedy decoding needs for token parity with the fp32 reference.
#   "tf32x3": error-compensated tensor-core GEMM -- A.B + A.B_lo + A_lo.B with TF32 multiplies and fp32
#             accumulation (lo = x - rn_tf32(x)): ~2^-22 relative operand error, i.e. fp32-class results at
#             tensor-core speed for the inference GEMMs (SURVEY.md 7.4-2 "TF32x3").
PRECISION = "tf32"
# Recurrent GEMM + gate math in one tcgen05 kernel (csrc/gru_step_tc.cu).  Correct (tests/test_gpu_kernels.py) but
# measured SLOWER than GEMM + gate kernel on B200 (note-GRU step 246 vs 148 us cold, tools/gru_step_bench.py): its
# row-per-thread epilogue reads gi / gi2 / h_prev and writes h / r|z|n / hn as 64-byte pieces per lane.  Off until
# the epilogue is staged through shared memory.
FUSED_GRU_STEP = False
# Second-generation fused step (persistent, two TMEM accumulators, all epilogue I/O as TMA boxes, double-buffered
# operand sets; csrc/gru_step_tc.cu gru_step_tma_kernel).  Validated on B200 in round 2: bit-level agreement with GEMM +
# gate kernel (max |dh| 7e-7) and 108 vs 147 us on the note-GRU step (16384 x 512); equal or slower on the 512-row
# recurrences (31.5 vs 29.5 us), hence the row threshold.
FUSED_GRU_STEP_TMA = True
FUSED_GRU_STEP_TMA_MIN_ROWS = 256
# Fold the per-step x-projection into the fused step as a second K segment (pd_gru_step_tmax): for the teacher-forced note
# GRU the (32*B, 16, 1536) projection of the note embeddings (1.6 GB at batch 512) is then never written or read -- the
# producing GEMM skips those columns (linear_split(skip_tail=)), each step multiplies its 128-wide embedding rows itself.
FUSED_GRU_STEP_X = True
# bf16 operands (fp32 accumulate / gate math / state) for the recurrent GEMMs of the batch-sized recurrences in training --
# time GRU, encoder bi-GRUs, chord decoder: their per-step kernels are bound by launch + TMA-pipeline latency, which scales
# with the bytes and k-blocks of the main loop.  BASELINE configs[1] "bf16/fp32-accum"; gradient error of these operands
# alone vs the fp32 reference: 1.0e-3 (tolerance 1e-2; measured on the CPU emulation, all-TF32: 7e-4).
BF16_RECURRENT = True
BF16_RECURRENT_MAX_ROWS = 4096
# tile width of the bf16 fused step for SHORT recurrences (<= BF16_STEP_SHORT_T steps: the encoders' bi-GRUs and the chord
# decoder, which run concurrently on forked streams): 64 = 64 CTAs per step, so two chains fit on the machine side by
# side; 32 = a full wave per step (what the 32-step time GRU, which runs alone, always uses).  Default = the measured-
# faster setting (tools/step_ab.py BF16_STEP_UNITS_SHORT=32 / 64).
BF16_STEP_UNITS_SHORT = 64       # B200, batch 512: 7.86 -> 7.80 ms/step (GPU call 36)
BF16_STEP_SHORT_T = 8


#: greedy decoding: the summary bi-GRU of the predicted notes visits its rows in sorted-length order (pd_gru128_fwd_perm)
GREEDY_SORT_SUMMARY_ROWS = True

#: the plain-TF32 greedy pass (training with tfr < 1) runs its note-GRU slot as one fused launch (ptvae._greedy_fast)
GREEDY_FUSED_TF32_STEP = True


def fold_x_ok(rows, H, x, w_x):
    """Can a recurrence over ``rows`` sequences take its x-projection inside the fused step kernel?  x (rows,T,K2)
    contiguous inputs, w_x (3H,K2) the matching W_ih columns (may be a column slice of a wider matrix)."""
    return (FUSED_GRU_STEP_TMA and FUSED_GRU_STEP_X and PRECISION == "tf32" and torch.is_tensor(x)
            and rows >= FUSED_GRU_STEP_TMA_MIN_ROWS and H % 64 == 0 and x.dim() == 3 and x.is_contiguous()
            and x.shape[2] % 4 == 0 and x.data_ptr() % 16 == 0 and w_x.stride(1) == 1 and w_x.stride(0) % 4 == 0
            and w_x.data_ptr() % 16 == 0 and w_x.shape == (3 * H, x.shape[2]))
_lo_cache = {}          # weight low parts, keyed by (data_ptr, version): static during a decode
TF32X3_MIN_ROWS = 512   # smaller 3xTF32 GEMMs are launch-latency bound: they run on the single-launch FFMA kernel


class precision:
    """``with ops.precision("fp32"): ...`` selects the GEMM arithmetic for the enclosed calls."""

    def __init__(self, mode):
        assert mode in ("tf32", "fp32", "tf32x3")
        self.mode = mode

    def __enter__(self):
        global PRECISION
        self.prev, PRECISION = PRECISION, self.mode
        _lo_cache.clear()       # weight hi/lo splits are only valid within one scope (pointers get reused)

    def __exit__(self, *exc):
        global PRECISION
        PRECISION = self.prev
        _lo_cache.clear()


def _split3(t, order, cache):
    """Operand of the single-launch 3xTF32 GEMM: (rows, 3*pad4(cols)) = [hi | hi | lo] (order 0, activations) or
    [hi | lo | hi] (order 1, weights; cached -- they do not change inside a decode), hi = rn_tf32(t), lo = t - hi."""
    key = (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), order)
    if cache and key in _lo_cache:
        return _lo_cache[key]
    rows, cols = t.shape
    out = torch.empty(rows, 3 * _pad4(cols), device=t.device, dtype=torch.float32)
    _call("pd_tf32_split3", _ptr(t), t.stride(0), rows, cols, _ptr(out), out.stride(0), order, _stream())
    if cache:
        if len(_lo_cache) > 256:
            _lo_cache.clear()
        _lo_cache[key] = out
    return out


def _gemm(a, sam, sak, b, sbk, sbn, out, bias, M, N, K, accumulate, a3=None, rows=None, pred=0):
    if rows is not None:
        # packed note level: slot-major rows with a device live-row table (rows = (cp int32 tensor, slot_rows)); pred 1 =
        # the rows of a / out, 2 = the contraction index.  Tensor-core kernel only (the packed path is TF32 training).
        cp, slot_rows = rows
        _call("pd_gemm_tf32_rows", _ptr(a), sam, sak, _ptr(b), sbk, sbn, _ptr(out), out.stride(0), _ptr(bias), M, N, K,
              int(accumulate), pred, _ptr(cp), slot_rows, _stream())
        return out
    name = "pd_gemm_f32"
    tc_ok = False
    if PRECISION != "fp32" and K >= 8 and N >= 16:
        lda = sak if sak != 1 else sam
        ldb = sbk if sbk != 1 else sbn
        tc_ok = (a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0 and lda % 4 == 0 and ldb % 4 == 0
                 and lda >= 4 and ldb >= 4)
    if (tc_ok and PRECISION == "tf32x3" and M >= TF32X3_MIN_ROWS and sak == 1 and sbk == 1 and a.dim() == 2
            and b.dim() == 2):
        # NT only (the inference GEMMs): the three TF32 products as ONE GEMM over the concatenated K = 3*pad4(K),
        # accumulated in fp32 in TMEM.  Small batches (M < 512) stay on the single-launch FFMA kernel: they are
        # launch-latency bound.
        a3, b3 = (_split3(a, 0, False) if a3 is None else a3), _split3(b, 1, True)
        _call("pd_gemm_tf32", _ptr(a3), a3.stride(0), 1, _ptr(b3), 1, b3.stride(0), _ptr(out), out.stride(0),
              _ptr(bias), M, N, a3.shape[1], int(accumulate), _stream())
        return out
    if tc_ok and PRECISION == "tf32":
        name = "pd_gemm_tf32"
        if SMALL_GEMM_CFG and M <= 1024 and N > 128:
            # tile configuration of the batch-sized (M = batch) recurrent GEMMs (tuning switch; 0 = library heuristic)
            _call("pd_gemm_tf32_cfg", _ptr(a), sam, sak, _ptr(b), sbk, sbn, _ptr(out), out.stride(0), _ptr(bias), M, N, K,
                  int(accumulate), int(SMALL_GEMM_CFG), _stream())
            return out
        if BG_GEMM_CFG and _on_wgrad_stream():
            # deferred weight-gradient GEMM: the "background" launch shape (see csrc/gemm_tc.cu launch())
            _call("pd_gemm_tf32_cfg", _ptr(a), sam, sak, _ptr(b), sbk, sbn, _ptr(out), out.stride(0), _ptr(bias), M, N, K,
                  int(accumulate), int(BG_GEMM_CFG), _stream())
            return out
    _call(name, _ptr(a), sam, sak, _ptr(b), sbk, sbn, _ptr(out), out.stride(0), _ptr(bias), M, N, K,
          int(accumulate), _stream())
    return out


def split3_act(x):
    """The [hi | hi | lo] operand of activations x for the 3xTF32 GEMM path, or None when that path does not apply
    (other precision mode, small batch).  Lets a caller split a tensor ONCE for several GEMMs (``gemm_nt(..., a3=)``)."""
    return _split3(x, 0, False) if split3_applies(x) else None


def gemm_nt(x, w, out, bias=None, accumulate=False, a3=None, rows=None):
    """out (M,N) (+)= x (M,K) @ w (N,K)^T (+ bias).  a3: ``split3_act(x)`` computed by the caller (optional).
    ``rows`` = (cp, slot_rows): x / out rows are slot-major rows of the packed note level; dead row tiles are skipped."""
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and out.shape == (M, N) and x.stride(1) == 1 and out.stride(1) == 1
    assert w.stride(1) == 1 or K == 1
    return _gemm(x, x.stride(0), 1, w, 1, w.stride(0), out, bias, M, N, K, accumulate, a3, rows, 1)


def gemm_nn(x, w, out, accumulate=False, rows=None):
    """out (M,N) (+)= x (M,K) @ w (K,N).  ``rows``: see gemm_nt."""
    M, K = x.shape
    N = w.shape[1]
    assert w.shape[0] == K and out.shape == (M, N) and x.stride(1) == 1 and w.stride(1) == 1
    return _gemm(x, x.stride(0), 1, w, w.stride(0), 1, out, None, M, N, K, accumulate, None, rows, 1)


def gemm_tn(a, b, out, accumulate=False, rows=None):
    """out (M,N) (+)= a (R,M)^T @ b (R,N).  ``rows`` = (cp, slot_rows): the R rows are slot-major rows of the packed note
    level; only live 32-row blocks are accumulated."""
    R, M = a.shape
    N = b.shape[1]
    assert b.shape[0] == R and out.shape == (M, N) and a.stride(1) == 1 and b.stride(1) == 1
    return _gemm(a, 1, a.stride(0), b, b.stride(0), 1, out, None, M, N, R, accumulate, None, rows, 2)


def colsum(x, out, accumulate=False, rows=None):
    M, N = x.shape
    if rows is not None:
        _call("pd_colsum_rows_f32", _ptr(x), x.stride(0), M, N, _ptr(out), int(accumulate), _ptr(rows[0]), rows[1], _stream())
        return out
    _call("pd_colsum_f32", _ptr(x), x.stride(0), M, N, _ptr(out), int(accumulate), _stream())
    return out


def sum_steps(x, T):
    """(R,T',C) strided -> (R,C): sum over the first T steps."""
    R, _, C = x.shape
    out = torch.empty(R, C, device=x.device, dtype=torch.float32)
    _call("pd_sum_steps_f32", _ptr(x), x.stride(0), x.stride(1), T, _ptr(out), out.stride(0), R, C, _stream())
    return out


def transpose(x):
    x = x.contiguous()
    out = torch.empty(x.shape[1], x.shape[0], device=x.device, dtype=x.dtype)
    _call("pd_transpose_f32", _ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    """y = x W^T + b on 2-D x with strided rows (nn.Linear; K7 of SURVEY.md 2.2)."""

    @staticmethod
    def forward(ctx, x, w, b, wg=None):
        x2, _ = _rows(_chk(x, "x"))
        y = _empty_rows(x2.shape[0], w.shape[0], x.device)
        K = x2.shape[1]
        ctx.k = K
        ctx.wg = wg
        if PRECISION == "tf32" and K % 4 and K >= 64 and x2.shape[0] >= 256:
            # an input width TMA cannot address (row pitch not a multiple of 16 bytes: fc1 of the texture encoder,
            # 290 -> 1000) would send all three GEMMs of the layer to the FFMA kernel; zero-padded copies of the two
            # small operands keep them on the tensor cores
            xp = torch.zeros(x2.shape[0], _pad4(K), device=x.device, dtype=torch.float32)
            wp = torch.zeros(w.shape[0], _pad4(K), device=x.device, dtype=torch.float32)
            xp[:, :K].copy_(x2)
            wp[:, :K].copy_(w)
            x2, w = xp, wp
        gemm_nt(x2, w, y, b)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.x_shape = x.shape
        if y.is_contiguous():
            return y.view(*x.shape[:-1], w.shape[0])
        lead = tuple(x.shape[:-1])
        strides, acc = [], y.stride(0)
        for d in reversed(lead):
            strides.append(acc)
            acc *= d
        return y.as_strided(lead + (w.shape[0],), tuple(reversed(strides)) + (1,))

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2, _ = _rows(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x2.shape, device=dy.device, dtype=torch.float32)
            gemm_nn(dy2, w, dx)
            dx = dx[:, :ctx.k].reshape(ctx.x_shape)           # (drops the zero-padding columns, if any)
        k = ctx.k
        if ctx.wg is not None:
            # weight / bias gradients leave the chain: buffers now, kernels on the weight-gradient stream (ops.defer)
            if ctx.needs_input_grad[1]:
                dw = torch.empty(w.shape[0], k, device=dy.device, dtype=torch.float32)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = torch.empty(w.shape[0], device=dy.device, dtype=torch.float32)

            def job():
                if dw is not None:
                    if k != w.shape[1]:
                        tmp = torch.empty(w.shape, device=dw.device, dtype=torch.float32)
                        gemm_tn(dy2, x2, tmp)
                        dw.copy_(tmp[:, :k])
                    else:
                        gemm_tn(dy2, x2, dw)
                if db is not None:
                    colsum(dy2, db)
            ctx.wg.set(job, keep=(dy2, x2))
            return dx, dw, db, None
        if ctx.needs_input_grad[1]:
            dw = torch.empty(w.shape, device=dy.device, dtype=torch.float32)
            gemm_tn(dy2, x2, dw)
            if k != w.shape[1]:
                dw = dw[:, :k].contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(w.shape[0], device=dy.device, dtype=torch.float32)
            colsum(dy2, db)
        return dx, dw, db, None


DEFER_MIN_ROWS = 256       # smaller layers keep their weight gradients in line (a tag node costs more than it hides)


def _n_rows(x):
    return x.numel() // max(1, x.shape[-1])


def linear(x, w, b=None):
    wg = None
    if _n_rows(x) >= DEFER_MIN_ROWS:
        w, b, wg = defer(w, b)
    return _Linear.apply(x, w, b, wg)


class GradSlab:
    """The (M, N) gradient matrix of a multi-head projection (``linear_split``), allocated by whichever backward
    touches it first.  Consumers that know their input came from ``linear_split`` (the GRU recurrences) write their
    input gradient straight into their column block, so the projection's backward needs no gather copies and the
    heads' input gradients never exist separately (no autograd add over the shared activations)."""

    def __init__(self):
        self.M = self.N = self.dev = self.buf = None

    def get(self):
        if self.buf is None:
            self.buf = _empty_rows(self.M, self.N, self.dev)
        return self.buf

    def block(self, off, n, lead, slot_major=False):
        """Column block [off, off+n) viewed as (*lead, n) (rows = prod(lead)).  ``slot_major``: lead = (B, T) over rows
        stored (t, b)-major (the packed note level): the view is (B, T, n) with strides (ld, B*ld, 1)."""
        part = self.get()[:, off:off + n]
        if slot_major:
            B, T = lead
            return part.as_strided((B, T, n), (part.stride(0), B * part.stride(0), 1))
        strides, acc = [], part.stride(0)
        for d in reversed(lead):
            strides.append(acc)
            acc *= d
        return part.as_strided(tuple(lead) + (n,), tuple(reversed(strides)) + (1,))


class _LinearSplit(torch.autograd.Function):
    """(y1, y2, ...) = split(x W^T + b, sizes): several heads that read the same activations as ONE GEMM.
    E.g. the pitch head and the (folded) duration-hidden projection both consume the note-GRU states
    (ptvae.py:336-343); as separate Linears their backward writes two (Q,512) input gradients that autograd then
    adds (1.5 GB of traffic at B = 512) and reads the states twice for the two weight gradients.  Likewise the note
    embeddings feed both directions of the summary bi-GRU and the note GRU (ptvae.py:446-453, :396)."""

    @staticmethod
    def forward(ctx, x, w, b, sizes, bias_cols, slab, skip_tail=0, wg=None, rows=None):
        ctx.wg = wg
        ctx.rows = rows         # packed note level: x rows are slot-major with a device live-row table (dead rows: no work)
        x2, _ = _rows(_chk(x, "x"))
        y = _empty_rows(x2.shape[0], w.shape[0], x.device)
        if rows is not None:
            assert not skip_tail
            gemm_nt(x2, w, y, b, rows=rows)
        elif skip_tail:
            # the last head is NOT computed (its consumer multiplies x itself, ops.fold_x_ok); its columns of y stay
            # uninitialised and only route the gradient: the backward still covers all heads
            n = w.shape[0] - skip_tail
            gemm_nt(x2, w[:n], y[:, :n], None if b is None else b[:n])
        else:
            gemm_nt(x2, w, y, b)
        ctx.save_for_backward(x2, w)
        ctx.sizes, ctx.bias_cols, ctx.slab, ctx.x_shape = sizes, bias_cols, slab, x.shape
        slab.M, slab.N, slab.dev = x2.shape[0], w.shape[0], x.device
        lead = tuple(x.shape[:-1])
        outs, off = [], 0
        for n in sizes:
            part = y[:, off:off + n]
            if len(lead) > 1:
                strides, acc = [], part.stride(0)
                for d in reversed(lead):
                    strides.append(acc)
                    acc *= d
                part = part.as_strided(lead + (n,), tuple(reversed(strides)) + (1,))
            outs.append(part)
            off += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *ds):
        x2, w = ctx.saved_tensors
        dy = ctx.slab.get()
        ctx.slab.buf = None                                   # the slab is consumed by this backward
        off = 0
        for n, d in zip(ctx.sizes, ds):
            part = dy[:, off:off + n]
            off += n
            if d is None:
                part.zero_()
            elif not (d.data_ptr() == part.data_ptr() and d.stride(-1) == 1 and d.stride(-2) == part.stride(0)):
                part.copy_(d.reshape(part.shape))             # a consumer that did not write into the slab
        dx = dw = db = None
        rows = ctx.rows
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x2.shape, device=dy.device, dtype=torch.float32)
            gemm_nn(dy, w, dx, rows=rows)                       # (packed: dead rows of dx stay unwritten)
            dx = dx.view(ctx.x_shape)
        nb = ctx.bias_cols
        if ctx.wg is not None:
            if ctx.needs_input_grad[1]:
                dw = torch.empty(w.shape, device=dy.device, dtype=torch.float32)
            if ctx.needs_input_grad[2]:
                db = torch.empty(w.shape[0], device=dy.device, dtype=torch.float32)

            def job():
                if dw is not None:
                    gemm_tn(dy, x2, dw, rows=rows)
                if db is not None:
                    db.zero_()
                    colsum(dy[:, :nb], db[:nb], rows=rows)
            ctx.wg.set(job, keep=(dy, x2))
            return dx, dw, db, None, None, None, None, None, None
        if ctx.needs_input_grad[1]:
            dw = torch.empty(w.shape, device=dy.device, dtype=torch.float32)
            gemm_tn(dy, x2, dw, rows=rows)
        if ctx.needs_input_grad[2]:
            db = torch.zeros(w.shape[0], device=dy.device, dtype=torch.float32)
            colsum(dy[:, :nb], db[:nb], rows=rows)
        return dx, dw, db, None, None, None, None, None, None


def linear_split(x, w, b, sizes, bias_cols=None, skip_tail=False, rows=None):
    """Heads of widths ``sizes`` over the same input as one GEMM; outputs are shaped (*x.shape[:-1], n).  Only the
    first ``bias_cols`` columns carry a trainable bias (default: all).  ``skip_tail``: do not compute the LAST head in the
    forward pass (its tensor is returned uninitialised, for gradient routing only; it must carry no bias)."""
    if isinstance(sizes, int):
        sizes = (sizes, w.shape[0] - sizes)
    sizes = tuple(sizes)
    assert sum(sizes) == w.shape[0]
    slab = GradSlab()
    nb = w.shape[0] if bias_cols is None else bias_cols
    assert not skip_tail or nb <= w.shape[0] - sizes[-1]
    wg = None
    if _n_rows(x) >= DEFER_MIN_ROWS:
        w, b, wg = defer(w, b)
    outs = _LinearSplit.apply(x, w, b, sizes, nb, slab, sizes[-1] if skip_tail else 0, wg, rows)
    off = 0
    for o, n in zip(outs, sizes):
        o._pd_slab = (slab, off, n)                           # lets slab-aware consumers write their gradient in place
        off += n
    return outs


class _MatMulNN(torch.autograd.Function):
    """c = a @ b for small weight-space products (a (M,K), b (K,N), both row-strided)."""

    @staticmethod
    def forward(ctx, a, b):
        c = torch.empty(a.shape[0], b.shape[1], device=a.device, dtype=torch.float32)
        gemm_nn(a, b, c)
        ctx.save_for_backward(a, b)
        return c

    @staticmethod
    def backward(ctx, dc):
        a, b = ctx.saved_tensors
        dc = dc.contiguous()
        da = torch.empty(a.shape, device=dc.device, dtype=torch.float32)
        db = torch.empty(b.shape, device=dc.device, dtype=torch.float32)
        gemm_nt(dc, b, da)             # da = dc b^T
        gemm_tn(a, dc, db)             # db = a^T dc
        return da, db


def matmul_nn(a, b):
    return _MatMulNN.apply(a, b)


# ------------------------------------------------------------------------------------------------
def _gates_fwd(gi, gi2, gh, hprev, hout, rzn, hn, lengths, t):
    B, H = hout.shape
    _call("pd_gru_gates_fwd", _ptr(gi), gi.stride(0), _ptr(gi2), 0 if gi2 is None else gi2.stride(0),
          _ptr(gh), gh.stride(0), _ptr(hprev), 0 if hprev is None else hprev.stride(0),
          _ptr(hout), hout.stride(0), _ptr(rzn), 0 if rzn is None else rzn.stride(0),
          _ptr(hn), 0 if hn is None else hn.stride(0), _ptr(lengths), t, B, H, _stream())


def gates_fwd_split3(gi, gi2, gh, h, h3):
    """In-place inference GRU step on h (B,H) that also emits h3 = [hi | hi | lo] of the new state (3xTF32 operand)."""
    B, H = h.shape
    _call("pd_gru_gates_fwd_split3", _ptr(gi), gi.stride(0), _ptr(gi2), 0 if gi2 is None else gi2.stride(0), _ptr(gh),
          gh.stride(0), _ptr(h), h.stride(0), _ptr(h), h.stride(0), None, 0, B, H, _ptr(h3), h3.stride(0), _stream())


# Greedy decode at >= 512 rows (tf32x3): recurrent 3xTF32 GEMM + gate math + split of the new state in ONE tcgen05 kernel
# (pd_gru_step_tma3) instead of GEMM + pd_gru_gates_fwd_split3 -- the (B,3H) h-projection never goes to HBM.
# Validated on B200 (kernel test + 1,024-segment token parity 0.9997 vs the CPU oracle) but NOT faster: with K = 3H = 1536 the
# step is bound by L2 -> SM operand traffic (2 GB per note slot at 16,384 segments), where the 128 x 192 fused tile re-reads
# the A operand 8x against 6x for the 256-wide GEMM tile: 213 us fused vs 138 us GEMM + 70 us gate kernel per slot
# (profiles/r02_decode16384_launches_*.txt), 266 vs 258 ms per 16,384-segment decode.  Off by default.
FUSED_DECODE_STEP = False


def gru_step_split3(a3, w3, b_hh, gi, gi2, h, h3out):
    """In-place inference GRU step on h (B,H): a3 = [hi|hi|lo] of h (as produced by the previous step), w3 = [hi|lo|hi] of
    W_hh (``weight_split3``); writes the new h and its split into h3out (a different buffer than a3)."""
    B, H = h.shape
    _call("pd_gru_step_tma3", _ptr(a3), a3.stride(0), _ptr(w3), w3.stride(0), _ptr(b_hh), _ptr(gi), gi.stride(0),
          _ptr(gi2), 0 if gi2 is None else gi2.stride(0), _ptr(h), h.stride(0), _ptr(h), h.stride(0), _ptr(h3out),
          h3out.stride(0), B, H, _stream())


def gru_step_split3x(a3, w3, x3, wx3, b_hh, gi2, h, h3out):
    """``gru_step_split3`` with the x-projection computed inside the kernel: x3 = [hi|hi|lo] of the step's input rows,
    wx3 = [hi|lo|hi] of the matching W_ih columns; gi2 carries the rest of the input projection incl. b_ih."""
    B, H = h.shape
    _call("pd_gru_step_tma3x", _ptr(a3), a3.stride(0), _ptr(w3), w3.stride(0), _ptr(x3), x3.stride(0), _ptr(wx3),
          wx3.stride(0), x3.shape[1], _ptr(b_hh), _ptr(gi2), gi2.stride(0), _ptr(h), h.stride(0), _ptr(h), h.stride(0),
          _ptr(h3out), h3out.stride(0), B, H, _stream())


def split3_into(x, out):
    """[hi | hi | lo] split of activations x (rows, cols) into a preallocated (rows, 3*pad4(cols)) buffer."""
    _call("pd_tf32_split3", _ptr(x), x.stride(0), x.shape[0], x.shape[1], _ptr(out), out.stride(0), 0, _stream())
    return out


# fold the note-embedding projection of the greedy decode into the fused step (second K segment): saves the gi GEMM launch
# and 200 MB of x-projection traffic per note slot at 16,384 segments
FUSED_DECODE_STEP_X = True


def weight_split3(w):
    """[hi | lo | hi] operand of a weight matrix for the 3xTF32 GEMMs (cached within the precision scope)."""
    return _split3(w, 1, True)


def fused_decode_step_ok(h, w_hh):
    return (FUSED_DECODE_STEP and PRECISION == "tf32x3" and h.dim() == 2 and h.shape[0] >= TF32X3_MIN_ROWS
            and h.shape[1] % 64 == 0 and w_hh.shape == (3 * h.shape[1], h.shape[1]) and h.stride(1) == 1
            and h.stride(0) % 4 == 0 and h.data_ptr() % 16 == 0)


def split3_applies(x):
    """Would ``split3_act(x)`` return an operand (3xTF32 tensor-core GEMM path) for this tensor?"""
    return (PRECISION == "tf32x3" and x.dim() == 2 and x.shape[0] >= TF32X3_MIN_ROWS and x.stride(1) == 1 and x.shape[1] >= 8
            and x.data_ptr() % 16 == 0 and x.stride(0) % 4 == 0 and x.stride(0) >= 4)


# Weight-resident GRU128 kernel (csrc/gru128_resident.cu) for the note-summary bi-GRU.  Its matvecs run on
# mma.sync, which B200 issues at about FFMA rate, so per row it only matches the tcgen05 GEMM + gate kernels --
# but it is ONE launch instead of 32 (fwd) / 48 (bwd) and it occupies every SM.  Measured (B200, 16 steps):
# 16384 sequences 423 vs 461 us forward, equal backward, but the teacher-forced step loses 0.3 ms because the
# two directions no longer overlap the rest of the step; 512 sequences (free-running training, small-batch
# decode) and the 3-pass tf32x3 decode win 6-12 %.  Hence the row limit in TF32 mode.
RESIDENT_GRU128 = True
RESIDENT_GRU128_MAX_ROWS_TF32 = 4096
# Length-SORTED rows (the packed note level): the 16 sequences of a CTA then have (nearly) the same length, the kernel's
# loop stops at the tile's longest sequence, so its work is sum(lengths) instead of ~rows x max -- it then wins at any size
RESIDENT_GRU128_SORTED = True
_rows_sorted = False


class rows_sorted:
    """``with ops.rows_sorted():`` the sequences handed to ``gru_sequence`` inside are sorted by length."""

    def __enter__(self):
        global _rows_sorted
        self.prev, _rows_sorted = _rows_sorted, True

    def __exit__(self, *exc):
        global _rows_sorted
        _rows_sorted = self.prev
        return False


def _resident128_ok(gi, gi2, h0, lengths, H):
    return (RESIDENT_GRU128 and H == 128 and lengths is not None and h0 is None and gi2 is None
            and (PRECISION == "tf32x3" or (PRECISION == "tf32" and (gi.shape[0] <= RESIDENT_GRU128_MAX_ROWS_TF32
                                                                     or (_rows_sorted and RESIDENT_GRU128_SORTED))))
            and gi.stride(2) == 1 and gi.stride(0) % 2 == 0
            and gi.stride(1) % 2 == 0 and gi.data_ptr() % 8 == 0)


def gru_sequence_nograd(gi, gi2, h0, w_hh, b_hh, lengths=None, reverse=False, save=None, n_steps=None, xsrc=None):
    """Run a GRU over precomputed input projections.  gi (B,T,3H) strided view, gi2 (B,3H) or None;
    ``n_steps`` (<= T) uses only the first slots of gi (the note GRU consumes 15 of the 16 embedded slots).
    ``xsrc`` = (x (B,T,K2), w_x (3H,K2)): gi is NOT read (it may be uninitialised); every step computes W_x x[:, t] inside
    the fused step kernel (``fold_x_ok`` must hold; gi2 carries the rest of the input projection incl. the bias).

    Returns h_all (B,n_steps,H).  ``save`` (dict) receives rzn / hn for the backward pass.
    """
    B, T, H3 = gi.shape
    T = T if n_steps is None else n_steps
    H = H3 // 3
    dev = gi.device
    h_all = torch.empty(B, T, H, device=dev, dtype=torch.float32)
    if xsrc is not None:
        x, w_x = xsrc
        assert gi2 is not None and h0 is not None and lengths is None and fold_x_ok(B, H, x, w_x)
        rzn = hn = None
        if save is not None:
            rzn = torch.empty(B, T, H3, device=dev, dtype=torch.float32)
            hn = torch.empty(B, T, H, device=dev, dtype=torch.float32)
            save["rzn"], save["hn"] = rzn, hn
        hprev = h0
        for t in (range(T - 1, -1, -1) if reverse else range(T)):
            xt = x[:, t]
            _call("pd_gru_step_tmax", _ptr(hprev), hprev.stride(0), _ptr(w_hh), w_hh.stride(0), _ptr(xt), x.stride(0),
                  _ptr(w_x), w_x.stride(0), x.shape[2], _ptr(b_hh), _ptr(gi2), gi2.stride(0), _ptr(h_all[:, t]),
                  h_all.stride(0), None if rzn is None else _ptr(rzn[:, t]), 0 if rzn is None else rzn.stride(0),
                  None if hn is None else _ptr(hn[:, t]), 0 if hn is None else hn.stride(0), B, H, _stream())
            hprev = h_all[:, t]
        return h_all
    if _resident128_ok(gi, gi2, h0, lengths, H):
        # weight-resident kernel: whole variable-length recurrence in one launch (csrc/gru128_resident.cu)
        rzn = hn = None
        if save is not None:
            rzn = torch.empty(B, T, H3, device=dev, dtype=torch.float32)
            hn = torch.empty(B, T, H, device=dev, dtype=torch.float32)
            save["rzn"], save["hn"], save["resident"] = rzn, hn, True
        _call("pd_gru128_fwd", _ptr(gi), gi.stride(0), gi.stride(1), _ptr(lengths), _ptr(w_hh), _ptr(b_hh),
              _ptr(h_all), h_all.stride(0), h_all.stride(1), _ptr(rzn), 0 if rzn is None else rzn.stride(0),
              0 if rzn is None else rzn.stride(1), _ptr(hn), 0 if hn is None else hn.stride(0),
              0 if hn is None else hn.stride(1), B, T, int(reverse), 3 if PRECISION == "tf32x3" else 1, _stream())
        return h_all
    rzn = hn = None
    if save is not None:
        rzn = torch.empty(B, T, H3, device=dev, dtype=torch.float32)
        hn = torch.empty(B, T, H, device=dev, dtype=torch.float32)
        save["rzn"], save["hn"] = rzn, hn
    order = list(range(T - 1, -1, -1) if reverse else range(T))
    # fused step (recurrent GEMM + gate math in one tcgen05 kernel) whenever TMA can address the operands
    fused_ok = ((FUSED_GRU_STEP or (FUSED_GRU_STEP_TMA and lengths is None and B >= FUSED_GRU_STEP_TMA_MIN_ROWS))
                and PRECISION == "tf32" and H % 64 == 0
                and w_hh.stride(1) == 1
                and w_hh.stride(0) % 4 == 0 and w_hh.data_ptr() % 16 == 0 and b_hh.data_ptr() % 16 == 0
                and gi.stride(2) == 1 and gi.stride(0) % 4 == 0 and gi.stride(1) % 4 == 0 and gi.data_ptr() % 16 == 0
                and (gi2 is None or (gi2.stride(1) == 1 and gi2.stride(0) % 4 == 0 and gi2.data_ptr() % 16 == 0)))

    def run(sl):
        _gru_steps_fwd(gi[sl], _sl(gi2, sl), _sl(h0, sl), w_hh, b_hh, _sl(lengths, sl), order, h_all[sl], _sl(rzn, sl),
                       _sl(hn, sl), fused_ok, save)
    _over_row_chunks(B, H3, run)
    return h_all


def _sl(t, sl):
    return None if t is None else t[sl]


# Optional row chunking of the big recurrences (OFF: measured slower).  The note GRU runs 15 steps over 32*B
# independent sequences; at B = 512 one step's h-projection is 100 MB, so step-major order streams ~630 MB per
# step through HBM.  Chunk-major order (all steps of a row chunk before the next chunk, a few chunks side by side on
# forked streams) was meant to keep a chunk's gh / gi2 / state in the 126 MB L2 across steps.  B200, graph-replayed
# step at B = 512 (tools/chunk_sweep.py): unchunked 17.78 ms; 50 MB chunks x 2 lanes 17.84; 25 MB x 3 18.18;
# 12 MB x 3 19.21; 6 MB x 4 19.73 -- the smaller GEMM / gate launches lose more than the L2 hits return.
ROW_CHUNK_BYTES = 0               # target size of one chunk's (rows, 3H) fp32 slab; 0 = never chunk
ROW_CHUNK_MIN_BYTES = 48 << 20    # recurrences whose slab is smaller than this run unchunked
ROW_CHUNK_LANES = 2


def _over_row_chunks(B, H3, run):
    """Call run(slice) over row chunks (concurrently on ROW_CHUNK_LANES forked streams) or once over all rows."""
    if not FORK_STREAMS or not ROW_CHUNK_BYTES or B * H3 * 4 < ROW_CHUNK_MIN_BYTES:
        run(slice(0, B))
        return
    rows = max(128, ROW_CHUNK_BYTES // (H3 * 4) // 128 * 128)
    starts = list(range(0, B, rows))

    def lane(l):
        def go():
            for s0 in starts[l::ROW_CHUNK_LANES]:
                run(slice(s0, min(B, s0 + rows)))
        return go
    fork_join([lane(l) for l in range(min(ROW_CHUNK_LANES, len(starts)))])


def to_bf16(x):
    """fp32 (rows, cols) with unit inner stride -> bf16 copy (round to nearest even)."""
    rows, cols = x.shape
    out = torch.empty(rows, cols, device=x.device, dtype=torch.bfloat16)
    _call("pd_f32_to_bf16", _ptr(x), x.stride(0), rows, cols, _ptr(out), out.stride(0), _stream())
    return out


def _gru_steps_fwd(gi, gi2, h0, w_hh, b_hh, lengths, order, h_all, rzn, hn, fused_ok, save=None):
    """All steps of the recurrence for one block of rows (every argument already row-sliced)."""
    B, H = h_all.shape[0], h_all.shape[2]
    gh = torch.empty(B, 3 * H, device=gi.device, dtype=torch.float32)
    hprev = h0
    # bf16 copies of W_hh and of the running state for the batch-sized recurrences in training (BF16_RECURRENT)
    use_bf16 = (fused_ok and BF16_RECURRENT and save is not None and rzn is not None and lengths is None
                and FUSED_GRU_STEP_TMA and FUSED_GRU_STEP_TMA_MIN_ROWS <= B <= BF16_RECURRENT_MAX_ROWS and H % 64 == 0
                and w_hh.stride(1) == 1)
    wb = hb = None
    if use_bf16:
        wb = to_bf16(w_hh)
        save["wb"] = wb                                   # the backward's dgh . W_hh GEMM multiplies the same copy
        hb = torch.empty(2, B, H, device=gi.device, dtype=torch.bfloat16)      # ping-pong: state in / state out
        hb_cur = 0
        hb_valid = False
    for t in order:
        if (fused_ok and hprev is not None and hprev.stride(1) == 1 and hprev.stride(0) % 4 == 0
                and hprev.data_ptr() % 16 == 0):
            if use_bf16:
                if not hb_valid:                          # first fused step: the incoming state has no bf16 copy yet
                    _call("pd_f32_to_bf16", _ptr(hprev), hprev.stride(0), B, H, _ptr(hb[hb_cur]), H, _stream())
                # short recurrences (the encoders' bi-GRUs, the chord decoder) run next to each other on forked streams:
                # 64-unit tiles (64 CTAs per step) let two of them share the machine; the 32-step time GRU runs alone
                units = 64 if (BF16_STEP_UNITS_SHORT == 64 and len(order) <= BF16_STEP_SHORT_T) else 32
                _call("pd_gru_step_tma_bf16_units", _ptr(hb[hb_cur]), H, _ptr(wb), wb.stride(0), _ptr(b_hh), _ptr(gi[:, t]),
                      gi.stride(0), _ptr(gi2), 0 if gi2 is None else gi2.stride(0), _ptr(hprev), hprev.stride(0),
                      _ptr(h_all[:, t]), h_all.stride(0), _ptr(hb[hb_cur ^ 1]), H, _ptr(rzn[:, t]), rzn.stride(0),
                      _ptr(hn[:, t]), hn.stride(0), B, H, units, _stream())
                hb_cur ^= 1
                hb_valid = True
                hprev = h_all[:, t]
                continue
            if FUSED_GRU_STEP_TMA and lengths is None and B >= FUSED_GRU_STEP_TMA_MIN_ROWS:
                _call("pd_gru_step_tma", _ptr(hprev), hprev.stride(0), _ptr(w_hh), w_hh.stride(0), _ptr(b_hh),
                      _ptr(gi[:, t]), gi.stride(0), _ptr(gi2), 0 if gi2 is None else gi2.stride(0),
                      _ptr(h_all[:, t]), h_all.stride(0), None if rzn is None else _ptr(rzn[:, t]),
                      0 if rzn is None else rzn.stride(0), None if hn is None else _ptr(hn[:, t]),
                      0 if hn is None else hn.stride(0), B, H, _stream())
                hprev = h_all[:, t]
                continue
            _call("pd_gru_step_tf32", _ptr(hprev), hprev.stride(0), _ptr(w_hh), w_hh.stride(0), _ptr(b_hh),
                  _ptr(gi[:, t]), gi.stride(0), _ptr(gi2), 0 if gi2 is None else gi2.stride(0),
                  _ptr(h_all[:, t]), h_all.stride(0), None if rzn is None else _ptr(rzn[:, t]),
                  0 if rzn is None else rzn.stride(0), None if hn is None else _ptr(hn[:, t]),
                  0 if hn is None else hn.stride(0), _ptr(lengths), t, B, H, _stream())
            hprev = h_all[:, t]
            continue
        if hprev is None:
            gemm_nt(gh[:, :0], w_hh[:, :0], gh, b_hh)          # K = 0: gh = b_hh
        else:
            gemm_nt(hprev, w_hh, gh, b_hh)
        _gates_fwd(gi[:, t], gi2, gh, hprev, h_all[:, t],
                   None if rzn is None else rzn[:, t], None if hn is None else hn[:, t], lengths, t)
        hprev = h_all[:, t]


def _gru_steps_bwd(dout, rzn, hn, h_all, h0, w_hh, lengths, order, dgi, dgh, dh0, wb=None):
    """BPTT over all steps for one block of rows; dh0 (or None) receives the gradient of the initial state.
    ``dout``: (B,T,H) gradient of every step's output, or (d (B,H), step): only that step's output was used."""
    B, T, H = h_all.shape
    dev = h_all.device
    final = dout if isinstance(dout, tuple) else None

    def dout_of(t):
        if final is None:
            return dout[:, t], dout.stride(0)
        return (final[0], final[0].stride(0)) if t == final[1] else (None, 0)
    # the recurrent gradient arrives in two pieces: dh*z (written by the gate kernel of the later step)
    # and dgh @ W_hh (written by that step's GEMM); the next gate kernel sums both with dout[:, t]
    bufs = torch.empty(4, B, H, device=dev, dtype=torch.float32)
    dz_a, dz_b, dm_a, dm_b = bufs[0], bufs[1], bufs[2], bufs[3]
    dz = dm = None
    st = _stream()
    # batch-sized recurrences run their dgh.W_hh GEMM split over K (zero-fill + red.add epilogue): the gate kernel of the
    # step clears the accumulator instead, one graph node less on every serial step
    fold_zero = (PRECISION == "tf32" and w_hh.stride(1) == 1 and w_hh.stride(0) % 4 == 0 and w_hh.data_ptr() % 16 == 0
                 and _lib.lib.pd_gemm_tf32_splits(B, H, 3 * H) > 1)
    dghb = torch.empty(B, 3 * H, device=dev, dtype=torch.bfloat16) if (wb is not None and fold_zero) else None
    for i in range(T - 1, -1, -1):
        t = order[i]
        hprev = h_all[:, order[i - 1]] if i > 0 else h0
        nz = dz_b if dz is dz_a else dz_a
        nm = dm_b if dm is dm_a else dm_a
        d_t, ld_t = dout_of(t)
        if dghb is not None and hprev is not None:
            # bf16 operands for dgh . W_hh (the forward multiplied the same bf16 W_hh): the gate kernel emits the bf16 copy
            _call("pd_gru_gates_bwd_zb", _ptr(dz), 0 if dz is None else dz.stride(0), _ptr(d_t),
                  ld_t, _ptr(dm), 0 if dm is None else dm.stride(0), _ptr(rzn[:, t]), rzn.stride(0),
                  _ptr(hn[:, t]), hn.stride(0), _ptr(hprev), hprev.stride(0),
                  _ptr(dgi[:, t]), dgi.stride(0), _ptr(dgh[:, t]), dgh.stride(0), _ptr(nz), nz.stride(0),
                  _ptr(lengths), t, B, H, _ptr(nm), nm.stride(0), _ptr(dghb), 3 * H, st)
            _call("pd_gemm_bf16", _ptr(dghb), 3 * H, 1, _ptr(wb), wb.stride(0), 1, _ptr(nm), nm.stride(0), None, B, H, 3 * H, 1, st)
            dz, dm = nz, nm
            continue
        if fold_zero and hprev is not None:
            _call("pd_gru_gates_bwd_z", _ptr(dz), 0 if dz is None else dz.stride(0), _ptr(d_t),
                  ld_t, _ptr(dm), 0 if dm is None else dm.stride(0), _ptr(rzn[:, t]), rzn.stride(0),
                  _ptr(hn[:, t]), hn.stride(0), _ptr(hprev), hprev.stride(0),
                  _ptr(dgi[:, t]), dgi.stride(0), _ptr(dgh[:, t]), dgh.stride(0), _ptr(nz), nz.stride(0),
                  _ptr(lengths), t, B, H, _ptr(nm), nm.stride(0), st)
            gemm_nn(dgh[:, t], w_hh, nm, accumulate=True)
            dz, dm = nz, nm
            continue
        _call("pd_gru_gates_bwd", _ptr(dz), 0 if dz is None else dz.stride(0), _ptr(d_t),
              ld_t, _ptr(dm), 0 if dm is None else dm.stride(0), _ptr(rzn[:, t]), rzn.stride(0),
              _ptr(hn[:, t]), hn.stride(0), _ptr(hprev), 0 if hprev is None else hprev.stride(0),
              _ptr(dgi[:, t]), dgi.stride(0), _ptr(dgh[:, t]), dgh.stride(0), _ptr(nz), nz.stride(0),
              None, 0, _ptr(lengths), t, B, H, st)
        dz = nz
        if hprev is not None:
            gemm_nn(dgh[:, t], w_hh, nm)                          # dgh W_hh
            dm = nm
        else:
            dm = None
    if dh0 is not None:
        if dm is None:
            dh0.copy_(dz)
        else:
            _call("pd_add_f32", _ptr(dz), _ptr(dm), dz.numel(), _ptr(dh0), st)


class _GruSeq(torch.autograd.Function):
    """GRU recurrence over precomputed x-projections; BPTT in the backward.  Serves the time / note /
    duration / chord GRUs and (with ``lengths``) the packed note-summary bi-GRU."""

    @staticmethod
    def forward(ctx, gi, gi2, h0, w_hh, b_hh, lengths, reverse, n_steps=None, slab=None, xsrc=None, wg=None, final_only=False):
        _chk(gi, "gi")
        ctx.slab = slab
        ctx.wg = wg
        ctx.final_only = final_only      # return only the state after the last processed step (a summariser's output):
        #                                  its backward then takes an (B,H) gradient -- no (B,T,H) tensor of zeros
        save = {}
        h_all = gru_sequence_nograd(gi, gi2, h0, w_hh, b_hh, lengths, reverse, save, n_steps, xsrc)
        ctx.save_for_backward(save["rzn"], save["hn"], h_all, h0, w_hh, lengths)
        ctx.wb = save.get("wb")           # bf16 copy of W_hh the forward multiplied (batch-sized recurrences), or None
        ctx.resident = bool(save.get("resident"))
        ctx.reverse = reverse
        ctx.has_gi2 = gi2 is not None
        ctx.t_full = gi.shape[1]
        if final_only:
            return h_all[:, 0 if reverse else h_all.shape[1] - 1].contiguous()
        return h_all

    @staticmethod
    def backward(ctx, dout):
        rzn, hn, h_all, h0, w_hh, lengths = ctx.saved_tensors
        B, T, H = h_all.shape
        dev = dout.device
        if not dout.is_contiguous():
            dout = dout.contiguous()
        last = (0 if ctx.reverse else T - 1) if ctx.final_only else -1
        if ctx.slab is not None:       # gi is a head of a linear_split: its gradient goes straight into that op's slab
            dgi = ctx.slab[0].block(ctx.slab[1], ctx.slab[2], (B, ctx.t_full), slot_major=len(ctx.slab) > 3)
        else:
            dgi = torch.empty(B, ctx.t_full, 3 * H, device=dev, dtype=torch.float32)
        if ctx.t_full > T:
            dgi[:, T:].zero_()                     # unused input slots get no gradient
        dgh = torch.empty(B, T, 3 * H, device=dev, dtype=torch.float32)
        dgi2 = None        # gradient of the step-constant projection: summed over the steps after the loop
        order = list(range(T - 1, -1, -1) if ctx.reverse else range(T))
        want_dh0 = h0 is not None and ctx.needs_input_grad[2]
        dh0 = torch.empty(B, H, device=dev, dtype=torch.float32) if want_dh0 else None
        slab_rows = ctx.slab[4] if ctx.slab is not None and len(ctx.slab) > 4 else None
        if ctx.resident and (slab_rows is not None or last >= 0):
            _call("pd_gru128_bwd_rows", _ptr(dout), dout.stride(0), dout.stride(1) if last < 0 else 0, _ptr(h_all),
                  h_all.stride(0), h_all.stride(1), _ptr(rzn), rzn.stride(0), rzn.stride(1), _ptr(hn), hn.stride(0),
                  hn.stride(1), _ptr(lengths), _ptr(w_hh), _ptr(dgi), dgi.stride(0), dgi.stride(1), _ptr(dgh), dgh.stride(0),
                  dgh.stride(1), B, T, int(ctx.reverse), None if slab_rows is None else _ptr(slab_rows[0]), last, _stream())
        elif ctx.resident:
            _call("pd_gru128_bwd", _ptr(dout), dout.stride(0), dout.stride(1), _ptr(h_all), h_all.stride(0),
                  h_all.stride(1), _ptr(rzn), rzn.stride(0), rzn.stride(1), _ptr(hn), hn.stride(0), hn.stride(1),
                  _ptr(lengths), _ptr(w_hh), _ptr(dgi), dgi.stride(0), dgi.stride(1), _ptr(dgh), dgh.stride(0),
                  dgh.stride(1), B, T, int(ctx.reverse), _stream())
        else:
            def run(sl):
                _gru_steps_bwd(dout[sl] if last < 0 else (dout[sl], last), rzn[sl], hn[sl], h_all[sl], _sl(h0, sl), w_hh,
                               _sl(lengths, sl), order, dgi[sl], dgh[sl], _sl(dh0, sl), ctx.wb)
            _over_row_chunks(B, 3 * H, run)
        dgh_flat = dgh.view(B * T, 3 * H)
        db = torch.empty(3 * H, device=dev, dtype=torch.float32)
        has_gi2 = ctx.has_gi2
        if has_gi2:
            # one pass over dgi instead of a read-modify-write of (B,3H) in every step's gate kernel; the r and z
            # thirds of db_hh equal those of sum(dgi) (dgh differs from dgi only in the n gate)
            dgi2 = sum_steps(dgi, T)
            colsum(dgi2[:, :2 * H], db[:2 * H])        # (in line: dgi2 goes back to the engine, which may add into it)
        dw = torch.empty(w_hh.shape, device=dev, dtype=torch.float32)
        first = order[0]
        reverse = ctx.reverse

        seq_lengths = lengths if (lengths is not None and ctx.resident and dgh.is_contiguous()) else None

        def wgrads():
            if has_gi2:
                colsum(dgh_flat[:, 2 * H:], db[2 * H:])
            elif seq_lengths is not None:
                # length-masked recurrence (the note summariser): 3/4 of dgh's rows are the zeros of masked steps -- not read
                _call("pd_colsum_seq_f32", _ptr(dgh), 3 * H, B, T, 3 * H, _ptr(seq_lengths), _ptr(db), 0, _stream())
            else:
                colsum(dgh_flat, db)
            # dW_hh = sum_{b,t} dgh[b,t]^T h_prev[b,t] as ONE split-K GEMM over all (b,t) rows: h_prev of row r is
            # row r-1 (r+1 when reversed) of the flattened state buffer, except at each sequence's first step,
            # whose h_prev is h0 -- those rows are handled by a small GEMM and then zeroed in dgh.
            if h0 is not None:
                gemm_tn(dgh[:, first], h0, dw)
            if T > 1:
                dgh[:, first].zero_()
                h_flat = h_all.view(B * T, H)
                if reverse:
                    gemm_tn(dgh_flat[:-1], h_flat[1:], dw, accumulate=h0 is not None)
                else:
                    gemm_tn(dgh_flat[1:], h_flat[:-1], dw, accumulate=h0 is not None)
            elif h0 is None:
                dw.zero_()
        if ctx.wg is not None:
            ctx.wg.set(wgrads, keep=(dgh, h_all, h0))     # off the chain: on the weight-gradient stream (ops.defer)
        else:
            wgrads()
        return dgi, dgi2, dh0, dw, db, None, None, None, None, None, None, None


def gru_sequence(gi, gi2, h0, w_hh, b_hh, lengths=None, reverse=False, n_steps=None, xsrc=None, final_only=False):
    """Autograd-aware GRU over (B,T,3H) input projections; falls to the no-grad loop when nothing
    requires grad (inference).  ``xsrc``: see ``gru_sequence_nograd`` (gi then only routes the gradient).
    ``final_only``: return the (B,H) state after the last processed step (index 0 when reversed) instead of all states."""
    if xsrc is not None:
        xsrc = (xsrc[0].detach(), xsrc[1].detach())           # gradients flow through gi's producer (linear_split)
    if torch.is_grad_enabled() and (gi.requires_grad or w_hh.requires_grad or
                                    (h0 is not None and h0.requires_grad)):
        wg = None
        if gi.shape[0] * gi.shape[1] >= DEFER_MIN_ROWS:
            w_hh, b_hh, wg = defer(w_hh, b_hh)
        return _GruSeq.apply(gi, gi2, h0, w_hh, b_hh, lengths, reverse, n_steps, _slab_of(gi, gi.shape[-1]), xsrc, wg,
                             final_only)
    h = gru_sequence_nograd(gi, gi2, h0, w_hh, b_hh, lengths, reverse, None, n_steps, xsrc)
    return h[:, 0 if reverse else h.shape[1] - 1] if final_only else h


# ------------------------------------------------------------------------------------------------
# Packed note level (csrc/packed.cu).  In loss mode 3/4 of the teacher-forced note-level work of the reference is dead: a
# (segment, time step) row with k notes has k + 1 target tokens, the loss ignores the other slots of the 15 and no
# gradient leaves them.  The rows are therefore sorted by token count and every note-level buffer is slot-major, so the
# live rows of a slot are a prefix whose length is DEVICE data (``table``): kernels are launched for the full extent -- the
# step stays one CUDA graph for every batch -- and skip dead tiles.  Results on live positions are those of the full
# computation; dead positions are never written and never read.
PACKED_NOTES = True
PACK_NB = 17           # entries per block of the pack table: c[t] | cp[t] | 6 cp[t], t = 0..16 (csrc/packed.cu)


class Packed:
    """Row order of one batch: ``perm`` (sorted position -> row), ``inv``, the live-row ``table``, and the slot-major
    tokens / targets of the sorted rows."""

    def __init__(self, tok, lengths32):
        R = lengths32.numel()
        dev = tok.device
        self.R = R
        self.perm = torch.empty(R, device=dev, dtype=torch.int32)
        self.inv = torch.empty(R, device=dev, dtype=torch.int32)
        self.table = torch.zeros(64, device=dev, dtype=torch.int32)
        _call("pd_pack_order", _ptr(lengths32), R, _ptr(self.perm), _ptr(self.inv), _ptr(self.table), _stream())
        self.tok = torch.empty(16 * R, 6, device=dev, dtype=torch.int32)          # (16, R, 6)
        self.pitch_tgt = torch.empty(15 * R, device=dev, dtype=torch.int32)       # (15, R)
        self.dur_tgt = torch.empty(15 * R * 5, device=dev, dtype=torch.int32)     # (15, R, 5)
        self.lengths = torch.empty(R, device=dev, dtype=torch.int32)              # of the sorted rows (descending)
        _call("pd_pack_grid", _ptr(tok), _ptr(lengths32), _ptr(self.perm), R, _ptr(self.tok), _ptr(self.pitch_tgt),
              _ptr(self.dur_tgt), _ptr(self.lengths), _stream())

    def slot_major_tokens(self, tok):
        """Another token grid of the same rows (R*16, 6) int32 -- e.g. the tokens a scheduled-sampling pass fed -- in this
        order's slot-major layout (16*R, 6)."""
        R = self.R
        out = torch.empty(16 * R, 6, device=tok.device, dtype=torch.int32)
        scratch = torch.empty(15 * R * 6 + R, device=tok.device, dtype=torch.int32)     # targets / lengths of THAT grid: unused
        _call("pd_pack_grid", _ptr(tok), _ptr(self.lengths), _ptr(self.perm), R, _ptr(out), _ptr(scratch), _ptr(scratch[15 * R:]),
              _ptr(scratch[:R]), _stream())
        return out

    def rows(self, t0):
        """Live-row predicate (cp tensor, slot_rows) of slot-major buffers whose slot 0 needs tokens beyond position ``t0``:
        0 for the embedded tokens (slot n is live for rows with more than n tokens), 1 for the note-GRU states / logits
        (slot n predicts token n + 1)."""
        return (self.table[PACK_NB + t0:], self.R)


class _GatherRows(torch.autograd.Function):
    """y = x[idx] for a PERMUTATION idx of the rows (inverse ``inv``): backward is the inverse gather."""

    @staticmethod
    def forward(ctx, x, idx, inv):
        x2, _ = _rows(_chk(x, "x"))
        y = torch.empty(x2.shape[0], x2.shape[1], device=x.device, dtype=torch.float32)
        _call("pd_gather_rows_f32", _ptr(x2), x2.stride(0), _ptr(idx), x2.shape[0], x2.shape[1], _ptr(y), y.stride(0), _stream())
        ctx.save_for_backward(inv)
        ctx.x_shape = x.shape
        return y

    @staticmethod
    def backward(ctx, g):
        (inv,) = ctx.saved_tensors
        g2, _ = _rows(g)
        d = torch.empty(g2.shape[0], g2.shape[1], device=g.device, dtype=torch.float32)
        _call("pd_gather_rows_f32", _ptr(g2), g2.stride(0), _ptr(inv), g2.shape[0], g2.shape[1], _ptr(d), d.stride(0), _stream())
        return d.view(ctx.x_shape), None, None


def gather_rows(x, idx, inv):
    return _GatherRows.apply(x, idx, inv)


class _NoteGruPacked(torch.autograd.Function):
    """Teacher-forced note GRU (ptvae.py:396-398 for all 15 slots) over length-sorted rows, slot-major.
    emb (16,R,K2) embedded ground-truth tokens, w_x (3H,K2) their W_ih columns, gi_s (R,3H) the step-constant rest of the
    input projection incl. b_ih, h0 (R,H).  Returns the states (15,R,H); slot n is computed for the first table[cp][n+1]
    rows only (fused step kernel with the x-projection as a second K segment), the rest of the buffer is never written."""

    @staticmethod
    def forward(ctx, emb, w_x, gi_s, h0, w_hh, b_hh, table, wg):
        T, R, H = emb.shape[0] - 1, emb.shape[1], h0.shape[1]
        dev = emb.device
        h_all = torch.empty(T, R, H, device=dev, dtype=torch.float32)
        rzn = torch.empty(T, R, 3 * H, device=dev, dtype=torch.float32)
        hn = torch.empty(T, R, H, device=dev, dtype=torch.float32)
        hprev = h0
        for n in range(T):
            _call("pd_gru_step_tmax_rows", _ptr(hprev), hprev.stride(0), _ptr(w_hh), w_hh.stride(0), _ptr(emb[n]), emb.stride(1),
                  _ptr(w_x), w_x.stride(0), emb.shape[2], _ptr(b_hh), _ptr(gi_s), gi_s.stride(0), _ptr(h_all[n]), H,
                  _ptr(rzn[n]), 3 * H, _ptr(hn[n]), H, R, H, _ptr(table[PACK_NB + n + 1:]), _stream())
            hprev = h_all[n]
        ctx.save_for_backward(emb, w_x, h0, w_hh, rzn, hn, h_all, table)
        ctx.wg = wg
        return h_all

    @staticmethod
    def backward(ctx, dout):
        emb, w_x, h0, w_hh, rzn, hn, h_all, table = ctx.saved_tensors
        T, R, H = h_all.shape
        K2 = emb.shape[2]
        dev = dout.device
        if not dout.is_contiguous():
            dout = dout.contiguous()
        st = _stream()
        # recurrent gradient carriers (dh*z of the later step, dgh.W_hh of the later step), ping-pong.  Zero-filled ONCE:
        # a row enters the live prefix at its last token and has no later step, so it must read zeros there
        bufs = torch.zeros(4, R, H, device=dev, dtype=torch.float32)
        dz_a, dz_b, dm_a, dm_b = bufs[0], bufs[1], bufs[2], bufs[3]
        dgi = torch.empty(T, R, 3 * H, device=dev, dtype=torch.float32)
        dgh = torch.empty(T, R, 3 * H, device=dev, dtype=torch.float32)
        dz = dm = None
        for n in range(T - 1, -1, -1):
            hprev = h_all[n - 1] if n > 0 else h0
            nz = dz_b if dz is dz_a else dz_a
            nm = dm_b if dm is dm_a else dm_a
            cp_n = table[PACK_NB + n + 1:]
            _call("pd_gru_gates_bwd_rows", _ptr(dz), 0 if dz is None else H, _ptr(dout[n]), H, _ptr(dm), 0 if dm is None else H,
                  _ptr(rzn[n]), 3 * H, _ptr(hn[n]), H, _ptr(hprev), hprev.stride(0), _ptr(dgi[n]), 3 * H, _ptr(dgh[n]), 3 * H,
                  _ptr(nz), H, R, H, _ptr(cp_n), st)
            gemm_nn(dgh[n], w_hh, nm, rows=(cp_n, R))                 # dgh W_hh for the live row tiles
            dz, dm = nz, nm
        dh0 = torch.empty(R, H, device=dev, dtype=torch.float32)
        _call("pd_add_f32", _ptr(dz), _ptr(dm), dz.numel(), _ptr(dh0), st)
        rows1 = (table[PACK_NB + 1:], R)
        # gradient of the step-constant projection: the live slots of every row, one pass
        dgi2 = torch.empty(R, 3 * H, device=dev, dtype=torch.float32)
        _call("pd_sum_slots_rows_f32", _ptr(dgi), R * 3 * H, 3 * H, T, _ptr(rows1[0]), _ptr(dgi2), 3 * H, R, 3 * H, st)
        db = torch.empty(3 * H, device=dev, dtype=torch.float32)
        colsum(dgi2[:, :2 * H], db[:2 * H])       # r / z thirds of db_hh = those of sum(dgi); in line (dgi2 goes to the engine)
        # input gradient of the embedded tokens: zero where no live step consumed them (slot 15 is never an input)
        demb = torch.zeros(T + 1, R, K2, device=dev, dtype=torch.float32)
        dgi_flat, dgh_flat = dgi.view(T * R, 3 * H), dgh.view(T * R, 3 * H)
        gemm_nn(dgi_flat, w_x, demb[:T].view(T * R, K2), rows=rows1)
        dwx = torch.empty(w_x.shape, device=dev, dtype=torch.float32)
        dw = torch.empty(w_hh.shape, device=dev, dtype=torch.float32)

        def wgrads():
            colsum(dgh_flat[:, 2 * H:], db[2 * H:], rows=rows1)
            gemm_tn(dgi_flat, emb[:T].view(T * R, K2), dwx, rows=rows1)
            # dW_hh: h_prev of (slot n, row r) is (slot n - 1, row r) -- R rows earlier in the slot-major buffer -- and h0
            # for slot 0
            gemm_tn(dgh[0], h0, dw, rows=rows1)
            if T > 1:
                gemm_tn(dgh_flat[R:], h_all.view(T * R, H)[:-R], dw, accumulate=True, rows=(table[PACK_NB + 2:], R))
        if ctx.wg is not None:
            ctx.wg.set(wgrads, keep=(dgi, dgh, emb, h_all, h0))
        else:
            wgrads()
        return demb, dwx, dgi2, dh0, dw, db, None, None


def note_gru_packed(emb, w_x, gi_s, h0, w_hh, b_hh, table):
    """See ``_NoteGruPacked``.  emb (16,R,K2) contiguous; requires the tensor-core (tf32) precision mode."""
    wg = None
    w_x, w_hh, b_hh, wg = defer(w_x, w_hh, b_hh)
    return _NoteGruPacked.apply(emb, w_x, gi_s, h0, w_hh, b_hh, table, wg)


def packed_ok(R, H, K2):
    """Can a batch of R (segment, time step) rows take the packed note level?  TF32 tensor-core mode, TMA-addressable
    shapes, and R a multiple of 128 (batch a multiple of 4): the kernels skip work per 128-row tile, and a tile must not
    straddle two note slots -- rows of a dead slot inside a live tile would be written with values computed from
    unwritten inputs.  Other batch sizes take the dense path."""
    return (PACKED_NOTES and PRECISION == "tf32" and FUSED_GRU_STEP_TMA and R % 128 == 0 and H % 64 == 0 and K2 % 4 == 0
            and torch.is_grad_enabled())


# ------------------------------------------------------------------------------------------------
def grid_prepare(x):
    """x (B,32,16,6) int64 -> tok int32 (B*512,6), lengths int32 (B*32), pitch_tgt (B*480), dur_tgt (B*2400)."""
    _chk(x, "x")
    x = x.contiguous()
    B = x.shape[0]
    dev = x.device
    tok = torch.empty(B * 512, 6, device=dev, dtype=torch.int32)
    lengths = torch.empty(B * 32, device=dev, dtype=torch.int32)
    pt = torch.empty(B * 480, device=dev, dtype=torch.int32)
    dt = torch.empty(B * 2400, device=dev, dtype=torch.int32)
    _call("pd_grid_prepare", _ptr(x), B * 32, _ptr(tok), _ptr(lengths), _ptr(pt), _ptr(dt), _stream())
    return tok, lengths, pt, dt


def pr_mat_to_grid(pr_mat):
    """Device-side batch construction: pr_mat (B,32,128) fp32 -> PianoTree grid x
    (B,32,16,6) int64, as converter.target_to_3dtarget does per item on the host (dataset.py:98-104).
    Returns (x, overflow) where overflow is a device int32 flag set if any step held more than 14 onsets."""
    pr = _chk(pr_mat, "pr_mat").contiguous()
    B = pr.shape[0]
    x = torch.empty(B, 32, 16, 6, device=pr.device, dtype=torch.int64)
    overflow = torch.zeros(1, device=pr.device, dtype=torch.int32)
    _call("pd_prmat_to_grid", _ptr(pr), B * 32, _ptr(x), _ptr(overflow), _stream())
    return x, overflow


def pack_tokens(tokens):
    """Decoded tokens (...,6) int32 -> compact (...,2) uint8 [pitch, 5 duration bits packed MSB-first] on the device."""
    tok = _chk(tokens, "tokens").contiguous()
    R = tok.numel() // 6
    out = torch.empty(tuple(tok.shape[:-1]) + (2,), device=tok.device, dtype=torch.uint8)
    _call("pd_pack_tokens", _ptr(tok), R, _ptr(out), _stream())
    return out


def unpack_tokens(packed):
    """Host-side inverse of ``pack_tokens``: uint8 ndarray (...,2) -> int64 ndarray (...,6), the reference's est_x layout."""
    import numpy as np
    packed = np.asarray(packed)
    out = np.empty(packed.shape[:-1] + (6,), dtype=np.int64)
    out[..., 0] = packed[..., 0]
    for b in range(5):
        out[..., 1 + b] = (packed[..., 1] >> (4 - b)) & 1
    return out


def augment_batch(pr_mat, chord14, shift):
    """Device-side batch augmentation + construction (dataset.py:67-120 per item on the host):
    transpose every segment by ``shift[b]`` semitones -- ``np.roll`` of the piano-roll along pitch
    (converter.py:65-68) and ``expand_chord(c, shift)`` (converter.py:150-164) -- then build the PianoTree grid.

    pr_mat (B,32,128) fp32, chord14 (B,8,14) fp32 [root, 12 chroma bits, bass], shift (B,) int.
    Returns (x (B,32,16,6) int64, c (B,8,36) fp32, pr_mat' (B,32,128) fp32, overflow flag) -- the model's inputs."""
    pr = _chk(pr_mat, "pr_mat").contiguous()
    ch = _chk(chord14, "chord14").contiguous().to(torch.float32)
    B = pr.shape[0]
    sh = shift.to(device=pr.device, dtype=torch.int32).contiguous()
    out = torch.empty_like(pr)
    _call("pd_roll_prmat", _ptr(pr), _ptr(sh), B, _ptr(out), _stream())
    rows_per_seg = ch.shape[1]
    c36 = torch.empty(B, rows_per_seg, 36, device=pr.device, dtype=torch.float32)
    _call("pd_expand_chord", _ptr(ch), _ptr(sh), B * rows_per_seg, rows_per_seg, _ptr(c36), _stream())
    x, overflow = pr_mat_to_grid(out)
    return x, c36, out, overflow


def slerp_path(z1, z2, count):
    """(B,D), (B,D) -> (B,count,D): model.py:218-242 ``interp_path`` for every pair, on the device."""
    a, b = _chk(z1, "z1").contiguous().to(torch.float32), _chk(z2, "z2").contiguous().to(torch.float32)
    B, D = a.shape
    out = torch.empty(B, count, D, device=a.device, dtype=torch.float32)
    _call("pd_slerp_path", _ptr(a), _ptr(b), B, D, count, _ptr(out), _stream())
    return out


def tokens_to_pr_mat(tokens):
    """Decoded tokens (B,32,15,6) int32 -> pr_mat (B,32,128) fp32 on device (ptvae.py:558-575 without the
    host loop and without moving the tokens off the GPU)."""
    tok = _chk(tokens, "tokens").to(torch.int32).contiguous()
    B = tok.shape[0]
    pr = torch.empty(B, 32, 128, device=tok.device, dtype=torch.float32)
    _call("pd_grid_to_prmat", _ptr(tok), B * 32, _ptr(pr), _stream())
    return pr


class _NoteEmbed(torch.autograd.Function):
    """note_embedding(multi-hot) as a 6-row gather-add (ptvae.py:299-313,:333); tok int32 (R,6)."""

    @staticmethod
    def forward(ctx, tok, w, b, out=None, wg=None, rows=None):
        ctx.rows = rows          # packed note level: tok rows are slot-major; only live rows send gradient to the table
        R = tok.shape[0]
        wt = transpose(w)                                  # (135,128): rows contiguous per pitch
        if out is None:
            out = torch.empty(R, 128, device=w.device, dtype=torch.float32)
        _call("pd_note_embed_fwd", _ptr(tok), R, _ptr(wt), _ptr(b), _ptr(out), out.stride(0), _stream())
        ctx.save_for_backward(tok)
        ctx.wg = wg
        return out

    @staticmethod
    def backward(ctx, g):
        (tok,) = ctx.saved_tensors
        g2, _ = _rows(g)
        dw = torch.empty(128, 135, device=g.device, dtype=torch.float32)
        db = torch.empty(128, device=g.device, dtype=torch.float32)

        rows = ctx.rows

        def wgrads():                                      # the whole backward is parameter gradients (tokens are ints)
            dwt = torch.zeros(135, 128, device=dw.device, dtype=torch.float32)
            db.zero_()
            if rows is not None:
                _call("pd_note_embed_bwd_rows", _ptr(tok), tok.shape[0], _ptr(g2), g2.stride(0), _ptr(dwt), _ptr(db),
                      _ptr(rows[0]), rows[1], _stream())
            else:
                _call("pd_note_embed_bwd", _ptr(tok), tok.shape[0], _ptr(g2), g2.stride(0), _ptr(dwt), _ptr(db), _stream())
            _call("pd_transpose_f32", _ptr(dwt), 135, 128, _ptr(dw), _stream())
        if ctx.wg is not None:
            ctx.wg.set(wgrads, keep=(g2, tok))
        else:
            wgrads()
        return None, dw, db, None, None, None


def note_embed(tok, w, b, rows=None):
    wg = None
    if torch.is_grad_enabled() and tok.shape[0] >= DEFER_MIN_ROWS:
        w, b, wg = defer(w, b)
    return _NoteEmbed.apply(tok, w, b, None, wg, rows)


class _TextureFrontend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pr_mat, w, b, wg=None):
        _chk(pr_mat, "pr_mat")
        pr = pr_mat.contiguous()
        B, C = pr.shape[0], w.shape[0]
        out = torch.empty(B, C, 8, 29, device=pr.device, dtype=torch.float32)
        wc = w.contiguous()
        if any(ctx.needs_input_grad):
            # training: the forward also records the pooled arg-max positions; the backward reads them instead of
            # recomputing the convolution windows
            amax = torch.empty(B, C, 8, 29, device=pr.device, dtype=torch.int8)
            _call("pd_texture_frontend_fwd_ix", _ptr(pr), _ptr(wc), _ptr(b), B, C, _ptr(out), _ptr(amax), _stream())
            ctx.save_for_backward(pr, amax)
        else:
            _call("pd_texture_frontend_fwd", _ptr(pr), _ptr(wc), _ptr(b), B, C, _ptr(out), _stream())
        ctx.wshape, ctx.bshape = w.shape, b.shape
        ctx.wg = wg
        return out

    @staticmethod
    def backward(ctx, g):
        pr, amax = ctx.saved_tensors
        g = g.contiguous()
        dw = torch.empty(ctx.wshape, device=g.device, dtype=torch.float32)
        db = torch.empty(ctx.bshape, device=g.device, dtype=torch.float32)

        def wgrads():                                      # (the piano-roll input needs no gradient)
            dw.zero_()
            db.zero_()
            _call("pd_texture_frontend_bwd_ix", _ptr(pr), _ptr(amax), pr.shape[0], ctx_c, _ptr(g), _ptr(dw), _ptr(db), _stream())
        ctx_c = ctx.wshape[0]
        if ctx.wg is not None:
            ctx.wg.set(wgrads, keep=(g, pr, amax))
        else:
            wgrads()
        return None, dw, db, None


def texture_frontend(pr_mat, w, b):
    wg = None
    if torch.is_grad_enabled() and pr_mat.shape[0] * 32 >= DEFER_MIN_ROWS:
        w, b, wg = defer(w, b)
    return _TextureFrontend.apply(pr_mat, w, b, wg)


class _MaskedCE(torch.autograd.Function):
    """mean CE over rows whose int32 target != ignore (nn.CrossEntropyLoss(ignore_index))."""

    @staticmethod
    def forward(ctx, logits, targets, ignore, slab=None):
        l2, _ = _rows(_chk(logits, "logits"))
        ctx.slab = slab
        acc = torch.empty(2, device=l2.device, dtype=torch.float32)
        loss = torch.empty((), device=l2.device, dtype=torch.float32)
        _call("pd_ce_fwd", _ptr(l2), l2.stride(0), _ptr(targets), l2.shape[0], l2.shape[1], ignore,
              _ptr(acc), _ptr(loss), _stream())
        ctx.save_for_backward(l2, targets, acc)
        ctx.ignore = ignore
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        l2, targets, acc = ctx.saved_tensors
        if ctx.slab is not None:        # logits are a head of a linear_split: write into that op's gradient slab
            d = ctx.slab[0].block(ctx.slab[1], ctx.slab[2], (l2.shape[0],))
        else:
            d = _empty_rows(l2.shape[0], l2.shape[1], l2.device)
        g = g.contiguous()
        _call("pd_ce_bwd", _ptr(l2), l2.stride(0), _ptr(targets), l2.shape[0], l2.shape[1], ctx.ignore,
              _ptr(acc), _ptr(g), _ptr(d), d.stride(0), _stream())
        if d.is_contiguous():
            return d.view(ctx.shape), None, None, None
        lead = tuple(ctx.shape[:-1])
        strides, a = [], d.stride(0)
        for k in reversed(lead):
            strides.append(a)
            a *= k
        return d.as_strided(lead + (l2.shape[1],), tuple(reversed(strides)) + (1,)), None, None, None


def _slab_of(t, width):
    slab = getattr(t, "_pd_slab", None)
    return slab if slab is not None and slab[2] == width else None


def slot_major_seq(g, T, R, rows=None):
    """A head of a ``linear_split`` over slot-major rows (T*R, n) as the (R, T, n) sequence view the GRU ops take; the
    gradient-slab tag follows (the recurrence's backward then writes its input gradient in place, slot-major).
    ``rows``: the live-row predicate the producing ``linear_split`` was given -- dead (row, step) entries of g are unwritten
    (the masked recurrences never read them) and their gradient need not be written."""
    v = g.view(T, R, -1).permute(1, 0, 2)
    tag = getattr(g, "_pd_slab", None)
    if tag is not None:
        v._pd_slab = tag + ("slot_major", rows)
    return v


def keep_slab(new, old):
    """Carry the linear_split gradient-slab tag of ``old`` over to a reshaped view of it."""
    tag = getattr(old, "_pd_slab", None)
    if tag is not None:
        new._pd_slab = tag
    return new


def masked_ce(logits, targets, ignore=-100):
    return _MaskedCE.apply(logits, targets, ignore, _slab_of(logits, logits.shape[-1]) if logits.dim() == 2 else None)


class _Exp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x).contiguous()
        y = torch.empty_like(x)
        _call("pd_exp_fwd", _ptr(x), x.numel(), _ptr(y), _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = g.contiguous()
        d = torch.empty_like(y)
        _call("pd_mul_f32", _ptr(g), _ptr(y), y.numel(), _ptr(d), _stream())
        return d


def exp(x):
    return _Exp.apply(x)


class _Reparam(torch.autograd.Function):
    """z = mu + std * eps (Normal.rsample, train_utils.py:33-34); eps None -> z = mu."""

    @staticmethod
    def forward(ctx, mu, sd, eps):
        mu, sd = _chk(mu).contiguous(), sd.contiguous()
        B, D = mu.shape
        z = torch.empty(B, D, device=mu.device, dtype=torch.float32)
        _call("pd_reparam_fwd", _ptr(mu), _ptr(sd), _ptr(eps), B, D, _ptr(z), z.stride(0), _stream())
        ctx.save_for_backward(eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        (eps,) = ctx.saved_tensors
        dz2, _ = _rows(dz)
        B, D = dz2.shape
        dmu = torch.empty(B, D, device=dz.device, dtype=torch.float32)
        dsd = torch.empty(B, D, device=dz.device, dtype=torch.float32)
        _call("pd_reparam_bwd", _ptr(dz2), dz2.stride(0), _ptr(eps), B, D, _ptr(dmu), _ptr(dsd), _stream())
        return dmu, dsd, None


def reparam(mu, sd, eps):
    return _Reparam.apply(mu, sd, eps)


class _KL(torch.autograd.Function):
    """mean over all elements of KL(N(mu, sd) || N(0,1))  (train_utils.py:45-49)."""

    @staticmethod
    def forward(ctx, mu, sd):
        mu, sd = _chk(mu).contiguous(), sd.contiguous()
        out = torch.empty((), device=mu.device, dtype=torch.float32)
        _call("pd_kl_fwd", _ptr(mu), _ptr(sd), mu.numel(), _ptr(out), _stream())
        ctx.save_for_backward(mu, sd)
        return out

    @staticmethod
    def backward(ctx, g):
        mu, sd = ctx.saved_tensors
        g = g.contiguous()
        dmu, dsd = torch.empty_like(mu), torch.empty_like(sd)
        _call("pd_kl_bwd", _ptr(mu), _ptr(sd), mu.numel(), _ptr(g), _ptr(dmu), _ptr(dsd), _stream())
        return dmu, dsd


def kl_std_normal(mu, sd):
    return _KL.apply(mu, sd)


# ------------------------------------------------------------------------------------------------
def greedy_pick(pitch, dur, n, tok_out, lens):
    """pitch (R,130), dur (R,5,2) logits -> tok_out int32 (R,6) view; updates lens (R,) int32."""
    p2, _ = _rows(pitch)
    d2 = dur.reshape(dur.shape[0], 10)
    if d2.stride(1) != 1:
        d2 = d2.contiguous()
    R = p2.shape[0]
    _call("pd_greedy_pick", _ptr(p2), p2.stride(0), _ptr(d2), d2.stride(0), R, n, _ptr(tok_out),
          tok_out.stride(0), _ptr(lens), _stream())


def greedy_pick_embed(pitch, dur, n, tok_out, lens, emb_wt, emb_b, emb_out):
    """``greedy_pick`` + the embedding of the picked tokens (``pd_note_embed_fwd``) in one launch: emb_out (R,128) view."""
    p2, _ = _rows(pitch)
    d2 = dur.reshape(dur.shape[0], 10)
    if d2.stride(1) != 1:
        d2 = d2.contiguous()
    R = p2.shape[0]
    _call("pd_greedy_pick_embed", _ptr(p2), p2.stride(0), _ptr(d2), d2.stride(0), R, n, _ptr(tok_out), tok_out.stride(0),
          _ptr(lens), _ptr(emb_wt), _ptr(emb_b), _ptr(emb_out), emb_out.stride(0), _stream())


def dur_token(logit):
    """(R,2) logits -> (R,5) feedback token with the 1 at index == argmax bit (ptvae.py:322-326)."""
    l2, _ = _rows(logit)
    tok = torch.empty(l2.shape[0], 5, device=l2.device, dtype=torch.float32)
    _call("pd_dur_token", _ptr(l2), l2.stride(0), l2.shape[0], _ptr(tok), _stream())
    return tok


class _DurDecode(torch.autograd.Function):
    """Fused 5-step duration GRU + head with greedy bit feedback (ptvae.py:353-367).  Forward: one
    weight-resident kernel.  Backward: one kernel for the per-note BPTT plus ONE tensor-core GEMM
    (GX^T . S) that produces every parameter gradient (layout in csrc/dur_decoder.cu)."""

    @staticmethod
    def forward(ctx, h0, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, slab=None, wg=None, rows=None):
        h2, _ = _rows(_chk(h0, "dur h0"))
        ctx.slab = slab
        ctx.wg = wg
        ctx.rows = rows          # packed note level: notes are slot-major rows; dead 16-note tiles are skipped
        Q = h2.shape[0]
        dev = h2.device
        logits = torch.empty(Q, 5, 2, device=dev, dtype=torch.float32)
        need = any(ctx.needs_input_grad)
        S = torch.empty(Q, 6, 72, device=dev, dtype=torch.float32) if need else None
        params = [t.contiguous() for t in (w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out)]
        ctx.tf32 = dur_mode()
        if rows is not None:
            assert ctx.tf32 == 1
            _call("pd_dur_decode_fwd_rows", _ptr(h2), h2.stride(0), Q, *[_ptr(t) for t in params], _ptr(logits), _ptr(S),
                  _ptr(rows[0]), rows[1], _stream())
        else:
            _call("pd_dur_decode_fwd", _ptr(h2), h2.stride(0), Q, *[_ptr(t) for t in params], _ptr(logits), _ptr(S),
                  ctx.tf32, _stream())
        ctx.save_for_backward(S, *params)
        ctx.h_shape = h0.shape
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        S, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out = ctx.saved_tensors
        Q = S.shape[0]
        dev = S.device
        dlogits = dlogits.contiguous()
        GX = torch.empty(Q, 6, 264, device=dev, dtype=torch.float32)
        if ctx.slab is not None:
            dh0 = ctx.slab[0].block(ctx.slab[1], ctx.slab[2], (Q,))
        else:
            dh0 = torch.empty(Q, 64, device=dev, dtype=torch.float32)
        rows = ctx.rows
        # rows of the GX^T . S contraction: 6 per note, so the live-row table in units of 6 (third block of the pack table)
        rows6 = None if rows is None else (rows[0][PACK_NB:], 6 * rows[1])
        if rows is not None:
            _call("pd_dur_decode_bwd_rows", _ptr(S), _ptr(dlogits), Q, *[_ptr(t) for t in (w_ih, b_ih, w_hh, b_hh, sos, w_out,
                                                                                          b_out)],
                  _ptr(GX), _ptr(dh0), dh0.stride(0), _ptr(rows[0]), rows[1], _stream())
        else:
            _call("pd_dur_decode_bwd", _ptr(S), _ptr(dlogits), Q, *[_ptr(t) for t in (w_ih, b_ih, w_hh, b_hh, sos, w_out,
                                                                                     b_out)],
                  _ptr(GX), _ptr(dh0), dh0.stride(0), ctx.tf32, _stream())
        if ctx.wg is not None:
            # every parameter gradient comes from the GX^T . S GEMM: buffers now, the GEMM and its unpacking on the
            # weight-gradient stream (ops.defer)
            outs = [torch.empty(t.shape, device=dev, dtype=torch.float32) for t in (w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out)]

            def job():
                G = torch.empty(264, 72, device=dev, dtype=torch.float32)
                gemm_tn(GX.view(Q * 6, 264), S.view(Q * 6, 72), G, rows=rows6)
                gi_rows = torch.cat([G[0:128], G[192:256]], 0)
                for o, v in zip(outs, (gi_rows[:, 64:69], gi_rows[:, 69], G[0:192, 0:64], G[0:192, 69],
                                       w_ih.t() @ gi_rows[:, 70], G[256:258, 0:64], G[256:258, 69])):
                    o.copy_(v)
            ctx.wg.set(job, keep=(GX, S))
            return (dh0.view(ctx.h_shape), *outs, None, None, None)
        G = torch.empty(264, 72, device=dev, dtype=torch.float32)
        gemm_tn(GX.view(Q * 6, 264), S.view(Q * 6, 72), G, rows=rows6)
        gi_rows = torch.cat([G[0:128], G[192:256]], 0)            # [dr | dz | dn] x S columns
        dw_hh, db_hh = G[0:192, 0:64], G[0:192, 69]
        dw_ih, db_ih = gi_rows[:, 64:69], gi_rows[:, 69]
        dsos = w_ih.t() @ gi_rows[:, 70]                           # (5,) from the step-0 input-gate grads
        dw_out, db_out = G[256:258, 0:64], G[256:258, 69]
        return (dh0.view(ctx.h_shape), dw_ih.contiguous(), db_ih.contiguous(), dw_hh.contiguous(),
                db_hh.contiguous(), dsos, dw_out.contiguous(), db_out.contiguous(), None, None, None)


#: below this many notes per call the TF32 (one warp per 16 notes) duration decoder is latency bound -- a 512-note slot
#: of the training-time greedy pass takes 21 us, 27 % of a free-running step -- and the fp32 FFMA kernel, which spreads a
#: tile over six warps, is used instead (0: never)
DUR_FFMA_MAX_NOTES = 0


def dur_mode(n_notes=None):
    """Arithmetic flag of the duration-decoder kernels for the current precision scope (csrc/dur_decoder.cu)."""
    if PRECISION == "tf32" and n_notes is not None and n_notes < DUR_FFMA_MAX_NOTES:
        return 0
    return {"fp32": 0, "tf32": 1, "tf32x3": 3}[PRECISION]


def dur_decode(h0, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, rows=None):
    slab = _slab_of(h0, 64) if h0.dim() == 2 else None            # (before tagging: tags are new tensor objects)
    wg = None
    if torch.is_grad_enabled() and _n_rows(h0) >= DEFER_MIN_ROWS:
        w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, wg = defer(w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out)
    return _DurDecode.apply(h0, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, slab, wg, rows)


class _SelectRows(torch.autograd.Function):
    """out = flag ? a : b with the decision read from DEVICE memory (one int32): scheduled sampling whose
    teacher-forcing plan is data, not host control flow, so one captured CUDA graph serves every ratio."""

    @staticmethod
    def forward(ctx, a, b, flag):
        a2, _ = _rows(_chk(a, "a"))
        b2, _ = _rows(_chk(b, "b"))
        out = torch.empty(a2.shape[0], a2.shape[1], device=a2.device, dtype=torch.float32)
        _call("pd_select_rows", _ptr(a2), a2.stride(0), _ptr(b2), b2.stride(0), _ptr(flag), _ptr(out), out.stride(0),
              a2.shape[0], a2.shape[1], _stream())
        ctx.save_for_backward(flag)
        ctx.shapes = (a.shape, b.shape)
        return out.view(a.shape)

    @staticmethod
    def backward(ctx, g):
        (flag,) = ctx.saved_tensors
        g2, _ = _rows(g)
        da = torch.empty(g2.shape, device=g.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        db = torch.empty(g2.shape, device=g.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        _call("pd_select_rows_bwd", _ptr(g2), g2.stride(0), _ptr(flag), _ptr(da), 0 if da is None else da.stride(0),
              _ptr(db), 0 if db is None else db.stride(0), g2.shape[0], g2.shape[1], _stream())
        return (None if da is None else da.view(ctx.shapes[0]), None if db is None else db.view(ctx.shapes[1]), None)


def select_rows(a, b, flag):
    """``flag`` (int32 device tensor, 1 element): nonzero -> a, zero -> b.  a, b: same shape, unit inner stride."""
    return _SelectRows.apply(a, b, flag)


def chord_feedback(root, chroma, bass):
    """(B,12), (B,12,2), (B,12) logits -> (B,36) feedback token (batch-union one-hots, ptvae.py:73-78)."""
    B = root.shape[0]
    r2, _ = _rows(root)
    b2, _ = _rows(bass)
    c2 = chroma.reshape(B, 24)
    if c2.stride(1) != 1:
        c2 = c2.contiguous()
    flags = torch.empty(24, device=root.device, dtype=torch.float32)
    tok = torch.empty(B, 36, device=root.device, dtype=torch.float32)
    _call("pd_chord_feedback", _ptr(r2), r2.stride(0), _ptr(c2), c2.stride(0), _ptr(b2), b2.stride(0), B,
          _ptr(flags), _ptr(tok), tok.stride(0), _stream())
    return tok


def chord_targets(c):
    c = _chk(c, "c").contiguous()
    rows = c.shape[0] * c.shape[1]
    dev = c.device
    root = torch.empty(rows, device=dev, dtype=torch.int32)
    chroma = torch.empty(rows * 12, device=dev, dtype=torch.int32)
    bass = torch.empty(rows, device=dev, dtype=torch.int32)
    _call("pd_chord_targets", _ptr(c), rows, _ptr(root), _ptr(chroma), _ptr(bass), _stream())
    return root, chroma, bass
