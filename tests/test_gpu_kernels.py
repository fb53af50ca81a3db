"""-m gpu: every CUDA kernel of libpolydis_b200, called through the C-ABI, against the numpy
restatement of the same entry point (tests/cpu_backend.py) on seeded inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.cpu_backend import CpuBackend

CPU = CpuBackend()


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _p(t):
    return None if t is None else t.data_ptr()


def _cuda_like(a):
    """GPU copy of ``a`` that keeps its strides: strided views (a step slice of a (B,T,K) buffer, a column block of a wider
    weight) are rebuilt over a copy of their base -- ``.cuda()`` alone would compact them while the test passes the
    original row strides to the kernel."""
    if a.is_contiguous() or a._base is None or not a._base.is_contiguous():
        return a.cuda()
    return a._base.cuda().as_strided(a.shape, a.stride(), a.storage_offset() - a._base.storage_offset())


def _both(name, make_args):
    """make_args(device) -> (args list with tensors, outputs list of tensors).  Runs the CUDA entry on
    GPU tensors and the numpy emulation on CPU copies; returns [(gpu_out, cpu_out)]."""
    from polydis_b200 import _lib
    torch.manual_seed(0)
    args, outs = make_args()
    gargs = [_cuda_like(a) if torch.is_tensor(a) else a for a in args]
    gouts = [gargs[next(i for i, a in enumerate(args) if a is o)] for o in outs]
    st = torch.cuda.current_stream().cuda_stream
    _lib.call(name, *[_p(a) if torch.is_tensor(a) else a for a in gargs[:-1]], st)
    torch.cuda.synchronize()
    getattr(CPU, name)(*[_p(a) if torch.is_tensor(a) else a for a in args[:-1]], None)
    return [(g.cpu(), o) for g, o in zip(gouts, outs)]


@pytest.mark.parametrize("M,N,K,layout,acc,bias", [
    (4, 3072, 36, "nt", 0, 1), (513, 130, 512, "nt", 0, 1), (96, 1000, 290, "nt", 0, 1),
    (257, 300, 77, "nn", 1, 0), (1536, 512, 4000, "tn", 0, 0), (64, 5, 30000, "tn", 1, 0),
    (1, 192, 0, "nt", 0, 1), (2048, 1536, 128, "nt", 0, 0), (300, 64, 642, "nn", 0, 0),
])
def test_gemm_f32(M, N, K, layout, acc, bias):
    _dev()

    def mk():
        lda_pad, ldc = 8, N + 4
        if layout == "tn":
            A = torch.randn(K, M + lda_pad); sam, sak = 1, M + lda_pad
        else:
            A = torch.randn(M, K + lda_pad); sam, sak = K + lda_pad, 1
        if layout == "nt":
            Bm = torch.randn(N, K + 4); sbk, sbn = 1, K + 4
        else:
            Bm = torch.randn(K, N + 4); sbk, sbn = N + 4, 1
        C = torch.randn(M, ldc)
        b = torch.randn(N) if bias else None
        return [A, sam, sak, Bm, sbk, sbn, C, ldc, b, M, N, K, acc, None], [C]
    (g, c), = _both("pd_gemm_f32", mk)
    tol = 2e-5 * max(1.0, np.sqrt(K))
    assert torch.allclose(g, c, atol=tol, rtol=1e-4), float((g - c).abs().max())


@pytest.mark.parametrize("M,N,K,layout,acc,bias", [
    (4, 3072, 36, "nt", 0, 1), (513, 136, 512, "nt", 1, 1), (2048, 1536, 128, "nt", 0, 0),
    (257, 300, 77, "nn", 1, 0), (16384, 512, 1536, "nn", 0, 0), (1536, 512, 40000, "tn", 0, 0),
    (64, 128, 30001, "tn", 1, 0), (130, 512, 7777, "tn", 0, 0), (1000, 1000, 292, "nt", 0, 1),
])
def test_gemm_tf32_tcgen05(M, N, K, layout, acc, bias):
    """tcgen05 kernel (TF32 operands, fp32 accumulate) vs the fp32 restatement: error bounded by TF32
    rounding of the operands (2^-11 each) -- checked relative to sqrt(K) * |a||b| scale."""
    _dev()
    r4 = lambda v: (v + 3) // 4 * 4          # TMA needs row strides that are multiples of 16 bytes

    def mk():
        ldc = N + 4
        if layout == "tn":
            A = torch.randn(K, r4(M) + 8); sam, sak = 1, r4(M) + 8
        else:
            A = torch.randn(M, r4(K) + 8); sam, sak = r4(K) + 8, 1
        if layout == "nt":
            Bm = torch.randn(N, r4(K) + 4); sbk, sbn = 1, r4(K) + 4
        else:
            Bm = torch.randn(K, r4(N) + 4); sbk, sbn = r4(N) + 4, 1
        C = torch.randn(M, ldc)
        b = torch.randn(N) if bias else None
        return [A, sam, sak, Bm, sbk, sbn, C, ldc, b, M, N, K, acc, None], [C]
    (g, c), = _both("pd_gemm_tf32", mk)
    # TF32 round-to-nearest: per-product relative error ~4e-4 rms -> output error ~4e-4*sqrt(K) rms
    tol = 3e-3 * np.sqrt(K) + 1e-5
    assert torch.allclose(g, c, atol=tol, rtol=0), float((g - c).abs().max())
    assert float((g - c).abs().mean()) < 0.6e-3 * np.sqrt(K) + 1e-5


@pytest.mark.parametrize("M,N,K,layout,acc,bias,dbg", [
    # the routes bench.py times at batch 512: persistent kernel (cfg 925641), TMA-store epilogue (dbg 0) and the
    # plain-store epilogue (dbg bit 4); N of the step's GEMMs (136 = padded pitch head, 194 = pitch + duration-hidden
    # heads, 1536 = note-GRU gates, 2304 = merged note projections), M not a multiple of 128
    (16384, 1536, 512, "nt", 0, 1, 0), (16384, 1536, 512, "nt", 0, 1, 4), (16500, 194, 512, "nt", 0, 1, 0),
    (16500, 194, 512, "nt", 0, 1, 4), (20001, 136, 512, "nt", 0, 1, 0), (8200, 2304, 128, "nt", 0, 1, 0),
    (8200, 2304, 128, "nt", 0, 0, 4), (16384, 512, 1536, "nn", 0, 0, 0), (16390, 512, 194, "nn", 0, 0, 0),
    (16390, 512, 194, "nn", 1, 0, 0), (9000, 1536, 512, "nt", 1, 1, 0),
])
def test_gemm_tf32_persistent_routes(M, N, K, layout, acc, bias, dbg):
    """Direct test of the persistent tcgen05 kernel (two TMEM accumulators, TMA-store / plain-store epilogues): the
    heuristic only selects it for >= 296 tiles, which no other kernel test reaches."""
    _dev()
    r4 = lambda v: (v + 3) // 4 * 4

    def mk():
        ldc = r4(N) + 4
        A = torch.randn(M, r4(K) + 8); sam, sak = r4(K) + 8, 1
        if layout == "nt":
            Bm = torch.randn(N, r4(K) + 4); sbk, sbn = 1, r4(K) + 4
        else:
            Bm = torch.randn(K, r4(N) + 4); sbk, sbn = r4(N) + 4, 1
        C = torch.randn(M, ldc)
        b = torch.randn(N) if bias else None
        return [A, sam, sak, Bm, sbk, sbn, C, ldc, b, M, N, K, acc, dbg * 1000000 + 925641, None], [C]
    (g, c), = _both("pd_gemm_tf32_cfg", mk)
    tol = 3e-3 * np.sqrt(K) + 1e-5
    assert torch.allclose(g[:, :N], c[:, :N], atol=tol, rtol=0), float((g[:, :N] - c[:, :N]).abs().max())
    assert float((g[:, :N] - c[:, :N]).abs().mean()) < 0.6e-3 * np.sqrt(K) + 1e-5
    # the padding of C beyond N must be untouched -- except that a TMA bulk store clips at 16-byte granularity: when N is
    # not a multiple of 4 the columns up to the next multiple of 4 may receive zeros (documented in polydis_b200.h)
    assert torch.equal(g[:, r4(N):], c[:, r4(N):])
    pad = g[:, N:r4(N)]
    assert bool(((pad == c[:, N:r4(N)]) | (pad == 0)).all())


def test_gemm_tf32_heuristic_picks_persistent_for_the_step_shapes():
    """The shape the bench's dominant GEMM uses ([32*512 x 512].[512 x 1536]) through the heuristic entry point."""
    _dev()

    def mk():
        A, Bm, C, b = torch.randn(16384, 512), torch.randn(1536, 512), torch.zeros(16384, 1536), torch.randn(1536)
        return [A, 512, 1, Bm, 1, 512, C, 1536, b, 16384, 1536, 512, 0, None], [C]
    (g, c), = _both("pd_gemm_tf32", mk)
    assert torch.allclose(g, c, atol=3e-3 * np.sqrt(512), rtol=0)


def test_gemm_tf32_rejects_unaligned_operands():
    """Row strides that TMA cannot address are refused (-22) so the host routes to the FFMA kernel."""
    _dev()
    from polydis_b200 import _lib
    A, Bm, C = torch.randn(64, 130).cuda(), torch.randn(32, 130).cuda(), torch.zeros(64, 32).cuda()
    with pytest.raises(RuntimeError):
        _lib.call("pd_gemm_tf32", A.data_ptr(), 130, 1, Bm.data_ptr(), 1, 130, C.data_ptr(), 32, None, 64, 32, 130, 0,
                  torch.cuda.current_stream().cuda_stream)


def test_colsum_seq_skips_masked_steps():
    """Length-aware column sum (bias gradients of the masked note-summary GRU): only (sequence, step) rows below the
    sequence's length are read -- the others are poisoned with NaN here."""
    _dev()
    rng = np.random.RandomState(3)
    R, T, N = 3000, 16, 384
    lengths = torch.from_numpy(rng.randint(0, 18, R).astype(np.int32))          # incl. 0 and > T
    x = torch.randn(R, T, N)
    dead = torch.arange(T)[None, :] >= lengths[:, None]
    x[dead] = float("nan")
    for acc in (0, 1):
        def mk():
            out = torch.full((N,), 0.5)
            return ([x, N, R, T, N, lengths, out, acc, None], [out])
        (g, c), = _both("pd_colsum_seq_f32", mk)
        assert bool(torch.isfinite(g).all())
        assert torch.allclose(g, c, rtol=1e-4, atol=2e-3), float((g - c).abs().max())


def test_colsum_and_transpose():
    _dev()
    (g, c), = _both("pd_colsum_f32",
                    lambda: (lambda X, o: ([X, 200, 5000, 192, o, 1, None], [o]))(torch.randn(5000, 200), torch.randn(192)))
    assert torch.allclose(g, c, atol=2e-3, rtol=1e-4)
    (g, c), = _both("pd_colsum_f32",          # odd width / stride: scalar kernel; ragged row split
                    lambda: (lambda X, o: ([X, 203, 4999, 190, o, 0, None], [o]))(torch.randn(4999, 203), torch.randn(190)))
    assert torch.allclose(g, c, atol=2e-3, rtol=1e-4)
    (g, c), = _both("pd_colsum_f32",          # vector kernel on a column slice of a wider matrix
                    lambda: (lambda X, o: ([X, 384, 30001, 256, o, 0, None], [o]))(torch.randn(30001, 384), torch.zeros(256)))
    assert torch.allclose(g, c, atol=5e-3, rtol=1e-4)
    (g, c), = _both("pd_transpose_f32",
                    lambda: (lambda X, o: ([X, 135, 128, o, None], [o]))(torch.randn(135, 128), torch.zeros(128, 135)))
    assert torch.equal(g, c)
    for T in (1, 2, 15):
        (g, c), = _both("pd_sum_steps_f32",
                        lambda: (lambda X, o: ([X, 16 * 388, 388, T, o, 384, 77, 384, None], [o]))(torch.randn(77, 16, 388), torch.zeros(77, 384)))
        assert torch.allclose(g, c, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("B,H,masked,bcast,hprev", [(7, 64, False, False, True), (33, 1024, False, True, True),
                                                   (130, 128, True, False, True), (130, 128, True, False, False)])
def test_gru_gates_fwd_bwd(B, H, masked, bcast, hprev):
    _dev()
    t = 3

    def mk_f():
        gi, gh = torch.randn(B, 2, 3 * H), torch.randn(B, 3 * H)
        gi2 = torch.randn(B, 3 * H) if bcast else None
        hp = torch.randn(B, 2, H) if hprev else None
        ho, rzn, hn = torch.zeros(B, 2, H), torch.zeros(B, 3 * H), torch.zeros(B, H)
        ln = torch.randint(1, 8, (B,), dtype=torch.int32) if masked else None
        return ([gi, 6 * H, gi2, 3 * H, gh, 3 * H, hp, 2 * H, ho, 2 * H, rzn, 3 * H, hn, H, ln, t, B, H, None],
                [ho, rzn, hn])
    for g, c in _both("pd_gru_gates_fwd", mk_f):
        assert torch.allclose(g, c, atol=2e-6, rtol=1e-5), float((g - c).abs().max())

    def mk_b():
        dh, dh2 = torch.randn(B, H), torch.randn(B, 2, H)
        rzn = torch.rand(B, 3 * H) * 0.98 + 0.01
        rzn[:, 2 * H:] = rzn[:, 2 * H:] * 2 - 1
        hn, hp = torch.randn(B, H), (torch.randn(B, H) if hprev else None)
        dgi, dgh, dhp = torch.zeros(B, 2, 3 * H), torch.zeros(B, 3 * H), torch.zeros(B, H)
        dgi2 = torch.randn(B, 3 * H) if bcast else None
        ln = torch.randint(1, 8, (B,), dtype=torch.int32) if masked else None
        dh3 = torch.randn(B, H) if bcast else None
        return ([dh, H, dh2, 2 * H, dh3, H, rzn, 3 * H, hn, H, hp, H, dgi, 6 * H, dgh, 3 * H, dhp, H, dgi2, 3 * H, ln, t,
                 B, H, None], [dgi, dgh, dhp] + ([dgi2] if bcast else []))
    for g, c in _both("pd_gru_gates_bwd", mk_b):
        assert torch.allclose(g, c, atol=2e-6, rtol=1e-5), float((g - c).abs().max())


def _tokens(R):
    tok = torch.zeros(R, 6, dtype=torch.int32)
    tok[:, 0] = torch.randint(0, 131, (R,))
    tok[:, 1:] = torch.randint(0, 3, (R, 5))
    return tok


def test_grid_prepare_and_note_embed():
    _dev()
    from polydis_b200.synth import synth_batch
    x = torch.from_numpy(synth_batch(5, 2)[0])
    n = 5 * 32

    def mk():
        tok, ln = torch.zeros(n * 16, 6, dtype=torch.int32), torch.zeros(n, dtype=torch.int32)
        pt, dt = torch.zeros(n * 15, dtype=torch.int32), torch.zeros(n * 75, dtype=torch.int32)
        return [x, n, tok, ln, pt, dt, None], [tok, ln, pt, dt]
    for g, c in _both("pd_grid_prepare", mk):
        assert torch.equal(g, c)
    R = 3001

    def mk_e():
        out = torch.zeros(R, 2, 128)
        return [_tokens(R), R, torch.randn(135, 128), torch.randn(128), out, 256, None], [out]
    (g, c), = _both("pd_note_embed_fwd", mk_e)
    assert torch.allclose(g, c, atol=1e-5)

    def mk_b():
        dwt, db = torch.randn(135, 128), torch.randn(128)
        return [_tokens(R), R, torch.randn(R, 128), 128, dwt, db, None], [dwt, db]
    for g, c in _both("pd_note_embed_bwd", mk_b):
        assert torch.allclose(g, c, atol=2e-3, rtol=1e-4), float((g - c).abs().max())


def test_greedy_pick_dur_token_chord():
    _dev()
    R = 1000

    def mk():
        p = torch.randn(R, 132)
        p[::7, 129] = 50.0                     # force some EOS
        p[5, 3] = p[5, 90] = 60.0              # tie -> first index
        d = torch.randn(R, 10)
        d[::3, 0] = d[::3, 1]                  # tie -> bit 0
        tok, lens = torch.zeros(R, 6, dtype=torch.int32), torch.zeros(R, dtype=torch.int32)
        lens[::14] = 2
        return [p, 132, d, 10, R, 4, tok, 6, lens, None], [tok, lens]
    for g, c in _both("pd_greedy_pick", mk):
        assert torch.equal(g, c)

    def mk15():
        tok, lens = torch.zeros(R, 6, dtype=torch.int32), torch.zeros(R, dtype=torch.int32)
        lens[::2] = 3
        return [torch.randn(R, 130), 130, torch.randn(R, 10), 10, R, 15, tok, 6, lens, None], [tok, lens]
    for g, c in _both("pd_greedy_pick", mk15):
        assert torch.equal(g, c)
    (g, c), = _both("pd_dur_token", lambda: (lambda t: ([torch.randn(R, 2), 2, R, t, None], [t]))(torch.zeros(R, 5)))
    assert torch.equal(g, c)

    def mke():           # pick + embedding of the picked token in one launch == pd_greedy_pick then pd_note_embed_fwd
        p = torch.randn(R, 132)
        p[::7, 129] = 50.0
        d = torch.randn(R, 10)
        tok, lens = torch.zeros(R, 6, dtype=torch.int32), torch.zeros(R, dtype=torch.int32)
        ev = torch.zeros(R, 16, 128)[:, 3]               # a slot of the (R,16,128) predicted-note buffer
        return [p, 132, d, 10, R, 4, tok, 6, lens, torch.randn(135, 128), torch.randn(128), ev, 16 * 128, None], [tok, lens, ev]
    (gt, ct), (gl, cl), (ge, ce) = _both("pd_greedy_pick_embed", mke)
    assert torch.equal(gt, ct) and torch.equal(gl, cl)
    assert torch.allclose(ge, ce, atol=1e-6, rtol=0), float((ge - ce).abs().max())

    def mkc():
        tok = torch.zeros(9, 36)
        return ([torch.randn(9, 12), 12, torch.randn(9, 24), 24, torch.randn(9, 12), 12, 9, torch.zeros(24), tok, 36,
                 None], [tok])
    (g, c), = _both("pd_chord_feedback", mkc)
    assert torch.equal(g, c)
    from polydis_b200.synth import synth_batch
    cc = torch.from_numpy(synth_batch(6, 1)[1])

    def mkt():
        r, ch, b = (torch.zeros(48, dtype=torch.int32), torch.zeros(48 * 12, dtype=torch.int32),
                    torch.zeros(48, dtype=torch.int32))
        return [cc, 48, r, ch, b, None], [r, ch, b]
    for g, c in _both("pd_chord_targets", mkt):
        assert torch.equal(g, c)


@pytest.mark.parametrize("NB", [6, 80])
def test_texture_frontend(NB):
    """(NB = 80: 640 (sample, band) items -- the persistent backward CTAs each own several)"""
    _dev()
    from polydis_b200.synth import synth_batch
    pr = torch.from_numpy(synth_batch(NB, 4)[2])

    def mk():
        out = torch.zeros(NB, 10, 8, 29)
        return [pr, torch.randn(10, 48) * 0.2, torch.randn(10) * 0.1, NB, 10, out, None], [out]
    (g, c), = _both("pd_texture_frontend_fwd", mk)
    assert torch.allclose(g, c, atol=1e-4), float((g - c).abs().max())

    def mkb():
        dw, db = torch.zeros(10, 48), torch.zeros(10)
        return [pr, torch.randn(10, 48) * 0.2, torch.randn(10) * 0.1, NB, 10, torch.randn(NB, 10, 8, 29), dw, db,
                None], [dw, db]
    for g, c in _both("pd_texture_frontend_bwd", mkb):
        assert torch.allclose(g, c, atol=5e-3 * np.sqrt(NB / 6), rtol=1e-4), float((g - c).abs().max())
    # training form: the forward records the pooled arg-max, the backward reads it
    wt, bt = torch.randn(10, 48) * 0.2, torch.randn(10) * 0.1

    def mki():
        out, am = torch.zeros(NB, 10, 8, 29), torch.zeros(NB, 10, 8, 29, dtype=torch.int8)
        return [pr, wt, bt, NB, 10, out, am, None], [out, am]
    (go, co), (ga, ca) = _both("pd_texture_frontend_fwd_ix", mki)
    assert torch.allclose(go, co, atol=1e-4) and float((ga != ca).float().mean()) < 1e-3      # (exact ties aside)
    gout = torch.randn(NB, 10, 8, 29)

    def mkbi():
        dw, db = torch.zeros(10, 48), torch.zeros(10)
        return [pr, ca, NB, 10, gout, dw, db, None], [dw, db]
    for g, c in _both("pd_texture_frontend_bwd_ix", mkbi):
        assert torch.allclose(g, c, atol=5e-3 * np.sqrt(NB / 6), rtol=1e-4), float((g - c).abs().max())


@pytest.mark.parametrize("R,C,ignore", [(4001, 130, 130), (9000, 2, 2), (333, 12, -100)])
def test_cross_entropy(R, C, ignore):
    _dev()
    tgt = torch.randint(0, C + (1 if ignore >= 0 else 0), (R,), dtype=torch.int32)
    if ignore >= 0:
        tgt[tgt == C] = ignore

    def mk():
        acc, loss = torch.zeros(2), torch.zeros(1)
        return [torch.randn(R, C + 2) * 3, C + 2, tgt, R, C, ignore, acc, loss, None], [acc, loss]
    (ga, ca), (gl, cl) = _both("pd_ce_fwd", mk)
    assert torch.allclose(gl, cl, rtol=1e-5) and ga[1] == ca[1]

    def mkb():
        acc = torch.tensor([0.0, float((tgt != ignore).sum())])
        d = torch.zeros(R, C)
        return [torch.randn(R, C + 2) * 3, C + 2, tgt, R, C, ignore, acc, torch.tensor([0.7]), d, C, None], [d]
    (g, c), = _both("pd_ce_bwd", mkb)
    assert torch.allclose(g, c, atol=1e-8, rtol=1e-4)


def test_posterior_kernels():
    _dev()
    B, D = 37, 256
    (g, c), = _both("pd_exp_fwd", lambda: (lambda y: ([torch.randn(B * D), B * D, y, None], [y]))(torch.zeros(B * D)))
    assert torch.allclose(g, c, rtol=1e-6)
    (g, c), = _both("pd_mul_f32", lambda: (lambda y: ([torch.randn(B * D), torch.randn(B * D), B * D, y, None], [y]))(
        torch.zeros(B * D)))
    assert torch.allclose(g, c)
    (g, c), = _both("pd_reparam_fwd", lambda: (lambda z: (
        [torch.randn(B, D), torch.rand(B, D), torch.randn(B, D), B, D, z, 2 * D, None], [z]))(torch.zeros(B, 2 * D)))
    assert torch.allclose(g, c, atol=1e-6)
    for g, c in _both("pd_reparam_bwd", lambda: (lambda a, b: (
            [torch.randn(B, 2 * D), 2 * D, torch.randn(B, D), B, D, a, b, None], [a, b]))(torch.zeros(B, D), torch.zeros(B, D))):
        assert torch.allclose(g, c, atol=1e-6)
    (g, c), = _both("pd_kl_fwd", lambda: (lambda o: (
        [torch.randn(B * D), torch.rand(B * D) + 0.5, B * D, o, None], [o]))(torch.zeros(1)))
    assert torch.allclose(g, c, rtol=1e-4)
    for g, c in _both("pd_kl_bwd", lambda: (lambda a, b: (
            [torch.randn(B * D), torch.rand(B * D) + 0.5, B * D, torch.tensor([1.3]), a, b, None], [a, b]))(
            torch.zeros(B * D), torch.zeros(B * D))):
        assert torch.allclose(g, c, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("Q,tf32", [(1, 0), (33, 0), (5000, 0), (33, 1), (5000, 1), (33, 3), (5000, 3)])
def test_dur_decoder_fused(Q, tf32):
    """Weight-resident fused duration GRU (fwd + bwd) vs the numpy restatement; tf32=1 runs the recurrent
    matvecs on the tensor cores (operands rounded to TF32, ~1e-3 relative)."""
    _dev()
    tol = {0: 2e-5, 1: 3e-3, 3: 3e-5}[tf32]
    par = lambda: [torch.randn(192, 5) * 0.3, torch.randn(192) * 0.1, torch.randn(192, 64) * 0.2, torch.randn(192) * 0.1,
                   torch.rand(5), torch.randn(2, 64) * 0.3, torch.randn(2) * 0.1]

    def mk():
        lg, S = torch.zeros(Q, 5, 2), torch.zeros(Q, 6, 72)
        return [torch.randn(Q, 2, 64), 128, Q] + par() + [lg, S, tf32, None], [lg, S]
    (gl, cl), (gs, cs) = _both("pd_dur_decode_fwd", mk)
    same_tok = (gs[:, :, 64:69] == cs[:, :, 64:69]).all(-1).all(-1)
    assert same_tok.float().mean() > (0.97 if tf32 == 1 else 0.995)   # a near-tied bit may flip by rounding
    assert torch.allclose(gl[same_tok], cl[same_tok], atol=tol), float((gl - cl)[same_tok].abs().max())
    assert torch.allclose(gs[same_tok], cs[same_tok], atol=tol)

    # backward on a state buffer produced by the emulation (so both sides see identical tokens)
    S = cs.clone()

    def mkb():
        GX, dh0 = torch.zeros(Q, 6, 264), torch.zeros(Q, 2, 64)
        return [S, torch.randn(Q, 5, 2), Q] + par() + [GX, dh0, 128, tf32, None], [GX, dh0]
    for g, c in _both("pd_dur_decode_bwd", mkb):
        if g.dim() == 3 and g.shape[-1] == 64:
            g, c = g[:, 0], c[:, 0]
        assert torch.allclose(g, c, atol=(5e-3 if tf32 else 3e-5), rtol=1e-4), float((g - c).abs().max())   # bwd: 3 -> TF32


@pytest.mark.parametrize("Q", [1, 33, 512, 2048])
def test_dur_decoder_small_inference_kernel(Q):
    """TF32 calls without saves of <= 2048 notes (a greedy note slot): the four-warps-per-tile kernel (W_hh fragments in
    registers, state and head partials exchanged through shared memory) against the numpy restatement and against the
    one-warp-per-tile kernel (pd_dur_quad_max_notes(0)); the duration head's input rows sit in a wider buffer at an even
    column offset, as in the decoder."""
    _dev()
    from polydis_b200 import _lib
    par = lambda: [torch.randn(192, 5) * 0.3, torch.randn(192) * 0.1, torch.randn(192, 64) * 0.2, torch.randn(192) * 0.1,
                   torch.rand(5), torch.randn(2, 64) * 0.3, torch.randn(2) * 0.1]

    def mk():
        lg = torch.zeros(Q, 5, 2)
        return [torch.randn(Q, 196)[:, 130:194], 196, Q] + par() + [lg, None, 1, None], [lg]
    outs = []
    for limit in (2048, 0):
        assert _lib.lib.pd_dur_quad_max_notes(limit) == 0
        try:
            (gl, cl), = _both("pd_dur_decode_fwd", mk)
        finally:
            _lib.lib.pd_dur_quad_max_notes(2048)
        outs.append(gl)
        bits_g, bits_c = gl[:, :, 1] > gl[:, :, 0], cl[:, :, 1] > cl[:, :, 0]
        same = (bits_g == bits_c).all(-1)
        assert same.float().mean() > 0.97                      # a near-tied bit may flip by TF32 rounding
        assert torch.allclose(gl[same], cl[same], atol=3e-3), float((gl - cl)[same].abs().max())
    same = ((outs[0][:, :, 1] > outs[0][:, :, 0]) == (outs[1][:, :, 1] > outs[1][:, :, 0])).all(-1)
    assert same.float().mean() > 0.99                          # same TF32 operands, different summation order
    assert torch.allclose(outs[0][same], outs[1][same], atol=2e-5), float((outs[0] - outs[1])[same].abs().max())


def test_prmat_grid_conversions():
    _dev()
    from polydis_b200.synth import synth_batch
    x, c, pr = synth_batch(7, 13)
    pr[0, 3, 20:37] = 2                                    # one overflowing step

    def mk():
        out, ovf = torch.zeros(7 * 32, 16, 6, dtype=torch.int64), torch.zeros(1, dtype=torch.int32)
        return [torch.from_numpy(pr), 7 * 32, out, ovf, None], [out, ovf]
    for g, c_ in _both("pd_prmat_to_grid", mk):
        assert torch.equal(g, c_)
    tok = torch.from_numpy(x[:, :, 1:, :].astype(np.int32)).contiguous()

    def mk2():
        out = torch.zeros(7 * 32, 128)
        return [tok, 7 * 32, out, None], [out]
    (g, c_), = _both("pd_grid_to_prmat", mk2)
    assert torch.equal(g, c_)


def test_optimizer_tail_kernels():
    _dev()
    n = 100003
    (g, c), = _both("pd_sumsq_f32", lambda: (lambda o: ([torch.randn(n), n, o, None], [o]))(torch.ones(1)))
    assert torch.allclose(g, c, rtol=1e-4)

    def mk():
        p, m, v = torch.randn(n), torch.randn(n) * 0.1, torch.rand(n) * 0.01
        return ([p, torch.randn(n), m, v, n, torch.tensor([4.0e5]), torch.tensor([7], dtype=torch.int32), 1e-3, 0.9999,
                 1e-5, 0.9, 0.999, 1e-8, 1.0, None], [p, m, v])
    for g, c in _both("pd_adam_clip_step", mk):
        assert torch.allclose(g, c, atol=1e-7, rtol=1e-5)
    (g, c), = _both("pd_counter_inc", lambda: (lambda o: ([o, None], [o]))(torch.tensor([41], dtype=torch.int32)))
    assert int(g) == 42 and int(c) == 42


@pytest.mark.parametrize("fused", [True, False])
def test_graphed_train_step_matches_eager(fused):
    """CUDA-graph replay of the whole training step (fwd + bwd + clip + Adam) == the same step issued eagerly from
    identical weights, with the reparameterisation noise injected (GraphedTrainStep(inject_eps=True)): two consecutive
    steps (the second depends on the first update) and the parameters afterwards."""
    dev = _dev()
    import random
    from polydis_b200.graphs import GraphedTrainStep
    from polydis_b200.model import DisentangleVAE
    from polydis_b200.optim import FusedClipAdam
    from polydis_b200.synth import synth_batch
    from polydis_b200.weights import make_state_dict
    B = 8
    x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 31))
    torch.manual_seed(7)
    eps = [(torch.randn(B, 256, device=dev), torch.randn(B, 256, device=dev)) for _ in range(2)]
    losses, finals = [], []
    for graphed in (False, True):
        m = DisentangleVAE.init_model(device=dev)
        m.load_state_dict(make_state_dict(2))
        m.to(dev).train()
        params = list(m.parameters())
        if fused:
            opt = FusedClipAdam(params, lr=1e-3, clip=1.0, lr_gamma=0.9999, lr_min=1e-5)
        else:
            opt = torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True)
        random.seed(7)
        out = []
        if graphed:
            # warm-up + capture run real steps on the first batch; restore_after_capture puts weights / moments back
            step = GraphedTrainStep(m, opt, B, warmup=2, inject_eps=True).capture(x, c, pr)
            for e in eps:
                step.eps[0].copy_(e[0]); step.eps[1].copy_(e[1])
                out.append(step(x, c, pr).clone())
        else:
            for e in eps:
                opt.zero_grad()
                l = m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5), eps=e)
                l[0].backward()
                if fused:
                    opt.reducer.finish()
                else:
                    torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)
                opt.step()
                out.append(torch.stack([v.detach() for v in l]))
        torch.cuda.synchronize()
        losses.append(torch.stack(out).cpu())
        finals.append([p.detach().clone().cpu() for p in params])
    # identical inputs, weights and noise: only the summation order of split-K / atomic reductions differs.  Step 1
    # therefore agrees to fp32 reassociation noise; Adam's g / (|g| + eps) turns that noise into lr-sized differences on
    # elements whose gradient is ~0, so step 2 and the final weights are held to a correspondingly looser bound.
    assert torch.allclose(losses[0][0], losses[1][0], rtol=2e-5, atol=1e-6), (losses[0][0], losses[1][0])
    assert torch.allclose(losses[0][1], losses[1][1], rtol=1e-3, atol=1e-5), (losses[0][1], losses[1][1])
    assert float(losses[0][1][0]) < float(losses[0][0][0])            # and the update reduced the loss
    n_bad = sum(int(((a - b_).abs() > 4e-4).sum()) for a, b_ in zip(*finals))
    n_all = sum(a.numel() for a in finals[0])
    assert n_bad <= 1e-3 * n_all, (n_bad, n_all)


def test_graphed_train_step_prefetch_pipeline():
    """prefetch / step_prefetched: the batch staged from pinned host memory is the one the replay trains on, and
    the next batch is copied behind it."""
    dev = _dev()
    from polydis_b200.graphs import GraphedTrainStep
    from polydis_b200.model import DisentangleVAE
    from polydis_b200.synth import synth_batch
    host = [[torch.from_numpy(a).pin_memory() for a in synth_batch(8, seed)] for seed in (31, 32)]
    m = DisentangleVAE.init_model(device=dev).to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True, capturable=True)
    g = GraphedTrainStep(m, opt, 8, warmup=1).capture(*[t.to(dev) for t in host[0]])
    g.prefetch(*host[1])
    l1 = g.step_prefetched(next_batch=host[0])
    torch.cuda.synchronize()
    assert torch.equal(g.x.cpu(), host[1][0]) and torch.equal(g.pr.cpu(), host[1][2]) and torch.isfinite(l1).all()
    l2 = g.step_prefetched()
    torch.cuda.synchronize()
    assert torch.equal(g.x.cpu(), host[0][0]) and torch.equal(g.c.cpu(), host[0][1]) and torch.isfinite(l2).all()


def test_tf32x3_gemm_is_fp32_class():
    """Error-compensated 3xTF32 GEMM (tensor cores) vs fp64: relative error ~1e-6, i.e. fp32-class, against
    ~3e-4 for plain TF32."""
    dev = _dev()
    from polydis_b200 import ops
    torch.manual_seed(0)
    A, W, b = torch.randn(700, 512, device=dev), torch.randn(194, 512, device=dev), torch.randn(194, device=dev)
    ref = A.double() @ W.double().t() + b.double()
    errs = {}
    for mode in ("tf32", "tf32x3", "fp32"):
        out = torch.empty(700, 196, device=dev)[:, :194]
        with ops.precision(mode):
            ops.gemm_nt(A, W, out, b)
        errs[mode] = float((out.double() - ref).abs().max() / ref.abs().max())
    assert errs["tf32x3"] < 3e-6 and errs["fp32"] < 1e-6 and errs["tf32"] > 10 * errs["tf32x3"], errs


@pytest.mark.parametrize("M,N,K,layout,acc", [(513, 136, 512, "nt", 1), (2048, 1536, 128, "nt", 0), (257, 304, 80, "nn", 1),
                                              (1536, 512, 40000, "tn", 0), (64, 128, 30008, "tn", 1)])
def test_gemm_bf16_tcgen05(M, N, K, layout, acc):
    """bf16-operand tcgen05 GEMM (kind::f16, fp32 accumulate) and the fp32->bf16 conversion kernel vs the numpy
    restatement (products of bf16 values are exact in fp32; only the summation order differs)."""
    _dev()
    from polydis_b200 import _lib
    r8 = lambda v: (v + 7) // 8 * 8
    torch.manual_seed(1)
    if layout == "tn":
        A32 = torch.randn(K, r8(M) + 8); sam, sak = 1, r8(M) + 8
    else:
        A32 = torch.randn(M, r8(K) + 8); sam, sak = r8(K) + 8, 1
    if layout == "nt":
        B32 = torch.randn(N, r8(K) + 8); sbk, sbn = 1, r8(K) + 8
    else:
        B32 = torch.randn(K, r8(N) + 8); sbk, sbn = r8(N) + 8, 1
    outs = []
    for X in (A32, B32):
        def mk(X=X):
            o = torch.zeros(X.shape, dtype=torch.int16)
            return [X, X.shape[1], X.shape[0], X.shape[1], o, X.shape[1], None], [o]
        (g, c), = _both("pd_f32_to_bf16", mk)
        assert torch.equal(g, c)
        assert torch.equal(g.view(torch.bfloat16), X.to(torch.bfloat16))
        outs.append(c)

    def mkg():
        C = torch.randn(M, N + 4)
        return [outs[0], sam, sak, outs[1], sbk, sbn, C, N + 4, torch.randn(N), M, N, K, acc, None], [C]
    (g, c), = _both("pd_gemm_bf16", mkg)
    assert torch.allclose(g, c, atol=2e-5 * max(1.0, np.sqrt(K)), rtol=1e-4), float((g - c).abs().max())


@pytest.mark.parametrize("B,H,masked,bcast,save", [(7, 64, False, False, True), (300, 512, False, True, True),
                                                   (130, 128, True, False, True), (513, 1024, False, True, False)])
def test_gru_step_fused_tcgen05(B, H, masked, bcast, save):
    """Fused recurrent GEMM + gate epilogue vs GEMM-then-gates in numpy (TF32 operand rounding on the GEMM)."""
    _dev()
    t = 3

    def mk():
        hp, w, b = torch.randn(B, 2, H) * 0.5, torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
        gi = torch.randn(B, 2, 3 * H)
        gi2 = torch.randn(B, 3 * H) if bcast else None
        ho = torch.zeros(B, 2, H)
        rzn, hn = (torch.zeros(B, 3 * H), torch.zeros(B, H)) if save else (None, None)
        ln = torch.randint(1, 8, (B,), dtype=torch.int32) if masked else None
        return ([hp, 2 * H, w, H, b, gi, 6 * H, gi2, 3 * H, ho, 2 * H, rzn, 3 * H, hn, H, ln, t, B, H, None],
                [ho] + ([rzn, hn] if save else []))
    for g, c in _both("pd_gru_step_tf32", mk):
        assert torch.allclose(g, c, atol=3e-3, rtol=0), float((g - c).abs().max())


@pytest.mark.parametrize("R,T,reverse,passes", [(5, 16, 0, 1), (1000, 16, 1, 1), (333, 16, 0, 3), (64, 7, 1, 3)])
def test_gru128_resident(R, T, reverse, passes):
    """Weight-resident variable-length GRU (fwd: TF32 / 3xTF32 matvecs; bwd: TF32) vs the numpy restatement."""
    _dev()
    H = 128
    torch.manual_seed(4)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
    lengths = torch.randint(1, T + 1, (R,), dtype=torch.int32)
    lengths[0] = T

    def mk():
        gi = torch.randn(R, T, 3 * H)
        h_all, rzn, hn = torch.zeros(R, T, H), torch.zeros(R, T, 3 * H), torch.zeros(R, T, H)
        return ([gi, T * 3 * H, 3 * H, lengths, w, b, h_all, T * H, H, rzn, T * 3 * H, 3 * H, hn, T * H, H, R, T, reverse,
                 passes, None], [h_all, rzn, hn])
    res = _both("pd_gru128_fwd", mk)
    tol = 2e-5 if passes == 3 else 4e-3
    for g, c in res:
        assert torch.allclose(g, c, atol=tol, rtol=0), float((g - c).abs().max())
    h_all, rzn, hn = (c for _, c in res)                 # emulation outputs feed the backward on both sides

    def mkb():
        dgi, dgh = torch.zeros(R, T, 3 * H), torch.zeros(R, T, 3 * H)
        return ([torch.randn(R, T, H), T * H, H, h_all, T * H, H, rzn, T * 3 * H, 3 * H, hn, T * H, H, lengths, w, dgi,
                 T * 3 * H, 3 * H, dgh, T * 3 * H, 3 * H, R, T, reverse, None], [dgi, dgh])
    for g, c in _both("pd_gru128_bwd", mkb):
        assert torch.allclose(g, c, atol=5e-3, rtol=1e-3), float((g - c).abs().max())


@pytest.mark.parametrize("R,reverse,passes", [(1000, 0, 1), (333, 1, 3), (21, 0, 1)])
def test_gru128_resident_visits_rows_in_a_given_order(R, reverse, passes):
    """pd_gru128_fwd_perm: the rows are processed in the order of a permutation (greedy decoding sorts them by predicted
    length: a 16-row tile runs to its longest sequence) -- every output row is bit-for-bit what pd_gru128_fwd writes."""
    _dev()
    from polydis_b200 import _lib
    H, T = 128, 16
    torch.manual_seed(6)
    w, b = (torch.randn(3 * H, H) / np.sqrt(H)).cuda(), (torch.randn(3 * H) * 0.1).cuda()
    lengths = torch.randint(0, T + 1, (R,), dtype=torch.int32).cuda()
    gi = torch.randn(R, T, 3 * H).cuda()
    perm, inv, table = (torch.zeros(n, dtype=torch.int32).cuda() for n in (R, R, 64))
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("pd_pack_order", lengths.data_ptr(), R, perm.data_ptr(), inv.data_ptr(), table.data_ptr(), st)
    outs = []
    for use_perm in (False, True):
        h_all, rzn, hn = (torch.zeros(R, T, n).cuda() for n in (H, 3 * H, H))
        args = [gi.data_ptr(), T * 3 * H, 3 * H, lengths.data_ptr(), w.data_ptr(), b.data_ptr(), h_all.data_ptr(), T * H, H,
                rzn.data_ptr(), T * 3 * H, 3 * H, hn.data_ptr(), T * H, H, R, T, reverse, passes]
        _lib.call("pd_gru128_fwd_perm" if use_perm else "pd_gru128_fwd", *(args + ([perm.data_ptr()] if use_perm else [])), st)
        torch.cuda.synchronize()
        outs.append((h_all, rzn, hn))
    ls = lengths[perm.long()]
    assert bool((ls[:-1] >= ls[1:]).all())
    for a, c in zip(*outs):
        assert torch.equal(a, c)


@pytest.mark.parametrize("order", [0, 1])
def test_tf32_split3(order):
    """[hi | hi | lo] / [hi | lo | hi] operand builder of the single-launch 3xTF32 GEMM (zero padded to K % 4 == 0)."""
    _dev()
    (g, c), = _both("pd_tf32_split3",
                    lambda: (lambda X, o: ([X, 136, 77, 130, o, 400, order, None], [o]))(torch.randn(77, 136), torch.full((77, 400), 7.0)))
    assert torch.equal(g, c)
    kp = 132
    hi, x = c[:, :kp], torch.zeros(77, kp)
    lo = c[:, kp:2 * kp] if order else c[:, 2 * kp:3 * kp]
    assert torch.equal((hi.view(torch.int32) & 0x1FFF), torch.zeros(77, kp, dtype=torch.int32))   # 10-bit mantissa
    assert float(lo.abs().max()) < 2 ** -10 * 6 and torch.equal(c[:, 3 * kp:], torch.full((77, 400 - 3 * kp), 7.0))


def test_gru_gates_fwd_split3():
    """Inference gate kernel that also emits the [hi | hi | lo] operand of the new state."""
    _dev()
    B, H = 70, 128

    def mk():
        hp, h, h3 = torch.randn(B, H), torch.zeros(B, H), torch.zeros(B, 3 * H + 4)
        return ([torch.randn(B, 3 * H), 3 * H, torch.randn(B, 3 * H), 3 * H, torch.randn(B, 3 * H), 3 * H, hp, H, h, H, None, 0, B, H,
                 h3, 3 * H + 4, None], [h, h3])
    (gh_, ch_), (g3, c3) = _both("pd_gru_gates_fwd_split3", mk)
    assert torch.allclose(gh_, ch_, atol=2e-6)
    hi, lo = g3[:, :H], g3[:, 2 * H:3 * H]
    assert torch.equal(hi, g3[:, H:2 * H]) and torch.equal((hi + lo), gh_)          # exact split of the GPU state
    assert torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros(B, H, dtype=torch.int32))
    # (hi itself may differ from the emulation's by one TF32 ulp where the two states straddle a rounding boundary)
    assert torch.allclose(hi + lo, c3[:, :H] + c3[:, 2 * H:3 * H], atol=3e-6) and torch.equal(g3[:, 3 * H:], c3[:, 3 * H:])


@pytest.mark.parametrize("B,H,bcast,save", [(300, 128, True, True), (4100, 512, True, True), (129, 64, False, False),
                                             (512, 1024, True, True), (512, 512, False, True)])
def test_gru_step_tma(B, H, bcast, save):
    """Persistent fused GRU step with TMA epilogue I/O (pd_gru_step_tma) vs the numpy restatement."""
    _dev()
    torch.manual_seed(5)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1

    def mk():
        ho = torch.zeros(B, H)
        rzn = torch.zeros(B, 3 * H) if save else None
        hn = torch.zeros(B, H) if save else None
        return ([torch.randn(B, H), H, w, H, b, torch.randn(B, 3 * H), 3 * H, torch.randn(B, 3 * H) if bcast else None, 3 * H, ho, H,
                 rzn, 3 * H, hn, H, B, H, None], [ho] + ([rzn, hn] if save else []))
    for g, c in _both("pd_gru_step_tma", mk):
        assert torch.allclose(g, c, atol=4e-3, rtol=0), float((g - c).abs().max())


@pytest.mark.parametrize("units", [0, 32, 64])
@pytest.mark.parametrize("B,H,bcast", [(512, 1024, True), (512, 512, False), (300, 128, True), (2048, 1024, False)])
def test_gru_step_tma_bf16_operands(B, H, bcast, units):
    """Fused step with bf16 copies of h_prev / W_hh as the tcgen05 operands (kind::f16, fp32 accumulate; the batch-sized
    recurrences in training): both sides multiply the SAME bf16 values, so the comparison is tight; the kernel also emits
    the bf16 copy of the new state."""
    _dev()
    torch.manual_seed(5)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
    hp = torch.randn(B, H)
    wb, hb = w.to(torch.bfloat16), hp.to(torch.bfloat16)

    def mk():
        ho, hbo = torch.zeros(B, H), torch.zeros(B, H, dtype=torch.bfloat16)
        rzn, hn = torch.zeros(B, 3 * H), torch.zeros(B, H)
        return ([hb, H, wb, H, b, torch.randn(B, 3 * H), 3 * H, torch.randn(B, 3 * H) if bcast else None, 3 * H, hp, H, ho, H,
                 hbo, H, rzn, 3 * H, hn, H, B, H] + ([units] if units else []) + [None], [ho, hbo, rzn, hn])
    # units: 0 = the default entry point (32-unit tiles); 32 / 64 = pd_gru_step_tma_bf16_units
    (gh, ch), (gb, cb), (gr, cr), (gn, cn) = _both("pd_gru_step_tma_bf16_units" if units else "pd_gru_step_tma_bf16", mk)
    assert torch.allclose(gh, ch, atol=2e-5, rtol=0), float((gh - ch).abs().max())
    assert torch.allclose(gr, cr, atol=2e-5, rtol=0) and torch.allclose(gn, cn, atol=2e-5, rtol=0)
    assert float((gb.float() - cb.float()).abs().max()) <= 2 ** -7           # (one bf16 ulp where h' straddles a boundary)
    assert float((gb.float() - gh).abs().max()) <= 2 ** -8 * float(gh.abs().max())


def test_gru_gates_bwd_zb_bf16_copy():
    """Gate-gradient kernel that also clears the split-K accumulator and writes the bf16 copy of dgh."""
    _dev()
    B, H = 512, 1024

    def mk():
        dgi, dgh, dhp = torch.zeros(B, 3 * H), torch.zeros(B, 3 * H), torch.zeros(B, H)
        zo, gb = torch.full((B, H), 3.0), torch.zeros(B, 3 * H, dtype=torch.bfloat16)
        rzn = torch.rand(B, 3 * H) * 0.9 + 0.05
        return ([torch.randn(B, H), H, torch.randn(B, H), H, torch.randn(B, H), H, rzn, 3 * H, torch.randn(B, H), H,
                 torch.randn(B, H), H, dgi, 3 * H, dgh, 3 * H, dhp, H, None, 0, B, H, zo, H, gb, 3 * H, None], [dgi, dgh, dhp, zo, gb])
    res = _both("pd_gru_gates_bwd_zb", mk)
    for g, c in res[:4]:
        assert torch.allclose(g, c, atol=1e-5, rtol=1e-5)
    assert float(res[3][0].abs().max()) == 0.0
    assert float((res[4][0].float() - res[4][1].float()).abs().max()) <= 2 ** -7 * float(res[1][0].abs().max())


@pytest.mark.parametrize("B,H,bcast", [(700, 128, True), (4100, 512, True), (513, 1024, False)])
def test_gru_step_tma3_split3_inference(B, H, bcast):
    """Fused inference step of the 3xTF32 decode path (pd_gru_step_tma3): [hi|hi|lo] . [hi|lo|hi] on tcgen05, expf / tanhf
    gates, in-place state update and the split of the new state -- fp32-class agreement with the fp64-accumulated
    restatement, and an exact hi + lo == h split."""
    _dev()
    torch.manual_seed(6)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
    hp0 = torch.randn(B, H) * 0.5

    def split3(x, order):
        out = torch.zeros(x.shape[0], 3 * x.shape[1])
        CPU.pd_tf32_split3(x.data_ptr(), x.shape[1], x.shape[0], x.shape[1], out.data_ptr(), 3 * x.shape[1], order, None)
        return out
    a3, w3 = split3(hp0, 0), split3(w, 1)

    def mk():
        h, h3 = torch.zeros(B, H), torch.zeros(B, 3 * H)       # (the model updates the state in place; covered by the
        return ([a3, 3 * H, w3, 3 * H, b, torch.randn(B, 3 * H), 3 * H,   #  decode parity tests)
                 torch.randn(B, 3 * H) if bcast else None, 3 * H, hp0, H, h, H, h3, 3 * H, B, H, None], [h, h3])
    (gh_, ch_), (g3, c3) = _both("pd_gru_step_tma3", mk)
    # K = 3H products of the error-compensated form: error ~ sqrt(K) * 2^-22 -- fp32-class (single-pass TF32: ~1e-3)
    assert torch.allclose(gh_, ch_, atol=3e-5, rtol=0), float((gh_ - ch_).abs().max())
    hi, lo = g3[:, :H], g3[:, 2 * H:]
    assert torch.equal(hi, g3[:, H:2 * H]) and torch.equal(hi + lo, gh_)
    assert torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros(B, H, dtype=torch.int32))


@pytest.mark.parametrize("B,H,K,save", [(300, 128, 128, True), (4100, 512, 128, True), (129, 64, 36, False),
                                        (512, 512, 128, False), (512, 512, 128, True), (1300, 512, 128, False)])
def test_gru_step_tmax_folded_x_projection(B, H, K, save):
    """Training form of the fused step with the x-projection as a second K segment (pd_gru_step_tmax): x rows with a wide
    stride (slot n of a (B,16,K) embedding buffer), w_x a column slice of a wider W_ih, saved gates for the backward.
    Row counts that leave half the SMs without a 64-unit tile (<= 74 tiles: the 512-row greedy slot) take 32-unit tiles."""
    _dev()
    torch.manual_seed(9)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
    wih = torch.randn(3 * H, 64 + K) / np.sqrt(K)                       # W_ih; the step uses columns 64..64+K
    x_all = torch.randn(B, 4, K)                                        # step 2 of a (B,4,K) buffer

    def mk():
        ho = torch.zeros(B, H)
        rzn = torch.zeros(B, 3 * H) if save else None
        hn = torch.zeros(B, H) if save else None
        return ([torch.randn(B, H), H, w, H, x_all[:, 2], 4 * K, wih[:, 64:], 64 + K, K, b, torch.randn(B, 3 * H), 3 * H, ho, H,
                 rzn, 3 * H, hn, H, B, H, None], [ho] + ([rzn, hn] if save else []))
    for g, c in _both("pd_gru_step_tmax", mk):
        assert torch.allclose(g, c, atol=6e-3, rtol=0), float((g - c).abs().max())


@pytest.mark.parametrize("B,H,K", [(700, 128, 128), (4100, 512, 128), (513, 512, 36)])
def test_gru_step_tma3x_folded_x_projection(B, H, K):
    """pd_gru_step_tma3x: the fused 3xTF32 inference step with the x-projection as a second K segment of the main loop
    (r / z products accumulate onto the h-projection, the n product into its own TMEM block)."""
    _dev()
    torch.manual_seed(8)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
    wx = torch.randn(3 * H, K) / np.sqrt(K)
    hp0, x0 = torch.randn(B, H) * 0.5, torch.randn(B, K)

    def split3(x, order):
        kp = (x.shape[1] + 3) // 4 * 4
        out = torch.zeros(x.shape[0], 3 * kp)
        CPU.pd_tf32_split3(x.data_ptr(), x.shape[1], x.shape[0], x.shape[1], out.data_ptr(), 3 * kp, order, None)
        return out
    a3, w3, x3, wx3 = split3(hp0, 0), split3(w, 1), split3(x0, 0), split3(wx, 1)
    K2 = x3.shape[1]

    def mk():
        h, h3 = torch.zeros(B, H), torch.zeros(B, 3 * H)
        return ([a3, 3 * H, w3, 3 * H, x3, K2, wx3, K2, K2, b, torch.randn(B, 3 * H), 3 * H, hp0, H, h, H, h3, 3 * H, B, H, None],
                [h, h3])
    (gh_, ch_), (g3, c3) = _both("pd_gru_step_tma3x", mk)
    assert torch.allclose(gh_, ch_, atol=3e-5, rtol=0), float((gh_ - ch_).abs().max())
    assert torch.equal(g3[:, :H] + g3[:, 2 * H:], gh_) and torch.equal(g3[:, :H], g3[:, H:2 * H])


def test_select_rows_kernels():
    """Device-flag row select (scheduled sampling with the plan as device data) and its gradient routing."""
    _dev()
    for flag in (0, 1):
        def mk2():
            a, out = torch.randn(37, 132), torch.zeros(37, 128)
            return [a, 132, torch.randn(37, 128), 128, torch.tensor([flag], dtype=torch.int32), out, 128, 37, 128, None], [out]
        (g, c), = _both("pd_select_rows", mk2)
        assert torch.equal(g, c)

        def mkb():
            da, db = torch.ones(37, 128), torch.ones(37, 128)
            return [torch.randn(37, 128), 128, torch.tensor([flag], dtype=torch.int32), da, 128, db, 128, 37, 128, None], [da, db]
        for g, c in _both("pd_select_rows_bwd", mkb):
            assert torch.equal(g, c)


def test_graphed_train_step_device_plan():
    """One CUDA graph for every teacher-forcing ratio: decisions uploaded per step, selected on the device."""
    dev = _dev()
    import random
    from polydis_b200.graphs import GraphedTrainStep
    from polydis_b200.model import DisentangleVAE
    from polydis_b200.synth import synth_batch
    x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(8, 31))
    m = DisentangleVAE.init_model(device=dev).to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=0.0, fused=True, capturable=True)
    random.seed(3)
    g = GraphedTrainStep(m, opt, 8, tfr=(0.5, 0.5, 0.5), warmup=1, device_plan=True).capture(x, c, pr)
    l_mixed = float(g(x, c, pr)[0])
    g.set_tfr(1., 1., 1.)
    l_tf = float(g(x, c, pr)[0])
    ref = float(m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))[0])
    assert np.isfinite(l_mixed) and abs(l_tf - ref) < 5e-2 * abs(ref)      # same decisions, different noise draw
