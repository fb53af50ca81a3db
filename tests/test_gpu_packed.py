"""-m gpu: the packed note level (csrc/packed.cu and the *_rows entry points: length-sorted rows, slot-major buffers,
device live-row table) -- every entry point through the C-ABI against its numpy restatement in tests/cpu_backend.py.
Buffers that a kernel only partly writes start from the same contents on both sides, so "dead rows stay untouched" is
part of what is compared."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.test_gpu_kernels import _both, _dev


def _table(R, counts):
    """pack table for explicit live counts c[t] (t = 0..16)."""
    t = torch.zeros(64, dtype=torch.int32)
    for i, c in enumerate(counts):
        cp = min(R, (c + 127) // 128 * 128)
        t[i], t[17 + i], t[34 + i] = c, cp, 6 * cp
    return t


@pytest.mark.parametrize("R", [192, 4096, 16384])
def test_pack_order_and_grid(R):
    _dev()
    rng = np.random.RandomState(R)
    lengths = torch.from_numpy(np.where(rng.rand(R) < 0.4, 2, rng.randint(3, 17, R)).astype(np.int32))
    tok = torch.from_numpy(rng.randint(0, 131, (R, 16, 6)).astype(np.int32))

    def mk():
        o = [torch.zeros(R, dtype=torch.int32), torch.zeros(R, dtype=torch.int32), torch.zeros(64, dtype=torch.int32)]
        return ([lengths, R] + o + [None], o)
    (gp, cp_), (gi, ci), (gt, ct) = _both("pd_pack_order", mk)
    assert torch.equal(gp, cp_) and torch.equal(gi, ci) and torch.equal(gt, ct)
    ls = lengths[cp_.long()]
    assert bool((ls[:-1] >= ls[1:]).all()) and int(ct[0]) == R and int(ct[17]) == R
    perm = cp_

    def mk3():
        o = [torch.zeros(16 * R * 6, dtype=torch.int32), torch.zeros(15 * R, dtype=torch.int32),
             torch.zeros(15 * R * 5, dtype=torch.int32), torch.zeros(R, dtype=torch.int32)]
        return ([tok, lengths, perm, R] + o + [None], o)
    for g, c in _both("pd_pack_grid", mk3):
        assert torch.equal(g, c)


@pytest.mark.parametrize("M,N,K,layout,slot_rows,counts,bias", [
    (15 * 512, 196, 512, "nt", 512, [512, 300, 129, 128, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], 1),     # one-shot kernel
    (15 * 8192, 200, 512, "nt", 8192, [8192, 5000, 4097, 2000, 600, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0], 1),   # persistent, TMA store
    (15 * 8192, 512, 200, "nn", 8192, [8192, 5000, 4097, 2000, 600, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0], 0),   # persistent NN
    (4096, 512, 1536, "nn", 4096, [1000], 0),                                                       # per-step dgh.W_hh
    (15 * 192, 128, 1536, "nn", 192, [192, 100, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 5], 0),          # slots not tile-aligned
])
def test_gemm_rows_m_side(M, N, K, layout, slot_rows, counts, bias):
    """pred = 1: dead 128-row tiles of A / C are skipped, their C rows keep their contents."""
    _dev()
    tab = _table(slot_rows, counts)

    def mk():
        a = torch.randn(M, K)
        b = torch.randn(N, K) if layout == "nt" else torch.randn(K, N)
        ldc = (N + 3) // 4 * 4
        c = torch.full((M, ldc), 7.0)
        bs = torch.randn(N) if bias else None
        sb = (1, K) if layout == "nt" else (N, 1)
        return ([a, K, 1, b, sb[0], sb[1], c, ldc, bs, M, N, K, 0, 1, tab[17:], slot_rows, None], [c])
    for g, c in _both("pd_gemm_tf32_rows", mk):
        g, c = g[:, :N], c[:, :N]
        assert torch.allclose(g, c, atol=8e-2 * np.sqrt(K / 512), rtol=1e-2), float((g - c).abs().max())
        assert bool((c == 7.0).any()) or min(counts) > 0


@pytest.mark.parametrize("M,N,Kslots,slot_rows,counts,acc", [
    (1536, 512, 14, 1024, [1024, 1000, 513, 200, 64, 1, 0, 0, 0, 0, 0, 0, 0, 0], 1),
    (194, 512, 15, 2048, [2048, 1500, 1000, 400, 100, 33, 0, 0, 0, 0, 0, 0, 0, 0, 0], 0),
    (264, 72, 15, 6 * 192, [6 * 192, 6 * 128, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], 0),
])
def test_gemm_rows_k_side(M, N, Kslots, slot_rows, counts, acc):
    """pred = 2 (weight gradients): only live 32-row k-blocks of the slot-major contraction index are accumulated."""
    _dev()
    K = Kslots * slot_rows
    tab = torch.zeros(64, dtype=torch.int32)
    for i, c in enumerate(counts):
        tab[i] = min(slot_rows, (c + 31) // 32 * 32)

    def mk():
        lda = (M + 3) // 4 * 4                             # (TMA needs 16-byte row pitches: the model's slabs are padded too)
        a = torch.randn(K, lda)[:, :M]
        b = torch.randn(K, N)
        # dead rows hold garbage the kernel must never touch
        live = (torch.arange(K) % slot_rows) < tab[(torch.arange(K) // slot_rows).long()]
        a[~live] = float("nan")
        b[~live] = float("nan")
        c = torch.full((M, N), 3.0) if acc else torch.zeros(M, N)
        return ([a, 1, lda, b, N, 1, c, N, None, M, N, K, acc, 2, tab, slot_rows, None], [c])
    for g, c in _both("pd_gemm_tf32_rows", mk):
        assert bool(torch.isfinite(g).all())
        scale = np.sqrt(max(1, sum(counts)))
        assert torch.allclose(g, c, atol=3e-3 * scale, rtol=1e-2), float((g - c).abs().max())


@pytest.mark.parametrize("B,H,K2,nrows", [(4100, 512, 128, 1300), (300, 128, 128, 300), (4096, 512, 128, 0), (520, 64, 36, 129)])
def test_gru_step_tmax_rows(B, H, K2, nrows):
    _dev()
    torch.manual_seed(3)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
    wx = torch.randn(3 * H, K2) / np.sqrt(K2)
    nr = torch.tensor([nrows], dtype=torch.int32)

    def mk():
        ho, rzn, hn = torch.full((B, H), 5.0), torch.full((B, 3 * H), 5.0), torch.full((B, H), 5.0)
        return ([torch.randn(B, H), H, w, H, torch.randn(B, K2), K2, wx, K2, K2, b, torch.randn(B, 3 * H), 3 * H, ho, H, rzn, 3 * H,
                 hn, H, B, H, nr, None], [ho, rzn, hn])
    for g, c in _both("pd_gru_step_tmax_rows", mk):
        assert torch.allclose(g, c, atol=6e-3, rtol=0), float((g - c).abs().max())


@pytest.mark.parametrize("B,H,nrows", [(700, 512, 384), (256, 128, 0), (512, 512, 512)])
def test_gru_gates_bwd_rows(B, H, nrows):
    _dev()
    nr = torch.tensor([nrows], dtype=torch.int32)

    def mk():
        dgi, dgh, dhp = torch.full((B, 3 * H), 2.0), torch.full((B, 3 * H), 2.0), torch.zeros(B, H)
        rzn = torch.rand(B, 3 * H) * 0.9 + 0.05
        return ([torch.randn(B, H), H, torch.randn(B, H), H, torch.randn(B, H), H, rzn, 3 * H, torch.randn(B, H), H,
                 torch.randn(B, H), H, dgi, 3 * H, dgh, 3 * H, dhp, H, B, H, nr, None], [dgi, dgh, dhp])
    for g, c in _both("pd_gru_gates_bwd_rows", mk):
        assert torch.allclose(g, c, atol=1e-5, rtol=1e-5), float((g - c).abs().max())


def test_dur_decode_rows_fwd_bwd():
    _dev()
    R, slots = 512, 15
    Q = R * slots
    tab = _table(R, [512, 300, 130, 17, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    torch.manual_seed(5)
    par = [torch.randn(192, 5) * 0.3, torch.randn(192) * 0.1, torch.randn(192, 64) * 0.2, torch.randn(192) * 0.1, torch.rand(5),
           torch.randn(2, 64) * 0.3, torch.randn(2) * 0.1]
    h0 = torch.randn(Q, 64)

    def mk():
        lg, S = torch.full((Q, 5, 2), 9.0), torch.full((Q, 6, 72), 9.0)
        return ([h0, 64, Q] + par + [lg, S, tab[17:], R, None], [lg, S])
    (gl, cl), (gS, cS) = _both("pd_dur_decode_fwd_rows", mk)
    live = (torch.arange(Q) % R) < tab[17:][(torch.arange(Q) // R).long()]
    # greedy feedback inside: near-tied logits may pick the other bit on the two sides -- compare rows whose bits agree
    same = ((gl[..., 1] > gl[..., 0]) == (cl[..., 1] > cl[..., 0])).all(1)
    assert float(same[live].float().mean()) > 0.995
    assert torch.allclose(gl[same], cl[same], atol=3e-3) and torch.allclose(gS[same], cS[same], atol=3e-3)
    assert bool((gl[~live] == 9.0).all()) and bool((gS[~live] == 9.0).all())
    S = cS.clone()
    dlog = torch.randn(Q, 5, 2) * 0.1

    def mk2():
        GX, dh0 = torch.full((Q, 6, 264), 4.0), torch.full((Q, 64), 4.0)
        return ([S, dlog, Q] + par + [GX, dh0, 64, tab[17:], R, None], [GX, dh0])
    for g, c in _both("pd_dur_decode_bwd_rows", mk2):
        assert torch.allclose(g, c, atol=2e-3, rtol=1e-2), float((g - c).abs().max())
        assert bool((g[~live] == 4.0).all())


def test_row_reductions_gather_and_embed_bwd():
    _dev()
    R, T, C = 1024, 15, 1536
    tab = _table(R, [1024, 1024, 700, 300, 129, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    x = torch.randn(T, R, C)
    live = torch.arange(R)[None, :] < tab[18:18 + T][:, None]
    x[~live] = float("nan")

    def mk():
        o = torch.zeros(R, C)
        return ([x, R * C, C, T, tab[18:], o, C, R, C, None], [o])
    for g, c in _both("pd_sum_slots_rows_f32", mk):
        assert torch.allclose(g, c, atol=1e-4, rtol=1e-5)

    def mk2():
        o = torch.zeros(196)
        y = torch.randn(T * R, 196)
        y[~live.reshape(-1)] = float("nan")
        return ([y, 196, T * R, 194, o, 0, tab[18:], R, None], [o])
    for g, c in _both("pd_colsum_rows_f32", mk2):
        assert torch.allclose(g[:194], c[:194], atol=2e-2, rtol=1e-4) and bool(torch.isfinite(g[:194]).all())
    idx = torch.randperm(R).to(torch.int32)

    def mk3():
        o = torch.zeros(R, 256)
        return ([torch.randn(R, 256), 256, idx, R, 256, o, 256, None], [o])
    for g, c in _both("pd_gather_rows_f32", mk3):
        assert torch.equal(g, c)
    rng = np.random.RandomState(1)
    tok = torch.from_numpy(np.concatenate([rng.randint(0, 131, (16 * R, 1)), rng.randint(0, 2, (16 * R, 5))], 1).astype(np.int32))
    tab0 = _table(R, [1024, 1024, 700, 300, 129, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    live16 = (torch.arange(R)[None, :] < tab0[17:33][:, None]).reshape(-1)

    def mk4():
        g_ = torch.randn(16 * R, 128)
        g_[~live16] = float("nan")
        dwt, db = torch.zeros(135, 128), torch.zeros(128)
        return ([tok, 16 * R, g_, 128, dwt, db, tab0[17:], R, None], [dwt, db])
    for g, c in _both("pd_note_embed_bwd_rows", mk4):
        assert bool(torch.isfinite(g).all()) and torch.allclose(g, c, atol=2e-3, rtol=1e-4), float((g - c).abs().max())


@pytest.mark.parametrize("reverse,final", [(0, False), (1, False), (0, True), (1, True)])
def test_gru128_bwd_rows_sorted_lengths(reverse, final):
    """Resident summary-GRU backward over length-sorted rows with a slot-major dgi slab: masked (row, step) entries of dgi
    are zero-filled only inside the padded live prefix of each step, everything else keeps its contents."""
    _dev()
    from tests.cpu_backend import CpuBackend
    cpu = CpuBackend()
    R, T, H = 512, 16, 128
    torch.manual_seed(4)
    w, b = torch.randn(3 * H, H) / np.sqrt(H), torch.randn(3 * H) * 0.1
    rng = np.random.RandomState(3)
    lengths = torch.from_numpy(np.sort(np.where(rng.rand(R) < 0.4, 2, rng.randint(3, 12, R)))[::-1].astype(np.int32).copy())
    tab = _table(R, [int((lengths > t).sum()) for t in range(17)])
    gi = torch.randn(R, T, 3 * H)
    h_all, rzn, hn = torch.zeros(R, T, H), torch.zeros(R, T, 3 * H), torch.zeros(R, T, H)
    cpu.pd_gru128_fwd(gi.data_ptr(), T * 3 * H, 3 * H, lengths.data_ptr(), w.data_ptr(), b.data_ptr(), h_all.data_ptr(), T * H, H,
                      rzn.data_ptr(), T * 3 * H, 3 * H, hn.data_ptr(), T * H, H, R, T, reverse, 1, None)

    def mkb():
        slab = torch.full((T, R, 3 * H), 6.0)                        # slot-major storage, (R,T,.) strides (3H, R*3H)
        dgh = torch.zeros(R, T, 3 * H)
        # final: only the summary (the state after the last processed step) carried gradient -> an (R,H) dout
        dout = torch.randn(R, H) if final else torch.randn(R, T, H)
        step = (0 if reverse else T - 1) if final else -1
        return ([dout, H if final else T * H, 0 if final else H, h_all, T * H, H, rzn, T * 3 * H, 3 * H, hn, T * H, H, lengths,
                 w, slab, 3 * H, R * 3 * H, dgh, T * 3 * H, 3 * H, R, T, reverse, tab[17:], step, None], [slab, dgh])
    (gs, cs), (gh, ch) = _both("pd_gru128_bwd_rows", mkb)
    assert torch.allclose(gs, cs, atol=5e-3, rtol=1e-3) and torch.allclose(gh, ch, atol=5e-3, rtol=1e-3)
    assert bool((cs == 6.0).any())
