"""Host-side logic of the drop-in model on the numpy C-ABI emulation (tests/cpu_backend.py) against
the golden fixtures written by the unmodified reference.  No GPU, no CUDA kernels: this pins the
orchestration (hoisted projections, batched teacher-forced phases, BPTT, strides, random-draw order).
"""
import os
import random

import numpy as np
import pytest
import torch

from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict, STATE_DICT_SPEC
from tests import cpu_backend
from tests.golden.make_golden_probe import probe_indices


def _model(seed, gain=1.0, eos_bias=0.0):
    from polydis_b200.model import DisentangleVAE
    m = DisentangleVAE.init_model(device=torch.device("cpu"))
    m.load_state_dict(make_state_dict(seed, gain=gain, eos_bias=eos_bias))
    return m


def test_state_dict_keys_match_reference_contract():
    from polydis_b200.model import DisentangleVAE
    m = DisentangleVAE.init_model(device=torch.device("cpu"))
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in STATE_DICT_SPEC]
    assert all(tuple(sd[k].shape) == s for k, s, _ in STATE_DICT_SPEC)


@pytest.mark.parametrize("tag", ["tf111", "tf000", "tf555", "tf111-chunked", "tf555-devplan", "tf111-devplan",
                                 "tf000-batched", "tf555-batched", "tf111-fusedstep", "tf111-deferall", "tf555-deferall",
                                 "tf111-nodefer"])
def test_training_matches_reference_golden(golden_dir, monkeypatch, tag):
    be = cpu_backend.install(monkeypatch)
    if tag.endswith("-deferall") or tag.endswith("-nodefer"):
        # weight gradients as deferred jobs behind tag nodes (ops.defer) at EVERY site incl. the batch-sized layers, or
        # nowhere (the in-line computation): same gradients either way
        from polydis_b200 import ops
        monkeypatch.setattr(ops, "DEFER_MIN_ROWS", 1)
        monkeypatch.setattr(ops, "DEFER_WGRAD", tag.endswith("-deferall"))
        tag = tag[:tag.rindex("-")]
    fused = tag.endswith("-fusedstep")
    if fused:       # fused recurrent step kernels for every recurrence + the note GRU's x-projection folded into the step
        from polydis_b200 import ops
        monkeypatch.setattr(ops, "FUSED_GRU_STEP_TMA_MIN_ROWS", 1)
        monkeypatch.setattr(ops, "BF16_RECURRENT", False)      # (exact-logic check; the bf16 operand path has its own test)
        tag = tag[:-len("-fusedstep")]
    if tag.endswith("-chunked"):        # row-chunked recurrences (ops._over_row_chunks): 128-row chunks, ragged tail
        from polydis_b200 import ops
        monkeypatch.setattr(ops, "ROW_CHUNK_MIN_BYTES", 0)
        monkeypatch.setattr(ops, "ROW_CHUNK_BYTES", 1)
        monkeypatch.setattr(ops, "RESIDENT_GRU128", False)
        tag = tag[:-len("-chunked")]
    if tag.endswith("-batched"):        # scheduled sampling as greedy pass + batched teacher-forced phases on mixed inputs
        from polydis_b200.ptvae import PtvaeDecoder
        monkeypatch.setattr(PtvaeDecoder, "batched_sampling", True)
        tag = tag[:-len("-batched")]
    devplan = tag.endswith("-devplan")
    tag = tag[:-len("-devplan")] if devplan else tag
    g = np.load(os.path.join(golden_dir, f"train_{tag}.npz"))
    B = int(g["B"])
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(B, int(g["data_seed"])))
    m = _model(int(g["w_seed"]), float(g["gain"]), float(g["eos_bias"]))
    m.train()
    random.seed(int(g["rng_seed"]))
    eps = (torch.from_numpy(g["eps_chd"]), torch.from_numpy(g["eps_rhy"]))
    if True:
        if devplan:     # teacher-forcing decisions as device data (one CUDA graph for every ratio): same draws, same result
            plan = torch.tensor(m.draw_plan(*[float(v) for v in g["tfr"]]), dtype=torch.int32)
            out = m.run(x, c, pr, *[float(v) for v in g["tfr"]], eps=eps, plan_dev=plan)
        else:
            out = m.run(x, c, pr, *[float(v) for v in g["tfr"]], eps=eps)
        if fused:
            assert be.calls.count("pd_gru_step_tmax") == 15 and be.calls.count("pd_gru_step_tma") > 40
        losses = m.loss_function(x, c, *out, 0.1, (1, 0.5))
        np.testing.assert_allclose([float(v.detach()) for v in losses], g["losses"], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out[0].detach().numpy(), g["pitch"], atol=2e-5)
        np.testing.assert_allclose(out[1].detach().numpy(), g["dur"], atol=2e-5)
        np.testing.assert_allclose(out[3].scale.detach().numpy(), g["std_rhy"], atol=2e-5)
        np.testing.assert_allclose(out[5].detach().numpy(), g["chroma"], atol=2e-5)
    losses[0].backward()
    params = dict(m.named_parameters())
    for i, (name, _, _) in enumerate(STATE_DICT_SPEC):
        gr = params[name].grad.reshape(-1).double()
        assert abs(float(gr.norm()) - g["grad_norm"][i]) <= 1e-3 * g["grad_norm"][i] + 1e-9, name
        got = gr[torch.from_numpy(probe_indices(name, gr.numel()))].numpy()
        np.testing.assert_allclose(got, g["grad_probe"][i], rtol=5e-3,
                                   atol=1e-4 * g["grad_norm"][i] + 1e-10, err_msg=name)
    # python's random stream must be left exactly where the reference leaves it (487 draws)
    random.seed(int(g["rng_seed"]))
    for _ in range(487):
        random.random()
    expect = random.random()
    random.seed(int(g["rng_seed"]))
    m.run(x[:1], c[:1], pr[:1], *[float(v) for v in g["tfr"]], eps=(eps[0][:1], eps[1][:1]))
    assert random.random() == expect


def test_bf16_recurrent_operands_meet_the_gradient_tolerance(golden_dir, monkeypatch):
    """Batch-sized recurrences with bf16 copies of W_hh / h / dgh as GEMM operands (ops.BF16_RECURRENT: fp32 accumulate,
    gate math and state): host logic (ping-pong state copies, W_hh copy shared with the backward) and the numerical cost
    of the operand rounding alone -- the emulation is otherwise fp32-exact -- against the reference golden: losses well
    inside 1e-3, every gradient inside the 1e-2 tolerance (measured worst: 1e-3)."""
    be = cpu_backend.install(monkeypatch)
    from polydis_b200 import ops
    monkeypatch.setattr(ops, "FUSED_GRU_STEP_TMA_MIN_ROWS", 1)
    g = np.load(os.path.join(golden_dir, "train_tf111.npz"))
    B = int(g["B"])
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(B, int(g["data_seed"])))
    m = _model(int(g["w_seed"]), float(g["gain"]), float(g["eos_bias"]))
    m.train()
    random.seed(int(g["rng_seed"]))
    eps = (torch.from_numpy(g["eps_chd"]), torch.from_numpy(g["eps_rhy"]))
    losses = m.loss(x, c, pr, 1., 1., 1., eps=eps)
    losses[0].backward()
    # time GRU 31 + encoders 4 x 7 + chord decoder 8 fused steps (first steps without an initial state stay unfused)
    assert be.calls.count("pd_gru_step_tma_bf16_units") >= 60 and be.calls.count("pd_gru_gates_bwd_zb") >= 60
    assert be.calls.count("pd_gemm_bf16") == be.calls.count("pd_gru_gates_bwd_zb")
    np.testing.assert_allclose([float(v.detach()) for v in losses], g["losses"], rtol=3e-4, atol=1e-6)
    params = dict(m.named_parameters())
    worst = 0.0
    for i, (name, _, _) in enumerate(STATE_DICT_SPEC):
        gr = params[name].grad.reshape(-1).double()
        assert abs(float(gr.norm()) - g["grad_norm"][i]) <= 5e-3 * g["grad_norm"][i] + 1e-9, name
        got = gr[torch.from_numpy(probe_indices(name, gr.numel()))].numpy()
        err = np.linalg.norm(got - g["grad_probe"][i]) / (np.linalg.norm(g["grad_probe"][i]) + 1e-12)
        worst = max(worst, err)
        assert err <= 5e-3, (name, err)
    assert worst > 1e-5            # (the bf16 path really ran)


@pytest.mark.parametrize("B,defer,tfr", [(8, True, 1.), (4, False, 1.), (12, True, 1.), (8, True, 0.), (8, False, 0.5)])
def test_packed_loss_mode_equals_dense_path(monkeypatch, B, defer, tfr):
    """Loss mode through the packed note level (rows sorted by token count, slot-major buffers, dead note slots skipped)
    gives the losses and all 81 gradients of the dense path (which the test above pins to the reference goldens).  Every
    ``torch.empty`` is NaN-filled, so a dead row leaking into any result fails the comparison.  Batches that are not a
    multiple of 4 must fall back to the dense path."""
    be = cpu_backend.install(monkeypatch)
    from polydis_b200 import ops
    monkeypatch.setattr(ops, "DEFER_WGRAD", defer)
    monkeypatch.setattr(ops, "BF16_RECURRENT", False)          # exact comparison: no operand rounding on either side
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(B, 77))
    torch.manual_seed(3)
    eps = (torch.randn(B, 256), torch.randn(B, 256))
    res = {}
    for packed in (False, True):
        monkeypatch.setattr(ops, "PACKED_NOTES", packed)
        m = _model(1, 2.0, 3.0)
        m.train()
        random.seed(11)
        be.calls.clear()
        with monkeypatch.context() as mp:
            if packed:
                cpu_backend.poison_empty(mp)
            losses = m('train', x, c, pr, tfr1=tfr, tfr2=tfr, tfr3=tfr, beta=0.1, weights=(1, 0.5), eps=eps)
            losses[0].backward()
        # (tfr < 1: scheduled sampling as a greedy pass + the batched phases over mixed inputs, packed as well)
        assert (be.calls.count("pd_gru_step_tmax_rows") == 15) == packed
        res[packed] = ([float(v.detach()) for v in losses], {n: p.grad.clone() for n, p in m.named_parameters()})
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=1e-5, atol=1e-7)
    for n, g in res[False][1].items():
        gp = res[True][1][n]
        assert bool(torch.isfinite(gp).all()), n
        assert float((gp - g).abs().max()) <= 2e-5 * float(g.abs().max()) + 1e-9, n
    monkeypatch.setattr(ops, "PACKED_NOTES", True)
    m = _model(1)
    be.calls.clear()
    m('train', x[:3], c[:3], pr[:3], tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))
    assert "pd_pack_order" not in be.calls            # 96 rows: a 128-row tile would straddle note slots -> dense path


@pytest.mark.parametrize("tag", ["w0", "w1", "w1-3xtf32"])
def test_greedy_tokens_match_reference_golden(golden_dir, monkeypatch, tag):
    cpu_backend.install(monkeypatch)
    if tag.endswith("-3xtf32"):     # force the large-batch route: [hi|hi|lo].[hi|lo|hi] GEMMs, operands split once per
        from polydis_b200 import ops  # state (by the gate kernel), 3-pass duration decoder flag
        monkeypatch.setattr(ops, "TF32X3_MIN_ROWS", 1)
        tag = tag[:-len("-3xtf32")]
    g = np.load(os.path.join(golden_dir, f"infer_{tag}.npz"))
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(int(g["B"]), int(g["data_seed"])))
    m = _model(int(g["w_seed"]), float(g["gain"]), float(g["eos_bias"]))
    est = m.inference(pr, c, sample=False)
    assert est.dtype == np.int64 and est.shape == (int(g["B"]), 32, 15, 6)
    assert (est == g["est_x"]).mean() >= 0.999
    est2 = m.swap(pr, pr, c, c, True, True)
    assert np.array_equal(est, est2)


def test_forward_mode_dispatch(monkeypatch):
    cpu_backend.install(monkeypatch)
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(2, 5))
    m = _model(4)
    random.seed(0)
    a = m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))
    assert len(a) == 11 and all(v.dim() == 0 for v in a)
    assert len(m('run', x, c, pr, 1., 1., 1.)) == 7
    assert m(2, pr, c, False).shape == (2, 32, 15, 6)
    with pytest.raises(NotImplementedError):
        m('bogus')


def test_linear_split_gradients(monkeypatch):
    """ops.linear_split (several heads as one GEMM with a shared gradient slab) against plain torch autograd:
    slab-aware consumer (GRU recurrence writes its input gradient in place), ordinary consumer (gather copy) and
    an unused head (zero block); bias gradient restricted to the leading ``bias_cols`` columns."""
    cpu_backend.install(monkeypatch)
    from polydis_b200 import ops
    torch.manual_seed(3)
    R, T, K, H = 9, 4, 12, 8
    x = torch.randn(R, T, K, requires_grad=True)
    w = torch.randn(3 * H + 6 + 5, K, requires_grad=True)
    b = torch.randn(3 * H + 6 + 5, requires_grad=True)
    w_hh, b_hh = torch.randn(3 * H, H, requires_grad=True), torch.randn(3 * H, requires_grad=True)
    lengths = torch.tensor([4, 1, 3, 2, 4, 4, 1, 2, 3], dtype=torch.int32)

    def loss_ours():
        gi, mid, _unused = ops.linear_split(x, w, b, (3 * H, 6, 5), bias_cols=3 * H + 6)
        h = ops.gru_sequence(gi, None, None, w_hh, b_hh, lengths, False)
        return (h[:, -1] ** 2).sum() + (mid * torch.arange(6.0)).sin().sum()

    def loss_ref():
        y = x @ w.t() + b
        gi, mid = y[..., :3 * H], y[..., 3 * H:3 * H + 6]
        gru = torch.nn.GRU(K, H, batch_first=True)            # only its cell math is used
        h = torch.zeros(R, H)
        for t in range(T):
            gh = h @ w_hh.t() + b_hh
            r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
            hn = (1 - z) * n + z * h
            h = torch.where((t < lengths)[:, None], hn, h)
        del gru
        return (h ** 2).sum() + (mid * torch.arange(6.0)).sin().sum()

    grads = []
    for fn in (loss_ours, loss_ref):
        for p in (x, w, b, w_hh, b_hh):
            p.grad = None
        l = fn()
        l.backward()
        grads.append([float(l.detach())] + [p.grad.clone() for p in (x, w, b, w_hh, b_hh)])
    assert abs(grads[0][0] - grads[1][0]) < 1e-4 * abs(grads[1][0])
    for name, g, r in zip("x w b w_hh b_hh".split(), grads[0][1:], grads[1][1:]):
        if name == "b":                                        # columns past bias_cols carry no bias gradient by contract
            assert torch.equal(g[3 * H + 6:], torch.zeros(5))
            g, r = g[:3 * H + 6], r[:3 * H + 6]
        assert torch.allclose(g, r, atol=2e-4, rtol=1e-4), name
