"""Remaining public surface of the drop-in model on the numpy C-ABI
emulation: weighted duration loss, sampling wrappers, interpolation, checkpoint load, output formatting."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import polydis_oracle as O
from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict
from tests import cpu_backend


def _model(seed=4, **kw):
    from polydis_b200.model import DisentangleVAE
    m = DisentangleVAE.init_model(device=torch.device("cpu"))
    m.load_state_dict(make_state_dict(seed, **kw))
    return m


def test_weighted_duration_loss_matches_torch(monkeypatch):
    cpu_backend.install(monkeypatch)
    m = _model()
    x = torch.from_numpy(synth_batch(2, 8)[0])
    torch.manual_seed(0)
    pitch = torch.randn(2, 32, 15, 130, requires_grad=True)
    dur = torch.randn(2, 32, 15, 5, 2, requires_grad=True)
    loss, pl, dl = m.decoder.recon_loss(x, pitch, dur, weights=(1, 0.5), weighted_dur=True)
    ce = torch.nn.functional.cross_entropy
    w = [1, 0.6, 0.4, 0.3, 0.3]
    ref_p = ce(pitch.detach().reshape(-1, 130), x[:, :, 1:, 0].reshape(-1), ignore_index=130)
    ref_d = sum(w[k] * ce(dur.detach().reshape(-1, 5, 2)[:, k], x[:, :, 1:, 1:].reshape(-1, 5)[:, k], ignore_index=2)
                for k in range(5))
    assert abs(float(pl) - float(ref_p)) < 1e-5 and abs(float(dl) - float(ref_d)) < 1e-5
    assert abs(float(loss) - float(ref_p + 0.5 * ref_d)) < 1e-5
    loss.backward()
    assert pitch.grad.abs().sum() > 0 and dur.grad.abs().sum() > 0


def test_sampling_wrappers_and_interp(monkeypatch):
    cpu_backend.install(monkeypatch)
    m = _model(gain=2.0, eos_bias=0.75)
    sd = make_state_dict(4, gain=2.0, eos_bias=0.75)
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(3, 9))
    # posterior_sample with frozen chord + texture latents == mean decoding == swap
    a = m.posterior_sample(pr, c, scale=None, sample_chd=False, sample_txt=False)
    assert np.array_equal(a, m.swap(pr, pr, c, c, True, True))
    assert np.array_equal(a, O.inference(sd, pr, c))
    # scale = 0 collapses the posterior onto its mean
    assert np.array_equal(m.posterior_sample(pr, c, scale=0.0), a)
    # prior_sample always rsamples (model.py:174-184): from N(0, scale) for the replaced latents, from the
    # posterior otherwise -- with both latents replaced and scale 0, z = 0 for every segment
    z0 = m.prior_sample(pr, c, sample_chd=True, sample_rhy=True, scale=0.0)
    assert z0.shape == (3, 32, 15, 6) and np.array_equal(z0[0], z0[1]) and np.array_equal(z0[0], z0[2])
    torch.manual_seed(1)
    s1 = m.inference(pr, c, sample=True)
    assert s1.shape == (3, 32, 15, 6)
    # interpolation end points decode like the two sources (distinct pairs: the slerp of the reference
    # divides by sin(0) for identical end points, model.py:222-229)
    pr1, c1, pr2, c2 = pr[:2], c[:2], pr[:2].flip(0), c[:2].flip(0)
    it = m.interp(pr1, c1, pr2, c2, interp_chd=True, interp_rhy=True, int_count=3)
    assert it.shape == (2, 3, 32, 15, 6)
    assert (it[:, 0] == a[:2]).mean() > 0.99
    assert (it[:, -1] == a[:2][::-1]).mean() > 0.99
    assert np.array_equal(m.gt_sample(x), x[:, :, 1:].numpy())


def test_load_model_strips_dataparallel_prefix(monkeypatch, tmp_path):
    cpu_backend.install(monkeypatch)
    sd = make_state_dict(6)
    path = os.path.join(tmp_path, "ckpt.pt")
    torch.save({"module." + k: v for k, v in sd.items()}, path)
    m = _model(seed=1)
    m.load_model(path, map_location="cpu")
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())


def test_output_to_numpy_and_grid_to_pr(monkeypatch):
    cpu_backend.install(monkeypatch)
    m = _model()
    x, c, pr = synth_batch(1, 10)
    torch.manual_seed(0)
    p, d = torch.randn(1, 32, 15, 130), torch.randn(1, 32, 15, 5, 2)
    est, pn, dn = m.decoder.output_to_numpy(p, d)
    assert est.shape == (1, 32, 15, 6) and np.array_equal(est[..., 0], p.argmax(-1).numpy())
    assert np.array_equal(est[..., 1:], d.argmax(-1).numpy()) and pn.shape == (1, 32, 15, 130)
    # ground-truth grid -> piano-roll round trip (first 10 notes of a step, durations clipped at the bar end)
    roll, notes = m.decoder.grid_to_pr_and_notes(x[0], bpm=60., start=0.)
    assert roll.shape == (32, 128)
    t, pch = np.nonzero(pr[0])
    assert all(roll[a, b] == min(int(pr[0][a, b]), 32 - a) for a, b in zip(t, pch))
    assert len(notes) == int((roll > 0).sum())


def test_device_batch_construction_and_roundtrip(monkeypatch):
    """pr_mat -> grid equals the reference converter's grid; tokens -> pr_mat (8f-4) inverts it
    for steps with at most 10 notes."""
    cpu_backend.install(monkeypatch)
    from polydis_b200 import ops
    x, c, pr = synth_batch(3, 12)
    gx, ovf = ops.pr_mat_to_grid(torch.from_numpy(pr))
    assert int(ovf) == 0 and np.array_equal(gx.numpy(), x)
    back = ops.tokens_to_pr_mat(torch.from_numpy(x[:, :, 1:, :].astype(np.int32)))
    assert np.array_equal(back.numpy(), pr)
    crowded = np.zeros((1, 32, 128), np.float32)
    crowded[0, 2, 30:46] = 3
    _, ovf = ops.pr_mat_to_grid(torch.from_numpy(crowded))
    assert int(ovf) == 1


def test_ptvae_encoder_matches_torch_restatement(monkeypatch):
    """PtvaeEncoder against nn.GRU + pack_padded_sequence on the same weights, fwd + bwd."""
    cpu_backend.install(monkeypatch)
    from torch.nn.utils.rnn import pack_padded_sequence
    from polydis_b200.ptvae import PtvaeEncoder
    torch.manual_seed(0)
    enc = PtvaeEncoder(device="cpu")
    x = torch.from_numpy(synth_batch(2, 14)[0])
    dist, emb, lengths = enc(x)
    (dist.mean.sum() + dist.scale.sum()).backward()

    sd = {k: v.detach().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    g1 = torch.nn.GRU(128, 256, batch_first=True, bidirectional=True)
    g2 = torch.nn.GRU(512, 512, batch_first=True, bidirectional=True)
    g1.load_state_dict({k.split(".", 1)[1]: v for k, v in sd.items() if k.startswith("enc_notes_gru.")})
    g2.load_state_dict({k.split(".", 1)[1]: v for k, v in sd.items() if k.startswith("enc_time_gru.")})
    mh = O.grid_multihot(x)
    e = torch.nn.functional.linear(mh, sd["note_embedding.weight"], sd["note_embedding.bias"])
    ln = O.grid_lengths(x)
    pk = pack_padded_sequence(e.view(-1, 16, 128), ln.view(-1), batch_first=True, enforce_sorted=False)
    n = g1(pk)[-1].transpose(0, 1).reshape(2, 32, 512)
    h = g2(n)[-1].transpose(0, 1).reshape(2, 1024)
    mu = torch.nn.functional.linear(h, sd["linear_mu.weight"], sd["linear_mu.bias"])
    sdv = torch.nn.functional.linear(h, sd["linear_std.weight"], sd["linear_std.bias"]).exp()
    assert torch.allclose(dist.mean, mu, atol=2e-5) and torch.allclose(dist.scale, sdv, atol=2e-5)
    assert torch.equal(lengths, ln) and torch.allclose(emb, e, atol=1e-6)
    (mu.sum() + sdv.sum()).backward()
    ref_g = {"enc_notes_gru." + k: p.grad for k, p in g1.named_parameters()}
    ref_g.update({"enc_time_gru." + k: p.grad for k, p in g2.named_parameters()})
    for name, p in enc.named_parameters():
        ref = ref_g[name] if name in ref_g else sd[name].grad
        assert torch.allclose(p.grad, ref, atol=2e-4, rtol=1e-3), name


def test_aux_rows_match_reference_golden(monkeypatch, golden_dir):
    """Rows either side of the path against vectors written by the unmodified reference
    (tests/golden/make_golden_aux.py): on-device augmentation (np.roll + expand_chord + grid), interp_path slerp,
    weighted duration loss -- here through the numpy emulation of the C ABI (the -m gpu twin runs the kernels)."""
    import os
    cpu_backend.install(monkeypatch)
    from polydis_b200 import ops
    g = np.load(os.path.join(golden_dir, "aux.npz"))
    _, _, pr = synth_batch(int(g["B"]), int(g["data_seed"]))
    x, c36, pr_s, ovf = ops.augment_batch(torch.from_numpy(pr), torch.from_numpy(g["chord14"]), torch.from_numpy(g["shifts"]))
    assert np.array_equal(pr_s.numpy(), g["pr_shift"]) and np.array_equal(c36.numpy(), g["c36"])
    assert np.array_equal(x.numpy(), g["grid"]) and int(ovf) == 0
    paths = ops.slerp_path(torch.from_numpy(g["z1"]), torch.from_numpy(g["z2"]), 7).numpy()
    assert np.allclose(paths, g["paths"], atol=2e-6, rtol=1e-5)
    m = _model()
    xs = torch.from_numpy(synth_batch(int(g["B"]), int(g["data_seed"]))[0][:2])
    for flag, key in ((True, "wd_losses"), (False, "ud_losses")):
        got = m.decoder.recon_loss(xs, torch.from_numpy(g["wd_pitch"]), torch.from_numpy(g["wd_dur"]), (1, 0.5), flag)
        assert np.allclose([float(v) for v in got], g[key], rtol=2e-6)


def test_ptvae_encoder_matches_reference_golden(monkeypatch, golden_dir):
    """PtvaeEncoder against the reference's own module (fixture from tests/golden/make_golden_aux.py)."""
    import os
    cpu_backend.install(monkeypatch)
    from polydis_b200 import ops
    from polydis_b200.ptvae import PtvaeEncoder
    from polydis_b200.weights import make_ptvae_encoder_state, PTVAE_ENCODER_SPEC
    g = np.load(os.path.join(golden_dir, "aux.npz"))
    enc = PtvaeEncoder(device="cpu")
    enc.load_state_dict(make_ptvae_encoder_state(5, gain=1.5))
    with ops.precision("fp32"):
        dist, emb, lengths = enc(torch.from_numpy(synth_batch(3, 808)[0]))
        (dist.mean.sum() + dist.scale.sum()).backward()
    assert np.allclose(dist.mean.detach().numpy(), g["enc_mu"], atol=2e-5)
    assert np.allclose(dist.scale.detach().numpy(), g["enc_std"], atol=2e-5, rtol=2e-5)
    assert np.array_equal(lengths.numpy(), g["enc_lens"])
    params = dict(enc.named_parameters())
    for i, (name, _, _) in enumerate(PTVAE_ENCODER_SPEC):
        gn = float(params[name].grad.double().norm())
        assert abs(gn - g["enc_grad_norm"][i]) <= 1e-3 * g["enc_grad_norm"][i] + 1e-9, name


def test_vectorised_synth_has_the_recipe_distribution():
    """``synth_batch`` (vectorised, PCG64) draws from the distribution of the BASELINE / SURVEY recipe
    (``synth_batch_recipe``: RandomState loop): active-step rate, notes per step, pitch range, duration law, chord stats."""
    import numpy as np
    from polydis_b200.synth import synth_batch, synth_batch_recipe
    xa, ca, pa = synth_batch(256, 11)
    xb, cb, pb = synth_batch_recipe(256, 11)
    assert xa.shape == xb.shape and ca.shape == cb.shape and pa.shape == pb.shape and xb.dtype == np.int64
    for pr in (pa, pb):
        assert pr[:, :, :36].sum() == 0 and pr[:, :, 96:].sum() == 0
        assert (pr.max(-1) <= (32 - np.arange(32))[None, :]).all()
    na, nb = (pa != 0).sum(-1), (pb != 0).sum(-1)
    assert abs((na > 0).mean() - (nb > 0).mean()) < 0.02 and abs((na > 0).mean() - 0.6) < 0.02
    assert abs(na[na > 0].mean() - nb[nb > 0].mean()) < 0.12 and na.max() <= 8 and nb.max() <= 8
    assert abs(pa[pa != 0].mean() - pb[pb != 0].mean()) < 0.25            # mean duration
    assert abs(ca[:, :, 12:24].mean() - cb[:, :, 12:24].mean()) < 0.02
    assert (ca[:, :, :12].sum(-1) == 1).all() and (cb[:, :, 24:].sum(-1) == 1).all()
    # the grid builder is shared: token counts per step follow from the note counts (SOS + notes + EOS)
    assert ((xb[..., 0] != 130).sum(-1) == nb + 2).all()
