"""-m gpu: the SURVEY.md 8 rows that round 1 only covered on the CPU emulation, on the CUDA library:
a14 weighted duration loss, a20 posterior_sample / prior_sample / interp, f1 on-device augmentation + batch
construction, f3 PtvaeEncoder, f4 device slerp -- against vectors written by the unmodified reference
(tests/golden/aux.npz, made by tests/golden/make_golden_aux.py) and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict, make_ptvae_encoder_state, PTVAE_ENCODER_SPEC


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _model(dev, seed, gain=1.0, eos_bias=0.0):
    from polydis_b200.model import DisentangleVAE
    m = DisentangleVAE.init_model(device=dev)
    m.load_state_dict(make_state_dict(seed, gain=gain, eos_bias=eos_bias))
    return m.to(dev)


def test_augmentation_and_batch_construction_on_device(golden_dir):
    """f1: np.roll transposition + expand_chord + PianoTree grid, bit-exact against the reference's converter."""
    dev = _dev()
    from polydis_b200 import ops
    g = np.load(os.path.join(golden_dir, "aux.npz"))
    _, _, pr = synth_batch(int(g["B"]), int(g["data_seed"]))
    x, c36, pr_s, ovf = ops.augment_batch(torch.from_numpy(pr).to(dev), torch.from_numpy(g["chord14"]).to(dev),
                                          torch.from_numpy(g["shifts"]).to(dev))
    assert np.array_equal(pr_s.cpu().numpy(), g["pr_shift"])
    assert np.array_equal(c36.cpu().numpy(), g["c36"])
    assert np.array_equal(x.cpu().numpy(), g["grid"]) and int(ovf) == 0
    # the constructed batch feeds the model directly
    m = _model(dev, 4)
    losses = m('train', x, c36, pr_s, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5))
    assert all(torch.isfinite(v) for v in losses)


def test_slerp_and_interp_on_device(golden_dir):
    """f4 / a20: interp_path on the device against the reference's numpy; interp end points decode like the sources."""
    dev = _dev()
    from oracle import polydis_oracle as O
    from polydis_b200 import ops
    g = np.load(os.path.join(golden_dir, "aux.npz"))
    paths = ops.slerp_path(torch.from_numpy(g["z1"]).to(dev), torch.from_numpy(g["z2"]).to(dev), 7).cpu().numpy()
    assert np.allclose(paths, g["paths"], atol=2e-6, rtol=1e-5), np.abs(paths - g["paths"]).max()
    m = _model(dev, 4, 2.0, 0.75)
    sd = make_state_dict(4, gain=2.0, eos_bias=0.75)
    _, c, pr = (torch.from_numpy(a) for a in synth_batch(4, 9))
    ref = O.inference(sd, pr, c)
    prd, cd = pr.to(dev), c.to(dev)
    it = m.interp(prd[:2], cd[:2], prd[2:], cd[2:], interp_chd=True, interp_rhy=True, int_count=4)
    assert it.shape == (2, 4, 32, 15, 6) and it.dtype == np.int64
    assert (it[:, 0] == ref[:2]).mean() >= 0.999 and (it[:, -1] == ref[2:]).mean() >= 0.999
    # only one latent interpolated: the other stays at the first source
    it2 = m.interp(prd[:2], cd[:2], prd[2:], cd[2:], interp_chd=False, interp_rhy=True, int_count=3)
    assert (it2[:, 0] == ref[:2]).mean() >= 0.999


def test_sampling_wrappers_on_device():
    """a20: posterior_sample / prior_sample against the oracle's decode of the same latents."""
    dev = _dev()
    from oracle import polydis_oracle as O
    m = _model(dev, 4, 2.0, 0.75)
    sd = make_state_dict(4, gain=2.0, eos_bias=0.75)
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(3, 9))
    prd, cd = pr.to(dev), c.to(dev)
    ref = O.inference(sd, pr, c)
    a = m.posterior_sample(prd, cd, scale=None, sample_chd=False, sample_txt=False)
    assert (a == ref).mean() >= 0.999
    assert (m.posterior_sample(prd, cd, scale=0.0) == ref).mean() >= 0.999      # scale 0 collapses onto the mean
    z0 = m.prior_sample(prd, cd, sample_chd=True, sample_rhy=True, scale=0.0)    # z = 0 for every segment
    assert z0.shape == (3, 32, 15, 6) and np.array_equal(z0[0], z0[1]) and np.array_equal(z0[0], z0[2])
    # a real sample: decode of z = mu + std * eps must equal the oracle's decode of the same z
    torch.manual_seed(5)
    e1, e2 = torch.randn(3, 256), torch.randn(3, 256)
    got = m.inference(prd, cd, sample=True, eps=(e1.to(dev), e2.to(dev)))
    assert (got == O.inference(sd, pr, c, e1, e2)).mean() >= 0.999
    assert got.shape == (3, 32, 15, 6) and (got != ref).any()
    assert np.array_equal(m.gt_sample(x.to(dev)), x[:, :, 1:].numpy())


def test_weighted_duration_loss_on_device(golden_dir):
    """a14: recon_loss(weighted_dur=True) (ptvae.py:512-527) against the reference's value, forward and gradient flow."""
    dev = _dev()
    g = np.load(os.path.join(golden_dir, "aux.npz"))
    m = _model(dev, 4)
    xs = torch.from_numpy(synth_batch(int(g["B"]), int(g["data_seed"]))[0][:2]).to(dev)
    pitch = torch.from_numpy(g["wd_pitch"]).to(dev).requires_grad_(True)
    dur = torch.from_numpy(g["wd_dur"]).to(dev).requires_grad_(True)
    for flag, key in ((True, "wd_losses"), (False, "ud_losses")):
        got = m.decoder.recon_loss(xs, pitch, dur, (1, 0.5), flag)
        assert np.allclose([float(v) for v in got], g[key], rtol=1e-5), ([float(v) for v in got], g[key])
    got[0].backward()
    loss_w = m.decoder.recon_loss(xs, pitch, dur, (1, 0.5), True)[0]
    pitch.grad = dur.grad = None
    loss_w.backward()
    # per-bit weights [1, .6, .4, .3, .3]: the gradient of bit k scales with its weight; compare with torch autograd
    pr_, dr_ = torch.from_numpy(g["wd_pitch"]).requires_grad_(True), torch.from_numpy(g["wd_dur"]).requires_grad_(True)
    xc = xs.cpu()
    ce_p = torch.nn.CrossEntropyLoss(ignore_index=130)(pr_.view(-1, 130), xc[:, :, 1:, 0].reshape(-1))
    w = [1, 0.6, 0.4, 0.3, 0.3]
    ce_d = sum(w[k] * torch.nn.CrossEntropyLoss(ignore_index=2)(dr_[..., k, :].reshape(-1, 2), xc[:, :, 1:, 1 + k].reshape(-1))
               for k in range(5))
    (ce_p + 0.5 * ce_d).backward()
    assert torch.allclose(pitch.grad.cpu(), pr_.grad, atol=1e-7, rtol=1e-4)
    assert torch.allclose(dur.grad.cpu(), dr_.grad, atol=1e-7, rtol=1e-4)


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_ptvae_encoder_on_device(golden_dir, prec):
    """f3: PtvaeEncoder (packed note-level bi-GRU(256) + time-level bi-GRU(512)) against the reference's own module
    on seeded weights: posterior mean / std, lengths, and the gradient norm of every parameter."""
    dev = _dev()
    from polydis_b200 import ops
    from polydis_b200.ptvae import PtvaeEncoder
    g = np.load(os.path.join(golden_dir, "aux.npz"))
    enc = PtvaeEncoder(device=dev)
    enc.load_state_dict(make_ptvae_encoder_state(5, gain=1.5))
    enc.to(dev)
    x = torch.from_numpy(synth_batch(3, 808)[0]).to(dev)
    with ops.precision(prec):
        dist, emb, lengths = enc(x)
        (dist.mean.sum() + dist.scale.sum()).backward()
    tol = 2e-5 if prec == "fp32" else 3e-3
    assert np.allclose(dist.mean.detach().cpu().numpy(), g["enc_mu"], atol=tol)
    assert np.allclose(dist.scale.detach().cpu().numpy(), g["enc_std"], atol=tol, rtol=tol)
    assert np.array_equal(lengths.cpu().numpy(), g["enc_lens"])
    assert np.allclose(emb.detach().double().sum(-1).cpu().numpy(), g["enc_emb_sum"], atol=1e-4)
    params = dict(enc.named_parameters())
    for i, (name, _, _) in enumerate(PTVAE_ENCODER_SPEC):
        gn = float(params[name].grad.double().norm())
        assert abs(gn - g["enc_grad_norm"][i]) <= (1e-3 if prec == "fp32" else 1e-2) * g["enc_grad_norm"][i] + 1e-9, name
