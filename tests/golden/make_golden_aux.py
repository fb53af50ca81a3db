"""Golden vectors for the rows either side of the hot path, written by the UNMODIFIED reference --
build-container only (same import recipe as make_golden.py):

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden_aux.py

  aux.npz  * augmentation: ``converter.augment_pr`` (np.roll along pitch) + ``converter.expand_chord(c, shift)`` for
             every shift in [-6, 5] (dataset.py:67-120) and the PianoTree grid of the shifted piano-roll
             (``converter.target_to_3dtarget`` as called at dataset.py:98-104);
           * ``DisentangleVAE.interp_path`` (model.py:218-242) for a few latent pairs;
           * ``PtvaeDecoder.recon_loss(weighted_dur=True)`` (ptvae.py:512-527) on seeded logits;
           * ``PtvaeEncoder`` (ptvae.py:125-215) posterior mean / std, lengths and per-parameter gradient norms.
"""
import os
import sys
import types

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
pm = types.ModuleType("pretty_midi"); pm.Note = lambda *a, **k: a; sys.modules["pretty_midi"] = pm
tb = types.ModuleType("tensorboardX"); tb.SummaryWriter = object; sys.modules["tensorboardX"] = tb

import numpy as np
import torch

from converter import augment_pr, expand_chord, target_to_3dtarget     # reference
from model import DisentangleVAE                                       # reference
from ptvae import PtvaeEncoder                                         # reference

from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_ptvae_encoder_state, PTVAE_ENCODER_SPEC


def main():
    rng = np.random.RandomState(2024)
    B = 12
    x, c, pr = synth_batch(B, 555)
    shifts = np.arange(-6, 6).astype(np.int32)                          # shift_low=-6, shift_high=5 (dataset.py)
    chord14 = np.zeros((B, 8, 14), np.float32)
    chord14[..., 0] = rng.randint(0, 12, (B, 8))
    chord14[..., 1:13] = (rng.rand(B, 8, 12) < 0.3)
    chord14[..., 13] = rng.randint(0, 12, (B, 8))
    pr_shift = np.stack([augment_pr(pr[b], int(shifts[b])) for b in range(B)]).astype(np.float32)
    c36 = np.stack([np.stack([expand_chord(chord14[b, s], int(shifts[b])) for s in range(8)]) for b in range(B)])
    grid = np.stack([target_to_3dtarget(pr_shift[b], max_note_count=16, max_pitch=128, min_pitch=0, pitch_pad_ind=130,
                                        pitch_sos_ind=128, pitch_eos_ind=129) for b in range(B)]).astype(np.int64)
    m = DisentangleVAE.init_model(device=torch.device("cpu"))
    z1 = rng.randn(5, 256).astype(np.float32)
    z2 = rng.randn(5, 256).astype(np.float32)
    z2[4] = 3.0 * z1[4] + 0.01 * z2[4]                                  # nearly parallel pair
    paths = np.stack([m.interp_path(a, b, 7).numpy() for a, b in zip(z1, z2)])
    # weighted duration loss
    torch.manual_seed(3)
    xs = torch.from_numpy(x[:2])
    pitch = torch.randn(2, 32, 15, 130)
    dur = torch.randn(2, 32, 15, 5, 2)
    wl = [float(v) for v in m.decoder.recon_loss(xs, pitch, dur, (1, 0.5), True)]
    ul = [float(v) for v in m.decoder.recon_loss(xs, pitch, dur, (1, 0.5), False)]
    # PtvaeEncoder (ptvae.py:125-215): posterior + gradient norms of mu.sum() + std.sum()
    enc = PtvaeEncoder(device=torch.device("cpu"))
    enc.load_state_dict(make_ptvae_encoder_state(5, gain=1.5))
    xe = torch.from_numpy(synth_batch(3, 808)[0])
    dist, emb, lens = enc(xe)
    (dist.mean.sum() + dist.scale.sum()).backward()
    pe = dict(enc.named_parameters())
    enc_gn = np.array([float(pe[n].grad.double().norm()) for n, _, _ in PTVAE_ENCODER_SPEC])
    enc_gs = np.array([float(pe[n].grad.double().sum()) for n, _, _ in PTVAE_ENCODER_SPEC])
    np.savez_compressed(os.path.join(HERE, "aux.npz"), enc_mu=dist.mean.detach().numpy(),
                        enc_std=dist.scale.detach().numpy(), enc_lens=lens.numpy(), enc_grad_norm=enc_gn,
                        enc_grad_sum=enc_gs, enc_emb_sum=emb.detach().double().sum(-1).numpy(), data_seed=555, B=B, shifts=shifts, chord14=chord14,
                        pr_shift=pr_shift, c36=c36.astype(np.float32), grid=grid, z1=z1, z2=z2, paths=paths,
                        wd_pitch=pitch.numpy(), wd_dur=dur.numpy(), wd_losses=np.array(wl), ud_losses=np.array(ul))
    print("wrote aux.npz", wl, ul)


if __name__ == "__main__":
    main()
