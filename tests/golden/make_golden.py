"""Generate the golden fixtures by running the UNMODIFIED reference (CPU) -- build-container only.

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden.py

Imports /root/reference (read-only) with stub ``pretty_midi`` / ``tensorboardX`` modules (neither is
touched by the hot path; both are missing offline), loads seeded weights
(``polydis_b200.weights.make_state_dict``) into ``DisentangleVAE.init_model(cpu)`` and records, for
seeded synthetic inputs (``polydis_b200.synth.synth_batch``):

  train_*.npz   the 11 losses, full pitch/dur logits, chord logits, mu/std of both posteriors and,
                for every one of the 81 parameters, gradient L2 norm + sum + 48 probed entries
  infer_*.npz   greedy ``est_x`` tokens from ``model.inference`` / ``swap``-style mean decoding
  grid.npz      the reference converter's PianoTree grid for the synthetic piano-rolls

/root/reference does not exist on the GPU box, so nothing at test time imports it; the fixtures and
this script are what travels.
"""
import os
import random
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

from converter import target_to_3dtarget            # reference
from model import DisentangleVAE                     # reference

from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict, STATE_DICT_SPEC

from tests.golden.make_golden_probe import probe_indices


def ref_model(seed, gain=1.0, eos_bias=0.0):
    m = DisentangleVAE.init_model(device=torch.device("cpu"))
    m.load_state_dict(make_state_dict(seed, gain=gain, eos_bias=eos_bias))
    return m


def draw_eps(B, seed):
    torch.manual_seed(seed)
    return torch.empty(B, 256).normal_(), torch.empty(B, 256).normal_()


def make_train(tag, B, data_seed, w_seed, tfr, rng_seed, gain=1.0, eos_bias=0.0):
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(B, data_seed))
    m = ref_model(w_seed, gain, eos_bias)
    e1, e2 = draw_eps(B, rng_seed)
    torch.manual_seed(rng_seed)
    random.seed(rng_seed)
    m.train()
    out = m.run(x, c, pr, *tfr)
    losses = m.loss_function(x, c, *out, 0.1, (1, 0.5))
    losses[0].backward()
    rec = dict(B=B, data_seed=data_seed, w_seed=w_seed, tfr=np.array(tfr), rng_seed=rng_seed,
               gain=gain, eos_bias=eos_bias,
               eps_chd=e1.numpy(), eps_rhy=e2.numpy(),
               losses=np.array([float(v) for v in losses], dtype=np.float64),
               pitch=out[0].detach().numpy(), dur=out[1].detach().numpy(),
               mu_chd=out[2].mean.detach().numpy(), std_chd=out[2].scale.detach().numpy(),
               mu_rhy=out[3].mean.detach().numpy(), std_rhy=out[3].scale.detach().numpy(),
               root=out[4].detach().numpy(), chroma=out[5].detach().numpy(),
               bass=out[6].detach().numpy())
    params = dict(m.named_parameters())
    norms, sums, probes = [], [], []
    for name, _, _ in STATE_DICT_SPEC:
        g = params[name].grad.reshape(-1).double()
        norms.append(float(g.norm()))
        sums.append(float(g.sum()))
        probes.append(g[torch.from_numpy(probe_indices(name, g.numel()))].numpy())
    rec.update(grad_norm=np.array(norms), grad_sum=np.array(sums), grad_probe=np.stack(probes))
    np.savez_compressed(os.path.join(HERE, f"train_{tag}.npz"), **rec)
    print(tag, "losses", np.round(rec["losses"], 5))


def make_infer(tag, B, data_seed, w_seed, gain, eos_bias):
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(B, data_seed))
    m = ref_model(w_seed, gain, eos_bias)
    random.seed(0)
    est = m.inference(pr, c, sample=False)
    np.savez_compressed(os.path.join(HERE, f"infer_{tag}.npz"), B=B, data_seed=data_seed,
                        w_seed=w_seed, gain=gain, eos_bias=eos_bias, est_x=est.astype(np.int16))
    print(tag, "eos frac", float((est[..., 0] == 129).mean()), "distinct pitches",
          len(np.unique(est[..., 0])))


def make_grid():
    x, c, pr = synth_batch(6, 11)
    ref = np.stack([target_to_3dtarget(p, max_note_count=16, max_pitch=128, min_pitch=0,
                                       pitch_pad_ind=130, pitch_sos_ind=128, pitch_eos_ind=129)
                    for p in pr])
    assert np.array_equal(ref, x), "vectorised grid builder disagrees with the reference converter"
    np.savez_compressed(os.path.join(HERE, "grid.npz"), pr_mat=pr.astype(np.int8), x=ref.astype(np.int16),
                        c=c.astype(np.int8))
    print("grid ok", ref.shape)


if __name__ == "__main__":
    torch.set_num_threads(8)
    make_grid()
    make_train("tf111", 4, 0, 0, (1.0, 1.0, 1.0), 5)
    make_train("tf000", 3, 1, 1, (0.0, 0.0, 0.0), 6)
    make_train("tf555", 3, 2, 2, (0.5, 0.5, 0.5), 7, gain=2.0, eos_bias=0.75)
    make_infer("w0", 8, 3, 3, 1.0, 0.0)          # default init: near-degenerate decodes
    make_infer("w1", 8, 3, 3, 2.0, 0.75)         # gain-2 weights: diverse tokens, varying lengths
