"""CPU oracle for the PolyDis hot path -- TEST INFRASTRUCTURE, not product code.

A plain-PyTorch fp32 restatement of the reference's training forward / loss and greedy inference,
written functionally over a state dict.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module; the product package
never does (it fails loudly without its CUDA library instead).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference from
/root/reference (CPU, torch 2.11) on seeded inputs + weights and stores its losses, logits,
gradient probes and greedy tokens in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this file against those vectors.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so reference-run fixtures are the pin.

The restatement keeps the reference's op granularity (one ``aten::gru`` call per recurrent step,
dense multi-hot embedding, per-step Python loops) so that timing it on host cores is a fair
stand-in for the reference's CPU path, and keeps every numerics-relevant quirk:
  * TextureEncoder ``.view(bs, 8, -1)`` memory reinterpretation          ptvae.py:114
  * ``Normal(mu, exp(linear_var))`` -- exp output used as the std         ptvae.py:27-28,120-121
  * KL averaged over all B*256 elements                                   train_utils.py:45-49
  * chord-decoder argmax feedback is the batch UNION of one-hots          ptvae.py:73-77
  * duration feedback token has its 1 at index == bit value               ptvae.py:322-326
  * dur_hid_linear consumes raw pitch logits                              ptvae.py:349-352
  * predicted lengths exclude EOS; GT lengths include it                  ptvae.py:415-416,292-297
  * PAD notes embed as bias + 2*sum(W[:,130:135])                          ptvae.py:309-312
  * python ``random`` is consumed 14x per time step for tfr2, once per time step (but the last)
    for tfr1, and 8x for tfr3                                             ptvae.py:420,476,81
"""
import random as _random
from collections import namedtuple

import torch
from torch.nn.utils.rnn import pack_padded_sequence

TeacherPlan = namedtuple("TeacherPlan", ["tf_note", "tf_time", "tf_chd"])

N_STEP, N_NOTE, N_DUR = 32, 16, 5
P_SOS, P_EOS, P_PAD, D_PAD, P_RANGE = 128, 129, 130, 2, 130


def draw_plan(tfr1, tfr2, tfr3, training=True, rng=_random):
    """Consume python's ``random`` exactly as one reference forward does and return the decisions.

    Order (ptvae.py:460-486 with :395-424 nested, then :58-83): for each of 32 time steps, 14 draws
    against tfr2 (note slots 1..14; slot 15 breaks before drawing), then -- except after the last
    step -- one draw against tfr1; finally (training only) 8 draws against tfr3.
    """
    tf_note, tf_time, tf_chd = [], [], []
    for t in range(N_STEP):
        tf_note.append([rng.random() < tfr2 for _ in range(N_NOTE - 2)])
        if t < N_STEP - 1:
            tf_time.append(rng.random() < tfr1)
    if training:
        tf_chd = [rng.random() < tfr3 for _ in range(N_STEP // 4)]
    return TeacherPlan(tf_note, tf_time, tf_chd)


def _lin(sd, name, x):
    return torch.nn.functional.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _gru_w(sd, name, reverse=False):
    s = "_reverse" if reverse else ""
    return [sd[f"{name}.weight_ih_l0{s}"], sd[f"{name}.weight_hh_l0{s}"],
            sd[f"{name}.bias_ih_l0{s}"], sd[f"{name}.bias_hh_l0{s}"]]


def _gru_step(sd, name, x, h):
    """One batch-first uni-directional aten::gru call on a length-1 sequence. x (B,1,I), h (1,B,H)."""
    return torch._VF.gru(x, h, _gru_w(sd, name), True, 1, 0.0, False, False, True)


def _bigru_last(sd, name, x):
    """Bi-GRU over a padded batch-first sequence; returns final hiddens (B, 2H) [fwd | bwd]."""
    w = _gru_w(sd, name) + _gru_w(sd, name, True)
    h0 = x.new_zeros(2, x.size(0), w[1].size(1))
    hn = torch._VF.gru(x, h0, w, True, 1, 0.0, False, True, True)[1]
    return hn.transpose(0, 1).reshape(x.size(0), -1)


def _bigru_last_packed(sd, name, x, lengths):
    """Bi-GRU over variable-length sequences (pack_padded_sequence, unsorted); final hiddens (B,2H)."""
    w = _gru_w(sd, name) + _gru_w(sd, name, True)
    pk = pack_padded_sequence(x, lengths.reshape(-1).cpu(), batch_first=True, enforce_sorted=False)
    h0 = x.new_zeros(2, x.size(0), w[1].size(1))
    hn = torch._VF.gru(pk.data, pk.batch_sizes, h0.index_select(1, pk.sorted_indices), w,
                       True, 1, 0.0, False, True)[1]
    hn = hn.index_select(1, pk.unsorted_indices)
    return hn.transpose(0, 1).reshape(x.size(0), -1)


# ----------------------------------------------------------------------------------------------
# encoders                                                                    ptvae.py:11-29,90-122
def chord_encoder(sd, c):
    h = _bigru_last(sd, "chd_encoder.gru", c)
    return _lin(sd, "chd_encoder.linear_mu", h), _lin(sd, "chd_encoder.linear_var", h).exp()


def texture_encoder(sd, pr_mat):
    bs = pr_mat.size(0)
    y = torch.nn.functional.conv2d(pr_mat.unsqueeze(1), sd["rhy_encoder.cnn.0.weight"],
                                   sd["rhy_encoder.cnn.0.bias"], stride=(4, 1))
    y = torch.nn.functional.max_pool2d(torch.relu(y), (1, 4), (1, 4))
    y = y.contiguous().view(bs, 8, -1)              # reinterpretation, not a transpose
    y = _lin(sd, "rhy_encoder.fc2", _lin(sd, "rhy_encoder.fc1", y))
    h = _bigru_last(sd, "rhy_encoder.gru", y)
    return _lin(sd, "rhy_encoder.linear_mu", h), _lin(sd, "rhy_encoder.linear_var", h).exp()


# ----------------------------------------------------------------------------------------------
# PianoTree decoder                                                           ptvae.py:292-496
def grid_lengths(x):
    return N_NOTE - (x[..., 0] == P_PAD).sum(-1)


def grid_multihot(x):
    oh = torch.zeros(x.shape[:-1] + (P_RANGE + 1,), dtype=torch.float32)
    oh.scatter_(-1, x[..., :1], 1.0)
    return torch.cat([oh[..., :P_RANGE], x[..., 1:].float()], -1)


def embed_grid(sd, x):
    return _lin(sd, "decoder.note_embedding", grid_multihot(x)), grid_lengths(x)


def _token_embed(sd, pitch_idx, dur_idx):
    tok = torch.zeros(pitch_idx.size(0), P_RANGE + N_DUR)
    tok[torch.arange(pitch_idx.size(0)), pitch_idx] = 1.0
    tok[:, P_RANGE:] = dur_idx.float()
    return _lin(sd, "decoder.note_embedding", tok)


def _decode_one_note(sd, h_note):
    """h_note (B,1,512) -> pitch logits (B,130), dur logits (B,5,2).        ptvae.py:336-368"""
    bs = h_note.size(0)
    pitch = _lin(sd, "decoder.pitch_out_linear", h_note).squeeze(1)
    dh = _lin(sd, "decoder.dur_hid_linear",
              torch.cat([h_note.transpose(0, 1), pitch.unsqueeze(0)], -1))
    tok = sd["decoder.dur_sos_token"].repeat(bs, 1).unsqueeze(1)
    durs = []
    for k in range(N_DUR):
        tok, dh = _gru_step(sd, "decoder.dec_dur_gru", tok, dh)
        d = _lin(sd, "decoder.dur_out_linear", tok).squeeze(1)
        durs.append(d)
        if k == N_DUR - 1:
            break
        nxt = torch.zeros(bs, N_DUR)
        nxt[torch.arange(bs), d.argmax(1)] = 1.0      # 1 at index == bit value (0/1)
        tok = nxt.unsqueeze(1)
    return pitch, torch.stack(durs, 1)


def _decode_step_notes(sd, summary, gt_notes, inference, tf_row):
    """summary (B,1,1024); gt_notes (B,16,128) or None.                     ptvae.py:370-428"""
    bs = summary.size(0)
    h = _lin(sd, "decoder.dec_time_to_notes_hid", summary.transpose(0, 1))
    if inference:
        sos = torch.zeros(P_RANGE + N_DUR)
        sos[P_SOS] = 1.0
        sos[P_RANGE:] = 2.0
        tok = _lin(sd, "decoder.note_embedding", sos).repeat(bs, 1).unsqueeze(1)
    else:
        tok = gt_notes[:, 0].unsqueeze(1)
    pred = torch.zeros(bs, N_NOTE, 128)
    pred[:, 0] = tok.squeeze(1)
    lens = torch.zeros(bs)
    pitches, durs = [], []
    for n in range(1, N_NOTE):
        out, h = _gru_step(sd, "decoder.dec_notes_gru", torch.cat([summary, tok], -1), h)
        p, d = _decode_one_note(sd, out)
        pitches.append(p)
        durs.append(d)
        p_idx, d_idx = p.argmax(1), d.argmax(2)
        emb = _token_embed(sd, p_idx, d_idx)
        pred[:, n] = emb
        lens[(p_idx == P_EOS) & (lens == 0)] = n
        if n == N_NOTE - 1:
            break
        if inference or not tf_row[n - 1]:
            tok = emb.unsqueeze(1)
        else:
            tok = gt_notes[:, n].unsqueeze(1)
    lens[lens == 0] = N_NOTE - 1
    return torch.stack(pitches, 1), torch.stack(durs, 1), pred, lens


def pianotree_decoder(sd, z, inference, emb_x, lengths, plan):
    bs = z.size(0)
    h = _lin(sd, "decoder.z2dec_hid_linear", z).unsqueeze(0)
    z_in = _lin(sd, "decoder.z2dec_in_linear", z).unsqueeze(1)
    if not inference:
        summ = _bigru_last_packed(sd, "decoder.dec_notes_emb_gru",
                                  emb_x.reshape(-1, N_NOTE, 128), lengths).view(bs, N_STEP, 256)
    tok = sd["decoder.dec_init_input"].repeat(bs, 1).unsqueeze(1)
    pitches, durs = [], []
    for t in range(N_STEP):
        out, h = _gru_step(sd, "decoder.dec_time_gru", torch.cat([tok, z_in], -1), h)
        p, d, pred, plen = _decode_step_notes(sd, out, None if inference else emb_x[:, t],
                                              inference, plan.tf_note[t])
        pitches.append(p)
        durs.append(d)
        if t == N_STEP - 1:
            break
        if plan.tf_time[t] and not inference:
            tok = summ[:, t].unsqueeze(1)
        else:
            tok = _bigru_last_packed(sd, "decoder.dec_notes_emb_gru", pred, plen).unsqueeze(1)
    return torch.stack(pitches, 1), torch.stack(durs, 1)


# ----------------------------------------------------------------------------------------------
# chord decoder                                                                ptvae.py:51-87
def chord_decoder(sd, z_chd, c, tf_chd):
    bs = z_chd.size(0)
    h = _lin(sd, "chd_decoder.z2dec_hid", z_chd).unsqueeze(0)
    z_in = _lin(sd, "chd_decoder.z2dec_in", z_chd).unsqueeze(1)
    tok = sd["chd_decoder.init_input"].repeat(bs, 1).unsqueeze(1)
    roots, chromas, basses = [], [], []
    for t in range(N_STEP // 4):
        out, h = _gru_step(sd, "chd_decoder.gru", torch.cat([tok, z_in], -1), h)
        r = _lin(sd, "chd_decoder.root_out", out)
        ch = _lin(sd, "chd_decoder.chroma_out", out).view(bs, 1, 12, 2)
        b = _lin(sd, "chd_decoder.bass_out", out)
        roots.append(r)
        chromas.append(ch)
        basses.append(b)
        # batch-union one-hot feedback: every sample sees every sample's argmax
        t_root = torch.zeros(12)
        t_root[r.argmax(-1).reshape(-1)] = 1.0
        t_bass = torch.zeros(12)
        t_bass[b.argmax(-1).reshape(-1)] = 1.0
        tok = torch.cat([t_root.expand(bs, 1, 12), ch.argmax(-1).float(),
                         t_bass.expand(bs, 1, 12)], -1)
        if tf_chd[t]:
            tok = c[:, t].unsqueeze(1)
    return torch.cat(roots, 1), torch.cat(chromas, 1), torch.cat(basses, 1)


# ----------------------------------------------------------------------------------------------
# losses                                                   ptvae.py:498-529, model.py:57-90
def kl_to_std_normal(mu, std):
    return (-std.log() + 0.5 * (std * std + mu * mu) - 0.5).mean()


def recon_loss(x, pitch_logits, dur_logits, weights=(1, 0.5)):
    ce = torch.nn.functional.cross_entropy
    pl = ce(pitch_logits.reshape(-1, P_RANGE), x[:, :, 1:, 0].reshape(-1), ignore_index=P_PAD)
    dl = ce(dur_logits.reshape(-1, 2), x[:, :, 1:, 1:].reshape(-1), ignore_index=D_PAD)
    return weights[0] * pl + weights[1] * dl, pl, dl


def chord_loss(c, root, chroma, bass):
    ce = torch.nn.functional.cross_entropy
    lr = ce(root.reshape(-1, 12), c[:, :, 0:12].argmax(-1).reshape(-1))
    lc = ce(chroma.reshape(-1, 2), c[:, :, 12:24].long().reshape(-1))
    lb = ce(bass.reshape(-1, 12), c[:, :, 24:].argmax(-1).reshape(-1))
    return lr + lc + lb, lr, lc, lb


# ----------------------------------------------------------------------------------------------
# model-level entry points                                                     model.py:42-149
def run(sd, x, c, pr_mat, plan, eps_chd, eps_rhy):
    """-> pitch (B,32,15,130), dur (B,32,15,5,2), (mu,std) chd, (mu,std) rhy, root, chroma, bass."""
    emb, lengths = embed_grid(sd, x)
    mu_c, sd_c = chord_encoder(sd, c)
    mu_r, sd_r = texture_encoder(sd, pr_mat)
    z_c = mu_c + sd_c * eps_chd
    z_r = mu_r + sd_r * eps_rhy
    pitch, dur = pianotree_decoder(sd, torch.cat([z_c, z_r], -1), False, emb, lengths, plan)
    root, chroma, bass = chord_decoder(sd, z_c, c, plan.tf_chd)
    return pitch, dur, (mu_c, sd_c), (mu_r, sd_r), root, chroma, bass


def loss_function(x, c, pitch, dur, dist_chd, dist_rhy, root, chroma, bass, beta, weights):
    rl, pl, dl = recon_loss(x, pitch, dur, weights)
    kc, kr = kl_to_std_normal(*dist_chd), kl_to_std_normal(*dist_rhy)
    cl, lr, lc, lb = chord_loss(c, root, chroma, bass)
    kl = kc + kr
    return rl + beta * kl + cl, rl, pl, dl, kl, kc, kr, cl, lr, lc, lb


def loss(sd, x, c, pr_mat, plan, eps_chd, eps_rhy, beta=0.1, weights=(1, 0.5)):
    return loss_function(x, c, *run(sd, x, c, pr_mat, plan, eps_chd, eps_rhy), beta, weights)


def greedy_decode(sd, z_chd, z_rhy):
    """-> est_x (B,32,15,6) int64 numpy (argmax pitch, 5 argmax dur bits).  model.py:124-131"""
    plan = TeacherPlan([[False] * 14] * N_STEP, [False] * (N_STEP - 1), [])
    with torch.no_grad():
        p, d = pianotree_decoder(sd, torch.cat([z_chd, z_rhy], -1), True, None, None, plan)
        return torch.cat([p.argmax(-1, keepdim=True), d.argmax(-1)], -1).numpy(), p, d


def inference(sd, pr_mat, c, eps_chd=None, eps_rhy=None):
    """Greedy inference; ``eps`` None = posterior means (what ``swap`` uses).   model.py:133-149"""
    with torch.no_grad():
        mu_c, sd_c = chord_encoder(sd, c)
        mu_r, sd_r = texture_encoder(sd, pr_mat)
        z_c = mu_c if eps_chd is None else mu_c + sd_c * eps_chd
        z_r = mu_r if eps_rhy is None else mu_r + sd_r * eps_rhy
    return greedy_decode(sd, z_c, z_r)[0]
