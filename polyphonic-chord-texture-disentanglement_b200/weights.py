"""State-dict contract of ``DisentangleVAE.init_model()`` and seeded random weights.

``STATE_DICT_SPEC`` lists the 81 tensors (name, shape, fan) a reference checkpoint holds
(model.py:244-265 builds RnnEncoder(36,1024,256), TextureEncoder(256,1024,256,10),
RnnDecoder(z_dim=256), PtvaeDecoder(dec_dur_hid_size=64, z_size=512); names are torch's
``named_parameters()`` of those modules, ptvae.py:11-122,218-290).  The drop-in model keeps the same
keys so ``load_model`` (amc_dl/torch_plus/module.py:46-53) works on reference checkpoints.

``make_state_dict(seed)`` draws weights from the same distributions torch's default initialisers
use (U(-1/sqrt(fan), 1/sqrt(fan)); ``torch.rand`` for the three free parameters) but from a private
CPU generator in key order, so the same tensors can be rebuilt on any box and loaded into either
implementation.  ``sharpen`` scales the two output heads so greedy-token margins are O(1)
(fixture W1 of SURVEY.md 7.4-2).
"""
import math
import torch


def _gru(prefix, inp, hid, bidir):
    out = []
    for suf in ([""] + (["_reverse"] if bidir else [])):
        out += [(f"{prefix}.weight_ih_l0{suf}", (3 * hid, inp), hid),
                (f"{prefix}.weight_hh_l0{suf}", (3 * hid, hid), hid),
                (f"{prefix}.bias_ih_l0{suf}", (3 * hid,), hid),
                (f"{prefix}.bias_hh_l0{suf}", (3 * hid,), hid)]
    return out


def _lin(prefix, inp, out):
    return [(f"{prefix}.weight", (out, inp), inp), (f"{prefix}.bias", (out,), inp)]


def state_dict_spec(chd_size=256, txt_size=256, num_channel=10, dur_hid=64):
    z = chd_size + txt_size
    s = []
    s += _gru("chd_encoder.gru", 36, 1024, True)
    s += _lin("chd_encoder.linear_mu", 2048, chd_size) + _lin("chd_encoder.linear_var", 2048, chd_size)
    s += [("rhy_encoder.cnn.0.weight", (num_channel, 1, 4, 12), 48),
          ("rhy_encoder.cnn.0.bias", (num_channel,), 48)]
    s += _lin("rhy_encoder.fc1", num_channel * 29, 1000) + _lin("rhy_encoder.fc2", 1000, 256)
    s += _gru("rhy_encoder.gru", 256, 1024, True)
    s += _lin("rhy_encoder.linear_mu", 2048, txt_size) + _lin("rhy_encoder.linear_var", 2048, txt_size)
    s += [("decoder.dec_init_input", (256,), 0), ("decoder.dur_sos_token", (5,), 0)]
    s += _lin("decoder.note_embedding", 135, 128)
    s += _lin("decoder.z2dec_hid_linear", z, 1024) + _lin("decoder.z2dec_in_linear", z, 256)
    s += _gru("decoder.dec_notes_emb_gru", 128, 128, True)
    s += _gru("decoder.dec_time_gru", 512, 1024, False)
    s += _lin("decoder.dec_time_to_notes_hid", 1024, 512)
    s += _gru("decoder.dec_notes_gru", 1152, 512, False)
    s += _lin("decoder.pitch_out_linear", 512, 130)
    s += _gru("decoder.dec_dur_gru", 5, dur_hid, False)
    s += _lin("decoder.dur_hid_linear", 642, dur_hid) + _lin("decoder.dur_out_linear", dur_hid, 2)
    s += [("chd_decoder.init_input", (36,), 0)]
    s += _lin("chd_decoder.z2dec_hid", chd_size, 512) + _lin("chd_decoder.z2dec_in", chd_size, 256)
    s += _gru("chd_decoder.gru", 292, 512, False)
    s += _lin("chd_decoder.root_out", 512, 12) + _lin("chd_decoder.chroma_out", 512, 24)
    s += _lin("chd_decoder.bass_out", 512, 12)
    return s


STATE_DICT_SPEC = state_dict_spec()


def make_state_dict(seed=0, sharpen=1.0, gain=1.0, eos_bias=0.0):
    """Seeded fp32 CPU state dict with the reference's keys/shapes (81 tensors, 27,310,079 elems).

    ``gain`` multiplies every matrix / conv weight (gain 2 puts the recurrences in a regime where
    greedy tokens differ between samples and steps -- a much harder token-parity fixture than the
    default init, whose decodes barely depend on the input); ``eos_bias`` is added to the EOS pitch
    logit bias so decoded steps end at varying lengths.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * int(seed) + 17)
    sd = {}
    for name, shape, fan in STATE_DICT_SPEC:
        u = torch.rand(shape, generator=g, dtype=torch.float32)
        if fan:
            b = 1.0 / math.sqrt(fan)
            u = (2.0 * u - 1.0) * b
        sd[name] = u
    if gain != 1.0:
        for k in sd:
            if sd[k].dim() >= 2:
                sd[k] = sd[k] * float(gain)
    if eos_bias != 0.0:
        sd["decoder.pitch_out_linear.bias"][129] += float(eos_bias)
    if sharpen != 1.0:
        for k in ("decoder.pitch_out_linear.weight", "decoder.pitch_out_linear.bias",
                  "decoder.dur_out_linear.weight", "decoder.dur_out_linear.bias"):
            sd[k] = sd[k] * float(sharpen)
    return sd


#: state dict of the alternative texture encoder ``PtvaeEncoder`` (ptvae.py:125-215 of the reference)
PTVAE_ENCODER_SPEC = (_lin("note_embedding", 135, 128) + _gru("enc_notes_gru", 128, 256, True)
                      + _gru("enc_time_gru", 512, 512, True) + _lin("linear_mu", 1024, 512) + _lin("linear_std", 1024, 512))


def make_ptvae_encoder_state(seed=0, gain=1.0):
    """Seeded weights for ``PtvaeEncoder`` with the reference's keys / shapes (same recipe as ``make_state_dict``)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7000003 * int(seed) + 29)
    sd = {}
    for name, shape, fan in PTVAE_ENCODER_SPEC:
        u = (2.0 * torch.rand(shape, generator=g, dtype=torch.float32) - 1.0) / math.sqrt(fan)
        sd[name] = u * float(gain) if u.dim() >= 2 else u
    return sd
