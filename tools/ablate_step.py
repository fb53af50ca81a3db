"""Where does the (stream-overlapped, graph-replayed) training step spend its time?  Time the step with one
component at a time replaced by a cheap stand-in (results are then wrong -- timing only)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops, ptvae
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch
dev = torch.device("cuda:0")
B = 512
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 0))

def timed(name, patch=None, unpatch=None):
    torch.manual_seed(0); random.seed(0)
    m = DisentangleVAE.init_model(device=dev).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
    if patch: patch(m)
    g = GraphedTrainStep(m, opt, B, warmup=2).capture(x, c, pr)
    for _ in range(3): g(x, c, pr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): g(x, c, pr)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1)/10:7.2f} ms/step", flush=True)
    if unpatch: unpatch()
    del g, m, opt

timed("full step")
orig_sum = ptvae.PtvaeDecoder._summarize
def p1(m):
    ptvae.PtvaeDecoder._summarize = lambda self, notes, l: ops.linear(notes[:, 0], self.dec_notes_emb_gru.weight_ih_l0[:256, :], None) + 0 * notes.sum((1, 2))[:, None]
timed("no note-summary bi-GRU", p1, lambda: setattr(ptvae.PtvaeDecoder, "_summarize", orig_sum))
orig_durs = ptvae.PtvaeDecoder._decode_durs
def p2(m):
    ptvae.PtvaeDecoder._decode_durs = lambda self, h, p, folded=None: ops.linear(h, self.dur_hid_linear.weight[:10, :512], None).view(-1, 5, 2)
timed("no duration decoder", p2, lambda: setattr(ptvae.PtvaeDecoder, "_decode_durs", orig_durs))
orig_enc = ptvae._bigru_final
def p3(m):
    def fake(gru, xx, lengths=None):
        w_ih = gru.weight_ih_l0
        return ops.linear(xx[:, 0], w_ih[:2 * gru.hidden_size], None)
    ptvae._bigru_final = fake
timed("no bi-GRUs at all (encoders + summary)", p3, lambda: setattr(ptvae, "_bigru_final", orig_enc))
ops.FORK_STREAMS = False
timed("full step, single stream")
