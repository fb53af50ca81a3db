import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polydis_b200 import ops
from polydis_b200.model import DisentangleVAE
from polydis_b200.graphs import GraphedTrainStep
from polydis_b200.synth import synth_batch
from polydis_b200.weights import make_state_dict
dev = torch.device("cuda:0")
B = 8
x, c, pr = (torch.from_numpy(a).to(dev) for a in synth_batch(B, 31))
torch.manual_seed(7)
eps = (torch.randn(B, 256, device=dev), torch.randn(B, 256, device=dev))

def fresh():
    m = DisentangleVAE.init_model(device=dev)
    m.load_state_dict(make_state_dict(2))
    return m.to(dev).train()

def show(tag, l):
    print(f"{tag:34s}", " ".join(f"{float(v):.6f}" for v in l), flush=True)

for defer in (True,):
    ops.DEFER_WGRAD = bool(defer)
    ops.DBG_WS_SKIP = {defer[5:]} if isinstance(defer, str) else set()
    m = fresh(); random.seed(7)
    l = m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5), eps=eps)
    torch.cuda.synchronize(); show(f"eager fwd only defer={defer}", l)
    l[0].backward(); torch.cuda.synchronize()
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters()))
    print("   grad norm", float(gn))
    with torch.no_grad():
        l = m('train', x, c, pr, tfr1=1., tfr2=1., tfr3=1., beta=0.1, weights=(1, 0.5), eps=eps)
    torch.cuda.synchronize(); show(f"eager no_grad defer={defer}", l)
    m = fresh(); random.seed(7)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
    dbg_store = ops._dbg = {}
    g = GraphedTrainStep(m, opt, B, warmup=2, inject_eps=True).capture(x, c, pr)
    g.eps[0].copy_(eps[0]); g.eps[1].copy_(eps[1])
    ops._dbg = None
    m3 = fresh()
    with torch.no_grad():
        w_eff3, b_eff3 = m3.decoder._dur_hid_folded()
        ref = dict(w_ph=torch.cat([m3.decoder.pitch_out_linear.weight, w_eff3], 0), b_ph=torch.cat([m3.decoder.pitch_out_linear.bias, b_eff3], 0),
                   w_eff=w_eff3, b_eff=b_eff3,
                   w_heads=torch.cat([m3.chd_decoder.root_out.weight, m3.chd_decoder.chroma_out.weight, m3.chd_decoder.bass_out.weight], 0),
                   b_heads=torch.cat([m3.chd_decoder.root_out.bias, m3.chd_decoder.chroma_out.bias, m3.chd_decoder.bass_out.bias], 0))
    dbg = dict(dbg_store)
    print({k: (v.data_ptr(), tuple(v.shape)) for k, v in dbg.items()})
    show(f"graph step 1 defer={defer}", g(x, c, pr).clone())
    # NOTE: after the replay the weights have been updated by one step; the captured weight-space tensors still hold what the
    # replay's forward computed (from the pre-update weights = fresh)
    for k in ref:
        print("   ", k, "max |graph tensor - fresh|", float((dbg[k].detach() - ref[k]).abs().max()))
    m2 = fresh()
    d = max(float((p - q).abs().max()) for p, q in zip(m.parameters(), m2.parameters()))
    print("   max |param - fresh| after 1 graph step", d)
