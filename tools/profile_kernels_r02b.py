"""Round 2, late kernels: one launch each of the greedy-slot kernels at the batch-512 shapes of the free-running training
step (four-warps-per-tile duration decoder vs the one-warp-per-tile one, pick + embed, 32-unit x-folded fused step), the
parallel pack_order and the length-aware column sum, for `ncu --set full --profile-from-start off`.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02b_kernels \
        python tools/profile_kernels_r02b.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from polydis_b200 import ops, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
P = lambda t: None if t is None else t.data_ptr()
st = lambda: torch.cuda.current_stream().cuda_stream
B, H, E = 512, 512, 128
heads = torch.randn(B, 196, device=dev)
par = [torch.randn(192, 5, device=dev) * 0.3, torch.randn(192, device=dev) * 0.1, torch.randn(192, 64, device=dev) * 0.2,
       torch.randn(192, device=dev) * 0.1, torch.rand(5, device=dev), torch.randn(2, 64, device=dev) * 0.3,
       torch.randn(2, device=dev) * 0.1]
dlog = torch.empty(B, 5, 2, device=dev)
tok = torch.zeros(B, 6, device=dev, dtype=torch.int32)
lens = torch.zeros(B, device=dev, dtype=torch.int32)
emb_wt, emb_b = torch.randn(135, 128, device=dev), torch.randn(128, device=dev)
pred = torch.empty(B, 16, E, device=dev)
h_n, h_nb = torch.randn(B, H, device=dev) * 0.3, torch.empty(B, H, device=dev)
wn_hh = torch.randn(3 * H, H, device=dev) * 0.03
wn_ih = torch.randn(3 * H, 1024 + E, device=dev) * 0.03
bn_hh = torch.randn(3 * H, device=dev) * 0.1
gi_s = torch.randn(B, 3 * H, device=dev)
pred.normal_()
R = 16384
rng = np.random.RandomState(0)
lengths = torch.from_numpy(np.where(rng.rand(R) < 0.4, 2, rng.randint(3, 17, R)).astype(np.int32)).to(dev)
perm, inv, table = (torch.zeros(n, device=dev, dtype=torch.int32) for n in (R, R, 64))
dgh = torch.randn(R, 16, 384, device=dev)
db = torch.empty(384, device=dev)


def run():
    for limit in (2048, 0):          # four warps per tile, then one warp per tile
        _lib.lib.pd_dur_quad_max_notes(limit)
        _lib.call("pd_dur_decode_fwd", P(heads[:, 130:194]), 196, B, *[P(t) for t in par], P(dlog), None, 1, st())
    _lib.lib.pd_dur_quad_max_notes(2048)
    ops.greedy_pick_embed(heads[:, :130], dlog, 4, tok, lens, emb_wt, emb_b, pred[:, 4])
    w_x = wn_ih[:, 1024:]
    _lib.call("pd_gru_step_tmax", P(h_n), H, P(wn_hh), H, P(pred[:, 3]), 16 * E, P(w_x), w_x.stride(0), E, P(bn_hh), P(gi_s),
              3 * H, P(h_nb), H, None, 0, None, 0, B, H, st())
    _lib.call("pd_pack_order", P(lengths), R, P(perm), P(inv), P(table), st())
    _lib.call("pd_colsum_seq_f32", P(dgh), 384, R, 16, 384, P(lengths), P(db), 0, st())
    _lib.call("pd_colsum_f32", P(dgh), 384, R * 16, 384, P(db), 0, st())


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
