"""Small invocations of the kernels added after the first sanitizer pass (pruned EMD sweeps, pruned FPS, two-phase three_nn,
128-wide ball query, auction_match, select_top_k, fused EMD cost), for compute-sanitizer.  python tools/sanitize_targets.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfnet_b200 import ops, tf_approxmatch, tf_auctionmatch, tf_grouping, tf_interpolate, tf_sampling

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(11)
rnd = lambda b, n: (torch.rand((b, n, 3), generator=g) - 0.5).to(dev)
x1, x2 = rnd(1, 4100), rnd(1, 4096)
m = tf_approxmatch.approx_match(x1, x2)                 # pruned sweeps (Morton sort, masks), ragged sizes
c, _ = ops.emd_cost_op(x1, x2, False)
cg, g1, g2 = ops.emd_cost_grad_op(x1, x2)              # matrix-free cost + both gradients (split sweeps, fused pass 3 + pass 1)
mr_ = ops.approx_match_op(x1[:, :700].contiguous(), x2[:, :515].contiguous(), 1)   # exact mode, ragged (plain-load staging: 515 % 4 != 0)
mp_ = ops.emd_cost_grad_op(x1, x2, 8)                  # compacted-gather pruned sweeps (Hilbert order, masks)
ms_ = ops.emd_cost_grad_op(x1[:, :1030].contiguous(), x2[:, :900].contiguous(), 4)  # split sums + epilogue kernels
mg_ = ops.match_cost_grad_op(x1[:, :1028].contiguous(), x2[:, :700].contiguous(), m[:, :700, :1028].contiguous())   # one-pass gradient
from rfnet_b200 import losses
raw, new = rnd(2, 900), rnd(2, 300)
dec = torch.tensor([0.05], device=dev, requires_grad=True)
nw = new.clone().requires_grad_(True)
losses.merge_layer(raw, nw, dec).sum().backward()      # fused merge_layer forward + backward (scatter plan)
f32, i32 = torch.float32, torch.int32
a1, a2 = rnd(3, 2048), rnd(3, 5000)
o = [torch.empty((3, 2048), dtype=f32, device=dev), torch.empty((3, 2048), dtype=i32, device=dev), torch.empty((3, 5000), dtype=f32, device=dev),
     torch.empty((3, 5000), dtype=i32, device=dev), torch.empty((3, 2048, 3), device=dev), torch.empty((3, 5000, 3), device=dev), torch.empty(4, device=dev)]
ws = torch.empty(ops.nn_distance_workspace_bytes(3, 2048, 5000), dtype=torch.uint8, device=dev)
ops.raw_chamfer_step(a1, a2, torch.ones((3, 2048), device=dev), torch.ones((3, 5000), device=dev), *o, ws)   # fused chamfer step
# filtered nn search: ragged sizes (plain-load resolve), a lattice (exact warp scans), repeated points (duplicate search, dead items)
for a, c_ in ((rnd(2, 2049), rnd(2, 4099)), (torch.floor(rnd(2, 2048) * 8) / 8, torch.floor(rnd(2, 4096) * 8) / 8),
              (rnd(2, 300).repeat(1, 8, 1), rnd(2, 4096))):
    ops.nn_distance_exact_scans(a.contiguous(), c_.contiguous())
    ops.nn_distance_op(a.contiguous(), c_.contiguous(), True, False)
x = rnd(2, 5003)
idx = tf_sampling.farthest_point_sample(300, x)         # pruned FPS, ragged last cluster
q = tf_sampling.gather_point(x, idx)
bi, bc = tf_grouping.query_ball_point(0.15, 16, x, q)
d3, i3 = tf_interpolate.three_nn(x, q)
ml, mr = tf_auctionmatch.auction_match(rnd(2, 600), rnd(2, 600))
si, so = tf_grouping.select_top_k(5, torch.rand((2, 7, 300), generator=g).to(dev))
# list-based select_top_k: vector and scalar row copies, two thresholds per lane, list overflow (plain sort of the row), whole-row kernel
for n_, k_ in ((1000, 40), (999, 7), (2048, 129)):
    tf_grouping.select_top_k(k_, torch.rand((2, 9, n_), generator=g).to(dev))
tf_grouping.select_top_k(16, (torch.rand((1, 9, 1200), generator=g) > 0.9).float().to(dev))
# staged three_interpolate (a 16-channel slice of the known points in shared memory) + the planned / standalone gradients
kn, un = rnd(2, 16), rnd(2, 8192)
d3s, i3s = tf_interpolate.three_nn(un, kn)
feats = torch.rand((2, 16, 32), generator=g).to(dev).requires_grad_(True)
tf_interpolate.three_interpolate(feats, i3s, torch.rand((2, 8192, 3), generator=g).to(dev)).sum().backward()
pts = torch.rand((2, 5003, 8), generator=g).to(dev).requires_grad_(True)
tf_grouping.group_point(pts, bi).sum().backward()     # CSR gradient with prefetched entry numbers
torch.cuda.synchronize()
print("sanitize targets ok", float(m.sum()), float(c.sum()), int(idx.sum()), int(bc.sum()), int(i3.sum()), int(ml.sum()), int(si.sum()))
