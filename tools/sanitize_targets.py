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
mr_ = ops.approx_match_op(x1[:, :700].contiguous(), x2[:, :515].contiguous(), 1)   # reference-order mode, ragged
x = rnd(2, 5003)
idx = tf_sampling.farthest_point_sample(300, x)         # pruned FPS, ragged last cluster
q = tf_sampling.gather_point(x, idx)
bi, bc = tf_grouping.query_ball_point(0.15, 16, x, q)
d3, i3 = tf_interpolate.three_nn(x, q)
ml, mr = tf_auctionmatch.auction_match(rnd(2, 600), rnd(2, 600))
si, so = tf_grouping.select_top_k(5, torch.rand((2, 7, 300), generator=g).to(dev))
torch.cuda.synchronize()
print("sanitize targets ok", float(m.sum()), float(c.sum()), int(idx.sum()), int(bc.sum()), int(i3.sum()), int(ml.sum()), int(si.sum()))
