"""Single-shot workloads for ncu captures (see profiles/README.md).  python tools/profile_targets.py {emd16k|emd2k|fps|config4}"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfnet_b200 import ops, tf_approxmatch, tf_grouping, tf_interpolate, tf_sampling

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(3)
rnd = lambda b, n: (torch.rand((b, n, 3), generator=g) - 0.5).to(dev)
what = sys.argv[1] if len(sys.argv) > 1 else "emd16k"
if what in ("emd16k", "emd2k"):
    b, n = (4, 16384) if what == "emd16k" else (32, 2048)
    x1, x2 = rnd(b, n), rnd(b, n)
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for _ in range(2):
        m = ops.approx_match_op(x1, x2, flags)
        c = tf_approxmatch.match_cost(x1, x2, m)
        ops.match_cost_grad_op(x1, x2, m)              # one-pass gradient from a stored matrix
        del m
        ops.emd_cost_grad_op(x1, x2, flags)            # matrix-free cost + both gradients
elif what == "fps":
    x = rnd(32, 16384)
    for _ in range(2):
        tf_sampling.farthest_point_sample(2048, x)
elif what == "config4":
    x = rnd(32, 16384)
    for _ in range(2):
        idx = tf_sampling.farthest_point_sample(2048, x)
        q = tf_sampling.gather_point(x, idx)
        bi, _ = tf_grouping.query_ball_point(0.1, 32, x, q)
        tf_grouping.group_point(x, bi)
        f = torch.randn((32, 16384, 64), device=dev)
        grp = tf_grouping.group_point(f, bi)
        d, i3 = tf_interpolate.three_nn(x, q)
        w = torch.rand((32, 16384, 3), device=dev)
        feats = torch.randn((32, 2048, 64), device=dev)
        tf_interpolate.three_interpolate(feats, i3, w)
        ops.group_point_grad_op(f, bi, grp)
        ops.three_interpolate_grad_op(feats, i3, w, f)
torch.cuda.synchronize()
print("done", what)
