"""Generates ref_gpu.npz from the REFERENCE's own CUDA kernels (oracle/_ref/libref_gpu.so: tf_nndistance.cu,
tf_approxmatch.cu, tf_sampling_g.cu, tf_grouping_g.cu recompiled unchanged for sm_100a) on a B200:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/ref_gpu.npz'   then copy the file to tests/golden/

These pin the ops that have NO CPU implementation in the reference (FPS, ball query) and the GPU flavour of approx_match
(10 levels, float, (b,m,n) layout) to the reference's actual GPU results.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(424242)
u = lambda *s: (rng.random(s, dtype=np.float32) - 0.5).astype(np.float32)
T = lambda a: torch.from_numpy(a).to(dev)
out = {}

for tag, (b, n, m) in {"a": (2, 500, 1100), "b": (1, 1030, 64)}.items():
    x1, x2 = u(b, n, 3), u(b, m, 3)
    d1, i1, d2, i2 = ref.run_gpu("NnDistance", [T(x1), T(x2)], [((b, n), torch.float32), ((b, n), torch.int32), ((b, m), torch.float32), ((b, m), torch.int32)])
    out.update({"nn_%s_%s" % (tag, k): v for k, v in dict(xyz1=x1, xyz2=x2, dist1=d1.cpu().numpy(), idx1=i1.cpu().numpy(), dist2=d2.cpu().numpy(), idx2=i2.cpu().numpy()).items()})

for tag, (b, n, m) in {"a": (2, 256, 256), "b": (1, 600, 200)}.items():
    x1, x2 = u(b, n, 3), u(b, m, 3)
    (match,) = ref.run_gpu("ApproxMatch", [T(x1), T(x2)], [((b, m, n), torch.float32)])
    (cost,) = ref.run_gpu("MatchCost", [T(x1), T(x2), match], [((b,), torch.float32)])
    g1, g2 = ref.run_gpu("MatchCostGrad", [T(x1), T(x2), match], [((b, n, 3), torch.float32), ((b, m, 3), torch.float32)])
    out.update({"emd_%s_%s" % (tag, k): v for k, v in dict(xyz1=x1, xyz2=x2, match=match.cpu().numpy(), cost=cost.cpu().numpy(), grad1=g1.cpu().numpy(), grad2=g2.cpu().numpy()).items()})

for tag, (b, n, m) in {"a": (2, 3000, 128), "b": (1, 16384, 512), "c": (3, 700, 700)}.items():
    x = u(b, n, 3)
    if tag == "c":
        x[:, 350:] = x[:, :350]  # duplicated points: exercises the block's tie rule
    (idx,) = ref.run_gpu("FarthestPointSample", [T(x)], [((b, m), torch.int32)], attrs={"npoint": m})
    out.update({"fps_%s_inp" % tag: x, "fps_%s_idx" % tag: idx.cpu().numpy()})

for tag, (b, n, m, ns, r) in {"a": (1, 128, 8, 32, 0.3), "b": (2, 4000, 100, 32, 0.1), "c": (1, 2000, 50, 8, 0.05)}.items():
    x1, x2 = u(b, n, 3), u(b, m, 3)
    rad = np.array([r], np.float32)
    idx, cnt = ref.run_gpu("QueryBallPoint", [T(x1), T(x2), T(rad)], [((b, m, ns), torch.int32), ((b, m), torch.int32)], attrs={"nsample": ns}, zero_outputs=True)
    out.update({"ball_%s_%s" % (tag, k): v for k, v in dict(xyz1=x1, xyz2=x2, radius=rad, idx=idx.cpu().numpy(), cnt=cnt.cpu().numpy()).items()})

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_gpu.npz"
os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
