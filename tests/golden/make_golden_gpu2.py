"""Generates ref_gpu2.npz: SelectionSort and AuctionMatch outputs of the REFERENCE's own CUDA kernels (tf_grouping_g.cu and
tf_auctionmatch_g.cu inside oracle/_ref/libref_gpu.so, recompiled unchanged for sm_100a) on a B200:

    gpurun -- 'python tests/golden/make_golden_gpu2.py gpurun_out/ref_gpu2.npz'   then copy the file to tests/golden/

Neither op has a CPU implementation in the reference; these vectors pin oracle/rfnet_oracle.c (rfo_selection_sort,
rfo_auction_match) to what the reference kernels actually produce.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(777)
u = lambda *s: (rng.random(s, dtype=np.float32) - 0.5).astype(np.float32)
T = lambda a: torch.from_numpy(a).to(dev)
out = {}

for tag, (b, m, n, k) in {"a": (2, 7, 40, 5), "b": (1, 300, 257, 32), "c": (2, 3, 64, 64)}.items():
    d = rng.random((b, m, n), dtype=np.float32)
    if tag == "b":
        d[:, :, 100:140] = d[:, :, 20:60]  # equal values: first-index tie rule
    outi, o = ref.run_gpu("SelectionSort", [T(d)], [((b, m, n), torch.int32), ((b, m, n), torch.float32)], attrs={"k": k})
    out.update({"sel_%s_dist" % tag: d, "sel_%s_k" % tag: np.array([k], np.int32), "sel_%s_outi" % tag: outi.cpu().numpy(), "sel_%s_out" % tag: o.cpu().numpy()})

# n = 4096 takes the reference kernel's register-prefetch path (n == blockDim.x * 8), 2048 / 1024 / 300 the 4-, 2- and 1-wide loops
for tag, (b, n) in {"a": (2, 300), "b": (1, 1024), "c": (1, 2048), "d": (1, 4096), "e": (3, 64)}.items():
    x1, x2 = u(b, n, 3), u(b, n, 3)
    ml, mr = ref.run_gpu("AuctionMatch", [T(x1), T(x2)], [((b, n), torch.int32), ((b, n), torch.int32)])
    torch.cuda.synchronize()
    out.update({"auc_%s_xyz1" % tag: x1, "auc_%s_xyz2" % tag: x2, "auc_%s_matchl" % tag: ml.cpu().numpy(), "auc_%s_matchr" % tag: mr.cpu().numpy()})

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_gpu2.npz"
os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
