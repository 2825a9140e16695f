"""Generates tests/golden/ref_cpu.npz from the REFERENCE's own CPU OpKernels (oracle/_ref/libref_cpu.so, i.e.
/root/reference/**/tf_*.cpp compiled unmodified).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors; these fixtures pin the oracle -- and through it the CUDA
kernels -- to outputs of the reference itself on seeded inputs, and travel to machines where the reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build, ref  # noqa: E402

build(with_ref=True)
rng = np.random.default_rng(20251017)
u = lambda *s: (rng.random(s, dtype=np.float32) - 0.5).astype(np.float32)
out = {}

# BASELINE config 1 in miniature + ragged shape: NnDistance / NnDistanceGrad (pc_distance/tf_nndistance.cpp:60-163)
for tag, (b, n, m) in {"a": (2, 96, 160), "b": (1, 257, 31)}.items():
    x1, x2 = u(b, n, 3), u(b, m, 3)
    d1, i1, d2, i2 = ref.nn_distance(x1, x2)
    g1, g2 = rng.standard_normal((b, n)).astype(np.float32), rng.standard_normal((b, m)).astype(np.float32)
    gx1, gx2 = ref.nn_distance_grad(x1, x2, g1, i1, g2, i2)
    out.update({"nn_%s_%s" % (tag, k): v for k, v in dict(xyz1=x1, xyz2=x2, dist1=d1, idx1=i1, dist2=d2, idx2=i2, gd1=g1, gd2=g2, gxyz1=gx1, gxyz2=gx2).items()})

# ApproxMatch / MatchCost / MatchCostGrad CPU kernels (pc_distance/tf_approxmatch.cpp:23-140); match in the CPU kernel's
# native (b, n, m) element order (SURVEY.md 8c divergence 2)
for tag, (b, n, m) in {"a": (2, 64, 64), "b": (1, 96, 48)}.items():
    x1, x2 = u(b, n, 3), u(b, m, 3)
    match = ref.approx_match(x1, x2)
    cost = ref.match_cost(x1, x2, match)
    g1, g2 = ref.match_cost_grad(x1, x2, match)
    out.update({"emd_%s_%s" % (tag, k): v for k, v in dict(xyz1=x1, xyz2=x2, match_nm=match.reshape(b, n, m), cost=cost, grad1=g1, grad2=g2).items()})

# ThreeNN / ThreeInterpolate / ThreeInterpolateGrad (tf_ops/interpolation/tf_interpolate.cpp:60-153), incl. the shapes of
# the reference's own test (tf_interpolate_op_test.py:9-21)
for tag, (b, n, m, c) in {"a": (1, 128, 8, 16), "b": (2, 200, 77, 5)}.items():
    x1, x2 = u(b, n, 3), u(b, m, 3)
    dist, idx = ref.three_nn(x1, x2)
    pts, w = rng.standard_normal((b, m, c)).astype(np.float32), rng.random((b, n, 3), dtype=np.float32)
    o = ref.three_interpolate(pts, idx, w)
    go = rng.standard_normal((b, n, c)).astype(np.float32)
    gp = ref.three_interpolate_grad(pts, idx, w, go)
    out.update({"interp_%s_%s" % (tag, k): v for k, v in dict(xyz1=x1, xyz2=x2, dist=dist, idx=idx, points=pts, weight=w, out=o, grad_out=go, grad_points=gp).items()})

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cpu.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
