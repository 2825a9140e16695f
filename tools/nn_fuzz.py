"""Randomised agreement check of the two nn_distance kernels: filtered search (default) vs direct kernel, all four outputs bit for bit,
over random shapes (sizes that do not divide tiles / groups / vector widths) and input families.  python tools/nn_fuzz.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rfnet_b200 import ops

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
dev = torch.device("cuda:0")


def family(kind, b, n):
    x = rng.random((b, n, 3), dtype=np.float32) - 0.5
    if kind == "shifted":
        x += np.float32(rng.uniform(-500, 500))
    elif kind == "scaled":
        x *= np.float32(10.0 ** rng.uniform(-12, 8))
    elif kind == "lattice":
        x = np.floor(x * 32) / 32
    elif kind == "padded":
        k = int(rng.integers(1, max(2, n // 3)))
        pick = rng.integers(0, k, size=(b, n - k))
        x = np.concatenate([x[:, :k], np.take_along_axis(x[:, :k], pick[..., None].repeat(3, axis=2), axis=1)], axis=1)
    elif kind == "plane":
        x[..., 2] = np.float32(0.25)
    elif kind == "clusters":
        centres = rng.random((b, 7, 3), dtype=np.float32) - 0.5
        x = centres[:, rng.integers(0, 7, size=n)] + (x * np.float32(1e-3))
    return np.ascontiguousarray(x, dtype=np.float32)


kinds = ["uniform", "shifted", "scaled", "lattice", "padded", "plane", "clusters"]
bad = 0
scans_total = queries_total = 0
for t in range(cases):
    b = int(rng.integers(1, 7))
    n = int(rng.integers(1024, 9000)) if rng.random() < 0.8 else int(rng.choice([1024, 2048, 4096, 8192, 16384]))
    m = int(rng.integers(1024, 9000)) if rng.random() < 0.8 else int(rng.choice([1024, 2048, 4096, 8192, 16384]))
    while 2.0 * b * n * m < 2 ** 24:
        b += 1
    k1, k2 = rng.choice(kinds), rng.choice(kinds)
    x1, x2 = torch.from_numpy(family(k1, b, n)).to(dev), torch.from_numpy(family(k2, b, m)).to(dev)
    if rng.random() < 0.15:
        x2 = x1[:, : min(n, m)].clone() if min(n, m) >= 1024 else x2      # coincident clouds
        m = x2.shape[1]
    unfused = bool(rng.random() < 0.3)
    d1, i1, d2, i2, scans = ops.nn_distance_exact_scans(x1, x2, unfused)
    e = ops.nn_distance_op(x1, x2, unfused, True)
    ok = torch.equal(d1, e[0]) and torch.equal(i1, e[1]) and torch.equal(d2, e[2]) and torch.equal(i2, e[3])
    scans_total += scans
    queries_total += b * (n + m)
    if not ok:
        bad += 1
        print("MISMATCH case %d: b=%d n=%d m=%d %s/%s unfused=%s: %d + %d + %d + %d entries differ" % (
            t, b, n, m, k1, k2, unfused, int((d1 != e[0]).sum()), int((i1 != e[1]).sum()), int((d2 != e[2]).sum()), int((i2 != e[3]).sum())), flush=True)
print("nn_fuzz: %d cases, %d mismatching, %.2f %% of queries scanned exactly overall" % (cases, bad, 100.0 * scans_total / max(1, queries_total)))
sys.exit(1 if bad else 0)
