"""CPU tier: PCD read/write, resample_pcd and the results.csv writer (the host-side formats around the operator path)."""
import csv
import os

import numpy as np

from rfnet_b200 import io_util


def test_pcd_roundtrip_binary_and_ascii(tmp_path):
    rng = np.random.default_rng(0)
    pts = (rng.random((1234, 3)) - 0.5).astype(np.float32)
    for binary in (True, False):
        path = os.path.join(tmp_path, "c_%d.pcd" % binary)
        io_util.save_pcd(path, pts, binary=binary)
        back = io_util.read_pcd(path)
        assert back.shape == (1234, 3) and back.dtype == np.float64
        assert np.array_equal(back.astype(np.float32), pts)


def test_pcd_with_extra_fields_and_comments(tmp_path):
    path = os.path.join(tmp_path, "x.pcd")
    with open(path, "w") as f:
        f.write("# comment\nVERSION .7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA ascii\n"
                "0.5 1.5 -2 4.2e6\n1 2 3 0\n")
    assert np.array_equal(io_util.read_pcd(path), np.array([[0.5, 1.5, -2.0], [1.0, 2.0, 3.0]]))


def test_resample_pcd_matches_reference_semantics():
    rng = np.random.default_rng(1)
    pcd = rng.random((100, 3))
    assert np.array_equal(io_util.resample_pcd(pcd, 40), pcd[:40])              # drop: keeps the first n
    up = io_util.resample_pcd(pcd, 250, rng=np.random.default_rng(2))
    assert up.shape == (250, 3) and np.array_equal(up[:100], pcd)                # duplicate: originals first, then random repeats
    assert all(any(np.array_equal(p, q) for q in pcd) for p in up[100:110])


def test_results_csv(tmp_path):
    w = io_util.ResultsCsv(str(tmp_path))
    w.add("02691156/a", 0.004, 0.002)
    w.add("02691156/b", 0.006, 0.004)
    w.add("03001627/c", 0.010, 0.003)
    means = w.close()
    rows = list(csv.reader(open(os.path.join(tmp_path, "results.csv"))))
    assert rows[0] == ["id", "cd", "emd"] and len(rows) == 4
    assert abs(means["02691156"][0] - 0.005) < 1e-12 and abs(means["03001627"][1] - 0.003) < 1e-12
