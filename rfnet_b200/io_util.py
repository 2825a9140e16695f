"""Host-side formats either side of the operator path: the PCD files recon_test.py reads
(io_util.py:7-15 uses open3d, which is not a dependency here), resample_pcd (data_util.py:8-13), and the results.csv the
test script writes (recon_test.py:42-44,68,92-100).  Pure numpy / stdlib -- no GPU work happens in this module.
"""
import csv
import os

import numpy as np

_PCD_TYPES = {("F", 4): np.float32, ("F", 8): np.float64, ("U", 1): np.uint8, ("U", 2): np.uint16, ("U", 4): np.uint32,
              ("I", 1): np.int8, ("I", 2): np.int16, ("I", 4): np.int32}


def read_pcd(filename):
    """-> (n, 3) float64 array of x, y, z, like the reference's read_pcd.  Supports DATA ascii and DATA binary."""
    with open(filename, "rb") as f:
        header = {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: no DATA line in PCD header" % filename)
            text = line.decode("ascii", "replace").strip()
            if not text or text.startswith("#"):
                continue
            key, _, rest = text.partition(" ")
            header[key.upper()] = rest.split()
            if key.upper() == "DATA":
                break
        fields = header["FIELDS"]
        sizes = [int(v) for v in header["SIZE"]]
        types = header["TYPE"]
        counts = [int(v) for v in header.get("COUNT", ["1"] * len(fields))]
        npts = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header.get("HEIGHT", ["1"])[0])
        for need in ("x", "y", "z"):
            if need not in fields:
                raise ValueError("%s: PCD has no '%s' field" % (filename, need))
        mode = header["DATA"][0].lower()
        if mode == "ascii":
            cols, off = {}, 0
            for name, c in zip(fields, counts):
                cols[name] = off
                off += c
            data = np.loadtxt(f, dtype=np.float64, ndmin=2) if npts else np.zeros((0, off))
            return np.ascontiguousarray(data[:npts, [cols["x"], cols["y"], cols["z"]]], dtype=np.float64)
        if mode == "binary":
            dt = np.dtype([(name, _PCD_TYPES[(t, s)], (c,) if c > 1 else ()) for name, t, s, c in zip(fields, types, sizes, counts)])
            rec = np.frombuffer(f.read(dt.itemsize * npts), dtype=dt, count=npts)
            return np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float64)
        raise ValueError("%s: PCD DATA mode '%s' is not supported (ascii and binary are)" % (filename, mode))


def save_pcd(filename, points, binary=True):
    """Write (n, 3) points as a PCD v0.7 file with float32 x y z fields (what open3d's write_point_cloud produces)."""
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    n = pts.shape[0]
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
            "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (n, n, "binary" if binary else "ascii"))
    with open(filename, "wb") as f:
        f.write(head.encode("ascii"))
        if binary:
            f.write(pts.tobytes())
        else:
            for p in pts:
                f.write(("%.9g %.9g %.9g\n" % (p[0], p[1], p[2])).encode("ascii"))


def resample_pcd(pcd, n, rng=None):
    """Drop or duplicate points so that pcd has exactly n points (data_util.py:8-13; the reference uses np.random)."""
    idx = np.arange(pcd.shape[0])
    if idx.shape[0] < n:
        extra = (rng.integers(pcd.shape[0], size=n - pcd.shape[0]) if rng is not None else np.random.randint(pcd.shape[0], size=n - pcd.shape[0]))
        idx = np.concatenate([idx, extra])
    return pcd[idx[:n]]


class ResultsCsv:
    """results.csv of recon_test.py: header id,cd,emd; one row per model; per-synset means at the end (recon_test.py:92-100)."""

    def __init__(self, results_dir):
        os.makedirs(results_dir, exist_ok=True)
        self._file = open(os.path.join(results_dir, "results.csv"), "w", newline="")
        self._writer = csv.writer(self._file)
        self._writer.writerow(["id", "cd", "emd"])
        self.cd_per_cat, self.emd_per_cat, self.count = {}, {}, 0

    def add(self, model_id, cd, emd):
        self._writer.writerow([model_id, cd, emd])
        synset = model_id.split("/")[0]
        self.cd_per_cat.setdefault(synset, []).append(float(cd))
        self.emd_per_cat.setdefault(synset, []).append(float(emd))
        self.count += 1

    def close(self):
        self._file.close()
        return {k: (float(np.mean(v)), float(np.mean(self.emd_per_cat[k]))) for k, v in self.cd_per_cat.items()}
