"""numpy front-end of librfnet_oracle.so (the plain-C restatement in rfnet_oracle.c).  TEST INFRASTRUCTURE ONLY.

Function names and argument orders follow the reference's Python wrappers (tf_ops/*/tf_*.py, pc_distance/tf_*.py) so
parity tests read like calls into the reference.
"""
import ctypes
import os

import numpy as np

from . import HERE, build

_lib = None
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "librfnet_oracle.so")
        src = os.path.join(HERE, "rfnet_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build(with_ref=False)
        _lib = ctypes.CDLL(path)
        _lib.rfo_ball_threshold.restype = ctypes.c_float
        _lib.rfo_ball_threshold.argtypes = [ctypes.c_float]
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def _check_xyz(*arrs):
    b = arrs[0].shape[0]
    for a in arrs:
        assert a.ndim == 3 and a.shape[2] == 3 and a.shape[0] == b, a.shape


def nn_distance(xyz1, xyz2, fused=True):
    """-> dist1 (b,n), idx1 (b,n), dist2 (b,m), idx2 (b,m).  tf_ops/CD/tf_nndistance.py:9-19."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    _check_xyz(xyz1, xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1, i1 = np.empty((b, n), np.float32), np.empty((b, n), np.int32)
    d2, i2 = np.empty((b, m), np.float32), np.empty((b, m), np.int32)
    lib().rfo_nn_distance(b, n, p1, m, p2, d1.ctypes.data_as(_f32p), i1.ctypes.data_as(_i32p),
                          d2.ctypes.data_as(_f32p), i2.ctypes.data_as(_i32p), int(bool(fused)))
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    """-> grad_xyz1, grad_xyz2.  tf_ops/CD/tf_nndistance.py:26-32."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    g1, pg1 = _f(grad_dist1)
    g2, pg2 = _f(grad_dist2)
    i1, pi1 = _i(idx1)
    i2, pi2 = _i(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    o1, o2 = np.empty_like(xyz1), np.empty_like(xyz2)
    lib().rfo_nn_distance_grad(b, n, p1, m, p2, pg1, pi1, pg2, pi2, o1.ctypes.data_as(_f32p), o2.ctypes.data_as(_f32p))
    return o1, o2


def approx_match(xyz1, xyz2, start_level=7):
    """GPU contract: match (b, m, n), 10 levels by default.  pc_distance/tf_approxmatch.py:10-18."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    _check_xyz(xyz1, xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.empty((b, m, n), np.float32)
    lib().rfo_approxmatch(b, n, m, p1, p2, match.ctypes.data_as(_f32p), int(start_level))
    return match


def approx_match_f64(xyz1, xyz2, start_level=7):
    """Same algorithm in double precision: the yardstick for the float32 noise band of approx_match on these inputs."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.empty((b, m, n), np.float32)
    lib().rfo_approxmatch_f64(b, n, m, p1, p2, match.ctypes.data_as(_f32p), int(start_level))
    return match


def approx_match_noise_band(xyz1, xyz2, match_f32=None):
    """max |float32 oracle - float64 oracle| relative to the largest entry: how far two correct float32 evaluations of
    approx_match may be expected to drift apart on these inputs."""
    m32 = approx_match(xyz1, xyz2) if match_f32 is None else match_f32
    m64 = approx_match_f64(xyz1, xyz2)
    return float(np.abs(m32 - m64).max() / max(float(np.abs(m64).max()), 1e-30))


def approx_match_tolerance(xyz1, xyz2, match_f32=None, floor=1e-4, factor=8.0, cap=5e-2):
    """Tolerance (relative to the largest entry) for comparing two float32 evaluations of approx_match on these inputs:
    `factor` x the measured float32 noise band, never tighter than `floor` (BASELINE.json's 1e-4) nor looser than `cap`."""
    return min(cap, max(floor, factor * approx_match_noise_band(xyz1, xyz2, match_f32)))


def approx_match_cpu_twin(xyz1, xyz2):
    """The reference's CPU variant: 11 levels, double accumulation, returned in its native (b, n, m) element order."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.empty((b, n, m), np.float32)
    lib().rfo_approxmatch_cpu_twin(b, n, m, p1, p2, match.ctypes.data_as(_f32p))
    return match


def match_cost(xyz1, xyz2, match):
    """match is (b, m, n).  -> cost (b,).  pc_distance/tf_approxmatch.py:27-36."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    match, pm = _f(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    assert match.shape == (b, m, n)
    cost = np.empty((b,), np.float32)
    lib().rfo_matchcost(b, n, m, p1, p2, pm, cost.ctypes.data_as(_f32p))
    return cost


def match_cost_grad(xyz1, xyz2, match):
    """-> grad1 (b,n,3), grad2 (b,m,3), before the grad_cost scaling.  pc_distance/tf_approxmatch.py:44-50."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    match, pm = _f(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1, g2 = np.empty_like(xyz1), np.empty_like(xyz2)
    lib().rfo_matchcostgrad(b, n, m, p1, p2, pm, g1.ctypes.data_as(_f32p), g2.ctypes.data_as(_f32p))
    return g1, g2


def farthest_point_sample(npoint, inp):
    """-> idx (b, npoint) int32.  tf_ops/sampling/tf_sampling.py:48-56."""
    inp, p = _f(inp)
    _check_xyz(inp)
    b, n, _ = inp.shape
    idx = np.empty((b, npoint), np.int32)
    lib().rfo_farthest_point_sample(b, n, int(npoint), p, idx.ctypes.data_as(_i32p))
    return idx


def gather_point(inp, idx):
    inp, p = _f(inp)
    idx, pi = _i(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.empty((b, m, 3), np.float32)
    lib().rfo_gather_point(b, n, m, p, pi, out.ctypes.data_as(_f32p))
    return out


def gather_point_grad(inp, idx, out_g):
    inp = np.asarray(inp)
    idx, pi = _i(idx)
    out_g, pg = _f(out_g)
    b, n, _ = inp.shape
    m = idx.shape[1]
    inp_g = np.empty((b, n, 3), np.float32)
    lib().rfo_gather_point_grad(b, n, m, pg, pi, inp_g.ctypes.data_as(_f32p))
    return inp_g


def query_ball_point(radius, nsample, xyz1, xyz2, fill_empty=0):
    """xyz1 dataset (b,n,3), xyz2 queries (b,m,3) -> idx (b,m,nsample), pts_cnt (b,m).  tf_grouping.py:8-20."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    r = np.array([radius], np.float32)
    idx = np.empty((b, m, nsample), np.int32)
    cnt = np.empty((b, m), np.int32)
    lib().rfo_query_ball_point(b, n, m, r.ctypes.data_as(_f32p), int(nsample), p1, p2, idx.ctypes.data_as(_i32p),
                               cnt.ctypes.data_as(_i32p), int(fill_empty))
    return idx, cnt


def group_point(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    lib().rfo_group_point(b, n, c, m, ns, pp, pi, out.ctypes.data_as(_f32p))
    return out


def group_point_grad(points, idx, grad_out):
    points = np.asarray(points)
    idx, pi = _i(idx)
    grad_out, pg = _f(grad_out)
    b, n, c = points.shape
    _, m, ns = idx.shape
    g = np.empty((b, n, c), np.float32)
    lib().rfo_group_point_grad(b, n, c, m, ns, pg, pi, g.ctypes.data_as(_f32p))
    return g


def knn_point(k, xyz1, xyz2):
    """xyz1 dataset (b,n,3), xyz2 queries (b,m,3) -> val (b,m,k) negated squared distances, idx (b,m,k).  tf_grouping.py:48-73."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    val = np.empty((b, m, k), np.float32)
    idx = np.empty((b, m, k), np.int32)
    lib().rfo_knn_point(b, n, m, int(k), p1, p2, val.ctypes.data_as(_f32p), idx.ctypes.data_as(_i32p))
    return val, idx


def select_top_k(k, dist):
    """dist (b,m,n) -> idx (b,m,n), dist_out (b,m,n): first k entries of every row are the k smallest, ascending.  tf_grouping.py:22-31."""
    dist, pd = _f(dist)
    b, m, n = dist.shape
    outi = np.empty((b, m, n), np.int32)
    out = np.empty((b, m, n), np.float32)
    lib().rfo_selection_sort(b, n, m, int(k), pd, outi.ctypes.data_as(_i32p), out.ctypes.data_as(_f32p))
    return outi, out


def auction_match(xyz1, xyz2):
    """xyz1, xyz2 (b,n,3) -> matchl (b,n): object of every xyz1 point, matchr (b,n): bidder of every xyz2 point.  tf_auctionmatch.py:11-20."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    _check_xyz(xyz1, xyz2)
    assert xyz1.shape == xyz2.shape
    b, n, _ = xyz1.shape
    matchl = np.empty((b, n), np.int32)
    matchr = np.empty((b, n), np.int32)
    lib().rfo_auction_match(b, n, p1, p2, matchl.ctypes.data_as(_i32p), matchr.ctypes.data_as(_i32p))
    return matchl, matchr


def three_nn(xyz1, xyz2, fused=False):
    """xyz1 unknown (b,n,3), xyz2 known (b,m,3) -> dist (b,n,3), idx (b,n,3).  tf_interpolate.py:8-18."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().rfo_three_nn(b, n, m, p1, p2, dist.ctypes.data_as(_f32p), idx.ctypes.data_as(_i32p), int(bool(fused)))
    return dist, idx


def three_interpolate(points, idx, weight):
    """points (b,m,c), idx/weight (b,n,3) -> (b,n,c).  tf_interpolate.py:20-28."""
    points, pp = _f(points)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    lib().rfo_three_interpolate(b, m, c, n, pp, pi, pw, out.ctypes.data_as(_f32p))
    return out


def three_interpolate_grad(points, idx, weight, grad_out):
    points = np.asarray(points)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    grad_out, pg = _f(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    g = np.empty((b, m, c), np.float32)
    lib().rfo_three_interpolate_grad(b, n, c, m, pg, pi, pw, g.ctypes.data_as(_f32p))
    return g


def ball_threshold(r):
    return float(lib().rfo_ball_threshold(float(np.float32(r))))


def nn_filter_bound_ratio(q, c, origin, fused=True):
    """max over all (query, candidate) pairs of |s + |q-o|^2 - d2_ref| / E for the expanded form the filtered NN search of
    rfnet_b200 scans (float32, the kernel's operation order) -- a numerical check of its error bound (<= 1 means it holds);
    test infrastructure only."""
    q, pq = _f(q)
    c, pc = _f(c)
    o, po = _f(np.asarray(origin, dtype=np.float32))
    fn = lib().rfo_nn_filter_bound_ratio
    fn.restype = ctypes.c_double
    return float(fn(q.shape[0], pq, c.shape[0], pc, po, int(bool(fused))))


def nn_filter_certificate_check(q, c, origin, fused=True):
    """(violations, certified pairs, tightest certified relative gap) of the filtered search's certificate over all (query, best
    candidate, other candidate) triples -- see rfo_nn_filter_certificate_violations; test infrastructure only."""
    q, pq = _f(q)
    c, pc = _f(c)
    o, po = _f(np.asarray(origin, dtype=np.float32))
    fn = lib().rfo_nn_filter_certificate_violations
    fn.restype = ctypes.c_long
    cert, tight = ctypes.c_long(0), ctypes.c_double(0.0)
    bad = fn(q.shape[0], pq, c.shape[0], pc, po, int(bool(fused)), ctypes.byref(cert), ctypes.byref(tight))
    return int(bad), int(cert.value), float(tight.value)
