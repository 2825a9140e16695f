"""GPU parity of query_ball_point / group_point / three_nn / three_interpolate (+grads)."""
import numpy as np
import pytest
import torch

from conftest import cloud
from oracle import port, ref

pytestmark = pytest.mark.gpu


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("b,n,m,ns,r", [(1, 128, 8, 32, 0.3), (2, 1000, 77, 16, 0.1), (2, 5000, 300, 32, 0.1), (3, 2050, 129, 5, 0.05),
                                        (2, 300, 40, 64, 10.0), (2, 300, 40, 8, 1e-6), (1, 4500, 33, 1, 0.2)])
def test_query_ball_point_bit_exact(cuda, rng, b, n, m, ns, r):
    """idx and pts_cnt bit-exact.  (1,128,8,32,0.3) is the shape of the reference's own test, tf_grouping_op_test.py:9-25.
    r = 10 covers every point in the ball (early exit), r = 1e-6 covers empty rows (defined as 0 in both)."""
    from rfnet_b200 import tf_grouping
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    x2[:, : min(m, 5)] = x1[:, : min(m, 5)]  # queries that coincide with dataset points (d = 0 -> max(.,1e-20) branch)
    wi, wc = port.query_ball_point(r, ns, x1, x2, fill_empty=0)
    gi, gc = tf_grouping.query_ball_point(r, ns, t(x1, cuda), t(x2, cuda))
    assert np.array_equal(gc.cpu().numpy(), wc)
    assert np.array_equal(gi.cpu().numpy(), wi)


@pytest.mark.parametrize("b,n,m,ns,r,kind", [(2, 16384, 2048, 32, 0.1, "cube"), (2, 5000, 777, 16, 0.07, "sphere"), (1, 3000, 300, 64, 0.3, "cube"),
                                             (1, 4096, 200, 8, 2.5, "cube"), (1, 4096, 100, 32, 1e-5, "dup"), (1, 2048, 64, 4, 0.05, "outside"),
                                             (1, 8000, 500, 32, 0.02, "line")])
def test_query_ball_point_grid_equals_scan(cuda, b, n, m, ns, r, kind):
    """From 2048 dataset points the op bins the dataset in a uniform grid and only tests the cells a ball touches; the
    result must be IDENTICAL to the full scan (RFNET_BALL_NO_GRID=1): volumes, surfaces, a radius larger than the cloud (every
    query overflows the hit buffer and is redone by the scan kernel), a radius below the coordinate resolution with duplicated
    points, queries outside the dataset's bounding box, and a degenerate (collinear) cloud."""
    from rfnet_b200 import tf_grouping
    g = torch.Generator(device="cpu").manual_seed(7 + n + m)
    x1 = torch.rand((b, n, 3), generator=g) - 0.5
    x2 = torch.rand((b, m, 3), generator=g) - 0.5
    if kind == "sphere":
        x1 = 0.5 * x1 / x1.norm(dim=-1, keepdim=True)
        x2 = 0.5 * x2 / x2.norm(dim=-1, keepdim=True)
    elif kind == "dup":
        x1[:, n // 2:] = x1[:, : n - n // 2]
        x2 = x1[:, :m].clone()
    elif kind == "outside":
        x2 = x2 * 3.0
    elif kind == "line":
        x1[:, :, 1:] = 0.25
        x2[:, :, 1:] = 0.25
    x1, x2 = x1.to(cuda), x2.to(cuda)
    gi, gc = tf_grouping.query_ball_point(r, ns, x1, x2)
    si, sc = torch.ops.rfnet.query_ball_point(x1, x2, torch.tensor([r], dtype=torch.float32, device=cuda), ns, False)   # no workspace: scan
    assert torch.equal(gc, sc) and torch.equal(gi, si)
    assert int(gc.min()) >= 0 and int(gc.max()) <= ns


def test_ball_threshold_is_exact(cuda):
    """The sqrt-free predicate: on distances straddling r the kernel must agree with max(sqrtf(d2),1e-20f) < r."""
    from rfnet_b200 import tf_grouping
    r = np.float32(0.1)
    # candidates at distance r*(1+k*2^-23) along x from the query, k = -40..40: d2 values are adjacent floats around r^2
    ks = np.arange(-40, 41)
    xs = (r * (1 + ks * 2.0 ** -23)).astype(np.float32)
    x1 = np.zeros((1, len(ks), 3), np.float32)
    x1[0, :, 0] = xs
    x2 = np.zeros((1, 1, 3), np.float32)
    wi, wc = port.query_ball_point(float(r), 128, x1, x2)
    gi, gc = tf_grouping.query_ball_point(float(r), 128, t(x1, cuda), t(x2, cuda))
    assert np.array_equal(gc.cpu().numpy(), wc) and np.array_equal(gi.cpu().numpy(), wi)
    assert 0 < wc[0, 0] < len(ks)


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
def test_query_ball_point_vs_reference_cuda_kernel(cuda, rng):
    from rfnet_b200 import tf_grouping
    b, n, m, ns = 4, 16384, 512, 32
    x1, x2 = t(cloud(rng, b, n), cuda), t(cloud(rng, b, m), cuda)
    r = torch.tensor([0.1], device=cuda)
    wi, wc = ref.run_gpu("QueryBallPoint", [x1, x2, r], [((b, m, ns), torch.int32), ((b, m), torch.int32)], attrs={"nsample": ns}, zero_outputs=True)
    gi, gc = tf_grouping.query_ball_point(r, ns, x1, x2)
    assert torch.equal(gc, wc) and torch.equal(gi, wi)


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
def test_config4_full_shape_vs_reference_cuda_kernels(cuda):
    """BASELINE config 4 at full size (B=32, dataset 16384, queries = the 2048 FPS points, r=0.1, nsample=32): idx and pts_cnt
    of EVERY query equal the reference CUDA kernel's (tf_grouping_g.cu:3-36), grid and scan variants; group_point (c=3, c=64)
    and its gradient equal the reference kernels' (:40-78; the gradient within float-atomic rounding)."""
    from rfnet_b200 import ops, tf_grouping, tf_sampling
    g = torch.Generator(device="cpu").manual_seed(11)
    b, n, m, ns = 32, 16384, 2048, 32
    x = (torch.rand((b, n, 3), generator=g) - 0.5).to(cuda)
    q = tf_sampling.gather_point(x, tf_sampling.farthest_point_sample(m, x))
    r = torch.tensor([0.1], device=cuda)
    wi, wc = ref.run_gpu("QueryBallPoint", [x, q, r], [((b, m, ns), torch.int32), ((b, m), torch.int32)], attrs={"nsample": ns}, zero_outputs=True)
    gi, gc = tf_grouping.query_ball_point(r, ns, x, q)
    assert torch.equal(gc, wc) and torch.equal(gi, wi)
    si, sc = torch.ops.rfnet.query_ball_point(x, q, r, ns, False)
    assert torch.equal(sc, wc) and torch.equal(si, wi)
    for c in (3, 64):
        pts = torch.randn((b, n, c), generator=g).to(cuda)
        (wg,) = ref.run_gpu("GroupPoint", [pts, wi], [((b, m, ns, c), torch.float32)])
        assert torch.equal(tf_grouping.group_point(pts, gi), wg)
        go = torch.randn((b, m, ns, c), generator=g).to(cuda)
        (wgg,) = ref.run_gpu("GroupPointGrad", [pts, wi, go], [((b, n, c), torch.float32)], zero_outputs=True)
        gg = ops.group_point_grad_op(pts, gi, go)
        assert float((gg - wgg).abs().max()) <= 1e-5 * float(wgg.abs().max())


@pytest.mark.parametrize("c", [1, 3, 16, 64, 6])
def test_group_point_and_grad(cuda, rng, c):
    from rfnet_b200 import ops, tf_grouping
    b, n, m, ns = 2, 300, 50, 9
    pts = rng.standard_normal((b, n, c)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m, ns)).astype(np.int32)
    got = tf_grouping.group_point(t(pts, cuda), t(idx, cuda)).cpu().numpy()
    assert np.array_equal(got, port.group_point(pts, idx))
    go = rng.standard_normal((b, m, ns, c)).astype(np.float32)
    want = port.group_point_grad(pts, idx, go)
    gg = ops.group_point_grad_op(t(pts, cuda), t(idx, cuda), t(go, cuda)).cpu().numpy()
    assert np.array_equal(gg, want)    # atomic-free scatter in the CPU prototype's order (query_ball_point.cpp:69-84): bit-exact


def test_group_point_gradient_check_like_reference(cuda):
    """tf_grouping_op_test.py:9-25: query_ball_point(0.3, 32) on (1,128,3)/(1,8,3) + group_point on (1,128,16);
    numeric-vs-analytic gradient error of GroupPointGrad < 1e-4."""
    from rfnet_b200 import tf_grouping
    r = np.random.default_rng(0)
    pts = torch.from_numpy(r.random((1, 128, 16)).astype(np.float32)).to(cuda)
    x1 = torch.from_numpy(r.random((1, 128, 3)).astype(np.float32)).to(cuda)
    x2 = torch.from_numpy(r.random((1, 8, 3)).astype(np.float32)).to(cuda)
    idx, _ = tf_grouping.query_ball_point(0.3, 32, x1, x2)
    w = torch.from_numpy(r.standard_normal((1, 8, 32, 16)).astype(np.float32)).to(cuda)
    p = pts.clone().requires_grad_(True)
    (tf_grouping.group_point(p, idx) * w).sum().backward()
    # the op is linear in points: the exact gradient is the scatter of w
    want = torch.zeros_like(pts).double()
    want.index_put_((torch.zeros_like(idx).long().reshape(-1), idx.long().reshape(-1)), w.double().reshape(-1, 16), accumulate=True)
    assert float((p.grad.double() - want).abs().max()) < 1e-4


@pytest.mark.parametrize("b,n,m", [(1, 128, 8), (2, 1000, 300), (2, 4100, 2049), (1, 50, 2), (1, 10, 1), (2, 300, 5003), (1, 700, 16384)])
def test_three_nn_bit_exact(cuda, rng, b, n, m):
    from rfnet_b200 import tf_interpolate
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    x2[:, m // 2:] = x2[:, : m - m // 2]  # duplicates: ties resolved by insertion order
    wd, wi = port.three_nn(x1, x2, fused=False)
    gd, gi = tf_interpolate.three_nn(t(x1, cuda), t(x2, cuda))
    assert np.array_equal(gi.cpu().numpy(), wi)
    assert np.array_equal(gd.cpu().numpy(), wd)


@pytest.mark.parametrize("b,n,m,kind", [(2, 16384, 2048, "cube"), (2, 5000, 4096, "sphere"), (1, 3000, 512, "dup"), (1, 2000, 1000, "outside"),
                                        (1, 4000, 3000, "line"), (1, 1000, 600, "same")])
def test_three_nn_grid_equals_scan(cuda, b, n, m, kind):
    """With 512..4096 known points three_nn searches a uniform grid (own cell, then shells, until the third distance is closer
    than any unvisited cell); distances AND indices must be identical to the full scan (RFNET_THREENN_NO_GRID=1), including
    which of several equidistant points comes first."""
    from rfnet_b200 import tf_interpolate
    g = torch.Generator(device="cpu").manual_seed(90 + n + m)
    x1 = torch.rand((b, n, 3), generator=g) - 0.5
    x2 = torch.rand((b, m, 3), generator=g) - 0.5
    if kind == "sphere":
        x1 = 0.5 * x1 / x1.norm(dim=-1, keepdim=True)
        x2 = 0.5 * x2 / x2.norm(dim=-1, keepdim=True)
    elif kind == "dup":
        x2[:, m // 2:] = x2[:, : m - m // 2]          # every known point twice: ties between indices
        x1[:, :200] = x2[:, :200]                     # zero distances
    elif kind == "outside":
        x1 = x1 * 4.0
    elif kind == "line":
        x2[:, :, 1:] = 0.1
    elif kind == "same":
        x2[:] = x2[:, :1]                              # all known points coincide: one cell, all distances equal
    x1, x2 = x1.to(cuda), x2.to(cuda)
    gd, gi = tf_interpolate.three_nn(x1, x2)
    sd, si = torch.ops.rfnet.three_nn(x1, x2, False)   # no workspace: the scan kernel
    assert torch.equal(gi, si) and torch.equal(gd, sd)


def test_three_nn_adversarial_order(cuda, rng):
    """Candidates sorted by DECREASING distance from the query cluster: every group improves the top 3, the per-query group
    list overflows and the kernel's exact re-scan path runs.  Still bit-exact."""
    from rfnet_b200 import tf_interpolate
    m = 3000
    x1 = (cloud(rng, 1, 700) * 0.01).astype(np.float32)                       # queries bunched at the origin
    x2 = cloud(rng, 1, m)
    order = np.argsort(-(x2[0] ** 2).sum(-1))
    x2 = np.ascontiguousarray(x2[:, order])
    wd, wi = port.three_nn(x1, x2, fused=False)
    gd, gi = tf_interpolate.three_nn(t(x1, cuda), t(x2, cuda))
    assert np.array_equal(gi.cpu().numpy(), wi) and np.array_equal(gd.cpu().numpy(), wd)


@pytest.mark.parametrize("c", [1, 5, 16, 64])
def test_three_interpolate_and_grad(cuda, rng, c):
    from rfnet_b200 import ops, tf_interpolate
    b, n, m = 2, 700, 90
    pts = rng.standard_normal((b, m, c)).astype(np.float32)
    idx = rng.integers(0, m, size=(b, n, 3)).astype(np.int32)
    w = rng.random((b, n, 3)).astype(np.float32)
    got = tf_interpolate.three_interpolate(t(pts, cuda), t(idx, cuda), t(w, cuda)).cpu().numpy()
    assert np.array_equal(got, port.three_interpolate(pts, idx, w))  # same unfused operation order -> bit-exact
    go = rng.standard_normal((b, n, c)).astype(np.float32)
    want = port.three_interpolate_grad(pts, idx, w, go)
    gg = ops.three_interpolate_grad_op(t(pts, cuda), t(idx, cuda), t(w, cuda), t(go, cuda)).cpu().numpy()
    assert np.array_equal(gg, want)    # same (j, u) order and unfused arithmetic as threeinterpolate_grad_cpu: bit-exact


@pytest.mark.parametrize("b,n,m,c", [(2, 6000, 24, 32), (3, 5001, 16, 16), (1, 60000, 300, 64), (2, 9000, 64, 48)])
def test_three_interpolate_staged_slices(cuda, rng, b, n, m, c):
    """Shapes that take three_interpolate_staged_kernel (rows of >= 16 channels, many unknown points per known point): a CTA keeps a
    16-channel slice of one cloud's known points in shared memory and streams a range of unknown points that may straddle clouds.
    Same expression as the plain gather kernel and the reference's CPU code: bit-exact, gradient included."""
    from rfnet_b200 import ops, tf_interpolate
    pts = rng.standard_normal((b, m, c)).astype(np.float32)
    idx = rng.integers(0, m, size=(b, n, 3)).astype(np.int32)
    idx[:, ::7, 1] = idx[:, ::7, 0]                              # repeated neighbours
    w = rng.random((b, n, 3)).astype(np.float32)
    got = tf_interpolate.three_interpolate(t(pts, cuda), t(idx, cuda), t(w, cuda)).cpu().numpy()
    assert np.array_equal(got, port.three_interpolate(pts, idx, w))
    go = rng.standard_normal((b, n, c)).astype(np.float32)
    gg = ops.three_interpolate_grad_op(t(pts, cuda), t(idx, cuda), t(w, cuda), t(go, cuda)).cpu().numpy()
    assert np.array_equal(gg, port.three_interpolate_grad(pts, idx, w, go))


def test_three_interpolate_gradient_check_like_reference(cuda):
    """tf_interpolate_op_test.py:9-21: points (1,8,16), xyz1 (1,128,3), xyz2 (1,8,3), weights 1/3; gradient error < 1e-4."""
    from rfnet_b200 import tf_interpolate
    r = np.random.default_rng(1)
    pts = torch.from_numpy(r.random((1, 8, 16)).astype(np.float32)).to(cuda)
    x1 = torch.from_numpy(r.random((1, 128, 3)).astype(np.float32)).to(cuda)
    x2 = torch.from_numpy(r.random((1, 8, 3)).astype(np.float32)).to(cuda)
    _, idx = tf_interpolate.three_nn(x1, x2)
    w = torch.full((1, 128, 3), 1.0 / 3.0, device=cuda)
    up = torch.from_numpy(r.standard_normal((1, 128, 16)).astype(np.float32)).to(cuda)
    p = pts.clone().requires_grad_(True)
    (tf_interpolate.three_interpolate(p, idx, w) * up).sum().backward()
    want = torch.zeros((1, 8, 16), device=cuda, dtype=torch.float64)
    for u in range(3):
        want.index_put_((torch.zeros(128, dtype=torch.long, device=cuda), idx[0, :, u].long()), (up[0].double() / 3.0), accumulate=True)
    assert float((p.grad.double() - want).abs().max()) < 1e-4


def test_config4_pipeline_properties(cuda):
    """BASELINE config 4 at full size: FPS 16384->2048, ball query r=0.1 k=32, group (c=3), three_nn + interpolate (c=64).
    Checked through size-independent properties (the small-shape tests above are the bit-exact ones)."""
    from rfnet_b200 import tf_grouping, tf_interpolate, tf_sampling
    g = torch.Generator(device="cpu").manual_seed(11)
    b, n, m, ns = 32, 16384, 2048, 32
    x = (torch.rand((b, n, 3), generator=g) - 0.5).to(cuda)
    q = tf_sampling.gather_point(x, tf_sampling.farthest_point_sample(m, x))
    idx, cnt = tf_grouping.query_ball_point(0.1, ns, x, q)
    assert int(cnt.min()) >= 1 and int(cnt.max()) <= ns           # every query is itself a dataset point
    grouped = tf_grouping.group_point(x, idx)                      # (b, m, ns, 3)
    d = (grouped - q[:, :, None, :]).norm(dim=-1)
    assert float(d.max()) < 0.1 + 1e-6                             # everything returned lies inside the ball
    asc = idx[:, :, 1:] >= idx[:, :, :-1]
    valid = torch.arange(1, ns, device=cuda)[None, None, :] < cnt[:, :, None]
    assert bool((asc | ~valid).all())                              # hits are in ascending dataset order
    dist, i3 = tf_interpolate.three_nn(x, q)
    assert bool((dist[..., 0] <= dist[..., 1]).all() and (dist[..., 1] <= dist[..., 2]).all())
    w = 1.0 / torch.clamp(dist, min=1e-10)
    w = w / w.sum(-1, keepdim=True)
    feats = torch.randn((b, m, 64), generator=torch.Generator(device="cpu").manual_seed(1)).to(cuda)
    out = tf_interpolate.three_interpolate(feats, i3, w)
    ref_out = (torch.gather(feats, 1, i3[..., 0].long()[..., None].expand(-1, -1, 64)) * w[..., 0:1]
               + torch.gather(feats, 1, i3[..., 1].long()[..., None].expand(-1, -1, 64)) * w[..., 1:2]
               + torch.gather(feats, 1, i3[..., 2].long()[..., None].expand(-1, -1, 64)) * w[..., 2:3])
    assert torch.allclose(out, ref_out, rtol=1e-5, atol=1e-6)


def test_group_point_grad_heavy_collisions_and_atomic_path(cuda, rng):
    """All rows gather the same few points (segment lengths in every class of the CSR sort); then the workspace-free
    (float-reduction) path of the C ABI on the same data, equal within rounding."""
    import ctypes
    from rfnet_b200 import _lib, ops
    b, n, m, ns, c = 2, 50, 300, 20, 8
    pts = rng.standard_normal((b, n, c)).astype(np.float32)
    idx = (rng.integers(0, 3, size=(b, m, ns)) * 7).astype(np.int32)          # only points 0, 7, 14 are ever gathered
    go = rng.standard_normal((b, m, ns, c)).astype(np.float32)
    want = port.group_point_grad(pts, idx, go)
    got = ops.group_point_grad_op(t(pts, cuda), t(idx, cuda), t(go, cuda))
    assert np.array_equal(got.cpu().numpy(), want)
    out = torch.empty((b, n, c), device=cuda)
    lib = _lib.load()
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    gi, gg = t(idx, cuda), t(go, cuda)
    rc = lib.rfnet_group_point_grad(b, n, c, m, ns, p(gg), p(gi), p(out), ctypes.c_void_p(0), 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert np.allclose(out.cpu().numpy(), want, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("b,n,m,k", [(2, 100, 33, 1), (2, 500, 200, 3), (1, 3000, 257, 16), (2, 2100, 130, 32), (1, 40, 5, 7)])
def test_knn_point_kernel_bit_exact(cuda, rng, b, n, m, k):
    """3-d knn_point without the (b,m,n) matrix: values (negated squared distances) and indices equal the oracle's
    top_k(-dist), including the lower-index-first tie rule on duplicated points."""
    from rfnet_b200 import tf_grouping
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    x1[:, n // 2:] = x1[:, : n - n // 2]          # duplicates -> ties
    wv, wi = port.knn_point(k, x1, x2)
    gv, gi = tf_grouping.knn_point(k, t(x1, cuda), t(x2, cuda))
    assert gi.dtype == torch.int32 and np.array_equal(gi.cpu().numpy(), wi)
    assert np.array_equal(gv.cpu().numpy(), wv)
    # agrees with the reference's framework formulation as well
    d = ((t(x1, cuda)[:, None] - t(x2, cuda)[:, :, None]) ** 2).sum(-1)
    tv, _ = torch.topk(-d, k=k, dim=-1)
    assert torch.allclose(gv, tv, rtol=1e-5, atol=1e-7)


def test_knn_point_gradient_and_errors(cuda, rng):
    """val = top_k(-dist) is differentiable in the reference (framework ops); the kernel path registers the same gradient:
    d val / d query = 2 (x - q), d val / d dataset[idx] = -2 (x - q), scattered with the atomic-free group_point gradient."""
    from rfnet_b200 import tf_grouping
    b, n, m, k = 2, 300, 50, 6
    x1 = t(cloud(rng, b, n), cuda).requires_grad_(True)
    x2 = t(cloud(rng, b, m), cuda).requires_grad_(True)
    wgt = torch.randn((b, m, k), generator=torch.Generator(device="cpu").manual_seed(3)).to(cuda)
    val, idx = tf_grouping.knn_point(k, x1, x2)
    (val * wgt).sum().backward()
    g1, g2 = x1.grad.clone(), x2.grad.clone()
    x1.grad = x2.grad = None
    d = ((x1[:, None] - x2[:, :, None]) ** 2).sum(-1)                         # the reference's formulation
    tv, ti = torch.topk(-d, k=k, dim=-1)
    assert torch.equal(ti.to(torch.int32), idx)
    (tv * wgt).sum().backward()
    assert torch.allclose(g1, x1.grad, rtol=1e-5, atol=1e-6) and torch.allclose(g2, x2.grad, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError, match="3-d points"):
        tf_grouping.knn_point(4, torch.zeros((1, 8, 5), device=cuda), torch.zeros((1, 2, 5), device=cuda))
    with pytest.raises(ValueError, match="k <= min"):
        tf_grouping.knn_point(40, torch.zeros((1, 64, 3), device=cuda), torch.zeros((1, 2, 3), device=cuda))


@pytest.mark.parametrize("c", [3, 16])
def test_scatter_plans_equal_unplanned_gradients(cuda, rng, c):
    """A scatter plan (the inverted idx, built once) gives the same bits as the gradient that builds its own, for group_point,
    gather_point and three_interpolate; autograd builds the plan at forward time and shares it between calls on the same idx."""
    from rfnet_b200 import ops, tf_grouping, tf_interpolate, tf_sampling
    b, n, m, ns = 2, 400, 150, 8
    pts = rng.standard_normal((b, n, c)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m, ns)).astype(np.int32)
    go = rng.standard_normal((b, m, ns, c)).astype(np.float32)
    tp, ti, tg = t(pts, cuda), t(idx, cuda), t(go, cuda)
    want = ops.group_point_grad_op(tp, ti, tg)
    plan = ops.scatter_plan_op(ti.reshape(b, -1), n)
    assert torch.equal(ops.group_point_grad_planned_op(tg, plan, n), want)
    assert np.array_equal(want.cpu().numpy(), port.group_point_grad(pts, idx, go))
    # autograd: two gathers through the same idx share one plan; results unchanged
    a1 = tp.clone().requires_grad_(True)
    a2 = tp.clone().requires_grad_(True)
    (tf_grouping.group_point(a1, ti) * tg).sum().backward()
    n_plans = len(ops._PLAN_CACHE)
    (tf_grouping.group_point(a2, ti) * tg).sum().backward()
    assert len(ops._PLAN_CACHE) == n_plans                                    # cache hit: no new plan
    assert torch.equal(a1.grad, want) and torch.equal(a2.grad, want)
    # three_interpolate: idx (b, n3, 3) into m known points
    n3 = 500
    i3 = rng.integers(0, m, size=(b, n3, 3)).astype(np.int32)
    w3 = rng.random((b, n3, 3)).astype(np.float32)
    feats = rng.standard_normal((b, m, c)).astype(np.float32)
    g3 = rng.standard_normal((b, n3, c)).astype(np.float32)
    want3 = ops.three_interpolate_grad_op(t(feats, cuda), t(i3, cuda), t(w3, cuda), t(g3, cuda))
    plan3 = ops.scatter_plan_op(t(i3, cuda).reshape(b, -1), m)
    assert torch.equal(ops.three_interpolate_grad_planned_op(t(g3, cuda), t(w3, cuda), plan3, m), want3)
    assert np.array_equal(want3.cpu().numpy(), port.three_interpolate_grad(feats, i3, w3, g3))
    f = t(feats, cuda).requires_grad_(True)
    (tf_interpolate.three_interpolate(f, t(i3, cuda), t(w3, cuda)) * t(g3, cuda)).sum().backward()
    assert torch.equal(f.grad, want3)
    # gather_point: group_point's gradient with one sample per row
    if c == 3:
        gi = rng.integers(0, n, size=(b, m)).astype(np.int32)
        og = rng.standard_normal((b, m, 3)).astype(np.float32)
        x = tp.clone().requires_grad_(True)
        tf_sampling.gather_point(x, t(gi, cuda)).backward(t(og, cuda))
        assert np.array_equal(x.grad.cpu().numpy(), port.gather_point_grad(pts, gi, og))


def test_scatter_gradient_every_segment_length_class(cuda, rng):
    """Segments of the inverted index are sorted by one thread (<= 16 entries), one warp (<= 1024), one CTA (bitonic, <= 8192)
    or left in claim order (longer): the gradient must equal the sequential oracle in every sorted class, and the long classes
    must not be slow (the first version's O(L^2) global-memory sort took tens of milliseconds at L = 8192)."""
    import time
    from rfnet_b200 import ops
    b, n, c = 2, 64, 4
    lengths = [1, 5, 16, 17, 300, 1024, 1025, 5000, 8192]
    rows = sum(lengths)
    idx = np.concatenate([np.full(L, t_, np.int32) for t_, L in enumerate(lengths)])
    perm = rng.permutation(rows)
    idx = np.stack([idx[perm], idx[rng.permutation(rows)]])[:, :, None]      # (b, rows, 1): target t_ receives lengths[t_] rows
    go = rng.standard_normal((b, rows, 1, c)).astype(np.float32)
    pts = np.zeros((b, n, c), np.float32)
    want = port.group_point_grad(pts, idx, go)
    tp, ti, tg = t(pts, cuda), t(idx, cuda), t(go, cuda)
    got = ops.group_point_grad_op(tp, ti, tg)
    assert np.array_equal(got.cpu().numpy(), want)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ops.group_point_grad_op(tp, ti, tg)
    torch.cuda.synchronize()
    assert (time.perf_counter() - t0) / 5 < 2e-3, "a long-segment sort is slow again"
