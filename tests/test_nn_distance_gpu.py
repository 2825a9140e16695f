"""GPU parity of nn_distance / its gradient against the CPU oracle (bit-exact indices AND distances)."""
import numpy as np
import pytest
import torch

from conftest import cloud
from oracle import port, ref

pytestmark = pytest.mark.gpu


def _run(cuda, x1, x2, unfused=False, direct=False):
    from rfnet_b200 import tf_nndistance
    out = tf_nndistance.nn_distance(torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda), unfused=unfused, direct=direct)
    return [o.cpu().numpy() for o in out]


# (b, n, m): ragged sizes, tiles that do not divide, TMA-aligned (m % 4 == 0) and unaligned rows, n or m tiny
SHAPES = [(1, 1, 1), (2, 5, 3), (3, 64, 1024), (2, 300, 257), (2, 1000, 1029), (1, 2048, 2048), (2, 2049, 4100), (4, 515, 8192), (1, 4096, 37)]


# the filtered search (default) takes over from 2^24 pairs; the last shapes are there for it
SHAPES_FILTER = [(2, 2049, 4100), (4, 515, 8192), (3, 3000, 2999), (1, 5000, 16390)]


@pytest.mark.parametrize("b,n,m", SHAPES + SHAPES_FILTER[2:])
@pytest.mark.parametrize("unfused", [False, True])
@pytest.mark.parametrize("direct", [False, True])
def test_nn_distance_bit_exact(cuda, rng, b, n, m, unfused, direct):
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    want = port.nn_distance(x1, x2, fused=not unfused)
    got = _run(cuda, x1, x2, unfused, direct)
    for name, g, w in zip(("dist1", "idx1", "dist2", "idx2"), got, want):
        assert g.dtype == w.dtype and g.shape == w.shape
        assert np.array_equal(g, w), "%s differs at %d positions" % (name, (g != w).sum())


def _padded(r, b, n, distinct):
    """a cloud of `distinct` points repeated to n rows the way the reference's resample_pcd pads short scans (data_util.py:8-13)"""
    base = (r.random((b, distinct, 3), dtype=np.float32) - 0.5)
    pick = r.integers(0, distinct, size=(b, n - distinct))
    return np.concatenate([base, np.take_along_axis(base, pick[..., None].repeat(3, axis=2), axis=1)], axis=1)


ADVERSARIAL = {
    # name -> (x1, x2) builder for (rng, b, n, m)
    "shifted_by_1000": lambda r, b, n, m: (cloud(r, b, n) + 1000.0, cloud(r, b, m) + 1000.0),
    "scale_1e6": lambda r, b, n, m: (cloud(r, b, n) * 1e6, cloud(r, b, m) * 1e6),
    "scale_1e-15": lambda r, b, n, m: (cloud(r, b, n) * 1e-15, cloud(r, b, m) * 1e-15),
    "scale_1e-22_underflow": lambda r, b, n, m: (cloud(r, b, n) * 1e-22, cloud(r, b, m) * 1e-22),
    "lattice_exact_ties": lambda r, b, n, m: (np.floor(cloud(r, b, n) * 16) / 16, np.floor(cloud(r, b, m) * 16) / 16),
    "padded_candidates_and_queries": lambda r, b, n, m: (_padded(r, b, n, 300), _padded(r, b, m, 777)),
    "same_cloud_twice": lambda r, b, n, m: (lambda x: (x[:, :n].copy(), x[:, :m].copy()))(cloud(r, b, max(n, m))),
    "tight_cluster_far_away": lambda r, b, n, m: (cloud(r, b, n) * 1e-3 + np.float32(37.5), cloud(r, b, m) * 1e-3 + np.float32(37.5)),
    "surface_with_outliers": lambda r, b, n, m: (np.concatenate([cloud(r, b, n - 3) * np.float32([1, 1, 1e-4]), np.full((b, 3, 3), 5e3, np.float32)], 1),
                                                  cloud(r, b, m) * np.float32([1, 1, 1e-4])),
}


@pytest.mark.parametrize("kind", sorted(ADVERSARIAL))
@pytest.mark.parametrize("b,n,m", [(2, 2049, 4100), (4, 4096, 16384), (1, 16384, 16384)])
def test_filtered_search_equals_direct_search(cuda, kind, b, n, m):
    """nn_filter_kernel (the default) against nn_search_kernel (the reference expression for every pair): all four outputs bit for
    bit, on inputs built to break the certification -- clouds far from the origin, tiny / huge scales, exact ties between distinct
    points, repeated points, coincident clouds -- in both distance contracts; plus the oracle on a slice."""
    r = np.random.default_rng(len(kind) * 1000 + n)
    x1, x2 = ADVERSARIAL[kind](r, b, n, m)
    x1, x2 = np.ascontiguousarray(x1, dtype=np.float32), np.ascontiguousarray(x2, dtype=np.float32)
    for unfused in (True, False):
        got = _run(cuda, x1, x2, unfused, direct=False)
        want = _run(cuda, x1, x2, unfused, direct=True)
        for name, g, w in zip(("dist1", "idx1", "dist2", "idx2"), got, want):
            assert np.array_equal(g, w), "%s %s: %d of %d differ" % (kind, name, (g != w).sum(), g.size)
    w = port.nn_distance(x1[:1, :40], x2[:1], fused=True)   # `got` is the fused run here
    assert np.array_equal(got[0][0, :40], w[0][0]) and np.array_equal(got[1][0, :40], w[1][0])


def test_filtered_search_certifies_almost_everything(cuda):
    """How often the filtered search falls back to the exact warp scan ((query, item) pairs over queries): rare on generic clouds,
    wherever they sit and however they are padded; exercised heavily by a lattice (which is what the test above relies on)."""
    from rfnet_b200 import ops
    r = np.random.default_rng(5)
    b, n, m = 4, 2048, 16384

    def frac(x1, x2):
        t1, t2 = torch.from_numpy(np.ascontiguousarray(x1, dtype=np.float32)).to(cuda), torch.from_numpy(np.ascontiguousarray(x2, dtype=np.float32)).to(cuda)
        d1, i1, d2, i2, scans = ops.nn_distance_exact_scans(t1, t2)
        e = ops.nn_distance_op(t1, t2, False, True)
        assert torch.equal(d1, e[0]) and torch.equal(i1, e[1]) and torch.equal(d2, e[2]) and torch.equal(i2, e[3])
        return scans / float(b * (n + m))

    cases = {"uniform": (cloud(r, b, n), cloud(r, b, m)), "shifted": (cloud(r, b, n) + 100.0, cloud(r, b, m) + 100.0),
             "padded_small": (_padded(r, b, n, 300), cloud(r, b, m)), "padded_large": (cloud(r, b, m)[:, :n], _padded(r, b, m, 2000))}
    got = {k: frac(*v) for k, v in cases.items()}
    print("exact-scan fractions:", got)
    assert all(v < 0.02 for v in got.values()), got
    assert frac(np.floor(cloud(r, b, n) * 16) / 16, np.floor(cloud(r, b, m) * 16) / 16) > 0.5


def test_nn_distance_ties_pick_lowest_index(cuda, rng):
    # duplicated candidates (data_util.resample_pcd duplicates points): ties must go to the first index
    base = cloud(rng, 2, 200)
    x2 = np.concatenate([base, base, base[:, :37]], axis=1)  # every point appears 2-3 times
    x1 = cloud(rng, 2, 333)
    x1[:, :50] = base[:, :50]  # exact hits, distance 0
    want = port.nn_distance(x1, x2, fused=True)
    got = _run(cuda, x1, x2)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert (got[1][:, :50] == np.arange(50)).all()


def test_nn_distance_integer_grid_many_ties(cuda):
    # points on a coarse integer grid: many exactly equal distances
    r = np.random.default_rng(7)
    x1 = r.integers(-3, 4, size=(3, 700, 3)).astype(np.float32)
    x2 = r.integers(-3, 4, size=(3, 1500, 3)).astype(np.float32)
    want = port.nn_distance(x1, x2, fused=True)
    got = _run(cuda, x1, x2)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_nn_distance_large_property(cuda):
    """BASELINE config 2 shape (B=32, 2048 vs 16384): too big for the scalar oracle in seconds, so check properties:
    the reported distance is exactly d2(query, candidate[idx]) and no candidate of a random sample is closer or ties with
    a lower index; a sampled subset of queries is checked against the oracle in full."""
    g = torch.Generator(device="cpu").manual_seed(2)
    b, n, m = 32, 2048, 16384
    x1 = (torch.rand((b, n, 3), generator=g) - 0.5)
    x2 = (torch.rand((b, m, 3), generator=g) - 0.5)
    d1, i1, d2, i2 = _run(cuda, x1.numpy(), x2.numpy())
    assert i1.min() >= 0 and i1.max() < m and i2.min() >= 0 and i2.max() < n
    # oracle on a slice: clouds 0 and 31, first 64 queries of each direction
    for c in (0, 31):
        w = port.nn_distance(x1[c:c + 1, :64].numpy(), x2[c:c + 1].numpy(), fused=True)
        assert np.array_equal(d1[c, :64], w[0][0]) and np.array_equal(i1[c, :64], w[1][0])
        w = port.nn_distance(x2[c:c + 1, :64].numpy(), x1[c:c + 1].numpy(), fused=True)
        assert np.array_equal(d2[c, :64], w[0][0]) and np.array_equal(i2[c, :64], w[1][0])
    # min property over everything, in float64 with a tolerance far below the gaps (exactness is covered above)
    x1d, x2d = x1.double().to(cuda), x2.double().to(cuda)
    full = torch.cdist(x1d, x2d) ** 2
    assert np.allclose(full.min(dim=2).values.cpu().numpy(), d1, rtol=1e-5, atol=1e-9)
    assert np.allclose(full.min(dim=1).values.cpu().numpy(), d2, rtol=1e-5, atol=1e-9)


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
@pytest.mark.parametrize("b,n,m", [(32, 2048, 16384), (4, 16384, 16384), (32, 16384, 16384)])
def test_nn_distance_full_shapes_vs_reference_cuda_kernel(cuda, b, n, m):
    """BASELINE config 2 (B=32, 2048 vs 16384), the north-star shape (B=32, 16384^2) and its 8-GPU shard (B=4): ALL four
    outputs bit-identical with the reference's own CUDA kernel (tf_ops/CD/tf_nndistance_g.cu == pc_distance/tf_nndistance.cu,
    recompiled unchanged for sm_100a) -- every distance and every index of every query, in both directions; then the gradient
    for those indices against the reference's NnDistanceGrad kernel (float atomics: order-dependent last bits -> 1e-5)."""
    from rfnet_b200 import ops, tf_nndistance
    g = torch.Generator(device="cpu").manual_seed(20 + b + n)
    x1 = (torch.rand((b, n, 3), generator=g) - 0.5).to(cuda)
    x2 = (torch.rand((b, m, 3), generator=g) - 0.5).to(cuda)
    want = ref.run_gpu("NnDistance", [x1, x2], [((b, n), torch.float32), ((b, n), torch.int32), ((b, m), torch.float32), ((b, m), torch.int32)])
    got = tf_nndistance.nn_distance(x1, x2)
    for gt, w in zip(got, want):
        assert torch.equal(gt, w)
    gd1 = torch.rand((b, n), generator=g).to(cuda)
    gd2 = torch.rand((b, m), generator=g).to(cuda)
    w1, w2 = ref.run_gpu("NnDistanceGrad", [x1, x2, gd1, want[1], gd2, want[3]], [((b, n, 3), torch.float32), ((b, m, 3), torch.float32)])
    for det in (False, True):
        g1, g2 = ops.nn_distance_grad_op(x1, x2, gd1, got[1], gd2, got[3], det)
        assert float((g1 - w1).abs().max()) <= 1e-5 * float(w1.abs().max())
        assert float((g2 - w2).abs().max()) <= 1e-5 * float(w2.abs().max())


def test_nn_distance_grad_many_queries_share_one_neighbour(cuda, rng):
    """Segments of every length class of the CSR sort (<=32 shuffle, <=1024 shared memory, <=8192 odd-even): many points of
    xyz1 have the same nearest neighbour in a tiny xyz2."""
    from rfnet_b200 import ops
    b, n, m = 2, 6000, 3
    x1 = cloud(rng, b, n)
    x2 = np.stack([np.array([[0.4, 0.4, 0.4], [-0.45, -0.45, -0.45], [0.0, 0.49, -0.49]], np.float32)] * b)
    _, i1, _, i2 = port.nn_distance(x1, x2)
    assert np.bincount(i1[0]).max() > 1024
    g1 = rng.standard_normal((b, n)).astype(np.float32)
    g2 = rng.standard_normal((b, m)).astype(np.float32)
    want = port.nn_distance_grad(x1, x2, g1, i1, g2, i2)
    t = lambda a: torch.from_numpy(a).to(cuda)
    got = ops.nn_distance_grad_op(t(x1), t(x2), t(g1), t(i1), t(g2), t(i2), True)
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)


def test_nn_distance_grad_atomic_path_without_workspace(cuda, rng):
    """C ABI with workspace == NULL: the reference GPU formulation (float reductions); equal within rounding."""
    from rfnet_b200 import ops
    b, n, m = 2, 700, 300
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    _, i1, _, i2 = port.nn_distance(x1, x2)
    g1 = rng.standard_normal((b, n)).astype(np.float32)
    g2 = rng.standard_normal((b, m)).astype(np.float32)
    want = port.nn_distance_grad(x1, x2, g1, i1, g2, i2)
    t = lambda a: torch.from_numpy(a).to(cuda)
    o1, o2 = torch.empty((b, n, 3), device=cuda), torch.empty((b, m, 3), device=cuda)
    ops.raw_nn_distance_grad(t(x1), t(x2), t(g1), t(i1), t(g2), t(i2), o1, o2, None)
    torch.cuda.synchronize()
    for g, w in zip((o1, o2), want):
        assert np.allclose(g.cpu().numpy(), w, rtol=1e-5, atol=1e-5 * np.abs(w).max())


@pytest.mark.parametrize("b,n,m", [(2, 5, 3), (2, 300, 257), (3, 1024, 2050)])
def test_nn_distance_grad(cuda, rng, b, n, m):
    from rfnet_b200 import ops
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    _, i1, _, i2 = port.nn_distance(x1, x2)
    g1 = rng.standard_normal((b, n)).astype(np.float32)
    g2 = rng.standard_normal((b, m)).astype(np.float32)
    want = port.nn_distance_grad(x1, x2, g1, i1, g2, i2)
    t = lambda a: torch.from_numpy(a).to(cuda)
    got = ops.nn_distance_grad_op(t(x1), t(x2), t(g1), t(i1), t(g2), t(i2), True)
    for g, w in zip(got, want):
        # the atomic-free scatter sums in the reference's sequential order (tf_nndistance.cpp:126-163): bit-exact
        assert np.array_equal(g.cpu().numpy(), w)
    again = ops.nn_distance_grad_op(t(x1), t(x2), t(g1), t(i1), t(g2), t(i2), True)
    assert all(torch.equal(a, c) for a, c in zip(got, again))      # and reproducible run to run
    fast = ops.nn_distance_grad_op(t(x1), t(x2), t(g1), t(i1), t(g2), t(i2))   # default: float reductions, the reference GPU formulation
    for g, w in zip(fast, want):
        assert np.allclose(g.cpu().numpy(), w, rtol=1e-5, atol=1e-5 * np.abs(w).max())


def test_nn_distance_autograd_matches_reference_formula(cuda, rng):
    from rfnet_b200 import losses
    x1 = torch.from_numpy(cloud(rng, 2, 400)).to(cuda).requires_grad_(True)
    x2 = torch.from_numpy(cloud(rng, 2, 300)).to(cuda).requires_grad_(True)
    loss, _ = losses.chamfer_big(x1, x2)
    loss.backward()
    # same loss built from oracle indices with plain torch ops
    a, c = x1.detach().cpu().double().requires_grad_(True), x2.detach().cpu().double().requires_grad_(True)
    _, i1, _, i2 = port.nn_distance(a.detach().float().numpy(), c.detach().float().numpy())
    i1, i2 = torch.from_numpy(i1).long(), torch.from_numpy(i2).long()
    d1 = ((a - torch.gather(c, 1, i1[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    d2 = ((c - torch.gather(a, 1, i2[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    ref = (d1.sqrt().mean() + d2.sqrt().mean()) / 2
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert np.allclose(x1.grad.cpu().numpy(), a.grad.numpy(), rtol=1e-4, atol=1e-7)
    assert np.allclose(x2.grad.cpu().numpy(), c.grad.numpy(), rtol=1e-4, atol=1e-7)


def test_nn_distance_empty_and_errors(cuda):
    from rfnet_b200 import tf_nndistance
    z = torch.zeros((0, 5, 3), device=cuda)
    out = tf_nndistance.nn_distance(z, torch.zeros((0, 7, 3), device=cuda))
    assert [tuple(o.shape) for o in out] == [(0, 5), (0, 5), (0, 7), (0, 7)]
    with pytest.raises(ValueError, match="3d point set xyz1"):
        tf_nndistance.nn_distance(torch.zeros((1, 5, 2), device=cuda), torch.zeros((1, 7, 3), device=cuda))
    with pytest.raises(ValueError, match="same batch size"):
        tf_nndistance.nn_distance(torch.zeros((1, 5, 3), device=cuda), torch.zeros((2, 7, 3), device=cuda))


def test_chamfer_partial_sums(cuda, rng):
    from rfnet_b200 import losses, ops
    d1 = torch.from_numpy(rng.random((7, 333), dtype=np.float32)).to(cuda)
    d2 = torch.from_numpy(rng.random((7, 1999), dtype=np.float32)).to(cuda)
    got = ops.chamfer_partial_sums_op(d1, d2).cpu().numpy()
    want = losses.chamfer_partial_sums(d1, d2).cpu().numpy()
    assert got[1] == 7 * 333 and got[3] == 7 * 1999
    assert np.allclose(got, want, rtol=1e-5)
    assert np.allclose(got[0], np.sqrt(d1.cpu().numpy().astype(np.float64)).sum(), rtol=1e-5)
    # deterministic: fixed summation order
    assert np.array_equal(got, ops.chamfer_partial_sums_op(d1, d2).cpu().numpy())


def test_host_pipeline_matches_direct_ops(cuda, rng):
    """rfnet_b200.host.ChamferHostPipeline: pinned host buffers in, pinned host buffers out, overlapped copies."""
    from rfnet_b200.host import ChamferHostPipeline
    b, n, m = 3, 700, 1500
    pipe = ChamferHostPipeline(b, n, m, cuda, depth=2, deterministic=True)
    batches = [(torch.from_numpy(cloud(rng, b, n)).pin_memory(), torch.from_numpy(cloud(rng, b, m)).pin_memory()) for _ in range(5)]
    for h1, h2 in batches:
        slot = pipe.submit(h1, h2)
        out = pipe.wait(slot)
        want = port.nn_distance(h1.numpy(), h2.numpy())
        assert np.array_equal(out["dist1"].numpy(), want[0]) and np.array_equal(out["idx1"].numpy(), want[1])
        assert np.array_equal(out["dist2"].numpy(), want[2]) and np.array_equal(out["idx2"].numpy(), want[3])
        g1, g2 = port.nn_distance_grad(h1.numpy(), h2.numpy(), np.full((b, n), 0.5 / (b * n), np.float32), want[1],
                                       np.full((b, m), 0.5 / (b * m), np.float32), want[3])
        assert np.allclose(out["grad1"].numpy(), g1, rtol=1e-5, atol=1e-9) and np.allclose(out["grad2"].numpy(), g2, rtol=1e-5, atol=1e-9)
        assert np.allclose(out["sums"].numpy()[0], np.sqrt(want[0].astype(np.float64)).sum(), rtol=1e-5)
    pipe.drain()


@pytest.mark.parametrize("b,n,m", [(2, 300, 257), (3, 2048, 5000), (32, 2048, 16384), (4, 16384, 16384), (1, 1, 7)])
def test_chamfer_step_equals_separate_calls(cuda, b, n, m):
    """rfnet_chamfer_step (search + one epilogue + one reduction) against nn_distance, nn_distance_grad and
    chamfer_partial_sums called one after the other: dist / idx bit-identical, gradients and sums to float rounding."""
    from rfnet_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(b * 7 + n)
    x1 = (torch.rand((b, n, 3), generator=g) - 0.5).to(cuda)
    x2 = (torch.rand((b, m, 3), generator=g) - 0.5).to(cuda)
    gd1, gd2 = torch.rand((b, n), generator=g).to(cuda), torch.rand((b, m), generator=g).to(cuda)
    d1, i1, d2, i2 = ops.nn_distance_op(x1, x2)
    w1, w2 = ops.nn_distance_grad_op(x1, x2, gd1, i1, gd2, i2, True)          # deterministic reference-order gradient
    wsum = ops.chamfer_partial_sums_op(d1, d2)
    f32, i32 = torch.float32, torch.int32
    o = dict(d1=torch.empty((b, n), dtype=f32, device=cuda), i1=torch.empty((b, n), dtype=i32, device=cuda),
             d2=torch.empty((b, m), dtype=f32, device=cuda), i2=torch.empty((b, m), dtype=i32, device=cuda),
             g1=torch.empty((b, n, 3), dtype=f32, device=cuda), g2=torch.empty((b, m, 3), dtype=f32, device=cuda), s=torch.empty(4, device=cuda))
    ws = torch.empty(ops.nn_distance_workspace_bytes(b, n, m), dtype=torch.uint8, device=cuda)
    for _ in range(2):   # twice: the call must re-initialise everything it accumulates into
        ops.raw_chamfer_step(x1, x2, gd1, gd2, o["d1"], o["i1"], o["d2"], o["i2"], o["g1"], o["g2"], o["s"], ws)
    assert torch.equal(o["d1"], d1) and torch.equal(o["i1"], i1) and torch.equal(o["d2"], d2) and torch.equal(o["i2"], i2)
    assert float((o["g1"] - w1).abs().max()) <= 1e-5 * float(w1.abs().max()) and float((o["g2"] - w2).abs().max()) <= 1e-5 * float(w2.abs().max())
    assert torch.allclose(o["s"], wsum, rtol=1e-5)


@pytest.mark.parametrize("b,nr,nn", [(2, 300, 512), (4, 16384, 2048), (1, 7, 1)])
def test_merge_layer_fused_forward_and_gradients(cuda, b, nr, nn):
    """rfnet::merge_layer (one directed search + one epilogue) against the reference's own formulation (vv_recon.py:132-139:
    NnDistance, GroupPoint, framework arithmetic) built from the drop-in ops: same points, and the same gradients w.r.t. the
    raw cloud, the new points and the trained decfactor."""
    from rfnet_b200 import losses, tf_grouping, tf_nndistance
    g = torch.Generator(device="cpu").manual_seed(nr + nn)
    raw0 = (torch.rand((b, nr, 3), generator=g) - 0.5).to(cuda)
    new0 = (torch.rand((b, nn, 3), generator=g) - 0.5).to(cuda)
    wgt = torch.randn((b, nn, 3), generator=g).to(cuda)
    res = []
    for fused in (True, False):
        raw, new = raw0.clone().requires_grad_(True), new0.clone().requires_grad_(True)
        dec = torch.tensor([0.07], device=cuda, requires_grad=True)
        if fused:
            out = losses.merge_layer(raw, new, dec)
        else:
            _, _, _, idx2 = tf_nndistance.nn_distance(raw, new)
            grouped = tf_grouping.group_point(raw, idx2.unsqueeze(-1))
            diff = grouped - new.unsqueeze(2)
            ratio = torch.exp(-(diff * diff).sum(-1, keepdim=True) / (1e-8 + dec * dec))
            out = new + (ratio * diff).sum(2)
        (out * wgt).sum().backward()
        res.append((out.detach(), raw.grad, new.grad, dec.grad))
    for a, w in zip(res[0], res[1]):
        assert torch.allclose(a, w, rtol=2e-5, atol=1e-6), float((a - w).abs().max())
