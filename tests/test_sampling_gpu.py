"""GPU parity of farthest_point_sample / gather_point (+grad): indices bit-exact against the oracle and the reference CUDA kernel."""
import numpy as np
import pytest
import torch

from conftest import cloud
from oracle import port, ref

pytestmark = pytest.mark.gpu


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# (b, n, m): up to 16384 points the pruned one-CTA kernel (1/2/4/8/16/32 register slots per lane), above it the cluster
# kernel; m > n (repeats), n < 512, tiny
SHAPES = [(1, 1, 1), (2, 7, 3), (3, 100, 150), (2, 512, 64), (2, 513, 600), (2, 3000, 32), (2, 3000, 700), (4, 4097, 128), (2, 9000, 520), (2, 16384, 256), (2, 16384, 255),
          (1, 20000, 64), (40, 1024, 64),
          (1, 40000, 24)]   # the last one exceeds 8 CTAs x 512 threads x 8 points: generic one-CTA kernel with global scratch


@pytest.mark.parametrize("b,n,m", SHAPES)
def test_fps_indices_bit_exact(cuda, rng, b, n, m):
    from rfnet_b200 import tf_sampling
    x = cloud(rng, b, n)
    want = port.farthest_point_sample(m, x)
    got = tf_sampling.farthest_point_sample(m, t(x, cuda)).cpu().numpy()
    assert got.dtype == np.int32 and np.array_equal(got, want), "first mismatch at %s" % (np.argwhere(got != want)[:1],)


def test_fps_duplicated_points_tie_rule(cuda, rng):
    """resample_pcd (data_util.py:8-13) duplicates points; equal distances must resolve like the reference's 512-thread
    block: lowest (k mod 512), then lowest k."""
    from rfnet_b200 import tf_sampling
    base = cloud(rng, 2, 700)
    x = np.concatenate([base, base, base[:, :300]], axis=1)   # n = 1700, every point 2-3 times
    want = port.farthest_point_sample(900, x)
    got = tf_sampling.farthest_point_sample(900, t(x, cuda)).cpu().numpy()
    assert np.array_equal(got, want)
    grid = np.random.default_rng(3).integers(-2, 3, size=(2, 1500, 3)).astype(np.float32)  # massive ties
    want = port.farthest_point_sample(200, grid)
    got = tf_sampling.farthest_point_sample(200, t(grid, cuda)).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("b,n,m,kind", [(3, 16384, 2048, "cube"), (2, 16384, 1500, "sphere"), (4, 3000, 1024, "grid"), (2, 700, 700, "dup"), (2, 100, 600, "cube")])
def test_fps_pruned_equals_cluster_kernel(cuda, b, n, m, kind):
    """The pruned kernel (Morton clusters whose box is out of reach of a pick are skipped) against the cluster kernel that
    updates every point at every pick (RFNET_FPS_NO_PRUNE=1): identical indices, on volumes, surfaces, lattices (massive
    ties) and duplicated points."""
    from rfnet_b200 import tf_sampling
    g = torch.Generator(device="cpu").manual_seed(50 + n + m)
    x = torch.rand((b, n, 3), generator=g) - 0.5
    if kind == "sphere":
        x = x / x.norm(dim=-1, keepdim=True)
    elif kind == "grid":
        x = torch.round(x * 6)
    elif kind == "dup":
        x[:, n // 2:] = x[:, : n - n // 2]
    x = x.to(cuda)
    pruned = tf_sampling.farthest_point_sample(m, x)
    full = torch.ops.rfnet.farthest_point_sample(x, m, False)    # no workspace: the cluster kernel
    assert torch.equal(pruned, full)


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
@pytest.mark.parametrize("b,n,m", [(32, 3000, 32), (4, 16384, 2048), (32, 5000, 500)])
def test_fps_vs_reference_cuda_kernel(cuda, rng, b, n, m):
    from rfnet_b200 import tf_sampling
    x = t(cloud(rng, b, n), cuda)
    (want,) = ref.run_gpu("FarthestPointSample", [x], [((b, m), torch.int32)], attrs={"npoint": m})
    got = tf_sampling.farthest_point_sample(m, x)
    assert torch.equal(got, want)


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
def test_fps_config4_full_shape_vs_reference_cuda_kernel(cuda):
    """BASELINE config 4 at full size (B=32, 16384 -> 2048): every index of every cloud equals the reference CUDA kernel's
    (tf_sampling_g.cu:105-170), for the pruned kernel and for the cluster kernel."""
    from rfnet_b200 import tf_sampling
    g = torch.Generator(device="cpu").manual_seed(9)
    x = (torch.rand((32, 16384, 3), generator=g) - 0.5).to(cuda)
    (want,) = ref.run_gpu("FarthestPointSample", [x], [((32, 2048), torch.int32)], attrs={"npoint": 2048})
    assert torch.equal(tf_sampling.farthest_point_sample(2048, x), want)
    assert torch.equal(torch.ops.rfnet.farthest_point_sample(x, 2048, False), want)


def test_fps_full_size_properties(cuda):
    """BASELINE config 4 (B=32, 16384 -> 2048): indices valid and unique per cloud, first index 0, and the sequence is
    the greedy one: each pick is at maximal distance from the previous picks (checked in float64 on one cloud)."""
    from rfnet_b200 import tf_sampling
    g = torch.Generator(device="cpu").manual_seed(9)
    x = (torch.rand((32, 16384, 3), generator=g) - 0.5).to(cuda)
    idx = tf_sampling.farthest_point_sample(2048, x)
    assert idx.shape == (32, 2048) and int(idx.min()) >= 0 and int(idx.max()) < 16384
    assert bool((idx[:, 0] == 0).all())
    for c in range(32):
        assert idx[c].unique().numel() == 2048
    p = x[0].double()
    sel = idx[0].long()
    mind = torch.full((16384,), float("inf"), device=cuda, dtype=torch.float64)
    for j in range(1, 200):
        mind = torch.minimum(mind, ((p - p[sel[j - 1]]) ** 2).sum(-1))
        assert mind[sel[j]] >= mind.max() * (1 - 1e-6)


def test_gather_point_and_grad(cuda, rng):
    from rfnet_b200 import ops, tf_sampling
    b, n, m = 3, 500, 777
    x = cloud(rng, b, n)
    idx = rng.integers(0, n, size=(b, m)).astype(np.int32)   # with repeats
    got = tf_sampling.gather_point(t(x, cuda), t(idx, cuda)).cpu().numpy()
    assert np.array_equal(got, port.gather_point(x, idx))
    og = rng.standard_normal((b, m, 3)).astype(np.float32)
    want = port.gather_point_grad(x, idx, og)
    gg = ops.gather_point_grad_op(t(x, cuda), t(idx, cuda), t(og, cuda)).cpu().numpy()
    assert np.array_equal(gg, want)   # atomic-free, ascending-row order == the sequential oracle
    # autograd path
    xt = t(x, cuda).requires_grad_(True)
    tf_sampling.gather_point(xt, t(idx, cuda)).backward(t(og, cuda))
    assert np.allclose(xt.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-6)


def test_fps_errors(cuda):
    from rfnet_b200 import tf_sampling
    with pytest.raises(ValueError, match="positive npoint"):
        tf_sampling.farthest_point_sample(0, torch.zeros((1, 5, 3), device=cuda))
    with pytest.raises(ValueError, match="inp shape"):
        tf_sampling.farthest_point_sample(2, torch.zeros((1, 5, 2), device=cuda))
