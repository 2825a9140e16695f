"""GPU parity of auction_match (tf_ops/emd) and select_top_k (tf_ops/grouping SelectionSort) against the CPU oracle and
the reference's own CUDA kernels.  Both are integer-valued results: bit-exact."""
import numpy as np
import pytest
import torch

from conftest import cloud
from oracle import port, ref

pytestmark = pytest.mark.gpu


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# 1-, 2-, 4-, 8- and 16-objects-per-thread instantiations, ragged and exact multiples of the 512-thread block
@pytest.mark.parametrize("b,n", [(3, 1), (2, 2), (4, 37), (2, 512), (2, 700), (1, 1024), (1, 1500), (1, 2048), (1, 3000)])
def test_auction_match_vs_oracle(cuda, rng, b, n):
    from rfnet_b200 import tf_auctionmatch
    x1, x2 = cloud(rng, b, n), cloud(rng, b, n)
    wl, wr = port.auction_match(x1, x2)
    ml, mr = tf_auctionmatch.auction_match(t(x1, cuda), t(x2, cuda))
    assert ml.dtype == torch.int32 and tuple(ml.shape) == (b, n) and tuple(mr.shape) == (b, n)
    assert np.array_equal(ml.cpu().numpy(), wl) and np.array_equal(mr.cpu().numpy(), wr)


@pytest.mark.parametrize("n", [4096, 5000, 8192])
def test_auction_match_large_properties(cuda, n):
    """The reference's maximum (4096) and beyond it: a perfect matching, matchl / matchr mutually inverse, and a cost no
    worse than matching every point to its nearest neighbour + n * tolerance_max would allow (loose sanity bound)."""
    from rfnet_b200 import tf_auctionmatch
    g = torch.Generator(device="cpu").manual_seed(n)
    x1 = (torch.rand((2, n, 3), generator=g) - 0.5).to(cuda)
    x2 = (torch.rand((2, n, 3), generator=g) - 0.5).to(cuda)
    ml, mr = tf_auctionmatch.auction_match(x1, x2)
    ar = torch.arange(n, device=cuda, dtype=torch.int32)
    for i in range(2):
        assert torch.equal(torch.sort(ml[i]).values, ar)
        assert torch.equal(mr[i][ml[i].long()], ar)
    matched = torch.gather(x2, 1, ml.long()[..., None].expand(-1, -1, 3))
    cost = (x1 - matched).norm(dim=-1).sum(dim=1)
    lower = torch.cdist(x1, x2).min(dim=2).values.sum(dim=1)     # every point to its nearest: a lower bound of any matching
    assert bool((cost >= lower - 1e-3).all()) and bool((cost <= lower + n * 1.0).all())
    if n == 4096:
        wl, _ = port.auction_match(x1[:1].cpu().numpy(), x2[:1].cpu().numpy())
        assert np.array_equal(ml[:1].cpu().numpy(), wl)


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
@pytest.mark.parametrize("b,n", [(2, 300), (1, 1024), (1, 2048), (1, 4096)])
def test_auction_match_vs_reference_cuda_kernel(cuda, rng, b, n):
    """The reference's AuctionMatchKernel (tf_auctionmatch_g.cu, recompiled for sm_100a) on the same clouds: identical
    assignment.  n = 4096 / 2048 / 1024 / 300 take its four differently unrolled bid loops (_g.cu:59,155,193,216)."""
    from rfnet_b200 import tf_auctionmatch
    x1, x2 = t(cloud(rng, b, n), cuda), t(cloud(rng, b, n), cuda)
    wl, wr = ref.run_gpu("AuctionMatch", [x1, x2], [((b, n), torch.int32), ((b, n), torch.int32)])
    ml, mr = tf_auctionmatch.auction_match(x1, x2)
    assert torch.equal(ml, wl) and torch.equal(mr, wr)


def test_auction_match_errors_and_emd_func(cuda, rng):
    from rfnet_b200 import losses, tf_auctionmatch
    with pytest.raises(ValueError, match="shape must match"):
        tf_auctionmatch.auction_match(torch.zeros((1, 8, 3), device=cuda), torch.zeros((1, 9, 3), device=cuda))
    with pytest.raises(ValueError, match="at most"):
        tf_auctionmatch.auction_match(torch.zeros((1, 9000, 3), device=cuda), torch.zeros((1, 9000, 3), device=cuda))
    ml, mr = tf_auctionmatch.auction_match(torch.zeros((2, 0, 3), device=cuda), torch.zeros((2, 0, 3), device=cuda))
    assert tuple(ml.shape) == (2, 0)
    # emd_func (vv_recon.py:365-380) forward + gradient w.r.t. the prediction
    b, n = 2, 400
    p, g = cloud(rng, b, n), cloud(rng, b, n)
    pred = t(p, cuda).requires_grad_(True)
    loss = losses.emd_func(pred, t(g, cuda))
    loss.backward()
    wl, _ = port.auction_match(p, g)
    matched = np.take_along_axis(g, wl[..., None].astype(np.int64), axis=1)
    dist = np.sqrt(((p - matched) ** 2).sum(-1)).mean(-1)
    radius = np.sqrt(((p - p.mean(axis=1, keepdims=True)) ** 2).sum(-1).max(-1))
    assert abs(loss.item() - float((dist / radius).mean())) <= 1e-5 * abs(float((dist / radius).mean()))
    assert pred.grad is not None and bool(torch.isfinite(pred.grad).all()) and float(pred.grad.abs().sum()) > 0


@pytest.mark.parametrize("b,m,n,k", [(2, 7, 40, 5), (1, 300, 257, 32), (2, 3, 64, 64), (1, 5, 1, 1), (1, 50, 2048, 16), (1, 3, 6000, 8), (1, 2, 30000, 4)])
def test_select_top_k_vs_oracle(cuda, rng, b, m, n, k):
    """Rows up to 5120 entries are sorted in shared memory, longer ones in place in global memory; both are bit-exact with
    the reference's swap sequence, tail included."""
    from rfnet_b200 import tf_grouping
    d = rng.random((b, m, n), dtype=np.float32)
    if n > 100:
        d[:, :, 60:90] = d[:, :, 10:40]   # equal values: the first of the current arrangement wins, as in the reference
    wi, wo = port.select_top_k(k, d)
    outi, out = tf_grouping.select_top_k(k, t(d, cuda))
    assert outi.dtype == torch.int32
    assert np.array_equal(outi.cpu().numpy(), wi) and np.array_equal(out.cpu().numpy(), wo)


@pytest.mark.parametrize("kind", ["quantised", "constant", "two_values", "sorted", "reversed", "signed_zeros", "nan", "inf"])
@pytest.mark.parametrize("n,k", [(2048, 32), (1000, 33), (999, 128), (4096, 129), (385, 64), (384, 7), (50, 50)])
def test_select_top_k_adversarial_rows(cuda, rng, kind, n, k):
    """The list-based kernel (k <= 128) on rows built to break it: masses of equal values (list overflow -> plain sort of the row), equal
    values straddling the threshold, monotone rows, -0.0 / +0.0, NaN and infinities; k = 129 takes the whole-row kernel.  Bit-exact with the
    oracle's swap sequence, tail included."""
    from rfnet_b200 import tf_grouping
    b, m = 2, 37
    d = rng.random((b, m, n), dtype=np.float32)
    if kind == "quantised":
        d = np.floor(d * 8).astype(np.float32) / 8
    elif kind == "constant":
        d[:] = np.float32(0.25)
        d[0, :, n // 2] = 0.125
    elif kind == "two_values":
        d = (d > 0.97).astype(np.float32)           # ~3 % ones, the rest zeros: every minimum is a tie
    elif kind == "sorted":
        d = np.sort(d, axis=-1)
    elif kind == "reversed":
        d = np.sort(d, axis=-1)[..., ::-1].copy()
    elif kind == "signed_zeros":
        d[..., ::3] = 0.0
        d[..., 1::7] = -0.0
    elif kind == "nan":
        d[..., 3] = np.nan                           # inside [0, k) for k > 3: stays where it is
        d[0, :, min(n - 1, 200)] = np.nan
        d[1, 5, :] = np.nan
    elif kind == "inf":
        d[..., ::5] = np.inf
        d[..., 2::11] = -np.inf
    wi, wo = port.select_top_k(k, d)
    outi, out = tf_grouping.select_top_k(k, t(d, cuda))
    assert np.array_equal(outi.cpu().numpy(), wi)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), wo.view(np.uint32))   # bit patterns: NaN and the sign of zero included


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
def test_select_top_k_vs_reference_cuda_kernel(cuda, rng):
    from rfnet_b200 import tf_grouping
    b, m, n, k = 2, 500, 333, 20
    d = t(rng.random((b, m, n), dtype=np.float32), cuda)
    wi, wo = ref.run_gpu("SelectionSort", [d], [((b, m, n), torch.int32), ((b, m, n), torch.float32)], attrs={"k": k})
    outi, out = tf_grouping.select_top_k(k, d)
    assert torch.equal(outi, wi) and torch.equal(out, wo)
    with pytest.raises(ValueError, match="positive k"):
        tf_grouping.select_top_k(0, d)
    # knn_point's documented alternative (tf_grouping.py:66-68): top-k of the materialised matrix == the knn kernel
    x1, x2 = t(cloud(rng, 1, 400), cuda), t(cloud(rng, 1, 90), cuda)
    dist = ((x1[:, None, :, :] - x2[:, :, None, :]) ** 2).sum(-1)
    si, so = tf_grouping.select_top_k(8, dist)
    val, idx = tf_grouping.knn_point(8, x1, x2)
    assert torch.equal(si[..., :8], idx)
