"""CPU tier: the Python drop-in surface -- names, argument orders, shape/dtype inference (meta tensors), error messages,
and the absence of any CPU fallback."""
import inspect

import pytest
import torch

from rfnet_b200 import losses, ops, tf_approxmatch, tf_auctionmatch, tf_grouping, tf_interpolate, tf_nndistance, tf_sampling


def meta(*shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device="meta")


def test_signatures_match_reference_wrappers():
    # tf_ops/CD/tf_nndistance.py:9, pc_distance/tf_approxmatch.py:10,27, tf_sampling.py:29,48, tf_grouping.py:8,33,48, tf_interpolate.py:8,19
    assert list(inspect.signature(tf_nndistance.nn_distance).parameters)[:2] == ["xyz1", "xyz2"]
    # approx_match keeps the reference's two positional arguments; exact is an optional keyword (the bit-exact parity mode)
    assert list(inspect.signature(tf_approxmatch.approx_match).parameters)[:2] == ["xyz1", "xyz2"]
    assert inspect.signature(tf_approxmatch.approx_match).parameters["exact"].default is False
    assert list(inspect.signature(tf_approxmatch.match_cost).parameters) == ["xyz1", "xyz2", "match"]
    assert list(inspect.signature(tf_sampling.farthest_point_sample).parameters) == ["npoint", "inp"]
    assert list(inspect.signature(tf_sampling.gather_point).parameters) == ["inp", "idx"]
    assert list(inspect.signature(tf_grouping.query_ball_point).parameters) == ["radius", "nsample", "xyz1", "xyz2"]
    assert list(inspect.signature(tf_grouping.group_point).parameters) == ["points", "idx"]
    assert list(inspect.signature(tf_grouping.knn_point).parameters) == ["k", "xyz1", "xyz2"]
    assert list(inspect.signature(tf_interpolate.three_nn).parameters) == ["xyz1", "xyz2"]
    assert list(inspect.signature(tf_interpolate.three_interpolate).parameters) == ["points", "idx", "weight"]
    # tf_grouping.py:22 (select_top_k), tf_ops/emd/tf_auctionmatch.py:11
    assert list(inspect.signature(tf_grouping.select_top_k).parameters) == ["k", "dist"]
    assert list(inspect.signature(tf_auctionmatch.auction_match).parameters) == ["xyz1", "xyz2"]


def test_output_shapes_and_dtypes_via_meta_tensors():
    b, n, m, c, ns = 4, 50, 70, 16, 8
    d1, i1, d2, i2 = torch.ops.rfnet.nn_distance(meta(b, n, 3), meta(b, m, 3))
    assert (d1.shape, i1.shape, d2.shape, i2.shape) == ((b, n), (b, n), (b, m), (b, m))
    assert (d1.dtype, i1.dtype, d2.dtype, i2.dtype) == (torch.float32, torch.int32, torch.float32, torch.int32)
    g1, g2 = torch.ops.rfnet.nn_distance_grad(meta(b, n, 3), meta(b, m, 3), meta(b, n), meta(b, n, dtype=torch.int32), meta(b, m), meta(b, m, dtype=torch.int32))
    assert g1.shape == (b, n, 3) and g2.shape == (b, m, 3)
    assert torch.ops.rfnet.approx_match(meta(b, n, 3), meta(b, m, 3)).shape == (b, m, n)   # (batch, #query, #dataset)
    assert torch.ops.rfnet.match_cost(meta(b, n, 3), meta(b, m, 3), meta(b, m, n)).shape == (b,)
    g1, g2 = torch.ops.rfnet.match_cost_grad(meta(b, n, 3), meta(b, m, 3), meta(b, m, n))
    assert g1.shape == (b, n, 3) and g2.shape == (b, m, 3)
    idx = torch.ops.rfnet.farthest_point_sample(meta(b, n, 3), 9)
    assert idx.shape == (b, 9) and idx.dtype == torch.int32
    assert torch.ops.rfnet.gather_point(meta(b, n, 3), meta(b, 9, dtype=torch.int32)).shape == (b, 9, 3)
    assert torch.ops.rfnet.gather_point_grad(meta(b, n, 3), meta(b, 9, dtype=torch.int32), meta(b, 9, 3)).shape == (b, n, 3)
    qi, qc = torch.ops.rfnet.query_ball_point(meta(b, n, 3), meta(b, m, 3), meta(1), ns)
    assert qi.shape == (b, m, ns) and qc.shape == (b, m) and qi.dtype == qc.dtype == torch.int32
    assert torch.ops.rfnet.group_point(meta(b, n, c), meta(b, m, ns, dtype=torch.int32)).shape == (b, m, ns, c)
    assert torch.ops.rfnet.group_point_grad(meta(b, n, c), meta(b, m, ns, dtype=torch.int32), meta(b, m, ns, c)).shape == (b, n, c)
    dist, i3 = torch.ops.rfnet.three_nn(meta(b, n, 3), meta(b, m, 3))
    assert dist.shape == i3.shape == (b, n, 3) and i3.dtype == torch.int32
    assert torch.ops.rfnet.three_interpolate(meta(b, m, c), meta(b, n, 3, dtype=torch.int32), meta(b, n, 3)).shape == (b, n, c)
    assert torch.ops.rfnet.three_interpolate_grad(meta(b, m, c), meta(b, n, 3, dtype=torch.int32), meta(b, n, 3), meta(b, n, c)).shape == (b, m, c)
    outi, out = torch.ops.rfnet.selection_sort(meta(b, m, n), 5)
    assert outi.shape == out.shape == (b, m, n) and outi.dtype == torch.int32 and out.dtype == torch.float32
    ml, mr = torch.ops.rfnet.auction_match(meta(b, n, 3), meta(b, n, 3))
    assert ml.shape == mr.shape == (b, n) and ml.dtype == mr.dtype == torch.int32
    cost, kept = torch.ops.rfnet.emd_cost(meta(b, n, 3), meta(b, m, 3), True)
    assert cost.shape == (b,) and kept.shape == (b, m, n)
    cost, kept = torch.ops.rfnet.emd_cost(meta(b, n, 3), meta(b, m, 3), False)
    assert cost.shape == (b,) and kept.numel() == 0
    cost, g1, g2 = torch.ops.rfnet.emd_cost_grad(meta(b, n, 3), meta(b, m, 3))
    assert cost.shape == (b,) and g1.shape == (b, n, 3) and g2.shape == (b, m, 3)
    assert torch.ops.rfnet.approx_match(meta(b, n, 3), meta(b, m, 3), 1).shape == (b, m, n)
    val, ki = torch.ops.rfnet.knn_point(meta(b, n, 3), meta(b, m, 3), 4)
    assert val.shape == ki.shape == (b, m, 4) and ki.dtype == torch.int32


def test_no_cpu_fallback():
    """CPU tensors must be rejected, not silently computed: there is no kernel registered for the CPU backend."""
    x, y = torch.zeros(1, 4, 3), torch.zeros(1, 5, 3)
    for call in (lambda: tf_nndistance.nn_distance(x, y), lambda: tf_approxmatch.approx_match(x, y),
                 lambda: tf_sampling.farthest_point_sample(2, x), lambda: tf_interpolate.three_nn(x, y),
                 lambda: tf_grouping.group_point(x, torch.zeros(1, 2, 2, dtype=torch.int32)),
                 lambda: tf_auctionmatch.auction_match(x, x), lambda: tf_grouping.select_top_k(2, torch.zeros(1, 3, 4)),
                 lambda: tf_approxmatch.emd_cost(x, y), lambda: losses.emd_func(x, x)):
        with pytest.raises((NotImplementedError, RuntimeError, ValueError)):
            call()


def test_product_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "rfnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "librfnet_oracle" not in text and "libref_" not in text, f


def test_knn_point_has_no_framework_fallback():
    """The reference's knn_point is framework code (tf_grouping.py:64-73).  Ours is a CUDA kernel for 3-d points and k <= 32 and
    NOTHING else: CPU tensors, feature vectors and larger k are errors, not a silent eager path."""
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.rand(2, 30, 3, generator=g), torch.rand(2, 7, 3, generator=g)
    with pytest.raises((NotImplementedError, RuntimeError)):
        tf_grouping.knn_point(4, x1, x2)                                   # CPU tensors: no kernel registered
    with pytest.raises(ValueError, match="3-d points"):
        tf_grouping.knn_point(4, torch.rand(2, 30, 5), torch.rand(2, 7, 5))
    with pytest.raises(ValueError, match="k <= min"):
        tf_grouping.knn_point(33, torch.rand(2, 64, 3), torch.rand(2, 7, 3))


def test_shard_bounds_cover_batch_exactly():
    for B in (1, 7, 32, 64):
        for W in (1, 2, 3, 4, 8):
            spans = [losses.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
