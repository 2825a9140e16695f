"""GPU tier: the host-buffer entry points of the C ABI (what a CPU-side caller such as the reference's DEVICE_CPU OpKernels
would bind), called through ctypes with plain numpy buffers -- no torch tensors cross the boundary."""
import ctypes

import numpy as np
import pytest

from conftest import cloud
from oracle import port

pytestmark = pytest.mark.gpu


def p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_nn_distance_host_matches_oracle(cuda, rng):
    from rfnet_b200 import _lib
    lib = _lib.load()
    for (b, n, m) in [(2, 300, 257), (3, 2048, 5000)]:
        x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
        d1, i1 = np.empty((b, n), np.float32), np.empty((b, n), np.int32)
        d2, i2 = np.empty((b, m), np.float32), np.empty((b, m), np.int32)
        rc = lib.rfnet_nn_distance_host(0, b, n, p(x1), m, p(x2), p(d1), p(i1), p(d2), p(i2), 0)
        assert rc == 0, lib.rfnet_error_string(rc)
        want = port.nn_distance(x1, x2, fused=True)
        for g, w in zip((d1, i1, d2, i2), want):
            assert np.array_equal(g, w)
        # the unfused flag reproduces the reference's CPU kernel (tf_nndistance.cpp:21-43) bit for bit
        rc = lib.rfnet_nn_distance_host(0, b, n, p(x1), m, p(x2), p(d1), p(i1), p(d2), p(i2), 1)
        assert rc == 0
        for g, w in zip((d1, i1, d2, i2), port.nn_distance(x1, x2, fused=False)):
            assert np.array_equal(g, w)


def test_emd_host_matches_oracle(cuda, rng):
    from rfnet_b200 import _lib
    lib = _lib.load()
    b, n = 2, 200
    x1, x2 = cloud(rng, b, n), cloud(rng, b, n)
    match = np.empty((b, n, n), np.float32)
    cost = np.empty((b,), np.float32)
    rc = lib.rfnet_emd_host(0, b, n, n, p(x1), p(x2), p(match), p(cost))
    assert rc == 0, lib.rfnet_error_string(rc)
    want = port.approx_match(x1, x2)
    assert np.abs(match - want).max() <= port.approx_match_tolerance(x1, x2, want) * want.max()
    assert np.allclose(cost, port.match_cost(x1, x2, want), rtol=1e-4)
    # match pointer may be NULL: only the cost comes back
    cost2 = np.empty((b,), np.float32)
    rc = lib.rfnet_emd_host(0, b, n, n, p(x1), p(x2), ctypes.c_void_p(0), p(cost2))
    assert rc == 0 and np.array_equal(cost, cost2)


def test_bad_arguments_return_error_codes(cuda):
    from rfnet_b200 import _lib
    lib = _lib.load()
    z = ctypes.c_void_p(0)
    assert lib.rfnet_nn_distance(1, 4, z, 4, z, z, z, z, z, z, 0, 0, z) == 1          # cudaErrorInvalidValue: null pointers
    assert lib.rfnet_nn_distance(-1, 4, z, 4, z, z, z, z, z, z, 0, 0, z) == 1
    assert lib.rfnet_query_ball_point(1, 4, 4, z, 0, z, z, z, z, z) == 1               # nsample must be positive (tf_grouping.cpp:75)
    assert lib.rfnet_nn_distance(0, 4, z, 4, z, z, z, z, z, z, 0, 0, z) == 0           # empty batch is a no-op
    assert b"invalid" in lib.rfnet_error_string(1)


def test_sharded_losses_single_rank_equal_plain_losses(cuda, rng):
    """world_size 1: the sharded loss must equal chamfer_big / earth_mover in value and gradient (the multi-rank reduction
    itself is covered on CPU with gloo in tests/test_sharding_gloo.py)."""
    import torch
    from rfnet_b200 import losses
    a = torch.from_numpy(cloud(rng, 3, 300)).to(cuda)
    c = torch.from_numpy(cloud(rng, 3, 300)).to(cuda)
    a1, c1 = a.clone().requires_grad_(True), c.clone().requires_grad_(True)
    a2, c2 = a.clone().requires_grad_(True), c.clone().requires_grad_(True)
    l1, _ = losses.chamfer_big(a1, c1)
    l2, _ = losses.sharded_chamfer_big(a2, c2)
    l1.backward(); l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-6 * abs(l1.item())
    assert torch.allclose(a1.grad, a2.grad, rtol=1e-5, atol=1e-9) and torch.allclose(c1.grad, c2.grad, rtol=1e-5, atol=1e-9)
    a3, c3 = a.clone().requires_grad_(True), c.clone().requires_grad_(True)
    a4, c4 = a.clone().requires_grad_(True), c.clone().requires_grad_(True)
    e1 = losses.earth_mover(a3, c3)
    e2 = losses.sharded_earth_mover(a4, c4)
    e1.backward(); e2.backward()
    assert abs(e1.item() - e2.item()) <= 1e-6 * abs(e1.item())
    assert torch.allclose(a3.grad, a4.grad, rtol=1e-5, atol=1e-9)
