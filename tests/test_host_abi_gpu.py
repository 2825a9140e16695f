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
    rc = lib.rfnet_emd_host(0, b, n, n, p(x1), p(x2), p(match), p(cost), 0)
    assert rc == 0, lib.rfnet_error_string(rc)
    want = port.approx_match(x1, x2)
    assert np.abs(match - want).max() <= port.approx_match_tolerance(x1, x2, want) * want.max()
    assert np.allclose(cost, port.match_cost(x1, x2, want), rtol=1e-4)
    # match pointer may be NULL: only the cost comes back
    cost2 = np.empty((b,), np.float32)
    rc = lib.rfnet_emd_host(0, b, n, n, p(x1), p(x2), ctypes.c_void_p(0), p(cost2), 0)
    assert rc == 0 and np.array_equal(cost, cost2)


def test_bad_arguments_return_error_codes(cuda):
    from rfnet_b200 import _lib
    lib = _lib.load()
    z = ctypes.c_void_p(0)
    assert lib.rfnet_nn_distance(1, 4, z, 4, z, z, z, z, z, z, 0, 0, z) == 1          # cudaErrorInvalidValue: null pointers
    assert lib.rfnet_nn_distance(-1, 4, z, 4, z, z, z, z, z, z, 0, 0, z) == 1
    assert lib.rfnet_query_ball_point(1, 4, 4, z, 0, z, z, z, z, z, 0, z) == 1               # nsample must be positive (tf_grouping.cpp:75)
    assert lib.rfnet_nn_distance(0, 4, z, 4, z, z, z, z, z, z, 0, 0, z) == 0           # empty batch is a no-op
    assert b"invalid" in lib.rfnet_error_string(1)
    # entry points added for tf_ops/emd and select_top_k
    assert lib.rfnet_auction_match(1, 9000, z, z, z, z, z) == 1                          # beyond RFNET_AUCTION_MAX_POINTS
    assert lib.rfnet_auction_match(1, 8, z, z, z, z, z) == 1                             # null pointers
    assert lib.rfnet_auction_match(0, 8, z, z, z, z, z) == 0 and lib.rfnet_auction_match(3, 0, z, z, z, z, z) == 0
    assert lib.rfnet_selection_sort(1, 8, 2, 0, z, z, z, z) == 1                         # k must be positive (tf_grouping.cpp:117)
    assert lib.rfnet_selection_sort(1, 0, 2, 3, z, z, z, z) == 0                         # empty rows: nothing to do
    assert lib.rfnet_emd_cost(1, 8, 8, z, z, z, z, z, 0, 0, z) == 1                      # cost pointer required
    assert lib.rfnet_emd_cost_grad(1, 8, 8, z, z, z, z, z, z, 0, 0, z) == 1
    assert lib.rfnet_emd_cost_grad_workspace_bytes(2, 4096, 4096) >= lib.rfnet_emd_cost_workspace_bytes(2, 4096, 4096)
    assert lib.rfnet_emd_cost_workspace_bytes(2, 4096, 4096) > lib.rfnet_approxmatch_workspace_bytes(2, 4096, 4096) > 0


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


def test_ops_are_cuda_graph_capturable(cuda, rng):
    """calls are stream-ordered with no host sync and no allocation inside the library, so a whole
    forward+backward step can be captured in a CUDA graph and replayed on new data."""
    import torch
    from rfnet_b200 import ops
    b, n, m = 4, 1024, 4096          # large enough to take the split/atomic-merge path of nn_distance
    x1 = torch.from_numpy(cloud(rng, b, n)).to(cuda)
    x2 = torch.from_numpy(cloud(rng, b, m)).to(cuda)
    f32, i32 = torch.float32, torch.int32
    d1, i1 = torch.empty((b, n), dtype=f32, device=cuda), torch.empty((b, n), dtype=i32, device=cuda)
    d2, i2 = torch.empty((b, m), dtype=f32, device=cuda), torch.empty((b, m), dtype=i32, device=cuda)
    g1, g2 = torch.empty_like(x1), torch.empty_like(x2)
    gd1, gd2 = torch.ones((b, n), device=cuda), torch.ones((b, m), device=cuda)
    sums = torch.empty(4, device=cuda)
    ws = torch.empty(ops.nn_distance_workspace_bytes(b, n, m), dtype=torch.uint8, device=cuda)
    fps_idx = torch.empty((b, 64), dtype=i32, device=cuda)

    def body():
        ops.raw_nn_distance(x1, x2, d1, i1, d2, i2, ws)
        ops.raw_nn_distance_grad(x1, x2, gd1, i1, gd2, i2, g1, g2, ws)     # workspace => atomic-free CSR scatter, also capturable
        ops.raw_chamfer_partial_sums(d1, d2, sums, ws)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()                        # warm-up outside capture (lazy module loading)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    # replay on NEW data written into the captured input buffers
    y1, y2 = cloud(rng, b, n), cloud(rng, b, m)
    x1.copy_(torch.from_numpy(y1)); x2.copy_(torch.from_numpy(y2))
    graph.replay()
    torch.cuda.synchronize()
    want = port.nn_distance(y1, y2)
    assert np.array_equal(d1.cpu().numpy(), want[0]) and np.array_equal(i1.cpu().numpy(), want[1])
    assert np.array_equal(d2.cpu().numpy(), want[2]) and np.array_equal(i2.cpu().numpy(), want[3])
    w1, w2 = port.nn_distance_grad(y1, y2, np.ones((b, n), np.float32), want[1], np.ones((b, m), np.float32), want[3])
    assert np.array_equal(g1.cpu().numpy(), w1) and np.array_equal(g2.cpu().numpy(), w2)


def test_model_level_callers(cuda, rng):
    """sampling / merge_layer / re_chamfer / zero_groupnear (vv_recon.py:67-70,132-139,171-193,414-419) over the drop-in ops,
    against the same formulas evaluated with the CPU oracle's indices."""
    import torch
    from rfnet_b200 import losses
    b, n_raw, n_new = 2, 300, 512
    raw, new = cloud(rng, b, n_raw), cloud(rng, b, n_new)
    traw, tnew = torch.from_numpy(raw).to(cuda), torch.from_numpy(new).to(cuda)
    # sampling == FPS + gather
    got = losses.sampling(32, traw).cpu().numpy()
    assert np.array_equal(got, port.gather_point(raw, port.farthest_point_sample(32, raw)))
    # merge_layer
    _, _, _, i2 = port.nn_distance(raw, new)
    nearest = np.take_along_axis(raw, i2[..., None].astype(np.int64), axis=1)
    diff = nearest - new
    want = new + np.exp(-(diff ** 2).sum(-1, keepdims=True) / (1e-8 + 0.05 ** 2)) * diff
    assert np.allclose(losses.merge_layer(traw, tnew, 0.05).cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    # re_chamfer: 8 slices of 64 points
    gt, pred = cloud(rng, b, 512), cloud(rng, b, 512)
    vals = []
    for i in range(8):
        d1, _, d2, _ = port.nn_distance(pred[:, i * 64:(i + 1) * 64], gt[:, i * 64:(i + 1) * 64])
        vals.append((np.sqrt(d1).mean() + np.sqrt(d2).mean()) / 2)
    got = losses.re_chamfer(torch.from_numpy(gt).to(cuda), torch.from_numpy(pred).to(cuda)).item()
    assert abs(got - np.mean(vals)) <= 1e-5 * abs(np.mean(vals))
    # zero_groupnear
    cens = cloud(rng, b, 16)
    outmat = rng.standard_normal((b, 16, 8, 3)).astype(np.float32) * 0.01
    _, _, dist, _ = port.nn_distance(cens, raw)
    want = max(0.0, float((outmat ** 2).sum(-1).mean() - 0.4 * dist.mean()))
    got = losses.zero_groupnear(torch.from_numpy(cens).to(cuda), traw, torch.from_numpy(outmat).to(cuda)).item()
    assert abs(got - want) <= 1e-5 * max(abs(want), 1e-6)


def test_64bit_offsets_beyond_int32(cuda):
    """b*n*m = 9 * 16384^2 = 2.4e9 > 2^31: the reference's approxmatch indexes with int and cannot run this
    (tf_approxmatch.cu:15); every cloud of the batched call must equal the same cloud run alone -- bit for bit, since the
    sweep plan does not depend on the batch."""
    import torch
    from rfnet_b200 import tf_approxmatch
    g = torch.Generator(device="cpu").manual_seed(17)
    b, n = 9, 16384
    x1 = (torch.rand((b, n, 3), generator=g) - 0.5).to(cuda)
    x2 = (torch.rand((b, n, 3), generator=g) - 0.5).to(cuda)
    match = tf_approxmatch.approx_match(x1, x2)
    assert match.numel() > 2 ** 31
    cost = tf_approxmatch.match_cost(x1, x2, match)
    for c in (0, 8):
        alone = tf_approxmatch.approx_match(x1[c:c + 1].contiguous(), x2[c:c + 1].contiguous())
        assert torch.equal(match[c], alone[0])
        c_alone = tf_approxmatch.match_cost(x1[c:c + 1].contiguous(), x2[c:c + 1].contiguous(), alone)
        assert cost[c].item() == c_alone.item()
    del match


def test_recon_test_loop_over_pcd_files(cuda, tmp_path):
    """`bench.py --workload recon_loss --data-dir` runs the reference's recon_test.py loss loop (:46-68) on PCD files
    through io_util.read_pcd / resample_pcd and writes results.csv with the reference's columns; values equal the CPU oracle's."""
    import csv
    import json
    import os
    import subprocess
    import sys
    from rfnet_b200 import io_util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.default_rng(5)
    data = os.path.join(tmp_path, "pcn")
    ids = ["02691156/a1", "02691156/b2", "03001627/c3"]
    clouds = {}
    for mid in ids:
        for sub, npts in (("partial", 700), ("complete", 2048), ("completion", 2048)):
            os.makedirs(os.path.join(data, sub, os.path.dirname(mid)), exist_ok=True)
            pts = (rng.random((npts, 3)) - 0.5).astype(np.float32)
            io_util.save_pcd(os.path.join(data, sub, mid + ".pcd"), pts, binary=(sub != "partial"))
            clouds[(mid, sub)] = pts
    lst = os.path.join(tmp_path, "list.txt")
    open(lst, "w").write("\n".join(ids) + "\n")
    out = os.path.join(tmp_path, "results")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "recon_loss", "--data-dir", data, "--list-path", lst,
                        "--results-dir", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["config"]["models"] == 3 and line["config"]["outputs_synthesised"] == 0
    rows = list(csv.reader(open(os.path.join(out, "results.csv"))))
    assert rows[0] == ["id", "cd", "emd"] and [r_[0] for r_ in rows[1:]] == ids
    for r_ in rows[1:]:
        gt, output = clouds[(r_[0], "complete")][None], clouds[(r_[0], "completion")][None]
        d1, _, d2, _ = port.nn_distance(output, gt)
        want_cd = (np.sqrt(d1).mean() + np.sqrt(d2).mean()) / 2                     # chamfer_big(output, gt), recon_test.py:27
        assert abs(float(r_[1]) - want_cd) <= 1e-5 * want_cd
    assert set(line["per_category_cd_emd"]) == {"02691156", "03001627"}
