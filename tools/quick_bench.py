"""Ad-hoc GPU timing of our ops next to the reference CUDA kernels (oracle/_ref/libref_gpu.so).  Development tool."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rfnet_b200 import ops, tf_nndistance, tf_approxmatch, tf_sampling, tf_grouping, tf_interpolate
from oracle import ref

dev = torch.device("cuda:0")


def timeit(fn, warm=3, iters=10, reps=3):
    """ms per call, `iters` calls back to back between two events (host launch overhead overlaps GPU work); min over reps."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / iters)
    return float(np.median(ts)), float(np.min(ts))


def rnd(b, n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand((b, n, 3), generator=g) - 0.5).to(dev)


which = sys.argv[1:] or ["nn"]
have_ref = ref.available("gpu")
if "nn" in which:
    for (b, n, m) in [(32, 2048, 16384), (32, 16384, 16384), (4, 16384, 16384), (32, 64, 1024), (32, 1024, 16384)]:
        x1, x2 = rnd(b, n, 1), rnd(b, m, 2)
        med, mn = timeit(lambda: tf_nndistance.nn_distance(x1, x2))
        pairs = 2.0 * b * n * m
        line = "nn_distance b=%d n=%d m=%d: ours %.3f ms (min %.3f) = %.1f Gpairs/s (%.1f%% of 6.2e12)" % (b, n, m, med, mn, pairs / mn / 1e6, pairs / mn / 1e6 / 6200 * 100)
        if have_ref:
            rmed, rmn = timeit(lambda: ref.run_gpu("NnDistance", [x1, x2], [((b, n), torch.float32), ((b, n), torch.int32), ((b, m), torch.float32), ((b, m), torch.int32)]), iters=5)
            line += " | reference kernel %.3f ms = %.1f Gpairs/s | speedup %.2fx" % (rmn, pairs / rmn / 1e6, rmn / mn)
        print(line, flush=True)
if "emd" in which:
    for (b, n) in [(32, 2048), (4, 2048), (32, 1024), (32, 64), (4, 16384), (8, 16384), (1, 16384), (32, 16384)]:
        x1, x2 = rnd(b, n, 3), rnd(b, n, 4)
        med, mn = timeit(lambda: tf_approxmatch.approx_match(x1, x2), warm=2, iters=5)
        match = tf_approxmatch.approx_match(x1, x2)
        cmed, cmn = timeit(lambda: tf_approxmatch.match_cost(x1, x2, match), warm=2, iters=5)
        gmed, gmn = timeit(lambda: ops.match_cost_grad_op(x1, x2, match), warm=2, iters=5)
        _, xmn = timeit(lambda: ops.approx_match_op(x1, x2, 1), warm=2, iters=5)
        _, smn = timeit(lambda: ops.approx_match_op(x1, x2, 4), warm=2, iters=5)
        _, pmn = timeit(lambda: ops.approx_match_op(x1, x2, 8), warm=2, iters=5)
        line = "emd b=%d n=m=%d: approx_match %.3f ms (%.1f clouds/s) [exact mode %.3f ms, split sums %.3f ms, pruned %.3f ms], match_cost %.3f ms, grad %.3f ms" % (b, n, mn, b / mn * 1e3, xmn, smn, pmn, cmn, gmn)
        if have_ref and b * n * n < 2 ** 31:
            rmed, rmn = timeit(lambda: ref.run_gpu("ApproxMatch", [x1, x2], [((b, n, n), torch.float32)]), warm=1, iters=2)
            line += " | reference approxmatch %.3f ms (%.1f clouds/s) speedup %.1fx" % (rmn, b / rmn * 1e3, rmn / mn)
        print(line, flush=True)
        del match
if "fps" in which:
    for (b, n, m) in [(32, 16384, 2048), (4, 16384, 2048), (32, 3000, 32), (32, 16384, 1024)]:
        x = rnd(b, n, 5)
        med, mn = timeit(lambda: tf_sampling.farthest_point_sample(m, x), warm=2, iters=5)
        line = "fps b=%d n=%d m=%d: ours %.3f ms (%.2f us/iter)" % (b, n, m, mn, mn * 1e3 / m)
        if have_ref:
            rmed, rmn = timeit(lambda: ref.run_gpu("FarthestPointSample", [x], [((b, m), torch.int32)], attrs={"npoint": m}), warm=1, iters=3)
            line += " | reference %.3f ms speedup %.1fx" % (rmn, rmn / mn)
        print(line, flush=True)
if "group" in which:
    b, n, m, ns = 32, 16384, 2048, 32
    x = rnd(b, n, 6)
    idx = tf_sampling.farthest_point_sample(m, x)
    q = tf_sampling.gather_point(x, idx)
    med, mn = timeit(lambda: tf_grouping.query_ball_point(0.1, ns, x, q))
    line = "query_ball_point b=%d n=%d m=%d ns=%d: ours %.3f ms" % (b, n, m, ns, mn)
    r = torch.tensor([0.1], device=dev)
    if have_ref:
        rmed, rmn = timeit(lambda: ref.run_gpu("QueryBallPoint", [x, q, r], [((b, m, ns), torch.int32), ((b, m), torch.int32)], attrs={"nsample": ns}), warm=1, iters=3)
        line += " | reference %.3f ms speedup %.1fx" % (rmn, rmn / mn)
    print(line, flush=True)
    gi, _ = tf_grouping.query_ball_point(0.1, ns, x, q)
    for c in (3, 64):
        pts = torch.randn((b, n, c), device=dev)
        med, mn = timeit(lambda: tf_grouping.group_point(pts, gi))
        byts = 4.0 * b * m * ns * (c + 1) + 4.0 * b * n * c
        line = "group_point c=%d: ours %.3f ms = %.0f GB/s" % (c, mn, byts / mn / 1e6)
        if have_ref:
            rmed, rmn = timeit(lambda: ref.run_gpu("GroupPoint", [pts, gi], [((b, m, ns, c), torch.float32)]), warm=1, iters=3)
            line += " | reference %.3f ms speedup %.1fx" % (rmn, rmn / mn)
        print(line, flush=True)
        go = torch.randn((b, m, ns, c), device=dev)
        med, mn = timeit(lambda: ops.group_point_grad_op(pts, gi, go))
        print("group_point_grad c=%d: ours %.3f ms" % (c, mn), flush=True)
    known = q
    med, mn = timeit(lambda: tf_interpolate.three_nn(x, known))
    print("three_nn b=%d n=%d m=%d: ours %.3f ms = %.1f Gpairs/s" % (b, n, m, mn, b * n * m / mn / 1e6), flush=True)
    d, i3 = tf_interpolate.three_nn(x, known)
    w = torch.rand((b, n, 3), device=dev)
    feats = torch.randn((b, m, 64), device=dev)
    med, mn = timeit(lambda: tf_interpolate.three_interpolate(feats, i3, w))
    byts = 4.0 * b * n * (64 + 6) + 4.0 * b * m * 64
    print("three_interpolate c=64: ours %.3f ms = %.0f GB/s" % (mn, byts / mn / 1e6), flush=True)
    go = torch.randn((b, n, 64), device=dev)
    med, mn = timeit(lambda: ops.three_interpolate_grad_op(feats, i3, w, go))
    print("three_interpolate_grad c=64: ours %.3f ms" % mn, flush=True)
if "emdcost" in which:
    for (b, n) in [(32, 2048), (4, 16384), (8, 16384)]:
        x1, x2 = rnd(b, n, 3), rnd(b, n, 4)
        def chain():
            match = tf_approxmatch.approx_match(x1, x2)
            return tf_approxmatch.match_cost(x1, x2, match)
        _, cmn = timeit(chain, warm=2, iters=5)
        _, fmn = timeit(lambda: ops.emd_cost_op(x1, x2, False), warm=2, iters=5)
        _, kmn = timeit(lambda: ops.emd_cost_op(x1, x2, True), warm=2, iters=5)
        _, gmn = timeit(lambda: ops.emd_cost_grad_op(x1, x2), warm=2, iters=5)
        print("emd cost b=%d n=m=%d: approx_match + match_cost %.3f ms | fused, no matrix %.3f ms (%.1f clouds/s) | fused, matrix kept %.3f ms | cost + both grads, no matrix %.3f ms (%.1f clouds/s)"
              % (b, n, cmn, fmn, b / fmn * 1e3, kmn, gmn, b / gmn * 1e3), flush=True)
if "auction" in which:
    from rfnet_b200 import tf_auctionmatch
    for (b, n) in [(32, 256), (32, 1024), (148, 1024), (8, 2048), (8, 4096), (148, 4096)]:
        x1, x2 = rnd(b, n, 5), rnd(b, n, 6)
        _, mn = timeit(lambda: tf_auctionmatch.auction_match(x1, x2), warm=1, iters=2, reps=2)
        line = "auction_match b=%d n=%d: ours %.2f ms (%.1f clouds/s)" % (b, n, mn, b / mn * 1e3)
        if have_ref and n <= 4096 and b <= 32:
            _, rmn = timeit(lambda: ref.run_gpu("AuctionMatch", [x1, x2], [((b, n), torch.int32), ((b, n), torch.int32)]), warm=1, iters=1, reps=2)
            line += " | reference kernel %.2f ms | speedup %.1fx" % (rmn, rmn / mn)
        print(line, flush=True)
if "select" in which:
    for (b, m, n, k) in [(32, 2048, 2048, 32), (8, 1024, 16384, 16)]:
        g = torch.Generator(device="cpu").manual_seed(9)
        d = torch.rand((b, m, n), generator=g).to(dev)
        _, mn = timeit(lambda: tf_grouping.select_top_k(k, d), warm=1, iters=3)
        line = "select_top_k b=%d m=%d n=%d k=%d: ours %.3f ms (%.0f GB/s over 12 B/entry)" % (b, m, n, k, mn, 12.0 * b * m * n / mn / 1e6)
        if have_ref:
            _, rmn = timeit(lambda: ref.run_gpu("SelectionSort", [d], [((b, m, n), torch.int32), ((b, m, n), torch.float32)], attrs={"k": k}), warm=1, iters=1, reps=2)
            line += " | reference kernel %.2f ms | speedup %.1fx" % (rmn, rmn / mn)
        print(line, flush=True)
