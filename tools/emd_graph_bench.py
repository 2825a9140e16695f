"""Is the small-batch EMD call bound by launches?  emd_cost_grad eager vs replayed as a CUDA graph (same kernels, same buffers)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfnet_b200 import _lib
lib = _lib.load()
g = torch.Generator(device="cpu").manual_seed(3)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for n, b in ((1024, 1), (2048, 1), (2048, 4), (2048, 8), (2048, 32), (16384, 1), (16384, 4)):
    x1 = (torch.rand((b, n, 3), generator=g) - 0.5).cuda()
    x2 = (torch.rand((b, n, 3), generator=g) - 0.5).cuda()
    cost = torch.empty(b, device="cuda"); g1 = torch.empty_like(x1); g2 = torch.empty_like(x2)
    wsb = lib.rfnet_emd_cost_grad_workspace_bytes(b, n, n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    def call(stream):
        rc = lib.rfnet_emd_cost_grad(b, n, n, p(x1), p(x2), p(cost), p(g1), p(g2), p(ws), wsb, 0, ctypes.c_void_p(stream.cuda_stream))
        assert rc == 0
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): call(s)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if n < 16384 else 5
        e0.record(s)
        for _ in range(reps): call(s)
        e1.record(s); s.synchronize()
        eager = e0.elapsed_time(e1) / reps
        ref = cost.clone()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            call(s)
        gr.replay(); s.synchronize()
        e0.record(s)
        for _ in range(reps): gr.replay()
        e1.record(s); s.synchronize()
        graph = e0.elapsed_time(e1) / reps
    print("n=%5d B=%2d: eager %.3f ms, graph replay %.3f ms (x%.2f), same cost: %s" % (n, b, eager, graph, eager / graph, torch.equal(ref, cost)), flush=True)
