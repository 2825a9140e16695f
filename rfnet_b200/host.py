"""Host-buffer front-end of the Chamfer path: what a caller that owns HOST memory uses (the reference's CPU-side callers
feed numpy batches through feed_dict, vv_recon.py:424-428, and read losses back).

``ChamferHostPipeline`` takes pinned host clouds, runs nn_distance forward + gradient + the loss partial sums on the
GPU, and returns the results in pinned host buffers.  Host->device copies, compute and device->host copies of
consecutive batches overlap on three CUDA streams (a ring of `depth` device/host slots); every batch still pays its own
copies -- they are just hidden behind the neighbouring batches' kernels, which is how a PCIe-attached B200 is fed.
"""
import torch

from . import ops


class ChamferHostPipeline:
    ALL_OUTPUTS = ("dist1", "idx1", "dist2", "idx2", "grad1", "grad2", "sums")

    def __init__(self, b, n, m, device, depth=3, grad_scale1=None, grad_scale2=None, outputs=ALL_OUTPUTS, deterministic=False):
        """outputs: which results are copied back to the host each step.  A training loop reads back only the loss
        ("sums": sum sqrt(dist1), count1, sum sqrt(dist2), count2 -> chamfer_big); gradients normally stay on the device."""
        self.b, self.n, self.m, self.dev, self.depth = b, n, m, torch.device(device), depth
        self.outputs = tuple(outputs)
        self.deterministic = bool(deterministic)   # atomic-free CSR scatter for the gradient (bit-reproducible, slower)
        dev = self.dev
        self.s_in, self.s_out, self.s_red = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.x1 = [torch.empty((b, n, 3), device=dev) for _ in range(depth)]
        self.x2 = [torch.empty((b, m, 3), device=dev) for _ in range(depth)]
        # device-side results and scratch, one set per slot: the steady-state loop allocates nothing and calls the C ABI directly
        f32, i32 = torch.float32, torch.int32
        self.dres = [dict(dist1=torch.empty((b, n), dtype=f32, device=dev), idx1=torch.empty((b, n), dtype=i32, device=dev),
                          dist2=torch.empty((b, m), dtype=f32, device=dev), idx2=torch.empty((b, m), dtype=i32, device=dev),
                          grad1=torch.empty((b, n, 3), dtype=f32, device=dev), grad2=torch.empty((b, m, 3), dtype=f32, device=dev),
                          sums=torch.empty((4,), dtype=f32, device=dev),
                          ws=torch.empty((ops.nn_distance_workspace_bytes(b, n, m),), dtype=torch.uint8, device=dev)) for _ in range(depth)]
        pin = lambda *s, dtype=torch.float32: torch.empty(s, dtype=dtype).pin_memory()
        self.out = [dict(dist1=pin(b, n), dist2=pin(b, m), idx1=pin(b, n, dtype=torch.int32), idx2=pin(b, m, dtype=torch.int32),
                         grad1=pin(b, n, 3), grad2=pin(b, m, 3), sums=pin(4)) for _ in range(depth)]
        self.gd1 = torch.full((b, n), 0.5 / (b * n) if grad_scale1 is None else grad_scale1, device=dev)
        self.gd2 = torch.full((b, m), 0.5 / (b * m) if grad_scale2 is None else grad_scale2, device=dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_compute = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.ev_red = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0
        self.h2d_bytes = b * (n + m) * 12
        per = {"dist1": b * n * 4, "idx1": b * n * 4, "dist2": b * m * 4, "idx2": b * m * 4, "grad1": b * n * 12, "grad2": b * m * 12, "sums": 16}
        self.d2h_bytes = sum(per[k] for k in self.outputs)

    def submit(self, h_xyz1, h_xyz2, reduce_fn=None):
        """Enqueue one batch (pinned host tensors).  Returns the slot whose `out` buffers will hold the results once
        `wait(slot)` returns.  `reduce_fn` (optional) is applied in place to the 4 loss partial sums (e.g. an all-reduce
        across ranks).  It runs on a SIDE stream that only the copy-out of the sums waits for: the collective's launch
        latency and the slowest rank never sit between two batches on the compute stream."""
        with torch.cuda.device(self.dev):
            return self._submit(h_xyz1, h_xyz2, reduce_fn)

    def _submit(self, h_xyz1, h_xyz2, reduce_fn):
        k = self.count % self.depth
        self.count += 1
        compute = torch.cuda.current_stream(self.dev)
        # the slot's device inputs are free once the compute that last read them has finished; its host outputs once
        # the previous copy-out has finished
        self.s_in.wait_event(self.ev_compute[k])
        with torch.cuda.stream(self.s_in):
            self.x1[k].copy_(h_xyz1, non_blocking=True)
            self.x2[k].copy_(h_xyz2, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        compute.wait_event(self.ev_in[k])
        compute.wait_event(self.ev_out[k])   # the slot's device results were last read by the copy-out of batch count - depth
        r = self.dres[k]
        if self.deterministic:
            ops.raw_nn_distance(self.x1[k], self.x2[k], r["dist1"], r["idx1"], r["dist2"], r["idx2"], r["ws"])
            ops.raw_nn_distance_grad(self.x1[k], self.x2[k], self.gd1, r["idx1"], self.gd2, r["idx2"], r["grad1"], r["grad2"], r["ws"])
            ops.raw_chamfer_partial_sums(r["dist1"], r["dist2"], r["sums"], r["ws"])
        else:
            # search + one epilogue (unpack, gradient, sqrt partial sums) + one reduction (+ the preparation kernel of the filtered search): four launches
            ops.raw_chamfer_step(self.x1[k], self.x2[k], self.gd1, self.gd2, r["dist1"], r["idx1"], r["dist2"], r["idx2"], r["grad1"], r["grad2"],
                                 r["sums"], r["ws"])
        self.ev_compute[k].record(compute)
        self.s_out.wait_event(self.ev_compute[k])
        if reduce_fn is not None:
            self.s_red.wait_event(self.ev_compute[k])
            with torch.cuda.stream(self.s_red):
                reduce_fn(r["sums"])          # in place
                self.ev_red[k].record(self.s_red)
            self.s_out.wait_event(self.ev_red[k])
        with torch.cuda.stream(self.s_out):
            o = self.out[k]
            for name in self.outputs:
                o[name].copy_(r[name], non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.last = r                         # device-side results of the last batch
        return k

    def wait(self, slot):
        self.ev_out[slot].synchronize()
        return self.out[slot]

    def drain(self):
        self.s_red.synchronize()
        self.s_out.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
