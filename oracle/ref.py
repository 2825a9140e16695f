"""Front-end of oracle/_ref/libref_{cpu,gpu}.so: the reference's own kernels.  TEST INFRASTRUCTURE ONLY.

``libref_cpu.so`` holds the reference CPU OpKernels (NnDistance, NnDistanceGrad, ApproxMatch, MatchCost, MatchCostGrad,
ThreeNN, ThreeInterpolate, ThreeInterpolateGrad) compiled with the reference's own flags; ``libref_gpu.so`` additionally
holds the reference CUDA kernels recompiled unchanged for sm_100a (the GPU oracle, and "the kernel to beat").

Ops are run by name through the same Compute() the TF executor would call (oracle/ref_driver.cpp).  CPU ops take numpy
arrays; GPU ops take torch CUDA tensors and launch on the legacy default stream exactly as the reference does.
"""
import ctypes
import os

import numpy as np

from . import HERE

_libs = {}
F32, I32 = 1, 3


def available(kind="cpu"):
    return os.path.exists(os.path.join(HERE, "_ref", "libref_%s.so" % kind))


def lib(kind="cpu"):
    if kind not in _libs:
        path = os.path.join(HERE, "_ref", "libref_%s.so" % kind)
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing: run `make -C oracle ref` where /root/reference exists" % path)
        l = ctypes.CDLL(path)
        l.ref_last_error.restype = ctypes.c_char_p
        _libs[kind] = l
    return _libs[kind]


def has_kernel(op, device, kind="cpu"):
    return bool(lib(kind).ref_has_kernel(op.encode(), device.encode()))


def _run(kind, op, device, in_ptrs, in_dtypes, in_shapes, attrs, out_ptrs, out_caps):
    n_in, n_out = len(in_ptrs), len(out_ptrs)
    dims = (ctypes.c_longlong * (4 * max(n_in, 1)))()
    for i, s in enumerate(in_shapes):
        for k, d in enumerate(s):
            dims[4 * i + k] = d
    names = [k.encode() for k in attrs]
    out_rank = (ctypes.c_int * max(n_out, 1))()
    out_dims = (ctypes.c_longlong * (4 * max(n_out, 1)))()
    rc = lib(kind).ref_run_op(
        op.encode(), device.encode(), n_in,
        (ctypes.c_void_p * max(n_in, 1))(*in_ptrs), (ctypes.c_int * max(n_in, 1))(*in_dtypes),
        (ctypes.c_int * max(n_in, 1))(*[len(s) for s in in_shapes]), dims,
        len(names), (ctypes.c_char_p * max(len(names), 1))(*names), (ctypes.c_int * max(len(names), 1))(*[int(v) for v in attrs.values()]),
        n_out, (ctypes.c_void_p * max(n_out, 1))(*out_ptrs), (ctypes.c_longlong * max(n_out, 1))(*out_caps), out_rank, out_dims)
    if rc != 0:
        raise ValueError("InvalidArgument: " + lib(kind).ref_last_error().decode())
    return [tuple(out_dims[4 * i + k] for k in range(out_rank[i])) for i in range(n_out)]


def run_cpu(op, inputs, outputs, attrs=None):
    """inputs: numpy arrays (float32/int32).  outputs: list of (shape, dtype) -> list of numpy arrays."""
    ins = [np.ascontiguousarray(a) for a in inputs]
    for a in ins:
        assert a.dtype in (np.float32, np.int32), a.dtype
    outs = [np.empty(s, dt) for s, dt in outputs]
    shapes = _run("cpu", op, "CPU", [a.ctypes.data for a in ins], [F32 if a.dtype == np.float32 else I32 for a in ins],
                  [a.shape for a in ins], attrs or {}, [o.ctypes.data for o in outs], [o.nbytes for o in outs])
    for o, s in zip(outs, shapes):
        assert tuple(o.shape) == tuple(s), (op, o.shape, s)
    return outs


def run_gpu(op, inputs, outputs, attrs=None, zero_outputs=False):
    """inputs: contiguous torch CUDA tensors.  outputs: list of (shape, torch dtype) -> list of CUDA tensors.
    Runs on the legacy default stream (== torch's default stream); the caller's current stream must be that one."""
    import torch
    ins = [t.contiguous() for t in inputs]
    dev = ins[0].device
    alloc = torch.zeros if zero_outputs else torch.empty
    outs = [alloc(s, dtype=dt, device=dev) for s, dt in outputs]
    assert torch.cuda.current_stream(dev) == torch.cuda.default_stream(dev)
    shapes = _run("gpu", op, "GPU", [t.data_ptr() for t in ins], [F32 if t.dtype == torch.float32 else I32 for t in ins],
                  [tuple(t.shape) for t in ins], attrs or {}, [o.data_ptr() for o in outs],
                  [o.numel() * o.element_size() for o in outs])
    for o, s in zip(outs, shapes):
        assert tuple(o.shape) == tuple(s), (op, o.shape, s)
    return outs


# ---------------------------------------------------------------------------------------------------------------------
# Convenience wrappers with the reference's Python signatures (CPU kernels, numpy in/out).
# ---------------------------------------------------------------------------------------------------------------------
def nn_distance(xyz1, xyz2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    return run_cpu("NnDistance", [xyz1, xyz2], [((b, n), np.float32), ((b, n), np.int32), ((b, m), np.float32), ((b, m), np.int32)])


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    return run_cpu("NnDistanceGrad", [xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2],
                   [(xyz1.shape, np.float32), (xyz2.shape, np.float32)])


def approx_match(xyz1, xyz2):
    """Reference CPU ApproxMatch: tensor shaped (b, m, n) whose ELEMENT ORDER is (b, n, m) (SURVEY.md 8c divergence 2)."""
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    return run_cpu("ApproxMatch", [xyz1, xyz2], [((b, m, n), np.float32)])[0]


def match_cost(xyz1, xyz2, match):
    return run_cpu("MatchCost", [xyz1, xyz2, match], [((xyz1.shape[0],), np.float32)])[0]


def match_cost_grad(xyz1, xyz2, match):
    # matchcostgrad_cpu only zeroes grad1[:, :, 0] (tf_approxmatch.cpp:108-109): pre-zero the buffers ourselves.
    ins = [np.ascontiguousarray(a) for a in (xyz1, xyz2, match)]
    outs = [np.zeros(xyz1.shape, np.float32), np.zeros(xyz2.shape, np.float32)]
    _run("cpu", "MatchCostGrad", "CPU", [a.ctypes.data for a in ins], [F32] * 3, [a.shape for a in ins], {},
         [o.ctypes.data for o in outs], [o.nbytes for o in outs])
    return outs


def three_nn(xyz1, xyz2):
    b, n, _ = xyz1.shape
    return run_cpu("ThreeNN", [xyz1, xyz2], [((b, n, 3), np.float32), ((b, n, 3), np.int32)])


def three_interpolate(points, idx, weight):
    b, m, c = points.shape
    n = idx.shape[1]
    return run_cpu("ThreeInterpolate", [points, idx, weight], [((b, n, c), np.float32)])[0]


def three_interpolate_grad(points, idx, weight, grad_out):
    return run_cpu("ThreeInterpolateGrad", [points, idx, weight, grad_out], [(points.shape, np.float32)])[0]
