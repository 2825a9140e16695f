"""PyTorch custom ops (namespace ``rfnet::``) over the C ABI of librfnet_ops.so.

One op per TensorFlow op the reference registers (REGISTER_OP in pc_distance/tf_nndistance.cpp:3-18,
pc_distance/tf_approxmatch.cpp:7-21, tf_ops/sampling/tf_sampling.cpp:14-63, tf_ops/grouping/tf_grouping.cpp:14-64,
tf_ops/interpolation/tf_interpolate.cpp:12-46): same inputs, attrs, output tuples, dtypes and shape checks; the checks
raise ``ValueError`` with the reference's InvalidArgument message.  Autograd mirrors each ``RegisterGradient`` /
``NoGradient`` of the reference's Python wrappers.  torch is plumbing here (device memory, streams, autograd graph); all
compute is in the CUDA library and there is no other path: CPU tensors are rejected.
"""
import ctypes

import torch

from . import _lib

_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None and t.numel() > 0 else _vp(0)


def _stream(t):
    return _vp(torch.cuda.current_stream(t.device).cuda_stream)


def _require(cond, msg):
    if not cond:
        raise ValueError(msg)


def _cuda_f32(name, t):
    _require(isinstance(t, torch.Tensor) and t.is_cuda, "%s must be a CUDA tensor (rfnet_b200 has no CPU path)" % name)
    _require(t.dtype == torch.float32, "%s must be float32" % name)
    return t.contiguous()


def _cuda_i32(name, t):
    _require(isinstance(t, torch.Tensor) and t.is_cuda, "%s must be a CUDA tensor (rfnet_b200 has no CPU path)" % name)
    _require(t.dtype == torch.int32, "%s must be int32" % name)
    return t.contiguous()


def _workspace(nbytes, device):
    return torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------------------ raw calls
# Allocation-free entry points over caller-owned CUDA tensors (contiguous, right dtype, all on the current device): what
# rfnet_b200.host uses in its steady-state loop, where the torch dispatcher and per-call allocations would dominate.
def raw_nn_distance(xyz1, xyz2, dist1, idx1, dist2, idx2, workspace, unfused=False, direct=False):
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    _lib.check(_lib.load().rfnet_nn_distance(b, n, _ptr(xyz1), m, _ptr(xyz2), _ptr(dist1), _ptr(idx1), _ptr(dist2), _ptr(idx2), _ptr(workspace),
                                             workspace.numel(), (1 if unfused else 0) | (2 if direct else 0), _stream(xyz1)), "rfnet_nn_distance")


def raw_nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, workspace=None):
    """workspace (uint8 tensor of rfnet_nn_distance_grad_workspace_bytes) selects the atomic-free deterministic scatter."""
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    _lib.check(_lib.load().rfnet_nn_distance_grad(b, n, _ptr(xyz1), m, _ptr(xyz2), _ptr(grad_dist1), _ptr(idx1), _ptr(grad_dist2), _ptr(idx2),
                                                  _ptr(grad_xyz1), _ptr(grad_xyz2), _ptr(workspace), 0 if workspace is None else workspace.numel(),
                                                  _stream(xyz1)), "rfnet_nn_distance_grad")


def raw_chamfer_partial_sums(dist1, dist2, sums4, workspace):
    _lib.check(_lib.load().rfnet_chamfer_partial_sums(dist1.shape[0], dist1.shape[1], dist2.shape[1], _ptr(dist1), _ptr(dist2), _ptr(sums4),
                                                      _ptr(workspace), workspace.numel(), _stream(dist1)), "rfnet_chamfer_partial_sums")


def raw_chamfer_step(xyz1, xyz2, grad_dist1, grad_dist2, dist1, idx1, dist2, idx2, grad_xyz1, grad_xyz2, sums4, workspace, unfused=False, direct=False):
    """nn_distance + NnDistanceGrad + the chamfer partial sums: one C-ABI call, three kernel launches (rfnet_chamfer_step)."""
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    _lib.check(_lib.load().rfnet_chamfer_step(b, n, _ptr(xyz1), m, _ptr(xyz2), _ptr(grad_dist1), _ptr(grad_dist2), _ptr(dist1), _ptr(idx1), _ptr(dist2),
                                              _ptr(idx2), _ptr(grad_xyz1), _ptr(grad_xyz2), _ptr(sums4), _ptr(workspace), workspace.numel(),
                                              (1 if unfused else 0) | (2 if direct else 0), _stream(xyz1)), "rfnet_chamfer_step")


def nn_distance_workspace_bytes(b, n, m):
    lib = _lib.load()
    return max(int(lib.rfnet_nn_distance_workspace_bytes(b, n, m)), int(lib.rfnet_chamfer_partial_sums_workspace_bytes()),
               int(lib.rfnet_nn_distance_grad_workspace_bytes(b, n, m)), int(lib.rfnet_chamfer_step_workspace_bytes(b, n, m)), 16)


# ------------------------------------------------------------------------------------------------------------ nn_distance
NN_UNFUSED, NN_DIRECT = 1, 2   # rfnet_ops.h: RFNET_NN_UNFUSED, RFNET_NN_DIRECT


def _nn_distance_call(xyz1, xyz2, flags, count_exact_scans=False):
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    dev = xyz1.device
    dist1 = torch.empty((b, n), dtype=torch.float32, device=dev)
    idx1 = torch.empty((b, n), dtype=torch.int32, device=dev)
    dist2 = torch.empty((b, m), dtype=torch.float32, device=dev)
    idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
    lib = _lib.load()
    wsb = lib.rfnet_nn_distance_workspace_bytes(b, n, m)
    ws = _workspace(wsb, dev)
    scans = torch.zeros((1,), dtype=torch.int64, device=dev) if count_exact_scans else None
    with torch.cuda.device(dev):
        _lib.check(lib.rfnet_nn_distance_stats(b, n, _ptr(xyz1), m, _ptr(xyz2), _ptr(dist1), _ptr(idx1), _ptr(dist2), _ptr(idx2), _ptr(ws), wsb,
                                               int(flags), _ptr(scans), _stream(xyz1)), "rfnet_nn_distance")
    return dist1, idx1, dist2, idx2, scans


def nn_distance_exact_scans(xyz1, xyz2, unfused=False):
    """Diagnostics: the outputs of nn_distance plus the number of (query, work item) pairs the filtered search could not certify and
    scanned with the reference expression (rfnet_nn_distance_stats)."""
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    d1, i1, d2, i2, scans = _nn_distance_call(xyz1, xyz2, NN_UNFUSED if unfused else 0, count_exact_scans=True)
    return d1, i1, d2, i2, int(scans.item())


@torch.library.custom_op("rfnet::nn_distance", mutates_args=(), device_types="cuda")
def nn_distance_op(xyz1: torch.Tensor, xyz2: torch.Tensor, unfused: bool = False, direct: bool = False) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    # unfused: the reference CPU build's distance expression; direct: evaluate the reference expression for every pair (nn_search_kernel)
    # instead of the filtered search (nn_filter_kernel) -- same outputs bit for bit, see csrc/nn_distance.cu
    # shape checks of NnDistanceGpuOp::Compute, pc_distance/tf_nndistance.cpp:175-182
    _require(xyz1.dim() == 3, "NnDistance requires xyz1 be of shape (batch,#points,3)")
    _require(xyz1.shape[2] == 3, "NnDistance only accepts 3d point set xyz1")
    _require(xyz2.dim() == 3, "NnDistance requires xyz2 be of shape (batch,#points,3)")
    _require(xyz2.shape[2] == 3, "NnDistance only accepts 3d point set xyz2")
    _require(xyz2.shape[0] == xyz1.shape[0], "NnDistance expects xyz1 and xyz2 have same batch size")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    return _nn_distance_call(xyz1, xyz2, (NN_UNFUSED if unfused else 0) | (NN_DIRECT if direct else 0))[:4]


@nn_distance_op.register_fake
def _(xyz1, xyz2, unfused=False, direct=False):
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    return (xyz1.new_empty((b, n)), xyz1.new_empty((b, n), dtype=torch.int32), xyz1.new_empty((b, m)), xyz1.new_empty((b, m), dtype=torch.int32))


@torch.library.custom_op("rfnet::nn_distance_grad", mutates_args=(), device_types="cuda")
def nn_distance_grad_op(xyz1: torch.Tensor, xyz2: torch.Tensor, grad_dist1: torch.Tensor, idx1: torch.Tensor, grad_dist2: torch.Tensor,
                        idx2: torch.Tensor, deterministic: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    # NnDistanceGradGpuOp::Compute, pc_distance/tf_nndistance.cpp:209-253.
    # deterministic=False: the reference GPU formulation (float reductions, tf_nndistance_g.cu:131-150), fastest.
    # deterministic=True : atomic-free CSR scatter, bit-exact with the reference CPU kernel's summation order (~4x the
    #                      cost of this small op; set RFNET_DETERMINISTIC=1 to make it the default for autograd too).
    _require(xyz1.dim() == 3, "NnDistanceGrad requires xyz1 be of shape (batch,#points,3)")
    _require(xyz1.shape[2] == 3, "NnDistanceGrad only accepts 3d point set xyz1")
    _require(xyz2.dim() == 3, "NnDistanceGrad requires xyz2 be of shape (batch,#points,3)")
    _require(xyz2.shape[2] == 3, "NnDistanceGrad only accepts 3d point set xyz2")
    _require(xyz2.shape[0] == xyz1.shape[0], "NnDistanceGrad expects xyz1 and xyz2 have same batch size")
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    _require(tuple(grad_dist1.shape) == (b, n), "NnDistanceGrad requires grad_dist1 be of shape(batch,#points)")
    _require(tuple(idx1.shape) == (b, n), "NnDistanceGrad requires idx1 be of shape(batch,#points)")
    _require(tuple(grad_dist2.shape) == (b, m), "NnDistanceGrad requires grad_dist2 be of shape(batch,#points)")
    _require(tuple(idx2.shape) == (b, m), "NnDistanceGrad requires idx2 be of shape(batch,#points)")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    grad_dist1, grad_dist2 = _cuda_f32("grad_dist1", grad_dist1), _cuda_f32("grad_dist2", grad_dist2)
    idx1, idx2 = _cuda_i32("idx1", idx1), _cuda_i32("idx2", idx2)
    g1, g2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    lib = _lib.load()
    wsb = lib.rfnet_nn_distance_grad_workspace_bytes(b, n, m) if deterministic else 0   # workspace => atomic-free deterministic scatter
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_nn_distance_grad(b, n, _ptr(xyz1), m, _ptr(xyz2), _ptr(grad_dist1), _ptr(idx1), _ptr(grad_dist2), _ptr(idx2),
                                              _ptr(g1), _ptr(g2), _ptr(ws) if wsb else _vp(0), wsb, _stream(xyz1)), "rfnet_nn_distance_grad")
    return g1, g2


@nn_distance_grad_op.register_fake
def _(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, deterministic=False):
    return torch.empty_like(xyz1), torch.empty_like(xyz2)


import os as _os
DETERMINISTIC_DEFAULT = _os.environ.get("RFNET_DETERMINISTIC", "0") not in ("", "0", "false", "False")


def _nn_distance_setup(ctx, inputs, output):
    xyz1, xyz2 = inputs[0], inputs[1]
    ctx.save_for_backward(xyz1, xyz2, output[1], output[3])


def _nn_distance_backward(ctx, grad_dist1, grad_idx1, grad_dist2, grad_idx2):
    # tf_ops/CD/tf_nndistance.py:26-32: idx grads are ignored; a missing upstream grad is zero
    xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
    if grad_dist1 is None:
        grad_dist1 = torch.zeros(idx1.shape, dtype=torch.float32, device=xyz1.device)
    if grad_dist2 is None:
        grad_dist2 = torch.zeros(idx2.shape, dtype=torch.float32, device=xyz1.device)
    g1, g2 = nn_distance_grad_op(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, DETERMINISTIC_DEFAULT)
    return g1, g2, None, None


nn_distance_op.register_autograd(_nn_distance_backward, setup_context=_nn_distance_setup)


@torch.library.custom_op("rfnet::chamfer_partial_sums", mutates_args=(), device_types="cuda")
def chamfer_partial_sums_op(dist1: torch.Tensor, dist2: torch.Tensor) -> torch.Tensor:
    """[sum sqrt(dist1), numel(dist1), sum sqrt(dist2), numel(dist2)]: the reductions chamfer_big (vv_recon.py:381-385)
    performs with framework ops, in one deterministic pass.  Not differentiable (use rfnet_b200.losses for training)."""
    _require(dist1.dim() == 2 and dist2.dim() == 2 and dist1.shape[0] == dist2.shape[0], "chamfer_partial_sums expects (batch,#points) distances")
    dist1, dist2 = _cuda_f32("dist1", dist1), _cuda_f32("dist2", dist2)
    out = torch.empty((4,), dtype=torch.float32, device=dist1.device)
    lib = _lib.load()
    wsb = lib.rfnet_chamfer_partial_sums_workspace_bytes()
    ws = _workspace(wsb, dist1.device)
    with torch.cuda.device(dist1.device):
        _lib.check(lib.rfnet_chamfer_partial_sums(dist1.shape[0], dist1.shape[1], dist2.shape[1], _ptr(dist1), _ptr(dist2), _ptr(out), _ptr(ws), wsb,
                                                  _stream(dist1)), "rfnet_chamfer_partial_sums")
    return out


@chamfer_partial_sums_op.register_fake
def _(dist1, dist2):
    return dist1.new_empty((4,))


# ------------------------------------------------------------------------------------------------------------ merge_layer
@torch.library.custom_op("rfnet::merge_layer", mutates_args=(), device_types="cuda")
def merge_layer_op(rawpts: torch.Tensor, newpts: torch.Tensor, decfactor: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    # vv_recon.py:132-139 (knum = 1) in one directed search + one epilogue: (refine_pts (b, n_new, 3), idx (b, n_new))
    _require(rawpts.dim() == 3 and rawpts.shape[2] == 3 and newpts.dim() == 3 and newpts.shape[2] == 3, "merge_layer expects (b,n,3) point sets")
    _require(rawpts.shape[0] == newpts.shape[0], "merge_layer expects rawpts and newpts have same batch size")
    _require(rawpts.shape[1] > 0, "merge_layer expects a non-empty raw cloud")
    rawpts, newpts, decfactor = _cuda_f32("rawpts", rawpts), _cuda_f32("newpts", newpts), _cuda_f32("decfactor", decfactor).reshape(-1)
    _require(decfactor.numel() == 1, "merge_layer expects a one-element decfactor (vv_recon.py:211)")
    b, nr, nn = rawpts.shape[0], rawpts.shape[1], newpts.shape[1]
    out = torch.empty_like(newpts)
    idx = torch.empty((b, nn), dtype=torch.int32, device=newpts.device)
    lib = _lib.load()
    wsb = lib.rfnet_merge_layer_workspace_bytes(b, nr, nn)
    ws = _workspace(wsb, newpts.device)
    with torch.cuda.device(newpts.device):
        _lib.check(lib.rfnet_merge_layer(b, nr, _ptr(rawpts), nn, _ptr(newpts), _ptr(decfactor), _ptr(out), _ptr(idx), _ptr(ws), wsb, _stream(newpts)),
                   "rfnet_merge_layer")
    return out, idx


@merge_layer_op.register_fake
def _(rawpts, newpts, decfactor):
    return torch.empty_like(newpts), newpts.new_empty((newpts.shape[0], newpts.shape[1]), dtype=torch.int32)


@torch.library.custom_op("rfnet::merge_layer_grad", mutates_args=(), device_types="cuda")
def merge_layer_grad_op(rawpts: torch.Tensor, newpts: torch.Tensor, decfactor: torch.Tensor, idx: torch.Tensor,
                        grad_out: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    # -> (grad_new (b,n_new,3), grad_raw_rows (b,n_new,3) to be scattered into rawpts by idx, partial sums of d/d decfactor)
    rawpts, newpts, grad_out = _cuda_f32("rawpts", rawpts), _cuda_f32("newpts", newpts), _cuda_f32("grad_out", grad_out)
    decfactor, idx = _cuda_f32("decfactor", decfactor).reshape(-1), _cuda_i32("idx", idx)
    b, nr, nn = rawpts.shape[0], rawpts.shape[1], newpts.shape[1]
    lib = _lib.load()
    g_new, g_rows = torch.empty_like(newpts), torch.empty_like(newpts)
    part = torch.zeros((max(int(lib.rfnet_merge_layer_grad_partials(b, nn)), 1),), dtype=torch.float32, device=newpts.device)
    with torch.cuda.device(newpts.device):
        _lib.check(lib.rfnet_merge_layer_grad(b, nr, _ptr(rawpts), nn, _ptr(newpts), _ptr(idx), _ptr(decfactor), _ptr(grad_out), _ptr(g_new), _ptr(g_rows),
                                              _ptr(part), _stream(newpts)), "rfnet_merge_layer_grad")
    return g_new, g_rows, part


@merge_layer_grad_op.register_fake
def _(rawpts, newpts, decfactor, idx, grad_out):
    return torch.empty_like(newpts), torch.empty_like(newpts), newpts.new_empty((1,))


def _merge_setup(ctx, inputs, output):
    rawpts, newpts, decfactor = inputs
    ctx.dec_shape = decfactor.shape
    ctx.save_for_backward(rawpts, newpts, decfactor, output[1])


def _merge_backward(ctx, grad_out, grad_idx):
    rawpts, newpts, decfactor, idx = ctx.saved_tensors
    g_new, g_rows, part = merge_layer_grad_op(rawpts, newpts, decfactor, idx, grad_out.contiguous())
    # scatter the rows into the raw cloud: GroupPointGrad with one sample per row (the reference's graph: group_point's gradient)
    g_raw = group_point_grad_planned_op(g_rows.unsqueeze(2), scatter_plan_cached(idx, rawpts.shape[1]), rawpts.shape[1])
    return g_raw, g_new, part.sum().reshape(ctx.dec_shape)


merge_layer_op.register_autograd(_merge_backward, setup_context=_merge_setup)


# ------------------------------------------------------------------------------------------------------------ approx_match
@torch.library.custom_op("rfnet::approx_match", mutates_args=(), device_types="cuda")
def approx_match_op(xyz1: torch.Tensor, xyz2: torch.Tensor, flags: int = 0) -> torch.Tensor:
    # ApproxMatchGpuOp::Compute, pc_distance/tf_approxmatch.cpp:145-173.  flags: EMD_EXACT | EMD_NO_PRUNE | EMD_SPLIT_SUMS (rfnet_ops.h)
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "ApproxMatch expects (batch_size,num_points,3) xyz1 shape")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0], "ApproxMatch expects (batch_size,num_points,3) xyz2 shape, and batch_size must match")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    match = torch.empty((b, m, n), dtype=torch.float32, device=xyz1.device)
    lib = _lib.load()
    wsb = lib.rfnet_approxmatch_workspace_bytes(b, n, m)
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_approxmatch(b, n, m, _ptr(xyz1), _ptr(xyz2), _ptr(match), _ptr(ws), wsb, int(flags), _stream(xyz1)), "rfnet_approxmatch")
    return match


@approx_match_op.register_fake
def _(xyz1, xyz2, flags=0):
    return xyz1.new_empty((xyz1.shape[0], xyz2.shape[1], xyz1.shape[1]))


def _check_match_args(op, xyz1, xyz2, match):
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "%s expects (batch_size,num_points,3) xyz1 shape" % op)
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0], "%s expects (batch_size,num_points,3) xyz2 shape, and batch_size must match" % op)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    _require(match.dim() == 3 and tuple(match.shape) == (b, m, n), "%s expects (batch_size,#query,#dataset) match shape" % op)
    return b, n, m


@torch.library.custom_op("rfnet::match_cost", mutates_args=(), device_types="cuda")
def match_cost_op(xyz1: torch.Tensor, xyz2: torch.Tensor, match: torch.Tensor) -> torch.Tensor:
    # MatchCostGpuOp::Compute, pc_distance/tf_approxmatch.cpp:201-229
    b, n, m = _check_match_args("MatchCost", xyz1, xyz2, match)
    xyz1, xyz2, match = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2), _cuda_f32("match", match)
    cost = torch.empty((b,), dtype=torch.float32, device=xyz1.device)
    lib = _lib.load()
    wsb = lib.rfnet_matchcost_workspace_bytes(b, n, m)
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_matchcost(b, n, m, _ptr(xyz1), _ptr(xyz2), _ptr(match), _ptr(cost), _ptr(ws), wsb, _stream(xyz1)), "rfnet_matchcost")
    return cost


@match_cost_op.register_fake
def _(xyz1, xyz2, match):
    return xyz1.new_empty((xyz1.shape[0],))


@torch.library.custom_op("rfnet::match_cost_grad", mutates_args=(), device_types="cuda")
def match_cost_grad_op(xyz1: torch.Tensor, xyz2: torch.Tensor, match: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    # MatchCostGradGpuOp::Compute, pc_distance/tf_approxmatch.cpp:262-294
    b, n, m = _check_match_args("MatchCostGrad", xyz1, xyz2, match)
    xyz1, xyz2, match = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2), _cuda_f32("match", match)
    g1, g2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    lib = _lib.load()
    wsb = lib.rfnet_matchcostgrad_workspace_bytes(b, n, m)
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_matchcostgrad(b, n, m, _ptr(xyz1), _ptr(xyz2), _ptr(match), _ptr(g1), _ptr(g2), _ptr(ws), wsb, _stream(xyz1)),
                   "rfnet_matchcostgrad")
    return g1, g2


@match_cost_grad_op.register_fake
def _(xyz1, xyz2, match):
    return torch.empty_like(xyz1), torch.empty_like(xyz2)


def _match_cost_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _match_cost_backward(ctx, grad_cost):
    # pc_distance/tf_approxmatch.py:44-50: grads scaled by grad_cost[:,None,None]; no gradient to match
    xyz1, xyz2, match = ctx.saved_tensors
    g1, g2 = match_cost_grad_op(xyz1, xyz2, match)
    s = grad_cost[:, None, None]
    return g1 * s, g2 * s, None


match_cost_op.register_autograd(_match_cost_backward, setup_context=_match_cost_setup)


EMD_EXACT = 1        # RFNET_EMD_EXACT: bit-identical to the reference CUDA binary
EMD_NO_PRUNE = 2     # RFNET_EMD_NO_PRUNE
EMD_SPLIT_SUMS = 4   # RFNET_EMD_SPLIT_SUMS: split every sum (more parallelism for tiny batches, different rounding order)
EMD_PRUNE = 8        # RFNET_EMD_PRUNE: exactly-pruned sweeps at the three sharpest levels (identical results)


@torch.library.custom_op("rfnet::emd_cost", mutates_args=(), device_types="cuda")
def emd_cost_op(xyz1: torch.Tensor, xyz2: torch.Tensor, keep_match: bool, flags: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    # ApproxMatch followed by MatchCost (vv_recon.py:396-399) in one call: (cost (b,), match (b,m,n) or an empty tensor).
    # Without keep_match the (b, m, n) matrix never reaches HBM.  Not differentiable: training uses emd_cost_grad_op.
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "ApproxMatch expects (batch_size,num_points,3) xyz1 shape")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0], "ApproxMatch expects (batch_size,num_points,3) xyz2 shape, and batch_size must match")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    cost = torch.empty((b,), dtype=torch.float32, device=xyz1.device)
    match = torch.empty((b, m, n) if keep_match else (0,), dtype=torch.float32, device=xyz1.device)
    lib = _lib.load()
    wsb = lib.rfnet_emd_cost_workspace_bytes(b, n, m)
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_emd_cost(b, n, m, _ptr(xyz1), _ptr(xyz2), _ptr(match) if keep_match else None, _ptr(cost), _ptr(ws), wsb, int(flags),
                                      _stream(xyz1)), "rfnet_emd_cost")
    return cost, match


@emd_cost_op.register_fake
def _(xyz1, xyz2, keep_match, flags=0):
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    return xyz1.new_empty((b,)), xyz1.new_empty((b, m, n) if keep_match else (0,))


@torch.library.custom_op("rfnet::emd_cost_grad", mutates_args=(), device_types="cuda")
def emd_cost_grad_op(xyz1: torch.Tensor, xyz2: torch.Tensor, flags: int = 0) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    # ApproxMatch -> MatchCost together with MatchCostGrad for that match (tf_approxmatch.py:44-50), no (b,m,n) matrix anywhere:
    # (cost (b,), grad1 (b,n,3), grad2 (b,m,3)); the gradients are d cost / d xyz for the match held constant.
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "ApproxMatch expects (batch_size,num_points,3) xyz1 shape")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0], "ApproxMatch expects (batch_size,num_points,3) xyz2 shape, and batch_size must match")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    cost = torch.empty((b,), dtype=torch.float32, device=xyz1.device)
    g1, g2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    lib = _lib.load()
    wsb = lib.rfnet_emd_cost_grad_workspace_bytes(b, n, m)
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_emd_cost_grad(b, n, m, _ptr(xyz1), _ptr(xyz2), _ptr(cost), _ptr(g1), _ptr(g2), _ptr(ws), wsb, int(flags), _stream(xyz1)),
                   "rfnet_emd_cost_grad")
    return cost, g1, g2


@emd_cost_grad_op.register_fake
def _(xyz1, xyz2, flags=0):
    return xyz1.new_empty((xyz1.shape[0],)), torch.empty_like(xyz1), torch.empty_like(xyz2)


def _emd_cost_grad_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1], output[2])


def _emd_cost_grad_backward(ctx, grad_cost, grad_g1, grad_g2):
    # the match is a constant of the loss (NoGradient('ApproxMatch'), tf_approxmatch.py:19); d cost = MatchCostGrad scaled by
    # grad_cost[:,None,None] (tf_approxmatch.py:50).  grad1 / grad2 themselves are not differentiated further.
    g1, g2 = ctx.saved_tensors
    s = grad_cost[:, None, None]
    return g1 * s, g2 * s, None


emd_cost_grad_op.register_autograd(_emd_cost_grad_backward, setup_context=_emd_cost_grad_setup)


# ------------------------------------------------------------------------------------------------------------ scatter plans
@torch.library.custom_op("rfnet::scatter_plan", mutates_args=(), device_types="cuda")
def scatter_plan_op(idx: torch.Tensor, n_targets: int) -> torch.Tensor:
    """The inverted index ("plan") of idx read as (b, rows) targets in [0, n_targets): what the atomic-free gradient scatters
    need, built once -- at forward time, off the backward's critical path -- and shared by every gradient through the same idx."""
    idx = _cuda_i32("idx", idx)
    b = idx.shape[0]
    rows = idx.numel() // max(b, 1)
    lib = _lib.load()
    nbytes = lib.rfnet_scatter_plan_bytes(b, n_targets, rows)
    plan = _workspace(nbytes, idx.device)
    with torch.cuda.device(idx.device):
        _lib.check(lib.rfnet_scatter_plan_build(b, n_targets, rows, _ptr(idx), _ptr(plan), nbytes, _stream(idx)), "rfnet_scatter_plan_build")
    return plan


@scatter_plan_op.register_fake
def _(idx, n_targets):
    return idx.new_empty((1,), dtype=torch.uint8)


@torch.library.custom_op("rfnet::group_point_grad_planned", mutates_args=(), device_types="cuda")
def group_point_grad_planned_op(grad_out: torch.Tensor, plan: torch.Tensor, n: int) -> torch.Tensor:
    # GroupPointGradGpuOp (tf_grouping.cpp:178-212) over a plan of its idx; grad_out (b, m, nsample, c) -> (b, n, c)
    grad_out = _cuda_f32("grad_out", grad_out)
    b, m, ns, c = grad_out.shape
    g = torch.empty((b, n, c), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.load().rfnet_group_point_grad_planned(b, n, c, m, ns, _ptr(grad_out), _ptr(plan), plan.numel(), _ptr(g), _stream(grad_out)),
                   "rfnet_group_point_grad_planned")
    return g


@group_point_grad_planned_op.register_fake
def _(grad_out, plan, n):
    return grad_out.new_empty((grad_out.shape[0], n, grad_out.shape[3]))


@torch.library.custom_op("rfnet::three_interpolate_grad_planned", mutates_args=(), device_types="cuda")
def three_interpolate_grad_planned_op(grad_out: torch.Tensor, weight: torch.Tensor, plan: torch.Tensor, m: int) -> torch.Tensor:
    # ThreeInterpolateGradOp (tf_interpolate.cpp:226-262) over a plan of its idx; grad_out (b, n, c) -> (b, m, c)
    grad_out, weight = _cuda_f32("grad_out", grad_out), _cuda_f32("weight", weight)
    b, n, c = grad_out.shape
    g = torch.empty((b, m, c), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.load().rfnet_three_interpolate_grad_planned(b, n, c, m, _ptr(grad_out), _ptr(weight), _ptr(plan), plan.numel(), _ptr(g),
                                                                    _stream(grad_out)), "rfnet_three_interpolate_grad_planned")
    return g


@three_interpolate_grad_planned_op.register_fake
def _(grad_out, weight, plan, m):
    return grad_out.new_empty((grad_out.shape[0], m, grad_out.shape[2]))


# plans of the idx tensors seen most recently: the xyz and the feature grouping of one layer share one.
# An entry keeps its idx tensor alive, so its address cannot be recycled while the entry exists; _version guards in-place edits.
_PLAN_CACHE = []
_PLAN_CACHE_SIZE = 4


def scatter_plan_cached(idx, n_targets):
    for e in _PLAN_CACHE:
        if e[0] is idx and e[1] == idx._version and e[2] == n_targets:
            return e[3]
    plan = scatter_plan_op(idx.reshape(idx.shape[0], -1), n_targets)
    _PLAN_CACHE.insert(0, (idx, idx._version, n_targets, plan))
    del _PLAN_CACHE[_PLAN_CACHE_SIZE:]
    return plan


# ------------------------------------------------------------------------------------------------------------ sampling
@torch.library.custom_op("rfnet::farthest_point_sample", mutates_args=(), device_types="cuda")
def farthest_point_sample_op(inp: torch.Tensor, npoint: int, pruned: bool = True) -> torch.Tensor:
    # FarthestPointSampleGpuOp, tf_ops/sampling/tf_sampling.cpp:95-123.  pruned=False: no workspace -> the cluster kernel (same indices)
    _require(npoint > 0, "FarthestPointSample expects positive npoint")
    _require(inp.dim() == 3 and inp.shape[2] == 3, "FarthestPointSample expects (batch_size,num_points,3) inp shape")
    inp = _cuda_f32("inp", inp)
    b, n = inp.shape[0], inp.shape[1]
    out = torch.empty((b, npoint), dtype=torch.int32, device=inp.device)
    lib = _lib.load()
    wsb = lib.rfnet_farthestpointsampling_workspace_bytes(b, n, npoint)
    if not pruned and n <= 32768:
        wsb = 0
    ws = _workspace(wsb, inp.device)
    with torch.cuda.device(inp.device):
        _lib.check(lib.rfnet_farthestpointsampling(b, n, npoint, _ptr(inp), _ptr(ws) if wsb else _vp(0), wsb, _ptr(out), _stream(inp)),
                   "rfnet_farthestpointsampling")
    return out


@farthest_point_sample_op.register_fake
def _(inp, npoint, pruned=True):
    return inp.new_empty((inp.shape[0], npoint), dtype=torch.int32)


@torch.library.custom_op("rfnet::gather_point", mutates_args=(), device_types="cuda")
def gather_point_op(inp: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    # GatherPointGpuOp, tf_ops/sampling/tf_sampling.cpp:126-148
    _require(inp.dim() == 3 and inp.shape[2] == 3, "GatherPoint expects (batch_size,num_points,3) inp shape")
    _require(idx.dim() == 2 and idx.shape[0] == inp.shape[0], "GatherPoint expects (batch_size,num_result) idx shape")
    inp, idx = _cuda_f32("inp", inp), _cuda_i32("idx", idx)
    b, n, m = inp.shape[0], inp.shape[1], idx.shape[1]
    out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
    with torch.cuda.device(inp.device):
        _lib.check(_lib.load().rfnet_gatherpoint(b, n, m, _ptr(inp), _ptr(idx), _ptr(out), _stream(inp)), "rfnet_gatherpoint")
    return out


@gather_point_op.register_fake
def _(inp, idx):
    return inp.new_empty((inp.shape[0], idx.shape[1], 3))


@torch.library.custom_op("rfnet::gather_point_grad", mutates_args=(), device_types="cuda")
def gather_point_grad_op(inp: torch.Tensor, idx: torch.Tensor, out_g: torch.Tensor) -> torch.Tensor:
    # GatherPointGradGpuOp, tf_ops/sampling/tf_sampling.cpp:151-178
    _require(inp.dim() == 3 and inp.shape[2] == 3, "GatherPointGradGpuOp expects (batch_size,num_points,3) inp")
    _require(idx.dim() == 2 and idx.shape[0] == inp.shape[0], "GatherPointGradGpuOp expects (batch_size,num_result) idx shape")
    b, n, m = inp.shape[0], inp.shape[1], idx.shape[1]
    _require(out_g.dim() == 3 and tuple(out_g.shape) == (b, m, 3), "GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape")
    idx, out_g = _cuda_i32("idx", idx), _cuda_f32("out_g", out_g)
    inp_g = torch.empty((b, n, 3), dtype=torch.float32, device=out_g.device)
    lib = _lib.load()
    wsb = lib.rfnet_scatteraddpoint_workspace_bytes(b, n, m)
    ws = _workspace(wsb, out_g.device)
    with torch.cuda.device(out_g.device):
        _lib.check(lib.rfnet_scatteraddpoint(b, n, m, _ptr(out_g), _ptr(idx), _ptr(inp_g), _ptr(ws) if wsb else _vp(0), wsb, _stream(out_g)),
                   "rfnet_scatteraddpoint")
    return inp_g


@gather_point_grad_op.register_fake
def _(inp, idx, out_g):
    return torch.empty_like(inp)


def _gather_setup(ctx, inputs, output):
    inp, idx = inputs
    ctx.n = inp.shape[1]
    ctx.save_for_backward(scatter_plan_cached(idx, inp.shape[1]))   # the inverted index, built while the forward runs


def _gather_backward(ctx, out_g):
    (plan,) = ctx.saved_tensors  # tf_ops/sampling/tf_sampling.py:43-47; gather_point's gradient is group_point's with nsample = 1
    return group_point_grad_planned_op(out_g.contiguous().unsqueeze(2), plan, ctx.n), None


gather_point_op.register_autograd(_gather_backward, setup_context=_gather_setup)


# ------------------------------------------------------------------------------------------------------------ grouping
@torch.library.custom_op("rfnet::query_ball_point", mutates_args=(), device_types="cuda")
def query_ball_point_op(xyz1: torch.Tensor, xyz2: torch.Tensor, radius: torch.Tensor, nsample: int, grid: bool = True) -> tuple[torch.Tensor, torch.Tensor]:
    # QueryBallPointGpuOp, tf_ops/grouping/tf_grouping.cpp:68-110 -- radius is a tensor input there too (:93-95).
    # grid=False: no workspace -> every query scans the whole dataset (same results)
    _require(nsample > 0, "QueryBallPoint expects positive nsample")
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3, "QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape.")
    _require(xyz2.shape[0] == xyz1.shape[0], "QueryBallPoint expects xyz1 and xyz2 have same batch size")
    xyz1, xyz2, radius = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2), _cuda_f32("radius", radius)
    _require(radius.numel() >= 1, "QueryBallPoint expects a radius")
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz1.device)
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    lib = _lib.load()
    wsb = lib.rfnet_query_ball_point_workspace_bytes(b, n, m) if grid else 0
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_query_ball_point(b, n, m, _ptr(radius), nsample, _ptr(xyz1), _ptr(xyz2), _ptr(idx), _ptr(cnt), _ptr(ws) if wsb else _vp(0), wsb,
                                              _stream(xyz1)), "rfnet_query_ball_point")
    return idx, cnt


@query_ball_point_op.register_fake
def _(xyz1, xyz2, radius, nsample, grid=True):
    b, m = xyz2.shape[0], xyz2.shape[1]
    return xyz1.new_empty((b, m, nsample), dtype=torch.int32), xyz1.new_empty((b, m), dtype=torch.int32)


@torch.library.custom_op("rfnet::group_point", mutates_args=(), device_types="cuda")
def group_point_op(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    # GroupPointGpuOp, tf_ops/grouping/tf_grouping.cpp:147-175
    _require(points.dim() == 3, "GroupPoint expects (batch_size, num_points, channel) points shape")
    _require(idx.dim() == 3 and idx.shape[0] == points.shape[0], "GroupPoint expects (batch_size, npoints, nsample) idx shape")
    points, idx = _cuda_f32("points", points), _cuda_i32("idx", idx)
    b, n, c = points.shape
    m, ns = idx.shape[1], idx.shape[2]
    out = torch.empty((b, m, ns, c), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().rfnet_group_point(b, n, c, m, ns, _ptr(points), _ptr(idx), _ptr(out), _stream(points)), "rfnet_group_point")
    return out


@group_point_op.register_fake
def _(points, idx):
    return points.new_empty((points.shape[0], idx.shape[1], idx.shape[2], points.shape[2]))


@torch.library.custom_op("rfnet::group_point_grad", mutates_args=(), device_types="cuda")
def group_point_grad_op(points: torch.Tensor, idx: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    # GroupPointGradGpuOp, tf_ops/grouping/tf_grouping.cpp:178-212
    _require(points.dim() == 3, "GroupPointGrad expects (batch_size, num_points, channel) points shape")
    _require(idx.dim() == 3 and idx.shape[0] == points.shape[0], "GroupPointGrad expects (batch_size, npoints, nsample) idx shape")
    b, n, c = points.shape
    m, ns = idx.shape[1], idx.shape[2]
    _require(grad_out.dim() == 4 and tuple(grad_out.shape) == (b, m, ns, c), "GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape")
    idx, grad_out = _cuda_i32("idx", idx), _cuda_f32("grad_out", grad_out)
    g = torch.empty((b, n, c), dtype=torch.float32, device=grad_out.device)
    lib = _lib.load()
    wsb = lib.rfnet_group_point_grad_workspace_bytes(b, n, c, m, ns)
    ws = _workspace(wsb, grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(lib.rfnet_group_point_grad(b, n, c, m, ns, _ptr(grad_out), _ptr(idx), _ptr(g), _ptr(ws) if wsb else _vp(0), wsb, _stream(grad_out)),
                   "rfnet_group_point_grad")
    return g


@group_point_grad_op.register_fake
def _(points, idx, grad_out):
    return torch.empty_like(points)


def _group_setup(ctx, inputs, output):
    points, idx = inputs
    ctx.n = points.shape[1]
    ctx.save_for_backward(scatter_plan_cached(idx, points.shape[1]))   # the inverted index, built while the forward runs


def _group_backward(ctx, grad_out):
    (plan,) = ctx.saved_tensors  # tf_ops/grouping/tf_grouping.py:42-46
    return group_point_grad_planned_op(grad_out.contiguous(), plan, ctx.n), None


group_point_op.register_autograd(_group_backward, setup_context=_group_setup)


@torch.library.custom_op("rfnet::knn_point", mutates_args=(), device_types="cuda")
def knn_point_op(xyz1: torch.Tensor, xyz2: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    """3-d k-nearest neighbours without materialising the (b,m,n) matrix; outputs as tf.nn.top_k(-dist) (tf_grouping.py:48-73)."""
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3 and xyz2.dim() == 3 and xyz2.shape[2] == 3, "knn_point kernel expects (b,n,3) and (b,m,3)")
    _require(xyz2.shape[0] == xyz1.shape[0], "knn_point expects xyz1 and xyz2 have same batch size")
    _require(0 < k <= 32 and k <= xyz1.shape[1], "knn_point kernel expects 0 < k <= min(32, ndataset)")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    val = torch.empty((b, m, k), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, m, k), dtype=torch.int32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(_lib.load().rfnet_knn_point(b, n, m, k, _ptr(xyz1), _ptr(xyz2), _ptr(val), _ptr(idx), _stream(xyz1)), "rfnet_knn_point")
    return val, idx


@knn_point_op.register_fake
def _(xyz1, xyz2, k):
    b, m = xyz2.shape[0], xyz2.shape[1]
    return xyz1.new_empty((b, m, k)), xyz1.new_empty((b, m, k), dtype=torch.int32)


def _knn_setup(ctx, inputs, output):
    xyz1, xyz2, _k = inputs
    ctx.save_for_backward(xyz1, xyz2, output[1])


def _knn_backward(ctx, grad_val, grad_idx):
    # val[b,j,t] = -|xyz1[b, idx[b,j,t]] - xyz2[b,j]|^2  (the reference's top_k(-dist) is differentiable in the same way):
    # d val / d xyz2[j] = +2 (x - q), d val / d xyz1[idx] = -2 (x - q); the scatter into xyz1 is group_point's gradient kernel
    xyz1, xyz2, idx = ctx.saved_tensors
    diff = group_point_op(xyz1, idx) - xyz2.unsqueeze(2)                 # (b, m, k, 3)
    w = (2.0 * grad_val).unsqueeze(-1) * diff
    g1 = group_point_grad_planned_op((-w).contiguous(), scatter_plan_cached(idx, xyz1.shape[1]), xyz1.shape[1])
    return g1, w.sum(2), None


knn_point_op.register_autograd(_knn_backward, setup_context=_knn_setup)


@torch.library.custom_op("rfnet::selection_sort", mutates_args=(), device_types="cuda")
def selection_sort_op(dist: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    # SelectionSortGpuOp, tf_ops/grouping/tf_grouping.cpp:113-143
    _require(k > 0, "SelectionSort expects positive k")
    _require(dist.dim() == 3, "SelectionSort expects (b,m,n) dist shape.")
    dist = _cuda_f32("dist", dist)
    b, m, n = dist.shape
    outi = torch.empty((b, m, n), dtype=torch.int32, device=dist.device)
    out = torch.empty((b, m, n), dtype=torch.float32, device=dist.device)
    with torch.cuda.device(dist.device):
        _lib.check(_lib.load().rfnet_selection_sort(b, n, m, k, _ptr(dist), _ptr(outi), _ptr(out), _stream(dist)), "rfnet_selection_sort")
    return outi, out


@selection_sort_op.register_fake
def _(dist, k):
    return dist.new_empty(dist.shape, dtype=torch.int32), dist.new_empty(dist.shape)


# ------------------------------------------------------------------------------------------------------------ auction_match
AUCTION_MAX_POINTS = 8192  # RFNET_AUCTION_MAX_POINTS (the reference: 4096, tf_auctionmatch.cpp:37)


@torch.library.custom_op("rfnet::auction_match", mutates_args=(), device_types="cuda")
def auction_match_op(xyz1: torch.Tensor, xyz2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    # AuctionMatchGpuOp::Compute, tf_ops/emd/tf_auctionmatch.cpp:28-59 (its messages say "ApproxMatch"; kept)
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "ApproxMatch expects (batch_size,num_points,3) xyz1 shape")
    _require(xyz1.shape[1] <= AUCTION_MAX_POINTS, "AuctionMatch handles at most %d dataset points" % AUCTION_MAX_POINTS)
    _require(xyz2.dim() == 3 and tuple(xyz2.shape) == tuple(xyz1.shape),
             "AuctionMatch expects (batch_size,num_points,3) xyz2 shape, and shape must match with xyz1")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    b, n = xyz1.shape[0], xyz1.shape[1]
    matchl = torch.empty((b, n), dtype=torch.int32, device=xyz1.device)
    matchr = torch.empty((b, n), dtype=torch.int32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(_lib.load().rfnet_auction_match(b, n, _ptr(xyz1), _ptr(xyz2), _ptr(matchl), _ptr(matchr), _stream(xyz1)), "rfnet_auction_match")
    return matchl, matchr


@auction_match_op.register_fake
def _(xyz1, xyz2):
    b, n = xyz1.shape[0], xyz1.shape[1]
    return xyz1.new_empty((b, n), dtype=torch.int32), xyz1.new_empty((b, n), dtype=torch.int32)


# ------------------------------------------------------------------------------------------------------------ interpolation
@torch.library.custom_op("rfnet::three_nn", mutates_args=(), device_types="cuda")
def three_nn_op(xyz1: torch.Tensor, xyz2: torch.Tensor, grid: bool = True) -> tuple[torch.Tensor, torch.Tensor]:
    # ThreeNNOp, tf_ops/interpolation/tf_interpolate.cpp:157-187.  grid=False: no workspace -> the scan kernel (same results)
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "ThreeNN expects (b,n,3) xyz1 shape")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3, "ThreeNN expects (b,m,3) xyz2 shape")
    _require(xyz2.shape[0] == xyz1.shape[0], "ThreeNN expects xyz1 and xyz2 have same batch size")
    xyz1, xyz2 = _cuda_f32("xyz1", xyz1), _cuda_f32("xyz2", xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    dist = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=xyz1.device)
    lib = _lib.load()
    wsb = lib.rfnet_three_nn_workspace_bytes(b, n, m) if grid else 0
    ws = _workspace(wsb, xyz1.device)
    with torch.cuda.device(xyz1.device):
        _lib.check(lib.rfnet_three_nn(b, n, m, _ptr(xyz1), _ptr(xyz2), _ptr(dist), _ptr(idx), _ptr(ws) if wsb else _vp(0), wsb, _stream(xyz1)), "rfnet_three_nn")
    return dist, idx


@three_nn_op.register_fake
def _(xyz1, xyz2, grid=True):
    b, n = xyz1.shape[0], xyz1.shape[1]
    return xyz1.new_empty((b, n, 3)), xyz1.new_empty((b, n, 3), dtype=torch.int32)


def _check_interp(op, points, idx, weight):
    _require(points.dim() == 3, "%s expects (b,m,c) points shape" % op)
    b = points.shape[0]
    _require(idx.dim() == 3 and idx.shape[0] == b and idx.shape[2] == 3, "%s expects (b,n,3) idx shape" % op)
    _require(weight.dim() == 3 and tuple(weight.shape) == tuple(idx.shape), "%s expects (b,n,3) weight shape" % op)


@torch.library.custom_op("rfnet::three_interpolate", mutates_args=(), device_types="cuda")
def three_interpolate_op(points: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    # ThreeInterpolateOp, tf_ops/interpolation/tf_interpolate.cpp:191-222
    _check_interp("ThreeInterpolate", points, idx, weight)
    points, idx, weight = _cuda_f32("points", points), _cuda_i32("idx", idx), _cuda_f32("weight", weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().rfnet_three_interpolate(b, m, c, n, _ptr(points), _ptr(idx), _ptr(weight), _ptr(out), _stream(points)),
                   "rfnet_three_interpolate")
    return out


@three_interpolate_op.register_fake
def _(points, idx, weight):
    return points.new_empty((points.shape[0], idx.shape[1], points.shape[2]))


@torch.library.custom_op("rfnet::three_interpolate_grad", mutates_args=(), device_types="cuda")
def three_interpolate_grad_op(points: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    # ThreeInterpolateGradOp, tf_ops/interpolation/tf_interpolate.cpp:226-262
    _check_interp("ThreeInterpolateGrad", points, idx, weight)
    b, m, c = points.shape
    n = idx.shape[1]
    _require(grad_out.dim() == 3 and tuple(grad_out.shape) == (b, n, c), "ThreeInterpolateGrad expects (b,n,c) grad_out shape")
    idx, weight, grad_out = _cuda_i32("idx", idx), _cuda_f32("weight", weight), _cuda_f32("grad_out", grad_out)
    g = torch.empty((b, m, c), dtype=torch.float32, device=grad_out.device)
    lib = _lib.load()
    wsb = lib.rfnet_three_interpolate_grad_workspace_bytes(b, n, c, m)
    ws = _workspace(wsb, grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(lib.rfnet_three_interpolate_grad(b, n, c, m, _ptr(grad_out), _ptr(idx), _ptr(weight), _ptr(g), _ptr(ws) if wsb else _vp(0), wsb,
                                                    _stream(grad_out)), "rfnet_three_interpolate_grad")
    return g


@three_interpolate_grad_op.register_fake
def _(points, idx, weight, grad_out):
    return torch.empty_like(points)


def _interp_setup(ctx, inputs, output):
    points, idx, weight = inputs
    ctx.m = points.shape[1]
    ctx.save_for_backward(weight, scatter_plan_cached(idx, points.shape[1]))   # the inverted index, built while the forward runs


def _interp_backward(ctx, grad_out):
    weight, plan = ctx.saved_tensors  # tf_ops/interpolation/tf_interpolate.py:29-34
    return three_interpolate_grad_planned_op(grad_out.contiguous(), weight, plan, ctx.m), None, None


three_interpolate_op.register_autograd(_interp_backward, setup_context=_interp_setup)
