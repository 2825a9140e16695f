"""The loss-level callers of the operators (vv_recon.py:381-399), plus their batch-sharded form.

``chamfer_big``, ``fidelity_loss`` and ``earth_mover`` are the reference's definitions on top of the drop-in ops.  The
``sharded_*`` functions are what one rank of a data-parallel job calls: every op is independent per cloud, so each rank
evaluates its own slice of the batch and the ONLY exchange is a sum of a few fp32 scalars (torch.distributed all_reduce:
NCCL over NVLink on GPUs, gloo in the CPU tests).  Gradients need no collective: d(loss)/d(points) is per cloud and the
1/(global batch) factor is applied locally.
"""
import torch
import torch.distributed as dist

from . import tf_approxmatch, tf_auctionmatch, tf_grouping, tf_nndistance, tf_sampling


def chamfer_big(pcd1, pcd2):
    """vv_recon.py:381-385 -> ((mean sqrt(dist1) + mean sqrt(dist2)) / 2, idx1)."""
    dist1, idx1, dist2, idx2 = tf_nndistance.nn_distance(pcd1, pcd2)
    d1 = torch.mean(torch.sqrt(dist1))
    d2 = torch.mean(torch.sqrt(dist2))
    return (d1 + d2) / 2, idx1


def fidelity_loss(pcd1, pcd2):
    """vv_recon.py:386-390 -> mean sqrt(dist1)."""
    dist1, _, _, _ = tf_nndistance.nn_distance(pcd1, pcd2)
    return torch.mean(torch.sqrt(dist1))


def earth_mover(pcd1, pcd2):
    """vv_recon.py:392-399 -> mean(cost / num_points); requires equal point counts."""
    assert pcd1.shape[1] == pcd2.shape[1]
    num_points = float(pcd1.shape[1])
    cost = tf_approxmatch.emd_cost(pcd1, pcd2)  # approx_match + match_cost in one call; match kept only for backward
    return torch.mean(cost / num_points)


def emd_func(pred, gt):
    """vv_recon.py:365-380: auction_match -> gather the matched GT points -> mean matched distance, normalised by the
    radius of the prediction around its centroid, averaged over the batch."""
    matchl_out, _ = tf_auctionmatch.auction_match(pred, gt)
    matched_out = tf_sampling.gather_point(gt, matchl_out)
    dist = torch.sqrt(((pred - matched_out) ** 2).sum(-1)).mean(-1)
    cens = pred.mean(dim=1, keepdim=True)
    radius = torch.sqrt(((pred - cens) ** 2).sum(-1).max(dim=-1).values)
    return (dist / radius).mean()


# ----------------------------------------------------------------------------------------------------------------------
# The other callers of the operators in the reference's model/loss code, as thin torch functions
# over the drop-in ops -- same arithmetic, same argument meaning.
# ----------------------------------------------------------------------------------------------------------------------
def sampling(npoint, xyz):
    """vv_recon.py:67-70 (use_type='f'): farthest point sampling followed by the gather of the picked points."""
    idx = tf_sampling.farthest_point_sample(npoint, xyz)
    return tf_sampling.gather_point(xyz, idx)


def merge_layer(rawpts, newpts, decfactor):
    """vv_recon.py:132-139: pull every new point towards its nearest raw point with weight exp(-d2 / (1e-8 + decfactor^2)).
    One fused operator (rfnet::merge_layer: a single directed nearest-neighbour search + one epilogue kernel) instead of the
    reference's NnDistance + GroupPoint + five framework ops; differentiable w.r.t. both clouds and decfactor."""
    from . import ops
    if not isinstance(decfactor, torch.Tensor):
        decfactor = torch.tensor([float(decfactor)], dtype=torch.float32, device=newpts.device)
    out, _ = ops.merge_layer_op(rawpts, newpts, decfactor)
    return out


def re_chamfer(gt, pred, part=8):
    """vv_recon.py:171-193: mean of chamfer_big over `part` equal slices of the point axis (pred slice vs gt slice)."""
    interval = gt.shape[1] // 8                                                     # the reference divides by 8, not by `part`
    total = 0.0
    for i in range(part):
        sl = slice(i * interval, (i + 1) * interval)
        total = total + chamfer_big(pred[:, sl].contiguous(), gt[:, sl].contiguous())[0]
    return total / part


def zero_groupnear(ptcens, rawpts, outmat):
    """vv_recon.py:414-419: relu(mean |outmat|^2 - 0.4 * mean nn-dist2(rawpts -> ptcens))."""
    _, _, dist, _ = tf_nndistance.nn_distance(ptcens, rawpts)
    inval = dist.mean()
    outval = (outmat * outmat).sum(-1).mean(-1).mean(-1).mean()
    return torch.relu(outval - 0.4 * inval)


def shard_bounds(global_batch, rank, world_size):
    """Contiguous slice [lo, hi) of the batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(global_batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_scalars(values, group=None):
    """Sum a 1-D tensor of partial sums over all ranks (in place when distributed is initialised)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group)
    return values


def chamfer_partial_sums(dist1, dist2):
    """[sum sqrt(dist1), count1, sum sqrt(dist2), count2] of a local shard, as one fp32 tensor."""
    return torch.stack([torch.sqrt(dist1).sum(), dist1.new_tensor(float(dist1.numel())),
                        torch.sqrt(dist2).sum(), dist2.new_tensor(float(dist2.numel()))])


def sharded_chamfer_big(pcd1_local, pcd2_local, group=None):
    """chamfer_big over a batch that is sharded across ranks: local nn_distance, then one 16-byte all-reduce.
    The returned loss is the GLOBAL value; its gradient w.r.t. the local clouds is the global loss's gradient."""
    dist1, idx1, dist2, idx2 = tf_nndistance.nn_distance(pcd1_local, pcd2_local)
    s1, s2 = torch.sqrt(dist1).sum(), torch.sqrt(dist2).sum()
    totals = all_reduce_scalars(torch.stack([s1.detach(), dist1.new_tensor(float(dist1.numel())), s2.detach(),
                                             dist2.new_tensor(float(dist2.numel()))]), group)
    n1, n2 = totals[1], totals[3]
    # value = global mean; gradient = local sum / global count  (straight-through on the reduced scalar)
    loss = ((totals[0] + (s1 - s1.detach())) / n1 + (totals[2] + (s2 - s2.detach())) / n2) / 2
    return loss, idx1


def sharded_earth_mover(pcd1_local, pcd2_local, group=None):
    """earth_mover over a sharded batch: local approx_match + match_cost, then one 8-byte all-reduce."""
    assert pcd1_local.shape[1] == pcd2_local.shape[1]
    num_points = float(pcd1_local.shape[1])
    cost = tf_approxmatch.emd_cost(pcd1_local, pcd2_local)
    s = (cost / num_points).sum()
    totals = all_reduce_scalars(torch.stack([s.detach(), cost.new_tensor(float(cost.numel()))]), group)
    return (totals[0] + (s - s.detach())) / totals[1]
