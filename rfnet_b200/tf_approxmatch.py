"""Drop-in for the reference's pc_distance/tf_approxmatch.py."""
from . import ops


def approx_match(xyz1, xyz2, exact=False):
    '''
input:
	xyz1 : batch_size * #dataset_points * 3
	xyz2 : batch_size * #query_points * 3
returns:
	match : batch_size * #query_points * #dataset_points

No gradient flows through approx_match (ops.NoGradient('ApproxMatch'), tf_approxmatch.py:19).
Every sum follows the reference CUDA kernel's order; exact=True also uses its non-flushing exponential everywhere, which
makes `match` bit-identical to the reference binary's (the parity mode, ~15 % slower).
    '''
    return ops.approx_match_op(xyz1.detach(), xyz2.detach(), ops.EMD_EXACT if exact else 0)


def match_cost(xyz1, xyz2, match):
    '''
input:
	xyz1 : batch_size * #dataset_points * 3
	xyz2 : batch_size * #query_points * 3
	match : batch_size * #query_points * #dataset_points
returns:
	cost : batch_size

Differentiable w.r.t. xyz1 and xyz2 (not match), as RegisterGradient('MatchCost') (tf_approxmatch.py:44-50).
    '''
    return ops.match_cost_op(xyz1, xyz2, match.detach())


def emd_cost(xyz1, xyz2, exact=False):
    '''
approx_match followed by match_cost, as every caller in the reference chains them (vv_recon.py:396-399), in one call:
	cost : batch_size
The (b, m, n) match matrix is never written to memory.  When a gradient can be asked for, cost and both MatchCostGrad
gradients come out of the same call (two passes that rebuild the matrix entries in registers); values equal
match_cost(xyz1, xyz2, approx_match(xyz1, xyz2)) and its gradient up to summation order.
    '''
    import torch
    flags = ops.EMD_EXACT if exact else 0
    if torch.is_grad_enabled() and (xyz1.requires_grad or xyz2.requires_grad):
        cost, _, _ = ops.emd_cost_grad_op(xyz1, xyz2, flags)
        return cost
    cost, _ = ops.emd_cost_op(xyz1, xyz2, False, flags)
    return cost
