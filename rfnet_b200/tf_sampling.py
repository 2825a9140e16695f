"""Drop-in for the reference's tf_ops/sampling/tf_sampling.py."""
from . import ops


def gather_point(inp, idx):
    '''
input:
    batch_size * ndataset * 3   float32
    batch_size * npoints        int32
returns:
    batch_size * npoints * 3    float32
    '''
    return ops.gather_point_op(inp, idx)


def farthest_point_sample(npoint, inp):
    '''
input:
    int32
    batch_size * ndataset * 3   float32
returns:
    batch_size * npoint         int32

Note the argument order (npoint first), as in the reference (tf_sampling.py:48-56).  Not differentiable.
    '''
    return ops.farthest_point_sample_op(inp.detach(), int(npoint))
