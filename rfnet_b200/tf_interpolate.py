"""Drop-in for the reference's tf_ops/interpolation/tf_interpolate.py."""
from . import ops


def three_nn(xyz1, xyz2):
    '''
    Input:
        xyz1: (b,n,3) float32 array, unknown points
        xyz2: (b,m,3) float32 array, known points
    Output:
        dist: (b,n,3) float32 array, distances to known points (SQUARED, ascending)
        idx: (b,n,3) int32 array, indices to known points
    '''
    return ops.three_nn_op(xyz1.detach(), xyz2.detach())


def three_interpolate(points, idx, weight):
    '''
    Input:
        points: (b,m,c) float32 array, known points
        idx: (b,n,3) int32 array, indices to known points
        weight: (b,n,3) float32 array, weights on known points
    Output:
        out: (b,n,c) float32 array, interpolated point values

    Differentiable w.r.t. points only (tf_interpolate.py:29-34).
    '''
    return ops.three_interpolate_op(points, idx, weight.detach())
