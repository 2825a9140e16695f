"""Drop-in for the reference's tf_ops/CD/tf_nndistance.py (== pc_distance/tf_nndistance.py)."""
from . import ops


def nn_distance(xyz1, xyz2, unfused=False, direct=False):
    '''
Computes the distance of nearest neighbors for a pair of point clouds
input: xyz1: (batch_size,#points_1,3)  the first point cloud
input: xyz2: (batch_size,#points_2,3)  the second point cloud
output: dist1: (batch_size,#point_1)   distance from first to second
output: idx1:  (batch_size,#point_1)   nearest neighbor from first to second
output: dist2: (batch_size,#point_2)   distance from second to first
output: idx2:  (batch_size,#point_2)   nearest neighbor from second to first

Distances are SQUARED, indices int32, ties go to the lowest index (tf_ops/CD/tf_nndistance_g.cu:28,38,118).
Differentiable w.r.t. xyz1 and xyz2 exactly as the reference's RegisterGradient('NnDistance') (tf_nndistance.py:26-32).
`unfused=True` evaluates d2 as the reference's CPU kernel does (no FMA contraction); the default is the GPU contract.
`direct=True` evaluates that expression for every pair (nn_search_kernel) instead of the default filtered search, which scans
an expanded form at half the arithmetic and evaluates the reference expression only where it decides the result -- same
outputs bit for bit; worth setting only for clouds full of DISTINCT points at exactly equal distances (lattices).
    '''
    return ops.nn_distance_op(xyz1, xyz2, unfused, direct)
