"""Drop-in for the reference's tf_ops/grouping/tf_grouping.py."""
import torch

from . import ops


def query_ball_point(radius, nsample, xyz1, xyz2):
    '''
    Input:
        radius: float32, ball search radius (python float or 1-element tensor)
        nsample: int32, number of points selected in each ball region
        xyz1: (batch_size, ndataset, 3) float32 array, input points
        xyz2: (batch_size, npoint, 3) float32 array, query points
    Output:
        idx: (batch_size, npoint, nsample) int32 array, indices to input points
        pts_cnt: (batch_size, npoint) int32 array, number of unique points in each local region

    Rows with no point inside the ball hold 0 (the reference leaves them uninitialised) and pts_cnt 0.
    '''
    if not isinstance(radius, torch.Tensor):
        radius = torch.tensor([float(radius)], dtype=torch.float32, device=xyz1.device)
    return ops.query_ball_point_op(xyz1.detach(), xyz2.detach(), radius.detach().reshape(-1), int(nsample))


def select_top_k(k, dist):
    '''
    Input:
        k: int32, number of k SMALLEST elements selected
        dist: (b,m,n) float32 array, distance matrix, m query points, n dataset points
    Output:
        idx: (b,m,n) int32 array, first k in n are indices to the top k
        dist_out: (b,m,n) float32 array, first k in n are the top k

    No gradient (ops.NoGradient('SelectionSort'), tf_grouping.py:32).
    '''
    return ops.selection_sort_op(dist.detach(), int(k))


def group_point(points, idx):
    '''
    Input:
        points: (batch_size, ndataset, channel) float32 array, points to sample from
        idx: (batch_size, npoint, nsample) int32 array, indices to points
    Output:
        out: (batch_size, npoint, nsample, channel) float32 array, values sampled from points
    '''
    return ops.group_point_op(points, idx)


def knn_point(k, xyz1, xyz2):
    '''
    Input:
        k: int32, number of k in k-nn search
        xyz1: (batch_size, ndataset, c) float32 array, input points
        xyz2: (batch_size, npoint, c) float32 array, query points
    Output:
        val: (batch_size, npoint, k) float32 array, NEGATED squared L2 distances (the reference returns top_k(-dist))
        idx: (batch_size, npoint, k) int32 array, indices to input points

    The reference implements this with framework ops only (tf.nn.top_k on a materialised (b,m,n) matrix, "ONLY SUPPORT
    CPU", tf_grouping.py:64-73).  Here it is the rfnet::knn_point kernel, which never builds the matrix: 3-d float32 CUDA
    points and k <= 32 (what the model uses).  Anything else is an error -- there is no framework fallback.
    val is differentiable w.r.t. both clouds (as top_k(-dist) is in the reference); ties keep the lowest index first.
    '''
    if xyz1.shape[-1] != 3 or xyz2.shape[-1] != 3:
        raise ValueError("knn_point: the kernel handles 3-d points (got %d-d); the reference's framework formulation for feature "
                         "vectors is not reproduced" % xyz1.shape[-1])
    if not 0 < int(k) <= min(32, xyz1.shape[1]):
        raise ValueError("knn_point: the kernel handles 0 < k <= min(32, ndataset) (got k=%d, ndataset=%d)" % (int(k), xyz1.shape[1]))
    return ops.knn_point_op(xyz1, xyz2, int(k))
