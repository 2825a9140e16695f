"""rfnet_b200 -- B200 (sm_100a) implementation of RFNet's point-cloud operators behind the reference's Python API.

Drop-in modules (same function names, argument orders and output tuples as the reference's TF1 wrappers):

    rfnet_b200.tf_nndistance   nn_distance                                   (tf_ops/CD/tf_nndistance.py, pc_distance/tf_nndistance.py)
    rfnet_b200.tf_approxmatch  approx_match, match_cost (+ fused emd_cost)   (pc_distance/tf_approxmatch.py)
    rfnet_b200.tf_auctionmatch auction_match                                 (tf_ops/emd/tf_auctionmatch.py)
    rfnet_b200.tf_sampling     farthest_point_sample, gather_point           (tf_ops/sampling/tf_sampling.py)
    rfnet_b200.tf_grouping     query_ball_point, group_point, knn_point,
                               select_top_k                                  (tf_ops/grouping/tf_grouping.py)
    rfnet_b200.tf_interpolate  three_nn, three_interpolate                   (tf_ops/interpolation/tf_interpolate.py)
    rfnet_b200.losses          chamfer_big, fidelity_loss, earth_mover,
                               emd_func, re_chamfer, merge_layer, ...        (vv_recon.py:365-419) + batch-sharded variants

Tensors are torch CUDA tensors.  All compute happens in librfnet_ops.so (C ABI: include/rfnet_ops.h); importing the op
modules fails loudly if that library has not been built -- there is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
