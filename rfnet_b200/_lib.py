"""ctypes binding of librfnet_ops.so (the C ABI declared in include/rfnet_ops.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librfnet_ops.so")

_i, _z, _p = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/rfnet_ops.h one to one (tests/test_abi.py checks the two agree).
SIGNATURES = {
    "rfnet_version": (_i, []),
    "rfnet_error_string": (ctypes.c_char_p, [_i]),
    "rfnet_nn_distance_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_nn_distance": (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _z, _i, _p]),
    "rfnet_nn_distance_stats": (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _z, _i, _p, _p]),
    "rfnet_nn_distance_plan": (_i, [_i, _i, _i, _i, _p]),
    "rfnet_nn_distance_grad_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_nn_distance_grad": (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _z, _p]),
    "rfnet_chamfer_partial_sums_workspace_bytes": (_z, []),
    "rfnet_chamfer_partial_sums": (_i, [_i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "rfnet_chamfer_step_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_chamfer_step": (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _z, _i, _p]),
    "rfnet_merge_layer_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_merge_layer": (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _p, _z, _p]),
    "rfnet_merge_layer_grad_partials": (_z, [_i, _i]),
    "rfnet_merge_layer_grad": (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "rfnet_approxmatch_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_approxmatch": (_i, [_i, _i, _i, _p, _p, _p, _p, _z, _i, _p]),
    "rfnet_matchcost_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_matchcost": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "rfnet_matchcostgrad_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_matchcostgrad": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _z, _p]),
    "rfnet_emd_cost_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_emd_cost": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _z, _i, _p]),
    "rfnet_emd_cost_grad_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_emd_cost_grad": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _z, _i, _p]),
    "rfnet_farthestpointsampling_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_farthestpointsampling": (_i, [_i, _i, _i, _p, _p, _z, _p, _p]),
    "rfnet_gatherpoint": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "rfnet_scatteraddpoint_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_scatteraddpoint": (_i, [_i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "rfnet_query_ball_point_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_query_ball_point": (_i, [_i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _z, _p]),
    "rfnet_group_point": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "rfnet_group_point_grad_workspace_bytes": (_z, [_i, _i, _i, _i, _i]),
    "rfnet_group_point_grad": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "rfnet_scatter_plan_bytes": (_z, [_i, _i, _i]),
    "rfnet_scatter_plan_build": (_i, [_i, _i, _i, _p, _p, _z, _p]),
    "rfnet_group_point_grad_planned": (_i, [_i, _i, _i, _i, _i, _p, _p, _z, _p, _p]),
    "rfnet_three_interpolate_grad_planned": (_i, [_i, _i, _i, _i, _p, _p, _p, _z, _p, _p]),
    "rfnet_knn_point": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "rfnet_selection_sort": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "rfnet_auction_match": (_i, [_i, _i, _p, _p, _p, _p, _p]),
    "rfnet_three_nn_workspace_bytes": (_z, [_i, _i, _i]),
    "rfnet_three_nn": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "rfnet_three_interpolate": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "rfnet_three_interpolate_grad_workspace_bytes": (_z, [_i, _i, _i, _i]),
    "rfnet_three_interpolate_grad": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "rfnet_nn_distance_host": (_i, [_i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _i]),
    "rfnet_emd_host": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _i]),
    "rfnet_probe_fp32": (_i, [_i, _p, _p, _p]),
}

_lib = None


class RfnetError(RuntimeError):
    pass


def load():
    """Load librfnet_ops.so.  Raises if it has not been built (python -m rfnet_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("rfnet_b200: %s is missing -- build it with `python -m rfnet_b200.build` "
                              "(there is no CPU or PyTorch fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code, what):
    if code != 0:
        msg = load().rfnet_error_string(code)
        raise RfnetError("%s failed: %s (%d)" % (what, msg.decode() if msg else "?", code))
