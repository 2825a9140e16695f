/* Force-included (nvcc -include) when compiling the reference's tf_ops/emd/tf_auctionmatch_g.cu for sm_100a.
 * The kernel uses the pre-Volta warp shuffle __shfl_down(var, delta, width) (tf_auctionmatch_g.cu:220-222,243-245), which
 * CUDA removed for sm_70+.  Both call sites are reached by converged lanes only (the whole warp, then lanes 0..15 of
 * warp 0), so shuffling over the currently active lanes is what the old intrinsic did there.  TEST INFRASTRUCTURE ONLY. */
#pragma once
#define __shfl_down(var, delta, width) __shfl_down_sync(__activemask(), var, delta, width)
