/*
 * rfnet_ops.h -- C ABI of librfnet_ops.so: B200 (sm_100a) kernels for RFNet's point-cloud operators.
 *
 * This is the drop-in boundary.  Every entry point replaces one of the C++ launcher functions that the reference's
 * TensorFlow OpKernels call (file:line given per function, paths relative to the reference repository).  Argument
 * ORDER and MEANING are the reference's; three things are added, always at the end:
 *     - scratch that the reference obtained with allocate_temp is passed as (workspace, workspace_bytes); ask
 *       rfnet_<op>_workspace_bytes() for the size.  Kernels never allocate.
 *     - a CUDA stream (the reference launched everything on the legacy default stream);
 *     - an int return: 0 on success, otherwise a cudaError_t value (1 = cudaErrorInvalidValue for bad arguments).
 *       Nothing is thrown across this boundary and nothing synchronises the host.
 *
 * All pointers are DEVICE pointers unless the function name ends in _host.  Clouds are row-major float32 (b, n, 3),
 * indices int32.  Offsets are 64-bit internally, so b*n*m may exceed 2^31 (the reference's approxmatch cannot:
 * pc_distance/tf_approxmatch.cu:15).  Calls are re-entrant and stream-ordered; there is no global state.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef RFNET_OPS_H_
#define RFNET_OPS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *rfnet_stream_t; /* == cudaStream_t */

/* Library version (major*10000 + minor*100 + patch) and a printable name for a return code. */
int rfnet_version(void);
const char *rfnet_error_string(int code);

/* ---------------------------------------------------------------------------------------------------------------
 * nn_distance (Chamfer).  Replaces NmDistanceKernelLauncher, pc_distance/tf_nndistance.cpp:168 (defined
 * tf_ops/CD/tf_nndistance_g.cu:127-130).  dist1[i,j] = min_k d2(xyz1[i,j], xyz2[i,k]), idx1 = first arg-min; same for
 * direction 2.  d2 is evaluated as fma(dz,dz, fma(dx,dx, dy*dy)) -- the contraction of the reference's GPU binary --
 * unless RFNET_NN_UNFUSED is set in flags, which gives ((dx*dx)+(dy*dy))+(dz*dz) as in the reference's CPU build
 * (pc_distance/tf_nndistance.cpp:21-43).  workspace: rfnet_nn_distance_workspace_bytes() bytes, 16-byte aligned (NULL only when that is 0).
 * flags: RFNET_NN_UNFUSED as above; RFNET_NN_DIRECT evaluates the distance expression for every pair (nn_search_kernel).  The
 * default from 2^24 pairs per call is the FILTERED search (nn_prepare_kernel + nn_filter_kernel, csrc/nn_distance.cu): every pair
 * is still visited, with the expanded form |c-o|^2 - 2(q-o).(c-o) at half the arithmetic, and the distance expression is
 * evaluated only for the group of 16 candidates proven to hold the nearest neighbour (or for the whole candidate range of a
 * query when no group can be certified).  Outputs are the same bits either way; the workspace is required for it.
 * Non-finite coordinates: a query whose distances are all NaN reports (+inf, 0) -- the reference reports (NaN-free seed
 * 1e38-style garbage / NaN depending on the build); three_nn ignores NaN distances exactly as the reference's strict '<'
 * does.  Results are only specified for finite inputs.
 * ------------------------------------------------------------------------------------------------------------- */
#define RFNET_NN_UNFUSED 1
#define RFNET_NN_DIRECT 2
size_t rfnet_nn_distance_workspace_bytes(int b, int n, int m);
int rfnet_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1, float *dist2,
                      int *idx2, void *workspace, size_t workspace_bytes, int flags, rfnet_stream_t stream);
/* Same call; exact_scans (device pointer to one 64-bit counter, may be NULL) is incremented once per (query, work item) that
 * the filtered search could not certify and scanned with the reference expression instead (diagnostics and tests). */
int rfnet_nn_distance_stats(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                            float *dist2, int *idx2, void *workspace, size_t workspace_bytes, int flags,
                            unsigned long long *exact_scans, rfnet_stream_t stream);

/* The launch plan rfnet_nn_distance would use for (b, n, m, flags) on the current device, without launching anything (host
 * arithmetic only; diagnostics and tests): out10 = { direct kernel?, queries per thread,  then per direction (xyz1 -> xyz2,
 * xyz2 -> xyz1): candidates per chunk, chunks per work item, work items per query tile, work items }. */
int rfnet_nn_distance_plan(int b, int n, int m, int flags, int *out10);

/* Replaces NmDistanceGradKernelLauncher, pc_distance/tf_nndistance.cpp:208 (tf_nndistance_g.cu:151-156).
 * grad_xyz1 / grad_xyz2 are fully overwritten (the launcher zero-fills them itself, as the reference does).
 *
 * GRADIENT SCATTERS (this function, rfnet_scatteraddpoint, rfnet_group_point_grad, rfnet_three_interpolate_grad):
 * with a workspace of rfnet_<op>_workspace_bytes() the scatter is ATOMIC-FREE and DETERMINISTIC -- the index list is
 * inverted with integer atomics only and every point's contributions are summed in ascending source order, which is the
 * summation order of the reference's sequential CPU kernels, so the result is bit-exact with them and reproducible run to
 * run.  With workspace == NULL the reference GPU formulation is used (float reductions after a zero-fill), whose last
 * bits depend on thread timing. */
size_t rfnet_nn_distance_grad_workspace_bytes(int b, int n, int m);
int rfnet_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                           const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                           float *grad_xyz2, void *workspace, size_t workspace_bytes, rfnet_stream_t stream);

/* Loss-level epilogue of the reference's chamfer_big / fidelity_loss (vv_recon.py:381-390), which the reference leaves to
 * framework ops: sums4 = { sum sqrt(dist1), b*n, sum sqrt(dist2), b*m } in a fixed summation order.  chamfer_big is then
 * (sums4[0]/sums4[1] + sums4[2]/sums4[3]) / 2, and across GPUs the four numbers are what gets all-reduced. */
size_t rfnet_chamfer_partial_sums_workspace_bytes(void);
int rfnet_chamfer_partial_sums(int b, int n, int m, const float *dist1, const float *dist2, float *sums4,
                               void *workspace, size_t workspace_bytes, rfnet_stream_t stream);

/* One training step of the reference's chamfer_big (vv_recon.py:381-385, forward and backward) in one call and four
 * kernel launches (preparation + search, one epilogue, one reduction; no memset): nn_distance, NnDistanceGrad for the given upstream gradients grad_dist1 (b,n) / grad_dist2 (b,m), and the
 * loss partial sums sums4 (as rfnet_chamfer_partial_sums).  Outputs are exactly those of the three separate calls:
 * dist/idx bit-identical; gradients by the reference's float-reduction formulation (last bits depend on thread timing). */
size_t rfnet_chamfer_step_workspace_bytes(int b, int n, int m);
int rfnet_chamfer_step(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                       const float *grad_dist2, float *dist1, int *idx1, float *dist2, int *idx2, float *grad_xyz1,
                       float *grad_xyz2, float *sums4, void *workspace, size_t workspace_bytes, int flags,
                       rfnet_stream_t stream);

/* merge_layer of the reference's model (vv_recon.py:132-139, always called with knum = 1): every new point is pulled towards
 * its nearest raw point,
 *     out[j] = new[j] + exp(-d2 / (1e-8 + decfactor^2)) * (raw[nn(j)] - new[j]),     d2 = |raw[nn(j)] - new[j]|^2,
 * nn(j) = NnDistance's idx2 for (raw, new).  The reference chains NnDistance (both directions), GroupPoint and five framework
 * ops; this is ONE directed search and one epilogue.  decfactor is a DEVICE pointer to one float (a trained variable there).
 * idx (b, n_new) is returned for the backward: rfnet_merge_layer_grad gives, for an upstream gradient grad_out (b, n_new, 3),
 * grad_new (b, n_new, 3), the rows grad_raw_rows (b, n_new, 3) to be scatter-added into raw by idx (rfnet_scatteraddpoint or a
 * scatter plan), and rfnet_merge_layer_grad_partials(b, n_new) partial sums of d out / d decfactor (add them up). */
size_t rfnet_merge_layer_workspace_bytes(int b, int n_raw, int n_new);
int rfnet_merge_layer(int b, int n_raw, const float *raw, int n_new, const float *newpts, const float *decfactor,
                      float *out, int *idx, void *workspace, size_t workspace_bytes, rfnet_stream_t stream);
size_t rfnet_merge_layer_grad_partials(int b, int n_new);
int rfnet_merge_layer_grad(int b, int n_raw, const float *raw, int n_new, const float *newpts, const int *idx,
                           const float *decfactor, const float *grad_out, float *grad_new, float *grad_raw_rows,
                           float *dec_partial, rfnet_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * approx_match / match_cost (EMD).  Replace approxmatchLauncher, matchcostLauncher, matchcostgradLauncher,
 * pc_distance/tf_approxmatch.cpp:141-143 (defined pc_distance/tf_approxmatch.cu:180-182,226-228,292-295).
 * match is (b, m, n): match[i, l, k] pairs xyz2[i,l] with xyz1[i,k] (tf_approxmatch.cu:152).  `temp` of the reference
 * ((b, 2(n+m)) floats, tf_approxmatch.cpp:168) becomes workspace.
 *
 * Every cloud's result depends on that cloud and on (n, m, flags) only -- never on the batch size or on the position in
 * the batch: evaluating a cloud alone, inside a batch of 32, or on another GPU of a sharded batch gives the same bits.
 *
 * flags (approxmatch, emd_cost, emd_cost_grad):
 *   0 (default)                every sum of the iteration is ONE chain over the candidates in ascending order: the reference
 *                              kernel's rounding order, term for term.  Differs from the reference binary only through the
 *                              flushing exponential (terms < 1.2e-38) and the derived exponentials of the final pass
 *                              (<= 1.1e-6 relative per matrix entry).
 *   RFNET_EMD_EXACT            additionally the reference's non-flushing __expf, no pruning, every exponential of the final
 *                              pass through the MUFU: match is bit-identical to the reference CUDA binary's.
 *   RFNET_EMD_PRUNE            force the exactly-pruned sweeps of the two sharpest levels (curve-ordered rows, per-cluster
 *                              candidate masks; terms that the flushing exponential makes exactly 0 are skipped) for clouds of
 *                              1024 .. 32768 points.  They are used by default wherever they pay (from ~32768 rows per call of
 *                              clouds with >= 2048 points); results are bit-identical either way.
 *   RFNET_EMD_NO_PRUNE         dense sweeps at every level (overrides RFNET_EMD_PRUNE).
 *   RFNET_EMD_SPLIT_SUMS       cut every sum into fixed-length pieces added in ascending order: more parallelism when the
 *                              call holds only one or two small clouds, at the price of a different rounding order (the
 *                              iteration is ill-conditioned: up to ~1e-3 of the largest entry).  Not combinable with EXACT.
 * ------------------------------------------------------------------------------------------------------------- */
#define RFNET_EMD_EXACT 1
#define RFNET_EMD_NO_PRUNE 2
#define RFNET_EMD_SPLIT_SUMS 4
#define RFNET_EMD_PRUNE 8
size_t rfnet_approxmatch_workspace_bytes(int b, int n, int m);
int rfnet_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, void *workspace,
                      size_t workspace_bytes, int flags, rfnet_stream_t stream);
size_t rfnet_matchcost_workspace_bytes(int b, int n, int m);
int rfnet_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *out,
                    void *workspace, size_t workspace_bytes, rfnet_stream_t stream);
size_t rfnet_matchcostgrad_workspace_bytes(int b, int n, int m);
int rfnet_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *grad1,
                        float *grad2, void *workspace, size_t workspace_bytes, rfnet_stream_t stream);
/* Fused approx_match + match_cost for the loss-level caller earth_mover (vv_recon.py:396-399: match = approx_match,
 * cost = match_cost, mean(cost / num_points)): the same sweeps as rfnet_approxmatch, then ONE pass that reduces
 * cost[i] = sum sqrt(d2) * match (b floats) while the matrix entries are still in registers.  `match` may be NULL:
 * the (b, m, n) matrix -- 1 GiB per cloud at 16384 x 16384 -- is then never written. */
size_t rfnet_emd_cost_workspace_bytes(int b, int n, int m);
int rfnet_emd_cost(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *cost,
                   void *workspace, size_t workspace_bytes, int flags, rfnet_stream_t stream);
/* The training form of the same chain: approx_match -> match_cost AND matchcostgrad (tf_approxmatch.py:44-50), with no
 * (b, m, n) matrix anywhere.  After the sweeps, two passes rebuild each matrix entry in registers from the per-level
 * factors: one reduces cost and grad1 (b, n, 3), the other grad2 (b, m, 3); grad1/grad2 are what rfnet_matchcostgrad
 * returns for the matrix rfnet_approxmatch would have written (before the grad_cost scaling of the Python wrapper). */
size_t rfnet_emd_cost_grad_workspace_bytes(int b, int n, int m);
int rfnet_emd_cost_grad(int b, int n, int m, const float *xyz1, const float *xyz2, float *cost, float *grad1,
                        float *grad2, void *workspace, size_t workspace_bytes, int flags, rfnet_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * sampling.  Replace farthestpointsamplingLauncher, gatherpointLauncher, scatteraddpointLauncher,
 * tf_ops/sampling/tf_sampling.cpp:94,125,150 (defined tf_ops/sampling/tf_sampling_g.cu:203-211).
 * FPS: out (b, m) int32, first index 0, ties broken exactly as the reference's 512-thread block does.  The workspace
 * (the reference's (32, n) temp, tf_sampling.cpp:115) holds the Morton order of the clouds for the pruned kernel
 * (n <= 16384, m >= 256) or the running distances of clouds too large for registers; with workspace == NULL clouds of up
 * to 32768 points run on the cluster kernel, which needs none.  Every path returns the same indices.
 * scatteraddpoint zero-fills inp_g itself (the reference's OpKernel did it, tf_sampling.cpp:174).
 * ------------------------------------------------------------------------------------------------------------- */
size_t rfnet_farthestpointsampling_workspace_bytes(int b, int n, int m);
int rfnet_farthestpointsampling(int b, int n, int m, const float *inp, void *workspace, size_t workspace_bytes,
                                int *out, rfnet_stream_t stream);
int rfnet_gatherpoint(int b, int n, int m, const float *inp, const int *idx, float *out, rfnet_stream_t stream);
size_t rfnet_scatteraddpoint_workspace_bytes(int b, int n, int m);
int rfnet_scatteraddpoint(int b, int n, int m, const float *out_g, const int *idx, float *inp_g, void *workspace,
                          size_t workspace_bytes, rfnet_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * grouping.  Replace queryBallPointLauncher, groupPointLauncher, groupPointGradLauncher,
 * tf_ops/grouping/tf_grouping.cpp:67,146,177 (defined tf_ops/grouping/tf_grouping_g.cu:125-141).
 * radius is a DEVICE pointer to one float, as in the reference (tf_grouping.cpp:93-95).  Rows with no point in the
 * ball are filled with 0 (the reference leaves them uninitialised) and get pts_cnt 0.
 * With a workspace (and 2048 <= n <= 32768) query_ball_point bins the dataset in a uniform grid and only tests the cells
 * a ball touches; with workspace == NULL every query scans the whole dataset.  Identical results either way.
 * groupPointGrad zero-fills grad_points itself (reference: tf_grouping.cpp:208).
 * ------------------------------------------------------------------------------------------------------------- */
size_t rfnet_query_ball_point_workspace_bytes(int b, int n, int m);
int rfnet_query_ball_point(int b, int n, int m, const float *radius, int nsample, const float *xyz1,
                           const float *xyz2, int *idx, int *pts_cnt, void *workspace, size_t workspace_bytes,
                           rfnet_stream_t stream);
int rfnet_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                      rfnet_stream_t stream);
size_t rfnet_group_point_grad_workspace_bytes(int b, int n, int c, int m, int nsample);
int rfnet_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                           float *grad_points, void *workspace, size_t workspace_bytes, rfnet_stream_t stream);

/* Scatter plans.  Every gradient of a gather (gather_point, group_point, three_interpolate) adds rows of an upstream gradient
 * into the points they were gathered from.  The atomic-free path inverts the index tensor first; that inverse -- a "plan" --
 * depends on idx only, so it can be built ONCE, off the critical path (e.g. at forward time), and shared by every gradient that
 * goes through the same idx (the xyz and the feature grouping of one layer).  idx is read as (b, rows) targets in
 * [0, n_targets): group_point: rows = m * nsample, n_targets = n; gather_point: rows = m; three_interpolate: rows = 3 n,
 * n_targets = m.  The *_planned gradients equal rfnet_<op>_grad with a workspace bit for bit. */
size_t rfnet_scatter_plan_bytes(int b, int n_targets, int rows);
int rfnet_scatter_plan_build(int b, int n_targets, int rows, const int *idx, void *plan, size_t plan_bytes,
                             rfnet_stream_t stream);
int rfnet_group_point_grad_planned(int b, int n, int c, int m, int nsample, const float *grad_out, const void *plan,
                                   size_t plan_bytes, float *grad_points, rfnet_stream_t stream);

/* knn_point for 3-d points, k <= 32.  The reference has no kernel for it: tf_ops/grouping/tf_grouping.py:48-73 materialises
 * the (b, m, n) distance matrix with framework ops and calls tf.nn.top_k(-dist) on the CPU.  xyz1 = dataset (b,n,3),
 * xyz2 = queries (b,m,3); val (b,m,k) = NEGATED squared distances in descending order (i.e. nearest first), idx (b,m,k). */
int rfnet_knn_point(int b, int n, int m, int k, const float *xyz1, const float *xyz2, float *val, int *idx,
                    rfnet_stream_t stream);

/* select_top_k.  Replaces selectionSortLauncher, tf_ops/grouping/tf_grouping.cpp:112 (defined tf_grouping_g.cu:129-132).
 * dist (b,m,n) -> outi (b,m,n), out (b,m,n): per row a copy of dist / 0..n-1 after k steps of selection sort with swaps,
 * so the first k entries are the k smallest (ascending, first index among equals) and the tail is permuted exactly as the
 * reference leaves it.  k > n is clamped to n (the reference reads out of bounds).  NaN entries behave as under the reference's
 * strict '<': never selected, and a NaN at position s < k stays there; -0.0 and +0.0 compare equal.  out / outi must not alias dist. */
int rfnet_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out, rfnet_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * auction_match.  Replaces AuctionMatchLauncher, tf_ops/emd/tf_auctionmatch.cpp:26 (defined tf_auctionmatch_g.cu:292-294).
 * xyz1, xyz2 (b,n,3) -> matchl (b,n): index into xyz2 assigned to every xyz1 point; matchr (b,n): the inverse.
 * Same bid sequence as the reference kernel (one bidder at a time, FIFO, tolerance 1e-4 -> 1e-2 -> 1 every 40 n bids).
 * The reference's `cost` temp (b*n*n floats, tf_auctionmatch.cpp:54) is not needed: no workspace at all.
 * n <= RFNET_AUCTION_MAX_POINTS (the reference stops at 4096, tf_auctionmatch.cpp:37).  Points left unassigned when
 * the auction gives up have matchl = -1 and the corresponding matchr = -1.
 * ------------------------------------------------------------------------------------------------------------- */
#define RFNET_AUCTION_MAX_POINTS 8192
int rfnet_auction_match(int b, int n, const float *xyz1, const float *xyz2, int *matchl, int *matchr,
                        rfnet_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * interpolation.  The reference has CPU code only: threenn_cpu, threeinterpolate_cpu, threeinterpolate_grad_cpu,
 * tf_ops/interpolation/tf_interpolate.cpp:60,107,131.  Same argument orders.  three_nn evaluates d2 UNFUSED, as the
 * reference's CPU build does, so indices match it bit for bit.  three_interpolate_grad zero-fills grad_points
 * (reference: tf_interpolate.cpp:258).
 * ------------------------------------------------------------------------------------------------------------- */
/* three_nn takes an optional workspace: with it (and 512 <= m <= 4096 known points) the search runs over a uniform grid of
 * the known cloud instead of scanning it; identical distances and indices. */
size_t rfnet_three_nn_workspace_bytes(int b, int n, int m);
int rfnet_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, void *workspace,
                   size_t workspace_bytes, rfnet_stream_t stream);
int rfnet_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                            float *out, rfnet_stream_t stream);
size_t rfnet_three_interpolate_grad_workspace_bytes(int b, int n, int c, int m);
int rfnet_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                 const float *weight, float *grad_points, void *workspace, size_t workspace_bytes,
                                 rfnet_stream_t stream);

int rfnet_three_interpolate_grad_planned(int b, int n, int c, int m, const float *grad_out, const float *weight,
                                         const void *plan, size_t plan_bytes, float *grad_points, rfnet_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Host-buffer entry points: what a CPU-side caller (e.g. the reference's DEVICE_CPU OpKernels,
 * pc_distance/tf_nndistance.cpp:60-81, pc_distance/tf_approxmatch.cpp:175-199) binds.  Pointers are HOST memory
 * (pinned memory makes the copies asynchronous); the call copies in, runs on `device`, copies out and returns when the
 * outputs are complete.  Device scratch comes from the CUDA stream-ordered pool of that device.
 * ------------------------------------------------------------------------------------------------------------- */
int rfnet_nn_distance_host(int device, int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1,
                           int *idx1, float *dist2, int *idx2, int flags);
int rfnet_emd_host(int device, int b, int n, int m, const float *xyz1, const float *xyz2, float *match_or_null,
                   float *cost, int flags);

/* ---------------------------------------------------------------------------------------------------------------
 * Measurement helper (used by bench.py only): runs a dependent-free FP32 FFMA2 stream on every SM for `iters` rounds
 * and returns lane-operations executed in *lane_ops, so the bench can report a MEASURED FP32 pipe peak next to the
 * architectural one.  Not part of the reference's interface.
 * ------------------------------------------------------------------------------------------------------------- */
int rfnet_probe_fp32(int iters, float *sink, unsigned long long *lane_ops, rfnet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RFNET_OPS_H_ */
