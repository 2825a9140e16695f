// query_ball_point, group_point and its gradient for sm_100a.
//
// Replaces query_ball_point_gpu / group_point_gpu / group_point_grad_gpu (tf_ops/grouping/tf_grouping_g.cu:3-78).
//
// Ball query design: the reference gives each query to one thread which streams the whole dataset from global memory
// (uncoalesced, one block per cloud).  Here a WARP owns a query: the dataset is staged tile by tile in shared memory, each
// lane tests four consecutive points of a 128-point group (two packed FP32x2 pairs), and ballots + popcounts append the hits IN INDEX ORDER, so "the first
// nsample points inside the ball" (tf_grouping_g.cu:17-31) is reproduced exactly, with early exit once the row is full.
//
// The predicate max(sqrtf(d2), 1e-20f) < r is evaluated without a square root: sqrt_rn is monotone, so it equals
// d2 < T with T = min{t : sqrt_rn(t) >= r}; T is found once per thread by stepping nextafter around r*r.

#include "common.cuh"
#include "pointgrid.cuh"
#include "rfnet_ops.h"
#include "segscatter.cuh"

namespace rfnet {

constexpr int BQ_WARPS = 8;
constexpr int BQ_QPW = 4;      // queries per warp
constexpr int BQ_TILE = 2048;  // dataset points per shared-memory tile (24 KiB, multiple of 128)

__device__ __forceinline__ float ball_threshold(float r) {
    if (!(r > 1e-20f)) return 0.0f;                          // max(.,1e-20f) < r can never hold (also r = NaN)
    if (r == __int_as_float(0x7f800000)) return r;           // every finite distance is inside
    float t = __fmul_rn(r, r);
    while (t > 0.0f && __fsqrt_rn(t) >= r) t = __uint_as_float(__float_as_uint(t) - 1u);  // down to sqrt(t) < r
    while (__fsqrt_rn(t) < r) t = __uint_as_float(__float_as_uint(t) + 1u);               // up to the first t with sqrt(t) >= r
    return t;
}

__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_kernel(int n, int m, const float* __restrict__ radius, int nsample,
                                                                   const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                                   int* __restrict__ idx, int* __restrict__ pts_cnt) {
    // dataset tile, structure-of-arrays and NEGATED: query + (-point) == query - point exactly (the operand order of
    // tf_grouping_g.cu:24), and each lane reads two consecutive points per LDS.64 for the packed distance
    __shared__ __align__(16) float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
    const int cloud = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* __restrict__ data = xyz1 + (size_t)cloud * n * 3;
    const float T = ball_threshold(radius[0]);
    const unsigned lt = (1u << lane) - 1u;

    const int qbase = (blockIdx.x * BQ_WARPS + warp) * BQ_QPW;
    float qx[BQ_QPW], qy[BQ_QPW], qz[BQ_QPW];
    int cnt[BQ_QPW], first[BQ_QPW];
#pragma unroll
    for (int u = 0; u < BQ_QPW; ++u) {
        const int j = qbase + u;
        const bool v = j < m;
        const float* q = xyz2 + ((size_t)cloud * m + (v ? j : 0)) * 3;
        qx[u] = q[0]; qy[u] = q[1]; qz[u] = q[2];
        cnt[u] = v ? 0 : nsample;  // out-of-range queries count as already full
        first[u] = 0;
    }

    for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
        const int len = min(BQ_TILE, n - t0);
        const int len64 = (len + 127) & ~127;   // padded to whole 128-point steps
        __syncthreads();
        for (int i = threadIdx.x; i < len64; i += BQ_WARPS * 32) {
            const bool v = i < len;
            const float* p = data + (size_t)(t0 + (v ? i : 0)) * 3;
            sx[i] = v ? -p[0] : __int_as_float(0x7f800000);  // padding sits at infinity: d2 = inf is never inside a ball
            sy[i] = v ? -p[1] : 0.f;
            sz[i] = v ? -p[2] : 0.f;
        }
        __syncthreads();
        bool all_full = true;
#pragma unroll
        for (int u = 0; u < BQ_QPW; ++u) {
            if (cnt[u] >= nsample) continue;  // warp-uniform
            int* __restrict__ row = idx + ((size_t)cloud * m + qbase + u) * nsample;
            const float2 QX = make_float2(qx[u], qx[u]), QY = make_float2(qy[u], qy[u]), QZ = make_float2(qz[u], qz[u]);
            for (int k0 = 0; k0 < len && cnt[u] < nsample; k0 += 128) {
                const int k = k0 + 4 * lane;  // this lane tests dataset points k .. k+3 (two packed pairs from one LDS.128 per axis)
                const float4 X = *reinterpret_cast<const float4*>(&sx[k]), Y = *reinterpret_cast<const float4*>(&sy[k]), Z = *reinterpret_cast<const float4*>(&sz[k]);
                const float2 da = sqdist3x2<true>(__fadd2_rn(make_float2(X.x, X.y), QX), __fadd2_rn(make_float2(Y.x, Y.y), QY), __fadd2_rn(make_float2(Z.x, Z.y), QZ));
                const float2 db = sqdist3x2<true>(__fadd2_rn(make_float2(X.z, X.w), QX), __fadd2_rn(make_float2(Y.z, Y.w), QY), __fadd2_rn(make_float2(Z.z, Z.w), QZ));
                const bool h0 = da.x < T, h1 = da.y < T, h2 = db.x < T, h3 = db.y < T;
                const unsigned any = __ballot_sync(0xffffffffu, h0 | h1 | h2 | h3);   // most steps have no hit: one vote
                if (any) {   // warp-uniform
                    const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                    const unsigned m2 = __ballot_sync(0xffffffffu, h2), m3 = __ballot_sync(0xffffffffu, h3);
                    if (cnt[u] == 0) {
                        // first hit in index order: lowest lane with a hit, then its lowest point
                        const int fl = __ffs(any) - 1;
                        const int fo = ((m0 >> fl) & 1u) ? 0 : (((m1 >> fl) & 1u) ? 1 : (((m2 >> fl) & 1u) ? 2 : 3));
                        first[u] = t0 + k0 + 4 * fl + fo;
                    }
                    // position in index order: every hit of a lower lane (all four of its points) comes first, then this lane's own in order
                    int pos = cnt[u] + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt);
                    if (h0) { if (pos < nsample) row[pos] = t0 + k; ++pos; }
                    if (h1) { if (pos < nsample) row[pos] = t0 + k + 1; ++pos; }
                    if (h2) { if (pos < nsample) row[pos] = t0 + k + 2; ++pos; }
                    if (h3 && pos < nsample) row[pos] = t0 + k + 3;
                    cnt[u] += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
                }
            }
            all_full = all_full && cnt[u] >= nsample;
        }
        if (__syncthreads_and(all_full)) break;
    }

#pragma unroll
    for (int u = 0; u < BQ_QPW; ++u) {
        const int j = qbase + u;
        if (j >= m) continue;
        const int c = min(cnt[u], nsample);
        int* __restrict__ row = idx + ((size_t)cloud * m + j) * nsample;
        // remaining slots repeat the first hit (tf_grouping_g.cu:26-29); rows without any hit are defined as 0 here
        for (int s = c + lane; s < nsample; s += 32) row[s] = first[u];
        if (lane == 0) pts_cnt[(size_t)cloud * m + j] = c;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Grid variant (used when the caller passes a workspace and the dataset has >= BG_MIN_POINTS points).
// The hits of a query all lie in the cells of a uniform grid that its ball touches; with cells at least one radius wide
// that is at most 3 x 3 x 3 of them.  ball_grid_kernel (one CTA per cloud) bins the dataset with a counting sort --
// cell = ((cx * Gy) + cy) * Gz + cz, G = min(32, extent / r) per axis -- and stores
// the cell offsets and the coordinates in cell order.  ball_query_grid_kernel gives a warp per query: it walks the <= 9 runs of
// z-consecutive cells (contiguous in the sorted order), tests the SAME predicate on the same fused distance as the scan
// kernel above, marks every hit in a per-warp bitmap over the dataset indices (shared memory), and reads the bitmap back in
// index order: "the first nsample points inside the ball in index order" (tf_grouping_g.cu:17-31) computed from ~n/37
// candidates instead of n, whatever the number of hits.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BG_MIN_POINTS = 2048;
constexpr int BG_MAX_POINTS = 32768;   // one bit per dataset point and warp in shared memory (4 KiB x 8 warps)
__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_grid_kernel(int n, int m, const float* __restrict__ radius, int nsample,
                                                                        const float* __restrict__ xyz2, const unsigned char* __restrict__ ws, size_t stride,
                                                                        int* __restrict__ idx, int* __restrict__ pts_cnt) {
    // one bit per dataset point and warp: hits are marked in whatever order the cells deliver them and read back in index order
    __shared__ unsigned bitmap[BQ_WARPS][BG_MAX_POINTS / 32];
    const int cloud = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BallGridView v = ball_grid_view(ws, stride, cloud, n);
    const BallGrid g = *v.g;
    const float r = radius[0];
    const float T = ball_threshold(r);
    const float rr = r * 1.0001f;   // the cell range is taken for a slightly larger ball: rounding can only add cells
    unsigned* __restrict__ bm = bitmap[warp];
    const int words = (n + 31) >> 5, rounds = (words + 31) >> 5;
    for (int w = lane; w < words; w += 32) bm[w] = 0u;
    __syncwarp();
    for (int u = 0; u < BQ_QPW; ++u) {
        const int j = (blockIdx.x * BQ_WARPS + warp) * BQ_QPW + u;
        if (j >= m) break;   // warp-uniform
        const float* q = xyz2 + ((size_t)cloud * m + j) * 3;
        const float qc[3] = {q[0], q[1], q[2]};
        int c0[3], c1[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float pad = rr + 1e-6f * (fabsf(qc[a]) + fabsf(g.lo[a]));   // + a few ulps of the coordinates, for radii near their resolution
            c0[a] = min(g.G[a] - 1, max(0, (int)floorf((qc[a] - pad - g.lo[a]) * g.scale[a])));
            c1[a] = min(g.G[a] - 1, max(0, (int)floorf((qc[a] + pad - g.lo[a]) * g.scale[a])));
        }
        // test the candidates of one run of z-consecutive cells, two warp-steps at a time: the loads are what this loop waits for
        auto scan_run = [&](int rs, int e) {
            for (int p0 = rs; p0 < e; p0 += 64) {
                const int pa = p0 + lane, pb = pa + 32;
                const bool va = pa < e, vb = pb < e;
                const float* ca = v.sorted_xyz + (size_t)(va ? pa : 0) * 3;
                const float* cb = v.sorted_xyz + (size_t)(vb ? pb : 0) * 3;
                const float ax = __ldg(ca), ay = __ldg(ca + 1), az = __ldg(ca + 2), bx = __ldg(cb), by = __ldg(cb + 1), bz = __ldg(cb + 2);
                // query + (-point): the operand order of the scan kernel (and of tf_grouping_g.cu:24)
                const float da = sqdist3<true>(__fadd_rn(-ax, qc[0]), __fadd_rn(-ay, qc[1]), __fadd_rn(-az, qc[2]));
                const float db = sqdist3<true>(__fadd_rn(-bx, qc[0]), __fadd_rn(-by, qc[1]), __fadd_rn(-bz, qc[2]));
                if (va && da < T) {
                    const int k = __ldg(v.sorted_idx + pa);
                    atomicOr(&bm[k >> 5], 1u << (k & 31));
                }
                if (vb && db < T) {
                    const int k = __ldg(v.sorted_idx + pb);
                    atomicOr(&bm[k >> 5], 1u << (k & 31));
                }
            }
        };
        if (c1[0] - c0[0] <= 2 && c1[1] - c0[1] <= 2) {
            // the usual case (cells are at least one radius wide): at most 3 x 3 runs; fetch all their bounds before any candidate
            int rs[9], re[9];
#pragma unroll
            for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 3; ++b2) {
                    const int cx = c0[0] + a2, cy = c0[1] + b2;
                    const bool ok = cx <= c1[0] && cy <= c1[1];
                    const int ca = ok ? (cx * g.G[1] + cy) * g.G[2] : 0;
                    rs[a2 * 3 + b2] = ok ? __ldg(v.cell_start + ca + c0[2]) : 0;
                    re[a2 * 3 + b2] = ok ? __ldg(v.cell_start + ca + c1[2] + 1) : 0;
                }
#pragma unroll
            for (int rI = 0; rI < 9; ++rI) scan_run(rs[rI], re[rI]);
        } else {   // a ball wider than three cells (radius close to the resolution of the coordinates): plain loops
            for (int cx = c0[0]; cx <= c1[0]; ++cx)
                for (int cy = c0[1]; cy <= c1[1]; ++cy) {
                    const int ca = (cx * g.G[1] + cy) * g.G[2];
                    scan_run(v.cell_start[ca + c0[2]], v.cell_start[ca + c1[2] + 1]);
                }
        }
        __syncwarp();
        // read the bitmap back in index order (word w belongs to lane w % 32 in round w / 32), clearing it for the next query
        int* __restrict__ row = idx + ((size_t)cloud * m + j) * nsample;
        int total = 0;
        for (int t = 0; t < rounds; ++t) {
            const int wi = t * 32 + lane;
            unsigned w = 0u;
            if (wi < words) { w = bm[wi]; bm[wi] = 0u; }
            if (!__any_sync(0xffffffffu, w != 0u)) continue;
            const int c = __popc(w);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            int off = total + incl - c;
            while (w && off < nsample) {
                const int bit = __ffs(w) - 1;
                w &= w - 1;
                row[off++] = wi * 32 + bit;
            }
            total += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        const int c = min(total, nsample);
        const int first = total ? row[0] : 0;   // rows without any hit are defined as 0 (the reference leaves them uninitialised)
        for (int s2 = c + lane; s2 < nsample; s2 += 32) row[s2] = first;
        if (lane == 0) pts_cnt[(size_t)cloud * m + j] = c;
    }
}

// out[i,j,k,:] = points[i, idx[i,j,k], :]                                         (tf_grouping_g.cu:40-57)
// grid.y = cloud, so all index arithmetic inside a cloud is 32-bit (64-bit divisions dominated the first version)
// GP_E output vectors per thread, GP_E * 256 consecutive vectors per CTA: all index loads of a thread are issued first,
// then all row loads, then the stores -- GP_E independent 16-byte gathers in flight per thread instead of one dependent
// idx -> row chain (0.62 -> see profiles/r2_gathers.txt of the measured copy bandwidth at c = 64).  Stores are streaming
// (st.global.cs): the output is written once and never re-read here, so it should not evict `points` from L2.
constexpr int GP_E = 4;
template <typename VEC>
__device__ __forceinline__ void store_streaming(VEC* p, VEC v);
template <>
__device__ __forceinline__ void store_streaming<float4>(float4* p, float4 v) { __stcs(p, v); }
template <>
__device__ __forceinline__ void store_streaming<float>(float* p, float v) { __stcs(p, v); }
template <typename VEC>
__global__ void __launch_bounds__(256) group_point_kernel(int n, int cv, unsigned rows_per_cloud, const VEC* __restrict__ points, const int* __restrict__ idx,
                                                         VEC* __restrict__ out) {
    const size_t cloud = blockIdx.y;
    const unsigned total = rows_per_cloud * (unsigned)cv;       // output vectors of this cloud (< 2^31, checked by the caller)
    const unsigned e0 = blockIdx.x * (256u * GP_E) + threadIdx.x;
    const int* __restrict__ id = idx + cloud * rows_per_cloud;
    const VEC* __restrict__ P = points + cloud * (size_t)n * cv;
    VEC* __restrict__ O = out + cloud * (size_t)total;
    unsigned l[GP_E];
    int src[GP_E];
#pragma unroll
    for (int u = 0; u < GP_E; ++u) {
        const unsigned e = e0 + u * 256u;
        const unsigned row = e / (unsigned)cv;
        l[u] = e - row * (unsigned)cv;
        src[u] = e < total ? __ldg(id + row) : -1;
    }
    VEC v[GP_E];
#pragma unroll
    for (int u = 0; u < GP_E; ++u)
        if (src[u] >= 0) v[u] = __ldg(P + (size_t)src[u] * cv + l[u]);
#pragma unroll
    for (int u = 0; u < GP_E; ++u)
        if (src[u] >= 0) store_streaming<VEC>(O + e0 + u * 256u, v[u]);
}

// grad_points[i, idx[i,j,k], :] += grad_out[i,j,k,:]  after zero-fill                (tf_grouping_g.cu:61-78, tf_grouping.cpp:208)
__global__ void group_point_grad_kernel(int n, int c, unsigned rows_per_cloud, const float* __restrict__ grad_out, const int* __restrict__ idx,
                                        float* __restrict__ grad_points) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned row = t / (unsigned)c;
    if (row >= rows_per_cloud) return;
    const unsigned l = t - row * (unsigned)c;
    const size_t cloud = blockIdx.y;
    const size_t grow = cloud * rows_per_cloud + row;
    atomicAdd(grad_points + (cloud * n + (size_t)idx[grow]) * c + l, grad_out[grow * c + l]);
}
// c % 4 == 0 and 16-byte aligned rows: one 128-bit reduction (REDG.ADD.F32x4) per four channels
__global__ void group_point_grad_v4_kernel(int n, int cv, unsigned rows_per_cloud, const float4* __restrict__ grad_out, const int* __restrict__ idx,
                                           float4* __restrict__ grad_points) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned row = t / (unsigned)cv;
    if (row >= rows_per_cloud) return;
    const unsigned l = t - row * (unsigned)cv;
    const size_t cloud = blockIdx.y;
    const size_t grow = cloud * rows_per_cloud + row;
    const float4 g = grad_out[grow * cv + l];
    float4* dst = grad_points + (cloud * n + (size_t)idx[grow]) * cv + l;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w) : "memory");
}

// Atomic-free gradient: grad_points[i, t, :] = sum of grad_out rows whose idx is t, in ascending row order (the order of
// the reference's CPU prototype, tf_ops/grouping/query_ball_point.cpp:69-84) -> deterministic, bit-exact with it.
template <typename VEC>
__device__ __forceinline__ VEC vec_add(VEC a, VEC b);
template <>
__device__ __forceinline__ float vec_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ float4 vec_add<float4>(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
template <typename VEC>
__device__ __forceinline__ VEC vec_zero();
template <>
__device__ __forceinline__ float vec_zero<float>() { return 0.f; }
template <>
__device__ __forceinline__ float4 vec_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <typename VEC>
__global__ void group_point_grad_seg_kernel(int n, int cv, unsigned R, const VEC* __restrict__ grad_out, const int* __restrict__ offset,
                                            const int* __restrict__ list, VEC* __restrict__ grad_points) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (target point, channel vector) of this cloud
    const unsigned i = t / (unsigned)cv;
    if (i >= (unsigned)n) return;
    const unsigned l = t - i * (unsigned)cv;
    const size_t cloud = blockIdx.y;
    const int beg = offset[cloud * (n + 1) + i], end = offset[cloud * (n + 1) + i + 1];
    const int* __restrict__ seg = list + cloud * R;
    const VEC* __restrict__ G = grad_out + cloud * (size_t)R * cv + l;
    VEC acc = vec_zero<VEC>();
    int e = beg;
    for (; e + 4 <= end; e += 4) {   // four source rows in flight; added in ascending order, as the sequential reference does
        const int s0 = seg[e], s1 = seg[e + 1], s2 = seg[e + 2], s3 = seg[e + 3];
        const VEC g0 = __ldg(G + (size_t)s0 * cv), g1 = __ldg(G + (size_t)s1 * cv), g2 = __ldg(G + (size_t)s2 * cv), g3 = __ldg(G + (size_t)s3 * cv);
        acc = vec_add<VEC>(vec_add<VEC>(vec_add<VEC>(vec_add<VEC>(acc, g0), g1), g2), g3);
    }
    for (; e < end; ++e) acc = vec_add<VEC>(acc, __ldg(G + (size_t)seg[e] * cv));
    grad_points[(cloud * n + i) * cv + l] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// selection_sort (select_top_k, tf_grouping_g.cu:83-123): per row of dist (b*m rows of n), copy the row and run k steps
// of selection sort WITH SWAPS, values and indices alike -- positions [0,k) end up holding the k smallest in ascending
// order, the tail holds the displaced entries exactly where the reference's swaps leave them.
// The reference gives a row to one thread (n*k serial global reads).  Here a WARP owns a row held in shared memory:
// each step is a lane-strided arg-min (first minimum: lowest position among equals, as the strict '<' scan from min = s
// gives) finished with a REDUX-free shuffle tree on (value, position), then lane 0 swaps.  Rows longer than the shared
// budget run the same code on the output row in global memory.
// ---------------------------------------------------------------------------------------------------------------
// k steps of selection sort with swaps on one row (values + positions, shared or global memory), by one warp: lane-strided arg-min
// (first minimum: lowest position among equals, as the strict '<' scan from min = s gives), shuffle tree on (value, position), lane 0
// swaps.  NaN as in the reference: nothing compares below a NaN, so a NaN at position s stays there (the step does nothing) and a NaN
// elsewhere is never selected.
__device__ __forceinline__ void ss_plain_sort(float* val, int* pos, int n, int k, int lane) {
    for (int s = 0; s < k; ++s) {
        const float vs = val[s];
        if (vs != vs) continue;   // warp-uniform
        float bv = 0.f;
        int bp = 0x7fffffff;   // "no element"
        for (int t = s + lane; t < n; t += 32) {
            const float v = val[t];
            if (v == v && (bp == 0x7fffffff || v < bv)) { bv = v; bp = t; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int op = __shfl_xor_sync(0xffffffffu, bp, off);
            // keep the smaller value; among equals the lower position; an empty lane never wins
            const bool take = op != 0x7fffffff && (bp == 0x7fffffff || ov < bv || (ov == bv && op < bp));
            if (take) { bv = ov; bp = op; }
        }
        if (lane == 0 && bp != s) {
            const float tv = val[bp]; val[bp] = val[s]; val[s] = tv;
            const int tp = pos[bp]; pos[bp] = pos[s]; pos[s] = tp;
        }
        __syncwarp();
    }
}

constexpr int SS_WARPS = 4;
__global__ void __launch_bounds__(SS_WARPS * 32) selection_sort_kernel(int n, int k, size_t rows, int smem_rows, const float* __restrict__ dist,
                                                                       int* __restrict__ outi, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row = (size_t)blockIdx.x * SS_WARPS + warp;
    if (row >= rows) return;
    const float* __restrict__ src = dist + row * n;
    float* o = out + row * n;
    int* oi = outi + row * n;
    float* val = smem_rows ? reinterpret_cast<float*>(ss_smem) + (size_t)warp * n : o;
    int* pos = smem_rows ? reinterpret_cast<int*>(ss_smem) + (size_t)SS_WARPS * n + (size_t)warp * n : oi;
    for (int i = lane; i < n; i += 32) {
        val[i] = src[i];
        pos[i] = i;
    }
    __syncwarp();
    ss_plain_sort(val, pos, n, k, lane);
    if (smem_rows) {
        for (int i = lane; i < n; i += 32) {
            o[i] = val[i];
            oi[i] = pos[i];
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// selection_sort for k <= 128: the same result without running the selection sort over the row.
//
// k steps of selection sort touch few positions: step s swaps position s with the position of the current minimum of [s, n).
// Let T be any value with at least k row entries <= T.  Every selected minimum is one of the k smallest entries, hence <= T, and
// an entry only ever moves to position s < k (selected) or to the position a selected entry just left (displaced).  So the set
//     P = [0, k)  U  { t : row[t] <= T }
// is closed under the swaps, every minimum search over [s, n) finds its answer inside P (ties included: ALL entries equal to
// the minimum are <= T, and the reference's strict '<' scan takes the one at the lowest current position), and everything
// outside P keeps its value and its index.  A warp therefore
//   (1) streams the row once: out = dist, outi = 0..n-1 (vector loads and stores), each lane keeping its J = ceil(k/32) + 1 smallest
//       values; T = the k-th smallest of those 32 J values (they are row entries, so at least k row entries are <= T), found by
//       bisection on their integer keys with ballots -- for a random row |P| - k is a handful beyond k (a first version took the
//       largest of the lanes' minima: ~6 % of the row);
//   (2) re-reads the row (cache hits) and compacts P in position order into shared memory (ballot + popcount);
//   (3) runs the k steps on that list: slot s is position s; arg-min over slots >= s on an order-preserving integer key (two REDUX:
//       smallest key, then the lowest slot holding it).  Lists of up to 128 slots keep their keys in registers (four slots per lane,
//       the displaced key travels by one shuffle, lane 0 keeps the slot -> entry map in shared memory); longer ones scan shared memory;
//   (4) writes back the slots whose entry changed.
// A row whose P exceeds SSF_CAP slots (masses of equal values) runs the plain selection sort on its output row instead.
// NaN: the reference never moves a NaN out of position s (nothing compares below it) and never selects one; reproduced by
// skipping step s when slot s holds a NaN and giving NaN the largest key.  -0.0 and +0.0 compare equal, as in the reference.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SSF_WARPS = 8;
constexpr int SSF_CAP = 256;     // slots per row list
constexpr int SSF_REG = 4;       // slots per lane of the register-resident step loop
constexpr int SSF_JMAX = 4;      // k <= 32 * SSF_JMAX

__device__ __forceinline__ unsigned ss_key(float v) {
    // order-preserving map float -> unsigned with -0 == +0 and every NaN on top
    if (v != v) return 0xffffffffu;
    const unsigned u = (unsigned)__float_as_int(v + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ss_unkey(unsigned key) {   // inverse of ss_key on non-NaN keys
    return __int_as_float((int)((key & 0x80000000u) ? (key ^ 0x80000000u) : ~key));
}

template <int J>
__device__ __forceinline__ void ss_insert(float (&a)[J], float v) {
    // a[0] <= ... <= a[J-1] keeps the J smallest values seen; NaN counts as +inf
    v = (v != v) ? __int_as_float(0x7f800000) : v;
#pragma unroll
    for (int i = J - 1; i > 0; --i) a[i] = fminf(a[i], fmaxf(a[i - 1], v));
    a[0] = fminf(a[0], v);
}

template <int J>   // values kept per lane: ceil(k / 32) + 1
__global__ void __launch_bounds__(SSF_WARPS * 32) selection_sort_fast_kernel(int n, int k, size_t rows, const float* __restrict__ dist,
                                                                             int* __restrict__ outi, float* __restrict__ out) {
    __shared__ int s_pos[SSF_WARPS][SSF_CAP];        // position (= original index) of list entry e
    __shared__ float s_val[SSF_WARPS][SSF_CAP];      // its value
    __shared__ unsigned s_key[SSF_WARPS][SSF_CAP];   // key of the entry currently in slot i (lists longer than 32 * SSF_REG only)
    __shared__ int s_id[SSF_WARPS][SSF_CAP];         // which entry slot i currently holds
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row = (size_t)blockIdx.x * SSF_WARPS + warp;
    if (row >= rows) return;
    const float* __restrict__ src = dist + row * n;
    float* o = out + row * n;
    int* oi = outi + row * n;
    const float inf = __int_as_float(0x7f800000);

    // (1) copy + threshold
    float a[J];
#pragma unroll
    for (int i = 0; i < J; ++i) a[i] = inf;
    const bool vec = (n & 3) == 0 && ((((uintptr_t)src) | ((uintptr_t)o) | ((uintptr_t)oi)) & 15u) == 0;
    if (vec) {
        const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
        float4* o4 = reinterpret_cast<float4*>(o);
        int4* oi4 = reinterpret_cast<int4*>(oi);
#pragma unroll 4
        for (int i = lane; i < (n >> 2); i += 32) {
            const float4 v = __ldg(s4 + i);
            o4[i] = v;
            oi4[i] = make_int4(4 * i, 4 * i + 1, 4 * i + 2, 4 * i + 3);
            ss_insert<J>(a, v.x); ss_insert<J>(a, v.y); ss_insert<J>(a, v.z); ss_insert<J>(a, v.w);
        }
    } else {
#pragma unroll 4
        for (int i = lane; i < n; i += 32) {
            const float v = __ldg(src + i);
            o[i] = v;
            oi[i] = i;
            ss_insert<J>(a, v);
        }
    }
    float T = inf;
    if (n > SSF_CAP) {   // every lane saw at least SSF_CAP / 32 >= J entries: the 32 J kept values are row entries, 32 J >= k + 32
        unsigned key[J];
#pragma unroll
        for (int i = 0; i < J; ++i) key[i] = ss_key(a[i]);
        // the smallest K with at least k keys <= K, i.e. the k-th smallest key
        unsigned lo = __reduce_min_sync(0xffffffffu, key[0]), hi = __reduce_max_sync(0xffffffffu, key[J - 1]);
        while (lo < hi) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            int c = 0;
#pragma unroll
            for (int i = 0; i < J; ++i) c += __popc(__ballot_sync(0xffffffffu, key[i] <= mid));
            if (c >= k) hi = mid; else lo = mid + 1;
        }
        T = ss_unkey(lo);   // +inf when fewer than k kept values are finite: everything is listed, the list overflows, plain sort
    }
    __syncwarp();

    // (2) the list P in position order
    int cnt = 0;
#pragma unroll 4
    for (int base = 0; base < n; base += 32) {
        const int t = base + lane;
        const float v = t < n ? __ldg(src + t) : inf;
        const bool in = t < n && (t < k || v <= T);
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int slot = cnt + __popc(bal & ((1u << lane) - 1u));
        if (in && slot < SSF_CAP) {
            s_pos[warp][slot] = t;
            s_val[warp][slot] = v;
        }
        cnt += __popc(bal);
    }
    __syncwarp();

    if (cnt > SSF_CAP) {
        // too many entries at or below the threshold: plain selection sort on the output row (ss_plain_sort, as selection_sort_kernel)
        ss_plain_sort(o, oi, n, k, lane);
        return;
    }
    for (int i = lane; i < cnt; i += 32) s_id[warp][i] = i;

    // (3) k steps on the list; slot s is position s because [0, k) leads the list
    if (cnt <= 32 * SSF_REG) {
        // keys in registers: lane holds slots lane + 32 r.  Only lane 0 touches s_id until the write-back.
        unsigned kr[SSF_REG];
#pragma unroll
        for (int r = 0; r < SSF_REG; ++r) kr[r] = lane + 32 * r < cnt ? ss_key(s_val[warp][lane + 32 * r]) : 0xffffffffu;
        __syncwarp();
        for (int s = 0; s < k; ++s) {
            unsigned bk = 0xffffffffu;
            int bs = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < SSF_REG; ++r) {
                const int slot = lane + 32 * r;
                if (slot >= s && kr[r] < bk) { bk = kr[r]; bs = slot; }   // ascending slot: the first of equals stays
            }
            const unsigned mk = __reduce_min_sync(0xffffffffu, bk);
            if (mk == 0xffffffffu) continue;   // only NaN left at or behind s: the reference's scan moves nothing
            const int ms = __reduce_min_sync(0xffffffffu, bk == mk ? bs : 0x7fffffff);
            const int rs = s >> 5, rm = ms >> 5;   // warp-uniform
            unsigned ksel = kr[0];
#pragma unroll
            for (int r = 1; r < SSF_REG; ++r) ksel = rs == r ? kr[r] : ksel;
            const unsigned ks = __shfl_sync(0xffffffffu, ksel, s & 31);
            if (ms == s || ks == 0xffffffffu) continue;   // already in place / a NaN at position s stays there
            if (lane == (ms & 31)) {
#pragma unroll
                for (int r = 0; r < SSF_REG; ++r) kr[r] = rm == r ? ks : kr[r];
            }
            if (lane == 0) {
                const int is = s_id[warp][s], im = s_id[warp][ms];
                s_id[warp][s] = im;
                s_id[warp][ms] = is;
            }
        }
        __syncwarp();
    } else {
        for (int i = lane; i < cnt; i += 32) s_key[warp][i] = ss_key(s_val[warp][i]);
        __syncwarp();
        for (int s = 0; s < k; ++s) {
            unsigned bk = 0xffffffffu;
            int bs = 0x7fffffff;
            for (int i = s + lane; i < cnt; i += 32) {
                const unsigned key = s_key[warp][i];
                if (key < bk || bs == 0x7fffffff) { bk = key; bs = i; }   // ascending i: the first of equals stays
            }
            const unsigned mk = __reduce_min_sync(0xffffffffu, bk);
            const int ms = __reduce_min_sync(0xffffffffu, (bk == mk && bs != 0x7fffffff) ? bs : 0x7fffffff);
            __syncwarp();   // every lane's reads of this step precede lane 0's writes (the reductions imply it; this states it)
            if (lane == 0 && ms != s) {
                const unsigned ks = s_key[warp][s];
                if (ks != 0xffffffffu) {   // a NaN at position s stays there (nothing compares below it)
                    const int is = s_id[warp][s], im = s_id[warp][ms];
                    s_key[warp][s] = mk; s_id[warp][s] = im;
                    s_key[warp][ms] = ks; s_id[warp][ms] = is;
                }
            }
            __syncwarp();
        }
    }

    // (4) write back what moved
    for (int i = lane; i < cnt; i += 32) {
        const int e = s_id[warp][i];
        if (e != i) {
            const int p = s_pos[warp][i];
            o[p] = s_val[warp][e];
            oi[p] = s_pos[warp][e];
        }
    }
}

template <int J>
static void launch_selection_sort_fast(int n, int k, size_t rows, const float* dist, int* outi, float* out, cudaStream_t s) {
    const size_t blocks = (rows + SSF_WARPS - 1) / SSF_WARPS;
    selection_sort_fast_kernel<J><<<(unsigned)blocks, SSF_WARPS * 32, 0, s>>>(n, k, rows, dist, outi, out);
}

}  // namespace rfnet

using namespace rfnet;

extern "C" size_t rfnet_query_ball_point_workspace_bytes(int b, int n, int m) {
    (void)m;
    if (b <= 0 || n < BG_MIN_POINTS || n > BG_MAX_POINTS) return 0;
    return ball_grid_stride(n) * (size_t)b;
}

extern "C" int rfnet_query_ball_point(int b, int n, int m, const float* radius, int nsample, const float* xyz1, const float* xyz2,
                                      int* idx, int* pts_cnt, void* workspace, size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && nsample > 0);  // tf_grouping.cpp:75
    if (b == 0 || m == 0) return 0;
    RFNET_CHECK_ARG(radius && xyz2 && idx && pts_cnt && (n == 0 || xyz1));
    RFNET_CHECK_ARG(b <= 65535);
    cudaStream_t s = (cudaStream_t)stream;
    const int per_block = BQ_WARPS * BQ_QPW;
    dim3 grid((unsigned)((m + per_block - 1) / per_block), (unsigned)b);
    if (workspace && n >= BG_MIN_POINTS && n <= BG_MAX_POINTS && workspace_bytes >= rfnet_query_ball_point_workspace_bytes(b, n, m)) {
        const size_t stride = ball_grid_stride(n);
        const size_t smem = sizeof(unsigned) * (BG_CELLS + BG_CELLS / 32);
        RFNET_CUDA(cudaFuncSetAttribute(ball_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ball_grid_kernel<<<b, 1024, smem, s>>>(n, radius, 32, xyz1, (unsigned char*)workspace, stride);
        ball_query_grid_kernel<<<grid, BQ_WARPS * 32, 0, s>>>(n, m, radius, nsample, xyz2, (const unsigned char*)workspace, stride, idx, pts_cnt);
        return launch_status();
    }
    ball_query_kernel<<<grid, BQ_WARPS * 32, 0, s>>>(n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
    return launch_status();
}

extern "C" int rfnet_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out,
                                 rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && c >= 0 && m >= 0 && nsample >= 0);
    const size_t rows = (size_t)b * m * nsample;
    if (rows == 0 || c == 0) return 0;
    RFNET_CHECK_ARG(n > 0 && points && idx && out);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t rpc = (size_t)m * nsample;
    RFNET_CHECK_ARG(b <= 65535 && rpc * (size_t)c < 0x7fffffffull);
    if (c % 4 == 0 && (((uintptr_t)points | (uintptr_t)out) & 15u) == 0) {
        dim3 grid((unsigned)((rpc * (c / 4) + 256 * GP_E - 1) / (256 * GP_E)), (unsigned)b);
        group_point_kernel<float4><<<grid, 256, 0, s>>>(n, c / 4, (unsigned)rpc, (const float4*)points, idx, (float4*)out);
    } else {
        dim3 grid((unsigned)((rpc * c + 256 * GP_E - 1) / (256 * GP_E)), (unsigned)b);
        group_point_kernel<float><<<grid, 256, 0, s>>>(n, c, (unsigned)rpc, points, idx, out);
    }
    return launch_status();
}

extern "C" size_t rfnet_group_point_grad_workspace_bytes(int b, int n, int c, int m, int nsample) {
    (void)c;
    if (b <= 0 || n <= 0) return 0;
    return seg::csr_bytes(b, n, (size_t)(m > 0 ? m : 0) * (nsample > 0 ? nsample : 0));
}

// the atomic-free sum over a built CSR (rfnet_group_point_grad with a workspace = build + this)
static int group_point_grad_from_csr(int b, int n, int c, size_t rpc, const float* grad_out, const seg::Csr& csr, float* grad_points, cudaStream_t s) {
    const bool vec = c % 4 == 0 && (((uintptr_t)grad_out | (uintptr_t)grad_points) & 15u) == 0;
    if (vec) {
        dim3 grid((unsigned)(((size_t)n * (c / 4) + 255) / 256), (unsigned)b);
        group_point_grad_seg_kernel<float4><<<grid, 256, 0, s>>>(n, c / 4, (unsigned)rpc, (const float4*)grad_out, csr.offset, csr.list, (float4*)grad_points);
    } else {
        dim3 grid((unsigned)(((size_t)n * c + 255) / 256), (unsigned)b);
        group_point_grad_seg_kernel<float><<<grid, 256, 0, s>>>(n, c, (unsigned)rpc, grad_out, csr.offset, csr.list, grad_points);
    }
    return launch_status();
}

extern "C" int rfnet_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx, float* grad_points,
                                      void* workspace, size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && c >= 0 && m >= 0 && nsample >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t rows = (size_t)b * m * nsample;
    const size_t rpc = (size_t)m * nsample;
    if ((size_t)b * n * c == 0) return 0;
    RFNET_CHECK_ARG(grad_points && (rows == 0 || (grad_out && idx)));
    RFNET_CHECK_ARG(b <= 65535 && rpc * (size_t)c < 0x7fffffffull && (size_t)n * c < 0x7fffffffull);
    const bool vec = c % 4 == 0 && (((uintptr_t)grad_out | (uintptr_t)grad_points) & 15u) == 0;
    if (workspace) {
        // atomic-free, deterministic path
        RFNET_CHECK_ARG(workspace_bytes >= rfnet_group_point_grad_workspace_bytes(b, n, c, m, nsample));
        seg::Csr csr = seg::csr_carve(workspace, b, n, rpc);
        const int rc = seg::csr_build(csr, b, n, rpc, idx, s);
        if (rc) return rc;
        return group_point_grad_from_csr(b, n, c, rpc, grad_out, csr, grad_points, s);
    }
    // no workspace: the reference's formulation (zero-fill + float reductions), order-dependent in the last bits
    RFNET_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * n * c, s));
    if (rows == 0) return 0;
    if (vec) {
        dim3 grid((unsigned)((rpc * (c / 4) + 255) / 256), (unsigned)b);
        group_point_grad_v4_kernel<<<grid, 256, 0, s>>>(n, c / 4, (unsigned)rpc, (const float4*)grad_out, idx, (float4*)grad_points);
    } else {
        dim3 grid((unsigned)((rpc * c + 255) / 256), (unsigned)b);
        group_point_grad_kernel<<<grid, 256, 0, s>>>(n, c, (unsigned)rpc, grad_out, idx, grad_points);
    }
    return launch_status();
}

// ---- scatter plans: the inverted index of an idx tensor, built once (e.g. while the forward gather runs) and reused by every
// gradient that scatters through the same idx.  idx is read as (b, rows) targets in [0, n_targets).
extern "C" size_t rfnet_scatter_plan_bytes(int b, int n_targets, int rows) {
    if (b <= 0 || n_targets <= 0 || rows < 0) return 0;
    return seg::csr_bytes(b, n_targets, (size_t)rows);
}

extern "C" int rfnet_scatter_plan_build(int b, int n_targets, int rows, const int* idx, void* plan, size_t plan_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n_targets >= 0 && rows >= 0);
    if (b == 0 || n_targets == 0) return 0;
    RFNET_CHECK_ARG(plan && plan_bytes >= rfnet_scatter_plan_bytes(b, n_targets, rows) && (rows == 0 || idx) && b <= 65535);
    seg::Csr csr = seg::csr_carve(plan, b, n_targets, (size_t)rows);
    return seg::csr_build(csr, b, n_targets, (size_t)rows, idx, (cudaStream_t)stream);
}

extern "C" int rfnet_group_point_grad_planned(int b, int n, int c, int m, int nsample, const float* grad_out, const void* plan, size_t plan_bytes,
                                              float* grad_points, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && c >= 0 && m >= 0 && nsample >= 0);
    const size_t rpc = (size_t)m * nsample;
    if ((size_t)b * n * c == 0) return 0;
    RFNET_CHECK_ARG(grad_points && plan && (rpc == 0 || grad_out) && rpc < 0x7fffffffull && plan_bytes >= rfnet_scatter_plan_bytes(b, n, (int)rpc));
    RFNET_CHECK_ARG(b <= 65535 && rpc * (size_t)c < 0x7fffffffull && (size_t)n * c < 0x7fffffffull);
    const seg::Csr csr = seg::csr_carve(const_cast<void*>(plan), b, n, rpc);
    return group_point_grad_from_csr(b, n, c, rpc, grad_out, csr, grad_points, (cudaStream_t)stream);
}

extern "C" int rfnet_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && k > 0);  // tf_grouping.cpp:117
    const size_t rows = (size_t)b * m;
    if (rows == 0 || n == 0) return 0;
    RFNET_CHECK_ARG(dist && outi && out);
    if (k > n) k = n;
    const size_t blocks = (rows + SS_WARPS - 1) / SS_WARPS;
    RFNET_CHECK_ARG(blocks <= 0x7fffffffull);
    if (k <= 32 * SSF_JMAX) {
        const int j = (k + 31) / 32 + 1;   // values kept per lane
        if (j == 2) launch_selection_sort_fast<2>(n, k, rows, dist, outi, out, (cudaStream_t)stream);
        else if (j == 3) launch_selection_sort_fast<3>(n, k, rows, dist, outi, out, (cudaStream_t)stream);
        else if (j == 4) launch_selection_sort_fast<4>(n, k, rows, dist, outi, out, (cudaStream_t)stream);
        else launch_selection_sort_fast<5>(n, k, rows, dist, outi, out, (cudaStream_t)stream);
        return launch_status();
    }
    const size_t smem = (size_t)SS_WARPS * n * 8;   // value + position per entry, one row per warp
    const int in_smem = smem <= 160 * 1024;
    if (in_smem && smem > 48 * 1024)
        RFNET_CUDA(cudaFuncSetAttribute(selection_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selection_sort_kernel<<<(unsigned)blocks, SS_WARPS * 32, in_smem ? smem : 0, (cudaStream_t)stream>>>(n, k, rows, in_smem, dist, outi, out);
    return launch_status();
}
