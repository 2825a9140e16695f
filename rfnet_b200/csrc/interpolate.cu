// three_nn, three_interpolate and its gradient for sm_100a.
//
// The reference has NO GPU code for these ops (tf_ops/interpolation/tf_interpolate.cpp registers DEVICE_CPU only,
// :187,222,262); these are the first GPU kernels for them and follow threenn_cpu / threeinterpolate_cpu /
// threeinterpolate_grad_cpu (tf_interpolate.cpp:60-153) bit for bit where order matters:
//   * three_nn evaluates d2 UNFUSED ((dx*dx + dy*dy) + dz*dz), as the reference's CPU build does, and keeps the three
//     smallest with the same strict-'<' insertion, so earlier indices win ties;
//   * three_interpolate evaluates (p1*w1 + p2*w2) + p3*w3 unfused.
// three_nn uses the same machinery as nn_distance: queries in registers as packed pairs, candidates broadcast from shared
// memory, a branch-free scan that only LISTS the candidate groups able to change a top 3, and an exact pass over the list.

#include "common.cuh"
#include "pointgrid.cuh"
#include "rfnet_ops.h"
#include "segscatter.cuh"

namespace rfnet {

constexpr int TN_THREADS = 128;
constexpr int TN_Q = 4;        // queries per thread (two packed pairs)
constexpr int TN_TILE = 2048;  // candidates per shared-memory tile

struct Top3 {
    float d1, d2, d3;
    int i1, i2, i3;
};
__device__ __forceinline__ void top3_insert(Top3& t, float d, int k) {  // tf_interpolate.cpp:76-92
    if (d < t.d1) {
        t.d3 = t.d2; t.i3 = t.i2; t.d2 = t.d1; t.i2 = t.i1; t.d1 = d; t.i1 = k;
    } else if (d < t.d2) {
        t.d3 = t.d2; t.i3 = t.i2; t.d2 = d; t.i2 = k;
    } else if (d < t.d3) {
        t.d3 = d; t.i3 = k;
    }
}

// the same insertion without branches (3 compares, 5 min/max, 5 selects): identical results, including which index stays
// ahead among equal distances (strict '<' everywhere)
// A NaN distance fails every '<' of the reference's insertion and is ignored there; fminf/fmaxf would instead shift the
// stored distances (fmaxf(d2, NaN) == d2).  NaN is therefore replaced by +inf, which also fails every strict compare
// against finite values and against the 1e40 -> +inf seeds.
__device__ __forceinline__ void top3_insert_bf(Top3& t, float d, int k) {
    d = (d != d) ? __int_as_float(0x7f800000) : d;
    const bool p1 = d < t.d1, p2 = d < t.d2, p3 = d < t.d3;
    t.i3 = p3 ? (p2 ? t.i2 : k) : t.i3;
    t.i2 = p2 ? (p1 ? t.i1 : k) : t.i2;
    t.i1 = p1 ? k : t.i1;
    t.d3 = fminf(t.d3, fmaxf(t.d2, d));
    t.d2 = fminf(t.d2, fmaxf(t.d1, d));
    t.d1 = fminf(t.d1, d);
}

// Two phases per shared-memory tile, so that the hot loop has no data-dependent branch (a top-3 insertion test fires in
// almost every 32-lane step: 128 queries per warp x ~16 insertions each over 2048 candidates):
//   scan     per group of 8 candidates: FUSED packed distances (3 packed FP32 issues per query pair and candidate), their
//            minimum g, and a branch-free record "g <= 3rd smallest group minimum so far (+1e-6 relative)" appended to a
//            per-query list in shared memory.  The 3rd smallest group minimum is an upper bound of the true 3rd smallest
//            distance, and fused / unfused evaluation differ by < 3e-7 relative, so every group holding a candidate
//            that the reference's scan would insert is on the list (about 3 ln(groups) entries per query).
//            Margin (u = 2^-24).  Both evaluations start from the SAME rounded differences dx, dy, dz; on those, the unfused
//            sum ((dx*dx)+(dy*dy))+(dz*dz) and the fused fma(dz,dz, fma(dx,dx, dy*dy)) are each within 3u relative of the exact
//            dx^2+dy^2+dz^2 (one product and at most two sum roundings per non-negative term).  The reference inserts
//            candidate c when unf(c) < d3, its running 3rd smallest unfused distance, and d3 <= G3_unf, the 3rd smallest
//            unfused group minimum of the groups before c's.  So fused(c) <= unf(c)(1+6u) < G3_unf (1+6u) <= G3_fused (1+12u),
//            and c's group minimum gm <= fused(c).  The list test is gm <= fl(G3_fused * 1.000001f) with 1.000001f = 1 + 18u
//            and one more rounding: >= G3_fused (1 + 16.9u) > G3_fused (1 + 12u): every such group is listed.
//   resolve  the listed groups, in scan order, go through the reference's exact code: UNFUSED distance, strict-'<'
//            insertion.  Unlisted groups cannot change the top 3, so the state after each tile is the reference's.
// A query whose list overflows (adversarially ordered candidates) re-scans the tile with the exact code.
constexpr int TN_MIN_CTAS = 4;
constexpr int TN_CAP = 32;     // list entries per query and tile (u8 group numbers)
constexpr int TN_GROUPS = TN_TILE / 8;
static_assert(TN_GROUPS <= 256, "group numbers are stored as bytes");

// exact code of the reference on one group of 8 candidates (SoA tile): unfused distance on packed candidate pairs
// (FADD2 / FMUL2, the two additions scalar so that nothing is contracted), strict-'<' insertion in candidate order
template <bool BRANCH_FREE>
__device__ __forceinline__ void tn_exact_group(Top3& t, float qx, float qy, float qz, const float* __restrict__ tx, const float* __restrict__ ty,
                                               const float* __restrict__ tz, int g, int base, bool valid) {
    const float4 xa = *reinterpret_cast<const float4*>(tx + g * 8), xb = *reinterpret_cast<const float4*>(tx + g * 8 + 4);
    const float4 ya = *reinterpret_cast<const float4*>(ty + g * 8), yb = *reinterpret_cast<const float4*>(ty + g * 8 + 4);
    const float4 za = *reinterpret_cast<const float4*>(tz + g * 8), zb = *reinterpret_cast<const float4*>(tz + g * 8 + 4);
    const float2 cx[4] = {make_float2(xa.x, xa.y), make_float2(xa.z, xa.w), make_float2(xb.x, xb.y), make_float2(xb.z, xb.w)};
    const float2 cy[4] = {make_float2(ya.x, ya.y), make_float2(ya.z, ya.w), make_float2(yb.x, yb.y), make_float2(yb.z, yb.w)};
    const float2 cz[4] = {make_float2(za.x, za.y), make_float2(za.z, za.w), make_float2(zb.x, zb.y), make_float2(zb.z, zb.w)};
    const float2 q2x = make_float2(qx, qx), q2y = make_float2(qy, qy), q2z = make_float2(qz, qz);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 d = sqdist3x2<false>(__fadd2_rn(q2x, make_float2(-cx[j].x, -cx[j].y)), __fadd2_rn(q2y, make_float2(-cy[j].x, -cy[j].y)),
                                          __fadd2_rn(q2z, make_float2(-cz[j].x, -cz[j].y)));
        if (BRANCH_FREE) {   // an invalid entry (this query's list is shorter than its neighbours') inserts +inf: a no-op
            top3_insert_bf(t, valid ? d.x : __int_as_float(0x7f800000), base + g * 8 + 2 * j);
            top3_insert_bf(t, valid ? d.y : __int_as_float(0x7f800000), base + g * 8 + 2 * j + 1);
        } else {
            if (d.x < t.d3) top3_insert(t, d.x, base + g * 8 + 2 * j);
            if (d.y < t.d3) top3_insert(t, d.y, base + g * 8 + 2 * j + 1);
        }
    }
}

__global__ void __launch_bounds__(TN_THREADS, TN_MIN_CTAS) three_nn_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                              float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ __align__(16) float tx[TN_TILE], ty[TN_TILE], tz[TN_TILE];   // candidates, SoA
    // [query][entry][thread]; one extra entry row per query takes the stores of an overflowing list
    __shared__ unsigned char list[TN_Q * (TN_CAP + 1) * TN_THREADS];
    const int cloud = blockIdx.y;
    const int tid = threadIdx.x;
    const float* __restrict__ qbase = xyz1 + (size_t)cloud * n * 3;
    const float* __restrict__ cbase = xyz2 + (size_t)cloud * m * 3;
    const int q0 = blockIdx.x * (TN_THREADS * TN_Q) + tid;

    float2 qx[TN_Q / 2], qy[TN_Q / 2], qz[TN_Q / 2];
    Top3 top[TN_Q];
#pragma unroll
    for (int h = 0; h < TN_Q / 2; ++h) {
        const int ia = q0 + (2 * h) * TN_THREADS, ib = ia + TN_THREADS;
        const bool va = ia < n, vb = ib < n;
        qx[h].x = va ? qbase[(size_t)ia * 3 + 0] : 0.f; qy[h].x = va ? qbase[(size_t)ia * 3 + 1] : 0.f; qz[h].x = va ? qbase[(size_t)ia * 3 + 2] : 0.f;
        qx[h].y = vb ? qbase[(size_t)ib * 3 + 0] : 0.f; qy[h].y = vb ? qbase[(size_t)ib * 3 + 1] : 0.f; qz[h].y = vb ? qbase[(size_t)ib * 3 + 2] : 0.f;
    }
    const float inf = __int_as_float(0x7f800000);  // (float)1e40 of the reference
    float ga[TN_Q], gb[TN_Q], gc[TN_Q];             // three smallest group minima so far (fused values), ascending
#pragma unroll
    for (int i = 0; i < TN_Q; ++i) {
        top[i].d1 = top[i].d2 = top[i].d3 = inf;
        top[i].i1 = top[i].i2 = top[i].i3 = 0;
        ga[i] = gb[i] = gc[i] = inf;
    }

    for (int t0 = 0; t0 < m; t0 += TN_TILE) {
        const int len = min(TN_TILE, m - t0);
        const int len8 = (len + 7) & ~7;
        __syncthreads();
        // AoS -> SoA; eight independent loads in flight per thread before the first store
#pragma unroll 1
        for (int i0 = 0; i0 < len * 3; i0 += TN_THREADS * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * TN_THREADS + tid;
                v[u] = i < len * 3 ? __ldg(cbase + (size_t)t0 * 3 + i) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * TN_THREADS + tid;
                const int c = i / 3, a = i - c * 3;
                if (i < len * 3) (a == 0 ? tx : (a == 1 ? ty : tz))[c] = v[u];
            }
        }
        for (int i = len + tid; i < len8; i += TN_THREADS) tx[i] = ty[i] = tz[i] = inf;  // padded candidates: d2 = inf, never inserted
        __syncthreads();
        unsigned lp[TN_Q];   // shared-memory address of the next list entry of each query (advances by TN_THREADS per record)
        unsigned sink[TN_Q];
        const unsigned lbase = smem_u32(list) + tid;
#pragma unroll
        for (int i = 0; i < TN_Q; ++i) {
            lp[i] = lbase + i * (TN_CAP + 1) * TN_THREADS;
            sink[i] = lp[i] + TN_CAP * TN_THREADS;
        }
        const int ngroups = len8 >> 3;
#pragma unroll 1
        for (int g = 0; g < ngroups; ++g) {
            const float4 xa = *reinterpret_cast<const float4*>(tx + g * 8), xb = *reinterpret_cast<const float4*>(tx + g * 8 + 4);
            const float4 ya = *reinterpret_cast<const float4*>(ty + g * 8), yb = *reinterpret_cast<const float4*>(ty + g * 8 + 4);
            const float4 za = *reinterpret_cast<const float4*>(tz + g * 8), zb = *reinterpret_cast<const float4*>(tz + g * 8 + 4);
            const float cx[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float cy[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
            const float cz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
            for (int h = 0; h < TN_Q / 2; ++h) {
                float2 d[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 dx = __fadd2_rn(qx[h], make_float2(-cx[j], -cx[j]));
                    const float2 dy = __fadd2_rn(qy[h], make_float2(-cy[j], -cy[j]));
                    const float2 dz = __fadd2_rn(qz[h], make_float2(-cz[j], -cz[j]));
                    d[j] = sqdist3x2<true>(dx, dy, dz);
                }
                const float gmin[2] = {fmin3(fmin3(d[0].x, d[1].x, d[2].x), fmin3(d[3].x, d[4].x, d[5].x), fminf(d[6].x, d[7].x)),
                                       fmin3(fmin3(d[0].y, d[1].y, d[2].y), fmin3(d[3].y, d[4].y, d[5].y), fminf(d[6].y, d[7].y))};
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int qi = 2 * h + s;
                    const float gm = gmin[s];
                    // record: predicated store + pointer bump, no branch.  Past TN_CAP entries the store lands in the spare row.
                    const unsigned a = min(lp[qi], sink[qi]);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %2, %3;\n\t@p st.shared.u8 [%1], %4;\n\t@p add.u32 %0, %0, %5;\n\t}"
                                 : "+r"(lp[qi]) : "r"(a), "f"(gm), "f"(gc[qi] * 1.000001f), "r"(g), "n"(TN_THREADS) : "memory");
                    gc[qi] = fminf(gc[qi], fmaxf(gb[qi], gm));
                    gb[qi] = fminf(gb[qi], fmaxf(ga[qi], gm));
                    ga[qi] = fminf(ga[qi], gm);
                }
            }
        }
        // resolve: the TN_Q lists of a thread advance together, so the four dependent insertion chains overlap
        int cn[TN_Q], emax = 0;
#pragma unroll
        for (int i = 0; i < TN_Q; ++i) {
            cn[i] = (int)((lp[i] - (lbase + i * (TN_CAP + 1) * TN_THREADS)) / TN_THREADS);
            if (cn[i] > TN_CAP) {   // overflow: exact re-scan of the tile for this query
                const float x = (i & 1) ? qx[i >> 1].y : qx[i >> 1].x, y = (i & 1) ? qy[i >> 1].y : qy[i >> 1].x, z = (i & 1) ? qz[i >> 1].y : qz[i >> 1].x;
                for (int g = 0; g < ngroups; ++g) tn_exact_group<false>(top[i], x, y, z, tx, ty, tz, g, t0, true);
                cn[i] = 0;
            }
            emax = max(emax, cn[i]);
        }
#pragma unroll 1
        for (int e = 0; e < emax; ++e) {
#pragma unroll
            for (int i = 0; i < TN_Q; ++i) {
                const float x = (i & 1) ? qx[i >> 1].y : qx[i >> 1].x, y = (i & 1) ? qy[i >> 1].y : qy[i >> 1].x, z = (i & 1) ? qz[i >> 1].y : qz[i >> 1].x;
                const bool v = e < cn[i];
                const int g = v ? (int)list[(i * (TN_CAP + 1) + e) * TN_THREADS + tid] : 0;
                tn_exact_group<true>(top[i], x, y, z, tx, ty, tz, g, t0, v);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TN_Q; ++i) {
        const int qi = q0 + i * TN_THREADS;
        if (qi < n) {
            const size_t o = ((size_t)cloud * n + qi) * 3;
            dist[o] = top[i].d1; dist[o + 1] = top[i].d2; dist[o + 2] = top[i].d3;
            idx[o] = top[i].i1; idx[o + 1] = top[i].i2; idx[o + 2] = top[i].i3;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// three_nn over a uniform grid (used when the caller passes a workspace and the known cloud has TG_MIN_POINTS ..
// TG_MAX_POINTS points).  The known points are binned by ball_grid_kernel (pointgrid.cuh, ~2 points per cell, <= 16 cells
// per axis); the grid (cell offsets, coordinates and indices in cell order) is staged in shared memory and every THREAD owns
// one query: it visits its own cell, then the shells of cells around it, keeping the three smallest (distance, index)
// pairs, and stops as soon as the third distance is strictly below the distance to the nearest face of the block of cells
// visited so far -- no unvisited point can then enter, not even on a tie.
// Same arithmetic as the scan kernel (unfused d2 of query + (-point)); the reference's "strict '<' while scanning in index
// order" is the lexicographic order on (distance, index), which is what the insertion here uses, so ties resolve identically.
// ~80 candidates per query instead of m.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TG_MIN_POINTS = 512;
constexpr int TG_MAX_POINTS = 4096;
constexpr int TG_GMAX = 16;
constexpr int TG_THREADS = 256;

__device__ __forceinline__ void top3_insert_lex(Top3& t, float d, int k) {
    const bool p1 = d < t.d1 || (d == t.d1 && k < t.i1);
    const bool p2 = d < t.d2 || (d == t.d2 && k < t.i2);
    const bool p3 = d < t.d3 || (d == t.d3 && k < t.i3);
    if (p1) {
        t.d3 = t.d2; t.i3 = t.i2; t.d2 = t.d1; t.i2 = t.i1; t.d1 = d; t.i1 = k;
    } else if (p2) {
        t.d3 = t.d2; t.i3 = t.i2; t.d2 = d; t.i2 = k;
    } else if (p3) {
        t.d3 = d; t.i3 = k;
    }
}

__global__ void __launch_bounds__(TG_THREADS) three_nn_grid_kernel(int n, int m, const float* __restrict__ xyz1, const unsigned char* __restrict__ ws,
                                                                   size_t stride, float* __restrict__ dist, int* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    const int cloud = blockIdx.y;
    const BallGridView v = ball_grid_view(ws, stride, cloud, m);
    const BallGrid g = *v.g;
    const int ncells = g.G[0] * g.G[1] * g.G[2];
    int* sStart = reinterpret_cast<int*>(tg_smem);                       // ncells + 1
    int* sIdx = sStart + (TG_GMAX * TG_GMAX * TG_GMAX + 1);              // m
    float* sXyz = reinterpret_cast<float*>(sIdx + m);                    // 3 m
    for (int i = threadIdx.x; i <= ncells; i += TG_THREADS) sStart[i] = v.cell_start[i];
    for (int i = threadIdx.x; i < m; i += TG_THREADS) sIdx[i] = v.sorted_idx[i];
    for (int i = threadIdx.x; i < 3 * m; i += TG_THREADS) sXyz[i] = v.sorted_xyz[i];
    __syncthreads();
    const int qi = blockIdx.x * TG_THREADS + threadIdx.x;
    if (qi >= n) return;
    const float* q = xyz1 + ((size_t)cloud * n + qi) * 3;
    const float qc[3] = {q[0], q[1], q[2]};
    int cq[3];
    float side[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        cq[a] = min(g.G[a] - 1, max(0, (int)floorf((qc[a] - g.lo[a]) * g.scale[a])));
        side[a] = g.scale[a] > 0.f ? 1.0f / g.scale[a] : 0.f;
    }
    const float inf = __int_as_float(0x7f800000);
    Top3 t;
    t.d1 = t.d2 = t.d3 = inf;
    t.i1 = t.i2 = t.i3 = 0;
    const int rmax = max(max(g.G[0], g.G[1]), g.G[2]);
    for (int rho = 0; rho < rmax; ++rho) {
        const int x0 = max(0, cq[0] - rho), x1 = min(g.G[0] - 1, cq[0] + rho);
        const int y0 = max(0, cq[1] - rho), y1 = min(g.G[1] - 1, cq[1] + rho);
        const int z0 = max(0, cq[2] - rho), z1 = min(g.G[2] - 1, cq[2] + rho);
        for (int cx = x0; cx <= x1; ++cx)
            for (int cy = y0; cy <= y1; ++cy) {
                const bool edge_xy = abs(cx - cq[0]) == rho || abs(cy - cq[1]) == rho;
                const int cbase = (cx * g.G[1] + cy) * g.G[2];
                // cells of this (cx, cy) column that belong to the shell: the whole z range on the rim, else only the two caps
                for (int part = 0; part < 2; ++part) {
                    int za, zb;
                    if (edge_xy) {
                        if (part) break;
                        za = z0; zb = z1;
                    } else {
                        const int zc = part ? cq[2] + rho : cq[2] - rho;
                        if (zc < 0 || zc >= g.G[2] || (part && rho == 0)) continue;
                        za = zb = zc;
                    }
                    const int s = sStart[cbase + za], e = sStart[cbase + zb + 1];
                    for (int p2 = s; p2 < e; ++p2) {
                        const float d = sqdist3<false>(__fadd_rn(qc[0], -sXyz[p2 * 3]), __fadd_rn(qc[1], -sXyz[p2 * 3 + 1]), __fadd_rn(qc[2], -sXyz[p2 * 3 + 2]));
                        top3_insert_lex(t, d, sIdx[p2]);
                    }
                }
            }
        // distance from the query to the nearest face of the visited block that has unvisited cells behind it
        float bound = inf;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (cq[a] - rho > 0) bound = fminf(bound, qc[a] - (g.lo[a] + (float)(cq[a] - rho) * side[a]));
            if (cq[a] + rho < g.G[a] - 1) bound = fminf(bound, (g.lo[a] + (float)(cq[a] + rho + 1) * side[a]) - qc[a]);
        }
        if (bound == inf) break;                       // the block covers the whole grid
        bound = fmaxf(bound, 0.f) * 0.99999f;          // float rounding of the face positions
        if (t.d3 < bound * bound) break;               // strictly closer than anything unvisited
    }
    const size_t o = ((size_t)cloud * n + qi) * 3;
    dist[o] = t.d1; dist[o + 1] = t.d2; dist[o + 2] = t.d3;
    idx[o] = t.i1; idx[o + 1] = t.i2; idx[o + 2] = t.i3;
}

// ---------------------------------------------------------------------------------------------------------------
// knn_point for 3-d points.  The reference builds the full (b, m, n) distance matrix with framework
// ops and calls tf.nn.top_k(-dist) "ONLY SUPPORT CPU" (tf_ops/grouping/tf_grouping.py:48-73); this kernel never
// materialises the matrix: one thread per query keeps its K best (distance, index) sorted in registers, candidates are
// broadcast from shared memory, and the insertion code runs only when a candidate beats the current K-th best.
// dist = ((dx*dx + dy*dy) + dz*dz) unfused; equal distances keep the lower index first, as top_k does.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 2048;
template <int K>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(int n, int m, int k, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                          float* __restrict__ val, int* __restrict__ idx) {
    __shared__ __align__(16) float tile[KNN_TILE * 3];
    const int cloud = blockIdx.y;
    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const float* __restrict__ data = xyz1 + (size_t)cloud * n * 3;
    const bool valid = q < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        const float* p = xyz2 + ((size_t)cloud * m + q) * 3;
        qx = p[0]; qy = p[1]; qz = p[2];
    }
    const float inf = __int_as_float(0x7f800000);
    float bd[K];
    int bi[K];
#pragma unroll
    for (int t = 0; t < K; ++t) { bd[t] = inf; bi[t] = 0; }
    for (int t0 = 0; t0 < n; t0 += KNN_TILE) {
        const int len = min(KNN_TILE, n - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < len * 3; i += KNN_THREADS) tile[i] = data[(size_t)t0 * 3 + i];
        __syncthreads();
        if (!valid) continue;
        for (int c = 0; c < len; ++c) {
            const float d = sqdist3<false>(tile[c * 3] - qx, tile[c * 3 + 1] - qy, tile[c * 3 + 2] - qz);
            if (d < bd[K - 1]) {
                // insert after every entry <= d (stable: earlier index first on ties), shifting the tail down
                float cd = d;
                int ci = t0 + c;
                bool shifting = false;  // once the new entry is placed, everything behind it moves down one slot unconditionally
#pragma unroll
                for (int t = 0; t < K; ++t) {
                    if (shifting || cd < bd[t]) {
                        const float td = bd[t]; const int ti = bi[t];
                        bd[t] = cd; bi[t] = ci;
                        cd = td; ci = ti;
                        shifting = true;
                    }
                }
            }
        }
    }
    if (!valid) return;
    float* v = val + ((size_t)cloud * m + q) * k;
    int* o = idx + ((size_t)cloud * m + q) * k;
#pragma unroll
    for (int t = 0; t < K; ++t)
        if (t < k) { v[t] = -bd[t]; o[t] = bi[t]; }   // top_k(-dist): values are the NEGATED squared distances (tf_grouping.py:72)
}

// out[i,j,l] = (p[i1,l]*w1 + p[i2,l]*w2) + p[i3,l]*w3        (tf_interpolate.cpp:107-127).  One thread per output vector.
template <typename VEC>
__device__ __forceinline__ VEC blend3(VEC a, VEC b, VEC c, float w1, float w2, float w3);
template <>
__device__ __forceinline__ float blend3<float>(float a, float b, float c, float w1, float w2, float w3) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a, w1), __fmul_rn(b, w2)), __fmul_rn(c, w3));
}
template <>
__device__ __forceinline__ float4 blend3<float4>(float4 a, float4 b, float4 c, float w1, float w2, float w3) {
    return make_float4(blend3<float>(a.x, b.x, c.x, w1, w2, w3), blend3<float>(a.y, b.y, c.y, w1, w2, w3),
                       blend3<float>(a.z, b.z, c.z, w1, w2, w3), blend3<float>(a.w, b.w, c.w, w1, w2, w3));
}
// grid.y = cloud: 32-bit index arithmetic inside a cloud.  TI_E output vectors per thread: the 3 * TI_E indices and weights
// are loaded first, then the 3 * TI_E rows, then blended and stored with streaming stores (written once, never re-read here).
constexpr int TI_E = 4;
template <typename VEC>
__device__ __forceinline__ void ti_store(VEC* p, VEC v);
template <>
__device__ __forceinline__ void ti_store<float4>(float4* p, float4 v) { __stcs(p, v); }
template <>
__device__ __forceinline__ void ti_store<float>(float* p, float v) { __stcs(p, v); }
template <typename VEC>
__global__ void __launch_bounds__(256) three_interpolate_kernel(int m, int cv, int n, const VEC* __restrict__ points, const int* __restrict__ idx,
                                                               const float* __restrict__ weight, VEC* __restrict__ out) {
    const size_t cloud = blockIdx.y;
    const unsigned total = (unsigned)n * (unsigned)cv;
    const unsigned e0 = blockIdx.x * (256u * TI_E) + threadIdx.x;
    const int* __restrict__ id = idx + cloud * (size_t)n * 3;
    const float* __restrict__ wt = weight + cloud * (size_t)n * 3;
    const VEC* __restrict__ P = points + cloud * (size_t)m * cv;
    VEC* __restrict__ O = out + cloud * (size_t)total;
    unsigned l[TI_E];
    int i0[TI_E], i1[TI_E], i2[TI_E];
    float w0[TI_E], w1[TI_E], w2[TI_E];
#pragma unroll
    for (int u = 0; u < TI_E; ++u) {
        const unsigned e = e0 + u * 256u;
        const unsigned j = e / (unsigned)cv;
        l[u] = e - j * (unsigned)cv;
        i0[u] = -1;
        if (e < total) {
            i0[u] = __ldg(id + j * 3); i1[u] = __ldg(id + j * 3 + 1); i2[u] = __ldg(id + j * 3 + 2);
            w0[u] = __ldg(wt + j * 3); w1[u] = __ldg(wt + j * 3 + 1); w2[u] = __ldg(wt + j * 3 + 2);
        }
    }
    VEC a[TI_E], b[TI_E], c[TI_E];
#pragma unroll
    for (int u = 0; u < TI_E; ++u)
        if (i0[u] >= 0) {
            a[u] = __ldg(P + (size_t)i0[u] * cv + l[u]);
            b[u] = __ldg(P + (size_t)i1[u] * cv + l[u]);
            c[u] = __ldg(P + (size_t)i2[u] * cv + l[u]);
        }
#pragma unroll
    for (int u = 0; u < TI_E; ++u)
        if (i0[u] >= 0) ti_store<VEC>(O + e0 + u * 256u, blend3<VEC>(a[u], b[u], c[u], w0[u], w1[u], w2[u]));
}

// Staged variant for wide rows (c % 16 == 0, known cloud slice fits shared memory).  The plain kernel above gathers every output
// vector's three source rows through L2: 3x the output volume in L2 reads (403 MB for 134 MB written at config 4; ncu: L2-bound,
// DRAM at 20 %).  Here a CTA owns a 16-channel slice of ONE cloud's known points in shared memory (m x 64 B) and streams a contiguous
// range of unknown points through it: the gathers are LDS.128, L2 carries the slice once per CTA, and HBM sees the algorithmic
// bytes.  The grid is persistent -- groups x slices CTAs, one per SM; the CTAs of a group walk the same rows at the same pace, each
// writing its 64-byte part of a row, so the 256-byte output rows complete in L2 close together.  Same expression, same bits.
constexpr int TIS_THREADS = 1024;
constexpr int TIS_CSV = 4;   // float4 per row slice
constexpr int TIS_U = 4;     // rows per thread and step: a CTA keeps 1024 rows (64 KB of output) in flight per memory round trip
__global__ void __launch_bounds__(TIS_THREADS, 1) three_interpolate_staged_kernel(int b, int m, int cv, int n, const float4* __restrict__ points,
                                                                                  const int* __restrict__ idx, const float* __restrict__ weight,
                                                                                  float4* __restrict__ out, unsigned long long rows_per_group) {
    extern __shared__ __align__(16) float4 tis_rows[];   // [m][TIS_CSV]
    const int nslice = cv / TIS_CSV;
    const int slice = blockIdx.x % nslice;
    const unsigned long long group = blockIdx.x / nslice;
    const unsigned long long total = (unsigned long long)b * (unsigned)n;
    unsigned long long r0 = group * rows_per_group;
    const unsigned long long r1 = min(total, r0 + rows_per_group);
    const int v = threadIdx.x & (TIS_CSV - 1), rt = threadIdx.x / TIS_CSV;
    constexpr int RPS = TIS_THREADS / TIS_CSV;   // rows per step and unroll slot
    while (r0 < r1) {
        const unsigned long long cloud = r0 / (unsigned)n;
        const unsigned long long cend = min(r1, (cloud + 1) * (unsigned)n);
        __syncthreads();   // the previous cloud's readers are done
        const float4* __restrict__ P = points + cloud * (size_t)m * cv + slice * TIS_CSV;
#pragma unroll 8
        for (int t = threadIdx.x; t < m * TIS_CSV; t += TIS_THREADS) tis_rows[t] = __ldg(P + (size_t)(t / TIS_CSV) * cv + (t & (TIS_CSV - 1)));
        __syncthreads();
        for (unsigned long long base = r0; base < cend; base += RPS * TIS_U) {
            int i0[TIS_U], i1[TIS_U], i2[TIS_U];
            float w0[TIS_U], w1[TIS_U], w2[TIS_U];
#pragma unroll
            for (int u = 0; u < TIS_U; ++u) {
                const unsigned long long r = base + u * RPS + rt;
                i0[u] = -1;
                if (r < cend) {
                    i0[u] = __ldg(idx + r * 3); i1[u] = __ldg(idx + r * 3 + 1); i2[u] = __ldg(idx + r * 3 + 2);
                    w0[u] = __ldg(weight + r * 3); w1[u] = __ldg(weight + r * 3 + 1); w2[u] = __ldg(weight + r * 3 + 2);
                }
            }
#pragma unroll
            for (int u = 0; u < TIS_U; ++u)
                if (i0[u] >= 0) {
                    const unsigned long long r = base + u * RPS + rt;
                    const float4 a = tis_rows[i0[u] * TIS_CSV + v], bb = tis_rows[i1[u] * TIS_CSV + v], c = tis_rows[i2[u] * TIS_CSV + v];
                    __stcs(out + r * cv + slice * TIS_CSV + v, blend3<float4>(a, bb, c, w0[u], w1[u], w2[u]));
                }
        }
        r0 = cend;
    }
}

// grad_points[i, i_t, l] += grad_out[i,j,l] * w_t  after zero-fill             (tf_interpolate.cpp:131-153, :258)
__global__ void three_interpolate_grad_kernel(int n, int c, int m, const float* __restrict__ grad_out, const int* __restrict__ idx,
                                              const float* __restrict__ weight, float* __restrict__ grad_points) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned j = t / (unsigned)c;
    if (j >= (unsigned)n) return;
    const unsigned l = t - j * (unsigned)c;
    const size_t cloud = blockIdx.y;
    const size_t row = cloud * n + j;
    const float g = grad_out[row * c + l];
    float* G = grad_points + cloud * (size_t)m * c + l;
#pragma unroll
    for (int u = 0; u < 3; ++u) atomicAdd(G + (size_t)idx[row * 3 + u] * c, __fmul_rn(g, weight[row * 3 + u]));
}
__global__ void three_interpolate_grad_v4_kernel(int n, int cv, int m, const float4* __restrict__ grad_out, const int* __restrict__ idx,
                                                 const float* __restrict__ weight, float4* __restrict__ grad_points) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned j = t / (unsigned)cv;
    if (j >= (unsigned)n) return;
    const unsigned l = t - j * (unsigned)cv;
    const size_t cloud = blockIdx.y;
    const size_t row = cloud * n + j;
    const float4 g = grad_out[row * cv + l];
    float4* G = grad_points + cloud * (size_t)m * cv + l;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const float w = weight[row * 3 + u];
        float4* dst = G + (size_t)idx[row * 3 + u] * cv;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__fmul_rn(g.x, w)), "f"(__fmul_rn(g.y, w)),
                     "f"(__fmul_rn(g.z, w)), "f"(__fmul_rn(g.w, w))
                     : "memory");
    }
}

// Atomic-free gradient.  Source entries are the 3n (j, u) pairs in row-major order; grad_points[t, :] accumulates
// grad_out[j, :] * w[j, u] over the entries that point at t in ascending (j, u) -- exactly the order of
// threeinterpolate_grad_cpu (tf_interpolate.cpp:131-153), unfused, so the result is bit-exact with the reference.
template <typename VEC>
__device__ __forceinline__ VEC vec_madd_unfused(VEC acc, VEC g, float w);
template <>
__device__ __forceinline__ float vec_madd_unfused<float>(float acc, float g, float w) { return __fadd_rn(acc, __fmul_rn(g, w)); }
template <>
__device__ __forceinline__ float4 vec_madd_unfused<float4>(float4 acc, float4 g, float w) {
    return make_float4(__fadd_rn(acc.x, __fmul_rn(g.x, w)), __fadd_rn(acc.y, __fmul_rn(g.y, w)), __fadd_rn(acc.z, __fmul_rn(g.z, w)),
                       __fadd_rn(acc.w, __fmul_rn(g.w, w)));
}
template <typename VEC>
__device__ __forceinline__ VEC vec_zero3();
template <>
__device__ __forceinline__ float vec_zero3<float>() { return 0.f; }
template <>
__device__ __forceinline__ float4 vec_zero3<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <typename VEC>
__global__ void three_interpolate_grad_seg_kernel(int n, int cv, int m, const VEC* __restrict__ grad_out, const float* __restrict__ weight,
                                                  const int* __restrict__ offset, const int* __restrict__ list, VEC* __restrict__ grad_points) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (known point, channel vector) of this cloud
    const unsigned i = t / (unsigned)cv;
    if (i >= (unsigned)m) return;
    const unsigned l = t - i * (unsigned)cv;
    const size_t cloud = blockIdx.y;
    const size_t R = (size_t)n * 3;
    const int beg = offset[cloud * (m + 1) + i], end = offset[cloud * (m + 1) + i + 1];
    const int* __restrict__ seg = list + cloud * R;
    const VEC* __restrict__ G = grad_out + cloud * (size_t)n * cv + l;
    const float* __restrict__ W = weight + cloud * R;
    VEC acc = vec_zero3<VEC>();
    // TIG_U contributions in flight, the entry numbers of the NEXT batch loaded before the current batch's rows are waited for (a
    // first version chained entry -> row loads four at a time: 13 dependent memory round trips for a segment of 24; ncu: long-scoreboard
    // bound at 30 % of DRAM).  Accumulated in ascending source order, as the sequential reference does.
    constexpr int TIG_U = 8;
    int cur[TIG_U];
#pragma unroll
    for (int u = 0; u < TIG_U; ++u) cur[u] = beg + u < end ? seg[beg + u] : -1;   // = 3*j + u
    for (int e = beg; e < end; e += TIG_U) {
        VEC g[TIG_U];
        float w[TIG_U];
#pragma unroll
        for (int u = 0; u < TIG_U; ++u)
            if (cur[u] >= 0) {
                g[u] = __ldg(G + (size_t)(cur[u] / 3) * cv);
                w[u] = __ldg(W + cur[u]);
            }
        int nxt[TIG_U];
#pragma unroll
        for (int u = 0; u < TIG_U; ++u) nxt[u] = e + TIG_U + u < end ? seg[e + TIG_U + u] : -1;
#pragma unroll
        for (int u = 0; u < TIG_U; ++u)
            if (cur[u] >= 0) acc = vec_madd_unfused<VEC>(acc, g[u], w[u]);
#pragma unroll
        for (int u = 0; u < TIG_U; ++u) cur[u] = nxt[u];
    }
    grad_points[(cloud * m + i) * cv + l] = acc;
}

}  // namespace rfnet

using namespace rfnet;

extern "C" size_t rfnet_three_nn_workspace_bytes(int b, int n, int m) {
    (void)n;
    if (b <= 0 || m < TG_MIN_POINTS || m > TG_MAX_POINTS) return 0;
    return ball_grid_stride(m) * (size_t)b;
}

extern "C" int rfnet_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx, void* workspace,
                              size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0 || n == 0) return 0;
    RFNET_CHECK_ARG(xyz1 && dist && idx && (m == 0 || xyz2) && b <= 65535);
    cudaStream_t s = (cudaStream_t)stream;
    if (workspace && m >= TG_MIN_POINTS && m <= TG_MAX_POINTS && workspace_bytes >= rfnet_three_nn_workspace_bytes(b, n, m)) {
        const size_t stride = ball_grid_stride(m);
        const size_t gsmem = sizeof(unsigned) * (BG_CELLS + BG_CELLS / 32);
        RFNET_CUDA(cudaFuncSetAttribute(ball_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        ball_grid_kernel<<<b, 1024, gsmem, s>>>(m, nullptr, TG_GMAX, xyz2, (unsigned char*)workspace, stride);
        const size_t smem = sizeof(int) * (TG_GMAX * TG_GMAX * TG_GMAX + 1 + (size_t)m) + sizeof(float) * 3 * (size_t)m;
        RFNET_CUDA(cudaFuncSetAttribute(three_nn_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)((n + TG_THREADS - 1) / TG_THREADS), (unsigned)b);
        three_nn_grid_kernel<<<grid, TG_THREADS, smem, s>>>(n, m, xyz1, (const unsigned char*)workspace, stride, dist, idx);
        return launch_status();
    }
    dim3 grid((unsigned)((n + TN_THREADS * TN_Q - 1) / (TN_THREADS * TN_Q)), (unsigned)b);
    three_nn_kernel<<<grid, TN_THREADS, 0, s>>>(n, m, xyz1, xyz2, dist, idx);
    return launch_status();
}

extern "C" int rfnet_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight, float* out,
                                       rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && c >= 0 && m >= 0);
    const size_t rows = (size_t)b * n;
    if (rows == 0 || c == 0) return 0;
    RFNET_CHECK_ARG(m > 0 && points && idx && weight && out);
    cudaStream_t s = (cudaStream_t)stream;
    RFNET_CHECK_ARG(b <= 65535 && (size_t)n * c < 0x7fffffffull);
    const bool aligned = (((uintptr_t)points | (uintptr_t)out) & 15u) == 0;
    const size_t slice_bytes = (size_t)m * TIS_CSV * sizeof(float4);
    if (aligned && c % (4 * TIS_CSV) == 0 && slice_bytes <= 200 * 1024) {
        // staged: worth it once a CTA streams several times more rows than it stages
        const int nslice = c / (4 * TIS_CSV);
        const int groups = num_sms() / nslice > 0 ? num_sms() / nslice : 1;
        const unsigned long long rpg = (rows + groups - 1) / groups;
        if (rpg >= 4ull * (unsigned)m) {
            RFNET_CUDA(cudaFuncSetAttribute(three_interpolate_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice_bytes));
            three_interpolate_staged_kernel<<<(unsigned)(groups * nslice), TIS_THREADS, slice_bytes, s>>>(b, m, c / 4, n, (const float4*)points, idx, weight,
                                                                                                     (float4*)out, rpg);
            return launch_status();
        }
    }
    if (c % 4 == 0 && aligned) {
        dim3 grid((unsigned)(((size_t)n * (c / 4) + 256 * TI_E - 1) / (256 * TI_E)), (unsigned)b);
        three_interpolate_kernel<float4><<<grid, 256, 0, s>>>(m, c / 4, n, (const float4*)points, idx, weight, (float4*)out);
    } else {
        dim3 grid((unsigned)(((size_t)n * c + 256 * TI_E - 1) / (256 * TI_E)), (unsigned)b);
        three_interpolate_kernel<float><<<grid, 256, 0, s>>>(m, c, n, points, idx, weight, out);
    }
    return launch_status();
}

extern "C" size_t rfnet_three_interpolate_grad_workspace_bytes(int b, int n, int c, int m) {
    (void)c;
    if (b <= 0 || m <= 0) return 0;
    return seg::csr_bytes(b, m, (size_t)(n > 0 ? n : 0) * 3);
}

static int three_interpolate_grad_from_csr(int b, int n, int c, int m, const float* grad_out, const float* weight, const seg::Csr& csr, float* grad_points,
                                           cudaStream_t s) {
    const bool vec = c % 4 == 0 && (((uintptr_t)grad_out | (uintptr_t)grad_points) & 15u) == 0;
    if (vec) {
        dim3 grid((unsigned)(((size_t)m * (c / 4) + 255) / 256), (unsigned)b);
        three_interpolate_grad_seg_kernel<float4><<<grid, 256, 0, s>>>(n, c / 4, m, (const float4*)grad_out, weight, csr.offset, csr.list, (float4*)grad_points);
    } else {
        dim3 grid((unsigned)(((size_t)m * c + 255) / 256), (unsigned)b);
        three_interpolate_grad_seg_kernel<float><<<grid, 256, 0, s>>>(n, c, m, grad_out, weight, csr.offset, csr.list, grad_points);
    }
    return launch_status();
}

// three_interpolate's gradient over a scatter plan of idx read as (b, 3n) targets in [0, m) (rfnet_scatter_plan_build)
extern "C" int rfnet_three_interpolate_grad_planned(int b, int n, int c, int m, const float* grad_out, const float* weight, const void* plan,
                                                    size_t plan_bytes, float* grad_points, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && c >= 0 && m >= 0);
    if ((size_t)b * m * c == 0) return 0;
    RFNET_CHECK_ARG(grad_points && plan && (n == 0 || (grad_out && weight)));
    RFNET_CHECK_ARG(b <= 65535 && (size_t)n * c < 0x7fffffffull && (size_t)m * c < 0x7fffffffull && (size_t)n * 3 < 0x7fffffffull);
    RFNET_CHECK_ARG(plan_bytes >= seg::csr_bytes(b, m, (size_t)n * 3));
    const seg::Csr csr = seg::csr_carve(const_cast<void*>(plan), b, m, (size_t)n * 3);
    return three_interpolate_grad_from_csr(b, n, c, m, grad_out, weight, csr, grad_points, (cudaStream_t)stream);
}

extern "C" int rfnet_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight,
                                            float* grad_points, void* workspace, size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && c >= 0 && m >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    if ((size_t)b * m * c == 0) return 0;
    const size_t rows = (size_t)b * n;
    RFNET_CHECK_ARG(grad_points && (rows == 0 || (grad_out && idx && weight)));
    RFNET_CHECK_ARG(b <= 65535 && (size_t)n * c < 0x7fffffffull && (size_t)m * c < 0x7fffffffull && (size_t)n * 3 < 0x7fffffffull);
    const bool vec = c % 4 == 0 && (((uintptr_t)grad_out | (uintptr_t)grad_points) & 15u) == 0;
    if (workspace) {
        RFNET_CHECK_ARG(workspace_bytes >= rfnet_three_interpolate_grad_workspace_bytes(b, n, c, m));
        seg::Csr csr = seg::csr_carve(workspace, b, m, (size_t)n * 3);
        const int rc = seg::csr_build(csr, b, m, (size_t)n * 3, idx, s);   // idx (b, n, 3) read as (b, 3n) source entries
        if (rc) return rc;
        return three_interpolate_grad_from_csr(b, n, c, m, grad_out, weight, csr, grad_points, s);
    }
    RFNET_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, s));
    if (rows == 0) return 0;
    if (vec) {
        dim3 grid((unsigned)(((size_t)n * (c / 4) + 255) / 256), (unsigned)b);
        three_interpolate_grad_v4_kernel<<<grid, 256, 0, s>>>(n, c / 4, m, (const float4*)grad_out, idx, weight, (float4*)grad_points);
    } else {
        dim3 grid((unsigned)(((size_t)n * c + 255) / 256), (unsigned)b);
        three_interpolate_grad_kernel<<<grid, 256, 0, s>>>(n, c, m, grad_out, idx, weight, grad_points);
    }
    return launch_status();
}

extern "C" int rfnet_knn_point(int b, int n, int m, int k, const float* xyz1, const float* xyz2, float* val, int* idx, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && k > 0 && k <= 32 && k <= (n > 0 ? n : 1));
    if (b == 0 || m == 0) return 0;
    RFNET_CHECK_ARG(xyz1 && xyz2 && val && idx && n > 0 && b <= 65535);
    dim3 grid((unsigned)((m + KNN_THREADS - 1) / KNN_THREADS), (unsigned)b);
    cudaStream_t s = (cudaStream_t)stream;
    if (k <= 1) knn_kernel<1><<<grid, KNN_THREADS, 0, s>>>(n, m, k, xyz1, xyz2, val, idx);
    else if (k <= 2) knn_kernel<2><<<grid, KNN_THREADS, 0, s>>>(n, m, k, xyz1, xyz2, val, idx);
    else if (k <= 4) knn_kernel<4><<<grid, KNN_THREADS, 0, s>>>(n, m, k, xyz1, xyz2, val, idx);
    else if (k <= 8) knn_kernel<8><<<grid, KNN_THREADS, 0, s>>>(n, m, k, xyz1, xyz2, val, idx);
    else if (k <= 16) knn_kernel<16><<<grid, KNN_THREADS, 0, s>>>(n, m, k, xyz1, xyz2, val, idx);
    else knn_kernel<32><<<grid, KNN_THREADS, 0, s>>>(n, m, k, xyz1, xyz2, val, idx);
    return launch_status();
}
