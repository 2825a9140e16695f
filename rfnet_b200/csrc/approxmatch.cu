// approx_match (approximate EMD transport plan), match_cost and match_cost gradient for sm_100a.
//
// Replaces approxmatch / matchcost / matchcostgrad1 / matchcostgrad2 (pc_distance/tf_approxmatch.cu:1-295).
//
// The reference runs ONE 512-thread block per cloud through 10 levels x 3 dependent n*m passes and read-modify-writes the
// n*m match matrix once per level (80 B of HBM traffic per pair).  Here:
//   * every pass is the same reduction  S[row] = sum_c exp(level * d2(row, c)) * w[c]  ("weighted exp-sum sweep") with the
//     roles of the two clouds swapped between passes; it runs on the whole chip: grid = clouds x row tiles x candidate
//     splits, rows in registers as packed pairs (FADD2/FMUL2/FFMA2), candidates + weights broadcast from shared memory,
//     one MUFU.EX2 per pair; the pass's update rule (ratioL / consumption+ratioR / remainL) is applied by the sweep itself
//     when a CTA sees all candidates, else partial sums go to a small buffer and a tiny epilogue kernel reduces them in
//     split order -- deterministic, no atomics;
//   * match is NOT accumulated level by level.  The per-level factors ratioL_j[k], ratioR_j[l] are kept (10*(n+m) floats per
//     cloud) and match[l,k] = sum_j e_j(k,l) * ratioL_j[k] * ratioR_j[l] is written ONCE at the end: 4 B/pair of HBM
//     traffic instead of 80, at the price of 9 more ex2 per pair (the j = -2 level has e = 1).
// exp(x) is ex2.approx(x * log2e) exactly as __expf in the reference; level * log2e is folded into one constant, which is
// bit-identical because level is a power of two.  d2 is the reference's fused expression.  Offsets are 64-bit.
#include <stdlib.h>

#include "common.cuh"
#include "morton.cuh"
#include "rfnet_ops.h"

namespace rfnet {

constexpr int EMD_LEVELS = 10;   // j = 7 .. -2   (tf_approxmatch.cu:21)
constexpr int EMD_THREADS = 128;
#ifndef EMD_TC_VALUE
#define EMD_TC_VALUE 256
#endif
constexpr int EMD_TC = EMD_TC_VALUE;  // candidates per shared-memory chunk (float4 each: 4 KiB)
constexpr float LOG2E = 1.4426950408889634f;
#ifndef EMD_UNROLL
#define EMD_UNROLL 2  // candidates per unrolled step of the sweep (tools/emd_tune.cu sweeps this)
#endif
#define EMD_PRAGMA_(x) _Pragma(#x)
#define EMD_PRAGMA_UNROLL(n) EMD_PRAGMA_(unroll n)

__host__ __device__ inline float emd_level(int li) {  // li = 0..9  ->  j = 7..-2 ; level = -4^j, 0 at j = -2
    const int j = 7 - li;
    if (j == -2) return 0.0f;
    float v = 1.0f;
    for (int i = 0; i < (j < 0 ? -j : j); ++i) v *= 4.0f;
    return j >= 0 ? -v : -1.0f / v;
}

// Workspace layout (floats), per call:
//   remainL [b*n] remainR [b*m] ratioL [b*n] ratioR [b*m]                      running state of the current level
//   facL [EMD_LEVELS][b*n]  facR [EMD_LEVELS][b*m]                             per-level factors for the final materialisation
//   partial [b * max(n,m) * max_split]                                         sweep partial sums
struct EmdWs {
    float *remainL, *remainR, *ratioL, *ratioR, *facL, *facR, *partial;
    int *perm1, *perm2;          // Morton order of xyz1 / xyz2 (pruned sweeps only)
    unsigned *maskA, *maskB;     // candidate masks: rows = xyz1 clusters vs xyz2 candidates (passes 1, 3) / the reverse (pass 2)
};
// pruned sweeps pay off (and their sort fits shared memory) for clouds of 4096 .. 32768 points
static bool emd_prune_enabled(int n, int m) { return n >= 4096 && m >= 4096 && n <= 32768 && m <= 32768; }
static size_t emd_mask_words(int b, int nr, int nc) { return (size_t)b * ((nr + 127) / 128) * 3 * ((nc + 31) / 32); }
static int emd_max_split(int b, int n, int m) {
    // enough CTAs per sweep to give every SM ~8 at the smallest batch; bounded by the number of candidate chunks
    const int rows = n < m ? n : m, cands = n < m ? m : n;
    (void)rows;
    int chunks = (cands + EMD_TC - 1) / EMD_TC;
    int s = chunks < 32 ? chunks : 32;
    return s < 1 ? 1 : s;
}
static size_t emd_ws_floats(int b, int n, int m) {
    const size_t bn = (size_t)b * n, bm = (size_t)b * m;
    size_t f = 2 * (bn + bm) + (size_t)EMD_LEVELS * (bn + bm) + (size_t)b * (n > m ? n : m) * emd_max_split(b, n, m);
    if (emd_prune_enabled(n, m)) f += bn + bm + emd_mask_words(b, n, m) + emd_mask_words(b, m, n);
    return f;
}
static EmdWs emd_carve(float* w, int b, int n, int m) {
    const size_t bn = (size_t)b * n, bm = (size_t)b * m;
    EmdWs s;
    s.remainL = w; w += bn;
    s.remainR = w; w += bm;
    s.ratioL = w; w += bn;
    s.ratioR = w; w += bm;
    s.facL = w; w += EMD_LEVELS * bn;
    s.facR = w; w += EMD_LEVELS * bm;
    s.partial = w; w += (size_t)b * (n > m ? n : m) * emd_max_split(b, n, m);
    s.perm1 = reinterpret_cast<int*>(w); w += bn;
    s.perm2 = reinterpret_cast<int*>(w); w += bm;
    s.maskA = reinterpret_cast<unsigned*>(w); w += emd_mask_words(b, n, m);
    s.maskB = reinterpret_cast<unsigned*>(w);
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// Weighted exp-sum sweep:  partial[cloud][split][row] = init + sum_{c in split, ascending} ex2(lvl2 * d2(row, c)) * w[c]
//   rows: (b, nr, 3), cands: (b, nc, 3), w: (b, nc).  grid.x = b * nrt * nsplit.  Q rows per thread (Q/2 packed pairs).
// The accumulation is the reference binary's, term by term and in candidate order:
//   passes 1 and 2:  acc = fma(e, w[c], acc)                       (acc starts at 1e-9 in pass 1: tf_approxmatch.cu:36)
//   pass 3 (PASS3):  acc = fma(rowfac[row] * e, w[c], acc)         (t = ratioL*e; suml = fma(t, ratioR, suml))
// so with nsplit == 1 every row sum is bit-identical to what the reference's thread computes.
// UNIT_E: the last level (j = -2) has level = 0, i.e. e = ex2(0 * d2) = 1 for every pair; the same accumulation chain is run
// without evaluating distances or exponentials (1 lane-op per pair instead of 8 + a MUFU).
// ---------------------------------------------------------------------------------------------------------------
// ---- the pass's update rule ("normalise" in pass 1, "saturate" in pass 2, "consume" in pass 3) ------------------
// EPI = 1: ratioL[k] = remainL[k] / (1e-9 + sum)                                  tf_approxmatch.cu:26-59 (the seed is in the sum)
// EPI = 2: consumption = min(remainR / (sumr + 1e-9), 1); ratioR = consumption * remainR; remainR -= sumr   :75-108
// EPI = 3: remainL[k] = max(0, remainL[k] - sum)                                  :127-160 (ratioL is applied per term in the sweep)
// When the candidates are NOT split over several CTAs a sweep thread owns the complete sum of its rows and applies the rule
// itself (fused: no partial-sum round trip, no epilogue launch); with splits the partial sums are reduced in split order by
// the small epilogue kernels further down.  (Fusing the split case too, through a last-CTA ticket, was measured slower.)
struct EmdEpi {
    float* remain;   // EPI 1, 3: remainL   EPI 2: remainR
    float* ratio;    // EPI 1: ratioL       EPI 2: ratioR
    float* fac;      // EPI 1: facL[level]  EPI 2: facR[level]
};
template <int EPI>
__device__ __forceinline__ void emd_apply(const EmdEpi& e, size_t idx, float sum) {
    if (EPI == 1) {
        const float r = e.remain[idx] / sum;
        e.ratio[idx] = r;
        e.fac[idx] = r;
    } else if (EPI == 2) {
        const float rem = e.remain[idx];
        const float sumr = sum * rem;
        const float consumption = fminf(rem / (sumr + 1e-9f), 1.0f);
        const float r = consumption * rem;
        e.ratio[idx] = r;
        e.fac[idx] = r;
        e.remain[idx] = fmaxf(0.0f, rem - sumr);
    } else {
        e.remain[idx] = fmaxf(0.0f, e.remain[idx] - sum);
    }
}

template <int Q, int EPI, bool UNIT_E>
__global__ void __launch_bounds__(EMD_THREADS) emd_sweep_kernel(int nr, int nc, int nrt, int nsplit, int cps, float lvl2, float init0,
                                                               const float* __restrict__ rows, const float* __restrict__ cands,
                                                               const float* __restrict__ w, const float* __restrict__ rowfac,
                                                               float* __restrict__ partial, EmdEpi epi) {
    constexpr bool PASS3 = EPI == 3;
    __shared__ __align__(16) float4 sC[EMD_TC];
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int split = bid % nsplit;
    const int tile = (bid / nsplit) % nrt;
    const int cloud = bid / (nsplit * nrt);
    const float* __restrict__ rbase = rows + (size_t)cloud * nr * 3;
    const float* __restrict__ cbase = cands + (size_t)cloud * nc * 3;
    const float* __restrict__ wbase = w + (size_t)cloud * nc;

    const int r0 = tile * (EMD_THREADS * Q) + tid;
    float2 rx[Q / 2], ry[Q / 2], rz[Q / 2], rf[Q / 2], acc[Q / 2];
    const float a0 = split == 0 ? init0 : 0.0f;
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int ia = r0 + (2 * h) * EMD_THREADS, ib = ia + EMD_THREADS;
        const bool va = ia < nr, vb = ib < nr;
        // rows are kept NEGATED so that (cand - row) is one FADD2 with a broadcast scalar: c + (-r) == c - r exactly
        rx[h].x = va ? -rbase[(size_t)ia * 3 + 0] : 0.f; ry[h].x = va ? -rbase[(size_t)ia * 3 + 1] : 0.f; rz[h].x = va ? -rbase[(size_t)ia * 3 + 2] : 0.f;
        rx[h].y = vb ? -rbase[(size_t)ib * 3 + 0] : 0.f; ry[h].y = vb ? -rbase[(size_t)ib * 3 + 1] : 0.f; rz[h].y = vb ? -rbase[(size_t)ib * 3 + 2] : 0.f;
        rf[h] = make_float2(1.f, 1.f);
        if (PASS3) {
            rf[h].x = va ? rowfac[(size_t)cloud * nr + ia] : 0.f;
            rf[h].y = vb ? rowfac[(size_t)cloud * nr + ib] : 0.f;
        }
        acc[h] = make_float2(a0, a0);
    }
    const float2 L2 = make_float2(lvl2, lvl2);

    const int c_begin = split * cps * EMD_TC;
    const int c_end = min(nc, c_begin + cps * EMD_TC);
    for (int c0 = c_begin; c0 < c_end; c0 += EMD_TC) {
        const int len = min(EMD_TC, c_end - c0);
        __syncthreads();
        for (int i = tid; i < EMD_TC; i += EMD_THREADS) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // padded candidates carry weight 0: fma(e, 0, acc) == acc
            if (i < len) {
                const float* c = cbase + (size_t)(c0 + i) * 3;
                v = make_float4(c[0], c[1], c[2], wbase[c0 + i]);
            }
            sC[i] = v;
        }
        __syncthreads();
        const int len2 = (len + EMD_UNROLL - 1) / EMD_UNROLL * EMD_UNROLL;  // padded entries carry weight 0 (EMD_TC % EMD_UNROLL == 0)
        EMD_PRAGMA_UNROLL(EMD_UNROLL)
        for (int k = 0; k < len2; ++k) {
            const float4 c = sC[k];
#pragma unroll
            for (int h = 0; h < Q / 2; ++h) {
                if (UNIT_E) {
                    // e == 1: fma(1, w, acc) == acc + w and fma(rf * 1, w, acc) == fma(rf, w, acc), bit for bit
                    acc[h] = PASS3 ? __ffma2_rn(rf[h], make_float2(c.w, c.w), acc[h]) : __fadd2_rn(acc[h], make_float2(c.w, c.w));
                } else {
                    const float2 dx = __fadd2_rn(rx[h], make_float2(c.x, c.x));  // the sign of the difference is irrelevant after squaring
                    const float2 dy = __fadd2_rn(ry[h], make_float2(c.y, c.y));
                    const float2 dz = __fadd2_rn(rz[h], make_float2(c.z, c.z));
                    const float2 a = __fmul2_rn(sqdist3x2<true>(dx, dy, dz), L2);
                    float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
                    if (PASS3) e = __fmul2_rn(rf[h], e);
                    acc[h] = __ffma2_rn(e, make_float2(c.w, c.w), acc[h]);
                }
            }
        }
    }
    if (nsplit == 1) {   // complete sums: apply the pass's rule here
#pragma unroll
        for (int h = 0; h < Q / 2; ++h) {
            const int ia = r0 + (2 * h) * EMD_THREADS, ib = ia + EMD_THREADS;
            if (ia < nr) emd_apply<EPI>(epi, (size_t)cloud * nr + ia, acc[h].x);
            if (ib < nr) emd_apply<EPI>(epi, (size_t)cloud * nr + ib, acc[h].y);
        }
        return;
    }
    float* __restrict__ out = partial + ((size_t)cloud * nsplit + split) * nr;
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int ia = r0 + (2 * h) * EMD_THREADS, ib = ia + EMD_THREADS;
        if (ia < nr) out[ia] = acc[h].x;
        if (ib < nr) out[ib] = acc[h].y;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Exact pruning of the three sharpest levels (j = 7, 6, 5: e = exp(-4^j d2) with 4^j = 16384, 4096, 1024).
// ex2.approx.ftz returns EXACTLY 0 once its argument is below -126, i.e. for d2 > 0.0053 / 0.021 / 0.085, and a zero term
// leaves the accumulator bit-identical (fma(0, w, acc) == acc).  At those levels almost every pair is such a no-op, so:
//   * morton_sort_kernel orders each cloud along a Morton curve (one CTA per cloud, counting sort in shared memory);
//     a warp of the pruned sweep then owns 128 CONSECUTIVE points of that order: a spatially tight cluster;
//   * emd_mask_kernel marks, per cluster and level, the candidates whose distance to the cluster's bounding box still
//     allows a non-zero term (with a safety margin of 4 in the exponent-2 argument, ~3 % in distance);
//   * emd_sweep_pruned_kernel is the sweep above restricted to marked candidates, visited in ASCENDING candidate order:
//     every row sum goes through the same sequence of non-trivial fma's as in the dense sweep, so the result is bit-for-bit
//     the dense one (tests/test_emd_gpu.py::test_pruned_sweeps_are_exact) -- only the row -> thread assignment changes.
// The masks depend on the points only: built once per call for both roles (rows = xyz1 / rows = xyz2), used by 9 sweeps.
// ---------------------------------------------------------------------------------------------------------------
constexpr int EMD_PRUNE_LEVELS = 3;
constexpr int EMD_CLUSTER = 128;        // rows of one warp of the pruned sweep (4 per lane)
constexpr float EMD_PRUNE_ARG = -130.0f;

// grid = (clusters of rows, clouds), 256 threads.  mask[((cloud * nclusters + cluster) * 3 + lev) * nwords + word]
__global__ void __launch_bounds__(256) emd_mask_kernel(int nr, int nc, int nwords, const float* __restrict__ rows, const int* __restrict__ perm,
                                                       const float* __restrict__ cands, float l0, float l1, float l2, unsigned* __restrict__ mask) {
    __shared__ float red[6][8];
    const int cloud = blockIdx.y, cluster = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    const int s = cluster * EMD_CLUSTER + tid;
    if (tid < EMD_CLUSTER && s < nr) {
        const float* p = rows + ((size_t)cloud * nr + perm[(size_t)cloud * nr + s]) * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) lo[a] = hi[a] = p[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = red[a][0], h = red[3 + a][0];
        for (int w2 = 1; w2 < 8; ++w2) { l = fminf(l, red[a][w2]); h = fmaxf(h, red[3 + a][w2]); }
        lo[a] = l;
        hi[a] = h;
    }
    unsigned* __restrict__ out = mask + ((size_t)cloud * gridDim.x + cluster) * EMD_PRUNE_LEVELS * nwords;
    const float* __restrict__ cb = cands + (size_t)cloud * nc * 3;
    for (int c = tid; c < nwords * 32; c += 256) {
        float d2 = inf;
        if (c < nc) {
            d2 = 0.f;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = cb[(size_t)c * 3 + a];
                const float d = fmaxf(fmaxf(lo[a] - v, v - hi[a]), 0.f);   // distance to the box along this axis
                d2 = fmaf(d, d, d2);
            }
        }
        const bool valid = c < nc;
        const unsigned w0 = __ballot_sync(0xffffffffu, valid && l0 * d2 > EMD_PRUNE_ARG);
        const unsigned w1 = __ballot_sync(0xffffffffu, valid && l1 * d2 > EMD_PRUNE_ARG);
        const unsigned w2 = __ballot_sync(0xffffffffu, valid && l2 * d2 > EMD_PRUNE_ARG);
        if (lane == 0) {
            out[c >> 5] = w0;
            out[nwords + (c >> 5)] = w1;
            out[2 * nwords + (c >> 5)] = w2;
        }
    }
}

// The dense sweep (Q = 4) restricted to the marked candidates of the warp's cluster.  `mask` points at the level's words
// of cluster 0 of cloud 0; clusters are EMD_PRUNE_LEVELS * nwords apart.
template <int EPI>
__global__ void __launch_bounds__(EMD_THREADS) emd_sweep_pruned_kernel(int nr, int nc, int nrt, int nsplit, int cps, float lvl2, float init0,
                                                                      const float* __restrict__ rows, const float* __restrict__ cands,
                                                                      const float* __restrict__ w, const float* __restrict__ rowfac,
                                                                      const int* __restrict__ perm, const unsigned* __restrict__ mask, int nwords,
                                                                      float* __restrict__ partial, EmdEpi epi) {
    constexpr int Q = 4;
    constexpr bool PASS3 = EPI == 3;
    __shared__ __align__(16) float4 sC[EMD_TC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bid = blockIdx.x;
    const int split = bid % nsplit;
    const int tile = (bid / nsplit) % nrt;
    const int cloud = bid / (nsplit * nrt);
    const float* __restrict__ rbase = rows + (size_t)cloud * nr * 3;
    const float* __restrict__ cbase = cands + (size_t)cloud * nc * 3;
    const float* __restrict__ wbase = w + (size_t)cloud * nc;
    const int* __restrict__ pbase = perm + (size_t)cloud * nr;
    const int nclusters = (nr + EMD_CLUSTER - 1) / EMD_CLUSTER;
    const int cluster = tile * (EMD_THREADS * Q / EMD_CLUSTER) + warp;
    const unsigned* __restrict__ mrow = mask + ((size_t)cloud * nclusters + min(cluster, nclusters - 1)) * EMD_PRUNE_LEVELS * nwords;

    // sorted position of this lane's rows: the warp's cluster is 128 consecutive positions, 32 per q
    const int s0 = cluster * EMD_CLUSTER + lane;
    int row[Q];
    float2 rx[Q / 2], ry[Q / 2], rz[Q / 2], rf[Q / 2], acc[Q / 2];
    const float a0 = split == 0 ? init0 : 0.0f;
#pragma unroll
    for (int q = 0; q < Q; ++q) row[q] = (s0 + 32 * q) < nr ? pbase[s0 + 32 * q] : -1;
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int ia = row[2 * h], ib = row[2 * h + 1];
        const bool va = ia >= 0, vb = ib >= 0;
        rx[h].x = va ? -rbase[(size_t)ia * 3 + 0] : 0.f; ry[h].x = va ? -rbase[(size_t)ia * 3 + 1] : 0.f; rz[h].x = va ? -rbase[(size_t)ia * 3 + 2] : 0.f;
        rx[h].y = vb ? -rbase[(size_t)ib * 3 + 0] : 0.f; ry[h].y = vb ? -rbase[(size_t)ib * 3 + 1] : 0.f; rz[h].y = vb ? -rbase[(size_t)ib * 3 + 2] : 0.f;
        rf[h] = make_float2(1.f, 1.f);
        if (PASS3) {
            rf[h].x = va ? rowfac[(size_t)cloud * nr + ia] : 0.f;
            rf[h].y = vb ? rowfac[(size_t)cloud * nr + ib] : 0.f;
        }
        acc[h] = make_float2(a0, a0);
    }
    const float2 L2 = make_float2(lvl2, lvl2);
    const int c_begin = split * cps * EMD_TC;
    const int c_end = min(nc, c_begin + cps * EMD_TC);
    for (int c0 = c_begin; c0 < c_end; c0 += EMD_TC) {
        const int len = min(EMD_TC, c_end - c0);
        __syncthreads();
        for (int i = tid; i < len; i += EMD_THREADS) {
            const float* c = cbase + (size_t)(c0 + i) * 3;
            sC[i] = make_float4(c[0], c[1], c[2], wbase[c0 + i]);
        }
        // this warp's mask words for the chunk (c0 is a multiple of 32): lane i holds word i
        const int wi0 = c0 >> 5;
        const unsigned myword = (lane < EMD_TC / 32 && wi0 + lane < nwords) ? mrow[wi0 + lane] : 0u;
        __syncthreads();
#pragma unroll 1
        for (int wi = 0; wi < EMD_TC / 32; ++wi) {
            unsigned bits = __shfl_sync(0xffffffffu, myword, wi);
            while (bits) {
                const int k = wi * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const float4 c = sC[k];
#pragma unroll
                for (int h = 0; h < Q / 2; ++h) {
                    const float2 dx = __fadd2_rn(rx[h], make_float2(c.x, c.x));
                    const float2 dy = __fadd2_rn(ry[h], make_float2(c.y, c.y));
                    const float2 dz = __fadd2_rn(rz[h], make_float2(c.z, c.z));
                    const float2 a = __fmul2_rn(sqdist3x2<true>(dx, dy, dz), L2);
                    float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
                    if (PASS3) e = __fmul2_rn(rf[h], e);
                    acc[h] = __ffma2_rn(e, make_float2(c.w, c.w), acc[h]);
                }
            }
        }
    }
    if (nsplit == 1) {
#pragma unroll
        for (int h = 0; h < Q / 2; ++h) {
            if (row[2 * h] >= 0) emd_apply<EPI>(epi, (size_t)cloud * nr + row[2 * h], acc[h].x);
            if (row[2 * h + 1] >= 0) emd_apply<EPI>(epi, (size_t)cloud * nr + row[2 * h + 1], acc[h].y);
        }
        return;
    }
    float* __restrict__ out = partial + ((size_t)cloud * nsplit + split) * nr;
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        if (row[2 * h] >= 0) out[row[2 * h]] = acc[h].x;
        if (row[2 * h + 1] >= 0) out[row[2 * h + 1]] = acc[h].y;
    }
}

// ---- epilogues: one thread per row; sum the split partials in fixed order, then the pass's update rule -----------
__device__ __forceinline__ float emd_sum_partials(const float* __restrict__ partial, size_t cloud, int nsplit, int nr, int r) {
    float s = partial[(cloud * nsplit) * nr + r];
    for (int sp = 1; sp < nsplit; ++sp) s += partial[(cloud * nsplit + sp) * nr + r];
    return s;
}
__global__ void emd_init_kernel(size_t bn, size_t bm, float multiL, float multiR, float* __restrict__ remainL, float* __restrict__ remainR) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < bn) remainL[t] = multiL;                     // tf_approxmatch.cu:17-20
    else if (t < bn + bm) remainR[t - bn] = multiR;
}
// pass 1 (tf_approxmatch.cu:26-59): ratioL[k] = remainL[k] / (1e-9 + sum)
__global__ void emd_epi1_kernel(int n, int nsplit, size_t bn, const float* __restrict__ partial, const float* __restrict__ remainL,
                                float* __restrict__ ratioL, float* __restrict__ facL) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bn) return;
    const float suml = emd_sum_partials(partial, t / n, nsplit, n, (int)(t % n));  // the 1e-9 seed is inside split 0
    const float r = remainL[t] / suml;
    ratioL[t] = r;
    facL[t] = r;
}
// pass 2 (tf_approxmatch.cu:75-108)
__global__ void emd_epi2_kernel(int m, int nsplit, size_t bm, const float* __restrict__ partial, float* __restrict__ remainR,
                                float* __restrict__ ratioR, float* __restrict__ facR) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bm) return;
    const float rem = remainR[t];
    const float sumr = emd_sum_partials(partial, t / m, nsplit, m, (int)(t % m)) * rem;
    const float consumption = fminf(rem / (sumr + 1e-9f), 1.0f);
    const float r = consumption * rem;
    ratioR[t] = r;
    facR[t] = r;
    remainR[t] = fmaxf(0.0f, rem - sumr);
}
// pass 3 (tf_approxmatch.cu:127-160), without the match write: remainL[k] = max(0, remainL[k] - ratioL[k] * sum_l e*ratioR[l])
__global__ void emd_epi3_kernel(int n, int nsplit, size_t bn, const float* __restrict__ partial, float* __restrict__ remainL) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bn) return;
    const float suml = emd_sum_partials(partial, t / n, nsplit, n, (int)(t % n));  // ratioL is applied per term in the sweep
    remainL[t] = fmaxf(0.0f, remainL[t] - suml);
}

// ---------------------------------------------------------------------------------------------------------------
// Final materialisation: match[cloud, l, k] = sum_j ex2(lvl2_j * d2(k,l)) * facL_j[k] * facR_j[l]   (k contiguous)
// CTA = 128 k's x MT_L l's; thread owns one k (point + 10 factors in registers), l's come from shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MT_THREADS = 128;
constexpr int MT_L = 64;
struct EmdLevels { float lvl2[EMD_LEVELS]; };
// CTA = 256 k's (two per thread, a packed pair) x MT_L l's.  Per l the thread reads the point and its 10 factors with four
// LDS.128 (shared by both k's) and runs the level loop on packed pairs: FMUL2 (argument), 2 MUFU.EX2, FMUL2 (x facL pair),
// FFMA2 (x facR broadcast, accumulate).  The accumulation is the reference's: acc = fma(facL * e, facR, acc), level by level.
// WRITE stores the matrix; COST also folds match_cost (tf_approxmatch.cu:183-225) into the same pass -- cost partial per CTA
// = sum sqrt(d2) * match over its tile -- so a loss that only needs the cost (earth_mover, vv_recon.py:396-399) never
// writes or re-reads the (b, m, n) matrix.
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool WRITE, bool COST>
__global__ void __launch_bounds__(MT_THREADS) emd_materialise_kernel(int n, int m, size_t bn, size_t bm, EmdLevels lv,
                                                                     const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                                     const float* __restrict__ facL, const float* __restrict__ facR,
                                                                     float* __restrict__ match, float* __restrict__ cost_partial) {
    __shared__ __align__(16) float4 sP[MT_L];       // x, y, z of xyz2[l]
    __shared__ __align__(16) float sF[MT_L][12];    // facR_j[l], j = 0..9 (+2 pad)
    __shared__ float sW[MT_THREADS / 32];
    const int cloud = blockIdx.z;
    const int ka = blockIdx.x * (2 * MT_THREADS) + threadIdx.x, kb = ka + MT_THREADS;
    const int l0 = blockIdx.y * MT_L;
    const int nl = min(MT_L, m - l0);
    for (int i = threadIdx.x; i < nl; i += MT_THREADS) {
        const float* q = xyz2 + ((size_t)cloud * m + l0 + i) * 3;
        sP[i] = make_float4(q[0], q[1], q[2], 0.f);
    }
    for (int i = threadIdx.x; i < nl * EMD_LEVELS; i += MT_THREADS) {
        const int l = i / EMD_LEVELS, j = i % EMD_LEVELS;
        sF[l][j] = facR[(size_t)j * bm + (size_t)cloud * m + l0 + l];
    }
    const bool va = ka < n, vb = kb < n;
    // rows negated (cand + (-row) == cand - row exactly); lanes past n sit at infinity so they never block a level skip
    const float inf = __int_as_float(0x7f800000);
    float2 nx = make_float2(-inf, -inf), ny = make_float2(0.f, 0.f), nz = make_float2(0.f, 0.f);
    float2 fl[EMD_LEVELS];
    if (va) { const float* p = xyz1 + ((size_t)cloud * n + ka) * 3; nx.x = -p[0]; ny.x = -p[1]; nz.x = -p[2]; }
    if (vb) { const float* p = xyz1 + ((size_t)cloud * n + kb) * 3; nx.y = -p[0]; ny.y = -p[1]; nz.y = -p[2]; }
#pragma unroll
    for (int j = 0; j < EMD_LEVELS; ++j) {
        fl[j].x = va ? facL[(size_t)j * bn + (size_t)cloud * n + ka] : 0.f;
        fl[j].y = vb ? facL[(size_t)j * bn + (size_t)cloud * n + kb] : 0.f;
    }
    __syncthreads();
    float* __restrict__ out = WRITE ? match + ((size_t)cloud * m + l0) * n : nullptr;
    float2 csum = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int l = 0; l < nl; ++l) {
        const float4 q = sP[l];
        const float4 f0 = *reinterpret_cast<const float4*>(&sF[l][0]), f1 = *reinterpret_cast<const float4*>(&sF[l][4]), f2 = *reinterpret_cast<const float4*>(&sF[l][8]);
        const float fr[EMD_LEVELS] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y};
        const float2 d2 = sqdist3x2<true>(__fadd2_rn(nx, make_float2(q.x, q.x)), __fadd2_rn(ny, make_float2(q.y, q.y)), __fadd2_rn(nz, make_float2(q.z, q.z)));
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < EMD_LEVELS - 1; ++j) {
            // ex2.approx.ftz returns exactly 0 below 2^-126: when that holds for the whole warp the level contributes
            // fma(0, ., acc) == acc and its MUFUs can be skipped (top levels: most pairs are farther than 0.07 / 0.15 apart)
            const float2 a = __fmul2_rn(d2, make_float2(lv.lvl2[j], lv.lvl2[j]));
            if (j < 3 && !__any_sync(0xffffffffu, a.x >= -126.0f || a.y >= -126.0f)) continue;
            const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
            acc = __ffma2_rn(__fmul2_rn(fl[j], e), make_float2(fr[j], fr[j]), acc);
        }
        acc = __ffma2_rn(fl[EMD_LEVELS - 1], make_float2(fr[EMD_LEVELS - 1], fr[EMD_LEVELS - 1]), acc);  // j = -2: level 0, e = 1
        if (WRITE) {
            if (va) out[(size_t)l * n + ka] = acc.x;
            if (vb) out[(size_t)l * n + kb] = acc.y;
        }
        // lanes past n carry d2 = inf and acc = 0: keep inf * 0 out of the sum
        if (COST) csum = __ffma2_rn(make_float2(va ? sqrt_approx(d2.x) : 0.f, vb ? sqrt_approx(d2.y) : 0.f), acc, csum);
    }
    if (COST) {
        float sum = warp_sum(csum.x + csum.y);
        if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < MT_THREADS / 32; ++i) t += sW[i];
            cost_partial[((size_t)cloud * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// match_cost: cost[i] = sum_{l,k} sqrtf(d2(k,l)) * match[i,l,k]                     (tf_approxmatch.cu:183-225)
// One streaming pass over match.  CTA = 256 k's x MC_L l's -> one partial per CTA, then a fixed-order final sum.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MC_THREADS = 256;
constexpr int MC_L = 128;
__global__ void __launch_bounds__(MC_THREADS) matchcost_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                               const float* __restrict__ match, float* __restrict__ partial) {
    __shared__ float sP[MC_L * 3];
    __shared__ float sW[MC_THREADS / 32];
    const int cloud = blockIdx.z;
    const int k = blockIdx.x * MC_THREADS + threadIdx.x;
    const int l0 = blockIdx.y * MC_L;
    const int nl = min(MC_L, m - l0);
    for (int i = threadIdx.x; i < nl * 3; i += MC_THREADS) sP[i] = xyz2[((size_t)cloud * m + l0) * 3 + i];
    __syncthreads();
    float sum = 0.f;
    if (k < n) {
        const float* p = xyz1 + ((size_t)cloud * n + k) * 3;
        const float x1 = p[0], y1 = p[1], z1 = p[2];
        const float* __restrict__ mp = match + ((size_t)cloud * m + l0) * n + k;
#pragma unroll 4
        for (int l = 0; l < nl; ++l) {
            const float d2 = sqdist3<true>(sP[l * 3] - x1, sP[l * 3 + 1] - y1, sP[l * 3 + 2] - z1);
            sum = __fmaf_rn(__fsqrt_rn(d2), __ldg(mp + (size_t)l * n), sum);
        }
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < MC_THREADS / 32; ++i) s += sW[i];
        partial[((size_t)cloud * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}
__global__ void reduce_partials_kernel(int per_cloud, const float* __restrict__ partial, float* __restrict__ out) {
    __shared__ float sW[8];
    const int cloud = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < per_cloud; i += blockDim.x) s += partial[(size_t)cloud * per_cloud + i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sW[i];
        out[cloud] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// match_cost gradient.                                                              (tf_approxmatch.cu:229-291)
//   grad1[k] = sum_l match[l,k] * (p1_k - p2_l) * rsqrt(max(d2, 1e-20))   thread per k, l-range per CTA, partials
//   grad2[l] = sum_k match[l,k] * (p2_l - p1_k) * rsqrt(max(d2, 1e-20))   warp per l, lanes stride k, shuffle reduce
// ---------------------------------------------------------------------------------------------------------------
constexpr int G1_THREADS = 128;
constexpr int G1_L = 512;
__global__ void __launch_bounds__(G1_THREADS) matchcostgrad1_kernel(int n, int m, int nlt, const float* __restrict__ xyz1,
                                                                    const float* __restrict__ xyz2, const float* __restrict__ match,
                                                                    float* __restrict__ partial) {
    __shared__ float sP[G1_L * 3];
    const int cloud = blockIdx.z;
    const int k = blockIdx.x * G1_THREADS + threadIdx.x;
    const int l0 = blockIdx.y * G1_L;
    const int nl = min(G1_L, m - l0);
    for (int i = threadIdx.x; i < nl * 3; i += G1_THREADS) sP[i] = xyz2[((size_t)cloud * m + l0) * 3 + i];
    __syncthreads();
    if (k >= n) return;
    const float* p = xyz1 + ((size_t)cloud * n + k) * 3;
    const float x1 = p[0], y1 = p[1], z1 = p[2];
    const float* __restrict__ mp = match + ((size_t)cloud * m + l0) * n + k;
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 4
    for (int l = 0; l < nl; ++l) {
        const float dx = x1 - sP[l * 3], dy = y1 - sP[l * 3 + 1], dz = z1 - sP[l * 3 + 2];
        const float s = __ldg(mp + (size_t)l * n) * rsqrtf(fmaxf(sqdist3<true>(dx, dy, dz), 1e-20f));
        gx = __fmaf_rn(dx, s, gx); gy = __fmaf_rn(dy, s, gy); gz = __fmaf_rn(dz, s, gz);
    }
    float* o = partial + (((size_t)cloud * nlt + blockIdx.y) * n + k) * 3;
    o[0] = gx; o[1] = gy; o[2] = gz;
}
__global__ void matchcostgrad1_reduce_kernel(int n, int nlt, size_t bn, const float* __restrict__ partial, float* __restrict__ grad1) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over bn*3
    if (t >= bn * 3) return;
    const size_t cloud = t / ((size_t)n * 3), r = t % ((size_t)n * 3);
    float s = 0.f;
    for (int i = 0; i < nlt; ++i) s += partial[((size_t)cloud * nlt + i) * n * 3 + r];
    grad1[t] = s;
}
constexpr int G2_WARPS = 8;
__global__ void __launch_bounds__(G2_WARPS * 32) matchcostgrad2_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                                       const float* __restrict__ match, float* __restrict__ grad2) {
    const int cloud = blockIdx.y;
    const int l = blockIdx.x * G2_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (l >= m) return;
    const float* q = xyz2 + ((size_t)cloud * m + l) * 3;
    const float x2 = q[0], y2 = q[1], z2 = q[2];
    const float* __restrict__ a = xyz1 + (size_t)cloud * n * 3;
    const float* __restrict__ mp = match + ((size_t)cloud * m + l) * n;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int k = lane; k < n; k += 32) {
        const float dx = x2 - a[(size_t)k * 3], dy = y2 - a[(size_t)k * 3 + 1], dz = z2 - a[(size_t)k * 3 + 2];
        const float s = __ldg(mp + k) * rsqrtf(fmaxf(sqdist3<true>(dx, dy, dz), 1e-20f));
        gx = __fmaf_rn(dx, s, gx); gy = __fmaf_rn(dy, s, gy); gz = __fmaf_rn(dz, s, gz);
    }
    gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
    if (lane == 0) {
        float* o = grad2 + ((size_t)cloud * m + l) * 3;
        o[0] = gx; o[1] = gy; o[2] = gz;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Vector variants of match_cost / match_cost_grad for n % 4 == 0 (rows of `match` 16-byte aligned): each thread owns four
// consecutive k, reads `match` as float4, does the distance algebra on packed pairs and uses one MUFU per element
// (sqrt.approx / rsqrt.approx; the reductions are float sums in a different order than the reference anyway, and the
// results are held to 1e-4).  These are single streaming passes over `match`: 4 algorithmic bytes per pair, HBM-bound.
// ---------------------------------------------------------------------------------------------------------------
struct Pts4 {  // four consecutive points, negated, as packed pairs (01) and (23)
    float2 nx01, nx23, ny01, ny23, nz01, nz23;
};
__device__ __forceinline__ Pts4 load_pts4_neg(const float* __restrict__ p) {  // p 16-byte aligned, 12 floats
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1), c = __ldg(reinterpret_cast<const float4*>(p) + 2);
    Pts4 r;
    r.nx01 = make_float2(-a.x, -a.w); r.nx23 = make_float2(-b.z, -c.y);
    r.ny01 = make_float2(-a.y, -b.x); r.ny23 = make_float2(-b.w, -c.z);
    r.nz01 = make_float2(-a.z, -b.y); r.nz23 = make_float2(-c.x, -c.w);
    return r;
}
constexpr int MV_THREADS = 128;
__global__ void __launch_bounds__(MV_THREADS) matchcost_v4_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                                  const float* __restrict__ match, float* __restrict__ partial) {
    __shared__ float sP[MC_L * 3];
    __shared__ float sW[MV_THREADS / 32];
    const int cloud = blockIdx.z;
    const int k4 = (blockIdx.x * MV_THREADS + threadIdx.x) * 4;
    const int l0 = blockIdx.y * MC_L;
    const int nl = min(MC_L, m - l0);
    for (int i = threadIdx.x; i < nl * 3; i += MV_THREADS) sP[i] = xyz2[((size_t)cloud * m + l0) * 3 + i];
    __syncthreads();
    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
    if (k4 < n) {
        const Pts4 q = load_pts4_neg(xyz1 + ((size_t)cloud * n + k4) * 3);
        const float4* __restrict__ mp = reinterpret_cast<const float4*>(match + ((size_t)cloud * m + l0) * n + k4);
        const size_t stride = (size_t)(n >> 2);
#pragma unroll 8
        for (int l = 0; l < nl; ++l) {
            const float4 mv = __ldg(mp + (size_t)l * stride);
            const float px = sP[l * 3], py = sP[l * 3 + 1], pz = sP[l * 3 + 2];
            const float2 d01 = sqdist3x2<true>(__fadd2_rn(q.nx01, make_float2(px, px)), __fadd2_rn(q.ny01, make_float2(py, py)), __fadd2_rn(q.nz01, make_float2(pz, pz)));
            const float2 d23 = sqdist3x2<true>(__fadd2_rn(q.nx23, make_float2(px, px)), __fadd2_rn(q.ny23, make_float2(py, py)), __fadd2_rn(q.nz23, make_float2(pz, pz)));
            s01 = __ffma2_rn(make_float2(sqrt_approx(d01.x), sqrt_approx(d01.y)), make_float2(mv.x, mv.y), s01);
            s23 = __ffma2_rn(make_float2(sqrt_approx(d23.x), sqrt_approx(d23.y)), make_float2(mv.z, mv.w), s23);
        }
    }
    float sum = warp_sum((s01.x + s01.y) + (s23.x + s23.y));
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < MV_THREADS / 32; ++i) t += sW[i];
        partial[((size_t)cloud * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}

// grad1 partials over an l-range: thread owns 4 k (12 accumulators), same partial layout as matchcostgrad1_kernel
__global__ void __launch_bounds__(MV_THREADS) matchcostgrad1_v4_kernel(int n, int m, int nlt, const float* __restrict__ xyz1,
                                                                       const float* __restrict__ xyz2, const float* __restrict__ match,
                                                                       float* __restrict__ partial) {
    __shared__ float sP[G1_L * 3];
    const int cloud = blockIdx.z;
    const int k4 = (blockIdx.x * MV_THREADS + threadIdx.x) * 4;
    const int l0 = blockIdx.y * G1_L;
    const int nl = min(G1_L, m - l0);
    for (int i = threadIdx.x; i < nl * 3; i += MV_THREADS) sP[i] = xyz2[((size_t)cloud * m + l0) * 3 + i];
    __syncthreads();
    if (k4 >= n) return;
    const Pts4 q = load_pts4_neg(xyz1 + ((size_t)cloud * n + k4) * 3);
    const float4* __restrict__ mp = reinterpret_cast<const float4*>(match + ((size_t)cloud * m + l0) * n + k4);
    const size_t stride = (size_t)(n >> 2);
    float2 gx01 = make_float2(0.f, 0.f), gx23 = gx01, gy01 = gx01, gy23 = gx01, gz01 = gx01, gz23 = gx01;
#pragma unroll 4
    for (int l = 0; l < nl; ++l) {
        const float4 mv = __ldg(mp + (size_t)l * stride);
        const float px = sP[l * 3], py = sP[l * 3 + 1], pz = sP[l * 3 + 2];
        // e = p2 - p1 (packed); grad1 accumulates (p1 - p2) * s = e * (-s)
        const float2 ex01 = __fadd2_rn(q.nx01, make_float2(px, px)), ey01 = __fadd2_rn(q.ny01, make_float2(py, py)), ez01 = __fadd2_rn(q.nz01, make_float2(pz, pz));
        const float2 ex23 = __fadd2_rn(q.nx23, make_float2(px, px)), ey23 = __fadd2_rn(q.ny23, make_float2(py, py)), ez23 = __fadd2_rn(q.nz23, make_float2(pz, pz));
        const float2 d01 = sqdist3x2<true>(ex01, ey01, ez01), d23 = sqdist3x2<true>(ex23, ey23, ez23);
        const float2 ns01 = make_float2(-mv.x * rsqrtf(fmaxf(d01.x, 1e-20f)), -mv.y * rsqrtf(fmaxf(d01.y, 1e-20f)));
        const float2 ns23 = make_float2(-mv.z * rsqrtf(fmaxf(d23.x, 1e-20f)), -mv.w * rsqrtf(fmaxf(d23.y, 1e-20f)));
        gx01 = __ffma2_rn(ex01, ns01, gx01); gy01 = __ffma2_rn(ey01, ns01, gy01); gz01 = __ffma2_rn(ez01, ns01, gz01);
        gx23 = __ffma2_rn(ex23, ns23, gx23); gy23 = __ffma2_rn(ey23, ns23, gy23); gz23 = __ffma2_rn(ez23, ns23, gz23);
    }
    float* o = partial + (((size_t)cloud * nlt + blockIdx.y) * n + k4) * 3;  // 12 consecutive floats, 16-byte aligned
    reinterpret_cast<float4*>(o)[0] = make_float4(gx01.x, gy01.x, gz01.x, gx01.y);
    reinterpret_cast<float4*>(o)[1] = make_float4(gy01.y, gz01.y, gx23.x, gy23.x);
    reinterpret_cast<float4*>(o)[2] = make_float4(gz23.x, gx23.y, gy23.y, gz23.y);
}

// grad2: one warp per l, 16 rows per CTA sharing shared-memory tiles of xyz1 (stored negated, structure-of-arrays, so the
// differences p2 - p1 are FADD2 with a broadcast scalar on packed pairs); lanes stride k in float4 chunks of `match`
constexpr int G2V_WARPS = 16;
constexpr int G2V_TILE = 1024;  // points of xyz1 per tile (12 KiB)
__global__ void __launch_bounds__(G2V_WARPS * 32) matchcostgrad2_v4_kernel(int n, int m, const float* __restrict__ xyz1,
                                                                           const float* __restrict__ xyz2, const float* __restrict__ match,
                                                                           float* __restrict__ grad2) {
    __shared__ __align__(16) float sx[G2V_TILE], sy[G2V_TILE], sz[G2V_TILE];
    const int cloud = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.x * G2V_WARPS + warp;
    const bool live = l < m;
    float x2 = 0.f, y2 = 0.f, z2 = 0.f;
    if (live) {
        const float* q = xyz2 + ((size_t)cloud * m + l) * 3;
        x2 = q[0]; y2 = q[1]; z2 = q[2];
    }
    const float2 X2 = make_float2(x2, x2), Y2 = make_float2(y2, y2), Z2 = make_float2(z2, z2);
    const float* __restrict__ a = xyz1 + (size_t)cloud * n * 3;
    const float* __restrict__ mrow = match + ((size_t)cloud * m + (live ? l : 0)) * n;
    float2 gx = make_float2(0.f, 0.f), gy = gx, gz = gx;
    for (int t0 = 0; t0 < n; t0 += G2V_TILE) {
        const int len = min(G2V_TILE, n - t0);  // multiple of 4
        __syncthreads();
        for (int i4 = threadIdx.x * 4; i4 < len; i4 += G2V_WARPS * 32 * 4) {  // four points per thread: 3 float4 in, 3 float4 out
            const float4* src = reinterpret_cast<const float4*>(a + (size_t)(t0 + i4) * 3);
            const float4 A = __ldg(src), B = __ldg(src + 1), C = __ldg(src + 2);
            *reinterpret_cast<float4*>(&sx[i4]) = make_float4(-A.x, -A.w, -B.z, -C.y);
            *reinterpret_cast<float4*>(&sy[i4]) = make_float4(-A.y, -B.x, -B.w, -C.z);
            *reinterpret_cast<float4*>(&sz[i4]) = make_float4(-A.z, -B.y, -C.x, -C.w);
        }
        __syncthreads();
        if (live) {
#pragma unroll 2
            for (int k4 = lane * 4; k4 < len; k4 += 128) {
                const float4 mv = __ldg(reinterpret_cast<const float4*>(mrow + t0 + k4));
                const float4 NX = *reinterpret_cast<const float4*>(&sx[k4]), NY = *reinterpret_cast<const float4*>(&sy[k4]), NZ = *reinterpret_cast<const float4*>(&sz[k4]);
                // e = p2 - p1 for the four points (packed pairs 01, 23)
                const float2 ex01 = __fadd2_rn(make_float2(NX.x, NX.y), X2), ex23 = __fadd2_rn(make_float2(NX.z, NX.w), X2);
                const float2 ey01 = __fadd2_rn(make_float2(NY.x, NY.y), Y2), ey23 = __fadd2_rn(make_float2(NY.z, NY.w), Y2);
                const float2 ez01 = __fadd2_rn(make_float2(NZ.x, NZ.y), Z2), ez23 = __fadd2_rn(make_float2(NZ.z, NZ.w), Z2);
                const float2 d01 = sqdist3x2<true>(ex01, ey01, ez01), d23 = sqdist3x2<true>(ex23, ey23, ez23);
                const float2 s01 = __fmul2_rn(make_float2(mv.x, mv.y), make_float2(rsqrtf(fmaxf(d01.x, 1e-20f)), rsqrtf(fmaxf(d01.y, 1e-20f))));
                const float2 s23 = __fmul2_rn(make_float2(mv.z, mv.w), make_float2(rsqrtf(fmaxf(d23.x, 1e-20f)), rsqrtf(fmaxf(d23.y, 1e-20f))));
                gx = __ffma2_rn(ex01, s01, gx); gy = __ffma2_rn(ey01, s01, gy); gz = __ffma2_rn(ez01, s01, gz);
                gx = __ffma2_rn(ex23, s23, gx); gy = __ffma2_rn(ey23, s23, gy); gz = __ffma2_rn(ez23, s23, gz);
            }
        }
    }
    const float sx_ = warp_sum(gx.x + gx.y), sy_ = warp_sum(gy.x + gy.y), sz_ = warp_sum(gz.x + gz.y);
    if (live && lane == 0) {
        float* o = grad2 + ((size_t)cloud * m + l) * 3;
        o[0] = sx_; o[1] = sy_; o[2] = sz_;
    }
}

// Sweep launch policy (numbers from tools/emd_tune.cu on B200, profiles/r1_emd_tune_*.txt).  The sweep is MUFU-bound and
// needs ~40 resident warps per SM to hide the ex2 latency; Q = 4 rows per thread is the sweet spot.
//   * No candidate split when the batch alone provides >= 6.5 CTAs per SM (e.g. B=32 x 16384 rows): every row sum then
//     follows the reference's order bit for bit; costs ~9 % against the split grid (77 % vs 84 % of MUFU peak).
//   * Otherwise (few or small clouds) the candidate range is split across up to 32 CTAs per row tile, ~20 CTAs per SM, and
//     the partials are summed in split order: deterministic, but a different rounding order than the reference's chain.
struct SweepPlan { int Q, nrt, nsplit, cps; };
static SweepPlan emd_plan(int b, int nr, int nc) {
    SweepPlan p;
    const int chunks = (nc + EMD_TC - 1) / EMD_TC;
    p.Q = nr >= EMD_THREADS * 4 ? 4 : 2;
    p.nrt = (nr + EMD_THREADS * p.Q - 1) / (EMD_THREADS * p.Q);
    if ((long)b * p.nrt * 2 >= 13L * kNumSMs) { p.nsplit = 1; p.cps = chunks; return p; }
    long want = ((long)kNumSMs * 20 + (long)b * p.nrt - 1) / ((long)b * p.nrt);
    int nsplit = (int)(want < 1 ? 1 : want);
    if (nsplit > chunks) nsplit = chunks;
    if (nsplit > 32) nsplit = 32;
    p.cps = (chunks + nsplit - 1) / nsplit;
    p.nsplit = (chunks + p.cps - 1) / p.cps;
    return p;
}
template <int Q, int EPI>
static void emd_sweep_q(const SweepPlan& p, unsigned grid, int nr, int nc, float lvl2, float init0, const float* rows, const float* cands,
                        const float* w, const float* rowfac, float* partial, const EmdEpi& epi, cudaStream_t s) {
    if (lvl2 == 0.0f)
        emd_sweep_kernel<Q, EPI, true><<<grid, EMD_THREADS, 0, s>>>(nr, nc, p.nrt, p.nsplit, p.cps, lvl2, init0, rows, cands, w, rowfac, partial, epi);
    else
        emd_sweep_kernel<Q, EPI, false><<<grid, EMD_THREADS, 0, s>>>(nr, nc, p.nrt, p.nsplit, p.cps, lvl2, init0, rows, cands, w, rowfac, partial, epi);
}
// nsplit_out == 1 means the sweep applied the pass's rule itself; otherwise the caller launches the epilogue kernel
template <int EPI>
static void emd_sweep(int b, int nr, int nc, float lvl2, float init0, const float* rows, const float* cands, const float* w, const float* rowfac,
                      float* partial, const EmdEpi& epi, int& nsplit_out, cudaStream_t s, const int* perm = nullptr, const unsigned* mask = nullptr) {
    const SweepPlan p = emd_plan(b, nr, nc);
    nsplit_out = p.nsplit;
    const unsigned grid = (unsigned)(b * p.nrt * p.nsplit);
    if (mask && p.Q == 4) {
        emd_sweep_pruned_kernel<EPI><<<grid, EMD_THREADS, 0, s>>>(nr, nc, p.nrt, p.nsplit, p.cps, lvl2, init0, rows, cands, w, rowfac, perm, mask,
                                                                 (nc + 31) / 32, partial, epi);
        return;
    }
    if (p.Q == 4) emd_sweep_q<4, EPI>(p, grid, nr, nc, lvl2, init0, rows, cands, w, rowfac, partial, epi, s);
    else emd_sweep_q<2, EPI>(p, grid, nr, nc, lvl2, init0, rows, cands, w, rowfac, partial, epi, s);
}

// sweeps (-> per-level factors in the workspace), then one pass that materialises the matrix and/or reduces the cost
static int emd_run(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* cost, float* ws_floats, cudaStream_t s) {
    const size_t bn = (size_t)b * n, bm = (size_t)b * m;
    EmdWs ws = emd_carve(ws_floats, b, n, m);
    const float multiL = n >= m ? 1.0f : (float)(m / n), multiR = n >= m ? (float)(n / m) : 1.0f;  // integer division, tf_approxmatch.cu:4-10
    emd_init_kernel<<<(unsigned)((bn + bm + 255) / 256), 256, 0, s>>>(bn, bm, multiL, multiR, ws.remainL, ws.remainR);
    EmdLevels lv;
    for (int li = 0; li < EMD_LEVELS; ++li) lv.lvl2[li] = emd_level(li) * LOG2E;
    // exact pruning of the three sharpest levels (see emd_sweep_pruned_kernel); RFNET_EMD_NO_PRUNE=1 forces the dense sweeps
    const char* no_prune = getenv("RFNET_EMD_NO_PRUNE");
    const bool prune = emd_prune_enabled(n, m) && !(no_prune && no_prune[0] == '1');
    if (prune) {
        { const int rc = morton_sort(b, n, m, xyz1, xyz2, ws.perm1, ws.perm2, s); if (rc) return rc; }
        emd_mask_kernel<<<dim3((unsigned)((n + EMD_CLUSTER - 1) / EMD_CLUSTER), (unsigned)b), 256, 0, s>>>(n, m, (m + 31) / 32, xyz1, ws.perm1, xyz2, lv.lvl2[0],
                                                                                                       lv.lvl2[1], lv.lvl2[2], ws.maskA);
        emd_mask_kernel<<<dim3((unsigned)((m + EMD_CLUSTER - 1) / EMD_CLUSTER), (unsigned)b), 256, 0, s>>>(m, n, (n + 31) / 32, xyz2, ws.perm2, xyz1, lv.lvl2[0],
                                                                                                       lv.lvl2[1], lv.lvl2[2], ws.maskB);
    }
    for (int li = 0; li < EMD_LEVELS; ++li) {
        const float lvl2 = lv.lvl2[li];
        int ns;
        const bool pl = prune && li < EMD_PRUNE_LEVELS;
        const unsigned* mA = pl ? ws.maskA + (size_t)li * ((m + 31) / 32) : nullptr;
        const unsigned* mB = pl ? ws.maskB + (size_t)li * ((n + 31) / 32) : nullptr;
        // pass 1: rows = xyz1 (k), candidates = xyz2 (l) weighted by remainR            -> ratioL, facL
        emd_sweep<1>(b, n, m, lvl2, 1e-9f, xyz1, xyz2, ws.remainR, nullptr, ws.partial, EmdEpi{ws.remainL, ws.ratioL, ws.facL + (size_t)li * bn}, ns, s,
                     ws.perm1, mA);
        if (ns > 1) emd_epi1_kernel<<<(unsigned)((bn + 255) / 256), 256, 0, s>>>(n, ns, bn, ws.partial, ws.remainL, ws.ratioL, ws.facL + (size_t)li * bn);
        // pass 2: rows = xyz2 (l), candidates = xyz1 (k) weighted by ratioL             -> ratioR, facR, remainR
        emd_sweep<2>(b, m, n, lvl2, 0.0f, xyz2, xyz1, ws.ratioL, nullptr, ws.partial, EmdEpi{ws.remainR, ws.ratioR, ws.facR + (size_t)li * bm}, ns, s,
                     ws.perm2, mB);
        if (ns > 1) emd_epi2_kernel<<<(unsigned)((bm + 255) / 256), 256, 0, s>>>(m, ns, bm, ws.partial, ws.remainR, ws.ratioR, ws.facR + (size_t)li * bm);
        // pass 3: rows = xyz1 (k), candidates = xyz2 (l) weighted by ratioR, row factor ratioL  -> remainL
        emd_sweep<3>(b, n, m, lvl2, 0.0f, xyz1, xyz2, ws.ratioR, ws.ratioL, ws.partial, EmdEpi{ws.remainL, nullptr, nullptr}, ns, s, ws.perm1, mA);
        if (ns > 1) emd_epi3_kernel<<<(unsigned)((bn + 255) / 256), 256, 0, s>>>(n, ns, bn, ws.partial, ws.remainL);
    }
    dim3 grid((unsigned)((n + 2 * MT_THREADS - 1) / (2 * MT_THREADS)), (unsigned)((m + MT_L - 1) / MT_L), (unsigned)b);
    RFNET_CHECK_ARG(grid.y <= 65535);
    float* cpart = ws_floats + emd_ws_floats(b, n, m);  // only carved when a cost is requested (rfnet_emd_cost_workspace_bytes)
    if (match && cost)
        emd_materialise_kernel<true, true><<<grid, MT_THREADS, 0, s>>>(n, m, bn, bm, lv, xyz1, xyz2, ws.facL, ws.facR, match, cpart);
    else if (cost)
        emd_materialise_kernel<false, true><<<grid, MT_THREADS, 0, s>>>(n, m, bn, bm, lv, xyz1, xyz2, ws.facL, ws.facR, nullptr, cpart);
    else
        emd_materialise_kernel<true, false><<<grid, MT_THREADS, 0, s>>>(n, m, bn, bm, lv, xyz1, xyz2, ws.facL, ws.facR, match, nullptr);
    if (cost) reduce_partials_kernel<<<b, 256, 0, s>>>((int)(grid.x * grid.y), cpart, cost);
    return launch_status();
}

}  // namespace rfnet

using namespace rfnet;

extern "C" size_t rfnet_approxmatch_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return emd_ws_floats(b, n, m) * sizeof(float);
}

extern "C" int rfnet_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, void* workspace,
                                 size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0 || n == 0 || m == 0) return 0;
    RFNET_CHECK_ARG(xyz1 && xyz2 && match && workspace && workspace_bytes >= rfnet_approxmatch_workspace_bytes(b, n, m));
    RFNET_CHECK_ARG(b <= 65535);
    return emd_run(b, n, m, xyz1, xyz2, match, nullptr, (float*)workspace, (cudaStream_t)stream);
}

extern "C" size_t rfnet_emd_cost_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    const size_t tiles = (size_t)((n + 2 * MT_THREADS - 1) / (2 * MT_THREADS)) * ((m + MT_L - 1) / MT_L);
    return (emd_ws_floats(b, n, m) + (size_t)b * tiles) * sizeof(float);
}

extern "C" int rfnet_emd_cost(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* cost, void* workspace,
                              size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0) return 0;
    RFNET_CHECK_ARG(cost);
    if (n == 0 || m == 0) {
        RFNET_CUDA(cudaMemsetAsync(cost, 0, sizeof(float) * b, (cudaStream_t)stream));
        return 0;
    }
    RFNET_CHECK_ARG(xyz1 && xyz2 && workspace && workspace_bytes >= rfnet_emd_cost_workspace_bytes(b, n, m));
    RFNET_CHECK_ARG(b <= 65535);
    return emd_run(b, n, m, xyz1, xyz2, match, cost, (float*)workspace, (cudaStream_t)stream);
}

extern "C" size_t rfnet_matchcost_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return sizeof(float) * (size_t)b * ((n + MC_THREADS - 1) / MC_THREADS) * ((m + MC_L - 1) / MC_L);
}

extern "C" int rfnet_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* out, void* workspace,
                               size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0) return 0;
    RFNET_CHECK_ARG(out);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        RFNET_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * b, s));
        return 0;
    }
    RFNET_CHECK_ARG(xyz1 && xyz2 && match && workspace && workspace_bytes >= rfnet_matchcost_workspace_bytes(b, n, m) && b <= 65535);
    dim3 grid((unsigned)((n + MC_THREADS - 1) / MC_THREADS), (unsigned)((m + MC_L - 1) / MC_L), (unsigned)b);
    RFNET_CHECK_ARG(grid.y <= 65535);
    const bool vec = (n % 4 == 0) && ((((uintptr_t)match | (uintptr_t)xyz1) & 15u) == 0);
    if (vec) {
        grid.x = (unsigned)((n + MV_THREADS * 4 - 1) / (MV_THREADS * 4));
        matchcost_v4_kernel<<<grid, MV_THREADS, 0, s>>>(n, m, xyz1, xyz2, match, (float*)workspace);
    } else {
        matchcost_kernel<<<grid, MC_THREADS, 0, s>>>(n, m, xyz1, xyz2, match, (float*)workspace);
    }
    reduce_partials_kernel<<<b, 256, 0, s>>>((int)(grid.x * grid.y), (const float*)workspace, out);
    return launch_status();
}

extern "C" size_t rfnet_matchcostgrad_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return sizeof(float) * 3 * (size_t)b * n * ((m + G1_L - 1) / G1_L);
}

extern "C" int rfnet_matchcostgrad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* grad1, float* grad2,
                                   void* workspace, size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        if (n) RFNET_CUDA(cudaMemsetAsync(grad1, 0, sizeof(float) * 3 * (size_t)b * n, s));
        if (m) RFNET_CUDA(cudaMemsetAsync(grad2, 0, sizeof(float) * 3 * (size_t)b * m, s));
        return 0;
    }
    RFNET_CHECK_ARG(xyz1 && xyz2 && match && grad1 && grad2 && workspace && workspace_bytes >= rfnet_matchcostgrad_workspace_bytes(b, n, m) && b <= 65535);
    const int nlt = (m + G1_L - 1) / G1_L;
    RFNET_CHECK_ARG(nlt <= 65535);
    const size_t bn = (size_t)b * n;
    const bool vec = (n % 4 == 0) && ((((uintptr_t)match | (uintptr_t)xyz1 | (uintptr_t)workspace) & 15u) == 0);
    if (vec) {
        dim3 g1((unsigned)((n + MV_THREADS * 4 - 1) / (MV_THREADS * 4)), (unsigned)nlt, (unsigned)b);
        matchcostgrad1_v4_kernel<<<g1, MV_THREADS, 0, s>>>(n, m, nlt, xyz1, xyz2, match, (float*)workspace);
    } else {
        dim3 g1((unsigned)((n + G1_THREADS - 1) / G1_THREADS), (unsigned)nlt, (unsigned)b);
        matchcostgrad1_kernel<<<g1, G1_THREADS, 0, s>>>(n, m, nlt, xyz1, xyz2, match, (float*)workspace);
    }
    matchcostgrad1_reduce_kernel<<<(unsigned)((bn * 3 + 255) / 256), 256, 0, s>>>(n, nlt, bn, (const float*)workspace, grad1);
    if (vec) {
        dim3 g2((unsigned)((m + G2V_WARPS - 1) / G2V_WARPS), (unsigned)b);
        matchcostgrad2_v4_kernel<<<g2, G2V_WARPS * 32, 0, s>>>(n, m, xyz1, xyz2, match, grad2);
    } else {
        dim3 g2((unsigned)((m + G2_WARPS - 1) / G2_WARPS), (unsigned)b);
        matchcostgrad2_kernel<<<g2, G2_WARPS * 32, 0, s>>>(n, m, xyz1, xyz2, match, grad2);
    }
    return launch_status();
}
