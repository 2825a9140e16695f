// approx_match (approximate EMD transport plan), match_cost and match_cost gradient for sm_100a.
//
// Replaces approxmatch / matchcost / matchcostgrad1 / matchcostgrad2 (pc_distance/tf_approxmatch.cu:1-295).
//
// The reference runs ONE 512-thread block per cloud through 10 levels x 3 dependent n*m passes and read-modify-writes the
// n*m match matrix once per level (80 B of HBM traffic per pair).  Here:
//   * every pass is the same reduction  S[row] = sum_c exp(level * d2(row, c)) * w[c]  ("weighted exp-sum sweep") with the
//     roles of the two clouds swapped between passes; it runs on the whole chip: grid = clouds x row tiles x candidate
//     splits, rows in registers as packed pairs (FADD2/FMUL2/FFMA2), candidates + weights broadcast from shared memory,
//     one MUFU.EX2 per pair-pass;
//   * pass 3 of level j and pass 1 of level j-1 have the same rows (xyz1), the same candidates (xyz2) and weights that are
//     both final once pass 2 of level j is done, so they run as ONE sweep that evaluates the distance once and keeps two
//     accumulators (21 sweeps instead of 30, a third less FP32 work; each chain of fma's is unchanged, bit for bit);
//   * the pass's update rule (ratioL / consumption+ratioR / remainL) is applied by the sweep itself (a thread owns the whole
//     sum of its row).  Nothing depends on the batch size: a cloud gets the same bits whether it is evaluated alone or inside
//     any batch, so sharding a batch over GPUs cannot change a result (with RFNET_EMD_SPLIT_SUMS the split plan is a function
//     of (n, m) only, for the same reason);
//   * match is NOT accumulated level by level.  The per-level factors ratioL_j[k], ratioR_j[l] are kept (10*(n+m) floats per
//     cloud) and match[l,k] = sum_j e_j(k,l) * ratioL_j[k] * ratioR_j[l] is written ONCE at the end (4 B/pair of HBM traffic
//     instead of 80) -- or never: the loss-level entry points reduce the cost and both gradients from the same factors.
// exp(x) is ex2.approx(x * log2e) exactly as __expf in the reference; level * log2e is folded into one constant, which is
// bit-identical because level is a power of two.  d2 is the reference's fused expression.  Offsets are 64-bit.
//
// Summation order.  By DEFAULT every row sum is ONE chain over the candidates in ascending order -- the reference thread's
// chain of fma's (tf_approxmatch.cu:26-160 as compiled: FMUL dy*dy, FFMA dx, FFMA dz, FMUL level, FMUL log2e, MUFU.EX2, FFMA
// accumulate) -- so the iteration follows the reference's rounding order term for term.  The default differs from the
// reference binary only by (a) the flushing exponential (terms below 1.2e-38, < 1e-29 absolute after weighting) and (b) the
// derived exponentials of the final pass (<= 1.1e-6 relative on a matrix entry).  RFNET_EMD_EXACT removes both (non-ftz
// exponential, no pruning, nine MUFU exponentials per entry): `match` is then BIT-IDENTICAL to the reference CUDA kernel's
// (tests/test_emd_gpu.py, at B=32 2048^2 and 16384^2 too).  RFNET_EMD_SPLIT_SUMS cuts each sum into fixed-length pieces
// (more parallelism for one or two small clouds; a different rounding order: measured 7e-4 of the largest entry).
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
// candidates per split (multiples of 32): tuned on B200 with tools/emd_tune.cu, profiles/r2_emd_tune.txt
#ifndef EMD_SPLIT_SMALL
#define EMD_SPLIT_SMALL 256    // clouds of up to EMD_SPLIT_SMALL_MAX candidates
#endif
#ifndef EMD_SPLIT_LARGE
#define EMD_SPLIT_LARGE 512
#endif
constexpr int EMD_SPLIT_SMALL_MAX = 4096;
constexpr int EMD_MAX_SPLITS = 64;
#define EMD_PRAGMA_(x) _Pragma(#x)
#define EMD_PRAGMA_UNROLL(n) EMD_PRAGMA_(unroll n)

__host__ __device__ inline float emd_level(int li) {  // li = 0..9  ->  j = 7..-2 ; level = -4^j, 0 at j = -2
    const int j = 7 - li;
    if (j == -2) return 0.0f;
    float v = 1.0f;
    for (int i = 0; i < (j < 0 ? -j : j); ++i) v *= 4.0f;
    return j >= 0 ? -v : -1.0f / v;
}

// ---------------------------------------------------------------------------------------------------------------
// Sweep plan: a function of the cloud sizes and the flags only (NOT of the batch size -- see the header).
// B200 tuning (profiles/r1_emd_tune_*.txt, r2_emd_tune.txt): the sweep is MUFU-bound and wants ~40 resident warps per SM;
// Q = 4 rows per thread with the candidate range cut into many short splits is the best or within 1 % of the best grid at
// every batch size from 1 to 32 clouds, so no batch-dependent choice is needed.
// ---------------------------------------------------------------------------------------------------------------
struct SweepPlan { int Q, nrt, nsplit, split_len; };
static int emd_split_len(int nc) {
    int sl = nc <= EMD_SPLIT_SMALL_MAX ? EMD_SPLIT_SMALL : EMD_SPLIT_LARGE;
    if ((nc + sl - 1) / sl > EMD_MAX_SPLITS) sl = (((nc + EMD_MAX_SPLITS - 1) / EMD_MAX_SPLITS) + 31) / 32 * 32;
    return sl;
}
// the split plan (RFNET_EMD_SPLIT_SUMS); without that flag every row sum is one chain and no plan is needed
static SweepPlan emd_plan(int nr, int nc) {
    SweepPlan p;
    p.Q = nr >= EMD_THREADS * 4 ? 4 : 2;
    p.nrt = (nr + EMD_THREADS * p.Q - 1) / (EMD_THREADS * p.Q);
    p.split_len = emd_split_len(nc);
    p.nsplit = (nc + p.split_len - 1) / p.split_len;
    return p;
}

// Workspace layout (floats), per call:
//   remainL [b*n] remainR [b*m] ratioL [b*n] ratioR [b*m]                      running state of the current level
//   facL [EMD_LEVELS][b*n]  facR [EMD_LEVELS][b*m]                             per-level factors for the final pass
//   partial                                                                    sweep partial sums (two sets for the fused sweep)
struct EmdWs {
    float *remainL, *remainR, *ratioL, *ratioR, *facL, *facR, *partial;
    int *perm1, *perm2;          // Morton order of xyz1 / xyz2 (pruned sweeps only)
    unsigned *maskA, *maskB;     // candidate masks: rows = xyz1 clusters vs xyz2 candidates (passes 1, 3) / the reverse (pass 2)
};
// pruned sweeps pay off (and their sort fits shared memory) for clouds of 4096 .. 32768 points
static bool emd_prune_enabled(int n, int m) { return n >= 4096 && m >= 4096 && n <= 32768 && m <= 32768; }
static size_t emd_mask_words(int b, int nr, int nc) { return (size_t)b * ((nr + 127) / 128) * 3 * ((nc + 31) / 32); }
static size_t emd_partial_floats(int b, int n, int m) {
    const size_t rows1 = 2 * (size_t)n * emd_plan(n, m).nsplit;   // rows = xyz1: the fused sweep keeps two sums per row
    const size_t rows2 = (size_t)m * emd_plan(m, n).nsplit;
    return (size_t)b * (rows1 > rows2 ? rows1 : rows2);
}
static size_t emd_ws_floats(int b, int n, int m) {
    const size_t bn = (size_t)b * n, bm = (size_t)b * m;
    size_t f = 2 * (bn + bm) + (size_t)EMD_LEVELS * (bn + bm) + emd_partial_floats(b, n, m);
    if (emd_prune_enabled(n, m)) f += bn + bm + emd_mask_words(b, n, m) + emd_mask_words(b, m, n);
    return (f + 3) & ~(size_t)3;
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
    s.partial = w; w += emd_partial_floats(b, n, m);
    s.perm1 = reinterpret_cast<int*>(w); w += bn;
    s.perm2 = reinterpret_cast<int*>(w); w += bm;
    s.maskA = reinterpret_cast<unsigned*>(w); w += emd_mask_words(b, n, m);
    s.maskB = reinterpret_cast<unsigned*>(w);
    return s;
}

template <bool EXACT>
__device__ __forceinline__ float emd_ex2(float x) { return EXACT ? ex2_approx_full(x) : ex2_approx(x); }

// ---------------------------------------------------------------------------------------------------------------
// Weighted exp-sum sweep:  sum[cloud][row] = init + sum_{c ascending} ex2(lvl2 * d2(row, c)) * w[c]
//   rows: (b, nr, 3), cands: (b, nc, 3), w: (b, nc).  grid.x = b * nrt * nsplit.  Q rows per thread (Q/2 packed pairs).
// The accumulation is the reference binary's, term by term and in candidate order:
//   MODE 1 (pass 1, tf_approxmatch.cu:26-59):   acc = fma(e, w[c], acc), acc starts at 1e-9     -> ratioL = remainL / acc
//   MODE 2 (pass 2, :75-108):                   acc = fma(e, w[c], acc)                          -> consumption, ratioR, remainR
//   MODE 3 (pass 3, :127-160):                  acc = fma(rowfac[row] * e, w[c], acc)            -> remainL = max(0, remainL - acc)
//   MODE 4 = MODE 3 at this level fused with MODE 1 at the NEXT level (lvl2b, weights wb = remainR): one distance per
//            pair, two exponentials, two chains; the rule of pass 3 is applied first, then pass 1's on the updated remainL.
// UNIT: the last level (j = -2) has level = 0, i.e. e = ex2(0 * d2) = 1 for every pair; the same chain is run without
// distances or exponentials.  In MODE 4 UNIT refers to the next level (the fused pair j = -1, -2).
// With nsplit == 1 the thread owns the complete sum of its rows and applies the rule itself; with splits the partial sums
// are reduced in split order by emd_epi_kernel.  (Fusing the split case too, through a last-CTA ticket, was measured slower.)
// PRUNED (Q = 4, MODE 1..3): see below.
// ---------------------------------------------------------------------------------------------------------------
struct SweepArgs {
    int nr, nc, nrt, nsplit, split_len, nwords;
    float lvl2, lvl2b, init0;
    const float *rows, *cands, *w, *wb, *rowfac;
    float *partial, *partial_b;
    float *remain, *ratio, *fac;   // MODE 1, 3, 4: remainL / ratioL / facL[level (MODE 4: next level)]   MODE 2: remainR / ratioR / facR[level]
    const int* perm;               // PRUNED: Morton order of the rows
    const unsigned* mask;          // PRUNED: the level's candidate-mask words of cluster 0 of cloud 0
};
template <int MODE>
__device__ __forceinline__ void emd_apply(float* __restrict__ remain, float* __restrict__ ratio, float* __restrict__ fac, size_t idx, float sum, float sumb) {
    if (MODE == 1) {
        const float r = remain[idx] / sum;
        ratio[idx] = r;
        fac[idx] = r;
    } else if (MODE == 2) {
        const float rem = remain[idx];
        const float sumr = sum * rem;
        const float consumption = fminf(rem / (sumr + 1e-9f), 1.0f);
        const float r = consumption * rem;
        ratio[idx] = r;
        fac[idx] = r;
        remain[idx] = fmaxf(0.0f, rem - sumr);
    } else if (MODE == 3) {
        remain[idx] = fmaxf(0.0f, remain[idx] - sum);
    } else {
        const float rem = fmaxf(0.0f, remain[idx] - sum);
        remain[idx] = rem;
        const float r = rem / sumb;
        ratio[idx] = r;
        fac[idx] = r;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Exact pruning of the three sharpest levels (j = 7, 6, 5: e = exp(-4^j d2) with 4^j = 16384, 4096, 1024).
// ex2.approx.ftz returns EXACTLY 0 once its argument is below -126, i.e. for d2 > 0.0053 / 0.021 / 0.085, and a zero term
// leaves the accumulator bit-identical (fma(0, w, acc) == acc).  At those levels almost every pair is such a no-op, so:
//   * morton_sort_kernel orders each cloud along a Morton curve (one CTA per cloud, counting sort in shared memory);
//     a warp of the pruned sweep then owns 128 CONSECUTIVE points of that order: a spatially tight cluster;
//   * emd_mask_kernel marks, per cluster and level, the candidates whose distance to the cluster's bounding box still
//     allows a non-zero term;
//   * the PRUNED sweep is the dense one restricted to marked candidates, visited in ASCENDING candidate order:
//     every row sum goes through the same sequence of non-trivial fma's as in the dense sweep, so the result is bit-for-bit
//     the dense one (tests/test_emd_gpu.py::test_pruned_sweeps_are_exact) -- only the row -> thread assignment changes.
// The masks depend on the points only: built once per call for both roles (rows = xyz1 / rows = xyz2), used by 9 sweeps.
//
// Margin.  A candidate is dropped when lvl2 * dbox2 <= EMD_PRUNE_ARG = -130, where dbox2 is the squared distance from the
// candidate to the cluster's bounding box, evaluated in float32 with fma's.  For every row p of the cluster the true
// d2(p, c) >= dbox2_exact, and the two float evaluations (3 subtractions, 3 products, 2 sums each, every operand < 4 in
// magnitude) differ from their exact values by less than 8 ulp relative, i.e. |lvl2*d2 - lvl2*dbox2| < 130 * 8 * 2^-24 <
// 1e-4 in the argument of the exponential.  The sweep's own argument for that pair is therefore below -130 + 1e-4 < -126,
// where ex2.approx.ftz is exactly 0: a margin of 4 units covers the rounding by four orders of magnitude.
// ---------------------------------------------------------------------------------------------------------------
constexpr int EMD_PRUNE_LEVELS = 3;
constexpr int EMD_CLUSTER = 128;        // rows of one warp of the pruned sweep (4 per lane)
constexpr float EMD_PRUNE_ARG = -130.0f;

template <int Q, int MODE, bool UNIT, bool PRUNED, bool EXACT, int NT = EMD_THREADS>
__global__ void __launch_bounds__(NT) emd_sweep_kernel(const __grid_constant__ SweepArgs a) {
    constexpr bool P3 = MODE == 3 || MODE == 4;
    constexpr bool DUAL = MODE == 4;
    static_assert(!PRUNED || (Q == 4 && !DUAL && !UNIT && !EXACT), "pruned sweeps: Q = 4, single level, ftz exponential");
    __shared__ __align__(16) float4 sC[EMD_TC];
    __shared__ float sB[DUAL ? EMD_TC : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bid = blockIdx.x;
    const int split = bid % a.nsplit;
    const int tile = (bid / a.nsplit) % a.nrt;
    const int cloud = bid / (a.nsplit * a.nrt);
    const int nr = a.nr, nc = a.nc;
    const float* __restrict__ rbase = a.rows + (size_t)cloud * nr * 3;
    const float* __restrict__ cbase = a.cands + (size_t)cloud * nc * 3;
    const float* __restrict__ wbase = a.w + (size_t)cloud * nc;
    const float* __restrict__ wbbase = DUAL ? a.wb + (size_t)cloud * nc : nullptr;

    // rows of this thread: dense = tile-strided; pruned = the warp's cluster of 128 consecutive Morton positions, 32 per q
    int row[Q];
    const unsigned* __restrict__ mrow = nullptr;
    if (PRUNED) {
        const int nclusters = (nr + EMD_CLUSTER - 1) / EMD_CLUSTER;
        const int cluster = tile * (NT * Q / EMD_CLUSTER) + warp;
        const int* __restrict__ pbase = a.perm + (size_t)cloud * nr;
        const int s0 = cluster * EMD_CLUSTER + lane;
#pragma unroll
        for (int q = 0; q < Q; ++q) row[q] = (s0 + 32 * q) < nr ? pbase[s0 + 32 * q] : -1;
        mrow = a.mask + ((size_t)cloud * nclusters + min(cluster, nclusters - 1)) * EMD_PRUNE_LEVELS * a.nwords;
    } else {
        const int r0 = tile * (NT * Q) + tid;
#pragma unroll
        for (int q = 0; q < Q; ++q) row[q] = (r0 + q * NT) < nr ? r0 + q * NT : -1;
    }
    float2 rx[Q / 2], ry[Q / 2], rz[Q / 2], rf[Q / 2], acc[Q / 2], accb[Q / 2];
    const float a0 = split == 0 ? a.init0 : 0.0f;
    const float b0 = split == 0 ? 1e-9f : 0.0f;   // pass 1 of the next level starts at 1e-9 (tf_approxmatch.cu:36)
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int ia = row[2 * h], ib = row[2 * h + 1];
        const bool va = ia >= 0, vb = ib >= 0;
        // rows are kept NEGATED so that (cand - row) is one FADD2 with a broadcast scalar: c + (-r) == c - r exactly
        rx[h].x = va ? -rbase[(size_t)ia * 3 + 0] : 0.f; ry[h].x = va ? -rbase[(size_t)ia * 3 + 1] : 0.f; rz[h].x = va ? -rbase[(size_t)ia * 3 + 2] : 0.f;
        rx[h].y = vb ? -rbase[(size_t)ib * 3 + 0] : 0.f; ry[h].y = vb ? -rbase[(size_t)ib * 3 + 1] : 0.f; rz[h].y = vb ? -rbase[(size_t)ib * 3 + 2] : 0.f;
        rf[h] = make_float2(1.f, 1.f);
        if (P3) {
            rf[h].x = va ? a.rowfac[(size_t)cloud * nr + ia] : 0.f;
            rf[h].y = vb ? a.rowfac[(size_t)cloud * nr + ib] : 0.f;
        }
        acc[h] = make_float2(a0, a0);
        accb[h] = make_float2(b0, b0);
    }
    const float2 L2 = make_float2(a.lvl2, a.lvl2), L2B = make_float2(a.lvl2b, a.lvl2b);

    auto step = [&](int k) {
        const float4 c = sC[k];
        const float2 cw = make_float2(c.w, c.w);
        float2 cwb = make_float2(0.f, 0.f);
        if (DUAL) { const float t = sB[k]; cwb = make_float2(t, t); }
#pragma unroll
        for (int h = 0; h < Q / 2; ++h) {
            if (UNIT && !DUAL) {
                // e == 1: fma(1, w, acc) == acc + w and fma(rf * 1, w, acc) == fma(rf, w, acc), bit for bit
                acc[h] = P3 ? __ffma2_rn(rf[h], cw, acc[h]) : __fadd2_rn(acc[h], cw);
            } else {
                const float2 dx = __fadd2_rn(rx[h], make_float2(c.x, c.x));  // the sign of the difference is irrelevant after squaring
                const float2 dy = __fadd2_rn(ry[h], make_float2(c.y, c.y));
                const float2 dz = __fadd2_rn(rz[h], make_float2(c.z, c.z));
                const float2 d2 = sqdist3x2<true>(dx, dy, dz);
                const float2 x = __fmul2_rn(d2, L2);
                float2 e = make_float2(emd_ex2<EXACT>(x.x), emd_ex2<EXACT>(x.y));
                if (P3) e = __fmul2_rn(rf[h], e);
                acc[h] = __ffma2_rn(e, cw, acc[h]);
                if (DUAL) {
                    if (UNIT) {
                        accb[h] = __fadd2_rn(accb[h], cwb);
                    } else {
                        const float2 xb = __fmul2_rn(d2, L2B);
                        accb[h] = __ffma2_rn(make_float2(emd_ex2<EXACT>(xb.x), emd_ex2<EXACT>(xb.y)), cwb, accb[h]);
                    }
                }
            }
        }
    };

    const int c_begin = split * a.split_len;
    const int c_end = min(nc, c_begin + a.split_len);
    for (int c0 = c_begin; c0 < c_end; c0 += EMD_TC) {
        const int len = min(EMD_TC, c_end - c0);
        __syncthreads();
        for (int i = tid; i < EMD_TC; i += NT) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // padded candidates carry weight 0: fma(e, 0, acc) == acc
            float vb = 0.f;
            if (i < len) {
                const float* c = cbase + (size_t)(c0 + i) * 3;
                v = make_float4(c[0], c[1], c[2], wbase[c0 + i]);
                if (DUAL) vb = wbbase[c0 + i];
            }
            sC[i] = v;
            if (DUAL) sB[i] = vb;
        }
        if (PRUNED) {
            // this warp's mask words for the chunk (c0 is a multiple of 32): lane i holds word i
            const int wi0 = c0 >> 5;
            const unsigned myword = (lane < EMD_TC / 32 && wi0 + lane < a.nwords) ? mrow[wi0 + lane] : 0u;
            __syncthreads();
#pragma unroll 1
            for (int wi = 0; wi < EMD_TC / 32; ++wi) {
                unsigned bits = __shfl_sync(0xffffffffu, myword, wi);
                if (wi * 32 >= len) bits = 0u;
                else if (len - wi * 32 < 32) bits &= (1u << (len - wi * 32)) - 1u;   // candidates past the split's end belong to the next split
                while (bits) {
                    const int k = wi * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    step(k);
                }
            }
        } else {
            __syncthreads();
            const int len2 = (len + EMD_UNROLL - 1) / EMD_UNROLL * EMD_UNROLL;  // padded entries carry weight 0 (EMD_TC % EMD_UNROLL == 0)
            EMD_PRAGMA_UNROLL(EMD_UNROLL)
            for (int k = 0; k < len2; ++k) step(k);
        }
    }
    if (a.nsplit == 1) {   // complete sums: apply the pass's rule here
#pragma unroll
        for (int h = 0; h < Q / 2; ++h) {
            if (row[2 * h] >= 0) emd_apply<MODE>(a.remain, a.ratio, a.fac, (size_t)cloud * nr + row[2 * h], acc[h].x, accb[h].x);
            if (row[2 * h + 1] >= 0) emd_apply<MODE>(a.remain, a.ratio, a.fac, (size_t)cloud * nr + row[2 * h + 1], acc[h].y, accb[h].y);
        }
        return;
    }
    const size_t po = ((size_t)cloud * a.nsplit + split) * nr;
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        if (row[2 * h] >= 0) { a.partial[po + row[2 * h]] = acc[h].x; if (DUAL) a.partial_b[po + row[2 * h]] = accb[h].x; }
        if (row[2 * h + 1] >= 0) { a.partial[po + row[2 * h + 1]] = acc[h].y; if (DUAL) a.partial_b[po + row[2 * h + 1]] = accb[h].y; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same sweep with ONE ROW PER THREAD and the candidates taken two at a time as the packed pair: the default kernel.
// Every row sum is a single sequential chain over all candidates in ascending order -- the reference thread's chain,
// acc = fma(e_l, w_l, acc) for l = 0, 1, 2, ... -- so results carry the reference's rounding order whatever the batch
// size, and (rows being independent) do not depend on how rows are spread over threads, CTAs or GPUs.
// Four times as many threads as the Q = 4 kernel for the same rows: 4 clouds of 16384 points already give every SM ~14
// warps, and the unrolled candidate loop keeps 8 independent exponentials per thread in flight.  FP32 cost per pair-pass
// is unchanged (7 packed issues + 2 scalar FFMA per two candidates against 8 packed issues per two rows).
// Shared memory holds candidate PAIRS as structure-of-arrays float4's: (x0, x1, y0, y1), (z0, z1, w0, w1).
// ---------------------------------------------------------------------------------------------------------------
constexpr int EMD_ROW_UNROLL = 4;   // candidate pairs per unrolled step
template <int MODE, bool UNIT, bool EXACT, int NT>
__global__ void __launch_bounds__(NT) emd_row_kernel(const __grid_constant__ SweepArgs a) {
    constexpr bool P3 = MODE == 3 || MODE == 4;
    constexpr bool DUAL = MODE == 4;
    __shared__ __align__(16) float4 sA[EMD_TC / 2];
    __shared__ __align__(16) float4 sBv[EMD_TC / 2];
    __shared__ __align__(8) float2 sWB[DUAL ? EMD_TC / 2 : 1];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.nrt;
    const int cloud = blockIdx.x / a.nrt;
    const int nr = a.nr, nc = a.nc;
    const float* __restrict__ cbase = a.cands + (size_t)cloud * nc * 3;
    const float* __restrict__ wbase = a.w + (size_t)cloud * nc;
    const float* __restrict__ wbbase = DUAL ? a.wb + (size_t)cloud * nc : nullptr;
    const int row = tile * NT + tid;
    const bool valid = row < nr;
    const size_t ridx = (size_t)cloud * nr + (valid ? row : 0);
    // the row NEGATED in both halves: (cand + (-row)) == cand - row exactly
    const float* rp = a.rows + ridx * 3;
    const float2 RX = make_float2(-rp[0], -rp[0]), RY = make_float2(-rp[1], -rp[1]), RZ = make_float2(-rp[2], -rp[2]);
    float rfs = 1.f;
    if (P3) rfs = a.rowfac[ridx];
    const float2 RF = make_float2(rfs, rfs);
    const float2 L2 = make_float2(a.lvl2, a.lvl2), L2B = make_float2(a.lvl2b, a.lvl2b);
    float acc = a.init0, accb = 1e-9f;   // pass 1 (also the fused one of the next level) starts at 1e-9 (tf_approxmatch.cu:36)

    for (int c0 = 0; c0 < nc; c0 += EMD_TC) {
        const int len = min(EMD_TC, nc - c0);
        __syncthreads();
        for (int i = tid; i < EMD_TC / 2; i += NT) {
            // padded candidates carry weight 0 and finite coordinates: fma(e, 0, acc) == acc
            float x0 = 0.f, y0 = 0.f, z0 = 0.f, w0 = 0.f, x1 = 0.f, y1 = 0.f, z1 = 0.f, w1 = 0.f, v0 = 0.f, v1 = 0.f;
            const int ca = 2 * i, cb = 2 * i + 1;
            if (ca < len) { const float* c = cbase + (size_t)(c0 + ca) * 3; x0 = c[0]; y0 = c[1]; z0 = c[2]; w0 = wbase[c0 + ca]; if (DUAL) v0 = wbbase[c0 + ca]; }
            if (cb < len) { const float* c = cbase + (size_t)(c0 + cb) * 3; x1 = c[0]; y1 = c[1]; z1 = c[2]; w1 = wbase[c0 + cb]; if (DUAL) v1 = wbbase[c0 + cb]; }
            sA[i] = make_float4(x0, x1, y0, y1);
            sBv[i] = make_float4(z0, z1, w0, w1);
            if (DUAL) sWB[i] = make_float2(v0, v1);
        }
        __syncthreads();
        const int npairs = (len + 1) / 2;
        const int np4 = (npairs + EMD_ROW_UNROLL - 1) / EMD_ROW_UNROLL * EMD_ROW_UNROLL;   // entries up to EMD_TC / 2 are padded with weight 0
        EMD_PRAGMA_UNROLL(EMD_ROW_UNROLL)
        for (int i = 0; i < np4; ++i) {
            const float4 A = sA[i], Bv = sBv[i];
            if (UNIT && !DUAL) {
                // e == 1: fma(1, w, acc) == acc + w and fma(rf * 1, w, acc) == fma(rf, w, acc), bit for bit
                if (P3) { acc = __fmaf_rn(rfs, Bv.z, acc); acc = __fmaf_rn(rfs, Bv.w, acc); }
                else { acc = __fadd_rn(acc, Bv.z); acc = __fadd_rn(acc, Bv.w); }
            } else {
                const float2 dx = __fadd2_rn(make_float2(A.x, A.y), RX);
                const float2 dy = __fadd2_rn(make_float2(A.z, A.w), RY);
                const float2 dz = __fadd2_rn(make_float2(Bv.x, Bv.y), RZ);
                const float2 d2 = sqdist3x2<true>(dx, dy, dz);
                const float2 x = __fmul2_rn(d2, L2);
                float2 e = make_float2(emd_ex2<EXACT>(x.x), emd_ex2<EXACT>(x.y));
                if (P3) e = __fmul2_rn(RF, e);
                acc = __fmaf_rn(e.x, Bv.z, acc);
                acc = __fmaf_rn(e.y, Bv.w, acc);
                if (DUAL) {
                    const float2 wb = sWB[i];
                    if (UNIT) {
                        accb = __fadd_rn(accb, wb.x);
                        accb = __fadd_rn(accb, wb.y);
                    } else {
                        const float2 xb = __fmul2_rn(d2, L2B);
                        accb = __fmaf_rn(emd_ex2<EXACT>(xb.x), wb.x, accb);
                        accb = __fmaf_rn(emd_ex2<EXACT>(xb.y), wb.y, accb);
                    }
                }
            }
        }
    }
    if (valid) emd_apply<MODE>(a.remain, a.ratio, a.fac, ridx, acc, accb);
}

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

// ---- epilogue: one thread per row; sum the split partials in fixed order, then the pass's update rule ------------
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
template <int MODE>
__global__ void emd_epi_kernel(int nr, int nsplit, size_t total, const float* __restrict__ partial, const float* __restrict__ partial_b,
                               float* __restrict__ remain, float* __restrict__ ratio, float* __restrict__ fac) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t cloud = t / nr;
    const int r = (int)(t % nr);
    const float sum = emd_sum_partials(partial, cloud, nsplit, nr, r);   // the seeds (1e-9) are inside split 0
    const float sumb = MODE == 4 ? emd_sum_partials(partial_b, cloud, nsplit, nr, r) : 0.f;
    emd_apply<MODE>(remain, ratio, fac, t, sum, sumb);
}

// ---------------------------------------------------------------------------------------------------------------
// Final pass over all pairs, from the stored per-level factors:
//     match[cloud, l, k] = sum_j ex2(lvl2_j * d2(k,l)) * facL_j[k] * facR_j[l]
// accumulated as the reference accumulates it, level by level: acc = fma(facL * e, facR, acc) (tf_approxmatch.cu:152, FFMA in
// the compiled kernel).  One kernel template serves every consumer of the matrix:
//   WRITE  store match (k contiguous)                                          -> rfnet_approxmatch
//   COST   cost[cloud] += sqrt(d2) * match   (tf_approxmatch.cu:183-225)       -> rfnet_emd_cost: the matrix never reaches HBM
//   GRAD   grad_own += match * (own - other) * rsqrt(max(d2, 1e-20))  (:229-291) -> rfnet_emd_cost_grad: no matrix in training
// ROLE 1: a thread owns two points k of xyz1 (a packed pair; their 10 factors facL in registers) and walks over the points l
//         of xyz2 (coordinates + 10 factors facR from shared memory, four LDS.128 per l);  GRAD gives grad1.
// ROLE 2: the same with the clouds swapped (own = xyz2, others = xyz1); GRAD gives grad2.  The product is still formed as
//         fma(facL * e, facR, acc), so both roles see bit-identical matrix entries.
// A CTA covers 256 own points x `ot_len` others (a function of the cloud size only); cost and gradient partials per CTA are
// summed in a fixed order afterwards: deterministic, batch-invariant, no atomics.
// sqrt(d2) * match is evaluated as d2 * rsqrt(max(d2, 1e-20)) * match: one MUFU.RSQ serves cost and gradient.
// DERIVED (default): e_{j+1} = exp(-4^(j+1) d2) = e_j^4, so only five of the nine exponentials go through the MUFU pipe
// (levels j = 6, 4, 2, 0, -1); the other four are two FMUL2 each.  Relative error of a derived factor: 4 x 2^-22 + 2 ulp
// ~ 1.1e-6, on non-negative terms, i.e. <= 1.1e-6 relative on every matrix entry, cost and gradient (the bar is 1e-4) --
// and the pass goes from MUFU-bound (18 MUFU / 30 packed FP32 per packed pair) to balanced (10 / 38).
// EXACT (RFNET_EMD_EXACT): nine non-ftz exponentials per pair, no skipping.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MT_THREADS = 128;
constexpr int MT_L = 64;
struct EmdLevels { float lvl2[EMD_LEVELS]; };
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
static int emd_pair_tile(int n_other) { return n_other <= 4096 ? 256 : 1024; }   // others per CTA (multiple of MT_L)

struct PairArgs {
    int n_own, n_oth, ot_len;
    size_t b_own, b_oth;                 // batch * points: stride between levels of the factor arrays
    const float *own, *oth;              // (b, n_own, 3), (b, n_oth, 3)
    const float *fac_own, *fac_oth;      // [EMD_LEVELS][b * n]
    float* match;                        // WRITE (ROLE 1 only): (b, n_oth, n_own)
    float* cost_partial;                 // COST: one float per CTA
    float* grad_partial;                 // GRAD: (b, gridDim.y, n_own, 3)
    EmdLevels lv;
};
template <int ROLE, bool WRITE, bool COST, bool GRAD, bool EXACT>
__global__ void __launch_bounds__(MT_THREADS) emd_pair_kernel(const __grid_constant__ PairArgs a) {
    static_assert(!WRITE || ROLE == 1, "the matrix is written k-contiguous: own = xyz1");
    __shared__ __align__(16) float4 sP[MT_L];       // x, y, z of the other point
    __shared__ __align__(16) float sF[MT_L][12];    // its factors, j = 0..9 (+2 pad)
    __shared__ float sW[MT_THREADS / 32];
    const int cloud = blockIdx.z;
    const int n_own = a.n_own, n_oth = a.n_oth;
    const int ka = blockIdx.x * (2 * MT_THREADS) + threadIdx.x, kb = ka + MT_THREADS;
    const bool va = ka < n_own, vb = kb < n_own;
    // own points negated (other + (-own) == other - own exactly); lanes past the end sit at infinity so they never block a level skip
    const float inf = __int_as_float(0x7f800000);
    float2 nx = make_float2(-inf, -inf), ny = make_float2(0.f, 0.f), nz = make_float2(0.f, 0.f);
    float2 fo[EMD_LEVELS];
    if (va) { const float* p = a.own + ((size_t)cloud * n_own + ka) * 3; nx.x = -p[0]; ny.x = -p[1]; nz.x = -p[2]; }
    if (vb) { const float* p = a.own + ((size_t)cloud * n_own + kb) * 3; nx.y = -p[0]; ny.y = -p[1]; nz.y = -p[2]; }
#pragma unroll
    for (int j = 0; j < EMD_LEVELS; ++j) {
        fo[j].x = va ? a.fac_own[(size_t)j * a.b_own + (size_t)cloud * n_own + ka] : 0.f;
        fo[j].y = vb ? a.fac_own[(size_t)j * a.b_own + (size_t)cloud * n_own + kb] : 0.f;
    }
    float2 csum = make_float2(0.f, 0.f), gx = csum, gy = csum, gz = csum;
    const int o_begin = blockIdx.y * a.ot_len;
    const int o_end = min(n_oth, o_begin + a.ot_len);
    for (int l0 = o_begin; l0 < o_end; l0 += MT_L) {
        const int nl = min(MT_L, o_end - l0);
        __syncthreads();
        for (int i = threadIdx.x; i < nl; i += MT_THREADS) {
            const float* q = a.oth + ((size_t)cloud * n_oth + l0 + i) * 3;
            sP[i] = make_float4(q[0], q[1], q[2], 0.f);
        }
        for (int i = threadIdx.x; i < nl * EMD_LEVELS; i += MT_THREADS) {
            const int l = i / EMD_LEVELS, j = i % EMD_LEVELS;
            sF[l][j] = a.fac_oth[(size_t)j * a.b_oth + (size_t)cloud * n_oth + l0 + l];
        }
        __syncthreads();
        float* __restrict__ out = WRITE ? a.match + ((size_t)cloud * n_oth + l0) * n_own : nullptr;
#pragma unroll 2
        for (int l = 0; l < nl; ++l) {
            const float4 q = sP[l];
            const float4 f0 = *reinterpret_cast<const float4*>(&sF[l][0]), f1 = *reinterpret_cast<const float4*>(&sF[l][4]), f2 = *reinterpret_cast<const float4*>(&sF[l][8]);
            const float fx[EMD_LEVELS] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y};
            const float2 dx = __fadd2_rn(nx, make_float2(q.x, q.x)), dy = __fadd2_rn(ny, make_float2(q.y, q.y)), dz = __fadd2_rn(nz, make_float2(q.z, q.z));
            const float2 d2 = sqdist3x2<true>(dx, dy, dz);
            float2 acc = make_float2(0.f, 0.f);
            // acc = fma(facL_j * e_j, facR_j, acc): facL belongs to xyz1 (own in ROLE 1, other in ROLE 2)
            auto term = [&](int j, float2 e) {
                const float2 o = make_float2(fx[j], fx[j]);
                acc = ROLE == 1 ? __ffma2_rn(__fmul2_rn(fo[j], e), o, acc) : __ffma2_rn(__fmul2_rn(o, e), fo[j], acc);
            };
            auto ex = [&](int j) {
                const float2 x = __fmul2_rn(d2, make_float2(a.lv.lvl2[j], a.lv.lvl2[j]));
                return make_float2(emd_ex2<EXACT>(x.x), emd_ex2<EXACT>(x.y));
            };
            if (EXACT) {
#pragma unroll
                for (int j = 0; j < EMD_LEVELS - 1; ++j) term(j, ex(j));
            } else {
                // levels 0 (j = 7) and 1 (j = 6): ex2.approx.ftz is exactly 0 below -126; when that holds at level 1 for the whole
                // warp both terms are fma(0, ., acc) == acc and are skipped (most pairs are farther than 0.15 apart)
                const float2 x1 = __fmul2_rn(d2, make_float2(a.lv.lvl2[1], a.lv.lvl2[1]));
                if (__any_sync(0xffffffffu, x1.x >= -126.0f || x1.y >= -126.0f)) {
                    const float2 e1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
                    const float2 s1 = __fmul2_rn(e1, e1);
                    term(0, __fmul2_rn(s1, s1));
                    term(1, e1);
                }
#pragma unroll
                for (int j = 3; j <= 7; j += 2) {
                    const float2 e = ex(j);
                    const float2 sq = __fmul2_rn(e, e);
                    term(j - 1, __fmul2_rn(sq, sq));
                    term(j, e);
                }
                term(8, ex(8));
            }
            {   // j = -2: level 0, e = 1
                const float2 o = make_float2(fx[EMD_LEVELS - 1], fx[EMD_LEVELS - 1]);
                acc = ROLE == 1 ? __ffma2_rn(fo[EMD_LEVELS - 1], o, acc) : __ffma2_rn(o, fo[EMD_LEVELS - 1], acc);
            }
            if (WRITE) {
                if (va) out[(size_t)l * n_own + ka] = acc.x;
                if (vb) out[(size_t)l * n_own + kb] = acc.y;
            }
            if (COST || GRAD) {
                // lanes past the end carry d2 = inf and acc = 0: rsqrt(inf) = 0 keeps them out of every sum (0 * inf never formed)
                const float2 s = __fmul2_rn(acc, make_float2(rsqrt_approx(fmaxf(d2.x, 1e-20f)), rsqrt_approx(fmaxf(d2.y, 1e-20f))));
                if (COST) csum = __ffma2_rn(s, make_float2(va ? d2.x : 0.f, vb ? d2.y : 0.f), csum);
                if (GRAD) {
                    // dx = other - own; the gradient wants (own - other) * s
                    const float2 ns = make_float2(-s.x, -s.y);
                    gx = __ffma2_rn(make_float2(va ? dx.x : 0.f, vb ? dx.y : 0.f), ns, gx);
                    gy = __ffma2_rn(dy, ns, gy);
                    gz = __ffma2_rn(dz, ns, gz);
                }
            }
        }
    }
    if (GRAD) {
        float* __restrict__ gp = a.grad_partial + (((size_t)cloud * gridDim.y + blockIdx.y) * n_own) * 3;
        if (va) { gp[(size_t)ka * 3 + 0] = gx.x; gp[(size_t)ka * 3 + 1] = gy.x; gp[(size_t)ka * 3 + 2] = gz.x; }
        if (vb) { gp[(size_t)kb * 3 + 0] = gx.y; gp[(size_t)kb * 3 + 1] = gy.y; gp[(size_t)kb * 3 + 2] = gz.y; }
    }
    if (COST) {
        float sum = warp_sum(csum.x + csum.y);
        if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < MT_THREADS / 32; ++i) t += sW[i];
            a.cost_partial[((size_t)cloud * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
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

// ---- sweep launchers -------------------------------------------------------------------------------------------
template <int MODE, int NT>
static void emd_launch_row(const SweepArgs& a, unsigned grid, bool unit, bool exact, cudaStream_t s) {
    if (exact) {
        if (unit) emd_row_kernel<MODE, true, true, NT><<<grid, NT, 0, s>>>(a);
        else emd_row_kernel<MODE, false, true, NT><<<grid, NT, 0, s>>>(a);
    } else {
        if (unit) emd_row_kernel<MODE, true, false, NT><<<grid, NT, 0, s>>>(a);
        else emd_row_kernel<MODE, false, false, NT><<<grid, NT, 0, s>>>(a);
    }
}
template <int Q, int MODE>
static void emd_launch_split(const SweepArgs& a, unsigned grid, bool unit, cudaStream_t s) {
    if (unit) emd_sweep_kernel<Q, MODE, true, false, false><<<grid, EMD_THREADS, 0, s>>>(a);
    else emd_sweep_kernel<Q, MODE, false, false, false><<<grid, EMD_THREADS, 0, s>>>(a);
}
// One sweep (+ its epilogue kernel when the candidates are split).  rows/cands: the two clouds in the roles of this pass.
// MODE 4: lvl2b / wb describe pass 1 of the next level.  mask != nullptr selects the pruned sweep (MODE 1..3, nr >= 512).
// Kernel choice never changes a result except through RFNET_EMD_SPLIT_SUMS (the only thing that reorders a sum):
//   default      one chain per row: emd_row_kernel (a thread per row), or the pruned sweep (a warp per 128-row cluster) at the sharp levels
//   SPLIT_SUMS   emd_sweep_kernel<Q> over candidate splits + emd_epi_kernel
template <int MODE>
static void emd_sweep(int b, int nr, int nc, int flags, float lvl2, float lvl2b, const float* rows, const float* cands, const float* w, const float* wb,
                      const float* rowfac, float* remain, float* ratio, float* fac, float* partial, const int* perm, const unsigned* mask,
                      cudaStream_t s) {
    SweepArgs a;
    a.nr = nr; a.nc = nc; a.nwords = (nc + 31) / 32;
    a.lvl2 = lvl2; a.lvl2b = lvl2b; a.init0 = MODE == 1 ? 1e-9f : 0.0f;
    a.rows = rows; a.cands = cands; a.w = w; a.wb = wb; a.rowfac = rowfac;
    a.partial = partial; a.partial_b = partial;
    a.remain = remain; a.ratio = ratio; a.fac = fac;
    a.perm = perm; a.mask = mask;
    const bool exact = (flags & RFNET_EMD_EXACT) != 0;
    const bool unit = MODE == 4 ? lvl2b == 0.0f : lvl2 == 0.0f;
    const int sms = num_sms();
    if (!(flags & RFNET_EMD_SPLIT_SUMS)) {
        a.nsplit = 1; a.split_len = (nc + 31) / 32 * 32;
        if (mask && MODE != 4) {
            // a warp per 128-row cluster; one-warp CTAs when four-warp CTAs would leave SMs idle
            const long clusters = (long)b * ((nr + EMD_CLUSTER - 1) / EMD_CLUSTER);
            constexpr int PM = MODE == 4 ? 1 : MODE;
            if (clusters >= 8L * sms) {
                a.nrt = (nr + 4 * EMD_CLUSTER - 1) / (4 * EMD_CLUSTER);
                emd_sweep_kernel<4, PM, false, true, false, 128><<<(unsigned)(b * a.nrt), 128, 0, s>>>(a);
            } else {
                a.nrt = (nr + EMD_CLUSTER - 1) / EMD_CLUSTER;
                emd_sweep_kernel<4, PM, false, true, false, 32><<<(unsigned)(b * a.nrt), 32, 0, s>>>(a);
            }
            return;
        }
        if ((long)b * nr >= 256L * sms) {
            a.nrt = (nr + 127) / 128;
            emd_launch_row<MODE, 128>(a, (unsigned)(b * a.nrt), unit, exact, s);
        } else {
            a.nrt = (nr + 63) / 64;
            emd_launch_row<MODE, 64>(a, (unsigned)(b * a.nrt), unit, exact, s);
        }
        return;
    }
    const SweepPlan p = emd_plan(nr, nc);
    a.nrt = p.nrt; a.nsplit = p.nsplit; a.split_len = p.split_len;
    a.partial_b = partial + (size_t)b * nr * p.nsplit;
    const unsigned grid = (unsigned)(b * p.nrt * p.nsplit);
    if (mask && p.Q == 4 && MODE != 4) {
        constexpr int PM = MODE == 4 ? 1 : MODE;
        emd_sweep_kernel<4, PM, false, true, false><<<grid, EMD_THREADS, 0, s>>>(a);
    } else if (p.Q == 4) {
        emd_launch_split<4, MODE>(a, grid, unit, s);
    } else {
        emd_launch_split<2, MODE>(a, grid, unit, s);
    }
    if (p.nsplit > 1) {
        const size_t total = (size_t)b * nr;
        emd_epi_kernel<MODE><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(nr, p.nsplit, total, a.partial, a.partial_b, remain, ratio, fac);
    }
}

// ---- final pass launcher ----------------------------------------------------------------------------------------
template <int ROLE, bool WRITE, bool COST, bool GRAD>
static void emd_launch_pair(const PairArgs& a, dim3 grid, bool exact, cudaStream_t s) {
    if (exact) emd_pair_kernel<ROLE, WRITE, COST, GRAD, true><<<grid, MT_THREADS, 0, s>>>(a);
    else emd_pair_kernel<ROLE, WRITE, COST, GRAD, false><<<grid, MT_THREADS, 0, s>>>(a);
}
static dim3 emd_pair_grid(int b, int n_own, int n_oth) {
    const int ot = emd_pair_tile(n_oth);
    return dim3((unsigned)((n_own + 2 * MT_THREADS - 1) / (2 * MT_THREADS)), (unsigned)((n_oth + ot - 1) / ot), (unsigned)b);
}
// floats of scratch behind the sweep workspace: cost partials of the ROLE-1 pass, then gradient partials of both roles
static size_t emd_cost_partials(int b, int n, int m) {
    const dim3 g = emd_pair_grid(b, n, m);
    return ((size_t)g.x * g.y * g.z + 3) & ~(size_t)3;
}
static size_t emd_grad_partials(int b, int n_own, int n_oth) {
    const dim3 g = emd_pair_grid(b, n_own, n_oth);
    return g.y > 1 ? (size_t)b * g.y * n_own * 3 : 0;   // a single tile of others writes the gradient directly
}
__global__ void emd_grad_reduce_kernel(int n, int nt, size_t bn, const float* __restrict__ partial, float* __restrict__ grad) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over bn*3
    if (t >= bn * 3) return;
    const size_t cloud = t / ((size_t)n * 3), r = t % ((size_t)n * 3);
    float s = 0.f;
    for (int i = 0; i < nt; ++i) s += partial[((size_t)cloud * nt + i) * n * 3 + r];
    grad[t] = s;
}

// sweeps (-> per-level factors in the workspace), then the final pass(es): matrix and/or cost and/or both gradients
static int emd_run(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* cost, float* grad1, float* grad2, float* ws_floats,
                   int flags, cudaStream_t s) {
    const size_t bn = (size_t)b * n, bm = (size_t)b * m;
    EmdWs ws = emd_carve(ws_floats, b, n, m);
    const bool exact = (flags & RFNET_EMD_EXACT) != 0;
    const float multiL = n >= m ? 1.0f : (float)(m / n), multiR = n >= m ? (float)(n / m) : 1.0f;  // integer division, tf_approxmatch.cu:4-10
    emd_init_kernel<<<(unsigned)((bn + bm + 255) / 256), 256, 0, s>>>(bn, bm, multiL, multiR, ws.remainL, ws.remainR);
    EmdLevels lv;
    for (int li = 0; li < EMD_LEVELS; ++li) lv.lvl2[li] = emd_level(li) * LOG2E;
    // exact pruning of the three sharpest levels (see emd_sweep_kernel<PRUNED>); relies on the ftz exponential's exact zeros
    const bool prune = emd_prune_enabled(n, m) && !(flags & (RFNET_EMD_NO_PRUNE | RFNET_EMD_EXACT));
    const int nwA = (m + 31) / 32, nwB = (n + 31) / 32;
    if (prune) {
        { const int rc = morton_sort(b, n, m, xyz1, xyz2, ws.perm1, ws.perm2, s); if (rc) return rc; }
        emd_mask_kernel<<<dim3((unsigned)((n + EMD_CLUSTER - 1) / EMD_CLUSTER), (unsigned)b), 256, 0, s>>>(n, m, nwA, xyz1, ws.perm1, xyz2, lv.lvl2[0],
                                                                                                       lv.lvl2[1], lv.lvl2[2], ws.maskA);
        emd_mask_kernel<<<dim3((unsigned)((m + EMD_CLUSTER - 1) / EMD_CLUSTER), (unsigned)b), 256, 0, s>>>(m, n, nwB, xyz2, ws.perm2, xyz1, lv.lvl2[0],
                                                                                                       lv.lvl2[1], lv.lvl2[2], ws.maskB);
    }
    auto maskA = [&](int li) -> const unsigned* { return prune && li < EMD_PRUNE_LEVELS ? ws.maskA + (size_t)li * nwA : nullptr; };
    auto maskB = [&](int li) -> const unsigned* { return prune && li < EMD_PRUNE_LEVELS ? ws.maskB + (size_t)li * nwB : nullptr; };
    // pass 1 of the first level: rows = xyz1 (k), candidates = xyz2 (l) weighted by remainR            -> ratioL, facL[0]
    emd_sweep<1>(b, n, m, flags, lv.lvl2[0], 0.f, xyz1, xyz2, ws.remainR, nullptr, nullptr, ws.remainL, ws.ratioL, ws.facL, ws.partial, ws.perm1, maskA(0), s);
    for (int li = 0; li < EMD_LEVELS; ++li) {
        const float lvl2 = lv.lvl2[li];
        // pass 2: rows = xyz2 (l), candidates = xyz1 (k) weighted by ratioL                           -> ratioR, facR[li], remainR
        emd_sweep<2>(b, m, n, flags, lvl2, 0.f, xyz2, xyz1, ws.ratioL, nullptr, nullptr, ws.remainR, ws.ratioR, ws.facR + (size_t)li * bm, ws.partial,
                     ws.perm2, maskB(li), s);
        // pass 3: rows = xyz1 (k), candidates = xyz2 (l) weighted by ratioR, row factor ratioL         -> remainL
        // fused with pass 1 of level li + 1 (weights remainR, already final)                          -> ratioL, facL[li + 1]
        // unless one of the two runs as a pruned sweep (each has its own, tighter candidate mask)
        const bool last = li == EMD_LEVELS - 1;
        if (!last && !maskA(li) && !maskA(li + 1)) {
            emd_sweep<4>(b, n, m, flags, lvl2, lv.lvl2[li + 1], xyz1, xyz2, ws.ratioR, ws.remainR, ws.ratioL, ws.remainL, ws.ratioL,
                         ws.facL + (size_t)(li + 1) * bn, ws.partial, nullptr, nullptr, s);
        } else {
            emd_sweep<3>(b, n, m, flags, lvl2, 0.f, xyz1, xyz2, ws.ratioR, nullptr, ws.ratioL, ws.remainL, nullptr, nullptr, ws.partial, ws.perm1, maskA(li), s);
            if (!last)
                emd_sweep<1>(b, n, m, flags, lv.lvl2[li + 1], 0.f, xyz1, xyz2, ws.remainR, nullptr, nullptr, ws.remainL, ws.ratioL,
                             ws.facL + (size_t)(li + 1) * bn, ws.partial, ws.perm1, maskA(li + 1), s);
        }
    }
    // final pass(es)
    float* extra = ws_floats + emd_ws_floats(b, n, m);   // carved by the *_workspace_bytes of the entry points that need it
    PairArgs pa;
    pa.n_own = n; pa.n_oth = m; pa.ot_len = emd_pair_tile(m); pa.b_own = bn; pa.b_oth = bm;
    pa.own = xyz1; pa.oth = xyz2; pa.fac_own = ws.facL; pa.fac_oth = ws.facR;
    pa.match = match; pa.cost_partial = extra; pa.grad_partial = nullptr; pa.lv = lv;
    const dim3 g1 = emd_pair_grid(b, n, m);
    RFNET_CHECK_ARG(g1.y <= 65535 && g1.z <= 65535);
    if (grad1) {
        float* gp1 = extra + emd_cost_partials(b, n, m);
        pa.grad_partial = g1.y > 1 ? gp1 : grad1;
        emd_launch_pair<1, false, true, true>(pa, g1, exact, s);
        if (g1.y > 1) emd_grad_reduce_kernel<<<(unsigned)((bn * 3 + 255) / 256), 256, 0, s>>>(n, (int)g1.y, bn, gp1, grad1);
        PairArgs pb = pa;
        pb.n_own = m; pb.n_oth = n; pb.ot_len = emd_pair_tile(n); pb.b_own = bm; pb.b_oth = bn;
        pb.own = xyz2; pb.oth = xyz1; pb.fac_own = ws.facR; pb.fac_oth = ws.facL; pb.match = nullptr; pb.cost_partial = nullptr;
        const dim3 g2 = emd_pair_grid(b, m, n);
        RFNET_CHECK_ARG(g2.y <= 65535);
        float* gp2 = gp1 + emd_grad_partials(b, n, m);
        pb.grad_partial = g2.y > 1 ? gp2 : grad2;
        emd_launch_pair<2, false, false, true>(pb, g2, exact, s);
        if (g2.y > 1) emd_grad_reduce_kernel<<<(unsigned)((bm * 3 + 255) / 256), 256, 0, s>>>(m, (int)g2.y, bm, gp2, grad2);
    } else if (match && cost) {
        emd_launch_pair<1, true, true, false>(pa, g1, exact, s);
    } else if (cost) {
        emd_launch_pair<1, false, true, false>(pa, g1, exact, s);
    } else {
        emd_launch_pair<1, true, false, false>(pa, g1, exact, s);
    }
    if (cost) reduce_partials_kernel<<<b, 256, 0, s>>>((int)(g1.x * g1.y), extra, cost);
    return launch_status();
}

}  // namespace rfnet

using namespace rfnet;

// EXACT excludes the split sums (a different rounding order) -- asking for both is a caller error
static bool emd_flags_ok(int flags) {
    if (flags & ~(RFNET_EMD_EXACT | RFNET_EMD_NO_PRUNE | RFNET_EMD_SPLIT_SUMS)) return false;
    return !((flags & RFNET_EMD_EXACT) && (flags & RFNET_EMD_SPLIT_SUMS));
}

extern "C" size_t rfnet_approxmatch_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return emd_ws_floats(b, n, m) * sizeof(float);
}

extern "C" int rfnet_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, void* workspace,
                                 size_t workspace_bytes, int flags, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0 || n == 0 || m == 0) return 0;
    RFNET_CHECK_ARG(xyz1 && xyz2 && match && workspace && workspace_bytes >= rfnet_approxmatch_workspace_bytes(b, n, m));
    RFNET_CHECK_ARG(b <= 65535 && emd_flags_ok(flags));
    return emd_run(b, n, m, xyz1, xyz2, match, nullptr, nullptr, nullptr, (float*)workspace, flags, (cudaStream_t)stream);
}

extern "C" size_t rfnet_emd_cost_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return (emd_ws_floats(b, n, m) + emd_cost_partials(b, n, m)) * sizeof(float);
}

extern "C" int rfnet_emd_cost(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* cost, void* workspace,
                              size_t workspace_bytes, int flags, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0) return 0;
    RFNET_CHECK_ARG(cost);
    if (n == 0 || m == 0) {
        RFNET_CUDA(cudaMemsetAsync(cost, 0, sizeof(float) * b, (cudaStream_t)stream));
        return 0;
    }
    RFNET_CHECK_ARG(xyz1 && xyz2 && workspace && workspace_bytes >= rfnet_emd_cost_workspace_bytes(b, n, m));
    RFNET_CHECK_ARG(b <= 65535 && emd_flags_ok(flags));
    return emd_run(b, n, m, xyz1, xyz2, match, cost, nullptr, nullptr, (float*)workspace, flags, (cudaStream_t)stream);
}

extern "C" size_t rfnet_emd_cost_grad_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return (emd_ws_floats(b, n, m) + emd_cost_partials(b, n, m) + emd_grad_partials(b, n, m) + emd_grad_partials(b, m, n)) * sizeof(float);
}

extern "C" int rfnet_emd_cost_grad(int b, int n, int m, const float* xyz1, const float* xyz2, float* cost, float* grad1, float* grad2,
                                   void* workspace, size_t workspace_bytes, int flags, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0) return 0;
    RFNET_CHECK_ARG(cost);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        RFNET_CUDA(cudaMemsetAsync(cost, 0, sizeof(float) * b, s));
        if (n) { RFNET_CHECK_ARG(grad1); RFNET_CUDA(cudaMemsetAsync(grad1, 0, sizeof(float) * 3 * (size_t)b * n, s)); }
        if (m) { RFNET_CHECK_ARG(grad2); RFNET_CUDA(cudaMemsetAsync(grad2, 0, sizeof(float) * 3 * (size_t)b * m, s)); }
        return 0;
    }
    RFNET_CHECK_ARG(xyz1 && xyz2 && grad1 && grad2 && workspace && workspace_bytes >= rfnet_emd_cost_grad_workspace_bytes(b, n, m));
    RFNET_CHECK_ARG(b <= 65535 && emd_flags_ok(flags));
    return emd_run(b, n, m, xyz1, xyz2, nullptr, cost, grad1, grad2, (float*)workspace, flags, s);
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
