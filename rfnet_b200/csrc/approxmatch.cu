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
// Split plan (RFNET_EMD_SPLIT_SUMS only): a function of the cloud sizes, NOT of the batch size -- see the header.
// B200 tuning (profiles/r2_emd_tune.txt): Q = 4 rows per thread with the candidate range cut into many short splits is the
// best or within 1 % of the best grid at every batch size from 1 to 32 clouds, so no batch-dependent choice is needed.
// ---------------------------------------------------------------------------------------------------------------
struct SweepPlan { int Q, nrt, nsplit, split_len; };
static int emd_split_len(int nc) {
    int sl = nc <= EMD_SPLIT_SMALL_MAX ? EMD_SPLIT_SMALL : EMD_SPLIT_LARGE;
    if ((nc + sl - 1) / sl > EMD_MAX_SPLITS) sl = (((nc + EMD_MAX_SPLITS - 1) / EMD_MAX_SPLITS) + 31) / 32 * 32;
    return sl;
}
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
//   partial                                                                    sweep partial sums (split sums only; two sets for the fused sweep)
//   pairA1/pairZ1, pairA2/pairZ2                                               the clouds as candidate PAIRS (see emd_row_kernel)
//   perm1, perm2, maskA, maskB                                                 Morton orders + candidate masks of the pruned levels
constexpr int EMD_PRUNE_LEVELS = 3;          // levels the masks are built for (layout of the mask words)
constexpr int EMD_PRUNED_SWEEP_LEVELS = 2;   // levels that run as pruned sweeps
constexpr int EMD_CLUSTER = 32;         // rows of one warp of the row kernel: one per lane
struct EmdWs {
    float *remainL, *remainR, *ratioL, *ratioR, *facL, *facR, *partial;
    float4 *pairA1, *pairA2;     // (x0, x1, y0, y1) of candidates 2i, 2i+1
    float2 *pairZ1, *pairZ2;     // (z0, z1)
    int *perm1, *perm2;          // Morton order of xyz1 / xyz2 (pruned sweeps only)
    unsigned *maskA, *maskB;     // candidate masks: rows = xyz1 clusters vs xyz2 candidates (passes 1, 3) / the reverse (pass 2)
};
// exact pruning pays off from about a thousand points per cloud; the Morton sort handles up to 32768
static bool emd_prune_enabled(int n, int m) { return n >= 1024 && m >= 1024 && n <= MORTON_SORT_MAX && m <= MORTON_SORT_MAX; }   // (2048 by default, see emd_run)
static size_t emd_mask_words(int b, int nr, int nc) { return (size_t)b * ((nr + EMD_CLUSTER - 1) / EMD_CLUSTER) * EMD_PRUNE_LEVELS * ((nc + 31) / 32); }
static int emd_npad(int n) { return (((n + 1) / 2) + 1) & ~1; }   // candidate pairs per cloud, padded to an even count (16-byte rows of float2)
static size_t emd_partial_floats(int b, int n, int m) {
    const size_t rows1 = 2 * (size_t)n * emd_plan(n, m).nsplit;   // rows = xyz1: the fused sweep keeps two sums per row
    const size_t rows2 = (size_t)m * emd_plan(m, n).nsplit;
    return ((size_t)b * (rows1 > rows2 ? rows1 : rows2) + 3) & ~(size_t)3;
}
static size_t emd_ws_floats(int b, int n, int m) {
    const size_t bn = (((size_t)b * n) + 3) & ~(size_t)3, bm = (((size_t)b * m) + 3) & ~(size_t)3;
    size_t f = 2 * (bn + bm) + (size_t)EMD_LEVELS * (bn + bm) + emd_partial_floats(b, n, m);
    f += 6 * (size_t)b * (emd_npad(n) + emd_npad(m));
    if (emd_prune_enabled(n, m)) f += bn + bm + emd_mask_words(b, n, m) + emd_mask_words(b, m, n);
    return (f + 3) & ~(size_t)3;
}
static EmdWs emd_carve(float* w, int b, int n, int m) {
    // every block starts on a 16-byte boundary (the workspace itself must be 16-byte aligned: cudaMalloc / torch give 256+)
    const size_t bn = (((size_t)b * n) + 3) & ~(size_t)3, bm = (((size_t)b * m) + 3) & ~(size_t)3;
    EmdWs s;
    s.remainL = w; w += bn;
    s.remainR = w; w += bm;
    s.ratioL = w; w += bn;
    s.ratioR = w; w += bm;
    s.facL = w; w += EMD_LEVELS * bn;
    s.facR = w; w += EMD_LEVELS * bm;
    s.partial = w; w += emd_partial_floats(b, n, m);
    s.pairA1 = reinterpret_cast<float4*>(w); w += 4 * (size_t)b * emd_npad(n);
    s.pairA2 = reinterpret_cast<float4*>(w); w += 4 * (size_t)b * emd_npad(m);
    s.pairZ1 = reinterpret_cast<float2*>(w); w += 2 * (size_t)b * emd_npad(n);
    s.pairZ2 = reinterpret_cast<float2*>(w); w += 2 * (size_t)b * emd_npad(m);
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
//   rows: (b, nr, 3), cands: (b, nc, 3), w: (b, nc).
// The accumulation is the reference binary's, term by term and in candidate order:
//   MODE 1 (pass 1, tf_approxmatch.cu:26-59):   acc = fma(e, w[c], acc), acc starts at 1e-9     -> ratioL = remainL / acc
//   MODE 2 (pass 2, :75-108):                   acc = fma(e, w[c], acc)                          -> consumption, ratioR, remainR
//   MODE 3 (pass 3, :127-160):                  acc = fma(rowfac[row] * e, w[c], acc)            -> remainL = max(0, remainL - acc)
//   MODE 4 = MODE 3 at this level fused with MODE 1 at the NEXT level (lvl2b, weights wb = remainR): one distance per
//            pair, two exponentials, two chains; the rule of pass 3 is applied first, then pass 1's on the updated remainL.
// UNIT: the last level (j = -2) has level = 0, i.e. e = ex2(0 * d2) = 1 for every pair; the same chain is run without
// distances or exponentials.  In MODE 4 UNIT refers to the next level (the fused pair j = -1, -2).
// ---------------------------------------------------------------------------------------------------------------
struct SweepArgs {
    int nr, nc, nrt, nsplit, split_len, nwords, npad, tma;
    float lvl2, lvl2b, init0;
    const float *rows, *cands, *w, *wb, *rowfac;
    const float4* pairA;           // row kernel: the candidates as pairs, (b, npad)
    const float2* pairZ;
    float *partial, *partial_b;    // split sums
    float *remain, *ratio, *fac;   // MODE 1, 3, 4: remainL / ratioL / facL[level (MODE 4: next level)]   MODE 2: remainR / ratioR / facR[level]
    const int* perm;               // pruned: Morton order of the rows
    const unsigned* mask;          // pruned: the level's candidate-mask words of cluster 0 of cloud 0
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
// emd_row_kernel -- the default sweep: ONE ROW PER THREAD, candidates taken two at a time as the packed pair.
// Every row sum is a single sequential chain over all candidates in ascending order -- the reference thread's chain,
// acc = fma(e_l, w_l, acc) for l = 0, 1, 2, ... -- so results carry the reference's rounding order whatever the batch
// size, and (rows being independent) do not depend on how rows are spread over threads, CTAs or GPUs.
// FP32 cost per pair-pass equals the row-packed kernel's (7 packed issues + 2 scalar FFMA per two candidates against 8
// packed issues per two rows), with four times as many threads for the same rows: 4 clouds of 16384 points already give
// every SM 14 warps.
// Candidates come pre-paired from emd_pairs_kernel -- (x0, x1, y0, y1), (z0, z1) per pair, written once per call -- and are
// staged chunk by chunk in shared memory by the TMA bulk-copy engine (cp.async.bulk + mbarrier, double buffered) together
// with the weight pairs: no thread spends an instruction on staging and the next chunk lands while this one is consumed.
// (Clouds whose size is not a multiple of 4 take plain loads: the weight rows are then not 16-byte aligned.)
//
// PRUNED: exact pruning of the three sharpest levels (j = 7, 6, 5: e = exp(-4^j d2) with 4^j = 16384, 4096, 1024).
// ex2.approx.ftz returns EXACTLY 0 once its argument is below -126, i.e. for d2 > 0.0053 / 0.021 / 0.085, and a zero term
// leaves the accumulator bit-identical (fma(0, w, acc) == acc).  At those levels almost every pair is such a no-op, so:
//   * morton_sort_kernel orders each cloud along a Morton curve (one CTA per cloud, counting sort in shared memory);
//     a warp then owns 32 CONSECUTIVE points of that order: a spatially tight cluster;
//   * emd_mask_kernel marks, per cluster and level, the candidates whose distance to the cluster's bounding box still
//     allows a non-zero term;
//   * the warp visits only candidate pairs with a marked member, in ASCENDING order: every row sum goes through the same
//     sequence of non-trivial fma's as in the dense sweep, so the result is bit-for-bit the dense one
//     (tests/test_emd_gpu.py::test_pruned_sweeps_are_exact) -- only the row -> thread assignment changes.
// The masks depend on the points only: built once per call for both roles (rows = xyz1 / rows = xyz2), used by 9 sweeps.
//
// Margin.  A candidate is dropped when lvl2 * dbox2 <= EMD_PRUNE_ARG = -130, where dbox2 is the squared distance from the
// candidate to the cluster's bounding box, evaluated in float32 with fma's.  For every row p of the cluster the true
// d2(p, c) >= dbox2_exact, and the two float evaluations (3 subtractions, 3 products, 2 sums each, every operand < 4 in
// magnitude) differ from their exact values by less than 8 ulp relative, i.e. |lvl2*d2 - lvl2*dbox2| < 130 * 8 * 2^-24 <
// 1e-4 in the argument of the exponential.  The sweep's own argument for that pair is therefore below -130 + 1e-4 < -126,
// where ex2.approx.ftz is exactly 0: a margin of 4 units covers the rounding by four orders of magnitude.
// ---------------------------------------------------------------------------------------------------------------
constexpr float EMD_PRUNE_ARG = -130.0f;
constexpr int ROW_CP = 128;         // candidate pairs per staged chunk
#ifndef EMD_ROW_UNROLL_VALUE
#define EMD_ROW_UNROLL_VALUE 8
#endif
constexpr int EMD_ROW_UNROLL = EMD_ROW_UNROLL_VALUE;   // candidate pairs per unrolled step (tools/emd_tune.cu)

template <int MODE, bool UNIT, bool EXACT, int NT, int R = 1>
__global__ void __launch_bounds__(NT) emd_row_kernel(const __grid_constant__ SweepArgs a) {
    // (the dependents are released at the END of this kernel: CTAs of the next sweep that become resident while this one runs are
    // placed on whatever SMs have room, and a sweep's CTAs live as long as the sweep -- measured 2.5x slower at 1 cloud of 16384 points)
    pdl_wait();   // everything this kernel reads was written by the kernels before it in the stream
    constexpr bool P3 = MODE == 3 || MODE == 4;
    constexpr bool DUAL = MODE == 4;
    constexpr bool COORDS = !UNIT || DUAL;      // a unit-level single sweep needs the weights only
    __shared__ __align__(128) float4 sA[2][ROW_CP];
    __shared__ __align__(16) float2 sZ[2][ROW_CP];
    __shared__ __align__(16) float2 sW[2][ROW_CP];
    __shared__ __align__(16) float2 sV[DUAL ? 2 : 1][DUAL ? ROW_CP : 2];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.nrt;
    const int cloud = blockIdx.x / a.nrt;
    const int nr = a.nr, nc = a.nc;
    const float* __restrict__ wbase = a.w + (size_t)cloud * nc;
    const float* __restrict__ vbase = DUAL ? a.wb + (size_t)cloud * nc : nullptr;
    const float4* __restrict__ Abase = a.pairA + (size_t)cloud * a.npad;
    const float2* __restrict__ Zbase = a.pairZ + (size_t)cloud * a.npad;
    // R rows per thread, each with its own chain(s): the candidate loads of a step are shared by the R rows (R = 2 halves the
    // shared-memory instructions per pair-pass: LDS and MUFU share the MIO queue, the top stall of the R = 1 kernel at small batches)
    bool valid[R];
    size_t ridx[R];
    float2 RX[R], RY[R], RZ[R], RF[R];
    float rfs[R], acc[R], accb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int pos = tile * (NT * R) + r * NT + tid;
        valid[r] = pos < nr;
        ridx[r] = (size_t)cloud * nr + (valid[r] ? pos : 0);
        // the row NEGATED in both halves: (cand + (-row)) == cand - row exactly
        const float* rp = a.rows + ridx[r] * 3;
        RX[r] = make_float2(-rp[0], -rp[0]); RY[r] = make_float2(-rp[1], -rp[1]); RZ[r] = make_float2(-rp[2], -rp[2]);
        rfs[r] = P3 ? a.rowfac[ridx[r]] : 1.f;
        RF[r] = make_float2(rfs[r], rfs[r]);
        acc[r] = a.init0;
        accb[r] = 1e-9f;   // pass 1 (also the fused one of the next level) starts at 1e-9 (tf_approxmatch.cu:36)
    }
    const float2 L2 = make_float2(a.lvl2, a.lvl2), L2B = make_float2(a.lvl2b, a.lvl2b);

    const int npairs = (nc + 1) / 2;
    const int nchunks = (npairs + ROW_CP - 1) / ROW_CP;
    const bool tma = a.tma != 0;
    if (tma) {
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    auto issue = [&](int ci) {   // thread 0 only: start the bulk copies of chunk ci (nc % 4 == 0: every size is a multiple of 16 bytes)
        const int p0 = ci * ROW_CP;
        const unsigned cp = (unsigned)min(ROW_CP, npairs - p0);
        const int bf = ci & 1;
        mbar_expect_tx(&bar[bf], cp * ((COORDS ? 24u : 0u) + 8u + (DUAL ? 8u : 0u)));
        if (COORDS) {
            tma_bulk_g2s(sA[bf], Abase + p0, cp * 16u, &bar[bf]);
            tma_bulk_g2s(sZ[bf], Zbase + p0, cp * 8u, &bar[bf]);
        }
        tma_bulk_g2s(sW[bf], wbase + 2 * p0, cp * 8u, &bar[bf]);
        if (DUAL) tma_bulk_g2s(sV[bf], vbase + 2 * p0, cp * 8u, &bar[bf]);
    };
    if (tma && tid == 0) issue(0);

    for (int ci = 0; ci < nchunks; ++ci) {
        const int p0 = ci * ROW_CP;
        const int cp = min(ROW_CP, npairs - p0);
        const int bf = tma ? (ci & 1) : 0;
        if (tma) {
            if (tid == 0 && ci + 1 < nchunks) issue(ci + 1);   // the other buffer was released by the barrier that ended chunk ci - 1
            mbar_wait(&bar[bf], (ci >> 1) & 1);
        } else {
            for (int i = tid; i < cp; i += NT) {
                const int ca = 2 * (p0 + i), cb = ca + 1;
                if (COORDS) { sA[0][i] = Abase[p0 + i]; sZ[0][i] = Zbase[p0 + i]; }
                // a phantom second member of the last pair (odd nc) carries weight 0: fma(e, 0, acc) == acc
                sW[0][i] = make_float2(wbase[ca], cb < nc ? wbase[cb] : 0.f);
                if (DUAL) sV[0][i] = make_float2(vbase[ca], cb < nc ? vbase[cb] : 0.f);
            }
            __syncthreads();
        }
        // (A hand-made software pipeline -- exponentials of group g + 1 issued before the fma chain of group g -- was measured
        // 25 % SLOWER than letting ptxas schedule the unrolled loop: the register copies and clamps cost more than the stalls.)
        auto step = [&](int i) {
            const float2 w = sW[bf][i];
            if (UNIT && !DUAL) {
                // e == 1: fma(1, w, acc) == acc + w and fma(rf * 1, w, acc) == fma(rf, w, acc), bit for bit
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (P3) { acc[r] = __fmaf_rn(rfs[r], w.x, acc[r]); acc[r] = __fmaf_rn(rfs[r], w.y, acc[r]); }
                    else { acc[r] = __fadd_rn(acc[r], w.x); acc[r] = __fadd_rn(acc[r], w.y); }
                }
            } else {
                const float4 A = sA[bf][i];
                const float2 Z = sZ[bf][i];
                float2 v = make_float2(0.f, 0.f);
                if (DUAL) v = sV[bf][i];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float2 dx = __fadd2_rn(make_float2(A.x, A.y), RX[r]);
                    const float2 dy = __fadd2_rn(make_float2(A.z, A.w), RY[r]);
                    const float2 dz = __fadd2_rn(Z, RZ[r]);
                    const float2 d2 = sqdist3x2<true>(dx, dy, dz);
                    const float2 x = __fmul2_rn(d2, L2);
                    float2 e = make_float2(emd_ex2<EXACT>(x.x), emd_ex2<EXACT>(x.y));
                    if (P3) e = __fmul2_rn(RF[r], e);
                    acc[r] = __fmaf_rn(e.x, w.x, acc[r]);
                    acc[r] = __fmaf_rn(e.y, w.y, acc[r]);
                    if (DUAL) {
                        if (UNIT) {
                            accb[r] = __fadd_rn(accb[r], v.x);
                            accb[r] = __fadd_rn(accb[r], v.y);
                        } else {
                            const float2 xb = __fmul2_rn(d2, L2B);
                            accb[r] = __fmaf_rn(emd_ex2<EXACT>(xb.x), v.x, accb[r]);
                            accb[r] = __fmaf_rn(emd_ex2<EXACT>(xb.y), v.y, accb[r]);
                        }
                    }
                }
            }
        };
        {
            EMD_PRAGMA_UNROLL(EMD_ROW_UNROLL)
            for (int i = 0; i < cp; ++i) step(i);
        }
        __syncthreads();   // everyone is done with this buffer before it is refilled
    }
    pdl_launch_dependents();
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (valid[r]) emd_apply<MODE>(a.remain, a.ratio, a.fac, ridx[r], acc[r], accb[r]);
}

// ---------------------------------------------------------------------------------------------------------------
// emd_pruned_kernel -- the same chain restricted to the candidate pairs that can contribute at a sharp level (RFNET_EMD_PRUNE).
// A warp owns one 32-row Morton cluster.  It walks the cluster's candidate mask 32 words at a time, COMPACTS the indices of
// the pairs with a marked member into a per-warp list (ascending), and every 32 listed pairs: each lane gathers one pair
// (coordinates + weights, three small loads, all in flight together) into a per-warp shared-memory tile, then the warp runs
// the dense, unrolled chain loop over the tile.  Work and traffic are proportional to the surviving pairs, the inner loop has
// the dense kernel's instruction-level parallelism, and there is no block-wide barrier.  (The first pruned variant staged
// every candidate and visited marked pairs one by one: it took as long as a dense sweep -- profiles/r2_emd_launches_*.csv.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int PG_WARPS = 4;
template <int MODE>
__global__ void __launch_bounds__(PG_WARPS * 32) emd_pruned_kernel(const __grid_constant__ SweepArgs a) {
    // (the dependents are released at the END of this kernel: CTAs of the next sweep that become resident while this one runs are
    // placed on whatever SMs have room, and a sweep's CTAs live as long as the sweep -- measured 2.5x slower at 1 cloud of 16384 points)
    pdl_wait();   // everything this kernel reads was written by the kernels before it in the stream
    constexpr bool P3 = MODE == 3;
    static_assert(MODE >= 1 && MODE <= 3, "one sharp level per pruned sweep");
    __shared__ int sList[PG_WARPS][64];
    __shared__ __align__(16) float4 sA[PG_WARPS][32];
    __shared__ __align__(8) float2 sZ[PG_WARPS][32];
    __shared__ __align__(8) float2 sW[PG_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nr = a.nr, nc = a.nc;
    const int nclusters = (nr + EMD_CLUSTER - 1) / EMD_CLUSTER;
    const int cpc = (nclusters + PG_WARPS - 1) / PG_WARPS;        // CTAs per cloud
    const int cloud = blockIdx.x / cpc;
    const int cluster = (blockIdx.x % cpc) * PG_WARPS + warp;
    if (cluster >= nclusters) return;                              // warp-uniform; no block barrier below
    const int pos = cluster * EMD_CLUSTER + lane;
    const bool valid = pos < nr;
    const int row = a.perm[(size_t)cloud * nr + (valid ? pos : cluster * EMD_CLUSTER)];
    const size_t ridx = (size_t)cloud * nr + row;
    const float* rp = a.rows + ridx * 3;
    const float2 RX = make_float2(-rp[0], -rp[0]), RY = make_float2(-rp[1], -rp[1]), RZ = make_float2(-rp[2], -rp[2]);
    float rfs = 1.f;
    if (P3) rfs = a.rowfac[ridx];
    const float2 RF = make_float2(rfs, rfs);
    const float2 L2 = make_float2(a.lvl2, a.lvl2);
    float acc = a.init0;
    const float* __restrict__ wbase = a.w + (size_t)cloud * nc;
    const float4* __restrict__ Abase = a.pairA + (size_t)cloud * a.npad;
    const float2* __restrict__ Zbase = a.pairZ + (size_t)cloud * a.npad;
    const unsigned* __restrict__ mrow = a.mask + ((size_t)cloud * nclusters + cluster) * EMD_PRUNE_LEVELS * a.nwords;
    int* list = sList[warp];

    auto flush = [&](int cnt) {   // gather the first cnt (<= 32) listed pairs, then chain them in order
        if (lane < cnt) {
            const int i = list[lane];
            sA[warp][lane] = __ldg(Abase + i);
            sZ[warp][lane] = __ldg(Zbase + i);
            const int ca = 2 * i, cb = ca + 1;
            sW[warp][lane] = make_float2(__ldg(wbase + ca), cb < nc ? __ldg(wbase + cb) : 0.f);   // a phantom member carries weight 0
        }
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const float4 A = sA[warp][k];
            const float2 Z = sZ[warp][k], w = sW[warp][k];
            const float2 dx = __fadd2_rn(make_float2(A.x, A.y), RX);
            const float2 dy = __fadd2_rn(make_float2(A.z, A.w), RY);
            const float2 dz = __fadd2_rn(Z, RZ);
            const float2 x = __fmul2_rn(sqdist3x2<true>(dx, dy, dz), L2);
            float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));   // the unmarked member of a listed pair gives an exact 0
            if (P3) e = __fmul2_rn(RF, e);
            acc = __fmaf_rn(e.x, w.x, acc);
            acc = __fmaf_rn(e.y, w.y, acc);
        }
        __syncwarp();
    };

    int count = 0;   // listed, not yet processed (warp-uniform, < 32 between words)
    for (int w0 = 0; w0 < a.nwords; w0 += 32) {
        const unsigned myword = (w0 + lane < a.nwords) ? mrow[w0 + lane] : 0u;
        unsigned nonzero = __ballot_sync(0xffffffffu, myword != 0u);
        while (nonzero) {
            const int wl = __ffs(nonzero) - 1;
            nonzero &= nonzero - 1;
            const unsigned bits = __shfl_sync(0xffffffffu, myword, wl);
            const unsigned pm = (bits | (bits >> 1)) & 0x55555555u;      // bit 2j set: pair j of this word has a marked member
            if (lane < 16 && ((pm >> (2 * lane)) & 1u))
                list[count + __popc(pm & ((1u << (2 * lane)) - 1u))] = (w0 + wl) * 16 + lane;
            count += __popc(pm);
            __syncwarp();
            if (count >= 32) {
                flush(32);
                const int rest = count - 32;                              // <= 15
                const int v = lane < rest ? list[32 + lane] : 0;
                __syncwarp();
                if (lane < rest) list[lane] = v;
                __syncwarp();
                count = rest;
            }
        }
    }
    if (count > 0) flush(count);
    pdl_launch_dependents();
    if (valid) emd_apply<MODE>(a.remain, a.ratio, a.fac, ridx, acc, 0.f);
}

// The clouds as candidate pairs, written once per call: A[i] = (x0, x1, y0, y1), Z[i] = (z0, z1) of candidates 2i, 2i+1;
// entries past the cloud's end are zero (finite coordinates; their weights are zero or never read).
__global__ void emd_pairs_kernel(int n, int npad, const float* __restrict__ xyz, float4* __restrict__ A, float2* __restrict__ Z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const size_t cloud = blockIdx.y;
    const float* __restrict__ p = xyz + cloud * (size_t)n * 3;
    const int ca = 2 * i, cb = 2 * i + 1;
    float x0 = 0.f, y0 = 0.f, z0 = 0.f, x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (ca < n) { x0 = p[ca * 3]; y0 = p[ca * 3 + 1]; z0 = p[ca * 3 + 2]; }
    if (cb < n) { x1 = p[cb * 3]; y1 = p[cb * 3 + 1]; z1 = p[cb * 3 + 2]; }
    A[cloud * npad + i] = make_float4(x0, x1, y0, y1);
    Z[cloud * npad + i] = make_float2(z0, z1);
}

// grid = (32-row clusters, clouds), 256 threads.  mask[((cloud * nclusters + cluster) * 3 + lev) * nwords + word]
__global__ void __launch_bounds__(256) emd_mask_kernel(int nr, int nc, int nwords, const float* __restrict__ rows, const int* __restrict__ perm,
                                                       const float* __restrict__ cands, float l0, float l1, float l2, unsigned* __restrict__ mask) {
    __shared__ float box[6];
    const int cloud = blockIdx.y, cluster = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const float inf = __int_as_float(0x7f800000);
    if (tid < 32) {   // bounding box of the cluster's (up to) 32 rows: one warp
        float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
        const int s = cluster * EMD_CLUSTER + tid;
        if (s < nr) {
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
            if (tid == 0) { box[a] = lo[a]; box[3 + a] = hi[a]; }
        }
    }
    __syncthreads();
    const float lo0 = box[0], lo1 = box[1], lo2 = box[2], hi0 = box[3], hi1 = box[4], hi2 = box[5];
    unsigned* __restrict__ out = mask + ((size_t)cloud * gridDim.x + cluster) * EMD_PRUNE_LEVELS * nwords;
    const float* __restrict__ cb = cands + (size_t)cloud * nc * 3;
    for (int c = tid; c < nwords * 32; c += 256) {
        float d2 = inf;
        if (c < nc) {
            const float vx = cb[(size_t)c * 3], vy = cb[(size_t)c * 3 + 1], vz = cb[(size_t)c * 3 + 2];
            const float dx = fmaxf(fmaxf(lo0 - vx, vx - hi0), 0.f), dy = fmaxf(fmaxf(lo1 - vy, vy - hi1), 0.f), dz = fmaxf(fmaxf(lo2 - vz, vz - hi2), 0.f);
            d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));   // distance to the box
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

// ---------------------------------------------------------------------------------------------------------------
// emd_sweep_kernel -- the split-sum variant (RFNET_EMD_SPLIT_SUMS): Q rows per thread as packed pairs, the candidate range
// of a sum cut into `nsplit` pieces handled by different CTAs (grid.x = b * nrt * nsplit); partial sums go to a small buffer
// and emd_epi_kernel adds them in split order and applies the pass's rule -- deterministic, no atomics, but a different
// rounding order than the reference's single chain.  Worth it only when a call holds one or two small clouds.
// (Fusing the epilogue through a last-CTA ticket was measured slower.)
// ---------------------------------------------------------------------------------------------------------------
template <int Q, int MODE, bool UNIT>
__global__ void __launch_bounds__(EMD_THREADS) emd_sweep_kernel(const __grid_constant__ SweepArgs a) {
    constexpr bool P3 = MODE == 3 || MODE == 4;
    constexpr bool DUAL = MODE == 4;
    __shared__ __align__(16) float4 sC[EMD_TC];
    __shared__ float sB[DUAL ? EMD_TC : 1];
    const int tid = threadIdx.x;
    const int bid = blockIdx.x;
    const int split = bid % a.nsplit;
    const int tile = (bid / a.nsplit) % a.nrt;
    const int cloud = bid / (a.nsplit * a.nrt);
    const int nr = a.nr, nc = a.nc;
    const float* __restrict__ rbase = a.rows + (size_t)cloud * nr * 3;
    const float* __restrict__ cbase = a.cands + (size_t)cloud * nc * 3;
    const float* __restrict__ wbase = a.w + (size_t)cloud * nc;
    const float* __restrict__ wbbase = DUAL ? a.wb + (size_t)cloud * nc : nullptr;
    int row[Q];
    const int r0 = tile * (EMD_THREADS * Q) + tid;
#pragma unroll
    for (int q = 0; q < Q; ++q) row[q] = (r0 + q * EMD_THREADS) < nr ? r0 + q * EMD_THREADS : -1;
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
    const int c_begin = split * a.split_len;
    const int c_end = min(nc, c_begin + a.split_len);
    for (int c0 = c_begin; c0 < c_end; c0 += EMD_TC) {
        const int len = min(EMD_TC, c_end - c0);
        __syncthreads();
        for (int i = tid; i < EMD_TC; i += EMD_THREADS) {
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
        __syncthreads();
        const int len2 = (len + EMD_UNROLL - 1) / EMD_UNROLL * EMD_UNROLL;  // padded entries carry weight 0 (EMD_TC % EMD_UNROLL == 0)
        EMD_PRAGMA_UNROLL(EMD_UNROLL)
        for (int k = 0; k < len2; ++k) {
            const float4 c = sC[k];
            const float2 cw = make_float2(c.w, c.w);
            float2 cwb = make_float2(0.f, 0.f);
            if (DUAL) { const float t = sB[k]; cwb = make_float2(t, t); }
#pragma unroll
            for (int h = 0; h < Q / 2; ++h) {
                if (UNIT && !DUAL) {
                    acc[h] = P3 ? __ffma2_rn(rf[h], cw, acc[h]) : __fadd2_rn(acc[h], cw);
                } else {
                    const float2 dx = __fadd2_rn(rx[h], make_float2(c.x, c.x));  // the sign of the difference is irrelevant after squaring
                    const float2 dy = __fadd2_rn(ry[h], make_float2(c.y, c.y));
                    const float2 dz = __fadd2_rn(rz[h], make_float2(c.z, c.z));
                    const float2 d2 = sqdist3x2<true>(dx, dy, dz);
                    const float2 x = __fmul2_rn(d2, L2);
                    float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                    if (P3) e = __fmul2_rn(rf[h], e);
                    acc[h] = __ffma2_rn(e, cw, acc[h]);
                    if (DUAL) {
                        if (UNIT) {
                            accb[h] = __fadd2_rn(accb[h], cwb);
                        } else {
                            const float2 xb = __fmul2_rn(d2, L2B);
                            accb[h] = __ffma2_rn(make_float2(ex2_approx(xb.x), ex2_approx(xb.y)), cwb, accb[h]);
                        }
                    }
                }
            }
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

// ---- epilogue of the split sums: one thread per row; add the split partials in fixed order, then the pass's rule ----
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
    float* grad2_partial;                // GRAD2: (b, gridDim.x * 4, n_oth, 3): one slice per warp (64 own points)
    EmdLevels lv;
};
template <int ROLE, bool WRITE, bool COST, bool GRAD, bool GRAD2, bool EXACT>
__global__ void __launch_bounds__(MT_THREADS) emd_pair_kernel(const __grid_constant__ PairArgs a) {
    // (the dependents are released at the END of this kernel: CTAs of the next sweep that become resident while this one runs are
    // placed on whatever SMs have room, and a sweep's CTAs live as long as the sweep -- measured 2.5x slower at 1 cloud of 16384 points)
    pdl_wait();   // everything this kernel reads was written by the kernels before it in the stream
    static_assert(!WRITE || ROLE == 1, "the matrix is written k-contiguous: own = xyz1");
    static_assert(!GRAD2 || (ROLE == 1 && GRAD), "the one-pass form: grad1 in registers, grad2 through the per-warp slab");
    constexpr int LB = GRAD2 ? 4 : 2;                 // others per unrolled batch
    __shared__ __align__(16) float sG2[GRAD2 ? MT_THREADS / 32 : 1][GRAD2 ? LB * 3 * 32 : 4];
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
#pragma unroll 1
        for (int lb = 0; lb < nl; lb += LB) {
#pragma unroll
        for (int u = 0; u < LB; ++u) {
            const int l = lb + u;
            if (l >= nl) {
                if (GRAD2) {   // a ragged tail still has to clear its slab rows
                    const int lane_ = threadIdx.x & 31, r_ = u * 3;
                    float* slab_ = sG2[threadIdx.x >> 5];
                    slab_[(r_ + 0) * 32 + lane_] = 0.f;
                    slab_[(r_ + 1) * 32 + lane_] = 0.f;
                    slab_[(r_ + 2) * 32 + lane_] = 0.f;
                }
                continue;
            }
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
                    const float2 dxg = make_float2(va ? dx.x : 0.f, vb ? dx.y : 0.f);
                    gx = __ffma2_rn(dxg, ns, gx);
                    gy = __ffma2_rn(dy, ns, gy);
                    gz = __ffma2_rn(dz, ns, gz);
                    if (GRAD2) {
                        // the other point's gradient, sum over k of match * (other - own) * rsqrt: this lane's two k's go to the
                        // warp's slab (one row of 32 lane partials per (other point, component)), reduced below every LB rows
                        const float2 px = __fmul2_rn(dxg, s), py = __fmul2_rn(dy, s), pz = __fmul2_rn(dz, s);
                        const int lane_ = threadIdx.x & 31, r_ = u * 3;
                        float* slab_ = sG2[threadIdx.x >> 5];
                        slab_[(r_ + 0) * 32 + lane_] = px.x + px.y;
                        slab_[(r_ + 1) * 32 + lane_] = py.x + py.y;
                        slab_[(r_ + 2) * 32 + lane_] = pz.x + pz.y;
                    }
                }
            }
        }
        if (GRAD2) {
            const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
            __syncwarp();
            {
                // all 32 lanes: lane (g, c) = (lane / 8, lane % 8) adds columns 4c..4c+3 of the three slab rows of other point lb + g (one
                // LDS.128 each, a quarter-warp reads one whole row: no bank conflicts), then three butterfly steps inside the 8-lane group
                // join the eight column chunks -- 8 issue slots per other point where 12 lanes walking 32 columns each cost 32.
                // A fixed order, the same for every call: deterministic and batch-invariant as before.
                static_assert(!GRAD2 || LB == 4, "four other points x eight column chunks = one warp");
                const int g_ = lane_ >> 3, c_ = lane_ & 7;
                const float4 v0 = *reinterpret_cast<const float4*>(&sG2[warp_][(3 * g_ + 0) * 32 + 4 * c_]);
                const float4 v1 = *reinterpret_cast<const float4*>(&sG2[warp_][(3 * g_ + 1) * 32 + 4 * c_]);
                const float4 v2 = *reinterpret_cast<const float4*>(&sG2[warp_][(3 * g_ + 2) * 32 + 4 * c_]);
                float t0 = (v0.x + v0.y) + (v0.z + v0.w), t1 = (v1.x + v1.y) + (v1.z + v1.w), t2 = (v2.x + v2.y) + (v2.z + v2.w);
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    t0 += __shfl_xor_sync(0xffffffffu, t0, off);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, off);
                    t2 += __shfl_xor_sync(0xffffffffu, t2, off);
                }
                if (c_ == 0 && lb + g_ < nl) {
                    float* gp2 = a.grad2_partial + (((size_t)cloud * (gridDim.x * (MT_THREADS / 32)) + blockIdx.x * (MT_THREADS / 32) + warp_) * n_oth + l0 + lb + g_) * 3;
                    gp2[0] = t0; gp2[1] = t1; gp2[2] = t2;
                }
            }
            __syncwarp();
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
    // (the dependents are released at the END of this kernel: CTAs of the next sweep that become resident while this one runs are
    // placed on whatever SMs have room, and a sweep's CTAs live as long as the sweep -- measured 2.5x slower at 1 cloud of 16384 points)
    pdl_wait();   // everything this kernel reads was written by the kernels before it in the stream
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

// ---------------------------------------------------------------------------------------------------------------
// match_cost gradient for a GIVEN matrix in ONE streaming pass over `match` (4 B per pair, the HBM floor); the reference
// reads it twice (matchcostgrad1 + matchcostgrad2, tf_approxmatch.cu:229-291) and so did the first version here.
// CTA = 128 consecutive k (a float4 of one matrix row per lane) x GF_LT rows l; warp w takes GF_LT / 8 consecutive rows
//   t(k,l) = match[l,k] * rsqrt(max(d2, 1e-20)) * (p2_l - p1_k)
//   grad1[k] = -sum_l t : 12 accumulators per lane across the whole l range, summed over the 8 warps at the end
//   grad2[l] = +sum_k t : the lane's 4-pair partial goes to a rotated (bank-conflict-free) per-WARP shared-memory slab, 8 rows
//              at a time; 24 lanes then add the 32 lane partials of one (l, component) each -- ~4 instructions per row and
//              lane where a shuffle tree would cost 30, and no block-wide barrier anywhere in the row loop (a first version
//              with a CTA-wide slab and two barriers per 64 rows drained the load pipeline at every barrier: 0.53 of HBM).
// Partials over l-ranges (grad1) and k-tiles (grad2) are summed in a fixed order by emd_grad_reduce_kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GF_THREADS = 256, GF_WARPS = 8, GF_KT = 128, GF_LT = 512, GF_UN = 8;
constexpr int GF_RPW = GF_LT / GF_WARPS;   // consecutive rows l per warp
__global__ void __launch_bounds__(GF_THREADS, 2) matchcostgrad_fused_kernel(int n, int m, int nkt, int nlt, const float* __restrict__ xyz1,
                                                                            const float* __restrict__ xyz2, const float* __restrict__ match,
                                                                            float* __restrict__ part1, float* __restrict__ part2) {
    // per-warp slab of lane partials: [row of the batch][component][lane, rotated by the row index so that both the lane-wise
    // writes and the row-wise reads are bank-conflict free].  Warps never wait for each other inside the row loop.
    __shared__ float sG[GF_WARPS][GF_UN * 3 * 32];
    const int cloud = blockIdx.z, kt = blockIdx.x, lt = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k4 = kt * GF_KT + lane * 4;
    const bool live = k4 < n;   // n % 4 == 0: a quad is either entirely inside or entirely outside
    Pts4 q;
    q.nx01 = q.nx23 = q.ny01 = q.ny23 = q.nz01 = q.nz23 = make_float2(0.f, 0.f);
    if (live) q = load_pts4_neg(xyz1 + ((size_t)cloud * n + k4) * 3);
    float2 ax01 = make_float2(0.f, 0.f), ax23 = ax01, ay01 = ax01, ay23 = ax01, az01 = ax01, az23 = ax01;   // sum_l t for the lane's four k
    const float* __restrict__ mbase = match + (size_t)cloud * m * n + k4;
    const float* __restrict__ p2 = xyz2 + (size_t)cloud * m * 3;
    float* __restrict__ slab = sG[warp];
    const int l_begin = lt * GF_LT + warp * GF_RPW;
    const int l_end = min(m, l_begin + GF_RPW);
#pragma unroll 1
    for (int l0 = l_begin; l0 < l_end; l0 += GF_UN) {
        float4 mv[GF_UN];
#pragma unroll
        for (int u = 0; u < GF_UN; ++u)   // all loads of the batch first: GF_UN x 16 B in flight per lane
            mv[u] = (live && l0 + u < l_end) ? __ldcs(reinterpret_cast<const float4*>(mbase + (size_t)(l0 + u) * n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < GF_UN; ++u) {
            const int l = min(l0 + u, m - 1);   // rows past the end carry match == 0: they add exact zeros
            const float px = __ldg(p2 + (size_t)l * 3), py = __ldg(p2 + (size_t)l * 3 + 1), pz = __ldg(p2 + (size_t)l * 3 + 2);
            const float2 ex01 = __fadd2_rn(q.nx01, make_float2(px, px)), ey01 = __fadd2_rn(q.ny01, make_float2(py, py)), ez01 = __fadd2_rn(q.nz01, make_float2(pz, pz));
            const float2 ex23 = __fadd2_rn(q.nx23, make_float2(px, px)), ey23 = __fadd2_rn(q.ny23, make_float2(py, py)), ez23 = __fadd2_rn(q.nz23, make_float2(pz, pz));
            const float2 d01 = sqdist3x2<true>(ex01, ey01, ez01), d23 = sqdist3x2<true>(ex23, ey23, ez23);
            const float2 s01 = __fmul2_rn(make_float2(mv[u].x, mv[u].y), make_float2(rsqrt_approx(fmaxf(d01.x, 1e-20f)), rsqrt_approx(fmaxf(d01.y, 1e-20f))));
            const float2 s23 = __fmul2_rn(make_float2(mv[u].z, mv[u].w), make_float2(rsqrt_approx(fmaxf(d23.x, 1e-20f)), rsqrt_approx(fmaxf(d23.y, 1e-20f))));
            ax01 = __ffma2_rn(ex01, s01, ax01); ay01 = __ffma2_rn(ey01, s01, ay01); az01 = __ffma2_rn(ez01, s01, az01);
            ax23 = __ffma2_rn(ex23, s23, ax23); ay23 = __ffma2_rn(ey23, s23, ay23); az23 = __ffma2_rn(ez23, s23, az23);
            const float2 gx = __ffma2_rn(ex01, s01, __fmul2_rn(ex23, s23)), gy = __ffma2_rn(ey01, s01, __fmul2_rn(ey23, s23)),
                         gz = __ffma2_rn(ez01, s01, __fmul2_rn(ez23, s23));
            const int r = u * 3;
            slab[(r + 0) * 32 + ((lane + r + 0) & 31)] = gx.x + gx.y;
            slab[(r + 1) * 32 + ((lane + r + 1) & 31)] = gy.x + gy.y;
            slab[(r + 2) * 32 + ((lane + r + 2) & 31)] = gz.x + gz.y;
        }
        __syncwarp();
        if (lane < GF_UN * 3 && l0 + lane / 3 < l_end) {   // one (row, component) per lane: add the 32 lane partials in lane order
            float t = 0.f;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) t += slab[lane * 32 + ((i + lane) & 31)];
            part2[(((size_t)cloud * nkt + kt) * m + l0) * 3 + lane] = t;   // GF_UN consecutive rows: 3 * GF_UN consecutive floats
        }
        __syncwarp();
    }
    // grad1: add the eight warps' accumulators (each saw different rows l of the same 128 k) and negate
    __shared__ float sW1[GF_WARPS][GF_KT * 3];
    float* sW = sW1[warp];
    sW[(lane * 4 + 0) * 3 + 0] = ax01.x; sW[(lane * 4 + 0) * 3 + 1] = ay01.x; sW[(lane * 4 + 0) * 3 + 2] = az01.x;
    sW[(lane * 4 + 1) * 3 + 0] = ax01.y; sW[(lane * 4 + 1) * 3 + 1] = ay01.y; sW[(lane * 4 + 1) * 3 + 2] = az01.y;
    sW[(lane * 4 + 2) * 3 + 0] = ax23.x; sW[(lane * 4 + 2) * 3 + 1] = ay23.x; sW[(lane * 4 + 2) * 3 + 2] = az23.x;
    sW[(lane * 4 + 3) * 3 + 0] = ax23.y; sW[(lane * 4 + 3) * 3 + 1] = ay23.y; sW[(lane * 4 + 3) * 3 + 2] = az23.y;
    __syncthreads();
    for (int i = tid; i < GF_KT * 3; i += GF_THREADS) {
        if (kt * GF_KT + i / 3 < n) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < GF_WARPS; ++w) t += sW1[w][i];
            part1[(((size_t)cloud * nlt + lt) * n + (size_t)kt * GF_KT) * 3 + i] = -t;
        }
    }
}

// ---- sweep launchers -------------------------------------------------------------------------------------------
template <int MODE, int NT>
static void emd_launch_row(const SweepArgs& a, unsigned grid, bool unit, bool exact, cudaStream_t s) {
    if (exact) {
        if (unit) launch_pdl(emd_row_kernel<MODE, true, true, NT>, dim3(grid), dim3(NT), 0, s, a);
        else launch_pdl(emd_row_kernel<MODE, false, true, NT>, dim3(grid), dim3(NT), 0, s, a);
    } else {
        if (unit) launch_pdl(emd_row_kernel<MODE, true, false, NT>, dim3(grid), dim3(NT), 0, s, a);
        else launch_pdl(emd_row_kernel<MODE, false, false, NT>, dim3(grid), dim3(NT), 0, s, a);
    }
}
template <int Q, int MODE>
static void emd_launch_split(const SweepArgs& a, unsigned grid, bool unit, cudaStream_t s) {
    if (unit) emd_sweep_kernel<Q, MODE, true><<<grid, EMD_THREADS, 0, s>>>(a);
    else emd_sweep_kernel<Q, MODE, false><<<grid, EMD_THREADS, 0, s>>>(a);
}
// One sweep.  rows/cands: the two clouds in the roles of this pass; pairA/pairZ: the candidates' pair layout.
// MODE 4: lvl2b / wb describe pass 1 of the next level.  mask != nullptr selects the pruned form (MODE 1..3).
// Kernel choice never changes a result except through RFNET_EMD_SPLIT_SUMS (the only thing that reorders a sum):
//   default      one chain per row: emd_row_kernel, 128 or 64 threads per CTA by the number of rows in the call
//   SPLIT_SUMS   emd_sweep_kernel<Q> over candidate splits + emd_epi_kernel (dense at every level)
template <int MODE>
static void emd_sweep(int b, int nr, int nc, int flags, float lvl2, float lvl2b, const float* rows, const float* cands, const float4* pairA,
                      const float2* pairZ, const float* w, const float* wb, const float* rowfac, float* remain, float* ratio, float* fac,
                      float* partial, const int* perm, const unsigned* mask, cudaStream_t s) {
    SweepArgs a;
    a.nr = nr; a.nc = nc; a.nwords = (nc + 31) / 32; a.npad = emd_npad(nc);
    a.lvl2 = lvl2; a.lvl2b = lvl2b; a.init0 = MODE == 1 ? 1e-9f : 0.0f;
    a.rows = rows; a.cands = cands; a.w = w; a.wb = wb; a.rowfac = rowfac;
    a.pairA = pairA; a.pairZ = pairZ;
    a.partial = partial; a.partial_b = partial;
    a.remain = remain; a.ratio = ratio; a.fac = fac;
    a.perm = perm; a.mask = mask;
    // bulk copies need 16-byte aligned weight rows: cloud stride nc * 4 bytes
    a.tma = (nc % 4 == 0) && ((((uintptr_t)w | (uintptr_t)(wb ? wb : w) | (uintptr_t)pairA | (uintptr_t)pairZ) & 15u) == 0);
    const bool exact = (flags & RFNET_EMD_EXACT) != 0;
    const bool unit = MODE == 4 ? lvl2b == 0.0f : lvl2 == 0.0f;
    if (!(flags & RFNET_EMD_SPLIT_SUMS)) {
        a.nsplit = 1; a.split_len = nc;
        if (mask != nullptr && MODE != 4) {
            // (the fused sweep is never pruned: each level has its own, tighter mask)
            constexpr int PM = MODE == 4 ? 1 : MODE;
            const int nclusters = (nr + EMD_CLUSTER - 1) / EMD_CLUSTER;
            a.nrt = (nclusters + PG_WARPS - 1) / PG_WARPS;
            launch_pdl(emd_pruned_kernel<PM>, dim3((unsigned)(b * a.nrt)), dim3(PG_WARPS * 32), 0, s, a);
            return;
        }
        // 64-thread CTAs beat 128-thread ones at every size (profiles/r2_emd_tune.txt); one-warp CTAs when even those leave fewer
        // than ~12 CTAs per SM (finer balance over the SMs, no block barrier)
        if ((long)b * nr >= 768L * num_sms()) {
            a.nrt = (nr + 63) / 64;
            emd_launch_row<MODE, 64>(a, (unsigned)(b * a.nrt), unit, exact, s);
        } else {
            a.nrt = (nr + 31) / 32;
            emd_launch_row<MODE, 32>(a, (unsigned)(b * a.nrt), unit, exact, s);
        }
        return;
    }
    const SweepPlan p = emd_plan(nr, nc);
    a.nrt = p.nrt; a.nsplit = p.nsplit; a.split_len = p.split_len;
    a.partial_b = partial + (size_t)b * nr * p.nsplit;
    const unsigned grid = (unsigned)(b * p.nrt * p.nsplit);
    if (p.Q == 4) emd_launch_split<4, MODE>(a, grid, unit, s);
    else emd_launch_split<2, MODE>(a, grid, unit, s);
    if (p.nsplit > 1) {
        const size_t total = (size_t)b * nr;
        emd_epi_kernel<MODE><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(nr, p.nsplit, total, a.partial, a.partial_b, remain, ratio, fac);
    }
}

// ---- final pass launcher ----------------------------------------------------------------------------------------
template <int ROLE, bool WRITE, bool COST, bool GRAD, bool GRAD2>
static void emd_launch_pair(const PairArgs& a, dim3 grid, bool exact, cudaStream_t s) {
    if (exact) launch_pdl(emd_pair_kernel<ROLE, WRITE, COST, GRAD, GRAD2, true>, grid, dim3(MT_THREADS), 0, s, a);
    else launch_pdl(emd_pair_kernel<ROLE, WRITE, COST, GRAD, GRAD2, false>, grid, dim3(MT_THREADS), 0, s, a);
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
static size_t emd_grad2_partials(int b, int n_own, int n_oth) {   // one slice per warp of the own-point tiles
    const dim3 g = emd_pair_grid(b, n_own, n_oth);
    return (size_t)b * g.x * (MT_THREADS / 32) * n_oth * 3;
}
__global__ void emd_grad_reduce_kernel(int n, int nt, size_t bn, const float* __restrict__ partial, float* __restrict__ grad) {
    // (the dependents are released at the END of this kernel: CTAs of the next sweep that become resident while this one runs are
    // placed on whatever SMs have room, and a sweep's CTAs live as long as the sweep -- measured 2.5x slower at 1 cloud of 16384 points)
    pdl_wait();   // everything this kernel reads was written by the kernels before it in the stream
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
    const bool split = (flags & RFNET_EMD_SPLIT_SUMS) != 0;
    if (!split) {   // the clouds as candidate pairs for the row kernel
        const int np1 = emd_npad(n), np2 = emd_npad(m);
        emd_pairs_kernel<<<dim3((unsigned)((np1 + 255) / 256), (unsigned)b), 256, 0, s>>>(n, np1, xyz1, ws.pairA1, ws.pairZ1);
        emd_pairs_kernel<<<dim3((unsigned)((np2 + 255) / 256), (unsigned)b), 256, 0, s>>>(m, np2, xyz2, ws.pairA2, ws.pairZ2);
    }
    // exact pruning of the three sharpest levels (see emd_row_kernel); relies on the flushing exponential's exact zeros
    // Pruned and dense sweeps give the same bits, so this choice may depend on the batch: the prologue (curve order + masks,
    // ~0.2 ms for 4 clouds of 16384 points) pays from about 32768 rows per call (measured, profiles/r2_timings_emd.txt: B=32
    // 2048^2 -5 %, B=4 16384^2 -12 %, B=32 16384^2 -14 %; B=4 2048^2 and B=32 1024^2 would lose 5-7 %).
    const bool big = n >= 2048 && m >= 2048 && ((n >= 4096 && m >= 4096) || (long)b * (n > m ? n : m) >= 32768);
    const bool prune = emd_prune_enabled(n, m) && (big || (flags & RFNET_EMD_PRUNE)) &&
                       !(flags & (RFNET_EMD_NO_PRUNE | RFNET_EMD_EXACT | RFNET_EMD_SPLIT_SUMS));
    const int nwA = (m + 31) / 32, nwB = (n + 31) / 32;
    if (prune) {
        { const int rc = morton_sort(b, n, m, xyz1, xyz2, ws.perm1, ws.perm2, s); if (rc) return rc; }
        emd_mask_kernel<<<dim3((unsigned)((n + EMD_CLUSTER - 1) / EMD_CLUSTER), (unsigned)b), 256, 0, s>>>(n, m, nwA, xyz1, ws.perm1, xyz2, lv.lvl2[0],
                                                                                                       lv.lvl2[1], lv.lvl2[2], ws.maskA);
        emd_mask_kernel<<<dim3((unsigned)((m + EMD_CLUSTER - 1) / EMD_CLUSTER), (unsigned)b), 256, 0, s>>>(m, n, nwB, xyz2, ws.perm2, xyz1, lv.lvl2[0],
                                                                                                       lv.lvl2[1], lv.lvl2[2], ws.maskB);
    }
    // levels 7 and 6 keep ~4 % / ~14 % of the pairs of a uniform 16384-point cloud; level 5 keeps ~58 %: not worth the gathers
    auto maskA = [&](int li) -> const unsigned* { return prune && li < EMD_PRUNED_SWEEP_LEVELS ? ws.maskA + (size_t)li * nwA : nullptr; };
    auto maskB = [&](int li) -> const unsigned* { return prune && li < EMD_PRUNED_SWEEP_LEVELS ? ws.maskB + (size_t)li * nwB : nullptr; };
    // pass 1 of the first level: rows = xyz1 (k), candidates = xyz2 (l) weighted by remainR            -> ratioL, facL[0]
    emd_sweep<1>(b, n, m, flags, lv.lvl2[0], 0.f, xyz1, xyz2, ws.pairA2, ws.pairZ2, ws.remainR, nullptr, nullptr, ws.remainL, ws.ratioL, ws.facL, ws.partial,
                 ws.perm1, maskA(0), s);
    for (int li = 0; li < EMD_LEVELS; ++li) {
        const float lvl2 = lv.lvl2[li];
        // pass 2: rows = xyz2 (l), candidates = xyz1 (k) weighted by ratioL                           -> ratioR, facR[li], remainR
        emd_sweep<2>(b, m, n, flags, lvl2, 0.f, xyz2, xyz1, ws.pairA1, ws.pairZ1, ws.ratioL, nullptr, nullptr, ws.remainR, ws.ratioR,
                     ws.facR + (size_t)li * bm, ws.partial, ws.perm2, maskB(li), s);
        // pass 3: rows = xyz1 (k), candidates = xyz2 (l) weighted by ratioR, row factor ratioL         -> remainL
        // fused with pass 1 of level li + 1 (weights remainR, already final)                          -> ratioL, facL[li + 1]
        // unless one of the two runs as a pruned sweep (each has its own, tighter candidate mask)
        const bool last = li == EMD_LEVELS - 1;
        if (!last && !maskA(li) && !maskA(li + 1)) {
            emd_sweep<4>(b, n, m, flags, lvl2, lv.lvl2[li + 1], xyz1, xyz2, ws.pairA2, ws.pairZ2, ws.ratioR, ws.remainR, ws.ratioL, ws.remainL, ws.ratioL,
                         ws.facL + (size_t)(li + 1) * bn, ws.partial, nullptr, nullptr, s);
        } else {
            emd_sweep<3>(b, n, m, flags, lvl2, 0.f, xyz1, xyz2, ws.pairA2, ws.pairZ2, ws.ratioR, nullptr, ws.ratioL, ws.remainL, nullptr, nullptr, ws.partial,
                         ws.perm1, maskA(li), s);
            if (!last)
                emd_sweep<1>(b, n, m, flags, lv.lvl2[li + 1], 0.f, xyz1, xyz2, ws.pairA2, ws.pairZ2, ws.remainR, nullptr, nullptr, ws.remainL, ws.ratioL,
                             ws.facL + (size_t)(li + 1) * bn, ws.partial, ws.perm1, maskA(li + 1), s);
        }
    }
    // final pass(es)
    float* extra = ws_floats + emd_ws_floats(b, n, m);   // carved by the *_workspace_bytes of the entry points that need it
    PairArgs pa;
    pa.n_own = n; pa.n_oth = m; pa.ot_len = emd_pair_tile(m); pa.b_own = bn; pa.b_oth = bm;
    pa.own = xyz1; pa.oth = xyz2; pa.fac_own = ws.facL; pa.fac_oth = ws.facR;
    pa.match = match; pa.cost_partial = extra; pa.grad_partial = nullptr; pa.grad2_partial = nullptr; pa.lv = lv;
    const dim3 g1 = emd_pair_grid(b, n, m);
    RFNET_CHECK_ARG(g1.y <= 65535 && g1.z <= 65535);
    if (grad1) {
        // ONE pass over all pairs: cost and grad1 in the registers of the thread that owns the xyz1 point, grad2 through per-warp
        // slabs (a first version ran a second pass with the roles swapped: twice the exponentials)
        float* gp1 = extra + emd_cost_partials(b, n, m);
        float* gp2 = gp1 + emd_grad_partials(b, n, m);
        pa.grad_partial = g1.y > 1 ? gp1 : grad1;
        pa.grad2_partial = gp2;
        emd_launch_pair<1, false, true, true, true>(pa, g1, exact, s);
        if (g1.y > 1) launch_pdl(emd_grad_reduce_kernel, dim3((unsigned)((bn * 3 + 255) / 256)), dim3(256), 0, s, n, (int)g1.y, bn, (const float*)gp1, grad1);
        launch_pdl(emd_grad_reduce_kernel, dim3((unsigned)((bm * 3 + 255) / 256)), dim3(256), 0, s, m, (int)(g1.x * (MT_THREADS / 32)), bm, (const float*)gp2, grad2);
    } else if (match && cost) {
        emd_launch_pair<1, true, true, false, false>(pa, g1, exact, s);
    } else if (cost) {
        emd_launch_pair<1, false, true, false, false>(pa, g1, exact, s);
    } else {
        emd_launch_pair<1, true, false, false, false>(pa, g1, exact, s);
    }
    if (cost) launch_pdl(reduce_partials_kernel, dim3(b), dim3(256), 0, s, (int)(g1.x * g1.y), (const float*)extra, cost);
    return launch_status();
}

}  // namespace rfnet

using namespace rfnet;

// EXACT excludes the split sums (a different rounding order) -- asking for both is a caller error
static bool emd_flags_ok(int flags) {
    if (flags & ~(RFNET_EMD_EXACT | RFNET_EMD_NO_PRUNE | RFNET_EMD_SPLIT_SUMS | RFNET_EMD_PRUNE)) return false;
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
    return (emd_ws_floats(b, n, m) + emd_cost_partials(b, n, m) + emd_grad_partials(b, n, m) + emd_grad2_partials(b, n, m)) * sizeof(float);
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
    const size_t legacy = sizeof(float) * 3 * (size_t)b * n * ((m + G1_L - 1) / G1_L);
    const size_t fused = sizeof(float) * 3 * (size_t)b * ((size_t)n * ((m + GF_LT - 1) / GF_LT) + (size_t)m * ((n + GF_KT - 1) / GF_KT));
    return legacy > fused ? legacy : fused;
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
    const size_t bn = (size_t)b * n, bm = (size_t)b * m;
    const bool vec = (n % 4 == 0) && ((((uintptr_t)match | (uintptr_t)xyz1 | (uintptr_t)workspace) & 15u) == 0);
    if (vec) {
        // one pass over the matrix: both gradients from a single read
        const int nkt = (n + GF_KT - 1) / GF_KT, nlt = (m + GF_LT - 1) / GF_LT;
        RFNET_CHECK_ARG(nlt <= 65535);
        float* part1 = (float*)workspace;
        float* part2 = part1 + 3 * bn * nlt;
        matchcostgrad_fused_kernel<<<dim3((unsigned)nkt, (unsigned)nlt, (unsigned)b), GF_THREADS, 0, s>>>(n, m, nkt, nlt, xyz1, xyz2, match, part1, part2);
        emd_grad_reduce_kernel<<<(unsigned)((bn * 3 + 255) / 256), 256, 0, s>>>(n, nlt, bn, part1, grad1);
        emd_grad_reduce_kernel<<<(unsigned)((bm * 3 + 255) / 256), 256, 0, s>>>(m, nkt, bm, part2, grad2);
        return launch_status();
    }
    const int nlt = (m + G1_L - 1) / G1_L;
    RFNET_CHECK_ARG(nlt <= 65535);
    dim3 g1((unsigned)((n + G1_THREADS - 1) / G1_THREADS), (unsigned)nlt, (unsigned)b);
    matchcostgrad1_kernel<<<g1, G1_THREADS, 0, s>>>(n, m, nlt, xyz1, xyz2, match, (float*)workspace);
    matchcostgrad1_reduce_kernel<<<(unsigned)((bn * 3 + 255) / 256), 256, 0, s>>>(n, nlt, bn, (const float*)workspace, grad1);
    dim3 g2((unsigned)((m + G2_WARPS - 1) / G2_WARPS), (unsigned)b);
    matchcostgrad2_kernel<<<g2, G2_WARPS * 32, 0, s>>>(n, m, xyz1, xyz2, match, grad2);
    return launch_status();
}
