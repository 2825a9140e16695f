// nn_distance (Chamfer nearest-neighbour search) and its gradient for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel of the reference (tf_ops/CD/tf_nndistance_g.cu:4-156).
//
// Two searches with bit-identical outputs (and one launch plan, pick_plan):
//   nn_search_kernel                      the reference distance expression for every pair (RFNET_NN_DIRECT, and small calls)
//   nn_prepare_kernel + nn_filter_kernel  the default: every pair visited with the expanded form |c|^2 - 2 q.c at half the FP32
//                                         work, the reference expression only where it decides the result (see there)
// Common design (not a port -- the reference scans 512-candidate tiles with one query per thread and merges tiles through a
// global read-modify-write):
//   * work item = (direction, cloud, tile of 128*Q queries, split of the candidate range); both directions in ONE launch;
//     grid sized so that >= ~30 items per SM exist even at 4 clouds per GPU;
//   * candidates are staged chunk by chunk in shared memory by the TMA bulk-copy engine (cp.async.bulk + mbarrier, double
//     buffered) -- raw (x,y,z) rows for the direct kernel (plain loads when rows are not 16-byte aligned), prepared
//     (x, y, z, |c|^2) rows for the filtered one;
//   * each thread keeps Q queries in registers as Q/2 packed pairs and evaluates two values per instruction with
//     Blackwell's packed FP32 pipe (FADD2 / FMUL2 / FFMA2); candidates are warp-broadcast LDS.128;
//   * min tracking is branch-free: eight (sixteen) values and the running best are folded with 3-input FMNMX3, one compare
//     + one predicated move remember the GROUP where the best strictly decreased; the index inside that group is recovered
//     once per work item by evaluating its distances (same operations, same bits).
//     (A first version branched to an index-selection slow path; ncu showed it taken in nearly every group, because 32
//     lanes x 8 candidates almost always contain a new running minimum near the start of an item.)
//   * splits are merged with a 64-bit atomicMin on (distance bits, index) keys, which is exactly the reference's
//     "smallest distance, lowest index" rule.
// The distance is evaluated in the reference's operand order (common.cuh: sqdist3x2), so indices are bit-exact.
#include "common.cuh"
#include "rfnet_ops.h"
#include "segscatter.cuh"

namespace rfnet {

#ifndef NN_THREADS_VALUE
#define NN_THREADS_VALUE 128
#endif
#ifndef NN_MIN_CTAS
#define NN_MIN_CTAS 5   // 96 registers per thread at 128 threads (tools/nn_tune.cu sweeps both)
#endif
constexpr int NN_THREADS = NN_THREADS_VALUE;
constexpr int NN_TC = 1024;  // max candidates per staged chunk (12 KiB per buffer)

// Adds (x, y, z) to a 12-byte row of a float buffer whose base is 8-byte aligned: every row has one 8-byte aligned pair -- (x,y)
// for even rows, (y,z) for odd rows -- which goes out as ONE vector reduction (red.global.add.v2.f32), the third value as a scalar
// one: 2 reductions per row instead of 3.  The kernel is bound by the issue of these reductions (ncu: mio_throttle).
__device__ __forceinline__ void red_add_row3(float* row, float x, float y, float z) {
    if ((reinterpret_cast<uintptr_t>(row) & 7u) == 0) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(row), "f"(x), "f"(y) : "memory");
        atomicAdd(row + 2, z);
    } else if ((reinterpret_cast<uintptr_t>(row + 1) & 7u) == 0) {
        atomicAdd(row, x);
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(row + 1), "f"(y), "f"(z) : "memory");
    } else {   // a base that is only 4-byte aligned
        atomicAdd(row, x); atomicAdd(row + 1, y); atomicAdd(row + 2, z);
    }
}
struct NNDir {
    const float* q;             // queries    (b, nq, 3)
    const float* c;             // candidates (b, nc, 3)
    float* dist;                // (b, nq)
    int* idx;                   // (b, nq)
    unsigned long long* keys;   // (b, nq) packed keys, used when nsplit > 1
    int nq, nc;
    int nqt;                    // query tiles per cloud
    int nsplit;                 // candidate splits per (cloud, tile)
    int cps;                    // chunks per split
    int chunk;                  // candidates per chunk: multiple of 8, <= NN_TC
    int items;                  // b * nqt * nsplit
    int tma;                    // candidate rows are 16-byte aligned -> bulk-copy path
    // filtered search only (nn_prepare_kernel writes both):
    const float4* cv;           // candidates as (x - ox, y - oy, z - oz, |c - o|^2), ncp rows per cloud (tail rows: |.|^2 = +inf)
    const float4* meta;         // per (cloud, split): (ox, oy, oz, max |c - o|^2) of the item's candidate range
    int ncp;                    // nc rounded up to whole groups
};
struct NNParams {
    NNDir d[2];
    unsigned long long* stats;  // optional (may be null): nn_filter_kernel counts the queries it had to scan exactly
};

template <int Q, bool FUSED>
__global__ void __launch_bounds__(NN_THREADS, NN_MIN_CTAS) nn_search_kernel(const NNParams p) {
    static_assert(Q % 2 == 0, "queries are processed as packed pairs");
    __shared__ __align__(128) float sC[2][NN_TC * 3];
    __shared__ __align__(8) uint64_t bar[2];

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int dir = bid >= p.d[0].items ? 1 : 0;
    if (dir) bid -= p.d[0].items;
    const NNDir& D = p.d[dir];

    const int split = bid % D.nsplit;
    const int tile = (bid / D.nsplit) % D.nqt;
    const int cloud = bid / (D.nsplit * D.nqt);
    const int nq = D.nq, nc = D.nc, chunk = D.chunk;

    const float* __restrict__ qbase = D.q + (size_t)cloud * nq * 3;
    const float* __restrict__ cbase = D.c + (size_t)cloud * nc * 3;

    // ---- queries into registers, as pairs (2h, 2h+1) -> (x: query tid + 2h*T, y: query tid + (2h+1)*T)
    const int q0 = tile * (NN_THREADS * Q) + tid;
    float2 qx[Q / 2], qy[Q / 2], qz[Q / 2];
    float best[Q];
    int bestk[Q];  // first candidate index of the 8-group in which `best` was attained
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int ia = q0 + (2 * h) * NN_THREADS, ib = ia + NN_THREADS;
        const bool va = ia < nq, vb = ib < nq;
        qx[h].x = va ? qbase[(size_t)ia * 3 + 0] : 0.f;
        qy[h].x = va ? qbase[(size_t)ia * 3 + 1] : 0.f;
        qz[h].x = va ? qbase[(size_t)ia * 3 + 2] : 0.f;
        qx[h].y = vb ? qbase[(size_t)ib * 3 + 0] : 0.f;
        qy[h].y = vb ? qbase[(size_t)ib * 3 + 1] : 0.f;
        qz[h].y = vb ? qbase[(size_t)ib * 3 + 2] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        best[i] = __int_as_float(0x7f800000);  // +inf: the first candidate always wins unless its distance is +inf too,
        bestk[i] = 0;                           // in which case index 0 stands -- same as the reference's k==0 seed
    }

    const int nchunks_total = (nc + chunk - 1) / chunk;
    const int first_chunk = split * D.cps;
    const int my_chunks = min(D.cps, nchunks_total - first_chunk);
    const bool tma = D.tma != 0;

    if (tma) {
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }

    auto issue = [&](int ci) {  // thread 0 only: start the bulk copy of chunk ci of this item
        const int start = (first_chunk + ci) * chunk;
        const int len = min(chunk, nc - start);
        const unsigned bytes = (unsigned)len * 12u;
        mbar_expect_tx(&bar[ci & 1], bytes);
        tma_bulk_g2s(sC[ci & 1], cbase + (size_t)start * 3, bytes, &bar[ci & 1]);
    };
    if (tma && tid == 0 && my_chunks > 0) issue(0);

    for (int ci = 0; ci < my_chunks; ++ci) {
        const int start = (first_chunk + ci) * chunk;
        const int len = min(chunk, nc - start);
        const int len8 = (len + 7) & ~7;
        float* sbuf = sC[ci & 1];

        if (tma) {
            if (tid == 0 && ci + 1 < my_chunks) issue(ci + 1);
            mbar_wait(&bar[ci & 1], (ci >> 1) & 1);
        } else {
            const float* __restrict__ src = cbase + (size_t)start * 3;
            for (int i = tid; i < len * 3; i += NN_THREADS) sbuf[i] = src[i];
        }
        if (!tma || len8 != len) {
            // pad the last group with +inf coordinates: their distances are +inf and can never win
            for (int i = len * 3 + tid; i < len8 * 3; i += NN_THREADS) sbuf[i] = __int_as_float(0x7f800000);
            __syncthreads();
        }

        const float4* __restrict__ c4 = reinterpret_cast<const float4*>(sbuf);
#pragma unroll 1
        for (int k = 0; k < len8; k += 8) {
            const float4 v0 = c4[(k >> 2) * 3 + 0], v1 = c4[(k >> 2) * 3 + 1], v2 = c4[(k >> 2) * 3 + 2];
            const float4 v3 = c4[(k >> 2) * 3 + 3], v4 = c4[(k >> 2) * 3 + 4], v5 = c4[(k >> 2) * 3 + 5];
            const float cx[8] = {v0.x, v0.w, v1.z, v2.y, v3.x, v3.w, v4.z, v5.y};
            const float cy[8] = {v0.y, v1.x, v1.w, v2.z, v3.y, v4.x, v4.w, v5.z};
            const float cz[8] = {v0.z, v1.y, v2.x, v2.w, v3.z, v4.y, v5.x, v5.w};
            const int kbase = start + k;
#pragma unroll
            for (int h = 0; h < Q / 2; ++h) {
                float2 d[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // (query - candidate) is the exact negation of the reference's (candidate - query); squares agree bit for bit
                    const float2 dx = __fadd2_rn(qx[h], make_float2(-cx[j], -cx[j]));
                    const float2 dy = __fadd2_rn(qy[h], make_float2(-cy[j], -cy[j]));
                    const float2 dz = __fadd2_rn(qz[h], make_float2(-cz[j], -cz[j]));
                    d[j] = sqdist3x2<FUSED>(dx, dy, dz);
                }
                // Branch-free min tracking: fold the eight distances and the running best with four FMNMX3; remember only
                // the GROUP in which the best strictly decreased (one predicated move).  The index inside the group is
                // recovered once per work item (resolve_group below).  Strict '<' keeps the FIRST group attaining the minimum.
                {
                    const float g = fmin3(fmin3(d[0].x, d[1].x, d[2].x), fmin3(d[3].x, d[4].x, d[5].x), fmin3(d[6].x, d[7].x, best[2 * h]));
                    bestk[2 * h] = (g < best[2 * h]) ? kbase : bestk[2 * h];
                    best[2 * h] = g;
                }
                {
                    const float g = fmin3(fmin3(d[0].y, d[1].y, d[2].y), fmin3(d[3].y, d[4].y, d[5].y), fmin3(d[6].y, d[7].y, best[2 * h + 1]));
                    bestk[2 * h + 1] = (g < best[2 * h + 1]) ? kbase : bestk[2 * h + 1];
                    best[2 * h + 1] = g;
                }
            }
        }
        __syncthreads();  // everyone is done with sbuf before the copy engine may overwrite it
    }

    // ---- results: recover the index inside the winning group by re-evaluating its (up to) eight distances -- the same
    // operations on the same operands as in the loop, hence the same bits -- and taking the first that equals `best`.
    // (best, bestk) go through shared memory (the candidate buffers are free now: the loop ended with a barrier) so that
    // this epilogue is a rolled loop and does not inflate the register budget of the main loop.
    float* sBest = sC[0];
    int* sBk = reinterpret_cast<int*>(sC[0]) + Q * NN_THREADS;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        sBest[i * NN_THREADS + tid] = best[i];
        sBk[i * NN_THREADS + tid] = bestk[i];
    }
#pragma unroll 1
    for (int i = 0; i < Q; ++i) {
        const int qi = q0 + i * NN_THREADS;
        if (qi >= nq) break;
        const float bd = sBest[i * NN_THREADS + tid];
        const int k0 = sBk[i * NN_THREADS + tid];
        const float qxs = qbase[(size_t)qi * 3 + 0], qys = qbase[(size_t)qi * 3 + 1], qzs = qbase[(size_t)qi * 3 + 2];
        const int lim = min(8, nc - k0);
        int j = 0;
#pragma unroll
        for (int t = 7; t >= 0; --t) {
            if (t < lim) {
                const float* __restrict__ c = cbase + (size_t)(k0 + t) * 3;
                const float dd = sqdist3<FUSED>(qxs - c[0], qys - c[1], qzs - c[2]);  // (query - candidate), as in the loop
                j = (dd == bd) ? t : j;
            }
        }
        const size_t o = (size_t)cloud * nq + qi;
        if (D.nsplit == 1) {
            D.dist[o] = bd;
            D.idx[o] = k0 + j;
        } else {
            atomicMin(&D.keys[o], pack_key(bd, k0 + j));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// nn_filter_kernel: the same all-pairs search at HALF the FP32-pipe work per pair, same bits out.
//
// Every (query, candidate) pair is still visited, but the scan evaluates the expanded form about the item's origin o
//     s(q,c) = |c-o|^2 - 2 (q-o).(c-o)          ( = d2(q,c) - |q-o|^2 in real arithmetic )
// as three packed FMAs per pair of queries (|c-o|^2 rides along as the candidate's 4th coordinate, nn_prepare_kernel), i.e.
// 3 FP32 lane-ops per pair where the reference expression needs 6 -- and uses it only to decide WHICH group of 16 candidates
// holds the nearest neighbour.  Per query the scan keeps the smallest group minimum b1, the group k1 it came from and the second
// smallest group minimum b2 (branch-free: 8 FMNMX3 / FMNMX for the group minimum, then 3 min/max + 1 compare + 1 select).  The
// certificate (nn_filter_certain below, with its derivation) turns b1 and b2 into a rigorous upper bound Y on the squared distance
// of the scan's best candidate and a lower bound X on that of every candidate outside its group, both in terms of the rounded
// differences the scan works with, and  sqrt(X) - sqrt(Y) > 9 * 2^-24 (|q-o| + max|c-o|)  proves that every candidate outside group k1
// has a reference distance STRICTLY above that of the candidate that produced b1, hence above the minimum of group k1: the exact
// arg-min (first index on ties) lies in group
// k1, whose G reference distances are then evaluated in the reference's operand order -- dist and idx come out bit for bit.
// Otherwise (two groups too close to separate: well under 1 % of the (query, item) pairs on unit-cube clouds; every query when
// distinct points sit at exactly equal distances) the warp scans the item's candidates for that query with the reference
// expression, 32 lanes wide -- or, if that happens to many of its queries, runs the direct loop for its whole tile
// (nn_direct_scan_warp).  Either way the item contributes the exact (distance, index) minimum of its candidate range; items
// are merged by the same 64-bit atomicMin as above.
// ---------------------------------------------------------------------------------------------------------------
#ifndef NNF_MIN_CTAS
#define NNF_MIN_CTAS 4
#endif
#ifndef NNF_GROUP
#define NNF_GROUP 16
#endif
constexpr int NNF_G = NNF_GROUP;   // candidates per tracked group: 8, 16 or 32
constexpr int NNF_DIRECT_PENDING = 24;   // uncertified queries per warp (of 32 x Q) from which the warp switches to the direct loop
static_assert(NNF_G == 8 || NNF_G == 16 || NNF_G == 32, "group size");

// The certificate.  b1 = s(c1) is the smallest scanned value of the item, b2 the smallest group minimum outside c1's group; q~ = fl(q - o),
// c~ = fl(c - o) are the ROUNDED differences the scan works with, u = 2^-24, L = |q~| + max|c~|.  Three error sources:
//   scan:       cn = fma(z,z, fma(x,x, y*y)) and s = fma(ax,x, fma(ay,y, fma(az,z, cn))), a = -2q~: six roundings, each relative to a
//               partial sum bounded by |c~|^2 + |a||c~|   =>  |s - (|c~|^2 - 2 q~.c~)| <= A := 6.1u (Cmax^2 + |q~| Cmax)
//   centring:   q~ - c~ = (q - c) + e, |e| <= 1.01u L      =>  | |q~ - c~| - |q - c| | <= 1.01u L
//   reference:  d2_ref = |q - c|^2 (1 + t), |t| <= 5.01u (three differences, three products, two sums, fused or not)
// With D~ = |q~ - c~|^2 = (|c~|^2 - 2q~.c~) + |q~|^2:  every c outside c1's group has D~(c) >= X := b2 - A + |q~|^2, and
// D~(c1) <= Y := b1 + A + |q~|^2.  Hence sqrt(d2_ref(c)) >= (1 - 2.51u)(sqrt(X) - 1.02u L) and sqrt(d2_ref(c1)) <= (1 + 2.51u)(sqrt(Y) + 1.02u L),
// and  sqrt(X) - sqrt(Y) > 9u L  (> 2.51u (sqrt(X) + sqrt(Y)) + 2.1u L, as sqrt(X), sqrt(Y) <= L (1 + ..))  implies d2_ref(c) > d2_ref(c1):
// the exact arg-min is in c1's group.  Every quantity below is rounded AGAINST the certificate (X down, Y up, L and A up), 1e-35
// covers gradual underflow in the scan (<= 2^-149 per rounding).  The test is about 3.5x sharper than a uniform bound
// E = 16u L^2 on |s + |q~|^2 - d2_ref| (the first version; tests/test_oracle.py checks both numerically: no violation in 1e7 certified
// pairs, near-ties included; the tightest certified pair differs by 3e-4 relative).
__device__ __forceinline__ bool nn_filter_certain(float b1, float b2, float rx, float ry, float rz, float cmax_up_sqrt) {
    const float INF = __int_as_float(0x7f800000);
    const float qn_lo = __fmaf_rd(rz, rz, __fmaf_rd(rx, rx, __fmul_rd(ry, ry)));
    const float qn_hi = __fmaf_ru(rz, rz, __fmaf_ru(rx, rx, __fmul_ru(ry, ry)));
    const float ql = __fsqrt_ru(qn_hi);
    const float L = __fadd_ru(ql, cmax_up_sqrt);
    const float A = __fmaf_ru(3.6359e-7f /* 6.1u, rounded up */, __fmaf_ru(cmax_up_sqrt, cmax_up_sqrt, __fmul_ru(ql, cmax_up_sqrt)), 1e-35f);
    const float X = __fadd_rd(__fadd_rd(b2, -A), qn_lo);
    const float Y = fmaxf(0.f, __fadd_ru(__fadd_ru(b1, A), qn_hi));
    // every comparison is false for NaN, so non-finite states fall through to the exact scan
    return X > 0.f && (__fadd_rd(__fsqrt_rd(X), -__fsqrt_ru(Y)) > __fmul_ru(5.3645e-7f /* 9u, rounded up */, L)) && (fabsf(b1) < INF) && (L < INF);
}

// Preparation of the candidates for the filtered search: one CTA per work-item range of a candidate cloud (at most NNP_RANGE
// rows -- the launch plan never makes an item longer).  It produces
//   * the range's origin o, the centre of its bounding box: the tolerance scales with (|q - o| + |c - o|)^2, so a cloud far from
//     the origin of its coordinate system would otherwise fail every certification; any o is valid, the roundings of q - o and
//     c - o are part of the error budget;
//   * the rows (x - ox, y - oy, z - oz, |c - o|^2), the cloud's last range padded to whole groups, and max |c - o|^2;
//   * the DUPLICATES.  Clouds padded to a fixed size by repeating points (the reference's resample_pcd, data_util.py:8-13)
//     give every query several candidates with bit-identical s, which no tolerance can tell apart.  But inside one item a
//     repeated point can never be the answer -- the copy with the lowest index has the same distance and wins the tie (copies
//     in different items are merged by the keys as ever) -- so the other copies leave the scan (|c|^2 = +inf).  Every point
//     bids its index for two slots of a hash table in shared memory (atomicMin); afterwards a point is a copy if the winner of
//     either slot is another point with equal coordinates.  Sets of equal points whose lowest index lost BOTH slots to other
//     points bid once more for a slot of a second, nearly empty table.  The lowest index of a set of equal points always
//     stays (a winner with equal coordinates would be a lower index); a copy that is still missed costs an exact scan, never
//     a wrong answer.
constexpr int NNP_THREADS = 512, NNP_K = 8, NNP_RANGE = NNP_THREADS * NNP_K;   // 4096 candidates per item at most
struct NNFill { unsigned* ptr; size_t words; unsigned value; };   // buffers the preparation launch fills on the side (keys, zeroed gradients)
struct NNPrep {
    const float* c[2];      // candidate clouds of direction 0 / 1 (b, nc, 3)
    float4* cv[2];
    float4* meta[2];        // per (cloud, range): (ox, oy, oz, max |c - o|^2)
    int nc[2], ncp[2], range[2], nsplit[2];
    unsigned slots[2];      // hash table size of the direction (power of two)
    unsigned blocks0;       // blocks of direction first_dir; the rest belong to the other direction
    int first_dir;
    NNFill fill[3];
};
// three slot numbers from one multiply-xor mix of the coordinate bits (the high bits of a product depend on all bits of its factor)
__device__ __forceinline__ unsigned nn_mix(float x, float y, float z) {
    return __float_as_uint(x) * 0x9E3779B1u ^ __float_as_uint(y) * 0x85EBCA77u ^ __float_as_uint(z) * 0xC2B2AE3Du;
}
__device__ __forceinline__ unsigned nn_slot1(unsigned h) { return h >> 17; }
__device__ __forceinline__ unsigned nn_slot2(unsigned h) { return (h * 0x27D4EB2Fu) >> 17; }
__device__ __forceinline__ unsigned nn_slot3(unsigned h) { return (h * 0x165667B1u) >> 19; }
__global__ void __launch_bounds__(NNP_THREADS) nn_prepare_kernel(const NNPrep p) {
    extern __shared__ __align__(16) unsigned tab[];
    __shared__ float sRed[NNP_THREADS / 32][6];
    __shared__ float sO[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();   // the search's CTAs may take the slots this grid frees (they wait for its completion themselves)
    pdl_wait();                // the previous call's kernels may still read what the fills below overwrite
    // ---- side job: this CTA's slice of the buffers the call needs filled (packed keys 0xff.., zeroed gradient outputs)
#pragma unroll 1
    for (int f = 0; f < 3; ++f) {
        const NNFill F = p.fill[f];
        if (F.words == 0) continue;
        const size_t per = ((F.words + gridDim.x - 1) / gridDim.x + 3) & ~(size_t)3;
        const size_t w0 = (size_t)blockIdx.x * per, w1 = w0 + per < F.words ? w0 + per : F.words;
        if ((reinterpret_cast<uintptr_t>(F.ptr) & 15u) == 0) {
            for (size_t i = w0 + (size_t)tid * 4; i < w1; i += (size_t)NNP_THREADS * 4) {
                if (i + 4 <= w1) *reinterpret_cast<uint4*>(F.ptr + i) = make_uint4(F.value, F.value, F.value, F.value);
                else for (size_t j = i; j < w1; ++j) F.ptr[j] = F.value;
            }
        } else {
            for (size_t i = w0 + tid; i < w1; i += NNP_THREADS) F.ptr[i] = F.value;
        }
    }
    unsigned blk = blockIdx.x;
    int d = p.first_dir;
    if (blk >= p.blocks0) { blk -= p.blocks0; d = 1 - d; }
    const int cloud = (int)(blk / (unsigned)p.nsplit[d]), split = (int)(blk % (unsigned)p.nsplit[d]);
    const int nc = p.nc[d];
    const int lo = split * p.range[d], hi = min(nc, lo + p.range[d]);
    const int hip = split == p.nsplit[d] - 1 ? p.ncp[d] : hi;   // the cloud's last range also writes the padding rows
    const float* __restrict__ src = p.c[d] + ((size_t)cloud * nc + lo) * 3;
    float4* __restrict__ dst = p.cv[d] + (size_t)cloud * p.ncp[d] + lo;
    const unsigned mask = p.slots[d] - 1, mask2 = mask >> 2;
    unsigned* __restrict__ tab2 = tab + mask + 1;
    const float INF = __int_as_float(0x7f800000);

    for (unsigned i = tid * 4; i < mask + 1 + ((mask + 1) >> 2); i += NNP_THREADS * 4)   // both tables
        *reinterpret_cast<uint4*>(tab + i) = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    float x[NNP_K], y[NNP_K], z[NNP_K];
    float lo0 = INF, lo1 = INF, lo2 = INF, hi0 = -INF, hi1 = -INF, hi2 = -INF;
#pragma unroll
    for (int k = 0; k < NNP_K; ++k) {
        const int li = tid + k * NNP_THREADS;
        const bool v = lo + li < hi;
        x[k] = v ? src[li * 3 + 0] : 0.f; y[k] = v ? src[li * 3 + 1] : 0.f; z[k] = v ? src[li * 3 + 2] : 0.f;
        if (v) {   // (fminf / fmaxf drop NaN)
            lo0 = fminf(lo0, x[k]); hi0 = fmaxf(hi0, x[k]);
            lo1 = fminf(lo1, y[k]); hi1 = fmaxf(hi1, y[k]);
            lo2 = fminf(lo2, z[k]); hi2 = fmaxf(hi2, z[k]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo0 = fminf(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); hi0 = fmaxf(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
        lo1 = fminf(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = fmaxf(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
        lo2 = fminf(lo2, __shfl_xor_sync(0xffffffffu, lo2, o)); hi2 = fmaxf(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
    }
    if (lane == 0) { sRed[warp][0] = lo0; sRed[warp][1] = lo1; sRed[warp][2] = lo2; sRed[warp][3] = hi0; sRed[warp][4] = hi1; sRed[warp][5] = hi2; }
    __syncthreads();   // (also: the tables are filled)
    if (tid < 3) {
        float a = INF, c = -INF;
        for (int w = 0; w < NNP_THREADS / 32; ++w) { a = fminf(a, sRed[w][tid]); c = fmaxf(c, sRed[w][3 + tid]); }
        float o = 0.5f * a + 0.5f * c;
        if (!(fabsf(o) < INF)) o = 0.f;   // no finite coordinate at all: every query ends in the exact scan anyway
        sO[tid] = o;
    }
    unsigned h[NNP_K];
#pragma unroll
    for (int k = 0; k < NNP_K; ++k) {
        const int li = tid + k * NNP_THREADS;
        h[k] = nn_mix(x[k], y[k], z[k]);
        if (lo + li < hi) {
            atomicMin(&tab[nn_slot1(h[k]) & mask], (unsigned)li);
            atomicMin(&tab[nn_slot2(h[k]) & mask], (unsigned)li);
        }
    }
    __syncthreads();
    // first round: copy / winner of a slot / neither.  The points that are neither -- a set of equal points whose lowest index lost
    // both slots to other points -- bid once more, for one slot of a second, nearly empty table (a quarter of the first).
    auto same = [&](unsigned w, int k) { return src[w * 3 + 0] == x[k] && src[w * 3 + 1] == y[k] && src[w * 3 + 2] == z[k]; };
    unsigned copies = 0u, again = 0u;
#pragma unroll
    for (int k = 0; k < NNP_K; ++k) {
        const int li = tid + k * NNP_THREADS;
        if (lo + li < hi) {
            const unsigned w1 = tab[nn_slot1(h[k]) & mask], w2 = tab[nn_slot2(h[k]) & mask];   // both <= li
            if ((w1 != (unsigned)li && same(w1, k)) || (w2 != (unsigned)li && same(w2, k))) copies |= 1u << k;
            else if (w1 != (unsigned)li && w2 != (unsigned)li) {
                again |= 1u << k;
                atomicMin(&tab2[nn_slot3(h[k]) & mask2], (unsigned)li);
            }
        }
    }
    __syncthreads();
    const float ox = sO[0], oy = sO[1], oz = sO[2];
    float cmax = 0.f;
#pragma unroll
    for (int k = 0; k < NNP_K; ++k) {
        const int li = tid + k * NNP_THREADS;
        if (lo + li < hip) {
            float4 v = make_float4(0.f, 0.f, 0.f, INF);   // padding rows and copies never produce a minimum
            if (lo + li < hi) {
                bool copy = (copies >> k) & 1u;
                if ((again >> k) & 1u) {
                    const unsigned w = tab2[nn_slot3(h[k]) & mask2];
                    copy = w != (unsigned)li && same(w, k);
                }
                v.x = x[k] - ox; v.y = y[k] - oy; v.z = z[k] - oz;
                if (!copy) {
                    v.w = __fmaf_rn(v.z, v.z, __fmaf_rn(v.x, v.x, __fmul_rn(v.y, v.y)));
                    cmax = fmaxf(cmax, v.w);   // drops NaN
                }
            }
            dst[li] = v;
        }
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(cmax));   // cmax >= +0: bits are monotone
    if (lane == 0) sRed[warp][0] = __uint_as_float(wmax);   // (sRed was last read two barriers ago)
    __syncthreads();
    if (tid == 0) {
        unsigned mx = 0u;
        for (int w = 0; w < NNP_THREADS / 32; ++w) mx = max(mx, __float_as_uint(sRed[w][0]));
        p.meta[d][(size_t)cloud * p.nsplit[d] + split] = make_float4(ox, oy, oz, __uint_as_float(mx));
    }
}

// The direct kernel's loop for one warp of nn_filter_kernel (see there): Q queries per lane against candidates [range_lo, range_hi)
// of the cloud, the reference expression for every pair, candidates broadcast from global memory.  Not inlined: it has its own
// register allocation and stays out of the way of the filtered scan.
template <int Q, bool FUSED>
__device__ __noinline__ void nn_direct_scan_warp(const float* __restrict__ qbase, const float* __restrict__ cbase, int q0, int nq, int range_lo, int range_hi,
                                                bool vec, bool direct_store, float* __restrict__ dist, int* __restrict__ idx, unsigned long long* __restrict__ keys) {
    const float INF = __int_as_float(0x7f800000);
    float2 rqx[Q / 2], rqy[Q / 2], rqz[Q / 2];
    float best[Q];
    int bestk[Q];
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int ia = q0 + (2 * h) * NN_THREADS, ib = ia + NN_THREADS;
        const bool va = ia < nq, vb = ib < nq;
        rqx[h].x = va ? qbase[(size_t)ia * 3 + 0] : 0.f; rqy[h].x = va ? qbase[(size_t)ia * 3 + 1] : 0.f; rqz[h].x = va ? qbase[(size_t)ia * 3 + 2] : 0.f;
        rqx[h].y = vb ? qbase[(size_t)ib * 3 + 0] : 0.f; rqy[h].y = vb ? qbase[(size_t)ib * 3 + 1] : 0.f; rqz[h].y = vb ? qbase[(size_t)ib * 3 + 2] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) { best[i] = INF; bestk[i] = range_lo; }
#pragma unroll 1
    for (int k = range_lo; k < range_hi; k += 8) {
        float cx[8], cy[8], cz[8];
        if (vec && k + 8 <= range_hi) {
            const float4* __restrict__ c4 = reinterpret_cast<const float4*>(cbase + (size_t)k * 3);
            const float4 v0 = __ldg(c4), v1 = __ldg(c4 + 1), v2 = __ldg(c4 + 2), v3 = __ldg(c4 + 3), v4 = __ldg(c4 + 4), v5 = __ldg(c4 + 5);
            cx[0] = v0.x; cy[0] = v0.y; cz[0] = v0.z; cx[1] = v0.w; cy[1] = v1.x; cz[1] = v1.y; cx[2] = v1.z; cy[2] = v1.w; cz[2] = v2.x;
            cx[3] = v2.y; cy[3] = v2.z; cz[3] = v2.w; cx[4] = v3.x; cy[4] = v3.y; cz[4] = v3.z; cx[5] = v3.w; cy[5] = v4.x; cz[5] = v4.y;
            cx[6] = v4.z; cy[6] = v4.w; cz[6] = v5.x; cx[7] = v5.y; cy[7] = v5.z; cz[7] = v5.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {   // rows past the end: +inf coordinates, +inf distances
                const bool in = k + j < range_hi;
                const float* __restrict__ c = cbase + (size_t)(in ? k + j : range_lo) * 3;
                cx[j] = in ? c[0] : INF; cy[j] = in ? c[1] : INF; cz[j] = in ? c[2] : INF;
            }
        }
#pragma unroll
        for (int h = 0; h < Q / 2; ++h) {
            float2 d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                d[j] = sqdist3x2<FUSED>(__fadd2_rn(rqx[h], make_float2(-cx[j], -cx[j])), __fadd2_rn(rqy[h], make_float2(-cy[j], -cy[j])),
                                        __fadd2_rn(rqz[h], make_float2(-cz[j], -cz[j])));
            const float g0 = fmin3(fmin3(d[0].x, d[1].x, d[2].x), fmin3(d[3].x, d[4].x, d[5].x), fmin3(d[6].x, d[7].x, best[2 * h]));
            bestk[2 * h] = (g0 < best[2 * h]) ? k : bestk[2 * h];
            best[2 * h] = g0;
            const float g1 = fmin3(fmin3(d[0].y, d[1].y, d[2].y), fmin3(d[3].y, d[4].y, d[5].y), fmin3(d[6].y, d[7].y, best[2 * h + 1]));
            bestk[2 * h + 1] = (g1 < best[2 * h + 1]) ? k : bestk[2 * h + 1];
            best[2 * h + 1] = g1;
        }
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {   // the first position of the winning 8-group that attains the minimum (same operations, same bits)
        const int qi = q0 + i * NN_THREADS;
        if (qi < nq) {
            const float qxs = qbase[(size_t)qi * 3 + 0], qys = qbase[(size_t)qi * 3 + 1], qzs = qbase[(size_t)qi * 3 + 2];
            const int lim = min(8, range_hi - bestk[i]);
            int j = 0;
#pragma unroll
            for (int t = 7; t >= 0; --t) {
                if (t < lim) {
                    const float* __restrict__ c = cbase + (size_t)(bestk[i] + t) * 3;
                    j = (sqdist3<FUSED>(qxs - c[0], qys - c[1], qzs - c[2]) == best[i]) ? t : j;
                }
            }
            if (direct_store) { dist[qi] = best[i]; idx[qi] = bestk[i] + j; }
            else atomicMin(&keys[qi], pack_key(best[i], bestk[i] + j));
        }
    }
}

template <int Q, bool FUSED>
__global__ void __launch_bounds__(NN_THREADS, NNF_MIN_CTAS) nn_filter_kernel(const NNParams p) {
    static_assert(Q % 2 == 0, "queries are processed as packed pairs");
    constexpr int G = NNF_G;
    __shared__ __align__(128) float4 sCv[2][NN_TC];   // (x, y, z, |c|^2) chunks relative to the cloud's origin, double buffered
    __shared__ __align__(8) uint64_t bar[2];

    pdl_launch_dependents();
    pdl_wait();   // nn_prepare_kernel's rows, origins and key fill
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int dir = bid >= p.d[0].items ? 1 : 0;
    if (dir) bid -= p.d[0].items;
    const NNDir& D = p.d[dir];

    const int split = bid % D.nsplit;
    const int tile = (bid / D.nsplit) % D.nqt;
    const int cloud = bid / (D.nsplit * D.nqt);
    const int nq = D.nq, nc = D.nc, chunk = D.chunk;

    const float* __restrict__ qbase = D.q + (size_t)cloud * nq * 3;
    const float* __restrict__ cbase = D.c + (size_t)cloud * nc * 3;
    const float4* __restrict__ cvbase = D.cv + (size_t)cloud * D.ncp;
    const int nchunks_total = (nc + chunk - 1) / chunk;
    const int first_chunk = split * D.cps;
    const int my_chunks = min(D.cps, nchunks_total - first_chunk);

    auto issue = [&](int ci) {  // thread 0 only: start the bulk copy of chunk ci of this item (whole groups: the rows are padded)
        const int start = (first_chunk + ci) * chunk;
        const int lenG = (min(chunk, nc - start) + G - 1) / G * G;
        const unsigned bytes = (unsigned)lenG * 16u;
        mbar_expect_tx(&bar[ci & 1], bytes);
        tma_bulk_g2s(sCv[ci & 1], cvbase + start, bytes, &bar[ci & 1]);
    };
    if (tid == 0) {   // the first chunk is on its way while the queries are loaded
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
        if (my_chunks > 0) issue(0);
    }
    const float4 meta = D.meta[(size_t)cloud * D.nsplit + split];
    const float ox = meta.x, oy = meta.y, oz = meta.z, cmax_cloud = meta.w;

    // ---- -2 * (query - origin) into registers, as pairs (2h, 2h+1) -> (x: query tid + 2h*T, y: query tid + (2h+1)*T)
    const int q0 = tile * (NN_THREADS * Q) + tid;
    float2 qx[Q / 2], qy[Q / 2], qz[Q / 2];
    float b1[Q], b2[Q];
    int k1[Q];
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int ia = q0 + (2 * h) * NN_THREADS, ib = ia + NN_THREADS;
        const bool va = ia < nq, vb = ib < nq;
        qx[h].x = va ? -2.0f * (qbase[(size_t)ia * 3 + 0] - ox) : 0.f;
        qy[h].x = va ? -2.0f * (qbase[(size_t)ia * 3 + 1] - oy) : 0.f;
        qz[h].x = va ? -2.0f * (qbase[(size_t)ia * 3 + 2] - oz) : 0.f;
        qx[h].y = vb ? -2.0f * (qbase[(size_t)ib * 3 + 0] - ox) : 0.f;
        qy[h].y = vb ? -2.0f * (qbase[(size_t)ib * 3 + 1] - oy) : 0.f;
        qz[h].y = vb ? -2.0f * (qbase[(size_t)ib * 3 + 2] - oz) : 0.f;
    }
    const float INF = __int_as_float(0x7f800000);
#pragma unroll
    for (int i = 0; i < Q; ++i) { b1[i] = INF; b2[i] = INF; k1[i] = 0; }
    __syncthreads();   // the barriers are initialised

    int item_live = 0;
    for (int ci = 0; ci < my_chunks; ++ci) {
        const int start = (first_chunk + ci) * chunk;
        const int lenG = (min(chunk, nc - start) + G - 1) / G * G;
        if (tid == 0 && ci + 1 < my_chunks) issue(ci + 1);
        mbar_wait(&bar[ci & 1], (ci >> 1) & 1);
        const float4* __restrict__ cv = sCv[ci & 1];
        int live = 0;   // does the chunk hold any candidate that takes part in the scan (not a copy, not padding)?
        for (int i = tid; i < lenG; i += NN_THREADS) live |= cv[i].w < INF;

#pragma unroll 1
        for (int k = 0; k < lenG; k += G) {
            float g[Q];
#pragma unroll
            for (int part = 0; part < G / 8; ++part) {
                float4 c[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) c[j] = cv[k + part * 8 + j];
#pragma unroll
                for (int h = 0; h < Q / 2; ++h) {
                    float2 s[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        s[j] = __ffma2_rn(qz[h], make_float2(c[j].z, c[j].z), make_float2(c[j].w, c[j].w));
                        s[j] = __ffma2_rn(qy[h], make_float2(c[j].y, c[j].y), s[j]);
                        s[j] = __ffma2_rn(qx[h], make_float2(c[j].x, c[j].x), s[j]);
                    }
                    if (part == 0) {
                        g[2 * h] = fmin3(fmin3(s[0].x, s[1].x, s[2].x), fmin3(s[3].x, s[4].x, s[5].x), fminf(s[6].x, s[7].x));
                        g[2 * h + 1] = fmin3(fmin3(s[0].y, s[1].y, s[2].y), fmin3(s[3].y, s[4].y, s[5].y), fminf(s[6].y, s[7].y));
                    } else {
                        g[2 * h] = fmin3(fmin3(g[2 * h], s[0].x, s[1].x), fmin3(s[2].x, s[3].x, s[4].x), fmin3(s[5].x, s[6].x, s[7].x));
                        g[2 * h + 1] = fmin3(fmin3(g[2 * h + 1], s[0].y, s[1].y), fmin3(s[2].y, s[3].y, s[4].y), fmin3(s[5].y, s[6].y, s[7].y));
                    }
                }
            }
            // smallest / second smallest group minimum and the group of the smallest (strict '<': the FIRST such group)
            const int kbase = start + k;
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                const float hi = fmaxf(g[i], b1[i]);
                k1[i] = (g[i] < b1[i]) ? kbase : k1[i];
                b1[i] = fminf(g[i], b1[i]);
                b2[i] = fminf(b2[i], hi);
            }
        }
        item_live |= __syncthreads_or(live);  // (barrier: everyone is done with sCv[ci&1] before the copy engine may overwrite it)
    }
    // An item whose candidates are all copies of points with lower indices (or padding) has nothing to contribute: each copy
    // loses the tie against its original in whichever item holds it.
    if (!item_live) {
        if (D.nsplit == 1)
            for (int i = 0; i < Q; ++i) {
                const int qi = q0 + i * NN_THREADS;
                if (qi < nq) { D.dist[(size_t)cloud * nq + qi] = INF; D.idx[(size_t)cloud * nq + qi] = 0; }
            }
        return;
    }

    // ---- per query: certify group k1 and evaluate it exactly, or queue the query for the exact warp scan.  The tracking
    // state goes through shared memory (the chunk buffers are free now) so that this epilogue is a rolled loop.
    float* sB1 = reinterpret_cast<float*>(&sCv[0][0]);
    float* sB2 = sB1 + Q * NN_THREADS;
    int* sK1 = reinterpret_cast<int*>(sB2 + Q * NN_THREADS);
    static_assert(3 * Q * NN_THREADS <= 2 * NN_TC * 4, "tracking state fits in the chunk buffers");
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        sB1[i * NN_THREADS + tid] = b1[i];
        sB2[i * NN_THREADS + tid] = b2[i];
        sK1[i * NN_THREADS + tid] = k1[i];
    }
    const int range_lo = first_chunk * chunk;
    const int range_hi = min(nc, range_lo + my_chunks * chunk);
    const float cmax_sqrt = __fsqrt_ru(__fmul_ru(cmax_cloud, 1.0000003f));   // max |c~|, rounded up (the prepared norms carry <= 3u)
    const bool vec = D.tma != 0;   // raw candidate rows are 16-byte aligned: G of them are 3G/4 vector loads
    auto emit = [&](int qi, float bd, int bi) {
        const size_t o = (size_t)cloud * nq + qi;
        if (D.nsplit == 1) {
            D.dist[o] = bd;
            D.idx[o] = bi;
        } else {
            atomicMin(&D.keys[o], pack_key(bd, bi));
        }
    };
    unsigned pending = 0;
#pragma unroll 1
    for (int i = 0; i < Q; ++i) {
        const int qi = q0 + i * NN_THREADS;
        if (qi >= nq) break;
        const float v1 = sB1[i * NN_THREADS + tid], v2 = sB2[i * NN_THREADS + tid];
        const int k0 = sK1[i * NN_THREADS + tid];
        const float qxs = qbase[(size_t)qi * 3 + 0], qys = qbase[(size_t)qi * 3 + 1], qzs = qbase[(size_t)qi * 3 + 2];
        const float rx = qxs - ox, ry = qys - oy, rz = qzs - oz;   // the same roundings as in the scan
        const bool certain = nn_filter_certain(v1, v2, rx, ry, rz, cmax_sqrt);
        if (certain) {
            // the G reference distances of the certified group, all independent (the loads and the arithmetic pipeline),
            // then their minimum and the FIRST position attaining it
            const int lim = min(G, range_hi - k0);
            float dd[G];
            if (vec && lim == G) {
                const float4* __restrict__ grp = reinterpret_cast<const float4*>(cbase + (size_t)k0 * 3);
                float r[G * 3];
#pragma unroll
                for (int t = 0; t < G * 3 / 4; ++t) {
                    const float4 v = __ldg(grp + t);
                    r[4 * t] = v.x; r[4 * t + 1] = v.y; r[4 * t + 2] = v.z; r[4 * t + 3] = v.w;
                }
#pragma unroll
                for (int t = 0; t < G; ++t) dd[t] = sqdist3<FUSED>(qxs - r[3 * t], qys - r[3 * t + 1], qzs - r[3 * t + 2]);   // (query - candidate), as nn_search_kernel
            } else {
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    const float* __restrict__ c = cbase + (size_t)(k0 + min(t, lim - 1)) * 3;
                    dd[t] = sqdist3<FUSED>(qxs - c[0], qys - c[1], qzs - c[2]);
                }
#pragma unroll
                for (int t = 0; t < G; ++t) dd[t] = (t < lim) ? dd[t] : INF;
            }
            float bd = fminf(dd[0], dd[1]);
#pragma unroll
            for (int t = 2; t < G; t += 2) bd = fmin3(bd, dd[t], dd[t + 1]);
            int bj = 0;
#pragma unroll
            for (int t = G - 1; t >= 0; --t) bj = (dd[t] == bd) ? t : bj;
            emit(qi, bd, k0 + bj);
        } else {
            pending |= 1u << i;
        }
    }
    const int lane = tid & 31;
    // ---- many uncertified queries in this warp (clouds full of distinct points at exactly equal distances, e.g. a lattice):
    // one query per warp step below would cost ~25 scan steps each, so the warp runs the direct kernel's loop instead -- all
    // Q queries of every lane against the item's candidates, the reference expression for every pair, candidates broadcast
    // from global memory -- at twice the cost of the filtered scan it replaces, whatever the number of ties.
    if (__reduce_add_sync(0xffffffffu, __popc(pending)) >= NNF_DIRECT_PENDING) {
        nn_direct_scan_warp<Q, FUSED>(qbase, cbase, q0, nq, range_lo, range_hi, vec, D.nsplit == 1, D.dist + (size_t)cloud * nq, D.idx + (size_t)cloud * nq,
                                      D.nsplit == 1 ? nullptr : D.keys + (size_t)cloud * nq);
        if (p.stats && pending) atomicAdd(p.stats, (unsigned long long)__popc(pending));
        pending = 0u;
    }
    // ---- exact scan of the item's candidate range for the queries that could not be certified, one query per warp step
    unsigned any = __ballot_sync(0xffffffffu, pending != 0);
    while (any) {
        const int srcl = __ffs(any) - 1;
        const unsigned pm = __shfl_sync(0xffffffffu, pending, srcl);
        const int i = __ffs(pm) - 1;
        const int qi = q0 - lane + srcl + i * NN_THREADS;
        const float qxs = qbase[(size_t)qi * 3 + 0], qys = qbase[(size_t)qi * 3 + 1], qzs = qbase[(size_t)qi * 3 + 2];
        float bd = INF;
        int bi = 0;
        for (int cidx = range_lo + lane; cidx < range_hi; cidx += 32) {
            const float* __restrict__ c = cbase + (size_t)cidx * 3;
            const float dd = sqdist3<FUSED>(qxs - c[0], qys - c[1], qzs - c[2]);
            if (dd < bd) { bd = dd; bi = cidx; }
        }
        unsigned long long key = pack_key(bd, bi);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if (lane == srcl) {
            emit(qi, __uint_as_float((unsigned)(key >> 32)), (int)(unsigned)(key & 0xffffffffull));
            pending &= pending - 1;
            if (p.stats) atomicAdd(p.stats, 1ull);
        }
        any = __ballot_sync(0xffffffffu, pending != 0);
    }
}

// keys of direction 0 (count0 entries) are followed by those of direction 1: one launch unpacks whichever were merged
__global__ void nn_unpack_keys_kernel(const unsigned long long* __restrict__ keys, float* __restrict__ dist1, int* __restrict__ idx1, size_t count0,
                                      float* __restrict__ dist2, int* __restrict__ idx2, size_t begin, size_t end) {
    pdl_wait();   // the search's keys
    const size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) {
        const unsigned long long k = keys[i];
        const float d = __uint_as_float((unsigned)(k >> 32));
        const int x = (int)(unsigned)(k & 0xffffffffull);
        if (i < count0) { dist1[i] = d; idx1[i] = x; }
        else { dist2[i - count0] = d; idx2[i - count0] = x; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient.  grad_self[j] = 2*gd[j]*(p_j - nn_j); grad_other[idx[j]] -= the same  (tf_nndistance_g.cu:131-150), both
// directions.  Phase A writes the own-point term for every point of both clouds with plain stores (so no memset is
// needed), phase B adds the scattered terms with red.global.add.f32.
// ---------------------------------------------------------------------------------------------------------------
__global__ void nn_grad_own_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                   const float* __restrict__ gd1, const int* __restrict__ idx1, const float* __restrict__ gd2,
                                   const int* __restrict__ idx2, float* __restrict__ g1, float* __restrict__ g2, size_t total1, size_t total2) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total1) {
        const size_t cloud = t / n;
        const int j2 = idx1[t];
        const float g = gd1[t] * 2.0f;
        const float* a = xyz1 + t * 3;
        const float* o = xyz2 + (cloud * m + j2) * 3;
        g1[t * 3 + 0] = g * (a[0] - o[0]);
        g1[t * 3 + 1] = g * (a[1] - o[1]);
        g1[t * 3 + 2] = g * (a[2] - o[2]);
    } else if (t < total1 + total2) {
        const size_t u = t - total1;
        const size_t cloud = u / m;
        const int j2 = idx2[u];
        const float g = gd2[u] * 2.0f;
        const float* a = xyz2 + u * 3;
        const float* o = xyz1 + (cloud * n + j2) * 3;
        g2[u * 3 + 0] = g * (a[0] - o[0]);
        g2[u * 3 + 1] = g * (a[1] - o[1]);
        g2[u * 3 + 2] = g * (a[2] - o[2]);
    }
}

__global__ void nn_grad_scatter_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                       const float* __restrict__ gd1, const int* __restrict__ idx1, const float* __restrict__ gd2,
                                       const int* __restrict__ idx2, float* __restrict__ g1, float* __restrict__ g2, size_t total1, size_t total2) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total1) {
        const size_t cloud = t / n;
        const int j2 = idx1[t];
        const float g = gd1[t] * 2.0f;
        const float* a = xyz1 + t * 3;
        const size_t oi = (cloud * m + j2) * 3;
        red_add_row3(g2 + oi, -(g * (a[0] - xyz2[oi + 0])), -(g * (a[1] - xyz2[oi + 1])), -(g * (a[2] - xyz2[oi + 2])));
    } else if (t < total1 + total2) {
        const size_t u = t - total1;
        const size_t cloud = u / m;
        const int j2 = idx2[u];
        const float g = gd2[u] * 2.0f;
        const float* a = xyz2 + u * 3;
        const size_t oi = (cloud * n + j2) * 3;
        red_add_row3(g1 + oi, -(g * (a[0] - xyz1[oi + 0])), -(g * (a[1] - xyz1[oi + 1])), -(g * (a[2] - xyz1[oi + 2])));
    }
}

// Atomic-free gradient (used when the caller provides a workspace): one thread per point of either cloud assembles that
// point's whole gradient in the reference's own summation order (NnDistanceGradOp::Compute, tf_nndistance.cpp:126-163):
//   point p of xyz1:  own term first, then  -= g2[k] * (xyz2[k] - xyz1[p])  for the k with idx2[k] == p, ascending k;
//   point q of xyz2:  -= g1[j] * (xyz1[j] - xyz2[q])  for the j with idx1[j] == q, ascending j, THEN += its own term
// with every product and sum rounded separately, so the result is bit-exact with the reference's CPU kernel and
// independent of thread timing.  Both index lists are inverted in ONE CSR over the n+m points of a cloud pair
// (seg::csr_build, integer atomics only): source r < m is row r of xyz2 and targets point idx2[r] of xyz1; source
// r >= m is row r-m of xyz1 and targets point n + idx1[r-m], i.e. a point of xyz2.
__global__ void nn_grad_combine_idx_kernel(int n, int m, const int* __restrict__ idx1, const int* __restrict__ idx2, int* __restrict__ comb) {
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (unsigned)(n + m)) return;
    const size_t cloud = blockIdx.y;
    comb[cloud * (n + m) + r] = r < (unsigned)m ? idx2[cloud * m + r] : n + idx1[cloud * n + (r - m)];
}
__global__ void nn_grad_seg_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2, const float* __restrict__ gd1,
                                   const int* __restrict__ idx1, const float* __restrict__ gd2, const int* __restrict__ idx2,
                                   const int* __restrict__ offset, const int* __restrict__ list, float* __restrict__ g1, float* __restrict__ g2) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)(n + m)) return;
    const size_t cloud = blockIdx.y;
    const float* __restrict__ A = xyz1 + cloud * (size_t)n * 3;
    const float* __restrict__ B = xyz2 + cloud * (size_t)m * 3;
    const int beg = offset[cloud * (n + m + 1) + t], end = offset[cloud * (n + m + 1) + t + 1];
    const int* __restrict__ seg = list + cloud * (size_t)(n + m);
    if (t < (unsigned)n) {
        const unsigned p = t;
        const float px = A[p * 3], py = A[p * 3 + 1], pz = A[p * 3 + 2];
        const int j2 = idx1[cloud * n + p];
        const float g = __fmul_rn(gd1[cloud * n + p], 2.0f);
        float ax = __fmul_rn(g, __fsub_rn(px, B[j2 * 3])), ay = __fmul_rn(g, __fsub_rn(py, B[j2 * 3 + 1])), az = __fmul_rn(g, __fsub_rn(pz, B[j2 * 3 + 2]));
        for (int e = beg; e < end; ++e) {
            const int k = seg[e];  // row of xyz2
            const float gk = __fmul_rn(gd2[cloud * m + k], 2.0f);
            ax = __fsub_rn(ax, __fmul_rn(gk, __fsub_rn(B[k * 3], px)));
            ay = __fsub_rn(ay, __fmul_rn(gk, __fsub_rn(B[k * 3 + 1], py)));
            az = __fsub_rn(az, __fmul_rn(gk, __fsub_rn(B[k * 3 + 2], pz)));
        }
        float* o = g1 + (cloud * n + p) * 3;
        o[0] = ax; o[1] = ay; o[2] = az;
    } else {
        const unsigned q = t - n;
        const float qx = B[q * 3], qy = B[q * 3 + 1], qz = B[q * 3 + 2];
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int e = beg; e < end; ++e) {
            const int j = seg[e] - m;  // row of xyz1
            const float gj = __fmul_rn(gd1[cloud * n + j], 2.0f);
            ax = __fsub_rn(ax, __fmul_rn(gj, __fsub_rn(A[j * 3], qx)));
            ay = __fsub_rn(ay, __fmul_rn(gj, __fsub_rn(A[j * 3 + 1], qy)));
            az = __fsub_rn(az, __fmul_rn(gj, __fsub_rn(A[j * 3 + 2], qz)));
        }
        const int j2 = idx2[cloud * m + q];
        const float g = __fmul_rn(gd2[cloud * m + q], 2.0f);
        ax = __fadd_rn(ax, __fmul_rn(g, __fsub_rn(qx, A[j2 * 3])));
        ay = __fadd_rn(ay, __fmul_rn(g, __fsub_rn(qy, A[j2 * 3 + 1])));
        az = __fadd_rn(az, __fmul_rn(g, __fsub_rn(qz, A[j2 * 3 + 2])));
        float* o = g2 + (cloud * m + q) * 3;
        o[0] = ax; o[1] = ay; o[2] = az;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Loss-level epilogue of chamfer_big / fidelity_loss (vv_recon.py:381-390): sums[0] = sum sqrt(dist1), sums[1] = #dist1,
// sums[2] = sum sqrt(dist2), sums[3] = #dist2.  Two tiny launches, fixed summation order (deterministic).
// ---------------------------------------------------------------------------------------------------------------
constexpr int CS_BLOCKS = 148;  // per direction
__global__ void __launch_bounds__(256) chamfer_sums_kernel(size_t n1, size_t n2, const float* __restrict__ dist1, const float* __restrict__ dist2,
                                                           float* __restrict__ partial) {
    __shared__ float sW[8];
    const int dir = blockIdx.x >= CS_BLOCKS;
    const int blk = blockIdx.x - dir * CS_BLOCKS;
    const float* __restrict__ d = dir ? dist2 : dist1;
    const size_t n = dir ? n2 : n1;
    float s = 0.f;
#pragma unroll 4
    for (size_t i = (size_t)blk * 256 + threadIdx.x; i < n; i += (size_t)CS_BLOCKS * 256) s += __fsqrt_rn(__ldg(d + i));
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += sW[i];
        partial[blockIdx.x] = t;
    }
}
__global__ void chamfer_sums_final_kernel(size_t n1, size_t n2, const float* __restrict__ partial, float* __restrict__ sums) {
    // two warps, one per direction; fixed lane-strided order followed by a fixed shuffle tree: deterministic
    const int dir = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float t = 0.f;
    for (int i = lane; i < CS_BLOCKS; i += 32) t += partial[dir * CS_BLOCKS + i];
    t = warp_sum(t);
    if (lane == 0) {
        sums[dir * 2] = t;
        sums[dir * 2 + 1] = (float)(dir ? n2 : n1);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
constexpr int NN_CTAS_PER_SM = NN_MIN_CTAS;  // resident CTAs per SM implied by the launch bounds

// One work item = one chunk of candidates for one query tile.  Every item pays a fixed price (query loads, pipeline fill,
// index resolution, key merge) worth roughly 48 candidates of scanning.  The hardware hands items to the resident CTA slots
// in blockIdx order -- direction 0's items, then direction 1's -- so the launch time is the makespan of list-scheduling
// two classes of equal-cost items on `slots` machines; nn_makespan evaluates that exactly (a handful of steps) and
// pick_plan takes, per direction, the chunk length (an even split of the candidate range, rounded up to a group) and the number
// of chunks per item that minimise it.
// With one shared power-of-two chunk the bench shape ran 2.77 waves of work in 3 rounds (7 % idle); see DESIGN.md 4.1.
static long nn_makespan(long slots, long cnt_a, long cost_a, long cnt_b, long cost_b) {
    // slot groups (time at which the slots of the group become free, how many), kept sorted by time; at most a few entries
    long gt[64], gc[64];
    int ng = 1;
    gt[0] = 0; gc[0] = slots;
    long end = 0;
    for (int cls = 0; cls < 2; ++cls) {
        long left = cls ? cnt_b : cnt_a;
        const long cost = cls ? cost_b : cost_a;
        while (left > 0) {
            // earliest group
            int e = 0;
            for (int i = 1; i < ng; ++i) if (gt[i] < gt[e]) e = i;
            const long use = gc[e] < left ? gc[e] : left;
            const long t1 = gt[e] + cost;
            left -= use;
            if (t1 > end) end = t1;
            if (use == gc[e]) { gt[e] = t1; }
            else { gc[e] -= use; if (ng < 64) { gt[ng] = t1; gc[ng] = use; ++ng; } }
            // merge groups with equal time
            for (int i = 0; i < ng; ++i)
                for (int j = i + 1; j < ng; ++j)
                    if (gt[i] == gt[j]) { gc[i] += gc[j]; gt[j] = gt[ng - 1]; gc[j] = gc[ng - 1]; --ng; --j; }
        }
    }
    return end;
}
struct NNPlan { int chunk0, cps0, chunk1, cps1; };
#ifdef NN_TUNE
static int g_force_chunk = 0, g_force_cps = 0;   // tools/nn_tune.cu sweeps these
#endif
// Cost model, in units of "one candidate scanned by one CTA": an item of `cps` chunks of `len` candidates costs
// cps * (len + per_chunk) + per_item.  direct kernel: per_chunk ~ 0 (the bulk copy is hidden), per_item ~ 48 (query loads,
// pipeline fill, index resolution, key merge).  filtered kernel: the scan is twice as fast, so the same overheads weigh
// double, each chunk pays a conversion pass and two block barriers, and each item one exact evaluation of a group per query.
static NNPlan pick_plan(int b, int n, int m, int Q, bool direct) {
    struct Memo { int b, n, m, Q, sms, direct; NNPlan plan; };
    static thread_local Memo memo = {0, 0, 0, 0, 0, 0, {0, 0, 0, 0}};   // a pure function of its arguments: the memo only saves host time
    const int sms = num_sms();
    if (memo.b == b && memo.n == n && memo.m == m && memo.Q == Q && memo.sms == sms && memo.direct == (int)direct) return memo.plan;
    const int TQ = NN_THREADS * Q;
    const long slots = (long)sms * (direct ? NN_CTAS_PER_SM : NNF_MIN_CTAS);
    const long per_chunk = direct ? 0 : 8, per_item = direct ? 48 : 430;
    const long tiles0 = (long)b * ((n + TQ - 1) / TQ), tiles1 = (long)b * ((m + TQ - 1) / TQ);
    const int grp = direct ? 8 : NNF_G;   // chunks are whole groups of the kernel that scans them
    auto lens = [grp](int nc, int* out) {   // even splits of nc into k pieces, rounded up to a group, between 256 and NN_TC candidates
        int cnt = 0, last = 0;
        for (int k = 1; k <= 64 && cnt < 64; ++k) {
            int len = ((nc + k - 1) / k + grp - 1) / grp * grp;
            if (len > NN_TC) continue;
            if (len < 256 && cnt > 0) break;
            if (len != last) out[cnt++] = len;
            last = len;
        }
        if (cnt == 0) out[cnt++] = NN_TC;
        return cnt;
    };
    int l0[64], l1[64];
    const int c0 = lens(m, l0), c1 = lens(n, l1);
    const int cps_opts[] = {1, 2, 3, 4, 6, 8};
    const int ncps = direct ? 1 : 6;   // the direct kernel's per-item price is small: one chunk per item spreads best
    long best = -1;
    NNPlan plan = {NN_TC, 1, NN_TC, 1};
    for (int i = 0; i < c0; ++i)
        for (int ci = 0; ci < ncps; ++ci)
            for (int j = 0; j < c1; ++j)
                for (int cj = 0; cj < ncps; ++cj) {
                    const long k0 = (m + l0[i] - 1) / l0[i], k1 = (n + l1[j] - 1) / l1[j];
                    const long p0 = cps_opts[ci] < k0 ? cps_opts[ci] : k0, p1 = cps_opts[cj] < k1 ? cps_opts[cj] : k1;
                    if ((ci > 0 && cps_opts[ci] > k0) || (cj > 0 && cps_opts[cj] > k1)) continue;   // same plan as a smaller option
                    if (!direct && (cps_opts[ci] * l0[i] > NNP_RANGE || cps_opts[cj] * l1[j] > NNP_RANGE)) continue;   // longer than the preparation reaches
                    const long s0 = (k0 + p0 - 1) / p0, s1 = (k1 + p1 - 1) / p1;
                    const long cost0 = p0 * ((m < l0[i] ? m : l0[i]) + per_chunk) + per_item;
                    const long cost1 = p1 * ((n < l1[j] ? n : l1[j]) + per_chunk) + per_item;
                    // direct kernel: the makespan of list-scheduling the two classes of items in blockIdx order (items of fixed duration).
                    // filtered kernel: its CTAs share the SM's pipes, so the launch behaves like a throughput problem -- total work over
                    // the slots -- plus a tail of about 0.6 of the longest item; per_item and the 0.6 are fitted to measured plans
                    // (profiles/r2_nn_filter_plans.txt: 12 forced plans x 3 shapes, the fixed-duration model misranked them by up to 14 %)
                    long t;
                    if (direct) {
                        t = nn_makespan(slots, tiles0 * s0, cost0, tiles1 * s1, cost1);
                    } else {
                        const long total = tiles0 * s0 * cost0 + tiles1 * s1 * cost1, longest = cost0 > cost1 ? cost0 : cost1;
                        const long per_slot = (total + slots - 1) / slots;
                        t = (per_slot > longest ? per_slot : longest) + longest * 6 / 10;
                    }
                    if (best < 0 || t < best) { best = t; plan = {l0[i], (int)p0, l1[j], (int)p1}; }
                }
    memo = {b, n, m, Q, sms, (int)direct, plan};
    return plan;
}

static void plan_direction(NNDir& D, int b, int nq, int nc, int Q, int chunk, int cps, bool direct) {
    D.nq = nq;
    D.nc = nc;
    const int TQ = NN_THREADS * Q;
    D.nqt = (nq + TQ - 1) / TQ;
    D.chunk = chunk;
    const int nchunks = (nc + chunk - 1) / chunk;
    D.cps = cps > 0 ? (cps < nchunks ? cps : nchunks) : nchunks;   // cps <= 0: the whole candidate range in one item ...
    if (!direct && D.cps * chunk > NNP_RANGE) D.cps = NNP_RANGE / chunk;   // ... as far as the preparation of the filtered search reaches
    D.nsplit = (nchunks + D.cps - 1) / D.cps;
    D.items = b * D.nqt * D.nsplit;
    D.tma = (nc % 4 == 0) && (((uintptr_t)D.c & 15u) == 0);
}

#ifdef NN_FORCE_Q
static int pick_q(int) { return NN_FORCE_Q; }   // tools/nn_tune.cu
#else
static int pick_q(int nq) { return nq >= 1024 ? 8 : (nq >= 512 ? 4 : 2); }
#endif

template <int Q>
static void launch_search(const NNParams& p, int grid, bool fused, bool direct, cudaStream_t s) {
    if (direct) {
        if (fused)
            nn_search_kernel<Q, true><<<grid, NN_THREADS, 0, s>>>(p);
        else
            nn_search_kernel<Q, false><<<grid, NN_THREADS, 0, s>>>(p);
    } else {
        if (fused)
            launch_pdl(nn_filter_kernel<Q, true>, dim3(grid), dim3(NN_THREADS), 0, s, p);
        else
            launch_pdl(nn_filter_kernel<Q, false>, dim3(grid), dim3(NN_THREADS), 0, s, p);
    }
}

}  // namespace rfnet

using namespace rfnet;

// workspace: [ keys of direction 0 (b*n) | keys of direction 1 (b*m) ] [ prepared candidates of direction 0 (xyz2) | of direction 1
// (xyz1) ] [ origin and max norm per (cloud, item range), direction 0 | direction 1 ]
static inline int nn_padded(int nc) { return (nc + NNF_G - 1) / NNF_G * NNF_G; }
static inline int nn_max_ranges(int nc) { return (nc + 255) / 256; }   // chunks are at least 256 candidates long
struct NNLayout { size_t cv0, cv1, meta0, meta1, total; };
static NNLayout nn_layout(int b, int n, int m) {
    NNLayout L;
    size_t o = ((size_t)b * ((size_t)n + (size_t)m) * sizeof(unsigned long long) + 15) & ~(size_t)15;
    L.cv0 = o; o += sizeof(float4) * (size_t)b * nn_padded(m);
    L.cv1 = o; o += sizeof(float4) * (size_t)b * nn_padded(n);
    L.meta0 = o; o += sizeof(float4) * (size_t)b * nn_max_ranges(m);
    L.meta1 = o; o += sizeof(float4) * (size_t)b * nn_max_ranges(n);
    L.total = o;
    return L;
}
extern "C" size_t rfnet_nn_distance_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return nn_layout(b, n, m).total;
}

static int nn_prepare_launch(const NNPrep& pp, unsigned blocks, unsigned slots, cudaStream_t s) {
    // two hash tables in dynamic shared memory: `slots` and slots / 4 words (at most 2.5 x NNP_RANGE words = 40 KiB)
    static_assert((2 * NNP_RANGE + NNP_RANGE / 2) * sizeof(unsigned) <= 48 * 1024, "fits the default dynamic shared-memory limit");
    launch_pdl(nn_prepare_kernel, dim3(blocks), dim3(NNP_THREADS), (slots + slots / 4) * sizeof(unsigned), s, pp);
    return 0;
}

// plan + key memset + search launch; *need0 / *need1 tell the caller which directions left their results as packed keys
// dirs: 1 = xyz1 queries against xyz2 only, 2 = xyz2 queries against xyz1 only, 3 = both
// The launch plan of a call: which kernel (return value: true = direct), and per direction the chunk length, chunks per item,
// items.  A pure function of (b, n, m, flags) and the device's SM count.
static bool nn_plan(NNParams& p, int b, int n, int m, int Q, int flags) {
    // small problems (< 2^27 pairs, a few tens of microseconds of work) are launch-bound: one item per query tile, results
    // written directly, no key merge and no extra launches
    const bool split = 2.0 * b * (double)n * (double)m >= 134217728.0;
    // below 2^24 pairs a call is a few microseconds of work and launch-bound: the direct kernel needs no preparation launch; and
    // batches of tiny clouds (under 1024 points on both sides) are all per-item overhead, of which the direct kernel has less
    const bool direct = (flags & RFNET_NN_DIRECT) != 0 || 2.0 * b * (double)n * (double)m < 16777216.0 || (n < 1024 && m < 1024);
    NNPlan plan = {NN_TC, 0, NN_TC, 0};
    if (split) plan = pick_plan(b, n, m, Q, direct);
#ifdef NN_TUNE
    if (split && g_force_chunk) plan.chunk0 = plan.chunk1 = g_force_chunk;
    if (split && g_force_cps) plan.cps0 = plan.cps1 = g_force_cps;
#endif
    plan_direction(p.d[0], b, n, m, Q, plan.chunk0, plan.cps0, direct);
    plan_direction(p.d[1], b, m, n, Q, plan.chunk1, plan.cps1, direct);
    return direct;
}

// Diagnostics / tests: the plan rfnet_nn_distance would use, without launching anything.
// out = { direct, Q,  chunk0, chunks_per_item0, splits0, items0,  chunk1, chunks_per_item1, splits1, items1 }
extern "C" int rfnet_nn_distance_plan(int b, int n, int m, int flags, int* out10) {
    RFNET_CHECK_ARG(b > 0 && n > 0 && m > 0 && out10);
    NNParams p;
    p.d[0].c = p.d[1].c = nullptr;   // (alignment of the clouds is not part of the plan)
    const int Q = pick_q(n < m ? n : m);
    const bool direct = nn_plan(p, b, n, m, Q, flags);
    out10[0] = direct ? 1 : 0; out10[1] = Q;
    for (int d = 0; d < 2; ++d) { out10[2 + 4 * d] = p.d[d].chunk; out10[3 + 4 * d] = p.d[d].cps; out10[4 + 4 * d] = p.d[d].nsplit; out10[5 + 4 * d] = p.d[d].items; }
    return 0;
}

static int nn_search_launch(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1, float* dist2, int* idx2, void* workspace,
                            size_t workspace_bytes, int flags, cudaStream_t s, bool* need0_out, bool* need1_out, int dirs = 3,
                            unsigned long long* stats = nullptr, const NNFill* side_fill = nullptr, int n_side_fill = 0, bool* side_filled = nullptr) {
    // side_fill: up to two buffers of the caller that need a constant fill before ITS next kernel; the preparation launch of the
    // filtered search does it on the side (*side_filled = true), otherwise the caller issues its own memsets
    const int Q = pick_q(n < m ? n : m);
    NNParams p;
    p.stats = stats;
    p.d[0].q = xyz1; p.d[0].c = xyz2; p.d[0].dist = dist1; p.d[0].idx = idx1;
    p.d[1].q = xyz2; p.d[1].c = xyz1; p.d[1].dist = dist2; p.d[1].idx = idx2;
    const bool fused = !(flags & RFNET_NN_UNFUSED);
    const bool direct = nn_plan(p, b, n, m, Q, flags);
    if (!(dirs & 1)) { p.d[0].items = 0; p.d[0].nsplit = 1; }   // a direction without items launches no CTA and merges no keys
    if (!(dirs & 2)) { p.d[1].items = 0; p.d[1].nsplit = 1; }
    unsigned long long* keys = (unsigned long long*)workspace;
    p.d[0].keys = keys;
    p.d[1].keys = keys ? keys + (size_t)b * n : nullptr;
    const bool need0 = p.d[0].nsplit > 1, need1 = p.d[1].nsplit > 1;
    if (!direct) {
        // filtered search: the candidate range of every work item is prepared once per call (origin, duplicates, |c|^2)
        RFNET_CHECK_ARG(workspace && workspace_bytes >= rfnet_nn_distance_workspace_bytes(b, n, m));
        const NNLayout L = nn_layout(b, n, m);
        char* w = (char*)workspace;
        NNPrep pp;
        pp.c[0] = xyz2; pp.c[1] = xyz1;
        pp.cv[0] = reinterpret_cast<float4*>(w + L.cv0); pp.cv[1] = reinterpret_cast<float4*>(w + L.cv1);
        pp.meta[0] = reinterpret_cast<float4*>(w + L.meta0); pp.meta[1] = reinterpret_cast<float4*>(w + L.meta1);
        unsigned slots_max = 0, blocks[2];
        for (int d = 0; d < 2; ++d) {
            const NNDir& D = p.d[d];
            pp.nc[d] = D.nc; pp.ncp[d] = nn_padded(D.nc); pp.range[d] = D.cps * D.chunk; pp.nsplit[d] = D.nsplit;
            unsigned sl = 64;
            while (sl < 2u * (unsigned)(pp.range[d] < D.nc ? pp.range[d] : D.nc)) sl <<= 1;   // two bids per point: half full at most
            pp.slots[d] = sl;
            blocks[d] = D.items ? (unsigned)b * D.nsplit : 0u;
            if (blocks[d] && sl > slots_max) slots_max = sl;
            p.d[d].cv = pp.cv[d]; p.d[d].ncp = pp.ncp[d]; p.d[d].meta = pp.meta[d];
        }
        pp.first_dir = blocks[0] ? 0 : 1;
        pp.blocks0 = blocks[pp.first_dir];
        // the packed keys (0xff..: every real key is smaller) and the caller's buffers are filled by the same launch
        for (int f = 0; f < 3; ++f) pp.fill[f] = NNFill{nullptr, 0, 0u};
        if (need0 || need1) {
            unsigned long long* first = need0 ? p.d[0].keys : p.d[1].keys;
            pp.fill[0] = NNFill{reinterpret_cast<unsigned*>(first), 2 * ((need0 ? (size_t)b * n : 0) + (need1 ? (size_t)b * m : 0)), 0xffffffffu};
        }
        for (int f = 0; f < n_side_fill && f < 2; ++f) pp.fill[1 + f] = side_fill[f];
        if (side_filled) *side_filled = true;
        { const int rc = nn_prepare_launch(pp, blocks[0] + blocks[1], slots_max, s); if (rc) return rc; }
    } else if (need0 || need1) {
        RFNET_CHECK_ARG(workspace && workspace_bytes >= rfnet_nn_distance_workspace_bytes(b, n, m));
        // the two key arrays are contiguous: one memset covers whichever directions are merged
        unsigned long long* first = need0 ? p.d[0].keys : p.d[1].keys;
        const size_t cnt = (need0 ? (size_t)b * n : 0) + (need1 ? (size_t)b * m : 0);
        RFNET_CUDA(cudaMemsetAsync(first, 0xff, sizeof(unsigned long long) * cnt, s));
    }
    const int grid = p.d[0].items + p.d[1].items;
    if (Q == 8) launch_search<8>(p, grid, fused, direct, s);
    else if (Q == 4) launch_search<4>(p, grid, fused, direct, s);
    else launch_search<2>(p, grid, fused, direct, s);
    *need0_out = need0;
    *need1_out = need1;
    return 0;
}

extern "C" int rfnet_nn_distance(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1, float* dist2,
                                 int* idx2, void* workspace, size_t workspace_bytes, int flags, rfnet_stream_t stream) {
    return rfnet_nn_distance_stats(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, workspace, workspace_bytes, flags, nullptr, stream);
}

extern "C" int rfnet_nn_distance_stats(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1, float* dist2,
                                       int* idx2, void* workspace, size_t workspace_bytes, int flags, unsigned long long* exact_scans,
                                       rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0 || (n == 0 && m == 0)) return 0;
    RFNET_CHECK_ARG(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        // no candidates: the reference's CPU kernel reports (0, 0) for every query of the non-empty side
        if (n) { RFNET_CUDA(cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)b * n, s)); RFNET_CUDA(cudaMemsetAsync(idx1, 0, sizeof(int) * (size_t)b * n, s)); }
        if (m) { RFNET_CUDA(cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)b * m, s)); RFNET_CUDA(cudaMemsetAsync(idx2, 0, sizeof(int) * (size_t)b * m, s)); }
        return 0;
    }
    bool need0 = false, need1 = false;
    { const int rc = nn_search_launch(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, workspace, workspace_bytes, flags, s, &need0, &need1, 3, exact_scans); if (rc) return rc; }
    if (need0 || need1) {
        const size_t count0 = (size_t)b * n;
        const size_t begin = need0 ? 0 : count0, end = need1 ? count0 + (size_t)b * m : count0;
        launch_pdl(nn_unpack_keys_kernel, dim3((unsigned)((end - begin + 255) / 256)), dim3(256), 0, s, (const unsigned long long*)workspace, dist1, idx1, count0, dist2, idx2, begin, end);
    }
    return launch_status();
}

namespace rfnet {
// ---------------------------------------------------------------------------------------------------------------
// One training step of chamfer_big (vv_recon.py:381-385 and its backward) in three launches: the search above, then ONE
// epilogue over all points of both clouds -- unpack the merged keys into dist/idx, the own-point and scattered gradient
// terms of NnDistanceGrad (tf_nndistance_g.cu:131-150, float reductions on zero-filled outputs as there), and the
// per-block partial sums of sqrt(dist) -- and a one-block reduction of those partials in a fixed order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CE_THREADS = 256;
__global__ void __launch_bounds__(CE_THREADS) chamfer_epilogue_kernel(int n, int m, size_t total1, size_t total2, unsigned blocks1, int keyed1, int keyed2,
                                                                      const unsigned long long* __restrict__ keys, const float* __restrict__ xyz1,
                                                                      const float* __restrict__ xyz2, const float* __restrict__ gd1,
                                                                      const float* __restrict__ gd2, float* __restrict__ dist1, int* __restrict__ idx1,
                                                                      float* __restrict__ dist2, int* __restrict__ idx2, float* __restrict__ g1,
                                                                      float* __restrict__ g2, float* __restrict__ partial) {
    __shared__ float sW[CE_THREADS / 32];
    pdl_launch_dependents();
    pdl_wait();   // the search's keys
    const int dir = blockIdx.x >= blocks1;
    const size_t t = (size_t)(blockIdx.x - (dir ? blocks1 : 0)) * CE_THREADS + threadIdx.x;
    const size_t total = dir ? total2 : total1;
    const int nq = dir ? m : n, nc = dir ? n : m;
    float root = 0.f;
    if (t < total) {
        float d;
        int j2;
        float* __restrict__ dist = dir ? dist2 : dist1;
        int* __restrict__ idx = dir ? idx2 : idx1;
        if (dir ? keyed2 : keyed1) {
            const unsigned long long k = keys[(dir ? total1 : 0) + t];
            d = __uint_as_float((unsigned)(k >> 32));
            j2 = (int)(unsigned)(k & 0xffffffffull);
            dist[t] = d;
            idx[t] = j2;
        } else {
            d = dist[t];
            j2 = idx[t];
        }
        root = __fsqrt_rn(d);
        const size_t cloud = t / nq;
        const float g = (dir ? gd2 : gd1)[t] * 2.0f;
        const float* __restrict__ a = (dir ? xyz2 : xyz1) + t * 3;
        const size_t oi = (cloud * nc + j2) * 3;
        const float* __restrict__ o = (dir ? xyz1 : xyz2) + oi;
        float* __restrict__ gown = (dir ? g2 : g1) + t * 3;
        float* __restrict__ goth = (dir ? g1 : g2) + oi;
        const float tx = g * (a[0] - o[0]), ty = g * (a[1] - o[1]), tz = g * (a[2] - o[2]);
        red_add_row3(gown, tx, ty, tz);
        red_add_row3(goth, -tx, -ty, -tz);
    }
    root = warp_sum(root);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = root;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < CE_THREADS / 32; ++i) s += sW[i];
        partial[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(512) chamfer_epilogue_final_kernel(size_t total1, size_t total2, unsigned blocks1, unsigned blocks2,
                                                                     const float* __restrict__ partial, float* __restrict__ sums) {
    // warps 0-7: direction 0, warps 8-15: direction 1; thread-strided partials, then a fixed tree: deterministic
    __shared__ float sW[16];
    pdl_wait();   // the epilogue's partial sums
    const int dir = threadIdx.x >> 8, tid = threadIdx.x & 255;
    const unsigned cnt = dir ? blocks2 : blocks1;
    const float* __restrict__ p = partial + (dir ? blocks1 : 0);
    float t = 0.f;
    for (unsigned i = tid; i < cnt; i += 256) t += p[i];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = t;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += sW[dir * 8 + i];
        sums[dir * 2] = s;
        sums[dir * 2 + 1] = (float)(dir ? total2 : total1);
    }
}

}  // namespace rfnet

namespace rfnet {
// ---------------------------------------------------------------------------------------------------------------
// merge_layer of the reference's model (vv_recon.py:132-139, called with knum = 1): every new point is pulled towards its
// nearest raw point,  out = new + exp(-d2 / (1e-8 + dec^2)) * (raw[nn] - new),  d2 = |raw[nn] - new|^2 summed as the framework
// does ((dx*dx + dy*dy) + dz*dz).  The reference chains NnDistance (both directions), GroupPoint and five framework ops;
// here: ONE directed search (new -> raw) and one epilogue.
// ---------------------------------------------------------------------------------------------------------------
__global__ void merge_layer_kernel(int n_raw, int n_new, size_t total, int keyed, const unsigned long long* __restrict__ keys,
                                   const float* __restrict__ raw, const float* __restrict__ newp, const float* __restrict__ dec,
                                   int* __restrict__ idx, float* __restrict__ out) {
    pdl_wait();   // the search's keys / indices
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int j;
    if (keyed) {
        j = (int)(unsigned)(keys[t] & 0xffffffffull);
        idx[t] = j;
    } else {
        j = idx[t];
    }
    const size_t cloud = t / n_new;
    const float* r = raw + (cloud * n_raw + j) * 3;
    const float* q = newp + t * 3;
    const float dx = __fsub_rn(r[0], q[0]), dy = __fsub_rn(r[1], q[1]), dz = __fsub_rn(r[2], q[2]);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float dc = dec[0];
    const float ratio = expf(-d2 / (1e-8f + dc * dc));
    out[t * 3 + 0] = __fadd_rn(q[0], __fmul_rn(ratio, dx));
    out[t * 3 + 1] = __fadd_rn(q[1], __fmul_rn(ratio, dy));
    out[t * 3 + 2] = __fadd_rn(q[2], __fmul_rn(ratio, dz));
}
// backward of the above for an upstream gradient g (b, n_new, 3), idx held constant (NoGradient of the index, as in the
// reference's graph).  With diff = raw[nn] - new, s = 1e-8 + dec^2, r = exp(-d2/s), a = <g, diff>:
//   g_new      = g (1 - r) + a (2 r / s) diff
//   g_raw_rows = g r - a (2 r / s) diff            (rows to be scatter-added into raw by idx: GroupPointGrad with nsample = 1)
//   g_dec     += a r d2 (2 dec / s^2)              (per-block partial sums, fixed order)
__global__ void __launch_bounds__(256) merge_layer_grad_kernel(int n_raw, int n_new, size_t total, const float* __restrict__ raw,
                                                               const float* __restrict__ newp, const int* __restrict__ idx, const float* __restrict__ dec,
                                                               const float* __restrict__ g, float* __restrict__ g_new, float* __restrict__ g_rows,
                                                               float* __restrict__ dec_partial) {
    __shared__ float sW[8];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float gd = 0.f;
    if (t < total) {
        const size_t cloud = t / n_new;
        const float* r = raw + (cloud * n_raw + idx[t]) * 3;
        const float* q = newp + t * 3;
        const float dx = r[0] - q[0], dy = r[1] - q[1], dz = r[2] - q[2];
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        const float dc = dec[0];
        const float s = 1e-8f + dc * dc;
        const float ratio = expf(-d2 / s);
        const float gx = g[t * 3], gy = g[t * 3 + 1], gz = g[t * 3 + 2];
        const float a = gx * dx + gy * dy + gz * dz;
        const float k = a * (2.0f * ratio / s);
        g_new[t * 3 + 0] = gx * (1.0f - ratio) + k * dx;
        g_new[t * 3 + 1] = gy * (1.0f - ratio) + k * dy;
        g_new[t * 3 + 2] = gz * (1.0f - ratio) + k * dz;
        g_rows[t * 3 + 0] = gx * ratio - k * dx;
        g_rows[t * 3 + 1] = gy * ratio - k * dy;
        g_rows[t * 3 + 2] = gz * ratio - k * dz;
        gd = a * ratio * d2 * (2.0f * dc / (s * s));
    }
    gd = warp_sum(gd);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = gd;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int i = 0; i < 8; ++i) v += sW[i];
        dec_partial[blockIdx.x] = v;
    }
}
}  // namespace rfnet

extern "C" size_t rfnet_merge_layer_workspace_bytes(int b, int n_raw, int n_new) { return rfnet_nn_distance_workspace_bytes(b, n_raw, n_new); }

extern "C" int rfnet_merge_layer(int b, int n_raw, const float* raw, int n_new, const float* newpts, const float* decfactor, float* out, int* idx,
                                 void* workspace, size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n_raw > 0 && n_new >= 0);
    if (b == 0 || n_new == 0) return 0;
    RFNET_CHECK_ARG(raw && newpts && decfactor && out && idx && workspace && workspace_bytes >= rfnet_merge_layer_workspace_bytes(b, n_raw, n_new));
    cudaStream_t s = (cudaStream_t)stream;
    bool need0 = false, need1 = false;
    // direction 2 only: queries = new points, candidates = raw points.  Distances are not needed (the epilogue recomputes d2 the
    // framework's way), so the dist output of the search goes to the front of the output buffer and is overwritten afterwards
    { const int rc = nn_search_launch(b, n_raw, raw, n_new, newpts, nullptr, nullptr, out, idx, workspace, workspace_bytes, 0, s, &need0, &need1, 2); if (rc) return rc; }
    const size_t total = (size_t)b * n_new;
    const unsigned long long* keys = (const unsigned long long*)workspace + (size_t)b * n_raw;
    launch_pdl(merge_layer_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, n_raw, n_new, total, need1 ? 1 : 0, keys, raw, newpts, decfactor, idx, out);
    return launch_status();
}

extern "C" size_t rfnet_merge_layer_grad_partials(int b, int n_new) { return ((size_t)(b > 0 ? b : 0) * (n_new > 0 ? n_new : 0) + 255) / 256; }

extern "C" int rfnet_merge_layer_grad(int b, int n_raw, const float* raw, int n_new, const float* newpts, const int* idx, const float* decfactor,
                                      const float* grad_out, float* grad_new, float* grad_raw_rows, float* dec_partial, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n_raw > 0 && n_new >= 0);
    if (b == 0 || n_new == 0) return 0;
    RFNET_CHECK_ARG(raw && newpts && idx && decfactor && grad_out && grad_new && grad_raw_rows && dec_partial);
    const size_t total = (size_t)b * n_new;
    merge_layer_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_raw, n_new, total, raw, newpts, idx, decfactor, grad_out,
                                                                                               grad_new, grad_raw_rows, dec_partial);
    return launch_status();
}

static size_t chamfer_step_partials(int b, int n, int m) {
    return ((size_t)b * n + CE_THREADS - 1) / CE_THREADS + ((size_t)b * m + CE_THREADS - 1) / CE_THREADS;
}

extern "C" size_t rfnet_chamfer_step_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return rfnet_nn_distance_workspace_bytes(b, n, m) + sizeof(float) * chamfer_step_partials(b, n, m);
}

extern "C" int rfnet_chamfer_step(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1, const float* grad_dist2,
                                  float* dist1, int* idx1, float* dist2, int* idx2, float* grad_xyz1, float* grad_xyz2, float* sums4, void* workspace,
                                  size_t workspace_bytes, int flags, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b > 0 && n > 0 && m > 0);
    RFNET_CHECK_ARG(xyz1 && xyz2 && grad_dist1 && grad_dist2 && dist1 && idx1 && dist2 && idx2 && grad_xyz1 && grad_xyz2 && sums4);
    RFNET_CHECK_ARG(workspace && workspace_bytes >= rfnet_chamfer_step_workspace_bytes(b, n, m));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t t1 = (size_t)b * n, t2 = (size_t)b * m;
    bool need0 = false, need1 = false, zeroed = false;
    const size_t key_bytes = rfnet_nn_distance_workspace_bytes(b, n, m);
    const NNFill zero[2] = {{reinterpret_cast<unsigned*>(grad_xyz1), 3 * t1, 0u}, {reinterpret_cast<unsigned*>(grad_xyz2), 3 * t2, 0u}};
    { const int rc = nn_search_launch(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, workspace, key_bytes, flags, s, &need0, &need1, 3, nullptr, zero, 2, &zeroed); if (rc) return rc; }
    if (!zeroed) {
        RFNET_CUDA(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * 3 * t1, s));
        RFNET_CUDA(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * 3 * t2, s));
    }
    float* partial = reinterpret_cast<float*>((char*)workspace + key_bytes);
    const unsigned blocks1 = (unsigned)((t1 + CE_THREADS - 1) / CE_THREADS), blocks2 = (unsigned)((t2 + CE_THREADS - 1) / CE_THREADS);
    launch_pdl(chamfer_epilogue_kernel, dim3(blocks1 + blocks2), dim3(CE_THREADS), 0, s, n, m, t1, t2, blocks1, need0 ? 1 : 0, need1 ? 1 : 0,
               (const unsigned long long*)workspace, xyz1, xyz2, grad_dist1, grad_dist2, dist1, idx1, dist2, idx2, grad_xyz1, grad_xyz2, partial);
    launch_pdl(chamfer_epilogue_final_kernel, dim3(1), dim3(512), 0, s, t1, t2, blocks1, blocks2, (const float*)partial, sums4);
    return launch_status();
}

extern "C" size_t rfnet_nn_distance_grad_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return sizeof(int) * (size_t)b * ((size_t)n + m) + 64 + seg::csr_bytes(b, n + m, (size_t)n + m);
}

extern "C" int rfnet_nn_distance_grad(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                                      const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1, float* grad_xyz2,
                                      void* workspace, size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0 || (n == 0 && m == 0)) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        if (n) RFNET_CUDA(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * 3 * (size_t)b * n, s));
        if (m) RFNET_CUDA(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * 3 * (size_t)b * m, s));
        return 0;
    }
    RFNET_CHECK_ARG(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && grad_xyz1 && grad_xyz2);
    if (workspace) {
        RFNET_CHECK_ARG(workspace_bytes >= rfnet_nn_distance_grad_workspace_bytes(b, n, m) && b <= 65535);
        int* comb = reinterpret_cast<int*>(workspace);
        const size_t comb_bytes = (sizeof(int) * (size_t)b * ((size_t)n + m) + 63) & ~(size_t)63;
        seg::Csr csr = seg::csr_carve((char*)workspace + comb_bytes, b, n + m, (size_t)n + m);
        dim3 grid((unsigned)(((size_t)n + m + 255) / 256), (unsigned)b);
        nn_grad_combine_idx_kernel<<<grid, 256, 0, s>>>(n, m, idx1, idx2, comb);
        const int rc = seg::csr_build(csr, b, n + m, (size_t)n + m, comb, s);
        if (rc) return rc;
        nn_grad_seg_kernel<<<grid, 256, 0, s>>>(n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, csr.offset, csr.list, grad_xyz1, grad_xyz2);
        return launch_status();
    }
    // no workspace: own terms by plain stores, scattered terms by float reductions (order-dependent in the last bits)
    const size_t t1 = (size_t)b * n, t2 = (size_t)b * m;
    const unsigned grid = (unsigned)((t1 + t2 + 255) / 256);
    nn_grad_own_kernel<<<grid, 256, 0, s>>>(n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, t1, t2);
    nn_grad_scatter_kernel<<<grid, 256, 0, s>>>(n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, t1, t2);
    return launch_status();
}

extern "C" size_t rfnet_chamfer_partial_sums_workspace_bytes(void) { return sizeof(float) * 2 * CS_BLOCKS; }

extern "C" int rfnet_chamfer_partial_sums(int b, int n, int m, const float* dist1, const float* dist2, float* sums4, void* workspace,
                                          size_t workspace_bytes, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && sums4 && workspace && workspace_bytes >= rfnet_chamfer_partial_sums_workspace_bytes());
    RFNET_CHECK_ARG(((size_t)b * n == 0 || dist1) && ((size_t)b * m == 0 || dist2));
    cudaStream_t s = (cudaStream_t)stream;
    chamfer_sums_kernel<<<2 * CS_BLOCKS, 256, 0, s>>>((size_t)b * n, (size_t)b * m, dist1, dist2, (float*)workspace);
    chamfer_sums_final_kernel<<<1, 64, 0, s>>>((size_t)b * n, (size_t)b * m, (const float*)workspace, sums4);
    return launch_status();
}
