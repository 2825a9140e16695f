// auction_match: 1:1 assignment of the points of xyz1 (bidders) to the points of xyz2 (objects) by Bertsekas' auction,
// as the reference's AuctionMatchKernel runs it (tf_ops/emd/tf_auctionmatch_g.cu:2-291): ONE bidder at a time, taken from
// a FIFO of unassigned bidders; the bid goes to the object minimising value = |p_i - q_j| + price[j]; the object's price
// rises by (second best - best + tolerance); a displaced owner re-enters the FIFO; after 40 n bids the tolerance grows
// 1e-4 -> 1e-2 -> 1 and after 40 n bids at tolerance 1 the auction stops.  The bid sequence is strictly sequential, so the
// result is a deterministic function of the inputs and this kernel reproduces it bid for bid.
//
// "Second best" as the reference computes it.  The reference merges its 512 per-thread (best, second best, object)
// triples with a shuffle-down tree whose losing branch reads `best` after overwriting it (tf_auctionmatch_g.cu:226-228,
// 249-251: best=b1; best2=fminf(best,b2)), so the second best that reaches the bid equals the best unless the winning
// object belongs to thread 0 of the block (objects 0, 512, 1024, ...); only then is it the true runner-up.  The price
// increment is therefore exactly `tolerance` for all other objects.  Matching the reference's assignments means matching
// this, so the rule is implemented as stated (oracle/rfnet_oracle.c emulates the tree literally and agrees bit for bit
// with the reference kernel's output, tests/golden/ref_gpu2.npz).  The same tree fixes which object wins among EQUAL
// values: the higher object inside a thread, then the lane (and then the warp) whose index is largest when read with its
// bits reversed; tie_pick() below.
//
// What is different from the reference kernel (one 512-thread block per cloud over an n x n cost matrix in global memory,
// two block barriers and a cost-row read per bid):
//   * no cost matrix: every thread keeps its R objects (coordinates, price, owner) in REGISTERS and evaluates the R
//     distances of a bid on the fly -- sqrt.rn(fma(dz,dz, fma(dx,dx, dy*dy))), the reference binary's exact expression
//     (SASS of tf_auctionmatch_g.cu:36 compiled for sm_100a) -- so `cost` (b*n*n floats, tf_auctionmatch.cpp:54) is gone
//     and a bid touches no global memory at all;
//   * values are non-negative floats, so the warp-level best / second-best are two REDUX.MIN on the raw bits and a ballot
//     instead of a 5-step shuffle tree carrying three registers;
//   * the per-warp results are double-buffered in shared memory and every thread reduces them redundantly and advances
//     the queue state in registers: ONE block barrier per bid instead of two;
//   * one CTA per cloud on its own SM (the reference caps the grid at 32 blocks).
// Objects are laid out over the threads as in the reference (object j belongs to thread j % 512), which is what makes the
// two rules above reproducible.  For 1024 <= n < 4096 other than 1024 and 2048 the reference reads past its rows
// (its 2- and 4-wide loops have no bound check); those n are handled like the others here.
#include "common.cuh"
#include "../../include/rfnet_ops.h"

namespace rfnet {

constexpr int AU_THREADS = 512;
constexpr int AU_WARPS = AU_THREADS / 32;
constexpr float AU_BIG = 1e38f;  // the reference's initial best / second best (tf_auctionmatch_g.cu:55)

struct __align__(16) AuBest {
    unsigned best, runner;  // raw bits of non-negative floats: unsigned order == float order
    int j, owner;
};

// Winner among lanes (or warps) holding the same value, as the shuffle-down tree decides it: the last merge (distance 1)
// prefers the odd side, the one before (distance 2) the side with bit 1 set, and so on.
__device__ __forceinline__ int tie_pick(unsigned mask) {
    if (mask & (mask - 1)) {
        if (mask & 0xAAAAAAAAu) mask &= 0xAAAAAAAAu;
        if (mask & 0xCCCCCCCCu) mask &= 0xCCCCCCCCu;
        if (mask & 0xF0F0F0F0u) mask &= 0xF0F0F0F0u;
        if (mask & 0xFF00FF00u) mask &= 0xFF00FF00u;
        if (mask & 0xFFFF0000u) mask &= 0xFFFF0000u;
    }
    return __ffs(mask) - 1;
}

template <int R>
__global__ void __launch_bounds__(AU_THREADS, 1) auction_kernel(int n, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                                int* __restrict__ matchl, int* __restrict__ matchr) {
    extern __shared__ __align__(16) unsigned char au_smem[];
    float* sP = reinterpret_cast<float*>(au_smem);                    // bidders' points, AoS (n * 3)
    int* queue = reinterpret_cast<int*>(sP + (size_t)n * 3);          // FIFO of unassigned bidders (n)
    __shared__ AuBest sBest[2][AU_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t cloud = blockIdx.x;
    const float* __restrict__ p1 = xyz1 + cloud * n * 3;
    const float* __restrict__ p2 = xyz2 + cloud * n * 3;
    for (int i = tid; i < n * 3; i += AU_THREADS) sP[i] = p1[i];
    for (int i = tid; i < n; i += AU_THREADS) {
        queue[i] = i;
        matchl[cloud * n + i] = -1;
    }
    // objects j = tid + AU_THREADS * r live in registers
    float ox[R], oy[R], oz[R], price[R];
    int owner[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = tid + AU_THREADS * r;
        const bool v = j < n;
        ox[r] = v ? p2[j * 3] : 0.f;
        oy[r] = v ? p2[j * 3 + 1] : 0.f;
        oz[r] = v ? p2[j * 3 + 2] : 0.f;
        price[r] = 0.f;
        owner[r] = -1;
    }
    __syncthreads();

    int qhead = 0, qlen = n, cnt = 0, cur = 0;
    const int cnt_max = 40 * n;
    float tolerance = 1e-4f;
    unsigned it = 0;
    while (qlen) {
        const float px = sP[cur * 3], py = sP[cur * 3 + 1], pz = sP[cur * 3 + 2];
        float best = AU_BIG, best2 = AU_BIG;
        int bestr = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float d = __fsqrt_rn(sqdist3<true>(px - ox[r], py - oy[r], pz - oz[r]));
            const float v = __fadd_rn(d, price[r]);
            if (tid + AU_THREADS * r < n) {          // an equal value replaces the incumbent (tf_auctionmatch_g.cu:205-212)
                if (best < v) {
                    best2 = fminf(best2, v);
                } else {
                    best2 = best;
                    best = v;
                    bestr = r;
                }
            }
        }
        // warp: best = min over lanes.  runner = what the bid uses as second best IF thread 0 wins: thread 0's own second
        // best against everyone else's best.
        const unsigned ub = __float_as_uint(best);
        const unsigned wb = __reduce_min_sync(0xffffffffu, ub);
        const int wl = tie_pick(__ballot_sync(0xffffffffu, ub == wb));
        const unsigned wr = __reduce_min_sync(0xffffffffu, tid == 0 ? __float_as_uint(best2) : ub);
        const unsigned par = it & 1u;
        if (lane == wl) {
            int own = owner[0];
#pragma unroll
            for (int r = 1; r < R; ++r) own = bestr == r ? owner[r] : own;
            AuBest e;
            e.best = wb;
            e.runner = wr;
            e.j = tid + AU_THREADS * bestr;
            e.owner = own;
            sBest[par][warp] = e;
        }
        __syncthreads();
        // every thread reduces the AU_WARPS entries (lane l reads entry l % AU_WARPS: duplicates do not change a min)
        const AuBest e = sBest[par][lane & (AU_WARPS - 1)];
        const unsigned gb = __reduce_min_sync(0xffffffffu, e.best);
        const int gw = tie_pick(__ballot_sync(0xffffffffu, e.best == gb) & ((1u << AU_WARPS) - 1u));   // lanes 0..15 hold entries 0..15
        const unsigned gr = __reduce_min_sync(0xffffffffu, e.runner);
        const int bestj = __shfl_sync(0xffffffffu, e.j, gw);
        const int old = __shfl_sync(0xffffffffu, e.owner, gw);
        const unsigned gb2 = (bestj & (AU_THREADS - 1)) == 0 ? gr : gb;   // see "second best" above
        const float delta = __fadd_rn(__fsub_rn(__uint_as_float(gb2), __uint_as_float(gb)), tolerance);   // tf_auctionmatch_g.cu:259
        // the owner thread of bestj raises its price and records the new owner
        if ((bestj & (AU_THREADS - 1)) == tid) {
            const int rr = bestj / AU_THREADS;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r == rr) {
                    price[r] = __fadd_rn(price[r], delta);
                    owner[r] = cur;
                }
            }
        }
        // queue bookkeeping, redundantly in every thread (tf_auctionmatch_g.cu:260-280)
        qhead++;
        qlen--;
        if (qhead >= n) qhead -= n;
        cnt++;
        int next = -1;
        if (old != -1) {
            int tail = qhead + qlen;
            if (tail >= n) tail -= n;
            if (qlen == 0) next = old;          // the displaced owner is the only bidder left: it bids next
            qlen++;
            if (tid == 0) queue[tail] = old;    // read by others at the earliest after the next barrier
        }
        if (cnt == cnt_max) {
            if (tolerance == 1.0f) qlen = 0;
            tolerance = fminf(1.0f, tolerance * 100);
            cnt = 0;
        }
        if (qlen) cur = next >= 0 ? next : queue[qhead];
        ++it;
    }
    // every object's owner -> matchr; inverse -> matchl (bidders left without an object keep -1)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = tid + AU_THREADS * r;
        if (j < n) {
            matchr[cloud * n + j] = owner[r];
            if (owner[r] >= 0) matchl[cloud * n + owner[r]] = j;
        }
    }
}

template <int R>
static int launch_auction(int b, int n, const float* xyz1, const float* xyz2, int* matchl, int* matchr, cudaStream_t s) {
    const size_t smem = (size_t)n * 16;
    if (smem > 48 * 1024) RFNET_CUDA(cudaFuncSetAttribute(auction_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auction_kernel<R><<<b, AU_THREADS, smem, s>>>(n, xyz1, xyz2, matchl, matchr);
    return launch_status();
}

}  // namespace rfnet

using namespace rfnet;

extern "C" int rfnet_auction_match(int b, int n, const float* xyz1, const float* xyz2, int* matchl, int* matchr, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && n <= RFNET_AUCTION_MAX_POINTS);
    if (b == 0 || n == 0) return 0;
    RFNET_CHECK_ARG(xyz1 && xyz2 && matchl && matchr);
    cudaStream_t s = (cudaStream_t)stream;
    const int per = (n + AU_THREADS - 1) / AU_THREADS;
    if (per <= 1) return launch_auction<1>(b, n, xyz1, xyz2, matchl, matchr, s);
    if (per <= 2) return launch_auction<2>(b, n, xyz1, xyz2, matchl, matchr, s);
    if (per <= 4) return launch_auction<4>(b, n, xyz1, xyz2, matchl, matchr, s);
    if (per <= 8) return launch_auction<8>(b, n, xyz1, xyz2, matchl, matchr, s);
    return launch_auction<16>(b, n, xyz1, xyz2, matchl, matchr, s);
}
