// Atomic-free, deterministic "scatter-add by index": the common shape of every gradient in this library
// (gather_point, group_point, three_interpolate, nn_distance all scatter rows of an upstream gradient into the points they
// were gathered from).  The reference does it with float atomicAdd after a memset (tf_sampling_g.cu:183-192,
// tf_grouping_g.cu:61-78, tf_nndistance_g.cu:131-150), so its sums depend on thread timing.
//
// Here the index list is inverted into a CSR ("which source rows point at target t") using INTEGER atomics only -- counts
// and slot claims, whose final state does not depend on order -- every segment is then sorted by source row, and the
// float sums run over each segment in ascending source order: the result is bit-reproducible run to run, and no float
// atomic is issued.  Segments longer than kSegSortMax keep the claim order (still correct, just not order-stable); that
// only happens when one point receives more than 8192 contributions.
// Sorting by segment length: <= 16 entries: insertion sort by one thread; <= 1024: rank sort by one warp through shared
// memory; <= 8192: bitonic network by one CTA in shared memory (O(L log^2 L): ~20 us for 8192 entries -- the first version
// ran an O(L^2) transposition sort in global memory there, several milliseconds when a clustered cloud gave a few points
// thousands of hits each).
#pragma once
#include "common.cuh"

namespace rfnet {
namespace seg {

constexpr int kSegSortMax = 8192;
constexpr int kSegSmem = 1024;  // segments up to this length are rank-sorted through shared memory (8 warps x 4 KiB per CTA)

struct Csr {
    int* offset;      // (b, n+1)  exclusive prefix of the per-target counts
    int* cursor;      // (b, n)    counts, then fill cursors
    int* list;        // (b, R)    source rows grouped by target, ascending inside a segment
    int* long_count;  // (1)       number of queued long segments   (directly before cursor: one memset clears the counters and cursor)
    int* vlong_count; // (1)       number of queued very long segments (kSegSmem < L <= kSegSortMax), directly before long_count
    int2* long_list;  // (b * n)   (cloud, target): long segments from the front, very long ones from the back
};
static inline size_t csr_bytes(int b, int n, size_t R) {
    return sizeof(int) * ((size_t)b * (n + 1) + 4 + (size_t)b * n + (size_t)b * R + 2 * (size_t)b * n) + 64;
}
static inline Csr csr_carve(void* ws, int b, int n, size_t R) {
    Csr c;
    int* p = reinterpret_cast<int*>(ws);
    c.long_list = reinterpret_cast<int2*>(p);  p += 2 * (size_t)b * n;      // 8-byte aligned: first
    c.offset = p;                              p += (size_t)b * (n + 1);
    c.vlong_count = p;                         p += 1;
    c.long_count = p;                          p += 1;
    c.cursor = p;                              p += (size_t)b * n;
    c.list = p;
    (void)R;
    return c;
}

// idx: (b, R) targets in [0, n).  grid (ceil(R/256), b)
static __global__ void csr_count_kernel(int n, unsigned R, const int* __restrict__ idx, int* __restrict__ cursor) {
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const size_t cloud = blockIdx.y;
    atomicAdd(cursor + cloud * n + idx[cloud * R + r], 1);
}

// one CTA of 1024 threads per cloud: exclusive scan of the counts (n arbitrary), cursor := offset.  Each thread owns a
// run of consecutive targets, so there is a single block-wide scan of 1024 partial sums whatever n is.
static __global__ void __launch_bounds__(1024) csr_scan_kernel(int n, int* __restrict__ cursor, int* __restrict__ offset) {
    pdl_wait();   // the previous kernel of the CSR build (programmatic dependent launch: only the launch latency overlaps)
    __shared__ int warp_tot[32];
    const size_t cloud = blockIdx.x;
    int* cnt = cursor + cloud * n;
    int* off = offset + cloud * (n + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (n + 1023) / 1024;
    const int beg = min(n, (int)threadIdx.x * per), end = min(n, beg + per);
    int local = 0;
    for (int i = beg; i < end; ++i) local += cnt[i];
    int x = local;  // inclusive scan of the thread sums inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        warp_tot[lane] = t;  // inclusive totals of the warps
    }
    __syncthreads();
    int run = (warp ? warp_tot[warp - 1] : 0) + x - local;  // exclusive prefix of this thread's run
    for (int i = beg; i < end; ++i) {
        const int v = cnt[i];
        off[i] = run;
        cnt[i] = run;  // becomes the fill cursor
        run += v;
    }
    if (threadIdx.x == 1023) off[n] = warp_tot[31];
}

static __global__ void csr_fill_kernel(int n, unsigned R, const int* __restrict__ idx, int* __restrict__ cursor, int* __restrict__ list) {
    pdl_wait();   // the previous kernel of the CSR build (programmatic dependent launch: only the launch latency overlaps)
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const size_t cloud = blockIdx.y;
    const int pos = atomicAdd(cursor + cloud * n + idx[cloud * R + r], 1);
    list[cloud * R + pos] = (int)r;
}

// Sort every segment ascending (entries are distinct source rows).  Short segments -- the overwhelming majority: a point is
// typically referenced a handful of times -- are insertion-sorted by ONE THREAD per target; longer ones are queued
// (integer atomic on a counter) for the warp-per-segment kernel below.  grid (ceil(n/256), b)
constexpr int kSegThread = 16;
static __global__ void csr_sort_short_kernel(int n, unsigned R, size_t queue_len, const int* __restrict__ offset, int* __restrict__ list,
                                             int* __restrict__ long_count, int* __restrict__ vlong_count, int2* __restrict__ long_list) {
    pdl_wait();   // the previous kernel of the CSR build (programmatic dependent launch: only the launch latency overlaps)
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)n) return;
    const size_t cloud = blockIdx.y;
    const int beg = offset[cloud * (n + 1) + t], end = offset[cloud * (n + 1) + t + 1];
    const int L = end - beg;
    if (L <= 1) return;
    int* seg = list + cloud * R + beg;
    if (L <= kSegThread) {
        for (int i = 1; i < L; ++i) {
            const int v = seg[i];
            int j = i - 1;
            while (j >= 0 && seg[j] > v) { seg[j + 1] = seg[j]; --j; }
            seg[j + 1] = v;
        }
        return;
    }
    if (L <= kSegSmem) long_list[atomicAdd(long_count, 1)] = make_int2((int)cloud, (int)t);
    else if (L <= kSegSortMax) long_list[queue_len - 1 - (size_t)atomicAdd(vlong_count, 1)] = make_int2((int)cloud, (int)t);
}

// persistent warps over the queue of long segments.  The rank of an entry is the number of smaller entries.
static __global__ void __launch_bounds__(256) csr_sort_long_kernel(int n, unsigned R, const int* __restrict__ offset, int* __restrict__ list,
                                                                   const int* __restrict__ long_count, const int2* __restrict__ long_list) {
    pdl_wait();   // the previous kernel of the CSR build (programmatic dependent launch: only the launch latency overlaps)
    __shared__ int seg_smem[8 * kSegSmem];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = *long_count;
    for (int w = blockIdx.x * 8 + warp; w < total; w += gridDim.x * 8) {
        const int2 ct = long_list[w];
        const size_t cloud = (size_t)ct.x;
        const int beg = offset[cloud * (n + 1) + ct.y], end = offset[cloud * (n + 1) + ct.y + 1];
        const int L = end - beg;
        int* seg = list + cloud * R + beg;
        int* s = seg_smem + warp * kSegSmem;   // L <= kSegSmem by construction of the queue
        for (int i = lane; i < L; i += 32) s[i] = seg[i];
        __syncwarp();
        for (int i = lane; i < L; i += 32) {
            const int v = s[i];
            int rank = 0;
            for (int o = 0; o < L; ++o) rank += (s[o] < v) ? 1 : 0;
            seg[rank] = v;
        }
        __syncwarp();
    }
}

// one CTA per very long segment (kSegSmem < L <= kSegSortMax): bitonic network over shared memory, padded with INT_MAX
static __global__ void __launch_bounds__(512) csr_sort_vlong_kernel(int n, unsigned R, size_t queue_len, const int* __restrict__ offset, int* __restrict__ list,
                                                                    const int* __restrict__ vlong_count, const int2* __restrict__ long_list) {
    pdl_wait();   // the previous kernel of the CSR build (programmatic dependent launch: only the launch latency overlaps)
    __shared__ int s[kSegSortMax];
    const int total = *vlong_count;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int2 ct = long_list[queue_len - 1 - (size_t)w];
        const size_t cloud = (size_t)ct.x;
        const int beg = offset[cloud * (n + 1) + ct.y], end = offset[cloud * (n + 1) + ct.y + 1];
        const int L = end - beg;
        int* seg = list + cloud * R + beg;
        int P = 1;
        while (P < L) P <<= 1;
        __syncthreads();
        for (int i = threadIdx.x; i < P; i += blockDim.x) s[i] = i < L ? seg[i] : 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += blockDim.x) {
                    const int x = i ^ j;
                    if (x > i) {
                        const int a = s[i], b2 = s[x];
                        const bool up = (i & k) == 0;
                        if ((a > b2) == up) { s[i] = b2; s[x] = a; }
                    }
                }
                __syncthreads();
            }
        for (int i = threadIdx.x; i < L; i += blockDim.x) seg[i] = s[i];
    }
}

// Builds the CSR for idx (b, R) -> targets [0, n).  All work is stream-ordered; `ws` must hold csr_bytes(b, n, R).
static inline int csr_build(Csr c, int b, int n, size_t R, const int* idx, cudaStream_t s) {
    if (b == 0 || n == 0) return 0;
    RFNET_CUDA(cudaMemsetAsync(c.vlong_count, 0, sizeof(int) * ((size_t)b * n + 2), s));  // vlong_count, long_count and cursor are adjacent
    if (R) {
        dim3 g((unsigned)((R + 255) / 256), (unsigned)b);
        csr_count_kernel<<<g, 256, 0, s>>>(n, (unsigned)R, idx, c.cursor);
    }
    launch_pdl(csr_scan_kernel, dim3(b), dim3(1024), 0, s, n, c.cursor, c.offset);
    if (R) {
        dim3 g((unsigned)((R + 255) / 256), (unsigned)b);
        launch_pdl(csr_fill_kernel, g, dim3(256), 0, s, n, (unsigned)R, idx, c.cursor, c.list);
        dim3 gs((unsigned)((n + 255) / 256), (unsigned)b);
        const size_t queue_len = (size_t)b * n;
        launch_pdl(csr_sort_short_kernel, gs, dim3(256), 0, s, n, (unsigned)R, queue_len, (const int*)c.offset, c.list, c.long_count, c.vlong_count, c.long_list);
        launch_pdl(csr_sort_long_kernel, dim3(num_sms()), dim3(256), 0, s, n, (unsigned)R, (const int*)c.offset, c.list, c.long_count, c.long_list);
        launch_pdl(csr_sort_vlong_kernel, dim3(num_sms()), dim3(512), 0, s, n, (unsigned)R, queue_len, (const int*)c.offset, c.list, c.vlong_count, c.long_list);
    }
    return launch_status();
}

}  // namespace seg
}  // namespace rfnet
