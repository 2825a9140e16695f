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
#pragma once
#include "common.cuh"

namespace rfnet {
namespace seg {

constexpr int kSegSortMax = 8192;
constexpr int kSegSmem = 1024;  // segments up to this length are rank-sorted through shared memory (8 warps x 4 KiB per CTA)

struct Csr {
    int* offset;  // (b, n+1)  exclusive prefix of the per-target counts
    int* cursor;  // (b, n)    counts, then fill cursors
    int* list;    // (b, R)    source rows grouped by target, ascending inside a segment
};
static inline size_t csr_bytes(int b, int n, size_t R) {
    return sizeof(int) * ((size_t)b * (n + 1) + (size_t)b * n + (size_t)b * R) + 64;
}
static inline Csr csr_carve(void* ws, int b, int n, size_t R) {
    Csr c;
    c.offset = reinterpret_cast<int*>(ws);
    c.cursor = c.offset + (size_t)b * (n + 1);
    c.list = c.cursor + (size_t)b * n;
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

// one CTA of 1024 threads per cloud: exclusive scan of the counts (n arbitrary), cursor := offset
static __global__ void __launch_bounds__(1024) csr_scan_kernel(int n, int* __restrict__ cursor, int* __restrict__ offset) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const size_t cloud = blockIdx.x;
    int* cnt = cursor + cloud * n;
    int* off = offset + cloud * (n + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? cnt[i] : 0;
        int x = v;  // inclusive scan inside the warp
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
        const int carry = carry_s;
        const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + x - v;
        if (i < n) {
            off[i] = excl;
            cnt[i] = excl;  // becomes the fill cursor
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = carry_s;
}

static __global__ void csr_fill_kernel(int n, unsigned R, const int* __restrict__ idx, int* __restrict__ cursor, int* __restrict__ list) {
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const size_t cloud = blockIdx.y;
    const int pos = atomicAdd(cursor + cloud * n + idx[cloud * R + r], 1);
    list[cloud * R + pos] = (int)r;
}

// one warp per target: sort its segment ascending.  Entries are distinct, so the rank of an entry is the number of
// smaller entries.  grid (ceil(n/8), b), 256 threads; dynamic smem = 8 warps * kSegSmem ints
static __global__ void __launch_bounds__(256) csr_sort_kernel(int n, unsigned R, const int* __restrict__ offset, int* __restrict__ list) {
    extern __shared__ int seg_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp;
    if (t >= n) return;
    const size_t cloud = blockIdx.y;
    const int beg = offset[cloud * (n + 1) + t], end = offset[cloud * (n + 1) + t + 1];
    const int L = end - beg;
    if (L <= 1 || L > kSegSortMax) return;
    int* seg = list + cloud * R + beg;
    if (L <= 32) {
        const int v = lane < L ? seg[lane] : 0x7fffffff;
        int rank = 0;
#pragma unroll 8
        for (int o = 0; o < 32; ++o) rank += (__shfl_sync(0xffffffffu, v, o) < v) ? 1 : 0;
        __syncwarp();
        if (lane < L) seg[rank] = v;
        return;
    }
    if (L <= kSegSmem) {
        int* s = seg_smem + warp * kSegSmem;
        for (int i = lane; i < L; i += 32) s[i] = seg[i];
        __syncwarp();
        for (int i = lane; i < L; i += 32) {
            const int v = s[i];
            int rank = 0;
            for (int o = 0; o < L; ++o) rank += (s[o] < v) ? 1 : 0;
            seg[rank] = v;
        }
        return;
    }
    // long segment (kSegSmem < L <= kSegSortMax): in-place odd-even transposition sort over global memory, cooperative
    // across the warp -- O(L^2 / 64) steps per lane, only reached when one point collects thousands of contributions
    for (int pass = 0; pass < L; ++pass) {
        for (int i = (pass & 1) + 2 * lane; i + 1 < L; i += 64) {
            const int a = seg[i], b2 = seg[i + 1];
            if (a > b2) { seg[i] = b2; seg[i + 1] = a; }
        }
        __syncwarp();
    }
}

// Builds the CSR for idx (b, R) -> targets [0, n).  All work is stream-ordered; `ws` must hold csr_bytes(b, n, R).
static inline int csr_build(Csr c, int b, int n, size_t R, const int* idx, cudaStream_t s) {
    if (b == 0 || n == 0) return 0;
    RFNET_CUDA(cudaMemsetAsync(c.cursor, 0, sizeof(int) * (size_t)b * n, s));
    if (R) {
        dim3 g((unsigned)((R + 255) / 256), (unsigned)b);
        csr_count_kernel<<<g, 256, 0, s>>>(n, (unsigned)R, idx, c.cursor);
    }
    csr_scan_kernel<<<b, 1024, 0, s>>>(n, c.cursor, c.offset);
    if (R) {
        dim3 g((unsigned)((R + 255) / 256), (unsigned)b);
        csr_fill_kernel<<<g, 256, 0, s>>>(n, (unsigned)R, idx, c.cursor, c.list);
        dim3 gs((unsigned)((n + 7) / 8), (unsigned)b);
        csr_sort_kernel<<<gs, 256, 8 * kSegSmem * sizeof(int), s>>>(n, (unsigned)R, c.offset, c.list);
    }
    return launch_status();
}

}  // namespace seg
}  // namespace rfnet
