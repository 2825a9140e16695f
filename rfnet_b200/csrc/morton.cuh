// Space-filling-curve ordering of point clouds: one CTA per cloud, 15-bit cell codes (5 bits per axis over the cloud's bounding
// box), counting sort in shared memory.  Used to give warps spatially tight sets of points (approx_match's pruned sweeps,
// the pruned FPS).
// The curve is HILBERT's (Skilling's transpose algorithm, 5 bits x 3 axes), not the Z-curve the file is named after: consecutive
// Hilbert cells are always face neighbours, so ANY window of 64 consecutive cells has a bounding box of at most 128 cells
// (mean 105), whereas a Z-curve window that straddles an octant boundary spans up to 5120 cells (mean 273, 99th percentile
// 1623).  A pruned kernel waits for its slowest warp, so the tail is what matters: with Z-order the clusters that straddle
// a jump kept nearly every candidate.  Every consumer is exact for ANY permutation; only speed depends on the curve.
#pragma once
#include "common.cuh"

namespace rfnet {

constexpr int MORTON_SORT_MAX = 32768;

__host__ __device__ __forceinline__ unsigned morton_spread5(unsigned v) {  // 5 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8);
}

// Hilbert index of a cell (x, y, z), 5 bits per axis: J. Skilling, "Programming the Hilbert curve" (2004), AxesToTranspose,
// followed by the bit interleave of the transposed form (x most significant inside each triple).
__host__ __device__ __forceinline__ unsigned hilbert_code5(unsigned x, unsigned y, unsigned z) {
    unsigned X[3] = {x, y, z};
#pragma unroll
    for (unsigned Q = 16u; Q > 1u; Q >>= 1) {
        const unsigned P = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const unsigned t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    unsigned t = 0u;
#pragma unroll
    for (unsigned Q = 16u; Q > 1u; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1u;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (morton_spread5(X[0]) << 2) | (morton_spread5(X[1]) << 1) | morton_spread5(X[2]);
}

// grid = (clouds, 2): y = 0 orders xyz1 (n points) into perm1, y = 1 orders xyz2 (m points) into perm2.
// Counting sort on the 15-bit cell code: histogram of the 32768 cells with shared-memory integer atomics, block-wide
// exclusive scan (every thread owns 32 consecutive cells), scatter through per-cell cursors, then each thread puts the few
// points of each of its cells in ascending index order, so the permutation does not depend on the order in which the
// atomics landed (cells holding more than MORTON_CELL_SORT points keep the arrival order; nothing downstream depends on
// the order inside a cell -- the pruned kernels are exact for ANY permutation -- it only keeps runs reproducible).
// ~10x faster than the bitonic network this replaced (113 us for 16384 points).
constexpr int MORTON_CELLS = 32768;
constexpr int MORTON_CELL_SORT = 64;
static __global__ void __launch_bounds__(1024) morton_sort_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                               int* __restrict__ perm1, int* __restrict__ perm2) {
    extern __shared__ unsigned cell[];   // MORTON_CELLS counters, then cursors; one pad word per 32 (see CI)
    __shared__ float red[6][32];
    __shared__ unsigned wtot[32];
    const int np = blockIdx.y == 0 ? n : m;
    const float* __restrict__ pts = (blockIdx.y == 0 ? xyz1 : xyz2) + (size_t)blockIdx.x * np * 3;
    int* __restrict__ perm = (blockIdx.y == 0 ? perm1 : perm2) + (size_t)blockIdx.x * np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto CI = [](unsigned c) -> unsigned { return c + (c >> 5); };   // thread t owns cells 32t..32t+31: stride 33 words, conflict-free
    for (int i = tid; i < MORTON_CELLS + MORTON_CELLS / 32; i += 1024) cell[i] = 0u;
    // bounding box
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = tid; i < np; i += 1024)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pts[(size_t)i * 3 + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
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
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = red[a][0], h = red[3 + a][0];
        for (int w2 = 1; w2 < 32; ++w2) { l = fminf(l, red[a][w2]); h = fmaxf(h, red[3 + a][w2]); }
        lo[a] = l;
        scale[a] = h > l ? 31.999f / (h - l) : 0.f;
    }
    auto code = [&](int i) -> unsigned {
        unsigned c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = min(31u, (unsigned)fmaxf(0.f, (pts[(size_t)i * 3 + a] - lo[a]) * scale[a]));
        return hilbert_code5(c[0], c[1], c[2]);
    };
    for (int i = tid; i < np; i += 1024) atomicAdd(&cell[CI(code(i))], 1u);
    __syncthreads();
    // exclusive scan over the cells in code order: thread t owns cells [32 t, 32 t + 32)
    unsigned sum = 0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) sum += cell[CI(tid * 32 + k)];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned t = wtot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        wtot[lane] = t;
    }
    __syncthreads();
    const unsigned base = (warp ? wtot[warp - 1] : 0u) + incl - sum;   // first position of this thread's first cell
    {
        unsigned run = base;   // counters -> cursors, in place (a thread only touches its own cells)
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const unsigned c = cell[CI(tid * 32 + k)];
            cell[CI(tid * 32 + k)] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int i = tid; i < np; i += 1024) perm[atomicAdd(&cell[CI(code(i))], 1u)] = i;
    __syncthreads();
    // ascending index order inside every cell: after the scatter a cursor is the END of its cell = the start of the next
    int beg = (int)base;
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
        const int end = (int)cell[CI(tid * 32 + k)];
        const int L = end - beg;
        if (L >= 2 && L <= MORTON_CELL_SORT) {
            for (int a = beg + 1; a < end; ++a) {
                const int v = perm[a];
                int q = a - 1;
                while (q >= beg && perm[q] > v) { perm[q + 1] = perm[q]; --q; }
                perm[q + 1] = v;
            }
        }
        beg = end;
    }
}

// dynamic shared memory the sort needs for clouds of up to `nmax` points
static inline size_t morton_sort_smem(int nmax) {
    (void)nmax;
    return (size_t)(MORTON_CELLS + MORTON_CELLS / 32) * sizeof(unsigned);
}
// perm1[cloud][i] = index of the i-th point of xyz1 along the curve (and perm2 / xyz2 / m when xyz2 != nullptr)
static inline int morton_sort(int b, int n, int m, const float* xyz1, const float* xyz2, int* perm1, int* perm2, cudaStream_t s) {
    const size_t smem = morton_sort_smem(xyz2 ? (n > m ? n : m) : n);
    if (smem > 40 * 1024) RFNET_CUDA(cudaFuncSetAttribute(morton_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    morton_sort_kernel<<<dim3((unsigned)b, xyz2 ? 2u : 1u), 1024, smem, s>>>(n, m, xyz1, xyz2, perm1, perm2);
    return launch_status();
}

}  // namespace rfnet
