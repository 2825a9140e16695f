// Morton (Z-curve) ordering of point clouds: one CTA per cloud, 15-bit codes (5 bits per axis over the cloud's bounding box)
// packed with the 15-bit point index into 32-bit keys, bitonic sort in shared memory.  Used to give warps spatially tight
// sets of points (approx_match's pruned sweeps, the pruned FPS).  Up to MORTON_SORT_MAX points per cloud.
#pragma once
#include "common.cuh"

namespace rfnet {

constexpr int MORTON_SORT_MAX = 32768;

__device__ __forceinline__ unsigned morton_spread5(unsigned v) {  // 5 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8);
}

// grid = (clouds, 2): y = 0 sorts xyz1 (n points) into perm1, y = 1 sorts xyz2 (m points) into perm2.
static __global__ void __launch_bounds__(1024) morton_sort_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                               int* __restrict__ perm1, int* __restrict__ perm2) {
    extern __shared__ unsigned sort_keys[];
    __shared__ float red[6][32];
    const int np = blockIdx.y == 0 ? n : m;
    const float* __restrict__ pts = (blockIdx.y == 0 ? xyz1 : xyz2) + (size_t)blockIdx.x * np * 3;
    int* __restrict__ perm = (blockIdx.y == 0 ? perm1 : perm2) + (size_t)blockIdx.x * np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int npad = 1;
    while (npad < np) npad <<= 1;
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
    for (int i = tid; i < npad; i += 1024) {
        unsigned key = 0xffffffffu;
        if (i < np) {
            unsigned c[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) c[a] = min(31u, (unsigned)fmaxf(0.f, (pts[(size_t)i * 3 + a] - lo[a]) * scale[a]));
            key = ((morton_spread5(c[0]) | (morton_spread5(c[1]) << 1) | (morton_spread5(c[2]) << 2)) << 15) | (unsigned)i;
        }
        sort_keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += 1024) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
                const int p2 = i | j;
                const unsigned a = sort_keys[i], c = sort_keys[p2];
                const bool up = (i & k) == 0;
                if ((a > c) == up) { sort_keys[i] = c; sort_keys[p2] = a; }
            }
            __syncthreads();
        }
    for (int i = tid; i < np; i += 1024) perm[i] = (int)(sort_keys[i] & 0x7fffu);
}


// dynamic shared memory the sort needs for clouds of up to `nmax` points
static inline size_t morton_sort_smem(int nmax) {
    int np2 = 1;
    while (np2 < nmax) np2 <<= 1;
    return (size_t)np2 * sizeof(unsigned);
}
// perm1[cloud][i] = index of the i-th point of xyz1 along the curve (and perm2 / xyz2 / m when xyz2 != nullptr)
static inline int morton_sort(int b, int n, int m, const float* xyz1, const float* xyz2, int* perm1, int* perm2, cudaStream_t s) {
    const size_t smem = morton_sort_smem(xyz2 ? (n > m ? n : m) : n);
    if (smem > 40 * 1024) RFNET_CUDA(cudaFuncSetAttribute(morton_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    morton_sort_kernel<<<dim3((unsigned)b, xyz2 ? 2u : 1u), 1024, smem, s>>>(n, m, xyz1, xyz2, perm1, perm2);
    return launch_status();
}

}  // namespace rfnet
