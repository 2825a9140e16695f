// Uniform grid over a point cloud, built by one CTA per cloud with a counting sort (histogram of <= 32^3 cells in shared
// memory, block-wide exclusive scan, scatter): cell offsets, the permutation and the coordinates in cell order go to a
// caller-provided workspace.  Used by the grid variants of query_ball_point (grouping.cu) and three_nn (interpolate.cu).
#pragma once
#include "common.cuh"

namespace rfnet {

constexpr int BG_CELLS = 32768;        // 32^3

struct BallGrid {
    float lo[3], scale[3];   // cell coordinate along axis a: min(G[a] - 1, (int)((p[a] - lo[a]) * scale[a]))
    int G[3], pad;
};
static inline size_t ball_grid_stride(int n) {   // bytes of workspace per cloud
    return (sizeof(BallGrid) + sizeof(int) * (BG_CELLS + 1) + sizeof(int) * (size_t)n + sizeof(float) * 3 * (size_t)n + 63) & ~(size_t)63;
}
struct BallGridView {
    const BallGrid* g;
    const int* cell_start;     // BG_CELLS + 1
    const int* sorted_idx;     // n: dataset indices in cell order
    const float* sorted_xyz;   // n x 3 in cell order
};
__device__ __forceinline__ BallGridView ball_grid_view(const unsigned char* ws, size_t stride, int cloud, int n) {
    const unsigned char* p = ws + stride * cloud;
    BallGridView v;
    v.g = reinterpret_cast<const BallGrid*>(p);                p += sizeof(BallGrid);
    v.cell_start = reinterpret_cast<const int*>(p);            p += sizeof(int) * (BG_CELLS + 1);
    v.sorted_idx = reinterpret_cast<const int*>(p);            p += sizeof(int) * (size_t)n;
    v.sorted_xyz = reinterpret_cast<const float*>(p);
    return v;
}

// radius != nullptr: cells at least one radius wide (ball query).  radius == nullptr: about two points per cell (nearest-
// neighbour search).  At most gmax cells per axis (gmax <= 32).
static __global__ void __launch_bounds__(1024) ball_grid_kernel(int n, const float* __restrict__ radius, int gmax, const float* __restrict__ xyz1,
                                                                unsigned char* __restrict__ ws, size_t stride) {
    extern __shared__ unsigned cell[];   // BG_CELLS counters, then cursors; one pad word per 32 (CI)
    __shared__ float red[6][32];
    __shared__ unsigned wtot[32];
    __shared__ BallGrid sg;
    const int cloud = blockIdx.x;
    const float* __restrict__ pts = xyz1 + (size_t)cloud * n * 3;
    const BallGridView v = ball_grid_view(ws, stride, cloud, n);
    BallGrid* gout = const_cast<BallGrid*>(v.g);
    int* cell_start = const_cast<int*>(v.cell_start);
    int* sorted_idx = const_cast<int*>(v.sorted_idx);
    float* sorted_xyz = const_cast<float*>(v.sorted_xyz);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto CI = [](unsigned c) -> unsigned { return c + (c >> 5); };
    for (int i = tid; i < BG_CELLS + BG_CELLS / 32; i += 1024) cell[i] = 0u;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = tid; i < n; i += 1024)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float x = pts[(size_t)i * 3 + a];
            lo[a] = fminf(lo[a], x);
            hi[a] = fmaxf(hi[a], x);
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
    if (tid == 0) {
        float l3[3], e3[3];
        for (int a = 0; a < 3; ++a) {
            float l = red[a][0], h = red[3 + a][0];
            for (int w2 = 1; w2 < 32; ++w2) { l = fminf(l, red[a][w2]); h = fmaxf(h, red[3 + a][w2]); }
            l3[a] = l;
            e3[a] = h - l;
        }
        // cell side: the radius, or the side of a cube holding ~2 points if the points filled their bounding box uniformly
        float r = 0.f;
        if (radius) {
            r = radius[0];
        } else {
            float vol = 1.f;
            int dims = 0;
            for (int a = 0; a < 3; ++a)
                if (e3[a] > 0.f) { vol *= e3[a]; ++dims; }
            if (dims) r = powf(vol * 2.0f / (float)max(n, 1), 1.0f / (float)dims);
        }
        for (int a = 0; a < 3; ++a) {
            const float l = l3[a], ext = e3[a];
            int G = 1;
            if (r > 0.f && ext > 0.f && ext / r < 1e6f) G = max(1, min(gmax, (int)(ext / r)));   // cells at least one side wide
            else if (r > 0.f && ext > 0.f) G = gmax;
            sg.lo[a] = l;
            sg.scale[a] = ext > 0.f ? (float)G / ext : 0.f;
            sg.G[a] = G;
        }
        sg.pad = 0;
        *gout = sg;
    }
    __syncthreads();
    const BallGrid g = sg;
    auto code = [&](int i) -> unsigned {
        unsigned c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = (unsigned)min(g.G[a] - 1, max(0, (int)((pts[(size_t)i * 3 + a] - g.lo[a]) * g.scale[a])));
        return (c[0] * g.G[1] + c[1]) * g.G[2] + c[2];
    };
    for (int i = tid; i < n; i += 1024) atomicAdd(&cell[CI(code(i))], 1u);
    __syncthreads();
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
    const unsigned base = (warp ? wtot[warp - 1] : 0u) + incl - sum;
    {
        unsigned run = base;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const unsigned c = cell[CI(tid * 32 + k)];
            cell[CI(tid * 32 + k)] = run;
            cell_start[tid * 32 + k] = (int)run;
            run += c;
        }
        if (tid == 1023) cell_start[BG_CELLS] = (int)run;   // == n
    }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) sorted_idx[atomicAdd(&cell[CI(code(i))], 1u)] = i;
    __syncthreads();
    // (the order of the points inside a cell is whatever the atomics produced: the query kernel sorts its hits anyway)
    for (int p2 = tid; p2 < n; p2 += 1024) {
        const int i = sorted_idx[p2];
        sorted_xyz[(size_t)p2 * 3 + 0] = pts[(size_t)i * 3 + 0];
        sorted_xyz[(size_t)p2 * 3 + 1] = pts[(size_t)i * 3 + 1];
        sorted_xyz[(size_t)p2 * 3 + 2] = pts[(size_t)i * 3 + 2];
    }
}


}  // namespace rfnet
