// nn_distance (Chamfer nearest-neighbour search) and its gradient for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel of the reference (tf_ops/CD/tf_nndistance_g.cu:4-156).
//
// Design (not a port -- the reference scans 512-candidate tiles with one query per thread and merges tiles through a
// global read-modify-write):
//   * work item = (direction, cloud, tile of 128*Q queries, split of the candidate range); both directions in ONE launch;
//     grid sized so that >= ~30 items per SM exist even at 4 clouds per GPU;
//   * candidates are staged chunk by chunk in shared memory, raw (x,y,z) rows, by the TMA bulk-copy engine
//     (cp.async.bulk + mbarrier, double buffered) -- plain loads when rows are not 16-byte aligned;
//   * each thread keeps Q queries in registers as Q/2 packed pairs and evaluates two distances per instruction with
//     Blackwell's packed FP32 pipe (FADD2 / FMUL2 / FFMA2); candidates are warp-broadcast LDS.128;
//   * min tracking costs < 1 instruction per pair: eight distances and the running best are folded with four 3-input
//     FMNMX3, one compare decides whether the rare slow path (find the FIRST index attaining the new minimum) runs;
//   * splits are merged with a 64-bit atomicMin on (distance bits, index) keys, which is exactly the reference's
//     "smallest distance, lowest index" rule.
// The distance is evaluated in the reference's operand order (common.cuh: sqdist3x2), so indices are bit-exact.
#include "common.cuh"
#include "rfnet_ops.h"

namespace rfnet {

constexpr int NN_THREADS = 128;
constexpr int NN_TC = 1024;  // max candidates per staged chunk (12 KiB per buffer)

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
};
struct NNParams {
    NNDir d[2];
};

template <int Q, bool FUSED>
__global__ void __launch_bounds__(NN_THREADS) nn_search_kernel(const NNParams p) {
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
    int besti[Q];
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
        besti[i] = 0;                           // in which case index 0 stands -- same as the reference's k==0 seed
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
                {
                    const float g = fmin3(fmin3(d[0].x, d[1].x, d[2].x), fmin3(d[3].x, d[4].x, d[5].x), fmin3(d[6].x, d[7].x, best[2 * h]));
                    if (g < best[2 * h]) {  // rare: a strictly smaller distance appeared; take the FIRST index attaining it
                        best[2 * h] = g;
                        int j = 7;
#pragma unroll
                        for (int t = 6; t >= 0; --t) j = (d[t].x == g) ? t : j;
                        besti[2 * h] = kbase + j;
                    }
                }
                {
                    const float g = fmin3(fmin3(d[0].y, d[1].y, d[2].y), fmin3(d[3].y, d[4].y, d[5].y), fmin3(d[6].y, d[7].y, best[2 * h + 1]));
                    if (g < best[2 * h + 1]) {
                        best[2 * h + 1] = g;
                        int j = 7;
#pragma unroll
                        for (int t = 6; t >= 0; --t) j = (d[t].y == g) ? t : j;
                        besti[2 * h + 1] = kbase + j;
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with sbuf before the copy engine may overwrite it
    }

    // ---- results
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int qi = q0 + i * NN_THREADS;
        if (qi < nq) {
            const size_t o = (size_t)cloud * nq + qi;
            if (D.nsplit == 1) {
                D.dist[o] = best[i];
                D.idx[o] = besti[i];
            } else {
                atomicMin(&D.keys[o], pack_key(best[i], besti[i]));
            }
        }
    }
}

__global__ void nn_unpack_keys_kernel(const unsigned long long* __restrict__ keys, float* __restrict__ dist, int* __restrict__ idx, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const unsigned long long k = keys[i];
        dist[i] = __uint_as_float((unsigned)(k >> 32));
        idx[i] = (int)(unsigned)(k & 0xffffffffull);
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
        atomicAdd(&g2[oi + 0], -(g * (a[0] - xyz2[oi + 0])));
        atomicAdd(&g2[oi + 1], -(g * (a[1] - xyz2[oi + 1])));
        atomicAdd(&g2[oi + 2], -(g * (a[2] - xyz2[oi + 2])));
    } else if (t < total1 + total2) {
        const size_t u = t - total1;
        const size_t cloud = u / m;
        const int j2 = idx2[u];
        const float g = gd2[u] * 2.0f;
        const float* a = xyz2 + u * 3;
        const size_t oi = (cloud * n + j2) * 3;
        atomicAdd(&g1[oi + 0], -(g * (a[0] - xyz1[oi + 0])));
        atomicAdd(&g1[oi + 1], -(g * (a[1] - xyz1[oi + 1])));
        atomicAdd(&g1[oi + 2], -(g * (a[2] - xyz1[oi + 2])));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static void plan_direction(NNDir& D, int b, int nq, int nc, int Q, long target_items) {
    D.nq = nq;
    D.nc = nc;
    const int TQ = NN_THREADS * Q;
    D.nqt = (nq + TQ - 1) / TQ;
    int chunk = NN_TC;
    while (chunk > 256 && (long)b * D.nqt * ((nc + chunk - 1) / chunk) < target_items) chunk >>= 1;
    D.chunk = chunk;
    const int nchunks = (nc + chunk - 1) / chunk;
    long cps = ((long)b * D.nqt * nchunks) / target_items;
    if (cps < 1) cps = 1;
    if (cps > nchunks) cps = nchunks;
    D.cps = (int)cps;
    D.nsplit = (nchunks + D.cps - 1) / D.cps;
    D.items = b * D.nqt * D.nsplit;
    D.tma = (nc % 4 == 0) && (((uintptr_t)D.c & 15u) == 0);
}

static int pick_q(int nq) { return nq >= 1024 ? 8 : (nq >= 512 ? 4 : 2); }

template <int Q>
static void launch_search(const NNParams& p, int grid, bool fused, cudaStream_t s) {
    if (fused)
        nn_search_kernel<Q, true><<<grid, NN_THREADS, 0, s>>>(p);
    else
        nn_search_kernel<Q, false><<<grid, NN_THREADS, 0, s>>>(p);
}

}  // namespace rfnet

using namespace rfnet;

extern "C" size_t rfnet_nn_distance_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return (size_t)b * ((size_t)n + (size_t)m) * sizeof(unsigned long long);
}

extern "C" int rfnet_nn_distance(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1, float* dist2,
                                 int* idx2, void* workspace, size_t workspace_bytes, int flags, rfnet_stream_t stream) {
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
    const int Q = pick_q(n < m ? n : m);
    NNParams p;
    p.d[0].q = xyz1; p.d[0].c = xyz2; p.d[0].dist = dist1; p.d[0].idx = idx1;
    p.d[1].q = xyz2; p.d[1].c = xyz1; p.d[1].dist = dist2; p.d[1].idx = idx2;
    const long target = (long)kNumSMs * 16;  // per direction
    plan_direction(p.d[0], b, n, m, Q, target);
    plan_direction(p.d[1], b, m, n, Q, target);
    unsigned long long* keys = (unsigned long long*)workspace;
    p.d[0].keys = keys;
    p.d[1].keys = keys ? keys + (size_t)b * n : nullptr;
    const bool need0 = p.d[0].nsplit > 1, need1 = p.d[1].nsplit > 1;
    if (need0 || need1) {
        RFNET_CHECK_ARG(workspace && workspace_bytes >= rfnet_nn_distance_workspace_bytes(b, n, m));
        if (need0) RFNET_CUDA(cudaMemsetAsync(p.d[0].keys, 0xff, sizeof(unsigned long long) * (size_t)b * n, s));
        if (need1) RFNET_CUDA(cudaMemsetAsync(p.d[1].keys, 0xff, sizeof(unsigned long long) * (size_t)b * m, s));
    }
    const int grid = p.d[0].items + p.d[1].items;
    const bool fused = !(flags & RFNET_NN_UNFUSED);
    if (Q == 8) launch_search<8>(p, grid, fused, s);
    else if (Q == 4) launch_search<4>(p, grid, fused, s);
    else launch_search<2>(p, grid, fused, s);
    if (need0) {
        const size_t cnt = (size_t)b * n;
        nn_unpack_keys_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(p.d[0].keys, dist1, idx1, cnt);
    }
    if (need1) {
        const size_t cnt = (size_t)b * m;
        nn_unpack_keys_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(p.d[1].keys, dist2, idx2, cnt);
    }
    return launch_status();
}

extern "C" int rfnet_nn_distance_grad(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                                      const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1, float* grad_xyz2,
                                      rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if (b == 0 || (n == 0 && m == 0)) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        if (n) RFNET_CUDA(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * 3 * (size_t)b * n, s));
        if (m) RFNET_CUDA(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * 3 * (size_t)b * m, s));
        return 0;
    }
    RFNET_CHECK_ARG(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && grad_xyz1 && grad_xyz2);
    const size_t t1 = (size_t)b * n, t2 = (size_t)b * m;
    const unsigned grid = (unsigned)((t1 + t2 + 255) / 256);
    nn_grad_own_kernel<<<grid, 256, 0, s>>>(n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, t1, t2);
    nn_grad_scatter_kernel<<<grid, 256, 0, s>>>(n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, t1, t2);
    return launch_status();
}
