// farthest_point_sample, gather_point, gather_point gradient for sm_100a.
//
// Replaces farthestpointsamplingKernel / gatherpointKernel / scatteraddpointKernel (tf_ops/sampling/tf_sampling_g.cu:105-192).
//
// FPS design: the reference runs one 512-thread block per cloud, keeps the running min-distance in global memory and does
// a 9-level shared-memory tree per pick.  Here a thread-block CLUSTER owns a cloud: every CTA keeps its slice of the
// points AND their running distances in registers, a pick is a warp-shuffle arg-max, one DSMEM store + one remote
// mbarrier arrive per warp into every CTA of the cluster, a local mbarrier wait, and a second shuffle arg-max -- no
// global traffic and no hardware cluster barrier inside the loop (the coordinates of the winner come from a
// shared-memory copy of the cloud).
//
// Exactness: d2 is the reference's fused expression (common.cuh sqdist3<true>); the arg-max reproduces the reference's
// tie rule (tf_sampling_g.cu:146-163): among equal maxima the lowest (k mod 512) wins, then the lowest k.  This is encoded
// in a 64-bit key (distance bits | inverted tie rank) so the arg-max is a plain integer max.
#include <cooperative_groups.h>


#include "common.cuh"
#include "morton.cuh"
#include "rfnet_ops.h"

namespace cg = cooperative_groups;

namespace rfnet {

constexpr int FPS_THREADS = 512;
constexpr int FPS_WARPS = FPS_THREADS / 32;
constexpr int FPS_MAX_CLUSTER = 8;
constexpr int FPS_REF_BLOCK = 512;  // the reference's block size, which defines the tie rule

__device__ __forceinline__ unsigned long long fps_key(float d, int k) {
    // larger d wins; for equal d the smaller (k % 512, k / 512) wins  ->  invert the tie rank so that max() picks it
    const unsigned tie = ((unsigned)(k % FPS_REF_BLOCK) << 23) | (unsigned)(k / FPS_REF_BLOCK);
    return ((unsigned long long)__float_as_uint(d) << 32) | (0xffffffffu - tie);
}
__device__ __forceinline__ int fps_key_index(unsigned long long key) {
    const unsigned tie = 0xffffffffu - (unsigned)(key & 0xffffffffull);
    return (int)((tie & 0x7fffffu) * FPS_REF_BLOCK + (tie >> 23));
}
// 64-bit warp max with two REDUX.MAX (32-bit hardware warp reductions) instead of a 5-step shuffle tree: first the high
// words, then the low words of the lanes that hold the winning high word.
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

// ---- cluster-scope mbarrier helpers (one arrival per remote warp, waited on locally): cheaper than the hardware cluster
// barrier, which has to collect every thread of every CTA
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t remote_addr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote_addr), "l"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_acquire(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// One cluster per cloud.  P = points per thread (register resident).  SMEM_CLOUD: a copy of the whole cloud lives in
// dynamic shared memory (n*12 bytes) so the winner's coordinates are one LDS away; otherwise they are read from global.
//
// One pick = (1) update the P running distances and keep the thread's own maximum -- all points of a thread share
// k mod 512 and are visited in increasing k, so a strict '>' reproduces the reference's per-thread scan
// (tf_sampling_g.cu:146-149) -- (2) warp arg-max of the 64-bit keys (two REDUX), per-warp maxima to shared memory, CTA
// barrier, (3) warp 0 reduces them and its lanes 0..C-1 store the CTA's key into slot [rank] of CTA `lane` and arrive
// (release, cluster scope) on that CTA's mbarrier, (4) everybody waits (acquire) on the local mbarrier, which expects C
// arrivals, and takes the maximum of the C keys.  Slots, per-warp keys and barriers are double buffered by pick parity: a
// CTA can only publish pick j+2 after every CTA of the cluster has published j+1, i.e. after everyone has read pick j.
template <int P, bool SMEM_CLOUD>
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_cluster_kernel(int n, int m, const float* __restrict__ inp, int* __restrict__ out) {
    extern __shared__ __align__(16) float s_cloud[];
    __shared__ __align__(8) unsigned long long wkeys[2][FPS_WARPS];          // per-warp maxima of this CTA
    __shared__ __align__(8) unsigned long long slots[2][FPS_MAX_CLUSTER];    // per-CTA maxima of the whole cluster
    __shared__ __align__(8) uint64_t bars[2];

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ pts = inp + (size_t)cloud * n * 3;
    int* __restrict__ idxs = out + (size_t)cloud * m;

    if (tid == 0) {
        mbar_init(&bars[0], (unsigned)C);   // one arrival per CTA of the cluster and pick
        mbar_init(&bars[1], (unsigned)C);
        mbar_fence_init();
    }
    if (SMEM_CLOUD) {
        for (int i = tid; i < n * 3; i += FPS_THREADS) s_cloud[i] = pts[i];
    }
    // this thread's points: k = rank*per_cta + tid + i*FPS_THREADS
    const int per_cta = (n + C - 1) / C;
    const int k0 = rank * per_cta;
    const int kend = min(n, k0 + per_cta);
    float px[P], py[P], pz[P], td[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int k = k0 + tid + i * FPS_THREADS;
        const bool v = k < kend;
        px[i] = v ? pts[(size_t)k * 3 + 0] : 0.f;
        py[i] = v ? pts[(size_t)k * 3 + 1] : 0.f;
        pz[i] = v ? pts[(size_t)k * 3 + 2] : 0.f;
        td[i] = v ? 1e38f : -1.0f;  // tf_sampling_g.cu:119; slots past the slice can never win (every real distance is >= 0)
    }
    if (rank == 0 && tid == 0) idxs[0] = 0;
    // remote addresses that lane r < C of warp 0 publishes to (CTA r)
    const int dst = lane < C ? lane : 0;
    const uint32_t r_slot0 = map_to_rank(smem_u32(&slots[0][rank]), dst);
    const uint32_t r_slot1 = map_to_rank(smem_u32(&slots[1][rank]), dst);
    const uint32_t r_bar0 = map_to_rank(smem_u32(&bars[0]), dst);
    const uint32_t r_bar1 = map_to_rank(smem_u32(&bars[1]), dst);
    cluster.sync();  // barriers initialised and cloud copies complete in every CTA before anyone publishes

    int old = 0;
    for (int j = 1; j < m; ++j) {
        float lx, ly, lz;
        if (SMEM_CLOUD) {
            lx = s_cloud[old * 3 + 0]; ly = s_cloud[old * 3 + 1]; lz = s_cloud[old * 3 + 2];
        } else {
            lx = __ldg(pts + (size_t)old * 3 + 0); ly = __ldg(pts + (size_t)old * 3 + 1); lz = __ldg(pts + (size_t)old * 3 + 2);
        }
        float bestv = -1.0f;
        int bi = 0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const float d = sqdist3<true>(px[i] - lx, py[i] - ly, pz[i] - lz);
            td[i] = fminf(d, td[i]);
            if (td[i] > bestv) { bestv = td[i]; bi = i; }
        }
        unsigned long long key = bestv >= 0.0f ? fps_key(bestv, k0 + tid + bi * FPS_THREADS) : 0ull;  // 0 < every real key
        key = warp_max_u64(key);
        const int par = j & 1;
        if (lane == 0) wkeys[par][warp] = key;
        __syncthreads();
        if (warp == 0) {
            // CTA maximum, then ONE store + ONE remote arrive per destination CTA (a first version let all 16 warps publish:
            // 64 remote arrivals per barrier and pick serialised on the barrier word)
            unsigned long long ck = lane < FPS_WARPS ? wkeys[par][lane] : 0ull;
            ck = warp_max_u64(ck);
            if (lane < C) {
                st_cluster_u64(par ? r_slot1 : r_slot0, ck);
                mbar_arrive_remote_release(par ? r_bar1 : r_bar0);
            }
        }
        // bars[par] is used by picks par, par+2, ... (bars[1] first at j=1, bars[0] first at j=2): use number (j-1)/2, phase parity its low bit
        mbar_wait_cluster_acquire(&bars[par], (unsigned)(((j - 1) >> 1) & 1));
        unsigned long long best = slots[par][0];
        for (int r = 1; r < C; ++r) {
            const unsigned long long kk = slots[par][r];
            best = kk > best ? kk : best;
        }
        old = fps_key_index(best);
        if (rank == 0 && tid == 0) idxs[j] = old;
    }
    cluster.sync();  // no CTA may exit while a peer can still store into its shared memory
}

// ---------------------------------------------------------------------------------------------------------------
// Pruned FPS: one CTA per cloud, no cluster exchange, and most points are not touched by a pick.
// A pick p can only lower the running distance temp[x] = min(temp[x], d2(x, p)) of points with d2(x, p) < temp[x].  The
// cloud is ordered along a Morton curve (morton.cuh) and cut into clusters of 32 consecutive points -- one register slot
// across a warp -- each with a bounding box and the maximum running distance of its points.  A cluster whose box is
// farther from p than that maximum (dmin2(box, p) >= max temp, with a 1e-6 relative margin for the rounding of the two
// expressions) cannot change: it is skipped, exactly.
// Margin (u = 2^-24).  The kernel's d2(x, p) = fma(dz,dz, fma(dx,dx, dy*dy)) on rounded differences is within 5u relative of
// the exact squared distance (dy^2 term: 2u from the difference, u from the product, u from each fma; the others fewer; all
// terms are non-negative).  dmin2 is the same expression on the per-axis distances to the box: also within 5u of the exact
// box distance, which is <= the exact distance to any point x of the cluster.  Hence d2_computed(x, p) >= dmin2 * (1 - 10u).
// The test skips when fl(dmin2 * 0.999999f) >= cmax; 0.999999f = 1 - 17u and the product adds one rounding, so a skipped
// cluster has dmin2 * (1 - 16u) >= cmax >= temp[x], i.e. d2_computed(x, p) >= dmin2 * (1 - 10u) > temp[x] for all its points:
// min(temp[x], d2) == temp[x], bit for bit.  16u against the 10u needed.
// After a few dozen picks that is all but a handful of the 512
// clusters, so a pick costs one box test per lane, the update of the few live clusters, two REDUX arg-maxes and ONE block
// barrier (measured 0.75 us against 0.93 us for the cluster-wide exchange above; the rest is the latency of the chained warp votes).
// Running distances, tie ranks and the cluster summaries live in registers; the Morton-ordered coordinates in shared memory
// (12 B x n <= 192 KiB).  Values, arg-max and tie rule (lowest k mod 512, then lowest k: tf_sampling_g.cu:146-163) are those
// of fps_cluster_kernel, hence of the reference: same d2 expression, same keys.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FPSP_MAX_POINTS = FPS_THREADS * 32;   // 16384

struct __align__(16) FpsEntry {
    unsigned long long key;
    int pos, pad;
};

template <int P>   // register slots (clusters) per warp, P * 512 >= n
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_pruned_kernel(int n, int m, const float* __restrict__ inp, const int* __restrict__ perm_all,
                                                                    int* __restrict__ out) {
    extern __shared__ __align__(16) float sxyz[];            // Morton-ordered coordinates, AoS
    __shared__ FpsEntry entries[2][FPS_WARPS];
    const int cloud = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ pts = inp + (size_t)cloud * n * 3;
    const int* __restrict__ perm = perm_all + (size_t)cloud * n;
    int* __restrict__ idxs = out + (size_t)cloud * m;

#pragma unroll 4
    for (int pos = tid; pos < n; pos += FPS_THREADS) {
        const int k = perm[pos];
        sxyz[pos * 3 + 0] = pts[(size_t)k * 3 + 0];
        sxyz[pos * 3 + 1] = pts[(size_t)k * 3 + 1];
        sxyz[pos * 3 + 2] = pts[(size_t)k * 3 + 2];
    }
    __syncthreads();

    // slot s of this warp = cluster s * 16 + warp = positions (s * 16 + warp) * 32 + lane.  Consecutive clusters -- spatial
    // neighbours, which a pick tends to hit together -- belong to different warps, so the live clusters of a pick spread
    // over the 16 warps instead of serialising on one (the first version did: 82 % of its stall samples were the barrier)
    float td[P];                 // running distance of this lane's point in every slot (-1: no point)
    unsigned inv[P];             // 0xffffffff - tie rank of that point (fps_key's low word)
    const float inf = __int_as_float(0x7f800000);
    float blo[3] = {inf, inf, inf}, bhi[3] = {-inf, -inf, -inf};   // lane s keeps the box of slot s
    float cmax = -1.0f;                                           // ... its maximum running distance (-1: empty)
    unsigned long long ckey = 0ull;                               // ... the key of the point attaining it
    int cpos = 0;                                                 // ... and the lane that holds that point
#pragma unroll
    for (int s = 0; s < P; ++s) {
        const int pos = (s * FPS_WARPS + warp) * 32 + lane;
        const bool v = pos < n;
        td[s] = v ? 1e38f : -1.0f;   // tf_sampling_g.cu:119
        inv[s] = 0u;
        float x = 0.f, y = 0.f, z = 0.f;
        if (v) {
            const int k = perm[pos];
            inv[s] = (unsigned)(fps_key(0.0f, k) & 0xffffffffull);
            x = sxyz[pos * 3]; y = sxyz[pos * 3 + 1]; z = sxyz[pos * 3 + 2];
        }
        float lo3[3] = {v ? x : inf, v ? y : inf, v ? z : inf}, hi3[3] = {v ? x : -inf, v ? y : -inf, v ? z : -inf};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo3[a] = fminf(lo3[a], __shfl_xor_sync(0xffffffffu, lo3[a], o));
                hi3[a] = fmaxf(hi3[a], __shfl_xor_sync(0xffffffffu, hi3[a], o));
            }
        if (lane == s) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { blo[a] = lo3[a]; bhi[a] = hi3[a]; }
            cmax = lo3[0] <= hi3[0] ? 1e38f : -1.0f;   // non-empty cluster: every point starts at 1e38
        }
    }
    if (tid == 0) idxs[0] = 0;
    float lx = pts[0], ly = pts[1], lz = pts[2];   // the first pick is point 0 (tf_sampling_g.cu:121-122)

    for (int j = 1; j < m; ++j) {
        // (1) which of this warp's clusters can the pick change?
        const float ax = fmaxf(fmaxf(blo[0] - lx, lx - bhi[0]), 0.f), ay = fmaxf(fmaxf(blo[1] - ly, ly - bhi[1]), 0.f),
                    az = fmaxf(fmaxf(blo[2] - lz, lz - bhi[2]), 0.f);
        const float dmin2 = fmaf(az, az, fmaf(ay, ay, ax * ax));
        const unsigned live = __ballot_sync(0xffffffffu, lane < P && !(dmin2 * 0.999999f >= cmax));
        // (2) update the live clusters (warp-uniform branches; slots are register indices, hence the unrolled scan)
#pragma unroll
        for (int s0 = 0; s0 < P; s0 += 8) {
            if (((live >> s0) & 0xffu) == 0u) continue;
#pragma unroll
            for (int s = s0; s < s0 + 8 && s < P; ++s) {
                if (!((live >> s) & 1u)) continue;
                const int pos = (s * FPS_WARPS + warp) * 32 + lane;
                const float x = sxyz[pos * 3], y = sxyz[pos * 3 + 1], z = sxyz[pos * 3 + 2];   // in bounds of the allocation (P * 512 points)
                const float d = sqdist3<true>(x - lx, y - ly, z - lz);
                const float t = fminf(d, td[s]);          // -1 stays -1
                td[s] = t;
                const int tb = __float_as_int(t);
                const int vb = __reduce_max_sync(0xffffffffu, tb);      // non-negative floats order as signed ints; -1.0f is negative
                const unsigned ml = __reduce_max_sync(0xffffffffu, tb == vb ? inv[s] : 0u);
                const int wl = __ffs(__ballot_sync(0xffffffffu, tb == vb && inv[s] == ml)) - 1;
                if (lane == s) {
                    cmax = __int_as_float(vb);
                    ckey = vb >= 0 ? (((unsigned long long)(unsigned)vb << 32) | ml) : 0ull;
                    cpos = wl;
                }
            }
        }
        // (3) arg-max over the clusters: warp, then block (entries double-buffered by pick parity: one barrier per pick)
        const unsigned long long wk = warp_max_u64(lane < P ? ckey : 0ull);
        const int wlane = __ffs(__ballot_sync(0xffffffffu, lane < P && ckey == wk)) - 1;
        const int wpos = (wlane * FPS_WARPS + warp) * 32 + __shfl_sync(0xffffffffu, cpos, wlane);
        const int par = j & 1;
        if (lane == 0) {
            FpsEntry e;
            e.key = wk;
            e.pos = wpos;
            e.pad = 0;
            entries[par][warp] = e;
        }
        __syncthreads();
        const FpsEntry e = entries[par][lane & (FPS_WARPS - 1)];
        const unsigned long long best = warp_max_u64(e.key);
        const int bl = __ffs(__ballot_sync(0xffffffffu, e.key == best)) - 1;
        const int bpos = __shfl_sync(0xffffffffu, e.pos, bl);
        lx = sxyz[bpos * 3]; ly = sxyz[bpos * 3 + 1]; lz = sxyz[bpos * 3 + 2];
        if (tid == 0) idxs[j] = fps_key_index(best);
    }
}

template <int P>
static int launch_fps_pruned(int b, int n, int m, const float* inp, const int* perm, int* out, cudaStream_t s) {
    const size_t smem = (size_t)P * FPS_THREADS * 12;   // whole slots, so that every lane's read stays inside the allocation
    if (smem > 40 * 1024) RFNET_CUDA(cudaFuncSetAttribute(fps_pruned_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // + static entries
    fps_pruned_kernel<P><<<b, FPS_THREADS, smem, s>>>(n, m, inp, perm, out);
    return launch_status();
}

// Fallback for very large clouds (n > 8 * 512 * 8): one CTA per cloud, running distances in the workspace.
__global__ void __launch_bounds__(FPS_THREADS) fps_generic_kernel(int n, int m, const float* __restrict__ inp, float* __restrict__ temp, int* __restrict__ out) {
    __shared__ unsigned long long part[2][FPS_WARPS];
    const int cloud = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ pts = inp + (size_t)cloud * n * 3;
    float* __restrict__ td = temp + (size_t)cloud * n;
    int* __restrict__ idxs = out + (size_t)cloud * m;
    for (int k = tid; k < n; k += FPS_THREADS) td[k] = 1e38f;
    if (tid == 0) idxs[0] = 0;
    int old = 0;
    for (int j = 1; j < m; ++j) {
        const float lx = __ldg(pts + (size_t)old * 3), ly = __ldg(pts + (size_t)old * 3 + 1), lz = __ldg(pts + (size_t)old * 3 + 2);
        unsigned long long key = 0;
        for (int k = tid; k < n; k += FPS_THREADS) {
            const float d = sqdist3<true>(pts[(size_t)k * 3] - lx, pts[(size_t)k * 3 + 1] - ly, pts[(size_t)k * 3 + 2] - lz);
            const float d2 = fminf(d, td[k]);
            td[k] = d2;
            const unsigned long long kk = fps_key(d2, k);
            key = kk > key ? kk : key;
        }
        key = warp_max_u64(key);
        const int par = j & 1;
        if (lane == 0) part[par][warp] = key;
        __syncthreads();
        unsigned long long best = lane < FPS_WARPS ? part[par][lane] : 0ull;
        best = warp_max_u64(best);
        old = fps_key_index(best);
        if (tid == 0) idxs[j] = old;
    }
}

// out[i,j,:] = inp[i,idx[i,j],:]                                              (tf_sampling_g.cu:172-181)
__global__ void gather_point_kernel(int n, int m, size_t total, const float* __restrict__ inp, const int* __restrict__ idx, float* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per output float
    if (t >= total * 3) return;
    const size_t row = t / 3;
    const int c = (int)(t - row * 3);
    const size_t cloud = row / m;
    out[t] = __ldg(inp + (cloud * n + (size_t)idx[row]) * 3 + c);
}

// inp_g[i,idx[i,j],:] += out_g[i,j,:]  after zero-fill                          (tf_sampling_g.cu:183-192, tf_sampling.cpp:174)
__global__ void scatteradd_point_kernel(int n, int m, size_t total, const float* __restrict__ out_g, const int* __restrict__ idx, float* __restrict__ inp_g) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total * 3) return;
    const size_t row = t / 3;
    const int c = (int)(t - row * 3);
    const size_t cloud = row / m;
    atomicAdd(inp_g + (cloud * n + (size_t)idx[row]) * 3 + c, out_g[t]);
}

template <int P>
static int launch_fps_cluster(int b, int n, int m, int C, bool smem_cloud, const float* inp, int* out, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * C));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = smem_cloud ? (size_t)n * 12 : 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (smem_cloud) {
        RFNET_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
        RFNET_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<P, true>, n, m, inp, out));
    } else {
        RFNET_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<P, false>, n, m, inp, out));
    }
    return 0;
}

}  // namespace rfnet

using namespace rfnet;

extern "C" size_t rfnet_farthestpointsampling_workspace_bytes(int b, int n, int m) {
    (void)m;
    if (b <= 0 || n <= 0) return 0;
    // only the large-cloud fallback needs scratch (the reference always needs (32, n) floats: tf_sampling.cpp:115)
    if (n <= FPSP_MAX_POINTS) return sizeof(int) * (size_t)b * n;   // Morton order of every cloud (pruned kernel)
    return n > FPS_MAX_CLUSTER * FPS_THREADS * 8 ? sizeof(float) * (size_t)b * n : 0;
}

extern "C" int rfnet_farthestpointsampling(int b, int n, int m, const float* inp, void* workspace, size_t workspace_bytes, int* out,
                                           rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0);
    if (m <= 0 || b == 0) return 0;  // tf_sampling_g.cu:106-107
    RFNET_CHECK_ARG(n > 0 && inp && out);
    cudaStream_t s = (cudaStream_t)stream;
    // pruned single-CTA kernel whenever the cloud fits one SM's shared memory and a workspace for the Morton order was given
    // (workspace == NULL selects the cluster kernel: the tests compare the two that way)
    // (worth its prologue -- sort, gather, boxes: ~50 us -- from a few hundred picks on)
    if (n <= FPSP_MAX_POINTS && m >= 256 && workspace && workspace_bytes >= sizeof(int) * (size_t)b * n) {
        int* perm = (int*)workspace;
        int rc = morton_sort(b, n, 0, inp, nullptr, perm, nullptr, s);
        if (rc) return rc;
        const int P = (n + FPS_THREADS - 1) / FPS_THREADS;
        if (P <= 1) rc = launch_fps_pruned<1>(b, n, m, inp, perm, out, s);
        else if (P <= 2) rc = launch_fps_pruned<2>(b, n, m, inp, perm, out, s);
        else if (P <= 4) rc = launch_fps_pruned<4>(b, n, m, inp, perm, out, s);
        else if (P <= 8) rc = launch_fps_pruned<8>(b, n, m, inp, perm, out, s);
        else if (P <= 16) rc = launch_fps_pruned<16>(b, n, m, inp, perm, out, s);
        else rc = launch_fps_pruned<32>(b, n, m, inp, perm, out, s);
        return rc;
    }
    if (n > FPS_MAX_CLUSTER * FPS_THREADS * 8) {
        RFNET_CHECK_ARG(workspace && workspace_bytes >= sizeof(float) * (size_t)b * n);
        fps_generic_kernel<<<b, FPS_THREADS, 0, s>>>(n, m, inp, (float*)workspace, out);
        return launch_status();
    }
    // cluster size: enough CTAs that a slice fits 8 points/thread, more when few clouds would leave SMs idle
    int C = 1;
    while (C < FPS_MAX_CLUSTER && (n + C - 1) / C > FPS_THREADS * 8) C <<= 1;
    while (C < FPS_MAX_CLUSTER && (long)b * C * 2 <= num_sms() && (n + C - 1) / C > FPS_THREADS) C <<= 1;
    const int per_cta = (n + C - 1) / C;
    const int P = (per_cta + FPS_THREADS - 1) / FPS_THREADS;
    const bool smem_cloud = (size_t)n * 12 <= 200 * 1024;
    int rc;
    if (P <= 1) rc = launch_fps_cluster<1>(b, n, m, C, smem_cloud, inp, out, s);
    else if (P <= 2) rc = launch_fps_cluster<2>(b, n, m, C, smem_cloud, inp, out, s);
    else if (P <= 4) rc = launch_fps_cluster<4>(b, n, m, C, smem_cloud, inp, out, s);
    else rc = launch_fps_cluster<8>(b, n, m, C, smem_cloud, inp, out, s);
    if (rc) return rc;
    return launch_status();
}

extern "C" int rfnet_gatherpoint(int b, int n, int m, const float* inp, const int* idx, float* out, rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    const size_t total = (size_t)b * m;
    if (total == 0) return 0;
    RFNET_CHECK_ARG(n > 0 && inp && idx && out);
    gather_point_kernel<<<(unsigned)((total * 3 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, m, total, inp, idx, out);
    return launch_status();
}

extern "C" size_t rfnet_scatteraddpoint_workspace_bytes(int b, int n, int m) { return rfnet_group_point_grad_workspace_bytes(b, n, 3, m, 1); }

// gather_point's gradient is group_point's with c = 3 and one sample per row: same atomic-free segmented sum when a
// workspace is given, the reference's zero-fill + float reductions otherwise.
extern "C" int rfnet_scatteraddpoint(int b, int n, int m, const float* out_g, const int* idx, float* inp_g, void* workspace, size_t workspace_bytes,
                                     rfnet_stream_t stream) {
    return rfnet_group_point_grad(b, n, 3, m, 1, out_g, idx, inp_g, workspace, workspace_bytes, stream);
}
