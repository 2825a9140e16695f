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
    return n > FPS_MAX_CLUSTER * FPS_THREADS * 8 ? sizeof(float) * (size_t)b * n : 0;
}

extern "C" int rfnet_farthestpointsampling(int b, int n, int m, const float* inp, void* workspace, size_t workspace_bytes, int* out,
                                           rfnet_stream_t stream) {
    RFNET_CHECK_ARG(b >= 0 && n >= 0);
    if (m <= 0 || b == 0) return 0;  // tf_sampling_g.cu:106-107
    RFNET_CHECK_ARG(n > 0 && inp && out);
    cudaStream_t s = (cudaStream_t)stream;
    if (n > FPS_MAX_CLUSTER * FPS_THREADS * 8) {
        RFNET_CHECK_ARG(workspace && workspace_bytes >= rfnet_farthestpointsampling_workspace_bytes(b, n, m));
        fps_generic_kernel<<<b, FPS_THREADS, 0, s>>>(n, m, inp, (float*)workspace, out);
        return launch_status();
    }
    // cluster size: enough CTAs that a slice fits 8 points/thread, more when few clouds would leave SMs idle
    int C = 1;
    while (C < FPS_MAX_CLUSTER && (n + C - 1) / C > FPS_THREADS * 8) C <<= 1;
    while (C < FPS_MAX_CLUSTER && (long)b * C * 2 <= kNumSMs && (n + C - 1) / C > FPS_THREADS) C <<= 1;
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
