// Shared device helpers for the rfnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "rfnet_b200 kernels are written for sm_100a (Blackwell) only"
#endif

namespace rfnet {

constexpr int kNumSMsB200 = 148;  // B200; only the fallback of num_sms() when the attribute query fails

// Multiprocessor count of the CURRENT device, queried once per device (launch heuristics scale grids with it).
static inline int num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsB200;
    int v = cache[dev];
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMsB200;
        cache[dev] = v;   // benign race: every thread writes the same value
    }
    return v;
}

#define RFNET_CHECK_ARG(cond) \
    do {                      \
        if (!(cond)) return (int)cudaErrorInvalidValue; \
    } while (0)

#define RFNET_CUDA(call)                     \
    do {                                     \
        cudaError_t _e = (call);             \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

static inline int launch_status() { return (int)cudaGetLastError(); }

// Launch with programmatic stream serialization: the kernel must call pdl_wait() before it touches anything the previous kernel
// of the stream wrote (or anything that kernel still reads and this one overwrites).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------------------------
// Squared distance in the reference's operand order.  Explicit intrinsics so -fmad never decides the contraction.
//   fused   : fma(dz,dz, fma(dx,dx, dy*dy))     what nvcc emits for x*x+y*y+z*z in every reference CUDA kernel
//   unfused : ((dx*dx)+(dy*dy))+(dz*dz)         the reference's CPU build (g++ -O2 without FMA)
// ---------------------------------------------------------------------------------------------------------------
template <bool FUSED>
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz) {
    if (FUSED) return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Packed FP32x2 version (Blackwell FADD2/FMUL2/FFMA2): two distances per instruction, each lane IEEE-rn, so the results
// are bit-identical to sqdist3 on the two halves.
// NOTE (ptxas 12.9): a packed mul.rn.f32x2 feeding a packed add.rn.f32x2 -- or even fma.rn.f32x2(x, 1, y) -- IS contracted
// into FFMA2 despite the .rn qualifiers (checked in SASS, and caught by the unfused parity test).  The unfused variant
// therefore does its two additions with scalar __fadd_rn, which ptxas never contracts.
template <bool FUSED>
__device__ __forceinline__ float2 sqdist3x2(float2 dx, float2 dy, float2 dz) {
    if (FUSED) return __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
    const float2 px = __fmul2_rn(dx, dx), py = __fmul2_rn(dy, dy), pz = __fmul2_rn(dz, dz);
    return make_float2(__fadd_rn(__fadd_rn(px.x, py.x), pz.x), __fadd_rn(__fadd_rn(px.y, py.y), pz.y));
}

// 3-input min (FMNMX3 on sm_100).
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// 2^x on the MUFU pipe.  The reference's __expf compiles to ex2.approx.f32 WITHOUT .ftz, which ptxas expands to
// FSETP + FMUL + MUFU.EX2 + FMUL (halve the argument / square the result below 2^-126 so that denormal results survive).
// The .ftz form is the bare MUFU.EX2: identical bits whenever the result is a normal number, and 0 instead of a denormal
// (< 1.2e-38) otherwise.  Dropped terms are below 1.2e-38 x weight (<= 1e9), i.e. < 1e-29 absolute in sums that carry a 1e-9
// regulariser and in match entries: far inside every tolerance, but NOT bit-identical to the reference wherever such a term
// exists -- the reference-order mode of approx_match therefore uses ex2_approx_full.  The ftz form buys ~25 % of the sweep's
// FMA-pipe time and makes the exact pruning of the sharp levels possible (exact zeros).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// The reference's form (denormal results kept): used only where bit-for-bit agreement with its binary is asked for.
__device__ __forceinline__ float ex2_approx_full(float x) {
    float y;
    asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP): contiguous global -> shared, completion on an mbarrier.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Packed (distance, index) key: for d >= +0 the float bit pattern is monotone as an unsigned integer, so an integer min
// over keys yields "smallest distance, then smallest index" -- the reference's first-minimum rule -- exactly.
__device__ __forceinline__ unsigned long long pack_key(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}

// Programmatic dependent launch (griddepcontrol): a kernel launched with launch_pdl() may become resident while the kernel before
// it in the stream drains; pdl_wait() blocks until that kernel has completed and its writes are visible (a no-op for a kernel
// launched the ordinary way), pdl_launch_dependents() lets the NEXT kernel's CTAs take the slots this grid frees.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rfnet
