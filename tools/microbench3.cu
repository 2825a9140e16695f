// Latency of dependent warp collectives on sm_100a: REDUX (uniform-datapath reduction), SHFL butterfly, VOTE.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb3 tools/microbench3.cu && /tmp/mb3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_redux(unsigned* out, long long* cyc, int iters) {
    unsigned v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __reduce_max_sync(0xffffffffu, v + threadIdx.x) ^ (unsigned)i;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl5(unsigned* out, long long* cyc, int iters) {
    unsigned v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        v += threadIdx.x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        v ^= (unsigned)i;
    }
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_vote(unsigned* out, long long* cyc, int iters) {
    unsigned v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __ballot_sync(0xffffffffu, (v >> (i & 7)) & 1u) + threadIdx.x;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl1(unsigned* out, long long* cyc, int iters) {
    unsigned v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __shfl_sync(0xffffffffu, v, (i + 1) & 31) + threadIdx.x;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(unsigned* out, long long* cyc, int iters) {
    __shared__ unsigned s[32];
    unsigned v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v; __syncthreads(); v += s[(threadIdx.x + i) & 15]; }
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    unsigned* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int it = 10000; long long h;
    auto run = [&](const char* name, void (*k)(unsigned*, long long*, int), int threads) {
        k<<<1, threads>>>(out, cyc, it); k<<<1, threads>>>(out, cyc, it); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-34s %4d threads: %.1f cycles per iteration\n", name, threads, (double)h / it);
    };
    run("REDUX.MAX chain", k_redux, 32); run("REDUX.MAX chain", k_redux, 512);
    run("5-step SHFL butterfly max chain", k_shfl5, 32); run("5-step SHFL butterfly max chain", k_shfl5, 512);
    run("VOTE.BALLOT chain", k_vote, 32); run("VOTE.BALLOT chain", k_vote, 512);
    run("SHFL.IDX chain", k_shfl1, 32); run("SHFL.IDX chain", k_shfl1, 512);
    run("STS + BAR.SYNC + LDS chain", k_bar, 512);
    return 0;
}
