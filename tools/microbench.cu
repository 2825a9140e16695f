// Pipe-throughput microbenchmarks for B200 (sm_100a): which instruction mixes the Chamfer / EMD inner loops can sustain.
// Build+run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench.cu && /tmp/mb
// Output: one line per mix with warp-instructions/clk/SM and FP32 lane-ops/clk/SM (peak: 4 issue/clk/SM, 128 lanes/clk/SM).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int ITERS = 4096;
constexpr int U = 8;  // independent chains per thread

// MODE 0: scalar FFMA            MODE 1: FFMA2                 MODE 2: FADD2+FMUL2+FFMA2 (3:1:2, the NN mix)
// MODE 3: NN mix + FMNMX3 (24:4) MODE 4: MUFU.EX2 only         MODE 5: 4 FFMA2 : 1 MUFU (EMD mix, packed)
// MODE 6: 8 FFMA : 1 MUFU (EMD mix, scalar)                     MODE 7: scalar FADD/FMUL/FFMA NN mix (6 per pair) + FMNMX3
template <int MODE>
__global__ void k(float* out, float seed, long long* clocks) {
    float2 a[U];
    float s[U];
#pragma unroll
    for (int i = 0; i < U; i++) { a[i] = make_float2(seed + i, seed - i); s[i] = seed * i; }
    float2 c = make_float2(seed * 0.5f, seed * 0.25f);
    float best = 1e30f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < U; i++) { a[i].x = __fmaf_rn(a[i].x, c.x, c.y); a[i].y = __fmaf_rn(a[i].y, c.x, c.y); }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < U; i++) a[i] = __ffma2_rn(a[i], c, c);
        } else if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < U; i++) {
                float2 dx = __fadd2_rn(a[i], make_float2(-c.x, -c.x));
                float2 dy = __fadd2_rn(a[i], make_float2(-c.y, -c.y));
                float2 dz = __fadd2_rn(a[i], make_float2(-s[i], -s[i]));
                a[i] = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
            }
            if (MODE == 3) {
                float g0 = fmin3(fmin3(a[0].x, a[1].x, a[2].x), fmin3(a[3].x, a[4].x, a[5].x), fmin3(a[6].x, a[7].x, best));
                float g1 = fmin3(fmin3(a[0].y, a[1].y, a[2].y), fmin3(a[3].y, a[4].y, a[5].y), fmin3(a[6].y, a[7].y, best));
                best = fminf(g0, g1);
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < U; i++) { a[i].x = ex2(a[i].x); a[i].y = ex2(a[i].y); }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < U; i++) {
                float2 t = __ffma2_rn(a[i], c, c);
                t = __ffma2_rn(t, c, a[i]);
                t = __ffma2_rn(t, t, c);
                t = __ffma2_rn(t, c, t);
                a[i].x = ex2(t.x); a[i].y = ex2(t.y);
            }
        } else if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < U; i++) {
                float t = a[i].x;
#pragma unroll
                for (int r = 0; r < 8; r++) t = __fmaf_rn(t, c.x, c.y);
                a[i].x = ex2(t);
            }
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < U; i++) {
                float dx = a[i].x - c.x, dy = a[i].x - c.y, dz = a[i].x - s[i];
                a[i].x = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            }
            best = fmin3(fmin3(a[0].x, a[1].x, a[2].x), fmin3(a[3].x, a[4].x, a[5].x), fmin3(a[6].x, a[7].x, best));
        }
    }
    long long t1 = clock64();
    float r = best;
#pragma unroll
    for (int i = 0; i < U; i++) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double instr_per_iter, double laneops_per_iter, int warps_per_sm) {
    int threads = 32 * warps_per_sm, blocks = 148;
    float* out; long long* clk;
    cudaMalloc(&out, sizeof(float) * threads * blocks); cudaMalloc(&clk, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; i++) cyc += h[i]; cyc /= 148;
    double wi = instr_per_iter * ITERS * warps_per_sm / cyc;             // warp-instr / clk / SM
    double lo = laneops_per_iter * ITERS * warps_per_sm * 32 / cyc;      // fp32 lane-ops / clk / SM
    printf("%-34s warps/SM=%2d  cycles=%9.0f  ms=%.3f  MHz~%.0f  warp-instr/clk/SM=%.2f  fp32-laneops/clk/SM=%.1f\n", name, warps_per_sm, cyc, ms,
           cyc / (ms * 1e3), wi, lo);
    cudaFree(out); cudaFree(clk);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("FFMA scalar", 2 * U, 2 * U, w);
        run<1>("FFMA2 packed", U, 2 * U, w);
        run<2>("NN mix FADD2/FMUL2/FFMA2", 6 * U, 12 * U, w);
        run<3>("NN mix packed + FMNMX3", 6 * U + 9, 12 * U, w);
        run<7>("NN mix scalar + FMNMX3", 6 * U + 4, 6 * U, w);
        run<4>("MUFU.EX2", 2 * U, 0, w);
        run<5>("EMD mix 4 FFMA2 : 2 MUFU", 6 * U, 8 * U, w);
        run<6>("EMD mix 8 FFMA : 1 MUFU", 9 * U, 8 * U, w);
        printf("\n");
    }
    return 0;
}
