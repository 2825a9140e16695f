// Per-instruction throughput probes (B200): which packed-FP32 forms run at full rate, and how ALU/MUFU ops co-issue.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr int ITERS = 2048;
constexpr int U = 8;
template <int MODE>
__global__ void k(float* out, float seed, float zero, long long* clocks) {
    float2 a[U], b[U];
#pragma unroll
    for (int i = 0; i < U; i++) { a[i] = make_float2(seed + i, seed - i); b[i] = make_float2(seed * i, seed + 2 * i); }
    float2 c = make_float2(seed * 0.5f, seed * 0.25f);
    float2 z2 = make_float2(zero, zero);
    float best = 1e30f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < U; i++) {
                if (MODE == 0) a[i] = __fadd2_rn(a[i], b[i]);                                   // FADD2 pair+pair
                if (MODE == 1) a[i] = __fadd2_rn(a[i], make_float2(-c.x, -c.x));                 // FADD2 pair + (-scalar)
                if (MODE == 2) a[i] = __fmul2_rn(a[i], b[i]);                                   // FMUL2
                if (MODE == 3) a[i] = __ffma2_rn(a[i], b[i], c);                                // FFMA2 3 distinct
                if (MODE == 4) a[i] = __ffma2_rn(a[i], a[i], z2);                               // FFMA2 as square (runtime 0)
                if (MODE == 5) { a[i].x = a[i].x + b[i].x; a[i].y = a[i].y + b[i].y; }          // FADD scalar x2
                if (MODE == 6) { a[i].x = a[i].x * b[i].x; a[i].y = a[i].y * b[i].y; }          // FMUL scalar x2
            }
        }
        if (MODE == 7 || MODE == 8 || MODE == 9) {   // NN group: 8 candidates x 1 query pair: 24 FADD2 + 8 sq + 16 FFMA2, + 8 FMNMX3 + 2 FSETP
#pragma unroll
            for (int i = 0; i < U; i++) {
                float2 dx = __fadd2_rn(b[i], make_float2(-c.x, -c.x));
                float2 dy = __fadd2_rn(b[i], make_float2(-c.y, -c.y));
                float2 dz = __fadd2_rn(b[i], make_float2(-a[i].x, -a[i].x));
                float2 sq = (MODE == 8) ? __ffma2_rn(dy, dy, z2) : __fmul2_rn(dy, dy);
                a[i] = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, sq));
            }
            if (MODE != 9) {
                float g0 = fmin3(fmin3(a[0].x, a[1].x, a[2].x), fmin3(a[3].x, a[4].x, a[5].x), fmin3(a[6].x, a[7].x, best));
                float g1 = fmin3(fmin3(a[0].y, a[1].y, a[2].y), fmin3(a[3].y, a[4].y, a[5].y), fmin3(a[6].y, a[7].y, best));
                best = fminf(g0, g1);
            }
        }
        if (MODE == 10) {  // EMD pair-pass, packed: 3 FADD2, FMUL2, 2 FFMA2, FMUL2, 2 MUFU, FFMA2 per 2 pairs
#pragma unroll
            for (int i = 0; i < U; i++) {
                float2 dx = __fadd2_rn(b[i], make_float2(c.x, c.x));
                float2 dy = __fadd2_rn(b[i], make_float2(c.y, c.y));
                float2 dz = __fadd2_rn(b[i], make_float2(best, best));
                float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
                float2 ar = __fmul2_rn(d, z2);
                float2 e = make_float2(ex2(ar.x), ex2(ar.y));
                a[i] = __ffma2_rn(e, c, a[i]);
            }
        }
        if (MODE == 11) {  // same with squares as FFMA2(runtime zero) and the level multiply folded into an FFMA2
#pragma unroll
            for (int i = 0; i < U; i++) {
                float2 dx = __fadd2_rn(b[i], make_float2(c.x, c.x));
                float2 dy = __fadd2_rn(b[i], make_float2(c.y, c.y));
                float2 dz = __fadd2_rn(b[i], make_float2(best, best));
                float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, z2)));
                float2 ar = __ffma2_rn(d, z2, z2);
                float2 e = make_float2(ex2(ar.x), ex2(ar.y));
                a[i] = __ffma2_rn(e, c, a[i]);
            }
        }
    }
    long long t1 = clock64();
    float r = best;
#pragma unroll
    for (int i = 0; i < U; i++) r += a[i].x + a[i].y + b[i].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, double instr_per_iter, double laneops_per_iter, int w) {
    int threads = 32 * w, blocks = 148;
    float* out; long long* clk;
    cudaMalloc(&out, sizeof(float) * threads * blocks); cudaMalloc(&clk, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.0f, clk);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.0f, clk);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; i++) cyc += h[i]; cyc /= 148;
    printf("%-44s warps/SM=%2d  warp-instr/clk/SM=%.2f  fp32-laneops/clk/SM=%.1f  clk/instr/SMSP=%.2f\n", name, w, instr_per_iter * ITERS * w / cyc,
           laneops_per_iter * ITERS * w * 32 / cyc, cyc / (instr_per_iter * ITERS * w / 4));
    cudaFree(out); cudaFree(clk);
}
int main() {
    for (int w : {16, 32}) {
        run<0>("FADD2 pair+pair", 4 * U, 8 * U, w);
        run<1>("FADD2 pair+(-scalar)", 4 * U, 8 * U, w);
        run<2>("FMUL2", 4 * U, 8 * U, w);
        run<3>("FFMA2 (3 regs)", 4 * U, 8 * U, w);
        run<4>("FFMA2 a*a+runtime0", 4 * U, 8 * U, w);
        run<5>("FADD scalar", 8 * U, 8 * U, w);
        run<6>("FMUL scalar", 8 * U, 8 * U, w);
        run<7>("NN group packed (FMUL2 sq) + FMNMX3", 6 * U + 9, 12 * U, w);
        run<8>("NN group packed (FFMA2 sq) + FMNMX3", 6 * U + 9, 12 * U, w);
        run<9>("NN group packed (FMUL2 sq), no min", 6 * U, 12 * U, w);
        run<10>("EMD pair-pass packed (2 FMUL2)", 10 * U, 16 * U, w);
        run<11>("EMD pair-pass packed (all FFMA2)", 10 * U, 16 * U, w);
        printf("\n");
    }
    return 0;
}
