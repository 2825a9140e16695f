// FFMA2 operand forms used by nn_filter_kernel (B200): does a scalar-broadcast multiplier / addend run at full rate?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/microbench4 tools/microbench4.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
constexpr int ITERS = 20000;
constexpr int NC = 1024;
// MODE 0: a = fma2(b, bc(s), a)     pair, scalar, pair(dst)
// MODE 1: a = fma2(a, bc(s), bc(t)) pair, scalar, scalar
// MODE 2: a = fma2(b, c2, a)        pair, pair, pair(dst)   (c2 a loop-invariant pair)
// MODE 3: filter group, scalar-broadcast candidates from shared memory (the kernel's loop), 4 query pairs, G = 8, with tracking
// MODE 4: as 3 without the min tracking (results folded by 1 FMNMX3 per 2 values only ... still ALU) -- skipped
// MODE 5: filter group with candidates stored DUPLICATED in shared memory ((x,x,y,y),(z,z,w,w)): pair,pair,pair operands
// MODE 6: MODE 3 with G = 16
template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, float seed, float zero, long long* clocks) {
    __shared__ __align__(16) float4 sC[NC * 2];
    for (int i = threadIdx.x; i < NC * 2; i += blockDim.x) sC[i] = make_float4(seed * i, seed + i, seed - i, 3.f + i);
    __syncthreads();
    constexpr int U = 8;
    float2 a[U], b[U];
#pragma unroll
    for (int i = 0; i < U; i++) { a[i] = make_float2(seed + i, seed - i); b[i] = make_float2(seed * i, seed + 2 * i); }
    float2 c = make_float2(seed * 0.5f, seed * 0.25f);
    float b1[8], b2[8]; int k1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { b1[i] = 1e30f; b2[i] = 1e30f; k1[i] = 0; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE <= 2) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int i = 0; i < U; i++) {
                    if (MODE == 0) a[i] = __ffma2_rn(b[i], make_float2(c.x, c.x), a[i]);
                    if (MODE == 1) a[i] = __ffma2_rn(a[i], make_float2(c.x, c.x), make_float2(c.y, c.y));
                    if (MODE == 2) a[i] = __ffma2_rn(b[i], c, a[i]);
                }
            }
        } else {
            constexpr int G = (MODE == 6 || MODE == 7) ? 16 : 8;
            const int kb = (it * G) & (NC - 1);
            float g[8];
#pragma unroll
            for (int part = 0; part < G / 8; ++part) {
                float4 cc[8], cd[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    cc[j] = sC[(MODE == 5 ? 2 : 1) * (kb + part * 8 + j)];
                    if (MODE == 5) cd[j] = sC[2 * (kb + part * 8 + j) + 1];
                }
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    float2 s[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (MODE == 5) {
                            s[j] = __ffma2_rn(a[h], make_float2(cd[j].x, cd[j].y), make_float2(cd[j].z, cd[j].w));
                            s[j] = __ffma2_rn(b[h], make_float2(cc[j].z, cc[j].w), s[j]);
                            s[j] = __ffma2_rn(b[h + 4], make_float2(cc[j].x, cc[j].y), s[j]);
                        } else {
                            s[j] = __ffma2_rn(a[h], make_float2(cc[j].z, cc[j].z), make_float2(cc[j].w, cc[j].w));
                            s[j] = __ffma2_rn(b[h], make_float2(cc[j].y, cc[j].y), s[j]);
                            s[j] = __ffma2_rn(b[h + 4], make_float2(cc[j].x, cc[j].x), s[j]);
                        }
                    }
                    if (part == 0) {
                        g[2 * h] = fmin3(fmin3(s[0].x, s[1].x, s[2].x), fmin3(s[3].x, s[4].x, s[5].x), fminf(s[6].x, s[7].x));
                        g[2 * h + 1] = fmin3(fmin3(s[0].y, s[1].y, s[2].y), fmin3(s[3].y, s[4].y, s[5].y), fminf(s[6].y, s[7].y));
                    } else {
                        g[2 * h] = fmin3(fmin3(g[2 * h], s[0].x, s[1].x), fmin3(s[2].x, s[3].x, s[4].x), fmin3(s[5].x, s[6].x, s[7].x));
                        g[2 * h + 1] = fmin3(fmin3(g[2 * h + 1], s[0].y, s[1].y), fmin3(s[2].y, s[3].y, s[4].y), fmin3(s[5].y, s[6].y, s[7].y));
                    }
                }
            }
            if (MODE == 7) {
#pragma unroll
                for (int i = 0; i < 8; i++) b1[i] = fminf(g[i], b1[i]);
            } else
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float hi = fmaxf(g[i], b1[i]);
                k1[i] = (g[i] < b1[i]) ? kb : k1[i];
                b1[i] = fminf(g[i], b1[i]);
                b2[i] = fminf(b2[i], hi);
            }
        }
    }
    long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < U; i++) r += a[i].x + a[i].y + b[i].x + b1[i] + b2[i] + k1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, double ffma2_per_iter, int ctas) {
    int threads = 128, blocks = 148 * ctas;
    float* out; long long* clk;
    cudaMalloc(&out, sizeof(float) * threads * blocks); cudaMalloc(&clk, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.0f, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.0f, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;   // whole-kernel time at the (unthrottled) 1965 MHz boost clock
    // per SMSP: ctas warps (4 warps per CTA, one per SMSP)
    printf("%-64s CTAs/SM=%d  clk per FFMA2 per SMSP=%.2f\n", name, ctas, cyc / (ffma2_per_iter * ITERS * ctas));
    cudaFree(out); cudaFree(clk);
}
int main() {
    for (int ctas : {3, 4, 5}) {
        run<0>("FFMA2 pair*scalar+pair", 32, ctas);
        run<1>("FFMA2 pair*scalar+scalar", 32, ctas);
        run<2>("FFMA2 pair*pair+pair", 32, ctas);
        run<3>("filter group G=8, scalar operands, tracking", 96, ctas);
        run<6>("filter group G=16, scalar operands, tracking", 192, ctas);
        run<5>("filter group G=8, duplicated (pair) operands, tracking", 96, ctas);
        run<7>("filter group G=16, scalar operands, min tree only", 192, ctas);
        printf("\n");
    }
    return 0;
}
