// Tuning / self-check harness for the nn_distance kernels: times rfnet_nn_distance in both modes (filtered search = default,
// RFNET_NN_DIRECT = the reference expression for every pair), checks that the two agree bit for bit, and reports how many
// (query, item) pairs the filtered search had to scan exactly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I rfnet_b200/csrc [-DNN_THREADS_VALUE=128 -DNN_MIN_CTAS=5
//        -DNNF_MIN_CTAS=4 -DNNF_GROUP=16] -o tools/bin/nn_tune tools/nn_tune.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#define NN_TUNE 1
#include "../rfnet_b200/csrc/nn_distance.cu"

int main(int argc, char** argv) {
    if (argc > 1) rfnet::g_force_chunk = atoi(argv[1]);   // 0 = planner's choice
    if (argc > 2) rfnet::g_force_cps = atoi(argv[2]);
    const bool quick = argc > 3;
    printf("force chunk=%d cps=%d\n", rfnet::g_force_chunk, rfnet::g_force_cps);
    const int shapes[][3] = {{32, 2048, 16384}, {32, 16384, 16384}, {4, 16384, 16384}, {8, 2048, 2048}, {3, 1000, 777}};
    const float offsets[] = {0.f, 10.f, -1.f, -2.f, -3.f};   // 0: unit cube at the origin; 10: shifted by 10; -1: 1/16 lattice (exact ties); -2: surface z = f(x,y) at scale 1e3; -3: 300 distinct points repeated to fill the cloud (resample_pcd padding)
    printf("NN_THREADS=%d NN_MIN_CTAS=%d NNF_MIN_CTAS=%d NNF_GROUP=%d\n", NN_THREADS, NN_MIN_CTAS, NNF_MIN_CTAS, NNF_G);
    for (auto& sh : shapes)
        for (float off : offsets) {
            const int b = sh[0], n = sh[1], m = sh[2];
            if (off != 0.f && b * (size_t)n * m > 4e9) continue;
            if (quick && (off != 0.f || (n < 16384 && m < 16384))) continue;
            if (off < 0.f && (size_t)b * n * m > 2e9) continue;
            std::vector<float> h1((size_t)b * n * 3), h2((size_t)b * m * 3);
            unsigned s = 777;
            auto gen = [&](std::vector<float>& h) {
                for (size_t i = 0; i < h.size(); ++i) {
                    s = s * 1664525u + 1013904223u;
                    float v = (s >> 8) / 16777216.0f - 0.5f;
                    if (off > 0.f) v += off;
                    if (off == -1.f) v = floorf(v * 16.f) / 16.f;
                    if (off == -2.f) v = (i % 3 == 2) ? 1e3f * (h[i - 1] * 1e-3f * h[i - 2] * 1e-3f) : 1e3f * v;
                    h[i] = v;
                }
            };
            gen(h1); gen(h2);
            if (off == -3.f) {
                for (int c = 0; c < b; ++c)
                    for (int i = 300; i < n; ++i) { s = s * 1664525u + 1013904223u; const int src = (s >> 8) % 300; for (int d = 0; d < 3; ++d) h1[((size_t)c * n + i) * 3 + d] = h1[((size_t)c * n + src) * 3 + d]; }
            }
            float *x1, *x2, *d1, *d2, *e1, *e2; int *i1, *i2, *j1, *j2; void* ws; unsigned long long* st;
            const size_t wsb = rfnet_nn_distance_workspace_bytes(b, n, m);
            cudaMalloc(&x1, h1.size() * 4); cudaMalloc(&x2, h2.size() * 4); cudaMalloc(&d1, (size_t)b * n * 4); cudaMalloc(&d2, (size_t)b * m * 4);
            cudaMalloc(&i1, (size_t)b * n * 4); cudaMalloc(&i2, (size_t)b * m * 4); cudaMalloc(&ws, wsb ? wsb : 16);
            cudaMalloc(&e1, (size_t)b * n * 4); cudaMalloc(&e2, (size_t)b * m * 4); cudaMalloc(&j1, (size_t)b * n * 4); cudaMalloc(&j2, (size_t)b * m * 4);
            cudaMalloc(&st, 8); cudaMemset(st, 0, 8);
            cudaMemcpy(x1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(x2, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice);
            cudaEvent_t ev0, ev1; cudaEventCreate(&ev0); cudaEventCreate(&ev1);
            float ms[2];
            for (int mode = 0; mode < 2; ++mode) {
                const int flags = mode ? RFNET_NN_DIRECT : 0;
                float* od1 = mode ? e1 : d1; float* od2 = mode ? e2 : d2; int* oi1 = mode ? j1 : i1; int* oi2 = mode ? j2 : i2;
                for (int i = 0; i < 3; ++i) rfnet_nn_distance(b, n, x1, m, x2, od1, oi1, od2, oi2, ws, wsb, flags, 0);
                cudaEventRecord(ev0);
                const int it = 10;
                for (int i = 0; i < it; ++i) rfnet_nn_distance(b, n, x1, m, x2, od1, oi1, od2, oi2, ws, wsb, flags, 0);
                cudaEventRecord(ev1); cudaEventSynchronize(ev1);
                cudaEventElapsedTime(&ms[mode], ev0, ev1); ms[mode] /= it;
            }
            rfnet_nn_distance_stats(b, n, x1, m, x2, d1, i1, d2, i2, ws, wsb, 0, st, 0);
            unsigned long long scans = 0; cudaMemcpy(&scans, st, 8, cudaMemcpyDeviceToHost);
            std::vector<float> a1((size_t)b * n), c1((size_t)b * n), a2((size_t)b * m), c2((size_t)b * m);
            std::vector<int> p1((size_t)b * n), r1((size_t)b * n), p2((size_t)b * m), r2((size_t)b * m);
            cudaMemcpy(a1.data(), d1, a1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c1.data(), e1, c1.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(a2.data(), d2, a2.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data(), e2, c2.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(p1.data(), i1, p1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(r1.data(), j1, r1.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(p2.data(), i2, p2.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(r2.data(), j2, r2.size() * 4, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < a1.size(); ++i) bad += (memcmp(&a1[i], &c1[i], 4) != 0) || p1[i] != r1[i];
            for (size_t i = 0; i < a2.size(); ++i) bad += (memcmp(&a2[i], &c2[i], 4) != 0) || p2[i] != r2[i];
            const double pairs = 2.0 * b * n * m;
            const cudaError_t err = cudaDeviceSynchronize();
            printf("  b=%d n=%d m=%d off=%g: filter %.3f ms %.0f Gpairs/s (%.1f%% of the 6-op FP32 roofline) | direct %.3f ms %.0f Gpairs/s (%.1f%%) | x%.2f | "
                   "mismatches %zu | exact scans %llu (%.3f%% of queries) | %s\n",
                   b, n, m, off, ms[0], pairs / ms[0] / 1e6, pairs * 6 / (ms[0] * 1e-3) / (148.0 * 128 * 1.965e9) * 100, ms[1], pairs / ms[1] / 1e6,
                   pairs * 6 / (ms[1] * 1e-3) / (148.0 * 128 * 1.965e9) * 100, ms[1] / ms[0], bad, scans, 100.0 * scans / ((double)b * (n + m)),
                   cudaGetErrorString(err));
            cudaFree(x1); cudaFree(x2); cudaFree(d1); cudaFree(d2); cudaFree(i1); cudaFree(i2); cudaFree(ws);
            cudaFree(e1); cudaFree(e2); cudaFree(j1); cudaFree(j2); cudaFree(st);
        }
    return 0;
}
