// Tuning harness for the EMD sweep kernel: times emd_sweep_kernel<Q> directly for several (Q, nsplit) at fixed shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I rfnet_b200/csrc -DEMD_UNROLL=2 -o /tmp/emd_tune tools/emd_tune.cu
#include <cstdio>
#include <vector>
#include "../rfnet_b200/csrc/approxmatch.cu"

template <int Q>
float run(int b, int n, int nsplit, float lvl2, const float* x1, const float* x2, const float* w, float* partial) {
    const int nrt = (n + EMD_THREADS * Q - 1) / (EMD_THREADS * Q);
    const int chunks = (n + EMD_TC - 1) / EMD_TC;
    const int cps = (chunks + nsplit - 1) / nsplit;
    nsplit = (chunks + cps - 1) / cps;
    const unsigned grid = (unsigned)(b * nrt * nsplit);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) emd_sweep_kernel<Q, 1, false><<<grid, EMD_THREADS>>>(n, n, nrt, nsplit, cps, lvl2, 1e-9f, x1, x2, w, nullptr, partial, EmdEpi{w, partial, partial});
    cudaEventRecord(e0);
    const int it = 5;
    for (int i = 0; i < it; ++i) emd_sweep_kernel<Q, 1, false><<<grid, EMD_THREADS>>>(n, n, nrt, nsplit, cps, lvl2, 1e-9f, x1, x2, w, nullptr, partial, EmdEpi{w, partial, partial});
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    const double pp = (double)b * n * n;
    printf("  Q=%d nsplit=%2d grid=%6u (%.1f CTAs/SM): %8.3f ms  %.2f Tpair-pass/s  MUFU %.1f%%\n", Q, nsplit, grid, grid / 148.0, ms, pp / ms / 1e9,
           100.0 * pp / (ms * 1e-3) / (148.0 * 16 * 1.965e9));
    return ms;
}

int main() {
    const int shapes[3][2] = {{32, 2048}, {4, 16384}, {32, 16384}};
    printf("EMD_UNROLL=%d EMD_TC=%d\n", EMD_UNROLL, EMD_TC);
    for (auto& sh : shapes) {
        const int b = sh[0], n = sh[1];
        std::vector<float> h((size_t)b * n * 3), hw((size_t)b * n, 1.0f);
        unsigned s = 12345;
        for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (s >> 8) / 16777216.0f - 0.5f; }
        float *x1, *x2, *w, *partial;
        cudaMalloc(&x1, h.size() * 4); cudaMalloc(&x2, h.size() * 4); cudaMalloc(&w, hw.size() * 4); cudaMalloc(&partial, (size_t)b * n * 32 * 4);
        cudaMemcpy(x1, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (s >> 8) / 16777216.0f - 0.5f; }
        cudaMemcpy(x2, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
        printf("b=%d n=m=%d, level -64\n", b, n);
        const float lvl2 = -64.0f * LOG2E;
        for (int ns : {1, 2, 4, 8, 16, 32}) {
            if (ns > (n + EMD_TC - 1) / EMD_TC) continue;
            run<2>(b, n, ns, lvl2, x1, x2, w, partial);
            run<4>(b, n, ns, lvl2, x1, x2, w, partial);
            run<8>(b, n, ns, lvl2, x1, x2, w, partial);
        }
        cudaFree(x1); cudaFree(x2); cudaFree(w); cudaFree(partial);
    }
    return 0;
}
