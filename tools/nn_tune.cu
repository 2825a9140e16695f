// Tuning harness for nn_search_kernel: compile with -DNN_THREADS_VALUE=.. -DNN_MIN_CTAS=.. and time rfnet_nn_distance.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I rfnet_b200/csrc -DNN_THREADS_VALUE=128 -DNN_MIN_CTAS=5 -o /tmp/nn_tune tools/nn_tune.cu
#include <cstdio>
#include <vector>
#include "../rfnet_b200/csrc/nn_distance.cu"

int main() {
    const int shapes[3][3] = {{32, 2048, 16384}, {32, 16384, 16384}, {4, 16384, 16384}};
    printf("NN_THREADS=%d NN_MIN_CTAS=%d\n", NN_THREADS, NN_MIN_CTAS);
    for (auto& sh : shapes) {
        const int b = sh[0], n = sh[1], m = sh[2];
        std::vector<float> h1((size_t)b * n * 3), h2((size_t)b * m * 3);
        unsigned s = 777;
        for (auto& v : h1) { s = s * 1664525u + 1013904223u; v = (s >> 8) / 16777216.0f - 0.5f; }
        for (auto& v : h2) { s = s * 1664525u + 1013904223u; v = (s >> 8) / 16777216.0f - 0.5f; }
        float *x1, *x2, *d1, *d2; int *i1, *i2; void* ws;
        const size_t wsb = rfnet_nn_distance_workspace_bytes(b, n, m);
        cudaMalloc(&x1, h1.size() * 4); cudaMalloc(&x2, h2.size() * 4); cudaMalloc(&d1, (size_t)b * n * 4); cudaMalloc(&d2, (size_t)b * m * 4);
        cudaMalloc(&i1, (size_t)b * n * 4); cudaMalloc(&i2, (size_t)b * m * 4); cudaMalloc(&ws, wsb);
        cudaMemcpy(x1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(x2, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) rfnet_nn_distance(b, n, x1, m, x2, d1, i1, d2, i2, ws, wsb, 0, 0);
        cudaEventRecord(e0);
        const int it = 10;
        for (int i = 0; i < it; ++i) rfnet_nn_distance(b, n, x1, m, x2, d1, i1, d2, i2, ws, wsb, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= it;
        const double pairs = 2.0 * b * n * m;
        printf("  b=%d n=%d m=%d: %.3f ms  %.0f Gpairs/s  %.1f%% of FP32 lane peak\n", b, n, m, ms, pairs / ms / 1e6, pairs * 6 / (ms * 1e-3) / (148.0 * 128 * 1.965e9) * 100);
        cudaFree(x1); cudaFree(x2); cudaFree(d1); cudaFree(d2); cudaFree(i1); cudaFree(i2); cudaFree(ws);
    }
    return 0;
}
