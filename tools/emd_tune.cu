// Tuning harness for the EMD sweep kernel: times emd_sweep_kernel directly for several (Q, candidates per split) at fixed
// shapes, for the single-level sweep (MODE 1) and the fused pass-3 + pass-1 sweep (MODE 4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I rfnet_b200/csrc -o /tmp/emd_tune tools/emd_tune.cu
#include <cstdio>
#include <vector>
#include "../rfnet_b200/csrc/approxmatch.cu"

template <int Q, int MODE>
float run(int b, int n, int split_len, float lvl2, const float* x1, const float* x2, const float* w, float* partial, float* scratch) {
    SweepArgs a;
    a.nr = n; a.nc = n; a.nrt = (n + EMD_THREADS * Q - 1) / (EMD_THREADS * Q);
    a.split_len = split_len; a.nsplit = (n + split_len - 1) / split_len; a.nwords = (n + 31) / 32;
    a.lvl2 = lvl2; a.lvl2b = lvl2 * 0.25f; a.init0 = 1e-9f;
    a.rows = x1; a.cands = x2; a.w = w; a.wb = w; a.rowfac = w;
    a.partial = partial; a.partial_b = partial + (size_t)b * n * a.nsplit;
    a.remain = scratch; a.ratio = scratch + (size_t)b * n; a.fac = scratch + 2 * (size_t)b * n;
    a.perm = nullptr; a.mask = nullptr;
    const unsigned grid = (unsigned)(b * a.nrt * a.nsplit);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) emd_sweep_kernel<Q, MODE, false><<<grid, EMD_THREADS>>>(a);
    cudaEventRecord(e0);
    const int it = 5;
    for (int i = 0; i < it; ++i) emd_sweep_kernel<Q, MODE, false><<<grid, EMD_THREADS>>>(a);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    const double pp = (double)b * n * n * (MODE == 4 ? 2 : 1);
    printf("  mode %d Q=%d split_len=%5d nsplit=%3d grid=%6u (%5.1f CTAs/SM): %8.3f ms  %.2f Tpair-pass/s  MUFU %.1f%%\n", MODE, Q, split_len, a.nsplit, grid,
           grid / 148.0, ms, pp / ms / 1e9, 100.0 * pp / (ms * 1e-3) / (148.0 * 16 * 1.965e9));
    return ms;
}

template <int MODE, int NT, int R = 1>
float run_row(int b, int n, float lvl2, const float* x1, const float* x2, const float* w, float* scratch, float4* pA, float2* pZ) {
    SweepArgs a;
    a.nr = n; a.nc = n; a.nrt = (n + NT * R - 1) / (NT * R); a.nsplit = 1; a.split_len = n; a.nwords = (n + 31) / 32; a.npad = emd_npad(n); a.tma = (n % 4 == 0);
    a.lvl2 = lvl2; a.lvl2b = lvl2 * 0.25f; a.init0 = 1e-9f;
    a.rows = x1; a.cands = x2; a.w = w; a.wb = w; a.rowfac = w; a.pairA = pA; a.pairZ = pZ;
    a.partial = nullptr; a.partial_b = nullptr;
    a.remain = scratch; a.ratio = scratch + (size_t)b * n; a.fac = scratch + 2 * (size_t)b * n;
    a.perm = nullptr; a.mask = nullptr;
    const unsigned grid = (unsigned)(b * a.nrt);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) emd_row_kernel<MODE, false, false, NT, R><<<grid, NT>>>(a);
    cudaEventRecord(e0);
    const int it = 5;
    for (int i = 0; i < it; ++i) emd_row_kernel<MODE, false, false, NT, R><<<grid, NT>>>(a);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    const double pp = (double)b * n * n * (MODE == 4 ? 2 : 1);
    printf("  mode %d ROW kernel NT=%3d rows/thread %d unroll %d   grid=%6u (%5.1f CTAs/SM): %8.3f ms  %.2f Tpair-pass/s  MUFU %.1f%%\n", MODE, NT, R, EMD_ROW_UNROLL, grid,
           grid / 148.0, ms, pp / ms / 1e9, 100.0 * pp / (ms * 1e-3) / (148.0 * 16 * 1.965e9));
    return ms;
}

int main() {
    const int shapes[7][2] = {{32, 2048}, {4, 2048}, {32, 1024}, {32, 4096}, {1, 16384}, {4, 16384}, {32, 16384}};
    printf("EMD_UNROLL=%d EMD_TC=%d\n", EMD_UNROLL, EMD_TC);
    for (auto& sh : shapes) {
        const int b = sh[0], n = sh[1];
        std::vector<float> h((size_t)b * n * 3), hw((size_t)b * n, 1.0f);
        unsigned s = 12345;
        for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (s >> 8) / 16777216.0f - 0.5f; }
        float *x1, *x2, *w, *partial, *scratch;
        cudaMalloc(&x1, h.size() * 4); cudaMalloc(&x2, h.size() * 4); cudaMalloc(&w, hw.size() * 4);
        cudaMalloc(&partial, (size_t)b * n * 2 * 256 * 4); cudaMalloc(&scratch, (size_t)b * n * 3 * 4);
        cudaMemcpy(x1, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (s >> 8) / 16777216.0f - 0.5f; }
        cudaMemcpy(x2, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(scratch, 0, (size_t)b * n * 3 * 4);
        printf("b=%d n=m=%d, level -64\n", b, n);
        const float lvl2 = -64.0f * LOG2E;
        float4* pA; float2* pZ;
        cudaMalloc(&pA, (size_t)b * emd_npad(n) * 16); cudaMalloc(&pZ, (size_t)b * emd_npad(n) * 8);
        emd_pairs_kernel<<<dim3((unsigned)((emd_npad(n) + 255) / 256), (unsigned)b), 256>>>(n, emd_npad(n), x2, pA, pZ);
        run_row<1, 64>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        run_row<1, 32>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        run_row<1, 64, 2>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        run_row<1, 32, 2>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        run_row<4, 64>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        run_row<4, 32>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        run_row<4, 64, 2>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        run_row<4, 32, 2>(b, n, lvl2, x1, x2, w, scratch, pA, pZ);
        cudaFree(pA); cudaFree(pZ);
        for (int sl : {512}) {
            if (sl > n || (n + sl - 1) / sl > 256) continue;
            run<2, 1>(b, n, sl, lvl2, x1, x2, w, partial, scratch);
            run<4, 1>(b, n, sl, lvl2, x1, x2, w, partial, scratch);
            run<2, 4>(b, n, sl, lvl2, x1, x2, w, partial, scratch);
            run<4, 4>(b, n, sl, lvl2, x1, x2, w, partial, scratch);
        }
        cudaFree(x1); cudaFree(x2); cudaFree(w); cudaFree(partial); cudaFree(scratch);
    }
    return 0;
}
