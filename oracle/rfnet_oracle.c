/*
 * rfnet_oracle.c -- CPU restatement of the reference's point-cloud operators.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA kernels in rfnet_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (rfnet_b200/) never links, loads or falls back to it.
 *
 * Parity status: PINNED.  The reference holds no golden vectors for this path, so the oracle is
 * pinned against the reference itself: tests/test_oracle_vs_ref.py compares every function below with the reference's
 * own OpKernels compiled unmodified into oracle/_ref/ (CPU kernels here; the reference CUDA kernels, recompiled for
 * sm_100a, on the GPU box), and tests/golden/ holds outputs generated from oracle/_ref by tests/golden/make_golden.py.
 *
 * Each function cites the reference lines it restates.  Paths are relative to the reference repository root.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (see Makefile).  Contraction is OFF so that every fused multiply-add in
 * this file is an explicit fmaf(): "fused" mode reproduces what nvcc emits for the reference's CUDA kernels
 * (mul(dy,dy); fma(dx,dx,.); fma(dz,dz,.) -- checked in the sm_100a PTX of tf_nndistance_g.cu, tf_sampling_g.cu,
 * tf_grouping_g.cu and tf_approxmatch.cu), "unfused" mode reproduces the reference's CPU build (g++ -O2, no -mfma).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RFO_API __attribute__((visibility("default")))

/* Squared distance between candidate c and query q, evaluated the way the reference evaluates x*x+y*y+z*z with
 * x = c.x - q.x etc.:  fused   = fma(dz,dz, fma(dx,dx, dy*dy))      (GPU kernels, nvcc -fmad=true)
 *                      unfused = ((dx*dx) + (dy*dy)) + (dz*dz)      (CPU kernels, g++ without FMA)            */
static inline float sqdist(const float *c, const float *q, int fused) {
    float dx = c[0] - q[0], dy = c[1] - q[1], dz = c[2] - q[2];
    if (fused) return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
    return (dx * dx + dy * dy) + dz * dz;
}

/* ------------------------------------------------------------------------------------------------------------------
 * nn_distance.   pc_distance/tf_nndistance.cpp:21-43 (nnsearch, CPU) and tf_ops/CD/tf_nndistance_g.cu:4-126 (GPU).
 * For every query j of cloud i: smallest squared distance to the m candidates and the FIRST index attaining it
 * (strict '<', seeded by k==0: .cpp:34, _g.cu:28,38,118).
 * ---------------------------------------------------------------------------------------------------------------- */
RFO_API void rfo_nnsearch(int b, int n, int m, const float *queries, const float *cands, float *dist, int *idx, int fused) {
    for (int i = 0; i < b; i++) {
        const float *Q = queries + (size_t)i * n * 3, *C = cands + (size_t)i * m * 3;
        for (int j = 0; j < n; j++) {
            float best = 0.0f;
            int besti = 0;
            for (int k = 0; k < m; k++) {
                float d = sqdist(C + 3 * k, Q + 3 * j, fused);
                if (k == 0 || d < best) { best = d; besti = k; }
            }
            dist[(size_t)i * n + j] = best;
            idx[(size_t)i * n + j] = besti;
        }
    }
}

/* NnDistanceOp::Compute calls the search once per direction: tf_nndistance.cpp:79-80; GPU launcher _g.cu:127-130. */
RFO_API void rfo_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                             float *dist2, int *idx2, int fused) {
    rfo_nnsearch(b, n, m, xyz1, xyz2, dist1, idx1, fused);
    rfo_nnsearch(b, m, n, xyz2, xyz1, dist2, idx2, fused);
}

/* NnDistanceGrad.  pc_distance/tf_nndistance.cpp:122-163 (CPU, sequential) / tf_nndistance_g.cu:131-156 (GPU, atomics).
 * g = 2*grad_dist[j]; grad_self[j] += g*(p - nn); grad_other[nn_idx] -= g*(p - nn); both directions.               */
static void nn_grad_one_dir(int n, int m, const float *P, const float *O, const float *gd, const int *idx, float *gP, float *gO) {
    (void)m;
    for (int j = 0; j < n; j++) {
        int j2 = idx[j];
        float g = gd[j] * 2;
        for (int c = 0; c < 3; c++) {
            float t = g * (P[j * 3 + c] - O[j2 * 3 + c]);
            gP[j * 3 + c] += t;
            gO[j2 * 3 + c] -= t;
        }
    }
}
RFO_API void rfo_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                                  const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1, float *grad_xyz2) {
    memset(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3);
    memset(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3);
    for (int i = 0; i < b; i++) {
        const float *A = xyz1 + (size_t)i * n * 3, *B = xyz2 + (size_t)i * m * 3;
        float *gA = grad_xyz1 + (size_t)i * n * 3, *gB = grad_xyz2 + (size_t)i * m * 3;
        nn_grad_one_dir(n, m, A, B, grad_dist1 + (size_t)i * n, idx1 + (size_t)i * n, gA, gB);
        nn_grad_one_dir(m, n, B, A, grad_dist2 + (size_t)i * m, idx2 + (size_t)i * m, gB, gA);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * approx_match, GPU contract.  pc_distance/tf_approxmatch.cu:1-179.
 * float arithmetic, levels j = start_level .. -2 (the GPU runs start_level = 7, .cu:21; the CPU twin 8, .cpp:31),
 * level = -4^j except 0 at j = -2 (.cu:22-25), match laid out (m, n): match[l*n + k] (.cu:152).
 * multiL/multiR use integer division (.cu:4-10).  __expf(x) is ex2.approx(x*log2e); expf() is used here, the two
 * differ by a few ulp which the 1e-4 EMD tolerance absorbs.
 * ---------------------------------------------------------------------------------------------------------------- */
RFO_API void rfo_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, int start_level) {
    float *remainL = malloc(sizeof(float) * (size_t)(n + m) * 2);
    float *remainR = remainL + n, *ratioL = remainR + m, *ratioR = ratioL + n;
    float multiL, multiR;
    if (n >= m) { multiL = 1; multiR = (float)(n / m); } else { multiL = (float)(m / n); multiR = 1; }
    for (int i = 0; i < b; i++) {
        const float *A = xyz1 + (size_t)i * n * 3, *B = xyz2 + (size_t)i * m * 3;
        float *M = match + (size_t)i * n * m;
        memset(M, 0, sizeof(float) * (size_t)n * m);
        for (int k = 0; k < n; k++) remainL[k] = multiL;
        for (int l = 0; l < m; l++) remainR[l] = multiR;
        for (int j = start_level; j >= -2; j--) {
            float level = -powf(4.0f, (float)j);
            if (j == -2) level = 0;
            /* The multiply-adds below are fused exactly where nvcc fuses them in the reference binary (sm_100a PTX of
             * tf_approxmatch.cu: fma(e, remainR, suml); fma(e, ratioL, sumr); t = ratioL*e, fma(t, ratioR, match), fma(t, ratioR, suml)). */
            /* pass 1 (.cu:26-59): ratioL[k] = remainL[k] / (1e-9 + sum_l e(k,l) * remainR[l]) */
            for (int k = 0; k < n; k++) {
                float suml = 1e-9f;
                for (int l = 0; l < m; l++) suml = fmaf(expf(level * sqdist(B + 3 * l, A + 3 * k, 1)), remainR[l], suml);
                ratioL[k] = remainL[k] / suml;
            }
            /* pass 2 (.cu:75-108): column sums with ratioL, consumption clamp, remainR update */
            for (int l = 0; l < m; l++) {
                float sumr = 0;
                for (int k = 0; k < n; k++) sumr = fmaf(expf(level * sqdist(B + 3 * l, A + 3 * k, 1)), ratioL[k], sumr);
                sumr *= remainR[l];
                float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
            }
            /* pass 3 (.cu:127-160): match += e * ratioL[k] * ratioR[l]; remainL update */
            for (int k = 0; k < n; k++) {
                float suml = 0;
                for (int l = 0; l < m; l++) {
                    float t = ratioL[k] * expf(level * sqdist(B + 3 * l, A + 3 * k, 1));
                    M[(size_t)l * n + k] = fmaf(t, ratioR[l], M[(size_t)l * n + k]);
                    suml = fmaf(t, ratioR[l], suml);
                }
                remainL[k] = fmaxf(0.0f, remainL[k] - suml);
            }
        }
    }
    free(remainL);
}

/* approx_match in DOUBLE precision (same algorithm and level schedule as rfo_approxmatch, exp() in double, no 1-ulp noise).
 * Not a contract of the reference -- a yardstick: |rfo_approxmatch - rfo_approxmatch_f64| is the float32 rounding noise
 * band of this ill-conditioned iteration on a given input, against which the parity tests scale their tolerance. */
RFO_API void rfo_approxmatch_f64(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, int start_level) {
    double *remainL = malloc(sizeof(double) * (size_t)(n + m) * 2);
    double *remainR = remainL + n, *ratioL = remainR + m, *ratioR = ratioL + n;
    double *M = malloc(sizeof(double) * (size_t)n * m);
    double multiL, multiR;
    if (n >= m) { multiL = 1; multiR = (double)(n / m); } else { multiL = (double)(m / n); multiR = 1; }
    for (int i = 0; i < b; i++) {
        const float *A = xyz1 + (size_t)i * n * 3, *B = xyz2 + (size_t)i * m * 3;
        for (size_t t = 0; t < (size_t)n * m; t++) M[t] = 0;
        for (int k = 0; k < n; k++) remainL[k] = multiL;
        for (int l = 0; l < m; l++) remainR[l] = multiR;
        for (int j = start_level; j >= -2; j--) {
            double level = j == -2 ? 0.0 : -pow(4.0, (double)j);
#define D2(k, l) (((double)B[3*(l)]-A[3*(k)])*((double)B[3*(l)]-A[3*(k)]) + ((double)B[3*(l)+1]-A[3*(k)+1])*((double)B[3*(l)+1]-A[3*(k)+1]) + ((double)B[3*(l)+2]-A[3*(k)+2])*((double)B[3*(l)+2]-A[3*(k)+2]))
            for (int k = 0; k < n; k++) {
                double suml = 1e-9;
                for (int l = 0; l < m; l++) suml += exp(level * D2(k, l)) * remainR[l];
                ratioL[k] = remainL[k] / suml;
            }
            for (int l = 0; l < m; l++) {
                double sumr = 0;
                for (int k = 0; k < n; k++) sumr += exp(level * D2(k, l)) * ratioL[k];
                sumr *= remainR[l];
                double consumption = fmin(remainR[l] / (sumr + 1e-9), 1.0);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmax(0.0, remainR[l] - sumr);
            }
            for (int k = 0; k < n; k++) {
                double suml = 0;
                for (int l = 0; l < m; l++) {
                    double w = exp(level * D2(k, l)) * ratioL[k] * ratioR[l];
                    M[(size_t)l * n + k] += w;
                    suml += w;
                }
                remainL[k] = fmax(0.0, remainL[k] - suml);
            }
#undef D2
        }
        for (size_t t = 0; t < (size_t)n * m; t++) match[(size_t)i * n * m + t] = (float)M[t];
    }
    free(M);
    free(remainL);
}

/* approx_match, CPU twin.  pc_distance/tf_approxmatch.cpp:23-84: double accumulation, 11 levels (j = 8..-2),
 * match laid out (n, m): match[k*m + l] (.cpp:44,75).  Keeps the reference's n*m double weight matrix.          */
RFO_API void rfo_approxmatch_cpu_twin(int b, int n, int m, const float *xyz1, const float *xyz2, float *match) {
    int big = n > m ? n : m;
    double *satl = malloc(sizeof(double) * (size_t)(n + 3 * m + n));
    double *satr = satl + n, *ss = satr + m, *ss2 = ss + m, *rows = ss2 + m;
    double *w = malloc(sizeof(double) * (size_t)n * m);
    for (int i = 0; i < b; i++) {
        const float *A = xyz1 + (size_t)i * n * 3, *B = xyz2 + (size_t)i * m * 3;
        float *M = match + (size_t)i * n * m;
        for (int k = 0; k < n; k++) satl[k] = (double)(big / n);
        for (int l = 0; l < m; l++) satr[l] = (double)(big / m);
        memset(M, 0, sizeof(float) * (size_t)n * m);
        for (int j = 8; j >= -2; j--) {
            double level = -powf(4.0f, (float)j);
            if (j == -2) level = 0;
            for (int k = 0; k < n; k++) {
                double x1 = A[k * 3], y1 = A[k * 3 + 1], z1 = A[k * 3 + 2];
                for (int l = 0; l < m; l++) {
                    double x2 = B[l * 3], y2 = B[l * 3 + 1], z2 = B[l * 3 + 2];
                    /* .cpp:44: expf() of a double argument narrows to float first */
                    w[(size_t)k * m + l] = expf((float)(level * ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2)))) * satr[l];
                }
            }
            for (int l = 0; l < m; l++) ss[l] = 1e-9;
            for (int k = 0; k < n; k++) {
                double s = 1e-9;
                for (int l = 0; l < m; l++) s += w[(size_t)k * m + l];
                for (int l = 0; l < m; l++) w[(size_t)k * m + l] = w[(size_t)k * m + l] / s * satl[k];
                for (int l = 0; l < m; l++) ss[l] += w[(size_t)k * m + l];
            }
            for (int l = 0; l < m; l++) { double r = satr[l] / ss[l]; ss[l] = r < 1.0 ? r : 1.0; }
            for (int l = 0; l < m; l++) ss2[l] = 0;
            for (int k = 0; k < n; k++) {
                double s = 0;
                for (int l = 0; l < m; l++) {
                    w[(size_t)k * m + l] *= ss[l];
                    s += w[(size_t)k * m + l];
                    ss2[l] += w[(size_t)k * m + l];
                }
                rows[k] = s;
            }
            for (int k = 0; k < n; k++) { double v = satl[k] - rows[k]; satl[k] = v > 0 ? v : 0; }
            for (size_t t = 0; t < (size_t)n * m; t++) M[t] = (float)((double)M[t] + w[t]);   /* float += double, .cpp:75 */
            for (int l = 0; l < m; l++) { double v = satr[l] - ss2[l]; satr[l] = v > 0 ? v : 0; }
        }
    }
    free(w);
    free(satl);
}

/* match_cost, GPU contract.  pc_distance/tf_approxmatch.cu:183-225: cost[i] = sum_{k,l} sqrtf(d2(k,l)) * match[l*n+k].
 * The reference reduces per-thread partial sums with a tree; here accumulation is in double so the oracle is the
 * more accurate side of a 1e-4 comparison.                                                                      */
RFO_API void rfo_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *cost) {
    for (int i = 0; i < b; i++) {
        const float *A = xyz1 + (size_t)i * n * 3, *B = xyz2 + (size_t)i * m * 3, *M = match + (size_t)i * n * m;
        double s = 0;
        for (int l = 0; l < m; l++)
            for (int k = 0; k < n; k++) s += (double)(sqrtf(sqdist(B + 3 * l, A + 3 * k, 1)) * M[(size_t)l * n + k]);
        cost[i] = (float)s;
    }
}

/* match_cost grad, GPU contract.  pc_distance/tf_approxmatch.cu:229-295.
 * grad1[k] = sum_l match[l*n+k] * (p1_k - p2_l) * rsqrt(max(d2, 1e-20))   (matchcostgrad1, .cu:270-291)
 * grad2[l] = sum_k match[l*n+k] * (p2_l - p1_k) * rsqrt(max(d2, 1e-20))   (matchcostgrad2, .cu:229-269)        */
RFO_API void rfo_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *grad1, float *grad2) {
    double *g1 = malloc(sizeof(double) * (size_t)n * 3);
    for (int i = 0; i < b; i++) {
        const float *A = xyz1 + (size_t)i * n * 3, *B = xyz2 + (size_t)i * m * 3, *M = match + (size_t)i * n * m;
        for (int t = 0; t < n * 3; t++) g1[t] = 0;
        for (int l = 0; l < m; l++) {
            double g2[3] = {0, 0, 0};
            for (int k = 0; k < n; k++) {
                float dx = A[3 * k] - B[3 * l], dy = A[3 * k + 1] - B[3 * l + 1], dz = A[3 * k + 2] - B[3 * l + 2];
                float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                float s = M[(size_t)l * n + k] * (1.0f / sqrtf(fmaxf(d2, 1e-20f)));
                g1[3 * k] += dx * s; g1[3 * k + 1] += dy * s; g1[3 * k + 2] += dz * s;
                g2[0] -= dx * s; g2[1] -= dy * s; g2[2] -= dz * s;
            }
            for (int c = 0; c < 3; c++) grad2[((size_t)i * m + l) * 3 + c] = (float)g2[c];
        }
        for (int t = 0; t < n * 3; t++) grad1[(size_t)i * n * 3 + t] = (float)g1[t];
    }
    free(g1);
}

/* ------------------------------------------------------------------------------------------------------------------
 * farthest_point_sample.  tf_ops/sampling/tf_sampling_g.cu:105-170 (GPU only; there is no reference CPU code).
 * idx[0] = 0; running min-distance temp[k] starts at 1e38 (:118-119); each round updates temp against the last pick
 * (fused d2, :142-145) and takes the arg-max.  Tie rule of the 512-thread block (:146-163): thread t = k mod 512 keeps
 * its first strict maximum starting from (best=-1, besti=0); the tree keeps the LOWER slot unless strictly smaller,
 * so among equal maxima the lowest thread id wins, and within a thread the lowest k.
 * ---------------------------------------------------------------------------------------------------------------- */
RFO_API void rfo_farthest_point_sample(int b, int n, int m, const float *inp, int *idx) {
    enum { BLOCK = 512 };
    if (m <= 0) return;
    float *temp = malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < b; i++) {
        const float *P = inp + (size_t)i * n * 3;
        int old = 0;
        idx[(size_t)i * m] = 0;
        for (int k = 0; k < n; k++) temp[k] = 1e38f;
        for (int j = 1; j < m; j++) {
            float tbest[BLOCK];
            int tbesti[BLOCK];
            for (int t = 0; t < BLOCK; t++) { tbest[t] = -1; tbesti[t] = 0; }
            for (int k = 0; k < n; k++) {
                float d = sqdist(P + 3 * k, P + 3 * old, 1);
                float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                int t = k % BLOCK;
                if (d2 > tbest[t]) { tbest[t] = d2; tbesti[t] = k; }
            }
            float best = tbest[0];
            int besti = tbesti[0];
            for (int t = 1; t < BLOCK; t++)
                if (best < tbest[t]) { best = tbest[t]; besti = tbesti[t]; }
            old = besti;
            idx[(size_t)i * m + j] = old;
        }
    }
    free(temp);
}

/* gather_point / its gradient.  tf_ops/sampling/tf_sampling_g.cu:172-192 (+ cudaMemset at tf_sampling.cpp:174). */
RFO_API void rfo_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
    for (int i = 0; i < b; i++)
        for (int j = 0; j < m; j++) {
            int a = idx[(size_t)i * m + j];
            for (int c = 0; c < 3; c++) out[((size_t)i * m + j) * 3 + c] = inp[((size_t)i * n + a) * 3 + c];
        }
}
RFO_API void rfo_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < m; j++) {
            int a = idx[(size_t)i * m + j];
            for (int c = 0; c < 3; c++) inp_g[((size_t)i * n + a) * 3 + c] += out_g[((size_t)i * m + j) * 3 + c];
        }
}

/* ------------------------------------------------------------------------------------------------------------------
 * query_ball_point.  tf_ops/grouping/tf_grouping_g.cu:3-36 (GPU) / tf_ops/grouping/query_ball_point.cpp:19-47 (CPU
 * prototype, no pts_cnt).  xyz1 = dataset (b,n,3), xyz2 = queries (b,m,3).  Scan dataset in index order, keep the first
 * nsample with max(sqrtf(d2),1e-20f) < radius[0]; the first hit fills the whole row (:26-29).  Rows without a hit are
 * left untouched by the reference (allocator garbage); this oracle writes fill_empty there so tests can mask them.
 * d2 is fused with (query - dataset) differences (:24, PTX above).
 * ---------------------------------------------------------------------------------------------------------------- */
RFO_API void rfo_query_ball_point(int b, int n, int m, const float *radius, int nsample, const float *xyz1, const float *xyz2,
                                  int *idx, int *pts_cnt, int fill_empty) {
    float r = radius[0];
    for (int i = 0; i < b; i++) {
        const float *D = xyz1 + (size_t)i * n * 3, *Q = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < m; j++) {
            int *row = idx + ((size_t)i * m + j) * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; k++) {
                float d = fmaxf(sqrtf(sqdist(Q + 3 * j, D + 3 * k, 1)), 1e-20f);
                if (d < r) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; l++) row[l] = k;
                    row[cnt++] = k;
                }
            }
            if (cnt == 0)
                for (int l = 0; l < nsample; l++) row[l] = fill_empty;
            pts_cnt[(size_t)i * m + j] = cnt;
        }
    }
}

/* group_point / grad.  tf_ops/grouping/tf_grouping_g.cu:40-78 (+ cudaMemset tf_grouping.cpp:208);
 * CPU prototypes tf_ops/grouping/query_ball_point.cpp:52-84.                                                    */
RFO_API void rfo_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out) {
    for (int i = 0; i < b; i++)
        for (size_t t = 0; t < (size_t)m * nsample; t++) {
            int ii = idx[(size_t)i * m * nsample + t];
            memcpy(out + ((size_t)i * m * nsample + t) * c, points + ((size_t)i * n + ii) * c, sizeof(float) * c);
        }
}
RFO_API void rfo_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
    for (int i = 0; i < b; i++)
        for (size_t t = 0; t < (size_t)m * nsample; t++) {
            int ii = idx[(size_t)i * m * nsample + t];
            for (int l = 0; l < c; l++) grad_points[((size_t)i * n + ii) * c + l] += grad_out[((size_t)i * m * nsample + t) * c + l];
        }
}

/* ------------------------------------------------------------------------------------------------------------------
 * three_nn / three_interpolate (+grad).  tf_ops/interpolation/tf_interpolate.cpp:60-153 -- CPU-only ops in the reference.
 * three_nn: the reference's CPU build evaluates d2 UNFUSED in float ((x2-x1)^2 + ... with float operands, widened to
 * double only for the comparisons, :75), keeps the three smallest with strict '<' insertion (:76-92), init 1e40 / 0.
 * xyz1 = unknown (b,n,3) are the queries, xyz2 = known (b,m,3) the candidates.  `fused` selects the contraction so the
 * same oracle can also be asked what a GPU-style evaluation would give.
 * ---------------------------------------------------------------------------------------------------------------- */
RFO_API void rfo_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, int fused) {
    for (int i = 0; i < b; i++) {
        const float *U = xyz1 + (size_t)i * n * 3, *K = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; j++) {
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; k++) {
                double d = sqdist(K + 3 * k, U + 3 * j, fused);
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
            float *dd = dist + ((size_t)i * n + j) * 3;
            int *ii = idx + ((size_t)i * n + j) * 3;
            dd[0] = (float)b1; dd[1] = (float)b2; dd[2] = (float)b3;   /* (float)1e40 == +inf, as in the reference */
            ii[0] = i1; ii[1] = i2; ii[2] = i3;
        }
    }
}

/* tf_interpolate.cpp:107-127: out[j,:] = p[i1,:]*w1 + p[i2,:]*w2 + p[i3,:]*w3 (left to right, unfused on the CPU). */
RFO_API void rfo_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight, float *out) {
    for (int i = 0; i < b; i++)
        for (int j = 0; j < n; j++) {
            const float *w = weight + ((size_t)i * n + j) * 3;
            const int *id = idx + ((size_t)i * n + j) * 3;
            const float *P = points + (size_t)i * m * c;
            for (int l = 0; l < c; l++)
                out[((size_t)i * n + j) * c + l] = (P[(size_t)id[0] * c + l] * w[0] + P[(size_t)id[1] * c + l] * w[1]) + P[(size_t)id[2] * c + l] * w[2];
        }
}

/* tf_interpolate.cpp:131-153 (+ memset :258): grad_points[i_t,:] += grad_out[j,:] * w_t, j ascending. */
RFO_API void rfo_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * m * c);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < n; j++) {
            const float *w = weight + ((size_t)i * n + j) * 3;
            const int *id = idx + ((size_t)i * n + j) * 3;
            float *G = grad_points + (size_t)i * m * c;
            for (int l = 0; l < c; l++) {
                float g = grad_out[((size_t)i * n + j) * c + l];
                G[(size_t)id[0] * c + l] += g * w[0];
                G[(size_t)id[1] * c + l] += g * w[1];
                G[(size_t)id[2] * c + l] += g * w[2];
            }
        }
}

/* knn_point for 3-d points.  tf_ops/grouping/tf_grouping.py:48-73: dist = reduce_sum((xyz1 - xyz2)^2, -1) over the (b, m, n)
 * broadcast, then tf.nn.top_k(-dist, k): values are the negated distances, largest first; equal values keep the lower index
 * first.  xyz1 = dataset (b,n,3), xyz2 = queries (b,m,3). */
RFO_API void rfo_knn_point(int b, int n, int m, int k, const float *xyz1, const float *xyz2, float *val, int *idx) {
    float *d = malloc(sizeof(float) * (size_t)n);
    char *used = malloc((size_t)n);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < m; j++) {
            const float *Q = xyz2 + ((size_t)i * m + j) * 3;
            for (int t = 0; t < n; t++) { d[t] = sqdist(xyz1 + ((size_t)i * n + t) * 3, Q, 0); used[t] = 0; }
            for (int s = 0; s < k; s++) {
                int best = -1;
                for (int t = 0; t < n; t++)
                    if (!used[t] && (best < 0 || d[t] < d[best])) best = t;
                used[best] = 1;
                val[((size_t)i * m + j) * k + s] = -d[best];
                idx[((size_t)i * m + j) * k + s] = best;
            }
        }
    free(used);
    free(d);
}

/* Smallest float T such that sqrtf(T) >= r, i.e. (max(sqrtf(d2),1e-20f) < r)  <=>  (d2 < T) for r > 1e-20f.
 * Used by tests to check the sqrt-free ball-query predicate of the CUDA kernel. */
RFO_API float rfo_ball_threshold(float r) {
    float t = r * r;
    while (sqrtf(t) >= r && t > 0) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < r) t = nextafterf(t, INFINITY);
    return t;
}

/* ------------------------------------------------------------------------------------------------------------------
 * auction_match.   tf_ops/emd/tf_auctionmatch_g.cu:2-291 (GPU only in the reference; `n` <= 4096 there).
 * The kernel serves ONE bidder per loop iteration (the whole 512-thread block cooperates on it), so the auction is a
 * sequential algorithm and this is its restatement, INCLUDING the block's reduction tree, because the tree decides
 * the bid:
 *   cost(i,j) = sqrtf(d2(xyz1[i], xyz2[j]))   with d2 contracted as nvcc does (_g.cu:36; SASS: FMUL dy,dy; FFMA dx; FFMA dz)
 *   queue = 0..n-1, price = 0, matchr = -1, tolerance = 1e-4                                     (_g.cu:13-22,45-48)
 *   pop i; thread t scans its objects j = t, t+512, ... for (best, second best, bestj) of cost(i,j)+price[j], seeded
 *   with 1e38, an equal value replacing the incumbent (_g.cu:55-214: the four unrolled variants all reduce to this);
 *   the 512 triples are merged by a shuffle-down tree inside each warp and again across the 16 warps (_g.cu:216-256):
 *       merge(A = lower lane, B = lane + i):  if (A.best < B.best) { A.best2 = min(B.best, A.best2) }
 *                                             else { A.best = B.best; A.best2 = min(A.best, B.best2); A.bestj = B.bestj }
 *   -- in the else branch A.best has ALREADY been overwritten, so best2 becomes min(B.best, B.best2) = B.best: whenever the
 *   winning object does not belong to thread 0 the "second best" handed to the bid equals the best and
 *   delta = best2 - best + tolerance is exactly `tolerance`.  That is what the reference computes, so it is restated as is;
 *   price[bestj] += delta; the previous owner of bestj is pushed back; matchr[bestj] = i          (_g.cu:258-285)
 *   every 40*n bids: stop if tolerance == 1, else tolerance = min(1, tolerance*100)               (_g.cu:275-280)
 * The reference reads past the end of its rows when 1024 <= n < 4096 is not 1024 or 2048 (_g.cu:155-214 step by 2 or 4
 * block widths without a bound check); for those n this restatement simply skips the objects that do not exist.
 * Bidders left without an object when the auction gives up keep matchl = -1 (the reference writes matchl[-1] there).
 * ---------------------------------------------------------------------------------------------------------------- */
#define AUC_T 512
typedef struct { float best, best2; int bestj; } auc_triple;
static void auc_merge(auc_triple *a, const auc_triple *b) {   /* _g.cu:220-231 == :243-254, statement for statement */
    const float b1 = b->best, b2 = b->best2;
    const int bj = b->bestj;
    if (a->best < b1) {
        a->best2 = fminf(b1, a->best2);
    } else {
        a->best = b1;
        a->best2 = fminf(a->best, b2);
        a->bestj = bj;
    }
}
/* shuffle-down tree over `width` lanes (width = 32 or 16): lanes whose partner is outside keep reading themselves */
static void auc_tree(auc_triple *t, int width) {
    auc_triple old[32];
    for (int i = width >> 1; i > 0; i >>= 1) {
        memcpy(old, t, sizeof(auc_triple) * (size_t)width);
        for (int l = 0; l < width; l++) auc_merge(&t[l], &old[l + i < width ? l + i : l]);
    }
}
RFO_API void rfo_auction_match(int b, int n, const float *xyz1, const float *xyz2, int *matchl, int *matchr) {
    if (n <= 0) return;
    int *queue = (int *)malloc(sizeof(int) * (size_t)n);
    float *price = (float *)malloc(sizeof(float) * (size_t)n);
    auc_triple th[AUC_T], warps[32];
    for (int bno = 0; bno < b; bno++) {
        const float *P = xyz1 + (size_t)bno * n * 3, *Q = xyz2 + (size_t)bno * n * 3;
        int *ml = matchl + (size_t)bno * n, *mr = matchr + (size_t)bno * n;
        for (int j = 0; j < n; j++) { ml[j] = -1; mr[j] = -1; queue[j] = j; price[j] = 0.0f; }
        int qhead = 0, qlen = n, cnt = 0;
        float tolerance = 1e-4f;
        while (qlen) {
            const int i = queue[qhead];
            for (int t = 0; t < AUC_T; t++) {
                auc_triple r = {1e38f, 1e38f, 0};
                for (int j = t; j < n; j += AUC_T) {
                    /* x1 - x2 with 1 = bidder (buf), 2 = object: _g.cu:30-36 */
                    const float dx = P[i * 3] - Q[j * 3], dy = P[i * 3 + 1] - Q[j * 3 + 1], dz = P[i * 3 + 2] - Q[j * 3 + 2];
                    const float value = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy))) + price[j];
                    if (r.best < value) r.best2 = fminf(r.best2, value);      /* _g.cu:205-212 */
                    else { r.best2 = r.best; r.bestj = j; r.best = value; }
                }
                th[t] = r;
            }
            for (int w = 0; w < AUC_T / 32; w++) { auc_tree(th + 32 * w, 32); warps[w] = th[32 * w]; }
            auc_tree(warps, AUC_T / 32);
            const float best = warps[0].best, best2 = warps[0].best2;
            const int bestj = warps[0].bestj;
            const float delta = best2 - best + tolerance;
            qhead++; qlen--;
            if (qhead >= n) qhead -= n;
            const int old = mr[bestj];
            price[bestj] += delta;
            cnt++;
            if (old != -1) {
                int tail = qhead + qlen;
                qlen++;
                if (tail >= n) tail -= n;
                queue[tail] = old;
            }
            if (cnt == 40 * n) {
                if (tolerance == 1.0f) qlen = 0;
                tolerance = fminf(1.0f, tolerance * 100);
                cnt = 0;
            }
            mr[bestj] = i;
        }
        for (int j = 0; j < n; j++) if (mr[j] >= 0) ml[mr[j]] = j;
    }
    free(queue); free(price);
}

/* ------------------------------------------------------------------------------------------------------------------
 * selection_sort (select_top_k).   tf_ops/grouping/tf_grouping_g.cu:83-123.
 * out = copy of dist, outi = 0..n-1 per row; then k steps of selection sort WITH SWAPS: step s finds the first minimum
 * of out[s..n) (strict '<' from min = s) and swaps it into position s, values and indices alike.  The first k entries
 * are the k smallest in ascending order; the tail keeps the displaced entries where the swaps left them.
 * ---------------------------------------------------------------------------------------------------------------- */
RFO_API void rfo_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
    if (k > n) k = n;
    for (size_t row = 0; row < (size_t)b * m; row++) {
        const float *src = dist + row * n;
        float *o = out + row * n;
        int *oi = outi + row * n;
        for (int s = 0; s < n; s++) { o[s] = src[s]; oi[s] = s; }
        for (int s = 0; s < k; s++) {
            int min = s;
            for (int t = s + 1; t < n; t++) if (o[t] < o[min]) min = t;
            if (min != s) {
                const float tv = o[min]; o[min] = o[s]; o[s] = tv;
                const int ti = oi[min]; oi[min] = oi[s]; oi[s] = ti;
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Check of the error bound behind rfnet_b200's filtered nearest-neighbour search (csrc/nn_distance.cu: the uniform bound
 * E = 2^-20 (|q-o| + max|c-o|)^2 of its first version; the certificate in use, nn_filter_certain, is checked further below).
 * NOT a restatement of the reference: it replays, in the kernel's float32 operation order, the expanded form
 *     s = fma(-2(qx-ox), cx-ox, fma(-2(qy-oy), cy-oy, fma(-2(qz-oz), cz-oz, |c-o|^2))),  |c-o|^2 = fma(z,z, fma(x,x, y*y))
 * for n queries x m candidates about the origin o, and compares  s + |q-o|^2  (the squared norm taken in double from the
 * ROUNDED differences, as the derivation does) with the reference distance expression d2 (fused or unfused, sqdist above).
 * Returns max over all pairs of  |s + |q~|^2 - d2| / E,  E = 2^-20 (|q~| + max|c~|)^2 + 5e-31  -- the kernel certifies with 2E,
 * so the bound holds iff the result is <= 1.
 * ------------------------------------------------------------------------------------------------------------- */
double rfo_nn_filter_bound_ratio(int n, const float *q, int m, const float *c, const float *origin, int fused) {
    const float ox = origin[0], oy = origin[1], oz = origin[2];
    double cmax = 0.0;
    for (int j = 0; j < m; j++) {
        const float x = c[3 * j] - ox, y = c[3 * j + 1] - oy, z = c[3 * j + 2] - oz;
        const float cn = fmaf(z, z, fmaf(x, x, y * y));
        if ((double)cn > cmax) cmax = (double)cn;
    }
    double worst = 0.0;
    for (int i = 0; i < n; i++) {
        const float rx = q[3 * i] - ox, ry = q[3 * i + 1] - oy, rz = q[3 * i + 2] - oz;   /* the kernel's roundings */
        const float ax = -2.0f * rx, ay = -2.0f * ry, az = -2.0f * rz;
        const double qn = (double)rx * rx + (double)ry * ry + (double)rz * rz;
        const double L = sqrt(qn) + sqrt(cmax);
        const double E = ldexp(L * L, -20) + 5e-31;
        for (int j = 0; j < m; j++) {
            const float x = c[3 * j] - ox, y = c[3 * j + 1] - oy, z = c[3 * j + 2] - oz;
            const float cn = fmaf(z, z, fmaf(x, x, y * y));
            const float s = fmaf(ax, x, fmaf(ay, y, fmaf(az, z, cn)));
            const float d2 = sqdist(c + 3 * j, q + 3 * i, fused);
            const double r = fabs((double)s + qn - (double)d2) / E;
            if (r > worst) worst = r;
        }
    }
    return worst;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Check of the CERTIFICATE of the filtered search (csrc/nn_distance.cu: nn_filter_certain), replayed in double on the float32
 * values the kernel would see.  For every query: c1 = the candidate with the smallest scanned value s (b1), and for every other
 * candidate c (b2 := s(c)) the kernel's test
 *      sqrt(X) - sqrt(Y) > 9u L,   X = b2 - A + |q~|^2,  Y = max(0, b1 + A + |q~|^2),  A = 6.1u (Cmax^2 + |q~| Cmax),  L = |q~| + Cmax
 * is evaluated; whenever it holds, the reference distances must satisfy d2_ref(c) > d2_ref(c1).  Returns the number of violations
 * (must be 0) and, through *certified, how many pairs the test certified (so that a vacuous pass is visible); *tightest gets the
 * smallest (d2_ref(c) - d2_ref(c1)) / d2_ref(c1) over certified pairs.
 * ------------------------------------------------------------------------------------------------------------- */
long rfo_nn_filter_certificate_violations(int n, const float *q, int m, const float *c, const float *origin, int fused, long *certified,
                                          double *tightest) {
    const double u = ldexp(1.0, -24);
    const float ox = origin[0], oy = origin[1], oz = origin[2];
    float *s = (float *)malloc(sizeof(float) * (size_t)m), *d = (float *)malloc(sizeof(float) * (size_t)m);
    double cmax = 0.0;
    for (int j = 0; j < m; j++) {
        const float x = c[3 * j] - ox, y = c[3 * j + 1] - oy, z = c[3 * j + 2] - oz;
        const float cn = fmaf(z, z, fmaf(x, x, y * y));
        if ((double)cn > cmax) cmax = (double)cn;
    }
    const double Cm = sqrt(cmax * (1.0 + 4.0 * u));
    long bad = 0, cert = 0;
    double tight = 1e300;
    for (int i = 0; i < n; i++) {
        const float rx = q[3 * i] - ox, ry = q[3 * i + 1] - oy, rz = q[3 * i + 2] - oz;
        const float ax = -2.0f * rx, ay = -2.0f * ry, az = -2.0f * rz;
        const double qn = (double)rx * rx + (double)ry * ry + (double)rz * rz;
        const double ql = sqrt(qn), L = ql + Cm;
        const double A = 6.1 * u * (Cm * Cm + ql * Cm) + 1e-35;
        int j1 = 0;
        for (int j = 0; j < m; j++) {
            const float x = c[3 * j] - ox, y = c[3 * j + 1] - oy, z = c[3 * j + 2] - oz;
            const float cn = fmaf(z, z, fmaf(x, x, y * y));
            s[j] = fmaf(ax, x, fmaf(ay, y, fmaf(az, z, cn)));
            d[j] = sqdist(c + 3 * j, q + 3 * i, fused);
            if (s[j] < s[j1]) j1 = j;
        }
        const double Y0 = (double)s[j1] + A + qn, Y = Y0 > 0.0 ? Y0 : 0.0;
        for (int j = 0; j < m; j++) {
            if (j == j1) continue;
            const double X = (double)s[j] - A + qn;
            if (X > 0.0 && sqrt(X) - sqrt(Y) > 9.0 * u * L) {
                cert++;
                if (!(d[j] > d[j1])) bad++;
                else if (d[j1] > 0.f) { const double t = ((double)d[j] - (double)d[j1]) / (double)d[j1]; if (t < tight) tight = t; }
            }
        }
    }
    free(s); free(d);
    *certified = cert;
    *tightest = tight;
    return bad;
}
