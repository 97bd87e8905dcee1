// micro-kernel: the adjoint's data flow for config #2 with a fully static schedule (R = 2 rows per dst row)
#include <cuda_runtime.h>
#include <cstdio>
constexpr int NT = 256, TPG = 128, PT = 256, PH = 128, TD2 = 32, RG = 16, S2 = 4, R = 2, NB = 8;
constexpr int PLANE = TD2 * PH + 16;
template <int MODE>
__global__ void __launch_bounds__(256, 4) adj_micro(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ w2g, int n1, int m1, int tiles1)
{
    __shared__ float sT[2 * PLANE];
    __shared__ __align__(16) float sW2[2][48][4];
    int tile1 = blockIdx.x % tiles1, tile2 = blockIdx.x / tiles1;
    int tid = threadIdx.x, g = tid / TPG, ct = tid % TPG;
    int c0 = tile1 * 120, d0 = tile2 * TD2;
    int r1lo = 2 * c0;                        // src columns [r1lo, r1lo + 246)
    int ka = d0 + g * RG, kb = ka + RG, v0 = ka - 3;
    int i2lo = max(0, 2 * v0);                // src rows: 2 per dst row
    int nrow = 2 * (kb - v0); if (i2lo + nrow > n1) nrow = n1 - i2lo;
    int col = r1lo + 2 * ct;
    bool have = col + 1 < n1 && 2 * ct < 246;
    const float2* gp = reinterpret_cast<const float2*>(src + (size_t)i2lo * n1 + col);
    size_t step = n1 / 2;
    int togo = have ? nrow : 0;
    auto load = [&]() { float2 y = make_float2(0.f, 0.f); if (togo > 0) y = *gp; --togo; gp += step; return y; };
    float2 q[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) q[j] = load();
    for (int k = tid; k < 2 * 48 * 4; k += NT) (&sW2[0][0][0])[k] = w2g[k & 1023];
    __syncthreads();
    float A0[S2] = {0, 0, 0, 0}, A1[S2] = {0, 0, 0, 0};
    const float4* pw = reinterpret_cast<const float4*>(&sW2[g][0][0]);
    float* tp = sT + (ka - d0) * PH + ct;
    for (int v = v0; v < kb; v += NB / R) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            int vv = v + j / R;
            if (vv < kb) {
                float2 y = q[j]; q[j] = load();
                float4 w = *pw++;
                float wv[4] = {w.x, w.y, w.z, w.w};
                if (MODE >= 1) {
#pragma unroll
                    for (int s = 0; s < S2; ++s) { A0[(j / R + s) % S2] += wv[s] * y.x; A1[(j / R + s) % S2] += wv[s] * y.y; }
                } else { A0[0] += y.x; A1[0] += y.y; }
                if (j % R == R - 1) {
                    if (vv >= ka) { tp[0] = A0[(j / R) % S2]; tp[PLANE] = A1[(j / R) % S2]; tp += PH; }
                    A0[(j / R) % S2] = A1[(j / R) % S2] = 0.f;
                }
            }
        }
    }
    __syncthreads();
    if (MODE >= 2) {
        int kc = tid % 128, sl = tid / 128;
        if (kc < 120) {
            int k1 = c0 + kc; int e0 = 2 * kc;   // 8 taps: staged columns e0 .. e0+7
            float acc[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) acc[r] = 0.f;
            const float* pE = sT + sl * 16 * PH + (e0 >> 1);
            const float* pO = pE + PLANE;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float wk = w2g[(k1 * 8 + k) & 1023];
                const float* p = (k & 1) ? pO + k / 2 : pE + k / 2;
#pragma unroll
                for (int r = 0; r < 16; ++r) acc[r] += wk * p[r * PH];
            }
            float* d = dst + (size_t)(d0 + sl * 16) * m1 + k1;
#pragma unroll
            for (int r = 0; r < 16; ++r) d[(size_t)r * m1] = acc[r];
        }
    } else {
        if (tid == 0 && sT[5] == 12345.f) dst[0] = 1.f;
    }
}
int main()
{
    const int n = 8192, m = 4096; size_t bytes = (size_t)n * n * 4;
    float *a, *b, *out, *w; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, (size_t)m * m * 4 + 4096); cudaMalloc(&w, 4096);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes); cudaMemset(w, 0, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto f, const char* name) {
        float best = 1e9;
        for (int it = 0; it < 6; ++it) { const float* s = (it & 1) ? a : b; cudaEventRecord(e0); f(s); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < best) best = ms; }
        printf("%-60s %8.1f us  err=%s\n", name, best * 1e3, cudaGetErrorString(cudaGetLastError()));
    };
    time([&](const float* s) { adj_micro<0><<<35 * 128, 256>>>(s, out, w, n, m, 35); }, "loads + queue + stores to sT (no FMAs, no phase 2)");
    time([&](const float* s) { adj_micro<1><<<35 * 128, 256>>>(s, out, w, n, m, 35); }, "+ weights LDS + 8 FMAs per row");
    time([&](const float* s) { adj_micro<2><<<35 * 128, 256>>>(s, out, w, n, m, 35); }, "+ phase 2 (8 taps x 16 rows) + dst stores");
    return 0;
}
