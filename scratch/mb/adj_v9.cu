// adjoint micro-kernel, evolved step by step towards the library kernel: real (table-driven) geometry for config #2
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cmath>
#include <algorithm>
constexpr int NT = 256, TPG = 128, PT = 256, PH = 128, TD2 = 32, RG = 16, S1 = 4, S2 = 4, R = 2, NB = 8, MAXE = 12;
constexpr int PLANE = TD2 * PH + 16;
struct Args { const int *base1, *bp1, *bp2; const float *W1, *W2f, *src; float* dst; int n1, n2, m1, m2, tiles1, TD1; };
template <int MODE>
__global__ void __launch_bounds__(256, 4) adj_v9(const Args a)
{
    __shared__ float sT[2 * PLANE];
    __shared__ __align__(16) float sW2[2][40][4];
    const int tile1 = blockIdx.x % a.tiles1, tile2 = blockIdx.x / a.tiles1;
    const int tid = threadIdx.x, g = tid / TPG, ct = tid % TPG;
    const int c0 = tile1 * a.TD1, c1 = min(c0 + a.TD1, a.m1), d0 = tile2 * TD2;
    const int ka = d0 + g * RG, kb = ka + RG, v0 = ka - (S2 - 1);
    // round 1: every index lookup
    const int r1lo_raw = a.bp1[c0], r1hi = a.bp1[c1 - 1 + S1];
    const int i2lo = a.bp2[ka];
    const int kc = tid % 128, sl = tid / 128, k1 = c0 + kc;
    const bool have_col = k1 < c1;
    const int pe0 = a.bp1[min(k1, a.m1 - 1)], pe1 = a.bp1[min(k1, a.m1 - 1) + S1];
    const int r1lo = r1lo_raw & ~1, ncol = r1hi - r1lo;
    const int nrow = R * (kb - v0);
    const int col = r1lo + 2 * ct;
    const bool have = 2 * ct < ncol;
    const float2* gp = reinterpret_cast<const float2*>(a.src + (size_t)i2lo * a.n1 + col);
    const size_t step = a.n1 / 2;
    int togo = have ? nrow : 0;
    auto load = [&]() { float2 y = make_float2(0.f, 0.f); if (togo > 0) y = *gp; --togo; gp += step; return y; };
    // round 2: first rows, dim-2 weights, phase-2 taps (straight from global: they are consumed much later)
    float2 q[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) q[j] = load();
    for (int k = ct; k < nrow; k += TPG) *reinterpret_cast<float4*>(&sW2[g][k][0]) = *reinterpret_cast<const float4*>(a.W2f + (size_t)(i2lo + k) * S2);
    const int e0 = pe0 - r1lo, cnt = min(pe1 - pe0, MAXE);
    float tw[MAXE];
#pragma unroll
    for (int k = 0; k < MAXE; ++k) {
        tw[k] = 0.f;
        if (MODE >= 2 && have_col && k < cnt) { const int i1 = pe0 + k; tw[k] = a.W1[(size_t)i1 * S1 + (k1 - a.base1[i1])]; }
    }
    __syncthreads();
    float A0[S2] = {0, 0, 0, 0}, A1[S2] = {0, 0, 0, 0};
    const float4* pw = reinterpret_cast<const float4*>(&sW2[g][0][0]);
    float* tp = sT + (ka - d0) * PH + ct;
    for (int v = v0; v < kb; v += NB / R) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int vv = v + j / R;
            if (vv < kb) {
                const float2 y = q[j]; q[j] = load();
                const float4 w = *pw++;
                const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int s = 0; s < S2; ++s) { A0[(j / R + s) % S2] += wv[s] * y.x; A1[(j / R + s) % S2] += wv[s] * y.y; }
                if (j % R == R - 1) {
                    if (vv >= ka) { tp[0] = A0[(j / R) % S2]; tp[PLANE] = A1[(j / R) % S2]; tp += PH; }
                    A0[(j / R) % S2] = A1[(j / R) % S2] = 0.f;
                }
            }
        }
    }
    __syncthreads();
    if (MODE >= 2) {
        if (have_col) {
            float acc[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) acc[r] = 0.f;
            const int q0 = e0 & 1;
            const float* pE = sT + sl * 16 * PH + q0 * PLANE + (e0 >> 1);
            const float* pO = sT + sl * 16 * PH + (1 - q0) * PLANE + (e0 >> 1) + q0;
#pragma unroll
            for (int k = 0; k < MAXE; ++k) {
                if (k < cnt) {
                    const float* p = (k & 1) ? pO + (k - 1) / 2 : pE + k / 2;
#pragma unroll
                    for (int r = 0; r < 16; ++r) acc[r] += tw[k] * p[r * PH];
                }
            }
            float* d = a.dst + (size_t)(d0 + sl * 16) * a.m1 + k1;
#pragma unroll
            for (int r = 0; r < 16; ++r) d[(size_t)r * a.m1] = acc[r];
        }
    } else if (tid == 0 && sT[5] == 12345.f) a.dst[0] = 1.f;
}
int main()
{
    const int n = 8192, m = 4096; size_t bytes = (size_t)n * n * 4;
    std::vector<int> base(n), bp(m + 4), bpf(m + 4);
    for (int i = 0; i < n; ++i) base[i] = (int)std::floor(1.0 + i * 4095.0 / 8191.0) - 2;
    std::vector<int> bf(n); for (int i = 0; i < n; ++i) bf[i] = std::min(std::max(base[i], 0), m - 4);
    for (int j = 0; j < m + 4; ++j) { bp[j] = (int)(std::lower_bound(base.begin(), base.end(), j - 3) - base.begin()); bpf[j] = (int)(std::lower_bound(bf.begin(), bf.end(), j - 3) - bf.begin()); }
    std::vector<float> w(n * 4, 0.25f);
    int *dbase, *dbp, *dbpf; float *dW1, *dW2, *a, *b, *out;
    cudaMalloc(&dbase, n * 4); cudaMalloc(&dbp, (m + 4) * 4); cudaMalloc(&dbpf, (m + 4) * 4); cudaMalloc(&dW1, n * 16); cudaMalloc(&dW2, n * 16);
    cudaMemcpy(dbase, base.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dbp, bp.data(), (m + 4) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dbpf, bpf.data(), (m + 4) * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW1, w.data(), n * 16, cudaMemcpyHostToDevice); cudaMemcpy(dW2, w.data(), n * 16, cudaMemcpyHostToDevice);
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, (size_t)m * m * 4 + 4096);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    Args A{dbase, dbp, dbpf, dW1, dW2, a, out, n, n, m, m, 35, 120};
    auto time = [&](auto f, const char* name) {
        float best = 1e9;
        for (int it = 0; it < 6; ++it) { A.src = (it & 1) ? a : b; cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < best) best = ms; }
        printf("%-60s %8.1f us  err=%s\n", name, best * 1e3, cudaGetErrorString(cudaGetLastError()));
    };
    // interior tile rows only (1..126), like the regular launch of the library
    time([&]() { adj_v9<1><<<35 * 128, 256>>>(A); }, "table-driven geometry, phase 1 only");
    time([&]() { adj_v9<2><<<35 * 128, 256>>>(A); }, "table-driven geometry, phase 1 + 2");
    return 0;
}
