// micro-benchmark: read an 8192 x 8192 float array in the adjoint kernel's tile pattern
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int PF>
__global__ void __launch_bounds__(256, 4) tile_read(const float* __restrict__ src, float* __restrict__ out, int n1, int rows_per_group, int halo, int cols_per_tile, int tiles1)
{
    int tile1 = blockIdx.x % tiles1, tile2 = blockIdx.x / tiles1;
    int g = threadIdx.x / 128, ct = threadIdx.x % 128;
    int col = tile1 * cols_per_tile + 2 * ct;
    int row0 = tile2 * (2 * rows_per_group) + g * rows_per_group;
    int nrow = rows_per_group + halo;
    if (row0 + nrow > n1) nrow = n1 - row0;
    if (col + 1 >= n1) return;
    const float2* p = reinterpret_cast<const float2*>(src + (size_t)row0 * n1 + col);
    size_t step = n1 / 2;
    float2 acc[PF];
    for (int u = 0; u < PF; ++u) acc[u] = make_float2(0.f, 0.f);
    for (int r = 0; r < nrow; r += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (r + u < nrow) { float2 v = p[(size_t)(r + u) * step]; acc[u].x += v.x; acc[u].y += v.y; }
        }
    }
    float s = 0; for (int u = 0; u < PF; ++u) s += acc[u].x + acc[u].y;
    if (s == 12345.678f) out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void stream_read(const float4* __restrict__ src, float* out, size_t n4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float s = 0;
    for (; i < n4; i += stride) { float4 v = src[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 12345.678f) out[0] = s;
}
int main()
{
    const int n = 8192; size_t bytes = (size_t)n * n * 4;
    float *a, *b, *out; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, 1 << 24);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto f, const char* name) {
        float best = 1e9;
        for (int it = 0; it < 6; ++it) { const float* s = (it & 1) ? a : b; cudaEventRecord(e0); f(s); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < best) best = ms; }
        printf("%-50s %8.1f us  %7.1f GB/s\n", name, best * 1e3, bytes / best / 1e6);
    };
    time([&](const float* s) { stream_read<<<148 * 8, 512>>>((const float4*)s, out, bytes / 16); }, "stream float4 grid-stride");
    // adjoint pattern: 240 src cols per tile (120 dst), groups of 32 src rows + 6 halo
    time([&](const float* s) { tile_read<8><<<35 * 128, 256>>>(s, out, n, 32, 6, 240, 35); }, "tile 256thr: 1KB x (32+6) rows, PF8");
    time([&](const float* s) { tile_read<4><<<35 * 128, 256>>>(s, out, n, 32, 6, 240, 35); }, "tile 256thr: 1KB x (32+6) rows, PF4");
    time([&](const float* s) { tile_read<8><<<32 * 128, 256>>>(s, out, n, 32, 0, 256, 32); }, "tile 256thr: 1KB x 32 rows no halo, PF8");
    time([&](const float* s) { tile_read<8><<<32 * 32, 256>>>(s, out, n, 128, 0, 256, 32); }, "tile 256thr: 1KB x 128 rows no halo, PF8");
    time([&](const float* s) { tile_read<16><<<35 * 128, 256>>>(s, out, n, 32, 6, 240, 35); }, "tile 256thr: 1KB x (32+6) rows, PF16");
    return 0;
}
