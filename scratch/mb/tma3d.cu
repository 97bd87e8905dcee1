// minimal check of a 3-D tensor-map TMA box load (the staging step of the tiled affine kernel)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, float *out, int PB, int HB, int x0, int y0, int b)
{
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    float *box = (float *)raw;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(PB * HB * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                         "r"(s32(box)), "l"(&tmap), "r"(x0), "r"(y0), "r"(b), "r"(s32(&bar)) : "memory");
    }
    asm volatile("{\n .reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D;\n bra W;\nD:\n}" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < PB * HB; i += blockDim.x) out[i] = box[i];
}
int main(int argc, char **argv)
{
    const int x0 = argc > 1 ? atoi(argv[1]) : 37, y0 = argc > 2 ? atoi(argv[2]) : -2, promo = argc > 3 ? atoi(argv[3]) : 2, PBa = argc > 4 ? atoi(argv[4]) : 64;
    const int m1 = 512, m2 = 512, nb = 2, PB = PBa, HB = 43;
    std::vector<float> h((size_t)m1 * m2 * nb);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, PB * HB * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %d q=%d p=%p\n", (int)e, (int)q, p);
    CUtensorMap tm; memset(&tm, 0, sizeof tm);
    cuuint64_t gdim[3] = {m1, m2, nb}; cuuint64_t gstr[2] = {m1 * 4ull, (cuuint64_t)m1 * m2 * 4};
    cuuint32_t box[3] = {PB, HB, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    k<<<1, 128, PB * HB * 4>>>(tm, o, PB, HB, x0, y0, 1);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> ho(PB * HB);
    cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < HB; ++r2) for (int c = 0; c < PB; ++c) {
        int y = y0 + r2, x = x0 + c;
        float want = (y < 0 || y >= m2 || x >= m1 || x < 0) ? 0.f : h[(size_t)1 * m1 * m2 + (size_t)y * m1 + x];
        bad += ho[r2 * PB + c] != want;
    }
    printf("mismatches: %d\n", bad);
    return 0;
}
