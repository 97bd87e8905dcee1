// li_sep2d.cuh -- types and device helpers shared by the fused 2-D separable kernels (li_sep2d.cu, li_sep2d_adjs.cu).
#pragma once
#include <algorithm>
#include <vector>

#include "li_common.cuh"

namespace li {


// ---------------------------------------------------------------------------------------------
template <typename T, int V>
struct VecT;
template <>
struct VecT<float, 4> { using type = float4; };
template <>
struct VecT<float, 2> { using type = float2; };
template <>
struct VecT<float, 1> { using type = float; };
template <>
struct VecT<double, 2> { using type = double2; };
struct __align__(16) double4a { double x, y, z, w; };   // two 16-byte transactions
template <>
struct VecT<double, 4> { using type = double4a; };
template <>
struct VecT<double, 1> { using type = double; };

template <typename T, int V>
struct Pack {
    T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Pack<T, V> ld_pack(const T *p)
{
    using VT = typename VecT<T, V>::type;
    Pack<T, V> r;
    *reinterpret_cast<VT *>(r.v) = *reinterpret_cast<const VT *>(p);
    return r;
}

template <typename T, int V>
__device__ __forceinline__ void st_pack(T *p, const Pack<T, V> &r)
{
    using VT = typename VecT<T, V>::type;
    *reinterpret_cast<VT *>(p) = *reinterpret_cast<const VT *>(r.v);
}

template <typename T>
__device__ __forceinline__ T fmaT(T a, T b, T c);
template <>
__device__ __forceinline__ float fmaT<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <>
__device__ __forceinline__ double fmaT<double>(double a, double b, double c) { return fma(a, b, c); }

// packed pairs: FP32x2 instructions for float (sm_100), two scalar chains for double
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ double2 fma2(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ double2 mul2(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 add2(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

template <typename T>
struct SepArgs {
    const int32_t *base1, *base2;  // window bases (canonical, 0-based)
    const T *W1, *W2;              // [n*S]
    const int32_t *bp1, *bp2;      // adjoint only: bptr tables (see li_table.cu), [m+S]; bp2 is over FOLDED bases
    const T *W2f;                  // adjoint only: folded dim-2 weights (li_table.cu: fold_kernel)
    const int32_t *rowlist;        // adjoint only: tile rows handled by this launch (null: all, in order)
    const int32_t *bfold1;         // streaming adjoint: folded dim-1 bases
    const int32_t *irrp2;          // streaming adjoint: irregularity prefix sums of the dim-2 table (li_table.cu)
    const int32_t *cstart;         // streaming adjoint: first dst row of every chunk, [tiles2 + 1] (cost-balanced)
    const T *W1f;                  // streaming adjoint: folded dim-1 weights
    int n1, n2;                    // "rows" side (fine grid for an upsampling operator)
    int m1, m2;                    // "columns" side
    int tiles1, tiles2;
    int chunk0;                    // first chunk (along dim 2) of this launch: bands of the host-buffer pipeline
    int F1MAX, F2MAX;              // forward: footprint bounds; adjoint: F2MAX = max src rows feeding one row group
    int vec_ok;                    // forward: bit 0 = source rows 16-byte aligned (TMA), bit 1 = destination rows vector
                                   // aligned; tile adjoint: rows of the big array are 2-element-vector aligned
    T alpha, beta;
    const T *src;
    T *dst;
    // three dimensions (li_sep3d.cuh): the table along dim 3; n3 rows (fine side), m3 columns
    const int32_t *base3;          // window bases
    const T *W3;                   // forward: [n3*S3] weights
    const int32_t *bp3, *bfold3;   // adjoint: bptr table (rows feeding a column), folded bases
    const T *W3f;                  // adjoint: folded weights
    int n3, m3, S3;
    int nc3;                       // planes combined into one staged row (forward: S3; adjoint: most rows feeding a plane)
};


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes: multiple of 16, both sides 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T, int S>
__device__ __forceinline__ void ld_weights(const T *p, T (&w)[S])
{
    // rows of S weights are 16-byte aligned when S*sizeof(T) is a multiple of 16
    if constexpr ((S * sizeof(T)) % 16 == 0) {
        constexpr int V = 16 / sizeof(T);
#pragma unroll
        for (int k = 0; k < S / V; ++k) {
            const Pack<T, V> t = ld_pack<T, V>(p + k * V);
#pragma unroll
            for (int v = 0; v < V; ++v) w[k * V + v] = t.v[v];
        }
    } else {
#pragma unroll
        for (int s = 0; s < S; ++s) w[s] = p[s];
    }
}

static const int kSmemLimit = 200 * 1024;

template <typename KernT>
static int set_smem(KernT kern, size_t smem)
{
    if (smem > 48 * 1024) LI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return LI_OK;
}

// Banding of the fused kernels along dim 2 for the host-buffer entry points (sep2d_host_banded): the launchers consult
// this thread-local control block.  mode 1 = describe the chunks of the launch without launching (plan), mode 2 =
// launch only chunks [chunk0, chunk0 + nchunks).
struct BandCtl {
    int mode = 0;
    int chunk0 = 0, nchunks = 0;
    // filled by the plan: chunk k produces dst rows [row_begin[k], row_begin[k+1]) and reads src rows < src_end[k]
    std::vector<int> row_begin, src_end;
};
BandCtl *&sep2d_band();

// streaming adjoint (li_sep2d_adjs.cu): *done stays false when the kernel declines the problem
template <typename T>
int sep2d_adj_stream(SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st, bool *done);
int sep2d_num_sms();

}  // namespace li
