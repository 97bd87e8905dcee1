// li_tab_apply.cu -- table-backed operators along one dimension (K3-K6, generic form).
//
// Direct  : dst[p,i,q] = alpha * sum_s W[s,i] * src[p, J[s,i], q] + beta * dst[p,i,q]
//           reference __direct! (src/tabulated.jl:394-439), SparseInterpolator Direct
//           (src/sparse.jl:201-244), _apply_direct! (src/sparse.jl:719-774)
// Adjoint : dst[p,j,q] = beta * dst[p,j,q] + sum_{(i,s): J[s,i]=j} W[s,i] * (alpha * src[p,i,q])
//           reference __adjoint! (src/tabulated.jl:441-478), src/sparse.jl:246-275, :776-800
//
// These kernels work for ANY table (arbitrary J, e.g. SparseInterpolator with N-d positions):
// they gather through L1/L2.  Windowed tables of separable operators take the fused, staged
// kernels of li_sep2d.cu instead.  Threads always run along the contiguous dimension (the
// opposite of the reference's CPU loop nest, which puts the interpolated index innermost).
#include "li_common.cuh"

namespace li {

template <typename T>
__device__ __forceinline__ T fma_(T a, T b, T c);
template <>
__device__ __forceinline__ float fma_<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <>
__device__ __forceinline__ double fma_<double>(double a, double b, double c) { return fma(a, b, c); }

template <typename T>
__device__ __forceinline__ T finish(T a, T alpha, T beta, const T *dst)
{
    // beta == 0: dst is never read (src/tabulated.jl:408-425)
    if (beta == (T)0) return alpha == (T)1 ? a : alpha * a;
    return fma_(alpha, a, beta * (*dst));
}

// ---- Direct, interpolated dimension is the contiguous one (npre == 1) ------
// thread <-> output row i; J/W of the row are loaded once and reused for QB planes q.
template <typename T, int SC>
__global__ void __launch_bounds__(256)
tab_direct_d1_kernel(const int32_t *__restrict__ J, const T *__restrict__ W, int Srt, int nrows, int ncols,
                     int64_t npost, T alpha, T beta, const T *__restrict__ src, T *__restrict__ dst)
{
    const int S = SC ? SC : Srt;
    constexpr int SMAX = SC ? SC : LI_MAX_S;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    int j[SMAX];
    T w[SMAX];
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
        if (s < S) {
            j[s] = J[(int64_t)i * S + s];
            w[s] = W[(int64_t)i * S + s];
        }
    }
    for (int64_t q = blockIdx.y; q < npost; q += gridDim.y) {
        const T *sp = src + q * ncols;
        T a = (T)0;
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
            if (s < S) a = fma_(sp[j[s]], w[s], a);
        T *d = dst + q * nrows + i;
        *d = finish(a, alpha, beta, d);
    }
}

// ---- Direct, strided dimension (npre > 1) -----------------------------------
// thread <-> p (contiguous); a block handles RB consecutive rows i; J/W are warp-uniform.
template <typename T, int SC, int RB>
__global__ void __launch_bounds__(256)
tab_direct_dn_kernel(const int32_t *__restrict__ J, const T *__restrict__ W, int Srt, int nrows, int ncols,
                     int64_t npre, int64_t npost, T alpha, T beta, const T *__restrict__ src,
                     T *__restrict__ dst)
{
    const int S = SC ? SC : Srt;
    constexpr int SMAX = SC ? SC : LI_MAX_S;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npre) return;
    for (int i0 = blockIdx.y * RB; i0 < nrows; i0 += gridDim.y * RB)
    for (int64_t q = blockIdx.z; q < npost; q += gridDim.z) {
        const T *sp = src + p + q * npre * ncols;
        T *dp = dst + p + q * npre * nrows;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const int i = i0 + r;
            if (i < nrows) {
                T a = (T)0;
#pragma unroll
                for (int s = 0; s < SMAX; ++s)
                    if (s < S) a = fma_(sp[(int64_t)J[(int64_t)i * S + s] * npre], W[(int64_t)i * S + s], a);
                T *d = dp + (int64_t)i * npre;
                *d = finish(a, alpha, beta, d);
            }
        }
    }
}

// ---- Adjoint, deterministic gather over the transposed (CSC) table ------------
// Entries of a column are visited in ascending (row, tap) order == the order in which the
// reference's sequential scatter adds them, so the rounding sequence is the same.
template <typename T>
__global__ void __launch_bounds__(256)
tab_adjoint_gather_d1_kernel(const int32_t *__restrict__ colptr, const int32_t *__restrict__ tent,
                             const T *__restrict__ W, int S, int nrows, int ncols, int64_t npost, T alpha,
                             T beta, const T *__restrict__ src, T *__restrict__ dst)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const int e0 = colptr[j], e1 = colptr[j + 1];
    for (int64_t q = blockIdx.y; q < npost; q += gridDim.y) {
        const T *sp = src + q * nrows;
        T *d = dst + q * ncols + j;
        T acc = beta == (T)0 ? (T)0 : beta * (*d);
        for (int e = e0; e < e1; ++e) {
            const int ent = tent[e];
            const T x = alpha == (T)1 ? sp[ent / S] : alpha * sp[ent / S];
            acc = fma_(W[ent], x, acc);
        }
        *d = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
tab_adjoint_gather_dn_kernel(const int32_t *__restrict__ colptr, const int32_t *__restrict__ tent,
                             const T *__restrict__ W, int S, int nrows, int ncols, int64_t npre, int64_t npost,
                             T alpha, T beta, const T *__restrict__ src, T *__restrict__ dst)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npre) return;
    for (int j = blockIdx.y; j < ncols; j += gridDim.y) {
    const int e0 = colptr[j], e1 = colptr[j + 1];
    for (int64_t q = blockIdx.z; q < npost; q += gridDim.z) {
        const T *sp = src + p + q * npre * nrows;
        T *d = dst + p + ((int64_t)j + q * ncols) * npre;
        T acc = beta == (T)0 ? (T)0 : beta * (*d);
        for (int e = e0; e < e1; ++e) {
            const int ent = tent[e];
            const T v = sp[(int64_t)(ent / S) * npre];
            const T x = alpha == (T)1 ? v : alpha * v;
            acc = fma_(W[ent], x, acc);
        }
        *d = acc;
    }
    }
}

// ---- Adjoint, atomic scatter (benchmarked variant; not run-to-run deterministic) ----
template <typename T>
__global__ void __launch_bounds__(256)
tab_adjoint_atomic_kernel(const int32_t *__restrict__ J, const T *__restrict__ W, int S, int nrows, int ncols,
                          int64_t npre, int64_t npost, T alpha, const T *__restrict__ src, T *__restrict__ dst)
{
    const int64_t total = npre * nrows * npost;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t p = idx % npre;
        const int64_t r = idx / npre;
        const int i = (int)(r % nrows);
        const int64_t q = r / nrows;
        const T x = alpha == (T)1 ? src[idx] : alpha * src[idx];
        if (x == (T)0) continue;  // src/tabulated.jl:457
        T *dp = dst + p + q * npre * ncols;
        for (int s = 0; s < S; ++s) atomicAdd(dp + (int64_t)J[(int64_t)i * S + s] * npre, W[(int64_t)i * S + s] * x);
    }
}

// ---------------------------------------------------------------------------
template <typename T, int SC>
static int launch_direct(const li_table *t, int64_t npre, int64_t npost, T alpha, T beta, const T *src, T *dst,
                         cudaStream_t st)
{
    const int nrows = (int)t->nrows, ncols = (int)t->ncols;
    if (npre == 1) {
        dim3 grid((nrows + 255) / 256, (unsigned)std::min<int64_t>(npost, 65535));
        tab_direct_d1_kernel<T, SC><<<grid, 256, 0, st>>>(t->J, (const T *)t->W, t->S, nrows, ncols, npost, alpha,
                                                         beta, src, dst);
    } else {
        constexpr int RB = 4;
        const int threads = npre >= 256 ? 256 : (npre >= 128 ? 128 : (npre >= 64 ? 64 : 32));
        dim3 grid((unsigned)((npre + threads - 1) / threads), std::min((nrows + RB - 1) / RB, 65535),
                  (unsigned)std::min<int64_t>(npost, 65535));
        tab_direct_dn_kernel<T, SC, RB><<<grid, threads, 0, st>>>(t->J, (const T *)t->W, t->S, nrows, ncols, npre,
                                                                npost, alpha, beta, src, dst);
    }
    LI_CHECK_LAUNCH();
    return LI_OK;
}

template <typename T>
static int direct_T(const li_table *t, int64_t npre, int64_t npost, double alpha, double beta, const void *src,
                    void *dst, cudaStream_t st)
{
    const T a = (T)alpha, b = (T)beta;
    switch (t->S) {
    case 1: return launch_direct<T, 1>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 2: return launch_direct<T, 2>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 3: return launch_direct<T, 3>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 4: return launch_direct<T, 4>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 6: return launch_direct<T, 6>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 8: return launch_direct<T, 8>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    default: return launch_direct<T, 0>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    }
}

template <typename T>
static int adjoint_T(const li_table *t, int64_t npre, int64_t npost, double alpha, double beta, const void *src,
                     void *dst, int variant, cudaStream_t st)
{
    const int nrows = (int)t->nrows, ncols = (int)t->ncols;
    const T a = (T)alpha, b = (T)beta;
    if (variant == LI_VARIANT_ATOMIC) {
        int rc = vscale_impl(t->dtype, dst, npre * ncols * npost, beta, st);
        if (rc != LI_OK) return rc;
        const int64_t total = npre * nrows * npost;
        const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 32);
        tab_adjoint_atomic_kernel<T><<<blocks, 256, 0, st>>>(t->J, (const T *)t->W, t->S, nrows, ncols, npre, npost,
                                                            a, (const T *)src, (T *)dst);
        LI_CHECK_LAUNCH();
        return LI_OK;
    }
    int rc = table_ensure_csc(t, st);
    if (rc != LI_OK) return rc;
    if (npre == 1) {
        dim3 grid((ncols + 255) / 256, (unsigned)std::min<int64_t>(npost, 65535));
        tab_adjoint_gather_d1_kernel<T><<<grid, 256, 0, st>>>(t->colptr, t->tent, (const T *)t->W, t->S, nrows,
                                                             ncols, npost, a, b, (const T *)src, (T *)dst);
    } else {
        const int threads = npre >= 256 ? 256 : (npre >= 128 ? 128 : (npre >= 64 ? 64 : 32));
        dim3 grid((unsigned)((npre + threads - 1) / threads), std::min(ncols, 65535),
                  (unsigned)std::min<int64_t>(npost, 65535));
        tab_adjoint_gather_dn_kernel<T><<<grid, threads, 0, st>>>(t->colptr, t->tent, (const T *)t->W, t->S, nrows,
                                                                 ncols, npre, npost, a, b, (const T *)src, (T *)dst);
    }
    LI_CHECK_LAUNCH();
    return LI_OK;
}

int tab_apply_impl(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha, const void *src,
                   double beta, void *dst, int variant, cudaStream_t st)
{
    LI_REQUIRE(t && dst, LI_ERR_ARG, "null argument");
    LI_REQUIRE(npre >= 0 && npost >= 0, LI_ERR_DIM, "negative dimension");
    const int64_t nout = npre * npost * (adjoint ? t->ncols : t->nrows);
    if (nout == 0) return LI_OK;
    if (alpha == 0.0) return vscale_impl(t->dtype, dst, nout, beta, st);  // src/tabulated.jl:352-353
    LI_REQUIRE(src, LI_ERR_ARG, "null source");
    LI_REQUIRE(npre * npost * std::max(t->ncols, t->nrows) < (1LL << 62), LI_ERR_DIM, "array too large");
    if (!adjoint)
        return t->dtype == LI_F32 ? direct_T<float>(t, npre, npost, alpha, beta, src, dst, st)
                                  : direct_T<double>(t, npre, npost, alpha, beta, src, dst, st);
    return t->dtype == LI_F32 ? adjoint_T<float>(t, npre, npost, alpha, beta, src, dst, variant, st)
                              : adjoint_T<double>(t, npre, npost, alpha, beta, src, dst, variant, st);
}

}  // namespace li

extern "C" {

int li_tab_apply(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha, const void *src_dev,
                 double beta, void *dst_dev, int variant, void *stream)
{
    return li::tab_apply_impl(adjoint, t, npre, npost, alpha, src_dev, beta, dst_dev, variant, li::as_stream(stream));
}

}  // extern "C"

static int tab_apply_host_impl(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                               const void *src_host, double beta, void *dst_host, int variant, void *stream, bool sync)
{
    LI_REQUIRE(t && src_host && dst_host, LI_ERR_ARG, "null argument");
    cudaStream_t st = li::as_stream(stream);
    const size_t es = li::dtype_size(t->dtype);
    const size_t nin = (size_t)(npre * npost * (adjoint ? t->nrows : t->ncols)) * es;
    const size_t nout = (size_t)(npre * npost * (adjoint ? t->ncols : t->nrows)) * es;
    void *dsrc = nullptr, *ddst = nullptr;
    int rc = li::staging(0, nin ? nin : 1, &dsrc, stream);
    if (rc == LI_OK) rc = li::staging(1, nout ? nout : 1, &ddst, stream);
    if (rc != LI_OK) return rc;
    cudaError_t e = cudaMemcpyAsync(dsrc, src_host, nin, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && beta != 0.0) e = cudaMemcpyAsync(ddst, dst_host, nout, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = li::tab_apply_impl(adjoint, t, npre, npost, alpha, dsrc, beta, ddst, variant, st);
    if (e == cudaSuccess && rc == LI_OK) e = cudaMemcpyAsync(dst_host, ddst, nout, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        li::set_error("host-staged apply failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    return rc;
}

extern "C" {

int li_tab_apply_host(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                      const void *src_host, double beta, void *dst_host, int variant, void *stream)
{
    return tab_apply_host_impl(adjoint, t, npre, npost, alpha, src_host, beta, dst_host, variant, stream, true);
}

int li_tab_apply_host_async(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                            const void *src_host, double beta, void *dst_host, int variant, void *stream)
{
    return tab_apply_host_impl(adjoint, t, npre, npost, alpha, src_host, beta, dst_host, variant, stream, false);
}

}  // extern "C"
