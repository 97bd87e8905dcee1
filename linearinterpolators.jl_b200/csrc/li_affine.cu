// li_affine.cu -- 2-D interpolation through an affine coordinate transform (K9/K10).
//
// Reference: apply!(alpha, P, ker1, ker2, R, src, beta, dst), src/nonseparable.jl:147-264.
// For output pixel (i1, i2) (1-based):
//     x = R.xx*T(i1) + (R.xy*T(i2) + R.x),  y = R.yx*T(i1) + (R.yy*T(i2) + R.y)
//     (j1, w1) = getcoefs(ker1, lim1, x),   (j2, w2) = getcoefs(ker2, lim2, y)
//     Direct : dst[i1,i2] = sum_{s2} (sum_{s1} src[j1,j2]*w1)*w2           (:185,195)
//     Adjoint: dst[j1,j2] += w1 * (w2 * (alpha*src[i1,i2]))                  (:225-232)
// Coordinates and coefficients are evaluated with un-fused IEEE operations so the indices and
// weights are the reference's bit for bit; they are computed once per pixel and reused for all
// images of the batch (the reference loops over images outside the operator).
#include "li_common.cuh"
#include "li_kerdev.cuh"
#include "li_weights.cuh"

namespace li {

template <typename T>
struct AffineArgs {
    KerDev<T> k1, k2;
    T xx, xy, x, yx, yy, y;
    int n1, n2, m1, m2;
    int64_t nbatch;
    T alpha, beta;
};

template <typename T, int S1, int S2>
__device__ __forceinline__ bool affine_coefs(const AffineArgs<T> &a, int i1, int i2, int &b1, int &b2, T (&w1)[S1],
                                             T (&w2)[S2])
{
    using A = Ar<T>;
    const T pos1 = (T)(i1 + 1), pos2 = (T)(i2 + 1);
    const T off1 = A::add(A::mul(a.xy, pos2), a.x);
    const T off2 = A::add(A::mul(a.yy, pos2), a.y);
    const T px = A::add(A::mul(a.xx, pos1), off1);
    const T py = A::add(A::mul(a.yx, pos1), off2);
    bool ok = getcoefs<T, S1>(a.k1, a.m1, px, b1, w1);
    ok = getcoefs<T, S2>(a.k2, a.m2, py, b2, w2) && ok;
    return ok;
}

__device__ __forceinline__ int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// Direct: one thread per output pixel, looping over the images of its batch slice.
template <typename T, int S1, int S2>
__global__ void __launch_bounds__(256)
affine_direct_kernel(AffineArgs<T> a, const T *__restrict__ src, T *__restrict__ dst, int *__restrict__ err)
{
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i2 = blockIdx.y * blockDim.y + threadIdx.y;
    if (i1 >= a.n1 || i2 >= a.n2) return;
    T w1[S1], w2[S2];
    int b1, b2;
    if (!affine_coefs<T, S1, S2>(a, i1, i2, b1, b2, w1, w2)) {
        atomicExch(err, 1);
        return;
    }
    int j1[S1];
    int64_t o2[S2];
#pragma unroll
    for (int s = 0; s < S1; ++s) j1[s] = clampi(b1 + s, a.m1 - 1);
#pragma unroll
    for (int s = 0; s < S2; ++s) o2[s] = (int64_t)clampi(b2 + s, a.m2 - 1) * a.m1;
    const int64_t splane = (int64_t)a.m1 * a.m2, dplane = (int64_t)a.n1 * a.n2;
    const bool plain = a.alpha == (T)1 && a.beta == (T)0;
    for (int64_t b = blockIdx.z; b < a.nbatch; b += gridDim.z) {
        const T *sp = src + b * splane;
        T e = (T)0;
#pragma unroll
        for (int s2 = 0; s2 < S2; ++s2) {
            const T *col = sp + o2[s2];
            T sub = col[j1[0]] * w1[0];
#pragma unroll
            for (int s1 = 1; s1 < S1; ++s1) sub += col[j1[s1]] * w1[s1];
            e = s2 == 0 ? sub * w2[0] : e + sub * w2[s2];
        }
        T *d = dst + b * dplane + (int64_t)i2 * a.n1 + i1;
        // :185 / :195 -- the general branch reads dst even when beta == 0, like the reference
        *d = plain ? e : e * a.alpha + (*d) * a.beta;
    }
}

// Adjoint, atomic scatter variant (dst pre-scaled by beta).
template <typename T, int S1, int S2>
__global__ void __launch_bounds__(256)
affine_adjoint_atomic_kernel(AffineArgs<T> a, const T *__restrict__ src, T *__restrict__ dst, int *__restrict__ err)
{
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i2 = blockIdx.y * blockDim.y + threadIdx.y;
    if (i1 >= a.n1 || i2 >= a.n2) return;
    T w1[S1], w2[S2];
    int b1, b2;
    if (!affine_coefs<T, S1, S2>(a, i1, i2, b1, b2, w1, w2)) {
        atomicExch(err, 1);
        return;
    }
    int j1[S1];
    int64_t o2[S2];
#pragma unroll
    for (int s = 0; s < S1; ++s) j1[s] = clampi(b1 + s, a.m1 - 1);
#pragma unroll
    for (int s = 0; s < S2; ++s) o2[s] = (int64_t)clampi(b2 + s, a.m2 - 1) * a.m1;
    const int64_t splane = (int64_t)a.n1 * a.n2, dplane = (int64_t)a.m1 * a.m2;
    for (int64_t b = blockIdx.z; b < a.nbatch; b += gridDim.z) {
        T val = src[b * splane + (int64_t)i2 * a.n1 + i1];
        if (a.alpha != (T)1) val = a.alpha * val;
        T *dp = dst + b * dplane;
#pragma unroll
        for (int s2 = 0; s2 < S2; ++s2) {
            const T tmp = w2[s2] * val;
#pragma unroll
            for (int s1 = 0; s1 < S1; ++s1) atomicAdd(dp + o2[s2] + j1[s1], w1[s1] * tmp);
        }
    }
}

template <typename T, int S1, int S2>
static int launch_affine(int adjoint, const AffineArgs<T> &a, const T *src, T *dst, int *err, cudaStream_t st)
{
    dim3 block(64, 4);
    const int64_t pix_blocks = (int64_t)((a.n1 + 63) / 64) * ((a.n2 + 3) / 4);
    // enough z-slices to fill the machine for small images, few for large ones (weights are reused per slice)
    int64_t z = std::max<int64_t>(1, std::min<int64_t>(a.nbatch, (148 * 16 + pix_blocks - 1) / pix_blocks));
    dim3 grid((a.n1 + 63) / 64, (a.n2 + 3) / 4, (unsigned)std::min<int64_t>(z, 65535));
    LI_REQUIRE(grid.y <= 65535, LI_ERR_UNSUPPORTED, "image too tall for the affine kernel");
    if (adjoint)
        affine_adjoint_atomic_kernel<T, S1, S2><<<grid, block, 0, st>>>(a, src, dst, err);
    else
        affine_direct_kernel<T, S1, S2><<<grid, block, 0, st>>>(a, src, dst, err);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

template <typename T>
static int affine_T(int adjoint, const li_kernel_desc *ker1, const li_kernel_desc *ker2, const double R[6], int64_t n1,
                    int64_t n2, int64_t m1, int64_t m2, int64_t nbatch, double alpha, const void *src, double beta,
                    void *dst, int *err, cudaStream_t st)
{
    AffineArgs<T> a;
    int rc = make_kerdev<T>(ker1, m1, &a.k1);
    if (rc == LI_OK) rc = make_kerdev<T>(ker2, m2, &a.k2);
    if (rc != LI_OK) return rc;
    // convert(AffineTransform2D{T}, R): every coefficient rounded to T (src/nonseparable.jl:58)
    a.xx = (T)R[0], a.xy = (T)R[1], a.x = (T)R[2], a.yx = (T)R[3], a.yy = (T)R[4], a.y = (T)R[5];
    a.n1 = (int)n1, a.n2 = (int)n2, a.m1 = (int)m1, a.m2 = (int)m2;
    a.nbatch = nbatch;
    a.alpha = (T)alpha, a.beta = (T)beta;
    const int S1 = ker1->S, S2 = ker2->S;
#define LI_AFF(A1, A2)              \
    if (S1 == A1 && S2 == A2) return launch_affine<T, A1, A2>(adjoint, a, (const T *)src, (T *)dst, err, st);
    LI_AFF(1, 1) LI_AFF(1, 2) LI_AFF(1, 3) LI_AFF(1, 4)
    LI_AFF(2, 1) LI_AFF(2, 2) LI_AFF(2, 3) LI_AFF(2, 4)
    LI_AFF(3, 1) LI_AFF(3, 2) LI_AFF(3, 3) LI_AFF(3, 4)
    LI_AFF(4, 1) LI_AFF(4, 2) LI_AFF(4, 3) LI_AFF(4, 4)
    LI_AFF(6, 6) LI_AFF(8, 8)
#undef LI_AFF
    set_error("affine operator: kernel sizes (%d, %d) are not built", S1, S2);
    return LI_ERR_UNSUPPORTED;
}

}  // namespace li

extern "C" int li_affine2d_apply(int adjoint, const li_kernel_desc *ker1, const li_kernel_desc *ker2,
                                 const double R[6], int64_t n1, int64_t n2, int64_t m1, int64_t m2, int64_t nbatch,
                                 double alpha, const void *src_dev, double beta, void *dst_dev, int variant, int flags,
                                 void *stream)
{
    (void)variant;  // the adjoint currently has the atomic form only
    LI_REQUIRE(ker1 && ker2 && R && dst_dev, LI_ERR_ARG, "null argument");
    LI_REQUIRE(ker1->dtype == ker2->dtype, LI_ERR_ARG, "kernels must share one floating-point type");
    LI_REQUIRE(n1 >= 0 && n2 >= 0 && m1 >= 1 && m2 >= 1 && nbatch >= 0, LI_ERR_DIM, "bad dimensions");
    LI_REQUIRE(n1 < (1LL << 31) && n2 < (1LL << 31) && m1 < (1LL << 31) && m2 < (1LL << 31), LI_ERR_UNSUPPORTED,
               "dimension too large");
    cudaStream_t st = li::as_stream(stream);
    const int dtype = ker1->dtype;
    const int64_t nout = nbatch * (adjoint ? m1 * m2 : n1 * n2);
    if (nout == 0) return LI_OK;
    if (alpha == 0.0) return li::vscale_impl(dtype, dst_dev, nout, beta, st);  // :169-172
    LI_REQUIRE(src_dev, LI_ERR_ARG, "null source");
    if (n1 == 0 || n2 == 0) return li::vscale_impl(dtype, dst_dev, nout, beta, st);
    if (adjoint) {
        int rc = li::vscale_impl(dtype, dst_dev, nout, beta, st);  // :236
        if (rc != LI_OK) return rc;
    }
    static thread_local int *d_err = nullptr;
    if (!d_err) LI_CUDA(cudaMalloc(&d_err, sizeof(int)));
    LI_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    int rc = dtype == LI_F32 ? li::affine_T<float>(adjoint, ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev,
                                                   beta, dst_dev, d_err, st)
                             : li::affine_T<double>(adjoint, ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev,
                                                    beta, dst_dev, d_err, st);
    if (rc != LI_OK || (flags & LI_FLAG_ASYNC)) return rc;
    int h_err = 0;
    LI_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    LI_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
        li::set_error("InexactError: trunc(Int64, f) for a NaN or out-of-range coordinate");
        return LI_ERR_INEXACT;
    }
    return LI_OK;
}
