// li_affine.cu -- C entry point of the 2-D affine-transform operator (K9/K10); kernels: li_affine.cuh.
//
// Reference: apply!(alpha, P, ker1, ker2, R, src, beta, dst), src/nonseparable.jl:147-264.
#include <cstdlib>

#include "li_common.cuh"

namespace li {
template <typename T>
int affine_T(int adjoint, const li_kernel_desc *ker1, const li_kernel_desc *ker2, const double R[6], int64_t n1,
             int64_t n2, int64_t m1, int64_t m2, int64_t nbatch, double alpha, const void *src, double beta,
             void *dst, int *err, bool tiled, cudaStream_t st);
extern template int affine_T<float>(int, const li_kernel_desc *, const li_kernel_desc *, const double[6], int64_t, int64_t, int64_t,
                                    int64_t, int64_t, double, const void *, double, void *, int *, bool, cudaStream_t);
extern template int affine_T<double>(int, const li_kernel_desc *, const li_kernel_desc *, const double[6], int64_t, int64_t, int64_t,
                                     int64_t, int64_t, double, const void *, double, void *, int *, bool, cudaStream_t);
// gather-form adjoint (li_affine_gather.cuh): *done stays false when it declines the operator
template <typename T>
int affine_adjoint_gather(const li_kernel_desc *ker1, const li_kernel_desc *ker2, const double R[6], int64_t n1, int64_t n2, int64_t m1,
                          int64_t m2, int64_t nbatch, double alpha, const void *src, double beta, void *dst, int *err, cudaStream_t st,
                          bool *done);
extern template int affine_adjoint_gather<float>(const li_kernel_desc *, const li_kernel_desc *, const double[6], int64_t, int64_t, int64_t,
                                                 int64_t, int64_t, double, const void *, double, void *, int *, cudaStream_t, bool *);
extern template int affine_adjoint_gather<double>(const li_kernel_desc *, const li_kernel_desc *, const double[6], int64_t, int64_t, int64_t,
                                                  int64_t, int64_t, double, const void *, double, void *, int *, cudaStream_t, bool *);

// positions of an axis-aligned transform: x[i] = c * convert(T, i + 1) + off, product and sum rounded separately as in
// the reference (R.xx*pos1 + off1, src/nonseparable.jl:163-168)
template <typename T>
__global__ void axis_positions_kernel(T c, T off, int64_t n, T *__restrict__ x)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if constexpr (sizeof(T) == 4)
        x[i] = __fadd_rn(__fmul_rn(c, (T)(i + 1)), off);
    else
        x[i] = __dadd_rn(__dmul_rn(c, (T)(i + 1)), off);
}

// Aligned axes (R.xy == 0 and R.yx == 0: scaling + translation): the reference's FIXME "if axes are aligned, use
// separable interpolation" (src/nonseparable.jl:14).  The positions R.xx*i1 + (R.xy*i2 + R.x) do not depend on i2 (0*i2
// is exactly 0), R.yx*i1 + (R.yy*i2 + R.y) not on i1, so the coefficients are those of two per-dimension tables (bit
// for bit) and the operator is the separable one: HBM-bound fused kernels instead of the L1-bound tiled gather.
template <typename T>
static int affine_aligned(int adjoint, const li_kernel_desc *ker1, const li_kernel_desc *ker2, const double R[6], int64_t n1, int64_t n2,
                          int64_t m1, int64_t m2, int64_t nbatch, double alpha, const void *src, double beta, void *dst,
                          cudaStream_t st)
{
    void *x1 = nullptr;
    int rc = staging(8, (size_t)(n1 + n2) * sizeof(T), &x1, (void *)st);
    if (rc != LI_OK) return rc;
    T *px1 = (T *)x1, *px2 = px1 + n1;
    const T xy = (T)R[1], yx = (T)R[3];
    // off1 = R.xy*pos2 + R.x with R.xy == 0, off2 = R.yy*pos2 + R.y is the dim-2 position itself
    const T off1 = (T)(xy * (T)1) + (T)R[2];
    axis_positions_kernel<T><<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>((T)R[0], off1, n1, px1);
    axis_positions_kernel<T><<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>((T)R[4], (T)R[5], n2, px2);
    count_launch(2);
    LI_CUDA(cudaGetLastError());
    (void)yx;
    const li_kernel_desc *kers[2] = {ker1, ker2};
    const void *xs[2] = {px1, px2};
    const int64_t n[2] = {n1, n2}, m[2] = {m1, m2};
    return li_sep_otf_apply(adjoint, 2, kers, xs, n, m, nbatch, alpha, src, beta, dst, (void *)st);
}
}  // namespace li

extern "C" int li_affine2d_apply(int adjoint, const li_kernel_desc *ker1, const li_kernel_desc *ker2,
                                 const double R[6], int64_t n1, int64_t n2, int64_t m1, int64_t m2, int64_t nbatch,
                                 double alpha, const void *src_dev, double beta, void *dst_dev, int variant, int flags,
                                 void *stream)
{
    // direct: LI_VARIANT_AUTO = tiled kernel (source boxes staged in shared memory), LI_VARIANT_ATOMIC = one thread
    // per pixel straight on global memory (the first implementation, kept for comparison).  Adjoint: LI_VARIANT_AUTO /
    // LI_VARIANT_GATHER = the deterministic gather form (li_affine_gather.cuh) when it takes the operator, else --
    // and with LI_VARIANT_ATOMIC -- the atomic scatter
    const bool tiled = !adjoint && variant != LI_VARIANT_ATOMIC && !li::opt(li::OPT_AFFINE_PLAIN);
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
    if (R[1] == 0.0 && R[3] == 0.0 && variant != LI_VARIANT_ATOMIC && !li::opt(li::OPT_AFFINE_PLAIN) && !(li::opt(li::OPT_EXP) & 4)) {
        // (x2 = R.yx*pos1 + off2 with R.yx == 0: 0*pos1 + off2 == off2 unless off2 is -0, which getcoefs treats alike)
        if (li::opt(li::OPT_TRACE)) fprintf(stderr, "[li_b200] affine: aligned axes -> separable on-the-fly operator\n");
        return dtype == LI_F32 ? li::affine_aligned<float>(adjoint, ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev, beta, dst_dev, st)
                               : li::affine_aligned<double>(adjoint, ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev, beta, dst_dev, st);
    }
    static thread_local std::map<int, int *> d_errs;   // per device
    int *&d_err = d_errs[li::current_device()];
    if (!d_err) LI_CUDA(cudaMalloc(&d_err, sizeof(int)));
    LI_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    int rc = LI_OK;
    bool gathered = false;
    if (adjoint && variant != LI_VARIANT_ATOMIC && !li::opt(li::OPT_AFFINE_PLAIN)) {
        rc = dtype == LI_F32 ? li::affine_adjoint_gather<float>(ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev, beta, dst_dev,
                                                                d_err, st, &gathered)
                             : li::affine_adjoint_gather<double>(ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev, beta, dst_dev,
                                                                 d_err, st, &gathered);
        if (rc != LI_OK) return rc;
    }
    if (adjoint && !gathered) {
        rc = li::vscale_impl(dtype, dst_dev, nout, beta, st);  // :236
        if (rc != LI_OK) return rc;
    }
    if (!gathered)
    rc = dtype == LI_F32 ? li::affine_T<float>(adjoint, ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev,
                                                   beta, dst_dev, d_err, tiled, st)
                             : li::affine_T<double>(adjoint, ker1, ker2, R, n1, n2, m1, m2, nbatch, alpha, src_dev,
                                                    beta, dst_dev, d_err, tiled, st);
    if (rc != LI_OK || (flags & LI_FLAG_ASYNC)) return rc;
    int h_err = 0;
    LI_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    LI_CUDA(cudaStreamSynchronize(st));
    if (h_err >= 2) {
        li::set_error("internal: the gather-form affine adjoint ran out of room (%s)", h_err == 3 ? "candidate box" : "cell overflow list");
        return LI_ERR_UNSUPPORTED;
    }
    if (h_err) {
        li::set_error("InexactError: trunc(Int64, f) for a NaN or out-of-range coordinate");
        return LI_ERR_INEXACT;
    }
    return LI_OK;
}
