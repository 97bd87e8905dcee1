// li_normal.cu -- normal equations of a table-backed interpolator: A'*A, A'*W*A and regularize!
// (reference: AtA / AtWA / AtA! / AtWA!, src/sparse.jl:277-356; regularize / regularize!, src/sparse.jl:424-490).
//
// dst[j1, j2] = sum over the rows i and the tap pairs (s1, s2) with J[i, s1] = j1, J[i, s2] = j2 of
//               C[i, s2] * (w[i] * C[i, s1])          (w = 1 for AtA: C[i, s1] * C[i, s2])
// The reference scatters row by row.  Here a thread owns one row j1 of dst and walks the entries of column j1 of the
// transposed table (ascending (row, tap): the CSC arrays of the gather adjoint), so that every element of dst is summed
// by one thread in the reference's order (i, s1, s2): deterministic and equal to the sequential loops.  Products and
// sums go through round-to-nearest intrinsics (no contraction into FMA), as in the reference.
#include "li_common.cuh"
#include "li_weights.cuh"

namespace li {

template <typename T>
__global__ void atwa_kernel(const int32_t *__restrict__ colptr, const int32_t *__restrict__ tent, const int32_t *__restrict__ J,
                            const T *__restrict__ C, const T *__restrict__ w, int S, int64_t ncols, T *__restrict__ dst)
{
    using A = Ar<T>;
    const int64_t j1 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j1 >= ncols) return;
    for (int32_t k = colptr[j1]; k < colptr[j1 + 1]; ++k) {
        const int32_t e = tent[k];            // entry id = row * S + tap
        const int64_t i = e / S;
        const T c1 = C[e];
        const T wc1 = w ? A::mul(w[i], c1) : c1;
        for (int s2 = 0; s2 < S; ++s2) {
            const int64_t j2 = J[i * S + s2];
            T *p = dst + j1 + j2 * ncols;
            *p = A::add(*p, w ? A::mul(C[i * S + s2], wc1) : A::mul(c1, C[i * S + s2]));
        }
    }
}

// rho = max of the diagonal (one block; the matrix is ncols x ncols of a 1-D fit: small)
template <typename T>
__global__ void diag_max_kernel(const T *__restrict__ a, int64_t n, T *__restrict__ rho)
{
    __shared__ T sm[256];
    T m = a[0];
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const T d = a[j + j * n];
        m = d > m ? d : m;    // (max(rho, d) of the reference; NaN diagonals are not propagated there either)
    }
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x + s] > sm[threadIdx.x] ? sm[threadIdx.x + s] : sm[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *rho = sm[0];
}

// A += rho * (eps * I + mu * D'*D), D the first-order finite differences (src/sparse.jl:463-488): per element the
// additions of the reference in its order (eps term, then the mu terms)
template <typename T>
__global__ void regularize_kernel(T *__restrict__ a, int64_t n, T eps, T mu, const T *__restrict__ rhop)
{
    using A = Ar<T>;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const T rho = *rhop;
    T d = a[j + j * n];
    if (eps > (T)0) d = A::add(d, A::mul(eps, rho));
    if (mu > (T)0) {
        const T q = A::mul(rho, mu);
        if (n >= 2) {
            d = A::add(d, (j == 0 || j == n - 1) ? q : A::add(q, q));
            if (j >= 1) a[j + (j - 1) * n] = A::sub(a[j + (j - 1) * n], q);       // A[i+1, i] -= q
            if (j + 1 < n) a[j + (j + 1) * n] = A::sub(a[j + (j + 1) * n], q);    // A[i-1, i] -= q
        } else {
            d = A::add(d, q);
        }
    }
    a[j + j * n] = d;
}

}  // namespace li

extern "C" {

int li_table_atwa(const li_table *t, const void *w_dev, void *dst_dev, void *stream)
{
    LI_REQUIRE(t && dst_dev, LI_ERR_ARG, "null argument");
    cudaStream_t st = li::as_stream(stream);
    const int64_t n = t->ncols;
    const size_t es = li::dtype_size(t->dtype);
    LI_CUDA(cudaMemsetAsync(dst_dev, 0, (size_t)n * n * es, st));   // fill!(dst, zero(T))
    if (t->nrows == 0) return LI_OK;
    int rc = li::table_ensure_csc(t, st);
    if (rc != LI_OK) return rc;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (t->dtype == LI_F32)
        li::atwa_kernel<float><<<blocks, 128, 0, st>>>(t->colptr, t->tent, t->J, (const float *)t->W, (const float *)w_dev, t->S, n,
                                                      (float *)dst_dev);
    else
        li::atwa_kernel<double><<<blocks, 128, 0, st>>>(t->colptr, t->tent, t->J, (const double *)t->W, (const double *)w_dev, t->S, n,
                                                       (double *)dst_dev);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

int li_regularize(int dtype, void *A_dev, int64_t n, double eps, double mu, void *stream)
{
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(A_dev && n >= 1, LI_ERR_ARG, "bad argument");
    LI_REQUIRE(eps >= 0.0 && mu >= 0.0, LI_ERR_ARG, "the regularization levels must be nonnegative");   // the reference's @assert
    if (!(eps > 0.0) && !(mu > 0.0)) return LI_OK;
    cudaStream_t st = li::as_stream(stream);
    void *rho = nullptr;
    int rc = li::staging(9, 16, &rho, stream);
    if (rc != LI_OK) return rc;
    double rho_host = 0.0;
    if (dtype == LI_F32) {
        li::diag_max_kernel<float><<<1, 256, 0, st>>>((const float *)A_dev, n, (float *)rho);
        float r = 0.f;
        LI_CUDA(cudaMemcpyAsync(&r, rho, sizeof r, cudaMemcpyDeviceToHost, st));
        LI_CUDA(cudaStreamSynchronize(st));
        rho_host = r;
    } else {
        li::diag_max_kernel<double><<<1, 256, 0, st>>>((const double *)A_dev, n, (double *)rho);
        LI_CUDA(cudaMemcpyAsync(&rho_host, rho, sizeof rho_host, cudaMemcpyDeviceToHost, st));
        LI_CUDA(cudaStreamSynchronize(st));
    }
    li::count_launch();
    LI_REQUIRE(rho_host > 0.0, LI_ERR_ARG, "we have a problem!");   // (the reference's message, src/sparse.jl:462)
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (dtype == LI_F32)
        li::regularize_kernel<float><<<blocks, 128, 0, st>>>((float *)A_dev, n, (float)eps, (float)mu, (const float *)rho);
    else
        li::regularize_kernel<double><<<blocks, 128, 0, st>>>((double *)A_dev, n, eps, mu, (const double *)rho);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

}  // extern "C"
