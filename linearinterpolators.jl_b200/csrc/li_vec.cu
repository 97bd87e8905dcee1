// li_vec.cu -- the vector operations and the linear conjugate-gradient driver that iterative fitting runs around the
// interpolation operators: `x = conjgrad(H'*H, H'*y)` of test/benchmarks.jl:31-40 (LazyAlgebra's `vdot`, `vupdate!`,
// `vcombine!`, `conjgrad`; purpose stated at README.md:36-37).  Everything stays on the device: the scalars of the
// recurrence (rho, gamma, ...) live in device memory and are consumed by the next kernel, so an iteration costs no
// host round trip; the host only looks at the residual every `check_every` iterations.
//
// Dot products are accumulated in Float64 with a fixed reduction tree (a fixed grid of partial sums added up in
// index order by one block): deterministic run to run.
#include <math.h>

#include <algorithm>

#include "li_common.cuh"

namespace li {

constexpr int VD_BLOCKS = 148 * 4, VD_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    if (w == 0) {
        s = l < (blockDim.x >> 5) ? sh[l] : 0.0;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    return s;   // valid in warp 0
}

// partial[b] = sum over the block's grid-stride slice of x*y (Float64)
template <typename T>
__global__ void __launch_bounds__(VD_THREADS) vdot_partial_kernel(const T *__restrict__ x, const T *__restrict__ y, int64_t n,
                                                                   double *__restrict__ partial)
{
    __shared__ double sh[VD_THREADS / 32];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += (double)x[i] * (double)y[i];
    const double s = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// what the single-block epilogue does with the sum of the partials
enum { FIN_STORE = 0, FIN_RHO = 1, FIN_GAMMA = 2 };
// CG scalars (device): 0 rho, 1 rho_old, 2 gamma, 3 psi = alpha*rho/2, 4 psimax, 5 rho0, 6 breakdown flag
template <int MODE>
__global__ void __launch_bounds__(VD_THREADS) vdot_final_kernel(const double *__restrict__ partial, int np, double *__restrict__ out)
{
    __shared__ double sh[VD_THREADS / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) acc += partial[i];
    const double s = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        if (MODE == FIN_STORE) {
            out[0] = s;
        } else if (MODE == FIN_RHO) {   // rho_old <- rho, rho <- s
            out[1] = out[0];
            out[0] = s;
        } else {                        // gamma <- s; psi = rho^2 / (2 gamma)
            out[2] = s;
            if (!(s > 0.0)) out[6] = 1.0;   // operator not positive definite (LazyAlgebra: "operator is not positive definite")
            const double psi = s > 0.0 ? 0.5 * out[0] * out[0] / s : 0.0;
            out[3] = psi;
            out[4] = fmax(out[4], psi);
        }
    }
}

template <typename T>
__global__ void vupdate_kernel(T *__restrict__ y, T alpha, const T *__restrict__ x, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = fma(alpha, x[i], y[i]);
}

template <typename T>
__global__ void vcombine_kernel(T *dst, T alpha, const T *x, T beta, const T *y, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T bx = beta == (T)0 ? (T)0 : beta * y[i];   // beta == 0: y is never read
        dst[i] = fma(alpha, x[i], bx);
    }
}

// p <- r + (rho / rho_old) * p   (first iteration: p <- r)
template <typename T>
__global__ void cg_direction_kernel(T *__restrict__ p, const T *__restrict__ r, const double *__restrict__ sc, int first, int64_t n)
{
    const T beta = first ? (T)0 : (T)(sc[0] / sc[1]);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = first ? r[i] : fma(beta, p[i], r[i]);
}

// x += alpha p, r -= alpha q with alpha = rho / gamma; partial sums of the new r.r
template <typename T>
__global__ void __launch_bounds__(VD_THREADS) cg_step_kernel(T *__restrict__ x, T *__restrict__ r, const T *__restrict__ p, const T *__restrict__ q,
                                                              const double *__restrict__ sc, int64_t n, double *__restrict__ partial)
{
    __shared__ double sh[VD_THREADS / 32];
    const double g = sc[2];
    const T alpha = g > 0.0 ? (T)(sc[0] / g) : (T)0;
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        x[i] = fma(alpha, p[i], x[i]);
        const T ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        acc += (double)ri * (double)ri;
    }
    const double s = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

static unsigned grid_for(int64_t n) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + VD_THREADS - 1) / VD_THREADS, VD_BLOCKS)); }

template <typename T>
static int vdot_T(const T *x, const T *y, int64_t n, double *partial, double *out, int mode, cudaStream_t st)
{
    const unsigned blocks = grid_for(n);
    vdot_partial_kernel<T><<<blocks, VD_THREADS, 0, st>>>(x, y, n, partial);
    if (mode == FIN_STORE)
        vdot_final_kernel<FIN_STORE><<<1, VD_THREADS, 0, st>>>(partial, (int)blocks, out);
    else if (mode == FIN_RHO)
        vdot_final_kernel<FIN_RHO><<<1, VD_THREADS, 0, st>>>(partial, (int)blocks, out);
    else
        vdot_final_kernel<FIN_GAMMA><<<1, VD_THREADS, 0, st>>>(partial, (int)blocks, out);
    count_launch(2);
    LI_CUDA(cudaGetLastError());
    return LI_OK;
}

static int vdot_any(int dtype, const void *x, const void *y, int64_t n, double *partial, double *out, int mode, cudaStream_t st)
{
    return dtype == LI_F32 ? vdot_T<float>((const float *)x, (const float *)y, n, partial, out, mode, st)
                           : vdot_T<double>((const double *)x, (const double *)y, n, partial, out, mode, st);
}

// per (thread, device, stream) scratch: VD_BLOCKS partial sums + 8 scalars
static int vec_scratch(void *stream, double **partial, double **scalars)
{
    void *p = nullptr;
    int rc = staging(7, (VD_BLOCKS + 8) * sizeof(double), &p, stream);
    if (rc != LI_OK) return rc;
    *partial = (double *)p;
    *scalars = (double *)p + VD_BLOCKS;
    return LI_OK;
}

}  // namespace li

extern "C" {

int li_vdot_async(int dtype, const void *x_dev, const void *y_dev, int64_t n, double *result_dev, void *stream)
{
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(n >= 0 && result_dev && ((x_dev && y_dev) || n == 0), LI_ERR_ARG, "null argument");
    double *partial = nullptr, *sc = nullptr;
    int rc = li::vec_scratch(stream, &partial, &sc);
    if (rc != LI_OK) return rc;
    return li::vdot_any(dtype, x_dev, y_dev, n, partial, result_dev, li::FIN_STORE, li::as_stream(stream));
}

int li_vdot(int dtype, const void *x_dev, const void *y_dev, int64_t n, double *result_host, void *stream)
{
    LI_REQUIRE(result_host, LI_ERR_ARG, "null argument");
    double *partial = nullptr, *sc = nullptr;
    int rc = li::vec_scratch(stream, &partial, &sc);
    if (rc != LI_OK) return rc;
    rc = li_vdot_async(dtype, x_dev, y_dev, n, sc + 7, stream);
    if (rc != LI_OK) return rc;
    LI_CUDA(cudaMemcpyAsync(result_host, sc + 7, sizeof(double), cudaMemcpyDeviceToHost, li::as_stream(stream)));
    LI_CUDA(cudaStreamSynchronize(li::as_stream(stream)));
    return LI_OK;
}

int li_vupdate(int dtype, void *y_dev, double alpha, const void *x_dev, int64_t n, void *stream)
{
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(n >= 0 && ((x_dev && y_dev) || n == 0), LI_ERR_ARG, "null argument");
    if (n == 0 || alpha == 0.0) return LI_OK;
    cudaStream_t st = li::as_stream(stream);
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    if (dtype == LI_F32)
        li::vupdate_kernel<float><<<blocks, 256, 0, st>>>((float *)y_dev, (float)alpha, (const float *)x_dev, n);
    else
        li::vupdate_kernel<double><<<blocks, 256, 0, st>>>((double *)y_dev, alpha, (const double *)x_dev, n);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

int li_vcombine(int dtype, void *dst_dev, double alpha, const void *x_dev, double beta, const void *y_dev, int64_t n, void *stream)
{
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(n >= 0 && ((dst_dev && x_dev && (y_dev || beta == 0.0)) || n == 0), LI_ERR_ARG, "null argument");
    if (n == 0) return LI_OK;
    cudaStream_t st = li::as_stream(stream);
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    if (dtype == LI_F32)
        li::vcombine_kernel<float><<<blocks, 256, 0, st>>>((float *)dst_dev, (float)alpha, (const float *)x_dev, (float)beta,
                                                           (const float *)y_dev, n);
    else
        li::vcombine_kernel<double><<<blocks, 256, 0, st>>>((double *)dst_dev, alpha, (const double *)x_dev, beta, (const double *)y_dev, n);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

// Linear conjugate gradients for A x = b, A symmetric positive (semi-)definite, given as a callback that applies it
// to a device vector.  LazyAlgebra's conjgrad!: r = b - A x; p = r + (rho/rho_old) p; q = A p; alpha = rho / <p,q>;
// x += alpha p; r -= alpha q; convergence on the gradient norm sqrt(rho) <= max(gtol_abs, gtol_rel * sqrt(rho0)) and on
// the decrease of the objective psi = alpha*rho/2 <= ftol * max(psi).
int li_conjgrad(li_apply_fn A, void *ctx, int dtype, int64_t n, const void *b_dev, void *x_dev, int x_is_zero,
                const li_cg_options *opts, li_cg_info *info, void *stream)
{
    LI_REQUIRE(A && b_dev && x_dev, LI_ERR_ARG, "null argument");
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(n >= 1, LI_ERR_DIM, "empty system");
    li_cg_options o;
    o.maxiter = 50, o.ftol = 1e-8, o.gtol_abs = 0.0, o.gtol_rel = 0.0, o.check_every = 5;
    if (opts) o = *opts;
    if (o.check_every < 1) o.check_every = 1;
    cudaStream_t st = li::as_stream(stream);
    const size_t es = li::dtype_size(dtype);
    void *work = nullptr;   // r, p, q
    LI_CUDA(cudaMalloc(&work, 3 * (size_t)n * es));
    char *r = (char *)work, *p = r + (size_t)n * es, *q = p + (size_t)n * es;
    double *partial = nullptr, *sc = nullptr;
    int rc = li::vec_scratch(stream, &partial, &sc);
    li_cg_info inf;
    inf.iters = 0, inf.status = LI_CG_MAXITER, inf.rho0 = inf.rho = 0.0;
    auto fail = [&](int code) {
        cudaStreamSynchronize(st);
        cudaFree(work);
        if (info) *info = inf;
        return code;
    };
    if (rc != LI_OK) return fail(rc);
    const unsigned blocks = li::grid_for(n);
    cudaError_t e = cudaMemsetAsync(sc, 0, 8 * sizeof(double), st);
    // r = b - A x
    if (e == cudaSuccess) {
        if (x_is_zero) {
            e = cudaMemsetAsync(x_dev, 0, (size_t)n * es, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(r, b_dev, (size_t)n * es, cudaMemcpyDeviceToDevice, st);
        } else {
            rc = A(ctx, x_dev, q, stream);
            if (rc != LI_OK) return fail(rc);
            rc = li_vcombine(dtype, r, 1.0, b_dev, -1.0, q, n, stream);
            if (rc != LI_OK) return fail(rc);
        }
    }
    if (e != cudaSuccess) {
        li::set_error("conjgrad: %s", cudaGetErrorString(e));
        return fail(LI_ERR_CUDA);
    }
    rc = li::vdot_any(dtype, r, r, n, partial, sc, li::FIN_STORE, st);
    if (rc != LI_OK) return fail(rc);
    double h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto peek = [&]() -> int {
        LI_CUDA(cudaMemcpyAsync(h, sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        LI_CUDA(cudaStreamSynchronize(st));
        return LI_OK;
    };
    rc = peek();
    if (rc != LI_OK) return fail(rc);
    inf.rho0 = inf.rho = h[0];
    const double gtest = std::max(o.gtol_abs, o.gtol_rel * sqrt(h[0]));
    if (sqrt(h[0]) <= gtest || h[0] == 0.0) {
        inf.status = LI_CG_GTEST;
        return fail(LI_OK);
    }
    for (int k = 0; k < o.maxiter; ++k) {
        if (dtype == LI_F32) {
            li::cg_direction_kernel<float><<<blocks, li::VD_THREADS, 0, st>>>((float *)p, (const float *)r, sc, k == 0, n);
        } else {
            li::cg_direction_kernel<double><<<blocks, li::VD_THREADS, 0, st>>>((double *)p, (const double *)r, sc, k == 0, n);
        }
        li::count_launch();
        rc = A(ctx, p, q, stream);
        if (rc != LI_OK) return fail(rc);
        rc = li::vdot_any(dtype, p, q, n, partial, sc, li::FIN_GAMMA, st);
        if (rc != LI_OK) return fail(rc);
        if (dtype == LI_F32)
            li::cg_step_kernel<float><<<blocks, li::VD_THREADS, 0, st>>>((float *)x_dev, (float *)r, (const float *)p, (const float *)q, sc, n, partial);
        else
            li::cg_step_kernel<double><<<blocks, li::VD_THREADS, 0, st>>>((double *)x_dev, (double *)r, (const double *)p, (const double *)q, sc, n,
                                                                        partial);
        li::vdot_final_kernel<li::FIN_RHO><<<1, li::VD_THREADS, 0, st>>>(partial, (int)blocks, sc);
        li::count_launch(2);
        inf.iters = k + 1;
        if ((k + 1) % o.check_every == 0 || k + 1 == o.maxiter) {
            rc = peek();
            if (rc != LI_OK) return fail(rc);
            inf.rho = h[0];
            if (h[6] != 0.0) {
                inf.status = LI_CG_NOT_POSITIVE;
                li::set_error("conjgrad: operator is not positive definite");
                return fail(LI_ERR_ARG);
            }
            if (sqrt(h[0]) <= gtest) {
                inf.status = LI_CG_GTEST;
                break;
            }
            if (h[3] <= o.ftol * h[4]) {
                inf.status = LI_CG_FTEST;
                break;
            }
        }
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        li::set_error("conjgrad: %s", cudaGetErrorString(e));
        return fail(LI_ERR_CUDA);
    }
    return fail(LI_OK);
}

// the separable case without any callback into the host language: A = H'*H + mu*I, H = A_1 * ... * A_ndim
struct li_gram_ctx {
    int ndim;
    const li_table *const *tabs;
    const int64_t *dims_other;
    int64_t nbatch, nrows_total;
    double mu;
    int dtype;
    void *tmp, *ws;
    size_t wsb;
};

static int li_gram_apply(void *vctx, const void *x, void *y, void *stream)
{
    li_gram_ctx *c = (li_gram_ctx *)vctx;
    int rc = li_sep_apply(0, c->ndim, c->tabs, c->dims_other, c->nbatch, 1.0, x, 0.0, c->tmp, c->ws, c->wsb, LI_VARIANT_AUTO, stream);
    if (rc != LI_OK) return rc;
    // y = H' tmp (+ mu x)
    rc = li_sep_apply(1, c->ndim, c->tabs, c->dims_other, c->nbatch, 1.0, c->tmp, 0.0, y, c->ws, c->wsb, LI_VARIANT_AUTO, stream);
    if (rc != LI_OK || c->mu == 0.0) return rc;
    int64_t ncols_total = c->nbatch;
    for (int k = 0; k < c->ndim; ++k) ncols_total *= c->tabs[k] ? c->tabs[k]->ncols : c->dims_other[k];
    return li_vupdate(c->dtype, y, c->mu, x, ncols_total, stream);
}

int li_conjgrad_sep(int ndim, const li_table *const *tabs, const int64_t *dims_other, int64_t nbatch, double mu,
                    const void *y_dev, void *x_dev, int x_is_zero, const li_cg_options *opts, li_cg_info *info, void *stream)
{
    LI_REQUIRE(tabs && y_dev && x_dev && ndim >= 1 && ndim <= 3, LI_ERR_ARG, "bad argument");
    int dtype = -1;
    int64_t nin = nbatch, nout = nbatch;
    for (int k = 0; k < ndim; ++k) {
        if (tabs[k]) {
            dtype = tabs[k]->dtype;
            nin *= tabs[k]->ncols, nout *= tabs[k]->nrows;
        } else {
            LI_REQUIRE(dims_other, LI_ERR_DIM, "missing length of an untouched dimension");
            nin *= dims_other[k], nout *= dims_other[k];
        }
    }
    LI_REQUIRE(dtype >= 0, LI_ERR_ARG, "at least one table is required");
    cudaStream_t st = li::as_stream(stream);
    const size_t es = li::dtype_size(dtype);
    li_gram_ctx c;
    c.ndim = ndim, c.tabs = tabs, c.dims_other = dims_other, c.nbatch = nbatch, c.mu = mu, c.dtype = dtype;
    c.nrows_total = nout;
    size_t wa = 0, wb = 0;
    int rc = li_sep_workspace_bytes(0, ndim, tabs, dims_other, nbatch, &wa);
    if (rc == LI_OK) rc = li_sep_workspace_bytes(1, ndim, tabs, dims_other, nbatch, &wb);
    if (rc != LI_OK) return rc;
    c.wsb = std::max(wa, wb);
    void *buf = nullptr;   // tmp (rows side), b = H'y (columns side), workspace
    const size_t o1 = ((size_t)nout * es + 255) & ~(size_t)255, o2 = o1 + (((size_t)nin * es + 255) & ~(size_t)255);
    LI_CUDA(cudaMalloc(&buf, o2 + c.wsb + 256));
    c.tmp = buf;
    void *b = (char *)buf + o1;
    c.ws = c.wsb ? (char *)buf + o2 : nullptr;
    // b = H' y
    rc = li_sep_apply(1, ndim, tabs, dims_other, nbatch, 1.0, y_dev, 0.0, b, c.ws, c.wsb, LI_VARIANT_AUTO, stream);
    if (rc == LI_OK) rc = li_conjgrad(li_gram_apply, &c, dtype, nin, b, x_dev, x_is_zero, opts, info, stream);
    cudaStreamSynchronize(st);
    cudaFree(buf);
    return rc;
}

}  // extern "C"
