// li_sep.cu -- separable N-d operators H = A_1 * A_2 * ... * A_n and the on-the-fly
// front ends that build their coefficients on the device per call.
//
// Reference: LazyAlgebra composition of per-dimension operators
// (test/benchmarks.jl:27-29,44-49; H*x = A_1*(...*(A_n*x)), H'*y = A_n'*(...*(A_1'*y)))
// and the on-the-fly operators src/unidimensional.jl:95-245, src/separable.jl:56-162.
#include <algorithm>
#include <map>
#include <vector>

#include "li_common.cuh"

namespace li {

struct SepPlan {
    int npass = 0;
    int dim[3];           // dimension (0-based) of each pass, in application order
    int64_t npre[3], npost[3];
    int64_t out_elems[3];  // elements of the array produced by each pass
    int64_t in_elems = 0;
};

static int make_plan(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other, int64_t nbatch,
                     SepPlan *plan)
{
    LI_REQUIRE(ndim >= 1 && ndim <= 3, LI_ERR_UNSUPPORTED, "separable operators support 1 to 3 dimensions");
    LI_REQUIRE(tabs, LI_ERR_ARG, "null table list");
    LI_REQUIRE(nbatch >= 0, LI_ERR_DIM, "negative batch size");
    int64_t cur[3];
    int dtype = -1;
    for (int k = 0; k < ndim; ++k) {
        if (tabs[k]) {
            cur[k] = adjoint ? tabs[k]->nrows : tabs[k]->ncols;
            if (dtype < 0) dtype = tabs[k]->dtype;
            LI_REQUIRE(dtype == tabs[k]->dtype, LI_ERR_ARG, "tables of a separable operator must share one floating-point type");
        } else {
            LI_REQUIRE(dims_other && dims_other[k] >= 0, LI_ERR_DIM, "missing length of an untouched dimension");
            cur[k] = dims_other[k];
        }
    }
    LI_REQUIRE(dtype >= 0, LI_ERR_ARG, "at least one table is required");
    plan->in_elems = nbatch;
    for (int k = 0; k < ndim; ++k) plan->in_elems *= cur[k];
    plan->npass = 0;
    // Direct applies A_n first; Adjoint applies A_1' first.
    for (int step = 0; step < ndim; ++step) {
        const int k = adjoint ? step : ndim - 1 - step;
        if (!tabs[k]) continue;
        int64_t pre = 1, post = nbatch;
        for (int d = 0; d < k; ++d) pre *= cur[d];
        for (int d = k + 1; d < ndim; ++d) post *= cur[d];
        const int p = plan->npass++;
        plan->dim[p] = k;
        plan->npre[p] = pre;
        plan->npost[p] = post;
        cur[k] = adjoint ? tabs[k]->ncols : tabs[k]->nrows;
        plan->out_elems[p] = pre * post * cur[k];
    }
    return LI_OK;
}

template <typename T>
int sep3d_fused_T(int adjoint, const li_table *t1, const li_table *t2, const li_table *t3, int64_t nb, double alpha,
                  const void *src, double beta, void *dst, cudaStream_t st, bool *done);   // li_sep3d.cuh

static int sep_workspace(const SepPlan &plan, int dtype, size_t *bytes, size_t *half)
{
    // ping-pong buffers holding every intermediate (all but the last pass's output)
    int64_t m = 0;
    for (int p = 0; p + 1 < plan.npass; ++p) m = std::max(m, plan.out_elems[p]);
    const size_t one = ((size_t)m * dtype_size(dtype) + 255) & ~(size_t)255;
    *half = one;
    *bytes = plan.npass >= 3 ? 2 * one : (plan.npass == 2 ? one : 0);
    return LI_OK;
}

static int sep_apply_impl(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other,
                          int64_t nbatch, double alpha, const void *src, double beta, void *dst, void *workspace,
                          size_t workspace_bytes, int variant, cudaStream_t st)
{
    SepPlan plan;
    int rc = make_plan(adjoint, ndim, tabs, dims_other, nbatch, &plan);
    if (rc != LI_OK) return rc;
    int dtype = LI_F32;
    for (int k = 0; k < ndim; ++k)
        if (tabs[k]) dtype = tabs[k]->dtype;
    LI_REQUIRE(dst, LI_ERR_ARG, "null destination");
    if (plan.npass == 0 || plan.in_elems == 0) return LI_OK;
    const int64_t nout = plan.out_elems[plan.npass - 1];
    if (alpha == 0.0) return vscale_impl(dtype, dst, nout, beta, st);
    LI_REQUIRE(src, LI_ERR_ARG, "null source");

    // fused single-pass kernel for two windowed tables on dimensions 1 and 2
    if (ndim == 2 && tabs[0] && tabs[1] && variant != LI_VARIANT_ATOMIC) {
        bool done = false;
        rc = sep2d_fused_apply(adjoint, tabs[0], tabs[1], nbatch, alpha, src, beta, dst, st, &done);
        if (rc != LI_OK || done) return rc;
    }

    // ... also when a third, untouched dimension follows: it is one more batch dimension
    if (ndim == 3 && tabs[0] && tabs[1] && !tabs[2] && variant != LI_VARIANT_ATOMIC) {
        bool done = false;
        rc = sep2d_fused_apply(adjoint, tabs[0], tabs[1], dims_other[2] * nbatch, alpha, src, beta, dst, st, &done);
        if (rc != LI_OK || done) return rc;
    }

    // three windowed tables: one pass (li_sep3d.cuh); it declines what it cannot stage
    if (ndim == 3 && tabs[0] && tabs[1] && tabs[2] && variant != LI_VARIANT_ATOMIC && !opt(OPT_NO_FUSED) &&
        !(opt(OPT_EXP) & 2) && sep2d_fused_supported(tabs[0], tabs[1]) && sep2d_fused_supported(tabs[0], tabs[2])) {
        bool done = false;
        rc = dtype == LI_F32
                 ? sep3d_fused_T<float>(adjoint, tabs[0], tabs[1], tabs[2], nbatch, alpha, src, beta, dst, st, &done)
                 : sep3d_fused_T<double>(adjoint, tabs[0], tabs[1], tabs[2], nbatch, alpha, src, beta, dst, st, &done);
        if (rc != LI_OK || done) return rc;
    }

    size_t need = 0, half = 0;
    sep_workspace(plan, dtype, &need, &half);
    LI_REQUIRE(workspace_bytes >= need && (need == 0 || workspace), LI_ERR_ARG,
               "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);

    // three dimensions: the pass along dim 3 on its own, dims 1 and 2 fused with dim 3 as the batch (2 passes over
    // HBM instead of 3)
    if (ndim == 3 && tabs[0] && tabs[1] && tabs[2] && variant != LI_VARIANT_ATOMIC &&
        sep2d_fused_supported(tabs[0], tabs[1])) {
        const li_table *t1 = tabs[0], *t2 = tabs[1], *t3 = tabs[2];
        bool done = false;
        if (!adjoint) {
            // A3 first: [m1, m2, m3] -> ws [m1, m2, n3]; then A1*A2 on every slice
            rc = tab_apply_impl(0, t3, t1->ncols * t2->ncols, nbatch, 1.0, src, 0.0, workspace, variant, st);
            if (rc != LI_OK) return rc;
            rc = sep2d_fused_apply(0, t1, t2, t3->nrows * nbatch, alpha, workspace, beta, dst, st, &done);
            if (rc != LI_OK || done) return rc;
            // declined: finish with the remaining per-dimension passes
            void *ws2 = (char *)workspace + half;
            rc = tab_apply_impl(0, t2, t1->ncols, t3->nrows * nbatch, 1.0, workspace, 0.0, ws2, variant, st);
            if (rc != LI_OK) return rc;
            return tab_apply_impl(0, t1, 1, t2->nrows * t3->nrows * nbatch, alpha, ws2, beta, dst, variant, st);
        } else {
            // (A1*A2)' on every slice: [n1, n2, n3] -> ws [m1, m2, n3]; then A3'
            rc = sep2d_fused_apply(1, t1, t2, t3->nrows * nbatch, 1.0, src, 0.0, workspace, st, &done);
            if (rc != LI_OK) return rc;
            if (done) return tab_apply_impl(1, t3, t1->ncols * t2->ncols, nbatch, alpha, workspace, beta, dst, variant, st);
        }
    }
    const void *in = src;
    for (int p = 0; p < plan.npass; ++p) {
        const bool last = p == plan.npass - 1;
        void *out = last ? dst : (char *)workspace + (p & 1) * half;
        rc = tab_apply_impl(adjoint, tabs[plan.dim[p]], plan.npre[p], plan.npost[p], last ? alpha : 1.0, in,
                            last ? beta : 0.0, out, variant, st);
        if (rc != LI_OK) return rc;
        in = out;
    }
    return LI_OK;
}

}  // namespace li

namespace li {

int sep_accumulate_multi(int64_t n, const li_table *const *tabs, const void *const *src_dev, double alpha, double beta,
                         void *dst, int64_t row_lo, int64_t row_hi, cudaStream_t st, bool *done);   // li_comm.cu

template <typename T>
__global__ void sum_partials_kernel(T *__restrict__ dst, const T *__restrict__ part, int64_t n, int np)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T v = dst[i];
        for (int k = 0; k < np; ++k) v += part[(int64_t)k * n + i];   // fixed order: deterministic
        dst[i] = v;
    }
}

// Many fused 2-D applies accumulated into ONE destination (dst_k all equal, e.g. m = beta*m + alpha * sum_k H_k' y_k,
// the inner loop of an iterative fit): one apply is a short, latency-bound kernel and a chain of them on one stream
// leaves the machine mostly idle, so the frames are dealt round-robin to NSTR streams, each accumulating into its own
// partial result (stream 0 straight into dst), and the partials are added in a fixed order at the end.  The
// assignment of frames to streams is static, so the result is deterministic.  *ndone = number of leading frames
// that were accumulated (0 when the batch is not of that form; less than n when a fused kernel declined a frame --
// the caller continues from there one frame at a time).
int sep_accumulate_concurrent(int adjoint, int ndim, int64_t n, const li_table *const *tabs, const int64_t *dims_other,
                              int64_t nbatch, double alpha, const void *const *src_dev, double beta, void *const *dst_dev,
                              int variant, void *stream, int64_t *ndone)
{
    constexpr int NSTR = 4;
    *ndone = 0;
    // (beta must be 1: with any other value the in-order semantics rescales the sum at every frame)
    if (ndim != 2 || n < 2 || variant == LI_VARIANT_ATOMIC || alpha == 0.0 || beta != 1.0) return LI_OK;
    // adjoints into one destination: the multi-frame accumulate kernel (li_sep2d_multi.cuh), one launch for the batch
    if (adjoint && nbatch == 1) {
        bool same = true, full = true;
        for (int64_t k = 0; k < n && same; ++k) {
            same = dst_dev[k] == dst_dev[0];
            full = full && tabs[2 * k] && tabs[2 * k + 1];
        }
        if (same && full) {
            bool done = false;
            int rc = sep_accumulate_multi(n, tabs, src_dev, alpha, beta, dst_dev[0], 0, tabs[1]->ncols, as_stream(stream), &done);
            if (rc != LI_OK) return rc;
            if (done) {
                *ndone = n;
                return LI_OK;
            }
        }
    }
    if (n < 2 * NSTR) return LI_OK;
    for (int64_t k = 0; k < n; ++k) {
        const li_table *const *tk = tabs + k * ndim;
        if (dst_dev[k] != dst_dev[0] || !tk[0] || !tk[1] || !sep2d_fused_supported(tk[0], tk[1])) return LI_OK;
        if (tk[0]->dtype != tabs[0]->dtype) return LI_OK;
        // same destination shape for every frame
        const int64_t a0 = adjoint ? tk[0]->ncols : tk[0]->nrows, a1 = adjoint ? tk[1]->ncols : tk[1]->nrows;
        const int64_t b0 = adjoint ? tabs[0]->ncols : tabs[0]->nrows, b1 = adjoint ? tabs[1]->ncols : tabs[1]->nrows;
        if (a0 != b0 || a1 != b1) return LI_OK;
    }
    cudaStream_t st = as_stream(stream);
    const int dtype = tabs[0]->dtype;
    const int64_t nout = (adjoint ? tabs[0]->ncols * tabs[1]->ncols : tabs[0]->nrows * tabs[1]->nrows) * nbatch;
    if (nout <= 0) return LI_OK;
    void *part = nullptr;
    int rc = staging(3, (size_t)(NSTR - 1) * nout * dtype_size(dtype), &part, stream);
    if (rc != LI_OK) return rc;
    struct Side {
        cudaStream_t side[NSTR - 1] = {};
        cudaEvent_t ev_fork = nullptr, ev_join[NSTR - 1] = {};
    };
    static thread_local std::map<int, Side> sides;   // per device: streams and events belong to one
    Side &sd = sides[current_device()];
    cudaStream_t(&side)[NSTR - 1] = sd.side;
    cudaEvent_t &ev_fork = sd.ev_fork;
    cudaEvent_t(&ev_join)[NSTR - 1] = sd.ev_join;
    if (!ev_fork) {
        LI_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        for (int s = 0; s < NSTR - 1; ++s) {
            LI_CUDA(cudaStreamCreateWithFlags(&side[s], cudaStreamNonBlocking));
            LI_CUDA(cudaEventCreateWithFlags(&ev_join[s], cudaEventDisableTiming));
        }
    }
    // one-off table metadata is built on the caller's stream before the fork (those builds synchronise)
    LI_CUDA(cudaEventRecord(ev_fork, st));
    for (int s = 0; s < NSTR - 1; ++s) LI_CUDA(cudaStreamWaitEvent(side[s], ev_fork, 0));
    int64_t k = 0;
    for (; k < n; ++k) {
        const int s = (int)(k % NSTR);
        cudaStream_t sk = s == 0 ? st : side[s - 1];
        void *target = s == 0 ? dst_dev[0] : (void *)((char *)part + (size_t)(s - 1) * nout * dtype_size(dtype));
        const double bk = k >= NSTR ? 1.0 : (s == 0 ? beta : 0.0);   // first use of a target: dst keeps beta, partials start at 0
        bool ok = false;
        rc = sep2d_fused_apply(adjoint, tabs[k * ndim], tabs[k * ndim + 1], nbatch, alpha, src_dev[k], bk, target, sk, &ok);
        if (rc != LI_OK || !ok) break;   // (declined: the caller finishes the batch frame by frame)
    }
    for (int s = 0; s < NSTR - 1; ++s) {   // always join, also on errors
        cudaEventRecord(ev_join[s], side[s]);
        cudaStreamWaitEvent(st, ev_join[s], 0);
    }
    if (rc != LI_OK) return rc;
    const int np = (int)std::min<int64_t>(k, NSTR) - 1;   // partial accumulators in use
    if (np > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((nout + 255) / 256, 148 * 8);
        if (dtype == LI_F32)
            sum_partials_kernel<float><<<blocks, 256, 0, st>>>((float *)dst_dev[0], (const float *)part, nout, np);
        else
            sum_partials_kernel<double><<<blocks, 256, 0, st>>>((double *)dst_dev[0], (const double *)part, nout, np);
        LI_CHECK_LAUNCH();
    }
    *ndone = k;
    return LI_OK;
}

}  // namespace li

extern "C" {

int li_sep_workspace_bytes(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other,
                           int64_t nbatch, size_t *bytes)
{
    LI_REQUIRE(bytes, LI_ERR_ARG, "null argument");
    li::SepPlan plan;
    int rc = li::make_plan(adjoint, ndim, tabs, dims_other, nbatch, &plan);
    if (rc != LI_OK) return rc;
    int dtype = LI_F32;
    for (int k = 0; k < ndim; ++k)
        if (tabs[k]) dtype = tabs[k]->dtype;
    size_t half = 0;
    return li::sep_workspace(plan, dtype, bytes, &half);  // upper bound (the fused 2-D kernel needs none)
}

int li_sep_apply(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other, int64_t nbatch,
                 double alpha, const void *src_dev, double beta, void *dst_dev, void *workspace_dev,
                 size_t workspace_bytes, int variant, void *stream)
{
    return li::sep_apply_impl(adjoint, ndim, tabs, dims_other, nbatch, alpha, src_dev, beta, dst_dev, workspace_dev,
                              workspace_bytes, variant, li::as_stream(stream));
}

int li_sep_apply_batch(int adjoint, int ndim, int64_t n, const li_table *const *tabs, const int64_t *dims_other,
                       int64_t nbatch, double alpha, const void *const *src_dev, double beta, void *const *dst_dev,
                       int variant, void *stream)
{
    LI_REQUIRE(n >= 0 && tabs && (n == 0 || (src_dev && dst_dev)), LI_ERR_ARG, "null argument");
    cudaStream_t st = li::as_stream(stream);
    int64_t k0 = 0;
    {
        int rc = li::sep_accumulate_concurrent(adjoint, ndim, n, tabs, dims_other, nbatch, alpha, src_dev, beta, dst_dev,
                                               variant, stream, &k0);
        if (rc != LI_OK) return rc;
        if (k0 > 0) beta = 1.0;   // the destination already holds beta*dst + the first k0 frames
    }
    for (int64_t k = k0; k < n; ++k) {
        const li_table *const *tk = tabs + k * ndim;
        size_t need = 0;
        int rc = li_sep_workspace_bytes(adjoint, ndim, tk, dims_other, nbatch, &need);
        if (rc != LI_OK) return rc;
        void *ws = nullptr;
        // (the fused 2-D kernels need no workspace; this buffer is only touched by the multi-pass fallback)
        if (need && !(ndim == 2 && tk[0] && tk[1] && li::sep2d_fused_supported(tk[0], tk[1]))) {
            rc = li::staging(2, need, &ws, stream);
            if (rc != LI_OK) return rc;
        } else {
            need = 0;
        }
        rc = li::sep_apply_impl(adjoint, ndim, tk, dims_other, nbatch, alpha, src_dev[k], beta, dst_dev[k], ws, need,
                                variant, st);
        if (rc == LI_ERR_ARG && need == 0) {   // the fused kernel declined: retry with a workspace
            rc = li_sep_workspace_bytes(adjoint, ndim, tk, dims_other, nbatch, &need);
            if (rc == LI_OK) rc = li::staging(2, need, &ws, stream);
            if (rc == LI_OK)
                rc = li::sep_apply_impl(adjoint, ndim, tk, dims_other, nbatch, alpha, src_dev[k], beta, dst_dev[k], ws,
                                        need, variant, st);
        }
        if (rc != LI_OK) return rc;
    }
    return LI_OK;
}

}  // extern "C"

static int sep_apply_host_impl(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other,
                               int64_t nbatch, double alpha, const void *src_host, double beta, void *dst_host,
                               int variant, void *stream, bool sync)
{
    LI_REQUIRE(src_host && dst_host, LI_ERR_ARG, "null argument");
    cudaStream_t st = li::as_stream(stream);
    li::SepPlan plan;
    int rc = li::make_plan(adjoint, ndim, tabs, dims_other, nbatch, &plan);
    if (rc != LI_OK) return rc;
    int dtype = LI_F32;
    for (int k = 0; k < ndim; ++k)
        if (tabs[k]) dtype = tabs[k]->dtype;
    if (plan.npass == 0) return LI_OK;
    const size_t es = li::dtype_size(dtype);
    const size_t nin = (size_t)plan.in_elems * es, nout = (size_t)plan.out_elems[plan.npass - 1] * es;
    size_t wsb = 0, half = 0;
    li::sep_workspace(plan, dtype, &wsb, &half);
    if (ndim == 2 && tabs[0] && tabs[1] && li::sep2d_fused_supported(tabs[0], tabs[1]) && variant != LI_VARIANT_ATOMIC)
        wsb = 0;   // the fused kernel needs none; if it declines the call below retries with a workspace
    void *dsrc = nullptr, *ddst = nullptr, *ws = nullptr;
    rc = li::staging(0, nin ? nin : 1, &dsrc, stream);
    if (rc == LI_OK) rc = li::staging(1, nout ? nout : 1, &ddst, stream);
    if (rc == LI_OK && wsb) rc = li::staging(2, wsb, &ws, stream);
    if (rc != LI_OK) return rc;
    // large fused 2-D applies: uploads, kernels and downloads pipelined over bands of rows
    if (ndim == 2 && tabs[0] && tabs[1] && nbatch == 1 && beta == 0.0 && alpha != 0.0 && variant != LI_VARIANT_ATOMIC &&
        nin + nout >= (size_t)(32u << 20)) {
        bool banded = false;
        rc = li::sep2d_host_banded(adjoint, tabs[0], tabs[1], alpha, src_host, dst_host, dsrc, ddst, st, &banded);
        if (rc != LI_OK) return rc;
        if (banded) {
            if (sync) LI_CUDA(cudaStreamSynchronize(st));
            return LI_OK;
        }
    }
    cudaError_t e = cudaMemcpyAsync(dsrc, src_host, nin, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && beta != 0.0) e = cudaMemcpyAsync(ddst, dst_host, nout, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        rc = li::sep_apply_impl(adjoint, ndim, tabs, dims_other, nbatch, alpha, dsrc, beta, ddst, ws, wsb, variant, st);
        if (rc == LI_ERR_ARG && wsb == 0) {   // fused kernel declined
            li::sep_workspace(plan, dtype, &wsb, &half);
            rc = li::staging(2, wsb, &ws, stream);
            if (rc == LI_OK)
                rc = li::sep_apply_impl(adjoint, ndim, tabs, dims_other, nbatch, alpha, dsrc, beta, ddst, ws, wsb, variant, st);
        }
    }
    if (e == cudaSuccess && rc == LI_OK) e = cudaMemcpyAsync(dst_host, ddst, nout, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        li::set_error("host-staged separable apply failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    return rc;
}

extern "C" {

int li_sep_apply_host(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other, int64_t nbatch,
                      double alpha, const void *src_host, double beta, void *dst_host, int variant, void *stream)
{
    return sep_apply_host_impl(adjoint, ndim, tabs, dims_other, nbatch, alpha, src_host, beta, dst_host, variant, stream, true);
}

int li_sep_apply_host_async(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other,
                            int64_t nbatch, double alpha, const void *src_host, double beta, void *dst_host,
                            int variant, void *stream)
{
    return sep_apply_host_impl(adjoint, ndim, tabs, dims_other, nbatch, alpha, src_host, beta, dst_host, variant, stream, false);
}

// 1-D on-the-fly: coefficients are generated on the device for this call only.
int li_otf1d_apply(int adjoint, const li_kernel_desc *ker, const void *x_dev, int64_t n, int64_t len, double alpha,
                   const void *src_dev, double beta, void *dst_dev, int variant, void *stream)
{
    LI_REQUIRE(ker && dst_dev, LI_ERR_ARG, "null argument");
    cudaStream_t st = li::as_stream(stream);
    if (n == 0) return adjoint ? li::vscale_impl(ker->dtype, dst_dev, len, beta, st) : LI_OK;
    li_table *t = nullptr;
    int rc = li::table_create_impl(ker, x_dev, n, len, st, &t);
    if (rc != LI_OK) return rc;
    rc = li::tab_apply_impl(adjoint, t, 1, 1, alpha, src_dev, beta, dst_dev, variant, st);
    cudaError_t e = cudaStreamSynchronize(st);
    li_table_destroy(t);
    if (rc == LI_OK && e != cudaSuccess) {
        li::set_error("on-the-fly apply failed: %s", cudaGetErrorString(e));
        rc = LI_ERR_CUDA;
    }
    return rc;
}

int li_sep_otf_apply(int adjoint, int ndim, const li_kernel_desc *const *kers, const void *const *x_dev, const int64_t *n,
                     const int64_t *m, int64_t nbatch, double alpha, const void *src_dev, double beta, void *dst_dev,
                     void *stream)
{
    LI_REQUIRE(ndim >= 1 && ndim <= 3, LI_ERR_UNSUPPORTED, "on-the-fly separable operators support 1 to 3 dimensions");
    LI_REQUIRE(kers && x_dev && n && m && dst_dev, LI_ERR_ARG, "null argument");
    for (int k = 0; k < ndim; ++k) {
        LI_REQUIRE(kers[k], LI_ERR_ARG, "null kernel");
        LI_REQUIRE(kers[k]->dtype == kers[0]->dtype, LI_ERR_ARG, "kernels must share one floating-point type");
    }
    cudaStream_t st = li::as_stream(stream);
    li_table *t[3] = {nullptr, nullptr, nullptr};
    int rc = LI_OK;
    for (int k = 0; k < ndim && rc == LI_OK; ++k) rc = li::table_create_impl(kers[k], x_dev[k], n[k], m[k], st, &t[k]);
    void *ws = nullptr;
    size_t wsb = 0;
    // (the fused kernels need no workspace: try without, and only allocate the scratch of the per-dimension passes
    // when they decline)
    if (rc == LI_OK) {
        rc = li::sep_apply_impl(adjoint, ndim, t, nullptr, nbatch, alpha, src_dev, beta, dst_dev, nullptr, 0, LI_VARIANT_AUTO, st);
        if (rc == LI_ERR_ARG && strstr(li_last_error(), "workspace too small")) {
            rc = li_sep_workspace_bytes(adjoint, ndim, t, nullptr, nbatch, &wsb);
            if (rc == LI_OK && wsb) {
                cudaError_t e = cudaMalloc(&ws, wsb);
                if (e != cudaSuccess) {
                    li::set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
                    rc = LI_ERR_NOMEM;
                }
            }
            if (rc == LI_OK)
                rc = li::sep_apply_impl(adjoint, ndim, t, nullptr, nbatch, alpha, src_dev, beta, dst_dev, ws, wsb, LI_VARIANT_AUTO, st);
        }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(ws);
    for (int k = 0; k < ndim; ++k) li_table_destroy(t[k]);
    if (rc == LI_OK && e != cudaSuccess) {
        li::set_error("on-the-fly separable apply failed: %s", cudaGetErrorString(e));
        rc = LI_ERR_CUDA;
    }
    return rc;
}

int li_sep2d_otf_apply(int adjoint, const li_kernel_desc *ker1, const void *x1_dev, int64_t n1,
                       const li_kernel_desc *ker2, const void *x2_dev, int64_t n2, int64_t m1, int64_t m2,
                       int64_t nbatch, double alpha, const void *src_dev, double beta, void *dst_dev, void *stream)
{
    const li_kernel_desc *kers[2] = {ker1, ker2};
    const void *xs[2] = {x1_dev, x2_dev};
    const int64_t n[2] = {n1, n2}, m[2] = {m1, m2};
    return li_sep_otf_apply(adjoint, 2, kers, xs, n, m, nbatch, alpha, src_dev, beta, dst_dev, stream);
}

}  // extern "C"
