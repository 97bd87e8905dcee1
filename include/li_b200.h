/*
 * li_b200.h -- C ABI of libli_b200.so: the B200 (sm_100a) implementation of the
 * interpolation hot path of emmt/LinearInterpolators.jl.
 *
 * This is the drop-in boundary.  The reference is pure Julia and crosses no FFI
 * today; a maintainer replaces the *bodies* of the Julia methods cited below by
 * `ccall`s into these entry points (see INTEGRATION.md for the stubs), keeping the
 * LazyAlgebra signatures `vcreate(P, A, x, scratch)` and
 * `apply!(alpha, P, A, x, scratch, beta, y)` untouched.  Citations are
 * file:line into the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream);
 *   - arrays are column-major (Julia layout); `npre`/`npost` are the products of
 *     the dimensions before/after the interpolated one (CartesianIndices(predims)
 *     / (postdims) in src/tabulated.jl:355-357);
 *   - every function returns an li_status; li_last_error() gives the message of
 *     the last failure on the calling thread;
 *   - `*_dev` pointers are device pointers, `*_host` pointers are host pointers
 *     (pinned memory from li_host_alloc makes the copies asynchronous);
 *   - apply entry points are asynchronous on `stream` unless stated otherwise.
 *     There is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with LI_ERR_CUDA.
 */
#ifndef LI_B200_H
#define LI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LI_B200_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define LI_API __attribute__((visibility("default")))
#else
#define LI_API
#endif

typedef enum {
    LI_OK = 0,
    LI_ERR_ARG = 1,         /* ErrorException / ArgumentError / AssertionError in the reference */
    LI_ERR_DIM = 2,         /* DimensionMismatch (src/tabulated.jl:198, src/sparse.jl:624-651) */
    LI_ERR_INEXACT = 3,     /* InexactError of trunc(Int, f) (src/boundaries.jl:81,105,147) */
    LI_ERR_BOUNDS = 4,      /* "out of bound index in interpolation table" (src/tabulated.jl:212) */
    LI_ERR_UNSUPPORTED = 5, /* kernel family / size not built */
    LI_ERR_CUDA = 6,
    LI_ERR_NOMEM = 7,
    LI_ERR_NCCL = 8         /* NCCL missing (dlopen of libnccl.so.2 failed) or a collective failed */
} li_status;

typedef enum { LI_F32 = 0, LI_F64 = 1 } li_dtype;

/* InterpolationKernels 0.1.x kernel families (src/LinearInterpolators.jl:19-52). */
typedef enum {
    LI_RECTANGULAR = 0,       /* RectangularSpline,       S = 1 */
    LI_LINEAR = 1,            /* LinearSpline,            S = 2 */
    LI_QUADRATIC = 2,         /* QuadraticSpline,         S = 3 */
    LI_CUBIC = 3,             /* CubicSpline (B-spline),  S = 4 */
    LI_CATMULLROM = 4,        /* CatmullRomSpline,        S = 4 */
    LI_CARDINALCUBIC = 5,     /* CardinalCubicSpline(c),  S = 4, params[0] = c */
    LI_KEYS = 6,              /* KeysSpline(a),           S = 4, params[0] = a */
    LI_MITCHELLNETRAVALI = 7, /* MitchellNetravaliSpline(b, c), S = 4, params[0..1] */
    LI_LANCZOS = 8            /* LanczosKernel(S), S even, 2 <= S <= 16 */
} li_family;
/* family | LI_PRIME: the derivative kernel (RectangularSplinePrime ... LanczosKernelPrime,
 * src/LinearInterpolators.jl:23-41): getweights yields d/dt of the weights of the base kernel, everything else
 * (window, Flat / SafeFlat clamping of src/boundaries.jl:77-156, which does not look at the kernel) is unchanged. */
#define LI_PRIME 16

/* Boundary tag B of Kernel{T,S,B}; selects FlatLimits / SafeFlatLimits
 * (src/types.jl:27-52, src/boundaries.jl:37-42). */
typedef enum { LI_FLAT = 0, LI_SAFEFLAT = 1 } li_boundary;

/* Adjoint variants. */
typedef enum {
    LI_VARIANT_AUTO = 0,   /* deterministic gather */
    LI_VARIANT_GATHER = 1, /* deterministic: transposed table, ascending (row, tap) order */
    LI_VARIANT_ATOMIC = 2  /* red.global.add scatter; not run-to-run deterministic */
} li_variant;

#define LI_MAX_S 16

/* Value description of a Kernel{T,S,B} (the kernel types themselves stay in Julia). */
typedef struct {
    int32_t family;   /* li_family */
    int32_t S;        /* kernel size */
    int32_t boundary; /* li_boundary */
    int32_t dtype;    /* li_dtype: the kernel's floating-point type T */
    double params[4];
} li_kernel_desc;

typedef struct li_table li_table; /* opaque device-resident (J, W) table */

typedef struct {
    int32_t dtype, S;
    int64_t nrows, ncols;
    int32_t windowed; /* every row is clamp(base + s): eligible for the staged/fused kernels */
    int32_t monotone; /* window bases are non-decreasing */
} li_table_info;

/* ---- library / device ---------------------------------------------------- */
LI_API int li_version(void);
LI_API const char *li_last_error(void);
LI_API const char *li_status_string(int status);
LI_API int li_device_count(int *count);
LI_API int li_set_device(int device);
LI_API int li_get_device(int *device);
LI_API int li_num_sms(int *count);

/* ---- memory (replaces Array{T}(undef, dims) of vcreate when data lives on the
 *      device: src/tabulated.jl:305-327, src/sparse.jl:187-199,619-630,
 *      src/nonseparable.jl:65-75) ------------------------------------------- */
LI_API int li_malloc(void **dev_ptr, size_t bytes);
LI_API int li_free(void *dev_ptr);
LI_API int li_host_alloc(void **host_ptr, size_t bytes); /* pinned */
LI_API int li_host_free(void *host_ptr);
LI_API int li_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
LI_API int li_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
LI_API int li_memcpy_d2d(void *dst_dev, const void *src_dev, size_t bytes, void *stream);
LI_API int li_memset_zero(void *dst_dev, size_t bytes, void *stream);
LI_API int li_stream_create(void **stream);  /* a non-blocking cudaStream_t, for callers without a CUDA binding */
LI_API int li_stream_destroy(void *stream);
LI_API int li_stream_sync(void *stream);
/* vscale!(y, beta) of LazyAlgebra (call sites src/tabulated.jl:353,386,
 * src/sparse.jl:211,255,656,699, ...): beta == 0 fills zeros without reading. */
LI_API int li_vscale(int dtype, void *y_dev, int64_t n, double beta, void *stream);

/* ---- tables --------------------------------------------------------------
 * li_table_create: positions -> (J, W) on the device with Flat / SafeFlat
 * clamping.  Replaces __maketables (src/tabulated.jl:137-189) and _sparsecoefs
 * (src/sparse.jl:140-164), i.e. getcoefs (src/boundaries.jl:77-156) looped over
 * the rows.  `pos` holds nrows positions already converted to the kernel type T
 * (convert(T, X[i]), src/tabulated.jl:153) and may be a host or device pointer.
 * Synchronous (it reports LI_ERR_INEXACT). */
LI_API int li_table_create(const li_kernel_desc *ker, const void *pos, int pos_on_device, int64_t nrows,
                    int64_t ncols, void *stream, li_table **out);
/* li_table_import: adopt tables computed elsewhere (custom kernels evaluated in
 * Julia).  J is Int64, 1-based, S x nrows column-major exactly like A.J
 * (src/tabulated.jl:32) / A.J of the sparse operators (src/sparse.jl:44,522);
 * W like A.W / A.C.  Host pointers.  Performs the bound check of __fullcheck
 * (src/tabulated.jl:206-215) / _check (src/sparse.jl:166-185,802-817) once. */
LI_API int li_table_import(int dtype, int S, int64_t nrows, int64_t ncols, const int64_t *J_host,
                    const void *W_host, void *stream, li_table **out);
/* li_table_export: materialise coefficients(A) / columns(A) (src/tabulated.jl:43-44,
 * src/sparse.jl:67-68,527-528) on the host in the reference's layout. */
LI_API int li_table_export(const li_table *t, int64_t *J_host, void *W_host, void *stream);
LI_API int li_table_get_info(const li_table *t, li_table_info *info);
LI_API int li_table_destroy(li_table *t);

/* ---- table-backed operators ------------------------------------------------
 * dst <- alpha * P(A) * src + beta * dst along one dimension.
 * Direct  : src is [npre, ncols, npost], dst is [npre, nrows, npost]
 *           __direct! (src/tabulated.jl:394-439), SparseInterpolator Direct
 *           (src/sparse.jl:201-244, npre = npost = 1), _apply_direct!
 *           (src/sparse.jl:719-774).
 * Adjoint : src is [npre, nrows, npost], dst is [npre, ncols, npost]
 *           __adjoint! (src/tabulated.jl:441-478), src/sparse.jl:246-275,776-800.
 * alpha == 0 only scales dst by beta; beta == 0 never reads dst. */
LI_API int li_tab_apply(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                 const void *src_dev, double beta, void *dst_dev, int variant, void *stream);

/* Separable N-d operator H = A_1 * A_2 * ... * A_ndim (A_k interpolates dimension
 * k), the LazyAlgebra composition of test/benchmarks.jl:27-29,44-49.  A NULL entry
 * in `tabs` leaves that dimension untouched (its length is dims_in[k]).
 * Direct : src [ncols_1..ncols_n, nbatch] -> dst [nrows_1..nrows_n, nbatch];
 * Adjoint: the reverse.  Two tables on dimensions 1-2 (any pair of kernel sizes of
 * {2,3,4,6,8}^2; plus untouched / batch dimensions) and three tables on dimensions
 * 1-3 run as ONE fused kernel (one pass over HBM, no workspace: `workspace_dev` may
 * be NULL -- the call fails with LI_ERR_ARG "workspace too small" only if the fused
 * kernels decline the problem) when the tables are windowed and monotone; otherwise
 * per-dimension passes through `workspace`.  `dims_other[k]` is only read for NULL
 * tables. */
LI_API int li_sep_workspace_bytes(int adjoint, int ndim, const li_table *const *tabs,
                           const int64_t *dims_other, int64_t nbatch, size_t *bytes);
LI_API int li_sep_apply(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other,
                 int64_t nbatch, double alpha, const void *src_dev, double beta, void *dst_dev,
                 void *workspace_dev, size_t workspace_bytes, int variant, void *stream);

/* n independent separable applies in one call (one launch sequence, no per-call host overhead): apply k uses
 * tabs[k*ndim .. k*ndim+ndim-1], src_dev[k] and dst_dev[k].  Several applies may share one dst (beta = 1): they are
 * executed in order on `stream`, which is how `m = sum_k A_k' y_k` of iterative fitting (README: "A'*y") is
 * accumulated.  Intermediate workspace, if any apply needs it, is managed by the library.  A long batch of fused 2-D
 * applies into ONE dst with beta = 1 is dealt round-robin to a few internal streams with partial accumulators that
 * are summed in a fixed order at the end (deterministic; the sum is associated differently from the in-order one). */
LI_API int li_sep_apply_batch(int adjoint, int ndim, int64_t n, const li_table *const *tabs, const int64_t *dims_other,
                              int64_t nbatch, double alpha, const void *const *src_dev, double beta,
                              void *const *dst_dev, int variant, void *stream);

/* ---- operators with on-the-fly coefficients --------------------------------
 * 1-D: apply!(alpha, P, ker, x, src, beta, dst), src/unidimensional.jl:95-245.
 * Direct: src[len] -> dst[n]; Adjoint: src[n] -> dst[len].  x_dev: n positions of type T.
 * Synchronous (reports LI_ERR_INEXACT). */
LI_API int li_otf1d_apply(int adjoint, const li_kernel_desc *ker, const void *x_dev, int64_t n,
                   int64_t len, double alpha, const void *src_dev, double beta, void *dst_dev,
                   int variant, void *stream);
/* 2-D separable: apply!(alpha, P, ker1, x1, ker2, x2, src, beta, dst),
 * src/separable.jl:56-162.  Direct: src[m1,m2,nbatch] -> dst[n1,n2,nbatch].
 * Coefficients are generated on the device per call (no table survives the call).
 * Synchronous. */
/* N-d separable with on-the-fly coefficients, ndim = 1..3 (the generalisation the reference plans, src/separable.jl:14):
 * one kernel and one position vector per dimension; tables are generated on the device for this call, then the fused
 * 2-D / 3-D kernels run.  Direct: src[m_1..m_ndim, nbatch] -> dst[n_1..n_ndim, nbatch].  Synchronous. */
LI_API int li_sep_otf_apply(int adjoint, int ndim, const li_kernel_desc *const *kers, const void *const *x_dev,
                            const int64_t *n, const int64_t *m, int64_t nbatch, double alpha, const void *src_dev,
                            double beta, void *dst_dev, void *stream);
LI_API int li_sep2d_otf_apply(int adjoint, const li_kernel_desc *ker1, const void *x1_dev, int64_t n1,
                       const li_kernel_desc *ker2, const void *x2_dev, int64_t n2, int64_t m1,
                       int64_t m2, int64_t nbatch, double alpha, const void *src_dev, double beta,
                       void *dst_dev, void *stream);
/* 2-D affine: apply!(alpha, P, ker1, ker2, R, src, beta, dst),
 * src/nonseparable.jl:147-264, and the Mapping methods :77-101.
 * R = (xx, xy, x, yx, yy, y) of AffineTransform2D; rows = (n1, n2) is the output
 * size of the direct operator, cols = (m1, m2) its input size; nbatch images are
 * contiguous along a third dimension.  Synchronous unless flags & LI_FLAG_ASYNC. */
#define LI_FLAG_ASYNC 1 /* skip the InexactError read-back; errors surface at the next sync call */
LI_API int li_affine2d_apply(int adjoint, const li_kernel_desc *ker1, const li_kernel_desc *ker2,
                      const double R[6], int64_t n1, int64_t n2, int64_t m1, int64_t m2,
                      int64_t nbatch, double alpha, const void *src_dev, double beta,
                      void *dst_dev, int variant, int flags, void *stream);

/* ---- host-buffer front ends (Julia `Array` arguments) ----------------------
 * Same semantics with HOST src/dst: stage through the device (H2D, kernel, D2H)
 * on `stream` and synchronise.  Correct but PCIe-bound; still no CPU compute. */
LI_API int li_tab_apply_host(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                      const void *src_host, double beta, void *dst_host, int variant, void *stream);
LI_API int li_sep_apply_host(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other,
                      int64_t nbatch, double alpha, const void *src_host, double beta,
                      void *dst_host, int variant, void *stream);

/* Asynchronous variants: enqueue H2D copy, kernels and D2H copy on `stream` and return; the host buffers must be
 * page-locked (li_host_alloc) and stay valid until li_stream_sync(stream).  Device staging buffers are kept per
 * stream, so calls on different streams overlap (e.g. the D2H of a forward with the H2D of an adjoint). */
LI_API int li_tab_apply_host_async(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                                   const void *src_host, double beta, void *dst_host, int variant, void *stream);
LI_API int li_sep_apply_host_async(int adjoint, int ndim, const li_table *const *tabs, const int64_t *dims_other,
                                   int64_t nbatch, double alpha, const void *src_host, double beta,
                                   void *dst_host, int variant, void *stream);

/* ---- many frames accumulated into one model; the all-reduce over the GPUs of a node ----------------------------
 * Iterative fitting (README.md:36-37) evaluates  model <- beta*model + alpha * sum_k H_k' y_k  with H_k = A1_k * A2_k
 * the 2-D separable operator of frame k (composition of test/benchmarks.jl:27-29; adjoint loops
 * src/tabulated.jl:441-478).  The reference runs `apply!(alpha, Adjoint, H_k, y_k, false, 1, model)` frame after
 * frame on one core; here the local frames go through ONE multi-frame kernel and, with a communicator, the frames
 * are sharded over the ranks (one rank per GPU) and the rank-private sums are combined with an NCCL all-reduce
 * over NVLink -- the one exchange step of the path.  NCCL is loaded with dlopen at the first li_comm_* call. */
typedef struct li_comm li_comm;
#define LI_COMM_ID_BYTES 128
/* rank 0 creates an id and hands it to the other ranks by any out-of-band means (MPI, a file, torch.distributed) */
LI_API int li_comm_unique_id(void *id, size_t bytes);
/* one rank per GPU (one process, or one host thread, per GPU); binds to the CURRENT device; collective call */
LI_API int li_comm_init_rank(const void *id, int nranks, int rank, li_comm **out);
/* one process driving `ndev` GPUs (devices == NULL: 0..ndev-1): out[k] is the communicator of devices[k].
 * Use one host thread per device for the calls below, or bracket them with li_comm_group_start/end
 * (li_allreduce_sum only: li_adjoint_accumulate_allreduce must not be called inside a group). */
LI_API int li_comm_init_all(int ndev, const int *devices, li_comm **out);
LI_API int li_comm_info(const li_comm *comm, int *rank, int *nranks, int *device);
LI_API int li_comm_destroy(li_comm *comm);
LI_API int li_comm_group_start(void);
LI_API int li_comm_group_end(void);
LI_API int li_nccl_version(int *version);
/* buf <- sum over the ranks of buf (in place, `count` elements); comm == NULL or a single rank: no-op */
LI_API int li_allreduce_sum(li_comm *comm, int dtype, void *buf_dev, int64_t count, void *stream);
/* model[m1, m2] <- beta*model + alpha * sum over ALL ranks of sum over the nframes LOCAL frames of H_k' y_k.
 * tabs[2k], tabs[2k+1] are the dim-1 / dim-2 tables of local frame k (ncols = m1 / m2), src_dev[k] its data
 * [nrows_1, nrows_2].  comm == NULL: this GPU only.  A rank may own zero frames.  The model is produced in `nbands`
 * bands of rows; the all-reduce of a band overlaps the computation of the next one (nbands <= 16; 1 = a single
 * all-reduce at the end -- with one process per GPU the library's own two-shot all-reduce over NVLink peer memory (CUDA
 * IPC exchange buffers, rank-ordered sums: the same bits on every rank), else ncclAllReduce: the measured optimum on a node of B200s for config #5 -- every band is a launch over all the
 * local frames, and on 2 GPUs 2 / 4 bands took 3.2 / 2.9 ms against 2.1 ms with the 0.09 ms all-reduce exposed).  Deterministic for a given number of ranks, bands and frame order.  Asynchronous on
 * `stream` (the communicator's own stream is joined back into it before the call returns). */
LI_API int li_adjoint_accumulate_allreduce(li_comm *comm, int dtype, int64_t m1, int64_t m2, int64_t nframes,
                                           const li_table *const *tabs, const void *const *src_dev, double alpha,
                                           double beta, void *model_dev, int nbands, void *stream);

/* ---- vectors and the linear conjugate-gradient driver ------------------------------------------------------------
 * LazyAlgebra's vdot / vupdate! / vcombine! and conjgrad as used by `x = conjgrad(H'*H, H'*y; ftol, maxiter)`
 * (test/benchmarks.jl:31-40).  Dot products are accumulated in Float64 in a fixed order (deterministic). */
LI_API int li_vdot(int dtype, const void *x_dev, const void *y_dev, int64_t n, double *result_host, void *stream); /* blocking */
LI_API int li_vdot_async(int dtype, const void *x_dev, const void *y_dev, int64_t n, double *result_dev, void *stream);
LI_API int li_vupdate(int dtype, void *y_dev, double alpha, const void *x_dev, int64_t n, void *stream); /* y += alpha*x */
/* dst <- alpha*x + beta*y (dst may alias x or y; beta == 0 never reads y) */
LI_API int li_vcombine(int dtype, void *dst_dev, double alpha, const void *x_dev, double beta, const void *y_dev,
                       int64_t n, void *stream);
typedef int (*li_apply_fn)(void *ctx, const void *x_dev, void *y_dev, void *stream); /* y <- A*x, returns an li_status */
typedef struct {
    int32_t maxiter;     /* default 50 */
    int32_t check_every; /* the host reads the residual every so many iterations (default 5); no other round trip */
    double ftol;         /* stop when the decrease of the objective, alpha*rho/2, is <= ftol * its largest value */
    double gtol_abs, gtol_rel; /* stop when sqrt(rho) <= max(gtol_abs, gtol_rel*sqrt(rho0)), rho = |b - A*x|^2 */
} li_cg_options;
typedef enum { LI_CG_MAXITER = 0, LI_CG_FTEST = 1, LI_CG_GTEST = 2, LI_CG_NOT_POSITIVE = 3 } li_cg_status;
typedef struct {
    int32_t iters, status; /* li_cg_status */
    double rho0, rho;      /* squared residual norm at the start / at the last check */
} li_cg_info;
/* solves A*x = b for a symmetric positive (semi-)definite A given as a callback; x_is_zero != 0 starts from x = 0 */
LI_API int li_conjgrad(li_apply_fn A, void *ctx, int dtype, int64_t n, const void *b_dev, void *x_dev, int x_is_zero,
                       const li_cg_options *opts, li_cg_info *info, void *stream);
/* x <- argmin |H*x - y|^2 + mu*|x|^2 for the separable operator H = A_1*...*A_ndim of li_sep_apply: conjugate
 * gradients on (H'*H + mu*I) x = H'*y entirely inside the library (no callback into the host language) */
LI_API int li_conjgrad_sep(int ndim, const li_table *const *tabs, const int64_t *dims_other, int64_t nbatch, double mu,
                           const void *y_dev, void *x_dev, int x_is_zero, const li_cg_options *opts, li_cg_info *info,
                           void *stream);

/* ---- normal equations (src/sparse.jl:277-356, 424-490) ------------------------
 * li_table_atwa: dst[ncols, ncols] (column-major, device) <- A'*diag(w)*A, w[nrows] device weights or NULL for A'*A
 * (AtA! / AtWA!).  Every element is summed by one thread in the reference's (row, tap, tap) order: deterministic.
 * li_regularize: A <- A + rho*(eps*I + mu*D'*D), rho the largest diagonal element, D first-order differences
 * (regularize!); synchronises the stream to check rho > 0 (LI_ERR_ARG "we have a problem!" otherwise, as the reference). */
LI_API int li_table_atwa(const li_table *t, const void *w_dev, void *dst_dev, void *stream);
LI_API int li_regularize(int dtype, void *A_dev, int64_t n, double eps, double mu, void *stream);

/* ---- dispatch switches ------------------------------------------------------
 * Process-wide integer options, initialised once from the environment at load time (LI_B200_<NAME>): "no_fused",
 * "no_stream_dn", "no_bands", "affine_plain", "trace", "adj_debug", "no_multi", "exp".  They select slower cross-check
 * paths; the hot path never calls getenv(). */
LI_API int li_set_option(const char *name, int value);
LI_API int li_get_option(const char *name, int *value);

/* ---- counters --------------------------------------------------------------
 * Number of CUDA kernels this library has launched in this process (all entry
 * points); bench.py reports the difference over the timed region. */
LI_API int64_t li_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LI_B200_H */
