/*
 * li_oracle.h -- CPU oracle for the interpolation hot path of LinearInterpolators.jl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * `linearinterpolators.jl_b200/`, `libli_b200.so`) may include, link or call
 * anything in this directory; only `tests/`, `__graft_entry__.smoke()` and the
 * `cpu_baseline` / `--impl reference` legs of `bench.py` do.
 *
 * PARITY STATUS
 *   - Index tables (J), boundary handling, loop nests and alpha/beta algebra
 *     restate in-tree reference code (file:line cited on every function) and
 *     are pinned by the known answers of the reference's test/runtests.jl
 *     (restated in tests/test_oracle_reference_suite.py).
 *   - Weights: "parity unpinned".  `getweights` lives in the un-vendored
 *     dependency InterpolationKernels.jl (compat 0.1.2, i.e. [0.1.2, 0.2); no
 *     Manifest, so no exact pin) and Julia is not available in this image.
 *     The formulas below restate that package's published algorithm; they are
 *     pinned only to the tolerances of the reference's own tests, not to the
 *     last bit of a real Julia run.
 *
 * All functions use 1-based Int64 indices in their tables, exactly like the
 * reference (`J::Matrix{Int}`, src/tabulated.jl:32; `J::Vector{Int}`,
 * src/sparse.jl:44).  Arrays are column-major.  The code must be compiled with
 * -ffp-contract=off: Julia never contracts a*b+c.
 */
#ifndef LI_ORACLE_H
#define LI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel families of InterpolationKernels 0.1.x (re-exported by
 * src/LinearInterpolators.jl:19-52). */
enum {
    LIO_RECTANGULAR = 0,       /* S = 1 */
    LIO_LINEAR = 1,            /* S = 2 */
    LIO_QUADRATIC = 2,         /* S = 3 */
    LIO_CUBIC = 3,             /* S = 4, cubic B-spline */
    LIO_CATMULLROM = 4,        /* S = 4 */
    LIO_CARDINALCUBIC = 5,     /* S = 4, p[0] = c */
    LIO_KEYS = 6,              /* S = 4, p[0] = a */
    LIO_MITCHELLNETRAVALI = 7, /* S = 4, p[0] = b, p[1] = c */
    LIO_LANCZOS = 8            /* S even */
};

enum { LIO_FLAT = 0, LIO_SAFEFLAT = 1 };

enum { LIO_OK = 0, LIO_ERR_ARG = 1, LIO_ERR_INEXACT = 3 };

#define LIO_MAX_S 16

typedef struct {
    int32_t family;
    int32_t S;
    int32_t boundary;
    int32_t prime; /* != 0: the derivative kernel (`*Prime`, src/LinearInterpolators.jl:23-41): getweights yields d/dt */
    double p[2];
} lio_kernel;

/* sin() restated from the msun algorithm Julia's Base.sin follows, without FMA. */
double lio_sin_f64(double x);
float lio_sin_f32(float x);
double lio_cos_f64(double x);
float lio_cos_f32(float x);

#define LIO_DECLARE(T, SFX)                                                                       \
    void lio_limits_##SFX(const lio_kernel *k, int64_t len, T *inf, T *sup);                      \
    int lio_getweights_##SFX(const lio_kernel *k, T t, T *w);                                     \
    int lio_getcoefs_##SFX(const lio_kernel *k, int64_t len, T x, int64_t *j, T *w);              \
    int lio_AtWA_##SFX(int S, int64_t nrows, const int64_t *J, const T *C, int64_t ncols, const T *w, T *dst); \
    int lio_regularize_##SFX(T *A, int64_t n, T eps, T mu);                                       \
    int lio_maketable_##SFX(const lio_kernel *k, const T *x, int64_t nrows, int64_t ncols,        \
                            int64_t *J, T *W);                                                    \
    void lio_fractional_index_##SFX(const T *pos, int64_t n, T first, T step, T last,             \
                                    int64_t len, T *out);                                         \
    void lio_tab_direct_##SFX(int S, int64_t nrows, const int64_t *J, const T *W, int64_t npre,   \
                              int64_t ncols, int64_t npost, T alpha, const T *src, T beta,        \
                              T *dst, int nthreads);                                              \
    void lio_tab_adjoint_##SFX(int S, int64_t nrows, const int64_t *J, const T *W, int64_t npre,  \
                               int64_t ncols, int64_t npost, T alpha, const T *src, T beta,       \
                               T *dst, int skip_zero, int nthreads);                              \
    int lio_otf1d_direct_##SFX(const lio_kernel *k, const T *x, int64_t n, int64_t len, T alpha,  \
                               const T *src, T beta, T *dst);                                     \
    int lio_otf1d_adjoint_##SFX(const lio_kernel *k, const T *x, int64_t n, int64_t len, T alpha, \
                                const T *src, T beta, T *dst);                                    \
    int lio_sep2d_direct_##SFX(const lio_kernel *k1, const T *x1, int64_t n1, const lio_kernel *k2, \
                               const T *x2, int64_t n2, int64_t m1, int64_t m2, T alpha,          \
                               const T *src, T beta, T *dst);                                     \
    int lio_sep2d_adjoint_##SFX(const lio_kernel *k1, const T *x1, int64_t n1,                    \
                                const lio_kernel *k2, const T *x2, int64_t n2, int64_t m1,        \
                                int64_t m2, T alpha, const T *src, T beta, T *dst);               \
    int lio_affine2d_direct_##SFX(const lio_kernel *k1, const lio_kernel *k2, const T *R,         \
                                  int64_t n1, int64_t n2, int64_t m1, int64_t m2, T alpha,        \
                                  const T *src, T beta, T *dst, int nthreads);                    \
    int lio_affine2d_adjoint_##SFX(const lio_kernel *k1, const lio_kernel *k2, const T *R,        \
                                   int64_t n1, int64_t n2, int64_t m1, int64_t m2, T alpha,       \
                                   const T *src, T beta, T *dst);

LIO_DECLARE(float, f32)
LIO_DECLARE(double, f64)

int lio_max_threads(void);
void lio_set_weight_mode(int mode);   /* test helper: 0 = weights as they are, 1 = |w|, 2 = (w != 0) */
extern int lio_weight_mode_;

#ifdef __cplusplus
}
#endif
#endif
