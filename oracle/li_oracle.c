/*
 * li_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see li_oracle.h).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -fno-fast-math, OpenMP only
 * for the explicitly multi-threaded CPU-baseline entry points).
 */
#include "li_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* sin(x).  Julia's Base.sin is a pure-Julia port of the FreeBSD msun kernels
 * (k_sin.c, k_cos.c, e_rem_pio2.c, k_sindf.c, k_cosdf.c, e_rem_pio2f.c); it is
 * what InterpolationKernels' Lanczos kernel calls.  This is the same algorithm
 * with every a*b+c left un-fused (Julia's `muladd` only fuses on FMA hardware,
 * so this is the FMA-less evaluation).  Constants are given by their bit
 * patterns so no decimal parsing is involved.  Only the Cody-Waite branch of the
 * argument reduction is restated (|x| < 2^20*pi/2), far beyond the Lanczos
 * arguments |pi*x| <= pi*S/2. */
static double bits_f64(uint64_t u)
{
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
}
static uint64_t f64_bits(double d)
{
    uint64_t u;
    memcpy(&u, &d, sizeof u);
    return u;
}

#define DS1 bits_f64(0xBFC5555555555549ull) /* -1.66666666666666324348e-01 */
#define DS2 bits_f64(0x3F8111111110F8A6ull) /*  8.33333333332248946124e-03 */
#define DS3 bits_f64(0xBF2A01A019C161D5ull) /* -1.98412698298579493134e-04 */
#define DS4 bits_f64(0x3EC71DE357B1FE7Dull) /*  2.75573137070700676789e-06 */
#define DS5 bits_f64(0xBE5AE5E68A2B9CEBull) /* -2.50507602534068634195e-08 */
#define DS6 bits_f64(0x3DE5D93A5ACFD57Cull) /*  1.58969099521155010221e-10 */

#define DC1 bits_f64(0x3FA555555555554Cull) /*  4.16666666666666019037e-02 */
#define DC2 bits_f64(0xBF56C16C16C15177ull) /* -1.38888888888741095749e-03 */
#define DC3 bits_f64(0x3EFA01A019CB1590ull) /*  2.48015872894767294178e-05 */
#define DC4 bits_f64(0xBE927E4F809C52ADull) /* -2.75573143513906633035e-07 */
#define DC5 bits_f64(0x3E21EE9EBDB4B1C4ull) /*  2.08757232129817482790e-09 */
#define DC6 bits_f64(0xBDA8FAE9BE8838D4ull) /* -1.13596475577881948265e-11 */

#define INVPIO2 bits_f64(0x3FE45F306DC9C883ull) /* 6.36619772367581382433e-01 */
#define PIO2_1 bits_f64(0x3FF921FB54400000ull)  /* first 33 bits of pi/2 */
#define PIO2_1T bits_f64(0x3DD0B4611A626331ull) /* pi/2 - PIO2_1 */
#define PIO2_2 bits_f64(0x3DD0B4611A600000ull)  /* second 33 bits of pi/2 */
#define PIO2_2T bits_f64(0x3BA3198A2E037073ull) /* pi/2 - (PIO2_1 + PIO2_2) */
#define PIO2_3 bits_f64(0x3BA3198A2E000000ull)  /* third 33 bits of pi/2 */
#define PIO2_3T bits_f64(0x397B839A252049C1ull) /* pi/2 - (PIO2_1 + PIO2_2 + PIO2_3) */

static double ksin(double x, double y, int iy)
{
    double z = x * x;
    double w = z * z;
    double r = (DS2 + z * (DS3 + z * DS4)) + (z * w) * (DS5 + z * DS6);
    double v = z * x;
    if (!iy) return x + v * (DS1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * DS1);
}

static double kcos(double x, double y)
{
    double z = x * x;
    double w = z * z;
    double r = z * (DC1 + z * (DC2 + z * DC3)) + (w * w) * (DC4 + z * (DC5 + z * DC6));
    double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

/* Cody-Waite reduction with up to three 33-bit pieces of pi/2 (e_rem_pio2.c, "medium"). */
static int rem_pio2_medium(double x, double *y0, double *y1)
{
    double fn = rint(x * INVPIO2);
    int n = (int)fn;
    double r = x - fn * PIO2_1;
    double w = fn * PIO2_1T;
    int ex = (int)((f64_bits(x) >> 52) & 0x7ff);
    double y = r - w;
    int ey = (int)((f64_bits(y) >> 52) & 0x7ff);
    if (ex - ey > 16) { /* second iteration, good to 118 bits */
        double t = r;
        w = fn * PIO2_2;
        r = t - w;
        w = fn * PIO2_2T - ((t - r) - w);
        y = r - w;
        ey = (int)((f64_bits(y) >> 52) & 0x7ff);
        if (ex - ey > 49) { /* third iteration, 151 bits */
            t = r;
            w = fn * PIO2_3;
            r = t - w;
            w = fn * PIO2_3T - ((t - r) - w);
            y = r - w;
        }
    }
    *y0 = y;
    *y1 = (r - y) - w;
    return n;
}

double lio_sin_f64(double x)
{
    double ax = fabs(x);
    if (ax < 0.78539816339744830962 /* Float64(pi)/4 */) {
        if (ax < 1.4901161193847656e-08 /* sqrt(eps(Float64)) */) return x;
        return ksin(x, 0.0, 0);
    }
    if (!(ax < INFINITY)) return x - x; /* NaN for NaN/Inf */
    double y0, y1;
    int n = rem_pio2_medium(x, &y0, &y1) & 3;
    switch (n) {
    case 0:
        return ksin(y0, y1, 1);
    case 1:
        return kcos(y0, y1);
    case 2:
        return -ksin(y0, y1, 1);
    default:
        return -kcos(y0, y1);
    }
}

/* cos(x): the same reduction, the kernels of the next quadrant (msun s_cos.c). */
double lio_cos_f64(double x)
{
    double ax = fabs(x);
    if (ax < 0.78539816339744830962) return kcos(x, 0.0);
    if (!(ax < INFINITY)) return x - x;
    double y0, y1;
    int n = rem_pio2_medium(x, &y0, &y1) & 3;
    switch (n) {
    case 0:
        return kcos(y0, y1);
    case 1:
        return -ksin(y0, y1, 1);
    case 2:
        return -kcos(y0, y1);
    default:
        return ksin(y0, y1, 1);
    }
}

/* Float32: everything is evaluated in Float64 and rounded once (k_sindf/k_cosdf). */
#define FS1 bits_f64(0xBFC5555554CBAC77ull) /* -0x15555554cbac77.0p-55 */
#define FS2 bits_f64(0x3F811110896EFBB2ull) /*  0x111110896efbb2.0p-59 */
#define FS3 bits_f64(0xBF2A00F9E2CAE774ull) /* -0x1a00f9e2cae774.0p-65 */
#define FS4 bits_f64(0x3EC6CD878C3B46A7ull) /*  0x16cd878c3b46a7.0p-71 */
#define FC0 bits_f64(0xBFDFFFFFFD0C5E81ull) /* -0x1ffffffd0c5e81.0p-54 */
#define FC1 bits_f64(0x3FA55553E1053A42ull) /*  0x155553e1053a42.0p-57 */
#define FC2 bits_f64(0xBF56C087E80F1E27ull) /* -0x16c087e80f1e27.0p-62 */
#define FC3 bits_f64(0x3EF99342E0EE5069ull) /*  0x199342e0ee5069.0p-68 */
#define FPIO2_1 bits_f64(0x3FF921FB50000000ull)  /* first 25 bits of pi/2 */
#define FPIO2_1T bits_f64(0x3E5110B4611A6263ull) /* pi/2 - FPIO2_1 */

static float ksindf(double x)
{
    double z = x * x;
    double w = z * z;
    double r = FS3 + z * FS4;
    double s = z * x;
    return (float)((x + s * (FS1 + z * FS2)) + (s * w) * r);
}

static float kcosdf(double x)
{
    double z = x * x;
    double w = z * z;
    double r = FC2 + z * FC3;
    return (float)(((1.0 + z * FC0) + w * FC1) + (w * z) * r);
}

float lio_sin_f32(float x)
{
    float ax = fabsf(x);
    if (ax < 0.78539818525314331f /* Float32(pi)/4 */) {
        if (ax < 3.4526698300124393e-04f /* sqrt(eps(Float32)) */) return x;
        return ksindf((double)x);
    }
    if (!(ax < INFINITY)) return x - x;
    double xd = (double)x;
    double fn = rint(xd * INVPIO2);
    int n = (int)fn & 3;
    double y = (xd - fn * FPIO2_1) - fn * FPIO2_1T;
    switch (n) {
    case 0:
        return ksindf(y);
    case 1:
        return kcosdf(y);
    case 2:
        return -ksindf(y);
    default:
        return -kcosdf(y);
    }
}

float lio_cos_f32(float x)
{
    float ax = fabsf(x);
    if (ax < 0.78539818525314331f) return kcosdf((double)x);
    if (!(ax < INFINITY)) return x - x;
    double xd = (double)x;
    double fn = rint(xd * INVPIO2);
    int n = (int)fn & 3;
    double y = (xd - fn * FPIO2_1) - fn * FPIO2_1T;
    switch (n) {
    case 0:
        return kcosdf(y);
    case 1:
        return -ksindf(y);
    case 2:
        return -kcosdf(y);
    default:
        return ksindf(y);
    }
}

/* Error-scale helpers for the tests (tests/util.py): the on-the-fly and affine loops can be run with every weight
 * replaced by its magnitude (mode 1: "the same operator applied with |w| to |src|", the per-element error bound of
 * SURVEY.md Appendix F) or by 1 where it is non-zero (mode 2: the number of products summed into an element). */
int lio_weight_mode_ = 0;
void lio_set_weight_mode(int mode) { lio_weight_mode_ = mode; }

int lio_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
#define LIO_CAT_(a, b) a##_##b
#define LIO_CAT(a, b) LIO_CAT_(a, b)

#define T float
#define SFX f32
#define FN(name) LIO_CAT(name, f32)
#define NEXTAFTER nextafterf
#define FABS fabsf
#define RINT rintf
#define FLOOR floorf
#define SIN lio_sin_f32
#define COS lio_cos_f32
#include "li_oracle_impl.inc"
#undef T
#undef SFX
#undef FN
#undef NEXTAFTER
#undef FABS
#undef RINT
#undef FLOOR
#undef SIN
#undef COS

#define T double
#define SFX f64
#define FN(name) LIO_CAT(name, f64)
#define NEXTAFTER nextafter
#define FABS fabs
#define RINT rint
#define FLOOR floor
#define SIN lio_sin_f64
#define COS lio_cos_f64
#include "li_oracle_impl.inc"
