// li_weights.cuh -- device evaluation of interpolation coefficients.
//
// getcoefs (reference src/boundaries.jl:77-156) and the kernel weights of
// InterpolationKernels 0.1.x (call sites src/boundaries.jl:82,110,152), written so
// that index AND weight tables are reproducible to the last bit: every product/sum
// goes through a round-to-nearest intrinsic (__fmul_rn, __dadd_rn, ...), which the
// compiler never contracts into an FMA (Julia does not contract a*b+c either), no
// flush-to-zero, IEEE division, rint() for Julia's ties-to-even round().
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>

#include "li_b200.h"

namespace li {

template <typename T>
struct Ar;
template <>
struct Ar<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float floor_(float a) { return floorf(a); }
    static __device__ __forceinline__ float rint_(float a) { return rintf(a); }
    static __device__ __forceinline__ float abs_(float a) { return fabsf(a); }
};
template <>
struct Ar<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double floor_(double a) { return floor(a); }
    static __device__ __forceinline__ double rint_(double a) { return rint(a); }
    static __device__ __forceinline__ double abs_(double a) { return fabs(a); }
};

// ---------------------------------------------------------------------------
// sin(): the msun algorithm that Julia's Base.sin ports (k_sin/k_cos + Cody-Waite
// reduction; k_sindf/k_cosdf in double for Float32), evaluated without FMA.
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

__device__ inline double msun_ksin(double x, double y, bool have_y)
{
    const double S1 = __longlong_as_double(0xBFC5555555555549ull);
    const double S2 = __longlong_as_double(0x3F8111111110F8A6ull);
    const double S3 = __longlong_as_double(0xBF2A01A019C161D5ull);
    const double S4 = __longlong_as_double(0x3EC71DE357B1FE7Dull);
    const double S5 = __longlong_as_double(0xBE5AE5E68A2B9CEBull);
    const double S6 = __longlong_as_double(0x3DE5D93A5ACFD57Cull);
    double z = dmul(x, x);
    double w = dmul(z, z);
    double r = dadd(dadd(S2, dmul(z, dadd(S3, dmul(z, S4)))), dmul(dmul(z, w), dadd(S5, dmul(z, S6))));
    double v = dmul(z, x);
    if (!have_y) return dadd(x, dmul(v, dadd(S1, dmul(z, r))));
    return dsub(x, dsub(dsub(dmul(z, dsub(dmul(0.5, y), dmul(v, r))), y), dmul(v, S1)));
}

__device__ inline double msun_kcos(double x, double y)
{
    const double C1 = __longlong_as_double(0x3FA555555555554Cull);
    const double C2 = __longlong_as_double(0xBF56C16C16C15177ull);
    const double C3 = __longlong_as_double(0x3EFA01A019CB1590ull);
    const double C4 = __longlong_as_double(0xBE927E4F809C52ADull);
    const double C5 = __longlong_as_double(0x3E21EE9EBDB4B1C4ull);
    const double C6 = __longlong_as_double(0xBDA8FAE9BE8838D4ull);
    double z = dmul(x, x);
    double w = dmul(z, z);
    double r = dadd(dmul(z, dadd(C1, dmul(z, dadd(C2, dmul(z, C3))))),
                    dmul(dmul(w, w), dadd(C4, dmul(z, dadd(C5, dmul(z, C6))))));
    double hz = dmul(0.5, z);
    w = dsub(1.0, hz);
    return dadd(w, dadd(dsub(dsub(1.0, w), hz), dsub(dmul(z, r), dmul(x, y))));
}

__device__ inline int msun_exponent(double v)
{
    return (int)((__double_as_longlong(v) >> 52) & 0x7ff);
}

__device__ inline double li_sin(double x)
{
    const double INVPIO2 = __longlong_as_double(0x3FE45F306DC9C883ull);
    const double PIO2_1 = __longlong_as_double(0x3FF921FB54400000ull);
    const double PIO2_1T = __longlong_as_double(0x3DD0B4611A626331ull);
    const double PIO2_2 = __longlong_as_double(0x3DD0B4611A600000ull);
    const double PIO2_2T = __longlong_as_double(0x3BA3198A2E037073ull);
    const double PIO2_3 = __longlong_as_double(0x3BA3198A2E000000ull);
    const double PIO2_3T = __longlong_as_double(0x397B839A252049C1ull);
    double ax = fabs(x);
    if (ax < 0.78539816339744830962) {
        if (ax < 1.4901161193847656e-08) return x;
        return msun_ksin(x, 0.0, false);
    }
    if (!(ax < CUDART_INF)) return dsub(x, x);
    double fn = rint(dmul(x, INVPIO2));
    int n = (int)fn;
    double r = dsub(x, dmul(fn, PIO2_1));
    double w = dmul(fn, PIO2_1T);
    int ex = msun_exponent(x);
    double y = dsub(r, w);
    if (ex - msun_exponent(y) > 16) {
        double t = r;
        w = dmul(fn, PIO2_2);
        r = dsub(t, w);
        w = dsub(dmul(fn, PIO2_2T), dsub(dsub(t, r), w));
        y = dsub(r, w);
        if (ex - msun_exponent(y) > 49) {
            t = r;
            w = dmul(fn, PIO2_3);
            r = dsub(t, w);
            w = dsub(dmul(fn, PIO2_3T), dsub(dsub(t, r), w));
            y = dsub(r, w);
        }
    }
    double y1 = dsub(dsub(r, y), w);
    switch (n & 3) {
    case 0:
        return msun_ksin(y, y1, true);
    case 1:
        return msun_kcos(y, y1);
    case 2:
        return -msun_ksin(y, y1, true);
    default:
        return -msun_kcos(y, y1);
    }
}

__device__ inline float msun_ksindf(double x)
{
    const double S1 = __longlong_as_double(0xBFC5555554CBAC77ull);
    const double S2 = __longlong_as_double(0x3F811110896EFBB2ull);
    const double S3 = __longlong_as_double(0xBF2A00F9E2CAE774ull);
    const double S4 = __longlong_as_double(0x3EC6CD878C3B46A7ull);
    double z = dmul(x, x);
    double w = dmul(z, z);
    double r = dadd(S3, dmul(z, S4));
    double s = dmul(z, x);
    return __double2float_rn(dadd(dadd(x, dmul(s, dadd(S1, dmul(z, S2)))), dmul(dmul(s, w), r)));
}

__device__ inline float msun_kcosdf(double x)
{
    const double C0 = __longlong_as_double(0xBFDFFFFFFD0C5E81ull);
    const double C1 = __longlong_as_double(0x3FA55553E1053A42ull);
    const double C2 = __longlong_as_double(0xBF56C087E80F1E27ull);
    const double C3 = __longlong_as_double(0x3EF99342E0EE5069ull);
    double z = dmul(x, x);
    double w = dmul(z, z);
    double r = dadd(C2, dmul(z, C3));
    return __double2float_rn(dadd(dadd(dadd(1.0, dmul(z, C0)), dmul(w, C1)), dmul(dmul(w, z), r)));
}

__device__ inline float li_sin(float x)
{
    const double INVPIO2 = __longlong_as_double(0x3FE45F306DC9C883ull);
    const double PIO2_1 = __longlong_as_double(0x3FF921FB50000000ull);
    const double PIO2_1T = __longlong_as_double(0x3E5110B4611A6263ull);
    float ax = fabsf(x);
    if (ax < 0.78539818525314331f) {
        if (ax < 3.4526698300124393e-04f) return x;
        return msun_ksindf((double)x);
    }
    if (!(ax < CUDART_INF_F)) return __fsub_rn(x, x);
    double xd = (double)x;
    double fn = rint(dmul(xd, INVPIO2));
    int n = (int)fn & 3;
    double y = dsub(dsub(xd, dmul(fn, PIO2_1)), dmul(fn, PIO2_1T));
    switch (n) {
    case 0:
        return msun_ksindf(y);
    case 1:
        return msun_kcosdf(y);
    case 2:
        return -msun_ksindf(y);
    default:
        return -msun_kcosdf(y);
    }
}

// cos(): same reduction as li_sin, the kernels of the next quadrant (msun s_cos.c / s_cosf.c), without FMA.
__device__ inline double li_cos(double x)
{
    const double INVPIO2 = __longlong_as_double(0x3FE45F306DC9C883ull);
    const double PIO2_1 = __longlong_as_double(0x3FF921FB54400000ull);
    const double PIO2_1T = __longlong_as_double(0x3DD0B4611A626331ull);
    const double PIO2_2 = __longlong_as_double(0x3DD0B4611A600000ull);
    const double PIO2_2T = __longlong_as_double(0x3BA3198A2E037073ull);
    const double PIO2_3 = __longlong_as_double(0x3BA3198A2E000000ull);
    const double PIO2_3T = __longlong_as_double(0x397B839A252049C1ull);
    double ax = fabs(x);
    if (ax < 0.78539816339744830962) return msun_kcos(x, 0.0);
    if (!(ax < CUDART_INF)) return dsub(x, x);
    double fn = rint(dmul(x, INVPIO2));
    int n = (int)fn;
    double r = dsub(x, dmul(fn, PIO2_1));
    double w = dmul(fn, PIO2_1T);
    int ex = msun_exponent(x);
    double y = dsub(r, w);
    if (ex - msun_exponent(y) > 16) {
        double t = r;
        w = dmul(fn, PIO2_2);
        r = dsub(t, w);
        w = dsub(dmul(fn, PIO2_2T), dsub(dsub(t, r), w));
        y = dsub(r, w);
        if (ex - msun_exponent(y) > 49) {
            t = r;
            w = dmul(fn, PIO2_3);
            r = dsub(t, w);
            w = dsub(dmul(fn, PIO2_3T), dsub(dsub(t, r), w));
            y = dsub(r, w);
        }
    }
    double y1 = dsub(dsub(r, y), w);
    switch (n & 3) {
    case 0:
        return msun_kcos(y, y1);
    case 1:
        return -msun_ksin(y, y1, true);
    case 2:
        return -msun_kcos(y, y1);
    default:
        return msun_ksin(y, y1, true);
    }
}

__device__ inline float li_cos(float x)
{
    const double INVPIO2 = __longlong_as_double(0x3FE45F306DC9C883ull);
    const double PIO2_1 = __longlong_as_double(0x3FF921FB50000000ull);
    const double PIO2_1T = __longlong_as_double(0x3E5110B4611A6263ull);
    float ax = fabsf(x);
    if (ax < 0.78539818525314331f) return msun_kcosdf((double)x);
    if (!(ax < CUDART_INF_F)) return __fsub_rn(x, x);
    double xd = (double)x;
    double fn = rint(dmul(xd, INVPIO2));
    int n = (int)fn & 3;
    double y = dsub(dsub(xd, dmul(fn, PIO2_1)), dmul(fn, PIO2_1T));
    switch (n) {
    case 0:
        return msun_kcosdf(y);
    case 1:
        return -msun_ksindf(y);
    case 2:
        return -msun_kcosdf(y);
    default:
        return msun_ksindf(y);
    }
}

// ---------------------------------------------------------------------------
// Kernel description as the device sees it: every constant already rounded to T
// on the host (InterpolationKernels stores its constants as T fields).
template <typename T>
struct KerDev {
    int32_t family, S, boundary, deriv;   // deriv != 0: the derivative kernel (LI_PRIME)
    T c[8];      // family-specific constants
    T inf, sup;  // SafeFlatLimits (src/boundaries.jl:40-42)
};

// getweights(ker', t): the derivative kernels (`*Prime` of InterpolationKernels 0.1.x, re-exported at
// src/LinearInterpolators.jl:23-41): w'_s(t) = d/dt w_s(t), the weights being w_s(t) = ker(t - s + const).
template <typename T, int S>
__device__ __forceinline__ void getweights_prime(const KerDev<T> &k, T t, T (&w)[S])
{
    using A = Ar<T>;
    const T one = (T)1, two = (T)2, half = (T)0.5;
    if constexpr (S == 1) {
        w[0] = (T)0;
    } else if constexpr (S == 2) {
        w[0] = (T)-1;
        w[1] = one;
    } else if constexpr (S == 3) {
        w[0] = A::sub(t, half);
        w[1] = A::mul((T)-2, t);
        w[2] = A::add(t, half);
    } else {
        if (k.family == LI_LANCZOS) {
            // d/dx [b sin(pi x) sin(c x) / x^2] = b (pi cos(pi x) sin(c x) + c sin(pi x) cos(c x) - 2 sin(pi x) sin(c x) / x) / x^2
            constexpr int ctr = (S + 1) >> 1;
#pragma unroll
            for (int i = 1; i <= S; ++i) {
                T x = (i < ctr) ? A::add(t, (T)(ctr - i)) : (i == ctr ? t : A::sub(t, (T)(i - ctr)));
                T v;
                if (A::abs_(x) >= k.c[0] || x == (T)0) {
                    v = (T)0;
                } else {
                    const T px = A::mul(k.c[3], x), cx = A::mul(k.c[2], x);
                    const T s1 = li_sin(px), c1 = li_cos(px), s2 = li_sin(cx), c2 = li_cos(cx);
                    const T a1 = A::mul(A::mul(k.c[3], c1), s2);
                    const T a2 = A::mul(A::mul(k.c[2], s1), c2);
                    const T a3 = A::div(A::mul(A::mul(two, s1), s2), x);
                    v = A::div(A::mul(k.c[1], A::sub(A::add(a1, a2), a3)), A::mul(x, x));
                }
                w[i - 1] = v;
            }
        } else if constexpr (S == 4) {
            switch (k.family) {
            case LI_CUBIC: {
                T r = A::sub(one, t);
                w[0] = A::mul(A::mul((T)-0.5, r), r);
                w[1] = A::mul(A::sub(A::mul((T)1.5, t), two), t);
                w[2] = A::mul(A::sub(two, A::mul((T)1.5, r)), r);
                w[3] = A::mul(A::mul(half, t), t);
                break;
            }
            case LI_CATMULLROM: {
                w[0] = A::sub(A::mul(A::add(A::mul((T)-1.5, t), two), t), half);
                w[1] = A::mul(A::sub(A::mul((T)4.5, t), (T)5), t);
                w[2] = A::add(A::mul(A::add(A::mul((T)-4.5, t), (T)4), t), half);
                w[3] = A::mul(A::sub(A::mul((T)1.5, t), one), t);
                break;
            }
            case LI_CARDINALCUBIC: {
                const T p = k.c[0], q = k.c[1];
                T u = A::sub(one, t);
                T tu = A::mul(t, u);
                T umt = A::sub(u, t);
                T q1tu = A::mul(A::add(one, q), tu);
                w[0] = A::mul(p, A::mul(u, A::sub(u, A::add(t, t))));
                w[1] = A::sub(A::sub(A::mul(A::sub(u, A::mul(q, t)), umt), q1tu), one);
                w[2] = A::add(A::add(A::mul(A::sub(t, A::mul(q, u)), umt), q1tu), one);
                w[3] = A::mul(p, A::mul(t, A::sub(A::add(u, u), t)));
                break;
            }
            default: {
                // Keys / Mitchell-Netravali: P'(x) = (3 p3 x + 2 p2) x, Q'(x) = (3 q3 x + 2 q2) x + q1
                const T dp2 = A::mul(two, k.c[1]), dp3 = A::mul((T)3, k.c[2]);
                const T dq1 = k.c[4], dq2 = A::mul(two, k.c[5]), dq3 = A::mul((T)3, k.c[6]);
                auto P1 = [&](T x) { return A::mul(A::add(A::mul(dp3, x), dp2), x); };
                auto Q1 = [&](T x) { return A::add(A::mul(A::add(A::mul(dq3, x), dq2), x), dq1); };
                w[0] = Q1(A::add(t, one));
                w[1] = P1(t);
                w[2] = -P1(A::sub(one, t));
                w[3] = -Q1(A::sub(two, t));
                break;
            }
            }
        } else {
#pragma unroll
            for (int s = 0; s < S; ++s) w[s] = (T)0;  // unreachable: host validates (family, S)
        }
    }
}

// getweights(ker, t) for compile-time S (w fully in registers).
template <typename T, int S>
__device__ __forceinline__ void getweights(const KerDev<T> &k, T t, T (&w)[S])
{
    using A = Ar<T>;
    const T one = (T)1;
    if (k.deriv) {   // the derivative kernels: d/dt of the expressions below
        getweights_prime<T, S>(k, t, w);
        return;
    }
    if constexpr (S == 1) {
        w[0] = one;
    } else if constexpr (S == 2) {
        w[0] = A::sub(one, t);
        w[1] = t;
    } else if constexpr (S == 3) {
        // QuadraticSpline: c[0] = 1/sqrt(8), c[1] = 1/sqrt(2), c[2] = 3/4
        T c2t = A::mul(k.c[1], t);
        T q1 = A::sub(k.c[0], c2t);
        T q3 = A::add(k.c[0], c2t);
        w[0] = A::mul(q1, q1);
        w[1] = A::sub(k.c[2], A::mul(t, t));
        w[2] = A::mul(q3, q3);
    } else {
        if (k.family == LI_LANCZOS) {
            // c[0] = a = S/2, c[1] = b = a/pi^2, c[2] = c = pi/a, c[3] = T(pi)
            constexpr int ctr = (S + 1) >> 1;
#pragma unroll
            for (int i = 1; i <= S; ++i) {
                T x = (i < ctr) ? A::add(t, (T)(ctr - i)) : (i == ctr ? t : A::sub(t, (T)(i - ctr)));
                T v;
                if (A::abs_(x) >= k.c[0]) {
                    v = (T)0;
                } else if (x == (T)0) {
                    v = one;
                } else {
                    T num = A::mul(k.c[1], li_sin(A::mul(k.c[3], x)));
                    num = A::mul(num, li_sin(A::mul(k.c[2], x)));
                    v = A::div(num, A::mul(x, x));
                }
                w[i - 1] = v;
            }
        } else if constexpr (S == 4) {
            switch (k.family) {
            case LI_CUBIC: {
                // c[0] = 1/2, c[1] = 2/3, c[2] = 1/6
                T r = A::sub(one, t);
                T r2 = A::mul(r, r);
                T t2 = A::mul(t, t);
                w[0] = A::mul(A::mul(k.c[2], r2), r);
                w[1] = A::add(k.c[1], A::mul(A::sub(A::mul(k.c[0], t), one), t2));
                w[2] = A::add(k.c[1], A::mul(A::sub(A::mul(k.c[0], r), one), r2));
                w[3] = A::mul(A::mul(k.c[2], t2), t);
                break;
            }
            case LI_CATMULLROM: {
                T s = A::sub(one, t);
                T q = A::mul(A::mul((T)-0.5, t), s);
                T w1 = A::mul(q, s);
                T w4 = A::mul(q, t);
                T r = A::sub(w4, w1);
                w[0] = w1;
                w[1] = A::add(A::sub(s, w1), r);
                w[2] = A::sub(A::sub(t, w4), r);
                w[3] = w4;
                break;
            }
            case LI_CARDINALCUBIC: {
                // c[0] = p = (c-1)/2, c[1] = q = (c+1)/2
                T u = A::sub(one, t);
                T tu = A::mul(t, u);
                T ptu = A::mul(k.c[0], tu);
                w[0] = A::mul(ptu, u);
                w[1] = A::add(A::mul(A::sub(u, A::mul(k.c[1], t)), tu), u);
                w[2] = A::add(A::mul(A::sub(t, A::mul(k.c[1], u)), tu), t);
                w[3] = A::mul(ptu, t);
                break;
            }
            default: {
                // Keys / Mitchell-Netravali: c = (p0, p2, p3, q0, q1, q2, q3)
                const T p0 = k.c[0], p2 = k.c[1], p3 = k.c[2];
                const T q0 = k.c[3], q1 = k.c[4], q2 = k.c[5], q3 = k.c[6];
                auto P = [&](T x) {
                    return A::add(A::mul(A::mul(A::add(A::mul(p3, x), p2), x), x), p0);
                };
                auto Q = [&](T x) {
                    return A::add(A::mul(A::add(A::mul(A::add(A::mul(q3, x), q2), x), q1), x), q0);
                };
                w[0] = Q(A::add(t, one));
                w[1] = P(t);
                w[2] = P(A::sub(one, t));
                w[3] = Q(A::sub((T)2, t));
                break;
            }
            }
        } else {
#pragma unroll
            for (int s = 0; s < S; ++s) w[s] = (T)0;  // unreachable: host validates (family, S)
        }
    }
}

// getcoefs(ker, lim, x): returns false where Julia throws InexactError.
// `base` is the canonical 0-based window base: tap s reads column clamp(base + s, 0, len-1).
template <typename T, int S>
__device__ __forceinline__ bool getcoefs(const KerDev<T> &k, int len, T x, int &base, T (&w)[S])
{
    using A = Ar<T>;
    if (k.boundary == LI_SAFEFLAT) {
        if (x <= k.inf) {  // src/boundaries.jl:141-142: all j = first, w = (0,..,0,1)
#pragma unroll
            for (int s = 0; s < S; ++s) w[s] = (s == S - 1) ? (T)1 : (T)0;
            base = -(S - 1);
            return true;
        }
        if (x >= k.sup) {  // src/boundaries.jl:143-144: all j = last, w = (1,0,..,0)
#pragma unroll
            for (int s = 0; s < S; ++s) w[s] = (s == 0) ? (T)1 : (T)0;
            base = len - 1;
            return true;
        }
    }
    T f = (S & 1) ? A::rint_(x) : A::floor_(x);
    if (!(f >= (T)-9223372036854775808.0 && f < (T)9223372036854775808.0)) return false;
    constexpr int c = (S + 1) >> 1;
    long long j1 = (long long)f - c;  // 0-based index of tap 0 = (j_c - 1) - (c - 1); wraps like Int64
    // canonical form: any base <= -(S-1) clamps every tap to column 0, any base >= len-1 to len-1
    long long lo = -(long long)(S - 1), hi = (long long)len - 1;
    // guard the Int64 wrap-around of j_c - (c-1) near typemin (Julia wraps, then does not clamp)
    if ((long long)f < -9223372036854775807LL + 32) j1 = lo;
    base = (int)(j1 < lo ? lo : (j1 > hi ? hi : j1));
    getweights<T, S>(k, A::sub(x, f), w);
    return true;
}

}  // namespace li
