// li_kerdev.cuh -- host-side construction of the device kernel description.
//
// InterpolationKernels 0.1.x stores the constants of a Kernel{T,S,B} as fields of
// type T computed at construction; we do the same on the host (in Float64, rounded
// once to T) so the device only ever sees T values.  Limits follow
// src/boundaries.jl:37-42: prevfloat(T(2 - S/2)), nextfloat(T(len - 1 + S/2)).
#pragma once
#include <cmath>

#include "li_common.cuh"
#include "li_weights.cuh"

namespace li {

inline int expected_size(int family)
{
    switch (family) {
    case LI_RECTANGULAR: return 1;
    case LI_LINEAR: return 2;
    case LI_QUADRATIC: return 3;
    case LI_CUBIC:
    case LI_CATMULLROM:
    case LI_CARDINALCUBIC:
    case LI_KEYS:
    case LI_MITCHELLNETRAVALI: return 4;
    default: return -1;
    }
}

template <typename T>
int make_kerdev(const li_kernel_desc *d, int64_t len, KerDev<T> *k)
{
    LI_REQUIRE(d, LI_ERR_ARG, "null kernel description");
    LI_REQUIRE((sizeof(T) == 4) == (d->dtype == LI_F32), LI_ERR_ARG, "kernel dtype mismatch");
    LI_REQUIRE(d->boundary == LI_FLAT || d->boundary == LI_SAFEFLAT, LI_ERR_ARG, "unknown boundary conditions");
    LI_REQUIRE(len >= 1, LI_ERR_ARG, "dimension length must be at least 1");  // src/types.jl:30,49
    const int S = d->S;
    const int family = d->family & ~LI_PRIME;
    if (family == LI_LANCZOS) {
        LI_REQUIRE(S >= 2 && S <= LI_MAX_S && (S & 1) == 0, LI_ERR_UNSUPPORTED,
                   "LanczosKernel size must be even and in 2:%d (got %d)", LI_MAX_S, S);
    } else {
        LI_REQUIRE(expected_size(family) == S, LI_ERR_ARG, "kernel family %d cannot have size %d", family, S);
    }
    memset(k, 0, sizeof *k);
    k->family = family;
    k->deriv = (d->family & LI_PRIME) ? 1 : 0;
    k->S = S;
    k->boundary = d->boundary;
    const double pi = 3.14159265358979323846;
    switch (family) {
    case LI_QUADRATIC:
        k->c[0] = (T)0.35355339059327376220042218105242451964241796884424;
        k->c[1] = (T)0.70710678118654752440084436210484903928483593768847;
        k->c[2] = (T)3 / (T)4;
        break;
    case LI_CUBIC:
        k->c[0] = (T)1 / (T)2;
        k->c[1] = (T)2 / (T)3;
        k->c[2] = (T)1 / (T)6;
        break;
    case LI_CARDINALCUBIC:
        k->c[0] = (T)((d->params[0] - 1.0) / 2.0);
        k->c[1] = (T)((d->params[0] + 1.0) / 2.0);
        break;
    case LI_KEYS: {
        const double a = d->params[0];  // docs/src/notes.md:243-249
        k->c[0] = (T)1.0;
        k->c[1] = (T)(-a - 3.0);
        k->c[2] = (T)(a + 2.0);
        k->c[3] = (T)(-4.0 * a);
        k->c[4] = (T)(8.0 * a);
        k->c[5] = (T)(-5.0 * a);
        k->c[6] = (T)a;
        break;
    }
    case LI_MITCHELLNETRAVALI: {
        const double b = d->params[0], c = d->params[1];  // docs/src/notes.md:251-260
        k->c[0] = (T)((6.0 - 2.0 * b) / 6.0);
        k->c[1] = (T)((-18.0 + 12.0 * b + 6.0 * c) / 6.0);
        k->c[2] = (T)((12.0 - 9.0 * b - 6.0 * c) / 6.0);
        k->c[3] = (T)((8.0 * b + 24.0 * c) / 6.0);
        k->c[4] = (T)((-12.0 * b - 48.0 * c) / 6.0);
        k->c[5] = (T)((6.0 * b + 30.0 * c) / 6.0);
        k->c[6] = (T)((-b - 6.0 * c) / 6.0);
        break;
    }
    case LI_LANCZOS: {
        const double a = (double)S / 2.0;
        k->c[0] = (T)a;
        k->c[1] = (T)(a / (pi * pi));
        k->c[2] = (T)(pi / a);
        k->c[3] = (T)pi;
        break;
    }
    default:
        break;
    }
    const double half_s = (double)S / 2.0;
    const T lo = (T)(2.0 - half_s);
    const T hi = (T)((double)(len - 1) + half_s);
    k->inf = std::nextafter(lo, (T)-INFINITY);
    k->sup = std::nextafter(hi, (T)INFINITY);
    return LI_OK;
}

}  // namespace li
