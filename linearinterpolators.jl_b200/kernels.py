"""Value mirror of the InterpolationKernels 0.1.x kernel types ``Kernel{T,S,B}``.

The reference re-exports them (src/LinearInterpolators.jl:19-52) and only ever uses: the type
parameters ``T`` (floating-point type), ``S`` (size) and ``B`` (boundary tag ``Flat``/``SafeFlat``),
the conversions ``T(ker)`` / ``convert(Kernel{T}, ker)`` / ``B(ker)`` and ``getweights`` (which here
runs on the device).  Constructors accept the floating-point type, the parameters and the boundary
tag in flexible order, like the Julia ones (test/runtests.jl:16,76-78; test/demos.jl:22-25).
"""
from __future__ import annotations

import numpy as np

from . import _lib

RECTANGULAR, LINEAR, QUADRATIC, CUBIC, CATMULLROM, CARDINALCUBIC, KEYS, MITCHELLNETRAVALI, LANCZOS = range(9)


class Boundaries:
    """Abstract boundary tag; calling a tag on a kernel returns the kernel with these boundaries."""
    code = None

    def __new__(cls, ker=None):
        if ker is None:
            raise TypeError("boundary tags are used as `Flat` / `SafeFlat(ker)`")
        return ker.with_boundaries(cls)


class Flat(Boundaries):
    code = _lib.FLAT


class SafeFlat(Boundaries):
    code = _lib.SAFEFLAT


def _float_type(T):
    T = np.dtype(T)
    if T not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError(f"kernel floating-point type must be Float32 or Float64, got {T}")
    return T


def _split_args(args):
    T, B, params = None, None, []
    for a in args:
        if isinstance(a, type) and issubclass(a, Boundaries):
            B = a
        elif isinstance(a, (type, np.dtype)) or (isinstance(a, str) and a in ("float32", "float64")):
            T = _float_type(a)
        else:
            params.append(float(a))
    return T, B, params


class Kernel:
    """``Kernel{T,S,B}``."""
    family = None
    S = None
    nparams = 0
    defaults = ()
    is_prime = False        # the derivative kernel `ker'` (`*Prime` types of src/LinearInterpolators.jl:23-41)

    def __init__(self, *args):
        T, B, params = _split_args(args)
        self.T = _float_type(T if T is not None else np.float64)
        self.B = B if B is not None else Flat
        if len(params) == 0:
            params = list(self.defaults)
        if len(params) != self.nparams:
            raise _lib.ArgumentError(f"{type(self).__name__} takes {self.nparams} parameter(s)")
        self.params = tuple(params)

    # -- Julia-like queries
    def eltype(self):
        return self.T

    def boundaries(self):
        return self.B

    def __len__(self):
        return self.S

    def _clone(self, T=None, B=None):
        k = object.__new__(type(self))
        k.__dict__.update(self.__dict__)
        if T is not None:
            k.T = _float_type(T)
        if B is not None:
            k.B = B
        return k

    def with_eltype(self, T):
        """``T(ker)`` / ``convert(Kernel{T}, ker)``."""
        return self if _float_type(T) == self.T else self._clone(T=T)

    def with_boundaries(self, B):
        """``B(ker)``."""
        return self if B is self.B else self._clone(B=B)

    def desc(self) -> _lib.KernelDesc:
        d = _lib.KernelDesc()
        d.family, d.S, d.boundary = self.family | (_lib.PRIME if self.is_prime else 0), self.S, self.B.code
        d.dtype = _lib.F32 if self.T == np.dtype(np.float32) else _lib.F64
        for i in range(4):
            d.params[i] = self.params[i] if i < len(self.params) else 0.0
        return d

    @property
    def prime(self):
        """``ker'``: the derivative kernel (same size, parameters, type and boundaries)."""
        if self.is_prime:
            raise _lib.ArgumentError("the second derivative of a kernel is not available")
        k = object.__new__(_PRIME_OF[type(self)])
        k.__dict__.update(self.__dict__)
        return k

    def __repr__(self):
        p = ", ".join(f"{v:g}" for v in self.params)
        return f"{type(self).__name__}{{{self.T.name},{self.S},{self.B.__name__}}}({p})"


class RectangularSpline(Kernel):
    family, S = RECTANGULAR, 1


class LinearSpline(Kernel):
    family, S = LINEAR, 2


class QuadraticSpline(Kernel):
    family, S = QUADRATIC, 3


class CubicSpline(Kernel):
    family, S = CUBIC, 4


class CatmullRomSpline(Kernel):
    family, S = CATMULLROM, 4


class CardinalCubicSpline(Kernel):
    family, S, nparams = CARDINALCUBIC, 4, 1


class KeysSpline(Kernel):
    family, S, nparams = KEYS, 4, 1


class MitchellNetravaliSpline(Kernel):
    family, S, nparams, defaults = MITCHELLNETRAVALI, 4, 2, (1.0 / 3.0, 1.0 / 3.0)


class LanczosKernel(Kernel):
    """``LanczosKernel([T,] S [,B])``: the size is the (only) parameter and must be even."""
    family, nparams = LANCZOS, 1

    def __init__(self, *args):
        T, B, params = _split_args(args)
        if len(params) != 1 or int(params[0]) != params[0]:
            raise _lib.ArgumentError("LanczosKernel needs its (integer) size")
        S = int(params[0])
        if S < 2 or S % 2 or S > _lib.MAX_S:
            raise _lib.ArgumentError(f"LanczosKernel size must be even and in 2:{_lib.MAX_S}")
        self.T = _float_type(T if T is not None else np.float64)
        self.B = B if B is not None else Flat
        self.S = S
        self.params = ()


# the derivative kernels: same constructors as their base kernels
def _prime_class(base):
    return type(base.__name__.replace("Spline", "SplinePrime").replace("Kernel", "KernelPrime"), (base,),
                {"is_prime": True, "__doc__": f"Derivative of ``{base.__name__}`` (``{base.__name__}(...)'``)."})


RectangularSplinePrime = _prime_class(RectangularSpline)
LinearSplinePrime = _prime_class(LinearSpline)
QuadraticSplinePrime = _prime_class(QuadraticSpline)
CubicSplinePrime = _prime_class(CubicSpline)
CatmullRomSplinePrime = _prime_class(CatmullRomSpline)
CardinalCubicSplinePrime = _prime_class(CardinalCubicSpline)
KeysSplinePrime = _prime_class(KeysSpline)
MitchellNetravaliSplinePrime = _prime_class(MitchellNetravaliSpline)
LanczosKernelPrime = _prime_class(LanczosKernel)
_PRIME_OF = {c.__mro__[1]: c for c in (RectangularSplinePrime, LinearSplinePrime, QuadraticSplinePrime, CubicSplinePrime,
                                       CatmullRomSplinePrime, CardinalCubicSplinePrime, KeysSplinePrime,
                                       MitchellNetravaliSplinePrime, LanczosKernelPrime)}


def boundaries(ker: Kernel):
    return ker.B


def iscardinal(ker: Kernel) -> bool:
    """``iscardinal(ker)``: ker(0) = 1 and ker(k) = 0 at the other integers (interpolating kernels)."""
    if ker.is_prime:
        return False
    return ker.family in (RECTANGULAR, LINEAR, CATMULLROM, CARDINALCUBIC, KEYS, LANCZOS)


def isnormalized(ker: Kernel) -> bool:
    """``isnormalized(ker)``: the weights sum to one for every offset (the Lanczos kernel is not normalised)."""
    if ker.is_prime:
        return False
    return ker.family != LANCZOS

