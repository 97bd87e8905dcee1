"""li_b200 -- B200-native (sm_100a) forward/adjoint application of the interpolation operators of
emmt/LinearInterpolators.jl, behind the reference's own LazyAlgebra-style API.

The directory is named ``linearinterpolators.jl_b200``; it is imported as ``li_b200`` through the loader
module ``li_b200.py`` at the repository root.  Everything numerical happens in ``libli_b200.so``
(hand-written CUDA, C ABI in ``include/li_b200.h``); importing this package fails if that library has
not been built -- there is no CPU fallback.
"""
from . import _lib
from ._arrays import DeviceArray, HostStream, async_host, device_empty, pinned_empty, to_device, to_host
from ._lib import (ArgumentError, CudaError, DimensionMismatch, ErrorException, InexactError, LIB_PATH,
                   VARIANT_ATOMIC, VARIANT_AUTO, VARIANT_GATHER, get_option, launch_count, set_option)
from .kernels import (Boundaries, CardinalCubicSpline, CardinalCubicSplinePrime, CatmullRomSpline, CatmullRomSplinePrime,
                      CubicSpline, CubicSplinePrime, Flat, Kernel, KeysSpline, KeysSplinePrime, LanczosKernel,
                      LanczosKernelPrime, LinearSpline, LinearSplinePrime, MitchellNetravaliSpline,
                      MitchellNetravaliSplinePrime, QuadraticSpline, QuadraticSplinePrime, RectangularSpline,
                      RectangularSplinePrime, SafeFlat, boundaries, iscardinal, isnormalized)
from .mapping import (Adjoint, AdjointMapping, Composition, Direct, Gram, LinearMapping, Mapping, Operations,
                      Table, apply, apply_, apply_b, vcreate, vdot)
from .nonseparable import AffineTransform2D, TwoDimensionalTransformInterpolator, affine_apply, intercept, rotate
from .separable import sep2d_otf_apply, sep_apply_batch, sep_apply_tables, sep_otf_apply
from .sparse import (AtA, AtWA, Range, SparseInterpolator, SparseUnidimensionalInterpolator, fit, fractional_index,
                     regularize, regularize_)
from .tabulated import TabulatedInterpolator, coefficients, columns, ncols, nrows, width
from .unidimensional import otf1d_apply
from .distributed import (Comm, FrameBatch, accumulate_adjoint, adjoint_accumulate, bind_host_to_device, shard_range,
                          world_info)
from .solvers import CgResult, conjgrad, fit_separable, vcombine, vdot_device, vupdate


def getcoefs(ker, length, x):
    """``getcoefs(ker, limits(ker, len), x) -> (j_1..j_S, w_1..w_S)`` (src/boundaries.jl:77-156), evaluated on
    the device; indices are 1-based like the reference's."""
    t = Table.from_positions(ker, [x], int(length))
    return tuple(int(j) for j in t.J[:, 0]) + tuple(t.W[:, 0])


def version() -> int:
    return int(_lib.lib.li_version())
