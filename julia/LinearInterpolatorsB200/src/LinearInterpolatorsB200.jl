# LinearInterpolatorsB200 -- Julia host glue for libli_b200.so.
#
# UNTESTED IN THE BUILD IMAGE (Julia is not installed there).  Every method body is a `ccall` into the C ABI of
# include/li_b200.h; no interpolation arithmetic happens in Julia.  The LazyAlgebra signatures
# `vcreate(P, A, x, scratch)` / `apply!(α, P, A, x, scratch, β, y)` and the InterpolationKernels kernel types are
# the reference's own, so these operators drop in for the hot path of LinearInterpolators.jl.
module LinearInterpolatorsB200

export TabulatedInterpolator, SparseInterpolator, SparseUnidimensionalInterpolator,
    TwoDimensionalTransformInterpolator, DeviceArray, Comm, comm_unique_id, allreduce_sum!, adjoint_accumulate!,
    fit_separable, sep_apply!, sep_accumulate!, rows, columns, nrows, ncols, width, convert_coordinate, AtA, AtWA,
    regularize, regularize!

using InterpolationKernels
using TwoDimensional: AffineTransform2D
using LazyAlgebra
using LazyAlgebra.Foundations
import LazyAlgebra: apply!, vcreate, output_size, input_size, coefficients

const LIB = get(ENV, "LI_B200_LIB",
                normpath(joinpath(@__DIR__, "..", "..", "..", "linearinterpolators.jl_b200", "libli_b200.so")))

include("capi.jl")
include("boundaries.jl")
include("tabulated.jl")
include("sparse.jl")
include("unidimensional.jl")
include("separable.jl")
include("nonseparable.jl")
include("solvers.jl")

end # module
