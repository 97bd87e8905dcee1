# boundaries.jl -- Kernel{T,S,B} -> li_kernel_desc.  The boundary conditions themselves (limits, getcoefs: Flat and
# SafeFlat clamping, reference src/boundaries.jl) run on the device inside li_table_create and the on-the-fly
# operators; nothing is evaluated in Julia.

family_code(::RectangularSpline) = Int32(0)
family_code(::LinearSpline) = Int32(1)
family_code(::QuadraticSpline) = Int32(2)
family_code(::CubicSpline) = Int32(3)
family_code(::CatmullRomSpline) = Int32(4)
family_code(::CardinalCubicSpline) = Int32(5)
family_code(::KeysSpline) = Int32(6)
family_code(::MitchellNetravaliSpline) = Int32(7)
family_code(::LanczosKernel) = Int32(8)
# the derivative kernels (src/LinearInterpolators.jl:23-41): LI_PRIME = 16 or-ed into the family of the base kernel; the
# library then evaluates d/dt of the base weights (li_weights.cuh: getweights_prime)
const LI_PRIME = Int32(16)
family_code(::RectangularSplinePrime) = Int32(0) | LI_PRIME
family_code(::LinearSplinePrime) = Int32(1) | LI_PRIME
family_code(::QuadraticSplinePrime) = Int32(2) | LI_PRIME
family_code(::CubicSplinePrime) = Int32(3) | LI_PRIME
family_code(::CatmullRomSplinePrime) = Int32(4) | LI_PRIME
family_code(::CardinalCubicSplinePrime) = Int32(5) | LI_PRIME
family_code(::KeysSplinePrime) = Int32(6) | LI_PRIME
family_code(::MitchellNetravaliSplinePrime) = Int32(7) | LI_PRIME
family_code(::LanczosKernelPrime) = Int32(8) | LI_PRIME

family_params(::Kernel) = (0.0, 0.0, 0.0, 0.0)
family_params(k::CardinalCubicSpline) = (Float64(k.c), 0.0, 0.0, 0.0)
family_params(k::KeysSpline) = (Float64(k.a), 0.0, 0.0, 0.0)
family_params(k::MitchellNetravaliSpline) = (Float64(k.b), Float64(k.c), 0.0, 0.0)
family_params(k::CardinalCubicSplinePrime) = (Float64(k.c), 0.0, 0.0, 0.0)
family_params(k::KeysSplinePrime) = (Float64(k.a), 0.0, 0.0, 0.0)
family_params(k::MitchellNetravaliSplinePrime) = (Float64(k.b), Float64(k.c), 0.0, 0.0)

boundary_code(::Kernel{T,S,Flat}) where {T,S} = Int32(0)
boundary_code(::Kernel{T,S,SafeFlat}) where {T,S} = Int32(1)

kernel_desc(ker::Kernel{T,S}) where {T<:Floats,S} =
    KernelDesc(family_code(ker), Int32(S), boundary_code(ker), li_dtype(T), family_params(ker))

"""
    getcoefs(ker, len, x) -> j1, ..., jS, w1, ..., wS

same result as the reference's `getcoefs(ker, limits(ker, len), x)`, evaluated on the device through a one-row table.
"""
function getcoefs(ker::Kernel{T,S}, len::Integer, x::T) where {T<:Floats,S}
    t = export!(Table(ker, T[x], len))
    return (ntuple(s -> t.J[s, 1], Val(S))..., ntuple(s -> t.W[s, 1], Val(S))...)
end


"""
    convert_coordinate(T, c)

the reference's extension hook (src/boundaries.jl:65-74, test/runtests.jl:90-98): coordinates of any type (e.g.
Unitful quantities through src/init.jl) are converted to the kernel's floating-point type `T` before `getcoefs`;
packages extend it with their own methods.  Positions cross the C boundary as `T` values: `Table(ker, x, ncols)`
calls it element by element.
"""
convert_coordinate(::Type{T}, c::Real) where {T<:AbstractFloat} = convert(T, c)
getcoefs(ker::Kernel{T,S}, len::Integer, x::Real) where {T<:Floats,S} =
    getcoefs(ker, len, convert_coordinate(T, x))
