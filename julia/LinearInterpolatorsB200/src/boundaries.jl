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

family_params(::Kernel) = (0.0, 0.0, 0.0, 0.0)
family_params(k::CardinalCubicSpline) = (Float64(k.c), 0.0, 0.0, 0.0)
family_params(k::KeysSpline) = (Float64(k.a), 0.0, 0.0, 0.0)
family_params(k::MitchellNetravaliSpline) = (Float64(k.b), Float64(k.c), 0.0, 0.0)

boundary_code(::Kernel{T,S,Flat}) where {T,S} = Int32(0)
boundary_code(::Kernel{T,S,SafeFlat}) where {T,S} = Int32(1)

kernel_desc(ker::Kernel{T,S}) where {T<:Floats,S} =
    KernelDesc(family_code(ker), Int32(S), boundary_code(ker), li_dtype(T), family_params(ker))

"""
    getcoefs(ker, len, x) -> j1, ..., jS, w1, ..., wS

same result as the reference's `getcoefs(ker, limits(ker, len), x)`, evaluated on the device through a one-row table.
"""
function getcoefs(ker::Kernel{T,S}, len::Integer, x) where {T<:Floats,S}
    t = export!(Table(ker, T[convert(T, x)], len))
    return (ntuple(s -> t.J[s, 1], Val(S))..., ntuple(s -> t.W[s, 1], Val(S))...)
end
