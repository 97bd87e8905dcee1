# sparse.jl -- SparseInterpolator / SparseUnidimensionalInterpolator over the same device tables
# (reference: src/sparse.jl).  AtA / fit / regularize are out of scope (dense normal equations, not interpolation).

abstract type AbstractSparseInterpolator{T<:AbstractFloat} <: LinearMapping end
Base.eltype(::AbstractSparseInterpolator{T}) where {T} = T

# positions -> fractional indices, evaluated in T on the host exactly like the reference's closure
function fractional_index(::Type{T}, pos::AbstractArray{<:Real}, grd::AbstractRange) where {T<:AbstractFloat}
    c = (convert(T, first(grd)) + convert(T, last(grd))) / 2
    q = 1 / convert(T, step(grd))
    r = convert(T, 1 + length(grd)) / 2
    return T[q * (convert(T, p) - c) + r for p in vec(pos)]
end
fractional_index(::Type{T}, pos::AbstractArray{<:Real}) where {T<:AbstractFloat} = T[T(p) for p in vec(pos)]

struct SparseInterpolator{T<:AbstractFloat,S,N} <: AbstractSparseInterpolator{T}
    table::Table{T}
    nrows::Int
    ncols::Int
    dims::Dims{N}
end
(A::SparseInterpolator)(x::AbstractVector) = apply(A, x)
output_size(A::SparseInterpolator) = A.dims
input_size(A::SparseInterpolator) = (A.ncols,)
coefficients(A::AbstractSparseInterpolator) = vec(export!(A.table).W)
columns(A::AbstractSparseInterpolator) = vec(export!(A.table).J)

# src/sparse.jl:65-89: the operator as an explicit sparse matrix (host side, from the exported tables)
nrows(A::AbstractSparseInterpolator) = A.nrows
ncols(A::AbstractSparseInterpolator) = A.ncols
width(A::AbstractSparseInterpolator) = A.table.S
nrows(A::TabulatedInterpolator) = A.nrows
ncols(A::TabulatedInterpolator) = A.ncols
width(A::TabulatedInterpolator) = getfield(A, :table).S
function rows(A::Union{AbstractSparseInterpolator,TabulatedInterpolator})
    S, n = width(A), nrows(A)
    R = Vector{Int}(undef, S * n)
    k0 = 0
    for i in 1:n                       # R[k+s] = i, src/sparse.jl:75-82
        for s in 1:S
            R[k0 + s] = i
        end
        k0 += S
    end
    return R
end
import SparseArrays
SparseArrays.sparse(A::Union{AbstractSparseInterpolator,TabulatedInterpolator}) =
    SparseArrays.sparse(rows(A), vec(columns(A)), vec(coefficients(A)), nrows(A), ncols(A))   # src/sparse.jl:88-89

# ---- normal equations of a 1-D fit (src/sparse.jl:277-490): dense ncols x ncols matrices built on the device
const RGL_EPS = 1e-9
const RGL_MU = 0.0

"""
    AtA(A) / AtWA(A, w)

yield `A'*A` / `A'*diag(w)*A` (reference `AtA`, `AtWA`, src/sparse.jl:277-356) as a host matrix; a `DeviceArray` of
weights gives a `DeviceArray` (`li_table_atwa`: every element summed by one thread in the reference's order).
"""
AtA(A::SparseInterpolator{T}) where {T} = Array(_atwa(A, nothing))
AtWA(A::SparseInterpolator{T,S,N}, w::AbstractArray{T,N}) where {T,S,N} = Array(_atwa(A, DeviceArray(Array(w))))
AtWA(A::SparseInterpolator{T,S,N}, w::DeviceArray{T,N}) where {T,S,N} = _atwa(A, w)
function _atwa(A::SparseInterpolator{T}, w) where {T}
    w === nothing || size(w) == output_size(A) || throw(AssertionError("size(wgt) == output_size(A)"))
    dst = DeviceArray{T}(undef, (A.ncols, A.ncols))
    GC.@preserve A w dst check(ccall((:li_table_atwa, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                                     A.table.handle, w === nothing ? C_NULL : w.ptr, dst.ptr, C_NULL))
    check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return dst
end

"""
    regularize!(A, ϵ=1e-9, μ=0) / regularize(A, ϵ, μ)

`A + ρ*(ϵ*I + μ*D'*D)`, `ρ` the largest diagonal element (src/sparse.jl:424-490), in place on a `DeviceArray`
(`li_regularize`) or through the device for a host matrix.
"""
function regularize!(A::DeviceArray{T,2}, eps::Real=RGL_EPS, mu::Real=RGL_MU) where {T<:Floats}
    size(A, 1) == size(A, 2) || throw(AssertionError("size(A,1) == size(A,2)"))
    (eps ≥ 0 && mu ≥ 0) || throw(AssertionError("eps ≥ 0 && mu ≥ 0"))
    check(ccall((:li_regularize, LIB), Cint, (Cint, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Ptr{Cvoid}),
                li_dtype(T), A.ptr, size(A, 1), eps, mu, C_NULL))
    return A
end
regularize!(A::Matrix{T}, args...) where {T<:Floats} = copyto!(A, Array(regularize!(DeviceArray(A), args...)))
regularize(A::AbstractMatrix{T}, args...) where {T<:Floats} = regularize!(copyto!(Array{T}(undef, size(A)), A), args...)
function regularize(A::DeviceArray{T,2}, args...) where {T<:Floats}      # the regularized copy stays on the device
    B = DeviceArray{T}(undef, size(A))
    check(ccall((:li_memcpy_d2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), B.ptr, A.ptr, length(A) * sizeof(T), C_NULL))
    return regularize!(B, args...)
end

SparseInterpolator(ker::Kernel{T}, args...) where {T<:AbstractFloat} = SparseInterpolator{T}(ker, args...)
SparseInterpolator{T}(ker::Kernel, args...) where {T<:AbstractFloat} = SparseInterpolator{T}(T(ker), args...)
SparseInterpolator{T}(ker::Kernel{T,S}, pos::AbstractArray{<:Real,N}, grd::AbstractRange) where {T<:Floats,S,N} =
    SparseInterpolator{T,S,N}(Table(ker, fractional_index(T, pos, grd), length(grd)), length(pos), length(grd), size(pos))
SparseInterpolator{T}(ker::Kernel{T,S}, pos::AbstractArray{<:Real,N}, len::Integer) where {T<:Floats,S,N} =
    SparseInterpolator{T,S,N}(Table(ker, fractional_index(T, pos), len), length(pos), Int(len), size(pos))

vcreate(::Type{Direct}, A::SparseInterpolator{T}, x::HostOrDevice{T,1}, scratch::Bool=false) where {T} =
    similar_array(x, T, A.dims)
vcreate(::Type{Adjoint}, A::SparseInterpolator{T,S,N}, x::HostOrDevice{T,N}, scratch::Bool=false) where {T,S,N} =
    similar_array(x, T, (A.ncols,))

function apply!(α::Real, ::Type{P}, A::SparseInterpolator{T,S,N}, x::HostOrDevice{T}, scratch::Bool, β::Real,
                y::HostOrDevice{T}) where {P<:Union{Direct,Adjoint},T,S,N}
    inp, out = P === Direct ? (x, y) : (y, x)
    length(inp) == A.ncols || error("bad vector length (expecting $(A.ncols), got $(length(inp)))")
    size(out) == A.dims || error("bad output array size (expecting $(A.dims), got $(size(out)))")
    tab_apply!(P === Adjoint, A.table, 1, 1, α, x, β, y)
    return y
end

struct SparseUnidimensionalInterpolator{T<:AbstractFloat,S,D} <: AbstractSparseInterpolator{T}
    nrows::Int
    ncols::Int
    table::Table{T}
end
(A::SparseUnidimensionalInterpolator)(x) = apply(A, x)
Base.size(A::SparseUnidimensionalInterpolator) = (A.nrows, A.ncols)

SparseUnidimensionalInterpolator(ker::Kernel{T}, args...) where {T<:AbstractFloat} =
    SparseUnidimensionalInterpolator{T}(ker, args...)
SparseUnidimensionalInterpolator{T}(ker::Kernel, args...) where {T<:AbstractFloat} =
    SparseUnidimensionalInterpolator{T}(T(ker), args...)
function SparseUnidimensionalInterpolator{T}(ker::Kernel{T}, d::Integer, pos::AbstractVector{<:Real},
                                             len::Integer) where {T<:Floats}
    len ≥ 1 || throw(ArgumentError("invalid dimension length"))
    return SparseUnidimensionalInterpolator{T}(ker, d, pos, 1:Int(len))
end
function SparseUnidimensionalInterpolator{T}(ker::Kernel{T,S}, d::Integer, pos::AbstractVector{<:Real},
                                             grd::AbstractRange) where {T<:Floats,S}
    d ≥ 1 || throw(ArgumentError("invalid dimension of interpolation"))
    tbl = Table(ker, fractional_index(T, pos, grd), length(grd))
    return SparseUnidimensionalInterpolator{T,S,Int(d)}(length(pos), length(grd), tbl)
end

function _shape(A::SparseUnidimensionalInterpolator{T,S,D}, P, x::AbstractArray{<:Any,N}) where {T,S,D,N}
    1 ≤ D ≤ N || throw(DimensionMismatch("out of range dimension of interpolation"))
    nx, ny = P === Direct ? (A.ncols, A.nrows) : (A.nrows, A.ncols)
    size(x, D) == nx || throw(DimensionMismatch("dimension $D of `x` must be $nx"))
    return ntuple(i -> (i == D ? ny : size(x, i)), Val(N))
end
vcreate(::Type{P}, A::SparseUnidimensionalInterpolator{T}, x::HostOrDevice{T}, scratch::Bool=false) where {P<:Union{Direct,Adjoint},T} =
    similar_array(x, T, _shape(A, P, x))
function apply!(α::Real, ::Type{P}, A::SparseUnidimensionalInterpolator{T,S,D}, x::HostOrDevice{T,N}, scratch::Bool,
                β::Real, y::HostOrDevice{T,N}) where {P<:Union{Direct,Adjoint},T,S,D,N}
    size(y) == _shape(A, P, x) || throw(DimensionMismatch("`x` and `y` have incompatible dimensions"))
    tab_apply!(P === Adjoint, A.table, prod(size(x)[1:D-1]), prod(size(x)[D+1:N]), α, x, β, y)
    return y
end

# H = A1*A2 of per-dimension operators: one fused pass through li_sep_apply instead of LazyAlgebra's generic
# composition (which would allocate an intermediate and make two passes).
const PerDim = Union{TabulatedInterpolator,SparseUnidimensionalInterpolator}
interp_dim(A::TabulatedInterpolator) = A.d
interp_dim(::SparseUnidimensionalInterpolator{T,S,D}) where {T,S,D} = D
