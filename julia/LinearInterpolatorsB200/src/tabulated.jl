# tabulated.jl -- TabulatedInterpolator with device-resident tables (reference: src/tabulated.jl).

struct TabulatedInterpolator{T<:AbstractFloat,S,D} <: LinearMapping
    nrows::Int
    ncols::Int
    table::Table{T}
    d::D
end
(A::TabulatedInterpolator)(x::AbstractArray) = apply(A, x)
Base.getproperty(A::TabulatedInterpolator, f::Symbol) =
    f === :J ? export!(getfield(A, :table)).J : f === :W ? export!(getfield(A, :table)).W : getfield(A, f)
coefficients(A::TabulatedInterpolator) = A.W
columns(A::TabulatedInterpolator) = A.J

function TabulatedInterpolator(::Type{T}, d::D, ker::Kernel{T,S}, x::AbstractVector{<:Real},
                               ncols::Integer) where {T<:Floats,D<:Union{Nothing,Int},S}
    length(x) ≥ 1 || error("number of positions too small")
    ncols ≥ 1 || error("number of columns too small")
    return TabulatedInterpolator{T,S,D}(length(x), ncols, Table(ker, x, ncols), d)
end
TabulatedInterpolator(::Type{T}, d::D, ker::Kernel{T,S}, pos::Function, nrows::Integer,
                      ncols::Integer) where {T<:Floats,D<:Union{Nothing,Int},S} =
    (nrows ≥ 1 || error("number of rows too small");
     TabulatedInterpolator(T, d, ker, T[pos(i) for i in 1:nrows], ncols))
TabulatedInterpolator(::Type{T}, d::Union{Nothing,Int}, ker::Kernel, args...) where {T<:Floats} =
    TabulatedInterpolator(T, d, T(ker), args...)
TabulatedInterpolator(d::Union{Nothing,Int}, ker::Kernel{T}, pos::Function, nrows::Integer, ncols::Integer) where {T} =
    TabulatedInterpolator(float(T), d, ker, pos, nrows, ncols)
TabulatedInterpolator(d::Union{Nothing,Int}, ker::Kernel{Tk}, x::AbstractVector{Tx}, ncols::Integer) where {Tk,Tx} =
    TabulatedInterpolator(float(promote_type(Tk, Tx)), d, ker, x, ncols)
TabulatedInterpolator(T::DataType, ker::Kernel, args...) = TabulatedInterpolator(T, nothing, ker, args...)
TabulatedInterpolator(ker::Kernel, args...) = TabulatedInterpolator(nothing, ker, args...)
TabulatedInterpolator(T::DataType, d::Integer, ker::Kernel, args...) = TabulatedInterpolator(T, Int(d), ker, args...)
TabulatedInterpolator(d::Integer, ker::Kernel, args...) = TabulatedInterpolator(Int(d), ker, args...)

const RorC{T<:AbstractFloat} = Union{T,Complex{T}}
const HostOrDevice{T,N} = Union{Array{T,N},DeviceArray{T,N}}

_dim(A::TabulatedInterpolator{R,S,Int}, ::Val{N}) where {R,S,N} = A.d
_dim(A::TabulatedInterpolator{R,S,Nothing}, ::Val{1}) where {R,S} = 1
_dim(A::TabulatedInterpolator{R,S,Nothing}, ::Val{N}) where {R,S,N} = error("dimension to interpolate must be provided")

function _outdims(A, P, d::Int, src::AbstractArray{T,N}) where {T,N}
    1 ≤ d ≤ N || error("out of range dimension $d")
    n = P === Direct ? A.nrows : A.ncols
    return ntuple(i -> (i == d ? n : size(src, i)), Val(N))
end

vcreate(::Type{P}, A::TabulatedInterpolator{R}, src::HostOrDevice{T,N}, scratch::Bool=false) where {P<:Operations,R,T<:RorC{R},N} =
    similar_array(src, T, _outdims(A, P, _dim(A, Val(N)), src))
vcreate(::Type{P}, A::TabulatedInterpolator{R,S,Nothing}, d::Integer, src::HostOrDevice{T,N}, scratch::Bool=false) where {P<:Operations,R,S,T<:RorC{R},N} =
    similar_array(src, T, _outdims(A, P, Int(d), src))

apply!(α::Real, ::Type{P}, A::TabulatedInterpolator{R}, src::HostOrDevice{T,N}, scratch::Bool, β::Real,
       dst::HostOrDevice{T,N}) where {P<:Operations,R,T<:RorC{R},N} =
    _apply!(α, P, A, _dim(A, Val(N)), src, β, dst)
apply!(α::Real, ::Type{P}, A::TabulatedInterpolator{R,S,Nothing}, d::Integer, src::HostOrDevice{T,N}, scratch::Bool,
       β::Real, dst::HostOrDevice{T,N}) where {P<:Operations,R,S,T<:RorC{R},N} =
    _apply!(α, P, A, Int(d), src, β, dst)

function _apply!(α::Real, ::Type{P}, A::TabulatedInterpolator{R}, d::Int, src::AbstractArray{T,N}, β::Real,
                 dst::AbstractArray{T,N}) where {P<:Operations,R,T<:RorC{R},N}
    1 ≤ d ≤ N || error("out of range dimension $d")
    nin, nout = P === Direct ? (A.ncols, A.nrows) : (A.nrows, A.ncols)
    size(dst, d) == nout || throw(DimensionMismatch("dimension $d of destination must be $nout"))
    size(src, d) == nin || throw(DimensionMismatch("dimension $d of source must be $nin"))
    all(i -> i == d || size(src, i) == size(dst, i), 1:N) ||
        throw(DimensionMismatch("incompatible destination and source dimensions"))
    npre = prod(size(src)[1:d-1]) * (T <: Complex ? 2 : 1)   # Complex{R} data == R data with a leading dimension of 2
    npost = prod(size(src)[d+1:end])
    tab_apply!(P === Adjoint, A.table, npre, npost, α, src, β, dst)
    return dst
end
