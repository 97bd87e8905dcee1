# capi.jl -- thin ccall layer over include/li_b200.h: status codes -> the exceptions the reference throws,
# device arrays, table handles.

const LI_OK, LI_ERR_ARG, LI_ERR_DIM, LI_ERR_INEXACT, LI_ERR_BOUNDS, LI_ERR_UNSUPPORTED, LI_ERR_CUDA, LI_ERR_NOMEM =
    Cint.(0:7)
const Floats = Union{Float32,Float64}
li_dtype(::Type{Float32}) = Cint(0)
li_dtype(::Type{Float64}) = Cint(1)

last_error() = unsafe_string(ccall((:li_last_error, LIB), Cstring, ()))

function check(status::Integer)
    status == LI_OK && return nothing
    msg = last_error()
    status == LI_ERR_DIM && throw(DimensionMismatch(msg))
    status == LI_ERR_INEXACT && throw(InexactError(:trunc, Int, NaN))
    status == LI_ERR_NOMEM && throw(OutOfMemoryError())
    (status == LI_ERR_ARG || status == LI_ERR_BOUNDS) && throw(ErrorException(msg))
    throw(ErrorException("libli_b200: " * msg))
end

# li_kernel_desc
struct KernelDesc
    family::Int32
    S::Int32
    boundary::Int32
    dtype::Int32
    params::NTuple{4,Float64}
end

"""
    DeviceArray{T,N}(undef, dims)

column-major device array backed by `li_malloc` (no CUDA.jl): what `vcreate` returns when its argument lives on
the device.  Plain `Array` arguments are staged through the device by the `*_host` entry points instead.
"""
mutable struct DeviceArray{T,N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    function DeviceArray{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:li_malloc, LIB), Cint, (Ptr{Ptr{Cvoid}}, Csize_t), ref, prod(dims) * sizeof(T)))
        A = new{T,N}(ref[], dims)
        finalizer(a -> ccall((:li_free, LIB), Cint, (Ptr{Cvoid},), a.ptr), A)
        return A
    end
end
DeviceArray{T}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N} = DeviceArray{T,N}(undef, dims)
Base.size(A::DeviceArray) = A.dims
Base.unsafe_convert(::Type{Ptr{Cvoid}}, A::DeviceArray) = A.ptr
Base.getindex(::DeviceArray, I...) = error("DeviceArray lives on the GPU: copy it with Array(A)")
function DeviceArray(A::Array{T,N}) where {T,N}
    D = DeviceArray{T,N}(undef, size(A))
    check(ccall((:li_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), D.ptr, A, sizeof(A), C_NULL))
    check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return D
end
function Base.Array(D::DeviceArray{T,N}) where {T,N}
    A = Array{T,N}(undef, size(D))
    check(ccall((:li_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), A, D.ptr, sizeof(A), C_NULL))
    check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return A
end
similar_array(x::Array, ::Type{T}, dims) where {T} = Array{T}(undef, dims)
similar_array(x::DeviceArray, ::Type{T}, dims) where {T} = DeviceArray{T}(undef, Tuple(dims))

# Opaque li_table handle owned by the Julia operator.
mutable struct Table{T<:Floats}
    handle::Ptr{Cvoid}
    S::Int
    nrows::Int
    ncols::Int
    J::Union{Nothing,Matrix{Int}}   # lazily exported, reference layout (Int64, 1-based, S x nrows)
    W::Union{Nothing,Matrix{T}}
    function Table{T}(handle, S, nrows, ncols) where {T}
        t = new{T}(handle, S, nrows, ncols, nothing, nothing)
        finalizer(x -> ccall((:li_table_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), t)
        return t
    end
end

function Table(ker::Kernel{T,S}, x::AbstractVector, ncols::Integer) where {T<:Floats,S}
    pos = T[convert_coordinate(T, xi) for xi in x]   # convert_coordinate(T, X[i]) of the reference's __maketables
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    desc = Ref(kernel_desc(ker))
    check(ccall((:li_table_create, LIB), Cint,
                (Ptr{KernelDesc}, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                desc, pos, 0, length(pos), ncols, C_NULL, ref))
    return Table{T}(ref[], S, length(pos), Int(ncols))
end

function export!(t::Table{T}) where {T}
    if t.J === nothing
        J = Matrix{Int}(undef, t.S, t.nrows)
        W = Matrix{T}(undef, t.S, t.nrows)
        check(ccall((:li_table_export, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cvoid}),
                    t.handle, J, W, C_NULL))
        t.J, t.W = J, W
    end
    return t
end

# dst <- alpha*P(A)*src + beta*dst along one dimension (li_tab_apply / li_tab_apply_host)
function tab_apply!(adjoint::Bool, t::Table, npre::Integer, npost::Integer, α::Real, src, β::Real, dst)
    f = (src isa DeviceArray) ? :li_tab_apply : :li_tab_apply_host
    (src isa DeviceArray) == (dst isa DeviceArray) ||
        error("source and destination must both be host arrays or both be device arrays")
    if f === :li_tab_apply
        check(ccall((:li_tab_apply, LIB), Cint,
                    (Cint, Ptr{Cvoid}, Int64, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                    adjoint, t.handle, npre, npost, α, src, β, dst, 0, C_NULL))
        check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))   # Julia semantics: blocking
    else
        check(ccall((:li_tab_apply_host, LIB), Cint,
                    (Cint, Ptr{Cvoid}, Int64, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                    adjoint, t.handle, npre, npost, α, src, β, dst, 0, C_NULL))
    end
    return dst
end

# One Julia process per GPU (INTEGRATION.md, "Multi-GPU"): select the device before creating tables / arrays.
set_device(dev::Integer) = check(ccall((:li_set_device, LIB), Cint, (Cint,), dev))
function device_count()
    n = Ref{Cint}(0)
    check(ccall((:li_device_count, LIB), Cint, (Ptr{Cint},), n))
    return Int(n[])
end

"""
    Table(J, W, ncols)

adopt index / weight tables computed in Julia (user-defined kernels: anything `getcoefs` can produce), reference
layout `J::Matrix{Int}` 1-based and `W::Matrix{T}`, both `S x nrows`.  The bound check of `__fullcheck`
(src/tabulated.jl:206-215) is done once here, on the device, instead of at every `apply!`.
"""
function Table(J::Matrix{Int}, W::Matrix{T}, ncols::Integer) where {T<:Floats}
    size(J) == size(W) || throw(DimensionMismatch("J and W must have the same size"))
    S, nrows = size(J)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:li_table_import, LIB), Cint,
                (Cint, Cint, Int64, Int64, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                li_dtype(T), S, nrows, ncols, J, W, C_NULL, ref))
    return Table{T}(ref[], S, nrows, Int(ncols))
end

# vscale!(y, β) on a device array (β = 0: zero-fill without reading y, like LazyAlgebra's vscale!)
function vscale!(y::DeviceArray{T}, β::Real) where {T<:Floats}
    check(ccall((:li_vscale, LIB), Cint, (Cint, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}), li_dtype(T), y.ptr, length(y), β, C_NULL))
    return y
end
