# solvers.jl -- the exchange step (NCCL communicators behind the C ABI, multi-frame accumulate + all-reduce), LazyAlgebra's
# vector verbs on device arrays and the conjugate-gradient fit: what `conjgrad(H'*H, H'*y)` (test/benchmarks.jl:31-40,
# README.md:36-37) needs to stay on the GPU.

# ------------------------------------------------------------------------------------------------
# The one exchange step of the path: adjoints of many frames accumulated into a model shared by the GPUs of a node
# (README.md:36-37).  One Julia process per GPU; rank 0 creates the id and hands it over by any out-of-band means
# (MPI.jl bcast, a file).  NCCL lives inside libli_b200.so (dlopen at the first call): nothing else to load.
mutable struct Comm
    handle::Ptr{Cvoid}
    rank::Int
    nranks::Int
end
const COMM_ID_BYTES = 128
function comm_unique_id()
    id = Vector{UInt8}(undef, COMM_ID_BYTES)
    check(ccall((:li_comm_unique_id, LIB), Cint, (Ptr{Cvoid}, Csize_t), id, COMM_ID_BYTES))
    return id
end
function Comm(id::Vector{UInt8}, nranks::Integer, rank::Integer)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:li_comm_init_rank, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Ptr{Cvoid}}), id, nranks, rank, ref))
    c = Comm(ref[], Int(rank), Int(nranks))
    finalizer(x -> (x.handle == C_NULL || ccall((:li_comm_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), c)
    return c
end
function allreduce_sum!(comm::Comm, x::DeviceArray{T}) where {T<:Floats}
    check(ccall((:li_allreduce_sum, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}), comm.handle, li_dtype(T), x.ptr,
                length(x), C_NULL))
    check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return x
end

"""
    adjoint_accumulate!(model, ops, ys; α=1, β=0, comm=nothing, nbands=1)

`model <- β*model + α * Σ_ranks Σ_k H_k' * ys[k]` with `H_k = ops[k] = (A1_k, A2_k)` the per-dimension operators of
local frame `k`: the multi-frame accumulate kernel on this GPU, band by band, the NCCL all-reduce of a band
overlapping the next band (`li_adjoint_accumulate_allreduce`).  Replaces the reference's loop
`for k: apply!(1, Adjoint, H_k, y_k, false, 1, model)` (src/tabulated.jl:441-478 per factor).
"""
function adjoint_accumulate!(model::DeviceArray{T,2}, ops::AbstractVector, ys::AbstractVector{<:DeviceArray{T,2}};
                             α::Real=1, β::Real=0, comm::Union{Nothing,Comm}=nothing, nbands::Integer=1) where {T<:Floats}
    n = length(ops)
    n == length(ys) || throw(DimensionMismatch("one frame per operator"))
    handles = Ptr{Cvoid}[f.table.handle for op in ops for f in op]
    srcs = Ptr{Cvoid}[y.ptr for y in ys]
    GC.@preserve ops ys model begin
        check(ccall((:li_adjoint_accumulate_allreduce, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cdouble, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                    comm === nothing ? C_NULL : comm.handle, li_dtype(T), size(model, 1), size(model, 2), n, handles, srcs, α, β,
                    model.ptr, nbands, C_NULL))
        check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    end
    return model
end

# ------------------------------------------------------------------------------------------------
# LazyAlgebra's vector verbs on device arrays: with these, `conjgrad(H'*H, H'*y)` of test/benchmarks.jl:40 runs on
# DeviceArrays without leaving the GPU (one scalar read per dot product).
import LazyAlgebra: vdot, vupdate!, vcombine!, vcopy!, vcopy, vzero!, vnorm2
function vdot(x::DeviceArray{T}, y::DeviceArray{T}) where {T<:Floats}
    length(x) == length(y) || throw(DimensionMismatch("arrays must have the same number of elements"))
    r = Ref{Cdouble}(0)
    check(ccall((:li_vdot, LIB), Cint, (Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cvoid}), li_dtype(T), x.ptr, y.ptr,
                length(x), r, C_NULL))
    return T(r[])
end
vdot(::Type{R}, x::DeviceArray{T}, y::DeviceArray{T}) where {R<:AbstractFloat,T<:Floats} = R(vdot(x, y))
vnorm2(x::DeviceArray) = sqrt(vdot(x, x))
function vupdate!(y::DeviceArray{T}, α::Real, x::DeviceArray{T}) where {T<:Floats}
    check(ccall((:li_vupdate, LIB), Cint, (Cint, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Int64, Ptr{Cvoid}), li_dtype(T), y.ptr, α, x.ptr,
                length(y), C_NULL))
    return y
end
function vcombine!(dst::DeviceArray{T}, α::Real, x::DeviceArray{T}, β::Real, y::DeviceArray{T}) where {T<:Floats}
    check(ccall((:li_vcombine, LIB), Cint, (Cint, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Int64, Ptr{Cvoid}), li_dtype(T),
                dst.ptr, α, x.ptr, β, y.ptr, length(dst), C_NULL))
    return dst
end
vcopy!(dst::DeviceArray{T}, src::DeviceArray{T}) where {T<:Floats} = vcombine!(dst, 1, src, 0, src)
vcopy(x::DeviceArray{T,N}) where {T<:Floats,N} = vcopy!(DeviceArray{T,N}(undef, size(x)), x)
vzero!(x::DeviceArray{T}) where {T<:Floats} = vscale!(x, 0)
vcreate(x::DeviceArray{T,N}) where {T,N} = DeviceArray{T,N}(undef, size(x))

struct CGOptions
    maxiter::Int32
    check_every::Int32
    ftol::Float64
    gtol_abs::Float64
    gtol_rel::Float64
end
struct CGInfo
    iters::Int32
    status::Int32
    rho0::Float64
    rho::Float64
end

"""
    fit_separable(As, y; mu=0, ftol=1e-8, maxiter=50) -> x

`x = argmin |H*x - y|^2 + mu*|x|^2` for `H = As[1]*As[2]*...` (per-dimension operators): conjugate gradients on the
normal equations, entirely inside the library (`li_conjgrad_sep`: the fused kernels, device dot products, no callback
into Julia) -- the solve of test/benchmarks.jl:40, `conjgrad(H'*H, H'*y; ftol=1e-8, maxiter=50)`.
"""
function fit_separable(As::Tuple{Vararg{PerDim}}, y::DeviceArray{T,N}; mu::Real=0, ftol::Real=1e-8, maxiter::Integer=50,
                       check_every::Integer=5) where {T<:Floats,N}
    length(As) == N || throw(DimensionMismatch("one operator per dimension is required"))
    x = DeviceArray{T,N}(undef, ntuple(k -> As[k].ncols, Val(N)))
    handles = Ptr{Cvoid}[A.table.handle for A in As]
    dims = Int64[size(y)...]
    opts = Ref(CGOptions(maxiter, check_every, ftol, 0.0, 0.0))
    info = Ref(CGInfo(0, 0, 0.0, 0.0))
    GC.@preserve As y x begin
        check(ccall((:li_conjgrad_sep, LIB), Cint,
                    (Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{CGOptions}, Ptr{CGInfo}, Ptr{Cvoid}),
                    N, handles, dims, 1, mu, y.ptr, x.ptr, 1, opts, info, C_NULL))
    end
    return x
end
