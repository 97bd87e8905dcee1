# separable.jl -- 2-D separable apply! with on-the-fly coefficients (reference: src/separable.jl) and the fused
# product of per-dimension table operators.

function apply!(α::Real, ::Type{P}, ker1::Kernel{T,S1}, x1::AbstractVector{T}, ker2::Kernel{T,S2},
                x2::AbstractVector{T}, src::AbstractArray{T,2}, β::Real,
                dst::AbstractArray{T,2}) where {P<:Union{Direct,Adjoint},T<:Floats,S1,S2}
    n1, n2 = length(x1), length(x2)
    if P === Direct
        @assert size(dst) == (n1, n2)
        m1, m2 = size(src)
    else
        @assert size(src) == (n1, n2)
        m1, m2 = size(dst)
    end
    d1, d2 = DeviceArray(collect(x1)), DeviceArray(collect(x2))
    # device arrays are used in place; host arrays are staged (one copy each way)
    (src isa DeviceArray) == (dst isa DeviceArray) ||
        error("source and destination must both be host arrays or both be device arrays")
    dsrc = src isa DeviceArray ? src : DeviceArray(convert(Array{T}, src))
    ddst = dst isa DeviceArray ? dst : DeviceArray(convert(Array{T}, dst))
    k1, k2 = Ref(kernel_desc(ker1)), Ref(kernel_desc(ker2))
    check(ccall((:li_sep2d_otf_apply, LIB), Cint,
                (Cint, Ptr{KernelDesc}, Ptr{Cvoid}, Int64, Ptr{KernelDesc}, Ptr{Cvoid}, Int64, Int64, Int64, Int64,
                 Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}),
                P === Adjoint, k1, d1, n1, k2, d2, n2, m1, m2, 1, α, dsrc, β, ddst, C_NULL))
    check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    dst isa DeviceArray || copyto!(dst, Array(ddst))
    return dst
end

"""
    apply!(α, P, ker1, x1, ker2, x2, ker3, x3, src, β, dst)

three dimensions with on-the-fly coefficients (`li_sep_otf_apply`): the extension of src/separable.jl:56-162 the
reference plans (src/separable.jl:14).  Tables are generated on the device for this call, then the fused 3-D kernels run.
"""
function apply!(α::Real, ::Type{P}, ker1::Kernel{T}, x1::AbstractVector{T}, ker2::Kernel{T}, x2::AbstractVector{T},
                ker3::Kernel{T}, x3::AbstractVector{T}, src::AbstractArray{T,3}, β::Real,
                dst::AbstractArray{T,3}) where {P<:Union{Direct,Adjoint},T<:Floats}
    n = Int64[length(x1), length(x2), length(x3)]
    if P === Direct
        @assert size(dst) == Tuple(n)
        m = Int64[size(src)...]
    else
        @assert size(src) == Tuple(n)
        m = Int64[size(dst)...]
    end
    (src isa DeviceArray) == (dst isa DeviceArray) ||
        error("source and destination must both be host arrays or both be device arrays")
    ds = (DeviceArray(collect(x1)), DeviceArray(collect(x2)), DeviceArray(collect(x3)))
    dsrc = src isa DeviceArray ? src : DeviceArray(convert(Array{T}, src))
    ddst = dst isa DeviceArray ? dst : DeviceArray(convert(Array{T}, dst))
    descs = [kernel_desc(ker1), kernel_desc(ker2), kernel_desc(ker3)]
    GC.@preserve descs ds dsrc ddst begin
        kp = Ptr{Cvoid}[pointer(descs, k) for k in 1:3]
        xp = Ptr{Cvoid}[d.ptr for d in ds]
        check(ccall((:li_sep_otf_apply, LIB), Cint,
                    (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Int64, Cdouble, Ptr{Cvoid}, Cdouble,
                     Ptr{Cvoid}, Ptr{Cvoid}),
                    P === Adjoint, 3, kp, xp, n, m, 1, α, dsrc.ptr, β, ddst.ptr, C_NULL))
    end
    dst isa DeviceArray || copyto!(dst, Array(ddst))
    return dst
end

"""
    sep_apply!(α, P, (A1, A2[, A3]), x, β, y)

applies `A1*A2[*A3]` (or its adjoint), `Ak` interpolating dimension `k`, in one call of `li_sep_apply[_host]`
(a single fused pass over HBM for two dimensions).
"""
function sep_apply!(α::Real, ::Type{P}, As::Tuple{Vararg{PerDim}}, x::HostOrDevice{T,N}, β::Real,
                    y::HostOrDevice{T,N}) where {P<:Union{Direct,Adjoint},T,N}
    length(As) == N || throw(DimensionMismatch("one operator per dimension is required"))
    all(k -> interp_dim(As[k]) == k, 1:N) || error("operator k must interpolate dimension k")
    handles = Ptr{Cvoid}[A.table.handle for A in As]
    dims = Int64[size(x)...]
    if x isa DeviceArray
        need = Ref{Csize_t}(0)
        check(ccall((:li_sep_workspace_bytes, LIB), Cint, (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Ptr{Csize_t}),
                    P === Adjoint, N, handles, dims, 1, need))
        ws = DeviceArray{UInt8,1}(undef, (Int(need[]) + 1,))
        check(ccall((:li_sep_apply, LIB), Cint,
                    (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid},
                     Csize_t, Cint, Ptr{Cvoid}),
                    P === Adjoint, N, handles, dims, 1, α, x, β, y, ws, need[], 0, C_NULL))
        check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    else
        check(ccall((:li_sep_apply_host, LIB), Cint,
                    (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                    P === Adjoint, N, handles, dims, 1, α, x, β, y, 0, C_NULL))
    end
    return y
end

"""
    sep_accumulate!(P, ops, α, xs, β, y)

`y <- β*y + α * Σ_k P(H_k) * xs[k]` for separable operators `H_k = A1_k*A2_k*...` (tuples of table-backed factors)
acting on device arrays, in ONE library call (`li_sep_apply_batch`): the inner loop of an iterative fit,
`m += H_k' * y_k` over the observed frames.  With β = 1 long batches run on several internal streams and are summed
in a fixed order (deterministic).  β is applied by the first apply only.
"""
function sep_accumulate!(::Type{P}, ops::AbstractVector, α::Real, xs::AbstractVector{<:DeviceArray},
                         β::Real, y::DeviceArray) where {P<:Operations}
    n = length(ops)
    n == length(xs) || throw(DimensionMismatch("one source array per operator"))
    n == 0 && return y
    N = length(ops[1])
    handles = Ptr{Cvoid}[f.table.handle for op in ops for f in op]
    srcs = Ptr{Cvoid}[x.ptr for x in xs]
    dsts = fill(y.ptr, n)
    β == 1 || vscale!(y, β)     # li_sep_apply_batch applies its beta at every apply: scale once, then accumulate
    GC.@preserve ops xs y begin
        check(ccall((:li_sep_apply_batch, LIB), Cint,
                    (Cint, Cint, Int64, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Cdouble, Ptr{Ptr{Cvoid}}, Cdouble,
                     Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}),
                    P === Adjoint, N, n, handles, C_NULL, 1, α, srcs, 1.0, dsts, 0, C_NULL))
        check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    end
    return y
end


# ------------------------------------------------------------------------------------------------
# `H = A1*A2[*A3]` and `H'` through the reference's own verbs.
#
# LazyAlgebra builds a `Composition` for a product of mappings and applies its factors right to left with
# `vcreate`d temporaries (two passes over memory for A1*A2, four PCIe crossings for host arrays).  When every factor
# is a table-backed operator on its own dimension the product is ONE call of li_sep_apply[_host]: the methods below
# intercept `vcreate` / `apply!` for such compositions, so `H*x`, `H'*y`, `apply!(α, P, H, x, scratch, β, y)` and
# `conjgrad(H'*H, H'*y)` of the reference's users run the fused kernels unchanged.  Any other composition of these
# operators (two factors on one dimension, a factor without a fixed dimension) is applied factor by factor, like
# LazyAlgebra does.
const SepComposition = LazyAlgebra.Composition{N,<:NTuple{N,PerDim}} where {N}
factors(H::LazyAlgebra.Composition) = getfield(H, 1)          # the tuple of operands (field `ops` in LazyAlgebra 0.2)

function per_dimension(H::SepComposition, ndims::Int)
    slots = Vector{Any}(nothing, ndims)
    for A in factors(H)
        d = interp_dim(A)
        (d isa Int && 1 ≤ d ≤ ndims && slots[d] === nothing) || return nothing
        slots[d] = A
    end
    return slots
end

function _sep_outdims(H::SepComposition, ::Type{P}, x::AbstractArray{T,N}) where {P,T,N}
    dims = collect(size(x))
    for A in factors(H)
        d = interp_dim(A)
        d isa Int || error("dimension to interpolate must be provided")
        1 ≤ d ≤ N || error("out of range dimension $d")
        nin, nout = P === Direct ? (A.ncols, A.nrows) : (A.nrows, A.ncols)
        dims[d] == nin || throw(DimensionMismatch("dimension $d of source must be $nin"))
        dims[d] = nout
    end
    return Tuple(dims)
end

vcreate(::Type{P}, H::SepComposition, x::HostOrDevice{T,N}, scratch::Bool=false) where {P<:Union{Direct,Adjoint},T<:Floats,N} =
    similar_array(x, T, _sep_outdims(H, P, x))

function apply!(α::Real, ::Type{P}, H::SepComposition, x::HostOrDevice{T,N}, scratch::Bool, β::Real,
                y::HostOrDevice{T,N}) where {P<:Union{Direct,Adjoint},T<:Floats,N}
    size(y) == _sep_outdims(H, P, x) || throw(DimensionMismatch("incompatible destination and source dimensions"))
    slots = per_dimension(H, N)
    if slots !== nothing
        return sep_apply_tables!(α, P, slots, x, β, y)
    end
    # generic: factor by factor (right to left for Direct, left to right for Adjoint), alpha and beta in the last one
    fs = P === Direct ? reverse(collect(factors(H))) : collect(factors(H))
    z = x
    for (k, A) in enumerate(fs)
        if k == length(fs)
            apply!(α, P, A, z, false, β, y)
        else
            z = apply!(1, P, A, z, false, 0, vcreate(P, A, z, false))
        end
    end
    return y
end

# li_sep_apply[_host] with one (possibly missing) table per dimension
function sep_apply_tables!(α::Real, ::Type{P}, slots::Vector, x::HostOrDevice{T,N}, β::Real,
                           y::HostOrDevice{T,N}) where {P<:Union{Direct,Adjoint},T,N}
    handles = Ptr{Cvoid}[A === nothing ? C_NULL : A.table.handle for A in slots]
    dims = Int64[size(x)...]
    GC.@preserve slots x y begin
        if x isa DeviceArray
            # two tables on dimensions 1-2 or three tables on dimensions 1-3 normally run as ONE fused kernel that needs
            # no workspace: try that first (LI_ERR_ARG = 1 "workspace too small" when the fused kernels decline)
            st = ccall((:li_sep_apply, LIB), Cint,
                       (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid},
                        Csize_t, Cint, Ptr{Cvoid}),
                       P === Adjoint, N, handles, dims, 1, α, x, β, y, C_NULL, 0, 0, C_NULL)
            if st == 0
                check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
                return y
            end
            need = Ref{Csize_t}(0)
            check(ccall((:li_sep_workspace_bytes, LIB), Cint, (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Ptr{Csize_t}),
                        P === Adjoint, N, handles, dims, 1, need))
            ws = need[] > 0 ? DeviceArray{UInt8,1}(undef, (Int(need[]),)) : nothing
            check(ccall((:li_sep_apply, LIB), Cint,
                        (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid},
                         Csize_t, Cint, Ptr{Cvoid}),
                        P === Adjoint, N, handles, dims, 1, α, x, β, y, ws === nothing ? C_NULL : ws.ptr, need[], 0, C_NULL))
            check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
        else
            check(ccall((:li_sep_apply_host, LIB), Cint,
                        (Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                        P === Adjoint, N, handles, dims, 1, α, x, β, y, 0, C_NULL))
        end
    end
    return y
end
