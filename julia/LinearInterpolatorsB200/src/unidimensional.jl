# unidimensional.jl -- apply!(α, P, ker, x, src, β, dst) with coefficients evaluated on the device per call
# (reference: src/unidimensional.jl).

function apply!(α::Number, ::Type{P}, ker::Kernel{T,S}, x::AbstractArray{<:Any,N}, src::AbstractArray{T},
                β::Number, dst::AbstractArray{T}) where {P<:Union{Direct,Adjoint},T<:Floats,S,N}
    if P === Direct
        @assert size(dst) == size(x)
    else
        @assert size(src) == size(x)
    end
    len = P === Direct ? length(src) : length(dst)
    (src isa DeviceArray) == (dst isa DeviceArray) ||
        error("source and destination must both be host arrays or both be device arrays")
    dx = x isa DeviceArray ? x : DeviceArray(convert(Array{T}, x))
    dsrc = src isa DeviceArray ? src : DeviceArray(convert(Array{T}, src))      # (device arrays are used in place)
    ddst = dst isa DeviceArray ? dst : DeviceArray(convert(Array{T}, dst))
    desc = Ref(kernel_desc(ker))
    check(ccall((:li_otf1d_apply, LIB), Cint,
                (Cint, Ptr{KernelDesc}, Ptr{Cvoid}, Int64, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                P === Adjoint, desc, dx, length(x), len, α, dsrc, β, ddst, 0, C_NULL))
    check(ccall((:li_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    dst isa DeviceArray || copyto!(dst, Array(ddst))
    return dst
end
