# nonseparable.jl -- TwoDimensionalTransformInterpolator (reference: src/nonseparable.jl).  Only the six
# coefficients of the AffineTransform2D cross the C boundary.

struct TwoDimensionalTransformInterpolator{T<:AbstractFloat,K1<:Kernel{T},K2<:Kernel{T}} <: LinearMapping
    rows::Dims{2}
    cols::Dims{2}
    ker1::K1
    ker2::K2
    R::AffineTransform2D{T}
end
TwoDimensionalTransformInterpolator(rows::Dims{2}, cols::Dims{2}, ker::Kernel, R::AffineTransform2D) =
    TwoDimensionalTransformInterpolator(rows, cols, ker, ker, R)
function TwoDimensionalTransformInterpolator(rows::Dims{2}, cols::Dims{2}, ker1::Kernel{T1}, ker2::Kernel{T2},
                                             R::AffineTransform2D{Tr}) where {T1,T2,Tr}
    T = promote_type(T1, T2, Tr)
    k1, k2 = convert(Kernel{T}, ker1), convert(Kernel{T}, ker2)
    return TwoDimensionalTransformInterpolator{T,typeof(k1),typeof(k2)}(rows, cols, k1, k2, convert(AffineTransform2D{T}, R))
end
input_size(A::TwoDimensionalTransformInterpolator) = A.cols
output_size(A::TwoDimensionalTransformInterpolator) = A.rows
vcreate(::Type{Direct}, A::TwoDimensionalTransformInterpolator{T}, x::HostOrDevice{T,2}, scratch::Bool=false) where {T} =
    similar_array(x, T, A.rows)
vcreate(::Type{Adjoint}, A::TwoDimensionalTransformInterpolator{T}, x::HostOrDevice{T,2}, scratch::Bool=false) where {T} =
    similar_array(x, T, A.cols)

function apply!(α::Real, ::Type{P}, A::TwoDimensionalTransformInterpolator{T}, x::HostOrDevice{T,2}, scratch::Bool,
                β::Real, y::HostOrDevice{T,2}) where {P<:Union{Direct,Adjoint},T}
    @assert size(x) == (P === Direct ? A.cols : A.rows)
    @assert size(y) == (P === Direct ? A.rows : A.cols)
    return apply!(α, P, A.ker1, A.ker2, A.R, x, β, y)
end

function apply!(α::Real, ::Type{P}, ker1::Kernel{T}, ker2::Kernel{T}, R::AffineTransform2D{T}, src::HostOrDevice{T,2},
                β::Real, dst::HostOrDevice{T,2}) where {P<:Union{Direct,Adjoint},T<:Floats}
    (n1, n2), (m1, m2) = P === Direct ? (size(dst), size(src)) : (size(src), size(dst))
    dsrc = src isa DeviceArray ? src : DeviceArray(src)
    ddst = dst isa DeviceArray ? dst : DeviceArray(dst)
    coefs = Cdouble[R.xx, R.xy, R.x, R.yx, R.yy, R.y]
    k1, k2 = Ref(kernel_desc(ker1)), Ref(kernel_desc(ker2))
    check(ccall((:li_affine2d_apply, LIB), Cint,
                (Cint, Ptr{KernelDesc}, Ptr{KernelDesc}, Ptr{Cdouble}, Int64, Int64, Int64, Int64, Int64, Cdouble,
                 Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
                P === Adjoint, k1, k2, coefs, n1, n2, m1, m2, 1, α, dsrc, β, ddst, 0, 0, C_NULL))
    dst isa DeviceArray || copyto!(dst, Array(ddst))
    return dst
end
apply!(α::Real, ::Type{P}, ker::Kernel{T}, R::AffineTransform2D{T}, src::AbstractArray{T,2}, β::Real,
       dst::AbstractArray{T,2}) where {P<:Operations,T} = apply!(α, P, ker, ker, R, src, β, dst)
