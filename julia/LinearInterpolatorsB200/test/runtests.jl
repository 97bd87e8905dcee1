# Mirrors the reference's test/runtests.jl for the hot path; needs Julia >= 1.3, the three packages of Project.toml,
# a B200 and libli_b200.so (UNTESTED in the build image, which has no Julia).
using Test, LazyAlgebra, InterpolationKernels, TwoDimensional
using LinearInterpolatorsB200
using LinearInterpolatorsB200: getcoefs

@testset "tables and apply" begin
    x = range(-1, stop=1, length=301); xsub = x[1:3:end]
    y = cos.(2 .* x .* (x .+ 2)); ysub = y[1:3:end]
    t = range(1, stop=length(xsub), length=length(x))
    for K in (LinearSpline(), CatmullRomSpline(), LanczosKernel(4)), C in (Flat, SafeFlat)
        ker = C(K)
        A = TabulatedInterpolator(ker, t, length(xsub))
        @test maximum(abs.(A*ysub .- y)) ≤ 0.05
        S = SparseInterpolator(ker, t, length(xsub))
        @test S(ysub) ≈ A(ysub)
        u, v = randn(length(xsub)), randn(length(x))
        @test vdot(A*u, v) ≈ vdot(u, A'*v)
    end
    @test getcoefs(CatmullRomSpline(), 5, 0.5) == (1, 1, 1, 2, -1/16, 9/16, 9/16, -1/16)
end
