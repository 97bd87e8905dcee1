# bench.jl -- the reference's own timings of the hot path on the host CPU, for the "true reference" column of BASELINE.md.
#
#   julia -t 1 --project=<env with LinearInterpolators, LazyAlgebra> baseline/julia/bench.jl [small]
#
# The reference has no multi-threaded path (no Threads.@threads anywhere in its src/), so one thread is what it does.
# Prints one JSON line per configuration: samples per second of forward + adjoint, as bench.py counts them
# (a sample = one element on the interpolated side).  Untested here: the build image has no Julia.
using LinearInterpolators
using LazyAlgebra
using Printf

function best(f, reps)
    f()                                   # compile + warm
    t = Inf
    for _ in 1:reps
        t = min(t, @elapsed f())
    end
    return t
end

small = length(ARGS) >= 1 && ARGS[1] == "small"

# config #1: 1-D SparseInterpolator, CatmullRom, Float64, 10^4 -> 2*10^4 (test/benchmarks.jl shape)
let ncols = 10_000, nrows = 20_000
    ker = CatmullRomSpline(Float64)
    A = SparseInterpolator(ker, collect(range(1, ncols, length=nrows)), ncols)
    x, y = randn(ncols), randn(nrows)
    tf = best(() -> A * x, 200)
    ta = best(() -> A' * y, 200)
    @printf("{\"config\": 1, \"what\": \"SparseInterpolator CatmullRom Float64 1e4->2e4\", \"fwd_us\": %.3f, \"adj_us\": %.3f, \"Gsamples_s\": %.4f, \"threads\": %d}\n",
            1e6 * tf, 1e6 * ta, 2 * nrows / (tf + ta) / 1e9, Threads.nthreads())
end

# config #2: 2-D separable TabulatedInterpolator, Lanczos-4, Float32, 4096^2 -> 8192^2 (two passes: A1*A2 as LazyAlgebra composes them)
let M = small ? 1024 : 4096, N = 2M
    T = Float32
    ker = LanczosKernel(T, 4)
    xs = collect(range(1, M, length=N))
    A1 = TabulatedInterpolator(T, 1, ker, xs, M)
    A2 = TabulatedInterpolator(T, 2, ker, xs, M)
    H = A1 * A2
    x, y = randn(T, M, M), randn(T, N, N)
    tf = best(() -> H * x, 3)
    ta = best(() -> H' * y, 3)
    @printf("{\"config\": 2, \"what\": \"TabulatedInterpolator Lanczos-4 Float32 %d^2->%d^2, forward + adjoint\", \"fwd_ms\": %.3f, \"adj_ms\": %.3f, \"Gsamples_s\": %.4f, \"threads\": %d}\n",
            M, N, 1e3 * tf, 1e3 * ta, 2 * N^2 / (tf + ta) / 1e9, Threads.nthreads())
end
