# dump_goldens.jl -- run the REAL reference on the golden inputs and write its tables and results.
#
#   python baseline/julia/export_golden_inputs.py
#   julia --project=<env with LinearInterpolators, LazyAlgebra, InterpolationKernels> baseline/julia/dump_goldens.jl
#
# Output: baseline/_ref/goldens/<tag>.{J,W,dir,adj}.bin (raw little-endian, column-major as Julia stores them; J as
# Int64, 1-based) + cases.tsv (copied).  tests/test_golden.py compares the oracle and the CUDA path with these files
# when they exist: J and W bit for bit, applied values to 1e-6 (Float32) / 1e-12 (Float64).
# Untested here: the build image has no Julia (SURVEY.md section 8c).
using LinearInterpolators
using LazyAlgebra

const ROOT = normpath(joinpath(@__DIR__, "..", ".."))
const IN = joinpath(ROOT, "baseline", "_ref", "golden_inputs")
const OUT = joinpath(ROOT, "baseline", "_ref", "goldens")
mkpath(OUT)

readvec(T, path, n) = (a = Vector{T}(undef, n); read!(path, a); a)
readmat(T, path, n1, n2) = (a = Matrix{T}(undef, n1, n2); read!(path, a); a)

for line in eachline(joinpath(IN, "cases.tsv"))
    isempty(strip(line)) && continue
    tag, ctor, cond, tname, ncols_s, nrows_s, nb_s = split(line, '\t')
    T = tname == "Float32" ? Float32 : Float64
    ncols, nrows, nb = parse(Int, ncols_s), parse(Int, nrows_s), parse(Int, nb_s)
    # the kernel exactly as tests/util.py builds it: eltype T, boundary Flat or SafeFlat
    ker = eval(Meta.parse("let T = $tname; $ctor; end"))
    ker = cond == "SafeFlat" ? SafeFlat(ker) : Flat(ker)
    x = readvec(T, joinpath(IN, "$tag.x.bin"), nrows)
    src = readmat(T, joinpath(IN, "$tag.src.bin"), ncols, nb)
    y = readmat(T, joinpath(IN, "$tag.y.bin"), nrows, nb)
    local A
    try
        A = TabulatedInterpolator(T, 1, ker, x, ncols)     # src/tabulated.jl:89-100 (positions as a vector)
    catch err
        # hostile positions (NaN, |x| >= 2^63) make trunc(Int, .) throw InexactError: recorded, not an error of the dump
        write(joinpath(OUT, "$tag.error.txt"), sprint(showerror, err))
        continue
    end
    write(joinpath(OUT, "$tag.J.bin"), Array{Int64}(A.J))          # [S, nrows]
    write(joinpath(OUT, "$tag.W.bin"), Array{T}(A.W))
    write(joinpath(OUT, "$tag.dir.bin"), Array{T}(A * src))        # apply!(1, Direct, A, src, false, 0, dst)
    write(joinpath(OUT, "$tag.adj.bin"), Array{T}(A' * y))         # apply!(1, Adjoint, A, y, false, 0, dst)
end
cp(joinpath(IN, "cases.tsv"), joinpath(OUT, "cases.tsv"); force=true)
println("goldens written to ", OUT)
