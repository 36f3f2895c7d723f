# gen_golden.jl -- closes the "parity unpinned" gap wherever Julia is available.
# Runs the UNMODIFIED reference functions on the shipped meshes with the seeded inputs of
# this file and dumps α, β, Ζ, τ, Kuu, Kup, Muu, fu, fp (as CSC triples) as raw little-endian binary files (one
# directory per case) that tests/golden/compare_julia_golden.py checks the oracle (and, with --gpu, the CUDA path)
# against, bit for bit.
#
#   cd <reference checkout>;  julia --project=. /path/to/repo/julia/gen_golden.jl  out_dir
#
# Not runnable in this repository's build image (no Julia, no network).
using FCFV_NME23, SparseArrays

function dump_case(out, name, mesh, Formulation, variant)
    nel, nf = mesh.nel, mesh.nf
    VxDir, VyDir = copy(mesh.xf), -copy(mesh.yf)
    sex, sey = zeros(nel), zeros(nel)
    neu = [k .* mesh.xf .+ mesh.yf ./ k for k in 1:4]          # only correctly rounded IEEE operations: identical inputs on the Python side
    mesh.τe = 2.0 .* ones(nf); mesh.τ = zeros(nf)
    α, β, Ζ = ComputeFCFV(mesh, sex, sey, VxDir, VyDir, neu..., 2.0, Formulation)
    f = variant == 0 ? ElementAssemblyLoop : ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, ts = f(mesh, α, β, Ζ, VxDir, VyDir, neu..., Formulation)
    dir = joinpath(out, "$(name)_$(Formulation)_$(variant)")
    mkpath(dir)
    wr(key, a) = open(io -> write(io, a), joinpath(dir, key * ".bin"), "w")      # column-major, Float64 / Int64
    wr("alpha", α); wr("beta", β); wr("Z", Ζ); wr("tau", mesh.τ); wr("fu", fu); wr("fp", fp)
    for (tag, K) in (("Kuu", Kuu), ("Muu", Muu), ("Kup", Kup))
        wr(tag * "_colptr", K.colptr); wr(tag * "_rowval", K.rowval); wr(tag * "_nzval", K.nzval)
    end
    open(io -> println(io, "nel=$(nel) nf=$(nf) nnz_uu=$(nnz(Kuu)) nnz_muu=$(nnz(Muu)) nnz_up=$(nnz(Kup))"), joinpath(dir, "sizes.txt"), "w")
end

out = length(ARGS) > 0 ? ARGS[1] : "golden_julia"
mkpath(out)
for res in (:LowRes, :MedRes), F in (:Gradient, :SymmetricGradient), v in (0, 1)
    dump_case(out, String(res), LoadExternalMesh(res, [1.0 1e-2]), F, v)
end
