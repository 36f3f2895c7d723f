# gen_golden.jl -- closes the "parity unpinned" gap wherever Julia is available.
# Runs the UNMODIFIED reference functions on the shipped meshes with the seeded inputs of
# tests/cases.py and dumps α, β, Ζ, τ, Kuu, Kup, Muu, fu, fp (as CSC triples) to .mat files that
# tests/golden/compare_julia_golden.py can check the oracle (and hence the CUDA path) against.
#
#   cd <reference checkout>;  julia --project=. /path/to/repo/julia/gen_golden.jl  out_dir
#
# Not runnable in this repository's build image (no Julia, no network).
using FCFV_NME23, MAT, SparseArrays, Random

function dump(out, name, mesh, Formulation, variant)
    nel, nf = mesh.nel, mesh.nf
    VxDir, VyDir = copy(mesh.xf), -copy(mesh.yf)
    sex, sey = zeros(nel), zeros(nel)
    neu = [sin.(k .* mesh.xf .+ mesh.yf) for k in 1:4]          # deterministic, seed-free
    mesh.τe = 2.0 .* ones(nf); mesh.τ = zeros(nf)
    α, β, Ζ = ComputeFCFV(mesh, sex, sey, VxDir, VyDir, neu..., 2.0, Formulation)
    f = variant == 0 ? ElementAssemblyLoop : ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, ts = f(mesh, α, β, Ζ, VxDir, VyDir, neu..., Formulation)
    matwrite(joinpath(out, "$(name)_$(Formulation)_$(variant).mat"), Dict(
        "alpha" => α, "beta" => β, "Z" => Ζ, "tau" => mesh.τ, "fu" => fu, "fp" => fp,
        "Kuu_colptr" => Kuu.colptr, "Kuu_rowval" => Kuu.rowval, "Kuu_nzval" => Kuu.nzval,
        "Muu_colptr" => Muu.colptr, "Muu_rowval" => Muu.rowval, "Muu_nzval" => Muu.nzval,
        "Kup_colptr" => Kup.colptr, "Kup_rowval" => Kup.rowval, "Kup_nzval" => Kup.nzval))
end

out = length(ARGS) > 0 ? ARGS[1] : "golden_julia"
mkpath(out)
for res in (:LowRes, :MedRes), F in (:Gradient, :SymmetricGradient), v in (0, 1)
    dump(out, String(res), LoadExternalMesh(res, [1.0 1e-2]), F, v)
end
