# bench_reference.jl -- times the UNMODIFIED reference (single-threaded Julia) on the path this
# repository accelerates, for anyone who has Julia: ComputeFCFV + ElementAssemblyLoopNEW (incl. Sparsify)
# + ComputeResidualsFCFV_Stokes_o1, warm-up + best of 3, reported in Melements/s like bench.py.
#
#   cd <reference checkout>;  julia --project=. /path/to/repo/julia/bench_reference.jl
#
# bench.py --impl reference times the C restatement of the same algorithm instead (no Julia here).
using FCFV_NME23, Printf

function run(res)
    mesh = LoadExternalMesh(res, [1.0 1e-2])
    nel, nf = mesh.nel, mesh.nf
    VxDir, VyDir = copy(mesh.xf), -copy(mesh.yf)
    z, ze = zeros(nf), zeros(nel)
    mesh.τe = 2.0 .* ones(nf); mesh.τ = zeros(nf)
    best = Inf
    for rep in 1:4
        t = @elapsed begin
            α, β, Ζ = ComputeFCFV(mesh, ze, ze, VxDir, VyDir, z, z, z, z, 2.0, :SymmetricGradient)
            Kuu, Muu, Kup, fu, fp, ts = ElementAssemblyLoopNEW(mesh, α, β, Ζ, VxDir, VyDir, z, z, z, z, :SymmetricGradient)
            redirect_stdout(devnull) do
                ComputeResidualsFCFV_Stokes_o1(VxDir, VyDir, ze, mesh, α, β, Ζ, ze, ze, VxDir, VyDir, z, z, z, z, :SymmetricGradient)
            end
        end
        rep > 1 && (best = min(best, t))
    end
    @printf("%s: %d elements, %.3f s, %.4f Melements/s (1 thread)\n", res, nel, best, nel / best / 1e6)
end

run(:LowRes); run(:MedRes)
