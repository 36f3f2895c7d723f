# FCFV_B200.jl -- drop-in Julia shim: same names and signatures as the reference's exported functions
# (src/FCFV_NME23.jl:13,17), every floating-point operation done by libfcfv_b200.so on a B200.
#
#     include("julia/FCFV_B200.jl"); using .FCFV_B200        # instead of the three reference functions
#
# NOTE: Julia is not available in the build image of this repository, so this file has not been
# executed there; tests/ and bench.py drive the very same C entry points through Python ctypes
# (fcfv_nme23_b200/lib.py, api.py), which mirrors this file call for call.
module FCFV_B200

using SparseArrays, Printf

export ComputeFCFV, ComputeFCFV!, ElementAssemblyLoop, ElementAssemblyLoopNEW, ComputeResidualsFCFV_Stokes_o1, ComputeElementValues,
       SchurComplement, MeshGeometry!, PHResidual, PHUpdateP!, PHFusc, CenterPressure!, RenumberFacesFirstSeen!,
       RenumberElementsHilbert!, PrepareNumbering!, PinArrays, FusedPass, release!, NGPUS, CACHE_INPUTS, new_epoch!

const LIB = get(ENV, "FCFV_LIB", joinpath(@__DIR__, "..", "fcfv_nme23_b200", "libfcfv_b200.so"))

const GRADIENT, SYMMETRIC_GRADIENT = Cint(0), Cint(1)
const LOOP_OLD, LOOP_NEW = Cint(0), Cint(1)
const TAU_REFERENCE, TAU_FACE = Cint(0), Cint(1)
const KUU, KUP, MUU, KUUSC = Cint(0), Cint(1), Cint(2), Cint(3)
const CSC1_INT64 = Cint(1)
const A = 1.0                      # src/DiscretisationFCFV_Stokes.jl:1-2
# Muu is returned by the assembly loops and passed to StokesSolvers!, which never reads it (SolversFCFV_Stokes.jl:6).
# WANT_MUU[] = false returns an empty 2nf x 2nf matrix in its place and skips its assembly and export.
const WANT_MUU = Ref(true)

mutable struct Handle
    ptr::Ptr{Cvoid}
    nel::Int
    nf::Int
    orphans::Bool        # the mesh has faces no element references (their τ is only ever what the caller gives them)
end

# GPUs a mesh is spread over when its handle is created.  The reference's functions have no such argument and its
# drivers are single-process scripts, so the choice is made here: FCFV_B200.NGPUS[] = 8 (or ENV["FCFV_NGPUS"]) before
# the first call on a mesh.  With more than one GPU the handle partitions the mesh inside fcfv_set_mesh; every
# function below keeps taking and returning the global arrays.
const NGPUS = Ref(parse(Int, get(ENV, "FCFV_NGPUS", "1")))

# Upload cache (fcfv_upload_cache): the drivers pass the same arrays to ComputeFCFV, ElementAssemblyLoop[NEW] and
# ComputeResidualsFCFV_Stokes_o1 (SolKzFCFV.jl:150,155,205); with the cache on an array the GPU already mirrors -- same
# address, size, epoch and sampled fingerprint -- is not uploaded again, and α, β, Ζ, mesh.τ returned by ComputeFCFV cost
# nothing when they come back as arguments.  A new epoch starts at every ComputeFCFV.  A driver that modifies an input
# array IN PLACE between ComputeFCFV and the calls that follow it (none of the reference's does) calls
# FCFV_B200.new_epoch!(mesh) afterwards, or sets FCFV_B200.CACHE_INPUTS[] = false before the first call on the mesh.
const CACHE_INPUTS = Ref(true)
new_epoch!(h::Handle) = (ccall((:fcfv_new_epoch, LIB), Cint, (Ptr{Cvoid},), h.ptr); nothing)
new_epoch!(mesh) = haskey(HANDLES, mesh) ? new_epoch!(HANDLES[mesh]) : nothing
# converted copies die with the call: the cache must not remember their addresses
forget_temps!(h, pairs...) = (any(p -> !(p[1] === p[2]), pairs) && new_epoch!(h); nothing)

function Handle(device::Integer = 0; ngpus::Integer = NGPUS[])
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ngpus > 1 ? ccall((:fcfv_create_multi, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), r, ngpus) :
                     ccall((:fcfv_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device)
    rc == 0 || error("fcfv_create failed with status $rc: no usable CUDA device (there is no CPU fallback)")
    h = Handle(r[], 0, 0, true)
    finalizer(destroy!, h)
    return h
end

# fcfv_destroy accepts NULL, so destroying twice (release! followed by the finalizer) is harmless
function destroy!(h::Handle)
    p, h.ptr = h.ptr, C_NULL
    p == C_NULL || ccall((:fcfv_destroy, LIB), Cint, (Ptr{Cvoid},), p)
    return nothing
end

function check(h::Handle, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:fcfv_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))
    error("libfcfv_b200 status $rc: $msg")
end

form_id(F::Symbol) = F == :Gradient ? GRADIENT : F == :SymmetricGradient ? SYMMETRIC_GRADIENT :
                     error("Formulation must be :Gradient or :SymmetricGradient")

# One handle per mesh object.  Weak keys (FCFV_Mesh is mutable): a mesh that is no longer referenced does not keep
# its handle -- about 43 GB of device memory for 50 M elements -- alive; the handle's finalizer then frees it.
const HANDLES = WeakKeyDict{Any,Handle}()

"""    release!(mesh)

Frees the device memory held for `mesh` now (drivers that loop over meshes, e.g.
ViscousInclusionFCFV_DiscretisationErrors.jl).  The next call on `mesh` uploads it again."""
function release!(mesh)
    h = pop!(HANDLES, mesh, nothing)
    h === nothing || destroy!(h)
    return nothing
end

f64(x) = convert(Array{Float64}, x)
i64(x) = convert(Array{Int64}, x)

function new_handle()
    h = Handle()
    check(h, ccall((:fcfv_upload_cache, LIB), Cint, (Ptr{Cvoid}, Cint), h.ptr, CACHE_INPUTS[] ? 1 : 0))
    return h
end

# after the mesh has gone up (fcfv_set_mesh, or fcfv_compute_fcfv with the mesh arrays): sizes, orphan flag, numbering check
function adopt!(h::Handle, mesh)
    h.nel, h.nf = mesh.nel, mesh.nf
    orph = Ref{Cint}(1)
    check(h, ccall((:fcfv_has_orphan_faces, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}), h.ptr, orph))
    h.orphans = orph[] != 0
    HANDLES[mesh] = h          # mesh.τ goes up with the field refresh of the entry point that asked for the handle
    if mesh.nel >= 100_000            # the gathers of the GPU passes rely on numbering locality
        span, off = Ref{Float64}(0.0), Ref{Float64}(0.0)
        check(h, ccall((:fcfv_numbering_locality, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), h.ptr, span, off))
        if span[] > 0.05 || off[] > 0.05
            @warn "mesh numbering has little locality (element span $(round(span[], digits=2)), face offset $(round(off[], digits=2)) of the mesh): the GPU passes run 3-7x faster after FCFV_B200.PrepareNumbering!(mesh)"
        end
    end
    return h
end

function handle_for(mesh)
    haskey(HANDLES, mesh) && return HANDLES[mesh]
    h = new_handle()
    e2f, bc = i64(mesh.e2f), i64(mesh.bc)
    Ω, nx, ny, Γ, ke, δ = f64(mesh.Ω), f64(mesh.n_x), f64(mesh.n_y), f64(mesh.Γ), f64(mesh.ke), f64(mesh.δ)
    τe = ismissing(mesh.τe) ? zeros(mesh.nel) : f64(mesh.τe)       # may have nf entries (v2.jl:127)
    GC.@preserve e2f bc Ω nx ny Γ ke δ τe begin
        check(h, ccall((:fcfv_set_mesh, LIB), Cint,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
            h.ptr, mesh.nel, mesh.nf, e2f, bc, Ω, nx, ny, Γ, ke, δ, τe, length(τe)))
    end
    return adopt!(h, mesh)
end

# The drivers mutate mesh.ke / δ / τe / τ after the mesh is built (v2.jl:44,127; SolKzFCFV.jl:25,149):
# refresh those vectors at the top of every entry point.
# `tau = false`: in front of ComputeFCFV, which overwrites mesh.τ on every face that has an element (Discretisation:40).
function refresh!(h::Handle, mesh; tau::Bool = true)
    ke, δ = f64(mesh.ke), f64(mesh.δ)
    τe = ismissing(mesh.τe) ? zeros(mesh.nel) : f64(mesh.τe)
    τ = (ismissing(mesh.τ) || !tau) ? Float64[] : f64(mesh.τ)
    GC.@preserve ke δ τe τ begin
        check(h, ccall((:fcfv_update_fields, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
            h.ptr, ke, δ, τe, length(τe), isempty(τ) ? C_NULL : pointer(τ)))
    end
    forget_temps!(h, (mesh.ke, ke), (mesh.δ, δ), (mesh.τe, τe), ((ismissing(mesh.τ) || !tau) ? τ : mesh.τ, τ))
end

"""Ω, centroids, outward normals, Γ (and τ .= τr) from xv, yv, e2v, e2f, xf, yf -- the loops of
MakeTriangleMesh / LoadExternalMesh2 (CreateMeshFCFV.jl:142-156,180-211 / :300-314,337-367).
`numbering` 0: nodeA=[2 3 1], nodeB=[3 1 2]; 1: nodeA=[3 2 1], nodeB=[1 3 2]."""
function MeshGeometry!(mesh, τr; numbering::Integer = 0)
    h = Handle()
    nel, nf, nv = mesh.nel, mesh.nf, length(mesh.xv)
    e2v, e2f = i64(mesh.e2v), i64(mesh.e2f)
    xv, yv, xf, yf = f64(mesh.xv), f64(mesh.yv), f64(mesh.xf), f64(mesh.yf)
    Ω, xc, yc = zeros(nel), zeros(nel), zeros(nel)
    nx, ny, Γ = zeros(nel, 3), zeros(nel, 3), zeros(nel, 3)
    τ = ismissing(mesh.τ) ? zeros(nf) : f64(mesh.τ)
    GC.@preserve e2v e2f xv yv xf yf Ω xc yc nx ny Γ τ begin
        check(h, ccall((:fcfv_geometry, LIB), Cint,
            (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Cint, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h.ptr, nel, nf, nv, e2v, e2f, xv, yv, xf, yf, numbering, τr, Ω, xc, yc, nx, ny, Γ, τ))
    end
    mesh.Ω, mesh.xc, mesh.yc, mesh.n_x, mesh.n_y, mesh.Γ, mesh.τ = Ω, xc, yc, nx, ny, Γ, τ
    return mesh
end

"""src/DiscretisationFCFV_Stokes.jl:8-44.  One library call (fcfv_compute_fcfv): on the first call for a mesh object the mesh
goes up in front of the element kernel, later only its mutable fields ke, δ, τe (v2.jl:44,127; SolKzFCFV.jl:25,149); with
page-locked arrays (PinArrays) and a large mesh the call runs as an upload | kernel | download pipeline over element chunks."""
function ComputeFCFV(mesh, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, τr, Formulation)
    nel, nf = mesh.nel, mesh.nf
    return ComputeFCFV!(zeros(nel), zeros(nel, 2), zeros(nel, 2, 2), zeros(nf), mesh, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu,
                        τr, Formulation)
end

"""    ComputeFCFV!(α, β, Ζ, τ, mesh, sex, ..., Formulation)

The same into the caller's arrays (not in the reference): with inputs and outputs page-locked once (`PinArrays`), a large mesh runs
through the library's upload | kernel | download pipeline instead of one staged copy after the other.  `mesh.τ` becomes `τ`."""
function ComputeFCFV!(α::Array{Float64}, β::Array{Float64}, Ζ::Array{Float64}, τ::Array{Float64}, mesh, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu,
                      SxyNeu, SyxNeu, τr, Formulation)
    ismissing(mesh.τ) && (mesh.τ = zeros(mesh.nf))      # LoadExternalMesh leaves τ missing
    fresh = !haskey(HANDLES, mesh)
    h = fresh ? new_handle() : HANDLES[mesh]
    fresh || new_epoch!(h)              # first call of a driver pass: everything the driver may have changed goes up again
    nel, nf = mesh.nel, mesh.nf
    e2f, bc = fresh ? (i64(mesh.e2f), i64(mesh.bc)) : (Int64[], Int64[])
    Ω, nx, ny, Γ = fresh ? (f64(mesh.Ω), f64(mesh.n_x), f64(mesh.n_y), f64(mesh.Γ)) : (Float64[], Float64[], Float64[], Float64[])
    ke, δ, τin = f64(mesh.ke), f64(mesh.δ), f64(mesh.τ)
    τe = ismissing(mesh.τe) ? zeros(nel) : f64(mesh.τe)       # may have nf entries (v2.jl:127)
    (length(α) == nel && length(β) == 2nel && length(Ζ) == 4nel && length(τ) == nf) || error("ComputeFCFV!: output arrays of sizes nel, nel x 2, nel x 2 x 2, nf expected")
    sx, sy, vx, vy = f64(sex), f64(sey), f64(VxDir), f64(VyDir)
    sxx, syy, sxy, syx = f64(SxxNeu), f64(SyyNeu), f64(SxyNeu), f64(SyxNeu)
    pi64(a) = isempty(a) ? Ptr{Int64}(C_NULL) : pointer(a)
    pf64(a) = isempty(a) ? Ptr{Float64}(C_NULL) : pointer(a)
    # the Neumann arrays (received and ignored by the reference) go up while α, β, Ζ, τ come down
    GC.@preserve e2f bc Ω nx ny Γ ke δ τe τin sx sy vx vy sxx syy sxy syx α β Ζ τ begin
        check(h, ccall((:fcfv_compute_fcfv, LIB), Cint,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Cint, Float64,
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h.ptr, nel, nf, pi64(e2f), pi64(bc), pf64(Ω), pf64(nx), pf64(ny), pf64(Γ), ke, δ, τe, length(τe), τin,
            form_id(Formulation), A, sx, sy, vx, vy, sxx, syy, sxy, syx, α, β, Ζ, τ))
    end
    fresh && adopt!(h, mesh)
    forget_temps!(h, (mesh.ke, ke), (mesh.δ, δ), (mesh.τe, τe), (mesh.τ, τin),
                  (sex, sx), (sey, sy), (VxDir, vx), (VyDir, vy), (SxxNeu, sxx), (SyyNeu, syy), (SxyNeu, sxy), (SyxNeu, syx))
    mesh.τ = τ
    return α, β, Ζ
end

function export_csc(h::Handle, block::Cint)
    nr, nc, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(h, ccall((:fcfv_block_size, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}, Ref{Int64}, Ref{Int64}), h.ptr, block, nr, nc, nnz))
    colptr, rowval, nzval = Vector{Int64}(undef, nc[] + 1), Vector{Int64}(undef, nnz[]), Vector{Float64}(undef, nnz[])
    GC.@preserve colptr rowval nzval begin
        check(h, ccall((:fcfv_export, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                       h.ptr, block, CSC1_INT64, colptr, rowval, nzval))
    end
    return SparseMatrixCSC{Float64,Int64}(nr[], nc[], colptr, rowval, nzval)
end

function assemble(variant::Cint, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation)
    h = handle_for(mesh); refresh!(h, mesh)
    nel, nf = mesh.nel, mesh.nf
    a, b, z = f64(α), f64(β), f64(Ζ)
    vx, vy, sxx, syy, sxy, syx = f64(VxDir), f64(VyDir), f64(σxxNeu), f64(σyyNeu), f64(σxyNeu), f64(σyxNeu)
    n1, n2, n3, sec = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Float64}(0.0)
    GC.@preserve a b z vx vy sxx syy sxy syx begin
        check(h, ccall((:fcfv_assemble, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Float64}),
            h.ptr, variant, form_id(Formulation), A, a, b, z, vx, vy, sxx, syy, sxy, syx, WANT_MUU[] ? 1 : 0, n1, n2, n3, sec))
    end
    forget_temps!(h, (α, a), (β, b), (Ζ, z), (VxDir, vx), (VyDir, vy), (σxxNeu, sxx), (σyyNeu, syy), (σxyNeu, sxy), (σyxNeu, syx))
    Kuu, Kup = export_csc(h, KUU), export_csc(h, KUP)
    Muu = WANT_MUU[] ? export_csc(h, MUU) : spzeros(2nf, 2nf)
    fu, fp = zeros(2nf, 1), zeros(nel)                  # fu is a 2nf x 1 Matrix in the reference
    GC.@preserve fu fp check(h, ccall((:fcfv_export_rhs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h.ptr, fu, fp))
    return Kuu, Muu, Kup, fu, fp, sec[]                 # sec plays `tsparse`
end

"""src/DiscretisationFCFV_Stokes.jl:102-202"""
ElementAssemblyLoop(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation) =
    assemble(LOOP_OLD, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation)

"""src/DiscretisationFCFV_Stokes.jl:210-347"""
ElementAssemblyLoopNEW(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation) =
    assemble(LOOP_NEW, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation)

"""src/ComputeResidualsFCFV_Stokes.jl:5-108 -- prints the same six lines, returns nothing.
`tau_mode = :Reference` reads mesh.τ[e] like the reference (:17,54); `:Face` uses mesh.τ[e2f[e,i]]."""
function ComputeResidualsFCFV_Stokes_o1(Vxh, Vyh, Pe, mesh, ae, be, ze, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu,
                                        Formulation; tau_mode::Symbol = :Reference)
    h = handle_for(mesh); refresh!(h, mesh)
    vxh, vyh, pe, a, b, z = f64(Vxh), f64(Vyh), f64(Pe), f64(ae), f64(be), f64(ze)
    sx, sy, vx, vy = f64(sex), f64(sey), f64(VxDir), f64(VyDir)
    sxx, syy, sxy, syx = f64(SxxNeu), f64(SyyNeu), f64(SxyNeu), f64(SyxNeu)
    norms = zeros(6)
    GC.@preserve vxh vyh pe a b z sx sy vx vy sxx syy sxy syx norms begin
        check(h, ccall((:fcfv_residuals, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h.ptr, form_id(Formulation), tau_mode == :Face ? TAU_FACE : TAU_REFERENCE, vxh, vyh, pe, a, b, z, sx, sy, vx, vy,
            sxx, syy, sxy, syx, norms, C_NULL, C_NULL, C_NULL))
    end
    forget_temps!(h, (Vxh, vxh), (Vyh, vyh), (Pe, pe), (ae, a), (be, b), (ze, z), (sex, sx), (sey, sy), (VxDir, vx), (VyDir, vy),
                  (SxxNeu, sxx), (SyyNeu, syy), (SxyNeu, sxy), (SyxNeu, syx))
    @printf("Residual of local  equation 20a: %2.2e\n", norms[1])
    @printf("Residual of local  equation 20b: %2.2e\n", norms[2])
    @printf("Residual of local  equation 20c: %2.2e\n", norms[3])
    @printf("Residual of global equation 19a: %2.2e\n", norms[4])
    @printf("Residual of global equation 19a: %2.2e\n", norms[5])
    @printf("Residual of global equation 19b: %2.2e\n", norms[6])
    return
end

"""src/DiscretisationFCFV_Stokes.jl:52-94 -- same positional arguments; Pe, VxDir, VyDir, Formulation are unused there."""
function ComputeElementValues(mesh, Vxh, Vyh, Pe, α, β, Ζ, VxDir, VyDir, Formulation)
    h = handle_for(mesh); refresh!(h, mesh)
    nel = mesh.nel
    vxh, vyh, a, b, z = f64(Vxh), f64(Vyh), f64(α), f64(β), f64(Ζ)
    Vxe, Vye, τxxe, τyye, τxye = zeros(nel), zeros(nel), zeros(nel), zeros(nel), zeros(nel)
    GC.@preserve vxh vyh a b z Vxe Vye τxxe τyye τxye begin
        check(h, ccall((:fcfv_element_values, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h.ptr, vxh, vyh, a, b, z, Vxe, Vye, τxxe, τyye, τxye))
    end
    return Vxe, Vye, τxxe, τyye, τxye
end

# --- the Powell-Hestenes lines of StokesSolvers! that consume the device CSR (SolversFCFV_Stokes.jl:40-49)
"""Kuusc = Kuu .- Kup*(Kppi*Kpu) of StokesSolvers! (src/SolversFCFV_Stokes.jl:23-25), assembled on the GPU next to
Kuu / Kup.  Call after ElementAssemblyLoop[NEW] on the same mesh; `lu(Kuusc)` / `cholesky(Hermitian(Kuusc))` follow."""
function SchurComplement(mesh, penalty)
    h = handle_for(mesh); refresh!(h, mesh)
    n = Ref{Int64}(0)
    check(h, ccall((:fcfv_schur, LIB), Cint, (Ptr{Cvoid}, Float64, Ref{Int64}), h.ptr, penalty, n))
    return export_csc(h, KUUSC)
end

"""ru = fu - Kuu*u - Kup*p, rp = fp - Kpu*u, (norm(ru), norm(rp))"""
function PHResidual(mesh, u, p)
    h = handle_for(mesh)
    uu, pp = f64(vec(u)), f64(vec(p))
    ru, rp, nrm = zeros(2h.nf), zeros(h.nel), zeros(2)
    GC.@preserve uu pp ru rp nrm check(h, ccall((:fcfv_ph_residual, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h.ptr, uu, pp, ru, rp, nrm))
    return ru, rp, nrm[1], nrm[2]
end

"""p .+= (penalty .* ke ./ Ω) .* (fp .- Kpu*u)"""
function PHUpdateP!(mesh, penalty, u, p)
    h = handle_for(mesh)
    uu = f64(vec(u))
    GC.@preserve uu p check(h, ccall((:fcfv_ph_update_p, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}), h.ptr, penalty, uu, p))
    return p
end

"""fusc = fu .- Kup*(Kppi*fp .+ p)"""
function PHFusc(mesh, penalty, p)
    h = handle_for(mesh)
    pp, out = f64(vec(p)), zeros(2h.nf)
    GC.@preserve pp out check(h, ccall((:fcfv_ph_fusc, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}), h.ptr, penalty, pp, out))
    return out
end

"""Pe .= Pe .- mean(Pe) (src/SolversFCFV_Stokes.jl:20)"""
function CenterPressure!(mesh, Pe)
    h = handle_for(mesh)
    m = Ref{Float64}(0.0)
    GC.@preserve Pe check(h, ccall((:fcfv_center_pressure, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), h.ptr, Pe, m))
    return Pe
end

"""Mesh preparation: renumber the faces of `mesh` in the order the element loop first meets them (the numbering of the
reference's own `.dat` meshes, meshes/ReadMeshesDat.jl:36-88) -- the GPU kernels gather element data by face and are
several times faster when face ids follow the element order.  Permutes `e2f` and the face arrays `bc, xf, yf, τ` in
place; call it right after the mesh is created, before any face data is built from `xf, yf`.  Returns `new_of_old`."""
function RenumberFacesFirstSeen!(mesh)
    h = Handle()
    e2f = i64(mesh.e2f)
    perm = Vector{Int64}(undef, mesh.nf)
    e2f_new = Matrix{Int64}(undef, mesh.nel, 3)
    GC.@preserve e2f perm e2f_new check(h, ccall((:fcfv_first_seen_faces, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h.ptr, mesh.nel, mesh.nf, e2f, perm, e2f_new))
    mesh.e2f = e2f_new
    for name in (:bc, :xf, :yf, :τ)
        a = getfield(mesh, name)
        if !ismissing(a) && length(a) == mesh.nf
            b = similar(a); b[perm] = a; setfield!(mesh, name, b)
        end
    end
    release!(mesh)                    # connectivity changed: free the device copy, upload again on next use
    return perm
end

"""    RenumberElementsHilbert!(mesh) -> old_of_new

For a mesh whose ELEMENTS come in an order without spatial coherence: sorts them along a Hilbert curve through their
centroids (`mesh.xc, mesh.yc`) and permutes every element array of the mesh (`src/CreateMeshFCFV.jl:10-44`) and the
element ids stored in `f2e` / `e2e`.  Element-sized arrays built by the driver (`sex`, `sey`, `Pe`, ...) must be built
after the call, or permuted with the returned `old_of_new` (`x = x[old_of_new]`).  Follow with
`RenumberFacesFirstSeen!(mesh)` (or call `PrepareNumbering!`).  Triangle output does not need it."""
function RenumberElementsHilbert!(mesh)
    h = Handle()
    nel = mesh.nel
    xc, yc = f64(mesh.xc), f64(mesh.yc)
    order = Vector{Int64}(undef, nel)
    GC.@preserve xc yc order check(h, ccall((:fcfv_hilbert_elements, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), h.ptr, nel, xc, yc, order))
    new_of_old = invperm(order)
    for name in (:e2v, :e2f, :n_x, :n_y, :Γ, :e2e, :xc, :yc, :Ω, :τe, :ke, :λ, :phase, :δ, :N_x, :N_y)
        a = getfield(mesh, name)
        ismissing(a) && continue
        if size(a, 1) == nel
            setfield!(mesh, name, ndims(a) == 2 ? a[order, :] : a[order])
        elseif ndims(a) == 1 && length(a) > nel        # τe sized nf (ViscousInclusionFCFV_v2.jl:127)
            b = copy(a); b[1:nel] = a[order]; setfield!(mesh, name, b)
        end
    end
    for name in (:f2e, :e2e)
        a = getfield(mesh, name)
        ismissing(a) && continue
        setfield!(mesh, name, map(e -> e > 0 ? new_of_old[e] : e, a))
    end
    release!(mesh)
    return order
end

"""    PrepareNumbering!(mesh) -> (old_of_new_elements, new_of_old_faces)

Both preparation steps: Hilbert element order, then first-seen face numbering."""
function PrepareNumbering!(mesh)
    oe = RenumberElementsHilbert!(mesh)
    pf = RenumberFacesFirstSeen!(mesh)
    return oe, pf
end

mutable struct PinToken
    h::Handle
    arrays::Vector{Any}
end

"""    PinArrays(mesh_or_handle, arrays...) -> token

Optional.  Page-locks Julia arrays that are passed to the entries above again and again (`mesh.e2f`, `mesh.n_x`, ...,
`VxDir`, `sex`, outputs), so that their copies run at PCIe speed (≈ 53 GB/s) instead of through the library's staging
threads (≈ 35 GB/s for pageable memory).  Costs ≈ 0.12 s per GB once.  The token keeps
the arrays alive and un-pins them when it is finalised; do not `resize!` a pinned array."""
function PinArrays(h::Handle, arrays...)
    kept = Any[]
    for a in arrays
        (a isa Array && sizeof(a) > 0) || continue
        check(h, ccall((:fcfv_host_register, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), h.ptr, pointer(a), sizeof(a)))
        push!(kept, a)
    end
    t = PinToken(h, kept)
    finalizer(t) do tok
        for a in tok.arrays
            ccall((:fcfv_host_unregister, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), tok.h.ptr, pointer(a))
        end
    end
    return t
end
PinArrays(mesh, arrays...) = PinArrays(handle_for(mesh), arrays...)

"""    FusedPass(mesh, sex, sey, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Vxh, Vyh, Pe, Formulation; loop=:old, tau_mode=:Reference)

`ComputeFCFV` + `ElementAssemblyLoop[NEW]` + `ComputeResidualsFCFV_Stokes_o1` in one library call (`fcfv_pass`): every
array crosses PCIe once and `α, β, Ζ, τ` come down while the remaining inputs go up -- 1.5x (1.8x on a mesh that is
already resident) faster than the three calls through host buffers.  Returns `(α, β, Ζ, norms)`, sets `mesh.τ`; the
matrices stay on the GPU: `Kuu = FCFV_B200.export_csc(FCFV_B200.handle_for(mesh), FCFV_B200.KUU)` etc."""
function FusedPass(mesh, sex, sey, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Vxh, Vyh, Pe, Formulation;
                   loop::Symbol = :old, tau_mode::Symbol = :Reference)
    ismissing(mesh.τ) && (mesh.τ = zeros(mesh.nf))
    h = handle_for(mesh); refresh!(h, mesh)
    nel, nf = mesh.nel, mesh.nf
    sx, sy, vx, vy = f64(sex), f64(sey), f64(VxDir), f64(VyDir)
    sxx, syy, sxy, syx = f64(σxxNeu), f64(σyyNeu), f64(σxyNeu), f64(σyxNeu)
    vxh, vyh, pe = f64(Vxh), f64(Vyh), f64(Pe)
    α, β, Ζ, τ, norms = zeros(nel), zeros(nel, 2), zeros(nel, 2, 2), zeros(nf), zeros(6)
    n1, n2, n3 = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    GC.@preserve sx sy vx vy sxx syy sxy syx vxh vyh pe α β Ζ τ norms begin
        check(h, ccall((:fcfv_pass, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Cint, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ptr{Float64}),
            h.ptr, loop == :NEW ? LOOP_NEW : LOOP_OLD, form_id(Formulation), A, tau_mode == :Face ? TAU_FACE : TAU_REFERENCE,
            sx, sy, vx, vy, sxx, syy, sxy, syx, vxh, vyh, pe, α, β, Ζ, τ, WANT_MUU[] ? 1 : 0, n1, n2, n3, norms))
    end
    mesh.τ = τ
    return α, β, Ζ, norms
end

end # module
