"""The flow of the reference's drivers (examples/ViscousInclusionFCFV_v2.jl:100-150, StokesSolvers!
src/SolversFCFV_Stokes.jl:6-56 with solver = :PowellHestenesLU) on the B200 path, through the Python mirror of the
Julia shim:

    mesh -> ComputeFCFV -> ElementAssemblyLoop -> [Schur complement on the GPU] -> LU (host, SciPy standing in for
    UMFPACK: the factorisation stays outside the path) -> Powell-Hestenes iterations with the residual / update /
    right-hand-side lines on the GPU -> ComputeResidualsFCFV_Stokes_o1 -> ComputeElementValues

    python examples/viscous_inclusion.py [n]          # S(n) criss-cross mesh, default n = 64 (16 384 elements)

Prints the reference's residual lines; they drop to the Powell-Hestenes tolerance."""
import os
import sys
import time

import numpy as np
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcfv_nme23_b200 import api, mesh as M   # noqa: E402


def analytic(x, y, rc, ηm, ηc, er=-1.0, gr=0.0):
    """Velocity and pressure around a circular inclusion of radius rc and viscosity ηc in a matrix of viscosity ηm under
    far-field pure shear er and simple shear gr (Schmid & Podladchikov 2003, complex-potential solution; the
    reference evaluates the same formulas in src/EvalAnalDani.jl).  Coordinates relative to the inclusion centre."""
    Z = np.asarray(x, dtype=float) + 1j * np.asarray(y, dtype=float)
    A = ηm * (ηc - ηm) / (ηc + ηm)
    far = 1j * gr + 2.0 * er
    inside = np.abs(Z) <= rc
    Zo = np.where(inside, 1.0, Z)                          # placeholder inside: avoids 1/0, masked below
    V_in = (ηm / (ηc + ηm)) * far * np.conj(Z) - 0.5j * gr * Z
    phi = -0.5j * ηm * gr * Zo - far * A * rc ** 2 / Zo
    dphi = -0.5j * ηm * gr + far * A * rc ** 2 / Zo ** 2
    psi = (1j * gr - 2.0 * er) * ηm * Zo - far * A * rc ** 4 / Zo ** 3
    V_out = (phi - Zo * np.conj(dphi) - np.conj(psi)) / (2.0 * ηm)
    P_out = -2.0 * ηm * (ηc - ηm) / (ηc + ηm) * np.real(rc ** 2 / Zo ** 2 * far)
    V = np.where(inside, V_in, V_out)
    return V.real, V.imag, np.where(inside, 0.0, P_out)


def setup(n, η=(1.0, 1e-2), R=1.0 / 6.0, exact_bc=False):
    """Unit square, circular inclusion of viscosity η[1] (faces between phases tagged -1, README.md:24), pure shear
    on the boundary (v2.jl:103-108: Dirichlet values at face midpoints)."""
    m = M.MakeStructuredMesh(n, τr=1.0)                # geometry by the CUDA kernel
    inside = np.hypot(m.xc - 0.5, m.yc - 0.5) < R
    m.phase = np.where(inside, 2.0, 1.0)
    m.ke = np.where(inside, η[1], η[0])
    f = np.asarray(m.e2f) - 1
    kmin, kmax = np.full(m.nf, np.inf), np.full(m.nf, -np.inf)
    for i in range(3):
        np.minimum.at(kmin, f[:, i], m.ke); np.maximum.at(kmax, f[:, i], m.ke)
    m.bc = np.where((m.bc == 0) & (kmin != kmax), -1, m.bc).astype(np.int64)
    if exact_bc:        # the infinite-domain solution on the box boundary: it is then the exact solution inside
        VxDir, VyDir, _ = analytic(m.xf - 0.5, m.yf - 0.5, R, η[0], η[1], er=0.5)
    else:
        VxDir, VyDir = (m.xf - 0.5), -(m.yf - 0.5)
    return m, VxDir, VyDir


def solve(n=64, Formulation="SymmetricGradient", τr=2.0, γ=1e5, tol=1e-9, new_loop=False, verbose=True, exact_bc=False,
          η=(1.0, 1e-2), R=1.0 / 6.0):
    m, VxDir, VyDir = setup(n, η, R, exact_bc)
    nel, nf = m.nel, m.nf
    m.τe = τr * np.ones(nel)
    z, ze = np.zeros(nf), np.zeros(nel)
    t0 = time.perf_counter()
    ae, be, ze_ = api.ComputeFCFV(m, ze, ze, VxDir, VyDir, z, z, z, z, τr, Formulation)
    loop = api.ElementAssemblyLoopNEW if new_loop else api.ElementAssemblyLoop
    Kuu, Muu, Kup, fu, fp, tsparse = loop(m, ae, be, ze_, VxDir, VyDir, z, z, z, z, Formulation)
    Kuusc = api.SchurComplement(m, γ)                  # Kuu .- Kup*(Kppi*Kpu), SolversFCFV_Stokes.jl:23-25
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    lu = spla.splu(Kuusc.tocsc())                      # :27 (host factorisation, outside the path)
    t_fact = time.perf_counter() - t0
    u, p = np.zeros(2 * nf), np.zeros(nel)
    its = 0
    for its in range(1, 16):                           # :39-50
        ru, rp, nrm = api.ph_residual(m, u, p)
        if verbose:
            print(f"  --- iteration {its} --- \n  ||res.u|| = {nrm[0] / np.sqrt(2 * nf):2.2e}\n  ||res.p|| = {nrm[1] / np.sqrt(2 * nf):2.2e}")
        if nrm[0] / np.sqrt(2 * nf) < tol and nrm[1] / np.sqrt(2 * nf) < tol:
            break
        u = lu.solve(api.ph_fusc(m, γ, p))
        p = api.ph_update_p(m, γ, u, p)
    Vxh, Vyh, Pe = u[:nf].copy(), u[nf:].copy(), p
    norms = api.ComputeResidualsFCFV_Stokes_o1(Vxh, Vyh, Pe, m, ae, be, ze_, ze, ze, VxDir, VyDir, z, z, z, z, Formulation,
                                               tau_mode="FACE", verbose=verbose)
    Vxe, Vye, Txxe, Tyye, Txye = api.ComputeElementValues(m, Vxh, Vyh, Pe, ae, be, ze_, VxDir, VyDir, Formulation)
    if verbose:
        print(f"nel = {nel}, nnz(Kuu) = {Kuu.nnz}, nnz(Kup) = {Kup.nnz}; set-up + assembly + export {t_asm:.3f} s "
              f"(device CSR build {tsparse * 1e3:.2f} ms), LU {t_fact:.3f} s, {its} Powell-Hestenes iterations")
        print(f"min/max Vx {Vxe.min():+.4f} {Vxe.max():+.4f}   min/max P {Pe.min():+.4f} {Pe.max():+.4f}")
    out = dict(mesh=m, norms=norms, iterations=its, Vxe=Vxe, Vye=Vye, Pe=Pe, Txxe=Txxe, Tyye=Tyye, Txye=Txye)
    if exact_bc:        # L1 discretisation errors against the analytic solution (as ViscousInclusionFCFV_DiscretisationErrors.jl does)
        vx, vy, pa = analytic(m.xc - 0.5, m.yc - 0.5, R, η[0], η[1], er=0.5)
        w = m.Ω / m.Ω.sum()
        pe = Pe - (w * Pe).sum()
        pa = pa - (w * pa).sum()
        out["errors"] = dict(Vx=float((w * np.abs(Vxe - vx)).sum()), Vy=float((w * np.abs(Vye - vy)).sum()),
                             P=float((w * np.abs(pe - pa)).sum()))
        if verbose:
            print("L1 errors vs the analytic solution:", out["errors"])
    return out


if __name__ == "__main__":
    solve(int(sys.argv[1]) if len(sys.argv) > 1 else 64, exact_bc="--exact-bc" in sys.argv)
