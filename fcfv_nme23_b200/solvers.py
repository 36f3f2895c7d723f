"""Mirror of the reference's consumer of the assembled system, ``StokesSolvers!`` (src/SolversFCFV_Stokes.jl:6-56), for
the Python side: same arguments, same solver names, same printed lines.  The sparse factorisation stays on the host
(SciPy's SuperLU stands in for UMFPACK / CHOLMOD: it is outside the accelerated path); the lines of the Powell-Hestenes
loop that consume the device CSR -- Schur complement, residuals, right-hand side, pressure update (:23-25, :40-49) --
run on the GPU through ``fcfv_schur`` / ``fcfv_ph_*``.  Call after ``ElementAssemblyLoop[NEW]`` on the same mesh."""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import api

SOLVERS = ("CoupledBackslash", "PowellHestenesCholesky", "PowellHestenesLU")


def StokesSolvers(Vxh, Vyh, Pe, mesh, Kuu, Kup, fu, fp, M, solver, *, Kpu=None, Kpp=None, penalty=1e4, tol=1e-10, verbose=True):
    """Solves ``[Kuu Kup; Kpu Kpp] [u; p] = [fu; fp]`` and overwrites ``Vxh, Vyh, Pe`` in place (like the ``!`` function).
    ``M`` is accepted and, as in the reference, never read.  Returns the number of Powell-Hestenes iterations (0 for the
    coupled solve)."""
    solver = str(solver).lstrip(":")
    if solver not in SOLVERS:
        raise ValueError(f"solver must be one of {SOLVERS}")
    fu = np.asarray(fu, dtype=np.float64).ravel()
    fp = np.asarray(fp, dtype=np.float64).ravel()
    nVx = fu.size // 2
    say = print if verbose else (lambda *a, **k: None)
    say(f"Start solver {solver} for {fu.size + fp.size} DOFs")
    if solver == "CoupledBackslash":
        Kpu = (-Kup.T) if Kpu is None else Kpu
        Kpp = sp.csc_matrix((fp.size, fp.size)) if Kpp is None else Kpp
        K = sp.bmat([[Kuu, Kup], [Kpu, Kpp]], format="csc")
        xh = spla.spsolve(K, np.concatenate([fu, fp]))
        Vxh[:] = xh[:nVx]
        Vyh[:] = xh[nVx:2 * nVx]
        Pe[:], _ = api.center_pressure(mesh, xh[2 * nVx:])          # Pe .- mean(Pe), :20
        return 0
    t0 = time.perf_counter()
    Kuusc = api.SchurComplement(mesh, penalty)                     # :23-25 on the GPU
    Kf = spla.splu(Kuusc.tocsc())                                  # :27 / :30 (host)
    say(f"{'Cholesky' if solver.endswith('Cholesky') else 'LU'} took = {time.perf_counter() - t0:2.2e} s")
    u, p = np.zeros(fu.size), np.zeros(fp.size)
    its = 0
    for rit in range(1, 21):                                        # :39-50
        its = rit
        _, _, (nrmu, nrmp) = api.ph_residual(mesh, u, p)
        say(f"  --> Powell-Hestenes Iteration {rit:02d}\n  Momentum res.   = {nrmu / np.sqrt(fu.size):2.2e}\n"
            f"  Continuity res. = {nrmp / np.sqrt(fp.size):2.2e}")
        if nrmu / np.sqrt(fu.size) < tol and nrmp / np.sqrt(fu.size) < tol:     # sic: both by sqrt(length(ru)), :44
            break
        u = Kf.solve(api.ph_fusc(mesh, penalty, p))
        p = api.ph_update_p(mesh, penalty, u, p)
    Vxh[:] = u[:nVx]
    Vyh[:] = u[nVx:]
    Pe[:] = p
    return its
