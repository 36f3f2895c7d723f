"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  Parity: never compared with a Julia run (none is possible here); pinned bit for bit against
the reference's own source text executed through tests/julia_transliteration.py (see the header of ``fcfv_oracle.c``).

Function names and argument order follow the reference's Julia functions; meshes are any
object with the ``FCFV_Mesh`` attribute names (``nel, nf, e2f, bc, Ω, Γ, n_x, n_y, ke, δ, τ, τe``).
Sparse results come back as Julia-style CSC triples ``(colptr, rowval, nzval)``, 1-based int64.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
FORMULATIONS = {"Gradient": 0, "SymmetricGradient": 1}


def build(force=False):
    so = os.path.join(_HERE, "libfcfv_oracle.so")
    src = os.path.join(_HERE, "fcfv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libfcfv_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _form(f):
    return FORMULATIONS[str(f).lstrip(":")]


def _d(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def _i(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).ravel(order="F"))


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def geometry(nel, nf, e2v, e2f, xv, yv, xf, yf, numbering, tau_r):
    """a2 -- returns Ω, xc, yc, n_x, n_y, Γ (nel x 3 order='F'), τ (nf)."""
    Om, xc, yc = np.zeros(nel), np.zeros(nel), np.zeros(nel)
    nx, ny, G = np.zeros(3 * nel), np.zeros(3 * nel), np.zeros(3 * nel)
    tau = np.zeros(nf)
    a = [_i(e2v), _i(e2f), _d(xv), _d(yv), _d(xf), _d(yf)]
    rc = lib().oracle_geometry(C.c_int64(nel), C.c_int64(nf), *map(_p, a), C.c_int(numbering), C.c_double(tau_r),
                               *map(_p, (Om, xc, yc, nx, ny, G, tau)))
    assert rc == 0
    r = lambda x: x.reshape((nel, 3), order="F")
    return Om, xc, yc, r(nx), r(ny), r(G), tau


def mesh_geometry(mesh, tau_r):
    """Fills Ω, xc, yc, n_x, n_y, Γ, τ of a mesh in place with the oracle geometry."""
    mesh.Ω, mesh.xc, mesh.yc, mesh.n_x, mesh.n_y, mesh.Γ, mesh.τ = geometry(
        mesh.nel, mesh.nf, mesh.e2v, mesh.e2f, mesh.xv, mesh.yv, mesh.xf, mesh.yf, mesh.numbering, tau_r)
    return mesh


def ComputeFCFV(mesh, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, τr, Formulation, A=1.0):
    """a4 -- returns α (nel), β (nel x 2), Ζ (nel x 2 x 2); writes mesh.τ (allocated if None)."""
    nel, nf = mesh.nel, mesh.nf
    al, be, Z = np.zeros(nel), np.zeros(2 * nel), np.zeros(4 * nel)
    tau = np.zeros(nf) if mesh.τ is None else _d(mesh.τ).copy()
    a = [_i(mesh.e2f), _i(mesh.bc), _d(mesh.Ω), _d(mesh.Γ), _d(mesh.n_x), _d(mesh.n_y), _d(mesh.τe)[:nel].copy(),
         _d(sex), _d(sey), _d(VxDir), _d(VyDir)]
    rc = lib().oracle_compute_fcfv(C.c_int64(nel), C.c_int64(nf), *map(_p, a), C.c_int(_form(Formulation)),
                                   C.c_double(A), *map(_p, (al, be, Z, tau)))
    assert rc == 0
    mesh.τ = tau
    return al, be.reshape((nel, 2), order="F"), Z.reshape((nel, 2, 2), order="F")


def assembly_workspace(nel, nf):
    """Output arrays of one assembly at their capacity (Kuu/Muu 36*nel, Kup 6*nel entries).  bench.py's CPU legs
    allocate them once, OUTSIDE the timed region, and pass them as ``workspace=``: the wrapper's own allocations and
    trim copies are not part of the reference's algorithm and must not be charged to it."""
    cap_uu, cap_up = 36 * nel, 6 * nel
    ws = dict(uu=(np.zeros(2 * nf + 1, np.int64), np.zeros(cap_uu, np.int64), np.zeros(cap_uu)),
              mu=(np.zeros(2 * nf + 1, np.int64), np.zeros(cap_uu, np.int64), np.zeros(cap_uu)),
              up=(np.zeros(nel + 1, np.int64), np.zeros(cap_up, np.int64), np.zeros(cap_up)),
              fu=np.zeros(2 * nf), fp=np.zeros(nel), nel=nel, nf=nf)
    for t in (ws["uu"], ws["mu"], ws["up"]):     # touch every page now (np.zeros maps them lazily)
        for x in t:
            x.fill(0)
    return ws


def _assemble(variant, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A=1.0, workspace=None):
    nel, nf = mesh.nel, mesh.nf
    ws = workspace if workspace is not None else None
    if ws is not None:
        assert ws["nel"] == nel and ws["nf"] == nf
        uu, mu, up, fu, fp = ws["uu"], ws["mu"], ws["up"], ws["fu"], ws["fp"]
    else:
        cap_uu, cap_up = 36 * nel, 6 * nel
        uu = (np.zeros(2 * nf + 1, np.int64), np.zeros(cap_uu, np.int64), np.zeros(cap_uu))
        mu = (np.zeros(2 * nf + 1, np.int64), np.zeros(cap_uu, np.int64), np.zeros(cap_uu))
        up = (np.zeros(nel + 1, np.int64), np.zeros(cap_up, np.int64), np.zeros(cap_up))
        fu, fp = np.zeros(2 * nf), np.zeros(nel)
    nnz = [C.c_int64(0) for _ in range(3)]
    times = np.zeros(2)
    a = [_i(mesh.e2f), _i(mesh.bc), _d(mesh.Ω), _d(mesh.Γ), _d(mesh.n_x), _d(mesh.n_y), _d(mesh.ke), _d(mesh.δ),
         _d(mesh.τ), _d(α), _d(β), _d(Ζ), _d(VxDir), _d(VyDir), _d(σxxNeu), _d(σyyNeu), _d(σxyNeu), _d(σyxNeu)]
    rc = lib().oracle_assemble(
        C.c_int(variant), C.c_int(_form(Formulation)), C.c_double(A), C.c_int64(nel), C.c_int64(nf), *map(_p, a),
        _p(uu[0]), _p(uu[1]), _p(uu[2]), C.byref(nnz[0]),
        _p(mu[0]), _p(mu[1]), _p(mu[2]), C.byref(nnz[1]),
        _p(up[0]), _p(up[1]), _p(up[2]), C.byref(nnz[2]),
        _p(fu), _p(fp), _p(times))
    assert rc == 0
    if ws is not None:      # views into the workspace: valid until its next use
        trim = lambda t, n: (t[0], t[1][:n], t[2][:n])
    else:
        trim = lambda t, n: (t[0], t[1][:n].copy(), t[2][:n].copy())
    Kuu, Muu, Kup = trim(uu, nnz[0].value), trim(mu, nnz[1].value), trim(up, nnz[2].value)
    return Kuu, Muu, Kup, fu.reshape(-1, 1), fp, float(times[1]), float(times[0])


def ElementAssemblyLoop(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A=1.0, workspace=None):
    """a5+a7 -- returns Kuu, Muu, Kup (CSC triples), fu (2nf x 1), fp (nel), tsparse, tloop."""
    return _assemble(0, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A, workspace)


def ElementAssemblyLoopNEW(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A=1.0, workspace=None):
    """a6+a7."""
    return _assemble(1, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A, workspace)


def ComputeResidualsFCFV_Stokes_o1(Vxh, Vyh, Pe, mesh, ae, be, ze, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu,
                                   SyxNeu, Formulation, tau_mode="REFERENCE", arrays=False):
    """a8 -- returns the six RMS norms in print order; with ``arrays=True`` also a dict of
    F_nodes_x, F_nodes_y, F_glob2, F_eq1, F_eq2, F_eq3."""
    nel, nf = mesh.nel, mesh.nf
    norms = np.zeros(6)
    out = dict(F_nodes_x=np.zeros(nf), F_nodes_y=np.zeros(nf), F_glob2=np.zeros(nel), F_eq1=np.zeros(4 * nel),
               F_eq2=np.zeros(2 * nel), F_eq3=np.zeros(nel))
    a = [_i(mesh.e2f), _i(mesh.bc), _d(mesh.Ω), _d(mesh.Γ), _d(mesh.n_x), _d(mesh.n_y), _d(mesh.ke), _d(mesh.δ),
         _d(mesh.τ), _d(ae), _d(be), _d(ze), _d(Vxh), _d(Vyh), _d(Pe), _d(sex), _d(sey), _d(VxDir), _d(VyDir),
         _d(SxxNeu), _d(SyyNeu), _d(SxyNeu), _d(SyxNeu)]
    tm = {"REFERENCE": 0, "FACE": 1}[tau_mode]
    rc = lib().oracle_residuals(C.c_int(_form(Formulation)), C.c_int(tm), C.c_int64(nel), C.c_int64(nf),
                                *map(_p, a), _p(norms), *[_p(out[k]) for k in
                                                           ("F_nodes_x", "F_nodes_y", "F_glob2", "F_eq1", "F_eq2", "F_eq3")])
    assert rc == 0
    if arrays:
        out["F_eq1"] = out["F_eq1"].reshape((nel, 2, 2), order="F")
        out["F_eq2"] = out["F_eq2"].reshape((nel, 2), order="F")
        return norms, out
    return norms


def ph_residual(mesh, Kuu, Kup, fu, fp, u, p):
    """a9 -- ru, rp, (norm(ru), norm(rp)) as in SolversFCFV_Stokes.jl:40-42."""
    nel, nf = mesh.nel, mesh.nf
    ru, rp, nrm = np.zeros(2 * nf), np.zeros(nel), np.zeros(2)
    rc = lib().oracle_ph_residual(C.c_int64(nel), C.c_int64(nf), _p(Kuu[0]), _p(Kuu[1]), _p(Kuu[2]), _p(Kup[0]),
                                  _p(Kup[1]), _p(Kup[2]), _p(_d(fu)), _p(_d(fp)), _p(_d(u)), _p(_d(p)),
                                  _p(ru), _p(rp), _p(nrm))
    assert rc == 0
    return ru, rp, nrm


def ph_update_p(mesh, Kup, fp, penalty, u, p):
    """p .+= Kppi*(fp .- Kpu*u)  (SolversFCFV_Stokes.jl:23,49); returns the new p."""
    p = _d(p).copy()
    rc = lib().oracle_ph_update_p(C.c_int64(mesh.nel), _p(Kup[0]), _p(Kup[1]), _p(Kup[2]), _p(_d(fp)), _p(_d(mesh.ke)),
                                  _p(_d(mesh.Ω)), C.c_double(penalty), _p(_d(u)), _p(p))
    assert rc == 0
    return p


def ph_fusc(mesh, Kup, fu, fp, penalty, p):
    """fusc = fu - Kup*(Kppi*fp + p)  (SolversFCFV_Stokes.jl:47)."""
    out = np.zeros(2 * mesh.nf)
    rc = lib().oracle_ph_fusc(C.c_int64(mesh.nel), C.c_int64(mesh.nf), _p(Kup[0]), _p(Kup[1]), _p(Kup[2]),
                              _p(_d(fu)), _p(_d(fp)), _p(_d(mesh.ke)), _p(_d(mesh.Ω)), C.c_double(penalty), _p(_d(p)),
                              _p(out))
    assert rc == 0
    return out


def ComputeElementValues(mesh, Vxh, Vyh, Pe, α, β, Ζ, VxDir, VyDir, Formulation):
    """DiscretisationFCFV_Stokes.jl:52-94."""
    nel = mesh.nel
    o = [np.zeros(nel) for _ in range(5)]
    a = [_i(mesh.e2f), _i(mesh.bc), _d(mesh.Ω), _d(mesh.Γ), _d(mesh.n_x), _d(mesh.n_y), _d(mesh.ke), _d(mesh.δ),
         _d(mesh.τ), _d(α), _d(β), _d(Ζ), _d(Vxh), _d(Vyh)]
    rc = lib().oracle_element_values(C.c_int64(nel), *map(_p, a), *map(_p, o))
    assert rc == 0
    return tuple(o)


def schur_complement(mesh, Kuu, Kup, penalty):
    """Kuusc = Kuu .- Kup*(Kppi*Kpu), Kppi = spdiagm(penalty.*ke./Ω), Kpu = -Kup' (SolversFCFV_Stokes.jl:6,23-25).
    Kuu, Kup: Julia CSC triples; returns the CSC triple of Kuusc."""
    nel, n = mesh.nel, 2 * mesh.nf
    cap = len(Kuu[2]) + 36 * nel
    cp, rv, v = np.zeros(n + 1, np.int64), np.zeros(cap, np.int64), np.zeros(cap)
    nnz = C.c_int64(0)
    rc = lib().oracle_schur(C.c_int64(n), C.c_int64(nel), _p(_i(Kuu[0])), _p(_i(Kuu[1])), _p(_d(Kuu[2])), _p(_i(Kup[0])),
                            _p(_i(Kup[1])), _p(_d(Kup[2])), _p(_d(mesh.ke)), _p(_d(mesh.Ω)), C.c_double(penalty),
                            _p(cp), _p(rv), _p(v), C.byref(nnz))
    assert rc == 0
    return cp, rv[:nnz.value].copy(), v[:nnz.value].copy()


def csc_to_scipy(csc, shape):
    """Julia-style CSC triple -> scipy.sparse.csc_matrix (test helper)."""
    import scipy.sparse as sp
    colptr, rowval, nzval = csc
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=shape)
