"""Second, independent restatement of the reference path in vectorised NumPy/SciPy.

TEST INFRASTRUCTURE.  Written separately from ``oracle/fcfv_oracle.c`` (array-at-a-time
instead of scalar loops, SciPy's COO->CSC instead of a hand-written counting sort) so that
the two restatements check each other; both follow the Julia statements of
src/DiscretisationFCFV_Stokes.jl and src/ComputeResidualsFCFV_Stokes.jl.
NumPy evaluates ``a*b*c`` left to right without FMA, ``False*x`` is ``0.0*x`` (same sign rule
as Julia's strong zero for finite x), so values agree bitwise with the C oracle.
"""
import numpy as np
import scipy.sparse as sp


def _face_arrays(mesh):
    f = np.asarray(mesh.e2f) - 1                       # nel x 3, 0-based
    bc = np.asarray(mesh.bc)[f]                        # nel x 3
    return f, bc


def compute_fcfv(mesh, sex, sey, VxDir, VyDir, formulation, A=1.0):
    """DiscretisationFCFV_Stokes.jl:8-44."""
    nel = mesh.nel
    f, bc = _face_arrays(mesh)
    d = (bc == 1)
    G, nx, ny = np.asarray(mesh.Γ), np.asarray(mesh.n_x), np.asarray(mesh.n_y)
    te = np.asarray(mesh.τe)[:nel]
    vx, vy = np.asarray(VxDir)[f], np.asarray(VyDir)[f]
    alpha = np.zeros(nel)
    beta = np.zeros((nel, 2))
    Z = np.zeros((nel, 2, 2))
    beta[:, 0] += mesh.Ω * sex
    beta[:, 1] += mesh.Ω * sey
    tau = np.zeros(mesh.nf) if mesh.τ is None else np.array(mesh.τ, dtype=float)
    for i in range(3):
        dG = d[:, i] * G[:, i]
        if formulation == "Gradient":
            Z[:, 0, 0] += dG * nx[:, i] * vx[:, i]
            Z[:, 0, 1] += dG * nx[:, i] * vy[:, i]
            Z[:, 1, 0] += dG * ny[:, i] * vx[:, i]
            Z[:, 1, 1] += dG * ny[:, i] * vy[:, i]
        else:
            Z[:, 0, 0] += dG * (nx[:, i] * vx[:, i] + A * nx[:, i] * vx[:, i])
            Z[:, 0, 1] += dG * (nx[:, i] * vy[:, i] + A * ny[:, i] * vx[:, i])
            Z[:, 1, 0] += dG * (ny[:, i] * vx[:, i] + A * nx[:, i] * vy[:, i])
            Z[:, 1, 1] += dG * (ny[:, i] * vy[:, i] + A * ny[:, i] * vy[:, i])
        beta[:, 0] += dG * te * vx[:, i]
        beta[:, 1] += dG * te * vy[:, i]
        alpha += G[:, i] * te
    # last writer wins: element order, then face order
    for i in range(3):
        pass
    order_e = np.repeat(np.arange(nel), 3)
    tau[f.ravel()] = te[order_e]          # NumPy fancy assignment keeps the LAST duplicate
    return alpha, beta, Z, tau


def assemble(mesh, alpha, beta, Z, VxDir, VyDir, sxx, syy, sxy, syx, formulation, variant=0, A=1.0):
    """ElementAssemblyLoop (variant 0, :102-202) / ElementAssemblyLoopNEW (variant 1, :210-347)
    followed by Sparsify (:355-364).  Returns scipy CSC Kuu, Muu, Kup and dense fu, fp."""
    nel, nf = mesh.nel, mesh.nf
    f, bc = _face_arrays(mesh)
    G, nx, ny = np.asarray(mesh.Γ), np.asarray(mesh.n_x), np.asarray(mesh.n_y)
    tau = np.asarray(mesh.τ)[f]
    eta, Om = np.asarray(mesh.ke), np.asarray(mesh.Ω)
    ia = 1.0 / alpha
    rows, cols, vals, mvals = [], [], [], []
    if variant == 1:
        dl = np.asarray(mesh.δ)
        assert np.all(dl == 1.0), "vectorised restatement covers δ == 1 only"
        D11 = (eta * 1.0) / Om
        D12 = ((eta / dl) * 1.0) / Om
        D21, D22 = D12, D11
    else:
        h = eta * (1.0 / Om)
    fu = np.zeros(2 * nf)
    fp = np.zeros(nel)
    kup_r, kup_c, kup_v = [], [], []
    for i in range(3):
        bci = bc[:, i]
        J = 0.0 + (bci == -1) * 1.0
        for j in range(3):
            on = (bci != 1) & (bc[:, j] != 1)
            mG = on * -G[:, i]
            ninj = nx[:, i] * nx[:, j] + ny[:, i] * ny[:, j]
            c = J if formulation == "Gradient" else A
            t1 = ia * tau[:, i] * tau[:, j] * G[:, j]
            t3 = tau[:, i] * (1.0 if i == j else 0.0)
            if variant == 0:
                hG = h * G[:, j]
                vxx = mG * (t1 - hG * (ninj + c * nx[:, i] * nx[:, j]) - t3)
                vxy = mG * (-eta * (1.0 / Om) * G[:, j] * (c * ny[:, i] * nx[:, j]))   # row (i,x), col (j,y)
                vyx = mG * (-eta * (1.0 / Om) * G[:, j] * (c * nx[:, i] * ny[:, j]))   # row (i,y), col (j,x)
                vyy = mG * (t1 - hG * (ninj + c * ny[:, i] * ny[:, j]) - t3)
                m = mG * (t1 - hG * ninj - t3)
                mxx, myy = m, m
                r, cc = f[:, i], f[:, j]
            else:
                vxx = mG * (t1 - D11 * G[:, j] * (ninj + c * nx[:, i] * nx[:, j]) - t3)
                vyy = mG * (t1 - D22 * G[:, j] * (ninj + c * ny[:, i] * ny[:, j]) - t3)
                if formulation == "Gradient":
                    # values stored at [j,i], indices at [i,j] (:273-276 vs :299-302) => transpose
                    vyx_T = mG * (-D12 * G[:, j] * (c * nx[:, i] * ny[:, j]))  # Kuuv[j, i+nf] -> row (j,x), col (i,y)
                    vxy_T = mG * (-D21 * G[:, j] * (c * ny[:, i] * nx[:, j]))  # Kuuv[j+nf, i] -> row (j,y), col (i,x)
                    r, cc = f[:, j], f[:, i]
                    vxy, vyx = vyx_T, vxy_T
                else:
                    vxy = mG * (-D12 * G[:, j] * (c * ny[:, i] * nx[:, j]))    # Kuuv[i, j+nf]
                    vyx = mG * (-D21 * G[:, j] * (c * nx[:, i] * ny[:, j]))    # Kuuv[i+nf, j]
                    r, cc = f[:, i], f[:, j]
                mxx = mG * (t1 - D11 * G[:, j] * ninj - t3)
                myy = mG * (t1 - D22 * G[:, j] * ninj - t3)
            if i == j:
                dir1 = (bci == 1) * 1.0
                vxx = vxx + dir1; vyy = vyy + dir1; mxx = mxx + dir1; myy = myy + dir1
            rows += [r, r, r + nf, r + nf]
            cols += [cc, cc + nf, cc, cc + nf]
            vals += [vxx, vxy, vyx, vyy]
            # Muu is always written at [j,i]; NEW indexes at [i,j] => transposed there
            mr, mc = (f[:, i], f[:, j]) if variant == 0 else (f[:, j], f[:, i])
            mvals.append((mr, mc, mxx, myy))
        Xi = 0.0 + (bci == 2) * 1.0
        fi = f[:, i]
        tix = nx[:, i] * sxx[fi] + ny[:, i] * sxy[fi]
        tiy = nx[:, i] * syx[fi] + ny[:, i] * syy[fi]
        mGi = (bci != 1) * -G[:, i]
        if variant == 0:
            if formulation == "Gradient":
                nZx = nx[:, i] * (Z[:, 0, 0] + J * Z[:, 0, 0]) + ny[:, i] * (Z[:, 1, 0] + J * Z[:, 0, 1])
                nZy = nx[:, i] * (Z[:, 0, 1] + J * Z[:, 1, 0]) + ny[:, i] * (Z[:, 1, 1] + J * Z[:, 1, 1])
            else:
                nZx = nx[:, i] * Z[:, 0, 0] + ny[:, i] * Z[:, 1, 0]
                nZy = nx[:, i] * Z[:, 0, 1] + ny[:, i] * Z[:, 1, 1]
            fex = mGi * (-ia * tau[:, i] * beta[:, 0] + eta * (1.0 / Om) * nZx - tix * Xi)
            fey = mGi * (-ia * tau[:, i] * beta[:, 1] + eta * (1.0 / Om) * nZy - tiy * Xi)
        else:
            fex = mGi * (-ia * tau[:, i] * beta[:, 0] + (D11 * nx[:, i] * Z[:, 0, 0] + D11 * ny[:, i] * Z[:, 1, 0]) - tix * Xi)
            fey = mGi * (-ia * tau[:, i] * beta[:, 1] + (D22 * nx[:, i] * Z[:, 0, 1] + D22 * ny[:, i] * Z[:, 1, 1]) - tiy * Xi)
        np.add.at(fu, fi, 0.0 + ((bci != 1) * fex + (bci == 1) * VxDir[fi] * 1.0))
        np.add.at(fu, fi + nf, 0.0 + ((bci != 1) * fey + (bci == 1) * VyDir[fi] * 1.0))
        kup_r += [fi, fi + nf]
        kup_c += [np.arange(nel), np.arange(nel)]
        kup_v += [0.0 - (bci != 1) * G[:, i] * nx[:, i], 0.0 - (bci != 1) * G[:, i] * ny[:, i]]
        fp = fp - (bci == 1) * G[:, i] * (VxDir[fi] * nx[:, i] + VyDir[fi] * ny[:, i])

    def build(r, c, v, shape):
        M = sp.coo_matrix((np.concatenate(v), (np.concatenate(r), np.concatenate(c))), shape=shape).tocsc()
        M.sum_duplicates()
        M.eliminate_zeros()
        M.sort_indices()
        return M

    Kuu = build(rows, cols, vals, (2 * nf, 2 * nf))
    Kup = build(kup_r, kup_c, kup_v, (2 * nf, nel))
    mr = [t[0] for t in mvals] + [t[0] + nf for t in mvals]
    mc = [t[1] for t in mvals] + [t[1] + nf for t in mvals]
    mv = [t[2] for t in mvals] + [t[3] for t in mvals]
    Muu = build(mr, mc, mv, (2 * nf, 2 * nf))
    fu = np.where(fu == 0.0, 0.0, fu)     # Array(dropzeros(sparse(..))): stored -0.0 becomes +0.0
    return Kuu, Muu, Kup, fu, fp


def residual_arrays(mesh, Vxh, Vyh, Pe, ae, be, ze, sex, sey, VxDir, VyDir, Sxx, Syy, Sxy, Syx, formulation,
                    tau_mode="REFERENCE"):
    """ComputeResidualsFCFV_Stokes.jl:5-108, array-at-a-time."""
    nel, nf = mesh.nel, mesh.nf
    f, bc = _face_arrays(mesh)
    G, nx, ny = np.asarray(mesh.Γ), np.asarray(mesh.n_x), np.asarray(mesh.n_y)
    Om, eta, dl = np.asarray(mesh.Ω), np.asarray(mesh.ke), np.asarray(mesh.δ)
    tau_f = np.asarray(mesh.τ)
    tau = tau_f[f] if tau_mode == "FACE" else np.repeat(tau_f[:nel, None], 3, axis=1)
    n = np.stack([nx, ny], axis=-1)                                  # nel x 3 x 2
    uh = np.stack([np.asarray(Vxh)[f], np.asarray(Vyh)[f]], axis=-1)
    ud = np.stack([np.asarray(VxDir)[f], np.asarray(VyDir)[f]], axis=-1)
    nd = (bc != 1)
    ue = be / ae[:, None]
    Le = -1.0 / Om[:, None, None] * ze
    for i in range(3):
        w = nd[:, i]
        ue = ue + w[:, None] * (G[:, i] * tau[:, i])[:, None] * uh[:, i] / ae[:, None]
        M = n[:, i, :, None] * uh[:, i, None, :]
        if formulation == "SymmetricGradient":
            M = M + uh[:, i, :, None] * n[:, i, None, :]
        Le = Le - w[:, None, None] * (1.0 / Om * G[:, i])[:, None, None] * M
    D = np.empty((nel, 2, 2))
    D[:, 0, 0] = eta; D[:, 1, 1] = eta; D[:, 0, 1] = eta / dl; D[:, 1, 0] = eta / dl
    DL = D * Le
    F1 = Om[:, None, None] * Le
    F2 = np.stack([-sex * Om, -sey * Om], axis=-1)
    F3 = np.zeros(nel)
    Fg2 = np.zeros(nel)
    Fnx, Fny = np.zeros(nf), np.zeros(nf)
    for i in range(3):
        fi = f[:, i]
        Xi = (bc[:, i] == 2) * 1.0
        Ji = (bc[:, i] == -1) * 1.0
        t = np.stack([nx[:, i] * Sxx[fi] + ny[:, i] * Sxy[fi], nx[:, i] * Syx[fi] + ny[:, i] * Syy[fi]], axis=-1)
        u = np.where((bc[:, i] == 1)[:, None], ud[:, i], uh[:, i])
        nDL = np.einsum("ea,eab->eb", n[:, i], DL)
        acc = nDL + Pe[:, None] * n[:, i] + tau[:, i, None] * ue - tau[:, i, None] * u + t * Xi[:, None]
        if formulation == "Gradient":
            acc = acc + Ji[:, None] * np.einsum("ea,eba->eb", n[:, i], DL)
        Fg1 = G[:, i, None] * acc
        F2 = F2 + (G[:, i] * tau[:, i])[:, None] * ue - (G[:, i] * tau[:, i])[:, None] * u
        # dAi * u'*n = (dAi*u')*n (:75, :82); dAi * n * u' = (dAi*n)*u' (:77); dAi * (n*u' .+ u*n') (:79)
        Gu = G[:, i, None] * u
        un = Gu[:, 0] * n[:, i, 0] + Gu[:, 1] * n[:, i, 1]
        Fg2 = Fg2 + un
        if formulation == "SymmetricGradient":
            F1 = F1 + G[:, i, None, None] * (n[:, i, :, None] * u[:, None, :] + u[:, :, None] * n[:, i, None, :])
        else:
            F1 = F1 + (G[:, i, None] * n[:, i])[:, :, None] * u[:, None, :]
        F3 = F3 + un
        w = nd[:, i]
        np.add.at(Fnx, fi[w], Fg1[w, 0])
        np.add.at(Fny, fi[w], Fg1[w, 1])
    return dict(F_eq1=F1, F_eq2=F2, F_eq3=F3, F_nodes_x=Fnx, F_nodes_y=Fny, F_glob2=Fg2)


def residual_norms(arrs):
    r = lambda x: np.linalg.norm(np.ravel(x)) / np.sqrt(np.size(x))
    return np.array([r(arrs[k]) for k in ("F_eq1", "F_eq2", "F_eq3", "F_nodes_x", "F_nodes_y", "F_glob2")])


def element_values(mesh, Vxh, Vyh, alpha, beta, Z):
    """Independent NumPy restatement of ComputeElementValues (DiscretisationFCFV_Stokes.jl:52-94), vectorised over
    elements, faces visited in order i = 1,2,3 like the reference."""
    e2f = np.asarray(mesh.e2f, dtype=np.int64) - 1
    nd = (np.asarray(mesh.bc)[e2f] != 1).astype(float)            # (bc != 1), nel x 3
    tau = np.asarray(mesh.τ)[e2f]
    ux, uy = np.asarray(Vxh)[e2f], np.asarray(Vyh)[e2f]
    eta, Om, dl = np.asarray(mesh.ke), np.asarray(mesh.Ω), np.asarray(mesh.δ)
    G, nx, ny = np.asarray(mesh.Γ), np.asarray(mesh.n_x), np.asarray(mesh.n_y)
    Vx, Vy = beta[:, 0] / alpha, beta[:, 1] / alpha
    Txx, Tyy = eta / Om * Z[:, 0, 0], eta / Om * Z[:, 1, 1]
    Txy = eta / Om * 0.5 * (Z[:, 0, 1] + Z[:, 1, 0])
    for i in range(3):
        Vx = Vx + nd[:, i] * G[:, i] * tau[:, i] * ux[:, i] / alpha
        Vy = Vy + nd[:, i] * G[:, i] * tau[:, i] * uy[:, i] / alpha
        Txx = Txx + nd[:, i] * eta / Om * G[:, i] * nx[:, i] * ux[:, i]
        Tyy = Tyy + nd[:, i] * eta / Om * G[:, i] * ny[:, i] * uy[:, i]
        Txy = Txy + nd[:, i] * (eta / dl) * 0.5 * (1.0 / Om * G[:, i] * (nx[:, i] * uy[:, i] + ny[:, i] * ux[:, i]))
    return Vx, Vy, 2.0 * Txx, 2.0 * Tyy, 2.0 * Txy
