"""Seeded random unstructured meshes (Delaunay triangulations of random points) with random numbering, random
boundary / interface tags and random fields: irregular valence, faces in no particular order, faces no element
references, Neumann and interface faces anywhere.  The fixtures of the reference (`.mat`, `.dat`, S(n)) are all
regular in one way or another; these are not.  CPU: the kernels' arithmetic compiled for the host against the
oracle, bit for bit.  GPU: the CUDA path through the C ABI against the oracle."""
import numpy as np
import pytest
from scipy.spatial import Delaunay

import cases
from fcfv_nme23_b200 import mesh as M
from helpers import assert_csr_equal, csc_triple_to_csr
from test_host_emulation import emu_assemble, emu_local

SEEDS = [1, 2, 3, 4, 5, 6]
TOL = 1e-12


def random_case(seed, aniso=False):
    """(mesh, data): mesh with oracle geometry; the numbering, tags and fields are drawn from default_rng(seed)."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    npts = int(rng.integers(20, 160))
    pts = rng.random((npts, 2))
    if seed % 2 == 0:                                   # stretched domain: badly shaped triangles
        pts[:, 1] *= 0.05
    tri = Delaunay(pts)
    e2v = tri.simplices.astype(np.int64) + 1
    e2v = e2v[rng.permutation(e2v.shape[0])]            # element order without any coherence
    xv, yv = pts[:, 0].copy(), pts[:, 1].copy()
    e2f, xf, yf = M.number_faces_first_seen(xv, yv, e2v)
    nel, nf_used = e2v.shape[0], int(e2f.max())
    cnt = np.bincount(np.asarray(e2f).ravel() - 1, minlength=nf_used)
    bc = np.zeros(nf_used, dtype=np.int64)
    boundary = cnt == 1
    bc[boundary] = np.where(rng.random(boundary.sum()) < 0.7, 1, 2)          # Dirichlet / Neumann
    if seed == 3:
        bc[boundary] = 2                                                        # no Dirichlet face at all
    bc[~boundary] = np.where(rng.random((~boundary).sum()) < 0.15, -1, 0)     # material interfaces
    m = M._finish_square_mesh(xv, yv, e2v, e2f, xf, yf, bc)
    O.mesh_geometry(m, 1.0)
    # faces nobody references, then a random face numbering
    extra = int(rng.integers(0, 4))
    nf = nf_used + extra
    m.nf = nf
    m.bc = np.concatenate([m.bc, rng.integers(-1, 3, extra)]).astype(np.int64)
    m.xf = np.concatenate([m.xf, rng.random(extra)]); m.yf = np.concatenate([m.yf, rng.random(extra)])
    m.τ = np.full(nf, 7.25)                           # faces without elements keep this value
    M.renumber_faces(m, rng.permutation(nf) + 1)
    m.ke = np.exp(rng.normal(0.0, 2.0, nel))
    m.τe = np.exp(rng.normal(0.0, 1.0, nel)) * np.maximum(m.ke, 1.0)
    m.δ = np.where(rng.random(nel) < 0.5, 1.0, 1.0 + 3.0 * rng.random(nel)) if aniso else np.ones(nel)
    d = dict(sex=rng.standard_normal(nel), sey=rng.standard_normal(nel), VxDir=rng.standard_normal(nf),
             VyDir=rng.standard_normal(nf), neu=[rng.standard_normal(nf) for _ in range(4)],
             Vxh=rng.standard_normal(nf), Vyh=rng.standard_normal(nf), Pe=rng.standard_normal(nel))
    return m, d


def test_random_case_is_what_it_claims():
    m, d = random_case(4)
    f = np.asarray(m.e2f) - 1
    assert (np.bincount(f.ravel(), minlength=m.nf) <= 2).all()
    assert not np.array_equal(M.first_seen_face_permutation(m.e2f, m.nf), np.arange(1, m.nf + 1))
    assert set(np.unique(m.bc)) >= {-1, 0, 1, 2}
    assert np.all(m.Ω > 0) and np.all(m.Γ > 0)


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("form", ["Gradient", "SymmetricGradient"])
def test_emulation_matches_oracle(seed, variant, form, oracle, emu):
    m, d = random_case(seed)
    m2, _ = random_case(seed)
    nel, nf = m.nel, m.nf
    neu = d["neu"]
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, 1.0, form)
    ea, eb, eZ, etau = emu_local(emu, m2, d, form)
    assert np.array_equal(ea, a) and np.array_equal(eb, b) and np.array_equal(eZ, Z)
    used = np.zeros(nf, bool); used[np.asarray(m.e2f).ravel() - 1] = True
    assert np.array_equal(etau[used], m.τ[used])
    etau[~used] = m.τ[~used]
    fn = oracle.ElementAssemblyLoop if variant == 0 else oracle.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, form)
    m2.τ = etau
    euu, eup, emu_, efu, efp = emu_assemble(emu, m2, ea, eb, eZ, d, variant, form)
    assert_csr_equal(euu, csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu")
    assert_csr_equal(eup, csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup")
    assert_csr_equal(emu_, csc_triple_to_csr(Muu, (2 * nf, 2 * nf)), "Muu")
    assert np.array_equal(efu, fu.ravel()) and np.array_equal(efp, fp)
    ref_sc = oracle.schur_complement(m, Kuu, Kup, 1.0e3)
    suu, _, ssc, _, _ = emu_assemble(emu, m2, ea, eb, eZ, d, variant, form, schur=1.0e3)
    assert_csr_equal(suu, csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu (schur run)")
    assert_csr_equal(ssc, csc_triple_to_csr(ref_sc, (2 * nf, 2 * nf)), "Kuusc")


@pytest.mark.parametrize("seed", SEEDS[:3])
@pytest.mark.parametrize("form", ["Gradient", "SymmetricGradient"])
def test_emulation_matches_oracle_anisotropic(seed, form, oracle, emu):
    """delta != 1 on half of the elements (NEW loop; 2x2 inverses through LAPACK in the reference: <= 1e-12)."""
    m, d = random_case(seed, aniso=True)
    m2, _ = random_case(seed, aniso=True)
    nel, nf = m.nel, m.nf
    neu = d["neu"]
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, 1.0, form)
    ea, eb, eZ, etau = emu_local(emu, m2, d, form)
    Kuu, Muu, Kup, fu, fp, _, _ = oracle.ElementAssemblyLoopNEW(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, form)
    used = np.zeros(nf, bool); used[np.asarray(m.e2f).ravel() - 1] = True
    etau[~used] = m.τ[~used]
    m2.τ = etau
    euu, eup, emu_, efu, efp = emu_assemble(emu, m2, ea, eb, eZ, d, 1, form)
    assert_csr_equal(euu, csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu", bitwise=False, rtol=TOL)
    assert_csr_equal(eup, csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup")
    assert_csr_equal(emu_, csc_triple_to_csr(Muu, (2 * nf, 2 * nf)), "Muu", bitwise=False, rtol=TOL)
    assert np.allclose(efu, fu.ravel(), rtol=TOL, atol=0.0) and np.array_equal(efp, fp)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("form", ["Gradient", "SymmetricGradient"])
def test_cuda_matches_oracle(seed, variant, form, oracle):
    from fcfv_nme23_b200 import api, lib as L
    m, d = random_case(seed)
    g, _ = random_case(seed)
    nel, nf = m.nel, m.nf
    neu = d["neu"]
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, 1.0, form)
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, 1.0, form)
    assert np.array_equal(ga, a) and np.array_equal(gb, b) and np.array_equal(gZ, Z)
    assert np.array_equal(g.τ, m.τ)                  # faces without elements keep their value on both sides
    fn = oracle.ElementAssemblyLoop if variant == 0 else oracle.ElementAssemblyLoopNEW
    gfn = api.ElementAssemblyLoop if variant == 0 else api.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, form)
    gK, gM, gP, gfu, gfp, _ = gfn(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, form)
    h = api.handle_for(g)
    assert_csr_equal(h.export_csr(L.KUU), csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu")
    assert_csr_equal(h.export_csr(L.KUP), csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup")
    assert_csr_equal(h.export_csr(L.MUU), csc_triple_to_csr(Muu, (2 * nf, 2 * nf)), "Muu")
    assert np.array_equal(np.asarray(gfu).ravel(), fu.ravel()) and np.array_equal(gfp, fp)
    # the exported CSC is what Julia would hold
    assert np.array_equal(gK.indptr, Kuu[0] - 1) and np.array_equal(gK.indices, Kuu[1] - 1) and np.array_equal(gK.data, Kuu[2])
    ref_sc = oracle.schur_complement(m, Kuu, Kup, 1.0e3)
    api.SchurComplement(g, 1.0e3)
    assert_csr_equal(h.export_csr(L.KUUSC), csc_triple_to_csr(ref_sc, (2 * nf, 2 * nf)), "Kuusc")
    for tau_mode in ("REFERENCE", "FACE"):
        on = oracle.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"],
                                                   d["VyDir"], *neu, form, tau_mode=tau_mode)
        gn = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], g, ga, gb, gZ, d["sex"], d["sey"], d["VxDir"],
                                                d["VyDir"], *neu, form, tau_mode=tau_mode, verbose=False)
        big = on > 1e-9
        assert np.all(np.abs(gn[big] - on[big]) <= TOL * on[big]) and np.all(gn[~big] < 1e-9)
    ev = oracle.ComputeElementValues(m, d["Vxh"], d["Vyh"], d["Pe"], a, b, Z, d["VxDir"], d["VyDir"], form)
    gev = api.ComputeElementValues(g, d["Vxh"], d["Vyh"], d["Pe"], ga, gb, gZ, d["VxDir"], d["VyDir"], form)
    for k in range(5):
        assert np.array_equal(gev[k], ev[k])
    h.close()
