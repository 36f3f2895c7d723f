"""Parity of the CUDA path (through the C ABI, via the ctypes mirror of the Julia shim) against the
oracle on identical seeded inputs.  Contract (BASELINE.json north_star): CSR sparsity and face
numbering bit-exact; matrix, RHS and residual values within 1e-12 relative.  Design goal checked
here: alpha/beta/Z/tau/Kuu/Kup/Muu/fu/fp are BITWISE equal; residual norms within 1e-12."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import cases
from helpers import assert_csr_equal, csc_triple_to_csr, rel_inf

pytestmark = pytest.mark.gpu
TOL = 1e-12   # relative tolerance stated by north_star for floating-point values


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def api():
    from fcfv_nme23_b200 import api
    return api


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(cases.GOLDEN, "golden_oracle.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ["H1", "H2", "S6", "S17"])
def test_geometry_bitwise(name, api, oracle):
    m = cases.load_mesh(name)                      # geometry from the oracle
    ref = {k: getattr(m, k).copy() for k in ("Ω", "xc", "yc", "n_x", "n_y", "Γ")}
    for k in ref:
        setattr(m, k, None)
    m.τ = None
    api.mesh_geometry(m, 3.5)
    for k, v in ref.items():
        assert np.array_equal(getattr(m, k), v), k
    assert np.all(m.τ == 3.5)
    assert abs(m.Ω.sum() - 1.0) < 1e-12            # SolKzFCFV.jl:86 sanity check


@pytest.mark.parametrize("case", cases.all_cases(), ids=lambda c: c.name)
def test_local_assembly_residual(case, api, oracle, golden):
    m, d = cases.build(case)       # oracle side
    g, _ = cases.build(case)       # CUDA side (same seeded data)
    nel, nf = m.nel, m.nf
    neu = d["neu"]
    # ---- ComputeFCFV
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    assert np.array_equal(ga, a) and np.array_equal(gb, b) and np.array_equal(gZ, Z)
    assert np.array_equal(g.τ, m.τ)
    # ---- assembly
    ofn = oracle.ElementAssemblyLoop if case.variant == 0 else oracle.ElementAssemblyLoopNEW
    gfn = api.ElementAssemblyLoop if case.variant == 0 else api.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = ofn(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form)
    gKuu, gMuu, gKup, gfu, gfp, tsparse = gfn(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, case.form)
    assert tsparse > 0.0 and gfu.shape == (2 * nf, 1) and gfp.shape == (nel,)
    h = api.handle_for(g)
    from fcfv_nme23_b200 import lib as L
    for blk, ref, shape, what in ((L.KUU, Kuu, (2 * nf, 2 * nf), "Kuu"), (L.KUP, Kup, (2 * nf, nel), "Kup"),
                                  (L.MUU, Muu, (2 * nf, 2 * nf), "Muu")):
        # device CSR against CSR of the oracle's Julia-style CSC
        assert_csr_equal(h.export_csr(blk), csc_triple_to_csr(ref, shape), what)
        # Julia-facing CSC export: colptr / rowval / nzval identical to the oracle's SparseMatrixCSC fields
        cp, rv, v = h.export_csc(blk)
        assert cp.dtype == np.int64 and rv.dtype == np.int64
        assert np.array_equal(cp, ref[0]) and np.array_equal(rv, ref[1]) and np.array_equal(v, ref[2]), what + " CSC"
    assert np.array_equal(gfu, fu) and np.array_equal(gfp, fp)
    assert (gKuu != h.export_scipy_csc(L.KUU)).nnz == 0
    # ---- golden digests (pin the oracle; the CUDA path must hash to the same bytes)
    gold = golden[case.name]
    assert (gold["nnz_uu"], gold["nnz_up"], gold["nnz_muu"]) == (gKuu.nnz, gKup.nnz, gMuu.nnz)
    assert gold["local"] == digest(ga, gb.ravel(order="F"), gZ.ravel(order="F"), g.τ)
    assert gold["Kuu"] == digest(*h.export_csc(L.KUU)) and gold["Kup"] == digest(*h.export_csc(L.KUP))
    assert gold["Muu"] == digest(*h.export_csc(L.MUU)) and gold["rhs"] == digest(gfu, gfp)
    # ---- residuals
    for tau_mode in ("REFERENCE", "FACE"):
        norms, arr = oracle.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"],
                                                           d["VxDir"], d["VyDir"], *neu, case.form, tau_mode=tau_mode,
                                                           arrays=True)
        gn, garr = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], g, ga, gb, gZ, d["sex"], d["sey"],
                                                      d["VxDir"], d["VyDir"], *neu, case.form, tau_mode=tau_mode,
                                                      verbose=False, arrays=True)
        for k in range(6):
            if norms[k] > 1e-9:
                assert abs(gn[k] - norms[k]) <= TOL * abs(norms[k]), (tau_mode, k, gn[k], norms[k])
            else:
                assert gn[k] < 1e-9
        for k in ("F_nodes_x", "F_nodes_y", "F_glob2"):
            assert rel_inf(garr[k], arr[k]) <= TOL, (tau_mode, k)
    # ---- Schur complement fused into the assembly: fields of Julia's SparseMatrixCSC, bit for bit
    ref_sc = oracle.schur_complement(m, Kuu, Kup, 1.0e4)
    gsc = api.SchurComplement(g, 1.0e4)
    cp, rv, v = h.export_csc(L.KUUSC)
    assert np.array_equal(cp, ref_sc[0]) and np.array_equal(rv, ref_sc[1]) and np.array_equal(v, ref_sc[2]), "Kuusc CSC"
    assert gsc.nnz == len(ref_sc[2]) == gold["nnz_sc"] and gold["Kuusc"] == digest(cp, rv, v)
    assert_csr_equal(h.export_csr(L.KUU), csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu after schur")
    assert_csr_equal(h.export_csr(L.KUP), csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup after schur")
    # ---- ComputeElementValues (bitwise: same operation order, no BLAS on either side)
    ev = oracle.ComputeElementValues(m, d["Vxh"], d["Vyh"], d["Pe"], a, b, Z, d["VxDir"], d["VyDir"], case.form)
    gev = api.ComputeElementValues(g, d["Vxh"], d["Vyh"], d["Pe"], ga, gb, gZ, d["VxDir"], d["VyDir"], case.form)
    for k in range(5):
        assert np.array_equal(gev[k], ev[k]), ("element values", k)
    assert gold["elem_values"] == digest(*gev)
    # ---- Powell-Hestenes lines on the resident CSR
    u = np.concatenate([d["Vxh"], d["Vyh"]])
    ru, rp, nrm = oracle.ph_residual(m, Kuu, Kup, fu, fp, u, d["Pe"])
    gru, grp, gnrm = api.ph_residual(g, u, d["Pe"])
    assert rel_inf(gru, ru) <= TOL and rel_inf(grp, rp) <= TOL and rel_inf(gnrm, nrm) <= TOL
    assert rel_inf(api.ph_update_p(g, 1e4, u, d["Pe"]), oracle.ph_update_p(m, Kup, fp, 1e4, u, d["Pe"])) <= TOL
    assert rel_inf(api.ph_fusc(g, 1e4, d["Pe"]), oracle.ph_fusc(m, Kup, fu, fp, 1e4, d["Pe"])) <= TOL
    pc, mean = api.center_pressure(g, d["Pe"])          # Solvers:20, Pe .- mean(Pe)
    assert abs(mean - d["Pe"].mean()) <= TOL * max(1.0, abs(d["Pe"]).max()) and rel_inf(pc, d["Pe"] - d["Pe"].mean()) <= TOL
    h.close()


def test_edge_cases(api, oracle):
    from fcfv_nme23_b200 import lib as L
    from fcfv_nme23_b200.mesh import FCFV_Mesh
    # single triangle, all faces Dirichlet: Kuu = identity, Kup empty
    m = FCFV_Mesh(nel=1, nf=3, nv=3, nf_el=3, nn_el=3)
    m.e2v = np.asfortranarray(np.array([[1, 2, 3]], dtype=np.int64)); m.e2f = np.asfortranarray(np.array([[1, 2, 3]], dtype=np.int64))
    m.xv = np.array([0.0, 1.0, 0.0]); m.yv = np.array([0.0, 0.0, 1.0])
    m.xf = np.array([0.5, 0.0, 0.5]); m.yf = np.array([0.5, 0.5, 0.0])   # numbering 0: face1=(v2,v3) face2=(v3,v1) face3=(v1,v2)
    m.bc = np.array([1, 1, 1], dtype=np.int64); m.numbering = 0
    m.ke = np.ones(1); m.δ = np.ones(1); m.τe = np.ones(1)
    api.mesh_geometry(m, 1.0)
    assert abs(m.Ω[0] - 0.5) < 1e-15
    z3, z1 = np.zeros(3), np.zeros(1)
    V = np.array([1.0, 2.0, 3.0])
    a, b, Z = api.ComputeFCFV(m, z1, z1, V, -V, z3, z3, z3, z3, 1.0, ":Gradient")
    Kuu, Muu, Kup, fu, fp, _ = api.ElementAssemblyLoop(m, a, b, Z, V, -V, z3, z3, z3, z3, ":Gradient")
    assert Kuu.nnz == 6 and np.array_equal(Kuu.toarray(), np.eye(6)) and Kup.nnz == 0
    assert np.array_equal(fu.ravel(), np.concatenate([V, -V]))
    # invalid inputs fail loudly with a message, not a crash
    h = api.Handle(0)
    with pytest.raises(L.FCFVError):
        h.run_local("Gradient")                       # no mesh
    with pytest.raises(L.FCFVError):
        h.numbering_locality()                        # no mesh either
    x3, o3 = np.zeros(3), np.zeros(3, dtype=np.int64)
    assert h.lib.fcfv_hilbert_elements(h._h, 3, x3.ctypes.data_as(C.c_void_p), None, o3.ctypes.data_as(C.c_void_p)) != 0   # null array
    assert b"null array" in h.lib.fcfv_last_error(h._h)
    bad = cases.load_mesh("S6"); bad.e2f = bad.e2f.copy(); bad.e2f[0, 0] = bad.nf + 5
    with pytest.raises(L.FCFVError):
        h.set_mesh(bad)
    with pytest.raises(ValueError):
        api.formulation_id("Divergence")
    h.close()


@pytest.mark.parametrize("setting", ["solkz", "inclusion"])
def test_device_generator_matches_host(setting, api, oracle):
    """fcfv_generate_structured: connectivity identical to the host closed form (itself checked
    against the ReadMeshesDat first-seen rule), geometry bitwise equal to the oracle geometry."""
    from fcfv_nme23_b200 import mesh as M
    n = 13
    h = api.Handle(0)
    h.generate_structured(n, setting=setting, tau_r=10.0 if setting == "solkz" else 2.0, neumann_south=(setting == "inclusion"))
    dm = h.get_mesh()
    host = cases.load_mesh(f"S{n}")
    assert h.nel == host.nel and h.nf == host.nf
    assert np.array_equal(dm["e2f"], host.e2f)
    for k in ("Ω", "n_x", "n_y", "Γ"):
        assert np.array_equal(dm[k], getattr(host, k)), k
    bc = host.bc.copy()
    if setting == "inclusion":
        bc[(bc == 1) & (host.yf == 0.0)] = 2
        assert (dm["bc"] == -1).sum() > 0
    assert np.array_equal(dm["bc"] != 0, bc != 0) or setting == "inclusion"
    assert np.array_equal(dm["bc"][bc > 0], bc[bc > 0])
    h.close()


@pytest.mark.parametrize("setting,variant,form", [("solkz", 1, "SymmetricGradient"), ("inclusion", 0, "Gradient")])
def test_device_pipeline_vs_oracle(setting, variant, form, api, oracle):
    """Device-resident pipeline (what bench.py times) on a generated mesh: pull the generated inputs
    back, run the oracle on them, compare everything."""
    from fcfv_nme23_b200 import lib as L
    from fcfv_nme23_b200.mesh import FCFV_Mesh
    n = 24
    h = api.Handle(0)
    tau_r = 10.0 if setting == "solkz" else 2.0
    h.generate_structured(n, setting=setting, tau_r=tau_r, neumann_south=(setting == "inclusion"))
    h.run_local(form)
    h.run_assemble(variant, form, want_muu=True)
    ss, norms = h.run_residuals(form, "FACE")
    dm, dd = h.get_mesh(), h.get_data()
    m = FCFV_Mesh(nel=h.nel, nf=h.nf, nf_el=3)
    for k, v in dm.items():
        setattr(m, k, v)
    m.τ = None
    neu = [dd["sxx"], dd["syy"], dd["sxy"], dd["syx"]]
    a, b, Z = oracle.ComputeFCFV(m, dd["sex"], dd["sey"], dd["VxDir"], dd["VyDir"], *neu, tau_r, form)
    ga, gb, gZ, gtau = h.get_local()
    assert np.array_equal(ga, a) and np.array_equal(gb, b) and np.array_equal(gZ, Z) and np.array_equal(gtau, m.τ)
    ofn = oracle.ElementAssemblyLoop if variant == 0 else oracle.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = ofn(m, a, b, Z, dd["VxDir"], dd["VyDir"], *neu, form)
    nel, nf = m.nel, m.nf
    assert_csr_equal(h.export_csr(L.KUU), csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu")
    assert_csr_equal(h.export_csr(L.KUP), csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup")
    assert_csr_equal(h.export_csr(L.MUU), csc_triple_to_csr(Muu, (2 * nf, 2 * nf)), "Muu")
    gfu, gfp = h.export_rhs()
    assert np.array_equal(gfu, fu.ravel()) and np.array_equal(gfp, fp)
    on = oracle.ComputeResidualsFCFV_Stokes_o1(dd["Vxh"], dd["Vyh"], dd["Pe"], m, a, b, Z, dd["sex"], dd["sey"], dd["VxDir"],
                                               dd["VyDir"], *neu, form, tau_mode="FACE")
    assert rel_inf(norms, on) <= TOL
    h.close()


def _strip_face_global(n, row0, row1):
    """local -> global (0-based) face ids of a generated strip (fcfv_generate.cuh layout)."""
    rlo = row0 - 1 if row0 > 0 else 0
    rhi = row1 + 1 if row1 < n else n
    base = lambda qx, qy: 6 * (qy * n + qx) + (qx if qy == 0 else n) + qy + (1 if qx > 0 else 0)
    ids = []
    if rlo > 0:
        ids += [base(qx, rlo - 1) + (1 if rlo - 1 == 0 else 0) + 4 for qx in range(n)]
    ids += list(range(base(0, rlo), base(0, rhi) if rhi < n else 6 * n * n + 2 * n))
    return np.array(ids, dtype=np.int64), rlo, rhi


@pytest.mark.parametrize("setting,variant,form", [("solkz", 1, "SymmetricGradient"), ("inclusion", 1, "Gradient")])
def test_strips_reproduce_single_gpu_rows(setting, variant, form, api):
    """Element-partitioned strips with ghost rows: every owned row is bit-identical to the row the
    unpartitioned run produces, and the partial sums of squares add up to the global ones."""
    from fcfv_nme23_b200 import lib as L
    n, parts = 12, 3
    tau_r = 10.0 if setting == "solkz" else 2.0
    full = api.Handle(0)
    full.generate_structured(n, setting=setting, tau_r=tau_r)
    full.run_local(form); full.run_assemble(variant, form); ss_full, _ = full.run_residuals(form, "FACE")
    NF = full.nf
    rp, ci, v = full.export_csr(L.KUU)
    rpp, cip, vp = full.export_csr(L.KUP)
    fu_full, fp_full = full.export_rhs()
    ss_sum = np.zeros(6)
    rows_seen = np.zeros(2 * NF, dtype=int)
    for r in range(parts):
        row0, row1 = r * n // parts, (r + 1) * n // parts
        h = api.Handle(0)
        h.generate_structured(n, row0, row1, setting=setting, tau_r=tau_r)
        h.run_local(form); h.run_assemble(variant, form); ss, _ = h.run_residuals(form, "FACE")
        ss_sum += ss
        fg, rlo, rhi = _strip_face_global(n, row0, row1)
        assert fg.size == h.nf
        lrp, lci, lv = h.export_csr(L.KUU)
        lrpp, lcip, lvp = h.export_csr(L.KUP)
        lfu, lfp = h.export_rhs()
        nf = h.nf
        ebase = 4 * n * rlo
        for blockoff_l, blockoff_g in ((0, 0), (nf, NF)):
            for lf in range(nf):
                lr, gr = blockoff_l + lf, blockoff_g + fg[lf]
                cnt = lrp[lr + 1] - lrp[lr]
                if cnt == 0 and lfu[lr] == 0.0:
                    continue   # not owned here (owned rows always hold their diagonal)
                rows_seen[gr] += 1
                cols_l = lci[lrp[lr]:lrp[lr + 1]].astype(np.int64)
                cols_g = np.where(cols_l >= nf, fg[cols_l - nf] + NF, fg[np.minimum(cols_l, nf - 1)])
                assert np.array_equal(cols_g, ci[rp[gr]:rp[gr + 1]]), (r, lf)
                assert np.array_equal(lv[lrp[lr]:lrp[lr + 1]], v[rp[gr]:rp[gr + 1]])
                assert np.array_equal(lcip[lrpp[lr]:lrpp[lr + 1]] + ebase, cip[rpp[gr]:rpp[gr + 1]])
                assert np.array_equal(lvp[lrpp[lr]:lrpp[lr + 1]], vp[rpp[gr]:rpp[gr + 1]])
                assert lfu[lr] == fu_full[gr]
        own = slice(4 * n * (row0 - rlo), 4 * n * (row1 - rlo))
        assert np.array_equal(lfp[own], fp_full[4 * n * row0:4 * n * row1])
        h.close()
    assert np.all(rows_seen == 1), "every global row must be produced by exactly one strip"
    assert rel_inf(ss_sum, ss_full) <= TOL
    full.close()


def test_large_mesh_properties(api):
    """1M-element mesh (BASELINE config 3 size): size-independent properties on the device result."""
    from fcfv_nme23_b200 import lib as L
    n = 500
    h = api.Handle(0)
    h.generate_structured(n, setting="solkz", tau_r=10.0)
    assert h.nel == 4 * n * n and h.nf == 6 * n * n + 2 * n
    h.run_local("SymmetricGradient")
    h.run_assemble(L.LOOP_NEW, "SymmetricGradient")
    nnz_uu, nnz_up, _ = h.get_nnz()
    rp, ci, v = h.export_csr(L.KUU)
    assert rp[0] == 0 and rp[-1] == nnz_uu and np.all(np.diff(rp) >= 1) and np.diff(rp).max() <= 10
    # columns strictly ascending inside every row
    inner = np.ones(nnz_uu, dtype=bool); inner[rp[1:-1]] = False
    assert np.all(np.diff(ci.astype(np.int64))[inner[1:]] > 0)
    assert np.all(v != 0.0) and np.all(np.isfinite(v))
    # 24.78..25 nnz/element on S(n) for :SymmetricGradient, Kup ~ 5/element
    assert 24.9 < nnz_uu / h.nel < 25.0 and 4.99 < nnz_up / h.nel < 5.0
    # the symmetric formulation assembles a symmetric Kuu (pattern and values to round-off)
    import scipy.sparse as sp
    K = sp.csr_matrix((v, ci, rp), shape=(2 * h.nf, 2 * h.nf))
    D = (K - K.T).tocoo()
    assert np.abs(D.data).max() <= 1e-12 * np.abs(v).max()
    # identity 1: matrix-free residual == SpMV residual on non-Dirichlet rows, for arbitrary (u,p)
    dd, dm = h.get_data(), h.get_mesh()
    u = np.concatenate([dd["Vxh"], dd["Vyh"]])
    ru = np.empty(2 * h.nf); rp_ = np.empty(h.nel); nrm = np.zeros(2)
    import ctypes as C
    h.check(h.lib.fcfv_ph_residual(h._h, u.ctypes.data_as(C.c_void_p), dd["Pe"].ctypes.data_as(C.c_void_p),
                                   ru.ctypes.data_as(C.c_void_p), rp_.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p)))
    ss, norms = h.run_residuals("SymmetricGradient", "FACE")
    nd = dm["bc"] != 1
    fx2 = (ru[:h.nf][nd] ** 2).sum(); fy2 = (ru[h.nf:][nd] ** 2).sum()
    assert abs(fx2 - ss[3]) <= 1e-9 * ss[3] and abs(fy2 - ss[4]) <= 1e-9 * ss[4]
    assert abs((rp_ ** 2).sum() - ss[5]) <= 1e-9 * ss[5]
    # run-to-run determinism: a second assembly gives identical bytes
    h.run_assemble(L.LOOP_NEW, "SymmetricGradient")
    rp2, ci2, v2 = h.export_csr(L.KUU)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(v, v2)
    h.close()


@pytest.mark.parametrize("mesh_name,setting,variant,form,nparts", [
    ("LowRes", "inclusion", 0, "Gradient", 3),
    ("H1", "solkz", 1, "SymmetricGradient", 2),
    ("random:5", "random", 1, "SymmetricGradient", 3),      # tests/test_random_meshes.py: no numbering coherence at all
    ("random:2", "random", 0, "Gradient", 4),
])
def test_partitioned_mesh_on_gpu(mesh_name, setting, variant, form, nparts, api, oracle):
    """General (unstructured) element partition through the C ABI, ranks run one after the other on
    one GPU: union of owned rows == oracle rows of the unpartitioned mesh, bit for bit; partial
    sums of squares add up to the global norms (both tau modes, REFERENCE via tau_ref)."""
    from fcfv_nme23_b200 import lib as L
    from fcfv_nme23_b200.partition import partition_mesh
    case = cases.Case(mesh_name, setting, variant, form, tau_r=2.0 if setting == "inclusion" else 10.0)
    if setting == "random":
        from test_random_meshes import random_case
        gm, d = random_case(int(mesh_name.split(":")[1]))
    else:
        gm, d = cases.build(case)
    a, b, Z = oracle.ComputeFCFV(gm, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], case.tau_r, form)
    ofn = oracle.ElementAssemblyLoop if variant == 0 else oracle.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = ofn(gm, a, b, Z, d["VxDir"], d["VyDir"], *d["neu"], form)
    NF, NEL = gm.nf, gm.nel
    grp, gci, gv = csc_triple_to_csr(Kuu, (2 * NF, 2 * NF))
    grpp, gcip, gvp = csc_triple_to_csr(Kup, (2 * NF, NEL))
    norms = {tm: oracle.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], gm, a, b, Z, d["sex"], d["sey"], d["VxDir"],
                                                       d["VyDir"], *d["neu"], form, tau_mode=tm) for tm in ("REFERENCE", "FACE")}
    ss = {"REFERENCE": np.zeros(6), "FACE": np.zeros(6)}
    rows_seen = np.zeros(2 * NF, dtype=int)
    for r in range(nparts):
        P = partition_mesh(gm, nparts, r)      # gm.τ is the global τ after ComputeFCFV -> tau_ref
        m = P.mesh
        m.τ = None
        sf, se = P.scatter_face, P.scatter_elem
        neu = [sf(x) for x in d["neu"]]
        la, lb, lZ = api.ComputeFCFV(m, se(d["sex"]), se(d["sey"]), sf(d["VxDir"]), sf(d["VyDir"]), *neu, case.tau_r, form)
        assert np.array_equal(la[P.elem_owned], a[P.elem_global[P.elem_owned]])
        gfn = api.ElementAssemblyLoop if variant == 0 else api.ElementAssemblyLoopNEW
        h, nnz, sec = gfn(m, la, lb, lZ, sf(d["VxDir"]), sf(d["VyDir"]), *neu, form, export=False)
        lrp, lci, lv = h.export_csr(L.KUU)
        lrpp, lcip, lvp = h.export_csr(L.KUP)
        lfu, lfp = h.export_rhs()
        for lr, gr in zip(P.local_rows(), P.owned_rows()):
            rows_seen[gr] += 1
            assert np.array_equal(P.global_columns(lci[lrp[lr]:lrp[lr + 1]]), gci[grp[gr]:grp[gr + 1]])
            assert np.array_equal(lv[lrp[lr]:lrp[lr + 1]], gv[grp[gr]:grp[gr + 1]])
            assert np.array_equal(P.global_columns(lcip[lrpp[lr]:lrpp[lr + 1]], "up"), gcip[grpp[gr]:grpp[gr + 1]])
            assert np.array_equal(lvp[lrpp[lr]:lrpp[lr + 1]], gvp[grpp[gr]:grpp[gr + 1]])
            assert lfu[lr] == fu[gr, 0]
        # rows of faces owned elsewhere are empty
        notmine = np.setdiff1d(np.arange(2 * m.nf), P.local_rows())
        assert np.all(np.diff(lrp)[notmine] == 0)
        assert np.array_equal(lfp[P.elem_owned], fp[P.elem_global[P.elem_owned]])
        h.set_solution(sf(d["Vxh"]), sf(d["Vyh"]), se(d["Pe"]))
        for tm in ("REFERENCE", "FACE"):
            s, _ = h.run_residuals(form, tm)
            ss[tm] += s
        h.close()
    assert np.all(rows_seen[np.diff(grp) > 0] == 1)
    lens = np.array([4 * NEL, 2 * NEL, NEL, NF, NF, NEL], dtype=float)
    for tm in ("REFERENCE", "FACE"):
        got = np.sqrt(ss[tm]) / np.sqrt(lens)
        big = norms[tm] > 1e-9
        assert np.all(np.abs(got[big] - norms[tm][big]) <= TOL * norms[tm][big]), (tm, got, norms[tm])
        assert np.all(got[~big] < 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["Gradient", "SymmetricGradient"])
def test_anisotropic_delta_gpu(form, api, oracle):
    """delta != 1: ElementAssemblyLoopNEW through the ANISO kernels (full 2x2 D per element), residuals, element
    values and the Schur complement against the oracle (sparsity exact, values <= 1e-12: LAPACK inv in the reference)."""
    from fcfv_nme23_b200 import lib as L
    case = cases.Case("H1", "solkz", 1, form, tau_r=10.0)
    m, d = cases.build(case)
    g, _ = cases.build(case)
    for x in (m, g):
        x.δ = np.where(x.xc > 0.5, 1.0 + 0.75 * np.abs(np.sin(7.0 * x.yc)) + 0.25, 1.0)
    nel, nf = m.nel, m.nf
    neu = d["neu"]
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, form)
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, form)
    assert np.array_equal(ga, a) and np.array_equal(gb, b) and np.array_equal(gZ, Z)
    Kuu, Muu, Kup, fu, fp, _, _ = oracle.ElementAssemblyLoopNEW(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, form)
    gKuu, gMuu, gKup, gfu, gfp, _ = api.ElementAssemblyLoopNEW(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, form)
    h = api.handle_for(g)
    assert_csr_equal(h.export_csr(L.KUU), csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu", bitwise=False, rtol=TOL)
    assert_csr_equal(h.export_csr(L.KUP), csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup")
    assert_csr_equal(h.export_csr(L.MUU), csc_triple_to_csr(Muu, (2 * nf, 2 * nf)), "Muu", bitwise=False, rtol=TOL)
    assert rel_inf(gfu, fu) <= TOL and np.array_equal(gfp, fp)
    ref_sc = oracle.schur_complement(m, Kuu, Kup, 1.0e4)
    api.SchurComplement(g, 1.0e4)
    assert_csr_equal(h.export_csr(L.KUUSC), csc_triple_to_csr(ref_sc, (2 * nf, 2 * nf)), "Kuusc", bitwise=False, rtol=TOL)
    norms = oracle.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"],
                                                  d["VyDir"], *neu, form, tau_mode="FACE")
    gn = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], g, ga, gb, gZ, d["sex"], d["sey"], d["VxDir"],
                                            d["VyDir"], *neu, form, tau_mode="FACE", verbose=False)
    big = norms > 1e-9
    assert np.all(np.abs(gn[big] - norms[big]) <= TOL * norms[big])
    ev = oracle.ComputeElementValues(m, d["Vxh"], d["Vyh"], d["Pe"], a, b, Z, d["VxDir"], d["VyDir"], form)
    gev = api.ComputeElementValues(g, d["Vxh"], d["Vyh"], d["Pe"], ga, gb, gZ, d["VxDir"], d["VyDir"], form)
    for k in range(5):
        assert np.array_equal(gev[k], ev[k])
    h.close()


@pytest.mark.gpu
def test_int64_row_pointers(api, oracle, monkeypatch):
    """Meshes whose CSR capacity exceeds 2^31 entries switch to 64-bit row pointers (fcfv_index_width == 8).  The
    code path is forced on a small mesh and must give the same CSR / CSC as the oracle, Schur complement included."""
    from fcfv_nme23_b200 import lib as L
    monkeypatch.setenv("FCFV_FORCE_RP64", "1")
    case = cases.Case("H1", "solkz", 1, "SymmetricGradient", tau_r=10.0)
    m, d = cases.build(case)
    g, _ = cases.build(case)
    nel, nf = m.nel, m.nf
    neu = d["neu"]
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    Kuu, Muu, Kup, fu, fp, _, _ = oracle.ElementAssemblyLoopNEW(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form)
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    api.ElementAssemblyLoopNEW(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, case.form)
    h = api.handle_for(g)
    assert h.lib.fcfv_index_width(h._h) == 8
    rp, ci, v = h.export_csr(L.KUU)
    assert rp.dtype == np.int64
    assert_csr_equal((rp, ci, v), csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu")
    assert_csr_equal(h.export_csr(L.KUP), csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup")
    assert_csr_equal(h.export_csr(L.MUU), csc_triple_to_csr(Muu, (2 * nf, 2 * nf)), "Muu")
    ref_sc = oracle.schur_complement(m, Kuu, Kup, 1.0e4)
    api.SchurComplement(g, 1.0e4)
    cp, rv, vv = h.export_csc(L.KUUSC)
    assert np.array_equal(cp, ref_sc[0]) and np.array_equal(rv, ref_sc[1]) and np.array_equal(vv, ref_sc[2])
    u = np.concatenate([d["Vxh"], d["Vyh"]])
    ru, rp_, nrm = oracle.ph_residual(m, Kuu, Kup, fu, fp, u, d["Pe"])
    gru, grp, gnrm = api.ph_residual(g, u, d["Pe"])
    assert rel_inf(gru, ru) <= TOL and rel_inf(grp, rp_) <= TOL
    h.close()


@pytest.mark.gpu
def test_config4_50M_properties(api):
    """BASELINE config 4 size (S(3536), 50 013 184 elements) on one GPU: size-independent properties of the device
    result -- closed-form sizes, CSR row pointers, SURVEY §4 identity 1 through norms, run-to-run determinism."""
    import ctypes as C
    from fcfv_nme23_b200 import lib as L
    n = 3536
    h = api.Handle(0)
    h.generate_structured(n, setting="solkz", tau_r=10.0)
    assert h.nel == 4 * n * n == 50013184 and h.nf == 6 * n * n + 2 * n
    form = "SymmetricGradient"
    h.run_local(form); h.run_assemble(L.LOOP_NEW, form)
    ss, norms = h.run_residuals(form, "FACE")
    nnz_uu, nnz_up, _ = h.get_nnz()
    assert 24.99 < nnz_uu / h.nel < 25.0 and 4.999 < nnz_up / h.nel < 5.0
    w = h.lib.fcfv_index_width(h._h)
    rp = np.empty(2 * h.nf + 1, dtype=np.int64 if w == 8 else np.int32)
    h.check(h.lib.fcfv_export(h._h, L.KUU, L.CSR0, C.c_void_p(rp.ctypes.data), None, None))
    d = np.diff(rp.astype(np.int64))
    assert rp[0] == 0 and int(rp[-1]) == nnz_uu and d.min() >= 1 and d.max() <= 10
    # Dirichlet rows hold exactly the unit diagonal: 8n boundary faces x 2 components
    assert int((d == 1).sum()) == 2 * 4 * n
    del rp, d
    # identity 1: ||fu - Kuu u - Kup p||^2 = F_nodes sums + Dirichlet rows (VDir - u)^2 ;  ||fp - Kpu u||^2 = F_glob2 sum
    dd, dm = h.get_data(), h.get_mesh()
    u = np.concatenate([dd["Vxh"], dd["Vyh"]])
    nrm = np.zeros(2)
    h.check(h.lib.fcfv_ph_residual(h._h, C.c_void_p(u.ctypes.data), C.c_void_p(dd["Pe"].ctypes.data), None, None,
                                   C.c_void_p(nrm.ctypes.data)))
    bd = dm["bc"] == 1
    dir2 = ((dd["VxDir"][bd] - dd["Vxh"][bd]) ** 2).sum() + ((dd["VyDir"][bd] - dd["Vyh"][bd]) ** 2).sum()
    assert abs(nrm[0] ** 2 - (ss[3] + ss[4] + dir2)) <= 1e-9 * nrm[0] ** 2
    assert abs(nrm[1] ** 2 - ss[5]) <= 1e-9 * ss[5]
    del dd, dm, u
    # determinism
    h.run_assemble(L.LOOP_NEW, form)
    ss2, _ = h.run_residuals(form, "FACE")
    assert h.get_nnz()[:2] == (nnz_uu, nnz_up) and np.array_equal(ss, ss2)
    h.close()


@pytest.mark.gpu
def test_graph_pipeline_matches_direct_calls(api):
    """fcfv_run_pipeline (one replayed CUDA graph per pass) against the three direct calls: identical CSR bytes, nnz and
    residual sums; replays see updated fields (same device buffers), a new mesh size rebuilds the graph."""
    from fcfv_nme23_b200 import lib as L
    form, variant = "SymmetricGradient", 1
    for n in (9, 14):                                   # second size: buffers grow / mesh changes -> new graph
        h = api.Handle(0)
        h.generate_structured(n, setting="solkz", tau_r=10.0)
        h.run_local(form); h.run_assemble(variant, form)
        ss_ref, _ = h.run_residuals(form, "FACE")
        ref = h.export_csr(L.KUU); refp = h.export_csr(L.KUP); rhs = h.export_rhs(); nnz = h.get_nnz()
        for rep in range(3):                            # capture, then replays
            h.run_pipeline(variant, form, "FACE")
            h.sync()
            assert h.get_nnz() == nnz
            for a, b in zip(h.export_csr(L.KUU) + h.export_csr(L.KUP) + h.export_rhs(), ref + refp + rhs):
                assert np.array_equal(a, b)
            ss, _ = h.run_residuals(form, "FACE")
            assert np.array_equal(ss, ss_ref)
        # change a field in place: the replay must see it
        m = h.get_mesh()
        h.update_fields(ke=m["ke"] * 2.0)
        h.run_pipeline(variant, form, "FACE"); h.sync()
        got = h.export_csr(L.KUU)
        h.run_local(form); h.run_assemble(variant, form)
        exp = h.export_csr(L.KUU)
        assert all(np.array_equal(a, b) for a, b in zip(got, exp)) and not np.array_equal(got[2], ref[2])
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_module_constant_A_not_one_gpu(variant, api, oracle):
    """A = 0.9: the kernels' general path for the symmetric formulation (A == 1.0 runs a specialised instantiation)."""
    from fcfv_nme23_b200 import lib as L
    case = cases.Case("H1", "solkz", variant, "SymmetricGradient", tau_r=10.0)
    m, d = cases.build(case)
    g, _ = cases.build(case)
    nel, nf = m.nel, m.nf
    neu = d["neu"]
    A = 0.9
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form, A=A)
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form, A=A)
    assert np.array_equal(ga, a) and np.array_equal(gb, b) and np.array_equal(gZ, Z)
    ofn = oracle.ElementAssemblyLoop if variant == 0 else oracle.ElementAssemblyLoopNEW
    gfn = api.ElementAssemblyLoop if variant == 0 else api.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = ofn(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form, A=A)
    gfn(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, case.form, A=A)
    h = api.handle_for(g)
    assert_csr_equal(h.export_csr(L.KUU), csc_triple_to_csr(Kuu, (2 * nf, 2 * nf)), "Kuu")
    assert_csr_equal(h.export_csr(L.KUP), csc_triple_to_csr(Kup, (2 * nf, nel)), "Kup")
    assert_csr_equal(h.export_csr(L.MUU), csc_triple_to_csr(Muu, (2 * nf, 2 * nf)), "Muu")
    gfu, gfp = h.export_rhs()
    assert np.array_equal(gfu, fu.ravel()) and np.array_equal(gfp, fp)
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["S17", "LowRes", "H2"])
def test_first_seen_faces_device_matches_host(name, api):
    """fcfv_first_seen_faces (device) == mesh.first_seen_face_permutation (host) on shuffled meshes, and a mesh
    renumbered with it assembles the same matrix up to that permutation of rows and columns."""
    from fcfv_nme23_b200 import mesh as M
    m = cases.load_mesh(name)
    rng = np.random.default_rng(5)
    sh = rng.permutation(m.nf) + 1
    e2f_sh = np.asfortranarray(sh[np.asarray(m.e2f) - 1])
    perm, e2f_new = api.first_seen_faces(e2f_sh, m.nf)
    assert np.array_equal(perm, M.first_seen_face_permutation(e2f_sh, m.nf))
    assert np.array_equal(e2f_new, perm[e2f_sh - 1])
    # with unreferenced faces
    e2f = np.asfortranarray(np.array([[1, 4, 6], [4, 2, 7]], dtype=np.int64))
    p, _ = api.first_seen_faces(e2f, 8)
    assert p.tolist() == [1, 4, 6, 2, 7, 3, 5, 8]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["H1", "LowRes"])
def test_numbering_preparation_on_device(name, api):
    """fcfv_hilbert_elements + fcfv_first_seen_faces on a mesh whose elements and faces come in random order: the
    device results equal the host statements, and the prepared mesh assembles the same system bit for bit up to the
    permutation of rows and columns (uniform τe, so that the highest-element τ rule does not see the numbering)."""
    from fcfv_nme23_b200 import mesh as M
    form = "SymmetricGradient"
    case = cases.Case(name, "inclusion", 1, form, tau_r=2.0)
    m0, d = cases.build(case)
    nel, nf = m0.nel, m0.nf
    neu = d["neu"]
    a0, b0, Z0 = api.ComputeFCFV(m0, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, form)
    Kuu0, _, Kup0, fu0, fp0, _ = api.ElementAssemblyLoopNEW(m0, a0, b0, Z0, d["VxDir"], d["VyDir"], *neu, form)
    api.handle_for(m0).close()

    rng = np.random.default_rng(21)
    ms, _ = cases.build(case)
    old_e = rng.permutation(nel) + 1                       # element k of ms = element old_e[k] of m0
    pf = rng.permutation(nf) + 1                           # face f of m0 -> face pf[f] of ms
    M.renumber_elements(ms, old_e)
    M.renumber_faces(ms, pf)
    # preparation, device against host
    order = api.hilbert_elements(ms.xc, ms.yc)
    assert np.array_equal(order, M.hilbert_element_order(ms.xc, ms.yc))
    if name == "H1":                                     # degenerate boxes, ties, the empty mesh
        for x, y in ((np.zeros(5), np.arange(5.0)), (np.zeros(3), np.zeros(3)), (np.array([0.3, 0.3, 0.1, 0.3]), np.array([0.2, 0.2, 0.9, 0.2]))):
            assert np.array_equal(api.hilbert_elements(x, y), M.hilbert_element_order(x, y))
        assert api.hilbert_elements(np.zeros(0), np.zeros(0)).size == 0
    M.renumber_elements(ms, order)
    old_e = old_e[order - 1]
    perm, e2f_new = api.first_seen_faces(np.asfortranarray(ms.e2f), nf)
    assert np.array_equal(perm, M.first_seen_face_permutation(ms.e2f, nf))
    M.renumber_faces(ms, perm)
    assert np.array_equal(np.asarray(ms.e2f), e2f_new)
    pf = perm[pf - 1]
    jump = np.hypot(np.diff(ms.xc), np.diff(ms.yc))
    assert np.median(jump) < 4.0 * np.sqrt(np.median(ms.Ω))
    # the locality diagnostic tells the two numberings apart
    hs = api.Handle(0); hs.set_mesh(ms); good = hs.numbering_locality(); hs.close()
    mb, _ = cases.build(case)
    M.renumber_elements(mb, rng.permutation(nel) + 1); M.renumber_faces(mb, rng.permutation(nf) + 1)
    hb = api.Handle(0); hb.set_mesh(mb); bad = hb.numbering_locality(); hb.close()
    assert good[0] < 0.08 and good[1] < 0.08 and bad[0] > 0.25 and bad[1] > 0.25, (good, bad)

    def faces(x):
        out = np.empty_like(x); out[pf - 1] = x; return out
    sex, sey = d["sex"][old_e - 1], d["sey"][old_e - 1]
    fd = [faces(x) for x in (d["VxDir"], d["VyDir"], *neu)]
    a, b, Z = api.ComputeFCFV(ms, sex, sey, *fd, case.tau_r, form)
    assert np.array_equal(a, a0[old_e - 1]) and np.array_equal(b, b0[old_e - 1]) and np.array_equal(Z, Z0[old_e - 1])
    Kuu, _, Kup, fu, fp, _ = api.ElementAssemblyLoopNEW(ms, a, b, Z, *fd, form)
    api.handle_for(ms).close()
    dof = np.concatenate([pf - 1, nf + pf - 1])            # dof of m0 -> dof of ms
    back_uu = Kuu.tocsr()[dof][:, dof].tocsc()
    back_up = Kup.tocsr()[dof][:, np.argsort(old_e)].tocsc()
    for got, ref, tag in ((back_uu, Kuu0, "Kuu"), (back_up, Kup0, "Kup")):
        got.sort_indices(); ref = ref.copy(); ref.sort_indices()
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices), tag
        assert np.array_equal(got.data, ref.data), tag
    assert np.array_equal(np.asarray(fu).ravel()[dof], np.asarray(fu0).ravel())
    assert np.array_equal(np.asarray(fp).ravel(), np.asarray(fp0).ravel()[old_e - 1])


@pytest.mark.gpu
def test_host_register_roundtrip(api):
    """fcfv_host_register / fcfv_host_unregister: page-locking the caller's arrays changes nothing but the copy speed;
    registering twice or unregistering an unknown range is reported, not fatal."""
    case = cases.Case("H1", "solkz", 1, "SymmetricGradient", tau_r=10.0)
    m, d = cases.build(case)
    neu = d["neu"]
    a0, b0, Z0 = api.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    h = api.handle_for(m)
    arrs = [np.ascontiguousarray(d[k], dtype=np.float64) for k in ("sex", "sey", "VxDir", "VyDir")]
    h.host_register(*arrs)
    a1, b1, Z1 = api.ComputeFCFV(m, *arrs, *neu, case.tau_r, case.form)
    assert np.array_equal(a0, a1) and np.array_equal(b0, b1) and np.array_equal(Z0, Z1)
    with pytest.raises(RuntimeError):
        h.host_register(arrs[0])                   # already registered
    h.host_unregister(*arrs)
    with pytest.raises(RuntimeError):
        h.host_unregister(arrs[0])                 # not registered any more
    a2, _, _ = api.ComputeFCFV(m, *arrs, *neu, case.tau_r, case.form)   # the handle is still usable
    assert np.array_equal(a0, a2)
    h.close()


@pytest.mark.parametrize("form,new_loop", [("Gradient", False), ("SymmetricGradient", False), ("SymmetricGradient", True)])
def test_example_driver_solves_the_inclusion_problem(form, new_loop):
    """examples/viscous_inclusion.py: the whole driver flow on the GPU path (assembly, Schur complement, PH lines,
    residuals, element values) around a host LU.  The solve converges and every residual of the reference's six
    lines is at the Powell-Hestenes tolerance; the pure-shear symmetry of the problem shows in the fields."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("viscous_inclusion", os.path.join(os.path.dirname(__file__), "..", "examples",
                                                                                   "viscous_inclusion.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    r = ex.solve(16, Formulation=form, new_loop=new_loop, verbose=False)
    assert r["iterations"] <= 8
    assert r["norms"][0] < 1e-12 and r["norms"][1] < 1e-11 and np.all(r["norms"][2:] < 1e-7), r["norms"]
    m = r["mesh"]
    # pure shear around the centre: Vx is odd in (x - 1/2), pressure integrates to ~0 against the symmetry
    left, right = m.xc < 0.5, m.xc > 0.5
    assert abs(r["Vxe"][left].mean() + r["Vxe"][right].mean()) < 1e-8
    assert np.isfinite(r["Txxe"]).all() and np.abs(r["Txxe"]).max() > 0.1


def test_staged_pageable_copies_give_identical_results(monkeypatch, api, oracle):
    """Pageable host arrays above FCFV_STAGED_MIN_BYTES travel through the multi-threaded staging path (worker
    threads, page-locked chunks): with the threshold forced down to 1 KiB every array of a small mesh takes it --
    odd sizes, slices shorter than a chunk, empty slices -- and every result equals the direct path's, bit for bit.
    A second mesh is large enough for chunked slices (several chunks per worker)."""
    from fcfv_nme23_b200 import lib as L
    case = cases.Case("LowRes", "inclusion", 0, "Gradient")
    ref = {}
    monkeypatch.setenv("FCFV_SPARSE_FACE_DATA", "0")      # whole per-face arrays: they are part of what is staged here
    for tag, env in (("direct", "0"), ("staged", "1024")):
        monkeypatch.setenv("FCFV_STAGED_MIN_BYTES", env)
        m, d = cases.build(case)
        neu = d["neu"]
        a, b, Z = api.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, 2.0, case.form)
        K, Mu, P, fu, fp, _ = api.ElementAssemblyLoop(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form)
        n = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"], d["VyDir"],
                                               *neu, case.form, verbose=False)
        h = api.handle_for(m)
        assert (h.staged_copies > 20) == (tag == "staged"), (tag, h.staged_copies)
        ref[tag] = (a, b, Z, m.τ.copy(), K.indptr, K.indices, K.data, P.indptr, P.indices, P.data, Mu.data, np.asarray(fu), fp, n)
        h.close()
    for x, y in zip(ref["direct"], ref["staged"]):
        assert np.array_equal(x, y)
    # several chunks per worker: S(520), 1.08 M elements, arrays of 8.7 - 26 MB against the default 8 MiB threshold
    monkeypatch.delenv("FCFV_STAGED_MIN_BYTES")
    h = api.Handle(0)
    h.generate_structured(520, setting="solkz", tau_r=10.0)
    M0, D0 = h.get_mesh(), h.get_data()               # D2H of large arrays into fresh pageable memory: staged
    assert h.staged_copies > 0
    h.run_local("SymmetricGradient")
    al0 = h.get_local()[0]
    h.close()
    from fcfv_nme23_b200.mesh import FCFV_Mesh
    m = FCFV_Mesh(nel=M0["Ω"].size, nf=M0["bc"].size, nf_el=3)
    for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe"):
        setattr(m, k, M0[k])
    z = np.zeros(m.nf)
    a, b, Z = api.ComputeFCFV(m, D0["sex"], D0["sey"], D0["VxDir"], D0["VyDir"], z, z, z, z, 10.0, "SymmetricGradient")
    h2 = api.handle_for(m)
    assert h2.staged_copies > 10 and np.array_equal(a, al0)
    h2.close()


@pytest.mark.parametrize("solver", ["CoupledBackslash", "PowellHestenesLU"])
def test_stokes_solvers_mirror(solver, api, oracle):
    """solvers.StokesSolvers (mirror of StokesSolvers!, SolversFCFV_Stokes.jl:6-56) around the GPU lines: both solver
    families give the same velocity field on LowRes and the matrix-free residuals of the solution vanish."""
    from fcfv_nme23_b200 import solvers
    case = cases.Case("LowRes", "inclusion", 0, "Gradient")
    m, d = cases.build(case)
    z, ze = np.zeros(m.nf), np.zeros(m.nel)
    a, b, Z = api.ComputeFCFV(m, ze, ze, d["VxDir"], d["VyDir"], z, z, z, z, 2.0, case.form)
    Kuu, Muu, Kup, fu, fp, _ = api.ElementAssemblyLoop(m, a, b, Z, d["VxDir"], d["VyDir"], z, z, z, z, case.form)
    Vxh, Vyh, Pe = np.zeros(m.nf), np.zeros(m.nf), np.zeros(m.nel)
    its = solvers.StokesSolvers(Vxh, Vyh, Pe, m, Kuu, Kup, fu, fp, Muu, solver, penalty=1e5, tol=1e-8, verbose=False)
    assert its <= 6
    n = api.ComputeResidualsFCFV_Stokes_o1(Vxh, Vyh, Pe, m, a, b, Z, ze, ze, d["VxDir"], d["VyDir"], z, z, z, z, case.form,
                                           tau_mode="FACE", verbose=False)
    assert np.all(n < 1e-7), n
    dir_faces = m.bc == 1
    assert np.allclose(Vxh[dir_faces], d["VxDir"][dir_faces], atol=1e-9)
    if solver == "CoupledBackslash":
        assert abs(Pe.mean()) < 1e-10
    api.handle_for(m).close()
    with pytest.raises(ValueError):
        solvers.StokesSolvers(Vxh, Vyh, Pe, m, Kuu, Kup, fu, fp, Muu, "Jacobi")


@pytest.mark.parametrize("eta", [(1.0, 1e-2), (1.0, 10.0)], ids=["weak-inclusion", "strong-inclusion"])
def test_discretisation_error_converges_to_the_analytic_solution(eta):
    """Known answer from outside the code base: the analytic solution of a circular inclusion under pure shear (Schmid &
    Podladchikov 2003) imposed on the box boundary.  The GPU-assembled :SymmetricGradient system (host LU + PH lines)
    approaches it at the scheme's first order in L1, velocity and pressure, on S(16), S(32), S(64) (the mesh is not
    fitted to the circle); ElementAssemblyLoopNEW gives the same errors as ElementAssemblyLoop."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("viscous_inclusion", os.path.join(os.path.dirname(__file__), "..", "examples",
                                                                                   "viscous_inclusion.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    errs = [ex.solve(n, Formulation="SymmetricGradient", verbose=False, exact_bc=True, η=eta, tol=1e-10)["errors"] for n in (16, 32, 64)]
    for k in ("Vx", "Vy", "P"):
        assert errs[0][k] / errs[1][k] > 1.4 and errs[1][k] / errs[2][k] > 1.4, (k, errs)
    assert errs[2]["Vx"] < 1e-3 and errs[2]["P"] < 4e-2
    new = ex.solve(32, Formulation="SymmetricGradient", new_loop=True, verbose=False, exact_bc=True, η=eta, tol=1e-10)["errors"]
    assert all(abs(new[k] - errs[1][k]) <= 1e-5 * errs[1][k] for k in new), (new, errs[1])


@pytest.mark.parametrize("case", [cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True), cases.Case("H2", "solkz", 1, "SymmetricGradient", tau_r=10.0)],
                         ids=lambda c: c.name)
def test_fused_pass_equals_the_three_calls(case, api):
    """fcfv_pass (one call, every array uploaded once, outputs coming down while inputs go up) == fcfv_compute_local +
    fcfv_assemble + fcfv_residuals, bit for bit, with pageable and with registered host arrays."""
    from fcfv_nme23_b200 import lib as L
    m, d = cases.build(case)
    neu = d["neu"]
    a, b, Z = api.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    fn = api.ElementAssemblyLoop if case.variant == 0 else api.ElementAssemblyLoopNEW
    h, nnz, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form, want_muu=False, export=False)
    ref = [h.export_csr(L.KUU), h.export_csr(L.KUP), h.export_rhs()]
    n = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu,
                                           case.form, verbose=False)
    tau = m.τ.copy()
    h.close()
    for pinned in (False, True):
        g, _ = cases.build(case)
        hg = api.handle_for(g)
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, d["Vxh"], d["Vyh"], d["Pe"])]
        if pinned:
            hg.host_register(*arrs)
        a2, b2, Z2, n2, nnz2 = api.FusedPass(g, *arrs, case.form, loop="NEW" if case.variant else "old")
        assert np.array_equal(a2, a) and np.array_equal(b2, b) and np.array_equal(Z2, Z) and np.array_equal(g.τ, tau)
        assert nnz2[:2] == tuple(nnz)[:2] and np.array_equal(n2, n)
        got = [hg.export_csr(L.KUU), hg.export_csr(L.KUP), hg.export_rhs()]
        for x, y in zip(ref, got):
            for p, q in zip(x, y):
                assert np.array_equal(p, q)
        if pinned:
            hg.host_unregister(*arrs)
        hg.close()


def test_async_muu_then_schur_keeps_both_counts(api):
    """fcfv_run_assemble(want_muu) followed by fcfv_schur without fetching the counts in between: Muu and the Schur
    block report their own sizes (separate count slots) and a second fcfv_set_mesh with another size invalidates
    the data set for the first mesh."""
    import ctypes as C
    from fcfv_nme23_b200 import lib as L
    h = api.Handle(0)
    h.generate_structured(24, setting="inclusion", tau_r=2.0)
    form = "SymmetricGradient"
    h.run_local(form); h.run_assemble(L.LOOP_OLD, form, want_muu=True)
    nnz_ref = h.get_nnz()
    h.set_async(True)
    h.run_assemble(L.LOOP_OLD, form, want_muu=True)             # counts not fetched
    nsc = C.c_int64(0)
    h.check(h.lib.fcfv_schur(h._h, C.c_double(1.0e4), C.byref(nsc)))
    h.set_async(False)
    assert h.block_size(L.MUU)[2] == nnz_ref[2] > 0
    assert h.block_size(L.KUUSC)[2] == nsc.value > nnz_ref[2]
    rp, ci, v = h.export_csr(L.MUU)
    assert rp[-1] == nnz_ref[2] and len(v) == nnz_ref[2]
    # a mesh of another size: the residual must refuse to run on the old data
    m = cases.load_mesh("S6")
    m.τe = np.ones(m.nel); m.ke = np.ones(m.nel)
    h.set_mesh(m)
    rc = h.lib.fcfv_run_local(h._h, api.formulation_id(form), C.c_double(1.0))
    assert rc != 0 and b"not set" in h.lib.fcfv_last_error(h._h)
    h.close()
