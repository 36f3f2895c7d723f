"""Per-face boundary data through face lists (fcfv_sparse_face_data): VxDir / VyDir are read on Dirichlet faces only and
the Neumann arrays on Neumann faces only (Bool factors / branches of the reference: DiscretisationFCFV_Stokes.jl:27-37,
167-178,191-194, ComputeResidualsFCFV_Stokes.jl:57-66), so the library sends their values on those faces instead of
nf doubles.  Results must not depend on the mode, nor on what the arrays hold anywhere else."""
import ctypes as C

import numpy as np
import pytest

import cases
from helpers import assert_csr_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from fcfv_nme23_b200 import api
    return api


def three_calls(api, mesh, d, case):
    from fcfv_nme23_b200 import lib as L
    neu = d["neu"]
    a, b, Z = api.ComputeFCFV(mesh, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    loop = api.ElementAssemblyLoopNEW if case.variant else api.ElementAssemblyLoop
    h, nnz, _ = loop(mesh, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form, export=False)
    Kuu, Kup = h.export_csr(L.KUU), h.export_csr(L.KUP)
    fu, fp = h.export_rhs()
    n = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], mesh, a, b, Z, d["sex"], d["sey"], d["VxDir"], d["VyDir"],
                                           *neu, case.form, verbose=False)
    return dict(a=a, b=b, Z=Z, tau=mesh.τ.copy(), Kuu=Kuu, Kup=Kup, fu=fu, fp=fp, n=n)


def same(r0, r1):
    for k in ("a", "b", "Z", "tau", "fu", "fp", "n"):
        assert np.array_equal(r0[k], r1[k]), k
    assert_csr_equal(r0["Kuu"], r1["Kuu"], "Kuu")
    assert_csr_equal(r0["Kup"], r1["Kup"], "Kup")


def poison(d, mesh):
    """NaN wherever the reference's Bool factors / branches make the value irrelevant."""
    bc = np.asarray(mesh.bc)
    for k in ("VxDir", "VyDir"):
        d[k] = np.where(bc == 1, d[k], np.nan)
    d["neu"] = [np.where(bc == 2, s, np.nan) for s in d["neu"]]


CASES = [cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True), cases.Case("S17", "inclusion", 1, "Gradient", neumann=True),
         cases.Case("H2", "solkz", 1, "SymmetricGradient", tau_r=10.0), cases.Case("S40", "solkz", 0, "SymmetricGradient", tau_r=10.0)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_lists_give_the_dense_result(case, api, monkeypatch):
    res = {}
    for tag, mode, bad in (("dense", "0", False), ("listed", "2", False), ("listed-poisoned", "2", True), ("dense-poisoned", "0", True)):
        monkeypatch.setenv("FCFV_SPARSE_FACE_DATA", mode)
        m, d = cases.build(case)
        if bad:
            poison(d, m)
        res[tag] = three_calls(api, m, d, case)
        h = api.handle_for(m)
        act, nd, nn, ups = h.sparse_face_info()
        bc = np.asarray(m.bc)
        if mode == "2":
            assert act and nd == np.count_nonzero(bc == 1) and nn == np.count_nonzero(bc == 2) and ups > 0
        else:
            assert not act and nd == -1 and nn == -1 and ups == 0
        h.close()
    for tag in ("listed", "listed-poisoned", "dense-poisoned"):
        same(res["dense"], res[tag])


def test_only_the_listed_values_cross_the_link(api):
    case = cases.Case("S40", "inclusion", 1, "SymmetricGradient", neumann=True)
    m, d = cases.build(case)
    bc = np.asarray(m.bc)
    nd, nn = np.count_nonzero(bc == 1), np.count_nonzero(bc == 2)
    assert 0 < nn and 16 * (nd + nn) < m.nf         # auto mode takes the lists on this mesh
    h = api.Handle(0)
    h.set_mesh(m)
    assert h.sparse_face_info()[:3] == (True, nd, nn)
    ptr = lambda x: C.c_void_p(np.ascontiguousarray(x).ctypes.data)
    h.check(h.lib.fcfv_set_sources(h._h, ptr(d["sex"]), ptr(d["sey"])))           # fcfv_get_data wants every array set
    h.check(h.lib.fcfv_set_solution(h._h, ptr(d["Vxh"]), ptr(d["Vyh"]), ptr(d["Pe"])))
    up0 = h.transfer_stats()[0]
    h.check(h.lib.fcfv_set_face_data(h._h, ptr(d["VxDir"]), ptr(d["VyDir"]), *[ptr(s) for s in d["neu"]]))
    assert h.transfer_stats()[0] - up0 == 8 * (2 * nd + 4 * nn)
    got = h.get_data()
    for k, mask in (("VxDir", bc == 1), ("VyDir", bc == 1)):
        assert np.array_equal(got[k][mask], d[k][mask]) and not got[k][~mask].any()
    for k, s in zip(("sxx", "syy", "sxy", "syx"), d["neu"]):
        assert np.array_equal(got[k][bc == 2], s[bc == 2]) and not got[k][bc != 2].any()
    # a second upload after the values changed in place: the lists are always sent (no cache to go stale)
    d["VxDir"][bc == 1] += 1.0
    h.check(h.lib.fcfv_set_face_data(h._h, ptr(d["VxDir"]), None, None, None, None, None))
    assert np.array_equal(h.get_data()["VxDir"][bc == 1], d["VxDir"][bc == 1])
    # switching the mode on a handle that holds a mesh rebuilds / drops the lists
    h.sparse_face_data(0)
    assert h.sparse_face_info()[0] is False
    up1 = h.transfer_stats()[0]
    h.check(h.lib.fcfv_set_face_data(h._h, ptr(d["VxDir"]), None, None, None, None, None))
    assert h.transfer_stats()[0] - up1 == 8 * m.nf
    h.sparse_face_data(1)
    assert h.sparse_face_info()[:3] == (True, nd, nn)
    with pytest.raises(Exception):
        h.sparse_face_data(3)
    h.close()


def test_auto_mode_keeps_whole_arrays_when_the_lists_are_long(api):
    case = cases.Case("S6", "solkz", 1, "SymmetricGradient", tau_r=10.0)
    m, d = cases.build(case)
    assert 16 * np.count_nonzero(np.asarray(m.bc) == 1) > m.nf
    h = api.Handle(0)
    h.set_mesh(m)
    assert h.sparse_face_info()[0] is False
    h.close()


def test_generated_mesh_has_lists(api):
    h = api.Handle(0)
    h.generate_structured(64, setting="inclusion", tau_r=2.0, neumann_south=True)
    bc = h.get_mesh()["bc"]
    assert h.sparse_face_info()[:3] == (True, np.count_nonzero(bc == 1), np.count_nonzero(bc == 2))
    h.close()


@pytest.mark.parametrize("ngpus", [2, 3])
def test_multi_gpu_handle_with_lists(ngpus, api, monkeypatch):
    """Every part lists its own boundary faces; the values go from the GLOBAL arrays to the parts without any nf-sized
    scatter.  Same results as one GPU with whole arrays."""
    import test_multi_handle as T
    if T._ngpus() < ngpus:
        monkeypatch.setenv("FCFV_MULTI_OVERSUBSCRIBE", "1")
    case = cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True)
    monkeypatch.setenv("FCFV_SPARSE_FACE_DATA", "0")
    ref = T._run(case, 1)
    monkeypatch.setenv("FCFV_SPARSE_FACE_DATA", "2")
    m, d = cases.build(case)
    poison(d, m)
    got = T._run(case, ngpus, (m, d))
    for k in ("a", "b", "Z", "tau", "fu", "fp"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("Kuu", "Kup", "Muu", "Kuusc"):
        for x, y in zip(got[k], ref[k]):
            assert np.array_equal(x, y), k
    for mode in ("REFERENCE", "FACE"):
        n, r = got["norms"][mode], ref["norms"][mode]
        assert np.all(np.abs(n - r) <= 1e-12 * r.max()), mode
