"""Upload cache of the host-buffer entries (fcfv_upload_cache / fcfv_new_epoch / fcfv_transfer_stats) and the
ComputeFCFV entry with the reference's full argument list (fcfv_compute_local_ex): results identical to the uncached
calls, bytes actually saved, stale data impossible after a new epoch and caught by the fingerprint on wholesale changes."""
import numpy as np
import pytest

import cases
from helpers import assert_csr_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from fcfv_nme23_b200 import api
    return api


@pytest.fixture(autouse=True)
def whole_face_arrays(monkeypatch):
    """The byte counts below are those of whole per-face arrays; the face lists (tests/test_sparse_face_data.py) would
    replace the VDir / Neumann uploads by a few values each."""
    monkeypatch.setenv("FCFV_SPARSE_FACE_DATA", "0")


def three_calls(api, mesh, d, case, mutate=None):
    """The three calls of a reference driver through the mirror; `mutate(d)` runs between ComputeFCFV and the assembly."""
    from fcfv_nme23_b200 import lib as L
    neu = d["neu"]
    a, b, Z = api.ComputeFCFV(mesh, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    if mutate:
        mutate(d, mesh)
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


CASES = [cases.Case("S17", "solkz", 1, "SymmetricGradient", tau_r=10.0), cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True),
         cases.Case("S40", "solkz", 1, "SymmetricGradient", tau_r=10.0)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_cached_calls_identical_and_cheaper(case, api, monkeypatch):
    monkeypatch.setattr(api, "CACHE_INPUTS", False)
    m0, d0 = cases.build(case)
    r0 = three_calls(api, m0, d0, case)
    h0 = api.handle_for(m0)
    up0, down0, hit0 = h0.transfer_stats()
    assert hit0 == 0
    monkeypatch.setattr(api, "CACHE_INPUTS", True)
    m1, d1 = cases.build(case)
    r1 = three_calls(api, m1, d1, case)
    h1 = api.handle_for(m1)
    up1, down1, hit1 = h1.transfer_stats()
    same(r0, r1)
    nel, nf = m1.nel, m1.nf
    # VxDir / VyDir twice, the four Neumann arrays twice (prefetched by ComputeFCFV), sources once, alpha / beta / Z twice,
    # ke / delta / taue / tau in the two later field refreshes
    assert m1.τe is not None
    expect = 8 * (2 * 2 * nf + 2 * 4 * nf + 2 * nel + 2 * 7 * nel + 2 * (3 * nel + nf))
    # the handle is made inside the first ComputeFCFV (fcfv_compute_fcfv with the mesh arrays): the mesh fields go up once
    first = 0
    assert hit1 == expect + first, (hit1, expect, first)
    assert up0 - up1 == hit1 - 8 * 4 * nf       # without the cache ComputeFCFV does not send the Neumann arrays at all
    assert down0 == down1
    # a second pass on the same mesh object: a new epoch, everything but the mesh goes up again
    r2 = three_calls(api, m1, d1, case)
    same(r0, r2)
    up2, _, hit2 = h1.transfer_stats()
    assert hit2 == 2 * expect + first
    h0.close(); h1.close()


def test_wholesale_change_is_caught_by_the_fingerprint(api, monkeypatch):
    case = CASES[2]
    def mutate(d, mesh):
        d["VxDir"] *= 1.5            # in place: same address, same size, same epoch
        d["neu"][0][:] += 0.25
    monkeypatch.setattr(api, "CACHE_INPUTS", False)
    m0, d0 = cases.build(case)
    r0 = three_calls(api, m0, d0, case, mutate)
    monkeypatch.setattr(api, "CACHE_INPUTS", True)
    m1, d1 = cases.build(case)
    r1 = three_calls(api, m1, d1, case, mutate)
    for k in ("fu", "fp", "n"):
        assert np.array_equal(r0[k], r1[k]), k
    assert_csr_equal(r0["Kuu"], r1["Kuu"], "Kuu")
    api.handle_for(m0).close(); api.handle_for(m1).close()


def test_new_epoch_forces_the_upload(api, monkeypatch):
    """One entry changed in place is below the fingerprint's resolution on a large array: the documented contract is
    that the caller starts a new epoch -- after which the change must arrive."""
    case = CASES[2]

    def mutate(d, mesh):
        dirichlet = np.flatnonzero(np.asarray(mesh.bc) == 1)
        d["VxDir"][int(dirichlet[len(dirichlet) // 2])] += 1.0
        api.handle_for(mesh).new_epoch()
    monkeypatch.setattr(api, "CACHE_INPUTS", False)
    m0, d0 = cases.build(case)
    r0 = three_calls(api, m0, d0, case, mutate)
    monkeypatch.setattr(api, "CACHE_INPUTS", True)
    m1, d1 = cases.build(case)
    r1 = three_calls(api, m1, d1, case, mutate)
    same(r0, r1)
    dirichlet = np.flatnonzero(np.asarray(m1.bc) == 1)
    f = int(dirichlet[len(dirichlet) // 2])
    assert r1["fu"][f] == d1["VxDir"][f]            # Dirichlet row: fu = VDir, the changed value
    api.handle_for(m0).close(); api.handle_for(m1).close()


def test_compute_local_ex_matches_and_prefetches(api):
    import ctypes as C
    from fcfv_nme23_b200 import lib as L
    case = CASES[1]
    m, d = cases.build(case)
    nel, nf = m.nel, m.nf
    h = api.Handle(0)
    h.set_mesh(m)
    form = api.formulation_id(case.form)
    ptr = lambda x: C.c_void_p(x.ctypes.data)
    out0 = [np.empty(nel), np.empty(2 * nel), np.empty(4 * nel), np.empty(nf)]
    out1 = [np.empty(nel), np.empty(2 * nel), np.empty(4 * nel), np.empty(nf)]
    h.check(h.lib.fcfv_compute_local(h._h, form, 1.0, ptr(d["sex"]), ptr(d["sey"]), ptr(d["VxDir"]), ptr(d["VyDir"]), *map(ptr, out0)))
    h.upload_cache(True)
    h.check(h.lib.fcfv_compute_local_ex(h._h, form, 1.0, ptr(d["sex"]), ptr(d["sey"]), ptr(d["VxDir"]), ptr(d["VyDir"]),
                                        *map(ptr, d["neu"]), *map(ptr, out1)))
    for x, y in zip(out0, out1):
        assert np.array_equal(x, y)
    # the Neumann arrays are resident: the assembly with them (cache hits) equals the assembly told to use what is resident
    nnz = [C.c_int64(0) for _ in range(3)]
    up_before = h.transfer_stats()[0]
    h.check(h.lib.fcfv_assemble(h._h, case.variant, form, 1.0, None, None, None, ptr(d["VxDir"]), ptr(d["VyDir"]), *map(ptr, d["neu"]), 0,
                                C.byref(nnz[0]), C.byref(nnz[1]), C.byref(nnz[2]), None))
    assert h.transfer_stats()[0] == up_before          # nothing crossed the link
    fu1, fp1 = h.export_rhs()
    K1 = h.export_csr(L.KUU)
    h.upload_cache(False)
    h.check(h.lib.fcfv_assemble(h._h, case.variant, form, 1.0, None, None, None, ptr(d["VxDir"]), ptr(d["VyDir"]), *map(ptr, d["neu"]), 0,
                                C.byref(nnz[0]), C.byref(nnz[1]), C.byref(nnz[2]), None))
    assert h.transfer_stats()[0] == up_before + 8 * 6 * nf
    fu0, fp0 = h.export_rhs()
    assert np.array_equal(fu0, fu1) and np.array_equal(fp0, fp1)
    assert_csr_equal(K1, h.export_csr(L.KUU), "Kuu")
    h.close()
