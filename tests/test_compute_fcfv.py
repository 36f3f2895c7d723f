"""fcfv_compute_fcfv: ComputeFCFV(mesh, ...) as one library call -- mesh (or field refresh) + element kernel, as a pipeline
over element chunks when every array is page-locked.  Results are those of fcfv_set_mesh / fcfv_update_fields followed by
fcfv_compute_local_ex, bit for bit, on both routes; the calls that follow find the same device state."""
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


def pinned(a):
    """Page-locked copy of a NumPy array (column-major flattening, the reference's layout) as a 1-D NumPy view."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).pin_memory()
    return t, t.numpy()


def ptr(v):
    return None if v is None else C.c_void_p(v.ctypes.data)


def mesh_arrays(m):
    taue = m.τe if m.τe is not None else np.zeros(m.nel)
    return dict(e2f=np.asarray(m.e2f, dtype=np.int64), bc=np.asarray(m.bc, dtype=np.int64), Om=m.Ω, nx=m.n_x, ny=m.n_y, G=m.Γ, ke=m.ke, dl=m.δ,
                taue=np.asarray(taue, dtype=np.float64))


def rest_of_the_pass(api, h, case, d, loc):
    """Assembly + residuals on whatever the handle holds after ComputeFCFV."""
    from fcfv_nme23_b200 import lib as L
    form = api.formulation_id(case.form)
    nnz = [C.c_int64(0) for _ in range(3)]
    sig = [ptr(np.ascontiguousarray(s)) for s in d["neu"]]
    h.check(h.lib.fcfv_assemble(h._h, case.variant, form, 1.0, *[ptr(x) for x in loc], ptr(d["VxDir"]), ptr(d["VyDir"]), *sig, 1,
                                C.byref(nnz[0]), C.byref(nnz[1]), C.byref(nnz[2]), None))
    K, P, M = h.export_csr(L.KUU), h.export_csr(L.KUP), h.export_csr(L.MUU)
    fu, fp = h.export_rhs()
    norms = np.zeros(6)
    h.check(h.lib.fcfv_residuals(h._h, form, L.TAU_FACE, ptr(d["Vxh"]), ptr(d["Vyh"]), ptr(d["Pe"]), *[ptr(x) for x in loc],
                                 ptr(d["sex"]), ptr(d["sey"]), ptr(d["VxDir"]), ptr(d["VyDir"]), *sig, ptr(norms), None, None, None))
    return dict(K=K, P=P, M=M, fu=fu, fp=fp, norms=norms)


def same_pass(r0, r1):
    for k in ("K", "P", "M"):
        assert_csr_equal(r0[k], r1[k], k)
    for k in ("fu", "fp", "norms"):
        assert np.array_equal(r0[k], r1[k]), k


def build(case, random_seed=None):
    if random_seed is not None:
        from test_random_meshes import random_case
        return random_case(random_seed)
    return cases.build(case)


CASES = [(cases.Case("S17", "solkz", 1, "SymmetricGradient", tau_r=10.0), None),
         (cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True), None),
         (cases.Case("H2", "inclusion", 1, "Gradient"), None),
         (cases.Case("random", "random", 1, "SymmetricGradient", tau_r=2.0), 7)]     # faces no element references: tau_in is read


@pytest.mark.parametrize("sparse", ["0", "2"])
@pytest.mark.parametrize("case,seed", CASES, ids=lambda c: getattr(c, "name", str(c)))
def test_pipelined_call_equals_the_separate_calls(case, seed, sparse, api, monkeypatch):
    monkeypatch.setenv("FCFV_SPARSE_FACE_DATA", sparse)
    form = api.formulation_id(case.form)
    # ---- reference route: set_mesh, tau refresh, compute_local_ex, from pageable arrays
    m, d = build(case, seed)
    nel, nf = m.nel, m.nf
    tau0 = np.linspace(1.0, 2.0, nf)                 # mesh.tau as the caller holds it: survives on faces without elements
    h0 = api.Handle(0)
    h0.upload_cache(True)
    m.τ = tau0.copy()
    h0.set_mesh(m, tau=False)
    out0 = [np.empty(nel), np.empty(2 * nel), np.empty(4 * nel), np.empty(nf)]
    h0.check(h0.lib.fcfv_compute_local_ex(h0._h, form, 1.0, ptr(d["sex"]), ptr(d["sey"]), ptr(d["VxDir"]), ptr(d["VyDir"]),
                                          *[ptr(np.ascontiguousarray(s)) for s in d["neu"]], *[ptr(x) for x in out0]))
    assert h0.pipelined_calls == 0
    r0 = rest_of_the_pass(api, h0, case, d, out0[:3])
    # ---- one call, everything page-locked, pipeline forced on a small mesh
    monkeypatch.setenv("FCFV_PIPELINE_MIN_NEL", "1")
    keep = []
    def pin(a):
        t, v = pinned(a); keep.append(t); return v
    A = {k: pin(v) for k, v in mesh_arrays(m).items()}
    D = {k: pin(d[k]) for k in ("sex", "sey", "VxDir", "VyDir", "Vxh", "Vyh", "Pe")}
    D["neu"] = [pin(s) for s in d["neu"]]
    tin = pin(tau0)
    out1 = [pin(np.zeros(nel)), pin(np.zeros(2 * nel)), pin(np.zeros(4 * nel)), pin(np.zeros(nf))]
    h1 = api.Handle(0)
    h1.upload_cache(True)
    def call(h, mesh=True, A=A, D=D, out=out1):
        geo = [ptr(A[k]) for k in ("e2f", "bc", "Om", "nx", "ny", "G")] if mesh else [None] * 6
        h.check(h.lib.fcfv_compute_fcfv(h._h, nel, nf, *geo, ptr(A["ke"]), ptr(A["dl"]), ptr(A["taue"]), A["taue"].size, ptr(tin), form, 1.0,
                                        ptr(D["sex"]), ptr(D["sey"]), ptr(D["VxDir"]), ptr(D["VyDir"]), *[ptr(s) for s in D["neu"]],
                                        *[ptr(x) for x in out]))
    call(h1)
    h1.nel, h1.nf = nel, nf
    assert h1.pipelined_calls == 1
    for x, y in zip(out0, out1):
        assert np.array_equal(x, y)
    if seed is not None:                             # the random mesh has faces no element references: they keep the caller's tau
        flag = C.c_int(0)
        h1.check(h1.lib.fcfv_has_orphan_faces(h1._h, C.byref(flag)))
        assert flag.value == 1 and np.count_nonzero(out1[3] == tau0) > 0
    up = h1.transfer_stats()[0]
    r1 = rest_of_the_pass(api, h1, case, D, out1[:3])
    same_pass(r0, r1)
    # alpha / beta / Z (downloaded mirrors) and sex / sey (chunk uploads noted as mirrors) did not go up again
    sent = h1.transfer_stats()[0] - up
    # whole boundary arrays are mirrors since ComputeFCFV (cache hits); lists are short and sent by every call
    bcs = np.asarray(m.bc)
    face = 0 if sparse == "0" else 8 * 2 * (2 * np.count_nonzero(bcs == 1) + 4 * np.count_nonzero(bcs == 2))
    assert sent == 8 * (2 * nf + nel) + face, (sent, face)        # Vxh, Vyh, Pe (+ the boundary values of the two calls)
    # ---- second pass on the resident mesh with changed fields and sources: field refresh inside the call
    h1.new_epoch(); h0.new_epoch()
    A["ke"] *= 1.5; A["taue"] *= 0.5; D["sex"] += 0.25
    m.ke = m.ke * 1.5; m.τe = np.asarray(mesh_arrays(m)["taue"]) * 0.5; d["sex"] = d["sex"] + 0.25
    call(h1, mesh=False)
    assert h1.pipelined_calls == 2
    h0.update_fields(ke=m.ke, delta=m.δ, taue=m.τe)
    h0.check(h0.lib.fcfv_compute_local_ex(h0._h, form, 1.0, ptr(d["sex"]), ptr(d["sey"]), ptr(d["VxDir"]), ptr(d["VyDir"]),
                                          *[ptr(np.ascontiguousarray(s)) for s in d["neu"]], *[ptr(x) for x in out0]))
    for x, y in zip(out0, out1):
        assert np.array_equal(x, y)
    same_pass(rest_of_the_pass(api, h0, case, d, out0[:3]), rest_of_the_pass(api, h1, case, D, out1[:3]))
    h0.close(); h1.close()


def test_pageable_arrays_take_the_separate_calls(api, monkeypatch):
    """Plain NumPy arrays cannot be copied asynchronously: the call issues fcfv_set_mesh + fcfv_compute_local_ex."""
    monkeypatch.setenv("FCFV_PIPELINE_MIN_NEL", "1")
    case = CASES[0][0]
    m, d = cases.build(case)
    g, _ = cases.build(case)
    neu = d["neu"]
    a, b, Z = api.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    h = api.handle_for(m)
    assert h.pipelined_calls == 0 and h.nel == m.nel
    hg = api.Handle(0)
    hg.set_mesh(g)
    out = [np.empty(g.nel), np.empty(2 * g.nel), np.empty(4 * g.nel), np.empty(g.nf)]
    hg.check(hg.lib.fcfv_compute_local(hg._h, api.formulation_id(case.form), 1.0, ptr(d["sex"]), ptr(d["sey"]), ptr(d["VxDir"]),
                                       ptr(d["VyDir"]), *[ptr(x) for x in out]))
    assert np.array_equal(a, out[0]) and np.array_equal(b.reshape(-1, order="F"), out[1]) and np.array_equal(Z.reshape(-1, order="F"), out[2])
    assert np.array_equal(m.τ, out[3])
    # second ComputeFCFV on the same mesh object: resident route (field refresh inside the call)
    m.ke = m.ke * 2.0
    a2, b2, Z2 = api.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    assert np.array_equal(a2, a)            # alpha, beta, Z do not depend on ke
    h.close(); hg.close()


def test_argument_errors(api):
    h = api.Handle(0)
    z = np.zeros(8)
    rc = h.lib.fcfv_compute_fcfv(h._h, 0, 0, *[None] * 6, None, None, None, 0, None, 0, 1.0, ptr(z), ptr(z), ptr(z), ptr(z), *[None] * 4,
                                 *[None] * 4)
    assert rc != 0 and b"no mesh resident" in h.lib.fcfv_last_error(h._h)
    e2f = np.ones(3, dtype=np.int64)
    rc = h.lib.fcfv_compute_fcfv(h._h, 1, 3, ptr(e2f), None, *[None] * 4, None, None, None, 0, None, 0, 1.0, ptr(z), ptr(z), ptr(z), ptr(z),
                                 *[None] * 4, *[None] * 4)
    assert rc != 0 and b"null mesh array" in h.lib.fcfv_last_error(h._h)
    h.close()


def test_default_threshold_on_a_million_elements(api):
    """S(520): 1 081 600 elements, above the library's own pipeline threshold (2^20) -- eight chunks of 135 296 elements, arrays
    of 8.7-26 MB, no test hook: the page-locked call gives the results of the separate calls from pageable memory."""
    import torch
    from fcfv_nme23_b200.mesh import FCFV_Mesh
    h = api.Handle(0)
    h.generate_structured(520, setting="inclusion", tau_r=2.0, neumann_south=True)
    M0, D0 = h.get_mesh(), h.get_data()
    h.close()
    nel, nf = M0["Ω"].size, M0["bc"].size
    assert nel >= 1 << 20
    m = FCFV_Mesh(nel=nel, nf=nf, nf_el=3)
    for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe"):
        setattr(m, k, M0[k])
    neu = [D0[k] for k in ("sxx", "syy", "sxy", "syx")]
    a, b, Z = api.ComputeFCFV(m, D0["sex"], D0["sey"], D0["VxDir"], D0["VyDir"], *neu, 2.0, "Gradient")     # pageable: separate calls
    h0 = api.handle_for(m)
    assert h0.pipelined_calls == 0
    tau0 = m.τ.copy()
    h0.close()
    keep = []
    def pin(x):
        t, v = pinned(x); keep.append(t); return v
    A = {k: pin(M0[k]) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe")}
    D = {k: pin(v) for k, v in D0.items()}
    out = [pin(np.zeros(nel)), pin(np.zeros(2 * nel)), pin(np.zeros(4 * nel)), pin(np.zeros(nf))]
    h1 = api.Handle(0)
    h1.upload_cache(True)
    h1.check(h1.lib.fcfv_compute_fcfv(h1._h, nel, nf, *[ptr(A[k]) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe")], nel, None,
                                      api.formulation_id("Gradient"), 1.0, ptr(D["sex"]), ptr(D["sey"]), ptr(D["VxDir"]), ptr(D["VyDir"]),
                                      *[ptr(D[k]) for k in ("sxx", "syy", "sxy", "syx")], *[ptr(x) for x in out]))
    assert h1.pipelined_calls == 1
    assert np.array_equal(out[0], a) and np.array_equal(out[1], b.reshape(-1, order="F")) and np.array_equal(out[2], Z.reshape(-1, order="F"))
    assert np.array_equal(out[3], tau0)
    h1.close()


def test_fcfv_pass_takes_the_chunk_pipeline_with_page_locked_arrays(api, monkeypatch):
    """fcfv_pass (the three calls in one): its ComputeFCFV part runs as the chunk pipeline when the arrays allow it; same
    outputs as from pageable arrays."""
    from fcfv_nme23_b200 import lib as L
    case = cases.Case("H2", "inclusion", 1, "SymmetricGradient", neumann=True)
    form = api.formulation_id(case.form)
    res = []
    for pinned_arrays in (False, True):
        monkeypatch.setenv("FCFV_PIPELINE_MIN_NEL", "1")
        m, d = cases.build(case)
        nel, nf = m.nel, m.nf
        keep = []
        def conv(x):
            if not pinned_arrays:
                return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1, order="F"))
            t, v = pinned(np.asarray(x, dtype=np.float64)); keep.append(t); return v
        h = api.Handle(0)
        h.upload_cache(True)
        h.set_mesh(m)
        D = {k: conv(d[k]) for k in ("sex", "sey", "VxDir", "VyDir", "Vxh", "Vyh", "Pe")}
        S = [conv(s) for s in d["neu"]]
        out = [conv(np.zeros(nel)), conv(np.zeros(2 * nel)), conv(np.zeros(4 * nel)), conv(np.zeros(nf))]
        nnz = [C.c_int64(0) for _ in range(3)]
        norms = np.zeros(6)
        h.check(h.lib.fcfv_pass(h._h, case.variant, form, 1.0, L.TAU_FACE, ptr(D["sex"]), ptr(D["sey"]), ptr(D["VxDir"]), ptr(D["VyDir"]),
                                *[ptr(s) for s in S], ptr(D["Vxh"]), ptr(D["Vyh"]), ptr(D["Pe"]), *[ptr(x) for x in out], 1,
                                C.byref(nnz[0]), C.byref(nnz[1]), C.byref(nnz[2]), ptr(norms)))
        assert h.pipelined_calls == (1 if pinned_arrays else 0)
        fu, fp = h.export_rhs()
        res.append(dict(out=[x.copy() for x in out], K=h.export_csr(L.KUU), P=h.export_csr(L.KUP), M=h.export_csr(L.MUU), fu=fu, fp=fp,
                        norms=norms.copy(), nnz=[x.value for x in nnz]))
        h.close()
    a, b = res
    for x, y in zip(a["out"], b["out"]):
        assert np.array_equal(x, y)
    for k in ("K", "P", "M"):
        assert_csr_equal(a[k], b[k], k)
    assert np.array_equal(a["fu"], b["fu"]) and np.array_equal(a["fp"], b["fp"]) and np.array_equal(a["norms"], b["norms"]) and a["nnz"] == b["nnz"]


def test_mirror_with_registered_arrays_takes_the_pipeline(api, monkeypatch):
    """api.ComputeFCFV(..., out=(α, β, Ζ, τ)) on NumPy arrays that were page-locked with host_register: the call reaches the
    chunk pipeline through the mirror (what ComputeFCFV! + PinArrays give a Julia driver)."""
    monkeypatch.setenv("FCFV_PIPELINE_MIN_NEL", "1")
    case = cases.Case("H2", "solkz", 1, "SymmetricGradient", tau_r=10.0)
    m0, d0 = cases.build(case)
    ref = api.ComputeFCFV(m0, d0["sex"], d0["sey"], d0["VxDir"], d0["VyDir"], *d0["neu"], case.tau_r, case.form)
    assert api.handle_for(m0).pipelined_calls == 0
    m, d = cases.build(case)
    for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe"):
        v = np.asarray(getattr(m, k))
        setattr(m, k, np.asfortranarray(v, dtype=np.int64 if k in ("e2f", "bc") else np.float64))
    m.τ = np.zeros(m.nf)
    out = (np.zeros(m.nel), np.zeros((m.nel, 2), order="F"), np.zeros((m.nel, 2, 2), order="F"), np.zeros(m.nf))
    reg = api.Handle(0)
    arrays = [getattr(m, k) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe", "τ")] + [d[k] for k in ("sex", "sey", "VxDir", "VyDir")] + list(out)
    reg.host_register(*arrays)
    try:
        a, b, Z = api.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], case.tau_r, case.form, out=out)
        h = api.handle_for(m)
        assert h.pipelined_calls == 1 and a is out[0] and m.τ is out[3]
        for x, y in zip((a, b, Z), ref):
            assert np.array_equal(x, y)
        assert np.array_equal(m.τ, m0.τ)
        h.close()
    finally:
        reg.host_unregister(*arrays)
        reg.close()
    api.handle_for(m0).close()
