"""Several GPUs behind ONE handle (fcfv_create_multi): the reference-facing calls on the GLOBAL arrays, against the
same calls on one GPU.  alpha / beta / Z / tau, the Julia CSC fields of Kuu / Kup / Muu / Kuusc, fu, fp: bit for
bit; the six residual norms (one grouped NCCL all-reduce): 1e-12.

With fewer physical GPUs than parts the test runs in the library's oversubscription mode (parts share devices, the
all-reduce becomes a host sum): partitioner, scatter, ghost replication and export are exercised on any GPU box; the
NCCL path itself needs as many devices as parts (gpurun --gpus N)."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _run(case, ngpus, mesh_data=None):
    from fcfv_nme23_b200 import api, lib as L
    g, d = mesh_data if mesh_data is not None else cases.build(case)
    g.ngpus = ngpus
    neu = d["neu"]
    a, b, Z = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    fn = api.ElementAssemblyLoop if case.variant == 0 else api.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, ts = fn(g, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form)
    h = api.handle_for(g)
    assert h.lib.fcfv_num_devices(h._h) == ngpus and ts > 0.0
    out = dict(a=a, b=b, Z=Z, tau=g.τ.copy(), fu=fu, fp=fp, nnz=(Kuu.nnz, Muu.nnz, Kup.nnz))
    for name, blk in (("Kuu", L.KUU), ("Kup", L.KUP), ("Muu", L.MUU)):
        out[name] = h.export_csc(blk)
        out[name + "_csr"] = h.export_csr(blk)
    out["norms"] = {m: api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], g, a, b, Z, d["sex"], d["sey"], d["VxDir"],
                                                          d["VyDir"], *neu, case.form, tau_mode=m, verbose=False)
                    for m in ("REFERENCE", "FACE")}
    sc = api.SchurComplement(g, 1.0e4)
    out["Kuusc"] = h.export_csc(L.KUUSC)
    out["nnz_sc"] = sc.nnz
    # post-processing and the Powell-Hestenes lines (global arrays in and out on either kind of handle)
    rng = np.random.default_rng(11)
    u, p = rng.standard_normal(2 * g.nf), rng.standard_normal(g.nel)
    out["ev"] = api.ComputeElementValues(g, d["Vxh"], d["Vyh"], d["Pe"], a, b, Z, d["VxDir"], d["VyDir"], case.form)
    out["ph_res"] = api.ph_residual(g, u, p)
    out["ph_p"] = api.ph_update_p(g, 1.0e3, u, p)
    out["ph_fusc"] = api.ph_fusc(g, 1.0e3, p)
    out["center"] = api.center_pressure(g, d["Pe"])
    if ngpus > 1:
        info, sec = h.partition_info()
        assert sum(i[2] for i in info) == g.nel and all(i[0] >= i[2] and i[1] >= i[3] for i in info) and sec >= 0.0
        out["ghosts"] = sum(i[0] - i[2] for i in info)
    h.close()
    return out


def same_post_processing(got, ref):
    """Element values, ru / rp / p update / fusc: every entry is computed by the part that owns it from the same operands
    (bit for bit); norms and the pressure mean are sums taken in another order (1e-12)."""
    for x, y in zip(got["ev"], ref["ev"]):
        assert np.array_equal(x, y)
    assert np.array_equal(got["ph_res"][0], ref["ph_res"][0]) and np.array_equal(got["ph_res"][1], ref["ph_res"][1])
    assert np.all(np.abs(got["ph_res"][2] - ref["ph_res"][2]) <= TOL * ref["ph_res"][2])
    assert np.array_equal(got["ph_p"], ref["ph_p"]) and np.array_equal(got["ph_fusc"], ref["ph_fusc"])
    scale = max(1.0, np.abs(ref["center"][0]).max())
    assert np.all(np.abs(got["center"][0] - ref["center"][0]) <= TOL * scale) and abs(got["center"][1] - ref["center"][1]) <= TOL * scale


CASES = [cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True), cases.Case("H2", "solkz", 1, "SymmetricGradient", tau_r=10.0),
         cases.Case("S17", "inclusion", 1, "Gradient"), cases.Case("MedRes", "inclusion", 0, "SymmetricGradient")]


@pytest.mark.parametrize("ngpus", [2, 4, 8])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_multi_handle_equals_single_gpu(case, ngpus, monkeypatch):
    if _ngpus() < ngpus:
        monkeypatch.setenv("FCFV_MULTI_OVERSUBSCRIBE", "1")
    ref = _run(case, 1)
    got = _run(case, ngpus)
    for k in ("a", "b", "Z", "tau", "fu", "fp"):
        assert np.array_equal(got[k], ref[k]), k
    assert got["nnz"] == ref["nnz"] and got["nnz_sc"] == ref["nnz_sc"]
    for k in ("Kuu", "Kup", "Muu", "Kuusc", "Kuu_csr", "Kup_csr", "Muu_csr"):
        for x, y in zip(got[k], ref[k]):
            assert np.array_equal(x, y), k
    for m in ("REFERENCE", "FACE"):
        n, r = got["norms"][m], ref["norms"][m]
        big = r > 1e-9 * r.max()
        assert np.all(np.abs(n[big] - r[big]) <= TOL * r[big]) and np.all(np.abs(n[~big] - r[~big]) <= TOL * r.max()), m
    assert got["ghosts"] > 0
    same_post_processing(got, ref)


@pytest.mark.parametrize("ngpus", [2, 3])
def test_multi_handle_random_mesh(ngpus, monkeypatch):
    """Random Delaunay mesh, random numbering and tags, faces no element references."""
    if _ngpus() < ngpus:
        monkeypatch.setenv("FCFV_MULTI_OVERSUBSCRIBE", "1")
    from test_random_meshes import random_case
    case = cases.Case("random", "random", 1, "SymmetricGradient", tau_r=2.0)
    ref = _run(case, 1, random_case(5))
    got = _run(case, ngpus, random_case(5))
    for k in ("a", "b", "Z", "tau", "fu", "fp"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("Kuu", "Kup", "Muu", "Kuusc"):
        for x, y in zip(got[k], ref[k]):
            assert np.array_equal(x, y), k
    same_post_processing(got, ref)


@pytest.mark.parametrize("max_runs", ["0", "3"])
def test_multi_handle_gather_path(max_runs, monkeypatch):
    """The scatter goes straight from the caller's arrays when a part is a few runs of consecutive ids, and through a
    gathered temporary otherwise: both give the single-GPU result (FCFV_MULTI_MAX_RUNS forces the choice)."""
    if _ngpus() < 2:
        monkeypatch.setenv("FCFV_MULTI_OVERSUBSCRIBE", "1")
    case = CASES[1]
    ref = _run(case, 1)
    monkeypatch.setenv("FCFV_MULTI_MAX_RUNS", max_runs)
    got = _run(case, 2)
    for k in ("a", "b", "Z", "tau", "fu", "fp"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("Kuu", "Kup", "Muu", "Kuusc"):
        for x, y in zip(got[k], ref[k]):
            assert np.array_equal(x, y), k
    n, r = got["norms"]["FACE"], ref["norms"]["FACE"]
    assert np.all(np.abs(n - r) <= TOL * r.max())


def test_multi_handle_rejects_what_it_cannot_do(monkeypatch):
    import ctypes as C
    from fcfv_nme23_b200 import api, lib as L
    if _ngpus() < 2:
        monkeypatch.setenv("FCFV_MULTI_OVERSUBSCRIBE", "1")
    h = api.Handle(ngpus=2)
    with pytest.raises(L.FCFVError, match="multi-GPU"):
        h.generate_structured(8)
    p = C.c_void_p()
    assert h.lib.fcfv_device_csr(h._h, L.KUU, C.byref(p), None, None) != 0
    h.close()
