"""Golden vectors from the reference's own source (tests/golden/golden_reference.json, written by
tests/golden/make_reference_golden.py through the rule-table transliteration of /root/reference/src/*.jl) against
(1) the oracle, recomputed here, and (2) with -m gpu the CUDA path through the C ABI.  Bit for bit (SHA-256 of the
raw arrays) for alpha / beta / Z / tau, the CSC fields of Kuu / Muu / Kup, fu / fp and ComputeElementValues; 1e-12 for
the six residual norms.  Runs without the reference checkout."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

TOL = 1e-12
BITWISE_KEYS = ("local", "Kuu", "Muu", "Kup", "rhs", "elem_values")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


with open(os.path.join(cases.GOLDEN, "golden_reference.json")) as _fh:
    REF = {k: v for k, v in json.load(_fh).items() if not k.startswith("_")}
ALL = {c.name: c for c in cases.all_cases() + cases.large_cases()}


def _norms_close(got, ref):
    got, ref = np.asarray(got), np.asarray(ref)
    scale = max(ref.max(), got.max())
    big = ref > 1e-9 * scale
    return np.all(np.abs(got[big] - ref[big]) <= TOL * ref[big]) and np.all(np.abs(got[~big] - ref[~big]) <= TOL * scale)


def test_every_case_has_a_reference_record():
    assert set(ALL) == set(REF)
    # regression constants of SURVEY.md Appendix C (nnz after dropzeros), now confirmed by the reference's own code
    assert (REF["H3-solkz-NEW-SymmetricGradient"]["nnz_uu"], REF["H3-solkz-NEW-SymmetricGradient"]["nnz_up"]) == (412684, 82844)
    assert (REF["H4-solkz-NEW-SymmetricGradient"]["nnz_uu"], REF["H4-solkz-NEW-SymmetricGradient"]["nnz_up"]) == (1655548, 331724)
    assert (REF["LowRes-inclusion-OLD-Gradient"]["nnz_uu"], REF["LowRes-inclusion-OLD-Gradient"]["nnz_up"]) == (29546, 11314)
    assert REF["MedRes-inclusion-OLD-SymmetricGradient"]["nnz_uu"] == 236052


def test_stored_oracle_digests_equal_the_reference():
    """golden_oracle.json (what the oracle produced when its fixtures were written) holds the same bytes."""
    with open(os.path.join(cases.GOLDEN, "golden_oracle.json")) as fh:
        orc = json.load(fh)
    common = sorted(set(orc) & set(REF))
    assert len(common) >= 48
    for name in common:
        for k in BITWISE_KEYS + ("nel", "nf", "nnz_uu", "nnz_muu", "nnz_up"):
            assert orc[name][k] == REF[name][k], (name, k)
        assert _norms_close(orc[name]["norms_reference"], REF[name]["norms_reference"]), name


@pytest.mark.parametrize("name", sorted(ALL), ids=str)
def test_oracle_reproduces_the_reference(name, oracle):
    case, gold = ALL[name], REF[name]
    m, d = cases.build(case)
    O = oracle
    a, b, Z = O.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], case.tau_r, case.form)
    fn = O.ElementAssemblyLoop if case.variant == 0 else O.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *d["neu"], case.form)
    assert (gold["nel"], gold["nf"]) == (m.nel, m.nf)
    assert (gold["nnz_uu"], gold["nnz_muu"], gold["nnz_up"]) == (len(Kuu[2]), len(Muu[2]), len(Kup[2]))
    assert gold["local"] == digest(a, b.ravel(order="F"), Z.ravel(order="F"), m.τ)
    assert gold["Kuu"] == digest(*Kuu) and gold["Muu"] == digest(*Muu) and gold["Kup"] == digest(*Kup)
    assert gold["rhs"] == digest(fu, fp)
    ev = O.ComputeElementValues(m, d["Vxh"], d["Vyh"], d["Pe"], a, b, Z, d["VxDir"], d["VyDir"], case.form)
    assert gold["elem_values"] == digest(*ev)
    n = O.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"], d["VyDir"],
                                         *d["neu"], case.form, tau_mode="REFERENCE")
    assert _norms_close(n, gold["norms_reference"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ALL), ids=str)
def test_cuda_path_reproduces_the_reference(name):
    """The product path (C ABI, CUDA kernels) hashes to the bytes the reference's own code produced."""
    from fcfv_nme23_b200 import api, lib as L
    case, gold = ALL[name], REF[name]
    g, d = cases.build(case)
    neu = d["neu"]
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    gfn = api.ElementAssemblyLoop if case.variant == 0 else api.ElementAssemblyLoopNEW
    gKuu, gMuu, gKup, gfu, gfp, _ = gfn(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, case.form)
    h = api.handle_for(g)
    assert (gold["nnz_uu"], gold["nnz_muu"], gold["nnz_up"]) == (gKuu.nnz, gMuu.nnz, gKup.nnz)
    assert gold["local"] == digest(ga, gb.ravel(order="F"), gZ.ravel(order="F"), g.τ)
    assert gold["Kuu"] == digest(*h.export_csc(L.KUU)) and gold["Kup"] == digest(*h.export_csc(L.KUP))
    assert gold["Muu"] == digest(*h.export_csc(L.MUU)) and gold["rhs"] == digest(gfu, gfp)
    gev = api.ComputeElementValues(g, d["Vxh"], d["Vyh"], d["Pe"], ga, gb, gZ, d["VxDir"], d["VyDir"], case.form)
    assert gold["elem_values"] == digest(*gev)
    gn = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], g, ga, gb, gZ, d["sex"], d["sey"], d["VxDir"],
                                            d["VyDir"], *neu, case.form, tau_mode="REFERENCE", verbose=False)
    assert _norms_close(gn, gold["norms_reference"])
    h.close()


# ---- BASELINE config 3 size (1 M elements): CUDA path against the oracle, bit for bit ---------------------------
C3_CASES = [cases.Case("S500", "solkz", 1, "SymmetricGradient", tau_r=10.0),                  # what SolKzFCFV.jl runs
            cases.Case("S500", "inclusion", 0, "Gradient", neumann=True),                     # bc = -1 interface, bc = 2 south
            cases.Case("S500", "inclusion", 1, "SymmetricGradient", neumann=True)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", C3_CASES, ids=lambda c: c.name)
def test_config3_size_bitwise(case, oracle):
    from fcfv_nme23_b200 import api, lib as L
    from helpers import assert_csr_equal, csc_triple_to_csr
    m, d = cases.build(case)
    g, _ = cases.build(case)
    nel, nf, neu = m.nel, m.nf, d["neu"]
    assert nel == 1_000_000 and nf == 1_501_000
    if case.setting == "inclusion":
        assert (m.bc == -1).sum() > 0 and (m.bc == 2).sum() == 500
    a, b, Z = oracle.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    assert np.array_equal(ga, a) and np.array_equal(gb, b) and np.array_equal(gZ, Z) and np.array_equal(g.τ, m.τ)
    ofn = oracle.ElementAssemblyLoop if case.variant == 0 else oracle.ElementAssemblyLoopNEW
    gfn = api.ElementAssemblyLoop if case.variant == 0 else api.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = ofn(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form)
    gKuu, gMuu, gKup, gfu, gfp, _ = gfn(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, case.form)
    h = api.handle_for(g)
    for blk, ref, shape, what in ((L.KUU, Kuu, (2 * nf, 2 * nf), "Kuu"), (L.KUP, Kup, (2 * nf, nel), "Kup"),
                                  (L.MUU, Muu, (2 * nf, 2 * nf), "Muu")):
        assert_csr_equal(h.export_csr(blk), csc_triple_to_csr(ref, shape), what)
        cp, rv, v = h.export_csc(blk)
        assert np.array_equal(cp, ref[0]) and np.array_equal(rv, ref[1]) and np.array_equal(v, ref[2]), what + " CSC"
    assert np.array_equal(gfu.ravel(), np.asarray(fu).ravel()) and np.array_equal(gfp, fp)
    for tau_mode in ("REFERENCE", "FACE"):
        n = oracle.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"],
                                                  d["VyDir"], *neu, case.form, tau_mode=tau_mode)
        gn = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], g, ga, gb, gZ, d["sex"], d["sey"], d["VxDir"],
                                                d["VyDir"], *neu, case.form, tau_mode=tau_mode, verbose=False)
        assert _norms_close(gn, n), (tau_mode, gn, n)
    h.close()
