"""Small cases of every kernel family for compute-sanitizer (no torch import: only the library's own launches):

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py

LowRes (unstructured, interface faces) and H1 through ComputeFCFV, both assembly loops with the Muu third block, the
Schur third block, residuals with the optional arrays, element values, the Powell-Hestenes kernels, the CSC export,
the geometry kernel and the structured generator.  Results are compared with the oracle so that a sanitizer run is
also a parity run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import cases
from helpers import assert_csr_equal, csc_triple_to_csr
from oracle import oracle as O
from fcfv_nme23_b200 import api, lib as L


def one(case):
    m, d = cases.build(case)
    g, _ = cases.build(case)
    neu = d["neu"]
    nel, nf = m.nel, m.nf
    a, b, Z = O.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    ga, gb, gZ = api.ComputeFCFV(g, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *neu, case.tau_r, case.form)
    assert np.array_equal(ga, a) and np.array_equal(gb, b) and np.array_equal(gZ, Z) and np.array_equal(g.τ, m.τ)
    ofn = O.ElementAssemblyLoop if case.variant == 0 else O.ElementAssemblyLoopNEW
    gfn = api.ElementAssemblyLoop if case.variant == 0 else api.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = ofn(m, a, b, Z, d["VxDir"], d["VyDir"], *neu, case.form)
    gfn(g, ga, gb, gZ, d["VxDir"], d["VyDir"], *neu, case.form)                    # with the Muu third block
    h = api.handle_for(g)
    for blk, ref, shape, what in ((L.KUU, Kuu, (2 * nf, 2 * nf), "Kuu"), (L.KUP, Kup, (2 * nf, nel), "Kup"),
                                  (L.MUU, Muu, (2 * nf, 2 * nf), "Muu")):
        assert_csr_equal(h.export_csr(blk), csc_triple_to_csr(ref, shape), what)
        cp, rv, v = h.export_csc(blk)
        assert np.array_equal(cp, ref[0]) and np.array_equal(rv, ref[1]) and np.array_equal(v, ref[2])
    ref_sc = O.schur_complement(m, Kuu, Kup, 1.0e4)
    api.SchurComplement(g, 1.0e4)                                                  # Schur third block
    cp, rv, v = h.export_csc(L.KUUSC)
    assert np.array_equal(cp, ref_sc[0]) and np.array_equal(rv, ref_sc[1]) and np.array_equal(v, ref_sc[2])
    for tau_mode in ("REFERENCE", "FACE"):
        n = O.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"], d["VyDir"],
                                             *neu, case.form, tau_mode=tau_mode)
        gn, _ = api.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], g, ga, gb, gZ, d["sex"], d["sey"], d["VxDir"],
                                                   d["VyDir"], *neu, case.form, tau_mode=tau_mode, verbose=False, arrays=True)
        big = n > 1e-9
        assert np.all(np.abs(gn[big] - n[big]) <= 1e-12 * n[big])
    api.ComputeElementValues(g, d["Vxh"], d["Vyh"], d["Pe"], ga, gb, gZ, d["VxDir"], d["VyDir"], case.form)
    u = np.concatenate([d["Vxh"], d["Vyh"]])
    api.ph_residual(g, u, d["Pe"]); api.ph_update_p(g, 1e4, u, d["Pe"]); api.ph_fusc(g, 1e4, d["Pe"]); api.center_pressure(g, d["Pe"])
    h.close()
    print("ok", case.name, flush=True)


def main():
    for case in (cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True), cases.Case("LowRes", "inclusion", 1, "SymmetricGradient"),
                 cases.Case("H1", "solkz", 1, "SymmetricGradient", tau_r=10.0), cases.Case("H1", "solkz", 1, "Gradient", tau_r=10.0),
                 cases.Case("S6", "inclusion", 0, "SymmetricGradient", neumann=True)):
        one(case)
    m = cases.load_mesh("H1")
    api.mesh_geometry(m, 3.5)
    h = api.Handle(0)
    h.generate_structured(37, setting="solkz", tau_r=10.0)
    h.run_local("SymmetricGradient"); h.run_assemble(1, "SymmetricGradient"); h.run_residuals("SymmetricGradient", "FACE")
    for _ in range(3):
        h.run_pipeline(1, "SymmetricGradient", "FACE")                             # CUDA-graph replay path
    h.sync()
    h.close()
    print("sanitize cases done")


if __name__ == "__main__":
    main()
