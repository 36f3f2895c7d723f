"""Runs one test case through the transliterated reference (tests/julia_transliteration.py) and
returns its outputs as NumPy arrays in the oracle's conventions -- TEST INFRASTRUCTURE."""
from __future__ import annotations

import numpy as np

import julia_transliteration as T

_REF = None


def reference():
    global _REF
    if _REF is None:
        _REF = T.Reference()
    return _REF


def run_case(case, m, d, residual=True, elem_values=True):
    """m, d = cases.build(case).  Returns a dict: a, b, Z, tau, Kuu/Muu/Kup (colptr, rowval, nzval),
    fu, fp, norms (the six printed values, reference tau mode), ev (ComputeElementValues)."""
    R = reference()
    jm = T.jmesh(m)
    J = T.J
    neu = [J(x) for x in d["neu"]]
    sex, sey, VxDir, VyDir = J(d["sex"]), J(d["sey"]), J(d["VxDir"]), J(d["VyDir"])
    a, b, Z = R.ComputeFCFV(jm, sex, sey, VxDir, VyDir, *neu, case.tau_r, case.form)
    fn = R.ElementAssemblyLoop if case.variant == 0 else R.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _ = fn(jm, a, b, Z, VxDir, VyDir, *neu, case.form)
    out = dict(a=a.a, b=b.a, Z=Z.a, tau=jm.τ.a, Kuu=Kuu.triple(), Muu=Muu.triple(), Kup=Kup.triple(),
               fu=fu.a.reshape(-1, order="F"), fp=fp.a)
    if residual:
        R.jl.printed.clear()
        ret = R.ComputeResidualsFCFV_Stokes_o1(J(d["Vxh"]), J(d["Vyh"]), J(d["Pe"]), jm, a, b, Z, sex, sey, VxDir, VyDir,
                                               *neu, case.form)
        assert ret is None and len(R.jl.printed) == 6          # returns nothing, prints six lines (:101-107)
        out["norms"] = np.array([v[0] for _, v in R.jl.printed])
        out["formats"] = [f for f, _ in R.jl.printed]
    if elem_values:
        ev = R.ComputeElementValues(jm, J(d["Vxh"]), J(d["Vyh"]), J(d["Pe"]), a, b, Z, VxDir, VyDir, case.form)
        out["ev"] = tuple(x.a for x in ev)
    return out
