"""Golden vectors produced by the REFERENCE'S OWN SOURCE (run once in the build container):

    python tests/golden/make_reference_golden.py

Julia is not installed here, so the reference's functions are executed through the rule-table transliteration of
their source text (tests/julia_transliteration.py reads /root/reference/src/*.jl; nothing is retyped).  For every test
case this writes SHA-256 digests of ComputeFCFV's alpha / beta / Z / mesh.tau, of the CSC arrays of Kuu / Muu / Kup
returned by ElementAssemblyLoop[NEW] (-> Sparsify), of fu / fp and of ComputeElementValues' five arrays, the block
sizes and the six printed residual norms into golden_reference.json -- same record layout and digest function as
golden_oracle.json, so "oracle == reference" is `record_a[key] == record_b[key]`.  The file travels to boxes without
the reference checkout (tests/test_golden_reference.py: oracle on CPU, CUDA path with -m gpu)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    import cases
    import translit_run
    out = {}
    for case in cases.all_cases(mesh_dir=HERE) + cases.large_cases():
        t0 = time.time()
        m, d = cases.build(case, mesh_dir=HERE)
        r = translit_run.run_case(case, m, d)
        out[case.name] = dict(
            nel=int(m.nel), nf=int(m.nf), nnz_uu=int(len(r["Kuu"][2])), nnz_muu=int(len(r["Muu"][2])), nnz_up=int(len(r["Kup"][2])),
            local=digest(r["a"], r["b"].ravel(order="F"), r["Z"].ravel(order="F"), r["tau"]),
            Kuu=digest(*r["Kuu"]), Muu=digest(*r["Muu"]), Kup=digest(*r["Kup"]), rhs=digest(r["fu"], r["fp"]),
            elem_values=digest(*r["ev"]), norms_reference=[float(x) for x in r["norms"]])
        print(f"{case.name}: nnz_uu={out[case.name]['nnz_uu']} ({time.time() - t0:.1f}s)", flush=True)
    meta = {"_generator": "tests/golden/make_reference_golden.py: transliterated /root/reference/src/*.jl, executed with float64 scalars"}
    with open(os.path.join(HERE, "golden_reference.json"), "w") as fh:
        json.dump({**meta, **out}, fh, indent=1, sort_keys=True)
    print(f"wrote {len(out)} records")


if __name__ == "__main__":
    main()
