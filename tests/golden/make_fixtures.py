"""Generates the committed test fixtures.  Run once in the build container (needs /root/reference):

    python tests/golden/make_fixtures.py

1. Mesh inputs converted from the reference's shipped meshes (data, not source):
   MeshAdvancingFront{Low,Med}Res.mat -> .npz with the same keys; square_TRI_H{1,2}.dat copied,
   square_TRI_H{3,4}.dat gzip-compressed.
2. golden_oracle.json: SHA-256 digests + sizes of the ORACLE's outputs on those meshes.
   The reference has no golden vectors and Julia is not available here (parity unpinned), so
   these digests pin the oracle itself (regression) -- the GPU path is compared with the oracle
   bit for bit, hence with these digests.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("FCFV_REFERENCE", "/root/reference")


def convert_meshes():
    from scipy.io import loadmat
    for name in ("MeshAdvancingFrontLowRes", "MeshAdvancingFrontMedRes"):
        d = {k: v for k, v in loadmat(os.path.join(REF, "meshes", name + ".mat")).items() if not k.startswith("__")}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    for name in ("square_TRI_H1", "square_TRI_H2"):
        shutil.copyfile(os.path.join(REF, "meshes", name + ".dat"), os.path.join(HERE, name + ".dat"))
    import gzip
    for name in ("square_TRI_H3", "square_TRI_H4"):      # 0.7 / 2.8 MB of text: stored gzip-compressed (byte-exact)
        with open(os.path.join(REF, "meshes", name + ".dat"), "rb") as src, \
                gzip.GzipFile(os.path.join(HERE, name + ".dat.gz"), "wb", compresslevel=9, mtime=0) as dst:
            shutil.copyfileobj(src, dst)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def golden():
    import cases
    from oracle import oracle as O
    out = {}
    for case in cases.all_cases(mesh_dir=HERE):
        m, d = cases.build(case, mesh_dir=HERE)
        a, b, Z = O.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], case.tau_r, case.form)
        fn = O.ElementAssemblyLoop if case.variant == 0 else O.ElementAssemblyLoopNEW
        Kuu, Muu, Kup, fu, fp, _, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *d["neu"], case.form)
        nR = O.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"],
                                              d["VyDir"], *d["neu"], case.form, tau_mode="REFERENCE")
        nF = O.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"],
                                              d["VyDir"], *d["neu"], case.form, tau_mode="FACE")
        Ksc = O.schur_complement(m, Kuu, Kup, 1.0e4)
        ev = O.ComputeElementValues(m, d["Vxh"], d["Vyh"], d["Pe"], a, b, Z, d["VxDir"], d["VyDir"], case.form)
        out[case.name] = dict(
            nel=int(m.nel), nf=int(m.nf), nnz_uu=int(len(Kuu[2])), nnz_muu=int(len(Muu[2])), nnz_up=int(len(Kup[2])),
            local=digest(a, b.ravel(order="F"), Z.ravel(order="F"), m.τ),
            Kuu=digest(*Kuu), Muu=digest(*Muu), Kup=digest(*Kup), rhs=digest(fu, fp),
            nnz_sc=int(len(Ksc[2])), Kuusc=digest(*Ksc), elem_values=digest(*ev),
            norms_reference=[float(x) for x in nR], norms_face=[float(x) for x in nF])
    with open(os.path.join(HERE, "golden_oracle.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(f"wrote {len(out)} golden records")


if __name__ == "__main__":
    if os.path.isdir(REF):
        convert_meshes()          # mesh inputs come from the reference checkout; the digests only need the oracle
    golden()
