"""Host-side cost of the single-process multi-GPU handle (fcfv_create_multi): seconds per call of a driver pass on S(n).

    python tools/time_multi_handle.py [n] [ngpus]
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fcfv_nme23_b200 import api, lib as L


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    ngpus = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    g = api.Handle(0)
    g.generate_structured(n, 0, n, setting="solkz", tau_r=10.0)
    P, D = g.get_mesh(), g.get_data()
    nel, nf = g.nel, g.nf
    g.close()
    P = {k: np.ascontiguousarray(v.reshape(-1, order="F")) for k, v in P.items()}
    out = {"alpha": np.empty(nel), "beta": np.empty(2 * nel), "Z": np.empty(4 * nel), "tau": np.zeros(nf)}
    ptr = lambda a: C.c_void_p(a.ctypes.data)
    form = api.formulation_id("SymmetricGradient")
    for ng in (1, ngpus):
        h = api.Handle(ngpus=ng)
        h.upload_cache(True)
        lib = h.lib
        nnz = [C.c_int64(0) for _ in range(3)]
        norms = np.zeros(6)
        sig = [ptr(D[k]) for k in ("sxx", "syy", "sxy", "syx")]
        loc = [ptr(out[k]) for k in ("alpha", "beta", "Z")]
        def refresh():
            h.check(lib.fcfv_update_fields(h._h, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, ptr(out["tau"])))
        for rep in range(2):
            t = [time.perf_counter()]
            h.check(lib.fcfv_set_mesh(h._h, nel, nf, ptr(P["e2f"]), ptr(P["bc"]), ptr(P["Ω"]), ptr(P["n_x"]), ptr(P["n_y"]), ptr(P["Γ"]),
                                      ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel)); t.append(time.perf_counter())
            refresh()
            h.check(lib.fcfv_compute_local_ex(h._h, form, 1.0, ptr(D["sex"]), ptr(D["sey"]), ptr(D["VxDir"]), ptr(D["VyDir"]), *sig, *loc,
                                              ptr(out["tau"]))); t.append(time.perf_counter())
            refresh()
            h.check(lib.fcfv_assemble(h._h, 1, form, 1.0, *loc, ptr(D["VxDir"]), ptr(D["VyDir"]), *sig, 0, C.byref(nnz[0]), C.byref(nnz[1]),
                                      C.byref(nnz[2]), None)); t.append(time.perf_counter())
            refresh()
            h.check(lib.fcfv_residuals(h._h, form, L.TAU_FACE, ptr(D["Vxh"]), ptr(D["Vyh"]), ptr(D["Pe"]), *loc, ptr(D["sex"]), ptr(D["sey"]),
                                       ptr(D["VxDir"]), ptr(D["VyDir"]), *sig, ptr(norms), None, None, None)); t.append(time.perf_counter())
            d = np.diff(t)
            psec = h.partition_info()[1] if ng > 1 else 0.0
            print(f"ngpus={ng} S({n}) {nel} elements pass {rep}: set_mesh {d[0]:.3f} s (partition {psec:.3f}) | ComputeFCFV {d[1]:.3f} | "
                  f"assembly {d[2]:.3f} | residuals {d[3]:.3f} | total {t[-1] - t[0]:.3f} s = {nel / (t[-1] - t[0]) / 1e6:.1f} Melements/s  "
                  f"norms[3]={norms[3]:.6e}")
        h.close()


main()
