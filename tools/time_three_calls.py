"""Wall time of each of the three host-buffer calls of a driver pass (page-locked arrays, S(n)): where an e2e pass spends its time.

    python tools/time_three_calls.py [n]
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fcfv_nme23_b200 import api, lib as L


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3536
    h = api.Handle(0)
    h.generate_structured(n, setting="solkz", tau_r=10.0)
    P, D = h.get_mesh(), h.get_data()
    nel, nf = h.nel, h.nf
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).pin_memory()
    P = {k: pin(v) for k, v in P.items()}
    D = {k: pin(v) for k, v in D.items()}
    out = {"alpha": torch.empty(nel, dtype=torch.float64).pin_memory(), "beta": torch.empty(2 * nel, dtype=torch.float64).pin_memory(),
           "Z": torch.empty(4 * nel, dtype=torch.float64).pin_memory(), "tau": torch.zeros(nf, dtype=torch.float64).pin_memory()}
    ptr = lambda t: C.c_void_p(t.data_ptr())
    lib = h.lib
    form = api.formulation_id("SymmetricGradient")
    h.set_async(False); h.upload_cache(True)
    sig = [ptr(D[k]) for k in ("sxx", "syy", "sxy", "syx")]
    loc = [ptr(out[k]) for k in ("alpha", "beta", "Z")]
    nnz = [C.c_int64(0) for _ in range(3)]
    norms = np.zeros(6)
    def refresh():
        h.check(lib.fcfv_update_fields(h._h, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, ptr(out["tau"])))
    for with_mesh in (True, True, False, False):
        t = [time.perf_counter()]
        geo = [ptr(P[k]) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ")] if with_mesh else [None] * 6
        if not with_mesh:
            h.check(lib.fcfv_new_epoch(h._h))
        h.check(lib.fcfv_compute_fcfv(h._h, nel, nf, *geo, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, ptr(out["tau"]), form, 1.0,
                                      ptr(D["sex"]), ptr(D["sey"]), ptr(D["VxDir"]), ptr(D["VyDir"]), *sig, *loc, ptr(out["tau"])))
        t.append(time.perf_counter())
        refresh()
        h.check(lib.fcfv_assemble(h._h, 1, form, 1.0, *loc, ptr(D["VxDir"]), ptr(D["VyDir"]), *sig, 0, C.byref(nnz[0]), C.byref(nnz[1]),
                                  C.byref(nnz[2]), None))
        t.append(time.perf_counter())
        refresh()
        h.check(lib.fcfv_residuals(h._h, form, L.TAU_FACE, ptr(D["Vxh"]), ptr(D["Vyh"]), ptr(D["Pe"]), *loc, ptr(D["sex"]), ptr(D["sey"]),
                                   ptr(D["VxDir"]), ptr(D["VyDir"]), *sig, C.c_void_p(norms.ctypes.data), None, None, None))
        t.append(time.perf_counter())
        d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
        gb = lambda bytes_per_el: bytes_per_el * nel / 1e9
        print(f"{'with mesh' if with_mesh else 'resident '}: ComputeFCFV {d[0]:7.1f} ms | assembly {d[1]:6.1f} | residuals {d[2]:6.1f} | total {sum(d):7.1f} ms = "
              f"{nel / sum(d) / 1e3:6.1f} Melements/s   (up {gb(156 if with_mesh else 40):.1f} GB + {gb(32):.1f} GB, down {gb(68):.1f} GB)")
    h.close()


main()
