"""Host-buffer pass (fcfv_set_mesh + fcfv_compute_local + fcfv_assemble + fcfv_residuals) from PAGEABLE NumPy arrays
(what a Julia / NumPy caller has) vs the same arrays after fcfv_host_register.  Development measurement."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from fcfv_nme23_b200 import api, lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
h = api.Handle(0)
h.generate_structured(n, setting="solkz", tau_r=10.0)
nel, nf = h.nel, h.nf
M = {k: np.ascontiguousarray(v.reshape(-1, order="F")) for k, v in h.get_mesh().items()}
D = {k: np.ascontiguousarray(v) for k, v in h.get_data().items()}
out = dict(alpha=np.empty(nel), beta=np.empty(2 * nel), Z=np.empty(4 * nel), tau=np.empty(nf))
p = lambda a: C.c_void_p(a.ctypes.data)
lib, form = h.lib, api.formulation_id("SymmetricGradient")
norms = np.zeros(6)
nnz = [C.c_int64(0) for _ in range(3)]
sec = C.c_double(0.0)

def step():
    h.check(lib.fcfv_set_mesh(h._h, nel, nf, p(M["e2f"]), p(M["bc"]), p(M["Ω"]), p(M["n_x"]), p(M["n_y"]), p(M["Γ"]), p(M["ke"]),
                              p(M["δ"]), p(M["τe"]), nel))
    h.check(lib.fcfv_compute_local(h._h, form, 1.0, p(D["sex"]), p(D["sey"]), p(D["VxDir"]), p(D["VyDir"]), p(out["alpha"]),
                                   p(out["beta"]), p(out["Z"]), p(out["tau"])))
    h.check(lib.fcfv_assemble(h._h, 1, form, 1.0, None, None, None, p(D["VxDir"]), p(D["VyDir"]), p(D["sxx"]), p(D["syy"]),
                              p(D["sxy"]), p(D["syx"]), 0, C.byref(nnz[0]), C.byref(nnz[1]), C.byref(nnz[2]), C.byref(sec)))
    h.check(lib.fcfv_residuals(h._h, form, L.TAU_FACE, p(D["Vxh"]), p(D["Vyh"]), p(D["Pe"]), None, None, None, p(D["sex"]),
                               p(D["sey"]), p(D["VxDir"]), p(D["VyDir"]), p(D["sxx"]), p(D["syy"]), p(D["sxy"]), p(D["syx"]),
                               p(norms), None, None, None))

def timed(tag):
    step(); h.sync()
    t0 = time.perf_counter()
    for _ in range(3): step()
    h.sync()
    dt = (time.perf_counter() - t0) / 3
    gb = (168 * nel + 136 * nf + 8 * (7 * nel + nf)) / 1e9
    print(f"{tag:28s} {1e3 * dt:9.2f} ms/pass  {nel / dt / 1e6:8.1f} Mel/s  {gb / dt:6.1f} GB/s over PCIe  norms[1]={norms[1]:.6e}")

h.set_async(False)
timed("pageable arrays")
arrs = list(M.values()) + list(D.values()) + list(out.values())
t0 = time.perf_counter()
h.host_register(*arrs)
print(f"fcfv_host_register of {sum(a.nbytes for a in arrs) / 1e9:.2f} GB: {time.perf_counter() - t0:.2f} s (once)")
timed("registered arrays")
h.host_unregister(*arrs)
timed("pageable again")
h.close()
