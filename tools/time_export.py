import sys, time
sys.path.insert(0, '/root/repo')
from fcfv_nme23_b200 import api, lib as L
for n in (500, 1581):
    h = api.Handle(0)
    h.generate_structured(n, setting="solkz", tau_r=10.0)
    h.run_local("SymmetricGradient"); h.run_assemble(1, "SymmetricGradient"); h.sync()
    t0 = time.time(); cp, rv, v = h.export_csc(L.KUU); t1 = time.time()
    rp, ci, vv = h.export_csr(L.KUU); t2 = time.time()
    print(f"S({n}): nel {h.nel} nnz {len(v)}  CSC export {t1 - t0:.3f} s   CSR export {t2 - t1:.3f} s")
    h.close()
