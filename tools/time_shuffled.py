"""How much the path depends on numbering locality: S(n) as generated vs the same mesh with element and face ids
shuffled at random (worst case for the gathers).  Development measurement, results quoted in DESIGN.md."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fcfv_nme23_b200 import api
from fcfv_nme23_b200.mesh import FCFV_Mesh

form, variant = "SymmetricGradient", 1
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
g = api.Handle(0)
g.generate_structured(n, setting="solkz", tau_r=10.0)
M, D = g.get_mesh(), g.get_data()
nel, nf = g.nel, g.nf
g.close()

def run(tag, pe, pf):
    """pe: new element order (new e -> old e), pf: old face id -> new face id (0-based)."""
    m = FCFV_Mesh(nel=nel, nf=nf, nf_el=3)
    inv_f = np.empty(nf, dtype=np.int64); inv_f[pf] = np.arange(nf)          # new -> old
    m.e2f = np.asfortranarray(pf[M["e2f"][pe] - 1] + 1)
    m.bc = M["bc"][inv_f]
    for k in ("Ω", "ke", "δ", "τe"): setattr(m, k, M[k][pe].copy())
    for k in ("n_x", "n_y", "Γ"): setattr(m, k, np.asfortranarray(M[k][pe]))
    h = api.Handle(0)
    h.set_mesh(m)
    h.set_sources(D["sex"][pe], D["sey"][pe])
    h.set_face_data(*[D[k][inv_f] for k in ("VxDir", "VyDir", "sxx", "syy", "sxy", "syx")])
    h.set_solution(D["Vxh"][inv_f], D["Vyh"][inv_f], D["Pe"][pe])
    st = torch.cuda.ExternalStream(h.stream)
    h.set_async(True)
    def step():
        h.run_local(form); h.run_assemble(variant, form); h.run_residuals(form, "FACE", fetch=False)
    for _ in range(3): step()
    h.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): step()
    e1.record(st); h.sync()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag:28s} {ms:8.3f} ms/pass  {nel / ms / 1e3:8.1f} Mel/s  nnz {h.get_nnz()[:2]}")
    h.close()

rng = np.random.default_rng(0)
ident_e, ident_f = np.arange(nel), np.arange(nf)
run("as generated", ident_e, ident_f)
run("elements shuffled", rng.permutation(nel), ident_f)
run("faces shuffled", ident_e, rng.permutation(nf))
run("both shuffled", rng.permutation(nel), rng.permutation(nf))
# the remedy for arbitrary FACE numbering: fcfv_first_seen_faces (mesh preparation)
pf = rng.permutation(nf)
e2f_sh = np.asfortranarray(pf[M["e2f"] - 1] + 1)
new_of_old, _ = api.first_seen_faces(e2f_sh, nf)
run("faces shuffled + first-seen", ident_e, (new_of_old - 1)[pf])

# the remedy for an arbitrary ELEMENT order: fcfv_hilbert_elements, then first-seen faces.  Centroids of S(n):
# cell (i, j) of element e = e // 4, triangle k = e % 4 of the criss-cross cell.
from fcfv_nme23_b200 import mesh as MM
cell, k = np.arange(nel) // 4, np.arange(nel) % 4
dx = np.array([0.5, 5.0 / 6.0, 0.5, 1.0 / 6.0])[k]; dy = np.array([1.0 / 6.0, 0.5, 5.0 / 6.0, 0.5])[k]
xc, yc = (cell % n + dx) / n, (cell // n + dy) / n
order = api.hilbert_elements(xc, yc) - 1
new_of_old, _ = api.first_seen_faces(np.asfortranarray(M["e2f"][order]), nf)
run("hilbert elements + first-seen", order, new_of_old - 1)
pe = rng.permutation(nel)
order = pe[api.hilbert_elements(xc[pe], yc[pe]) - 1]
pf = rng.permutation(nf)
new_of_old, _ = api.first_seen_faces(np.asfortranarray(pf[M["e2f"][order] - 1] + 1), nf)
run("both shuffled + hilbert + f-s", order, (new_of_old - 1)[pf])
