"""Assembly with the optional third block (Muu, or the Powell-Hestenes Schur complement) against the plain one, S(n)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcfv_nme23_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1581
form, variant = "SymmetricGradient", 1
h = api.Handle(0)
h.generate_structured(n, setting="solkz", tau_r=10.0)
st = torch.cuda.ExternalStream(h.stream)
h.set_async(True)
h.run_local(form)

def timed(tag, fn):
    for _ in range(2): fn()
    h.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5): fn()
    e1.record(st); h.sync()
    ms = e0.elapsed_time(e1) / 5
    print(f"{tag:34s} {ms:8.3f} ms  {h.nel / ms / 1e3:8.1f} Mel/s  nnz {h.get_nnz()}")

timed("Kuu + Kup", lambda: h.run_assemble(variant, form))
timed("Kuu + Kup + Muu", lambda: h.run_assemble(variant, form, want_muu=True))
timed("Kuu + Kup + Kuusc (fused Schur)", lambda: h.check(h.lib.fcfv_run_assemble_schur(h._h, variant, api.formulation_id(form), 1.0, 1.0e4)))
h.close()
