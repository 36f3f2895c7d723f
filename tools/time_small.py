"""Small meshes (BASELINE configs 1-2 sizes): one pass as three direct calls vs the replayed CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcfv_nme23_b200 import api
form, variant = "SymmetricGradient", 1
for n in (16, 23, 50, 128, 256, 500):
    h = api.Handle(0)
    h.generate_structured(n, setting="solkz", tau_r=10.0)
    st = torch.cuda.ExternalStream(h.stream)
    h.set_async(True)
    def direct():
        h.run_local(form); h.run_assemble(variant, form); h.run_residuals(form, "FACE", fetch=False)
    def graph():
        h.run_pipeline(variant, form, "FACE")
    out = []
    for fn in (direct, graph):
        for _ in range(5): fn()
        h.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(200): fn()
        e1.record(st); h.sync()
        out.append(e0.elapsed_time(e1) / 200 * 1e3)
    print(f"S({n:3d}) nel {h.nel:8d}: direct {out[0]:8.1f} us/pass  graph {out[1]:8.1f} us/pass  x{out[0] / out[1]:.2f}")
    h.close()
