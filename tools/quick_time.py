"""Quick device timing of the resident pipeline on S(n) (development helper)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcfv_nme23_b200 import api, lib as L

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    form = sys.argv[2] if len(sys.argv) > 2 else "SymmetricGradient"
    variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    h = api.Handle(0)
    t0 = time.time(); h.generate_structured(n, setting="solkz", tau_r=10.0); h.sync(); print("gen", time.time() - t0, "s nel", h.nel, "bytes", h.device_bytes / 1e9)
    st = torch.cuda.ExternalStream(h.stream)
    h.set_async(True)
    def step():
        h.run_local(form); h.run_assemble(variant, form); h.run_residuals(form, "FACE", fetch=False)
    for _ in range(3): step()
    h.sync()
    names = ["local", "assemble", "residual"]
    fns = [lambda: h.run_local(form), lambda: h.run_assemble(variant, form), lambda: h.run_residuals(form, "FACE", fetch=False)]
    for nm, fn in zip(names, fns):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(st)
        for _ in range(5): fn()
        ev[1].record(st); h.sync()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print(f"{nm:9s} {ms:8.3f} ms  {h.nel / ms / 1e3:9.1f} Mel/s")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(st)
    for _ in range(5): step()
    ev[1].record(st); h.sync()
    ms = ev[0].elapsed_time(ev[1]) / 5
    nnz = h.get_nnz()
    bytes_alg = h.nel * (214 + 302 + 156 + 24 + 8) + h.nf * (60 + 16) + 12 * (nnz[0] + nnz[1]) + 8 * (2 * h.nf + 1)
    print(f"step      {ms:8.3f} ms  {h.nel / ms / 1e3:9.1f} Mel/s  nnz {nnz}  alg GB/s {bytes_alg / ms / 1e6:.0f}  dev GB {h.device_bytes / 1e9:.2f}")
    # per-kernel device times (CUDA events around every hot-path launch)
    h.sync(); h.set_timing(True)
    for _ in range(5): step()
    h.sync()
    print("  ".join(f"{k}={ms / max(c, 1) * (c / 5):.3f}" for k, (ms, c) in h.get_timers().items() if c))
    h.close()

main()
