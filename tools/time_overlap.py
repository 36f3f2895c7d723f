"""Would the residual kernels overlap usefully with the assembly passes?  Two handles (two streams, two copies of
S(n)) on one GPU: assembly on one, residuals on the other, back to back vs concurrently.  Development measurement."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcfv_nme23_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1581
form, variant = "SymmetricGradient", 1
A, B = api.Handle(0), api.Handle(0)
for h in (A, B):
    h.generate_structured(n, setting="solkz", tau_r=10.0)
    h.set_async(True)
    h.run_local(form); h.run_assemble(variant, form); h.run_residuals(form, "FACE", fetch=False)
    h.sync()
reps = 10

def wall(fn):
    A.sync(); B.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    A.sync(); B.sync()
    return (time.perf_counter() - t0) / reps * 1e3

def seq():
    for _ in range(reps):
        A.run_assemble(variant, form); A.run_residuals(form, "FACE", fetch=False)
def conc():
    for _ in range(reps):
        A.run_assemble(variant, form); B.run_residuals(form, "FACE", fetch=False)
def only_asm():
    for _ in range(reps): A.run_assemble(variant, form)
def only_res():
    for _ in range(reps): B.run_residuals(form, "FACE", fetch=False)

for tag, fn in (("assembly alone", only_asm), ("residuals alone", only_res), ("one stream, back to back", seq), ("two streams", conc)):
    wall(fn)
    print(f"{tag:28s} {wall(fn):8.3f} ms per pass")
