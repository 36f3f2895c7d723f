#!/bin/bash
# GPU suite + a 16 M-element pipeline pass under the bounds-checked build of the library (compute-sanitizer is closed on the pool)
set -u
OUT=gpurun_out; mkdir -p $OUT
LIBB=$PWD/fcfv_nme23_b200/variants/libfcfv_b200_bounds.so
FCFV_LIB=$LIBB timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_cabi.py::test_plain_c_client_on_gpu > $OUT/r2_bounds_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2_bounds_pytest.log
tail -6 $OUT/r2_bounds_pytest.log
FCFV_LIB=$LIBB timeout 600 python - <<'PY' 2>&1 | tee $OUT/r2_bounds_large.log
import sys; sys.path.insert(0, ".")
from fcfv_nme23_b200 import api, lib as L
for n, setting, form, variant, muu in ((2000, "solkz", "SymmetricGradient", 1, False), (1000, "inclusion", "Gradient", 0, True)):
    h = api.Handle(0)
    h.generate_structured(n, setting=setting, tau_r=10.0, neumann_south=(setting == "inclusion"))
    h.run_local(form); h.run_assemble(variant, form, want_muu=muu); ss, norms = h.run_residuals(form, "FACE")
    h.run_pipeline(variant, form, "FACE")
    print(f"S({n}) {setting} {form} loop {variant}: nel {h.nel} nnz {h.get_nnz()} violations so far {L.load().fcfv_debug_bounds_violations()}")
    h.close()
PY
