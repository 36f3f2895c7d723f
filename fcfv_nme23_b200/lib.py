"""ctypes binding of ``libfcfv_b200.so`` -- exactly the symbols of ``include/fcfv_b200.h``.

This is the stand-in for the Julia ``ccall`` shim (``julia/FCFV_B200.jl``) in an environment
without Julia.  There is no fallback: if the shared library is missing or CUDA is unavailable,
loading / ``fcfv_create`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FCFV_LIB") or os.path.join(_HERE, "libfcfv_b200.so")   # FCFV_LIB: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fcfv_b200.h")

FCFV_OK = 0
GRADIENT, SYMMETRIC_GRADIENT = 0, 1
LOOP_OLD, LOOP_NEW = 0, 1
TAU_REFERENCE, TAU_FACE = 0, 1
KUU, KUP, MUU, KUUSC = 0, 1, 2, 3
CSR0, CSC1_INT64 = 0, 1

_lib = None


class FCFVError(RuntimeError):
    pass


def build(force=False):
    """Compiles the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))] + [HEADER_PATH]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc")])
    return LIB_PATH


def declared_symbols():
    """Names of all functions declared in include/fcfv_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcfv_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and not os.environ.get("FCFV_LIB"):
        try:                                   # same image on the GPU boxes: nvcc is there, build in-tree once
            build()
        except Exception as exc:               # noqa: BLE001 -- reported below
            raise FCFVError(f"{LIB_PATH} is missing and could not be built ({exc}); there is no CPU fallback") from exc
    if not os.path.exists(LIB_PATH):
        raise FCFVError(f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i64, dbl, ci = C.c_void_p, C.c_int64, C.c_double, C.c_int
    P = C.POINTER
    sig = {
        "fcfv_create": [P(vp), ci],
        "fcfv_create_multi": [P(vp), ci],
        "fcfv_num_devices": [vp],
        "fcfv_partition_info": [vp, ci, P(i64), P(i64), P(i64), P(i64), P(dbl)],
        "fcfv_partition_part": [i64, i64, vp, ci, ci, P(i64), P(i64), vp, vp, vp, vp],
        "fcfv_destroy": [vp],
        "fcfv_version": [],
        "fcfv_set_mesh": [vp, i64, i64] + [vp] * 9 + [i64],
        "fcfv_update_fields": [vp, vp, vp, vp, i64, vp],
        "fcfv_set_ownership": [vp, vp, vp, i64, i64, vp],
        "fcfv_geometry": [vp, i64, i64, i64] + [vp] * 6 + [ci, dbl] + [vp] * 7,
        "fcfv_compute_local": [vp, ci, dbl] + [vp] * 8,
        "fcfv_assemble": [vp, ci, ci, dbl] + [vp] * 9 + [ci, P(i64), P(i64), P(i64), P(dbl)],
        "fcfv_index_width": [vp],
        "fcfv_block_size": [vp, ci, P(i64), P(i64), P(i64)],
        "fcfv_export": [vp, ci, ci, vp, vp, vp],
        "fcfv_export_rhs": [vp, vp, vp],
        "fcfv_device_csr": [vp, ci, P(vp), P(vp), P(vp)],
        "fcfv_device_rhs": [vp, P(vp), P(vp)],
        "fcfv_residuals": [vp, ci, ci] + [vp] * 17,
        "fcfv_ph_residual": [vp] + [vp] * 5,
        "fcfv_ph_update_p": [vp, dbl, vp, vp],
        "fcfv_ph_fusc": [vp, dbl, vp, vp],
        "fcfv_element_values": [vp] + [vp] * 10,
        "fcfv_schur": [vp, dbl, P(i64)],
        "fcfv_first_seen_faces": [vp, i64, i64, vp, vp, vp],
        "fcfv_hilbert_elements": [vp, i64, vp, vp, vp],
        "fcfv_pass": [vp, ci, ci, dbl, ci] + [vp] * 11 + [vp] * 4 + [ci, P(i64), P(i64), P(i64), vp],
        "fcfv_numbering_locality": [vp, P(dbl), P(dbl)],
        "fcfv_compute_local_ex": [vp, ci, dbl] + [vp] * 12,
        "fcfv_compute_fcfv": [vp, i64, i64] + [vp] * 9 + [i64, vp, ci, dbl] + [vp] * 12,
        "fcfv_upload_cache": [vp, ci],
        "fcfv_has_orphan_faces": [vp, P(ci)],
        "fcfv_new_epoch": [vp],
        "fcfv_transfer_stats": [vp, P(i64), P(i64), P(i64), ci],
        "fcfv_sparse_face_data": [vp, ci],
        "fcfv_sparse_face_info": [vp, P(ci), P(i64), P(i64), P(i64)],
        "fcfv_host_register": [vp, vp, i64],
        "fcfv_host_unregister": [vp, vp],
        "fcfv_run_pipeline": [vp, ci, ci, dbl, ci],
        "fcfv_center_pressure": [vp, vp, P(dbl)],
        "fcfv_run_assemble_schur": [vp, ci, ci, dbl, dbl],
        "fcfv_set_sources": [vp, vp, vp],
        "fcfv_set_face_data": [vp] + [vp] * 6,
        "fcfv_set_solution": [vp, vp, vp, vp],
        "fcfv_set_local": [vp, vp, vp, vp],
        "fcfv_run_local": [vp, ci, dbl],
        "fcfv_run_assemble": [vp, ci, ci, dbl, ci],
        "fcfv_run_residuals": [vp, ci, ci, vp, vp],
        "fcfv_get_local": [vp, vp, vp, vp, vp],
        "fcfv_get_nnz": [vp, P(i64), P(i64), P(i64)],
        "fcfv_set_async": [vp, ci],
        "fcfv_sync": [vp],
        "fcfv_set_timing": [vp, ci],
        "fcfv_num_timers": [],
        "fcfv_get_timers": [vp, vp, vp],
        "fcfv_generate_structured": [vp, i64, i64, i64, ci, ci, dbl, dbl, ci],
        "fcfv_mesh_size": [vp, P(i64), P(i64)],
        "fcfv_get_mesh": [vp] + [vp] * 10,
        "fcfv_get_data": [vp] + [vp] * 11,
        "fcfv_get_ownership": [vp, vp, vp, P(i64), P(i64)],
        "fcfv_comm_unique_id": [vp, vp],
        "fcfv_comm_init": [vp, vp, ci, ci],
        "fcfv_comm_allreduce_sum": [vp, vp, ci],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ci
    lib.fcfv_last_error.argtypes = [vp]
    lib.fcfv_last_error.restype = C.c_char_p
    lib.fcfv_timer_name.argtypes = [ci]
    lib.fcfv_timer_name.restype = C.c_char_p
    lib.fcfv_stream.argtypes = [vp]
    lib.fcfv_stream.restype = vp
    lib.fcfv_kernel_launches.argtypes = [vp]
    lib.fcfv_kernel_launches.restype = i64
    lib.fcfv_staged_copies.argtypes = [vp]
    lib.fcfv_staged_copies.restype = i64
    lib.fcfv_debug_bounds_violations.argtypes = []
    lib.fcfv_debug_bounds_violations.restype = i64
    lib.fcfv_pipelined_calls.argtypes = [vp]
    lib.fcfv_pipelined_calls.restype = i64
    lib.fcfv_device_bytes.argtypes = [vp]
    lib.fcfv_device_bytes.restype = i64
    _lib = lib
    return lib
