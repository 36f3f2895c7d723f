"""The C-ABI library: builds, loads, exports every symbol include/fcfv_b200.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from fcfv_nme23_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "run __graft_entry__.build() first"
    so = C.CDLL(L.LIB_PATH)
    declared = L.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(so, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (fcfv_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert so.fcfv_version() == 100


def test_binding_covers_header():
    lib = L.load()
    for s in L.declared_symbols():
        assert getattr(lib, s).restype is not None or s in ("fcfv_stream",), s


def test_shipped_library_is_not_the_bounds_checked_build():
    """fcfv_debug_bounds_violations: -1 = the index checks are compiled out (the build the benchmarks run); the checked
    variant (make variant NAME=bounds DEFS=-DFCFV_BOUNDS_CHECK=1) returns a count >= 0 and is run by tools/gpu_bounds.sh."""
    assert L.load().fcfv_debug_bounds_violations() == -1


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    lib = L.load()
    h = C.c_void_p()
    assert lib.fcfv_create(C.byref(h), 0) != 0 and not h.value     # FCFV_ERR_CUDA, no handle
    from fcfv_nme23_b200 import api
    with pytest.raises(L.FCFVError):
        api.Handle(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fcfv_nme23_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fcfv_nme23_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "fcfv_oracle" not in text, fn


def test_plain_c_client_links_and_runs():
    """A C99 translation unit including the header links against the library with gcc (C linkage, no C++ / torch types
    in the signatures); with a GPU it also runs the one-triangle case through the host-buffer entry points."""
    src = os.path.join(ROOT, "tests", "c_abi_client.c")
    exe = os.path.join(ROOT, "tests", "c_abi_client")
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", exe, src, "-L" + libdir, "-lfcfv_b200",
                           "-Wl,-rpath," + libdir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode in (0, 77), r.stdout + r.stderr
    try:
        import torch
        if torch.cuda.is_available():
            assert r.returncode == 0 and "C client OK" in r.stdout, r.stdout
    except ImportError:
        pass


@pytest.mark.gpu
def test_plain_c_client_on_gpu():
    """The C client's compute sequence (set_mesh -> compute_local -> assemble -> CSC export) on the device."""
    test_plain_c_client_links_and_runs()
    r = subprocess.run([os.path.join(ROOT, "tests", "c_abi_client")], capture_output=True, text=True)
    assert r.returncode == 0 and "C client OK" in r.stdout, r.stdout + r.stderr
