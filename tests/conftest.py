import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.join(ROOT, "tests")
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def emu():
    """Host build of the CUDA library's arithmetic (tests/host_emulation.cpp)."""
    import ctypes
    so = os.path.join(TESTS, "libhost_emulation.so")
    srcs = [os.path.join(TESTS, "host_emulation.cpp")] + [
        os.path.join(ROOT, "fcfv_nme23_b200", "csrc", f) for f in ("fcfv_math.cuh", "fcfv_kernels.cuh", "fcfv_generate.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        tmp = f"{so}.{os.getpid()}.tmp"          # several xdist workers may rebuild at once: each writes its own file,
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-o", tmp, srcs[0]])
        os.replace(tmp, so)                      # and the rename is atomic
    return ctypes.CDLL(so)


def pytest_sessionfinish(session, exitstatus):
    """With a bounds-checked build of the library (FCFV_LIB=.../libfcfv_b200_bounds.so, see include/fcfv_b200.h) the GPU suite
    ends with the number of out-of-range indices its kernels saw: anything but 0 fails the run."""
    if not _has_gpu():
        return
    try:
        from fcfv_nme23_b200 import lib as L
        if L._lib is None:
            return
        n = int(L._lib.fcfv_debug_bounds_violations())
    except Exception:
        return
    if n >= 0:
        print(f"\n[fcfv] bounds-checked build: {n} out-of-range indices seen by the kernels of this run")
        if n > 0:
            session.exitstatus = 1
