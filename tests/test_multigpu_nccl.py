"""Two ranks, two GPUs, NCCL: the only collectives of the path (all-reduce of the six residual sums of
squares, of the two Powell-Hestenes norms and of the pressure sum) against the single-GPU run of the
whole mesh.  Needs >= 2 CUDA devices (skipped otherwise; run with `gpurun --gpus 2`)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, FORM, VARIANT = 40, "SymmetricGradient", 1


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _rank_main(rank, world, q_id, q_out):
    sys.path.insert(0, ROOT)
    from fcfv_nme23_b200 import api, lib as L
    h = api.Handle(rank)
    row0, row1 = rank * N // world, (rank + 1) * N // world
    h.generate_structured(N, row0, row1, setting="solkz", tau_r=10.0)
    if rank == 0:
        uid = h.comm_unique_id()
        for _ in range(world - 1):
            q_id.put(uid)
    else:
        uid = q_id.get(timeout=120)
    h.comm_init(uid, rank, world)
    h.run_local(FORM)
    h.run_assemble(VARIANT, FORM)
    ss, norms = h.run_residuals(FORM, "FACE")                 # all-reduced over the ranks
    d = h.get_data()
    eo, fo, nel_g, nf_g = h.get_ownership()
    u = np.concatenate([d["Vxh"], d["Vyh"]])
    nrm = np.zeros(2)
    import ctypes as C
    h.check(h.lib.fcfv_ph_residual(h._h, C.c_void_p(u.ctypes.data), C.c_void_p(d["Pe"].ctypes.data), None, None,
                                   C.c_void_p(nrm.ctypes.data)))
    pe = d["Pe"].copy()
    mean = C.c_double(0.0)
    h.check(h.lib.fcfv_center_pressure(h._h, C.c_void_p(pe.ctypes.data), C.byref(mean)))
    q_out.put((rank, ss.tolist(), norms.tolist(), nrm.tolist(), float(mean.value), int(eo.sum()), int(fo.sum()), nel_g, nf_g))
    h.close()


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs two CUDA devices")
def test_two_rank_allreduce_matches_single_gpu():
    sys.path.insert(0, ROOT)
    from fcfv_nme23_b200 import api
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_rank_main, args=(r, world, q_id, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q_out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single GPU, whole mesh
    full = api.Handle(0)
    full.generate_structured(N, setting="solkz", tau_r=10.0)
    full.run_local(FORM); full.run_assemble(VARIANT, FORM)
    ss_full, norms_full = full.run_residuals(FORM, "FACE")
    d = full.get_data()
    import ctypes as C
    u = np.concatenate([d["Vxh"], d["Vyh"]])
    nrm_full = np.zeros(2)
    full.check(full.lib.fcfv_ph_residual(full._h, C.c_void_p(u.ctypes.data), C.c_void_p(d["Pe"].ctypes.data), None, None,
                                         C.c_void_p(nrm_full.ctypes.data)))
    full.close()
    rel = lambda a, b: np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300))
    assert sum(r[5] for r in res) == 4 * N * N and sum(r[6] for r in res) == 6 * N * N + 2 * N   # ownership is a partition
    for r in res:
        assert r[7] == 4 * N * N and r[8] == 6 * N * N + 2 * N
        assert rel(r[1], ss_full) <= 1e-12 and rel(r[2], norms_full) <= 1e-12        # both ranks hold the global values
        assert rel(r[3], nrm_full) <= 1e-12
        assert abs(r[4] - d["Pe"].mean()) <= 1e-12 * np.abs(d["Pe"]).max()
    assert res[0][1] == res[1][1]                                                    # identical bits on both ranks
