"""Multi-rank host logic on CPU: world_size 2 over gloo.  Each rank partitions the mesh, produces
the rows of the faces it owns (the ORACLE stands in for the CUDA kernels here -- the point is the
partitioning / ownership / ghost / tau-donor logic and the reduction plumbing), the ranks exchange
results with torch.distributed, and rank 0 checks that the union of owned rows is bit-identical to
the unpartitioned result and that the all-reduced sums of squares give the global norms."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mesh_name, setting, variant, form, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cases
        from helpers import csc_triple_to_csr
        from oracle import oracle as O
        from fcfv_nme23_b200.partition import partition_mesh
        case = cases.Case(mesh_name, setting, variant, form, tau_r=2.0 if setting == "inclusion" else 10.0)
        if setting == "random":
            from test_random_meshes import random_case
            gm, d = random_case(int(mesh_name.split(":")[1]))
        else:
            gm, d = cases.build(case)
        P = partition_mesh(gm, world, rank)
        m = P.mesh
        sf, se = P.scatter_face, P.scatter_elem
        neu = [sf(x) for x in d["neu"]]
        m.τ = None
        a, b, Z = O.ComputeFCFV(m, se(d["sex"]), se(d["sey"]), sf(d["VxDir"]), sf(d["VyDir"]), *neu, case.tau_r, form)
        fn = O.ElementAssemblyLoop if variant == 0 else O.ElementAssemblyLoopNEW
        Kuu, Muu, Kup, fu, fp, _, _ = fn(m, a, b, Z, sf(d["VxDir"]), sf(d["VyDir"]), *neu, form)
        rp, ci, v = csc_triple_to_csr(Kuu, (2 * m.nf, 2 * m.nf))
        rpp, cip, vp = csc_triple_to_csr(Kup, (2 * m.nf, m.nel))
        rows_l, rows_g = P.local_rows(), P.owned_rows()
        mine = {}
        for lr, gr in zip(rows_l, rows_g):
            mine[int(gr)] = (P.global_columns(ci[rp[lr]:rp[lr + 1]]).tolist(), v[rp[lr]:rp[lr + 1]].tolist(),
                             P.global_columns(cip[rpp[lr]:rpp[lr + 1]], "up").tolist(), vp[rpp[lr]:rpp[lr + 1]].tolist(),
                             float(fu[lr, 0]))
        fp_own = {int(g): float(x) for g, x in zip(P.elem_global[P.elem_owned], fp[P.elem_owned])}
        # residual sums of squares over owned elements / faces, all-reduced
        _, arr = O.ComputeResidualsFCFV_Stokes_o1(sf(d["Vxh"]), sf(d["Vyh"]), se(d["Pe"]), m, a, b, Z, se(d["sex"]), se(d["sey"]),
                                                  sf(d["VxDir"]), sf(d["VyDir"]), *neu, form, tau_mode="FACE", arrays=True)
        eo, fo = P.elem_owned, P.face_owned
        ss = torch.tensor([(arr["F_eq1"][eo] ** 2).sum(), (arr["F_eq2"][eo] ** 2).sum(), (arr["F_eq3"][eo] ** 2).sum(),
                           (arr["F_nodes_x"][fo] ** 2).sum(), (arr["F_nodes_y"][fo] ** 2).sum(), (arr["F_glob2"][eo] ** 2).sum()],
                          dtype=torch.float64)
        dist.all_reduce(ss)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, fp_own))
        if rank == 0:
            gm.τ = None
            A, B, ZZ = O.ComputeFCFV(gm, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], case.tau_r, form)
            GK, GM, GP, gfu, gfp, _, _ = fn(gm, A, B, ZZ, d["VxDir"], d["VyDir"], *d["neu"], form)
            grp, gci, gv = csc_triple_to_csr(GK, (2 * gm.nf, 2 * gm.nf))
            grpp, gcip, gvp = csc_triple_to_csr(GP, (2 * gm.nf, gm.nel))
            seen = {}
            fpall = {}
            for part, fpp in gathered:
                for k, val in part.items():
                    assert k not in seen, "row produced twice"
                    seen[k] = val
                fpall.update(fpp)
            # faces no element references belong to no rank: their rows are empty in the global system as well
            unseen = np.setdiff1d(np.arange(2 * gm.nf), np.fromiter(seen.keys(), dtype=np.int64))
            ok = bool(np.all(np.diff(grp)[unseen] == 0) and np.all(np.diff(grpp)[unseen] == 0) and np.all(gfu[unseen, 0] == 0.0))
            used = np.zeros(gm.nf, bool); used[np.asarray(gm.e2f).ravel() - 1] = True
            ok &= len(seen) == 2 * int(used.sum()) and len(fpall) == gm.nel
            for gr, (cols, vals, pc, pv, f) in seen.items():
                ok &= cols == gci[grp[gr]:grp[gr + 1]].tolist() and vals == gv[grp[gr]:grp[gr + 1]].tolist()
                ok &= pc == gcip[grpp[gr]:grpp[gr + 1]].tolist() and pv == gvp[grpp[gr]:grpp[gr + 1]].tolist()
                ok &= f == float(gfu[gr, 0])
            ok &= all(fpall[e] == float(gfp[e]) for e in range(gm.nel))
            gn = O.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], gm, A, B, ZZ, d["sex"], d["sey"], d["VxDir"],
                                                  d["VyDir"], *d["neu"], form, tau_mode="FACE")
            lens = np.array([4 * gm.nel, 2 * gm.nel, gm.nel, gm.nf, gm.nf, gm.nel], dtype=float)
            pn = np.sqrt(ss.numpy()) / np.sqrt(lens)
            big = gn > 1e-9
            ok &= bool(np.all(np.abs(pn[big] - gn[big]) <= 1e-12 * gn[big]) and np.all(pn[~big] < 1e-9))
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mesh_name,setting,variant,form", [
    ("LowRes", "inclusion", 0, "Gradient"),
    ("H1", "solkz", 1, "SymmetricGradient"),
    ("S6", "inclusion", 1, "Gradient"),
    ("random:5", "random", 1, "SymmetricGradient"),       # random Delaunay mesh, random numbering, faces without elements
])
def test_two_ranks_reproduce_global_rows(mesh_name, setting, variant, form):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mesh_name, setting, variant, form, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) is True
