#!/usr/bin/env python
"""bench.py -- throughput of the FCFV hot path (ComputeFCFV + assembly to device CSR + residuals).

    python bench.py --gpus 1 --steps K --warmup W            # this library on one B200
    torchrun ... bench.py --gpus N --steps K --warmup W      # one rank per GPU (driver launches it)
    python bench.py --impl reference ...                     # CPU arm: the reference's algorithm on host cores

Metric (BASELINE.json): Melements/s assembled+residual, FP64.  One step = one pass of the hot path
(local coefficients -> two-pass CSR assembly of Kuu/Kup/fu/fp -> residual norms) over one synthetic
mesh.  Workload at N=1: the 50M-element SolKz triangulation S(3536) of BASELINE config 4
(ElementAssemblyLoopNEW, :SymmetricGradient -- what examples/SolKzFCFV.jl runs), the mesh the
north_star target is quoted on; it fits one B200.  At N>1 the mesh is element-partitioned into
strips of quad rows with replicated ghost rows (weak scaling: 50M elements per rank); ranks
assemble their own rows with no data-path communication, NCCL all-reduces the six residual sums.

`value`   : device-resident pipeline, inputs already in HBM (timed with CUDA events on the
            library's stream, max over ranks).  Inputs (>40 GB) are far larger than L2 (126 MB).
`e2e`     : the same step through the reference-facing C-ABI calls with HOST (pinned) buffers:
            fcfv_compute_fcfv (mesh + ComputeFCFV) + fcfv_assemble + fcfv_residuals, H2D of every input
            and D2H of alpha/beta/Z/tau + nnz + norms inside the timed region.  The assembled CSR
            stays on the device (north_star: the GPU path ends at device-resident CSR).
`roofline`: dominant kernel (k_asm_fill), algorithmic bytes / CUDA-event time vs the measured HBM peak.
`cpu_baseline`: the oracle port (restated reference, C, 1 thread) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Melements/s assembled+residual (FP64)"
UNIT = "Melements/s"
FORM = "SymmetricGradient"
VARIANT = 1            # ElementAssemblyLoopNEW (examples/SolKzFCFV.jl:154)
N_DEFAULT = 3536       # S(3536): 50,013,184 elements (BASELINE config 4)
CPU_SAMPLE_N = 500     # S(500): 1,000,000 elements (BASELINE config 3) -- bounded CPU sample


_T0 = time.time()


def progress(msg):
    """Phase marker on stderr (rank 0): locates a stall in the driver's log."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


# BASELINE.json configs the bench can run: c4 (headline: SolKz, 50 M elements per GPU) and c5 (AnisoViscousInclusion-style:
# circular inclusion eta = [1, 1e-2] with interface faces tagged -1, south side Neumann, anisotropy factor delta != 1
# inside the inclusion -- the full 2x2 rheology path of ElementAssemblyLoopNEW --, tau_r = 2, 12.5 M elements per GPU,
# i.e. 100 M elements on 8 GPUs; examples/AnisoViscousInclusionFCFV_v0.jl:76-78,154-164)
CONFIGS = {
    "c4": dict(n1=N_DEFAULT, setting="solkz", tau_r=10.0, gen={}, delta_incl=None,
               name="SolKz S({n}) criss-cross triangulation, {nel} elements, ElementAssemblyLoopNEW :SymmetricGradient, tau_e=10*max(ke,1)"),
    "c5": dict(n1=1768, setting="inclusion", tau_r=2.0, gen=dict(eta_incl=1e-2, neumann_south=True), delta_incl=2.0,
               name="AnisoViscousInclusion-style S({n}) criss-cross triangulation, {nel} elements, eta=[1,1e-2], delta=2 in the inclusion, "
                    "interface faces bc=-1, south Neumann, ElementAssemblyLoopNEW :SymmetricGradient, tau_e=2"),
}
CONFIG = "c4"


def workload_name(n1, world, scaling):
    """Name of the workload both arms are quoted on (n1 = quads per side of the single-GPU mesh)."""
    n = int(round(n1 * world ** 0.5)) if scaling == "weak" else n1
    return CONFIGS[CONFIG]["name"].format(n=n, nel=4 * n * n)


def generate(h, n, row0, row1):
    """The rank's strip of the configured workload, generated on its GPU."""
    import numpy as np
    cfg = CONFIGS[CONFIG]
    h.generate_structured(n, row0, row1, setting=cfg["setting"], tau_r=cfg["tau_r"], **cfg["gen"])
    if cfg["delta_incl"] is not None:       # anisotropy factor inside the inclusion (the generator leaves delta = 1)
        import ctypes as C
        ke = np.empty(h.nel)
        h.check(h.lib.fcfv_get_mesh(h._h, None, None, None, None, None, None, C.c_void_p(ke.ctypes.data), None, None, None))
        h.update_fields(delta=np.where(ke != 1.0, cfg["delta_incl"], 1.0))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh-n", dest="n", type=int, default=None, help="quads per side of the per-rank-equivalent mesh (default 3536)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS), help="BASELINE.json config: c4 (headline) or c5")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=CPU_SAMPLE_N)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
def algorithmic_bytes(nel, nf, nnz_uu, nnz_up, rpw):
    """every distinct input array read once, every output written once (int32
    connectivity/columns, FP64 values).  Returns per-phase bytes for one step on this rank."""
    local = nel * (12 + 72 + 32) + nf * (4 + 16) + nel * 56 + nf * 8
    assembly = nel * (12 + 72 + 24 + 56) + nf * (4 + 8 + 16 + 32) + 12 * (nnz_uu + nnz_up) + 2 * rpw * (2 * nf + 1) + nf * 16 + nel * 8
    residual = nel * (12 + 72 + 24 + 56 + 16 + 8) + nf * (4 + 8 + 16 + 16 + 32) + 48
    return dict(local=local, assembly=assembly, residual=residual, total=local + assembly + residual)


def kernel_algorithmic_bytes(nel, nf, nnz_uu, nnz_up, rpw):
    """Bytes each hot kernel itself has to move once (its own inputs read once + its own outputs written once; face
    data that is only read on boundary faces is left out).  The fill kernel's figure is what ncu measures for it
    (29.9 GB at 50 M elements, profiles/*traffic*): its traffic is at the algorithmic minimum."""
    rec = 72 + 16 + 48            # per element: G, nx, ny | 1/alpha, c11 | beta, Z
    conn = 8 + 16 + 4 + 8         # per face: f2e, fconn, fbc, tau
    return {
        "k_local": nel * (12 + 4 + 72 + 32) + nel * 56 + nf * 8,
        "k_elem_coef": nel * 40,
        "k_asm_count": nel * rec + nf * conn,
        "k_asm_fill": nel * rec + nf * conn + 12 * (nnz_uu + nnz_up) + 2 * rpw * (2 * nf + 1) + nf * 16,
        "k_res_elem": nel * (12 + 4 + 4 + 72 + 24 + 56 + 16 + 8) + nf * (8 + 16),
        "k_res_late": nf * 4,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class NvmlClockSampler:
    """Same quantities through NVML (pynvml), every 10 ms: the timed region of the default run is ~0.13 s, which
    nvidia-smi's 100 ms loop covers with one to three rows.  Falls back to ClockSampler when NVML is not usable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.fallback, self.stop_flag = index, [], None, threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            try:
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self.dev = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            self.fallback = ClockSampler(index)

    def _loop(self):
        nv = self.nv
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)), int(reasons_fn(self.dev))))
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def start(self):
        if self.fallback:
            return self.fallback.start()
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.fallback:
            return self.fallback.stop()
        self.stop_flag.set()
        self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= r[1]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.REASONS if mask & bit], "samples": len(sm), "source": "nvml, 10 ms"}


def measured_traffic(kernel, nel):
    """DRAM bytes per launch of `kernel` from the newest committed ncu capture taken at this mesh size
    (profiles/*traffic*.json, written by tools/ncu_traffic.py); None when there is none."""
    import glob
    best = None
    for pth in sorted(glob.glob(os.path.join(ROOT, "profiles", "*traffic*.json"))):
        try:
            t = json.load(open(pth))
            if int(t["nel"]) == int(nel) and kernel in t["kernels"]:
                best = (float(t["kernels"][kernel]["dram_bytes"]), os.path.basename(pth))
        except Exception:
            pass
    return best


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
def cpu_port_throughput(mesh_arrays, data, reps=1, workspace=None):
    """Times the oracle port (C restatement of the reference's algorithm incl. its dense COO
    intermediates and sparse()/dropzeros, 1 thread) on host arrays.  Returns Melements/s.
    The ctypes wrapper's output arrays come from `workspace` (allocated by the caller, outside the timed
    region): only the C code's own allocations -- the reference's `zeros(...)` / `sparse` -- are timed."""
    from oracle import oracle as O
    from fcfv_nme23_b200.mesh import FCFV_Mesh
    O.lib()
    m = FCFV_Mesh(nel=int(mesh_arrays["Ω"].size), nf=int(mesh_arrays["bc"].size), nf_el=3)
    for k, v in mesh_arrays.items():
        setattr(m, k, v)
    if workspace is None:
        workspace = O.assembly_workspace(m.nel, m.nf)
    neu = [data["sxx"], data["syy"], data["sxy"], data["syx"]]
    best = None
    for _ in range(reps):
        m.τ = None
        t0 = time.perf_counter()
        a, b, Z = O.ComputeFCFV(m, data["sex"], data["sey"], data["VxDir"], data["VyDir"], *neu, 10.0, FORM)
        t1 = time.perf_counter()
        fn = O.ElementAssemblyLoopNEW if VARIANT == 1 else O.ElementAssemblyLoop
        out = fn(m, a, b, Z, data["VxDir"], data["VyDir"], *neu, FORM, workspace=workspace)
        t2 = time.perf_counter()
        O.ComputeResidualsFCFV_Stokes_o1(data["Vxh"], data["Vyh"], data["Pe"], m, a, b, Z, data["sex"], data["sey"],
                                         data["VxDir"], data["VyDir"], *neu, FORM, tau_mode="FACE")
        t3 = time.perf_counter()
        del out
        tot = t3 - t0
        if best is None or tot < best[0]:
            best = (tot, t1 - t0, t2 - t1, t3 - t2)
    return m.nel / best[0] / 1e6, best


def host_sample(n):
    """S(n) and its SolKz data built on the HOST (numpy + the oracle's geometry): the CPU arm never maps the
    product library.  Same fields as the device generator (tests/cases.py, checked against it in the GPU tests)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    m, d = cases.build(cases.Case(f"S{n}", "solkz", VARIANT, FORM, tau_r=10.0))
    mesh_arrays = {k: getattr(m, k) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe")}
    data = dict(sex=d["sex"], sey=d["sey"], VxDir=d["VxDir"], VyDir=d["VyDir"], sxx=d["neu"][0], syy=d["neu"][1],
                sxy=d["neu"][2], syx=d["neu"][3], Vxh=d["Vxh"], Vyh=d["Vyh"], Pe=d["Pe"])
    return mesh_arrays, data


# --------------------------------------------------------------------------------------------
def multi_gpu_parity(rank, world, local_rank, n=40):
    """Value proof for the N > 1 path, inside the bench run: a small S(n) mesh is split over the ranks of THIS
    job by the element partitioner (ghost elements replicated, face owner = rank of the highest adjacent element),
    every rank runs ComputeFCFV + assembly + residuals on its part with the library's NCCL all-reduce, and compares
    with the whole mesh on one GPU: every owned row of Kuu / Kup / fu and every owned fp entry bit for bit, the
    all-reduced sums of squares to 1e-12.  Returns the record for the JSON line (identical on all ranks)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from fcfv_nme23_b200 import api, lib as L
    from fcfv_nme23_b200.mesh import FCFV_Mesh
    from fcfv_nme23_b200.partition import partition_mesh
    full = api.Handle(local_rank)
    full.generate_structured(n, setting="solkz", tau_r=10.0)
    full.run_local(FORM); full.run_assemble(VARIANT, FORM)
    ss_full, _ = full.run_residuals(FORM, "FACE")
    grp, gci, gv = full.export_csr(L.KUU)
    grpp, gcip, gvp = full.export_csr(L.KUP)
    gfu, gfp = full.export_rhs()
    ma, d = full.get_mesh(), full.get_data()
    full.close()
    gm = FCFV_Mesh(nel=int(ma["Ω"].size), nf=int(ma["bc"].size), nf_el=3)
    for k, v in ma.items():
        setattr(gm, k, v)
    P = partition_mesh(gm, world, rank)
    h = api.Handle(local_rank)
    h.set_mesh(P.mesh)                                   # uploads the ownership masks as well
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(h.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    h.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    sf, se = P.scatter_face, P.scatter_elem
    h.set_sources(se(d["sex"]), se(d["sey"]))
    h.set_face_data(sf(d["VxDir"]), sf(d["VyDir"]), sf(d["sxx"]), sf(d["syy"]), sf(d["sxy"]), sf(d["syx"]))
    h.set_solution(sf(d["Vxh"]), sf(d["Vyh"]), se(d["Pe"]))
    h.run_local(FORM); h.run_assemble(VARIANT, FORM)
    ss, _ = h.run_residuals(FORM, "FACE")               # all-reduced by the library (NCCL)
    rp, ci, v = h.export_csr(L.KUU)
    rpp, cip, vp = h.export_csr(L.KUP)
    fu, fp = h.export_rhs()
    h.close()
    ok, rows = True, 0
    for lr, gr in zip(P.local_rows(), P.owned_rows()):
        ok &= np.array_equal(P.global_columns(ci[rp[lr]:rp[lr + 1]]), gci[grp[gr]:grp[gr + 1]])
        ok &= np.array_equal(v[rp[lr]:rp[lr + 1]], gv[grp[gr]:grp[gr + 1]])
        ok &= np.array_equal(P.global_columns(cip[rpp[lr]:rpp[lr + 1]], "up"), gcip[grpp[gr]:grpp[gr + 1]])
        ok &= np.array_equal(vp[rpp[lr]:rpp[lr + 1]], gvp[grpp[gr]:grpp[gr + 1]])
        ok &= fu[lr] == gfu[gr]
        rows += 1
    ok &= bool(np.array_equal(fp[P.elem_owned], gfp[P.elem_global[P.elem_owned]]))
    rel = float(np.max(np.abs(ss - ss_full) / np.maximum(np.abs(ss_full), 1e-300)))
    t = torch.tensor([1.0 if ok else 0.0, -rel], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    cnt = torch.tensor([rows, int(P.elem_owned.sum())], dtype=torch.int64, device="cuda")
    dist.all_reduce(cnt)
    used = np.zeros(gm.nf, bool); used[np.asarray(gm.e2f).ravel() - 1] = True
    complete = int(cnt[0].item()) == 2 * int(used.sum()) and int(cnt[1].item()) == gm.nel
    return {"mesh": f"S({n}), {gm.nel} elements over {world} ranks (host partitioner, ghost replication)",
            "bitwise_rows": bool(t[0].item() == 1.0 and complete), "rows_checked": int(cnt[0].item()),
            "max_rel": float(-t[1].item()), "tolerance": 1e-12,
            "checked": "Kuu/Kup rows (columns + values), fu, fp bit for bit vs the single-GPU run; 6 all-reduced sums of squares"}


# --------------------------------------------------------------------------------------------
def host_cpu():
    """(model name, logical cores) of the box the CPU arm runs on."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return model, os.cpu_count()


def run_reference(args, rank, world):
    """CPU arm: the reference's algorithm (oracle port; Julia itself is not installable here) on the
    host cores, each step one bounded sample S(cpu_n) of the workload."""
    if rank != 0:
        return
    n = args.cpu_n
    mesh_arrays, data = host_sample(n)
    from oracle import oracle as O
    ws = O.assembly_workspace(int(mesh_arrays["Ω"].size), int(mesh_arrays["bc"].size))
    for _ in range(min(args.warmup, 1)):
        cpu_port_throughput(mesh_arrays, data, workspace=ws)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_throughput(mesh_arrays, data, workspace=ws)
    dt = time.perf_counter() - t0
    nel = int(mesh_arrays["Ω"].size)
    val = nel * args.steps / dt / 1e6
    sample = f"S({n}) = {nel} elements per step, SolKz, ElementAssemblyLoopNEW :SymmetricGradient, restated reference (C, -O2, no FMA), not Julia"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N_DEFAULT if args.n is None else args.n, args.gpus, args.scaling),
                       "sample": f"each step = one bounded sample S({n}) of that workload on 1 host thread (the reference has no threading)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                             "host_cpu": host_cpu()[0], "host_cores": host_cpu()[1]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fcfv_nme23_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    global CONFIG
    CONFIG = args.config
    if CONFIG != "c4":
        args.no_cpu = True           # the CPU arm is quoted on the headline workload only
    n1 = args.n or CONFIGS[CONFIG]["n1"]
    if args.scaling == "weak":
        n = int(round(n1 * world ** 0.5))       # world * (4 n1^2) elements in total
    else:
        n = n1
    row0, row1 = rank * n // world, (rank + 1) * n // world
    h = api.Handle(local_rank)
    generate(h, n, row0, row1)
    owned_el = 4 * n * (row1 - row0)
    total_el = 4 * n * n
    if world > 1:
        # NCCL communicator of the library (residual-norm all-reduce on its own stream)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(h.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        h.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    st = torch.cuda.ExternalStream(h.stream)
    h.set_async(True)
    # Residual tau mode of the headline: FACE (the face value of tau, what is consistent with the assembled system and
    # what a partitioned mesh can evaluate locally).  The reference function's literal tau[e] (Residuals:17,54: the
    # FACE array indexed by the ELEMENT id) is timed beside it at N = 1 (`tau_modes`).
    TAU = "FACE"

    def step(tau_mode=TAU, hh=None):
        # one pass = local coefficients -> assembly -> residual sums (+ NCCL all-reduce at N > 1); fcfv_run_pipeline
        # issues exactly the launches of run_local + run_assemble + run_residuals, replayed as one CUDA graph
        (hh or h).run_pipeline(VARIANT, FORM, tau_mode)

    def barrier():
        h.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, reps):
        """ms for `reps` calls of fn: CUDA events on the library's stream, barrier + synchronize on both sides,
        max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st)
        barrier()
        t = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t

    progress(f"mesh generated ({h.nel} elements per rank), warm-up")
    for _ in range(args.warmup):
        step()
    barrier()
    progress("headline loop")
    # ---- headline: K steps, per-kernel timers OFF
    l0 = h.kernel_launches
    sampler = NvmlClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = h.kernel_launches - l0
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = total_el * args.steps / (ms * 1e-3) / 1e6
    # ---- per-kernel device times: a second, separate repetition with an event pair around every launch
    progress("per-kernel timers")
    h.set_timing(True)
    ksteps = max(1, min(args.steps, 5))
    for _ in range(ksteps):
        step()
    h.sync()
    timers = h.get_timers()
    h.set_timing(False)
    # ---- the same step with the reference's literal tau[e] in the residual
    ms_ref = timed(lambda: step("REFERENCE"), ksteps) / ksteps if world == 1 else None
    ss, norms = h.run_residuals(FORM, TAU)
    nnz_uu, nnz_up, _ = h.get_nnz()

    # ---- roofline of the dominant kernel -----------------------------------------------------
    rpw = h.lib.fcfv_index_width(h._h)
    ab = algorithmic_bytes(h.nel, h.nf, nnz_uu, nnz_up, rpw)
    kab = kernel_algorithmic_bytes(h.nel, h.nf, nnz_uu, nnz_up, rpw)
    peak, peak_src = measured_peak()
    kms = {k: v[0] / v[1] for k, v in timers.items() if v[1] > 0}          # ms per launch
    dom = max(kms, key=lambda k: kms[k])
    dom_ms = kms[dom]
    dom_bytes = kab.get(dom)
    traffic = measured_traffic(dom, h.nel)
    if dom_bytes is None:
        dom_bytes = traffic[0] if traffic else ab["assembly"]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    step_ms = ms / args.steps
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "basis": "bytes this kernel itself has to read and write once (kernel_algorithmic_bytes) / its CUDA-event time",
                "kernel_ms": dom_ms,
                "pipeline": {"algorithmic_bytes_per_step": ab["total"], "achieved": ab["total"] / (step_ms * 1e-3) / 1e9,
                             "frac": ab["total"] / (step_ms * 1e-3) / 1e9 / peak,
                             "bytes_per_element": ab["total"] / h.nel},
                "kernel_ms_per_step": {k: v[0] / ksteps for k, v in timers.items() if v[1] > 0},
                "kernel_timing": f"separate repetition of {ksteps} steps with per-kernel events; the headline loop runs without them"}
    # SURVEY 8(d) bytes of each phase against ALL kernels of the phase
    phases = {"local": ("k_local",), "assembly": ("k_elem_coef", "k_asm_count", "k_scan", "k_asm_fill", "other"),
              "residual": ("k_res_elem", "k_res_late", "k_reduce_partials")}
    for ph, ks in phases.items():
        pms = sum(timers[k][0] for k in ks if k in timers) / ksteps
        if pms > 0:
            roofline[ph + "_phase"] = {"ms_per_step": pms, "algorithmic_bytes": ab[ph], "achieved": ab[ph] / (pms * 1e-3) / 1e9,
                                       "frac": ab[ph] / (pms * 1e-3) / 1e9 / peak}
    # the SURVEY 8(d) attribution of round 1 (whole assembly phase's bytes over the fill kernel alone), kept for comparison
    if "k_asm_fill" in kms:
        roofline["frac_algorithmic_phase_bytes"] = ab["assembly"] / (kms["k_asm_fill"] * 1e-3) / 1e9 / peak
    if traffic:
        roofline["traffic_rate"] = {"achieved": traffic[0] / (dom_ms * 1e-3) / 1e9, "frac": traffic[0] / (dom_ms * 1e-3) / 1e9 / peak}
    tau_modes = {"FACE": {"ms_per_step": step_ms, "value": value}}
    if ms_ref is not None:
        tau_modes["REFERENCE"] = {"ms_per_step": ms_ref, "value": total_el / (ms_ref * 1e-3) / 1e6}

    # ---- N > 1: strong scaling of the single-GPU workload S(n1) in the same run ------------------
    strong = None
    if world > 1 and args.scaling == "weak":
        progress("strong-scaling point")
        hs = api.Handle(local_rank)
        generate(hs, n1, rank * n1 // world, (rank + 1) * n1 // world)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(hs.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        hs.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        hs.set_async(True)
        sts = torch.cuda.ExternalStream(hs.stream)
        for _ in range(max(3, args.warmup)):
            step(TAU, hs)
        hs.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); hs.sync(); dist.barrier()
        e0.record(sts)
        for _ in range(args.steps):
            step(TAU, hs)
        e1.record(sts)
        hs.sync(); barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_s = float(tt.item()) / args.steps
        strong = {"workload": workload_name(n1, 1, "strong"), "ms_per_step": ms_s, "value": 4 * n1 * n1 / (ms_s * 1e-3) / 1e6,
                  "unit": UNIT, "note": "the N = 1 mesh split over the N ranks (value / the N = 1 value / N = strong-scaling efficiency)"}
        hs.close()
    # ---- N > 1: value proof on the same communicator ------------------------------------------
    progress("multi-GPU parity record" if world > 1 else "e2e")
    parity = multi_gpu_parity(rank, world, local_rank) if world > 1 else None
    # ---- e2e through the host-buffer C ABI ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, h, rank, world, total_el, barrier, norms)
    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        mesh_arrays, data = host_sample(args.cpu_n)
        val, parts = cpu_port_throughput(mesh_arrays, data, reps=2)
        cpu = {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "host_cpu": host_cpu()[0], "host_cores": host_cpu()[1],
               "sample": f"S({args.cpu_n}) = {int(mesh_arrays['Ω'].size)} elements (BASELINE config 3 size), best of 2; "
                         f"local {parts[1]:.2f}s + assembly {parts[2]:.2f}s + residual {parts[3]:.2f}s; "
                         "restated reference (C, -O2, no FMA), not Julia"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n1, world, args.scaling),
                           "elements_per_rank": owned_el, "partition": f"{world} strips of quad rows + ghost rows" if world > 1 else "none",
                           "l2": "inputs (>40 GB per rank at the default size) exceed the 126 MB L2; no flush needed",
                           "nnz_uu_rank0": nnz_uu, "nnz_up_rank0": nnz_up, "rowptr_bytes": rpw,
                           "residual_tau_mode": TAU, "residual_norms": [float(x) for x in norms]},
                "tau_modes": tau_modes,
                "clocks": clocks, "gpu_launches": launches, "roofline": roofline}
        if strong is not None:
            line["strong"] = strong
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, hdev, rank, world, total_el, barrier, norms_device):
    """The same step through the reference-facing host-buffer calls, exactly as the shims (julia/FCFV_B200.jl,
    fcfv_nme23_b200/api.py) issue them for the three calls of a driver (SolKzFCFV.jl:150,155,205):

        ComputeFCFV            -> fcfv_compute_fcfv: [first call on a mesh: the mesh arrays | later: fcfv_new_epoch and the
                                  mutable fields ke, delta, tau_e] + sources / boundary data up, alpha / beta / Z / tau down
        ElementAssemblyLoopNEW -> fcfv_update_fields(...), fcfv_assemble(alpha, beta, Z, VDir, sigmaNeu)
        ComputeResiduals..._o1 -> fcfv_update_fields(...), fcfv_residuals(Vh, Pe, alpha, beta, Z, sources, VDir, sigmaNeu)

    with pinned HOST buffers, H2D and D2H inside the timed region; the handle's upload cache is on, as in the shims
    (an array the device already mirrors is not sent again).  Byte counts are the library's own transfer counters."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from fcfv_nme23_b200 import api, lib as L

    nel, nf = hdev.nel, hdev.nf
    mesh_arrays, data = hdev.get_mesh(), hdev.get_data()
    eo, fo, nel_g, nf_g = hdev.get_ownership()

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).pin_memory()
        return t
    P = {k: pin(v) for k, v in mesh_arrays.items()}
    P["eo"], P["fo"] = pin(eo), pin(fo)
    Dd = {k: pin(v) for k, v in data.items()}
    del mesh_arrays, data
    out = {"alpha": torch.empty(nel, dtype=torch.float64).pin_memory(), "beta": torch.empty(2 * nel, dtype=torch.float64).pin_memory(),
           "Z": torch.empty(4 * nel, dtype=torch.float64).pin_memory(), "tau": torch.zeros(nf, dtype=torch.float64).pin_memory()}
    ptr = lambda t: C.c_void_p(t.data_ptr())
    h = hdev   # reuse the handle (its buffers are already sized); ownership/comm state is kept
    lib = h.lib
    form = api.formulation_id(FORM)
    norms = np.zeros(6)
    nnz = [C.c_int64(0) for _ in range(3)]
    sec = C.c_double(0.0)

    def three_calls(h, P, Dd, out, nel, with_mesh=True, ownership=True, cached=True):
        def refresh(tau=True):
            h.check(lib.fcfv_update_fields(h._h, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, ptr(out["tau"]) if tau else None))
        sig = [ptr(Dd[k]) for k in ("sxx", "syy", "sxy", "syx")]
        loc = [ptr(out[k]) for k in ("alpha", "beta", "Z")]
        nf_ = P["bc"].numel()
        # ComputeFCFV = ONE library call (fcfv_compute_fcfv).  First call on an FCFV_Mesh object (or a driver that rebuilds
        # its mesh): the mesh goes up in front of the element kernel; later calls: only the mutable fields ke, delta, tau_e
        send_mesh = with_mesh
        if with_mesh and world > 1 and ownership:       # partitioned mesh of one rank: ownership masks go up between the
            h.check(lib.fcfv_set_mesh(h._h, nel, nf_, ptr(P["e2f"]), ptr(P["bc"]), ptr(P["Ω"]), ptr(P["n_x"]), ptr(P["n_y"]),   # mesh and the call
                                      ptr(P["Γ"]), ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel))
            h.check(lib.fcfv_set_ownership(h._h, ptr(P["eo"]), ptr(P["fo"]), nel_g, nf_g, None))
            send_mesh = False
        elif cached and not with_mesh:      # the shims skip this right after they have uploaded the mesh themselves
            h.check(lib.fcfv_new_epoch(h._h))
        geo = [ptr(P[k]) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ")] if send_mesh else [None] * 6
        # mesh.tau goes in as well: the library reads it only when the mesh has faces without elements (ComputeFCFV
        # overwrites every other entry)
        h.check(lib.fcfv_compute_fcfv(h._h, nel, nf_, *geo, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, ptr(out["tau"]), form, 1.0,
                                      ptr(Dd["sex"]), ptr(Dd["sey"]), ptr(Dd["VxDir"]), ptr(Dd["VyDir"]), *sig, *loc, ptr(out["tau"])))
        # ElementAssemblyLoopNEW
        refresh()
        h.check(lib.fcfv_assemble(h._h, VARIANT, form, 1.0, *loc, ptr(Dd["VxDir"]), ptr(Dd["VyDir"]), *sig, 0, C.byref(nnz[0]),
                                  C.byref(nnz[1]), C.byref(nnz[2]), C.byref(sec)))
        # ComputeResidualsFCFV_Stokes_o1
        refresh()
        h.check(lib.fcfv_residuals(h._h, form, L.TAU_FACE, ptr(Dd["Vxh"]), ptr(Dd["Vyh"]), ptr(Dd["Pe"]), *loc, ptr(Dd["sex"]),
                                   ptr(Dd["sey"]), ptr(Dd["VxDir"]), ptr(Dd["VyDir"]), *sig, C.c_void_p(norms.ctypes.data), None, None, None))

    def step(with_mesh=True, cached=True):
        three_calls(h, P, Dd, out, nel, with_mesh, True, cached)

    def timed_host(fn, reps):
        fn()
        barrier()
        h.transfer_stats(reset=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        h.sync()
        dt = time.perf_counter() - t0
        up, down, hit = h.transfer_stats(reset=True)
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            b = torch.tensor([up, down, hit], dtype=torch.int64, device="cuda")
            dist.all_reduce(b)                      # whole job: bytes of all ranks
            up, down, hit = (int(x) for x in b.tolist())
        return dt / reps, up // reps, down // reps, hit // reps

    h.set_async(False)
    h.upload_cache(True)
    steps = max(1, min(args.steps, 3))
    progress("e2e: three calls with set_mesh")
    dt, h2d, d2h, hit = timed_host(step, steps)
    # same inputs, same kernels: the norms that came back through the host-buffer calls are those of the device-resident pass
    nd = np.asarray(norms_device, dtype=np.float64)
    scale = np.maximum(np.abs(nd), 1e-300)
    check = {"max_rel_vs_device_pipeline": float(np.max(np.abs(norms - nd) / scale)), "tolerance": 1e-12,
             "nnz": [int(x.value) for x in nnz[:2]]}
    check["ok"] = bool(check["max_rel_vs_device_pipeline"] <= 1e-12)
    # same pass on a mesh that is already resident (second and later passes on one FCFV_Mesh object)
    progress("e2e: three calls, mesh resident")
    dt2, h2d2, d2h2, hit2 = timed_host(lambda: step(with_mesh=False), steps)
    # the three calls with the upload cache off (every call sends everything it is given: what round 1 measured,
    # with alpha / beta / Z now also going back up as the shims pass them)
    h.upload_cache(False)
    progress("e2e: three calls, cache off")
    dt3, h2d3, d2h3, _ = timed_host(lambda: step(cached=False), 1)
    h.upload_cache(True)
    # the same pass through the single-call entry fcfv_pass (a driver that replaces its three calls by one)
    def fused(with_mesh=True):
        if with_mesh:
            h.check(lib.fcfv_set_mesh(h._h, nel, nf, ptr(P["e2f"]), ptr(P["bc"]), ptr(P["Ω"]), ptr(P["n_x"]), ptr(P["n_y"]),
                                      ptr(P["Γ"]), ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel))
            if world > 1:
                h.check(lib.fcfv_set_ownership(h._h, ptr(P["eo"]), ptr(P["fo"]), nel_g, nf_g, None))
        h.check(lib.fcfv_new_epoch(h._h))
        h.check(lib.fcfv_update_fields(h._h, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, None))
        h.check(lib.fcfv_pass(h._h, VARIANT, form, 1.0, L.TAU_FACE, ptr(Dd["sex"]), ptr(Dd["sey"]), ptr(Dd["VxDir"]), ptr(Dd["VyDir"]),
                              ptr(Dd["sxx"]), ptr(Dd["syy"]), ptr(Dd["sxy"]), ptr(Dd["syx"]), ptr(Dd["Vxh"]), ptr(Dd["Vyh"]), ptr(Dd["Pe"]),
                              ptr(out["alpha"]), ptr(out["beta"]), ptr(out["Z"]), ptr(out["tau"]), 0, C.byref(nnz[0]), C.byref(nnz[1]),
                              C.byref(nnz[2]), C.c_void_p(norms.ctypes.data)))
    fused_res = {}
    for tag, wm in (("with_mesh", True), ("mesh_resident", False)):
        progress(f"e2e: fcfv_pass {tag}")
        dtf, upf, downf, _ = timed_host(lambda: fused(wm), steps)
        fused_res[tag] = {"value": total_el / dtf / 1e6, "ms_per_step": 1e3 * dtf, "h2d_bytes_per_step": int(upf), "d2h_bytes_per_step": int(downf)}
    fused_res["note"] = ("fcfv_pass: one call instead of three; not the headline because a driver has to change three lines to use it")
    # N > 1: the same three calls through ONE multi-GPU handle (fcfv_create_multi) in one process -- what a single-process
    # driver of the reference gets.  Rank 0 drives all GPUs of the box on a stand-alone S(1000) mesh (4 M elements: the
    # host-side partitioning and scatter of one process bound this figure, not the GPUs); the other ranks wait on a
    # host-side (gloo) barrier so that their GPUs are idle.  Partitioning and scatter are inside the timed region.
    single = None
    if world > 1:
        barrier()
        gl = dist.new_group(backend="gloo")
        if rank == 0:
            progress("e2e: one multi-GPU handle in one process")
            try:
                ns = min(1000, args.n or N_DEFAULT)
                hsm = api.Handle(0)
                hsm.generate_structured(ns, 0, ns, setting="solkz", tau_r=10.0)
                Ps = {k: pin(v) for k, v in hsm.get_mesh().items()}
                Ds = {k: pin(v) for k, v in hsm.get_data().items()}
                nels, nfs = hsm.nel, hsm.nf
                hsm.close()
                outs = {"alpha": torch.empty(nels, dtype=torch.float64).pin_memory(), "beta": torch.empty(2 * nels, dtype=torch.float64).pin_memory(),
                        "Z": torch.empty(4 * nels, dtype=torch.float64).pin_memory(), "tau": torch.zeros(nfs, dtype=torch.float64).pin_memory()}
                hm = api.Handle(ngpus=world)
                hm.upload_cache(True)
                three_calls(hm, Ps, Ds, outs, nels, True, False, True)
                t0 = time.perf_counter()
                for _ in range(2):
                    three_calls(hm, Ps, Ds, outs, nels, True, False, True)
                hm.sync()
                dts = (time.perf_counter() - t0) / 2
                info, psec = hm.partition_info()
                single = {"value": nels / dts / 1e6, "unit": UNIT, "ms_per_step": 1e3 * dts, "elements": int(nels), "gpus": world,
                          "partition_seconds": psec, "ghost_elements": int(sum(i[0] - i[2] for i in info)),
                          "note": "one process, one handle over all GPUs: fcfv_set_mesh partitions on the host (C++), every call "
                                  "scatters its global arrays; bound by the host-side gather and PCIe of one process"}
                hm.close()
            except Exception as ex:      # reported, not fatal: the per-rank figures above stand on their own
                single = {"error": str(ex)[:300]}
        dist.barrier(group=gl)
        barrier()
    return {"value": total_el / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "upload_cache_hit_bytes_per_step": int(hit), "steps": steps, "ms_per_step": 1e3 * dt,
            "mesh_resident": {"value": total_el / dt2 / 1e6, "ms_per_step": 1e3 * dt2, "h2d_bytes_per_step": int(h2d2),
                              "d2h_bytes_per_step": int(d2h2), "upload_cache_hit_bytes_per_step": int(hit2),
                              "note": "no fcfv_set_mesh (mesh uploaded once per FCFV_Mesh object)"},
            "cache_off": {"value": total_el / dt3 / 1e6, "ms_per_step": 1e3 * dt3, "h2d_bytes_per_step": int(h2d3), "d2h_bytes_per_step": int(d2h3),
                          "note": "same calls with fcfv_upload_cache(0): every call sends every array it is given"},
            "single_process_multi_gpu_handle": single,
            "fused_pass": fused_res,
            "result_check": check,
            "face_lists": dict(zip(("active", "dirichlet_faces", "neumann_faces", "list_uploads"), h.sparse_face_info())),
            "compute_fcfv_pipelined_calls": h.pipelined_calls,
            "api": "per pass what the shims issue for ComputeFCFV / ElementAssemblyLoopNEW / ComputeResidualsFCFV_Stokes_o1: "
                   "fcfv_compute_fcfv (with the mesh arrays: first call on a mesh object; mesh_resident: fcfv_new_epoch + the mutable "
                   "fields only) | fcfv_update_fields + fcfv_assemble | fcfv_update_fields + fcfv_residuals, with pinned host buffers "
                   "and the upload cache on; VxDir / VyDir / Neumann arrays cross the link as their values on the Dirichlet / "
                   "Neumann faces (fcfv_sparse_face_data); byte counts from fcfv_transfer_stats (all ranks); CSR left device-resident"}


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries print to stdout on their own (NCCL's version banner): keep fd 1 for the result line only
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
