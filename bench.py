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
            fcfv_set_mesh + fcfv_compute_local + fcfv_assemble + fcfv_residuals, H2D of every input
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


def workload_name(n1, world, scaling):
    """Name of the workload both arms are quoted on (n1 = quads per side of the single-GPU mesh)."""
    n = int(round(n1 * world ** 0.5)) if scaling == "weak" else n1
    return (f"SolKz S({n}) criss-cross triangulation, {4 * n * n} elements, ElementAssemblyLoopNEW :SymmetricGradient, "
            "tau_e=10*max(ke,1)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh-n", dest="n", type=int, default=None, help="quads per side of the per-rank-equivalent mesh (default 3536)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
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
def cpu_port_throughput(mesh_arrays, data, reps=1):
    """Times the oracle port (C restatement of the reference's algorithm incl. its dense COO
    intermediates and sparse()/dropzeros, 1 thread) on host arrays.  Returns Melements/s."""
    from oracle import oracle as O
    from fcfv_nme23_b200.mesh import FCFV_Mesh
    O.lib()
    m = FCFV_Mesh(nel=int(mesh_arrays["Ω"].size), nf=int(mesh_arrays["bc"].size), nf_el=3)
    for k, v in mesh_arrays.items():
        setattr(m, k, v)
    neu = [data["sxx"], data["syy"], data["sxy"], data["syx"]]
    best = None
    for _ in range(reps):
        m.τ = None
        t0 = time.perf_counter()
        a, b, Z = O.ComputeFCFV(m, data["sex"], data["sey"], data["VxDir"], data["VyDir"], *neu, 10.0, FORM)
        t1 = time.perf_counter()
        fn = O.ElementAssemblyLoopNEW if VARIANT == 1 else O.ElementAssemblyLoop
        out = fn(m, a, b, Z, data["VxDir"], data["VyDir"], *neu, FORM)
        t2 = time.perf_counter()
        O.ComputeResidualsFCFV_Stokes_o1(data["Vxh"], data["Vyh"], data["Pe"], m, a, b, Z, data["sex"], data["sey"],
                                         data["VxDir"], data["VyDir"], *neu, FORM, tau_mode="FACE")
        t3 = time.perf_counter()
        del out
        tot = t3 - t0
        if best is None or tot < best[0]:
            best = (tot, t1 - t0, t2 - t1, t3 - t2)
    return m.nel / best[0] / 1e6, best


def sample_inputs(n):
    """Host copies of S(n) and its SolKz data, produced by the device generator (setup, not timed)."""
    from fcfv_nme23_b200 import api
    h = api.Handle(0)
    h.generate_structured(n, setting="solkz", tau_r=10.0)
    mesh_arrays, data = h.get_mesh(), h.get_data()
    h.close()
    return mesh_arrays, data


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
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        mesh_arrays, data = sample_inputs(n)
    else:   # no GPU: build the sample on the host (slow path, tests only)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cases
        m, d = cases.build(cases.Case(f"S{n}", "solkz", VARIANT, FORM, tau_r=10.0))
        mesh_arrays = {k: getattr(m, k) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe")}
        data = dict(sex=d["sex"], sey=d["sey"], VxDir=d["VxDir"], VyDir=d["VyDir"], sxx=d["neu"][0], syy=d["neu"][1],
                    sxy=d["neu"][2], syx=d["neu"][3], Vxh=d["Vxh"], Vyh=d["Vyh"], Pe=d["Pe"])
    for _ in range(min(args.warmup, 1)):
        cpu_port_throughput(mesh_arrays, data)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_throughput(mesh_arrays, data)
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
    n1 = args.n or N_DEFAULT
    if args.scaling == "weak":
        n = int(round(n1 * world ** 0.5))       # world * (4 n1^2) elements in total
    else:
        n = n1
    row0, row1 = rank * n // world, (rank + 1) * n // world
    h = api.Handle(local_rank)
    h.generate_structured(n, row0, row1, setting="solkz", tau_r=10.0)
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

    def step():
        h.run_local(FORM)
        h.run_assemble(VARIANT, FORM)
        h.run_residuals(FORM, "FACE", fetch=False)

    def barrier():
        h.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    h.set_timing(True)
    l0 = h.kernel_launches
    sampler = NvmlClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = h.kernel_launches - l0
    timers = h.get_timers()
    h.set_timing(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    ss, norms = h.run_residuals(FORM, "FACE")
    nnz_uu, nnz_up, _ = h.get_nnz()
    value = total_el * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel -----------------------------------------------------
    rpw = h.lib.fcfv_index_width(h._h)
    ab = algorithmic_bytes(h.nel, h.nf, nnz_uu, nnz_up, rpw)
    peak, peak_src = measured_peak()
    dom = max((k for k in timers if timers[k][1] > 0), key=lambda k: timers[k][0])
    dom_ms = timers[dom][0] / timers[dom][1]
    phase_of = {"k_asm_fill": "assembly", "k_asm_count": "assembly", "k_res_elem": "residual", "k_local": "local"}
    dom_bytes = ab[phase_of.get(dom, "assembly")]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = measured_traffic(dom, h.nel)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "kernel_ms": dom_ms,
                "pipeline": {"algorithmic_bytes_per_step": ab["total"], "achieved": ab["total"] * args.steps / (ms * 1e-3) / 1e9,
                             "frac": ab["total"] * args.steps / (ms * 1e-3) / 1e9 / peak,
                             "bytes_per_element": ab["total"] / h.nel},
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timers.items() if v[1] > 0}}
    # the same algorithmic bytes against every kernel of the assembly phase (coefficients, count pass, scans, fill),
    # and the fill kernel's own measured DRAM traffic against its duration
    asm_ms = sum(timers[k][0] for k in ("k_elem_coef", "k_asm_count", "k_scan", "k_asm_fill") if k in timers) / args.steps
    if asm_ms > 0:
        roofline["assembly_phase"] = {"ms_per_step": asm_ms, "achieved": ab["assembly"] / (asm_ms * 1e-3) / 1e9,
                                      "frac": ab["assembly"] / (asm_ms * 1e-3) / 1e9 / peak}
    if traffic:
        roofline["traffic_rate"] = {"achieved": traffic[0] / (dom_ms * 1e-3) / 1e9, "frac": traffic[0] / (dom_ms * 1e-3) / 1e9 / peak}

    # ---- e2e through the host-buffer C ABI ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, h, rank, world, total_el, barrier)
    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        mesh_arrays, data = sample_inputs(args.cpu_n)
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
                           "residual_norms": [float(x) for x in norms]},
                "clocks": clocks, "gpu_launches": launches, "roofline": roofline}
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, hdev, rank, world, total_el, barrier):
    """Same step through fcfv_set_mesh / fcfv_compute_local / fcfv_assemble / fcfv_residuals with
    pinned HOST buffers (what the Julia shim calls), H2D and D2H inside the timed region."""
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
    out = {"alpha": torch.empty(nel, dtype=torch.float64).pin_memory(), "beta": torch.empty(2 * nel, dtype=torch.float64).pin_memory(),
           "Z": torch.empty(4 * nel, dtype=torch.float64).pin_memory(), "tau": torch.empty(nf, dtype=torch.float64).pin_memory()}
    ptr = lambda t: C.c_void_p(t.data_ptr())
    h = hdev   # reuse the handle (its buffers are already sized); ownership/comm state is kept
    lib = h.lib
    form = api.formulation_id(FORM)
    norms = np.zeros(6)
    nnz = [C.c_int64(0) for _ in range(3)]
    sec = C.c_double(0.0)

    def step(with_mesh=True):
        if with_mesh:
            h.check(lib.fcfv_set_mesh(h._h, nel, nf, ptr(P["e2f"]), ptr(P["bc"]), ptr(P["Ω"]), ptr(P["n_x"]), ptr(P["n_y"]),
                                      ptr(P["Γ"]), ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel))
            if world > 1:
                h.check(lib.fcfv_set_ownership(h._h, ptr(P["eo"]), ptr(P["fo"]), nel_g, nf_g, None))
        else:   # what the shim does on a mesh it has already uploaded: refresh the mutable fields only
            h.check(lib.fcfv_update_fields(h._h, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, None))
        h.check(lib.fcfv_compute_local(h._h, form, 1.0, ptr(Dd["sex"]), ptr(Dd["sey"]), ptr(Dd["VxDir"]), ptr(Dd["VyDir"]),
                                       ptr(out["alpha"]), ptr(out["beta"]), ptr(out["Z"]), ptr(out["tau"])))
        h.check(lib.fcfv_assemble(h._h, VARIANT, form, 1.0, None, None, None, ptr(Dd["VxDir"]), ptr(Dd["VyDir"]), ptr(Dd["sxx"]),
                                  ptr(Dd["syy"]), ptr(Dd["sxy"]), ptr(Dd["syx"]), 0, C.byref(nnz[0]), C.byref(nnz[1]),
                                  C.byref(nnz[2]), C.byref(sec)))
        h.check(lib.fcfv_residuals(h._h, form, L.TAU_FACE, ptr(Dd["Vxh"]), ptr(Dd["Vyh"]), ptr(Dd["Pe"]), None, None, None,
                                   ptr(Dd["sex"]), ptr(Dd["sey"]), ptr(Dd["VxDir"]), ptr(Dd["VyDir"]), ptr(Dd["sxx"]),
                                   ptr(Dd["syy"]), ptr(Dd["sxy"]), ptr(Dd["syx"]), C.c_void_p(norms.ctypes.data), None, None, None))

    h.set_async(False)
    # set_mesh 128 nel + 8 nf | compute_local 16 nel + 16 nf | assemble 48 nf | residuals 24 nel + 64 nf
    h2d = 168 * nel + 136 * nf + ((nel + nf) if world > 1 else 0)
    d2h = 8 * (7 * nel + nf) + 72 + 48 + 4     # alpha/beta/Z/tau + nnz totals + norms + validation flag
    steps = max(1, min(args.steps, 3))
    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    h.sync()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    # same pass on a mesh that is already resident (second and later calls on one FCFV_Mesh object)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(with_mesh=False)
    h.sync()
    dt2 = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt2], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt2 = float(t.item())
    # the same pass through the single-call entry fcfv_pass (a driver that replaces its three calls by one): every array
    # uploaded once, alpha/beta/Z/tau downloaded on a second stream while the remaining inputs go up
    def fused(with_mesh=True):
        if with_mesh:
            h.check(lib.fcfv_set_mesh(h._h, nel, nf, ptr(P["e2f"]), ptr(P["bc"]), ptr(P["Ω"]), ptr(P["n_x"]), ptr(P["n_y"]),
                                      ptr(P["Γ"]), ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel))
            if world > 1:
                h.check(lib.fcfv_set_ownership(h._h, ptr(P["eo"]), ptr(P["fo"]), nel_g, nf_g, None))
        else:
            h.check(lib.fcfv_update_fields(h._h, ptr(P["ke"]), ptr(P["δ"]), ptr(P["τe"]), nel, None))
        h.check(lib.fcfv_pass(h._h, VARIANT, form, 1.0, L.TAU_FACE, ptr(Dd["sex"]), ptr(Dd["sey"]), ptr(Dd["VxDir"]), ptr(Dd["VyDir"]),
                              ptr(Dd["sxx"]), ptr(Dd["syy"]), ptr(Dd["sxy"]), ptr(Dd["syx"]), ptr(Dd["Vxh"]), ptr(Dd["Vyh"]), ptr(Dd["Pe"]),
                              ptr(out["alpha"]), ptr(out["beta"]), ptr(out["Z"]), ptr(out["tau"]), 0, C.byref(nnz[0]), C.byref(nnz[1]),
                              C.byref(nnz[2]), C.c_void_p(norms.ctypes.data)))
    fused_res = {}
    for tag, wm in (("with_mesh", True), ("mesh_resident", False)):
        fused(wm)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fused(wm)
        h.sync()
        dtf = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dtf], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtf = float(t.item())
        fused_res[tag] = {"value": total_el * steps / dtf / 1e6, "ms_per_step": 1e3 * dtf / steps}
    fused_h2d = 24 * nel + 64 * nf
    fused_res["with_mesh"]["h2d_bytes_per_step"] = int(fused_h2d + 128 * nel + 8 * nf + ((nel + nf) if world > 1 else 0))
    fused_res["mesh_resident"]["h2d_bytes_per_step"] = int(fused_h2d + 24 * nel)
    fused_res["note"] = ("fcfv_pass: one call instead of three, each array uploaded once, D2H of alpha/beta/Z/tau concurrent with the "
                         "uploads; not the headline because a driver has to change three lines to use it")
    return {"value": total_el * steps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "fused_pass": fused_res,
            "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "mesh_resident": {"value": total_el * steps / dt2 / 1e6, "ms_per_step": 1e3 * dt2 / steps,
                              "h2d_bytes_per_step": int(h2d - (128 * nel + 8 * nf) + 24 * nel - ((nel + nf) if world > 1 else 0)),
                              "note": "fcfv_update_fields instead of fcfv_set_mesh (mesh uploaded once per FCFV_Mesh object)"},
            "api": "fcfv_set_mesh + fcfv_compute_local + fcfv_assemble + fcfv_residuals, pinned host buffers, CSR left device-resident"}


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
