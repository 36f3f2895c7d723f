"""Host-side mirror of the reference's operator interface for the ported path.

Same function names, argument order and meaning as the exported Julia functions
(src/FCFV_NME23.jl:13,17):

    ComputeFCFV(mesh, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, τr, Formulation)
    ElementAssemblyLoop / ElementAssemblyLoopNEW(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation)
    ComputeResidualsFCFV_Stokes_o1(Vxh, Vyh, Pe, mesh, ae, be, ze, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, Formulation)

Every floating-point operation happens in ``libfcfv_b200.so`` on the GPU (``lib.py``); this module
only marshals arrays (NumPy, order='F' like Julia) and wraps results (SciPy CSC matrices with
int64 indices play ``SparseMatrixCSC{Float64,Int64}``).  It is what ``julia/FCFV_B200.jl`` does
with ``ccall``.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib as _L

__all__ = [
    "Handle", "handle_for", "mesh_geometry", "ComputeFCFV", "ElementAssemblyLoop", "ElementAssemblyLoopNEW",
    "ComputeResidualsFCFV_Stokes_o1", "FusedPass", "ph_residual", "ph_update_p", "ph_fusc", "formulation_id",
]

_FORMS = {"Gradient": _L.GRADIENT, "SymmetricGradient": _L.SYMMETRIC_GRADIENT}
_TAU_MODES = {"REFERENCE": _L.TAU_REFERENCE, "FACE": _L.TAU_FACE}
A_DEFAULT = 1.0   # module-level constant `A = 1.0` of src/DiscretisationFCFV_Stokes.jl:1-2


def formulation_id(Formulation):
    name = str(Formulation).lstrip(":")
    if name not in _FORMS:
        raise ValueError(f"Formulation must be :Gradient or :SymmetricGradient, got {Formulation!r}")
    return _FORMS[name]


class _Arg:
    """Keeps the converted array alive for the duration of a call and yields its address."""

    def __init__(self):
        self.keep = []
        self.temp = False         # some argument had to be converted / copied: its address means nothing after the call

    def done(self, h):
        """After the call: temporaries die here, so the library's upload cache must not remember their addresses."""
        if self.temp:
            h.new_epoch()

    def d(self, a, n=None):       # float64 input
        return self._conv(a, np.float64, n)

    def i(self, a, n=None):       # int64 input
        return self._conv(a, np.int64, n)

    def u8(self, a, n=None):
        return self._conv(a, np.uint8, n)

    def _conv(self, a, dtype, n):
        if a is None:
            return None
        if isinstance(a, int):
            return C.c_void_p(a)
        if hasattr(a, "data_ptr"):   # torch tensor (host or device); caller guarantees layout
            if n is not None and a.numel() < n:
                raise ValueError(f"array has {a.numel()} entries, need {n}")
            self.keep.append(a)
            return C.c_void_p(a.data_ptr())
        arr = np.asarray(a, dtype=dtype)
        if arr.ndim > 1:
            arr = np.asfortranarray(arr)
            flat = arr.reshape(-1, order="F")
        else:
            flat = np.ascontiguousarray(arr)
        if n is not None and flat.size < n:
            raise ValueError(f"array has {flat.size} entries, need {n}")
        if not (isinstance(a, np.ndarray) and np.shares_memory(flat, a)):
            self.temp = True
        self.keep.append(flat)
        return C.c_void_p(flat.ctypes.data)


def _out(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Handle:
    """One GPU (``Handle(device)``) or several behind one handle (``Handle(ngpus=N)``: devices 0..N-1 of this
    process, the mesh is partitioned inside ``fcfv_set_mesh``; arguments and results stay the global arrays),
    its stream(s) and all device-resident state (``fcfv_t``)."""

    def __init__(self, device=0, ngpus=1):
        self.lib = _L.load()
        self._h = C.c_void_p()
        self.ngpus = int(ngpus)
        if self.ngpus > 1:
            rc = self.lib.fcfv_create_multi(C.byref(self._h), self.ngpus)
        else:
            rc = self.lib.fcfv_create(C.byref(self._h), int(device))
        if rc != _L.FCFV_OK:
            raise _L.FCFVError(f"fcfv_create(device={device}, ngpus={ngpus}) failed with status {rc}: no usable CUDA device "
                               "(this library has no CPU fallback)")
        self.device = int(device)
        self.nel = self.nf = 0

    def partition_info(self):
        """Per part of a multi-GPU handle: (nel_local, nf_local, nel_owned, nf_owned); and the partitioner's host seconds."""
        out, sec = [], C.c_double(0.0)
        for p in range(self.lib.fcfv_num_devices(self._h)):
            v = [C.c_int64(0) for _ in range(4)]
            self.check(self.lib.fcfv_partition_info(self._h, p, *[C.byref(x) for x in v], C.byref(sec)))
            out.append(tuple(x.value for x in v))
        return out, sec.value

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.fcfv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != _L.FCFV_OK:
            raise _L.FCFVError(f"status {rc}: {self.lib.fcfv_last_error(self._h).decode()}")

    @property
    def kernel_launches(self):
        return int(self.lib.fcfv_kernel_launches(self._h))

    def numbering_locality(self):
        """(element_span, face_offset) of the resident mesh, see ``fcfv_numbering_locality``: both <= 0.01 on meshes
        numbered with locality, ~1/3 on randomly numbered ones (then see ``first_seen_faces`` / ``hilbert_elements``)."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        self.check(self.lib.fcfv_numbering_locality(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    @property
    def staged_copies(self):
        """Transfers from / to pageable host memory that went through the multi-threaded staging path."""
        return int(self.lib.fcfv_staged_copies(self._h))

    @property
    def pipelined_calls(self):
        """``fcfv_compute_fcfv`` calls that ran as an upload | kernel | download pipeline over element chunks."""
        return int(self.lib.fcfv_pipelined_calls(self._h))

    @property
    def device_bytes(self):
        return int(self.lib.fcfv_device_bytes(self._h))

    @property
    def stream(self):
        return self.lib.fcfv_stream(self._h)

    def set_timing(self, flag):
        self.check(self.lib.fcfv_set_timing(self._h, int(bool(flag))))

    def get_timers(self):
        """{kernel name: (total ms, launches)} accumulated since set_timing(True)."""
        n = self.lib.fcfv_num_timers()
        ms, calls = np.zeros(n), np.zeros(n, dtype=np.int64)
        self.check(self.lib.fcfv_get_timers(self._h, _out(ms), _out(calls)))
        return {self.lib.fcfv_timer_name(k).decode(): (float(ms[k]), int(calls[k])) for k in range(n)}

    def set_async(self, flag):
        self.check(self.lib.fcfv_set_async(self._h, int(bool(flag))))

    def sync(self):
        self.check(self.lib.fcfv_sync(self._h))

    def host_register(self, *arrays):
        """Page-locks NumPy arrays that are passed to the host-buffer entries repeatedly (``fcfv_host_register``)."""
        for a in arrays:
            if a is not None and a.nbytes:
                self.check(self.lib.fcfv_host_register(self._h, C.c_void_p(a.ctypes.data), a.nbytes))

    def host_unregister(self, *arrays):
        for a in arrays:
            if a is not None and a.nbytes:
                self.check(self.lib.fcfv_host_unregister(self._h, C.c_void_p(a.ctypes.data)))

    # -- mesh ---------------------------------------------------------------------------------
    def set_mesh(self, mesh, tau=True):
        """``tau=False``: the caller is ComputeFCFV, which overwrites mesh.τ on every face that has an element -- mesh.τ then
        only goes up when the mesh has faces without elements."""
        a = _Arg()
        nel, nf = int(mesh.nel), int(mesh.nf)
        taue = mesh.τe if mesh.τe is not None else np.zeros(nel)
        ntaue = int(np.size(taue)) if not hasattr(taue, "numel") else int(taue.numel())
        self.check(self.lib.fcfv_set_mesh(
            self._h, nel, nf, a.i(mesh.e2f, 3 * nel), a.i(mesh.bc, nf), a.d(mesh.Ω, nel), a.d(mesh.n_x, 3 * nel),
            a.d(mesh.n_y, 3 * nel), a.d(mesh.Γ, 3 * nel), a.d(mesh.ke, nel), a.d(mesh.δ, nel), a.d(taue, nel), ntaue))
        self.nel, self.nf = nel, nf
        flag = C.c_int(0)
        self.check(self.lib.fcfv_has_orphan_faces(self._h, C.byref(flag)))
        self.has_orphans = bool(flag.value)      # faces no element references keep the tau the caller gave them
        if mesh.τ is not None and (tau or self.has_orphans):
            self.update_fields(tau=mesh.τ)
        if getattr(mesh, "elem_owned", None) is not None or getattr(mesh, "face_owned", None) is not None:
            self.set_ownership(mesh.elem_owned, mesh.face_owned, getattr(mesh, "nel_global", 0),
                               getattr(mesh, "nf_global", 0), getattr(mesh, "tau_ref", None))

    def compute_fcfv(self, mesh, send_mesh, Formulation, A, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, out=None):
        """``fcfv_compute_fcfv``: ComputeFCFV(mesh, ...) as one call -- the mesh (``send_mesh``) or only its mutable fields
        ke, δ, τe go up in front of the element kernel; α, β, Ζ and mesh.τ come back.  ``out = (α, β, Ζ, τ)``: caller's arrays
        (column-major; page-locked ones -- ``host_register`` -- let the call run as a pipeline)."""
        a = _Arg()
        nel, nf = int(mesh.nel), int(mesh.nf)
        taue = mesh.τe if mesh.τe is not None else (np.zeros(nel) if send_mesh else None)
        ntaue = 0 if taue is None else (int(np.size(taue)) if not hasattr(taue, "numel") else int(taue.numel()))
        if out is None:
            al, be, Z = np.empty(nel), np.empty((nel, 2), order="F"), np.empty((nel, 2, 2), order="F")
            tau = np.empty(nf)
        else:
            al, be, Z, tau = out
            for x, n in ((al, nel), (be, 2 * nel), (Z, 4 * nel), (tau, nf)):
                if x.dtype != np.float64 or x.size != n or not (x.flags.f_contiguous or x.flags.c_contiguous and x.ndim == 1):
                    raise ValueError("out arrays must be float64, column-major and of sizes nel, (nel, 2), (nel, 2, 2), nf")
        geo = ([a.i(mesh.e2f, 3 * nel), a.i(mesh.bc, nf), a.d(mesh.Ω, nel), a.d(mesh.n_x, 3 * nel), a.d(mesh.n_y, 3 * nel),
                a.d(mesh.Γ, 3 * nel)] if send_mesh else [None] * 6)
        # the Neumann arrays the reference's ComputeFCFV receives and ignores go up while the results come down
        self.check(self.lib.fcfv_compute_fcfv(
            self._h, nel, nf, *geo, a.d(mesh.ke, nel), a.d(mesh.δ, nel), a.d(taue, nel), ntaue, a.d(mesh.τ, nf),
            formulation_id(Formulation), float(A), a.d(sex, nel), a.d(sey, nel), a.d(VxDir, nf), a.d(VyDir, nf), a.d(SxxNeu, nf),
            a.d(SyyNeu, nf), a.d(SxyNeu, nf), a.d(SyxNeu, nf), _out(al), _out(be), _out(Z), _out(tau)))
        a.done(self)
        if send_mesh:
            self.nel, self.nf = nel, nf
            flag = C.c_int(0)
            self.check(self.lib.fcfv_has_orphan_faces(self._h, C.byref(flag)))
            self.has_orphans = bool(flag.value)
        return al, be, Z, tau

    def update_fields(self, ke=None, delta=None, taue=None, tau=None):
        a = _Arg()
        nt = 0 if taue is None else int(np.size(taue))
        self.check(self.lib.fcfv_update_fields(self._h, a.d(ke, self.nel), a.d(delta, self.nel), a.d(taue, self.nel),
                                               nt, a.d(tau, self.nf)))
        a.done(self)

    # -- upload cache / transfer counters ---------------------------------------------------------
    def upload_cache(self, on=True):
        """See ``fcfv_upload_cache``: skip uploads of host arrays the device already mirrors (same address, size, epoch
        and sampled fingerprint)."""
        self.check(self.lib.fcfv_upload_cache(self._h, int(bool(on))))

    def new_epoch(self):
        self.check(self.lib.fcfv_new_epoch(self._h))

    def transfer_stats(self, reset=False):
        """(h2d_bytes, d2h_bytes, cache_hit_bytes) moved / saved by this handle so far."""
        v = [C.c_int64(0) for _ in range(3)]
        self.check(self.lib.fcfv_transfer_stats(self._h, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), int(bool(reset))))
        return tuple(x.value for x in v)

    def sparse_face_data(self, mode=1):
        """See ``fcfv_sparse_face_data``: 0 = whole per-face boundary arrays go up, 1 (default) = only their values on
        the Dirichlet / Neumann faces when those are few, 2 = always."""
        self.check(self.lib.fcfv_sparse_face_data(self._h, int(mode)))

    def sparse_face_info(self):
        """(active, n_dirichlet, n_neumann, list_uploads) of the resident mesh."""
        act = C.c_int(0)
        v = [C.c_int64(0) for _ in range(3)]
        self.check(self.lib.fcfv_sparse_face_info(self._h, C.byref(act), C.byref(v[0]), C.byref(v[1]), C.byref(v[2])))
        return bool(act.value), v[0].value, v[1].value, v[2].value

    def set_ownership(self, elem_owned, face_owned, nel_global=0, nf_global=0, tau_ref=None):
        a = _Arg()
        self.check(self.lib.fcfv_set_ownership(self._h, a.u8(elem_owned, self.nel), a.u8(face_owned, self.nf),
                                               int(nel_global or 0), int(nf_global or 0), a.d(tau_ref, self.nel)))

    def geometry(self, nel, nf, nv, e2v, e2f, xv, yv, xf, yf, numbering, tau_r, tau=None):
        a = _Arg()
        Om, xc, yc = np.empty(nel), np.empty(nel), np.empty(nel)
        nx, ny, G = (np.empty((nel, 3), order="F") for _ in range(3))
        tau = np.zeros(nf) if tau is None else np.array(tau, dtype=np.float64)
        self.check(self.lib.fcfv_geometry(self._h, nel, nf, nv, a.i(e2v, 3 * nel), a.i(e2f, 3 * nel), a.d(xv, nv), a.d(yv, nv),
                                          a.d(xf, nf), a.d(yf, nf), int(numbering), float(tau_r), _out(Om), _out(xc), _out(yc),
                                          _out(nx), _out(ny), _out(G), _out(tau)))
        return Om, xc, yc, nx, ny, G, tau

    # -- device-resident pipeline -------------------------------------------------------------
    def set_sources(self, sex, sey):
        a = _Arg()
        self.check(self.lib.fcfv_set_sources(self._h, a.d(sex, self.nel), a.d(sey, self.nel)))

    def set_face_data(self, VxDir=None, VyDir=None, sxx=None, syy=None, sxy=None, syx=None):
        a = _Arg()
        self.check(self.lib.fcfv_set_face_data(self._h, *[a.d(x, self.nf) for x in (VxDir, VyDir, sxx, syy, sxy, syx)]))

    def set_solution(self, Vxh, Vyh, Pe):
        a = _Arg()
        self.check(self.lib.fcfv_set_solution(self._h, a.d(Vxh, self.nf), a.d(Vyh, self.nf), a.d(Pe, self.nel)))

    def set_local(self, alpha, beta, Z):
        a = _Arg()
        self.check(self.lib.fcfv_set_local(self._h, a.d(alpha, self.nel), a.d(beta, 2 * self.nel), a.d(Z, 4 * self.nel)))

    def run_local(self, formulation, A=A_DEFAULT):
        self.check(self.lib.fcfv_run_local(self._h, formulation_id(formulation), float(A)))

    def run_assemble(self, variant, formulation, A=A_DEFAULT, want_muu=False):
        self.check(self.lib.fcfv_run_assemble(self._h, int(variant), formulation_id(formulation), float(A), int(bool(want_muu))))

    def run_pipeline(self, variant, formulation, tau_mode="FACE", A=A_DEFAULT):
        """run_local + run_assemble + run_residuals as one replayed CUDA graph (launch-bound small meshes)."""
        self.check(self.lib.fcfv_run_pipeline(self._h, int(variant), formulation_id(formulation), float(A), _TAU_MODES[tau_mode]))

    def run_residuals(self, formulation, tau_mode="REFERENCE", fetch=True):
        if not fetch:
            self.check(self.lib.fcfv_run_residuals(self._h, formulation_id(formulation), _TAU_MODES[tau_mode], None, None))
            return None
        ss, nrm = np.zeros(6), np.zeros(6)
        self.check(self.lib.fcfv_run_residuals(self._h, formulation_id(formulation), _TAU_MODES[tau_mode], _out(ss), _out(nrm)))
        return ss, nrm

    def get_local(self, tau=True):
        nel, nf = self.nel, self.nf
        al, be, Z = np.empty(nel), np.empty((nel, 2), order="F"), np.empty((nel, 2, 2), order="F")
        t = np.empty(nf) if tau else None
        self.check(self.lib.fcfv_get_local(self._h, _out(al), _out(be), _out(Z), _out(t)))
        return al, be, Z, t

    def get_nnz(self):
        v = [C.c_int64(0) for _ in range(3)]
        self.check(self.lib.fcfv_get_nnz(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # -- export -------------------------------------------------------------------------------
    def block_size(self, block):
        v = [C.c_int64(0) for _ in range(3)]
        self.check(self.lib.fcfv_block_size(self._h, block, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def export_csr(self, block):
        """Device CSR as (rowptr, colidx, val), 0-based."""
        nrows, ncols, nnz = self.block_size(block)
        w = self.lib.fcfv_index_width(self._h)
        rp = np.empty(nrows + 1, dtype=np.int64 if w == 8 else np.int32)
        ci, v = np.empty(nnz, dtype=np.int32), np.empty(nnz)
        self.check(self.lib.fcfv_export(self._h, block, _L.CSR0, _out(rp), _out(ci), _out(v)))
        return rp, ci, v

    def export_csc(self, block):
        """Julia SparseMatrixCSC fields (colptr, rowval, nzval), 1-based int64."""
        nrows, ncols, nnz = self.block_size(block)
        cp, rv, v = np.empty(ncols + 1, dtype=np.int64), np.empty(nnz, dtype=np.int64), np.empty(nnz)
        self.check(self.lib.fcfv_export(self._h, block, _L.CSC1_INT64, _out(cp), _out(rv), _out(v)))
        return cp, rv, v

    def export_scipy_csc(self, block):
        import scipy.sparse as sp
        nrows, ncols, _ = self.block_size(block)
        cp, rv, v = self.export_csc(block)
        return sp.csc_matrix((v, rv - 1, cp - 1), shape=(nrows, ncols))

    def export_rhs(self):
        fu, fp = np.empty(2 * self.nf), np.empty(self.nel)
        self.check(self.lib.fcfv_export_rhs(self._h, _out(fu), _out(fp)))
        return fu, fp

    # -- synthetic meshes ---------------------------------------------------------------------
    def generate_structured(self, n, row0=0, row1=None, setting="solkz", tau_r=10.0, eta_incl=1e-2, neumann_south=False):
        row1 = n if row1 is None else row1
        sid = {"solkz": 0, "inclusion": 1}[setting]
        self.check(self.lib.fcfv_generate_structured(self._h, int(n), int(row0), int(row1), 1, sid, float(tau_r),
                                                     float(eta_incl), int(bool(neumann_south))))
        nel, nf = C.c_int64(0), C.c_int64(0)
        self.check(self.lib.fcfv_mesh_size(self._h, C.byref(nel), C.byref(nf)))
        self.nel, self.nf = nel.value, nf.value

    def get_mesh(self):
        """Resident mesh as a dict of NumPy arrays in the reference's layout."""
        nel, nf = self.nel, self.nf
        o = dict(e2f=np.empty((nel, 3), dtype=np.int64, order="F"), bc=np.empty(nf, dtype=np.int64), Ω=np.empty(nel),
                 n_x=np.empty((nel, 3), order="F"), n_y=np.empty((nel, 3), order="F"), Γ=np.empty((nel, 3), order="F"),
                 ke=np.empty(nel), δ=np.empty(nel), τe=np.empty(nel), τ=np.empty(nf))
        self.check(self.lib.fcfv_get_mesh(self._h, *[_out(o[k]) for k in ("e2f", "bc", "Ω", "n_x", "n_y", "Γ", "ke", "δ", "τe", "τ")]))
        return o

    def get_ownership(self):
        """(elem_owned, face_owned, nel_global, nf_global) of the resident (possibly partitioned) mesh."""
        eo, fo = np.empty(self.nel, dtype=np.uint8), np.empty(self.nf, dtype=np.uint8)
        ng, fg = C.c_int64(0), C.c_int64(0)
        self.check(self.lib.fcfv_get_ownership(self._h, _out(eo), _out(fo), C.byref(ng), C.byref(fg)))
        return eo, fo, ng.value, fg.value

    def get_data(self):
        nel, nf = self.nel, self.nf
        names = ("sex", "sey", "VxDir", "VyDir", "sxx", "syy", "sxy", "syx", "Vxh", "Vyh", "Pe")
        o = {k: np.empty(nel if k in ("sex", "sey", "Pe") else nf) for k in names}
        self.check(self.lib.fcfv_get_data(self._h, *[_out(o[k]) for k in names]))
        return o

    # -- communicator -------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        self.check(self.lib.fcfv_comm_unique_id(self._h, buf))
        return bytes(buf)

    def comm_init(self, id128, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(id128)
        self.check(self.lib.fcfv_comm_init(self._h, buf, int(rank), int(nranks)))


# ---------------------------------------------------------------------------------------------
# handle cache: one handle per mesh object, created on first use
# ---------------------------------------------------------------------------------------------
# GPUs a mesh is spread over when its handle is created (the reference's functions have no such argument):
# ``mesh.ngpus`` if set, else this default (environment FCFV_NGPUS, use_gpus(n)).
# The reference's drivers pass the same arrays to ComputeFCFV, ElementAssemblyLoop[NEW] and the residual function; with
# this switch on (default) a handle made by handle_for() skips uploads of arrays the GPU already mirrors.  A new epoch
# starts at every ComputeFCFV; arrays modified IN PLACE between ComputeFCFV and the calls that follow it need
# ``handle_for(mesh).new_epoch()`` (no driver of the reference does that).
CACHE_INPUTS = True
DEFAULT_GPUS = int(os.environ.get("FCFV_NGPUS", "1"))


def use_gpus(n):
    global DEFAULT_GPUS
    DEFAULT_GPUS = int(n)


def partition_part(e2f, nf, nparts, part):
    """Host-only: (elem_global, face_global, elem_owned, face_owned) of part `part` of `nparts` -- the partition a
    multi-GPU handle uses (1-based ascending global ids = the part's local numbering).  Needs no GPU."""
    lib = _L.load()
    e2f = np.asfortranarray(np.asarray(e2f, dtype=np.int64))
    nel = int(e2f.shape[0])
    ne, nfl = C.c_int64(0), C.c_int64(0)
    args = (nel, int(nf), C.c_void_p(e2f.ctypes.data), int(nparts), int(part))
    rc = lib.fcfv_partition_part(*args, C.byref(ne), C.byref(nfl), None, None, None, None)
    if rc != _L.FCFV_OK:
        raise _L.FCFVError(f"fcfv_partition_part failed with status {rc}")
    eg, fg = np.empty(ne.value, dtype=np.int64), np.empty(nfl.value, dtype=np.int64)
    eo, fo = np.empty(ne.value, dtype=np.uint8), np.empty(nfl.value, dtype=np.uint8)
    rc = lib.fcfv_partition_part(*args, C.byref(ne), C.byref(nfl), _out(eg), _out(fg), _out(eo), _out(fo))
    if rc != _L.FCFV_OK:
        raise _L.FCFVError(f"fcfv_partition_part failed with status {rc}")
    return eg, fg, eo, fo


def _new_handle(mesh, device=None) -> Handle:
    ngpus = int(getattr(mesh, "ngpus", None) or DEFAULT_GPUS)
    h = Handle(0 if device is None else device, ngpus=ngpus)
    h.upload_cache(CACHE_INPUTS)
    return h


def _warn_numbering(mesh, h):
    if h.ngpus == 1 and h.nel >= 100_000 and getattr(mesh, "elem_owned", None) is None:
        span, off = h.numbering_locality()
        if span > 0.05 or off > 0.05:
            import warnings
            warnings.warn(f"mesh numbering has little locality (element span {span:.2f}, face offset {off:.2f} of the mesh): "
                          "the GPU passes run 3-7x faster after mesh.renumber_elements(hilbert_elements) / "
                          "mesh.renumber_faces(first_seen_faces)", stacklevel=4)


def handle_for(mesh, device=None, tau=True) -> Handle:
    h = getattr(mesh, "_fcfv_handle", None)
    if h is None:
        h = _new_handle(mesh, device)
        h.set_mesh(mesh, tau=tau)
        mesh._fcfv_handle = h
        _warn_numbering(mesh, h)
    return h


def _refresh(mesh, h, tau=True):
    """The drivers mutate ke, δ, τe, τ after mesh creation (ViscousInclusionFCFV_v2.jl:44,127;
    SolKzFCFV.jl:25,149): re-upload those vectors at the top of every entry point.  ``tau=False``: in front of
    ComputeFCFV, which overwrites mesh.τ on every face that has an element (Discretisation:40)."""
    taue = mesh.τe if mesh.τe is not None else None
    h.update_fields(ke=mesh.ke, delta=mesh.δ, taue=taue, tau=mesh.τ if tau else None)


def mesh_geometry(mesh, τr, device=0):
    """Fills Ω, xc, yc, n_x, n_y, Γ and τ (= τr on every face) -- the geometry loops of
    MakeTriangleMesh / LoadExternalMesh2 (CreateMeshFCFV.jl:142-156,180-211 / :300-314,337-367)."""
    h = Handle(device)
    try:
        mesh.Ω, mesh.xc, mesh.yc, mesh.n_x, mesh.n_y, mesh.Γ, mesh.τ = h.geometry(
            mesh.nel, mesh.nf, mesh.nv, mesh.e2v, mesh.e2f, mesh.xv, mesh.yv, mesh.xf, mesh.yf, mesh.numbering, τr,
            tau=mesh.τ)
    finally:
        h.close()
    return mesh


def ComputeFCFV(mesh, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, τr, Formulation, A=A_DEFAULT, out=None):
    """src/DiscretisationFCFV_Stokes.jl:8-44.  Returns α (nel), β (nel x 2), Ζ (nel x 2 x 2) and
    overwrites ``mesh.τ`` (allocated when missing, as after ``LoadExternalMesh``).  The Neumann
    arrays and τr are accepted and unused, as in the reference.  ``out = (α, β, Ζ, τ)`` (optional, not in the reference):
    write into the caller's arrays instead of fresh ones -- with page-locked inputs and outputs (``Handle.host_register``) a large
    mesh then runs through the library's upload | kernel | download pipeline."""
    if mesh.τ is None:
        mesh.τ = np.zeros(mesh.nf)
    h = getattr(mesh, "_fcfv_handle", None)
    partitioned = getattr(mesh, "elem_owned", None) is not None or getattr(mesh, "face_owned", None) is not None
    send_mesh = h is None and not partitioned
    if h is None and partitioned:      # ownership masks go up between the mesh and the first call (fcfv_set_ownership)
        h = handle_for(mesh, tau=False)
    elif h is None:
        h = _new_handle(mesh)
    else:                    # first call of a driver pass: whatever the caller changed since the last one goes up again
        h.new_epoch()
    al, be, Z, tau = h.compute_fcfv(mesh, send_mesh, Formulation, A, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, out=out)
    if send_mesh:
        mesh._fcfv_handle = h
        _warn_numbering(mesh, h)
    mesh.τ = tau
    return al, be, Z


def _assemble(variant, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A, want_muu, export):
    h = handle_for(mesh)
    _refresh(mesh, h)
    nel, nf = h.nel, h.nf
    a = _Arg()
    nnz = [C.c_int64(0) for _ in range(3)]
    sec = C.c_double(0.0)
    h.check(h.lib.fcfv_assemble(h._h, variant, formulation_id(Formulation), float(A), a.d(α, nel), a.d(β, 2 * nel),
                                a.d(Ζ, 4 * nel), a.d(VxDir, nf), a.d(VyDir, nf), a.d(σxxNeu, nf), a.d(σyyNeu, nf),
                                a.d(σxyNeu, nf), a.d(σyxNeu, nf), int(bool(want_muu)), C.byref(nnz[0]), C.byref(nnz[1]),
                                C.byref(nnz[2]), C.byref(sec)))
    a.done(h)
    if not export:   # result stays on the device (fcfv_device_csr / Handle.export_*)
        return h, tuple(x.value for x in nnz), sec.value
    Kuu = h.export_scipy_csc(_L.KUU)
    Kup = h.export_scipy_csc(_L.KUP)
    Muu = h.export_scipy_csc(_L.MUU) if want_muu else None
    fu, fp = h.export_rhs()
    return Kuu, Muu, Kup, fu.reshape(-1, 1), fp, sec.value


def ElementAssemblyLoop(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A=A_DEFAULT,
                        want_muu=True, export=True):
    """src/DiscretisationFCFV_Stokes.jl:102-202 (+ Sparsify :355-364).  Returns
    ``Kuu, Muu, Kup`` (CSC, int64 indices), ``fu`` (2nf x 1), ``fp`` (nel), ``tsparse`` (device
    seconds of the CSR build).  ``export=False`` keeps everything on the GPU and returns
    ``(handle, (nnz_uu, nnz_up, nnz_muu), seconds)``."""
    return _assemble(_L.LOOP_OLD, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A, want_muu, export)


def ElementAssemblyLoopNEW(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A=A_DEFAULT,
                           want_muu=True, export=True):
    """src/DiscretisationFCFV_Stokes.jl:210-347 (+ Sparsify)."""
    return _assemble(_L.LOOP_NEW, mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation, A, want_muu, export)


def FusedPass(mesh, sex, sey, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Vxh, Vyh, Pe, Formulation, loop="old",
              tau_mode="REFERENCE", A=A_DEFAULT, want_muu=False):
    """``fcfv_pass``: ComputeFCFV + ElementAssemblyLoop[NEW] + ComputeResidualsFCFV_Stokes_o1 in one call, every host
    array uploaded once and the download of α, β, Ζ, τ overlapped with the uploads.  Returns ``(α, β, Ζ, norms, nnz)``;
    ``mesh.τ`` is set; the matrices stay on the GPU (``handle_for(mesh).export_scipy_csc(...)``, ``export_rhs()``)."""
    h = handle_for(mesh)
    _refresh(mesh, h)
    nel, nf = h.nel, h.nf
    a = _Arg()
    α, β, Ζ = np.empty(nel), np.empty((nel, 2), order="F"), np.empty((nel, 2, 2), order="F")
    τ, norms = np.empty(nf), np.zeros(6)
    nnz = [C.c_int64(0) for _ in range(3)]
    variant = _L.LOOP_NEW if str(loop).upper() == "NEW" else _L.LOOP_OLD
    h.check(h.lib.fcfv_pass(h._h, variant, formulation_id(Formulation), float(A), _TAU_MODES[tau_mode], a.d(sex, nel), a.d(sey, nel),
                            a.d(VxDir, nf), a.d(VyDir, nf), a.d(σxxNeu, nf), a.d(σyyNeu, nf), a.d(σxyNeu, nf), a.d(σyxNeu, nf),
                            a.d(Vxh, nf), a.d(Vyh, nf), a.d(Pe, nel), _out(α), _out(β), _out(Ζ), _out(τ), int(bool(want_muu)),
                            C.byref(nnz[0]), C.byref(nnz[1]), C.byref(nnz[2]), _out(norms)))
    mesh.τ = τ
    return α, β, Ζ, norms, tuple(int(v.value) for v in nnz)


_RES_LINES = (
    "Residual of local  equation 20a: %2.2e",
    "Residual of local  equation 20b: %2.2e",
    "Residual of local  equation 20c: %2.2e",
    "Residual of global equation 19a: %2.2e",
    "Residual of global equation 19a: %2.2e",
    "Residual of global equation 19b: %2.2e",
)


def ComputeResidualsFCFV_Stokes_o1(Vxh, Vyh, Pe, mesh, ae, be, ze, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu,
                                   SyxNeu, Formulation, tau_mode="REFERENCE", verbose=True, arrays=False):
    """src/ComputeResidualsFCFV_Stokes.jl:5-108.  Prints the same six lines (:101-106); the
    reference returns ``nothing``, this returns the six RMS norms (and, with ``arrays=True``,
    F_nodes_x, F_nodes_y, F_glob2).  ``tau_mode='REFERENCE'`` reads ``mesh.τ[e]`` like the
    reference (:17,54); ``'FACE'`` uses the face value ``mesh.τ[e2f[e,i]]``."""
    h = handle_for(mesh)
    _refresh(mesh, h)
    nel, nf = h.nel, h.nf
    a = _Arg()
    norms = np.zeros(6)
    Fx = np.empty(nf) if arrays else None
    Fy = np.empty(nf) if arrays else None
    Fg2 = np.empty(nel) if arrays else None
    h.check(h.lib.fcfv_residuals(h._h, formulation_id(Formulation), _TAU_MODES[tau_mode], a.d(Vxh, nf), a.d(Vyh, nf),
                                 a.d(Pe, nel), a.d(ae, nel), a.d(be, 2 * nel), a.d(ze, 4 * nel), a.d(sex, nel), a.d(sey, nel),
                                 a.d(VxDir, nf), a.d(VyDir, nf), a.d(SxxNeu, nf), a.d(SyyNeu, nf), a.d(SxyNeu, nf),
                                 a.d(SyxNeu, nf), _out(norms), _out(Fx), _out(Fy), _out(Fg2)))
    a.done(h)
    if verbose:
        for line, v in zip(_RES_LINES, norms):
            print(line % v)
    if arrays:
        return norms, dict(F_nodes_x=Fx, F_nodes_y=Fy, F_glob2=Fg2)
    return norms


def SchurComplement(mesh, penalty):
    """Kuusc = Kuu .- Kup*(Kppi*Kpu), Kppi = spdiagm(penalty.*ke./Ω), Kpu = -Kup' (src/SolversFCFV_Stokes.jl:6,23-25),
    produced by the assembly kernels themselves (one extra term per adjacent element) instead of a host sparse
    product.  Call after ElementAssemblyLoop[NEW] on the same mesh; returns a scipy CSC matrix with the fields of the
    SparseMatrixCSC Julia would build."""
    h = handle_for(mesh)
    _refresh(mesh, h)
    nnz = C.c_int64(0)
    h.check(h.lib.fcfv_schur(h._h, float(penalty), C.byref(nnz)))
    return h.export_scipy_csc(_L.KUUSC)


def ComputeElementValues(mesh, Vxh, Vyh, Pe, α, β, Ζ, VxDir, VyDir, Formulation):
    """src/DiscretisationFCFV_Stokes.jl:52-94 -> (Vxe, Vye, τxxe, τyye, τxye).  ``Pe``, ``VxDir``, ``VyDir`` and
    ``Formulation`` are accepted for signature parity; the reference does not read them."""
    h = handle_for(mesh)
    _refresh(mesh, h)
    nel, nf = h.nel, h.nf
    a = _Arg()
    out = [np.empty(nel) for _ in range(5)]
    h.check(h.lib.fcfv_element_values(h._h, a.d(Vxh, nf), a.d(Vyh, nf), a.d(α, nel), a.d(β, 2 * nel), a.d(Ζ, 4 * nel),
                                      *[_out(o) for o in out]))
    return tuple(out)


# ---------------------------------------------------------------------------------------------
# Powell-Hestenes lines of StokesSolvers! that consume the device CSR (SolversFCFV_Stokes.jl:40-49)
# ---------------------------------------------------------------------------------------------
def ph_residual(mesh, u, p):
    """ru = fu - Kuu*u - Kup*p, rp = fp - Kpu*u, (norm(ru), norm(rp))."""
    h = handle_for(mesh)
    a = _Arg()
    ru, rp, nrm = np.empty(2 * h.nf), np.empty(h.nel), np.zeros(2)
    h.check(h.lib.fcfv_ph_residual(h._h, a.d(u, 2 * h.nf), a.d(p, h.nel), _out(ru), _out(rp), _out(nrm)))
    return ru, rp, nrm


def ph_update_p(mesh, penalty, u, p):
    """p .+= (penalty.*ke./Ω) .* (fp .- Kpu*u); returns the new p."""
    h = handle_for(mesh)
    a = _Arg()
    pn = np.array(p, dtype=np.float64).ravel().copy()
    h.check(h.lib.fcfv_ph_update_p(h._h, float(penalty), a.d(u, 2 * h.nf), _out(pn)))
    return pn


def ph_fusc(mesh, penalty, p):
    """fusc = fu .- Kup*(Kppi*fp .+ p)."""
    h = handle_for(mesh)
    a = _Arg()
    out = np.empty(2 * h.nf)
    h.check(h.lib.fcfv_ph_fusc(h._h, float(penalty), a.d(p, h.nel), _out(out)))
    return out


def center_pressure(mesh, Pe):
    """Pe .- mean(Pe) (src/SolversFCFV_Stokes.jl:20); returns (Pe_centered, mean)."""
    h = handle_for(mesh)
    out = np.array(Pe, dtype=np.float64).ravel().copy()
    mean = C.c_double(0.0)
    h.check(h.lib.fcfv_center_pressure(h._h, _out(out), C.byref(mean)))
    return out, float(mean.value)


def first_seen_faces(e2f, nf, device=0):
    """``fcfv_first_seen_faces``: (new_of_old, e2f_new) of the reference's first-seen face numbering, computed on the
    GPU from ``e2f`` (nel x 3, 1-based).  See ``mesh.renumber_faces`` for applying it to a mesh."""
    e2f = np.asarray(e2f, dtype=np.int64)
    nel = e2f.shape[0]
    h = Handle(device)
    a = _Arg()
    perm = np.empty(nf, dtype=np.int64)
    out = np.empty((nel, 3), dtype=np.int64, order="F")
    try:
        h.check(h.lib.fcfv_first_seen_faces(h._h, nel, int(nf), a.i(e2f, 3 * nel), _out(perm), _out(out)))
    finally:
        h.close()
    return perm, out


def hilbert_elements(xc, yc, device=0):
    """``fcfv_hilbert_elements``: ``old_of_new`` (nel, 1-based) of the elements sorted along a Hilbert curve through
    their centroids, computed on the GPU.  ``mesh.renumber_elements`` applies it to a mesh."""
    xc, yc = np.ascontiguousarray(xc, dtype=np.float64).ravel(), np.ascontiguousarray(yc, dtype=np.float64).ravel()
    nel = xc.size
    h = Handle(device)
    a = _Arg()
    out = np.empty(nel, dtype=np.int64)
    try:
        h.check(h.lib.fcfv_hilbert_elements(h._h, nel, a.d(xc, nel), a.d(yc, nel), _out(out)))
    finally:
        h.close()
    return out
