"""Element partitioning of an ``FCFV_Mesh`` for one-process-per-GPU runs.

The reference is single-process; this is the host-side bookkeeping (integers only) that lets each
rank assemble the rows of the faces it owns **without communication**:

* elements are split into contiguous ranges (rank r owns ``[e0, e1)``);
* a face is owned by the rank that owns its highest-numbered adjacent element -- the same element
  whose ``τe`` the reference's last-writer rule puts into ``mesh.τ`` (DiscretisationFCFV_Stokes.jl:40);
* ghost elements are replicated: (A) the lower neighbour of every owned face, whose α/β/Ζ and
  geometry the row needs, and (B) the highest neighbour of every face of an owned-or-A element, so
  that the local last-writer rule reproduces the global ``mesh.τ`` on every column the rows touch;
* local ids are ascending in the global ids, so "highest-numbered", the column order inside a
  row and the summation order of the two contributions of a row are those of the global mesh:
  every owned row is bit-identical to the row of an unpartitioned run.

Only the six residual sums of squares (and the two Powell-Hestenes norms) cross ranks (NCCL
all-reduce of 6 doubles inside ``fcfv_run_residuals``).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass

import numpy as np

from .mesh import FCFV_Mesh

__all__ = ["Partition", "partition_mesh", "element_ranges"]


def element_ranges(nel, nparts):
    """Contiguous, balanced element ranges ``[(e0, e1), ...]`` (0-based, half-open)."""
    return [(r * nel // nparts, (r + 1) * nel // nparts) for r in range(nparts)]


def face_adjacency(e2f, nf):
    """lo[f], hi[f]: lowest / highest adjacent element (0-based), -1 for unreferenced faces."""
    f = np.asarray(e2f, dtype=np.int64) - 1
    nel = f.shape[0]
    e = np.repeat(np.arange(nel, dtype=np.int64)[:, None], 3, axis=1)
    hi = np.full(nf, -1, dtype=np.int64)
    lo = np.full(nf, np.iinfo(np.int64).max, dtype=np.int64)
    np.maximum.at(hi, f.ravel(), e.ravel())
    np.minimum.at(lo, f.ravel(), e.ravel())
    lo[hi < 0] = -1
    return lo, hi


@dataclass
class Partition:
    """Local mesh of one rank plus the maps back to the global numbering (all 0-based arrays)."""
    mesh: FCFV_Mesh
    rank: int
    nparts: int
    elem_global: np.ndarray      # local element -> global element
    face_global: np.ndarray      # local face -> global face
    elem_owned: np.ndarray       # bool
    face_owned: np.ndarray       # bool
    nel_global: int
    nf_global: int

    def scatter_elem(self, x):
        """Restricts a global per-element array (leading dimension nel) to the local elements."""
        return np.asfortranarray(np.asarray(x)[self.elem_global])

    def scatter_face(self, x):
        return np.ascontiguousarray(np.asarray(x)[self.face_global])

    def owned_rows(self):
        """Global row ids (0-based, x block then y block) of the rows this rank produces."""
        g = self.face_global[self.face_owned]
        return np.concatenate([g, g + self.nf_global])

    def local_rows(self):
        l = np.nonzero(self.face_owned)[0]
        return np.concatenate([l, l + self.mesh.nf])

    def global_columns(self, cols, ncols_kind="uu"):
        """Maps local column ids of a local CSR to global ones (Kuu/Muu: face dofs, Kup: elements)."""
        cols = np.asarray(cols, dtype=np.int64)
        if ncols_kind == "up":
            return self.elem_global[cols]
        nf = self.mesh.nf
        return np.where(cols >= nf, self.face_global[np.minimum(cols - nf, nf - 1)] + self.nf_global,
                        self.face_global[np.minimum(cols, nf - 1)])


def partition_mesh(mesh: FCFV_Mesh, nparts: int, rank: int) -> Partition:
    """Rank ``rank``'s share of ``mesh``: a contiguous range of elements, the lower neighbours of the faces it owns
    and the τ donors of every face it touches (module docstring).  A face belongs to the rank of its highest-numbered
    adjacent element; faces no element references belong to no rank (their rows are empty in the global system too)."""
    nel, nf = int(mesh.nel), int(mesh.nf)
    e2f = np.asarray(mesh.e2f, dtype=np.int64) - 1
    lo, hi = face_adjacency(mesh.e2f, nf)
    e0, e1 = element_ranges(nel, nparts)[rank]
    owned_f = (hi >= e0) & (hi < e1)
    in_range = lambda e: (e >= e0) & (e < e1)
    # (A) lower neighbours of owned faces
    A = lo[owned_f]
    A = np.unique(A[~in_range(A)])
    core = np.union1d(np.arange(e0, e1, dtype=np.int64), A)
    # (B) tau donors: highest neighbour of every face of an owned / A element
    B = hi[np.unique(e2f[core].ravel())]
    B = np.setdiff1d(B, core)
    elems = np.union1d(core, B)                       # sorted -> local ids ascending in global ids
    faces = np.unique(e2f[elems].ravel())
    m = copy.copy(mesh)
    if hasattr(m, "_fcfv_handle"):
        del m._fcfv_handle
    m.nel, m.nf = int(elems.size), int(faces.size)
    m.e2f = np.asfortranarray(np.searchsorted(faces, e2f[elems]) + 1)
    for k in ("Ω", "ke", "δ", "phase", "xc", "yc", "λ", "N_x", "N_y"):
        v = getattr(mesh, k)
        if v is not None:
            setattr(m, k, np.ascontiguousarray(np.asarray(v)[elems]))
    for k in ("n_x", "n_y", "Γ", "e2v"):
        v = getattr(mesh, k)
        if v is not None:
            setattr(m, k, np.asfortranarray(np.asarray(v)[elems]))
    for k in ("bc", "xf", "yf", "τ"):
        v = getattr(mesh, k)
        if v is not None:
            setattr(m, k, np.ascontiguousarray(np.asarray(v)[faces]))
    if mesh.τe is not None:
        m.τe = np.ascontiguousarray(np.asarray(mesh.τe)[:nel][elems])
    m.elem_owned = in_range(elems).astype(np.uint8)
    m.face_owned = owned_f[faces].astype(np.uint8)
    m.elem_global = elems + 1
    m.face_global = faces + 1
    m.nel_global, m.nf_global = nel, nf
    # FCFV_TAU_REFERENCE reads the FACE array at the ELEMENT's global id (Residuals:17,54)
    m.tau_ref = None
    if mesh.τ is not None and nf >= nel:
        m.tau_ref = np.ascontiguousarray(np.asarray(mesh.τ)[elems])
    return Partition(mesh=m, rank=rank, nparts=nparts, elem_global=elems, face_global=faces,
                     elem_owned=m.elem_owned.astype(bool), face_owned=m.face_owned.astype(bool),
                     nel_global=nel, nf_global=nf)
