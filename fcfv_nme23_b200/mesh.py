"""Host-side mesh container and mesh input for the FCFV hot path.

Mirrors the reference's ``FCFV_Mesh`` struct (src/CreateMeshFCFV.jl:10-44) and the
input side of ``LoadExternalMesh`` (:223-252) / ``LoadExternalMesh2`` (:260-371) and
``meshes/ReadMeshesDat.jl`` (:19-113).  Only integer bookkeeping and file parsing
live here; every floating-point quantity of the path (areas, normals, face lengths)
is produced by the CUDA library (``api.mesh_geometry`` -> ``fcfv_geometry``).

Conventions are the reference's (Julia): connectivity is **1-based**, ``nel x 3``
matrices are stored column-major (``order='F'``) so that ``X[e, i]`` has the same
memory offset ``i*nel + e`` the Julia array has.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

__all__ = [
    "FCFV_Mesh", "LoadExternalMesh", "LoadExternalMesh2", "MakeStructuredMesh",
    "read_dat", "number_faces_first_seen", "structured_connectivity",
]


@dataclass
class FCFV_Mesh:
    """Same field names as the reference struct; ``None`` plays ``missing``."""
    type: Optional[str] = None
    nel: Optional[int] = None
    nf: Optional[int] = None
    nv: Optional[int] = None
    nn_el: Optional[int] = None
    nf_el: Optional[int] = None
    xv: Optional[np.ndarray] = None
    yv: Optional[np.ndarray] = None
    xf: Optional[np.ndarray] = None
    yf: Optional[np.ndarray] = None
    bc: Optional[np.ndarray] = None      # face tag: 1 Dirichlet, 2 Neumann, -1 interface, 0 interior
    xc: Optional[np.ndarray] = None
    yc: Optional[np.ndarray] = None
    e2v: Optional[np.ndarray] = None     # nel x 3, 1-based
    e2f: Optional[np.ndarray] = None     # nel x 3, 1-based
    Ω: Optional[np.ndarray] = None
    n_x: Optional[np.ndarray] = None
    n_y: Optional[np.ndarray] = None
    Γ: Optional[np.ndarray] = None
    e2e: Optional[np.ndarray] = None
    f2e: Optional[np.ndarray] = None
    Γ_f: Optional[np.ndarray] = None
    Ω_f: Optional[np.ndarray] = None
    n_x_f: Optional[np.ndarray] = None
    n_y_f: Optional[np.ndarray] = None
    τ: Optional[np.ndarray] = None       # face stabilisation (written by ComputeFCFV)
    τe: Optional[np.ndarray] = None      # element stabilisation
    ke: Optional[np.ndarray] = None      # element viscosity
    λ: Optional[np.ndarray] = None
    phase: Optional[np.ndarray] = None
    δ: Optional[np.ndarray] = None       # anisotropy factor
    N_x: Optional[np.ndarray] = None
    N_y: Optional[np.ndarray] = None
    # --- not in the reference: element/face ownership for a partitioned mesh (parallel.py)
    elem_owned: Optional[np.ndarray] = None   # uint8 mask, None = all owned
    face_owned: Optional[np.ndarray] = None   # uint8 mask, None = all owned
    elem_global: Optional[np.ndarray] = None  # int64 1-based global element id of each local element
    face_global: Optional[np.ndarray] = None  # int64 1-based global face id of each local face
    # local face numbering used by the geometry kernel: 0 = Triangle o2 ordering
    # (CreateMeshFCFV.jl:161-162), 1 = square_TRI ordering (:319-320)
    numbering: int = 1


def _f(a, dtype):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


# ----------------------------------------------------------------------------------
# .mat advancing-front meshes (LoadExternalMesh, CreateMeshFCFV.jl:223-252)
# ----------------------------------------------------------------------------------
_RES_FILES = {
    "LowRes": "MeshAdvancingFrontLowRes",
    "MedRes": "MeshAdvancingFrontMedRes",
    "HighRes": "MeshAdvancingFrontHighRes",
}


def _mesh_dir(mesh_dir):
    if mesh_dir is not None:
        return mesh_dir
    return os.environ.get("FCFV_MESH_DIR", "./meshes")


def LoadExternalMesh(res, η, mesh_dir=None) -> FCFV_Mesh:
    """Reads an advancing-front mesh that already carries Γ, Ω, n (CreateMeshFCFV.jl:223-252).

    ``res`` is ``"LowRes" | "MedRes" | "HighRes"`` (a leading ':' is accepted).  The file is
    looked up as ``<mesh_dir>/<name>.mat`` (MAT v5, via scipy.io) or ``<name>.npz``
    (same keys; the form the test fixtures are committed in).  Like the reference it sets
    neither ``τ``, ``τe`` nor ``λ`` (``ComputeFCFV`` allocates ``τ`` when missing).
    """
    res = str(res).lstrip(":")
    if res not in _RES_FILES:
        raise ValueError(f"unknown mesh resolution {res!r}")
    base = os.path.join(_mesh_dir(mesh_dir), _RES_FILES[res])
    if os.path.exists(base + ".npz"):
        data = dict(np.load(base + ".npz"))
    elif os.path.exists(base + ".mat"):
        from scipy.io import loadmat  # file I/O only
        data = {k: v for k, v in loadmat(base + ".mat").items() if not k.startswith("__")}
    else:
        raise FileNotFoundError(base + ".{mat,npz}")
    m = FCFV_Mesh(type="InHouseMesher")
    m.nel = int(np.asarray(data["nel"]).ravel()[0])
    m.nf = int(np.asarray(data["nf"]).ravel()[0])
    m.nn_el = int(np.asarray(data["nn_el"]).ravel()[0])
    m.nf_el = int(np.asarray(data["nf_el"]).ravel()[0])
    for dst, src in (("xv", "xn"), ("yv", "yn"), ("xf", "xf"), ("yf", "yf"), ("xc", "xc"), ("yc", "yc")):
        setattr(m, dst, np.ascontiguousarray(np.asarray(data[src], dtype=np.float64).ravel()))
    m.nv = m.xv.size
    m.e2f = _f(data["e2f"], np.int64)
    m.e2v = _f(data["e2n"], np.int64)
    m.Γ = _f(data["Gamma"], np.float64)
    m.Ω = np.ascontiguousarray(np.asarray(data["Omega"], dtype=np.float64).ravel())
    m.n_x = _f(data["nx"], np.float64)
    m.n_y = _f(data["ny"], np.float64)
    m.phase = np.asarray(data["phase"], dtype=np.float64).ravel()
    m.bc = np.ascontiguousarray(np.asarray(data["bc"], dtype=np.int64).ravel())
    η = np.asarray(η, dtype=np.float64).ravel()
    m.ke = η[m.phase.astype(np.int64) - 1].copy()      # η[Int64.(mesh.phase)]  :247
    m.δ = np.ones(m.nel)
    m.N_x = np.zeros(m.nel)
    m.numbering = 0
    return m


# ----------------------------------------------------------------------------------
# square_TRI_H*.dat  (meshes/ReadMeshesDat.jl:19-113) + LoadExternalMesh2
# ----------------------------------------------------------------------------------
def read_dat(path):
    """Parses a ``square_TRI_H*.dat`` file: header ``nsd nel nv nextfaces``, ``nv`` coordinate
    rows, ``nel`` connectivity rows (1-based), boundary rows (ReadMeshesDat.jl:19-30)."""
    if not os.path.exists(path) and os.path.exists(path + ".gz"):      # the larger fixtures are stored compressed
        path += ".gz"
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rt") as fh:
            tok = fh.read().split()
    else:
        with open(path) as fh:
            tok = fh.read().split()
    nsd, nel, nv, next_ = (int(t) for t in tok[:4])
    if nsd != 2:
        raise ValueError("only 2-D meshes are supported")
    p = 4
    xy = np.array(tok[p:p + 2 * nv], dtype=np.float64).reshape(nv, 2)
    p += 2 * nv
    e2v = np.array(tok[p:p + 3 * nel], dtype=np.int64).reshape(nel, 3)
    return xy[:, 0].copy(), xy[:, 1].copy(), _f(e2v, np.int64)


def number_faces_first_seen(xv, yv, e2v):
    """Global face numbering of ReadMeshesDat.jl:36-88: walk elements in order; local faces are
    the vertex pairs (v1,v3), (v3,v2), (v1,v2); a face is identified by its midpoint
    ``0.5*(xa + xb), 0.5*(ya + yb)`` compared with ``==``; new midpoints get the next id.
    Returns ``e2f`` (nel x 3, 1-based, order='F'), ``xf``, ``yf``."""
    e2v = np.asarray(e2v)
    nel = e2v.shape[0]
    a = np.stack([e2v[:, 0], e2v[:, 2], e2v[:, 0]], axis=1) - 1   # xv1 | xv2(=v3) | xv1
    b = np.stack([e2v[:, 2], e2v[:, 1], e2v[:, 1]], axis=1) - 1   # xv2(=v3) | xv3(=v2) | xv3(=v2)
    mx = 0.5 * (xv[a] + xv[b])
    my = 0.5 * (yv[a] + yv[b])
    seen = {}
    e2f = np.zeros((nel, 3), dtype=np.int64)
    xf, yf = [], []
    for e in range(nel):
        for i in range(3):
            key = (mx[e, i], my[e, i])
            f = seen.get(key)
            if f is None:
                xf.append(key[0]); yf.append(key[1])
                f = len(xf)
                seen[key] = f
            e2f[e, i] = f
    return _f(e2f, np.int64), np.array(xf), np.array(yf)


def _boundary_tags_unit_square(xf, yf):
    """ReadMeshesDat.jl:92-100: bc = 1 where the midpoint lies on the unit-square boundary."""
    return ((xf == 0.0) | (xf == 1.0) | (yf == 0.0) | (yf == 1.0)).astype(np.int64)


def _finish_square_mesh(xv, yv, e2v, e2f, xf, yf, bc) -> FCFV_Mesh:
    m = FCFV_Mesh(type="UnstructTriangles")
    m.nel = int(e2f.shape[0])
    m.e2v = _f(e2v, np.int64)
    m.nv = int(xv.size)
    m.e2f = _f(e2f, np.int64)
    m.nf = int(e2f.max())
    m.xv, m.yv, m.xf, m.yf = xv, yv, xf, yf
    m.bc = np.ascontiguousarray(bc, dtype=np.int64)
    m.phase = np.ones(m.nel)
    m.ke = np.ones(m.nel)
    m.τ = np.zeros(m.nf)
    m.τe = np.zeros(m.nel)
    m.nn_el = 3
    m.nf_el = 3
    m.δ = np.ones(m.nel)
    m.N_x = np.zeros(m.nel)
    m.λ = np.zeros(m.nel)
    m.numbering = 1
    return m


def LoadExternalMesh2(nx=None, ny=None, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0, τr=1.0, inclusion=0,
                      R=0.0, BC=(1, 1, 1, 1), area=None, no_pts_incl=None, *, name="square_TRI_H1",
                      mesh_dir=None, geometry=True) -> FCFV_Mesh:
    """Mirror of ``LoadExternalMesh2`` (CreateMeshFCFV.jl:260-371).  The reference reads a
    ``.mat`` produced by ReadMeshesDat.jl from ``<name>.dat``; here the ``.dat`` is read directly
    and the converter's numbering rule is applied, then Ω, centroids, normals and Γ are computed by
    the CUDA geometry kernel (``geometry=True``; needs a GPU).  Like the reference, every
    positional argument except ``τr`` is ignored."""
    xv, yv, e2v = read_dat(os.path.join(_mesh_dir(mesh_dir), name + ".dat"))
    e2f, xf, yf = number_faces_first_seen(xv, yv, e2v)
    m = _finish_square_mesh(xv, yv, e2v, e2f, xf, yf, _boundary_tags_unit_square(xf, yf))
    if geometry:
        from .api import mesh_geometry
        mesh_geometry(m, τr)
    return m


# ----------------------------------------------------------------------------------
# Synthetic criss-cross triangulation S(n) of the unit square.
# Same topology as square_TRI_H*: n x n quads, each cut into 4 triangles around its centre.
# ----------------------------------------------------------------------------------
def structured_connectivity(n, row0=0, row1=None):
    """Closed-form connectivity of S(n) for quad rows ``row0 <= qy < row1`` (global ids).

    Quads run row-major (qy outer, qx inner); the 4 triangles of a quad are consecutive in the
    order bottom, right, top, left, each stored (corner, centre, corner) counter-clockwise, so
    local face 1 is the quad edge, 2 = (v3,centre), 3 = (v1,centre) under the square_TRI local
    numbering (CreateMeshFCFV.jl:319-320).  Vertices: (n+1)^2 corners row-major, then n^2
    centres.  Global face ids follow the first-seen rule of ReadMeshesDat.jl, which for this
    element order has the closed form used here (and by the device generator):

        base(qx,qy) = 6*(qy*n+qx) + (qy==0 ? qx : n) + qy + (qx>0)
        new faces of a quad, in order: [bottom if qy==0] d00 d10 right d11 top d01 [left if qx==0]

    Returns 1-based ``e2v``, ``e2f`` (both (4*n*(row1-row0)) x 3, order='F').
    """
    if row1 is None:
        row1 = n
    qy, qx = np.meshgrid(np.arange(row0, row1, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="ij")
    qx = qx.ravel(); qy = qy.ravel()
    c00 = qy * (n + 1) + qx
    c10 = c00 + 1
    c01 = c00 + (n + 1)
    c11 = c01 + 1
    ctr = (n + 1) * (n + 1) + qy * n + qx

    def base(qx, qy):
        return 6 * (qy * n + qx) + np.where(qy == 0, qx, n) + qy + (qx > 0)

    b = base(qx, qy)
    o = (qy == 0).astype(np.int64)
    d00 = b + o
    d10 = b + o + 1
    right = b + o + 2
    d11 = b + o + 3
    top = b + o + 4
    d01 = b + o + 5
    qym = np.maximum(qy - 1, 0)
    bottom = np.where(qy == 0, b, base(qx, qym) + (qym == 0) + 4)
    qxm = np.maximum(qx - 1, 0)
    left = np.where(qx == 0, b + o + 6, base(qxm, qy) + o + 2)

    nq = qx.size
    e2v = np.empty((nq, 4, 3), dtype=np.int64)
    e2f = np.empty((nq, 4, 3), dtype=np.int64)
    # bottom (c10, m, c00); right (c11, m, c10); top (c01, m, c11); left (c00, m, c01)
    for t, (v1, v3, f1, f2, f3) in enumerate((
            (c10, c00, bottom, d00, d10),
            (c11, c10, right, d10, d11),
            (c01, c11, top, d11, d01),
            (c00, c01, left, d01, d00))):
        e2v[:, t, 0] = v1; e2v[:, t, 1] = ctr; e2v[:, t, 2] = v3
        e2f[:, t, 0] = f1; e2f[:, t, 1] = f2; e2f[:, t, 2] = f3
    return _f(e2v.reshape(-1, 3) + 1, np.int64), _f(e2f.reshape(-1, 3) + 1, np.int64)


def structured_vertices(n):
    """Vertex coordinates of S(n): corner (i,j) -> (i/n, j/n); centre -> ((2i+1)/(2n), (2j+1)/(2n))."""
    i = np.arange(n + 1, dtype=np.float64)
    cx = np.tile(i / n, n + 1)
    cy = np.repeat(i / n, n + 1)
    k = (2.0 * np.arange(n, dtype=np.float64) + 1.0) / (2.0 * n)
    mx = np.tile(k, n)
    my = np.repeat(k, n)
    return np.concatenate([cx, mx]), np.concatenate([cy, my])


def MakeStructuredMesh(n, τr=1.0, *, geometry=True) -> FCFV_Mesh:
    """Synthetic mesh S(n) (nel = 4n^2, nf = 6n^2+2n, nv = (n+1)^2+n^2), the stand-in for the
    Triangle-generated Delaunay meshes of ``MakeTriangleMesh`` (mesh generation by the
    third-party C mesher is outside the ported path).  Face midpoints are ``0.5*(xa+xb)`` as in
    ReadMeshesDat.jl:50-55; boundary faces are tagged Dirichlet."""
    e2v, e2f = structured_connectivity(n)
    xv, yv = structured_vertices(n)
    nf = 6 * n * n + 2 * n
    xf = np.empty(nf); yf = np.empty(nf)
    a = np.stack([e2v[:, 0], e2v[:, 2], e2v[:, 0]], axis=1) - 1
    b = np.stack([e2v[:, 2], e2v[:, 1], e2v[:, 1]], axis=1) - 1
    xf[e2f - 1] = 0.5 * (xv[a] + xv[b])
    yf[e2f - 1] = 0.5 * (yv[a] + yv[b])
    m = _finish_square_mesh(xv, yv, e2v, e2f, xf, yf, _boundary_tags_unit_square(xf, yf))
    if geometry:
        from .api import mesh_geometry
        mesh_geometry(m, τr)
    return m


def first_seen_face_permutation(e2f, nf):
    """Host statement of what ``fcfv_first_seen_faces`` computes on the device: the face numbering of
    meshes/ReadMeshesDat.jl:36-88 (a face gets the next id when the element loop first meets it), from ``e2f`` alone.
    Returns ``new_of_old`` (nf, 1-based): faces no element references keep the ids after all the others."""
    e2f = np.asarray(e2f, dtype=np.int64)
    nel = e2f.shape[0]
    code = (4 * np.arange(nel, dtype=np.int64)[:, None] + np.arange(3, dtype=np.int64)[None, :]).ravel()   # 4*e + i
    first = np.full(nf, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, e2f.ravel() - 1, code)
    used = first != np.iinfo(np.int64).max
    new_of_old = np.empty(nf, dtype=np.int64)
    order = np.argsort(first[used], kind="stable")
    ids = np.flatnonzero(used)
    new_of_old[ids[order]] = np.arange(1, ids.size + 1)
    new_of_old[~used] = ids.size + np.arange(1, (~used).sum() + 1)
    return new_of_old


def renumber_faces(mesh: FCFV_Mesh, new_of_old):
    """Applies a face renumbering to a mesh in place: ``e2f`` and every face array it carries (``bc, xf, yf, τ``)."""
    p = np.asarray(new_of_old, dtype=np.int64) - 1
    mesh.e2f = _f(p[np.asarray(mesh.e2f, dtype=np.int64) - 1] + 1, np.int64)
    for name in ("bc", "xf", "yf", "τ"):
        a = getattr(mesh, name, None)
        if a is not None and np.ndim(a) == 1 and len(a) == mesh.nf:
            out = np.empty_like(np.asarray(a))
            out[p] = np.asarray(a)
            setattr(mesh, name, out)
    return mesh


HILBERT_BITS = 20


def hilbert_index(x, y, bits=HILBERT_BITS):
    """Index of the integer cells (x, y) along the Hilbert curve of a 2^bits x 2^bits grid (vectorised)."""
    x, y = np.asarray(x, dtype=np.int64).copy(), np.asarray(y, dtype=np.int64).copy()
    d = np.zeros_like(x)
    s = 1 << (bits - 1)
    while s > 0:
        rx, ry = ((x & s) != 0).astype(np.int64), ((y & s) != 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        x &= s - 1
        y &= s - 1
        flip = (ry == 0) & (rx == 1)
        x, y = np.where(flip, s - 1 - x, x), np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d


def hilbert_element_order(xc, yc):
    """Host statement of what ``fcfv_hilbert_elements`` computes on the device: ``old_of_new`` (nel, 1-based), the
    elements sorted along a Hilbert curve through their centroids (2^20 cells per direction over the bounding box,
    elements of one cell keep their order)."""
    xc, yc = np.asarray(xc, dtype=np.float64).ravel(), np.asarray(yc, dtype=np.float64).ravel()
    if xc.size == 0:
        return np.empty(0, dtype=np.int64)
    cells, top = float(1 << HILBERT_BITS), (1 << HILBERT_BITS) - 1

    def quantise(v):
        w = v.max() - v.min()
        s = cells / w if w > 0.0 else 0.0
        return np.clip(((v - v.min()) * s).astype(np.int64), 0, top)
    key = hilbert_index(quantise(xc), quantise(yc))
    return np.argsort(key, kind="stable").astype(np.int64) + 1


def renumber_elements(mesh: FCFV_Mesh, old_of_new):
    """Applies an element order to a mesh in place: every element array it carries (rows of ``e2v, e2f, n_x, n_y, Γ, e2e``,
    ``xc, yc, Ω, τe, ke, λ, phase, δ, N_x, N_y``) and the element ids stored in ``f2e`` / ``e2e``.  Face ids do not
    change; follow with ``renumber_faces(mesh, first_seen_face_permutation(mesh.e2f, mesh.nf))``."""
    o = np.asarray(old_of_new, dtype=np.int64) - 1
    nel = mesh.nel
    assert o.shape == (nel,)
    new_of_old = np.empty(nel + 1, dtype=np.int64)          # 1-based lookup; ids <= 0 (no neighbour) stay
    new_of_old[0] = 0
    new_of_old[o + 1] = np.arange(1, nel + 1)
    for name in ("e2v", "e2f", "n_x", "n_y", "Γ", "e2e", "xc", "yc", "Ω", "τe", "ke", "λ", "phase", "δ", "N_x", "N_y",
                 "elem_owned", "elem_global"):
        a = getattr(mesh, name, None)
        if a is None:
            continue
        a = np.asarray(a)
        if a.shape[0] == nel:
            b = a[o]
        elif a.ndim == 1 and a.shape[0] > nel:               # τe sized nf in some drivers: the first nel entries are the elements'
            b = a.copy()
            b[:nel] = a[o]
        else:
            continue
        setattr(mesh, name, np.asfortranarray(b) if b.ndim == 2 else b)
    for name in ("f2e", "e2e"):
        a = getattr(mesh, name, None)
        if a is not None:
            a = np.asarray(a, dtype=np.int64)
            setattr(mesh, name, np.asfortranarray(np.where(a > 0, new_of_old[np.maximum(a, 0)], a)))
    return mesh
