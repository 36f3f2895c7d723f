"""Host-side mesh logic: .dat parsing + first-seen face numbering (meshes/ReadMeshesDat.jl), the
closed-form numbering of S(n) against that rule, the device generator's strip layout (integer
logic compiled for the host), and the element partitioner.  Runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from fcfv_nme23_b200 import mesh as M
from fcfv_nme23_b200.partition import partition_mesh, face_adjacency, element_ranges


@pytest.mark.parametrize("name,nel,nv,nf,nbnd", [("H1", 1024, 545, 1568, 64), ("H2", 4096, 2113, 6208, 128)])
def test_dat_meshes(name, nel, nv, nf, nbnd):
    xv, yv, e2v = M.read_dat(os.path.join(cases.GOLDEN, f"square_TRI_{name}.dat"))
    assert (e2v.shape[0], xv.size) == (nel, nv)
    e2f, xf, yf = M.number_faces_first_seen(xv, yv, e2v)
    assert e2f.max() == nf == xf.size                         # nf = 6n^2 + 2n
    assert np.array_equal(np.unique(e2f), np.arange(1, nf + 1))
    assert e2f[0, 0] == 1 and e2f[0, 1] == 2 and e2f[0, 2] == 3   # first element sees three new faces
    counts = np.bincount(e2f.ravel())[1:]
    assert counts.max() == 2 and (counts == 1).sum() == nbnd      # boundary faces have one element
    m = cases.load_mesh(name)
    assert (m.bc == 1).sum() == nbnd and abs(m.Ω.sum() - 1.0) < 1e-12
    # outward normals: n . (face midpoint - centroid) >= 0, unit length
    f = m.e2f - 1
    dot = m.n_x * (m.xf[f] - m.xc[:, None]) + m.n_y * (m.yf[f] - m.yc[:, None])
    assert dot.min() >= 0 and np.abs(np.hypot(m.n_x, m.n_y) - 1).max() < 1e-15
    # sum_i Gamma_i n_i = 0 per element (closed polygon)
    assert np.abs((m.Γ * m.n_x).sum(axis=1)).max() < 1e-15 and np.abs((m.Γ * m.n_y).sum(axis=1)).max() < 1e-15


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16])
def test_structured_closed_form_matches_first_seen_rule(n):
    e2v, e2f = M.structured_connectivity(n)
    xv, yv = M.structured_vertices(n)
    e2f_ref, xf, yf = M.number_faces_first_seen(xv, yv, e2v)
    assert np.array_equal(e2f, e2f_ref)
    assert e2f.max() == 6 * n * n + 2 * n and e2v.max() == (n + 1) ** 2 + n * n
    m = M.MakeStructuredMesh(n, geometry=False)
    assert np.array_equal(m.xf, xf) and np.array_equal(m.yf, yf)
    assert (m.bc == 1).sum() == 4 * n


@pytest.mark.parametrize("n,row0,row1", [(5, 0, 5), (6, 0, 2), (6, 2, 4), (6, 4, 6), (9, 3, 4), (4, 1, 3)])
def test_device_strip_layout(n, row0, row1, emu):
    """Integer logic of fcfv_generate.cuh (make_strip_layout / quad_faces / tri_faces), compiled for
    the host: local connectivity == global closed form restricted to the strip, local face ids
    ascending in the global ids."""
    sizes = (C.c_longlong * 4)()
    emu.emu_strip_connectivity(C.c_longlong(n), C.c_longlong(row0), C.c_longlong(row1), sizes, None)
    nel, nf, rlo, rhi = (int(x) for x in sizes)
    assert (rlo, rhi) == (max(row0 - 1, 0), min(row1 + 1, n)) and nel == 4 * n * (rhi - rlo)
    loc = np.zeros(3 * nel, dtype=np.int32)
    emu.emu_strip_connectivity(C.c_longlong(n), C.c_longlong(row0), C.c_longlong(row1), sizes, loc.ctypes.data_as(C.c_void_p))
    loc = loc.reshape((nel, 3), order="F")
    _, e2f = M.structured_connectivity(n, rlo, rhi)               # global ids, 1-based
    glob_faces = np.unique(e2f)
    assert glob_faces.size == nf
    assert np.array_equal(loc, np.searchsorted(glob_faces, e2f))  # monotone local numbering


@pytest.mark.parametrize("mesh_name,nparts", [("LowRes", 3), ("H1", 4), ("S6", 2)])
def test_partition_invariants(mesh_name, nparts):
    m = cases.load_mesh(mesh_name)
    lo, hi = face_adjacency(m.e2f, m.nf)
    owned_rows = []
    for r in range(nparts):
        P = partition_mesh(m, nparts, r)
        e0, e1 = element_ranges(m.nel, nparts)[r]
        assert np.array_equal(P.elem_global[P.elem_owned], np.arange(e0, e1))
        assert np.all(np.diff(P.elem_global) > 0) and np.all(np.diff(P.face_global) > 0)
        # local connectivity maps back to the global one
        assert np.array_equal(P.face_global[P.mesh.e2f - 1], (m.e2f - 1)[P.elem_global])
        # both neighbours of every owned face are local, and so is the tau donor of every column
        fo = P.face_global[P.face_owned]
        assert np.all(np.isin(hi[fo], P.elem_global)) and np.all(np.isin(lo[fo], P.elem_global))
        core = np.isin(P.elem_global, np.union1d(np.arange(e0, e1), lo[fo]))
        cols = np.unique((m.e2f - 1)[P.elem_global[core]])
        assert np.all(np.isin(hi[cols], P.elem_global))
        owned_rows.append(fo)
    allrows = np.concatenate(owned_rows)
    assert np.array_equal(np.sort(allrows), np.nonzero(hi >= 0)[0])   # every face owned exactly once


def test_first_seen_face_permutation():
    """Host statement of fcfv_first_seen_faces: identity on meshes that are numbered first-seen (S(n), the .dat
    meshes), recovers that numbering from any shuffle, is a permutation on the advancing-front meshes."""
    import cases
    from fcfv_nme23_b200 import mesh as M
    for name in ("S6", "H1"):
        m = cases.load_mesh(name)
        assert np.array_equal(M.first_seen_face_permutation(m.e2f, m.nf), np.arange(1, m.nf + 1)), name
        sh = np.random.default_rng(3).permutation(m.nf) + 1
        e2f_sh = sh[np.asarray(m.e2f) - 1]
        p = M.first_seen_face_permutation(e2f_sh, m.nf)
        assert np.array_equal(p[e2f_sh - 1], np.asarray(m.e2f)), name
    m = cases.load_mesh("LowRes")
    p = M.first_seen_face_permutation(m.e2f, m.nf)
    assert np.array_equal(np.sort(p), np.arange(1, m.nf + 1)) and (p != np.arange(1, m.nf + 1)).any()
    # renumbering a mesh keeps it the same mesh: face data follows the faces
    bc0, xf0 = m.bc.copy(), m.xf.copy()
    e2f0 = np.asarray(m.e2f).copy()
    M.renumber_faces(m, p)
    assert np.array_equal(m.bc[np.asarray(m.e2f) - 1], bc0[e2f0 - 1]) and np.array_equal(m.xf[np.asarray(m.e2f) - 1], xf0[e2f0 - 1])
    assert np.array_equal(M.first_seen_face_permutation(m.e2f, m.nf), np.arange(1, m.nf + 1))
    # faces nobody references go last, in their old order
    e2f = np.array([[1, 4, 6], [4, 2, 7]], dtype=np.int64)
    assert M.first_seen_face_permutation(e2f, 8).tolist() == [1, 4, 6, 2, 7, 3, 5, 8]


def test_hilbert_element_order():
    """Host statement of fcfv_hilbert_elements: the curve visits every cell once and moves to an edge-neighbour
    each step; the element order is a permutation that keeps consecutive elements close; renumbering a mesh with it
    keeps it the same mesh and the oracle assembles the same matrix up to the row / column permutation."""
    import cases
    from fcfv_nme23_b200 import mesh as M
    from oracle import oracle as O
    n = 16
    X, Y = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    d = M.hilbert_index(X.ravel(), Y.ravel(), bits=4)
    assert np.array_equal(np.sort(d), np.arange(n * n))
    o = np.argsort(d)
    step = np.abs(np.diff(X.ravel()[o])) + np.abs(np.diff(Y.ravel()[o]))
    assert (step == 1).all()
    # the 20-bit index restricted to a coarse grid visits the coarse cells in the 4-bit curve's order
    big = M.hilbert_index(X.ravel() << 16, Y.ravel() << 16)
    assert np.array_equal(np.argsort(big), o)

    m = cases.load_mesh("S17")
    rng = np.random.default_rng(11)
    sh = rng.permutation(m.nel) + 1                  # an element order without any coherence
    ms = M.renumber_elements(cases.load_mesh("S17"), sh)
    order = M.hilbert_element_order(ms.xc, ms.yc)
    assert np.array_equal(np.sort(order), np.arange(1, m.nel + 1))
    mh = M.renumber_elements(ms, order)
    jump = np.hypot(np.diff(mh.xc), np.diff(mh.yc))
    assert np.median(jump) < 1.5 / 17 and jump.max() < 6.0 / 17       # neighbours in memory are neighbours in the plane
    M.renumber_faces(mh, M.first_seen_face_permutation(mh.e2f, mh.nf))
    # same mesh: element k of mh is element sh[order[k]-1] of m, with the same geometry on the same faces
    old = sh[order - 1] - 1
    assert np.array_equal(mh.Ω, m.Ω[old]) and np.array_equal(mh.n_x, m.n_x[old]) and np.array_equal(mh.Γ, m.Γ[old])
    assert np.array_equal(mh.xf[np.asarray(mh.e2f) - 1], m.xf[np.asarray(m.e2f)[old] - 1])
    if mh.f2e is not None:
        f2e = np.asarray(mh.f2e)
        for e in range(0, mh.nel, 37):
            for i in range(3):
                assert e + 1 in f2e[mh.e2f[e, i] - 1]
    # degenerate boxes and the empty mesh
    assert M.hilbert_element_order(np.zeros(5), np.arange(5.0)).tolist() == [1, 2, 3, 4, 5]
    assert M.hilbert_element_order(np.zeros(3), np.zeros(3)).tolist() == [1, 2, 3]
    assert M.hilbert_element_order([], []).size == 0
