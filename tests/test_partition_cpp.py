"""The C++ partitioner behind fcfv_create_multi (csrc/fcfv_multi.inl, exported host-only as fcfv_partition_part)
against the NumPy partitioner of the one-process-per-GPU path (fcfv_nme23_b200/partition.py), which the gloo tests
prove to reproduce the global rows: same local element / face sets, same ownership, for every part."""
import numpy as np
import pytest

import cases
from fcfv_nme23_b200 import api
from fcfv_nme23_b200.partition import partition_mesh


def _meshes():
    out = [("LowRes", cases.load_mesh("LowRes")), ("H1", cases.load_mesh("H1")), ("S17", cases.load_mesh("S17"))]
    from test_random_meshes import random_case
    for seed in (3, 5):
        out.append((f"random{seed}", random_case(seed)[0]))
    return out


@pytest.mark.parametrize("name,mesh", _meshes(), ids=lambda x: x if isinstance(x, str) else "")
@pytest.mark.parametrize("nparts", [1, 2, 3, 8])
def test_cpp_partition_equals_numpy_partition(name, mesh, nparts):
    owned_e = np.zeros(mesh.nel, dtype=np.int64)
    owned_f = np.zeros(mesh.nf, dtype=np.int64)
    for part in range(nparts):
        P = partition_mesh(mesh, nparts, part)
        eg, fg, eo, fo = api.partition_part(mesh.e2f, mesh.nf, nparts, part)
        assert np.array_equal(eg, P.elem_global + 1) and np.array_equal(fg, P.face_global + 1)
        assert np.array_equal(eo.astype(bool), P.elem_owned) and np.array_equal(fo.astype(bool), P.face_owned)
        owned_e[eg[eo.astype(bool)] - 1] += 1
        owned_f[fg[fo.astype(bool)] - 1] += 1
    used = np.zeros(mesh.nf, bool)
    used[np.asarray(mesh.e2f).ravel() - 1] = True
    assert np.all(owned_e == 1) and np.all(owned_f[used] == 1) and np.all(owned_f[~used] == 0)    # ownership is a partition


def test_cpp_partition_threaded_adjacency_on_a_larger_mesh():
    """Above 65 536 elements the face adjacency is built by several threads (compare-and-swap min / max): S(200) has
    160 000 elements, and a shuffled element order makes the ghost sets large."""
    from fcfv_nme23_b200 import mesh as M
    m = M.MakeStructuredMesh(200, geometry=False)
    rng = np.random.default_rng(1)
    for shuffled in (False, True):
        if shuffled:
            m.e2f = np.asfortranarray(np.asarray(m.e2f)[rng.permutation(m.nel)])
        for nparts, part in ((4, 0), (4, 2), (7, 6)):
            P = partition_mesh(m, nparts, part)
            eg, fg, eo, fo = api.partition_part(m.e2f, m.nf, nparts, part)
            assert np.array_equal(eg, P.elem_global + 1) and np.array_equal(fg, P.face_global + 1)
            assert np.array_equal(eo.astype(bool), P.elem_owned) and np.array_equal(fo.astype(bool), P.face_owned)


def test_cpp_partition_rejects_bad_face_ids():
    from fcfv_nme23_b200 import lib as L
    e2f = np.asfortranarray(np.array([[1, 2, 9]], dtype=np.int64))
    with pytest.raises(L.FCFVError):
        api.partition_part(e2f, 3, 2, 0)
