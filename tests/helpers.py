"""Comparison helpers for the test suites (TEST INFRASTRUCTURE)."""
import numpy as np
import scipy.sparse as sp


def csc_triple_to_csr(csc, shape):
    """Julia-style CSC (1-based) -> (rowptr, col, val) CSR 0-based, columns ascending in a row."""
    colptr, rowval, nzval = csc
    A = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=shape).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data


def assert_csr_equal(got, ref, what, bitwise=True, rtol=1e-12):
    rp, ci, v = got
    rrp, rci, rv = ref
    assert np.array_equal(np.asarray(rp, dtype=np.int64), rrp), f"{what}: rowptr differs"
    assert np.array_equal(np.asarray(ci, dtype=np.int64), rci), f"{what}: colidx differs"
    if bitwise:
        assert np.array_equal(v, rv), f"{what}: values differ (max abs {np.abs(v - rv).max() if len(v) else 0})"
    else:
        scale = np.abs(rv).max() if len(rv) else 1.0
        assert np.abs(v - rv).max() <= rtol * scale, f"{what}: values differ beyond {rtol}"


def rel_inf(x, ref):
    """||x - ref||_inf / ||ref||_inf  -- the tolerance definition of SURVEY.md §8(c)."""
    ref = np.asarray(ref, dtype=np.float64)
    s = np.abs(ref).max() if ref.size else 0.0
    d = np.abs(np.asarray(x, dtype=np.float64) - ref).max() if ref.size else 0.0
    return d / s if s > 0 else d
