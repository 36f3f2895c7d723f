"""The oracle against the REFERENCE'S OWN SOURCE TEXT, executed through the rule-table transliterator
of tests/julia_transliteration.py (Julia itself is not installed).  Needs the reference checkout
(/root/reference or $FCFV_REFERENCE); skipped elsewhere -- there the committed outputs of the same
run (tests/golden/golden_reference.json, tests/test_golden_reference.py) stand in."""
import numpy as np
import pytest

import cases
import julia_transliteration as T
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not T.available(), reason="reference checkout not present")

SMALL = cases.all_cases(small_only=True)


def _oracle(case, m, d):
    a, b, Z = O.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], case.tau_r, case.form)
    fn = O.ElementAssemblyLoop if case.variant == 0 else O.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *d["neu"], case.form)
    return a, b, Z, Kuu, Muu, Kup, fu, fp


@pytest.mark.parametrize("case", SMALL, ids=lambda c: c.name)
def test_oracle_equals_transliterated_reference(case):
    """ComputeFCFV, ElementAssemblyLoop[NEW] + Sparsify: every output bit for bit; the six residual norms to 1e-12;
    ComputeElementValues bit for bit."""
    import translit_run
    m, d = cases.build(case)
    r = translit_run.run_case(case, *cases.build(case))
    a, b, Z, Kuu, Muu, Kup, fu, fp = _oracle(case, m, d)
    assert np.array_equal(r["a"], a) and np.array_equal(r["b"], b) and np.array_equal(r["Z"], Z)
    assert np.array_equal(r["tau"], m.τ)                                   # last-writer rule (Discretisation:40)
    for name, got, ref in (("Kuu", r["Kuu"], Kuu), ("Muu", r["Muu"], Muu), ("Kup", r["Kup"], Kup)):
        for part, x, y in zip(("colptr", "rowval", "nzval"), got, ref):
            assert np.array_equal(x, y), f"{name}.{part} differs from the transliterated reference"
    assert np.array_equal(r["fu"], np.asarray(fu).ravel()) and np.array_equal(r["fp"], fp)
    # residuals: the reference's literal tau[e] mode; Julia goes through BLAS for the small products, so 1e-12.
    # Norms of quantities that vanish analytically (local equations: pure round-off) are compared on the scale of
    # the cancelling terms = the largest of the six norms.
    on = O.ComputeResidualsFCFV_Stokes_o1(d["Vxh"], d["Vyh"], d["Pe"], m, a, b, Z, d["sex"], d["sey"], d["VxDir"], d["VyDir"],
                                          *d["neu"], case.form, tau_mode="REFERENCE")
    scale = max(on.max(), r["norms"].max())
    big = on > 1e-9 * scale
    assert np.all(np.abs(r["norms"][big] - on[big]) <= 1e-12 * on[big])
    assert np.all(np.abs(r["norms"][~big] - on[~big]) <= 1e-12 * scale)
    ev = O.ComputeElementValues(m, d["Vxh"], d["Vyh"], d["Pe"], a, b, Z, d["VxDir"], d["VyDir"], case.form)
    for x, y in zip(r["ev"], ev):
        assert np.array_equal(x, y)


def test_print_formats_are_the_references():
    import translit_run
    case = SMALL[0]
    r = translit_run.run_case(case, *cases.build(case), elem_values=False)
    assert r["formats"][0] == "Residual of local  equation 20a: %2.2e\n"
    assert r["formats"][3] == r["formats"][4] == "Residual of global equation 19a: %2.2e\n"   # (sic) both say 19a


@pytest.mark.parametrize("mesh", ["H1", "S6"])
def test_geometry_loops(mesh):
    """Heron areas, centroids, normals, face lengths and the tau initialisation of LoadExternalMesh2
    (CreateMeshFCFV.jl:300-367) against the oracle geometry, bit for bit."""
    m = cases.load_mesh(mesh)
    R = T.Reference()
    jm = T.jmesh(m)
    R.geometry_external2(jm, 1.0)
    assert np.array_equal(jm.Ω.a, m.Ω) and np.array_equal(jm.xc.a, m.xc) and np.array_equal(jm.yc.a, m.yc)
    assert np.array_equal(jm.n_x.a, m.n_x) and np.array_equal(jm.n_y.a, m.n_y) and np.array_equal(jm.Γ.a, m.Γ)
    assert np.array_equal(jm.τ.a, np.ones(m.nf))


def test_geometry_loops_triangle_numbering():
    """MakeTriangleMesh's variant (CreateMeshFCFV.jl:142-211: nodeA = [2 3 1], nodeB = [3 1 2], 3 x nel vertex list)
    on a seeded Delaunay triangulation with faces at local position i joining vertices nodeA[i], nodeB[i]."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(3)
    pts = rng.random((60, 2))
    tri = Delaunay(pts).simplices.astype(np.int64)
    nel = tri.shape[0]
    edges, e2f = {}, np.zeros((nel, 3), dtype=np.int64)
    A, B = [1, 2, 0], [2, 0, 1]
    for e in range(nel):
        for i in range(3):
            key = tuple(sorted((tri[e, A[i]], tri[e, B[i]])))
            e2f[e, i] = edges.setdefault(key, len(edges)) + 1
    nf = len(edges)
    xf, yf = np.zeros(nf), np.zeros(nf)
    for (p, q), f in edges.items():
        xf[f], yf[f] = 0.5 * (pts[p, 0] + pts[q, 0]), 0.5 * (pts[p, 1] + pts[q, 1])
    Om, xc, yc, nx, ny, G, tau = O.geometry(nel, nf, tri + 1, e2f, pts[:, 0], pts[:, 1], xf, yf, 0, 3.5)
    import types
    jm = types.SimpleNamespace(nel=nel, nf=nf, nf_el=3, xv=T.J(pts[:, 0]), yv=T.J(pts[:, 1]), xf=T.J(xf), yf=T.J(yf),
                               e2v=T.JArr(np.asfortranarray(tri + 1)), e2v_t=T.JArr(np.asfortranarray((tri + 1).T)),
                               e2f=T.JArr(np.asfortranarray(e2f)), bc=T.JArr(np.zeros(nf, dtype=np.int64)), τ=T.J(np.zeros(nf)))
    T.Reference().geometry_triangle(jm, 3.5)
    assert np.array_equal(jm.Ω.a, Om) and np.array_equal(jm.xc.a, xc) and np.array_equal(jm.yc.a, yc)
    assert np.array_equal(jm.n_x.a, nx) and np.array_equal(jm.n_y.a, ny) and np.array_equal(jm.Γ.a, G)
    assert np.array_equal(jm.τ.a, np.full(nf, 3.5))


def test_anisotropic_delta_branch():
    """delta != 1 goes through two 2x2 LAPACK inverses in Julia (Discretisation:234-235): 1e-12 only."""
    import translit_run
    case = cases.Case("S6", "inclusion", 1, "SymmetricGradient")
    m, d = cases.build(case)
    m.δ = 1.0 + np.random.default_rng(5).random(m.nel)
    r = translit_run.run_case(case, m, d, residual=False, elem_values=False)
    m2, _ = cases.build(case)
    m2.δ = m.δ.copy()
    a, b, Z, Kuu, Muu, Kup, fu, fp = _oracle(case, m2, d)
    for got, ref in ((r["Kuu"], Kuu), (r["Muu"], Muu), (r["Kup"], Kup)):
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        assert np.abs(got[2] - ref[2]).max() <= 1e-12 * np.abs(ref[2]).max()
    assert np.abs(r["fu"] - np.asarray(fu).ravel()).max() <= 1e-12 * np.abs(fu).max()


@pytest.mark.parametrize("bad", ["x = a ? b : c", "while true", "y = f(x) do z", "x = a | b", "x = !a", "x = v[2:end]", "x = 3^0.5",
                                 "function f(a; k=1)", "@views x = y", "x .*= 2"])
def test_rewriter_refuses_constructs_outside_the_rule_table(bad):
    with pytest.raises(ValueError):
        T.translate([(1, bad)])


def test_rule_a4_strong_zero():
    """Python's bool*float is Julia's strong zero for finite x (rule A4)."""
    import math
    for x in (2.5, -2.5, 0.0, -0.0, 1e-320, -1e300):
        assert math.copysign(1.0, False * x) == math.copysign(1.0, x) and False * x == 0.0 and True * x == x
        assert math.copysign(1.0, x * False) == math.copysign(1.0, x)


def test_face_numbering_of_the_dat_converter():
    """meshes/ReadMeshesDat.jl:37-101 (first-seen face ids by exact midpoint comparison, boundary tags) on
    square_TRI_H1.dat against the host mirror the H-mesh fixtures go through: e2f, xf, yf, bc identical."""
    import os
    from fcfv_nme23_b200 import mesh as M
    xv, yv, e2v = M.read_dat(os.path.join(cases.GOLDEN, "square_TRI_H1.dat"))
    e2f, xf, yf = M.number_faces_first_seen(xv, yv, e2v)
    bc = M._boundary_tags_unit_square(xf, yf)
    R = T.Reference()
    je2f, jxf, jyf, jbc = R.number_faces(T.J(xv), T.J(yv), T.JArr(np.asfortranarray(e2v)), int(e2v.shape[0]))
    assert np.array_equal(je2f.a, e2f) and np.array_equal(jxf.a, xf) and np.array_equal(jyf.a, yf)
    assert np.array_equal(jbc.a, bc)


def _ph_line(text):
    """One statement of StokesSolvers! (src/SolversFCFV_Stokes.jl) as a Python statement over NumPy vectors and SciPy sparse
    matrices.  Rule table (anything else is refused): broadcast dots disappear (`.=`, `.-`, `.+`, `.*`, `./` act elementwise on
    vectors, as NumPy's operators do); `x .+= y` -> `x = x + y`; `A*x` of a sparse matrix is SciPy's `*` (matrix product);
    `A'` -> `A.T`; `spdiagm(v)` -> diagonal sparse matrix; `mesh.Ω` stays an attribute."""
    import re
    s = T._strip_comment(text).strip()
    if not s:
        return None
    m = re.match(r"^(\w+)\s*\.\+=\s*(.*)$", s)
    if m:
        s = f"{m.group(1)} = {m.group(1)} + ({m.group(2)})"
    s = s.replace(".=", "=")
    for op in (".-", ".+", ".*", "./"):
        s = s.replace(op, op[1])
    s = re.sub(r"(\w)'", r"\1.T", s)
    if re.search(r"[@!?:\[\]{}|&^]|\bfor\b|\bif\b|\bend\b", s):
        raise ValueError(f"construct outside the rule table: {text!r}")
    return s


@pytest.mark.parametrize("case", [cases.Case("LowRes", "inclusion", 0, "Gradient", neumann=True), cases.Case("H1", "solkz", 1, "SymmetricGradient", tau_r=10.0)],
                         ids=lambda c: c.name)
def test_powell_hestenes_lines_from_the_source_text(case):
    """a9: the lines of StokesSolvers! the GPU path replaces (SolversFCFV_Stokes.jl:6 Kpu, :23-25 Schur complement, :40-41 ru / rp,
    :47 fusc, :49 pressure update), read from the reference's text and executed with SciPy's sparse algebra, against the oracle's
    restatement: 1e-12 (sums are taken in another order); the Schur complement's stored pattern exactly."""
    import os
    import types
    import scipy.sparse as sp
    path = os.path.join(T.REFERENCE, "src", "SolversFCFV_Stokes.jl")
    m, d = cases.build(case)
    a, b, Z, Kuu3, Muu3, Kup3, fu, fp = _oracle(case, m, d)
    nel, nf = m.nel, m.nf
    rng = np.random.default_rng(7)
    u0, p0 = rng.standard_normal(2 * nf), rng.standard_normal(nel)
    penalty = 1.0e3
    env = dict(Kuu=O.csc_to_scipy(Kuu3, (2 * nf, 2 * nf)), Kup=O.csc_to_scipy(Kup3, (2 * nf, nel)), fu=np.asarray(fu, dtype=float).ravel(),
               fp=np.asarray(fp, dtype=float).ravel(), u=u0.copy(), p=p0.copy(), penalty=penalty, mesh=types.SimpleNamespace(ke=m.ke, Ω=m.Ω),
               spdiagm=lambda v: sp.diags(np.asarray(v)).tocsc(), ru=None, rp=None, fusc=None)
    sig = [ln for k, ln in T.extract_lines(path, 6, 6)][0]
    kpu = re_default(sig, "Kpu")                                   # the keyword default `Kpu=-Kup'`
    env["Kpu"] = eval(_ph_line(kpu), {}, env)
    for first, last in ((23, 25), (40, 41), (47, 47), (49, 49)):
        for k, ln in T.extract_lines(path, first, last):
            st = _ph_line(ln)
            if st:
                exec(st, {}, env)
    Kc, Pc = Kuu3, Kup3
    ru, rp, nrm = O.ph_residual(m, Kc, Pc, fu, fp, u0, p0)
    tol = lambda ref: 1e-12 * max(1.0, np.abs(ref).max())
    assert np.abs(env["ru"] - ru).max() <= tol(ru) and np.abs(env["rp"] - rp).max() <= tol(rp)
    assert abs(np.linalg.norm(env["ru"]) - nrm[0]) <= 1e-12 * nrm[0] and abs(np.linalg.norm(env["rp"]) - nrm[1]) <= 1e-12 * nrm[1]
    fusc = O.ph_fusc(m, Pc, fu, fp, penalty, p0)
    assert np.abs(env["fusc"] - fusc).max() <= tol(fusc)
    pn = O.ph_update_p(m, Pc, fp, penalty, u0, p0)
    assert np.abs(env["p"] - pn).max() <= tol(pn)
    sc = O.csc_to_scipy(O.schur_complement(m, Kc, Pc, penalty), (2 * nf, 2 * nf))
    ref = env["Kuusc"].tocsc()
    ref.eliminate_zeros()                                           # Julia's sparse `.-` stores no zeros (dropzeros semantics)
    ref.sort_indices(); sc.sort_indices()
    assert np.array_equal(ref.indptr, sc.indptr) and np.array_equal(ref.indices, sc.indices)
    assert np.abs(ref.data - sc.data).max() <= 1e-12 * np.abs(sc.data).max()


def re_default(signature, name):
    """`name=<expr>` out of a Julia keyword-argument list (up to the next top-level comma)."""
    import re
    m = re.search(rf"\b{name}\s*=\s*([^,;)]+)", signature)
    assert m, signature
    return m.group(1).strip()
