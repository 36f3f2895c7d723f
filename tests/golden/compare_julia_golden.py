"""Closes the "parity unpinned" gap wherever Julia is available.

    julia --project=<reference checkout> julia/gen_golden.jl golden_julia      # dumps the REAL reference's outputs
    python tests/golden/compare_julia_golden.py golden_julia [--gpu]           # checks the oracle (and the CUDA path)

Every case directory written by julia/gen_golden.jl holds raw little-endian arrays (Float64 / Int64, column-major):
alpha, beta, Z, tau, fu, fp and the CSC triples of Kuu, Muu, Kup.  The same inputs are rebuilt here (they only use
correctly rounded IEEE operations), the oracle -- and with --gpu the CUDA path through the C ABI -- is run on them and
every array is compared bit for bit.  Exit status 0 = identical everywhere.

    python tests/golden/compare_julia_golden.py --self-test DIR     # writes the ORACLE's outputs in the same format
                                                                    # (what CI uses to test this script without Julia)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = [(res, form, v) for res in ("LowRes", "MedRes") for form in ("Gradient", "SymmetricGradient") for v in (0, 1)]
FLOATS = ("alpha", "beta", "Z", "tau", "fu", "fp")


def inputs(res):
    """Mesh and fields of gen_golden.jl's dump_case()."""
    import cases
    m = cases.load_mesh(res)
    nel, nf = m.nel, m.nf
    d = dict(VxDir=m.xf.copy(), VyDir=-m.yf.copy(), sex=np.zeros(nel), sey=np.zeros(nel),
             neu=[k * m.xf + m.yf / k for k in (1.0, 2.0, 3.0, 4.0)])
    m.τe = 2.0 * np.ones(nf)
    m.τ = np.zeros(nf)
    return m, d


def run_oracle(res, form, variant):
    from oracle import oracle as O
    m, d = inputs(res)
    a, b, Z = O.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], 2.0, form)
    fn = O.ElementAssemblyLoop if variant == 0 else O.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *d["neu"], form)
    out = dict(alpha=a, beta=b.ravel(order="F"), Z=Z.ravel(order="F"), tau=m.τ, fu=fu.ravel(), fp=fp)
    for tag, K in (("Kuu", Kuu), ("Muu", Muu), ("Kup", Kup)):
        out[tag + "_colptr"], out[tag + "_rowval"], out[tag + "_nzval"] = K
    return out


def run_gpu(res, form, variant):
    from fcfv_nme23_b200 import api
    m, d = inputs(res)
    a, b, Z = api.ComputeFCFV(m, d["sex"], d["sey"], d["VxDir"], d["VyDir"], *d["neu"], 2.0, form)
    fn = api.ElementAssemblyLoop if variant == 0 else api.ElementAssemblyLoopNEW
    Kuu, Muu, Kup, fu, fp, _ = fn(m, a, b, Z, d["VxDir"], d["VyDir"], *d["neu"], form)
    api.handle_for(m).close()
    out = dict(alpha=a, beta=b.ravel(order="F"), Z=Z.ravel(order="F"), tau=m.τ, fu=np.asarray(fu).ravel(), fp=fp)
    for tag, K in (("Kuu", Kuu), ("Muu", Muu), ("Kup", Kup)):
        out[tag + "_colptr"], out[tag + "_rowval"], out[tag + "_nzval"] = K.indptr + 1, K.indices + 1, K.data
    return out


def write_case(directory, arrays):
    os.makedirs(directory, exist_ok=True)
    for k, v in arrays.items():
        dtype = np.float64 if (k in FLOATS or k.endswith("nzval")) else np.int64
        np.ascontiguousarray(v, dtype=dtype).tofile(os.path.join(directory, k + ".bin"))


def compare_case(directory, arrays, who):
    bad = 0
    for k, v in arrays.items():
        dtype = np.float64 if (k in FLOATS or k.endswith("nzval")) else np.int64
        ref = np.fromfile(os.path.join(directory, k + ".bin"), dtype=dtype)
        got = np.ascontiguousarray(v, dtype=dtype).ravel()
        if ref.shape != got.shape:
            print(f"  {who} {k}: size {got.size} != reference {ref.size}")
            bad += 1
        elif not np.array_equal(ref, got):
            i = int(np.flatnonzero(ref != got)[0])
            rel = float(np.max(np.abs(ref - got)) / max(np.max(np.abs(ref)), 1e-300)) if dtype == np.float64 else float("nan")
            print(f"  {who} {k}: {int((ref != got).sum())} of {ref.size} entries differ, first at {i}: {got[i]!r} vs reference {ref[i]!r}"
                  f" (max rel. deviation {rel:.2e})")
            bad += 1
    return bad


def main(argv):
    if len(argv) >= 2 and argv[0] == "--self-test":
        for res, form, v in CASES[:4]:
            write_case(os.path.join(argv[1], f"{res}_{form}_{v}"), run_oracle(res, form, v))
        return 0
    if not argv:
        print(__doc__)
        return 2
    root, gpu = argv[0], "--gpu" in argv
    bad = seen = 0
    for res, form, v in CASES:
        directory = os.path.join(root, f"{res}_{form}_{v}")
        if not os.path.isdir(directory):
            continue
        seen += 1
        n = compare_case(directory, run_oracle(res, form, v), "oracle")
        if gpu:
            n += compare_case(directory, run_gpu(res, form, v), "CUDA")
        print(f"{res} {form} loop {'NEW' if v else 'old'}: {'identical' if n == 0 else str(n) + ' arrays differ'}")
        bad += n
    if seen == 0:
        print(f"no case directory under {root}")
        return 2
    print("PARITY PINNED: every array identical to the Julia reference" if bad == 0 else f"{bad} arrays differ from the Julia reference")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
