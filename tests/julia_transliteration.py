"""Mechanical Julia -> Python transliteration of the REFERENCE'S OWN SOURCE TEXT -- TEST INFRASTRUCTURE.

Julia is not installed in this image, so the reference cannot be executed.  Instead of trusting a
hand-written restatement, this module reads the actual text of the reference's functions
(`/root/reference/src/*.jl`), rewrites it token by token with the rule table below and executes the
result with IEEE float64 scalars.  No arithmetic expression is retyped by a human: what runs is the
reference's own sequence of operations.  The oracle (`oracle/fcfv_oracle.c`) is then compared with
it -- bitwise for ComputeFCFV / ElementAssemblyLoop[NEW] / Sparsify / the geometry loops, 1e-12 for
the residual function (which goes through BLAS in Julia) -- in `tests/test_transliteration.py`, and
`tests/golden/make_translit_golden.py` stores its outputs as golden vectors that travel to boxes
without the reference checkout.

RULE TABLE (everything the rewriter does; each rule is a statement about Julia semantics):

  structure
   S1  `function f(a, b)` ... `end`        -> `def f(a, b):` + indented body (blocks are re-indented
                                             from the `for/if/elseif/else/end` keywords, not from the
                                             source's white space)
   S2  `for v=a:b` / `for v = a:b`          -> `for v in jl.rng(a, b):`   (inclusive 1-based range)
   S3  `if c` / `elseif c` / `else`         -> `if c:` / `elif c:` / `else:`
   S4  `# comment`                          -> dropped; trailing `;` dropped; `@inbounds` dropped
   S5  `x = @elapsed STMT`                  -> `STMT` then `x = 0.0` (a timing, not a value)
   S6  `@printf(fmt, v)`                    -> `jl.printf(fmt, v)` (records v, prints nothing)
   S7  `for i in eachindex(x)`              -> `for i in jl.rng(1, jl.length(x)):`
   S8  `a && b`, `a || b`, `true`, `false`  -> `a and b`, `a or b`, `True`, `False` (short-circuit operators bind
                                             looser than comparisons in both languages)
  scalars
   A1  `X^-1`                               -> `jl.inv(X)` = 1.0/X   (Base.literal_pow(^, x, Val(-1)) = inv(x))
   A2  `X^2`                                -> `jl.sq(X)`  = X*X     (Base.literal_pow(^, x, Val(2))  = x*x)
   A3  `:Name`                              -> `"Name"` (Symbols are only compared with ==)
   A4  Bool*Float64, Float64*Bool           -> Python's bool*float: `False*x == copysign(0.0, x)` for finite x,
                                             which is Julia's strong zero `ifelse(b, x, copysign(zero(x), x))`
   A6  `(p) & (q)` on Bools                 -> unchanged (only with both operands parenthesised: Python's `&` binds
                                             tighter than comparisons, Julia's like `*`)
   A5  n-ary `a*b*c`, `a+b+c`               -> unchanged: both languages fold left; unary minus binds like Julia's
                                             (`-a*b == (-a)*b`, `-a^-1 == -(a^-1)`); `1/2` is 0.5 in both
  arrays  (class JArr below: column-major, 1-based, Julia's broadcasting / product rules)
   B1  `zeros(...)`, `ones(...)`, `size`, `length`, `sqrt`, `inv`, `norm`, `sparse`, `dropzeros`, `Array`
                                            -> `jl.<same name>` (restated stdlib semantics, see each docstring)
   B2  `X[i,j]`, `X[e,:,:]`, `X[:]`          -> unchanged (JArr.__getitem__ is 1-based; `:` is Python's slice)
   B3  `[a b; c d]`, `[a; b]`, `[a b c]`     -> `jl.hvcat([[a, b], [c, d]])` etc. (a `[` that does not follow an
                                             identifier / `]` / `)` / `'` opens a literal)
   B4  postfix `'`                          -> `.adj` (adjoint; of a vector: a row that `*` treats like Julia's
                                             Adjoint{Vector}: row*vector is a scalar)
   B5  `.*`  `.+`  `.-`                      -> `@`  `+`  `-`  (JArr.__matmul__ is the ELEMENTWISE product, JArr.__mul__
                                             is Julia's `*`; precedence and associativity of the replaced operators
                                             are the same in both languages)
   B6  `X .= E`, `X[..] .+= E`              -> `jl.assign(X, E)`, `X[..] += E` (in-place broadcast assignment)
   B7  `x = []`, `push!(x, v)`              -> growable vector: `jl.hvcat([])`, `jl.push(x, v)` (meshes/ReadMeshesDat.jl only)

The rewriter refuses (raises) on any construct outside this table, so an edit of the reference that
introduces new syntax fails loudly instead of being silently mistranslated.
"""
from __future__ import annotations

import math
import os
import re
import types

import numpy as np

REFERENCE = os.environ.get("FCFV_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE, "src", "DiscretisationFCFV_Stokes.jl"))


# =================================================================================================
# Julia run-time emulation
# =================================================================================================
class JArr:
    """Julia Array emulation: column-major storage, 1-based indices, slices copy, linear indexing,
    broadcasting aligned on LEADING dimensions (a Vector is a column), `*` = matrix product."""
    __array_priority__ = 1000

    def __init__(self, a, adjvec=False):
        self.a = np.asarray(a)
        self.adjvec = adjvec          # Adjoint{T, Vector{T}}: stored with shape (1, n)

    # ---- shape helpers
    @property
    def shape(self):
        return self.a.shape

    def _index(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        out = []
        for k in idx:
            if isinstance(k, slice):
                if k != slice(None):
                    raise NotImplementedError("only `:` slices are transliterated")
                out.append(k)
            elif isinstance(k, JArr):
                raise NotImplementedError("array indices are not transliterated")
            else:
                ki = int(k)
                if ki != k:
                    raise IndexError(f"non-integer index {k}")
                if ki < 1:
                    raise IndexError(f"index {ki} < 1")
                out.append(ki - 1)
        return tuple(out)

    def __getitem__(self, idx):
        ix = self._index(idx)
        if len(ix) == 1 and self.a.ndim > 1:                 # linear indexing / X[:]
            flat = self.a.reshape(-1, order="F")
            return JArr(flat.copy()) if isinstance(ix[0], slice) else flat[ix[0]].item()
        if len(ix) != self.a.ndim:
            raise IndexError(f"{len(ix)} indices for a {self.a.ndim}-d array")
        r = self.a[ix]
        if all(not isinstance(k, slice) for k in ix):
            return r.item()
        return JArr(np.array(r))

    def __setitem__(self, idx, val):
        ix = self._index(idx)
        v = val.a if isinstance(val, JArr) else val
        if len(ix) == 1 and self.a.ndim > 1:
            if isinstance(ix[0], slice):
                self.a[...] = np.asarray(v).reshape(self.a.shape, order="F")
            else:
                i = np.unravel_index(ix[0], self.a.shape, order="F")
                self.a[i] = v
            return
        if len(ix) != self.a.ndim:
            raise IndexError(f"{len(ix)} indices for a {self.a.ndim}-d array")
        if all(not isinstance(k, slice) for k in ix):
            self.a[ix] = v
            return
        tgt = self.a[ix]
        v = np.asarray(v)
        if v.size != tgt.size and v.size != 1:
            raise ValueError(f"cannot assign {v.shape} to {tgt.shape}")
        self.a[ix] = v if v.size == 1 else v.reshape(tgt.shape, order="F")

    # ---- broadcasting (leading-dimension alignment)
    @staticmethod
    def _bc(x, y, op):
        xa = x.a if isinstance(x, JArr) else np.asarray(x)
        ya = y.a if isinstance(y, JArr) else np.asarray(y)
        nd = max(xa.ndim, ya.ndim)
        xa = xa.reshape(xa.shape + (1,) * (nd - xa.ndim))
        ya = ya.reshape(ya.shape + (1,) * (nd - ya.ndim))
        adj = (isinstance(x, JArr) and x.adjvec and (not isinstance(y, JArr) or y.adjvec or ya.size == 1)) or \
              (isinstance(y, JArr) and y.adjvec and (not isinstance(x, JArr) or xa.size == 1))
        return JArr(op(xa, ya), adjvec=bool(adj))

    def __add__(self, o): return JArr._bc(self, o, np.add)
    def __radd__(self, o): return JArr._bc(o, self, np.add)
    def __sub__(self, o): return JArr._bc(self, o, np.subtract)
    def __rsub__(self, o): return JArr._bc(o, self, np.subtract)
    def __matmul__(self, o): return JArr._bc(self, o, np.multiply)      # Julia `.*`
    def __rmatmul__(self, o): return JArr._bc(o, self, np.multiply)
    def __neg__(self): return JArr(-self.a, self.adjvec)

    def __truediv__(self, o):
        if isinstance(o, JArr):
            raise NotImplementedError("array / array is not transliterated")
        return JArr(self.a / o, self.adjvec)

    # ---- Julia `*`
    def __mul__(self, o):
        if not isinstance(o, JArr):
            return JArr(self.a * o, self.adjvec)
        A, B = self.a, o.a
        if self.adjvec and B.ndim == 1:                       # u' * n  -> scalar, left-to-right sum
            s = 0.0
            for k in range(B.shape[0]):
                s = s + A[0, k].item() * B[k].item() if k else A[0, 0].item() * B[0].item()
            return s
        A2 = A.reshape(A.shape[0], 1) if A.ndim == 1 else A
        B2 = B.reshape(B.shape[0], 1) if B.ndim == 1 else B
        if A2.shape[1] != B2.shape[0]:
            raise ValueError(f"DimensionMismatch: {A.shape} * {B.shape}")
        C = np.zeros((A2.shape[0], B2.shape[1]))
        for i in range(A2.shape[0]):
            for j in range(B2.shape[1]):
                s = A2[i, 0] * B2[0, j]
                for k in range(1, A2.shape[1]):
                    s = s + A2[i, k] * B2[k, j]
                C[i, j] = s
        if self.adjvec:                                       # u' * M -> Adjoint vector
            return JArr(C, adjvec=True)
        if B.ndim == 1:                                       # M * v -> Vector
            return JArr(C[:, 0].copy())
        return JArr(C)

    def __rmul__(self, o):
        return JArr(o * self.a, self.adjvec)

    @property
    def adj(self):
        if self.a.ndim == 1:
            return JArr(self.a.reshape(1, -1).copy(), adjvec=True)
        if self.adjvec:
            return JArr(self.a.reshape(-1).copy())
        return JArr(self.a.T.copy())


class JSparse:
    """SparseMatrixCSC{Float64,Int64}: colptr / rowval 1-based, rows ascending in a column."""

    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n, self.colptr, self.rowval, self.nzval = m, n, colptr, rowval, nzval

    def triple(self):
        return self.colptr, self.rowval, self.nzval


class _Runtime:
    """The `jl.` namespace of the transliterated code."""
    A = 1.0                       # src/DiscretisationFCFV_Stokes.jl:2 (module-level constant, checked in load())

    def __init__(self):
        self.printed = []

    @staticmethod
    def rng(a, b):
        return range(int(a), int(b) + 1)

    @staticmethod
    def inv(x):
        if isinstance(x, JArr):                               # LinearAlgebra.inv of a small matrix (LAPACK in Julia)
            return JArr(np.linalg.inv(x.a))
        return 1.0 / x

    @staticmethod
    def sq(x):
        return x * x

    @staticmethod
    def sqrt(x):
        return math.sqrt(x)

    @staticmethod
    def _dims(args):
        dt = np.float64
        if args and isinstance(args[0], type):
            dt, args = args[0], args[1:]
        if len(args) == 1 and isinstance(args[0], tuple):
            args = args[0]
        return dt, tuple(int(d) for d in args)

    @staticmethod
    def zeros(*args):
        dt, dims = _Runtime._dims(args)
        return JArr(np.zeros(dims, dtype=dt, order="F"))

    @staticmethod
    def ones(*args):
        dt, dims = _Runtime._dims(args)
        return JArr(np.ones(dims, dtype=dt, order="F"))

    @staticmethod
    def size(x, d=None):
        return x.a.shape if d is None else x.a.shape[d - 1]

    @staticmethod
    def length(x):
        return int(x.a.size)

    @staticmethod
    def norm(x):
        """LinearAlgebra.norm = 2-norm of the flattened array (BLAS nrm2 in Julia: compared at 1e-12)."""
        return math.sqrt(float(np.sum(np.square(x.a.reshape(-1, order="F")))))

    @staticmethod
    def hvcat(rows):
        """Julia literal: [a b; c d] -> Matrix, [a; b] -> Vector, [a b c] -> 1 x n Matrix."""
        if not rows:
            return JArr(np.zeros(0))
        if all(len(r) == 1 for r in rows) and len(rows) > 1:
            return JArr(np.array([r[0] for r in rows]))
        return JArr(np.array(rows))

    @staticmethod
    def assign(dst, src):
        """`dst .= src`"""
        s = src.a if isinstance(src, JArr) else np.asarray(src)
        dst.a[...] = s if s.size == 1 else s.reshape(dst.a.shape, order="F")

    @staticmethod
    def sparse(I, J, V, m, n):
        """SparseArrays.sparse(I, J, V, m, n): Float64 index vectors are converted to Int, duplicates (i, j) are
        combined with + in input order, the result is CSC with rows ascending inside a column and KEEPS entries
        whose sum is zero (dropzeros removes them afterwards)."""
        i = np.asarray(I.a).reshape(-1, order="F")
        j = np.asarray(J.a).reshape(-1, order="F")
        v = np.asarray(V.a, dtype=np.float64).reshape(-1, order="F")
        ii, jj = i.astype(np.int64), j.astype(np.int64)
        if not (np.array_equal(ii, i) and np.array_equal(jj, j)):
            raise ValueError("InexactError: non-integer index")
        m, n = int(m), int(n)
        if ii.size and (ii.min() < 1 or ii.max() > m or jj.min() < 1 or jj.max() > n):
            raise IndexError("sparse: index out of bounds")
        order = np.lexsort((ii, jj))                           # stable: input order inside a duplicate group
        ii, jj, v = ii[order], jj[order], v[order]
        new = np.ones(ii.size, dtype=bool)
        new[1:] = (ii[1:] != ii[:-1]) | (jj[1:] != jj[:-1])
        starts = np.nonzero(new)[0]
        vals = v[starts].copy()
        counts = np.diff(np.append(starts, ii.size))
        for extra in range(1, int(counts.max()) if counts.size else 1):   # sequential, in input order
            sel = counts > extra
            vals[sel] = vals[sel] + v[starts[sel] + extra]
        rows, cols = ii[starts], jj[starts]
        colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(cols - 1, minlength=n))]).astype(np.int64)
        return JSparse(m, n, colptr, rows.astype(np.int64), vals)

    @staticmethod
    def dropzeros(S):
        """Removes stored entries with iszero(v): +0.0 and -0.0."""
        keep = S.nzval != 0.0
        cols = np.repeat(np.arange(S.n), np.diff(S.colptr))
        colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(cols[keep], minlength=S.n))]).astype(np.int64)
        return JSparse(S.m, S.n, colptr, S.rowval[keep], S.nzval[keep])

    @staticmethod
    def Array(S):
        """Dense copy of a sparse matrix (used for fu: 2nf x 1)."""
        out = np.zeros((S.m, S.n), order="F")
        cols = np.repeat(np.arange(S.n), np.diff(S.colptr))
        out[S.rowval - 1, cols] = S.nzval
        return JArr(out)

    @staticmethod
    def push(x, v):
        """push!(x, v) on a Vector{Any} that only ever receives Float64 (ReadMeshesDat.jl:38-39,59-60)."""
        x.a = np.append(x.a.astype(np.float64), float(v))

    def printf(self, fmt, *vals):
        self.printed.append((fmt, vals))


# =================================================================================================
# source rewriting
# =================================================================================================
_IDENT = r"[^\W\d]\w*"      # unicode-aware identifier


def extract_function(path, name):
    """Lines (1-based numbers kept) of `function name(` ... matching `end`."""
    lines = open(path, encoding="utf-8").read().split("\n")
    start = next(k for k, ln in enumerate(lines) if re.match(rf"\s*function\s+{re.escape(name)}\s*\(", ln))
    depth, out = 0, []
    for k in range(start, len(lines)):
        code = _strip_comment(lines[k]).strip()
        out.append((k + 1, lines[k]))
        if re.match(r"(@inbounds\s+)?(function|for|if|while)\b", code):
            depth += 1
        elif re.match(r"end\b", code):
            depth -= 1
            if depth == 0:
                return out
    raise ValueError(f"unterminated function {name}")


def extract_lines(path, first, last):
    lines = open(path, encoding="utf-8").read().split("\n")
    return [(k, lines[k - 1]) for k in range(first, last + 1)]


def _strip_comment(line):
    out, instr = [], False
    for ch in line:
        if ch == '"':
            instr = not instr
        if ch == "#" and not instr:
            break
        out.append(ch)
    return "".join(out).rstrip()


def _match_back(s, close):
    """Index of the bracket that opens the one closing at s[close]."""
    pairs = {")": "(", "]": "["}
    depth = 0
    for k in range(close, -1, -1):
        if s[k] in ")]":
            depth += 1
        elif s[k] in "([":
            depth -= 1
            if depth == 0:
                if s[k] != pairs[s[close]]:
                    raise ValueError(f"mismatched brackets in: {s}")
                return k
    raise ValueError(f"unbalanced brackets in: {s}")


def _operand_start(s, end):
    """Start of the postfix-complete operand that ends just before s[end] (A1/A2)."""
    k = end - 1
    if s[k] == ")":
        k = _match_back(s, k)
        # function call / identifier in front of the parenthesis belongs to the operand
        m = re.search(rf"({_IDENT}(\.{_IDENT})*)$", s[:k])
        return m.start() if m else k
    if s[k] == "]":
        k = _match_back(s, k)
    m = re.search(rf"({_IDENT}(\.{_IDENT})*)$", s[:k] if s[end - 1] == "]" else s[:end])
    if not m:
        raise ValueError(f"cannot find the operand of ^ in: {s}")
    return m.start()


def _rewrite_pow(s):
    while True:
        m = re.search(r"\^\s*(-1|2)(?![\w.])", s)
        if not m:
            if "^" in s:
                raise ValueError(f"unsupported power in: {s}")
            return s
        st = _operand_start(s, m.start())
        fn = "jl.inv" if m.group(1) == "-1" else "jl.sq"
        s = s[:st] + f"{fn}({s[st:m.start()]})" + s[m.end():]


def _split_top(s, seps):
    """Splits s at top-level separators (not inside brackets)."""
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if depth == 0 and ch in seps:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return parts


def _rewrite_literals(s):
    """B3: array literals.  Scans left to right; a `[` is an index when it directly follows an identifier
    character, `]`, `)` or `'`, otherwise it opens a literal."""
    out, k = [], 0
    while k < len(s):
        ch = s[k]
        if ch == "[":
            prev = s[k - 1] if k > 0 else " "
            is_index = prev.isalnum() or prev in "_])'" or (ord(prev) > 127 and prev not in " ")
            depth, j = 0, k
            while True:
                if s[j] == "[":
                    depth += 1
                elif s[j] == "]":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            inner = s[k + 1:j]
            if is_index:
                out.append("[" + _rewrite_literals(inner) + "]")
            else:
                rows = []
                for row in _split_top(inner, ";"):
                    row = row.strip()
                    if not row:
                        continue
                    elems = [e for e in _split_top(re.sub(r"\s+", " ", row), " ") if e]
                    rows.append("[" + ", ".join(_rewrite_literals(e) for e in elems) + "]")
                out.append("jl.hvcat([" + ", ".join(rows) + "])")
            k = j + 1
        else:
            out.append(ch)
            k += 1
    return "".join(out)


_STDLIB = ("zeros", "ones", "size", "length", "sqrt", "inv", "norm", "sparse", "dropzeros", "Array")


def _rewrite_expr(s):
    """All expression-level rules on one statement (comments already removed)."""
    strings = []

    def stash(m):
        strings.append(m.group(0))
        return f"\x00{len(strings) - 1}\x00"
    s = re.sub(r'"(?:[^"\\]|\\.)*"', stash, s)                 # keep string literals out of the way
    t = s.replace("push!(", "push(").replace("&&", "").replace("||", "")
    t = re.sub(r"\)\s*&\s*\(", ")(", t)                                        # A6: parenthesised Bool & Bool
    if re.search(r"!(?!=)|\bend\b|\$|::|->|\bdo\b|\?|\||&", t):
        raise ValueError(f"construct outside the rule table: {s}")
    s = s.replace("&&", " and ").replace("||", " or ")                          # S8
    s = re.sub(r"(?<![\w.])true(?!\w)", "True", s)
    s = re.sub(r"(?<![\w.])false(?!\w)", "False", s)
    s = re.sub(r"(?<![\w.])push!\(", "jl.push(", s)                            # B7
    s = _rewrite_pow(s)                                                         # A1, A2
    s = re.sub(rf"(?<=[=\s(,]):({_IDENT})", r'"\1"', s)                         # A3
    s = _rewrite_literals(s)                                                    # B3
    s = re.sub(r"(?<=[\w\])'])'", ".adj", s)                                    # B4 (repeated quotes handled left to right)
    if re.search(r"\d\.[*+\-]", s):
        raise ValueError(f"ambiguous dotted operator after a number: {s}")
    s = s.replace(".*", "@").replace(".+", "+").replace(".-", "-")              # B5
    s = re.sub(r"(?<![\w.])(Float64|Int64|Int)(?![\w(])", lambda m: {"Float64": "np.float64", "Int64": "np.int64", "Int": "np.int64"}[m.group(1)], s)
    s = re.sub(rf"(?<![\w.])({'|'.join(_STDLIB)})\(", r"jl.\1(", s)             # B1
    s = re.sub(r"(?<![\w.])A(?![\w(\[])", "jl.A", s)                            # module constant A (Discretisation:2)
    s = re.sub(r"\x00(\d+)\x00", lambda m: strings[int(m.group(1))], s)
    return s


def translate(numbered_lines, header=None):
    """Python source of the given Julia lines.  `header`: a `def ...:` line for bare line ranges."""
    py, depth = [], 0
    if header:
        py.append(header)
        depth = 1

    def emit(text, d=None):
        py.append("    " * (depth if d is None else d) + text)

    last_was_header = bool(header)
    for no, raw in numbered_lines:
        code = _strip_comment(raw).strip()
        code = re.sub(r"^@inbounds\s+", "", code)                               # S4
        while code.endswith(";"):
            code = code[:-1].rstrip()
        if not code:
            continue
        try:
            m = re.match(r"function\s+(\w+)\s*\((.*)\)\s*$", code)
            if m:                                                               # S1
                if ";" in m.group(2) or "=" in m.group(2):
                    raise ValueError("keyword / default arguments are not transliterated")
                emit(f"def {m.group(1)}({m.group(2)}):")
                depth += 1
                last_was_header = True
                continue
            m = re.match(rf"for\s+({_IDENT})\s+in\s+eachindex\(({_IDENT})\)$", code)
            if m:                                                               # S7
                emit(f"for {m.group(1)} in jl.rng(1, jl.length({m.group(2)})):")
                depth += 1
                last_was_header = True
                continue
            m = re.match(rf"for\s+({_IDENT})\s*=\s*(.+?):(.+)$", code)
            if m:                                                               # S2
                emit(f"for {m.group(1)} in jl.rng({_rewrite_expr(m.group(2))}, {_rewrite_expr(m.group(3))}):")
                depth += 1
                last_was_header = True
                continue
            m = re.match(r"(if|elseif)\s+(.+)$", code)
            if m:                                                               # S3
                if m.group(1) == "if":
                    emit(f"if {_rewrite_expr(m.group(2))}:")
                    depth += 1
                else:
                    if last_was_header:
                        emit("pass")
                    emit(f"elif {_rewrite_expr(m.group(2))}:", depth - 1)
                last_was_header = True
                continue
            if code == "else":
                if last_was_header:
                    emit("pass")
                emit("else:", depth - 1)
                last_was_header = True
                continue
            if code == "end":
                if last_was_header:
                    emit("pass")
                depth -= 1
                last_was_header = False
                continue
            if re.match(r"(while|try|let|begin|struct|macro|do)\b", code):
                raise ValueError("block outside the rule table")
            last_was_header = False
            if code == "return":
                emit("return")
                continue
            m = re.match(r"return\s+(.+)$", code)
            if m:
                emit("return " + _rewrite_expr(m.group(1)))
                continue
            m = re.match(rf"({_IDENT})\s*=\s*@elapsed\s+(.+)$", code)
            if m:                                                               # S5
                emit(_rewrite_stmt(m.group(2)))
                emit(f"{m.group(1)} = 0.0")
                continue
            if code.startswith("@printf("):                                     # S6
                emit("jl.printf(" + _rewrite_expr(code[len("@printf("):]))
                continue
            if code.startswith("@"):
                raise ValueError("macro outside the rule table")
            emit("; ".join(_rewrite_stmt(p.strip()) for p in _split_top(code, ";") if p.strip()))
        except Exception as ex:
            raise ValueError(f"line {no}: {raw.strip()!r}: {ex}") from ex
    return "\n".join(py) + "\n"


def _rewrite_stmt(s):
    m = re.match(rf"({_IDENT})\s*\.=\s*(.+)$", s)                               # B6
    if m:
        return f"jl.assign({m.group(1)}, {_rewrite_expr(m.group(2))})"
    m = re.match(r"(.+?)\s*\.(\+|-)=\s*(.+)$", s)
    if m:
        return f"{_rewrite_expr(m.group(1))} {m.group(2)}= {_rewrite_expr(m.group(3))}"
    if re.search(r"\.[*/^%]?=", s):
        raise ValueError("unsupported broadcast assignment")
    return _rewrite_expr(s)


# =================================================================================================
# loading
# =================================================================================================
_DISC = "src/DiscretisationFCFV_Stokes.jl"
_RES = "src/ComputeResidualsFCFV_Stokes.jl"
_MESH = "src/CreateMeshFCFV.jl"


class Reference:
    """The transliterated reference functions, callable from Python with JArr arguments."""

    def __init__(self, root=REFERENCE):
        self.root = root
        self.jl = _Runtime()
        self.source = {}
        ns = {"jl": self.jl, "np": np}
        disc = os.path.join(root, _DISC)
        # module constant (Discretisation:1-2): the first non-comment line must be `A = 1.0`
        first = next(_strip_comment(ln).strip() for ln in open(disc, encoding="utf-8") if _strip_comment(ln).strip())
        m = re.match(r"A\s*=\s*([0-9.eE+-]+)$", first)
        if not m:
            raise ValueError(f"unexpected first statement of {_DISC}: {first}")
        self.jl.A = float(m.group(1))
        for name in ("ComputeFCFV", "ComputeElementValues", "ElementAssemblyLoop", "ElementAssemblyLoopNEW", "Sparsify"):
            self._load(ns, name, translate(extract_function(disc, name)))
        self._load(ns, "ComputeResidualsFCFV_Stokes_o1",
                   translate(extract_function(os.path.join(root, _RES), "ComputeResidualsFCFV_Stokes_o1")))
        # geometry loops of LoadExternalMesh2 (CreateMeshFCFV.jl:300-314 areas / centroids, :318-327 local numbering + arrays,
        # :337-367 normals) and of MakeTriangleMesh (:141-156, :160-162, :180-211), wrapped as functions of the mesh
        mp = os.path.join(root, _MESH)
        self._load(ns, "geometry_external2", translate(
            [(0, "nel = mesh.nel"), (0, "Ωe = zeros(nel)"), (0, "xc = zeros(nel)"), (0, "yc = zeros(nel)")]
            + extract_lines(mp, 300, 327) + extract_lines(mp, 337, 367), header="def geometry_external2(mesh, τr):"))
        # MakeTriangleMesh: Triangle's 3 x nel vertex list is the local `e2v`; the Voronoi neighbour lines (:185-188)
        # are mesh-generator specific and skipped
        self._load(ns, "geometry_triangle", translate(
            [(0, "nel = mesh.nel"), (0, "e2v = mesh.e2v_t"), (0, "Ωe = zeros(nel)"), (0, "xc = zeros(nel)"), (0, "yc = zeros(nel)")]
            + extract_lines(mp, 142, 170) + extract_lines(mp, 180, 184) + extract_lines(mp, 189, 211),
            header="def geometry_triangle(mesh, τr):"))
        # face numbering + boundary tags of the .dat -> .mat converter (meshes/ReadMeshesDat.jl:3-13, :37-101)
        rp = os.path.join(root, "meshes", "ReadMeshesDat.jl")
        self._load(ns, "IsItNew", translate(extract_function(rp, "IsItNew")))
        body = extract_lines(rp, 37, 101)
        if body[0][1].strip() != "ifac = 1" or body[-1][1].strip() != "end":
            raise ValueError("meshes/ReadMeshesDat.jl: unexpected layout")
        self._load(ns, "number_faces", translate(body + [(0, "return e2f, xf, yf, bc")], header="def number_faces(xv, yv, e2v, nel):"))
        self.ns = ns

    def _load(self, ns, name, src):
        self.source[name] = src
        exec(compile(src, f"<transliterated {name}>", "exec"), ns)

    def __getattr__(self, name):
        try:
            return self.__dict__["ns"][name]
        except KeyError:
            raise AttributeError(name)


# ---- conversion between the tests' FCFV_Mesh / numpy world and the JArr world --------------------
_MESH_FIELDS = ("e2f", "e2v", "bc", "Ω", "Γ", "n_x", "n_y", "ke", "δ", "τe", "τ", "xv", "yv", "xf", "yf", "xc", "yc")


def jmesh(m):
    """SimpleNamespace with the FCFV_Mesh field names of the reference (CreateMeshFCFV.jl:10-44), JArr fields."""
    j = types.SimpleNamespace(nel=int(m.nel), nf=int(m.nf), nf_el=3, nn_el=3)
    for k in _MESH_FIELDS:
        v = getattr(m, k, None)
        if v is not None:
            a = np.array(v, order="F")
            if k in ("e2f", "e2v", "bc"):
                a = a.astype(np.int64)
            setattr(j, k, JArr(a))
    if getattr(j, "τ", None) is None:          # LoadExternalMesh leaves mesh.τ missing; the shim allocates it
        j.τ = JArr(np.zeros(j.nf))
    return j


def J(x):
    return JArr(np.array(x, dtype=np.float64, order="F"))


def N(x):
    return x.a if isinstance(x, JArr) else x
