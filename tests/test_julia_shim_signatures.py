"""julia/FCFV_B200.jl cannot be executed here (Julia is not installed), so its `ccall`s are checked mechanically against
include/fcfv_b200.h: every bound symbol is declared, the return type is right, the argument-type tuple has the header's
arity and every Julia type is one that matches the C parameter type, and the call passes as many values as types."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# C parameter type (qualifiers stripped) -> Julia types a ccall may declare for it
C2J = {
    "fcfv_t*": {"Ptr{Cvoid}"},
    "fcfv_t**": {"Ref{Ptr{Cvoid}}"},
    "int": {"Cint"},
    "int64_t": {"Int64"},
    "double": {"Float64"},
    "double*": {"Ptr{Float64}", "Ref{Float64}", "Ptr{Cvoid}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}", "Ptr{Cvoid}"},
    "int*": {"Ref{Cint}", "Ptr{Cint}"},
    "uint8_t*": {"Ptr{UInt8}", "Ptr{Cvoid}"},
    "void*": {"Ptr{Cvoid}", "Ptr{Float64}", "Ptr{Int64}"},
    "void**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
}
RET = {"int": "Cint", "const char*": "Cstring", "void*": "Ptr{Cvoid}", "int64_t": "Int64"}


def header_decls():
    text = open(os.path.join(ROOT, "include", "fcfv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int64_t|int|const char\*|void\*)\s+(fcfv_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if params and params != "void":
            for prm in params.split(","):
                prm = re.sub(r"\bconst\b", "", prm).strip()
                t = re.match(r"(.*?)(\w+)$", prm).group(1).replace(" ", "")      # drop the parameter name
                types.append(t)
        decls[name] = (ret, types)
    return decls


def split_top(s):
    """Splits at commas that are not inside (), {} or []."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def shim_ccalls():
    text = open(os.path.join(ROOT, "julia", "FCFV_B200.jl")).read()
    text = re.sub(r"#[^\n]*", "", text)
    calls = []
    for m in re.finditer(r"ccall\(", text):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        parts = split_top(text[m.end():i - 1])
        sym = re.match(r"\(:(\w+),\s*LIB\)", parts[0]).group(1)
        argtypes = split_top(parts[2].strip()[1:-1]) if parts[2].strip() != "()" else []
        calls.append((sym, parts[1], argtypes, parts[3:]))
    return calls


def test_every_ccall_matches_the_header():
    decls = header_decls()
    calls = shim_ccalls()
    assert len(calls) >= 25
    for sym, ret, argtypes, values in calls:
        assert sym in decls, f"{sym} is not declared in include/fcfv_b200.h"
        cret, ctypes_ = decls[sym]
        assert ret == RET[cret], f"{sym}: return type {ret} for C {cret}"
        assert len(argtypes) == len(ctypes_), f"{sym}: {len(argtypes)} argument types, the header has {len(ctypes_)}"
        assert len(values) == len(argtypes), f"{sym}: {len(values)} values for {len(argtypes)} types"
        for k, (jt, ct) in enumerate(zip(argtypes, ctypes_)):
            assert ct in C2J, f"{sym}: unmapped C type {ct}"
            assert jt in C2J[ct], f"{sym} argument {k + 1}: Julia {jt} does not match C {ct}"


def test_the_reference_functions_are_exported():
    text = open(os.path.join(ROOT, "julia", "FCFV_B200.jl")).read()
    for fn in ("ComputeFCFV", "ElementAssemblyLoop", "ElementAssemblyLoopNEW", "ComputeResidualsFCFV_Stokes_o1", "ComputeElementValues"):
        assert re.search(rf"^\s*(function\s+{fn}\(|{fn}\()", text, flags=re.M), fn
    # same positional arguments as the reference (src/DiscretisationFCFV_Stokes.jl:8,102,210; src/ComputeResidualsFCFV_Stokes.jl:5)
    assert "function ComputeFCFV(mesh, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu, τr, Formulation)" in text
    assert "ElementAssemblyLoop(mesh, α, β, Ζ, VxDir, VyDir, σxxNeu, σyyNeu, σxyNeu, σyxNeu, Formulation)" in text
    assert ("function ComputeResidualsFCFV_Stokes_o1(Vxh, Vyh, Pe, mesh, ae, be, ze, sex, sey, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, "
            "SyxNeu,") in text
