"""Test cases shared by the CPU and GPU suites: mesh fixture x problem setting x assembly variant
x formulation, with seeded (numpy default_rng(0)) data so that oracle and CUDA path see identical
inputs.  TEST INFRASTRUCTURE (imports the oracle for the mesh geometry of the H-meshes)."""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from fcfv_nme23_b200 import mesh as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@dataclass(frozen=True)
class Case:
    mesh: str          # LowRes | MedRes | H1 | H2 | S<n>
    setting: str       # inclusion | solkz
    variant: int       # 0 ElementAssemblyLoop, 1 ElementAssemblyLoopNEW
    form: str          # Gradient | SymmetricGradient
    neumann: bool = False
    tau_r: float = 2.0

    @property
    def name(self):
        return f"{self.mesh}-{self.setting}-{'NEW' if self.variant else 'OLD'}-{self.form}{'-neu' if self.neumann else ''}"


def all_cases(mesh_dir=GOLDEN, small_only=False):
    out = []
    meshes = ["LowRes", "H1", "S6"] if small_only else ["LowRes", "MedRes", "H1", "H2", "S6", "S17"]
    for mesh in meshes:
        settings = ["inclusion"] if mesh.endswith("Res") else ["solkz", "inclusion"]
        for setting in settings:
            for variant in (0, 1):
                for form in ("Gradient", "SymmetricGradient"):
                    out.append(Case(mesh, setting, variant, form, neumann=False, tau_r=2.0 if setting == "inclusion" else 10.0))
    # Neumann south boundary (README.md:24) on two meshes
    for mesh in ("LowRes", "S6"):
        for variant in (0, 1):
            for form in ("Gradient", "SymmetricGradient"):
                out.append(Case(mesh, "inclusion", variant, form, neumann=True))
    return out


def large_cases():
    """BASELINE config 2 beyond H2: SolKz on square_TRI_H3 / H4 (the largest meshes the reference ships; H5 is missing)."""
    out = [Case("H3", "solkz", v, f, tau_r=10.0) for v in (0, 1) for f in ("Gradient", "SymmetricGradient")]
    out += [Case("H4", "solkz", 1, "SymmetricGradient", tau_r=10.0), Case("H4", "solkz", 0, "Gradient", tau_r=10.0)]
    return out


_mesh_cache = {}


def load_mesh(name, mesh_dir=GOLDEN):
    """Mesh with geometry.  .mat meshes carry their geometry; H/S meshes get it from the ORACLE
    geometry (a2) so that CPU-only tests have it; GPU tests check the CUDA geometry against it."""
    from oracle import oracle as O
    key = (name, mesh_dir)
    if key in _mesh_cache:
        return _copy_mesh(_mesh_cache[key])
    if name in ("LowRes", "MedRes"):
        m = M.LoadExternalMesh(name, [1.0, 1e-2], mesh_dir=mesh_dir)
    elif name.startswith("H"):
        m = M.LoadExternalMesh2(τr=1.0, name=f"square_TRI_{name}", mesh_dir=mesh_dir, geometry=False)
        O.mesh_geometry(m, 1.0)
    elif name.startswith("S"):
        m = M.MakeStructuredMesh(int(name[1:]), geometry=False)
        O.mesh_geometry(m, 1.0)
    else:
        raise KeyError(name)
    _mesh_cache[key] = m
    return _copy_mesh(m)


def _copy_mesh(m):
    import copy
    c = copy.copy(m)
    for k, v in vars(m).items():
        if isinstance(v, np.ndarray):
            setattr(c, k, v.copy(order="K"))
    if hasattr(c, "_fcfv_handle"):
        del c._fcfv_handle
    return c


def build(case: Case, mesh_dir=GOLDEN):
    """Returns (mesh, data) for a case.  Settings follow the drivers:
    inclusion: ke from phase / radius, τe = τr (ViscousInclusionFCFV_v2.jl:127, length nf);
    solkz: ke = exp(13.8 yc), sey = sin(1.6π yc) cos(3π xc), τe = τr·max(ke,1) (SolKzFCFV.jl:33,149)."""
    m = load_mesh(case.mesh, mesh_dir)
    nel, nf = m.nel, m.nf
    rng = np.random.default_rng(0)
    if case.setting == "inclusion":
        if not case.mesh.endswith("Res"):
            R = 1.0 / 6.0
            inside = np.hypot(m.xc - 0.5, m.yc - 0.5) < R
            m.ke = np.where(inside, 1e-2, 1.0)
            # interface faces: between elements of different viscosity
            f = np.asarray(m.e2f) - 1
            kmin = np.full(nf, np.inf); kmax = np.full(nf, -np.inf)
            for i in range(3):
                np.minimum.at(kmin, f[:, i], m.ke); np.maximum.at(kmax, f[:, i], m.ke)
            m.bc = np.where((m.bc == 0) & (kmin != kmax), -1, m.bc).astype(np.int64)
        m.τe = case.tau_r * np.ones(nf)          # v2.jl:127 -- length nf on purpose
        sex = np.zeros(nel); sey = np.zeros(nel)
        cx, cy = (0.0, 0.0) if case.mesh.endswith("Res") else (0.5, 0.5)
        VxDir = m.xf - cx; VyDir = -(m.yf - cy)  # pure shear
    else:
        m.ke = np.exp(13.8 * m.yc)
        m.τe = case.tau_r * np.maximum(m.ke, 1.0)
        sex = np.zeros(nel); sey = np.sin(1.6 * np.pi * m.yc) * np.cos(3 * np.pi * m.xc)
        VxDir = np.sin(np.pi * m.xf) * np.cos(np.pi * m.yf); VyDir = -np.cos(np.pi * m.xf) * np.sin(np.pi * m.yf)
    if case.neumann:
        south = (m.bc == 1) & (m.yf <= m.yf.min() + 1e-12)
        m.bc = np.where(south, 2, m.bc).astype(np.int64)
    m.τ = None                                     # as after LoadExternalMesh
    neu = [rng.standard_normal(nf) for _ in range(4)]
    data = dict(sex=sex, sey=sey, VxDir=VxDir, VyDir=VyDir, neu=neu,
                Vxh=rng.standard_normal(nf), Vyh=rng.standard_normal(nf), Pe=rng.standard_normal(nel))
    return m, data
