// fcfv_generate.cuh -- on-device generator of the synthetic criss-cross mesh S(n) and of the
// closed-form field settings the benchmark uses.  Host generation of 10^7..10^8
// elements would dominate a run; here a strip of quad rows is produced directly in HBM.
//
// Topology / numbering (identical to fcfv_nme23_b200/mesh.py: structured_connectivity, which is
// checked against the first-seen rule of the reference's meshes/ReadMeshesDat.jl:36-88):
//   quads row-major, triangles of a quad consecutive: bottom, right, top, left, each stored
//   (corner, centre, corner) counter-clockwise; local faces (v1,v3), (v3,centre), (v1,centre);
//   base(qx,qy) = 6*(qy*n+qx) + (qy==0 ? qx : n) + qy + (qx>0);
//   new faces of a quad in order: [bottom if qy==0] d00 d10 right d11 top d01 [left if qx==0].
// A strip holds quad rows [rlo, rhi) = owned rows [row0,row1) plus one ghost row on each side that
// has a neighbour.  Local element ids are consecutive; local face ids are ascending in the global
// id: first the n bottom edges of row rlo (they belong to row rlo-1), then the contiguous global
// range [base(0,rlo), base(0,rhi)).
#pragma once
#include "fcfv_math.cuh"

namespace fcfv {

struct StripLayout {
    long long n, row0, row1, rlo, rhi;
    long long nquads, nel, nf;
    long long fbase;      // global id of the first face first-seen in row rlo
    long long nbottom;    // n if rlo > 0 else 0
    long long ebase;      // global id of the first local element
};

FCFV_HD long long face_base(long long n, long long qx, long long qy) {
    return 6 * (qy * n + qx) + (qy == 0 ? qx : n) + qy + (qx > 0 ? 1 : 0);
}

inline StripLayout make_strip_layout(long long n, long long row0, long long row1) {
    StripLayout S;
    S.n = n; S.row0 = row0; S.row1 = row1;
    S.rlo = row0 > 0 ? row0 - 1 : 0;
    S.rhi = row1 < n ? row1 + 1 : n;
    S.nquads = (S.rhi - S.rlo) * n;
    S.nel = 4 * S.nquads;
    S.fbase = face_base(n, 0, S.rlo);
    S.nbottom = S.rlo > 0 ? n : 0;
    S.nf = S.nbottom + (face_base(n, 0, S.rhi) - S.fbase);
    S.ebase = 4 * n * S.rlo;
    return S;
}

// local ids of the 8 faces around quad (qx,qy): bottom d00 d10 right d11 top d01 left
FCFV_HD void quad_faces(const StripLayout& S, long long qx, long long qy, int f[8]) {
    const long long n = S.n;
    const long long b = face_base(n, qx, qy);
    const long long o = qy == 0 ? 1 : 0;
    const long long shift = S.nbottom - S.fbase;   // global -> local for ids >= fbase
    f[1] = (int)(b + o + shift);
    f[2] = (int)(b + o + 1 + shift);
    f[3] = (int)(b + o + 2 + shift);
    f[4] = (int)(b + o + 3 + shift);
    f[5] = (int)(b + o + 4 + shift);
    f[6] = (int)(b + o + 5 + shift);
    if (qy == 0) f[0] = (int)(b + shift);
    else if (qy == S.rlo) f[0] = (int)qx;   // top edge of the row below the strip
    else f[0] = (int)(face_base(n, qx, qy - 1) + (qy - 1 == 0 ? 1 : 0) + 4 + shift);
    if (qx == 0) f[7] = (int)(b + o + 6 + shift);
    else f[7] = (int)(face_base(n, qx - 1, qy) + o + 2 + shift);
}

// faces of triangle t (0 bottom, 1 right, 2 top, 3 left) from the 8 quad faces:
// bottom (bottom,d00,d10) right (right,d10,d11) top (top,d11,d01) left (left,d01,d00)
FCFV_HD void tri_faces(const int f[8], int t, int& f1, int& f2, int& f3) {
    if (t == 0)      { f1 = f[0]; f2 = f[1]; f3 = f[2]; }
    else if (t == 1) { f1 = f[3]; f2 = f[2]; f3 = f[4]; }
    else if (t == 2) { f1 = f[5]; f2 = f[4]; f3 = f[6]; }
    else             { f1 = f[7]; f2 = f[6]; f3 = f[1]; }
}

struct GenOut {
    int* e2f; int8_t* bc;
    double *Om, *nx, *ny, *G, *ke, *dl, *taue, *sex, *sey, *Pe;
    double *VxDir, *VyDir, *sxx, *syy, *sxy, *syx, *Vxh, *Vyh;
    uint8_t *elem_owned, *face_owned;
};

#if defined(__CUDACC__)

constexpr double kPi = 3.14159265358979323846264338328;

// smooth closed-form data evaluated at a face midpoint
__device__ __forceinline__ void face_fields(int setting, double x, double y, double& vxd, double& vyd, double& sxx,
                                            double& syy, double& sxy, double& syx, double& vxh, double& vyh) {
    if (setting == 0) {   // SolKz-like: smooth divergence-free Dirichlet field
        vxd = sin(kPi * x) * cos(kPi * y);
        vyd = -cos(kPi * x) * sin(kPi * y);
        sxx = 0.0; syy = 0.0; sxy = 0.0; syx = 0.0;
    } else {              // inclusion: pure shear, constant boundary stress
        vxd = x - 0.5;
        vyd = -(y - 0.5);
        sxx = 2.0; syy = -2.0; sxy = 0.0; syx = 0.0;
    }
    vxh = sin(2.0 * kPi * x) * cos(3.0 * kPi * y) + x;
    vyh = cos(2.0 * kPi * x) * sin(kPi * y) - y;
}

__global__ void k_gen_elements(StripLayout S, GenOut O, int setting, double tau_r, double eta_incl) {
    const long long le = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (le >= S.nel) return;
    const long long nel = S.nel, n = S.n;
    const long long ql = le >> 2;
    const int t = (int)(le & 3);
    const long long qy = S.rlo + ql / n, qx = ql % n;
    const double dn = (double)n;
    const double x0 = (double)qx / dn, x1 = (double)(qx + 1) / dn;
    const double y0 = (double)qy / dn, y1 = (double)(qy + 1) / dn;
    const double xm = (2.0 * (double)qx + 1.0) / (2.0 * dn), ym = (2.0 * (double)qy + 1.0) / (2.0 * dn);
    // (v1, centre, v3): bottom (c10,m,c00) right (c11,m,c10) top (c01,m,c11) left (c00,m,c01)
    double x[3], y[3];
    x[1] = xm; y[1] = ym;
    if (t == 0)      { x[0] = x1; y[0] = y0; x[2] = x0; y[2] = y0; }
    else if (t == 1) { x[0] = x1; y[0] = y1; x[2] = x1; y[2] = y0; }
    else if (t == 2) { x[0] = x0; y[0] = y1; x[2] = x1; y[2] = y1; }
    else             { x[0] = x0; y[0] = y0; x[2] = x0; y[2] = y1; }
    const double xf[3] = {0.5 * (x[0] + x[2]), 0.5 * (x[2] + x[1]), 0.5 * (x[0] + x[1])};
    const double yf[3] = {0.5 * (y[0] + y[2]), 0.5 * (y[2] + y[1]), 0.5 * (y[0] + y[1])};
    ElemGeom g;
    element_geometry(x, y, xf, yf, 1, g);
    int f[8];
    quad_faces(S, qx, qy, f);
    int f1, f2, f3;
    tri_faces(f, t, f1, f2, f3);
    O.e2f[le] = f1; O.e2f[nel + le] = f2; O.e2f[2 * nel + le] = f3;
    O.Om[le] = g.Om;
#pragma unroll
    for (int k = 0; k < 3; k++) { O.nx[k * nel + le] = g.nx[k]; O.ny[k * nel + le] = g.ny[k]; O.G[k * nel + le] = g.G[k]; }
    double ke, te, sy;
    if (setting == 0) {
        ke = exp((2.0 * 6.9) * g.yc);                           // EvalAnalSolKz.jl:450, B = 6.9
        sy = sin((1.6 * kPi) * g.yc) * cos((3.0 * kPi) * g.xc); // sey = -rho (SolKzFCFV.jl:33, EvalAnalSolKz.jl:437)
        te = tau_r * fmax(ke, 1.0);                             // SolKzFCFV.jl:149
    } else {
        const double dx = g.xc - 0.5, dy = g.yc - 0.5;
        ke = (sqrt(dx * dx + dy * dy) < 1.0 / 6.0) ? eta_incl : 1.0;
        sy = 0.0;
        te = tau_r;                                             // ViscousInclusionFCFV_v2.jl:127
    }
    O.ke[le] = ke; O.dl[le] = 1.0; O.taue[le] = te;
    O.sex[le] = 0.0; O.sey[le] = sy;
    O.Pe[le] = cos(kPi * g.xc) * sin(2.0 * kPi * g.yc);
    O.elem_owned[le] = (qy >= S.row0 && qy < S.row1) ? 1 : 0;
}

__device__ __forceinline__ void put_face(const GenOut& O, int f, int setting, double x, double y, int bc, bool owned) {
    double vxd, vyd, sxx, syy, sxy, syx, vxh, vyh;
    face_fields(setting, x, y, vxd, vyd, sxx, syy, sxy, syx, vxh, vyh);
    O.bc[f] = (int8_t)bc;
    O.VxDir[f] = vxd; O.VyDir[f] = vyd;
    O.sxx[f] = sxx; O.syy[f] = syy; O.sxy[f] = sxy; O.syx[f] = syx;
    O.Vxh[f] = vxh; O.Vyh[f] = vyh;
    O.face_owned[f] = owned ? 1 : 0;
}

// One thread per quad writes the data of the faces first seen in that quad (owner = the rank that
// owns the face's highest-numbered adjacent element, consistent with the tau last-writer rule).
__global__ void k_gen_faces(StripLayout S, GenOut O, int setting, int neumann_south) {
    const long long ql = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ql >= S.nquads) return;
    const long long n = S.n;
    const long long qy = S.rlo + ql / n, qx = ql % n;
    const double dn = (double)n;
    const double x0 = (double)qx / dn, x1 = (double)(qx + 1) / dn;
    const double y0 = (double)qy / dn, y1 = (double)(qy + 1) / dn;
    const double xm = (2.0 * (double)qx + 1.0) / (2.0 * dn), ym = (2.0 * (double)qy + 1.0) / (2.0 * dn);
    int f[8];
    quad_faces(S, qx, qy, f);
    const bool row_owned = qy >= S.row0 && qy < S.row1;
    const bool up_owned = (qy + 1 < n) ? (qy + 1 >= S.row0 && qy + 1 < S.row1) : row_owned;
    const int south = neumann_south ? 2 : 1;
    if (qy == 0) put_face(O, f[0], setting, 0.5 * (x1 + x0), 0.5 * (y0 + y0), south, row_owned);
    else if (qy == S.rlo) put_face(O, f[0], setting, 0.5 * (x1 + x0), 0.5 * (y0 + y0), 0, row_owned);
    put_face(O, f[1], setting, 0.5 * (x0 + xm), 0.5 * (y0 + ym), 0, row_owned);                          // d00
    put_face(O, f[2], setting, 0.5 * (x1 + xm), 0.5 * (y0 + ym), 0, row_owned);                          // d10
    put_face(O, f[3], setting, 0.5 * (x1 + x1), 0.5 * (y1 + y0), qx == n - 1 ? 1 : 0, row_owned);        // right
    put_face(O, f[4], setting, 0.5 * (x1 + xm), 0.5 * (y1 + ym), 0, row_owned);                          // d11
    put_face(O, f[5], setting, 0.5 * (x0 + x1), 0.5 * (y1 + y1), qy == n - 1 ? 1 : 0, up_owned);         // top
    put_face(O, f[6], setting, 0.5 * (x0 + xm), 0.5 * (y1 + ym), 0, row_owned);                          // d01
    if (qx == 0) put_face(O, f[7], setting, 0.5 * (x0 + x0), 0.5 * (y0 + y1), 1, row_owned);             // left
}

// inclusion setting: faces between elements of different viscosity are material interfaces (bc = -1)
__global__ void k_gen_interface(const int2* __restrict__ f2e, const double* __restrict__ ke, int8_t* __restrict__ bc,
                                int nf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int2 pr = f2e[f];
    if (pr.x >= 0 && pr.y != pr.x && bc[f] == 0 && ke[pr.x >> 2] != ke[pr.y >> 2]) bc[f] = -1;
}

__global__ void k_widen_e2f(const int* __restrict__ in, long long* __restrict__ out, long long n3) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n3) out[k] = (long long)in[k] + 1;
}

__global__ void k_widen_bc(const int8_t* __restrict__ in, long long* __restrict__ out, long long nf) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nf) out[k] = (long long)in[k];
}

#endif  // __CUDACC__

}  // namespace fcfv
