// fcfv_math.cuh -- per-element / per-face arithmetic of the FCFV hot path (FP64).
//
// Everything here is a pure function of a few values held in registers.  The CUDA kernels in
// fcfv_kernels.cuh call these from device code; tests/host_emulation.cpp compiles the same
// functions for the host so that the arithmetic (not the launch machinery) can be compared
// with the oracle on a machine without a GPU.
//
// Bitwise contract: operations are written in the evaluation order of the reference's Julia
// expressions (n-ary * and + fold left, x^-1 = 1/x then multiply, Bool*Float64 = strong zero)
// and the file must be compiled without FMA contraction (nvcc -fmad=false, gcc
// -ffp-contract=off).  The sparsity of the assembled blocks depends on exact zeros.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define FCFV_HD __host__ __device__ __forceinline__
#else
#define FCFV_HD inline
#endif

namespace fcfv {

constexpr int kGradient = 0;
constexpr int kSymmetricGradient = 1;
constexpr int kLoopOld = 0;
constexpr int kLoopNew = 1;
constexpr int kDeadKey = 0x7fffffff;
// optional third block of an assembly: none, Muu, or the Schur complement Kuu - Kup*(Kppi*Kpu)
constexpr int kExtraNone = 0;
constexpr int kExtraMuu = 1;
constexpr int kExtraSchur = 2;

// Julia: *(b::Bool, x::Float64) = ifelse(b, x, copysign(0.0, x))
FCFV_HD double bmul(bool b, double x) { return b ? x : copysign(0.0, x); }

// ---------------------------------------------------------------------------------------------
// a2 geometry of one element (src/CreateMeshFCFV.jl:142-156 / :300-314 and :180-211 / :337-367)
// ---------------------------------------------------------------------------------------------
struct ElemGeom {
    double Om, xc, yc;
    double G[3], nx[3], ny[3];
};

// x[k], y[k]: vertex k of the element (k = 0,1,2 as stored in e2v); xfm/yfm: midpoints of its faces.
FCFV_HD void element_geometry(const double x[3], const double y[3], const double xfm[3], const double yfm[3],
                              int numbering, ElemGeom& g) {
    const double a = sqrt((x[0] - x[1]) * (x[0] - x[1]) + (y[0] - y[1]) * (y[0] - y[1]));
    const double b = sqrt((x[1] - x[2]) * (x[1] - x[2]) + (y[1] - y[2]) * (y[1] - y[2]));
    const double c = sqrt((x[0] - x[2]) * (x[0] - x[2]) + (y[0] - y[2]) * (y[0] - y[2]));
    const double s = 0.5 * ((a + b) + c);
    g.Om = sqrt(((s * (s - a)) * (s - b)) * (s - c));
    g.xc = (1.0 / 3.0) * ((x[0] + x[1]) + x[2]);
    g.yc = (1.0 / 3.0) * ((y[0] + y[1]) + y[2]);
#pragma unroll
    for (int f = 0; f < 3; f++) {
        // numbering 0: nodeA=[2 3 1], nodeB=[3 1 2];  numbering 1: nodeA=[3 2 1], nodeB=[1 3 2] (1-based)
        const int va = numbering ? (2 - f) : ((f + 1) % 3);
        const int vb = numbering ? ((3 - f) % 3) : ((f + 2) % 3);
        const double dx = x[va] - x[vb];
        const double dy = y[va] - y[vb];
        const double Gi = sqrt(dx * dx + dy * dy);
        const double nx = -dy / Gi;
        const double ny = dx / Gi;
        const double vx = xfm[f] - g.xc;
        const double vy = yfm[f] - g.yc;
        const double dot = nx * vx + ny * vy;
        g.nx[f] = bmul(dot >= 0.0, nx) - bmul(dot < 0.0, nx);
        g.ny[f] = bmul(dot >= 0.0, ny) - bmul(dot < 0.0, ny);
        g.G[f] = Gi;
    }
}

// ---------------------------------------------------------------------------------------------
// a4 ComputeFCFV for one element (src/DiscretisationFCFV_Stokes.jl:15-41)
// Non-Dirichlet faces add (+-0.0) to accumulators that start at +0.0, which never changes them
// (finite data), so they are skipped and their Dirichlet values need not be loaded.
// ---------------------------------------------------------------------------------------------
struct ElemLocal {
    double alpha, b1, b2, Z11, Z21, Z12, Z22;
};

template <int FORM>
FCFV_HD void element_local(double Om, double te, double sex, double sey, const double G[3], const double nx[3],
                           const double ny[3], const bool dir[3], const double vx[3], const double vy[3], double A,
                           ElemLocal& o) {
    o.alpha = 0.0; o.Z11 = 0.0; o.Z21 = 0.0; o.Z12 = 0.0; o.Z22 = 0.0;
    o.b1 = 0.0 + Om * sex;
    o.b2 = 0.0 + Om * sey;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (dir[i]) {
            const double dG = G[i];
            if (FORM == kGradient) {
                o.Z11 += (dG * nx[i]) * vx[i];
                o.Z12 += (dG * nx[i]) * vy[i];
                o.Z21 += (dG * ny[i]) * vx[i];
                o.Z22 += (dG * ny[i]) * vy[i];
            } else {
                o.Z11 += dG * (nx[i] * vx[i] + (A * nx[i]) * vx[i]);
                o.Z12 += dG * (nx[i] * vy[i] + (A * ny[i]) * vx[i]);
                o.Z21 += dG * (ny[i] * vx[i] + (A * nx[i]) * vy[i]);
                o.Z22 += dG * (ny[i] * vy[i] + (A * ny[i]) * vy[i]);
            }
            o.b1 += (dG * te) * vx[i];
            o.b2 += (dG * te) * vy[i];
        }
        o.alpha += G[i] * te;
    }
}

// ---------------------------------------------------------------------------------------------
// a5/a6/a7 assembly, lane formulation.
//
// The rows (f,x) and (f,y) of a face f are produced by FOUR lanes: (slot, comp) with slot = which
// of the <=2 adjacent elements, comp = which velocity component the row belongs to.  A lane loads
// ONE element and computes, for that element's three faces g_j, the two values of row (f,comp):
//   vs[j] -> column (g_j, comp)      ("same" component block, carries alpha^-1 tau tau and tau delta)
//   vc[j] -> column (g_j, other)     ("cross" component block)
// plus its share of fu[(f,comp)] and its Kup[(f,comp), e] entry.  The two slot lanes of a row then
// merge: the column (f,.) appears in both elements and is summed (sparse() sums duplicates, lower
// element first), columns are ranked across both lanes, exact zeros are dropped (dropzeros).
// set_mesh rejects meshes in which two elements share more than one face, so the row face is the
// only duplicate column.
// ---------------------------------------------------------------------------------------------

// D of ElementAssemblyLoopNEW (:231-236).  delta == 1: ([ke ke/δ; ke/δ ke] .* 1.0)/Ω;
// otherwise inv(inv(M) .* Ω) with 2x2 partial-pivoting inverses (parity 1e-12 only).
FCFV_HD void inv2(const double M[4], double R[4]) {
    const double a = M[0], c = M[1], b = M[2], d = M[3];
    if (fabs(c) > fabs(a)) {
        const double l = a / c, u22 = b - l * d;
        const double i11 = 1.0 / c, i22 = 1.0 / u22, i12 = -(d * i22) * i11;
        const double x11 = i11 - i12 * l, x12 = i12, x21 = -i22 * l, x22 = i22;
        R[0] = x12; R[1] = x22; R[2] = x11; R[3] = x21;
    } else {
        const double l = c / a, u22 = d - l * b;
        const double i11 = 1.0 / a, i22 = 1.0 / u22, i12 = -(b * i22) * i11;
        R[0] = i11 - i12 * l; R[1] = -i22 * l; R[2] = i12; R[3] = i22;
    }
}

// delta != 1 branch (:234-235): kept out of line so that the common path stays small
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
void rheology_general(double ke, double de, double Om, double* D /* D11 D21 D12 D22 */) {
    double M[4] = {ke, ke / de, ke / de, ke}, Dv[4];
    inv2(M, Dv);
    for (int q = 0; q < 4; q++) Dv[q] = Dv[q] * Om;
    inv2(Dv, D);
}

FCFV_HD void rheology_new(double ke, double de, double Om, double& D11, double& D12, double& D21, double& D22) {
    if (de == 1.0) {
        D11 = (ke * 1.0) / Om;
        D12 = D11;   // ((ke / 1.0) * 1.0) / Om: x / 1.0 == x exactly, so the same quotient
        D21 = D12;
        D22 = D11;
    } else {
        double R[4];
        rheology_general(ke, de, Om, R);
        D11 = R[0]; D21 = R[1]; D12 = R[2]; D22 = R[3];
    }
}

// Per-element coefficients of the assembly, formed once per element and assembly (k_elem_coef)
// instead of once per adjacent face: 1/alpha and, old loop, ηe*Ωe^-1 (Discretisation:146) or,
// NEW loop, the rheology matrix D (:231-236).
template <int VARIANT>
FCFV_HD void element_coef(double alpha, double ke, double de, double Om, double& ia, double& c11, double& c12,
                          double& c21, double& c22) {
    ia = 1.0 / alpha;
    if (VARIANT == kLoopOld) {
        c11 = ke * (1.0 / Om);
        c12 = c21 = c22 = 0.0;
    } else {
        rheology_new(ke, de, Om, c11, c12, c21, c22);
    }
}

// register-friendly selection a[i] for a runtime i in 0..2 (no local-memory indexing)
template <typename T>
FCFV_HD T pick3(int i, const T a[3]) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }

// What one lane loads of its element (comp-relative: nr = normal component of the row's velocity
// component, no = the other one).
struct LaneElem {
    int g[3], bc[3];
    double tau[3], G[3], nr[3], no[3];
    double ia;             // 1/alpha
    double cs;             // (gamma*ke)/Omega of the Powell-Hestenes Schur complement (kExtraSchur only)
    double c;              // old loop: eta * (1/Omega)
    double Drr, Dro;       // D[r,r], D[r,o] (NEW loop)
    double br;             // beta[e, comp]
    double Z1r, Z2r;       // Z[e,1,comp], Z[e,2,comp]
    double Zr1, Zr2;       // Z[e,comp,1], Z[e,comp,2]  (old :Gradient interface terms only)
};

struct LaneRow {
    int key[3];            // column face ids, kDeadKey when dropped / merged into the partner
    double vs[3], vc[3];   // same / cross component values
    double ms[3];          // Muu values (same component block only)
    double ts[3], tc[3];   // terms of Kup*(Kppi*Kpu) in the same / cross component block (kExtraSchur)
    double fu, kp;         // RHS share, Kup value
    int e, i;              // element, local index of the row face in it
};

// value of loop pair (p = i_loop, q = j_loop) of the reference: "same" = u1u1/u2u2 entry of component
// r, "cross" = the off-diagonal-block entry this lane's row owns.  TRANSPOSED = NEW :Gradient storage.
template <int VARIANT, int FORM, bool TRANSPOSED>
FCFV_HD void pair_values(const LaneElem& E, int p, int q, int comp, double A, double& same, double& cross, double& msame) {
    const int bcp = pick3(p, E.bc), bcq = pick3(q, E.bc);
    const double Gp = pick3(p, E.G), Gq = pick3(q, E.G);
    const double npr = pick3(p, E.nr), npo = pick3(p, E.no), nqr = pick3(q, E.nr), nqo = pick3(q, E.no);
    const double taup = pick3(p, E.tau), tauq = pick3(q, E.tau);
    const bool on = (bcp != 1) & (bcq != 1);
    const double mG = bmul(on, -Gp);
    // ninj = ni_x*nj_x + ni_y*nj_y (x product first; a+b == b+a bitwise)
    const double ninj = comp == 0 ? (npr * nqr + npo * nqo) : (npo * nqo + npr * nqr);
    const double c = (FORM == kGradient) ? (0.0 + bmul(bcp == -1, 1.0)) : A;
    const double t1 = ((E.ia * taup) * tauq) * Gq;
    const double cr = TRANSPOSED ? (c * npr) * nqo : (c * npo) * nqr;
    if (VARIANT == kLoopOld) {
        const double t3 = taup * (0.0 + bmul(p == q, 1.0));
        const double hG = E.c * Gq;          // (ηe*Ωe^-1)*Γj
        const double mhG = (-E.c) * Gq;      // ((-ηe)*Ωe^-1)*Γj: rounding is sign-symmetric
        same = mG * ((t1 - hG * (ninj + (c * npr) * nqr)) - t3);
        cross = mG * (mhG * cr);
        msame = mG * ((t1 - hG * ninj) - t3);
    } else {
        const double t3 = bmul(p == q, taup);
        same = mG * ((t1 - (E.Drr * Gq) * (ninj + (c * npr) * nqr)) - t3);
        cross = mG * (((-E.Dro) * Gq) * cr);
        msame = mG * ((t1 - (E.Drr * Gq) * ninj) - t3);
    }
}

// Row (f, comp) contributions of one element.  i = local index of f in the element.
// vDir = Dirichlet value of component comp on f; Sr1, Sr2 = Neumann stress row of component comp
// (comp x: sxx, sxy; comp y: syx, syy) -- only read by the caller when the face is Neumann.
template <int VARIANT, int FORM, int EXTRA>
FCFV_HD void lane_row(const LaneElem& E, int e, int i, int comp, double A, double vDir, double Sr1, double Sr2,
                      LaneRow& R) {
    constexpr bool WANT_M = EXTRA == kExtraMuu;
    R.e = e; R.i = i;
    // Kup[(f,comp), e] (:180-185)
    const double kp_row = 0.0 - bmul(pick3(i, E.bc) != 1, pick3(i, E.G)) * pick3(i, E.nr);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double same, cross, msame, t0, t1;
        if (VARIANT == kLoopNew && FORM == kGradient) pair_values<VARIANT, FORM, true>(E, j, i, comp, A, same, cross, msame);
        else                                          pair_values<VARIANT, FORM, false>(E, i, j, comp, A, same, cross, msame);
        if (WANT_M && VARIANT == kLoopNew && FORM != kGradient)   // NEW stores Muu transposed in both formulations
            pair_values<VARIANT, FORM, true>(E, j, i, comp, A, t0, t1, msame);
        if (j == i) {   // Dirichlet rows :187-190
            const double one = bmul(pick3(i, E.bc) == 1, 1.0);
            same += one;
            msame += one;
        }
        R.key[j] = E.g[j];
        R.vs[j] = same; R.vc[j] = cross; R.ms[j] = WANT_M ? msame : 0.0;
        if (EXTRA == kExtraSchur) {
            // (Kup*(Kppi*Kpu))[(f,comp),(g_j,b)] = Kup[(f,comp),e] * (coef_e * Kpu[e,(g_j,b)]), Kpu = -Kup'
            // (SolversFCFV_Stokes.jl:6,23-25); one term per element
            const double Gj = bmul(E.bc[j] != 1, E.G[j]);
            const double kps = 0.0 - Gj * E.nr[j], kpc = 0.0 - Gj * E.no[j];
            R.ts[j] = kp_row * (E.cs * (-kps));
            R.tc[j] = kp_row * (E.cs * (-kpc));
        } else {
            R.ts[j] = R.tc[j] = 0.0;
        }
    }
    // RHS :167-178 / :310-323
    const int bci = pick3(i, E.bc);
    const double nir = pick3(i, E.nr), nio = pick3(i, E.no), Gi = pick3(i, E.G), ti = pick3(i, E.tau);
    const double nix = comp == 0 ? nir : nio, niy = comp == 0 ? nio : nir;
    const double Xi = 0.0 + bmul(bci == 2, 1.0);
    const double tir = nix * Sr1 + niy * Sr2;
    const double mGi = bmul(bci != 1, -Gi);
    double fe;
    if (VARIANT == kLoopOld) {
        double nZ;
        if (FORM == kGradient) {
            const double J = 0.0 + bmul(bci == -1, 1.0);
            nZ = nix * (E.Z1r + J * E.Zr1) + niy * (E.Z2r + J * E.Zr2);
        } else {
            nZ = nix * E.Z1r + niy * E.Z2r;
        }
        fe = mGi * (((((-E.ia) * ti) * E.br) + E.c * nZ) - tir * Xi);
    } else {
        fe = mGi * (((((-E.ia) * ti) * E.br) + ((E.Drr * nix) * E.Z1r + (E.Drr * niy) * E.Z2r)) - tir * Xi);
    }
    R.fu = 0.0 + (bmul(bci != 1, fe) + bmul(bci == 1, vDir) * 1.0);   // :191-192
    R.kp = kp_row;
}

// ---------------------------------------------------------------------------------------------
// Rows (f,x) and (f,y) of one face, merged.  With the row face moved to the front of each element record, so that it
// sits at local index 0, the <=5 distinct column faces are: candidate 0 = f itself (both elements
// contribute, summed lower element first like sparse() does), 1-2 = the other two faces of the
// lower element, 3-4 = those of the higher element.  Placement works in RANK space: rank(c) =
// number of candidate keys below key[c]; a row part is described by a 5-bit mask with bit rank(c)
// set when candidate c is stored (alive and != +-0.0, dropzeros), and the position of c inside the
// part is popc(mask & lt[c]) with lt[c] = (1 << rank(c)) - 1.
// ---------------------------------------------------------------------------------------------
struct FaceRows {          // what stays in registers of the rows of one face
    int key[5];
    int lt[5];
    double x0s, x0c, y0s, y0c, m0x, m0y;   // candidate 0 (the row face's own column) while the two elements are summed
    double t0xs, t0xc, t0ys, t0yc;           // same for the terms of Kup*(Kppi*Kpu) (kExtraSchur)
    double fux, fuy;
    int kpe[2];                                // Kup columns: lower / higher element
    int m_xs, m_xc, m_ys, m_yc;               // rank-space masks of stored Kuu entries
    int m_ex[4];                              // third block: Muu {x, y, -, -} or Schur {xs, xc, ys, yc}
    int m_kp;                                  // bit 0 kpx[0], 1 kpx[1], 2 kpy[0], 3 kpy[1]
};

// The values of one candidate column: xs/xc = row (f,x) in the x-column block ("same") / y-column
// block ("cross"), ys/yc = row (f,y) in the y-column / x-column block; ex = third block (Muu: x, y
// rows; Schur complement: xs, xc, ys, yc).
struct CandVals { double xs, xc, ys, yc, ex[4]; };

// They leave asm_face through a sink as soon as they are final: the count pass drops them, the fill
// pass and the host emulation keep them in a RowVals.
struct RowVals {
    double xs[5], xc[5], ys[5], yc[5], ex[4][5];
    double kpx[2], kpy[2];
};

struct NullSink {
    static constexpr bool kRanked = false;     // nothing is placed: the column ranks are not needed
    FCFV_HD void cand(int, const CandVals&) const {}
    FCFV_HD void kup(int, double, double) const {}
};

struct RowValsSink {
    static constexpr bool kRanked = true;
    RowVals* V;
    FCFV_HD void cand(int c, const CandVals& v) const {
        V->xs[c] = v.xs; V->xc[c] = v.xc; V->ys[c] = v.ys; V->yc[c] = v.yc;
#pragma unroll
        for (int q = 0; q < 4; q++) V->ex[q][c] = v.ex[q];
    }
    FCFV_HD void kup(int s, double kx, double ky) const { V->kpx[s] = kx; V->kpy[s] = ky; }
};

FCFV_HD void rows_clear(FaceRows& R) {
#pragma unroll
    for (int c = 0; c < 5; c++) { R.key[c] = kDeadKey; R.lt[c] = 0; }
    R.x0s = R.x0c = R.y0s = R.y0c = R.m0x = R.m0y = 0.0;
    R.t0xs = R.t0xc = R.t0ys = R.t0yc = 0.0;
    R.fux = R.fuy = 0.0;
    R.kpe[0] = R.kpe[1] = -1;
    R.m_xs = R.m_xc = R.m_ys = R.m_yc = R.m_kp = 0;
    R.m_ex[0] = R.m_ex[1] = R.m_ex[2] = R.m_ex[3] = 0;
}

// Column faces and their ranks are known before any element data arrives (face connectivity
// record): keys k[0..4], absent faces (< 0) rank last.  Masks start empty and are filled while the
// values are produced.
// RANKED = false (count pass): only the number of stored entries matters, any five distinct bits do.
template <bool RANKED>
FCFV_HD void rows_begin(FaceRows& R, const int k[5]) {
#pragma unroll
    for (int c = 0; c < 5; c++) R.key[c] = k[c] < 0 ? kDeadKey : k[c];
    // ranks from the 10 pairwise comparisons (ties only between dead keys: broken by index)
    int rks[5] = {0, 1, 2, 3, 4};
    if (RANKED) {
#pragma unroll
        for (int c = 0; c < 5; c++) rks[c] = 0;
#pragma unroll
        for (int c = 0; c < 5; c++)
#pragma unroll
            for (int d = c + 1; d < 5; d++) {
                const int t = (R.key[d] < R.key[c]) ? 1 : 0;
                rks[c] += t;
                rks[d] += 1 - t;
            }
    }
#pragma unroll
    for (int c = 0; c < 5; c++) R.lt[c] = (1 << rks[c]) - 1;
    R.kpe[0] = R.kpe[1] = -1;
    R.m_xs = R.m_xc = R.m_ys = R.m_yc = R.m_kp = 0;
    R.m_ex[0] = R.m_ex[1] = R.m_ex[2] = R.m_ex[3] = 0;
}

// marks the stored (alive and != +-0.0: dropzeros) entries of candidate c
template <int EXTRA>
FCFV_HD void rows_mark(FaceRows& R, int c, bool alive, const CandVals& v) {
    const int bit = R.lt[c] + 1;
    R.m_xs |= (alive && v.xs != 0.0) ? bit : 0;
    R.m_xc |= (alive && v.xc != 0.0) ? bit : 0;
    R.m_ys |= (alive && v.ys != 0.0) ? bit : 0;
    R.m_yc |= (alive && v.yc != 0.0) ? bit : 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (EXTRA == kExtraSchur || (EXTRA == kExtraMuu && q < 2)) R.m_ex[q] |= (alive && v.ex[q] != 0.0) ? bit : 0;
}

// candidate values of faces j = 1, 2 of an element (single contribution).  Schur complement:
// Kuu .- C keeps Kuu[i,j] - C[i,j] wherever either is stored and drops zeros (sparse broadcast);
// a missing operand is +0.0, which a - c reproduces.
template <int EXTRA>
FCFV_HD void cand_of(const LaneRow& Lx, const LaneRow& Ly, int j, CandVals& v) {
    v.xs = Lx.vs[j]; v.xc = Lx.vc[j]; v.ys = Ly.vs[j]; v.yc = Ly.vc[j];
    v.ex[0] = v.ex[1] = v.ex[2] = v.ex[3] = 0.0;
    if (EXTRA == kExtraMuu) { v.ex[0] = Lx.ms[j]; v.ex[1] = Ly.ms[j]; }
    if (EXTRA == kExtraSchur) {
        v.ex[0] = Lx.vs[j] - Lx.ts[j]; v.ex[1] = Lx.vc[j] - Lx.tc[j];
        v.ex[2] = Ly.vs[j] - Ly.ts[j]; v.ex[3] = Ly.vc[j] - Ly.tc[j];
    }
}

// Lanes (comp x, comp y) of the lower element: candidates 0..2, RHS, first Kup entry.
template <int EXTRA, typename Sink>
FCFV_HD void rows_put0(FaceRows& R, bool alive, const LaneRow& Lx, const LaneRow& Ly, const Sink& sink) {
    R.x0s = Lx.vs[0]; R.x0c = Lx.vc[0]; R.y0s = Ly.vs[0]; R.y0c = Ly.vc[0];
    R.m0x = EXTRA == kExtraMuu ? Lx.ms[0] : 0.0; R.m0y = EXTRA == kExtraMuu ? Ly.ms[0] : 0.0;
    R.t0xs = Lx.ts[0]; R.t0xc = Lx.tc[0]; R.t0ys = Ly.ts[0]; R.t0yc = Ly.tc[0];
#pragma unroll
    for (int j = 1; j < 3; j++) {
        CandVals v;
        cand_of<EXTRA>(Lx, Ly, j, v);
        rows_mark<EXTRA>(R, j, alive, v);
        sink.cand(j, v);
    }
    R.fux = Lx.fu; R.fuy = Ly.fu;
    R.kpe[0] = Lx.e;
    R.m_kp |= ((alive && Lx.kp != 0.0) ? 1 : 0) | ((alive && Ly.kp != 0.0) ? 4 : 0);
    sink.kup(0, Lx.kp, Ly.kp);
}

// Lanes of the higher element: its row-face column is summed into candidate 0 (lower element
// first, like sparse() and the sparse product do), its other two faces become candidates 3..4.
// With on == false (the face has no second element; the lanes then hold a discarded re-read of
// the first) nothing is added.
template <int EXTRA, typename Sink>
FCFV_HD void rows_put1(FaceRows& R, bool on, bool alive, const LaneRow& Lx, const LaneRow& Ly, const Sink& sink) {
    R.x0s = on ? R.x0s + Lx.vs[0] : R.x0s; R.x0c = on ? R.x0c + Lx.vc[0] : R.x0c;
    R.y0s = on ? R.y0s + Ly.vs[0] : R.y0s; R.y0c = on ? R.y0c + Ly.vc[0] : R.y0c;
    if (EXTRA == kExtraMuu) { R.m0x = on ? R.m0x + Lx.ms[0] : R.m0x; R.m0y = on ? R.m0y + Ly.ms[0] : R.m0y; }
    if (EXTRA == kExtraSchur) {
        R.t0xs = on ? R.t0xs + Lx.ts[0] : R.t0xs; R.t0xc = on ? R.t0xc + Lx.tc[0] : R.t0xc;
        R.t0ys = on ? R.t0ys + Ly.ts[0] : R.t0ys; R.t0yc = on ? R.t0yc + Ly.tc[0] : R.t0yc;
    }
#pragma unroll
    for (int j = 1; j < 3; j++) {
        CandVals v;
        cand_of<EXTRA>(Lx, Ly, j, v);
        rows_mark<EXTRA>(R, 2 + j, on && alive, v);
        sink.cand(2 + j, v);
    }
    R.fux = on ? R.fux + Lx.fu : R.fux; R.fuy = on ? R.fuy + Ly.fu : R.fuy;
    R.kpe[1] = Lx.e;
    R.m_kp |= ((on && alive && Lx.kp != 0.0) ? 2 : 0) | ((on && alive && Ly.kp != 0.0) ? 8 : 0);
    sink.kup(1, Lx.kp, Ly.kp);
}

// merged row-face column and RHS; alive: the face's rows are produced (owned face with an element)
template <int EXTRA, typename Sink>
FCFV_HD void rows_finish(FaceRows& R, bool alive, const Sink& sink) {
    if (R.fux == 0.0) R.fux = 0.0;   // Array(dropzeros(sparse)): -0.0 -> +0.0
    if (R.fuy == 0.0) R.fuy = 0.0;
    CandVals v;
    v.xs = R.x0s; v.xc = R.x0c; v.ys = R.y0s; v.yc = R.y0c;
    v.ex[0] = v.ex[1] = v.ex[2] = v.ex[3] = 0.0;
    if (EXTRA == kExtraMuu) { v.ex[0] = R.m0x; v.ex[1] = R.m0y; }
    if (EXTRA == kExtraSchur) { v.ex[0] = R.x0s - R.t0xs; v.ex[1] = R.x0c - R.t0xc; v.ex[2] = R.y0s - R.t0ys; v.ex[3] = R.y0c - R.t0yc; }
    rows_mark<EXTRA>(R, 0, alive, v);
    sink.cand(0, v);
}

FCFV_HD int popc6(int x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// stored entries of the rows: Kuu x / y, Kup x / y, Muu x / y
FCFV_HD int rows_nuu_x(const FaceRows& R) { return popc6(R.m_xs | (R.m_xc << 5)); }
FCFV_HD int rows_nuu_y(const FaceRows& R) { return popc6(R.m_ys | (R.m_yc << 5)); }
FCFV_HD int rows_nup_x(const FaceRows& R) { return (R.m_kp & 1) + ((R.m_kp >> 1) & 1); }
FCFV_HD int rows_nup_y(const FaceRows& R) { return ((R.m_kp >> 2) & 1) + ((R.m_kp >> 3) & 1); }
// third block: Muu rows hold one part, Schur rows two
template <int EXTRA>
FCFV_HD int rows_nex_x(const FaceRows& R) { return EXTRA == kExtraSchur ? popc6(R.m_ex[0] | (R.m_ex[1] << 5)) : popc6(R.m_ex[0]); }
template <int EXTRA>
FCFV_HD int rows_nex_y(const FaceRows& R) { return EXTRA == kExtraSchur ? popc6(R.m_ex[2] | (R.m_ex[3] << 5)) : popc6(R.m_ex[1]); }

// emit(position inside the row, column, value); rows are laid out [x columns | y columns].
// Positions are formed unconditionally so that the conditional part is just the store.
template <typename Emit>
FCFV_HD void rows_emit_uu_x(const FaceRows& R, const double xs[5], const double xc[5], int nf, Emit&& emit) {
    const int n_xs = popc6(R.m_xs);
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const int lt = R.lt[c], bit = lt + 1;
        const int p_xs = popc6(R.m_xs & lt), p_xc = n_xs + popc6(R.m_xc & lt);
        if (R.m_xs & bit) emit(p_xs, R.key[c], xs[c]);
        if (R.m_xc & bit) emit(p_xc, R.key[c] + nf, xc[c]);
    }
}

template <typename Emit>
FCFV_HD void rows_emit_uu_y(const FaceRows& R, const double ys[5], const double yc[5], int nf, Emit&& emit) {
    const int n_yc = popc6(R.m_yc);
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const int lt = R.lt[c], bit = lt + 1;
        const int p_yc = popc6(R.m_yc & lt), p_ys = n_yc + popc6(R.m_ys & lt);
        if (R.m_yc & bit) emit(p_yc, R.key[c], yc[c]);
        if (R.m_ys & bit) emit(p_ys, R.key[c] + nf, ys[c]);
    }
}

template <typename Emit>
FCFV_HD void rows_emit_uu(const FaceRows& R, const RowVals& V, int nf, Emit&& emit) {
    rows_emit_uu_x(R, V.xs, V.xc, nf, [&](int pos, int col, double v) { emit(0, pos, col, v); });
    rows_emit_uu_y(R, V.ys, V.yc, nf, [&](int pos, int col, double v) { emit(1, pos, col, v); });
}

// third block.  Muu: row (f,comp) only holds columns (g,comp).  Schur complement: laid out like Kuu.
template <int EXTRA, typename Emit>
FCFV_HD void rows_emit_ex(const FaceRows& R, const RowVals& V, int nf, Emit&& emit) {
    if (EXTRA == kExtraMuu) {
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const int lt = R.lt[c], bit = lt + 1;
            const int p_x = popc6(R.m_ex[0] & lt), p_y = popc6(R.m_ex[1] & lt);
            if (R.m_ex[0] & bit) emit(0, p_x, R.key[c], V.ex[0][c]);
            if (R.m_ex[1] & bit) emit(1, p_y, R.key[c] + nf, V.ex[1][c]);
        }
    } else {
        const int n_xs = popc6(R.m_ex[0]), n_yc = popc6(R.m_ex[3]);
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const int lt = R.lt[c], bit = lt + 1;
            const int p_xs = popc6(R.m_ex[0] & lt), p_xc = n_xs + popc6(R.m_ex[1] & lt);
            const int p_yc = popc6(R.m_ex[3] & lt), p_ys = n_yc + popc6(R.m_ex[2] & lt);
            if (R.m_ex[0] & bit) emit(0, p_xs, R.key[c], V.ex[0][c]);
            if (R.m_ex[1] & bit) emit(0, p_xc, R.key[c] + nf, V.ex[1][c]);
            if (R.m_ex[3] & bit) emit(1, p_yc, R.key[c], V.ex[3][c]);
            if (R.m_ex[2] & bit) emit(1, p_ys, R.key[c] + nf, V.ex[2][c]);
        }
    }
}

// columns = element ids, lower element first
template <typename Emit>
FCFV_HD void rows_emit_up(const FaceRows& R, const double kpx[2], const double kpy[2], Emit&& emit) {
    if (R.m_kp & 1) emit(0, 0, R.kpe[0], kpx[0]);
    if (R.m_kp & 2) emit(0, R.m_kp & 1, R.kpe[1], kpx[1]);
    if (R.m_kp & 4) emit(1, 0, R.kpe[0], kpy[0]);
    if (R.m_kp & 8) emit(1, (R.m_kp >> 2) & 1, R.kpe[1], kpy[1]);
}

// fp[e] :194 -- sequential over the three faces, Dirichlet faces only contribute.
FCFV_HD double element_fp(const double G[3], const double nx[3], const double ny[3], const bool dir[3],
                          const double vx[3], const double vy[3]) {
    double fp = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
        if (dir[i]) fp = fp - G[i] * (vx[i] * nx[i] + vy[i] * ny[i]);
    return fp;
}

// ---------------------------------------------------------------------------------------------
// a8 residuals of one element (src/ComputeResidualsFCFV_Stokes.jl:12-84).
// Parity is 1e-12 relative (the reference goes through BLAS for the small products).
// ---------------------------------------------------------------------------------------------
struct ElemRes {
    double Fg1[3][2];   // F_glob1[e,i,:]
    double Fg2;         // F_glob2[e]
    double F1[2][2];    // F_eq1[e,:,:]
    double F2[2];       // F_eq2[e,:]
    double F3;          // F_eq3[e]
};

// traction(i, tx, ty): t = n * sigma_Neu of face i (:59-61), only called for Neumann faces (it enters through
// t*Xi, :69) -- evaluated where it is used so that it occupies no registers on the other faces
template <int FORM, typename Traction>
FCFV_HD void element_residual(double Om, double eta, double dl, double alpha, double b1, double b2, double Z11,
                              double Z21, double Z12, double Z22, double sex, double sey, double Pe,
                              const double G[3], const double nx[3], const double ny[3], const int bc[3],
                              const double tau[3], const double ux[3], const double uy[3], Traction&& traction,
                              ElemRes& r) {
    // ux, uy: the face velocity the reference uses on face i -- VDir on Dirichlet faces (:62-66), the hybrid
    // unknown elsewhere (pass A only visits non-Dirichlet faces, :17)
    // pass A :12-32.  The reference divides by alpha eight times (:14,19); here one reciprocal and eight products: each
    // term moves by at most an ulp, far inside the 1e-12 the residual values are held to (unlike the assembly, nothing
    // downstream depends on exact zeros), and the element kernel loses a fifth of its instructions.
    const double ia = 1.0 / alpha;
    double ue[2] = {b1 * ia, b2 * ia};
    const double iOm = 1.0 / Om;          // -1.0/Ω == -(1.0/Ω) bit for bit: one division serves :14 and :22
    const double mi = -iOm;
    double L[2][2] = {{mi * Z11, mi * Z12}, {mi * Z21, mi * Z22}};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (bc[i] != 1) {
            const double n[2] = {nx[i], ny[i]}, u[2] = {ux[i], uy[i]};
            ue[0] = ue[0] + ((G[i] * tau[i]) * u[0]) * ia;
            ue[1] = ue[1] + ((G[i] * tau[i]) * u[1]) * ia;
            const double s = iOm * G[i];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    double m = n[a] * u[b];
                    if (FORM == kSymmetricGradient) m = m + u[a] * n[b];
                    L[a][b] = L[a][b] - s * m;
                }
        }
    }
    // pass B :42-84
    const double ed = dl == 1.0 ? eta : eta / dl;      // x / 1.0 == x: the isotropic case skips the division, same bits
    const double D[2][2] = {{eta, ed}, {ed, eta}};
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) r.F1[a][b] = Om * L[a][b];
    r.F2[0] = -sex * Om;
    r.F2[1] = -sey * Om;
    r.F3 = 0.0;
    r.Fg2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double n[2] = {nx[i], ny[i]};
        const double Xi = 0.0 + bmul(bc[i] == 2, 1.0);
        const double Ji = 0.0 + bmul(bc[i] == -1, 1.0);
        double t[2] = {0.0, 0.0};
        if (bc[i] == 2) traction(i, t[0], t[1]);
        const double u[2] = {ux[i], uy[i]};
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const double nDL = n[0] * (D[0][b] * L[0][b]) + n[1] * (D[1][b] * L[1][b]);
            double acc = (((nDL + Pe * n[b]) + tau[i] * ue[b]) - tau[i] * u[b]) + t[b] * Xi;
            if (FORM == kGradient) {
                const double nDLt = n[0] * (D[b][0] * L[b][0]) + n[1] * (D[b][1] * L[b][1]);
                acc = acc + Ji * nDLt;
            }
            r.Fg1[i][b] = G[i] * acc;
        }
        r.F2[0] = r.F2[0] + (G[i] * tau[i]) * ue[0];
        r.F2[1] = r.F2[1] + (G[i] * tau[i]) * ue[1];
        // dAi * u'*n = (dAi*u')*n (:75, :82); dAi * n * u' = (dAi*n)*u' (:77); dAi * (n*u' .+ u*n') (:79)
        const double un = (G[i] * u[0]) * n[0] + (G[i] * u[1]) * n[1];
        r.Fg2 = r.Fg2 + un;
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                if (FORM == kSymmetricGradient) r.F1[a][b] = r.F1[a][b] + G[i] * (n[a] * u[b] + u[a] * n[b]);
                else                            r.F1[a][b] = r.F1[a][b] + (G[i] * n[a]) * u[b];
            }
        r.F2[0] = r.F2[0] - (G[i] * tau[i]) * u[0];
        r.F2[1] = r.F2[1] - (G[i] * tau[i]) * u[1];
        r.F3 = r.F3 + un;
    }
}

// ---------------------------------------------------------------------------------------------
// ComputeElementValues for one element (src/DiscretisationFCFV_Stokes.jl:52-94): element velocity
// and deviatoric stress from the face velocities.  tau[i] = mesh.τ of the element's faces.
// ---------------------------------------------------------------------------------------------
struct ElemValues { double Vx, Vy, Txx, Tyy, Txy; };

FCFV_HD void element_values(double Om, double eta, double dl, double alpha, double b1, double b2, double Z11, double Z21,
                            double Z12, double Z22, const double G[3], const double nx[3], const double ny[3], const int bc[3],
                            const double tau[3], const double uhx[3], const double uhy[3], ElemValues& o) {
    const double D11 = eta, D12 = eta / dl, D22 = eta;
    o.Vx = b1 / alpha;
    o.Vy = b2 / alpha;
    o.Txx = (eta / Om) * Z11;
    o.Tyy = (eta / Om) * Z22;
    o.Txy = ((eta / Om) * 0.5) * (Z12 + Z21);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const bool nd = bc[i] != 1;
        o.Vx += ((bmul(nd, G[i]) * tau[i]) * uhx[i]) / alpha;
        o.Vy += ((bmul(nd, G[i]) * tau[i]) * uhy[i]) / alpha;
        o.Txx += (((bmul(nd, D11) / Om) * G[i]) * nx[i]) * uhx[i];
        o.Tyy += (((bmul(nd, D22) / Om) * G[i]) * ny[i]) * uhy[i];
        o.Txy += (bmul(nd, D12) * 0.5) * (((1.0 / Om) * G[i]) * (nx[i] * uhy[i] + ny[i] * uhx[i]));
    }
    o.Txx *= 2.0; o.Tyy *= 2.0; o.Txy *= 2.0;
}

}  // namespace fcfv
