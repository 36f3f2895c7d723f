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
// a5/a6 element matrix entries.  ElemAsm = what one element contributes, loaded once per
// (face row, adjacent element).
// ---------------------------------------------------------------------------------------------
struct ElemAsm {
    int g[3];        // global face ids (0-based)
    int bc[3];       // face tags
    double tau[3];   // face stabilisation
    double G[3], nx[3], ny[3];
    double eta, ia;  // viscosity, 1/alpha
    double iO;       // 1/Omega                (old loop)
    double D11, D12, D21, D22;                // (NEW loop)
    double b1, b2, Z11, Z21, Z12, Z22;
};

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

FCFV_HD void rheology_new(double ke, double de, double Om, double& D11, double& D12, double& D21, double& D22) {
    if (de == 1.0) {
        D11 = (ke * 1.0) / Om;
        D12 = ((ke / de) * 1.0) / Om;
        D21 = D12;
        D22 = D11;
    } else {
        double M[4] = {ke, ke / de, ke / de, ke}, Dv[4], R[4];
        inv2(M, Dv);
        for (int q = 0; q < 4; q++) Dv[q] = Dv[q] * Om;
        inv2(Dv, R);
        D11 = R[0]; D21 = R[1]; D12 = R[2]; D22 = R[3];
    }
}

// The four values the reference computes for loop indices (i, j) of one element, named by the
// normal product they carry:  vA ~ c*nix*njx (u1u1), vB ~ c*niy*njx, vC ~ c*nix*njy, vD ~ c*niy*njy.
// Old loop :146-154, NEW :273-276 / :288-291.  mA/mD: the Muu values (:157-158 / :295-296).
template <int VARIANT, int FORM, bool WANT_M>
FCFV_HD void kvals(const ElemAsm& E, int i, int j, double A, double& vA, double& vB, double& vC, double& vD,
                   double& mA, double& mD) {
    const bool on = (E.bc[i] != 1) & (E.bc[j] != 1);
    const double mG = bmul(on, -E.G[i]);
    const double nix = E.nx[i], niy = E.ny[i], njx = E.nx[j], njy = E.ny[j], Gj = E.G[j];
    const double ninj = nix * njx + niy * njy;
    const double c = (FORM == kGradient) ? (0.0 + bmul(E.bc[i] == -1, 1.0)) : A;
    const double t1 = ((E.ia * E.tau[i]) * E.tau[j]) * Gj;
    if (VARIANT == kLoopOld) {
        const double t3 = E.tau[i] * (0.0 + bmul(i == j, 1.0));
        const double hG = (E.eta * E.iO) * Gj;
        const double mhG = ((-E.eta) * E.iO) * Gj;
        vA = mG * ((t1 - hG * (ninj + (c * nix) * njx)) - t3);
        vB = mG * (mhG * ((c * niy) * njx));
        vC = mG * (mhG * ((c * nix) * njy));
        vD = mG * ((t1 - hG * (ninj + (c * niy) * njy)) - t3);
        if (WANT_M) { mA = mG * ((t1 - hG * ninj) - t3); mD = mA; }
    } else {
        const double t3 = bmul(i == j, E.tau[i]);
        vA = mG * ((t1 - (E.D11 * Gj) * (ninj + (c * nix) * njx)) - t3);
        vD = mG * ((t1 - (E.D22 * Gj) * (ninj + (c * niy) * njy)) - t3);
        if (FORM == kGradient) {   // slots [j,i+nf] (D12, nix*njy) and [j+nf,i] (D21, niy*njx)
            vC = mG * (((-E.D12) * Gj) * ((c * nix) * njy));
            vB = mG * (((-E.D21) * Gj) * ((c * niy) * njx));
        } else {                   // slots [i,j+nf] (D12, niy*njx) and [i+nf,j] (D21, nix*njy)
            vB = mG * (((-E.D12) * Gj) * ((c * niy) * njx));
            vC = mG * (((-E.D21) * Gj) * ((c * nix) * njy));
        }
        if (WANT_M) { mA = mG * ((t1 - (E.D11 * Gj) * ninj) - t3); mD = mG * ((t1 - (E.D22 * Gj) * ninj) - t3); }
    }
}

// One face row pair (rows (f,x) and (f,y)) before compaction: up to 6 candidate column faces
// (3 per adjacent element), their four values, Muu values, RHS and Kup entries.
struct FaceRows {
    int key[6];                         // column face id, kDeadKey when unused / merged away
    double xx[6], xy[6], yx[6], yy[6];  // (row x,col x) (row x,col y) (row y,col x) (row y,col y)
    double mxx[6], myy[6];
    double fx, fy;                      // fu contributions (summed over the adjacent elements)
    int pe[2];                          // Kup column (element id) per adjacent element, -1 unused
    double px[2], py[2];
};

// Contribution of adjacent element `slot` (local face i of E is the row face).
// Row/column convention: old loop and NEW-:SymmetricGradient store value(i,j) at (row i, col j);
// NEW-:Gradient stores value(i,j) at (row j, col i) (values written at [j,i], indices at [i,j],
// Discretisation:273-276 vs :299-302) and NEW stores Muu the same transposed way (:295-296).
template <int VARIANT, int FORM, bool WANT_M>
FCFV_HD void add_element_rows(const ElemAsm& E, int i, int slot, int e, double A, double VxDir_f, double VyDir_f,
                              double sxx, double syy, double sxy, double syx, FaceRows& R) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double vA, vB, vC, vD, mA = 0.0, mD = 0.0;
        double xx, xy, yx, yy, mxx = 0.0, myy = 0.0;
        if (VARIANT == kLoopNew && FORM == kGradient) {
            kvals<VARIANT, FORM, false>(E, j, i, A, vA, vB, vC, vD, mA, mD);   // loop (i=j, j=i)
            xx = vA; xy = vC; yx = vB; yy = vD;
        } else {
            kvals<VARIANT, FORM, false>(E, i, j, A, vA, vB, vC, vD, mA, mD);
            xx = vA; xy = vB; yx = vC; yy = vD;
        }
        if (WANT_M) {
            double t0, t1, t2, t3;
            if (VARIANT == kLoopNew) kvals<VARIANT, FORM, true>(E, j, i, A, t0, t1, t2, t3, mA, mD);
            else                     kvals<VARIANT, FORM, true>(E, i, j, A, t0, t1, t2, t3, mA, mD);
            mxx = mA; myy = mD;
        }
        if (j == i) {   // Dirichlet rows :187-190
            const double one = bmul(E.bc[i] == 1, 1.0);
            xx += one; yy += one;
            if (WANT_M) { mxx += one; myy += one; }
        }
        const int k = 3 * slot + j;
        R.key[k] = E.g[j];
        R.xx[k] = xx; R.xy[k] = xy; R.yx[k] = yx; R.yy[k] = yy;
        R.mxx[k] = mxx; R.myy[k] = myy;
    }
    // RHS :167-178 / :310-323
    const int bci = E.bc[i];
    const double nix = E.nx[i], niy = E.ny[i], Gi = E.G[i], ti = E.tau[i];
    const double Xi = 0.0 + bmul(bci == 2, 1.0);
    const double tix = nix * sxx + niy * sxy;
    const double tiy = nix * syx + niy * syy;
    const double mGi = bmul(bci != 1, -Gi);
    double feix, feiy;
    if (VARIANT == kLoopOld) {
        double nZx, nZy;
        if (FORM == kGradient) {
            const double J = 0.0 + bmul(bci == -1, 1.0);
            nZx = nix * (E.Z11 + J * E.Z11) + niy * (E.Z21 + J * E.Z12);
            nZy = nix * (E.Z12 + J * E.Z21) + niy * (E.Z22 + J * E.Z22);
        } else {
            nZx = nix * E.Z11 + niy * E.Z21;
            nZy = nix * E.Z12 + niy * E.Z22;
        }
        feix = mGi * (((((-E.ia) * ti) * E.b1) + (E.eta * E.iO) * nZx) - tix * Xi);
        feiy = mGi * (((((-E.ia) * ti) * E.b2) + (E.eta * E.iO) * nZy) - tiy * Xi);
    } else {
        feix = mGi * (((((-E.ia) * ti) * E.b1) + ((E.D11 * nix) * E.Z11 + (E.D11 * niy) * E.Z21)) - tix * Xi);
        feiy = mGi * (((((-E.ia) * ti) * E.b2) + ((E.D22 * nix) * E.Z12 + (E.D22 * niy) * E.Z22)) - tiy * Xi);
    }
    // fu[i,e] += (bci!=1)*feix + (bci==1)*VxDir*1e0  (from 0.0)  :191-192
    const double fxe = 0.0 + (bmul(bci != 1, feix) + bmul(bci == 1, VxDir_f) * 1.0);
    const double fye = 0.0 + (bmul(bci != 1, feiy) + bmul(bci == 1, VyDir_f) * 1.0);
    if (slot == 0) { R.fx = fxe; R.fy = fye; } else { R.fx = R.fx + fxe; R.fy = R.fy + fye; }
    // Kup :180-185
    R.pe[slot] = e;
    R.px[slot] = 0.0 - bmul(bci != 1, Gi) * nix;
    R.py[slot] = 0.0 - bmul(bci != 1, Gi) * niy;
}

FCFV_HD void face_rows_init(FaceRows& R) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
        R.key[k] = kDeadKey;
        R.xx[k] = R.xy[k] = R.yx[k] = R.yy[k] = 0.0;
        R.mxx[k] = R.myy[k] = 0.0;
    }
    R.fx = R.fy = 0.0;
    R.pe[0] = R.pe[1] = -1;
    R.px[0] = R.px[1] = R.py[0] = R.py[1] = 0.0;
}

// sparse(..) sums duplicate (row, col) pairs: the row face itself when the face has two
// elements (and any other face shared by both elements of a degenerate pair).  Input order of
// the reference = ascending element id = slot order; the second slot is folded into the first.
template <bool WANT_M>
FCFV_HD void face_rows_merge(FaceRows& R) {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 3; b < 6; b++)
            if (R.key[a] == R.key[b] && R.key[a] != kDeadKey) {
                R.xx[a] = R.xx[a] + R.xx[b]; R.xy[a] = R.xy[a] + R.xy[b];
                R.yx[a] = R.yx[a] + R.yx[b]; R.yy[a] = R.yy[a] + R.yy[b];
                if (WANT_M) { R.mxx[a] = R.mxx[a] + R.mxx[b]; R.myy[a] = R.myy[a] + R.myy[b]; }
                R.key[b] = kDeadKey;
            }
    // fu = Array(dropzeros(sparse(..))): a stored -0.0 comes back as +0.0
    if (R.fx == 0.0) R.fx = 0.0;
    if (R.fy == 0.0) R.fy = 0.0;
}

// Position of entry k among the non-zero entries of the same kind, in ascending column order.
FCFV_HD int sorted_pos(const int key[6], const bool nz[6], int k) {
    int p = 0;
#pragma unroll
    for (int m = 0; m < 6; m++) p += (m != k) & nz[m] & (key[m] < key[k]);
    return p;
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

template <int FORM>
FCFV_HD void element_residual(double Om, double eta, double dl, double alpha, double b1, double b2, double Z11,
                              double Z21, double Z12, double Z22, double sex, double sey, double Pe,
                              const double G[3], const double nx[3], const double ny[3], const int bc[3],
                              const double tau[3], const double uhx[3], const double uhy[3], const double udx[3],
                              const double udy[3], const double tix[3], const double tiy[3], ElemRes& r) {
    // pass A :12-32
    double ue[2] = {b1 / alpha, b2 / alpha};
    const double mi = -1.0 / Om;
    double L[2][2] = {{mi * Z11, mi * Z12}, {mi * Z21, mi * Z22}};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (bc[i] != 1) {
            const double n[2] = {nx[i], ny[i]}, u[2] = {uhx[i], uhy[i]};
            ue[0] = ue[0] + ((G[i] * tau[i]) * u[0]) / alpha;
            ue[1] = ue[1] + ((G[i] * tau[i]) * u[1]) / alpha;
            const double s = (1.0 / Om) * G[i];
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
    const double D[2][2] = {{eta, eta / dl}, {eta / dl, eta}};
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
        const double t[2] = {tix[i], tiy[i]};
        const double u[2] = {bc[i] == 1 ? udx[i] : uhx[i], bc[i] == 1 ? udy[i] : uhy[i]};
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
        const double un = u[0] * n[0] + u[1] * n[1];
        r.Fg2 = r.Fg2 + G[i] * un;
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                double m = n[a] * u[b];
                if (FORM == kSymmetricGradient) m = m + u[a] * n[b];
                r.F1[a][b] = r.F1[a][b] + G[i] * m;
            }
        r.F2[0] = r.F2[0] - (G[i] * tau[i]) * u[0];
        r.F2[1] = r.F2[1] - (G[i] * tau[i]) * u[1];
        r.F3 = r.F3 + G[i] * un;
    }
}

}  // namespace fcfv
