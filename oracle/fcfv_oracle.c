/*
 * fcfv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Literal, single-threaded CPU restatement (plain C, FP64 + Int64) of the hot
 * path of tduretz/FCFV_NME23.  It exists only as the checker for the CUDA path
 * (tests/, __graft_entry__.smoke()) and as the "port" CPU baseline timed by
 * bench.py.  Nothing under fcfv_nme23_b200/ may call into it.
 *
 * PARITY: the reference ships no tests, golden vectors or expected values for
 * this path and Julia is not available in the build image, so this file was
 * never compared with a Julia run ("parity unpinned" in that strict sense).
 * What pins it instead: tests/julia_transliteration.py reads the reference's
 * own source text (src/DiscretisationFCFV_Stokes.jl, ComputeResidualsFCFV_
 * Stokes.jl, the geometry loops of CreateMeshFCFV.jl, meshes/ReadMeshesDat.jl),
 * rewrites it with a documented rule table and executes it with IEEE float64
 * scalars; this oracle equals that run BIT FOR BIT for ComputeFCFV, both
 * assembly loops + Sparsify, ComputeElementValues, the geometry and the face
 * numbering, and to 1e-12 for the six residual norms, on 54 cases
 * (tests/test_transliteration.py live, tests/golden/golden_reference.json
 * committed).  Still resting on restatement alone: the Powell-Hestenes lines
 * of SolversFCFV_Stokes.jl (sparse mat-vec, sparse `.-`) and whatever the rule
 * table itself gets wrong about Julia semantics.
 * It follows the Julia sources statement by statement (file:line citations
 * are relative to the reference checkout) and honours Julia's evaluation order
 * (n-ary + and * fold left, x^-1 == inv(x), Bool*Float64 is a strong zero,
 * no FMA contraction).
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared  (see oracle/Makefile).
 * -ffp-contract=off matters: Julia/LLVM never fuses a*b+c on its own, and the
 * sparsity pattern after dropzeros depends on exact floating-point zeros.
 *
 * Array conventions are Julia's: column-major, 1-based indices stored in
 * int64 arrays.  X[e,i] of an nel x 3 matrix lives at X[(i-1)*nel + (e-1)];
 * Z[e,a,b] of nel x 2 x 2 at Z[(e-1) + nel*(a-1) + 2*nel*(b-1)].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int64_t i64;

#define GRADIENT 0
#define SYMMETRIC_GRADIENT 1

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Julia: *(x::Bool, y::Float64) = ifelse(x, y, copysign(zero(y), y)) */
static inline double bmul(int b, double x) { return b ? x : copysign(0.0, x); }

/* ------------------------------------------------------------------------ */
/* a2: mesh geometry -- src/CreateMeshFCFV.jl:142-156,161-163,180-211       */
/*     (MakeTriangleMesh) and :300-314,319-320,337-367 (LoadExternalMesh2).  */
/* numbering 0: nodeA=[2 3 1], nodeB=[3 1 2]; 1: nodeA=[3 2 1], nodeB=[1 3 2] */
/* ------------------------------------------------------------------------ */
int oracle_geometry(i64 nel, i64 nf, const i64 *e2v, const i64 *e2f,
                    const double *xv, const double *yv, const double *xf,
                    const double *yf, int numbering, double tau_r,
                    double *Omega, double *xc, double *yc, double *n_x,
                    double *n_y, double *Gamma, double *tau /* nf or NULL */) {
    static const int A0[3] = {2, 3, 1}, B0[3] = {3, 1, 2};
    static const int A1[3] = {3, 2, 1}, B1[3] = {1, 3, 2};
    const int *nodeA = numbering ? A1 : A0, *nodeB = numbering ? B1 : B0;
    (void)nf;
    for (i64 e = 0; e < nel; e++) { /* :142-156 / :300-314 */
        double x1 = xv[e2v[e] - 1], y1 = yv[e2v[e] - 1];
        double x2 = xv[e2v[nel + e] - 1], y2 = yv[e2v[nel + e] - 1];
        double x3 = xv[e2v[2 * nel + e] - 1], y3 = yv[e2v[2 * nel + e] - 1];
        double a = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
        double b = sqrt((x2 - x3) * (x2 - x3) + (y2 - y3) * (y2 - y3));
        double c = sqrt((x1 - x3) * (x1 - x3) + (y1 - y3) * (y1 - y3));
        double s = 0.5 * ((a + b) + c); /* 1/2*(a+b+c): 1/2 == 0.5 */
        Omega[e] = sqrt(((s * (s - a)) * (s - b)) * (s - c));
        xc[e] = (1.0 / 3.0) * ((x1 + x2) + x3);
        yc[e] = (1.0 / 3.0) * ((y1 + y2) + y3);
    }
    for (i64 e = 0; e < nel; e++) { /* :180-211 / :337-367 */
        for (int f = 0; f < 3; f++) {
            i64 nodei = e2f[f * nel + e] - 1;
            i64 vert1 = e2v[(nodeA[f] - 1) * nel + e] - 1;
            i64 vert2 = e2v[(nodeB[f] - 1) * nel + e] - 1;
            double dx = xv[vert1] - xv[vert2];
            double dy = yv[vert1] - yv[vert2];
            double Gi = sqrt(dx * dx + dy * dy);
            double nx = -dy / Gi;
            double ny = dx / Gi;
            double vx = xf[nodei] - xc[e];
            double vy = yf[nodei] - yc[e];
            double dot = nx * vx + ny * vy;
            /* ((dot>=0.0)*n_x - (dot<0.0)*n_x) */
            n_x[f * nel + e] = bmul(dot >= 0.0, nx) - bmul(dot < 0.0, nx);
            n_y[f * nel + e] = bmul(dot >= 0.0, ny) - bmul(dot < 0.0, ny);
            Gamma[f * nel + e] = Gi;
            if (tau) tau[nodei] = tau_r; /* :209 / :365 */
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* a4: ComputeFCFV -- src/DiscretisationFCFV_Stokes.jl:8-44                  */
/* ------------------------------------------------------------------------ */
int oracle_compute_fcfv(i64 nel, i64 nf, const i64 *e2f, const i64 *bcarr,
                        const double *Omega, const double *Gamma,
                        const double *n_x, const double *n_y,
                        const double *taue, const double *sex,
                        const double *sey, const double *VxDir,
                        const double *VyDir, int formulation, double A,
                        double *alpha, double *beta, double *Z, double *tau) {
    (void)nf;
    memset(alpha, 0, sizeof(double) * nel);     /* :10 */
    memset(beta, 0, sizeof(double) * 2 * nel);  /* :11 */
    memset(Z, 0, sizeof(double) * 4 * nel);     /* :12 */
    double *b1 = beta, *b2 = beta + nel;
    double *Z11 = Z, *Z21 = Z + nel, *Z12 = Z + 2 * nel, *Z22 = Z + 3 * nel;
    for (i64 e = 0; e < nel; e++) {
        b1[e] += Omega[e] * sex[e]; /* :16 */
        b2[e] += Omega[e] * sey[e]; /* :17 */
        for (int i = 0; i < 3; i++) {
            i64 nodei = e2f[i * nel + e] - 1;
            int d = (bcarr[nodei] == 1);
            double Gi = Gamma[i * nel + e];
            double nix = n_x[i * nel + e], niy = n_y[i * nel + e];
            double te = taue[e]; /* :25 */
            double vx = VxDir[nodei], vy = VyDir[nodei];
            double dG = bmul(d, Gi); /* (bc==1) * Γi, leftmost factor */
            if (formulation == GRADIENT) { /* :27-30 */
                Z11[e] += (dG * nix) * vx;
                Z12[e] += (dG * nix) * vy;
                Z21[e] += (dG * niy) * vx;
                Z22[e] += (dG * niy) * vy;
            } else { /* :32-35 */
                Z11[e] += dG * (nix * vx + (A * nix) * vx);
                Z12[e] += dG * (nix * vy + (A * niy) * vx);
                Z21[e] += dG * (niy * vx + (A * nix) * vy);
                Z22[e] += dG * (niy * vy + (A * niy) * vy);
            }
            b1[e] += (dG * te) * vx; /* :37 */
            b2[e] += (dG * te) * vy; /* :38 */
            alpha[e] += Gi * te;     /* :39 */
            tau[nodei] = te;         /* :40  (last writer wins) */
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* a7: SparseArrays.sparse(I,J,V,m,n) (combine = +) followed by dropzeros.   */
/* Restates the stdlib algorithm: counting sort into CSR (stable, input      */
/* order), combine duplicates per row in input order, counting-sort          */
/* transpose into CSC (rows ascending inside each column), then remove        */
/* stored entries equal to +-0.0.  I and J arrive as Float64 (Discretisation */
/* :105-107,110-111) and are converted to Int.  Returns nnz, -1 on OOM.       */
/* ------------------------------------------------------------------------ */
static i64 sparse_dropzeros(i64 ncoo, const double *If, const double *Jf,
                            const double *V, i64 m, i64 n, i64 *colptr,
                            i64 *rowval, double *nzval) {
    i64 *csrcolval = (i64 *)malloc(sizeof(i64) * (size_t)(ncoo ? ncoo : 1));
    double *csrnzval = (double *)malloc(sizeof(double) * (size_t)(ncoo ? ncoo : 1));
    i64 *klast = (i64 *)calloc((size_t)n + 1, sizeof(i64));
    i64 *cnt = (i64 *)calloc((size_t)n + 2, sizeof(i64));
    i64 *start = (i64 *)calloc((size_t)m + 2, sizeof(i64));
    i64 *fill = (i64 *)malloc(sizeof(i64) * ((size_t)m + 2));
    if (!csrcolval || !csrnzval || !klast || !cnt || !start || !fill) {
        free(start); free(fill); free(csrcolval); free(csrnzval); free(klast); free(cnt);
        return -1;
    }
    /* counting sort by row (stable): start[i] = first slot of row i (1-based) */
    for (i64 k = 0; k < ncoo; k++) start[(i64)If[k]]++; /* counts at [1..m] */
    {
        i64 run = 0;
        for (i64 i = 1; i <= m; i++) { i64 c = start[i]; start[i] = run; run += c; }
        start[m + 1] = run;
    }
    memcpy(fill, start, sizeof(i64) * ((size_t)m + 2));
    for (i64 k = 0; k < ncoo; k++) {
        i64 i = (i64)If[k];
        i64 p = fill[i]++;
        csrcolval[p] = (i64)Jf[k];
        csrnzval[p] = V[k];
    }
    /* combine duplicates inside each row, in input order; count per column */
    for (i64 i = 1; i <= m; i++) {
        i64 rs = start[i], re = start[i + 1], w = rs;
        for (i64 r = rs; r < re; r++) {
            i64 j = csrcolval[r];
            if (klast[j] > rs) { /* column already seen in this row (klast stores pos+1) */
                csrnzval[klast[j] - 1] = csrnzval[klast[j] - 1] + csrnzval[r];
            } else {
                klast[j] = w + 1;
                csrcolval[w] = j;
                csrnzval[w] = csrnzval[r];
                w++;
                cnt[j]++;
            }
        }
        fill[i] = w; /* new row end */
    }
    /* transpose into CSC: rows visited ascending => sorted inside each column */
    {
        i64 run = 0;
        for (i64 j = 1; j <= n; j++) { i64 c = cnt[j]; cnt[j] = run; run += c; }
    }
    i64 total = 0;
    for (i64 i = 1; i <= m; i++) total += fill[i] - start[i];
    i64 *trow = (i64 *)malloc(sizeof(i64) * (size_t)(total ? total : 1));
    double *tval = (double *)malloc(sizeof(double) * (size_t)(total ? total : 1));
    i64 *cstart = (i64 *)malloc(sizeof(i64) * ((size_t)n + 2));
    if (!trow || !tval || !cstart) {
        free(trow); free(tval); free(cstart); free(start); free(fill);
        free(csrcolval); free(csrnzval); free(klast); free(cnt);
        return -1;
    }
    memcpy(cstart, cnt, sizeof(i64) * ((size_t)n + 2));
    for (i64 i = 1; i <= m; i++)
        for (i64 r = start[i]; r < fill[i]; r++) {
            i64 j = csrcolval[r];
            i64 p = cnt[j]++;
            trow[p] = i;
            tval[p] = csrnzval[r];
        }
    /* dropzeros: remove stored +-0.0 (iszero) */
    i64 nnz = 0;
    for (i64 j = 1; j <= n; j++) {
        colptr[j - 1] = nnz + 1;
        i64 cs = cstart[j], ce = (j < n) ? cstart[j + 1] : total;
        for (i64 p = cs; p < ce; p++)
            if (tval[p] != 0.0) { rowval[nnz] = trow[p]; nzval[nnz] = tval[p]; nnz++; }
    }
    colptr[n] = nnz + 1;
    free(trow); free(tval); free(cstart); free(start); free(fill);
    free(csrcolval); free(csrnzval); free(klast); free(cnt);
    return nnz;
}

/* 2x2 inverse the way LAPACK getrf/getri does it (partial pivoting);       */
/* only used for delta != 1 (Discretisation:234-235), 1e-12-level parity.    */
static void inv2(const double M[4] /* col-major */, double R[4]) {
    double a = M[0], c = M[1], b = M[2], d = M[3]; /* [a b; c d] */
    if (fabs(c) > fabs(a)) { /* swap rows */
        double l = a / c, u22 = b - l * d;
        /* PA = LU with P swap: L=[1 0; l 1], U=[c d; 0 u22] */
        double i11 = 1.0 / c, i22 = 1.0 / u22, i12 = -(d * i22) * i11;
        /* U^-1 = [i11 i12; 0 i22]; A^-1 = U^-1 L^-1 P */
        double x11 = i11 - i12 * l, x12 = i12, x21 = -i22 * l, x22 = i22;
        /* times P (column swap) */
        R[0] = x12; R[1] = x22; R[2] = x11; R[3] = x21;
    } else {
        double l = c / a, u22 = d - l * b;
        double i11 = 1.0 / a, i22 = 1.0 / u22, i12 = -(b * i22) * i11;
        R[0] = i11 - i12 * l; R[1] = -i22 * l; R[2] = i12; R[3] = i22;
    }
}

/* ------------------------------------------------------------------------ */
/* a5 + a6 + a7: ElementAssemblyLoop (variant 0, Discretisation:102-202),    */
/* ElementAssemblyLoopNEW (variant 1, :210-347), Sparsify (:355-364).        */
/* Materialises the same dense COO intermediates as the reference.           */
/* Outputs are Julia-style CSC (1-based Int64).  Capacities: Kuu/Muu         */
/* 36*nel entries, Kup 6*nel.  times[0] = element loop, times[1] = Sparsify.  */
/* ------------------------------------------------------------------------ */
int oracle_assemble(int variant, int formulation, double A, i64 nel, i64 nfaces,
                    const i64 *e2f, const i64 *bcarr, const double *Omega,
                    const double *Gamma, const double *n_x, const double *n_y,
                    const double *ke, const double *delta, const double *tau,
                    const double *alpha, const double *beta, const double *Z,
                    const double *VxDir, const double *VyDir,
                    const double *sxxNeu, const double *syyNeu,
                    const double *sxyNeu, const double *syxNeu,
                    i64 *Kuu_colptr, i64 *Kuu_rowval, double *Kuu_nzval, i64 *Kuu_nnz,
                    i64 *Muu_colptr, i64 *Muu_rowval, double *Muu_nzval, i64 *Muu_nnz,
                    i64 *Kup_colptr, i64 *Kup_rowval, double *Kup_nzval, i64 *Kup_nnz,
                    double *fu_out /* 2nf */, double *fp /* nel */, double *times) {
    const int nf = 3; /* "nf" inside the loops is mesh.nf_el (:114) */
    double t0 = now_s();
    size_t n36 = (size_t)36 * nel, n6 = (size_t)6 * nel;
    double *Kuui = (double *)calloc(n36 ? n36 : 1, sizeof(double));
    double *Kuuj = (double *)calloc(n36 ? n36 : 1, sizeof(double));
    double *Kuuv = (double *)calloc(n36 ? n36 : 1, sizeof(double));
    double *Muuv = (double *)calloc(n36 ? n36 : 1, sizeof(double));
    double *fu = (double *)calloc(n6 ? n6 : 1, sizeof(double));
    double *Kupi = (double *)calloc(n6 ? n6 : 1, sizeof(double));
    double *Kupj = (double *)calloc(n6 ? n6 : 1, sizeof(double));
    double *Kupv = (double *)calloc(n6 ? n6 : 1, sizeof(double));
    if (!Kuui || !Kuuj || !Kuuv || !Muuv || !fu || !Kupi || !Kupj || !Kupv) {
        free(Kuui); free(Kuuj); free(Kuuv); free(Muuv); free(fu); free(Kupi); free(Kupj); free(Kupv);
        return -1;
    }
    memset(fp, 0, sizeof(double) * nel);
    const double *b1 = beta, *b2 = beta + nel;
    const double *Z11 = Z, *Z21 = Z + nel, *Z12 = Z + 2 * nel, *Z22 = Z + 3 * nel;
#define K3(X, r, c, e) X[(size_t)(r) + 6 * (size_t)(c) + 36 * (size_t)(e)]
#define K2(X, r, e) X[(size_t)(r) + 6 * (size_t)(e)]
    for (i64 e = 0; e < nel; e++) {
        double Oe = Omega[e], eta = ke[e];
        double D11 = 0, D12 = 0, D21 = 0, D22 = 0;
        if (variant == 1) { /* :231-236 */
            double de = delta[e];
            if (de == 1.0) {
                D11 = (ke[e] * 1.0) / Oe;            /* ([..] .* 1.0)/Ω */
                D12 = ((ke[e] / de) * 1.0) / Oe;
                D21 = ((ke[e] / de) * 1.0) / Oe;
                D22 = (ke[e] * 1.0) / Oe;
            } else {
                double M[4] = {ke[e], ke[e] / de, ke[e] / de, ke[e]}, Dv[4], R[4];
                inv2(M, Dv);
                for (int q = 0; q < 4; q++) Dv[q] = Dv[q] * Oe;
                inv2(Dv, R);
                D11 = R[0]; D21 = R[1]; D12 = R[2]; D22 = R[3];
            }
        }
        double ia = 1.0 / alpha[e]; /* α[e]^-1 == inv(α[e]) */
        double iO = 1.0 / Oe;       /* Ωe^-1 */
        for (int i = 0; i < nf; i++) {
            double nix = n_x[i * nel + e], niy = n_y[i * nel + e];
            i64 nodei = e2f[i * nel + e]; /* 1-based */
            i64 bci = bcarr[nodei - 1];
            double J = 0.0 + bmul(bci == -1, 1.0); /* ȷ :127 */
            double Gi = Gamma[i * nel + e];
            double ti = tau[nodei - 1];
            for (int j = 0; j < nf; j++) {
                double njx = n_x[j * nel + e], njy = n_y[j * nel + e];
                i64 nodej = e2f[j * nel + e];
                i64 bcj = bcarr[nodej - 1];
                double Gj = Gamma[j * nel + e];
                double tj = tau[nodej - 1];
                int on = (bci != 1) & (bcj != 1);
                double ninj = nix * njx + niy * njy;
                double mG = bmul(on, -Gi); /* on * -Γi */
                double c = (formulation == GRADIENT) ? J : A;
                double t1 = ((ia * ti) * tj) * Gj; /* α^-1*τi*τj*Γj */
                if (variant == 0) { /* :138,146-158 */
                    double dlt = 0.0 + bmul(i == j, 1.0);
                    double t3 = ti * dlt;
                    double hG = (eta * iO) * Gj;   /* ηe*Ωe^-1*Γj */
                    double mhG = ((-eta) * iO) * Gj; /* -ηe*Ωe^-1*Γj */
                    K3(Kuuv, j, i, e) = mG * ((t1 - hG * (ninj + (c * nix) * njx)) - t3);
                    K3(Kuuv, j + nf, i, e) = mG * (mhG * ((c * niy) * njx));
                    K3(Kuuv, j, i + nf, e) = mG * (mhG * ((c * nix) * njy));
                    K3(Kuuv, j + nf, i + nf, e) = mG * ((t1 - hG * (ninj + (c * niy) * njy)) - t3);
                    K3(Muuv, j, i, e) = mG * ((t1 - hG * ninj) - t3);
                    K3(Muuv, j + nf, i + nf, e) = mG * ((t1 - hG * ninj) - t3);
                    /* connectivity :161-164 */
                    K3(Kuui, j, i, e) = (double)nodei;              K3(Kuui, j + nf, i, e) = (double)nodei;
                    K3(Kuuj, j, i, e) = (double)nodej;              K3(Kuuj, j + nf, i, e) = (double)(nodej + nfaces);
                    K3(Kuui, j, i + nf, e) = (double)(nodei + nfaces); K3(Kuui, j + nf, i + nf, e) = (double)(nodei + nfaces);
                    K3(Kuuj, j, i + nf, e) = (double)nodej;         K3(Kuuj, j + nf, i + nf, e) = (double)(nodej + nfaces);
                } else { /* :265,273-276 / :288-291, :295-302 */
                    double t3 = bmul(i == j, ti); /* τi*δ with δ::Bool */
                    if (formulation == GRADIENT) {
                        K3(Kuuv, j, i, e) = mG * ((t1 - (D11 * Gj) * (ninj + (c * nix) * njx)) - t3);
                        K3(Kuuv, j, i + nf, e) = mG * (((-D12) * Gj) * ((c * nix) * njy));
                        K3(Kuuv, j + nf, i, e) = mG * (((-D21) * Gj) * ((c * niy) * njx));
                        K3(Kuuv, j + nf, i + nf, e) = mG * ((t1 - (D22 * Gj) * (ninj + (c * niy) * njy)) - t3);
                    } else {
                        K3(Kuuv, i, j, e) = mG * ((t1 - (D11 * Gj) * (ninj + (c * nix) * njx)) - t3);
                        K3(Kuuv, i, j + nf, e) = mG * (((-D12) * Gj) * ((c * niy) * njx));
                        K3(Kuuv, i + nf, j, e) = mG * (((-D21) * Gj) * ((c * nix) * njy));
                        K3(Kuuv, i + nf, j + nf, e) = mG * ((t1 - (D22 * Gj) * (ninj + (c * niy) * njy)) - t3);
                    }
                    K3(Muuv, j, i, e) = mG * ((t1 - (D11 * Gj) * ninj) - t3);
                    K3(Muuv, j + nf, i + nf, e) = mG * ((t1 - (D22 * Gj) * ninj) - t3);
                    K3(Kuui, i, j, e) = (double)nodei;              K3(Kuui, i + nf, j, e) = (double)(nodei + nfaces);
                    K3(Kuuj, i, j, e) = (double)nodej;              K3(Kuuj, i + nf, j, e) = (double)nodej;
                    K3(Kuui, i, j + nf, e) = (double)nodei;         K3(Kuui, i + nf, j + nf, e) = (double)(nodei + nfaces);
                    K3(Kuuj, i, j + nf, e) = (double)(nodej + nfaces); K3(Kuuj, i + nf, j + nf, e) = (double)(nodej + nfaces);
                }
            }
            /* RHS :167-178 / :310-323 */
            double Xi = 0.0 + bmul(bci == 2, 1.0);
            double tix = nix * sxxNeu[nodei - 1] + niy * sxyNeu[nodei - 1];
            double tiy = nix * syxNeu[nodei - 1] + niy * syyNeu[nodei - 1];
            double mGi = bmul(bci != 1, -Gi);
            double feix, feiy;
            if (variant == 0) {
                double nZx, nZy;
                if (formulation == GRADIENT) {
                    nZx = nix * (Z11[e] + J * Z11[e]) + niy * (Z21[e] + J * Z12[e]);
                    nZy = nix * (Z12[e] + J * Z21[e]) + niy * (Z22[e] + J * Z22[e]);
                } else {
                    nZx = nix * Z11[e] + niy * Z21[e];
                    nZy = nix * Z12[e] + niy * Z22[e];
                }
                feix = mGi * (((((-ia) * ti) * b1[e]) + (eta * iO) * nZx) - tix * Xi);
                feiy = mGi * (((((-ia) * ti) * b2[e]) + (eta * iO) * nZy) - tiy * Xi);
            } else {
                feix = mGi * (((((-ia) * ti) * b1[e]) + ((D11 * nix) * Z11[e] + (D11 * niy) * Z21[e])) - tix * Xi);
                feiy = mGi * (((((-ia) * ti) * b2[e]) + ((D22 * nix) * Z12[e] + (D22 * niy) * Z22[e])) - tiy * Xi);
            }
            /* up block :180-185 */
            K2(Kupv, i, e) -= bmul(bci != 1, Gi) * nix;
            K2(Kupv, i + nf, e) -= bmul(bci != 1, Gi) * niy;
            K2(Kupi, i, e) = (double)nodei;
            K2(Kupj, i, e) = (double)(e + 1);
            K2(Kupi, i + nf, e) = (double)(nodei + nfaces);
            K2(Kupj, i + nf, e) = (double)(e + 1);
            /* Dirichlet :187-192 */
            K3(Kuuv, i, i, e) += bmul(bci == 1, 1.0);
            K3(Kuuv, i + nf, i + nf, e) += bmul(bci == 1, 1.0);
            K3(Muuv, i, i, e) += bmul(bci == 1, 1.0);
            K3(Muuv, i + nf, i + nf, e) += bmul(bci == 1, 1.0);
            K2(fu, i, e) += bmul(bci != 1, feix) + bmul(bci == 1, VxDir[nodei - 1]) * 1.0;
            K2(fu, i + nf, e) += bmul(bci != 1, feiy) + bmul(bci == 1, VyDir[nodei - 1]) * 1.0;
            /* :194 */
            fp[e] -= bmul(bci == 1, Gi) * (VxDir[nodei - 1] * nix + VyDir[nodei - 1] * niy);
        }
    }
    double t1s = now_s();
    /* Sparsify :355-364 -- X[:] makes flat copies (index arrays twice).      */
    int rc = 0;
    {
        double *If = (double *)malloc(sizeof(double) * (n36 ? n36 : 1));
        double *Jf = (double *)malloc(sizeof(double) * (n36 ? n36 : 1));
        double *Vf = (double *)malloc(sizeof(double) * (n36 ? n36 : 1));
        double *one = (double *)malloc(sizeof(double) * (n6 ? n6 : 1));
        i64 *fcolptr = (i64 *)malloc(sizeof(i64) * 2);
        i64 *frow = (i64 *)malloc(sizeof(i64) * (n6 ? n6 : 1));
        double *fval = (double *)malloc(sizeof(double) * (n6 ? n6 : 1));
        if (!If || !Jf || !Vf || !one || !fcolptr || !frow || !fval) rc = -1;
        if (!rc) {
            for (size_t k = 0; k < n6; k++) one[k] = 1.0; /* _one :357 */
            memcpy(If, Kuui, sizeof(double) * n36); memcpy(Jf, Kuuj, sizeof(double) * n36);
            memcpy(Vf, Kuuv, sizeof(double) * n36);
            *Kuu_nnz = sparse_dropzeros((i64)n36, If, Jf, Vf, 2 * nfaces, 2 * nfaces, Kuu_colptr, Kuu_rowval, Kuu_nzval);
            memcpy(If, Kuui, sizeof(double) * n36); memcpy(Jf, Kuuj, sizeof(double) * n36);
            memcpy(Vf, Muuv, sizeof(double) * n36);
            *Muu_nnz = sparse_dropzeros((i64)n36, If, Jf, Vf, 2 * nfaces, 2 * nfaces, Muu_colptr, Muu_rowval, Muu_nzval);
            memcpy(If, Kupi, sizeof(double) * n6); memcpy(Jf, Kupj, sizeof(double) * n6);
            memcpy(Vf, Kupv, sizeof(double) * n6);
            *Kup_nnz = sparse_dropzeros((i64)n6, If, Jf, Vf, 2 * nfaces, nel, Kup_colptr, Kup_rowval, Kup_nzval);
            memcpy(If, Kupi, sizeof(double) * n6); memcpy(Vf, fu, sizeof(double) * n6);
            i64 fn = sparse_dropzeros((i64)n6, If, one, Vf, 2 * nfaces, 1, fcolptr, frow, fval);
            if (*Kuu_nnz < 0 || *Muu_nnz < 0 || *Kup_nnz < 0 || fn < 0) rc = -1;
            else {
                memset(fu_out, 0, sizeof(double) * 2 * nfaces); /* Array(sparse) */
                for (i64 p = 0; p < fn; p++) fu_out[frow[p] - 1] = fval[p];
            }
        }
        free(If); free(Jf); free(Vf); free(one); free(fcolptr); free(frow); free(fval);
    }
    double t2s = now_s();
    if (times) { times[0] = t1s - t0; times[1] = t2s - t1s; }
    free(Kuui); free(Kuuj); free(Kuuv); free(Muuv); free(fu); free(Kupi); free(Kupj); free(Kupv);
    return rc;
#undef K3
#undef K2
}

/* ------------------------------------------------------------------------ */
/* a8: ComputeResidualsFCFV_Stokes_o1 -- src/ComputeResidualsFCFV_Stokes.jl  */
/* :5-108.  tau_mode 0 = REFERENCE (mesh.τ[e], face array indexed by element */
/* id, :17,54), 1 = FACE (mesh.τ[e2f[e,i]]).  norms[6] = norm(F)/sqrt(len)   */
/* in print order (:101-106).  Optional outputs may be NULL.                  */
/* Small dense products are evaluated as plain left-to-right sums.            */
/* ------------------------------------------------------------------------ */
static double rms(const double *x, i64 n) {
    long double s = 0.0L;
    for (i64 k = 0; k < n; k++) s += (long double)x[k] * (long double)x[k];
    return n ? (double)(sqrtl(s) / sqrtl((long double)n)) : 0.0;
}

int oracle_residuals(int formulation, int tau_mode, i64 nel, i64 nfaces,
                     const i64 *e2f, const i64 *bcarr, const double *Omega,
                     const double *Gamma, const double *n_x, const double *n_y,
                     const double *ke, const double *delta, const double *tau,
                     const double *ae, const double *be, const double *ze,
                     const double *Vxh, const double *Vyh, const double *Pe,
                     const double *sex, const double *sey, const double *VxDir,
                     const double *VyDir, const double *SxxNeu,
                     const double *SyyNeu, const double *SxyNeu,
                     const double *SyxNeu, double *norms,
                     double *F_nodes_x_out, double *F_nodes_y_out,
                     double *F_glob2_out, double *F_eq1_out, double *F_eq2_out,
                     double *F_eq3_out) {
    size_t n = (size_t)(nel ? nel : 1);
    double *ue = (double *)calloc(2 * n, sizeof(double));
    double *Le = (double *)calloc(4 * n, sizeof(double));
    double *Fg1 = (double *)calloc(6 * n, sizeof(double)); /* (nel,3,2) */
    double *Fg2 = (double *)calloc(n, sizeof(double));
    double *Fe1 = (double *)calloc(4 * n, sizeof(double));
    double *Fe2 = (double *)calloc(2 * n, sizeof(double));
    double *Fe3 = (double *)calloc(n, sizeof(double));
    double *Fnx = (double *)calloc((size_t)(nfaces ? nfaces : 1), sizeof(double));
    double *Fny = (double *)calloc((size_t)(nfaces ? nfaces : 1), sizeof(double));
    if (!ue || !Le || !Fg1 || !Fg2 || !Fe1 || !Fe2 || !Fe3 || !Fnx || !Fny) {
        free(ue); free(Le); free(Fg1); free(Fg2); free(Fe1); free(Fe2); free(Fe3); free(Fnx); free(Fny);
        return -1;
    }
#define U(a) ue[(size_t)e + n * (a)]
#define L(a, b) Le[(size_t)e + n * (a) + 2 * n * (b)]
#define ZE(a, b) ze[(size_t)e + n * (a) + 2 * n * (b)]
#define F1(a, b) Fe1[(size_t)e + n * (a) + 2 * n * (b)]
    if (nel == 0) n = 0;
    /* pass A :12-32 */
    for (i64 e = 0; e < nel; e++) {
        double vole = Omega[e];
        U(0) = be[e] / ae[e];
        U(1) = be[nel + e] / ae[e];
        double mi = -1.0 / vole;
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) L(a, b) = mi * ZE(a, b);
        for (int i = 0; i < 3; i++) {
            i64 nodei = e2f[i * nel + e] - 1;
            double ti = tau_mode ? tau[nodei] : tau[e]; /* :17 */
            if (bcarr[nodei] != 1) {
                double Gi = Gamma[i * nel + e];
                double nn[2] = {n_x[i * nel + e], n_y[i * nel + e]};
                double u[2] = {Vxh[nodei], Vyh[nodei]};
                U(0) = U(0) + ((Gi * ti) * u[0]) / ae[e]; /* Γi*τi*u'[:]/ae[e] */
                U(1) = U(1) + ((Gi * ti) * u[1]) / ae[e];
                double s = (1.0 / vole) * Gi;
                for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
                    double m = nn[a] * u[b];
                    if (formulation == SYMMETRIC_GRADIENT) m = m + u[a] * nn[b];
                    L(a, b) = L(a, b) - s * m;
                }
            }
        }
    }
    /* pass B :42-84 */
    for (i64 e = 0; e < nel; e++) {
        double vole = Omega[e], eta = ke[e];
        double D[2][2] = {{eta, eta / delta[e]}, {eta / delta[e], eta}};
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) F1(a, b) = vole * L(a, b);
        Fe2[e] = -sex[e] * vole;
        Fe2[n + e] = -sey[e] * vole;
        Fe3[e] = 0.0;
        for (int i = 0; i < 3; i++) {
            i64 nodei = e2f[i * nel + e] - 1;
            double dAi = Gamma[i * nel + e];
            double taui = tau_mode ? tau[nodei] : tau[e]; /* :54 */
            double nn[2] = {n_x[i * nel + e], n_y[i * nel + e]};
            i64 bci = bcarr[nodei];
            double Xi = 0.0 + bmul(bci == 2, 1.0);
            double Ji = 0.0 + bmul(bci == -1, 1.0);
            double ti[2] = {nn[0] * SxxNeu[nodei] + nn[1] * SxyNeu[nodei],
                            nn[0] * SyxNeu[nodei] + nn[1] * SyyNeu[nodei]};
            double u[2];
            if (bci == 1) { u[0] = VxDir[nodei]; u[1] = VyDir[nodei]; }
            else          { u[0] = Vxh[nodei];   u[1] = Vyh[nodei]; }
            for (int b = 0; b < 2; b++) {
                /* (n'*(D.*Le))[b] = n1*D1b*L1b + n2*D2b*L2b */
                double nDL = nn[0] * (D[0][b] * L(0, b)) + nn[1] * (D[1][b] * L(1, b));
                double acc = (((nDL + Pe[e] * nn[b]) + taui * U(b)) - taui * u[b]) + ti[b] * Xi;
                if (formulation == GRADIENT) {
                    /* (n'*(D.*Le)')[b] = n1*D b1*L b1 + n2*D b2*L b2 */
                    double nDLt = nn[0] * (D[b][0] * L(b, 0)) + nn[1] * (D[b][1] * L(b, 1));
                    acc = acc + Ji * nDLt;
                }
                Fg1[(size_t)e + n * i + 3 * n * b] += dAi * acc;
            }
            Fe2[e] = Fe2[e] + (dAi * taui) * U(0);
            Fe2[n + e] = Fe2[n + e] + (dAi * taui) * U(1);
            /* dAi * u'*n parses as (dAi*u')*n (:75, :82): the row is scaled first */
            double un = (dAi * u[0]) * nn[0] + (dAi * u[1]) * nn[1];
            Fg2[e] = Fg2[e] + un;
            for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
                if (formulation == SYMMETRIC_GRADIENT)      /* dAi * (n*u' .+ u*n') (:79) */
                    F1(a, b) = F1(a, b) + dAi * (nn[a] * u[b] + u[a] * nn[b]);
                else                                        /* dAi * n * u' = (dAi*n)*u' (:77) */
                    F1(a, b) = F1(a, b) + (dAi * nn[a]) * u[b];
            }
            Fe2[e] = Fe2[e] - (dAi * taui) * u[0];
            Fe2[n + e] = Fe2[n + e] - (dAi * taui) * u[1];
            Fe3[e] = Fe3[e] + un;
        }
    }
    /* pass C :88-100 */
    for (i64 e = 0; e < nel; e++)
        for (int i = 0; i < 3; i++) {
            i64 nodei = e2f[i * nel + e] - 1;
            if (bcarr[nodei] != 1) {
                Fnx[nodei] += Fg1[(size_t)e + n * i];
                Fny[nodei] += Fg1[(size_t)e + n * i + 3 * n];
            }
        }
    norms[0] = rms(Fe1, 4 * nel);
    norms[1] = rms(Fe2, 2 * nel);
    norms[2] = rms(Fe3, nel);
    norms[3] = rms(Fnx, nfaces);
    norms[4] = rms(Fny, nfaces);
    norms[5] = rms(Fg2, nel);
    if (F_nodes_x_out) memcpy(F_nodes_x_out, Fnx, sizeof(double) * nfaces);
    if (F_nodes_y_out) memcpy(F_nodes_y_out, Fny, sizeof(double) * nfaces);
    if (F_glob2_out) memcpy(F_glob2_out, Fg2, sizeof(double) * nel);
    if (F_eq1_out) memcpy(F_eq1_out, Fe1, sizeof(double) * 4 * nel);
    if (F_eq2_out) memcpy(F_eq2_out, Fe2, sizeof(double) * 2 * nel);
    if (F_eq3_out) memcpy(F_eq3_out, Fe3, sizeof(double) * nel);
    free(ue); free(Le); free(Fg1); free(Fg2); free(Fe1); free(Fe2); free(Fe3); free(Fnx); free(Fny);
    return 0;
#undef U
#undef L
#undef ZE
#undef F1
}

/* ------------------------------------------------------------------------ */
/* a9: Powell-Hestenes residual and pressure update on Julia-style CSC       */
/* -- src/SolversFCFV_Stokes.jl:23-24,40-49.  Kpu = -Kup' (:6).               */
/* y = A*x for CSC walks columns ascending, so each y[row] accumulates its   */
/* products in ascending column order starting from 0.0.                     */
/* ------------------------------------------------------------------------ */
static void csc_mv(i64 m, i64 n, const i64 *colptr, const i64 *rowval,
                   const double *nzval, const double *x, double *y) {
    for (i64 i = 0; i < m; i++) y[i] = 0.0;
    for (i64 j = 0; j < n; j++)
        for (i64 p = colptr[j] - 1; p < colptr[j + 1] - 1; p++)
            y[rowval[p] - 1] += nzval[p] * x[j];
}

int oracle_ph_residual(i64 nel, i64 nfaces, const i64 *Kuu_colptr,
                       const i64 *Kuu_rowval, const double *Kuu_nzval,
                       const i64 *Kup_colptr, const i64 *Kup_rowval,
                       const double *Kup_nzval, const double *fu,
                       const double *fp, const double *u, const double *p,
                       double *ru, double *rp, double *nrm) {
    i64 m = 2 * nfaces;
    double *t1 = (double *)malloc(sizeof(double) * (size_t)(m ? m : 1));
    double *t2 = (double *)malloc(sizeof(double) * (size_t)(m ? m : 1));
    if (!t1 || !t2) { free(t1); free(t2); return -1; }
    csc_mv(m, m, Kuu_colptr, Kuu_rowval, Kuu_nzval, u, t1);
    csc_mv(m, nel, Kup_colptr, Kup_rowval, Kup_nzval, p, t2);
    for (i64 i = 0; i < m; i++) ru[i] = (fu[i] - t1[i]) - t2[i]; /* :40 */
    /* rp = fp - Kpu*u, Kpu = -Kup' : (Kpu*u)[e] = sum_r (-Kup[r,e])*u[r], r ascending */
    for (i64 e = 0; e < nel; e++) {
        double s = 0.0;
        for (i64 q = Kup_colptr[e] - 1; q < Kup_colptr[e + 1] - 1; q++)
            s += (-Kup_nzval[q]) * u[Kup_rowval[q] - 1];
        rp[e] = fp[e] - s; /* :41 */
    }
    long double su = 0.0L, sp = 0.0L;
    for (i64 i = 0; i < m; i++) su += (long double)ru[i] * ru[i];
    for (i64 e = 0; e < nel; e++) sp += (long double)rp[e] * rp[e];
    nrm[0] = (double)sqrtl(su); /* norm(ru) :42 */
    nrm[1] = (double)sqrtl(sp);
    free(t1); free(t2);
    return 0;
}

/* p .+= Kppi*(fp .- Kpu*u), coef = penalty.*ke./Ω (:23,49) */
int oracle_ph_update_p(i64 nel, const i64 *Kup_colptr, const i64 *Kup_rowval,
                       const double *Kup_nzval, const double *fp,
                       const double *ke, const double *Omega, double penalty,
                       const double *u, double *p) {
    for (i64 e = 0; e < nel; e++) {
        double s = 0.0;
        for (i64 q = Kup_colptr[e] - 1; q < Kup_colptr[e + 1] - 1; q++)
            s += (-Kup_nzval[q]) * u[Kup_rowval[q] - 1];
        double coef = (penalty * ke[e]) / Omega[e];
        p[e] = p[e] + coef * (fp[e] - s);
    }
    return 0;
}

/* fusc .= fu .- Kup*(Kppi*fp .+ p)  (:47) */
int oracle_ph_fusc(i64 nel, i64 nfaces, const i64 *Kup_colptr,
                   const i64 *Kup_rowval, const double *Kup_nzval,
                   const double *fu, const double *fp, const double *ke,
                   const double *Omega, double penalty, const double *p,
                   double *fusc) {
    i64 m = 2 * nfaces;
    double *w = (double *)malloc(sizeof(double) * (size_t)(nel ? nel : 1));
    double *t = (double *)malloc(sizeof(double) * (size_t)(m ? m : 1));
    if (!w || !t) { free(w); free(t); return -1; }
    for (i64 e = 0; e < nel; e++) w[e] = ((penalty * ke[e]) / Omega[e]) * fp[e] + p[e];
    csc_mv(m, nel, Kup_colptr, Kup_rowval, Kup_nzval, w, t);
    for (i64 i = 0; i < m; i++) fusc[i] = fu[i] - t[i];
    free(w); free(t);
    return 0;
}

/* ComputeElementValues -- src/DiscretisationFCFV_Stokes.jl:52-94 (next row) */
int oracle_element_values(i64 nel, const i64 *e2f, const i64 *bcarr,
                          const double *Omega, const double *Gamma,
                          const double *n_x, const double *n_y,
                          const double *ke, const double *delta,
                          const double *tau, const double *alpha,
                          const double *beta, const double *Z,
                          const double *Vxh, const double *Vyh, double *Vxe,
                          double *Vye, double *Txx, double *Tyy, double *Txy) {
    const double *Z11 = Z, *Z21 = Z + nel, *Z12 = Z + 2 * nel, *Z22 = Z + 3 * nel;
    for (i64 e = 0; e < nel; e++) {
        double eta = ke[e], D11 = eta, D12 = eta / delta[e], D22 = eta, O = Omega[e];
        Vxe[e] = beta[e] / alpha[e];
        Vye[e] = beta[nel + e] / alpha[e];
        Txx[e] = (eta / O) * Z11[e];
        Tyy[e] = (eta / O) * Z22[e];
        Txy[e] = ((eta / O) * 0.5) * (Z12[e] + Z21[e]);
        for (int i = 0; i < 3; i++) {
            i64 nodei = e2f[i * nel + e] - 1;
            int nd = bcarr[nodei] != 1;
            double Gi = Gamma[i * nel + e], nix = n_x[i * nel + e], niy = n_y[i * nel + e];
            double ti = tau[nodei];
            Vxe[e] += ((bmul(nd, Gi) * ti) * Vxh[nodei]) / alpha[e];
            Vye[e] += ((bmul(nd, Gi) * ti) * Vyh[nodei]) / alpha[e];
            Txx[e] += (((bmul(nd, D11) / O) * Gi) * nix) * Vxh[nodei];
            Tyy[e] += (((bmul(nd, D22) / O) * Gi) * niy) * Vyh[nodei];
            Txy[e] += (bmul(nd, D12) * 0.5) * (((1.0 / O) * Gi) * (nix * Vyh[nodei] + niy * Vxh[nodei]));
        }
        Txx[e] *= 2.0; Tyy[e] *= 2.0; Txy[e] *= 2.0;
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Schur complement of the Powell-Hestenes solver, literally as Julia's      */
/* SparseArrays evaluates  src/SolversFCFV_Stokes.jl:6,23-25 :               */
/*   Kpu   = -Kup'                      (adjoint, then unary minus)          */
/*   coef  = penalty.*mesh.ke./mesh.Ω   ((penalty*ke)/Ω)                     */
/*   Kppi  = spdiagm(coef)                                                   */
/*   Kuusc = Kuu .- Kup*(Kppi*Kpu)                                           */
/* SparseArrays semantics restated (Julia stdlib, version unpinned by the    */
/* reference's Project.toml):  A*B is Gustavson's column-by-column product   */
/* (for column j of B: for each stored B[k,j] in ascending k, for each stored*/
/* A[i,k]: the first product of a row is assigned, later ones are added),    */
/* numerical zeros stay stored, rows come out sorted;  A .- B is the         */
/* zero-preserving sparse broadcast: result entries over the union of the    */
/* stored patterns, f(a, 0.0) / f(0.0, b) where one side is missing, and a   */
/* result is only stored if it is not zero.                                  */
/* All matrices are Julia CSC triples (1-based colptr / rowval).             */
/* Kuu: n x n (n = 2 nf), Kup: n x nel.  Output capacity: nnz(Kuu)+36*nel.   */
/* ------------------------------------------------------------------------ */
int oracle_schur(i64 n, i64 nel, const i64 *uu_cp, const i64 *uu_rv, const double *uu_v,
                 const i64 *up_cp, const i64 *up_rv, const double *up_v,
                 const double *ke, const double *Omega, double penalty,
                 i64 *sc_cp, i64 *sc_rv, double *sc_v, i64 *sc_nnz) {
    const i64 nnzp = up_cp[nel] - 1;
    /* Kpu = -Kup' : nel x n, CSC.  Transpose by counting rows of Kup. */
    i64 *pu_cp = calloc((size_t)n + 2, sizeof(i64));
    i64 *pu_rv = malloc(sizeof(i64) * (size_t)(nnzp > 0 ? nnzp : 1));
    double *pu_v = malloc(sizeof(double) * (size_t)(nnzp > 0 ? nnzp : 1));
    i64 *next = malloc(sizeof(i64) * (size_t)(n + 1));
    double *coef = malloc(sizeof(double) * (size_t)(nel > 0 ? nel : 1));
    double *x = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    i64 *xb = calloc((size_t)(n > 0 ? n : 1), sizeof(i64));
    i64 *rows = malloc(sizeof(i64) * (size_t)(n > 0 ? n : 1));
    if (!pu_cp || !pu_rv || !pu_v || !next || !coef || !x || !xb || !rows) return 1;
    for (i64 k = 0; k < nnzp; k++) pu_cp[up_rv[k]]++;          /* count per row i (1-based) -> column i of Kpu */
    { i64 run = 1; for (i64 j = 1; j <= n; j++) { i64 c = pu_cp[j]; pu_cp[j] = run; run += c; } pu_cp[n + 1] = run; }
    for (i64 j = 1; j <= n; j++) next[j] = pu_cp[j];
    for (i64 e = 1; e <= nel; e++)                             /* ascending e keeps rows of Kpu sorted */
        for (i64 k = up_cp[e - 1]; k < up_cp[e]; k++) {
            const i64 i = up_rv[k - 1];
            const i64 p = next[i]++;
            pu_rv[p - 1] = e;
            pu_v[p - 1] = -up_v[k - 1];
        }
    for (i64 e = 0; e < nel; e++) coef[e] = (penalty * ke[e]) / Omega[e];
    /* B = Kppi*Kpu scales row e of Kpu: B[e,j] = coef[e]*Kpu[e,j]   (one product per entry) */
    /* C = Kup*B and Kuusc = Kuu .- C, column by column */
    i64 out = 0;
    sc_cp[0] = 1;
    for (i64 j = 1; j <= n; j++) {
        i64 nr = 0;
        for (i64 kb = pu_cp[j]; kb < pu_cp[j + 1]; kb++) {     /* stored B[e,j], ascending e */
            const i64 e = pu_rv[kb - 1];
            const double b = coef[e - 1] * pu_v[kb - 1];
            for (i64 ka = up_cp[e - 1]; ka < up_cp[e]; ka++) { /* stored Kup[i,e], ascending i */
                const i64 i = up_rv[ka - 1];
                const double prod = up_v[ka - 1] * b;
                if (xb[i - 1] != j) { xb[i - 1] = j; x[i - 1] = prod; rows[nr++] = i; }
                else x[i - 1] += prod;
            }
        }
        /* sort the rows of column j of C (insertion sort: <= 10 entries per column) */
        for (i64 a = 1; a < nr; a++) { i64 r = rows[a], b2 = a; while (b2 > 0 && rows[b2 - 1] > r) { rows[b2] = rows[b2 - 1]; b2--; } rows[b2] = r; }
        /* merge with column j of Kuu: zero-preserving broadcast of (a, c) -> a - c */
        i64 ka = uu_cp[j - 1], ea = uu_cp[j], kc = 0;
        while (ka < ea || kc < nr) {
            i64 ra = ka < ea ? uu_rv[ka - 1] : (n + 1), rc = kc < nr ? rows[kc] : (n + 1);
            i64 r; double v;
            if (ra == rc) { r = ra; v = uu_v[ka - 1] - x[rc - 1]; ka++; kc++; }
            else if (ra < rc) { r = ra; v = uu_v[ka - 1] - 0.0; ka++; }
            else { r = rc; v = 0.0 - x[rc - 1]; kc++; }
            if (v != 0.0) { sc_rv[out] = r; sc_v[out] = v; out++; }
        }
        sc_cp[j] = out + 1;
    }
    *sc_nnz = out;
    free(pu_cp); free(pu_rv); free(pu_v); free(next); free(coef); free(x); free(xb); free(rows);
    return 0;
}
