// host_emulation.cpp -- TEST INFRASTRUCTURE.  Compiles the per-element / per-face arithmetic of
// the CUDA library (fcfv_math.cuh and the gather/count helpers of fcfv_kernels.cuh) for the HOST
// and drives it with plain serial loops, so that tests without a GPU can compare the arithmetic
// the kernels execute against the oracle.  It replaces only the launch machinery (thread index ->
// loop index, block scan -> running sum); it is never loaded by the product package.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (see tests/conftest.py).
#include "../fcfv_nme23_b200/csrc/fcfv_kernels.cuh"
#include "../fcfv_nme23_b200/csrc/fcfv_generate.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace fcfv;

namespace {
struct HostMesh {
    std::vector<int> e2f; std::vector<int8_t> bc; std::vector<int2> f2e;
    std::vector<int4> fconn; std::vector<int> fbc, ebc;
    MeshView view;
};

void build_mesh(HostMesh& H, long long nel, long long nf, const long long* e2f, const long long* bc, const double* Om,
                const double* nx, const double* ny, const double* G, const double* ke, const double* dl,
                const double* taue, const double* tau) {
    H.e2f.resize(3 * nel); H.bc.resize(nf); H.f2e.assign(nf, int2{0x7fffffff, -1});
    for (long long k = 0; k < 3 * nel; k++) H.e2f[k] = (int)(e2f[k] - 1);
    for (long long f = 0; f < nf; f++) { long long v = bc[f]; H.bc[f] = (v == 1 || v == 2 || v == -1) ? (int8_t)v : 0; }
    std::vector<int> cnt(nf, 0);
    for (long long e = 0; e < nel; e++)
        for (int i = 0; i < 3; i++) {
            int f = H.e2f[i * nel + e], code = 4 * (int)e + i;
            if (code < H.f2e[f].x) H.f2e[f].x = code;
            if (code > H.f2e[f].y) H.f2e[f].y = code;
            cnt[f]++;
        }
    for (long long f = 0; f < nf; f++) if (!cnt[f]) H.f2e[f] = int2{-1, -1};
    H.ebc.resize(nel);
    for (long long e = 0; e < nel; e++) {   // k_ebc
        int p = 0;
        for (int j = 0; j < 3; j++) p |= (((int)H.bc[H.e2f[j * nel + e]] + 1) & 15) << (4 * j);
        H.ebc[e] = p;
    }
    H.fconn.resize(nf); H.fbc.resize(nf);
    for (long long f = 0; f < nf; f++)   // k_fconn
        face_conn_of(H.e2f.data(), H.bc.data(), (int)nel, (int)f, H.f2e[f], H.fconn[f], H.fbc[f]);
    MeshView& M = H.view;
    M.nel = (int)nel; M.nf = (int)nf; M.e2f = H.e2f.data(); M.bc = H.bc.data(); M.f2e = H.f2e.data(); M.fconn = H.fconn.data(); M.fbc = H.fbc.data(); M.ebc = H.ebc.data();
    M.Om = Om; M.G = G; M.nx = nx; M.ny = ny; M.ke = ke; M.dl = dl; M.taue = taue; M.tau = tau;
    M.tau_ref = nullptr;
}

// Serial stand-in for k_asm_count / k_asm_fill: thread = face becomes a loop iteration, the block
// scan becomes a running sum, the tile cache becomes a direct load; asm_face, the rank-space
// placement (rows_emit_*) and the counts are the kernels' own code.
// EXTRA = kExtraMuu: third block Muu; kExtraSchur: third block Kuu - Kup*(Kppi*Kpu) with penalty gamma
template <int VARIANT, int FORM, int EXTRA>
void assemble(const MeshView& M, const LocalView& L, const FaceDataView& D, double A, double gamma, long long* rp_uu, int* c_uu,
              double* v_uu, long long* rp_up, int* c_up, double* v_up, long long* rp_mu, int* c_mu, double* v_mu,
              double* fu, double* fp) {
    const int nf = M.nf;
    // k_elem_coef
    const int nel = M.nel;
    bool aniso = false;
    for (int e = 0; e < nel; e++) aniso = aniso || M.dl[e] != 1.0;
    std::vector<double> ia(nel), c11(nel), c12(nel), c21(nel), c22(nel), cs(nel);
    for (int e = 0; e < nel; e++) {
        element_coef<VARIANT>(L.alpha[e], M.ke[e], (VARIANT == kLoopNew && aniso) ? M.dl[e] : 1.0, M.Om[e], ia[e], c11[e], c12[e], c21[e], c22[e]);
        cs[e] = (gamma * M.ke[e]) / M.Om[e];
    }
    const CoefView Cf{ia.data(), c11.data(), aniso ? c12.data() : nullptr, aniso ? c21.data() : nullptr, aniso ? c22.data() : nullptr,
                      cs.data()};
    FaceConn C;
    auto load = [&](bool second, int e, int i, FaceElem& F) {
        if (aniso) load_face_elem<VARIANT, true, EXTRA>(M, L, Cf, C, second, e, i, F);
        else load_face_elem<VARIANT, false, EXTRA>(M, L, Cf, C, second, e, i, F);
    };
    std::vector<int> cnt[6];
    for (auto& c : cnt) c.assign(nf, 0);
    for (int f = 0; f < nf; f++) {   // count pass
        FaceRows R;
        unpack_conn(f, M.fconn[f], M.fbc[f], C);
        asm_face<VARIANT, FORM, EXTRA>(D, A, f, M.f2e[f], C, true, load, NullSink(), R);
        cnt[0][f] = rows_nuu_x(R); cnt[1][f] = rows_nuu_y(R);
        cnt[2][f] = rows_nup_x(R); cnt[3][f] = rows_nup_y(R);
        cnt[4][f] = rows_nex_x<EXTRA>(R); cnt[5][f] = rows_nex_y<EXTRA>(R);
    }
    long long* rps[3] = {rp_uu, rp_up, rp_mu};
    for (int m = 0; m < 3; m++) {
        long long run = 0;
        for (int f = 0; f < nf; f++) { rps[m][f] = run; run += cnt[2 * m][f]; }
        for (int f = 0; f < nf; f++) { rps[m][nf + f] = run; run += cnt[2 * m + 1][f]; }
        rps[m][2 * nf] = run;
    }
    for (int e = 0; e < M.nel; e++) fp[e] = 0.0;
    for (int f = 0; f < nf; f++) {   // k_asm_fp
        if (((M.fbc[f] & 15) - 1) != 1 || M.f2e[f].x < 0) continue;
        const int2 pr = M.f2e[f];
        for (int s2 = 0; s2 < 2; s2++) {
            if (s2 == 1 && pr.y == pr.x) break;
            const int code = s2 == 0 ? pr.x : pr.y;
            double v;
            if (face_fp(M, D, code >> 2, code & 3, v)) fp[code >> 2] = v;
        }
    }
    for (int f = 0; f < nf; f++) {   // fill pass (same placement arithmetic as k_asm_fill)
        FaceRows R;
        unpack_conn(f, M.fconn[f], M.fbc[f], C);
        RowVals V;
        asm_face<VARIANT, FORM, EXTRA>(D, A, f, M.f2e[f], C, true, load, RowValsSink{&V}, R);
        fu[f] = R.fux; fu[nf + f] = R.fuy;
        rows_emit_uu(R, V, nf, [&](int comp, int pos, int col, double v) {
            const long long p = rp_uu[(long long)comp * nf + f] + pos; v_uu[p] = v; c_uu[p] = col; });
        rows_emit_up(R, V.kpx, V.kpy, [&](int comp, int pos, int col, double v) {
            const long long p = rp_up[(long long)comp * nf + f] + pos; v_up[p] = v; c_up[p] = col; });
        rows_emit_ex<EXTRA>(R, V, nf, [&](int comp, int pos, int col, double v) {
            const long long p = rp_mu[(long long)comp * nf + f] + pos; v_mu[p] = v; c_mu[p] = col; });
    }
}
}  // namespace

extern "C" {

int emu_geometry(long long nel, const long long* e2v, const long long* e2f, const double* xv, const double* yv,
                 const double* xf, const double* yf, int numbering, double* Om, double* xc, double* yc, double* nx,
                 double* ny, double* G) {
    for (long long e = 0; e < nel; e++) {
        double x[3], y[3], xm[3], ym[3];
        for (int k = 0; k < 3; k++) {
            long long v = e2v[k * nel + e] - 1, f = e2f[k * nel + e] - 1;
            x[k] = xv[v]; y[k] = yv[v]; xm[k] = xf[f]; ym[k] = yf[f];
        }
        ElemGeom g;
        element_geometry(x, y, xm, ym, numbering, g);
        Om[e] = g.Om; xc[e] = g.xc; yc[e] = g.yc;
        for (int k = 0; k < 3; k++) { nx[k * nel + e] = g.nx[k]; ny[k * nel + e] = g.ny[k]; G[k * nel + e] = g.G[k]; }
    }
    return 0;
}

int emu_local(int formulation, double A, long long nel, long long nf, const long long* e2f, const long long* bc,
              const double* Om, const double* nx, const double* ny, const double* G, const double* taue,
              const double* sex, const double* sey, const double* VxDir, const double* VyDir, double* alpha,
              double* beta, double* Z, double* tau) {
    HostMesh H;
    build_mesh(H, nel, nf, e2f, bc, Om, nx, ny, G, Om, Om, taue, tau);
    const MeshView& M = H.view;
    for (long long f = 0; f < nf; f++) { int hi = M.f2e[f].y; if (hi >= 0) tau[f] = taue[hi >> 2]; }   // k_face_tau
    for (int e = 0; e < (int)nel; e++) {                                                              // k_local
        double g[3], n1[3], n2[3], vx[3], vy[3]; bool dir[3];
        for (int i = 0; i < 3; i++) {
            int f = M.e2f[i * nel + e];
            g[i] = G[i * nel + e]; n1[i] = nx[i * nel + e]; n2[i] = ny[i * nel + e];
            dir[i] = M.bc[f] == 1; vx[i] = dir[i] ? VxDir[f] : 0.0; vy[i] = dir[i] ? VyDir[f] : 0.0;
        }
        ElemLocal o;
        if (formulation == 0) element_local<kGradient>(Om[e], taue[e], sex[e], sey[e], g, n1, n2, dir, vx, vy, A, o);
        else element_local<kSymmetricGradient>(Om[e], taue[e], sex[e], sey[e], g, n1, n2, dir, vx, vy, A, o);
        alpha[e] = o.alpha; beta[e] = o.b1; beta[nel + e] = o.b2;
        Z[e] = o.Z11; Z[nel + e] = o.Z21; Z[2 * nel + e] = o.Z12; Z[3 * nel + e] = o.Z22;
    }
    return 0;
}

int emu_assemble(int variant, int formulation, double A, long long nel, long long nf, const long long* e2f,
                 const long long* bc, const double* Om, const double* nx, const double* ny, const double* G,
                 const double* ke, const double* dl, const double* tau, const double* alpha, const double* beta,
                 const double* Z, const double* VxDir, const double* VyDir, const double* sxx, const double* syy,
                 const double* sxy, const double* syx, long long* rp_uu, int* c_uu, double* v_uu, long long* rp_up,
                 int* c_up, double* v_up, long long* rp_mu, int* c_mu, double* v_mu, double* fu, double* fp) {
    HostMesh H;
    build_mesh(H, nel, nf, e2f, bc, Om, nx, ny, G, ke, dl, ke, tau);
    LocalView L{alpha, beta, Z};
    FaceDataView D{VxDir, VyDir, sxx, syy, sxy, syx};
#define ARGS H.view, L, D, A, 0.0, rp_uu, c_uu, v_uu, rp_up, c_up, v_up, rp_mu, c_mu, v_mu, fu, fp
    if (variant == 0 && formulation == 0) assemble<kLoopOld, kGradient, kExtraMuu>(ARGS);
    else if (variant == 0) assemble<kLoopOld, kSymmetricGradient, kExtraMuu>(ARGS);
    else if (formulation == 0) assemble<kLoopNew, kGradient, kExtraMuu>(ARGS);
    else assemble<kLoopNew, kSymmetricGradient, kExtraMuu>(ARGS);
#undef ARGS
    return 0;
}

// same with the Schur complement (penalty gamma) as third block
int emu_assemble_schur(int variant, int formulation, double A, double gamma, long long nel, long long nf, const long long* e2f,
                       const long long* bc, const double* Om, const double* nx, const double* ny, const double* G,
                       const double* ke, const double* dl, const double* tau, const double* alpha, const double* beta,
                       const double* Z, const double* VxDir, const double* VyDir, const double* sxx, const double* syy,
                       const double* sxy, const double* syx, long long* rp_uu, int* c_uu, double* v_uu, long long* rp_up,
                       int* c_up, double* v_up, long long* rp_mu, int* c_mu, double* v_mu, double* fu, double* fp) {
    HostMesh H;
    build_mesh(H, nel, nf, e2f, bc, Om, nx, ny, G, ke, dl, ke, tau);
    LocalView L{alpha, beta, Z};
    FaceDataView D{VxDir, VyDir, sxx, syy, sxy, syx};
#define ARGS H.view, L, D, A, gamma, rp_uu, c_uu, v_uu, rp_up, c_up, v_up, rp_mu, c_mu, v_mu, fu, fp
    if (variant == 0 && formulation == 0) assemble<kLoopOld, kGradient, kExtraSchur>(ARGS);
    else if (variant == 0) assemble<kLoopOld, kSymmetricGradient, kExtraSchur>(ARGS);
    else if (formulation == 0) assemble<kLoopNew, kGradient, kExtraSchur>(ARGS);
    else assemble<kLoopNew, kSymmetricGradient, kExtraSchur>(ARGS);
#undef ARGS
    return 0;
}

int emu_residuals(int formulation, int tau_mode, long long nel, long long nf, const long long* e2f, const long long* bc,
                  const double* Om, const double* nx, const double* ny, const double* G, const double* ke,
                  const double* dl, const double* tau, const double* alpha, const double* beta, const double* Z,
                  const double* Vxh, const double* Vyh, const double* Pe, const double* sex, const double* sey,
                  const double* VxDir, const double* VyDir, const double* sxx, const double* syy, const double* sxy,
                  const double* syx, double* sumsq, double* Fx, double* Fy, double* Fg2) {
    HostMesh H;
    build_mesh(H, nel, nf, e2f, bc, Om, nx, ny, G, ke, dl, ke, tau);
    const MeshView& M = H.view;
    std::vector<double> Fg(6 * nel);
    for (int q = 0; q < 6; q++) sumsq[q] = 0.0;
    for (int e = 0; e < (int)nel; e++) {   // k_res_elem
        double g[3], n1[3], n2[3], t[3], ux[3], uy[3]; int b[3], fid[3];
        for (int i = 0; i < 3; i++) {
            int f = M.e2f[i * nel + e];
            g[i] = G[i * nel + e]; n1[i] = nx[i * nel + e]; n2[i] = ny[i * nel + e]; b[i] = M.bc[f];
            t[i] = tau_mode == 0 ? tau[e] : tau[f];
            ux[i] = b[i] == 1 ? VxDir[f] : Vxh[f]; uy[i] = b[i] == 1 ? VyDir[f] : Vyh[f]; fid[i] = f;
        }
        auto traction = [&](int i, double& tx, double& ty) {
            const int f = fid[i];
            tx = n1[i] * sxx[f] + n2[i] * sxy[f]; ty = n1[i] * syx[f] + n2[i] * syy[f];
        };
        ElemRes r;
        if (formulation == 0)
            element_residual<kGradient>(Om[e], ke[e], dl[e], alpha[e], beta[e], beta[nel + e], Z[e], Z[nel + e], Z[2 * nel + e],
                                        Z[3 * nel + e], sex[e], sey[e], Pe[e], g, n1, n2, b, t, ux, uy, traction, r);
        else
            element_residual<kSymmetricGradient>(Om[e], ke[e], dl[e], alpha[e], beta[e], beta[nel + e], Z[e], Z[nel + e],
                                                 Z[2 * nel + e], Z[3 * nel + e], sex[e], sey[e], Pe[e], g, n1, n2, b, t, ux, uy,
                                                 traction, r);
        for (int i = 0; i < 3; i++) { Fg[(2 * i) * nel + e] = r.Fg1[i][0]; Fg[(2 * i + 1) * nel + e] = r.Fg1[i][1]; }
        Fg2[e] = r.Fg2;
        sumsq[0] += r.F1[0][0] * r.F1[0][0] + r.F1[1][0] * r.F1[1][0] + r.F1[0][1] * r.F1[0][1] + r.F1[1][1] * r.F1[1][1];
        sumsq[1] += r.F2[0] * r.F2[0] + r.F2[1] * r.F2[1];
        sumsq[2] += r.F3 * r.F3;
        sumsq[5] += r.Fg2 * r.Fg2;
    }
    for (int f = 0; f < (int)nf; f++) {    // k_res_face
        double fx = 0.0, fy = 0.0;
        int2 pr = M.f2e[f];
        if (pr.x >= 0 && M.bc[f] != 1) {
            fx += Fg[(2 * (pr.x & 3)) * nel + (pr.x >> 2)]; fy += Fg[(2 * (pr.x & 3) + 1) * nel + (pr.x >> 2)];
            if (pr.y != pr.x) { fx += Fg[(2 * (pr.y & 3)) * nel + (pr.y >> 2)]; fy += Fg[(2 * (pr.y & 3) + 1) * nel + (pr.y >> 2)]; }
        }
        Fx[f] = fx; Fy[f] = fy;
        sumsq[3] += fx * fx; sumsq[4] += fy * fy;
    }
    return 0;
}

int emu_element_values(long long nel, long long nf, const long long* e2f, const long long* bc, const double* Om, const double* nx,
                       const double* ny, const double* G, const double* ke, const double* dl, const double* tau,
                       const double* alpha, const double* beta, const double* Z, const double* Vxh, const double* Vyh,
                       double* out /* 5*nel */) {
    HostMesh H;
    build_mesh(H, nel, nf, e2f, bc, Om, nx, ny, G, ke, dl, ke, tau);
    const MeshView& M = H.view;
    for (int e = 0; e < (int)nel; e++) {   // k_elem_values
        double g[3], n1[3], n2[3], t[3], uhx[3], uhy[3]; int b[3];
        for (int i = 0; i < 3; i++) {
            int f = M.e2f[i * nel + e];
            g[i] = G[i * nel + e]; n1[i] = nx[i * nel + e]; n2[i] = ny[i * nel + e]; b[i] = M.bc[f]; t[i] = tau[f];
            uhx[i] = Vxh[f]; uhy[i] = Vyh[f];
        }
        ElemValues o;
        element_values(Om[e], ke[e], dl[e], alpha[e], beta[e], beta[nel + e], Z[e], Z[nel + e], Z[2 * nel + e], Z[3 * nel + e],
                       g, n1, n2, b, t, uhx, uhy, o);
        out[e] = o.Vx; out[nel + e] = o.Vy; out[2 * nel + e] = o.Txx; out[3 * nel + e] = o.Tyy; out[4 * nel + e] = o.Txy;
    }
    return 0;
}

// strip layout + local face ids of the device generator (integer logic only)
int emu_strip_connectivity(long long n, long long row0, long long row1, long long* sizes /* nel, nf, rlo, rhi */,
                           int* e2f /* 3*nel or NULL */) {
    StripLayout S = make_strip_layout(n, row0, row1);
    sizes[0] = S.nel; sizes[1] = S.nf; sizes[2] = S.rlo; sizes[3] = S.rhi;
    if (!e2f) return 0;
    for (long long le = 0; le < S.nel; le++) {
        long long ql = le >> 2; int t = (int)(le & 3);
        long long qy = S.rlo + ql / n, qx = ql % n;
        int f[8];
        quad_faces(S, qx, qy, f);
        int f1, f2, f3;
        tri_faces(f, t, f1, f2, f3);
        e2f[le] = f1; e2f[S.nel + le] = f2; e2f[2 * S.nel + le] = f3;
    }
    return 0;
}

}  // extern "C"
