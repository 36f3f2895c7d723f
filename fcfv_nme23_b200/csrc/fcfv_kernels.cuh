// fcfv_kernels.cuh -- CUDA kernels (sm_100a, FP64, no tensor cores: nothing here is a dense
// contraction; every kernel is bound by HBM traffic) for the FCFV hot path.
//
// Data layout in HBM (structure of arrays, exactly the reference's column-major matrices):
//   e2f  int32 [3][nel] 0-based      G, nx, ny  double [3][nel]     Om, ke, dl, taue  double [nel]
//   bc   int8  [nf]                  tau, VxDir, VyDir, s??Neu, Vxh, Vyh  double [nf]
//   f2e  int2  [nf]: (lo, hi) codes 4*e+i of the <=2 (element, local face) pairs of a face
//   alpha [nel], beta [2][nel], Z [4][nel] (Z11, Z21, Z12, Z22)
//   CSR:  rowptr int32|int64 [nrows+1], col int32 [cap], val double [cap]
//
// Thread mappings: element kernels = one thread per element (consecutive threads read
// consecutive entries of every SoA array: fully coalesced 8-byte loads); face kernels = one
// thread per face (coalesced face arrays, gathers of the <=2 adjacent elements hit L1/L2 because
// face ids follow element order).  Assembly is a gather by face row: no atomics on floating
// point, every (row, col) value is produced by one thread with a fixed operation order.
#pragma once
#include "fcfv_math.cuh"

#if !defined(__CUDACC__)
struct int2 { int x, y; };   // host build of the arithmetic (tests/host_emulation.cpp)
#endif

namespace fcfv {

struct MeshView {
    int nel, nf;
    const int* e2f;
    const int8_t* bc;
    const int2* f2e;
    const double *Om, *G, *nx, *ny, *ke, *dl, *taue;
    const double* tau;
    const uint8_t* elem_owned;   // nullable
    const uint8_t* face_owned;   // nullable
    const double* tau_ref;       // nullable (REFERENCE tau mode on a partition)
};

struct LocalView { const double *alpha, *beta, *Z; };
struct FaceDataView { const double *VxDir, *VyDir, *sxx, *syy, *sxy, *syx; };
struct SolutionView { const double *Vxh, *Vyh, *Pe; };

// Loads what element e contributes to the rows of its faces.
template <int VARIANT>
FCFV_HD void load_elem_asm(const MeshView& M, const LocalView& L, int e, ElemAsm& E) {
    const int nel = M.nel;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int g = M.e2f[j * nel + e];
        E.g[j] = g;
        E.bc[j] = M.bc[g];
        E.tau[j] = M.tau[g];
        E.G[j] = M.G[j * nel + e];
        E.nx[j] = M.nx[j * nel + e];
        E.ny[j] = M.ny[j * nel + e];
    }
    const double Om = M.Om[e];
    E.eta = M.ke[e];
    E.ia = 1.0 / L.alpha[e];
    if (VARIANT == kLoopOld) {
        E.iO = 1.0 / Om;
        E.D11 = E.D12 = E.D21 = E.D22 = 0.0;
    } else {
        E.iO = 0.0;
        rheology_new(E.eta, M.dl[e], Om, E.D11, E.D12, E.D21, E.D22);
    }
    E.b1 = L.beta[e]; E.b2 = L.beta[nel + e];
    E.Z11 = L.Z[e]; E.Z21 = L.Z[nel + e]; E.Z12 = L.Z[2 * nel + e]; E.Z22 = L.Z[3 * nel + e];
}

// Everything one face thread needs: both rows of face f, fu, Kup entries; and, as a by-product
// for the elements whose local face 0 is f, their fp value (so fp costs no extra pass).
template <int VARIANT, int FORM, bool WANT_M>
FCFV_HD void gather_face_rows(const MeshView& M, const LocalView& L, const FaceDataView& D, int f, double A,
                              FaceRows& R, double fp_out[2], int fp_elem[2]) {
    face_rows_init(R);
    fp_elem[0] = fp_elem[1] = -1;
    const int2 pr = M.f2e[f];
    if (pr.x < 0) return;
    const double vxf = D.VxDir[f], vyf = D.VyDir[f];
    const double sxx = D.sxx[f], syy = D.syy[f], sxy = D.sxy[f], syx = D.syx[f];
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
        const int code = slot == 0 ? pr.x : pr.y;
        if (slot == 1 && pr.y == pr.x) break;
        const int e = code >> 2, i = code & 3;
        ElemAsm E;
        load_elem_asm<VARIANT>(M, L, e, E);
        add_element_rows<VARIANT, FORM, WANT_M>(E, i, slot, e, A, vxf, vyf, sxx, syy, sxy, syx, R);
        if (i == 0 && (M.elem_owned == nullptr || M.elem_owned[e])) {
            bool dir[3]; double vx[3], vy[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                dir[j] = E.bc[j] == 1;
                vx[j] = dir[j] ? D.VxDir[E.g[j]] : 0.0;
                vy[j] = dir[j] ? D.VyDir[E.g[j]] : 0.0;
            }
            fp_out[slot] = element_fp(E.G, E.nx, E.ny, dir, vx, vy);
            fp_elem[slot] = e;
        }
    }
    face_rows_merge<WANT_M>(R);
}

// Non-zero flags and counts of the compacted rows.
struct RowCounts {
    bool nzxx[6], nzxy[6], nzyx[6], nzyy[6], nzmx[6], nzmy[6];
    int nxx, nxy, nyx, nyy, nmx, nmy;   // per kind
    bool pzx[2], pzy[2];
    int npx, npy;
};

template <bool WANT_M>
FCFV_HD void count_rows(const FaceRows& R, RowCounts& C) {
    C.nxx = C.nxy = C.nyx = C.nyy = C.nmx = C.nmy = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const bool alive = R.key[k] != kDeadKey;
        C.nzxx[k] = alive && R.xx[k] != 0.0; C.nxx += C.nzxx[k];
        C.nzxy[k] = alive && R.xy[k] != 0.0; C.nxy += C.nzxy[k];
        C.nzyx[k] = alive && R.yx[k] != 0.0; C.nyx += C.nzyx[k];
        C.nzyy[k] = alive && R.yy[k] != 0.0; C.nyy += C.nzyy[k];
        if (WANT_M) {
            C.nzmx[k] = alive && R.mxx[k] != 0.0; C.nmx += C.nzmx[k];
            C.nzmy[k] = alive && R.myy[k] != 0.0; C.nmy += C.nzmy[k];
        } else {
            C.nzmx[k] = C.nzmy[k] = false;
        }
    }
    C.npx = C.npy = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        C.pzx[s] = R.pe[s] >= 0 && R.px[s] != 0.0; C.npx += C.pzx[s];
        C.pzy[s] = R.pe[s] >= 0 && R.py[s] != 0.0; C.npy += C.pzy[s];
    }
}

#if defined(__CUDACC__)
// =============================================================================================
// kernels
// =============================================================================================
constexpr int kBlock = 128;          // threads per block for element / face kernels
constexpr int kMaxRow = 10;          // max entries of a Kuu row (5 faces x 2 components)
constexpr int kNumSums = 6;          // Kuu-x, Kuu-y, Kup-x, Kup-y, Muu-x, Muu-y

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Exclusive scan of one int per thread over a block of kBlock threads; returns the offset and the
// block total.  `sw` is shared scratch of kBlock/32 + 1 ints.  Deterministic (integers).
__device__ __forceinline__ int block_excl_scan(int v, int& total, int* sw) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v);
    if (lane == 31) sw[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; k++) {
        const int s = sw[k];
        if (k < w) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

// Sum of one double per thread over the block, fixed tree order (deterministic): result valid in
// thread 0.  `sd` is shared scratch of kBlock/32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sd) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sd[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kBlock / 32; k++) r += sd[k];
    }
    __syncthreads();
    return r;
}

// ---- mesh set-up ---------------------------------------------------------------------------
__global__ void k_convert_e2f(const long long* __restrict__ in, int* __restrict__ out, long long n3, int nf,
                              int* __restrict__ err) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n3) return;
    const long long v = in[k] - 1;
    if (v < 0 || v >= nf) { atomicOr(err, 1); out[k] = 0; return; }
    out[k] = (int)v;
}

__global__ void k_convert_bc(const long long* __restrict__ in, int8_t* __restrict__ out, int nf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const long long v = in[f];
    out[f] = (v == 1 || v == 2 || v == -1) ? (int8_t)v : (int8_t)0;   // other tags behave as interior
}

__global__ void k_f2e_init(int2* __restrict__ f2e, int* __restrict__ cnt, int nf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    f2e[f] = make_int2(0x7fffffff, -1);
    cnt[f] = 0;
}

__global__ void k_f2e_build(const int* __restrict__ e2f, int2* __restrict__ f2e, int* __restrict__ cnt, int nel) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int f = e2f[i * nel + e];
        const int code = 4 * e + i;
        atomicMin(&f2e[f].x, code);   // integer atomics: order-independent result
        atomicMax(&f2e[f].y, code);
        atomicAdd(&cnt[f], 1);
    }
}

__global__ void k_f2e_check(int2* __restrict__ f2e, const int* __restrict__ cnt, int nf, int* __restrict__ err) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int c = cnt[f];
    if (c > 2) atomicOr(err, 2);
    if (c == 0) f2e[f] = make_int2(-1, -1);
}

// ---- a3: tau[f] = taue[highest-numbered adjacent element] (Discretisation:40, last writer) ----
__global__ void k_face_tau(const int2* __restrict__ f2e, const double* __restrict__ taue, double* __restrict__ tau,
                           int nf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int hi = f2e[f].y;
    if (hi >= 0) tau[f] = taue[hi >> 2];
}

// ---- a2 geometry -------------------------------------------------------------------------------
__global__ void k_geometry(int nel, const int* __restrict__ e2v, const int* __restrict__ e2f,
                           const double* __restrict__ xv, const double* __restrict__ yv,
                           const double* __restrict__ xf, const double* __restrict__ yf, int numbering, double tau_r,
                           double* __restrict__ Om, double* __restrict__ xc, double* __restrict__ yc,
                           double* __restrict__ nx, double* __restrict__ ny, double* __restrict__ G,
                           double* __restrict__ tau) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
    double x[3], y[3], xm[3], ym[3];
    int fid[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int v = e2v[k * nel + e];
        x[k] = xv[v]; y[k] = yv[v];
        fid[k] = e2f[k * nel + e];
        xm[k] = xf[fid[k]]; ym[k] = yf[fid[k]];
    }
    ElemGeom g;
    element_geometry(x, y, xm, ym, numbering, g);
    Om[e] = g.Om;
    if (xc) xc[e] = g.xc;
    if (yc) yc[e] = g.yc;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        nx[k * nel + e] = g.nx[k]; ny[k * nel + e] = g.ny[k]; G[k * nel + e] = g.G[k];
        if (tau) tau[fid[k]] = tau_r;   // same value from every writer
    }
}

// ---- a4 ComputeFCFV --------------------------------------------------------------------------
template <int FORM>
__global__ void __launch_bounds__(kBlock)
k_local(MeshView M, const double* __restrict__ sex, const double* __restrict__ sey,
        const double* __restrict__ VxDir, const double* __restrict__ VyDir, double A, double* __restrict__ alpha,
        double* __restrict__ beta, double* __restrict__ Z) {
    const int e = blockIdx.x * kBlock + threadIdx.x;
    const int nel = M.nel;
    if (e >= nel) return;
    double G[3], nx[3], ny[3], vx[3], vy[3];
    bool dir[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int f = M.e2f[i * nel + e];
        G[i] = M.G[i * nel + e]; nx[i] = M.nx[i * nel + e]; ny[i] = M.ny[i * nel + e];
        dir[i] = M.bc[f] == 1;
        vx[i] = dir[i] ? VxDir[f] : 0.0;
        vy[i] = dir[i] ? VyDir[f] : 0.0;
    }
    ElemLocal o;
    element_local<FORM>(M.Om[e], M.taue[e], sex[e], sey[e], G, nx, ny, dir, vx, vy, A, o);
    alpha[e] = o.alpha;
    beta[e] = o.b1; beta[nel + e] = o.b2;
    Z[e] = o.Z11; Z[nel + e] = o.Z21; Z[2 * nel + e] = o.Z12; Z[3 * nel + e] = o.Z22;
}

// ---- a5/a6/a7 assembly: count pass ---------------------------------------------------------------
// sums[q*nblk + b] = number of stored entries block b will write for quantity q.
template <int VARIANT, int FORM, bool WANT_M>
__global__ void __launch_bounds__(kBlock)
k_asm_count(MeshView M, LocalView L, FaceDataView D, double A, int* __restrict__ sums, int nblk) {
    __shared__ int sw[kBlock / 32 + 1];
    const int f = blockIdx.x * kBlock + threadIdx.x;
    int c0 = 0, c1 = 0, c2 = 0;
    if (f < M.nf && (M.face_owned == nullptr || M.face_owned[f])) {
        FaceRows R; double fpv[2]; int fpe[2];
        gather_face_rows<VARIANT, FORM, WANT_M>(M, L, D, f, A, R, fpv, fpe);
        RowCounts C;
        count_rows<WANT_M>(R, C);
        c0 = (C.nxx + C.nxy) | ((C.nyx + C.nyy) << 16);
        c1 = C.npx | (C.npy << 16);
        c2 = C.nmx | (C.nmy << 16);
    }
    int t0, t1, t2;
    block_excl_scan(c0, t0, sw);
    block_excl_scan(c1, t1, sw);
    if (WANT_M) block_excl_scan(c2, t2, sw); else t2 = 0;
    if (threadIdx.x == 0) {
        const int b = blockIdx.x;
        sums[0 * nblk + b] = t0 & 0xffff; sums[1 * nblk + b] = t0 >> 16;
        sums[2 * nblk + b] = t1 & 0xffff; sums[3 * nblk + b] = t1 >> 16;
        sums[4 * nblk + b] = t2 & 0xffff; sums[5 * nblk + b] = t2 >> 16;
    }
}

// ---- scan of the block sums -> global bases (one block per quantity pair) ------------------------
// bases[q*nblk + b] = global offset of block b's first entry of quantity q inside its CSR
// (x rows first, then y rows).  totals[q] = entries of quantity q; totals[6..8] = nnz of Kuu/Kup/Muu.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
__global__ void __launch_bounds__(kScanThreads)
k_scan_bases(const int* __restrict__ sums, long long* __restrict__ bases, int nblk, long long* __restrict__ totals) {
    __shared__ long long swarp[kScanThreads / 32];
    __shared__ long long carry_s;
    const int q = blockIdx.x;   // 0..5
    const int* in = sums + (size_t)q * nblk;
    long long* out = bases + (size_t)q * nblk;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long start = 0; start < nblk; start += (long long)kScanThreads * kScanItems) {
        const long long i0 = start + (long long)threadIdx.x * kScanItems;
        int v[kScanItems];
        long long tsum = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) { v[k] = (i0 + k < nblk) ? in[i0 + k] : 0; tsum += v[k]; }
        long long inc = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) swarp[w] = inc;
        __syncthreads();
        long long wbase = 0, tot = 0;
        for (int k = 0; k < kScanThreads / 32; k++) { const long long s = swarp[k]; if (k < w) wbase += s; tot += s; }
        long long run = carry_s + wbase + inc - tsum;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) { if (i0 + k < nblk) out[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[q] = carry_s;
}

// y-row bases are offset by the x-row total of the same matrix; also emits nnz and the last rowptr entry.
template <typename RP>
__global__ void k_finish_bases(long long* __restrict__ bases, int nblk, long long* __restrict__ totals,
                               RP* __restrict__ rp_uu, RP* __restrict__ rp_up, RP* __restrict__ rp_mu, int nf,
                               int want_m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblk) {
        bases[1 * (size_t)nblk + i] += totals[0];
        bases[3 * (size_t)nblk + i] += totals[2];
        if (want_m) bases[5 * (size_t)nblk + i] += totals[4];
    }
    if (i == 0) {
        totals[6] = totals[0] + totals[1];
        totals[7] = totals[2] + totals[3];
        totals[8] = totals[4] + totals[5];
        rp_uu[2 * (size_t)nf] = (RP)(totals[0] + totals[1]);
        rp_up[2 * (size_t)nf] = (RP)(totals[2] + totals[3]);
        if (want_m) rp_mu[2 * (size_t)nf] = (RP)(totals[4] + totals[5]);
    }
}

// ---- assembly: fill pass ------------------------------------------------------------------------
// Recomputes the rows (cheaper than storing padded rows), takes the block-local exclusive scan of
// the row lengths, stages the block's compacted entries in shared memory and writes them to the
// CSR arrays as contiguous, fully coalesced segments (rows of a block are consecutive).
struct CsrOut { void* rowptr; int* col; double* val; };

template <typename RP>
__device__ __forceinline__ void flush_segment(const double* sval, const int* scol, int count, long long base,
                                              const CsrOut& O) {
    for (int k = threadIdx.x; k < count; k += kBlock) {
        O.val[base + k] = sval[k];
        O.col[base + k] = scol[k];
    }
}

template <int VARIANT, int FORM, bool WANT_M, typename RP>
__global__ void __launch_bounds__(kBlock)
k_asm_fill(MeshView M, LocalView L, FaceDataView D, double A, const long long* __restrict__ bases, int nblk,
           CsrOut Kuu, CsrOut Kup, CsrOut Muu, double* __restrict__ fu, double* __restrict__ fp) {
    __shared__ double sval[kMaxRow * kBlock];
    __shared__ int scol[kMaxRow * kBlock];
    __shared__ int sw[kBlock / 32 + 1];
    const int b = blockIdx.x;
    const int f = b * kBlock + threadIdx.x;
    const int nf = M.nf;
    const bool inrange = f < nf;
    const bool active = inrange && (M.face_owned == nullptr || M.face_owned[f]);
    FaceRows R; RowCounts C;
    double fpv[2]; int fpe[2] = {-1, -1};
    int c0 = 0, c1 = 0, c2 = 0;
    if (inrange) {
        // fp duty is independent of face ownership, so gather whenever the face exists locally
        gather_face_rows<VARIANT, FORM, WANT_M>(M, L, D, f, A, R, fpv, fpe);
        if (fpe[0] >= 0) fp[fpe[0]] = fpv[0];
        if (fpe[1] >= 0) fp[fpe[1]] = fpv[1];
    }
    if (active) {
        count_rows<WANT_M>(R, C);
        c0 = (C.nxx + C.nxy) | ((C.nyx + C.nyy) << 16);
        c1 = C.npx | (C.npy << 16);
        c2 = C.nmx | (C.nmy << 16);
        fu[f] = R.fx;
        fu[nf + f] = R.fy;
    } else if (inrange) {
        fu[f] = 0.0; fu[nf + f] = 0.0;
    }
    int t0, t1, t2 = 0;
    const int o0 = block_excl_scan(c0, t0, sw);
    const int o1 = block_excl_scan(c1, t1, sw);
    int o2 = 0;
    if (WANT_M) o2 = block_excl_scan(c2, t2, sw);
    const long long bux = bases[0 * (size_t)nblk + b], buy = bases[1 * (size_t)nblk + b];
    const long long bpx = bases[2 * (size_t)nblk + b], bpy = bases[3 * (size_t)nblk + b];
    if (inrange) {
        ((RP*)Kuu.rowptr)[f] = (RP)(bux + (o0 & 0xffff));
        ((RP*)Kuu.rowptr)[(size_t)nf + f] = (RP)(buy + (o0 >> 16));
        ((RP*)Kup.rowptr)[f] = (RP)(bpx + (o1 & 0xffff));
        ((RP*)Kup.rowptr)[(size_t)nf + f] = (RP)(bpy + (o1 >> 16));
        if (WANT_M) {
            ((RP*)Muu.rowptr)[f] = (RP)(bases[4 * (size_t)nblk + b] + (o2 & 0xffff));
            ((RP*)Muu.rowptr)[(size_t)nf + f] = (RP)(bases[5 * (size_t)nblk + b] + (o2 >> 16));
        }
    }
    // ---- Kuu x rows: columns (g,x) ascending, then (g,y) ascending
    if (active) {
        const int off = o0 & 0xffff;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (C.nzxx[k]) { const int p = off + sorted_pos(R.key, C.nzxx, k); sval[p] = R.xx[k]; scol[p] = R.key[k]; }
            if (C.nzxy[k]) { const int p = off + C.nxx + sorted_pos(R.key, C.nzxy, k); sval[p] = R.xy[k]; scol[p] = R.key[k] + nf; }
        }
    }
    __syncthreads();
    flush_segment<RP>(sval, scol, t0 & 0xffff, bux, Kuu);
    __syncthreads();
    // ---- Kuu y rows
    if (active) {
        const int off = o0 >> 16;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (C.nzyx[k]) { const int p = off + sorted_pos(R.key, C.nzyx, k); sval[p] = R.yx[k]; scol[p] = R.key[k]; }
            if (C.nzyy[k]) { const int p = off + C.nyx + sorted_pos(R.key, C.nzyy, k); sval[p] = R.yy[k]; scol[p] = R.key[k] + nf; }
        }
    }
    __syncthreads();
    flush_segment<RP>(sval, scol, t0 >> 16, buy, Kuu);
    __syncthreads();
    // ---- Kup: x rows in sval[0..), y rows in sval[2*kBlock..); columns = element ids (slot order ascending)
    if (active) {
        int px = o1 & 0xffff, py = 2 * kBlock + (o1 >> 16);
#pragma unroll
        for (int s = 0; s < 2; s++) {
            if (C.pzx[s]) { sval[px] = R.px[s]; scol[px] = R.pe[s]; px++; }
            if (C.pzy[s]) { sval[py] = R.py[s]; scol[py] = R.pe[s]; py++; }
        }
    }
    __syncthreads();
    flush_segment<RP>(sval, scol, t1 & 0xffff, bpx, Kup);
    flush_segment<RP>(sval + 2 * kBlock, scol + 2 * kBlock, t1 >> 16, bpy, Kup);
    if (WANT_M) {
        __syncthreads();
        // ---- Muu: block diagonal, x rows hold (g,x) columns only, y rows (g,y) only
        if (active) {
            const int offx = o2 & 0xffff, offy = 5 * kBlock + (o2 >> 16);
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (C.nzmx[k]) { const int p = offx + sorted_pos(R.key, C.nzmx, k); sval[p] = R.mxx[k]; scol[p] = R.key[k]; }
                if (C.nzmy[k]) { const int p = offy + sorted_pos(R.key, C.nzmy, k); sval[p] = R.myy[k]; scol[p] = R.key[k] + nf; }
            }
        }
        __syncthreads();
        flush_segment<RP>(sval, scol, t2 & 0xffff, bases[4 * (size_t)nblk + b], Muu);
        flush_segment<RP>(sval + 5 * kBlock, scol + 5 * kBlock, t2 >> 16, bases[5 * (size_t)nblk + b], Muu);
    }
}

// ---- a8 residuals ------------------------------------------------------------------------------
// Element kernel: ue, Le, the three face fluxes F_glob1[e,i,:] (stored for the face kernel), and
// the sums of squares of F_eq1, F_eq2, F_eq3, F_glob2 reduced per block (warp shuffles, fixed order).
// partial[q*nblk + b], q = 0:F_eq1 1:F_eq2 2:F_eq3 3:F_glob2.
template <int FORM>
__global__ void __launch_bounds__(kBlock)
k_res_elem(MeshView M, LocalView L, FaceDataView D, SolutionView S, const double* __restrict__ sex,
           const double* __restrict__ sey, int tau_mode, double* __restrict__ Fg, double* __restrict__ partial,
           int nblk, double* __restrict__ F_glob2_out) {
    __shared__ double sd[kBlock / 32];
    const int e = blockIdx.x * kBlock + threadIdx.x;
    const int nel = M.nel;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
    if (e < nel) {
        double G[3], nx[3], ny[3], tau[3], uhx[3], uhy[3], udx[3], udy[3], tix[3], tiy[3];
        int bc[3];
        const double tau_e = (tau_mode == 0) ? (M.tau_ref ? M.tau_ref[e] : M.tau[e < M.nf ? e : 0]) : 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int f = M.e2f[i * nel + e];
            G[i] = M.G[i * nel + e]; nx[i] = M.nx[i * nel + e]; ny[i] = M.ny[i * nel + e];
            bc[i] = M.bc[f];
            tau[i] = (tau_mode == 0) ? tau_e : M.tau[f];
            uhx[i] = S.Vxh[f]; uhy[i] = S.Vyh[f];
            udx[i] = udy[i] = 0.0; tix[i] = tiy[i] = 0.0;
            if (bc[i] == 1) { udx[i] = D.VxDir[f]; udy[i] = D.VyDir[f]; }
            if (bc[i] == 2) {   // tractions only enter through t*Xi (:59-61,69)
                tix[i] = nx[i] * D.sxx[f] + ny[i] * D.sxy[f];
                tiy[i] = nx[i] * D.syx[f] + ny[i] * D.syy[f];
            }
        }
        ElemRes r;
        element_residual<FORM>(M.Om[e], M.ke[e], M.dl[e], L.alpha[e], L.beta[e], L.beta[nel + e], L.Z[e],
                               L.Z[nel + e], L.Z[2 * nel + e], L.Z[3 * nel + e], sex[e], sey[e], S.Pe[e], G, nx, ny,
                               bc, tau, uhx, uhy, udx, udy, tix, tiy, r);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Fg[(size_t)(2 * i) * nel + e] = r.Fg1[i][0];
            Fg[(size_t)(2 * i + 1) * nel + e] = r.Fg1[i][1];
        }
        if (F_glob2_out) F_glob2_out[e] = r.Fg2;
        if (M.elem_owned == nullptr || M.elem_owned[e]) {
            s1 = ((r.F1[0][0] * r.F1[0][0] + r.F1[1][0] * r.F1[1][0]) + r.F1[0][1] * r.F1[0][1]) + r.F1[1][1] * r.F1[1][1];
            s2 = r.F2[0] * r.F2[0] + r.F2[1] * r.F2[1];
            s3 = r.F3 * r.F3;
            s4 = r.Fg2 * r.Fg2;
        }
    }
    const double t1 = block_sum(s1, sd), t2 = block_sum(s2, sd), t3 = block_sum(s3, sd), t4 = block_sum(s4, sd);
    if (threadIdx.x == 0) {
        const int b = blockIdx.x;
        partial[0 * (size_t)nblk + b] = t1; partial[1 * (size_t)nblk + b] = t2;
        partial[2 * (size_t)nblk + b] = t3; partial[3 * (size_t)nblk + b] = t4;
    }
}

// Face kernel (pass C :88-100): F_nodes[f] = sum of the <=2 element fluxes (non-Dirichlet faces).
__global__ void __launch_bounds__(kBlock)
k_res_face(MeshView M, const double* __restrict__ Fg, double* __restrict__ partial, int nblk,
           double* __restrict__ Fx_out, double* __restrict__ Fy_out) {
    __shared__ double sd[kBlock / 32];
    const int f = blockIdx.x * kBlock + threadIdx.x;
    const int nel = M.nel;
    double sx = 0.0, sy = 0.0;
    if (f < M.nf) {
        double Fx = 0.0, Fy = 0.0;
        const int2 pr = M.f2e[f];
        if (pr.x >= 0 && M.bc[f] != 1) {
            const int e0 = pr.x >> 2, i0 = pr.x & 3;
            Fx = Fx + Fg[(size_t)(2 * i0) * nel + e0];
            Fy = Fy + Fg[(size_t)(2 * i0 + 1) * nel + e0];
            if (pr.y != pr.x) {
                const int e1 = pr.y >> 2, i1 = pr.y & 3;
                Fx = Fx + Fg[(size_t)(2 * i1) * nel + e1];
                Fy = Fy + Fg[(size_t)(2 * i1 + 1) * nel + e1];
            }
        }
        if (Fx_out) Fx_out[f] = Fx;
        if (Fy_out) Fy_out[f] = Fy;
        if (M.face_owned == nullptr || M.face_owned[f]) { sx = Fx * Fx; sy = Fy * Fy; }
    }
    const double tx = block_sum(sx, sd), ty = block_sum(sy, sd);
    if (threadIdx.x == 0) {
        partial[0 * (size_t)nblk + blockIdx.x] = tx;
        partial[1 * (size_t)nblk + blockIdx.x] = ty;
    }
}

// Final reduction of per-block partial sums: block q sums partial[q*stride .. + n) in a fixed order.
__global__ void __launch_bounds__(1024)
k_reduce_partials(const double* __restrict__ partial, int n, int stride, double* __restrict__ out) {
    __shared__ double sd[32];
    const double* p = partial + (size_t)blockIdx.x * stride;
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += 1024) s += p[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) sd[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int k = 0; k < 32; k++) r += sd[k];
        out[blockIdx.x] = r;
    }
}

// ---- a9 Powell-Hestenes -------------------------------------------------------------------------
// ru = (fu - Kuu*u) - Kup*p, one thread per row, products accumulated in ascending column order
// (the order Julia's CSC mat-vec produces for each row), per-block sum of ru^2.
template <typename RP>
__global__ void __launch_bounds__(kBlock)
k_ph_ru(int nrows, const RP* __restrict__ rpu, const int* __restrict__ cu, const double* __restrict__ vu,
        const RP* __restrict__ rpp, const int* __restrict__ cp, const double* __restrict__ vp,
        const double* __restrict__ fu, const double* __restrict__ u, const double* __restrict__ p,
        double* __restrict__ ru, double* __restrict__ partial) {
    __shared__ double sd[kBlock / 32];
    const int r = blockIdx.x * kBlock + threadIdx.x;
    double sq = 0.0;
    if (r < nrows) {
        double s1 = 0.0, s2 = 0.0;
        for (RP k = rpu[r]; k < rpu[r + 1]; k++) s1 += vu[k] * u[cu[k]];
        for (RP k = rpp[r]; k < rpp[r + 1]; k++) s2 += vp[k] * p[cp[k]];
        const double v = (fu[r] - s1) - s2;
        if (ru) ru[r] = v;
        sq = v * v;
    }
    const double t = block_sum(sq, sd);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// (Kpu*u)[e] with Kpu = -Kup', matrix-free from the mesh arrays: rows ascending = x rows of the
// element's faces by face id, then y rows.  Entries that were dropped as exact zeros add +-0.0.
FCFV_HD double kpu_dot_u(const MeshView& M, int e, const double* u) {
    const int nel = M.nel, nf = M.nf;
    int g[3]; double kx[3], ky[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        g[i] = M.e2f[i * nel + e];
        const bool nd = M.bc[g[i]] != 1;
        kx[i] = 0.0 - bmul(nd, M.G[i * nel + e]) * M.nx[i * nel + e];
        ky[i] = 0.0 - bmul(nd, M.G[i * nel + e]) * M.ny[i * nel + e];
    }
    // order of the three faces by id
    int o0 = 0, o1 = 1, o2 = 2, t;
    if (g[o0] > g[o1]) { t = o0; o0 = o1; o1 = t; }
    if (g[o1] > g[o2]) { t = o1; o1 = o2; o2 = t; }
    if (g[o0] > g[o1]) { t = o0; o0 = o1; o1 = t; }
    double s = 0.0;
    s += (-kx[o0]) * u[g[o0]]; s += (-kx[o1]) * u[g[o1]]; s += (-kx[o2]) * u[g[o2]];
    s += (-ky[o0]) * u[nf + g[o0]]; s += (-ky[o1]) * u[nf + g[o1]]; s += (-ky[o2]) * u[nf + g[o2]];
    return s;
}

// mode 0: rp = fp - Kpu*u (+ per-block sum of rp^2);  mode 1: p += (gamma*ke/Om)*(fp - Kpu*u)
__global__ void __launch_bounds__(kBlock)
k_ph_rp(MeshView M, const double* __restrict__ fp, const double* __restrict__ u, int mode, double gamma,
        double* __restrict__ rp, double* __restrict__ p, double* __restrict__ partial) {
    __shared__ double sd[kBlock / 32];
    const int e = blockIdx.x * kBlock + threadIdx.x;
    double sq = 0.0;
    if (e < M.nel) {
        const double v = fp[e] - kpu_dot_u(M, e, u);
        if (mode == 0) { if (rp) rp[e] = v; sq = v * v; }
        else { const double coef = (gamma * M.ke[e]) / M.Om[e]; p[e] = p[e] + coef * v; }
    }
    if (mode == 0) {
        const double t = block_sum(sq, sd);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}

__global__ void k_ph_w(int nel, const double* __restrict__ ke, const double* __restrict__ Om,
                       const double* __restrict__ fp, const double* __restrict__ p, double gamma,
                       double* __restrict__ w) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nel) w[e] = ((gamma * ke[e]) / Om[e]) * fp[e] + p[e];
}

template <typename RP>
__global__ void k_ph_fusc(int nrows, const RP* __restrict__ rpp, const int* __restrict__ cp,
                          const double* __restrict__ vp, const double* __restrict__ fu, const double* __restrict__ w,
                          double* __restrict__ fusc) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double s = 0.0;
    for (RP k = rpp[r]; k < rpp[r + 1]; k++) s += vp[k] * w[cp[k]];
    fusc[r] = fu[r] - s;
}

#endif  // __CUDACC__

}  // namespace fcfv
