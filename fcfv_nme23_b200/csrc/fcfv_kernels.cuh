// fcfv_kernels.cuh -- CUDA kernels (sm_100a, FP64, no tensor cores: nothing here is a dense
// contraction; the path is bound by HBM traffic and gather latency) for the FCFV hot path.
//
// Data layout in HBM (structure of arrays, exactly the reference's column-major matrices):
//   e2f  int32 [3][nel] 0-based      G, nx, ny  double [3][nel]     Om, ke, dl, taue  double [nel]
//   bc   int8  [nf]                  tau, VxDir, VyDir, s??Neu, Vxh, Vyh  double [nf]
//   f2e  int2  [nf]: (lo, hi) codes 4*e+i of the <=2 (element, local face) pairs of a face
//   fconn int4 [nf], fbc int32 [nf]: the other faces of both adjacent elements and the bc tags of all
//        five column faces of a face row; ebc int32 [nel]: bc tags of an element's faces + "I am the
//        highest adjacent element" bits (all three static per mesh)
//   alpha [nel], beta [2][nel], Z [4][nel] (Z11, Z21, Z12, Z22); ia, c11 [nel] per-assembly coefficients
//   CSR:  rowptr int32|int64 [nrows+1], col int32 [cap], val double [cap]
//
// Thread mappings: element kernels = one thread per element (consecutive threads read
// consecutive entries of every SoA array: fully coalesced 8-byte loads); face kernels = one
// thread per face (coalesced face arrays, gathers of the <=2 adjacent elements hit L1/L2 because
// face ids follow element order).  Assembly is a gather by face row: no atomics on floating
// point, every (row, col) value is produced by one thread with a fixed operation order; the
// compacted rows leave the SM through shared-memory staging and TMA bulk stores.
#pragma once
#include "fcfv_math.cuh"

#if !defined(__CUDACC__)
struct int2 { int x, y; };   // host build of the arithmetic (tests/host_emulation.cpp)
struct int4 { int x, y, z, w; };
#endif

namespace fcfv {

// Bounds-checked build (make variant NAME=bounds DEFS=-DFCFV_BOUNDS_CHECK=1; compute-sanitizer is closed on the GPU pool): every
// staged shared-memory index, every global segment the fill pass flushes, every gathered element / face id and every flux slot
// is compared with the capacity of its array; violations are counted, not trapped (fcfv_debug_bounds_violations).
#ifndef FCFV_BOUNDS_CHECK
#define FCFV_BOUNDS_CHECK 0
#endif
#if FCFV_BOUNDS_CHECK && defined(__CUDACC__)
__device__ unsigned long long g_bounds_violations = 0ull;
FCFV_HD void check_index(long long i, long long n) {
#if defined(__CUDA_ARCH__)
    if (i < 0 || i >= n) atomicAdd(&g_bounds_violations, 1ull);
#else
    (void)i; (void)n;
#endif
}
#define FCFV_CHECK_INDEX(i, n) ::fcfv::check_index((long long)(i), (long long)(n))
#else
#define FCFV_CHECK_INDEX(i, n) do { } while (0)
#endif

struct MeshView {
    int nel, nf;
    const int* e2f;
    const int8_t* bc;
    const int2* f2e;
    const int4* fconn;           // per face: the other two faces of the lower (x, y) and higher (z, w) element, -1 when absent
    const int* fbc;              // per face: bc tags of (f, x, y, z, w) packed 4 bits each (tag + 1); bit 20: this rank owns the face
    const int* ebc;              // per element: bc tags of its three faces packed 4 bits each (tag + 1); bits 12..14: highest
                                 // adjacent element of local face j; bit 15: this rank owns the element
    const int* eadj;             // per element: where the residual kernel finds the flux of the element across local face j
                                 // (9 bits each, see k_ebc) + bits 27..29: this rank owns local face j
    const double *Om, *G, *nx, *ny, *ke, *dl, *taue;
    const double* tau;
    // ownership of a partitioned mesh lives in the tag words above (kOwnedFace, kOwnedElem, eadj bits 27..29), which the kernels
    // load anyway: no separate masks, no extra dependent loads
    const double* tau_ref;       // nullable (REFERENCE tau mode on a partition)
};

constexpr int kOwnedFace = 1 << 20;   // in fbc
constexpr int kOwnedElem = 1 << 15;   // in ebc
constexpr int kLateFace = 1 << 21;    // in fbc: F_nodes of this face is not formed inside k_res_elem (its two elements sit in
                                      // different 128-element tiles, or no element references it): k_res_late does it
constexpr int kOrphanFace = 1 << 22;  // in fbc: no element references the face
// eadj slot of local face j (bits 9j .. 9j+8): (thread of the neighbour in this element's 128-element tile) << 2 | (its local
// index of the shared face); local index 3 = no neighbour in the tile: thread field kAdjBoundary = boundary face (the
// element is the face's only one), 0 = the neighbour lives in another tile
constexpr int kAdjNone = 3;

constexpr int kAdjBoundary = 127;
constexpr int kAdjOwnedShift = 27;
struct LocalView { const double *alpha, *beta, *Z; };
// per-element assembly coefficients (k_elem_coef): 1/alpha, c11 (old: eta/Omega, NEW: D11) and, only
// when some delta != 1 (c12 != nullptr), the other three entries of D
// cs = (gamma*ke)/Omega of the Schur complement (nullptr unless that block is assembled)
struct CoefView { const double *ia, *c11, *c12, *c21, *c22, *cs; };
struct FaceDataView { const double *VxDir, *VyDir, *sxx, *syy, *sxy, *syx; };
struct SolutionView { const double *Vxh, *Vyh, *Pe; };

// Everything a face row needs of one adjacent element, loaded once and viewed per component.
struct FaceElem {
    int g[3], bc[3];
    double tau[3], G[3], nx[3], ny[3];
    double ia, c11, c12, c21, c22, cs;
    double b1, b2, Z11, Z21, Z12, Z22;
};

// The record of an element built for row face i holds the row face at position 0: local face j sits
// at position (j == i ? 0 : j == 0 ? i : j) -- a swap, not a rotation.  Nothing downstream depends on
// the order of the two other faces (columns are ranked by face id).

// Column faces of the rows of face f and their bc tags: k[0] = f, k[1..2] = the other faces of the
// lower element, k[3..4] = those of the higher element (static per mesh: fconn / fbc arrays).
struct FaceConn { int k[5], bc[5]; };

FCFV_HD int pack_bc5(const int bc[5]) {
    int p = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) p |= ((bc[c] + 1) & 15) << (4 * c);
    return p;
}

FCFV_HD void unpack_conn(int f, int4 o, int p, FaceConn& C) {
    C.k[0] = f; C.k[1] = o.x; C.k[2] = o.y; C.k[3] = o.z; C.k[4] = o.w;
#pragma unroll
    for (int c = 0; c < 5; c++) C.bc[c] = ((p >> (4 * c)) & 15) - 1;
}

// what k_fconn stores for face f (pr = f2e[f]); bc tags of absent faces are 0
FCFV_HD void face_conn_of(const int* e2f, const int8_t* bc, int nel, int f, int2 pr, int4& o, int& p) {
    int k[5] = {f, -1, -1, -1, -1};
    if (pr.x >= 0) {
        const int e = pr.x >> 2, i = pr.x & 3;
        k[1] = e2f[(i == 1 ? 0 : 1) * nel + e];
        k[2] = e2f[(i == 2 ? 0 : 2) * nel + e];
        if (pr.y != pr.x) {
            const int e1 = pr.y >> 2, i1 = pr.y & 3;
            k[3] = e2f[(i1 == 1 ? 0 : 1) * nel + e1];
            k[4] = e2f[(i1 == 2 ? 0 : 2) * nel + e1];
        }
    }
    int b[5];
#pragma unroll
    for (int c = 0; c < 5; c++) b[c] = k[c] >= 0 ? (int)bc[k[c]] : 0;
    o.x = k[1]; o.y = k[2]; o.z = k[3]; o.w = k[4];
    p = pack_bc5(b);
}

// The element's three faces are loaded with the row face (local index i) at position 0: every later
// index is then a compile-time constant (no local memory).  The values do not depend on the order
// in which the pairs (i, j) are visited.  Face ids and bc tags come from the face's connectivity
// record (slot 0: candidates 0,1,2; slot 1: 0,3,4), so that nothing here waits for e2f.
#ifndef FCFV_NATURAL_COUNT
#define FCFV_NATURAL_COUNT 1   // 1: the count pass loads Γ, n in natural face order (one SoA plane per request) and swaps in registers
#endif

template <int VARIANT, bool ANISO, int EXTRA, bool NATURAL = false>
FCFV_HD void load_face_elem(const MeshView& M, const LocalView& L, const CoefView& Cf, const FaceConn& C, bool second, int e, int i,
                            FaceElem& E) {
    const int nel = M.nel;
    FCFV_CHECK_INDEX(e, nel);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int g = k == 0 ? C.k[0] : (second ? C.k[2 + k] : C.k[k]);
        FCFV_CHECK_INDEX(g, M.nf);
        E.g[k] = g;
        E.bc[k] = k == 0 ? C.bc[0] : (second ? C.bc[2 + k] : C.bc[k]);
        E.tau[k] = M.tau[g];
    }
    if (NATURAL) {
        double G[3], nx[3], ny[3];
#pragma unroll
        for (int j = 0; j < 3; j++) { G[j] = M.G[j * nel + e]; nx[j] = M.nx[j * nel + e]; ny[j] = M.ny[j * nel + e]; }
        // record position 0 <- local face i, position i <- local face 0, the third stays
        E.G[0] = pick3(i, G); E.nx[0] = pick3(i, nx); E.ny[0] = pick3(i, ny);
        E.G[1] = i == 1 ? G[0] : G[1]; E.nx[1] = i == 1 ? nx[0] : nx[1]; E.ny[1] = i == 1 ? ny[0] : ny[1];
        E.G[2] = i == 2 ? G[0] : G[2]; E.nx[2] = i == 2 ? nx[0] : nx[2]; E.ny[2] = i == 2 ? ny[0] : ny[2];
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = k == 0 ? i : (k == i ? 0 : k);
            E.G[k] = M.G[j * nel + e];
            E.nx[k] = M.nx[j * nel + e];
            E.ny[k] = M.ny[j * nel + e];
        }
    }
    E.ia = Cf.ia[e];
    E.c11 = Cf.c11[e];
    E.cs = EXTRA == kExtraSchur ? Cf.cs[e] : 0.0;
    if (VARIANT == kLoopNew && ANISO) { E.c12 = Cf.c12[e]; E.c21 = Cf.c21[e]; E.c22 = Cf.c22[e]; }
    else { E.c12 = E.c11; E.c21 = E.c11; E.c22 = E.c11; }   // delta == 1 everywhere: D12 = D21 = D22 = D11 (old loop: unused)
    E.b1 = L.beta[e]; E.b2 = L.beta[nel + e];
    E.Z11 = L.Z[e]; E.Z21 = L.Z[nel + e]; E.Z12 = L.Z[2 * nel + e]; E.Z22 = L.Z[3 * nel + e];
}

// comp-relative view (register renaming only): nr = normal component of the row's velocity component
FCFV_HD void lane_view(const FaceElem& F, int comp, LaneElem& E) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        E.g[j] = F.g[j]; E.bc[j] = F.bc[j]; E.tau[j] = F.tau[j]; E.G[j] = F.G[j];
        E.nr[j] = comp == 0 ? F.nx[j] : F.ny[j];
        E.no[j] = comp == 0 ? F.ny[j] : F.nx[j];
    }
    E.ia = F.ia; E.c = F.c11; E.cs = F.cs;
    E.Drr = comp == 0 ? F.c11 : F.c22;
    E.Dro = comp == 0 ? F.c12 : F.c21;
    E.br = comp == 0 ? F.b1 : F.b2;
    E.Z1r = comp == 0 ? F.Z11 : F.Z12;   // Z[e,1,comp]
    E.Z2r = comp == 0 ? F.Z21 : F.Z22;   // Z[e,2,comp]
    E.Zr1 = comp == 0 ? F.Z11 : F.Z21;   // Z[e,comp,1]
    E.Zr2 = comp == 0 ? F.Z12 : F.Z22;   // Z[e,comp,2]
}

// Contributions of one adjacent element (already loaded, row face first, F) to rows (f,x), (f,y).
template <int VARIANT, int FORM, int EXTRA>
FCFV_HD void slot_lanes(const FaceElem& F, int e, double A, double vx, double vy, double sxx, double sxy,
                        double syx, double syy, LaneRow& Lx, LaneRow& Ly) {
    LaneElem E;
    lane_view(F, 0, E);
    lane_row<VARIANT, FORM, EXTRA>(E, e, 0, 0, A, vx, sxx, sxy, Lx);   // row face at index 0
    lane_view(F, 1, E);
    lane_row<VARIANT, FORM, EXTRA>(E, e, 0, 1, A, vy, syx, syy, Ly);
}

// fp[e] (Discretisation:194) is non-zero only for elements with a Dirichlet face; everything else
// keeps the +0.0 the array is initialised with.  It is produced by the thread of the element's
// FIRST Dirichlet face in local order (one writer per element) in a pass over the faces that
// skips every non-Dirichlet face at once (k_asm_fp).  Returns false when face f (local index i of
// element e) is not that face.
FCFV_HD bool face_fp(const MeshView& M, const FaceDataView& D, int e, int i, double& fp) {
    const int nel = M.nel;
    bool dir[3]; double G[3], nx[3], ny[3], dvx[3], dvy[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int g = M.e2f[j * nel + e];
        dir[j] = M.bc[g] == 1;
        G[j] = M.G[j * nel + e]; nx[j] = M.nx[j * nel + e]; ny[j] = M.ny[j * nel + e];
        dvx[j] = dir[j] ? D.VxDir[g] : 0.0;
        dvy[j] = dir[j] ? D.VyDir[g] : 0.0;
    }
    const bool first = pick3(i, dir) && !(i >= 1 && dir[0]) && !(i == 2 && dir[1]);
    if (!first) return false;
    fp = element_fp(G, nx, ny, dir, dvx, dvy);
    return true;
}

// Both rows of face f -- what one assembly thread computes; shared by the count pass, the fill
// pass and the host emulation.  pr = f2e[f], C = connectivity record of f; load(second, e, i, F)
// provides the record of element e with row face i in front (second: take face ids / tags of the
// higher element); the values leave through `sink` (see RowVals).
template <int VARIANT, int FORM, int EXTRA, typename Loader, typename Sink>
FCFV_HD void asm_face(const FaceDataView& D, double A, int f, int2 pr, const FaceConn& C, bool owned, Loader&& load,
                      const Sink& sink, FaceRows& R) {
    const int bcf = C.bc[0];
    const bool has0 = pr.x >= 0, has1 = has0 && pr.y != pr.x;
    if (has0 && owned) {
        rows_begin<Sink::kRanked>(R, C.k);
        double vx = 0.0, vy = 0.0, sxx = 0.0, sxy = 0.0, syx = 0.0, syy = 0.0;
        if (bcf == 1) { vx = D.VxDir[f]; vy = D.VyDir[f]; }
        if (bcf == 2) { sxx = D.sxx[f]; sxy = D.sxy[f]; syx = D.syx[f]; syy = D.syy[f]; }   // tractions only enter through t*Xi
        LaneRow Lx, Ly;
        {
            const int e = pr.x >> 2, i = pr.x & 3;
            FaceElem F;
            load(false, e, i, F);
            slot_lanes<VARIANT, FORM, EXTRA>(F, e, A, vx, vy, sxx, sxy, syx, syy, Lx, Ly);
            rows_put0<EXTRA>(R, true, Lx, Ly, sink);
        }
        {
            // No branch around the second element: a missing one re-reads the first and is discarded,
            // so that the loads of both records are independent of any branch and can be issued up front.
            const int code = has1 ? pr.y : pr.x;
            const int e = code >> 2, i = code & 3;
            FaceElem F;
            load(has1, e, i, F);
            slot_lanes<VARIANT, FORM, EXTRA>(F, e, A, vx, vy, sxx, sxy, syx, syy, Lx, Ly);
            rows_put1<EXTRA>(R, has1, true, Lx, Ly, sink);
        }
        rows_finish<EXTRA>(R, true, sink);
    } else {
        rows_clear(R);   // faces without elements, faces owned by another rank
    }
}

#if defined(__CUDACC__)
// =============================================================================================
// kernels
// =============================================================================================
constexpr int kBlock = 128;          // threads per block for element / face kernels
static_assert(kBlock == 128, "eadj slots and kLateFace are built for 128-element tiles");
#ifndef FCFV_ASM_BLOCK
#define FCFV_ASM_BLOCK 128
#endif
constexpr int kAsmBlock = FCFV_ASM_BLOCK;   // assembly: one thread per face (both rows); <= 128 (packed block counts)
constexpr int kAsmFaces = kAsmBlock;
constexpr int kMaxRow = 10;          // max entries of a Kuu row (5 faces x 2 components)
constexpr int kNumSums = 6;          // Kuu-x, Kuu-y, Kup-x, Kup-y, Muu-x, Muu-y
constexpr int kScanChunk = 4096;     // block sums scanned per chunk
#ifndef FCFV_RES_MINBLOCKS
#define FCFV_RES_MINBLOCKS 6         // resident 128-thread residual blocks per SM (80 registers, 56 bytes spilled: 2.4 % faster than 5 x 96)
#endif
#ifndef FCFV_ASM_COUNT_MINBLOCKS
#define FCFV_ASM_COUNT_MINBLOCKS 7   // resident 128-thread blocks per SM the register budget of the count pass is tuned for (<= 72 registers)
#endif
#ifndef FCFV_ASM_FILL_MINBLOCKS
#define FCFV_ASM_FILL_MINBLOCKS 4    // same for the fill pass (<= 128 registers)
#endif

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Exclusive scan of one int per thread over a block of NT threads; returns the offset and the
// block total.  `sw` is shared scratch of NT/32 ints.  Deterministic (integers).
template <int NT>
__device__ __forceinline__ int block_excl_scan(int v, int& total, int* sw) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v);
    if (lane == 31) sw[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; k++) {
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

// N sums at once with a single barrier (same operations per quantity as block_sum: identical results).  `sd`
// holds N * kBlock/32 doubles and must not be reused by the caller afterwards without a barrier.
template <int N>
__device__ __forceinline__ void block_sums(double (&v)[N], double* sd) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < N; q++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], d);
        if (lane == 0) sd[q * (kBlock / 32) + w] = v[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < N; q++) {
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < kBlock / 32; k++) r += sd[q * (kBlock / 32) + k];
            v[q] = r;
        }
    }
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

// Ids of the Dirichlet (bc == 1) and Neumann (bc == 2) faces, appended in any order (the host sorts the short lists);
// counts[0..1] keep counting past `cap`, which tells the host that a list did not fit.
__global__ void k_bc_face_lists(int nf, const int8_t* __restrict__ bc, int cap, int* __restrict__ dir_idx, int* __restrict__ neu_idx,
                                int* __restrict__ counts) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int t = bc[f];
    if (t == 1) { const int k = atomicAdd(&counts[0], 1); if (k < cap) dir_idx[k] = f; }
    else if (t == 2) { const int k = atomicAdd(&counts[1], 1); if (k < cap) neu_idx[k] = f; }
}

// dst[idx[k]] = packed[k]: the values of a per-face array on the listed faces
__global__ void k_scatter_faces(int n, const int* __restrict__ idx, const double* __restrict__ packed, double* __restrict__ dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[idx[k]] = packed[k];       // idx comes from k_bc_face_lists (< nf by construction)
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

// err bit 1: face referenced more than twice; bit 2: the two references come from the same element
// or the two elements share more than one face (the row gather assumes one shared column).
__global__ void k_f2e_check(int2* __restrict__ f2e, const int* __restrict__ cnt, const int* __restrict__ e2f, int nel,
                            int nf, int* __restrict__ err) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int c = cnt[f];
    if (c > 2) atomicOr(err, 2);
    if (c == 0) { f2e[f] = make_int2(-1, -1); atomicOr(err, 8); }      // bit 8: not an error, the mesh has faces without elements
    if (c == 2) {
        const int2 pr = f2e[f];
        const int e0 = pr.x >> 2, e1 = pr.y >> 2;
        int shared = 0;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) shared += e2f[a * nel + e0] == e2f[b * nel + e1];
        if (e0 == e1 || shared != 1) atomicOr(err, 4);
    }
}

// faces without elements: (-1, -1) (renumbering utility; k_f2e_check does it together with the validation)
__global__ void k_f2e_orphans(int2* __restrict__ f2e, const int* __restrict__ cnt, int nf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nf && cnt[f] == 0) f2e[f] = make_int2(-1, -1);
}

// per-face connectivity record of the assembly (static per mesh; rebuilt when bc or e2f change)
__global__ void k_fconn(const int* __restrict__ e2f, const int8_t* __restrict__ bc, const int2* __restrict__ f2e, int nel, int nf,
                        int4* __restrict__ fconn, int* __restrict__ fbc) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    int4 o; int p;
    face_conn_of(e2f, bc, nel, f, f2e[f], o, p);
    fconn[f] = o;
    const int2 pr = f2e[f];
    const bool late = pr.x < 0 || (pr.y != pr.x && ((pr.x >> 2) >> 7) != ((pr.y >> 2) >> 7));
    fbc[f] = p | kOwnedFace | (late ? kLateFace : 0) | (pr.x < 0 ? kOrphanFace : 0);        // everything is owned until fcfv_set_ownership says otherwise (k_pack_owned)
}

// per-element copy of the bc tags of its faces (static per mesh): the element kernels know the tags
// together with e2f instead of one dependent gather later.  Bits 12..14: the element is the
// highest-numbered one adjacent to its local face j, i.e. the last writer of mesh.τ[face]
// (Discretisation:40), so that ComputeFCFV can set τ from the element side.
__global__ void k_ebc(const int* __restrict__ e2f, const int8_t* __restrict__ bc, const int2* __restrict__ f2e, int nel,
                      int* __restrict__ ebc, int* __restrict__ eadj) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
    int p = 0, a = 7 << kAdjOwnedShift;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int f = e2f[j * nel + e];
        const int2 pr = f2e[f];
        p |= (((int)bc[f] + 1) & 15) << (4 * j);
        p |= ((pr.y >> 2) == e ? 1 : 0) << (12 + j);
        // the element on the other side of face j, as seen from the residual kernel's 128-element tile
        int slot;
        if (pr.x == pr.y) slot = (kAdjBoundary << 2) | kAdjNone;
        else {
            const int other = pr.x == 4 * e + j ? pr.y : pr.x;
            const int oe = other >> 2;
            slot = (oe >> 7) == (e >> 7) ? (((oe & 127) << 2) | (other & 3)) : kAdjNone;
        }
        a |= slot << (9 * j);
    }
    ebc[e] = p | kOwnedElem;
    eadj[e] = a;
}

// face ownership of a partitioned mesh as seen from the elements: bits 27..29 of eadj (nullptr: everything owned)
__global__ void k_pack_face_owned(int nel, const int* __restrict__ e2f, const uint8_t* __restrict__ face_owned, int* __restrict__ eadj) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
    int a = eadj[e] & ~(7 << kAdjOwnedShift);
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (face_owned == nullptr || face_owned[e2f[j * nel + e]]) a |= 1 << (kAdjOwnedShift + j);
    eadj[e] = a;
}

// ownership masks of a partitioned mesh -> bit kOwnedFace of fbc / kOwnedElem of ebc (nullptr: everything owned)
__global__ void k_pack_owned(int n, const uint8_t* __restrict__ owned, int bit, int* __restrict__ packed) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int v = packed[k] & ~bit;
    packed[k] = (owned == nullptr || owned[k]) ? (v | bit) : v;
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
// COEF: kLoopOld / kLoopNew = also write the per-element assembly coefficients 1/alpha and c11 of that loop (what
// k_elem_coef would compute from alpha in a pass of its own; delta == 1 everywhere), -1 = alpha / beta / Z only.
template <int FORM, int COEF>
__global__ void __launch_bounds__(kBlock)
k_local(MeshView M, const double* __restrict__ sex, const double* __restrict__ sey,
        const double* __restrict__ VxDir, const double* __restrict__ VyDir, double A, double* __restrict__ alpha,
        double* __restrict__ beta, double* __restrict__ Z, double* __restrict__ tau, double* __restrict__ ia,
        double* __restrict__ c11, int e0, int e1) {
    // elements [e0, e1): the whole mesh, or one chunk of a ComputeFCFV whose inputs are still arriving (fcfv_compute_fcfv)
    const int e = e0 + blockIdx.x * kBlock + threadIdx.x;
    const int nel = M.nel;
    if (e >= e1) return;
    double G[3], nx[3], ny[3], vx[3], vy[3];
    bool dir[3];
    int fid[3];
    const int tags = M.ebc[e];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int f = M.e2f[i * nel + e];
        G[i] = M.G[i * nel + e]; nx[i] = M.nx[i * nel + e]; ny[i] = M.ny[i * nel + e];
        dir[i] = (((tags >> (4 * i)) & 15) - 1) == 1;
        vx[i] = dir[i] ? VxDir[f] : 0.0;
        vy[i] = dir[i] ? VyDir[f] : 0.0;
        fid[i] = f;
    }
    const double te = M.taue[e];
    // mesh.τ[face] = τe of the last writer (:40) = the highest adjacent element: one writer per face, no race
#pragma unroll
    for (int i = 0; i < 3; i++)
        if ((tags >> (12 + i)) & 1) tau[fid[i]] = te;
    ElemLocal o;
    const double Om = M.Om[e];
    element_local<FORM>(Om, te, sex[e], sey[e], G, nx, ny, dir, vx, vy, A, o);
    if (COEF >= 0) {
        double a, d11, d12, d21, d22;
        element_coef<COEF < 0 ? kLoopOld : COEF>(o.alpha, M.ke[e], 1.0, Om, a, d11, d12, d21, d22);
        ia[e] = a; c11[e] = d11;
    }
    alpha[e] = o.alpha;
    beta[e] = o.b1; beta[nel + e] = o.b2;
    Z[e] = o.Z11; Z[nel + e] = o.Z21; Z[2 * nel + e] = o.Z12; Z[3 * nel + e] = o.Z22;
}

// ---- per-element assembly coefficients -------------------------------------------------------------
// flag |= 1 when some delta != 1 (the NEW loop then needs the full 2x2 D per element)
__global__ void k_delta_flag(const double* __restrict__ dl, int nel, int* __restrict__ flag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nel && dl[e] != 1.0) atomicOr(flag, 1);
}

template <int VARIANT>
__global__ void __launch_bounds__(256)
k_elem_coef(int nel, const double* __restrict__ alpha, const double* __restrict__ ke, const double* __restrict__ dl,
            const double* __restrict__ Om, double gamma, double* __restrict__ ia, double* __restrict__ c11, double* __restrict__ c12,
            double* __restrict__ c21, double* __restrict__ c22, double* __restrict__ cs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
    double a, d11, d12, d21, d22;
    element_coef<VARIANT>(alpha[e], ke[e], (VARIANT == kLoopNew && c12 != nullptr) ? dl[e] : 1.0, Om[e], a, d11, d12, d21, d22);
    ia[e] = a; c11[e] = d11;
    if (VARIANT == kLoopNew && c12 != nullptr) { c12[e] = d12; c21[e] = d21; c22[e] = d22; }
    if (cs != nullptr) cs[e] = (gamma * ke[e]) / Om[e];   // penalty.*mesh.ke./mesh.Ω (SolversFCFV_Stokes.jl:23)
}

// delta == 1 everywhere, no Schur block, nel even: two elements per thread with 16-byte loads / stores
template <int VARIANT>
__global__ void __launch_bounds__(256)
k_elem_coef2(int nel2, const double2* __restrict__ alpha, const double2* __restrict__ ke, const double2* __restrict__ Om,
             double2* __restrict__ ia, double2* __restrict__ c11) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nel2) return;
    const double2 a = alpha[k], e = ke[k], o = Om[k];
    double2 ra, rc;
    double d12, d21, d22;
    element_coef<VARIANT>(a.x, e.x, 1.0, o.x, ra.x, rc.x, d12, d21, d22);
    element_coef<VARIANT>(a.y, e.y, 1.0, o.y, ra.y, rc.y, d12, d21, d22);
    ia[k] = ra;
    c11[k] = rc;
}

// ---- a5/a6/a7 assembly --------------------------------------------------------------------------
// A block owns a tile of kAsmFaces consecutive faces; a thread owns one face = CSR rows (f,x) and
// (f,y).  It gathers the records of the <=2 adjacent elements (row face swapped to index 0, so
// that every later index is a compile-time constant), merges the shared column, ranks
// the <=5 column faces and drops exact zeros in registers (FaceRows).  Both passes run the same
// code; the count pass only reduces the row lengths of a block, the fill pass places the entries.
//
// The gathers go through L1/L2 (face ids follow element order, so neighbouring threads share lines).
constexpr int kHalf = kMaxRow * kAsmFaces + 8;        // staging capacity per component (+ alignment shift)
constexpr int kStageBytes = 2 * kHalf * (8 + 4);
// rows of the block's face of this thread
template <int VARIANT, int FORM, int EXTRA, bool ANISO, bool NATURAL, typename Sink>
__device__ __forceinline__ void tile_face_rows(const MeshView& M, const LocalView& L, const CoefView& Cf, const FaceDataView& D, double A,
                                               const Sink& sink, int& f, bool& inrange, FaceRows& R) {
    f = blockIdx.x * kAsmFaces + threadIdx.x;
    inrange = f < M.nf;
    const int2 pr = inrange ? M.f2e[f] : make_int2(-1, -1);
    FaceConn C;
    const int tags = inrange ? M.fbc[f] : 0x11111;
    unpack_conn(f, inrange ? M.fconn[f] : make_int4(-1, -1, -1, -1), tags, C);
    const bool owned = inrange && (tags & kOwnedFace) != 0;
    auto load = [&](bool second, int e, int i, FaceElem& F) { load_face_elem<VARIANT, ANISO, EXTRA, NATURAL>(M, L, Cf, C, second, e, i, F); };
    asm_face<VARIANT, FORM, EXTRA>(D, A, f, pr, C, owned, load, sink, R);
}

// fp: one thread per 4 faces (one 16-byte load of the packed tags), only Dirichlet faces do anything
__global__ void __launch_bounds__(256)
k_asm_fp(MeshView M, FaceDataView D, double* __restrict__ fp) {
    const int f0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (f0 >= M.nf) return;
    int tags[4] = {0x11111, 0x11111, 0x11111, 0x11111};
    if (f0 + 3 < M.nf) {
        const int4 t = *reinterpret_cast<const int4*>(M.fbc + f0);
        tags[0] = t.x; tags[1] = t.y; tags[2] = t.z; tags[3] = t.w;
    } else {
        for (int q = 0; q < 4 && f0 + q < M.nf; q++) tags[q] = M.fbc[f0 + q];
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (((tags[q] & 15) - 1) != 1) continue;
        const int2 pr = M.f2e[f0 + q];
        if (pr.x < 0) continue;
        for (int s2 = 0; s2 < 2; s2++) {
            if (s2 == 1 && pr.y == pr.x) break;
            const int code = s2 == 0 ? pr.x : pr.y;
            const int e = code >> 2, i = code & 3;
            if (!(M.ebc[e] & kOwnedElem)) continue;
            double v;
            if (face_fp(M, D, e, i, v)) fp[e] = v;
        }
    }
}

// the same over the list of Dirichlet faces (fcfv_sparse_face_data: a few thousand of the mesh's millions): one thread per
// listed face
__global__ void __launch_bounds__(256)
k_asm_fp_listed(MeshView M, FaceDataView D, const int* __restrict__ dir_idx, int n, double* __restrict__ fp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int f = dir_idx[k];
    const int2 pr = M.f2e[f];
    if (pr.x < 0) return;
    for (int s2 = 0; s2 < 2; s2++) {
        if (s2 == 1 && pr.y == pr.x) break;
        const int code = s2 == 0 ? pr.x : pr.y;
        const int e = code >> 2, i = code & 3;
        if (!(M.ebc[e] & kOwnedElem)) continue;
        double v;
        if (face_fp(M, D, e, i, v)) fp[e] = v;
    }
}

// row lengths of a face packed in one 64-bit word: Kuu x, Kuu y (11 bits: <= 1280 per block),
// Kup x, Kup y (9 bits: <= 256), third block x, y (11 bits)
__device__ __forceinline__ int cshift(int q) { return q == 0 ? 0 : q == 1 ? 11 : q == 2 ? 22 : q == 3 ? 31 : q == 4 ? 40 : 51; }
__device__ __forceinline__ int cfield(unsigned long long c, int q) {
    const unsigned wd = (q == 2 || q == 3) ? 9u : 11u;
    return (int)((c >> cshift(q)) & ((1u << wd) - 1u));
}
template <int EXTRA>
__device__ __forceinline__ unsigned long long rows_counts(const FaceRows& R) {
    unsigned long long c = (unsigned long long)rows_nuu_x(R) | ((unsigned long long)rows_nuu_y(R) << cshift(1)) |
                           ((unsigned long long)rows_nup_x(R) << cshift(2)) | ((unsigned long long)rows_nup_y(R) << cshift(3));
    if (EXTRA != kExtraNone) c |= ((unsigned long long)rows_nex_x<EXTRA>(R) << cshift(4)) | ((unsigned long long)rows_nex_y<EXTRA>(R) << cshift(5));
    return c;
}

// exclusive scan of one packed count per thread over the block; sw64: NT/32 entries
template <int NT>
__device__ __forceinline__ unsigned long long block_excl_scan64(unsigned long long v, unsigned long long& total,
                                                                unsigned long long* sw64) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, dd);
        if (lane >= dd) inc += t;
    }
    if (lane == 31) sw64[w] = inc;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; k++) {
        const unsigned long long s = sw64[k];
        if (k < w) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

// count pass: sums[q*nblk + b] = number of stored entries block b will write for quantity q
template <int VARIANT, int FORM, int EXTRA, bool ANISO>
__global__ void __launch_bounds__(kAsmBlock, FCFV_ASM_COUNT_MINBLOCKS)
k_asm_count(MeshView M, LocalView L, CoefView Cf, FaceDataView D, double A, int* __restrict__ sums, int nblk) {
    __shared__ unsigned long long sw64[kAsmBlock / 32];
    FaceRows R;
    int f; bool inrange;
    tile_face_rows<VARIANT, FORM, EXTRA, ANISO, FCFV_NATURAL_COUNT != 0>(M, L, Cf, D, A, NullSink(), f, inrange, R);
    unsigned long long tot;
    block_excl_scan64<kAsmBlock>(rows_counts<EXTRA>(R), tot, sw64);
    if (threadIdx.x == 0) {
        const int b = blockIdx.x;
#pragma unroll
        for (int q = 0; q < kNumSums; q++) sums[q * (size_t)nblk + b] = cfield(tot, q);
    }
}

// ---- scan of the block sums: chunk-local exclusive scan, then the (few) chunk totals -------------
// grid (nchunks, 6); local[q*nblk + b] = exclusive prefix of sums inside b's chunk.
__global__ void __launch_bounds__(1024)
k_scan_local(const int* __restrict__ sums, int* __restrict__ local, int nblk, int nchunks, long long* __restrict__ chunk_tot) {
    __shared__ int sw[32];
    const int q = blockIdx.y, c = blockIdx.x;
    const int* in = sums + (size_t)q * nblk;
    int* out = local + (size_t)q * nblk;
    const long long i0 = (long long)c * kScanChunk + (long long)threadIdx.x * 4;
    int v[4], tsum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = (i0 + k < nblk) ? in[i0 + k] : 0; tsum += v[k]; }
    int total;
    int run = block_excl_scan<1024>(tsum, total, sw);
#pragma unroll
    for (int k = 0; k < 4; k++) { if (i0 + k < nblk) out[i0 + k] = run; run += v[k]; }
    if (threadIdx.x == 0) chunk_tot[(size_t)q * nchunks + c] = total;
}

// one block: exclusive scan of the chunk totals per quantity; y-row bases are offset by the x-row
// total of the same matrix.  totals[0..5] per quantity, totals[6..9] nnz of Kuu/Kup/Muu/Kuusc; also writes
// the last row pointer of each CSR.
template <typename RP>
__global__ void k_scan_chunks(const long long* __restrict__ chunk_tot, long long* __restrict__ chunk_base, int nchunks,
                              long long* __restrict__ totals, RP* __restrict__ rp_uu, RP* __restrict__ rp_up,
                              RP* __restrict__ rp_mu, int nf, int want_m, int slot_m) {
    __shared__ long long tot[kNumSums];
    const int q = threadIdx.x;
    if (q < kNumSums) {
        long long run = 0;
        for (int c = 0; c < nchunks; c++) { chunk_base[(size_t)q * nchunks + c] = run; run += chunk_tot[(size_t)q * nchunks + c]; }
        tot[q] = run;
        totals[q] = run;
    }
    __syncthreads();
    if (q < kNumSums && (q & 1))
        for (int c = 0; c < nchunks; c++) chunk_base[(size_t)q * nchunks + c] += tot[q - 1];
    if (q == 0) {
        totals[6] = tot[0] + tot[1]; totals[7] = tot[2] + tot[3];
        if (want_m) totals[slot_m] = tot[4] + tot[5];      // 8: Muu, 9: Schur complement (separate slots: an unfetched
                                                           // Muu count survives a later fcfv_schur)
        rp_uu[2 * (size_t)nf] = (RP)(tot[0] + tot[1]);
        rp_up[2 * (size_t)nf] = (RP)(tot[2] + tot[3]);
        if (want_m) rp_mu[2 * (size_t)nf] = (RP)(tot[4] + tot[5]);
    }
}

// ---- fill pass ----------------------------------------------------------------------------------
// Recomputes the rows, takes the block-local exclusive scan of the row lengths, stages the block's
// compacted entries in shared memory (x rows in the lower half, y rows in the upper half) and
// writes them to the CSR arrays as two contiguous, fully coalesced segments.
struct CsrOut { void* rowptr; int* col; double* val; long long cap; };   // cap: entries col / val can hold (bounds-checked build)

// The segment [base, base+count) of a CSR array leaves shared memory through the TMA unit: ONE
// thread issues a bulk copy (cp.async.bulk shared -> global) per array, so the flush costs no LSU
// loads / stores and almost no instructions.  Bulk copies need 16-byte aligned addresses and
// sizes: entry p of the segment is staged at shared index p + (base & 1) for values and
// p + (base & 3) for columns, which aligns shared and global addresses together; the (<= 1 / <= 3)
// leading and trailing entries outside the aligned middle part are written by ordinary stores.
__device__ __forceinline__ void tma_store(void* gdst, const void* ssrc, int bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// makes this thread's shared-memory writes visible to the async proxy (before the barrier that
// precedes the bulk copy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// middle part by TMA (called by one thread)
__device__ __forceinline__ void flush_segment_bulk(const double* sval, const int* scol, int count, long long base, const CsrOut& O) {
    if (count > 0) { FCFV_CHECK_INDEX(base, O.cap); FCFV_CHECK_INDEX(base + count - 1, O.cap); }
    {
        const int sh = (int)(base & 1), end = sh + count;
        const int q0 = 2 * sh, q1 = end & ~1;
        if (q1 > q0) tma_store(O.val + (base - sh) + q0, sval + q0, (q1 - q0) * 8);
    }
    {
        const int sh = (int)(base & 3), end = sh + count;
        const int q0 = (sh + 3) & ~3, q1 = end & ~3;
        if (q1 > q0) tma_store(O.col + (base - sh) + q0, scol + q0, (q1 - q0) * 4);
    }
}

// leading / trailing entries: thread `t` (0..7) of the group handles at most one value and one column
__device__ __forceinline__ void flush_segment_edges(const double* sval, const int* scol, int count, long long base, const CsrOut& O,
                                                    int t) {
    {
        const int sh = (int)(base & 1), end = sh + count;
        double* gv = O.val + (base - sh);
        if (t == 0 && sh && count > 0) gv[1] = sval[1];
        if (t == 1 && (end & 1) && end - 1 >= sh + sh) gv[end - 1] = sval[end - 1];
    }
    {
        const int sh = (int)(base & 3), end = sh + count;
        int* gc = O.col + (base - sh);
        const int head_end = min((sh + 3) & ~3, end);
        if (t < 4) { const int q = sh + t; if (q < head_end) gc[q] = scol[q]; }
        const int tail_beg = max(end & ~3, head_end);
        if (t >= 4) { const int q = tail_beg + (t - 4); if (q < end) gc[q] = scol[q]; }
    }
}

constexpr int kHalfP = 2 * kAsmFaces + 8;              // Kup staging capacity per component
constexpr int kStageBytesP = 2 * kHalfP * (8 + 4);

template <int VARIANT, int FORM, int EXTRA, bool ANISO, typename RP>
__global__ void __launch_bounds__(kAsmBlock, FCFV_ASM_FILL_MINBLOCKS)
k_asm_fill(MeshView M, LocalView L, CoefView Cf, FaceDataView D, double A, const int* __restrict__ local, const long long* __restrict__ chunk_base,
           int nblk, int nchunks, CsrOut Kuu, CsrOut Kup, CsrOut Muu, double* __restrict__ fu) {
    __shared__ __align__(16) unsigned char raw[kStageBytes];      // Kuu (then Muu) staging
    __shared__ __align__(16) unsigned char rawp[kStageBytesP];    // Kup staging
    __shared__ unsigned long long sw64[kAsmBlock / 32];
    __shared__ long long sbase[kNumSums];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int nf = M.nf;
    // global offsets of this tile's segments: fetched by a few threads while the rows are computed,
    // published by the barriers of the block scan
    if (tid < ((EXTRA != kExtraNone) ? 6 : 4)) sbase[tid] = chunk_base[(size_t)tid * nchunks + b / kScanChunk] + local[(size_t)tid * nblk + b];
    FaceRows R;
    RowVals V;
    int f; bool inrange;
    constexpr bool WANT_M = EXTRA != kExtraNone;   // a third block (Muu or the Schur complement) is assembled
    tile_face_rows<VARIANT, FORM, EXTRA, ANISO, false>(M, L, Cf, D, A, RowValsSink{&V}, f, inrange, R);
    if (inrange) {
        fu[f] = R.fux;
        fu[(size_t)nf + f] = R.fuy;
    }
    unsigned long long tot;
    const unsigned long long off = block_excl_scan64<kAsmBlock>(rows_counts<EXTRA>(R), tot, sw64);
    const long long bux = sbase[0], buy = sbase[1], bpx = sbase[2], bpy = sbase[3];
    const long long bmx = WANT_M ? sbase[4] : 0, bmy = WANT_M ? sbase[5] : 0;
    if (inrange) {
        ((RP*)Kuu.rowptr)[f] = (RP)(bux + cfield(off, 0));
        ((RP*)Kuu.rowptr)[(size_t)nf + f] = (RP)(buy + cfield(off, 1));
        ((RP*)Kup.rowptr)[f] = (RP)(bpx + cfield(off, 2));
        ((RP*)Kup.rowptr)[(size_t)nf + f] = (RP)(bpy + cfield(off, 3));
        if (WANT_M) {
            ((RP*)Muu.rowptr)[f] = (RP)(bmx + cfield(off, 4));
            ((RP*)Muu.rowptr)[(size_t)nf + f] = (RP)(bmy + cfield(off, 5));
        }
    }
    double* sval = reinterpret_cast<double*>(raw);
    int* scol = reinterpret_cast<int*>(raw + (size_t)2 * kHalf * 8);
    double* pval = reinterpret_cast<double*>(rawp);
    int* pcol = reinterpret_cast<int*>(rawp + (size_t)2 * kHalfP * 8);
    // ---- stage Kuu (x rows in the lower half of the buffer, y rows in the upper half) and Kup
    {
        double* vx = sval + cfield(off, 0) + (int)(bux & 1);
        double* vy = sval + kHalf + cfield(off, 1) + (int)(buy & 1);
        int* cx = scol + cfield(off, 0) + (int)(bux & 3);
        int* cy = scol + kHalf + cfield(off, 1) + (int)(buy & 3);
        rows_emit_uu(R, V, nf, [&](int comp, int pos, int col, double v) {
            FCFV_CHECK_INDEX((comp == 0 ? vx : vy) + pos - (sval + (comp == 0 ? 0 : kHalf)), kHalf);
            FCFV_CHECK_INDEX((comp == 0 ? cx : cy) + pos - (scol + (comp == 0 ? 0 : kHalf)), kHalf);
            FCFV_CHECK_INDEX(col, 2 * nf);
            if (comp == 0) { vx[pos] = v; cx[pos] = col; }
            else { vy[pos] = v; cy[pos] = col; }
        });
    }
    {
        double* vx = pval + cfield(off, 2) + (int)(bpx & 1);
        double* vy = pval + kHalfP + cfield(off, 3) + (int)(bpy & 1);
        int* cx = pcol + cfield(off, 2) + (int)(bpx & 3);
        int* cy = pcol + kHalfP + cfield(off, 3) + (int)(bpy & 3);
        rows_emit_up(R, V.kpx, V.kpy, [&](int comp, int pos, int col, double v) {   // columns = element ids, ascending
            FCFV_CHECK_INDEX((comp == 0 ? vx : vy) + pos - (pval + (comp == 0 ? 0 : kHalfP)), kHalfP);
            FCFV_CHECK_INDEX((comp == 0 ? cx : cy) + pos - (pcol + (comp == 0 ? 0 : kHalfP)), kHalfP);
            FCFV_CHECK_INDEX(col, M.nel);
            if (comp == 0) { vx[pos] = v; cx[pos] = col; }
            else { vy[pos] = v; cy[pos] = col; }
        });
    }
    fence_async_smem();
    __syncthreads();
    // ---- flush: thread 0 drives the TMA unit, threads 32..63 write the unaligned edge entries
    if (tid == 0) {
        flush_segment_bulk(sval, scol, cfield(tot, 0), bux, Kuu);
        flush_segment_bulk(sval + kHalf, scol + kHalf, cfield(tot, 1), buy, Kuu);
        flush_segment_bulk(pval, pcol, cfield(tot, 2), bpx, Kup);
        flush_segment_bulk(pval + kHalfP, pcol + kHalfP, cfield(tot, 3), bpy, Kup);
        tma_store_commit();
    } else if (tid >= 32 && tid < 64) {
        const int seg = (tid - 32) >> 3, t = (tid - 32) & 7;
        if (seg == 0) flush_segment_edges(sval, scol, cfield(tot, 0), bux, Kuu, t);
        else if (seg == 1) flush_segment_edges(sval + kHalf, scol + kHalf, cfield(tot, 1), buy, Kuu, t);
        else if (seg == 2) flush_segment_edges(pval, pcol, cfield(tot, 2), bpx, Kup, t);
        else flush_segment_edges(pval + kHalfP, pcol + kHalfP, cfield(tot, 3), bpy, Kup, t);
    }
    if (WANT_M) {
        // ---- third block (Muu or the Schur complement) reuses the Kuu buffer
        if (tid == 0) tma_store_wait_read();
        __syncthreads();
        double* vx = sval + cfield(off, 4) + (int)(bmx & 1);
        double* vy = sval + kHalf + cfield(off, 5) + (int)(bmy & 1);
        int* cx = scol + cfield(off, 4) + (int)(bmx & 3);
        int* cy = scol + kHalf + cfield(off, 5) + (int)(bmy & 3);
        rows_emit_ex<EXTRA>(R, V, nf, [&](int comp, int pos, int col, double v) {
            FCFV_CHECK_INDEX((comp == 0 ? vx : vy) + pos - (sval + (comp == 0 ? 0 : kHalf)), kHalf);
            FCFV_CHECK_INDEX((comp == 0 ? cx : cy) + pos - (scol + (comp == 0 ? 0 : kHalf)), kHalf);
            FCFV_CHECK_INDEX(col, 2 * nf);
            if (comp == 0) { vx[pos] = v; cx[pos] = col; }
            else { vy[pos] = v; cy[pos] = col; }
        });
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            flush_segment_bulk(sval, scol, cfield(tot, 4), bmx, Muu);
            flush_segment_bulk(sval + kHalf, scol + kHalf, cfield(tot, 5), bmy, Muu);
            tma_store_commit();
        } else if (tid >= 32 && tid < 48) {
            const int seg = (tid - 32) >> 3, t = (tid - 32) & 7;
            if (seg == 0) flush_segment_edges(sval, scol, cfield(tot, 4), bmx, Muu, t);
            else flush_segment_edges(sval + kHalf, scol + kHalf, cfield(tot, 5), bmy, Muu, t);
        }
    }
    if (tid == 0) tma_store_wait_read();   // shared memory must stay valid until the TMA unit has read it
}

// ---- a8 residuals ------------------------------------------------------------------------------
// Element kernel: ue, Le, the three face fluxes F_glob1[e,i,:] and the sums of squares of F_eq1, F_eq2, F_eq3, F_glob2
// reduced per block (warp shuffles, fixed order).  Pass C of the reference (:88-100, F_nodes[f] = sum of the fluxes of
// the <= 2 elements of a non-Dirichlet face) is formed HERE for every face whose elements sit in the same 128-element
// tile: the fluxes are exchanged through shared memory, the face's highest element adds them (lower element first, like
// the reference's element loop) and squares the result.  Only the fluxes of faces cut by a tile boundary go to memory
// (Fg, [face][side: lower / higher element][component] -- the two contributions of a face share one 32-byte sector) for
// k_res_late.  On a locality-numbered mesh that is one face in six.
// partial[q*stride + b]: q = 0 F_eq1, 1 F_eq2, 2 F_eq3, 3 F_glob2, 4 / 5 F_nodes x / y (this kernel's blocks first, then
// those of k_res_late).
// OUT: the optional arrays F_glob2 / F_nodes are written (second, rare launch); the hot variant carries no face ids.
template <int FORM, bool OUT>
__global__ void __launch_bounds__(kBlock, FCFV_RES_MINBLOCKS)
k_res_elem(MeshView M, LocalView L, FaceDataView D, SolutionView S, const double* __restrict__ sex,
           const double* __restrict__ sey, int tau_mode, double* __restrict__ Fg, double* __restrict__ partial,
           int stride, double* __restrict__ F_glob2_out, double* __restrict__ Fx_out, double* __restrict__ Fy_out) {
    __shared__ double sd[6 * (kBlock / 32)];
    __shared__ double2 sF[3 * kBlock];
    const int e = blockIdx.x * kBlock + threadIdx.x;
    const int nel = M.nel;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, sx = 0.0, sy = 0.0;
    int fid[3] = {0, 0, 0};
    int tags = 0, adj = 0;
    if (e < nel) {
        double G[3], nx[3], ny[3], tau[3], ux[3], uy[3];
        int bc[3];
        const double tau_e = (tau_mode == 0) ? (M.tau_ref ? M.tau_ref[e] : M.tau[e < M.nf ? e : 0]) : 0.0;
        tags = M.ebc[e];
        adj = M.eadj[e];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int f = M.e2f[i * nel + e];
            fid[i] = f;
            G[i] = M.G[i * nel + e]; nx[i] = M.nx[i * nel + e]; ny[i] = M.ny[i * nel + e];
            bc[i] = ((tags >> (4 * i)) & 15) - 1;
            tau[i] = (tau_mode == 0) ? tau_e : M.tau[f];
            ux[i] = bc[i] == 1 ? D.VxDir[f] : S.Vxh[f];   // Dirichlet faces take VDir (:62-66)
            uy[i] = bc[i] == 1 ? D.VyDir[f] : S.Vyh[f];
        }
        auto traction = [&](int i, double& tx, double& ty) {   // Neumann faces only
            const int f = pick3(i, fid);
            const double n1 = pick3(i, nx), n2 = pick3(i, ny);
            tx = n1 * D.sxx[f] + n2 * D.sxy[f];
            ty = n1 * D.syx[f] + n2 * D.syy[f];
        };
        ElemRes r;
        element_residual<FORM>(M.Om[e], M.ke[e], M.dl[e], L.alpha[e], L.beta[e], L.beta[nel + e], L.Z[e],
                               L.Z[nel + e], L.Z[2 * nel + e], L.Z[3 * nel + e], sex[e], sey[e], S.Pe[e], G, nx, ny,
                               bc, tau, ux, uy, traction, r);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double2 v = make_double2(r.Fg1[i][0], r.Fg1[i][1]);
            sF[3 * threadIdx.x + i] = v;
            FCFV_CHECK_INDEX(fid[i], M.nf);
            if (((adj >> (9 * i)) & 511) == kAdjNone)          // cut face: slot of this side (the highest element is side 1)
                reinterpret_cast<double2*>(Fg)[(size_t)2 * fid[i] + ((tags >> (12 + i)) & 1)] = v;
        }
        if (OUT && F_glob2_out) F_glob2_out[e] = r.Fg2;
        if (tags & kOwnedElem) {
            s1 = ((r.F1[0][0] * r.F1[0][0] + r.F1[1][0] * r.F1[1][0]) + r.F1[0][1] * r.F1[0][1]) + r.F1[1][1] * r.F1[1][1];
            s2 = r.F2[0] * r.F2[0] + r.F2[1] * r.F2[1];
            s3 = r.F3 * r.F3;
            s4 = r.Fg2 * r.Fg2;
        }
    }
    __syncthreads();
    if (e < nel) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int slot = (adj >> (9 * i)) & 511;
            if (slot == kAdjNone || !((tags >> (12 + i)) & 1)) continue;       // cut face, or the other element forms the sum
            double Fx = 0.0, Fy = 0.0;
            if ((((tags >> (4 * i)) & 15) - 1) != 1) {                           // :89 Dirichlet faces stay zero
                if ((slot & 3) != kAdjNone) {                                    // lower element first
                    const double2 lo = sF[3 * (slot >> 2) + (slot & 3)];
                    Fx = Fx + lo.x; Fy = Fy + lo.y;
                }
                const double2 me = sF[3 * threadIdx.x + i];
                Fx = Fx + me.x; Fy = Fy + me.y;
            }
            if (OUT && Fx_out) Fx_out[fid[i]] = Fx;
            if (OUT && Fy_out) Fy_out[fid[i]] = Fy;
            if ((adj >> (kAdjOwnedShift + i)) & 1) { sx = sx + Fx * Fx; sy = sy + Fy * Fy; }
        }
    }
    double t[6] = {s1, s2, s3, s4, sx, sy};
    block_sums<6>(t, sd);
    if (threadIdx.x == 0) {
        const int b = blockIdx.x;
#pragma unroll
        for (int q = 0; q < 6; q++) partial[q * (size_t)stride + b] = t[q];
    }
}

// Faces k_res_elem leaves out (kLateFace: cut by a tile boundary, or referenced by no element): one thread per four faces,
// one 16-byte load of the packed tags, then one 32-byte sector per cut face (both fluxes).  partial rows 4 / 5 from block
// offset `b0`.
constexpr int kLateBlock = 128;
constexpr int kLateFaces = 4;
__global__ void __launch_bounds__(kLateBlock)
k_res_late(MeshView M, const double* __restrict__ Fg, double* __restrict__ partial, int stride, int b0,
           double* __restrict__ Fx_out, double* __restrict__ Fy_out) {
    __shared__ double sd[2 * (kLateBlock / 32)];
    const long long f0 = (long long)kLateFaces * (blockIdx.x * kLateBlock + threadIdx.x);
    double sx = 0.0, sy = 0.0;
    if (f0 < M.nf) {
        int tags[kLateFaces] = {0, 0, 0, 0};
        if (f0 + kLateFaces <= M.nf) {
            const int4 t = *reinterpret_cast<const int4*>(M.fbc + f0);
            tags[0] = t.x; tags[1] = t.y; tags[2] = t.z; tags[3] = t.w;
        } else {
            for (int q = 0; q < kLateFaces && f0 + q < M.nf; q++) tags[q] = M.fbc[f0 + q];
        }
#pragma unroll
        for (int q = 0; q < kLateFaces; q++) {
            if (!(tags[q] & kLateFace)) continue;
            const long long f = f0 + q;
            double Fx = 0.0, Fy = 0.0;
            if (!(tags[q] & kOrphanFace) && ((tags[q] & 15) - 1) != 1) {      // :89 Dirichlet faces stay zero
                const double2 v0 = reinterpret_cast<const double2*>(Fg)[2 * f];
                const double2 v1 = reinterpret_cast<const double2*>(Fg)[2 * f + 1];
                Fx = Fx + v0.x; Fy = Fy + v0.y;       // lower element first
                Fx = Fx + v1.x; Fy = Fy + v1.y;
            }
            if (Fx_out) Fx_out[f] = Fx;
            if (Fy_out) Fy_out[f] = Fy;
            if (tags[q] & kOwnedFace) { sx = sx + Fx * Fx; sy = sy + Fy * Fy; }
        }
    }
    double t[2] = {sx, sy};
    block_sums<2>(t, sd);
    if (threadIdx.x == 0) {
        partial[4 * (size_t)stride + b0 + blockIdx.x] = t[0];
        partial[5 * (size_t)stride + b0 + blockIdx.x] = t[1];
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

// The six residual sums in ONE launch: grid (kRedSlices, 6).  Block (r, q) sums slice r of quantity q in a
// fixed order into scratch; the last block of a quantity to finish (ticket counter) adds the kRedSlices slice
// sums in index order -- the result does not depend on which block that is.
// q: 0 F_eq1, 1 F_eq2, 2 F_eq3 (element partial rows 0..2), 3 F_nodes_x, 4 F_nodes_y (face partial rows 4, 5),
// 5 F_glob2 (element partial row 3) = print order of the reference (:101-106).
constexpr int kRedSlices = 64;
__global__ void __launch_bounds__(256)
k_reduce_residuals(const double* __restrict__ partial, int nbe, int nbf, int stride, double* __restrict__ scratch,
                   unsigned int* __restrict__ tickets, double* __restrict__ out) {
    __shared__ double sd[8];
    __shared__ bool last;
    const int q = blockIdx.y, r = blockIdx.x;
    const int row = q < 3 ? q : (q < 5 ? q + 1 : 3);
    const int n = (q == 3 || q == 4) ? nbf : nbe;
    const double* p = partial + (size_t)row * stride;
    const int chunk = (n + kRedSlices - 1) / kRedSlices;
    const int lo = r * chunk, hi = min(n, lo + chunk);
    double s = 0.0;
    for (int k = lo + threadIdx.x; k < hi; k += 256) s += p[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) sd[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += sd[k];
        scratch[q * kRedSlices + r] = t;
        __threadfence();
        last = atomicAdd(&tickets[q], 1u) == kRedSlices - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (int k = 0; k < kRedSlices; k++) t += ((volatile double*)scratch)[q * kRedSlices + k];
        out[q] = t;
        tickets[q] = 0;
    }
}

// ---- ComputeElementValues (Discretisation:52-94) -------------------------------------------------
// out: Vxe, Vye, Txxe, Tyye, Txye as five consecutive arrays of nel doubles
__global__ void __launch_bounds__(kBlock)
k_elem_values(MeshView M, LocalView L, SolutionView S, double* __restrict__ out) {
    const int e = blockIdx.x * kBlock + threadIdx.x;
    const int nel = M.nel;
    if (e >= nel) return;
    double G[3], nx[3], ny[3], tau[3], uhx[3], uhy[3];
    int bc[3];
    const int tags = M.ebc[e];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int f = M.e2f[i * nel + e];
        G[i] = M.G[i * nel + e]; nx[i] = M.nx[i * nel + e]; ny[i] = M.ny[i * nel + e];
        bc[i] = ((tags >> (4 * i)) & 15) - 1; tau[i] = M.tau[f];
        uhx[i] = S.Vxh[f]; uhy[i] = S.Vyh[f];
    }
    ElemValues o;
    element_values(M.Om[e], M.ke[e], M.dl[e], L.alpha[e], L.beta[e], L.beta[nel + e], L.Z[e], L.Z[nel + e], L.Z[2 * nel + e],
                   L.Z[3 * nel + e], G, nx, ny, bc, tau, uhx, uhy, o);
    out[e] = o.Vx; out[(size_t)nel + e] = o.Vy; out[2 * (size_t)nel + e] = o.Txx; out[3 * (size_t)nel + e] = o.Tyy;
    out[4 * (size_t)nel + e] = o.Txy;
}

// ---- CSR -> Julia CSC (SparseMatrixCSC{Float64,Int64}: 1-based colptr / rowval, rows ascending in a column) ----
// The reference's consumers (sparse product, lu, cholesky) take CSC, so the export transposes on the
// device: count per column (integer atomics), exclusive scan, scatter with an atomic cursor, then an
// insertion sort of every column by row (<= 10 entries for Kuu, <= 6 for Kup).  Rows are unique inside a
// column, so the sorted result does not depend on the order in which the scatter ran.
template <typename RP>
__global__ void k_csc_count(int nrows, const RP* __restrict__ rowptr, const int* __restrict__ col, unsigned long long* __restrict__ cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    for (RP k = rowptr[r]; k < rowptr[r + 1]; k++) atomicAdd(&cnt[col[k]], 1ULL);
}

// colptr1[j] = 1 + exclusive prefix (written in place over cnt by the caller's scan); cursor starts at 0
template <typename RP>
__global__ void k_csc_scatter(int nrows, const RP* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                              const long long* __restrict__ colptr1, unsigned int* __restrict__ cursor,
                              long long* __restrict__ rowval, double* __restrict__ nzval) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    for (RP k = rowptr[r]; k < rowptr[r + 1]; k++) {
        const int c = col[k];
        const long long p = (colptr1[c] - 1) + atomicAdd(&cursor[c], 1u);
        rowval[p] = (long long)r + 1;
        nzval[p] = val[k];
    }
}

__global__ void k_csc_sort(int ncols, const long long* __restrict__ colptr1, long long* __restrict__ rowval, double* __restrict__ nzval) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const long long lo = colptr1[c] - 1, hi = colptr1[c + 1] - 1;
    for (long long a = lo + 1; a < hi; a++) {
        const long long r = rowval[a];
        const double v = nzval[a];
        long long b = a;
        while (b > lo && rowval[b - 1] > r) { rowval[b] = rowval[b - 1]; nzval[b] = nzval[b - 1]; b--; }
        rowval[b] = r; nzval[b] = v;
    }
}

// in-place exclusive scan of n counts -> 1-based column pointers (+ the total at [n]); chunks of 1024*4
__global__ void __launch_bounds__(1024)
k_scan64_local(unsigned long long* __restrict__ a, long long n, unsigned long long* __restrict__ chunk_tot) {
    __shared__ unsigned long long sw[32];
    const long long i0 = (long long)blockIdx.x * 4096 + (long long)threadIdx.x * 4;
    unsigned long long v[4], t = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = (i0 + k < n) ? a[i0 + k] : 0ULL; t += v[k]; }
    unsigned long long total;
    unsigned long long run = block_excl_scan64<1024>(t, total, sw);
#pragma unroll
    for (int k = 0; k < 4; k++) { if (i0 + k < n) a[i0 + k] = run; run += v[k]; }
    if (threadIdx.x == 0) chunk_tot[blockIdx.x] = total;
}

__global__ void k_scan64_chunks(unsigned long long* __restrict__ chunk_tot, int nchunks) {   // one thread: few thousand chunks
    unsigned long long run = 0;
    for (int c = 0; c < nchunks; c++) { const unsigned long long t = chunk_tot[c]; chunk_tot[c] = run; run += t; }
    chunk_tot[nchunks] = run;
}

__global__ void k_scan64_finish(unsigned long long* __restrict__ a, long long n, const unsigned long long* __restrict__ chunk_base, int nchunks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = a[i] + chunk_base[i / 4096] + 1ULL;
    if (i == n) a[n] = chunk_base[nchunks] + 1ULL;
}

// ---- first-seen face renumbering (mesh preparation) ---------------------------------------------------------
// The kernels gather the records of the elements next to consecutive faces: they are fast when face ids follow
// the element order, as in the reference's own meshes (meshes/ReadMeshesDat.jl:36-88 numbers a face when the
// element loop first meets it).  For a mesh numbered otherwise this computes that numbering from e2f alone:
// key(f) = 4*e + i of the first (element, local face) that references f; new id = rank of the key.
__global__ void k_fs_mark(const int2* __restrict__ f2e, int nf, unsigned long long* __restrict__ flag) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nf && f2e[f].x >= 0) flag[f2e[f].x] = 1ULL;        // keys are unique: one writer per entry
}

// rank1[code] = 1-based rank after the scan; faces without elements keep the ids after all the others
__global__ void k_fs_perm(const int2* __restrict__ f2e, int nf, const unsigned long long* __restrict__ rank1, long long nused,
                          const unsigned long long* __restrict__ orphan_rank1, long long* __restrict__ new_of_old) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int code = f2e[f].x;
    new_of_old[f] = code >= 0 ? (long long)rank1[code] : nused + (long long)orphan_rank1[f];
}

__global__ void k_fs_orphan(const int2* __restrict__ f2e, int nf, unsigned long long* __restrict__ flag) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nf) flag[f] = f2e[f].x < 0 ? 1ULL : 0ULL;
}

__global__ void k_fs_apply(const int* __restrict__ e2f, long long n3, const long long* __restrict__ new_of_old, long long* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n3) out[k] = new_of_old[e2f[k]];
}

// ---- numbering locality (diagnostic) ---------------------------------------------------------------------------
// Per-block sums over the faces that have an element of |e_hi - e_lo| / nel (how far apart in memory the two
// elements of a face are), |f/nf - e_lo/nel| (how far the face ids stray from the element order) and the count.
__global__ void __launch_bounds__(kBlock)
k_locality(const int2* __restrict__ f2e, int nf, int nel, double* __restrict__ partial, int nblk) {
    __shared__ double sd[3 * (kBlock / 32)];
    const int f = blockIdx.x * kBlock + threadIdx.x;
    double t[3] = {0.0, 0.0, 0.0};
    if (f < nf) {
        const int2 pr = f2e[f];
        if (pr.x >= 0) {
            const int e0 = pr.x >> 2, e1 = pr.y >> 2;
            t[0] = (double)(e1 > e0 ? e1 - e0 : e0 - e1) / (double)nel;
            const double d = (double)f / (double)nf - (double)e0 / (double)nel;
            t[1] = d < 0.0 ? -d : d;
            t[2] = 1.0;
        }
    }
    block_sums<3>(t, sd);
    if (threadIdx.x == 0)
        for (int q = 0; q < 3; q++) partial[(size_t)q * nblk + blockIdx.x] = t[q];
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
        if (mode == 0) { if (rp) rp[e] = v; sq = (M.ebc[e] & kOwnedElem) ? v * v : 0.0; }   // ghosts are counted by their owner
        else { const double coef = (gamma * M.ke[e]) / M.Om[e]; p[e] = p[e] + coef * v; }
    }
    if (mode == 0) {
        const double t = block_sum(sq, sd);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}

// Pe .-= mean(Pe) (SolversFCFV_Stokes.jl:20): per-block sums over owned elements, then p[e] -= sum / n
__global__ void __launch_bounds__(kBlock)
k_sum_owned(int nel, const double* __restrict__ p, const int* __restrict__ ebc, double* __restrict__ partial) {
    __shared__ double sd[kBlock / 32];
    const int e = blockIdx.x * kBlock + threadIdx.x;
    const double v = (e < nel && (ebc[e] & kOwnedElem)) ? p[e] : 0.0;
    const double t = block_sum(v, sd);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void k_sub_mean(int nel, double* __restrict__ p, const double* __restrict__ sum, double n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nel) p[e] = p[e] - sum[0] / n;
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
