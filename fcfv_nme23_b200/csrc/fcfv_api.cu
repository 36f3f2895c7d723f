// fcfv_api.cu -- C ABI of libfcfv_b200.so (see include/fcfv_b200.h for the contract).
//
// Host side of the library: owns the GPU buffers, moves the caller's arrays, launches the kernels
// of fcfv_kernels.cuh on one stream.  There is no CPU implementation of any step in here: if
// CUDA is unavailable every call fails with FCFV_ERR_CUDA.
#include "../../include/fcfv_b200.h"
#include "fcfv_kernels.cuh"
#include "fcfv_generate.cuh"
#include "fcfv_prep.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <initializer_list>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace fcfv;

struct Id128 { char b[128]; };   // ncclUniqueId

enum { T_LOCAL = 0, T_COEF, T_COUNT, T_SCAN, T_FILL, T_RES_ELEM, T_RES_FACE, T_REDUCE, T_PH, T_OTHER, kNumTimers };
static const char* const kTimerNames[kNumTimers] = {"k_local", "k_elem_coef", "k_asm_count", "k_scan", "k_asm_fill",
                                                    "k_res_elem", "k_res_late", "k_reduce_partials", "k_ph", "other"};
struct TimerRec { int id; cudaEvent_t a, b; };

namespace {

// The host array a device buffer currently mirrors (upload cache, fcfv_upload_cache): address, size, a sampled
// fingerprint of the contents and the epoch the record was made in.
struct HostMirror {
    const void* ptr = nullptr;
    size_t bytes = 0;
    uint64_t fp = 0;
    long long epoch = -1;
};

struct DBuf {   // device buffer that only grows
    void* p = nullptr;
    size_t cap = 0;
    HostMirror hm;
};

struct Csr {
    DBuf rowptr, col, val;
    long long nrows = 0, ncols = 0, nnz = -1;   // nnz < 0: not assembled
};

// ---- NCCL through dlopen: only the four entry points the path needs -------------------------
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

}  // namespace

// Copies between PAGEABLE host memory and the device, for callers that hand over plain Julia / NumPy arrays: the
// driver's own path stages them through one buffer with one thread (about 12 GB/s up, 3-4 GB/s down into fresh
// pages).  Here up to kStageThreads workers (not more than the host has hardware threads) each take a contiguous slice, copy it through their own pair of page-locked
// buffers and their own stream, so that the host-side memcpy (and the page faults of a fresh destination) run in
// parallel and overlap the DMA (8 workers x 2 MiB chunks: 33 GB/s up instead of 12, measured; smaller chunks lose
// to per-copy overhead; 16 workers on a 16-core host: another 6 % on a whole pass, 4 MiB chunks lose 8 %).  Page-locked or
// registered memory does not come through here.
#ifndef FCFV_STAGE_THREADS
#define FCFV_STAGE_THREADS 16
#endif
#ifndef FCFV_STAGE_CHUNK_KB
#define FCFV_STAGE_CHUNK_KB 2048
#endif
constexpr int kStageThreads = FCFV_STAGE_THREADS;
constexpr size_t kStageChunk = (size_t)FCFV_STAGE_CHUNK_KB << 10;
struct Stager {
    bool ready = false;
    cudaStream_t s[kStageThreads] = {};
    void* pin[kStageThreads][2] = {};
    cudaEvent_t ev[kStageThreads][2] = {};
    cudaEvent_t done[kStageThreads] = {};
    cudaEvent_t begin = nullptr;
};

struct fcfv_multi;

struct fcfv_handle {
    fcfv_multi* multi = nullptr;    // non-null: a multi-GPU handle (fcfv_create_multi, fcfv_multi.inl); the device state below is unused
    int device = 0;
    Stager stager;
    cudaStream_t stream_out = nullptr;   // device-to-host copies of fcfv_pass, concurrent with its uploads
    cudaEvent_t ev_out = nullptr;
    long long staged_copies = 0;      // copies that went through the Stager (diagnostics / tests)
    int stage_share = 1;              // parts of a multi-GPU handle copy concurrently: each takes its share of the host's threads
    size_t staged_min_bytes = 8u << 20;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    bool async = false;
    long long launches = 0;
    size_t bytes = 0;

    long long nel = 0, nf = 0;
    long long nel_global = 0, nf_global = 0;
    bool have_mesh = false, have_local = false, have_face_data = false, have_sources = false, have_solution = false;
    bool rp64 = false;
    bool has_orphans = false;         // some face is referenced by no element
    bool have_elem_owned = false, have_face_owned = false, have_tau_ref = false;

    // mesh
    DBuf e2f, bc, f2e, fconn, fbc, ebc, eadj, Om, G, nx, ny, ke, dl, taue, tau, elem_owned, face_owned, tau_ref;
    // data
    DBuf sex, sey, VxDir, VyDir, sxx, syy, sxy, syx, Vxh, Vyh, Pe;
    // local coefficients
    DBuf alpha, beta, Z;
    // per-element assembly coefficients (k_elem_coef): ia, c11 and, when some delta != 1, c12 c21 c22
    DBuf coef, coefx, coefs;
    bool delta_dirty = true, aniso = false;
    int coef_variant = -1;       // loop whose 1/alpha, c11 in `coef` match the resident alpha / ke (written by k_local in the pipeline), -1: none
    // assembly
    Csr Kuu, Kup, Muu, Kuusc;
    int last_variant = -1, last_form = -1;   // parameters of the last assembly (fcfv_schur reuses them)
    double last_A = 1.0;
    DBuf fu, fp, sums, bases, chunks, totals;
    long long totals_host[10] = {0};
    bool totals_valid = false;
    // residual
    DBuf Fg, partial, sumsq, red_scratch, red_tickets;
    // scratch
    DBuf tmp, tmp2, err_flag, phw;
    // per-kernel device timers (fcfv_set_timing): event pairs resolved lazily
    bool timing = false;
    std::vector<TimerRec> timer_recs;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
    double timer_ms[kNumTimers] = {0};
    long long timer_calls[kNumTimers] = {0};
    // CUDA graph of one pass (fcfv_run_pipeline): replayed while nothing it captured has changed
    cudaGraphExec_t graph_exec = nullptr;
    long long graph_launches = 0;
    bool graph_disabled = false;          // capture failed with a communicator attached: run directly
    long long graph_key[7] = {-1, -1, -1, -1, -1, -1, -1};   // variant, formulation, tau_mode + flags, alloc epoch, nel, nf, bits of A
    long long alloc_epoch = 0;                           // bumped whenever a device buffer is (re)allocated
    // upload cache (fcfv_upload_cache / fcfv_new_epoch) and transfer counters (fcfv_transfer_stats)
    bool cache_on = false;
    long long epoch = 0;
    long long h2d_bytes = 0, d2h_bytes = 0, cache_hit_bytes = 0, cache_hits = 0;
    // boundary-face lists (fcfv_sparse_face_data): VxDir / VyDir are only ever read on Dirichlet faces and the four Neumann
    // arrays on Neumann faces (Bool factors in the reference: Discretisation:17,31,167-178, Residuals:59-66), so a host array
    // of nf values crosses PCIe as the few values on those faces
    int sparse_mode = 1;              // 0 dense uploads, 1 lists when they are short (auto), 2 always lists
    bool sparse_faces = false;        // the lists below describe the resident mesh
    DBuf dir_idx, neu_idx, face_counts, fpack;
    std::vector<int> dir_list, neu_list;
    double* fstage = nullptr;         // page-locked, 6 slots of fstage_cap values
    size_t fstage_cap = 0;
    cudaEvent_t fstage_ev[6] = {};
    bool fstage_busy[6] = {};
    long long sparse_uploads = 0;
    long long pipelined_calls = 0;    // fcfv_compute_fcfv calls that ran as a chunk pipeline (diagnostics / tests)
    // nccl
    NcclApi nccl;
    void* comm = nullptr;
    int nranks = 1, rank = 0;
};

namespace {

// multi-GPU layer (fcfv_multi.inl): the same entry points on the global arrays, scattered over the parts
int multi_sync(fcfv_t* h);
void multi_destroy(fcfv_t* h);
int multi_set_mesh(fcfv_t*, int64_t, int64_t, const int64_t*, const int64_t*, const double*, const double*, const double*, const double*,
                   const double*, const double*, const double*, int64_t);
int multi_update_fields(fcfv_t*, const double*, const double*, const double*, int64_t, const double*);
int multi_set_sources(fcfv_t*, const double*, const double*);
int multi_set_face_data(fcfv_t*, const double*, const double*, const double*, const double*, const double*, const double*);
int set_face_data_listed(fcfv_t* h, const double* const src[6], const int64_t* faces);
int multi_set_solution(fcfv_t*, const double*, const double*, const double*);
int multi_set_local(fcfv_t*, const double*, const double*, const double*);
int multi_run_local(fcfv_t*, int, double);
int multi_get_local(fcfv_t*, double*, double*, double*, double*);
int multi_run_assemble(fcfv_t*, int, int, double, int, double);
int multi_fetch_totals(fcfv_t*);
int multi_export(fcfv_t*, int, int, void*, void*, double*);
int multi_export_rhs(fcfv_t*, double*, double*);
int multi_run_residuals(fcfv_t*, int, int, double*, double*);
int multi_unsupported(fcfv_t*, const char*);
int multi_element_values(fcfv_t*, const double*, const double*, const double*, const double*, const double*, double*, double*, double*, double*,
                         double*);
int multi_ph_residual(fcfv_t*, const double*, const double*, double*, double*, double*);
int multi_ph_update_p(fcfv_t*, double, const double*, double*);
int multi_ph_fusc(fcfv_t*, double, const double*, double*);
int multi_center_pressure(fcfv_t*, double*, double*);
int multi_has_orphans(fcfv_t*);
double multi_asm_seconds(fcfv_t*);
fcfv_t* multi_part0(fcfv_t*);
template <typename F> void multi_each(fcfv_t* h, F&& fn);

int fail(fcfv_t* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, FCFV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define TRY(expr)            \
    do {                     \
        int rc_ = (expr);    \
        if (rc_) return rc_; \
    } while (0)

#define LAUNCH(h, kernel, grid, block, ...)                            \
    do {                                                               \
        kernel<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);      \
        (h)->launches++;                                               \
    } while (0)

cudaEvent_t pool_event(fcfv_t* h) {
    if (h->events_used == h->event_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->event_pool.push_back(e);
    }
    return h->event_pool[h->events_used++];
}

// Folds finished event pairs into the per-kernel totals (stream must be synchronised).
void resolve_timers(fcfv_t* h) {
    for (const TimerRec& r : h->timer_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { h->timer_ms[r.id] += ms; h->timer_calls[r.id]++; }
    }
    h->timer_recs.clear();
    h->events_used = 0;
}

// LAUNCH with the kernel bracketed by two events on the launching stream when timing is on.
#define TLAUNCH(h, tid, kernel, grid, block, ...)                                        \
    do {                                                                                 \
        if ((h)->timing) {                                                               \
            TimerRec r_{tid, pool_event(h), pool_event(h)};                              \
            cudaEventRecord(r_.a, (h)->stream);                                          \
            kernel<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);                    \
            cudaEventRecord(r_.b, (h)->stream);                                          \
            (h)->timer_recs.push_back(r_);                                               \
        } else {                                                                         \
            kernel<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);                    \
        }                                                                                \
        (h)->launches++;                                                                 \
    } while (0)

int ensure(fcfv_t* h, DBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 8;
    if (b.cap >= bytes) return FCFV_OK;
    if (b.p) { CU(cudaFree(b.p)); h->bytes -= b.cap; b.p = nullptr; b.cap = 0; }
    b.hm.epoch = -1;        // fresh storage mirrors nothing
    // only buffers the handle owns can be referenced by the captured pipeline graph; the temporaries of the
    // export / geometry / renumbering calls live on the caller's stack and must not invalidate it
    const char* hb = reinterpret_cast<const char*>(h);
    const char* bb = reinterpret_cast<const char*>(&b);
    if (bb >= hb && bb < hb + sizeof(*h)) h->alloc_epoch++;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(h, FCFV_ERR_ALLOC, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = bytes;
    h->bytes += bytes;
    return FCFV_OK;
}

void release(fcfv_t* h, DBuf& b) {
    if (b.p) { cudaFree(b.p); h->bytes -= b.cap; }
    b.p = nullptr; b.cap = 0;
}

int stager_init(fcfv_t* h) {
    Stager& S = h->stager;
    if (S.ready) return FCFV_OK;
    CU(cudaEventCreateWithFlags(&S.begin, cudaEventDisableTiming));
    for (int t = 0; t < kStageThreads; t++) {
        CU(cudaStreamCreateWithFlags(&S.s[t], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&S.done[t], cudaEventDisableTiming));
        for (int k = 0; k < 2; k++) {
            CU(cudaHostAlloc(&S.pin[t][k], kStageChunk, cudaHostAllocDefault));
            CU(cudaEventCreateWithFlags(&S.ev[t][k], cudaEventDisableTiming));
        }
    }
    S.ready = true;
    return FCFV_OK;
}

void stager_destroy(fcfv_t* h) {
    Stager& S = h->stager;
    if (S.begin) cudaEventDestroy(S.begin);
    for (int t = 0; t < kStageThreads; t++) {
        if (S.s[t]) { cudaStreamSynchronize(S.s[t]); cudaStreamDestroy(S.s[t]); }
        if (S.done[t]) cudaEventDestroy(S.done[t]);
        for (int k = 0; k < 2; k++) {
            if (S.pin[t][k]) cudaFreeHost(S.pin[t][k]);
            if (S.ev[t][k]) cudaEventDestroy(S.ev[t][k]);
        }
    }
    S = Stager();
}

// one worker: slice [lo, hi) of the transfer through its two page-locked buffers
cudaError_t stage_slice(int device, Stager& S, int t, char* dev, char* host, size_t lo, size_t hi, bool to_device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    const size_t n = (hi - lo + kStageChunk - 1) / kStageChunk;
    auto len = [&](size_t i) { return std::min(kStageChunk, hi - (lo + i * kStageChunk)); };
    if (to_device) {
        for (size_t i = 0; i < n; i++) {
            const int k = (int)(i & 1);
            if (i >= 2 && (e = cudaEventSynchronize(S.ev[t][k])) != cudaSuccess) return e;     // buffer k drained
            memcpy(S.pin[t][k], host + lo + i * kStageChunk, len(i));
            if ((e = cudaMemcpyAsync(dev + lo + i * kStageChunk, S.pin[t][k], len(i), cudaMemcpyHostToDevice, S.s[t])) != cudaSuccess) return e;
            if ((e = cudaEventRecord(S.ev[t][k], S.s[t])) != cudaSuccess) return e;
        }
        // the last two chunks are still being read by the DMA engine: the buffers belong to the next copy only after that
        if ((e = cudaStreamSynchronize(S.s[t])) != cudaSuccess) return e;
    } else {
        auto fetch = [&](size_t i) -> cudaError_t {
            const int k = (int)(i & 1);
            cudaError_t r = cudaMemcpyAsync(S.pin[t][k], dev + lo + i * kStageChunk, len(i), cudaMemcpyDeviceToHost, S.s[t]);
            return r != cudaSuccess ? r : cudaEventRecord(S.ev[t][k], S.s[t]);
        };
        if (n > 0 && (e = fetch(0)) != cudaSuccess) return e;
        for (size_t i = 0; i < n; i++) {
            if (i + 1 < n && (e = fetch(i + 1)) != cudaSuccess) return e;      // the other buffer was consumed in the previous turn
            if ((e = cudaEventSynchronize(S.ev[t][i & 1])) != cudaSuccess) return e;
            memcpy(host + lo + i * kStageChunk, S.pin[t][i & 1], len(i));
        }
    }
    return cudaEventRecord(S.done[t], S.s[t]);
}

int copy_staged(fcfv_t* h, void* dev, void* host, size_t bytes, bool to_device) {
    TRY(stager_init(h));
    Stager& S = h->stager;
    // the workers' streams start after everything already queued on the handle's stream (producers of a D2H source,
    // earlier readers of an H2D destination) ...
    CU(cudaEventRecord(S.begin, h->stream));
    for (int t = 0; t < kStageThreads; t++) CU(cudaStreamWaitEvent(S.s[t], S.begin, 0));
    cudaError_t err[kStageThreads];
    std::thread th[kStageThreads];
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = (int)std::max(1u, std::min((unsigned)kStageThreads, (hw ? hw : 1u) / (unsigned)std::max(1, h->stage_share)));
    const size_t per = ((bytes / nt) + 255) & ~(size_t)255;
    for (int t = 0; t < kStageThreads; t++) {
        const size_t lo = t < nt ? std::min(bytes, per * t) : bytes, hi = t >= nt - 1 ? bytes : std::min(bytes, per * (t + 1));
        err[t] = cudaSuccess;
        auto work = [&, t, lo, hi]() { if (hi > lo) err[t] = stage_slice(h->device, S, t, (char*)dev, (char*)host, lo, hi, to_device);
                                       else err[t] = cudaEventRecord(S.done[t], S.s[t]); };
        try {
            th[t] = std::thread(work);
        } catch (...) {        // no thread to be had (resource limits): this slice runs here; nothing may cross the C boundary
            work();
        }
    }
    for (int t = 0; t < kStageThreads; t++)
        if (th[t].joinable()) th[t].join();
    for (int t = 0; t < kStageThreads; t++)
        if (err[t] != cudaSuccess) return fail(h, FCFV_ERR_CUDA, "staged host copy failed: %s", cudaGetErrorString(err[t]));
    // ... and the handle's stream continues after them
    for (int t = 0; t < kStageThreads; t++) CU(cudaStreamWaitEvent(h->stream, S.done[t], 0));
    h->staged_copies++;
    return FCFV_OK;
}

bool is_pageable_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeUnregistered;
}

bool is_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice;
}

// Copies between any combination of host/device pointers (UVA resolves the direction).  Large transfers from / to
// pageable host memory go through the Stager; like the driver's own pageable path they have completed on the host
// side when the call returns.
int copy(fcfv_t* h, void* dst, const void* src, size_t bytes) {
    if (bytes == 0 || dst == src) return FCFV_OK;       // dst == src: the data is already where it belongs (multi-GPU scatter)
    const bool dd = is_device(dst), sd = is_device(src);
    if (dd && !sd) h->h2d_bytes += (long long)bytes;       // what actually crosses PCIe (fcfv_transfer_stats)
    if (sd && !dd) h->d2h_bytes += (long long)bytes;
    if (bytes >= h->staged_min_bytes) {
        if (dd && is_pageable_host(src)) return copy_staged(h, dst, const_cast<void*>(src), bytes, true);
        if (sd && is_pageable_host(dst)) return copy_staged(h, const_cast<void*>(src), dst, bytes, false);
    }
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream));
    return FCFV_OK;
}

int upload(fcfv_t* h, DBuf& b, const void* src, size_t bytes) {
    TRY(ensure(h, b, bytes));
    b.hm.epoch = -1;
    return copy(h, b.p, src, bytes);
}

// ---- upload cache -------------------------------------------------------------------------------
// A driver of the reference hands the same arrays to ComputeFCFV, ElementAssemblyLoop[NEW] and
// ComputeResidualsFCFV_Stokes_o1 (VxDir / VyDir three times, the Neumann arrays and the sources twice, and alpha /
// beta / Z come back as arguments after they were returned, SolKzFCFV.jl:150-205).  With the cache on, a device
// buffer remembers which host array it mirrors; an upload of the same address and size inside the same epoch is
// skipped when a sampled fingerprint of the host contents still matches.  The caller vouches for "not modified in
// place inside an epoch" (fcfv_new_epoch starts a new one; fcfv_set_mesh does so implicitly); the fingerprint is a
// safety net that catches wholesale changes, not a proof.  Off by default.
uint64_t fingerprint(const void* p, size_t bytes) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    uint64_t hsh = 0x9E3779B97F4A7C15ull ^ (uint64_t)bytes;
    auto mix = [&](uint64_t v) { hsh ^= v + 0x9E3779B97F4A7C15ull + (hsh << 6) + (hsh >> 2); hsh *= 0xFF51AFD7ED558CCDull; hsh ^= hsh >> 33; };
    const size_t nw = bytes / 8;
    auto word = [&](size_t k) { uint64_t v; memcpy(&v, c + 8 * k, 8); return v; };
    if (nw <= 8192) {
        for (size_t k = 0; k < nw; k++) mix(word(k));
    } else {
        for (size_t k = 0; k < 256; k++) { mix(word(k)); mix(word(nw - 1 - k)); }
        const size_t stride = nw / 4096;
        for (size_t k = 0; k < 4096; k++) mix(word(k * stride + (k * 7) % stride));     // one word per 1/4096th, at varying offsets
    }
    for (size_t k = 8 * nw; k < bytes; k++) mix(c[k]);
    return hsh;
}

bool cache_hit(fcfv_t* h, const DBuf& b, const void* src, size_t bytes) {
    return h->cache_on && b.hm.epoch == h->epoch && b.hm.ptr == src && b.hm.bytes == bytes && b.cap >= bytes && !is_device(src) &&
           fingerprint(src, bytes) == b.hm.fp;
}

// host -> device copy of a caller's array into a handle-owned buffer, skipped when the buffer already mirrors it
int upload_cached(fcfv_t* h, DBuf& b, const void* src, size_t bytes) {
    if (bytes == 0) return ensure(h, b, bytes);
    if (cache_hit(h, b, src, bytes)) { h->cache_hits++; h->cache_hit_bytes += (long long)bytes; return FCFV_OK; }
    TRY(ensure(h, b, bytes));
    b.hm.epoch = -1;
    TRY(copy(h, b.p, src, bytes));
    if (h->cache_on && !is_device(src)) b.hm = HostMirror{src, bytes, fingerprint(src, bytes), h->epoch};
    return FCFV_OK;
}

// after a COMPLETED device -> host copy of a whole buffer: the destination is a mirror of it
void mirror_download(fcfv_t* h, DBuf& b, const void* dst, size_t bytes) {
    if (h->cache_on && dst && bytes && !is_device(dst)) b.hm = HostMirror{dst, bytes, fingerprint(dst, bytes), h->epoch};
}

void new_epoch(fcfv_t* h) { h->epoch++; }

int sync(fcfv_t* h) {
    CU(cudaStreamSynchronize(h->stream));
    return FCFV_OK;
}

int finish(fcfv_t* h) { return h->async ? FCFV_OK : sync(h); }

inline int nblocks(long long n, int block = kBlock) { return (int)std::max<long long>(1, (n + block - 1) / block); }

template <typename T> T* P(const DBuf& b) { return (T*)b.p; }

MeshView mesh_view(fcfv_t* h) {
    MeshView M;
    M.nel = (int)h->nel; M.nf = (int)h->nf;
    M.e2f = P<int>(h->e2f); M.bc = P<int8_t>(h->bc); M.f2e = P<int2>(h->f2e);
    M.fconn = P<int4>(h->fconn); M.fbc = P<int>(h->fbc); M.ebc = P<int>(h->ebc); M.eadj = P<int>(h->eadj);
    M.Om = P<double>(h->Om); M.G = P<double>(h->G); M.nx = P<double>(h->nx); M.ny = P<double>(h->ny);
    M.ke = P<double>(h->ke); M.dl = P<double>(h->dl); M.taue = P<double>(h->taue); M.tau = P<double>(h->tau);
    M.tau_ref = h->have_tau_ref ? P<double>(h->tau_ref) : nullptr;
    return M;
}
LocalView local_view(fcfv_t* h) { return LocalView{P<double>(h->alpha), P<double>(h->beta), P<double>(h->Z)}; }
FaceDataView face_view(fcfv_t* h) {
    return FaceDataView{P<double>(h->VxDir), P<double>(h->VyDir), P<double>(h->sxx),
                        P<double>(h->syy), P<double>(h->sxy), P<double>(h->syx)};
}
SolutionView sol_view(fcfv_t* h) { return SolutionView{P<double>(h->Vxh), P<double>(h->Vyh), P<double>(h->Pe)}; }

int check_form(fcfv_t* h, int formulation) {
    if (formulation != FCFV_GRADIENT && formulation != FCFV_SYMMETRIC_GRADIENT)
        return fail(h, FCFV_ERR_INVALID, "formulation must be FCFV_GRADIENT or FCFV_SYMMETRIC_GRADIENT (got %d)", formulation);
    return FCFV_OK;
}

int need(fcfv_t* h, bool ok, const char* what) {
    if (!ok) return fail(h, FCFV_ERR_STATE, "%s", what);
    return FCFV_OK;
}

// Builds the face -> element map and validates the connectivity (device side, integer atomics).
int build_f2e(fcfv_t* h) {
    const long long nel = h->nel, nf = h->nf;
    TRY(ensure(h, h->f2e, sizeof(int2) * nf));
    TRY(ensure(h, h->tmp2, sizeof(int) * nf));
    LAUNCH(h, k_f2e_init, nblocks(nf, 256), 256, P<int2>(h->f2e), P<int>(h->tmp2), (int)nf);
    LAUNCH(h, k_f2e_build, nblocks(nel, 256), 256, P<int>(h->e2f), P<int2>(h->f2e), P<int>(h->tmp2), (int)nel);
    LAUNCH(h, k_f2e_check, nblocks(nf, 256), 256, P<int2>(h->f2e), P<int>(h->tmp2), P<int>(h->e2f), (int)nel, (int)nf,
           P<int>(h->err_flag));
    return FCFV_OK;
}

// Face connectivity records of the assembly; call after f2e and bc are final.
int build_fconn(fcfv_t* h) {
    const long long nel = h->nel, nf = h->nf;
    TRY(ensure(h, h->fconn, sizeof(int4) * nf));
    TRY(ensure(h, h->fbc, sizeof(int) * nf));
    LAUNCH(h, k_fconn, nblocks(nf, 256), 256, P<int>(h->e2f), P<int8_t>(h->bc), P<int2>(h->f2e), (int)nel, (int)nf, P<int4>(h->fconn),
           P<int>(h->fbc));
    TRY(ensure(h, h->ebc, sizeof(int) * nel));
    TRY(ensure(h, h->eadj, sizeof(int) * nel));
    LAUNCH(h, k_ebc, nblocks(nel, 256), 256, P<int>(h->e2f), P<int8_t>(h->bc), P<int2>(h->f2e), (int)nel, P<int>(h->ebc), P<int>(h->eadj));
    return FCFV_OK;
}

// Lists of the Dirichlet / Neumann faces of the resident mesh (device bc).  They are short on every mesh of the reference
// (boundary faces: O(sqrt(nf))); when one of them is not (more than nf/16 or 2^18 faces in auto mode) the per-face data
// keeps going up whole.
int build_face_lists(fcfv_t* h) {
    h->sparse_faces = false;
    h->dir_list.clear(); h->neu_list.clear();
    const long long nf = h->nf;
    if (h->sparse_mode == 0 || nf == 0) return FCFV_OK;
    const long long cap = h->sparse_mode == 2 ? nf : std::min<long long>(nf / 16, 1LL << 18);
    if (cap < 1) return FCFV_OK;
    TRY(ensure(h, h->dir_idx, sizeof(int) * cap));
    TRY(ensure(h, h->neu_idx, sizeof(int) * cap));
    TRY(ensure(h, h->face_counts, sizeof(int) * 2));
    CU(cudaMemsetAsync(h->face_counts.p, 0, sizeof(int) * 2, h->stream));
    LAUNCH(h, k_bc_face_lists, nblocks(nf, 256), 256, (int)nf, P<int8_t>(h->bc), (int)cap, P<int>(h->dir_idx), P<int>(h->neu_idx),
           P<int>(h->face_counts));
    int cnt[2] = {0, 0};
    CU(cudaMemcpyAsync(cnt, h->face_counts.p, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
    TRY(sync(h));
    if (cnt[0] > cap || cnt[1] > cap) return FCFV_OK;
    struct L { std::vector<int>* v; DBuf* d; int n; } lists[2] = {{&h->dir_list, &h->dir_idx, cnt[0]}, {&h->neu_list, &h->neu_idx, cnt[1]}};
    for (L& l : lists) {
        l.v->resize((size_t)l.n);
        if (l.n == 0) continue;
        CU(cudaMemcpyAsync(l.v->data(), l.d->p, sizeof(int) * (size_t)l.n, cudaMemcpyDeviceToHost, h->stream));
        TRY(sync(h));
        std::sort(l.v->begin(), l.v->end());        // ascending ids: a deterministic list and a forward walk over the host arrays
        CU(cudaMemcpyAsync(l.d->p, l.v->data(), sizeof(int) * (size_t)l.n, cudaMemcpyHostToDevice, h->stream));
        TRY(sync(h));
    }
    const size_t need = (size_t)std::max(1, std::max(cnt[0], cnt[1]));
    if (need > h->fstage_cap) {
        for (int k = 0; k < 6; k++)
            if (h->fstage_busy[k]) { CU(cudaEventSynchronize(h->fstage_ev[k])); h->fstage_busy[k] = false; }
        if (h->fstage) { CU(cudaFreeHost(h->fstage)); h->fstage = nullptr; h->fstage_cap = 0; }
        const size_t cap2 = need + need / 2 + 64;
        if (cudaHostAlloc((void**)&h->fstage, sizeof(double) * 6 * cap2, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();         // no page-locked memory to be had: the boundary arrays keep going up whole
            h->fstage = nullptr;
            return FCFV_OK;
        }
        h->fstage_cap = cap2;
    }
    TRY(ensure(h, h->fpack, sizeof(double) * 6 * h->fstage_cap));
    for (int k = 0; k < 6; k++)
        if (!h->fstage_ev[k]) CU(cudaEventCreateWithFlags(&h->fstage_ev[k], cudaEventDisableTiming));
    h->sparse_faces = true;
    return FCFV_OK;
}

// Array `which` (0 VxDir, 1 VyDir, 2..5 sxx, syy, sxy, syx) of the per-face data from a HOST array through the list of the
// faces it is read on: gather into a page-locked slot, one small copy, scatter on the device.  The rest of the device
// array keeps what it held (zeros after allocation); no kernel reads it.
// map: local -> global face id when src is the GLOBAL array of a multi-GPU handle (this handle is one of its parts).
int upload_listed(fcfv_t* h, DBuf& b, const double* src, int which, const int64_t* map = nullptr) {
    const std::vector<int>& list = which < 2 ? h->dir_list : h->neu_list;
    const DBuf& idx = which < 2 ? h->dir_idx : h->neu_idx;
    b.hm.epoch = -1;
    const size_t n = list.size();
    h->sparse_uploads++;
    if (n == 0) return FCFV_OK;
    double* slot = h->fstage + (size_t)which * h->fstage_cap;
    if (h->fstage_busy[which]) { CU(cudaEventSynchronize(h->fstage_ev[which])); h->fstage_busy[which] = false; }
    if (map) for (size_t k = 0; k < n; k++) slot[k] = src[map[list[k]]];
    else     for (size_t k = 0; k < n; k++) slot[k] = src[list[k]];
    double* pack = P<double>(h->fpack) + (size_t)which * h->fstage_cap;
    CU(cudaMemcpyAsync(pack, slot, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    h->h2d_bytes += (long long)(sizeof(double) * n);
    CU(cudaEventRecord(h->fstage_ev[which], h->stream));
    h->fstage_busy[which] = true;
    LAUNCH(h, k_scatter_faces, nblocks((long long)n, 256), 256, (int)n, P<int>(idx), pack, P<double>(b));
    return FCFV_OK;
}

int alloc_mesh_dependent(fcfv_t* h) {
    const long long nel = h->nel, nf = h->nf;
    TRY(ensure(h, h->tau, sizeof(double) * nf));
    TRY(ensure(h, h->alpha, sizeof(double) * nel));
    TRY(ensure(h, h->beta, sizeof(double) * 2 * nel));
    TRY(ensure(h, h->Z, sizeof(double) * 4 * nel));
    // 64-bit row pointers once the CSR capacity can exceed 2^31 entries (FCFV_FORCE_RP64=1 forces them: test hook)
    const char* force64 = getenv("FCFV_FORCE_RP64");
    h->rp64 = 20 * nf >= (1LL << 31) || (force64 && force64[0] == '1');
    return FCFV_OK;
}

int read_err_flag(fcfv_t* h, int* flag) {
    CU(cudaMemcpyAsync(flag, h->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    return sync(h);
}

}  // namespace

// =================================================================================================
// All entry points are declared extern "C" in include/fcfv_b200.h; the definitions below inherit
// that linkage.

int fcfv_version(void) { return 100; }

int fcfv_create(fcfv_t** out, int device) {
    if (!out) return FCFV_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return FCFV_ERR_CUDA;   // no CPU fallback
    if (device < 0 || device >= count) return FCFV_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return FCFV_ERR_CUDA;
    fcfv_t* h = new (std::nothrow) fcfv_handle();
    if (!h) return FCFV_ERR_ALLOC;
    h->device = device;
    if (const char* v = getenv("FCFV_STAGED_MIN_BYTES")) h->staged_min_bytes = (size_t)strtoull(v, nullptr, 10);   // 0 = never
    if (h->staged_min_bytes == 0) h->staged_min_bytes = ~(size_t)0;
    if (const char* v = getenv("FCFV_SPARSE_FACE_DATA"))                      // initial fcfv_sparse_face_data mode (test hook)
        if (v[0] >= '0' && v[0] <= '2' && v[1] == 0) h->sparse_mode = v[0] - '0';
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
        delete h;
        return FCFV_ERR_CUDA;
    }
    if (ensure(h, h->err_flag, sizeof(int)) || ensure(h, h->totals, sizeof(long long) * 16) ||
        ensure(h, h->sumsq, sizeof(double) * 8) || ensure(h, h->red_scratch, sizeof(double) * 6 * kRedSlices) ||
        ensure(h, h->red_tickets, sizeof(unsigned int) * 8)) {
        delete h;
        return FCFV_ERR_ALLOC;
    }
    cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream);
    cudaMemsetAsync(h->red_tickets.p, 0, sizeof(unsigned int) * 8, h->stream);
    *out = h;
    return FCFV_OK;
}

int fcfv_destroy(fcfv_t* h) {
    if (!h) return FCFV_OK;
    if (h->multi) { multi_destroy(h); delete h; return FCFV_OK; }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    // the pipeline graph may hold captured NCCL work: a communicator is only released once every graph that captured
    // its operations is gone (ncclCommDestroy waits for that), so the graph goes first
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; cudaStreamSynchronize(h->stream); }
    if (h->comm && h->nccl.CommDestroy) h->nccl.CommDestroy(h->comm);
    DBuf* all[] = {&h->e2f, &h->bc, &h->f2e, &h->fconn, &h->fbc, &h->ebc, &h->eadj, &h->Om, &h->G, &h->nx, &h->ny, &h->ke, &h->dl, &h->taue, &h->tau,
                   &h->elem_owned, &h->face_owned, &h->tau_ref, &h->sex, &h->sey, &h->VxDir, &h->VyDir, &h->sxx,
                   &h->syy, &h->sxy, &h->syx, &h->Vxh, &h->Vyh, &h->Pe, &h->alpha, &h->beta, &h->Z,
                   &h->Kuu.rowptr, &h->Kuu.col, &h->Kuu.val, &h->Kup.rowptr, &h->Kup.col, &h->Kup.val,
                   &h->Muu.rowptr, &h->Muu.col, &h->Muu.val, &h->Kuusc.rowptr, &h->Kuusc.col, &h->Kuusc.val, &h->fu, &h->fp, &h->sums, &h->bases, &h->chunks, &h->totals,
                   &h->Fg, &h->partial, &h->sumsq, &h->red_scratch, &h->red_tickets, &h->tmp, &h->tmp2, &h->err_flag, &h->phw, &h->coef, &h->coefx, &h->coefs,
                   &h->dir_idx, &h->neu_idx, &h->face_counts, &h->fpack};
    for (DBuf* b : all) release(h, *b);
    if (h->fstage) cudaFreeHost(h->fstage);
    for (cudaEvent_t e : h->fstage_ev)
        if (e) cudaEventDestroy(e);
    stager_destroy(h);
    if (h->stream_out) { cudaStreamSynchronize(h->stream_out); cudaStreamDestroy(h->stream_out); }
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
    return FCFV_OK;
}

const char* fcfv_last_error(fcfv_t* h) { return h ? h->err.c_str() : "null handle"; }

int fcfv_set_async(fcfv_t* h, int async) {
    if (!h) return FCFV_ERR_INVALID;
    h->async = async != 0;
    return FCFV_OK;
}

int fcfv_sync(fcfv_t* h) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_sync(h);
    return sync(h);
}

int fcfv_upload_cache(fcfv_t* h, int on) {
    if (!h) return FCFV_ERR_INVALID;
    h->cache_on = on != 0;
    new_epoch(h);
    return FCFV_OK;
}

int fcfv_sparse_face_data(fcfv_t* h, int mode) {
    if (!h) return FCFV_ERR_INVALID;
    if (mode < 0 || mode > 2) return fail(h, FCFV_ERR_INVALID, "fcfv_sparse_face_data: mode must be 0 (off), 1 (auto) or 2 (always)");
    h->sparse_mode = mode;
    if (h->multi) {      // every part keeps its own lists (of its local faces)
        int rc = FCFV_OK;
        multi_each(h, [&](fcfv_t* q) {
            q->sparse_mode = mode;
            if (rc == FCFV_OK && q->have_mesh && cudaSetDevice(q->device) == cudaSuccess) rc = build_face_lists(q);
            if (rc != FCFV_OK) h->err = q->err;
        });
        return rc;
    }
    if (h->have_mesh) { CU(cudaSetDevice(h->device)); TRY(build_face_lists(h)); }
    return FCFV_OK;
}

int fcfv_sparse_face_info(fcfv_t* h, int* active, int64_t* n_dirichlet, int64_t* n_neumann, int64_t* uploads) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) {      // lists in use when every part has them; lengths summed over the parts (ghost faces count once per part)
        bool all = true;
        int64_t nd = 0, nn = 0, up = 0;
        multi_each(h, [&](fcfv_t* q) { all = all && q->sparse_faces; nd += (int64_t)q->dir_list.size(); nn += (int64_t)q->neu_list.size(); up += q->sparse_uploads; });
        if (active) *active = all ? 1 : 0;
        if (n_dirichlet) *n_dirichlet = all ? nd : -1;
        if (n_neumann) *n_neumann = all ? nn : -1;
        if (uploads) *uploads = up;
        return FCFV_OK;
    }
    if (active) *active = h->sparse_faces ? 1 : 0;
    if (n_dirichlet) *n_dirichlet = h->sparse_faces ? (int64_t)h->dir_list.size() : -1;
    if (n_neumann) *n_neumann = h->sparse_faces ? (int64_t)h->neu_list.size() : -1;
    if (uploads) *uploads = h->sparse_uploads;
    return FCFV_OK;
}

int fcfv_new_epoch(fcfv_t* h) {
    if (!h) return FCFV_ERR_INVALID;
    new_epoch(h);
    return FCFV_OK;
}

int fcfv_transfer_stats(fcfv_t* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* cache_hit_bytes, int reset) {
    if (!h) return FCFV_ERR_INVALID;
    long long up = h->h2d_bytes, down = h->d2h_bytes, hit = h->cache_hit_bytes;
    if (h->multi) multi_each(h, [&](fcfv_t* q) { up += q->h2d_bytes; down += q->d2h_bytes; hit += q->cache_hit_bytes; });
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    if (cache_hit_bytes) *cache_hit_bytes = hit;
    if (reset) {
        h->h2d_bytes = h->d2h_bytes = h->cache_hit_bytes = 0;
        if (h->multi) multi_each(h, [&](fcfv_t* q) { q->h2d_bytes = q->d2h_bytes = q->cache_hit_bytes = 0; });
    }
    return FCFV_OK;
}

int fcfv_host_register(fcfv_t* h, void* ptr, int64_t bytes) {
    if (!h) return FCFV_ERR_INVALID;
    if (!ptr || bytes <= 0) return fail(h, FCFV_ERR_INVALID, "fcfv_host_register: empty range");
    CU(cudaSetDevice(h->device));
    const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();      // not sticky: do not let a later launch check pick it up
        return fail(h, FCFV_ERR_CUDA, "fcfv_host_register: %s", cudaGetErrorString(e));
    }
    return FCFV_OK;
}

int fcfv_host_unregister(fcfv_t* h, void* ptr) {
    if (!h) return FCFV_ERR_INVALID;
    if (!ptr) return fail(h, FCFV_ERR_INVALID, "fcfv_host_unregister: null pointer");
    CU(cudaSetDevice(h->device));
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(h, FCFV_ERR_CUDA, "fcfv_host_unregister: %s", cudaGetErrorString(e));
    }
    return FCFV_OK;
}

int fcfv_set_timing(fcfv_t* h, int on) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) { int rc = FCFV_OK; multi_each(h, [&](fcfv_t* q) { if (!rc) rc = fcfv_set_timing(q, on); }); return rc; }
    TRY(sync(h));
    resolve_timers(h);
    h->timing = on != 0;
    for (int k = 0; k < kNumTimers; k++) { h->timer_ms[k] = 0.0; h->timer_calls[k] = 0; }
    return FCFV_OK;
}

int fcfv_num_timers(void) { return kNumTimers; }
const char* fcfv_timer_name(int k) { return (k >= 0 && k < kNumTimers) ? kTimerNames[k] : ""; }

int fcfv_get_timers(fcfv_t* h, double* ms, int64_t* calls) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) {      // per kernel: the slowest device's time, the first device's launch count
        int rc = FCFV_OK;
        bool first = true;
        multi_each(h, [&](fcfv_t* q) {
            double m[kNumTimers]; int64_t c[kNumTimers];
            if (rc || (rc = fcfv_get_timers(q, m, c))) return;
            for (int k = 0; k < kNumTimers; k++) {
                if (ms) ms[k] = first ? m[k] : std::max(ms[k], m[k]);
                if (calls && first) calls[k] = c[k];
            }
            first = false;
        });
        return rc;
    }
    TRY(sync(h));
    resolve_timers(h);
    for (int k = 0; k < kNumTimers; k++) {
        if (ms) ms[k] = h->timer_ms[k];
        if (calls) calls[k] = h->timer_calls[k];
    }
    return FCFV_OK;
}

void* fcfv_stream(fcfv_t* h) { return !h ? nullptr : (void*)(h->multi ? multi_part0(h)->stream : h->stream); }
int64_t fcfv_kernel_launches(fcfv_t* h) {
    if (!h) return 0;
    int64_t n = h->launches;
    if (h->multi) multi_each(h, [&](fcfv_t* q) { n += q->launches; });
    return n;
}
int64_t fcfv_staged_copies(fcfv_t* h) {
    if (!h) return 0;
    int64_t n = h->staged_copies;
    if (h->multi) multi_each(h, [&](fcfv_t* q) { n += q->staged_copies; });
    return n;
}
int64_t fcfv_pipelined_calls(fcfv_t* h) { return h ? (int64_t)h->pipelined_calls : 0; }
int64_t fcfv_debug_bounds_violations(void) {
#if FCFV_BOUNDS_CHECK
    // summed over every device this process has used (the counter is a per-device symbol)
    int n = 0, cur = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || cudaGetDevice(&cur) != cudaSuccess) return -2;
    long long total = 0;
    for (int d = 0; d < n; d++) {
        unsigned long long v = 0;
        if (cudaSetDevice(d) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
            cudaMemcpyFromSymbol(&v, g_bounds_violations, sizeof v) != cudaSuccess) { cudaGetLastError(); continue; }
        total += (long long)v;
    }
    cudaSetDevice(cur);
    return total;
#else
    return -1;      // not a bounds-checked build
#endif
}
int64_t fcfv_device_bytes(fcfv_t* h) {
    if (!h) return 0;
    int64_t n = (int64_t)h->bytes;
    if (h->multi) multi_each(h, [&](fcfv_t* q) { n += (int64_t)q->bytes; });
    return n;
}

// -------------------------------------------------------------------------------------------------
namespace {
// First half of fcfv_set_mesh: connectivity and tags (everything the face -> element map, the connectivity records, the
// tag words and the boundary-face lists are built from).  have_mesh stays false until the fields have arrived.
int set_mesh_connectivity(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, const int64_t* bc) {
    CU(cudaSetDevice(h->device));
    if (nel < 0 || nf < 0 || nel >= (1LL << 29) || nf >= (1LL << 30))
        return fail(h, FCFV_ERR_INVALID, "mesh size out of range (nel=%lld < 2^29, nf=%lld < 2^30 required)", (long long)nel, (long long)nf);
    h->have_mesh = false; h->have_local = false;
    new_epoch(h);        // a new mesh: nothing that is resident mirrors a host array any more
    h->coef_variant = -1;
    // data arrays are sized for the mesh they were set on: a new mesh invalidates them (sources, face data and
    // solution have to be set again; their buffers may be too small for the new nel / nf)
    if (nel != h->nel || nf != h->nf) h->have_sources = h->have_face_data = h->have_solution = false;
    h->delta_dirty = true;
    h->Kuu.nnz = h->Kup.nnz = h->Muu.nnz = h->Kuusc.nnz = -1;
    h->nel = nel; h->nf = nf; h->nel_global = nel; h->nf_global = nf;
    h->have_elem_owned = h->have_face_owned = h->have_tau_ref = false;
    // connectivity: int64 1-based -> int32 0-based on the device
    TRY(upload(h, h->tmp, e2f, sizeof(int64_t) * 3 * nel));
    TRY(ensure(h, h->e2f, sizeof(int) * 3 * nel));
    CU(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
    LAUNCH(h, k_convert_e2f, nblocks(3 * nel, 256), 256, P<long long>(h->tmp), P<int>(h->e2f), (long long)(3 * nel), (int)nf,
           P<int>(h->err_flag));
    TRY(upload(h, h->tmp, bc, sizeof(int64_t) * nf));
    TRY(ensure(h, h->bc, (size_t)nf));
    LAUNCH(h, k_convert_bc, nblocks(nf, 256), 256, P<long long>(h->tmp), P<int8_t>(h->bc), (int)nf);
    TRY(alloc_mesh_dependent(h));
    CU(cudaMemsetAsync(h->tau.p, 0, sizeof(double) * nf, h->stream));
    TRY(build_f2e(h));
    int flag = 0;
    TRY(read_err_flag(h, &flag));
    h->has_orphans = (flag & 8) != 0;        // faces no element references: their tau is only ever what the caller uploads
    flag &= 7;
    if (!flag) { TRY(build_fconn(h)); TRY(build_face_lists(h)); TRY(sync(h)); }
    if (flag & 1) return fail(h, FCFV_ERR_INVALID, "e2f contains a face id outside 1..nf");
    if (flag & 2) return fail(h, FCFV_ERR_INVALID, "a face is referenced by more than two (element, local face) pairs");
    if (flag & 4) return fail(h, FCFV_ERR_INVALID, "two elements share more than one face (or an element lists a face twice)");
    return FCFV_OK;
}
}  // namespace

int fcfv_set_mesh(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, const int64_t* bc, const double* Omega,
                  const double* nx, const double* ny, const double* Gamma, const double* ke, const double* delta,
                  const double* taue, int64_t ntaue) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_set_mesh(h, nel, nf, e2f, bc, Omega, nx, ny, Gamma, ke, delta, taue, ntaue);
    if (!e2f || !bc || !Omega || !nx || !ny || !Gamma || !ke || !delta)
        return fail(h, FCFV_ERR_INVALID, "fcfv_set_mesh: null mesh array");
    if (taue && ntaue < nel) return fail(h, FCFV_ERR_INVALID, "taue has %lld entries, need at least nel=%lld", (long long)ntaue, (long long)nel);
    TRY(set_mesh_connectivity(h, nel, nf, e2f, bc));
    TRY(upload(h, h->Om, Omega, sizeof(double) * nel));
    TRY(upload(h, h->nx, nx, sizeof(double) * 3 * nel));
    TRY(upload(h, h->ny, ny, sizeof(double) * 3 * nel));
    TRY(upload(h, h->G, Gamma, sizeof(double) * 3 * nel));
    // the mutable fields are remembered by the upload cache: the field refresh of the first ComputeFCFV on this mesh
    // (fcfv_update_fields in the same epoch) does not send them a second time
    TRY(upload_cached(h, h->ke, ke, sizeof(double) * nel));
    TRY(upload_cached(h, h->dl, delta, sizeof(double) * nel));
    TRY(ensure(h, h->taue, sizeof(double) * nel));
    if (taue) TRY(upload_cached(h, h->taue, taue, sizeof(double) * nel));
    else { h->taue.hm.epoch = -1; CU(cudaMemsetAsync(h->taue.p, 0, sizeof(double) * nel, h->stream)); }
    TRY(sync(h));
    h->have_mesh = true;
    return FCFV_OK;
}

int fcfv_update_fields(fcfv_t* h, const double* ke, const double* delta, const double* taue, int64_t ntaue,
                       const double* tau) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_update_fields: no mesh set"));
    if (h->multi) return multi_update_fields(h, ke, delta, taue, ntaue, tau);
    CU(cudaSetDevice(h->device));
    if (taue && ntaue < h->nel) return fail(h, FCFV_ERR_INVALID, "taue has %lld entries, need at least nel", (long long)ntaue);
    h->coef_variant = -1;
    if (ke) TRY(upload_cached(h, h->ke, ke, sizeof(double) * h->nel));
    if (delta) {
        const long long hits = h->cache_hits;
        TRY(upload_cached(h, h->dl, delta, sizeof(double) * h->nel));
        if (h->cache_hits == hits) h->delta_dirty = true;
    }
    if (taue) TRY(upload_cached(h, h->taue, taue, sizeof(double) * h->nel));
    if (tau) TRY(upload_cached(h, h->tau, tau, sizeof(double) * h->nf));
    return finish(h);
}

int fcfv_set_ownership(fcfv_t* h, const uint8_t* elem_owned, const uint8_t* face_owned, int64_t nel_global,
                       int64_t nf_global, const double* tau_ref) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_set_ownership");
    TRY(need(h, h->have_mesh, "fcfv_set_ownership: no mesh set"));
    CU(cudaSetDevice(h->device));
    h->have_elem_owned = elem_owned != nullptr;
    h->have_face_owned = face_owned != nullptr;
    h->have_tau_ref = tau_ref != nullptr;
    if (elem_owned) TRY(upload(h, h->elem_owned, elem_owned, (size_t)h->nel));
    if (face_owned) TRY(upload(h, h->face_owned, face_owned, (size_t)h->nf));
    if (tau_ref) TRY(upload(h, h->tau_ref, tau_ref, sizeof(double) * h->nel));
    h->nel_global = nel_global > 0 ? nel_global : h->nel;
    h->nf_global = nf_global > 0 ? nf_global : h->nf;
    // the kernels read the masks from the tag words they load anyway
    LAUNCH(h, k_pack_owned, nblocks(h->nel, 256), 256, (int)h->nel, elem_owned ? P<uint8_t>(h->elem_owned) : nullptr, kOwnedElem, P<int>(h->ebc));
    LAUNCH(h, k_pack_owned, nblocks(h->nf, 256), 256, (int)h->nf, face_owned ? P<uint8_t>(h->face_owned) : nullptr, kOwnedFace, P<int>(h->fbc));
    LAUNCH(h, k_pack_face_owned, nblocks(h->nel, 256), 256, (int)h->nel, P<int>(h->e2f), face_owned ? P<uint8_t>(h->face_owned) : nullptr, P<int>(h->eadj));
    CU(cudaGetLastError());
    return finish(h);
}

// -------------------------------------------------------------------------------------------------
int fcfv_geometry(fcfv_t* h, int64_t nel, int64_t nf, int64_t nv, const int64_t* e2v, const int64_t* e2f,
                  const double* xv, const double* yv, const double* xf, const double* yf, int numbering, double tau_r,
                  double* Omega, double* xc, double* yc, double* nx, double* ny, double* Gamma, double* tau) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) {      // stateless utility: runs on the first device
        const int rc = fcfv_geometry(multi_part0(h), nel, nf, nv, e2v, e2f, xv, yv, xf, yf, numbering, tau_r, Omega, xc, yc, nx, ny, Gamma, tau);
        if (rc) h->err = multi_part0(h)->err;
        return rc;
    }
    CU(cudaSetDevice(h->device));
    if (nel < 0 || nf < 0 || nv < 0 || nel >= (1LL << 29) || nf >= (1LL << 30) || nv >= (1LL << 31))
        return fail(h, FCFV_ERR_INVALID, "fcfv_geometry: size out of range");
    if (!e2v || !e2f || !xv || !yv || !xf || !yf) return fail(h, FCFV_ERR_INVALID, "fcfv_geometry: null input");
    if (numbering != 0 && numbering != 1) return fail(h, FCFV_ERR_INVALID, "fcfv_geometry: numbering must be 0 or 1");
    DBuf d_e2v, d_e2f, d_xv, d_yv, d_xf, d_yf, d_out, d_tau;
    int rc = FCFV_OK;
    auto body = [&]() -> int {
        CU(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
        TRY(upload(h, h->tmp, e2v, sizeof(int64_t) * 3 * nel));
        TRY(ensure(h, d_e2v, sizeof(int) * 3 * nel));
        LAUNCH(h, k_convert_e2f, nblocks(3 * nel, 256), 256, P<long long>(h->tmp), P<int>(d_e2v), (long long)(3 * nel), (int)nv, P<int>(h->err_flag));
        TRY(upload(h, h->tmp, e2f, sizeof(int64_t) * 3 * nel));
        TRY(ensure(h, d_e2f, sizeof(int) * 3 * nel));
        LAUNCH(h, k_convert_e2f, nblocks(3 * nel, 256), 256, P<long long>(h->tmp), P<int>(d_e2f), (long long)(3 * nel), (int)nf, P<int>(h->err_flag));
        TRY(upload(h, d_xv, xv, sizeof(double) * nv));
        TRY(upload(h, d_yv, yv, sizeof(double) * nv));
        TRY(upload(h, d_xf, xf, sizeof(double) * nf));
        TRY(upload(h, d_yf, yf, sizeof(double) * nf));
        TRY(ensure(h, d_out, sizeof(double) * 12 * nel));
        double* o = P<double>(d_out);
        double *dOm = o, *dxc = o + nel, *dyc = o + 2 * nel, *dnx = o + 3 * nel, *dny = o + 6 * nel, *dG = o + 9 * nel;
        double* dtau = nullptr;
        if (tau) {
            TRY(ensure(h, d_tau, sizeof(double) * nf));
            TRY(copy(h, d_tau.p, tau, sizeof(double) * nf));   // faces without elements keep their value
            dtau = P<double>(d_tau);
        }
        LAUNCH(h, k_geometry, nblocks(nel), kBlock, (int)nel, P<int>(d_e2v), P<int>(d_e2f), P<double>(d_xv), P<double>(d_yv),
               P<double>(d_xf), P<double>(d_yf), numbering, tau_r, dOm, dxc, dyc, dnx, dny, dG, dtau);
        int flag = 0;
        TRY(read_err_flag(h, &flag));
        if (flag) return fail(h, FCFV_ERR_INVALID, "fcfv_geometry: e2v / e2f entry out of range");
        if (Omega) TRY(copy(h, Omega, dOm, sizeof(double) * nel));
        if (xc) TRY(copy(h, xc, dxc, sizeof(double) * nel));
        if (yc) TRY(copy(h, yc, dyc, sizeof(double) * nel));
        if (nx) TRY(copy(h, nx, dnx, sizeof(double) * 3 * nel));
        if (ny) TRY(copy(h, ny, dny, sizeof(double) * 3 * nel));
        if (Gamma) TRY(copy(h, Gamma, dG, sizeof(double) * 3 * nel));
        if (tau) TRY(copy(h, tau, dtau, sizeof(double) * nf));
        return sync(h);
    };
    rc = body();
    cudaStreamSynchronize(h->stream);
    for (DBuf* b : {&d_e2v, &d_e2f, &d_xv, &d_yv, &d_xf, &d_yf, &d_out, &d_tau}) release(h, *b);
    return rc;
}

// -------------------------------------------------------------------------------------------------
namespace {
// in-place: counts -> 1-based exclusive ranks (+ total + 1 at [n])
int scan64_ranks(fcfv_t* h, unsigned long long* a, long long n, DBuf& chunks) {
    const int nchunks = (int)((n + 4095) / 4096);
    TRY(ensure(h, chunks, sizeof(unsigned long long) * (nchunks + 1)));
    LAUNCH(h, k_scan64_local, nchunks, 1024, a, n, P<unsigned long long>(chunks));
    LAUNCH(h, k_scan64_chunks, 1, 1, P<unsigned long long>(chunks), nchunks);
    LAUNCH(h, k_scan64_finish, nblocks(n + 1, 256), 256, a, n, P<unsigned long long>(chunks), nchunks);
    return FCFV_OK;
}
}  // namespace

int fcfv_first_seen_faces(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, int64_t* new_of_old, int64_t* e2f_new) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) {
        const int rc = fcfv_first_seen_faces(multi_part0(h), nel, nf, e2f, new_of_old, e2f_new);
        if (rc) h->err = multi_part0(h)->err;
        return rc;
    }
    CU(cudaSetDevice(h->device));
    if (nel < 0 || nf < 0 || nel >= (1LL << 29) || nf >= (1LL << 30)) return fail(h, FCFV_ERR_INVALID, "fcfv_first_seen_faces: size out of range");
    if (!e2f) return fail(h, FCFV_ERR_INVALID, "fcfv_first_seen_faces: null e2f");
    DBuf d_e2f, d_f2e, d_cnt, d_flag, d_orph, d_perm, d_out, d_chunks;
    auto body = [&]() -> int {
        CU(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
        TRY(upload(h, h->tmp, e2f, sizeof(int64_t) * 3 * nel));
        TRY(ensure(h, d_e2f, sizeof(int) * 3 * nel));
        LAUNCH(h, k_convert_e2f, nblocks(3 * nel, 256), 256, P<long long>(h->tmp), P<int>(d_e2f), (long long)(3 * nel), (int)nf, P<int>(h->err_flag));
        TRY(ensure(h, d_f2e, sizeof(int2) * nf));
        TRY(ensure(h, d_cnt, sizeof(int) * nf));
        LAUNCH(h, k_f2e_init, nblocks(nf, 256), 256, P<int2>(d_f2e), P<int>(d_cnt), (int)nf);
        LAUNCH(h, k_f2e_build, nblocks(nel, 256), 256, P<int>(d_e2f), P<int2>(d_f2e), P<int>(d_cnt), (int)nel);
        LAUNCH(h, k_f2e_orphans, nblocks(nf, 256), 256, P<int2>(d_f2e), P<int>(d_cnt), (int)nf);
        int flag = 0;
        TRY(read_err_flag(h, &flag));
        if (flag) return fail(h, FCFV_ERR_INVALID, "fcfv_first_seen_faces: e2f contains a face id outside 1..nf");
        const long long ncode = 4 * nel;
        TRY(ensure(h, d_flag, sizeof(unsigned long long) * (ncode + 1)));
        TRY(ensure(h, d_orph, sizeof(unsigned long long) * (nf + 1)));
        CU(cudaMemsetAsync(d_flag.p, 0, sizeof(unsigned long long) * (ncode + 1), h->stream));
        LAUNCH(h, k_fs_mark, nblocks(nf, 256), 256, P<int2>(d_f2e), (int)nf, P<unsigned long long>(d_flag));
        LAUNCH(h, k_fs_orphan, nblocks(nf, 256), 256, P<int2>(d_f2e), (int)nf, P<unsigned long long>(d_orph));
        TRY(scan64_ranks(h, P<unsigned long long>(d_flag), ncode, d_chunks));
        TRY(scan64_ranks(h, P<unsigned long long>(d_orph), nf, d_chunks));
        unsigned long long total1 = 0;
        CU(cudaMemcpyAsync(&total1, P<unsigned long long>(d_flag) + ncode, sizeof total1, cudaMemcpyDeviceToHost, h->stream));
        TRY(sync(h));
        TRY(ensure(h, d_perm, sizeof(long long) * std::max<long long>(nf, 1)));
        LAUNCH(h, k_fs_perm, nblocks(nf, 256), 256, P<int2>(d_f2e), (int)nf, P<unsigned long long>(d_flag), (long long)(total1 - 1),
               P<unsigned long long>(d_orph), P<long long>(d_perm));
        if (new_of_old) TRY(copy(h, new_of_old, d_perm.p, sizeof(long long) * nf));
        if (e2f_new) {
            TRY(ensure(h, d_out, sizeof(long long) * std::max<long long>(3 * nel, 1)));
            LAUNCH(h, k_fs_apply, nblocks(3 * nel, 256), 256, P<int>(d_e2f), (long long)(3 * nel), P<long long>(d_perm), P<long long>(d_out));
            TRY(copy(h, e2f_new, d_out.p, sizeof(long long) * 3 * nel));
        }
        CU(cudaGetLastError());
        return sync(h);
    };
    const int rc = body();
    cudaStreamSynchronize(h->stream);
    for (DBuf* b : {&d_e2f, &d_f2e, &d_cnt, &d_flag, &d_orph, &d_perm, &d_out, &d_chunks}) release(h, *b);
    return rc;
}

int fcfv_hilbert_elements(fcfv_t* h, int64_t nel, const double* xc, const double* yc, int64_t* old_of_new) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) {
        const int rc = fcfv_hilbert_elements(multi_part0(h), nel, xc, yc, old_of_new);
        if (rc) h->err = multi_part0(h)->err;
        return rc;
    }
    CU(cudaSetDevice(h->device));
    if (nel < 0 || nel >= (1LL << 31)) return fail(h, FCFV_ERR_INVALID, "fcfv_hilbert_elements: size out of range");
    if (!xc || !yc || !old_of_new) return fail(h, FCFV_ERR_INVALID, "fcfv_hilbert_elements: null array");
    DBuf d_x, d_y, d_out, d_tmp;
    auto body = [&]() -> int {
        TRY(upload(h, d_x, xc, sizeof(double) * nel));
        TRY(upload(h, d_y, yc, sizeof(double) * nel));
        TRY(ensure(h, d_out, sizeof(long long) * std::max<long long>(nel, 1)));
        size_t bytes = 0;
        CU(hilbert_element_order(h->stream, (int)nel, P<double>(d_x), P<double>(d_y), P<long long>(d_out), nullptr, 0, &bytes));
        TRY(ensure(h, d_tmp, bytes));
        CU(hilbert_element_order(h->stream, (int)nel, P<double>(d_x), P<double>(d_y), P<long long>(d_out), d_tmp.p, d_tmp.cap, nullptr));
        TRY(copy(h, old_of_new, d_out.p, sizeof(long long) * nel));
        return sync(h);
    };
    const int rc = body();
    cudaStreamSynchronize(h->stream);
    for (DBuf* b : {&d_x, &d_y, &d_out, &d_tmp}) release(h, *b);
    return rc;
}

// -------------------------------------------------------------------------------------------------
int fcfv_set_sources(fcfv_t* h, const double* sex, const double* sey) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_set_sources: no mesh set"));
    if (h->multi) return multi_set_sources(h, sex, sey);
    CU(cudaSetDevice(h->device));
    if (!sex || !sey) return fail(h, FCFV_ERR_INVALID, "fcfv_set_sources: null array");
    TRY(upload_cached(h, h->sex, sex, sizeof(double) * h->nel));
    TRY(upload_cached(h, h->sey, sey, sizeof(double) * h->nel));
    h->have_sources = true;
    return finish(h);
}

int fcfv_set_face_data(fcfv_t* h, const double* VxDir, const double* VyDir, const double* sxxNeu, const double* syyNeu,
                       const double* sxyNeu, const double* syxNeu) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_set_face_data: no mesh set"));
    if (h->multi) return multi_set_face_data(h, VxDir, VyDir, sxxNeu, syyNeu, sxyNeu, syxNeu);
    CU(cudaSetDevice(h->device));
    const size_t b = sizeof(double) * h->nf;
    DBuf* dst[6] = {&h->VxDir, &h->VyDir, &h->sxx, &h->syy, &h->sxy, &h->syx};
    const double* src[6] = {VxDir, VyDir, sxxNeu, syyNeu, sxyNeu, syxNeu};
    for (int k = 0; k < 6; k++) {
        const bool fresh = dst[k]->cap < b;
        TRY(ensure(h, *dst[k], b));
        if (src[k] && h->sparse_faces && !is_device(src[k])) {
            // only the faces the array is read on (Dirichlet / Neumann) cross the link
            if (fresh || !h->have_face_data) CU(cudaMemsetAsync(dst[k]->p, 0, b, h->stream));
            TRY(upload_listed(h, *dst[k], src[k], k));
        } else if (src[k]) TRY(upload_cached(h, *dst[k], src[k], b));
        else if (fresh || !h->have_face_data) { dst[k]->hm.epoch = -1; CU(cudaMemsetAsync(dst[k]->p, 0, b, h->stream)); }
    }
    h->have_face_data = true;
    return finish(h);
}

namespace {
// fcfv_set_face_data of one part of a multi-GPU handle straight from the caller's GLOBAL arrays (faces: local -> global ids):
// only possible when the part lists its boundary faces (sparse_faces); NULL = keep what is resident
int set_face_data_listed(fcfv_t* h, const double* const src[6], const int64_t* faces) {
    CU(cudaSetDevice(h->device));
    const size_t b = sizeof(double) * h->nf;
    DBuf* dst[6] = {&h->VxDir, &h->VyDir, &h->sxx, &h->syy, &h->sxy, &h->syx};
    for (int k = 0; k < 6; k++) {
        const bool fresh = dst[k]->cap < b;
        TRY(ensure(h, *dst[k], b));
        if (fresh || !h->have_face_data) { dst[k]->hm.epoch = -1; CU(cudaMemsetAsync(dst[k]->p, 0, b, h->stream)); }
        if (src[k]) TRY(upload_listed(h, *dst[k], src[k], k, faces));
    }
    h->have_face_data = true;
    return finish(h);
}
}  // namespace

int fcfv_set_solution(fcfv_t* h, const double* Vxh, const double* Vyh, const double* Pe) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_set_solution: no mesh set"));
    if (h->multi) return multi_set_solution(h, Vxh, Vyh, Pe);
    CU(cudaSetDevice(h->device));
    if (!Vxh || !Vyh || !Pe) return fail(h, FCFV_ERR_INVALID, "fcfv_set_solution: null array");
    TRY(upload_cached(h, h->Vxh, Vxh, sizeof(double) * h->nf));
    TRY(upload_cached(h, h->Vyh, Vyh, sizeof(double) * h->nf));
    TRY(upload_cached(h, h->Pe, Pe, sizeof(double) * h->nel));
    h->have_solution = true;
    return finish(h);
}

int fcfv_set_local(fcfv_t* h, const double* alpha, const double* beta, const double* Z) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_set_local: no mesh set"));
    if (h->multi) return multi_set_local(h, alpha, beta, Z);
    CU(cudaSetDevice(h->device));
    if (!alpha || !beta || !Z) return fail(h, FCFV_ERR_INVALID, "fcfv_set_local: null array");
    TRY(upload_cached(h, h->alpha, alpha, sizeof(double) * h->nel));
    TRY(upload_cached(h, h->beta, beta, sizeof(double) * 2 * h->nel));
    TRY(upload_cached(h, h->Z, Z, sizeof(double) * 4 * h->nel));
    h->coef_variant = -1;
    h->have_local = true;
    return finish(h);
}

// -------------------------------------------------------------------------------------------------
namespace {
template <int FORM, int COEF>
void launch_local(fcfv_t* h, double A, long long e0 = 0, long long e1 = -1) {
    const MeshView M = mesh_view(h);
    double* c = P<double>(h->coef);
    if (e1 < 0) e1 = h->nel;
    // mesh.τ is set by the element kernel (one writer per face: its highest adjacent element)
    TLAUNCH(h, T_LOCAL, (k_local<FORM, COEF>), nblocks(e1 - e0), kBlock, M, P<double>(h->sex), P<double>(h->sey), P<double>(h->VxDir),
            P<double>(h->VyDir), A, P<double>(h->alpha), P<double>(h->beta), P<double>(h->Z), P<double>(h->tau), COEF >= 0 ? c : nullptr,
            COEF >= 0 ? c + (size_t)h->nel : nullptr, (int)e0, (int)e1);
}

// coef_variant: FCFV_LOOP_OLD / FCFV_LOOP_NEW = the assembly of that loop follows (fcfv_run_pipeline): the element kernel
// writes its per-element coefficients as well, which saves the coefficient pass; -1 = plain ComputeFCFV
int run_local_impl(fcfv_t* h, int formulation, double A, int coef_variant) {
    TRY(check_form(h, formulation));
    TRY(need(h, h->have_mesh, "fcfv_run_local: no mesh set"));
    TRY(need(h, h->have_sources && h->have_face_data, "fcfv_run_local: sources / face data not set"));
    CU(cudaSetDevice(h->device));
    if (coef_variant >= 0 && (h->delta_dirty || h->aniso || h->coef.cap < sizeof(double) * 2 * (size_t)h->nel)) coef_variant = -1;
    const bool g = formulation == FCFV_GRADIENT;
    if (coef_variant == FCFV_LOOP_OLD) g ? launch_local<kGradient, kLoopOld>(h, A) : launch_local<kSymmetricGradient, kLoopOld>(h, A);
    else if (coef_variant == FCFV_LOOP_NEW) g ? launch_local<kGradient, kLoopNew>(h, A) : launch_local<kSymmetricGradient, kLoopNew>(h, A);
    else g ? launch_local<kGradient, -1>(h, A) : launch_local<kSymmetricGradient, -1>(h, A);
    h->coef_variant = coef_variant;
    return FCFV_OK;
}
}  // namespace

int fcfv_run_local(fcfv_t* h, int formulation, double A) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(check_form(h, formulation));
    if (h->multi) return multi_run_local(h, formulation, A);
    TRY(run_local_impl(h, formulation, A, -1));
    CU(cudaGetLastError());
    h->alpha.hm.epoch = h->beta.hm.epoch = h->Z.hm.epoch = h->tau.hm.epoch = -1;     // rewritten on the device
    h->have_local = true;
    return finish(h);
}

int fcfv_get_local(fcfv_t* h, double* alpha, double* beta, double* Z, double* tau) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_get_local(h, alpha, beta, Z, tau);
    TRY(need(h, h->have_local, "fcfv_get_local: local coefficients not computed"));
    CU(cudaSetDevice(h->device));
    if (alpha) TRY(copy(h, alpha, h->alpha.p, sizeof(double) * h->nel));
    if (beta) TRY(copy(h, beta, h->beta.p, sizeof(double) * 2 * h->nel));
    if (Z) TRY(copy(h, Z, h->Z.p, sizeof(double) * 4 * h->nel));
    if (tau) TRY(copy(h, tau, h->tau.p, sizeof(double) * h->nf));
    TRY(sync(h));
    // the caller's arrays now mirror the device buffers: handing them back (ElementAssemblyLoop(mesh, ae, be, ze, ...),
    // mesh.τ in the field refresh) costs no upload
    mirror_download(h, h->alpha, alpha, sizeof(double) * h->nel);
    mirror_download(h, h->beta, beta, sizeof(double) * 2 * h->nel);
    mirror_download(h, h->Z, Z, sizeof(double) * 4 * h->nel);
    mirror_download(h, h->tau, tau, sizeof(double) * h->nf);
    return FCFV_OK;
}

int fcfv_compute_local(fcfv_t* h, int formulation, double A, const double* sex, const double* sey, const double* VxDir,
                       const double* VyDir, double* alpha, double* beta, double* Z, double* tau) {
    return fcfv_compute_local_ex(h, formulation, A, sex, sey, VxDir, VyDir, nullptr, nullptr, nullptr, nullptr, alpha, beta, Z, tau);
}

namespace {
// alpha / beta / Z / tau leave on the handle's second stream (after everything queued on the main stream so far), so
// that uploads issued on the main stream afterwards run concurrently: PCIe is full duplex.  Pageable destinations go
// through the staged path instead (complete on return).  fcfv_local_downloaded() closes the transfer.
int download_local_async(fcfv_t* h, double* alpha, double* beta, double* Z, double* tau) {
    if (!h->stream_out) {
        CU(cudaStreamCreateWithFlags(&h->stream_out, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(h->ev_out, h->stream));
    CU(cudaStreamWaitEvent(h->stream_out, h->ev_out, 0));
    struct Out { double* dst; const void* src; size_t bytes; } outs[4] = {
        {alpha, h->alpha.p, sizeof(double) * (size_t)h->nel}, {beta, h->beta.p, sizeof(double) * 2 * (size_t)h->nel},
        {Z, h->Z.p, sizeof(double) * 4 * (size_t)h->nel}, {tau, h->tau.p, sizeof(double) * (size_t)h->nf}};
    for (const Out& o : outs) {
        if (!o.dst || o.bytes == 0) continue;
        if (is_device(o.dst) || is_pageable_host(o.dst)) { TRY(copy(h, o.dst, o.src, o.bytes)); continue; }
        CU(cudaMemcpyAsync(o.dst, o.src, o.bytes, cudaMemcpyDeviceToHost, h->stream_out));
        h->d2h_bytes += (long long)o.bytes;
    }
    return FCFV_OK;
}

int local_downloaded(fcfv_t* h, double* alpha, double* beta, double* Z, double* tau) {
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->stream_out));
    mirror_download(h, h->alpha, alpha, sizeof(double) * h->nel);
    mirror_download(h, h->beta, beta, sizeof(double) * 2 * h->nel);
    mirror_download(h, h->Z, Z, sizeof(double) * 4 * h->nel);
    mirror_download(h, h->tau, tau, sizeof(double) * h->nf);
    return FCFV_OK;
}
}  // namespace

// ComputeFCFV with the reference's full argument list (DiscretisationFCFV_Stokes.jl:8 also receives the four Neumann
// arrays and never reads them).  Here they are put to use: while alpha / beta / Z / tau come down, the Neumann arrays
// go up on the other direction of the link, and the assembly / residual calls that follow find them resident
// (upload cache on) -- the download costs no wall time beyond what the later uploads would have taken anyway.
int fcfv_compute_local_ex(fcfv_t* h, int formulation, double A, const double* sex, const double* sey, const double* VxDir,
                          const double* VyDir, const double* sxxNeu, const double* syyNeu, const double* sxyNeu, const double* syxNeu,
                          double* alpha, double* beta, double* Z, double* tau) {
    if (!h) return FCFV_ERR_INVALID;
    const bool was_async = h->async;
    h->async = true;
    int rc = fcfv_set_sources(h, sex, sey);
    if (!rc) {
        if (!VxDir || !VyDir) rc = fail(h, FCFV_ERR_INVALID, "fcfv_compute_local: null Dirichlet array");
        else rc = fcfv_set_face_data(h, VxDir, VyDir, nullptr, nullptr, nullptr, nullptr);
    }
    if (!rc) rc = fcfv_run_local(h, formulation, A);
    if (rc || h->multi) {
        h->async = was_async;
        if (rc) return rc;
        return fcfv_get_local(h, alpha, beta, Z, tau);
    }
    auto tail = [&]() -> int {
        TRY(download_local_async(h, alpha, beta, Z, tau));
        // without the cache the later calls would send them again anyway: nothing to gain
        if (h->cache_on && (sxxNeu || syyNeu || sxyNeu || syxNeu)) TRY(fcfv_set_face_data(h, nullptr, nullptr, sxxNeu, syyNeu, sxyNeu, syxNeu));
        return FCFV_OK;
    };
    rc = tail();
    h->async = was_async;
    const int rc2 = local_downloaded(h, alpha, beta, Z, tau);     // both streams are drained whatever happened above
    return rc ? rc : rc2;
}

namespace {
bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

constexpr int kLocalChunks = 8;

// the chunk pipeline needs a mesh worth the trouble and arrays that can be copied asynchronously (page-locked / registered)
bool chunk_pipeline_ok(fcfv_t* h, long long nel, std::initializer_list<const void*> arrays) {
    long long min_nel = 1LL << 20;          // below that a pass is a few milliseconds whichever way it is issued
    if (const char* v = getenv("FCFV_PIPELINE_MIN_NEL")) min_nel = atoll(v);
    if (h->multi || nel < min_nel || nel < 2 * kBlock) return false;
    for (const void* p : arrays)
        if (p && !is_pinned_host(p)) return false;
    return true;
}

// One ComputeFCFV with everything it needs still on the host, as a pipeline over element chunks: while chunk c of the
// per-element inputs goes up, the element kernel runs on chunk c-1 and alpha / beta / Z of chunk c-2 come down -- both
// directions of the link stay busy and the kernel costs nothing.  The element kernel reads, per element, only
// element-indexed arrays (geometry, tau_e, sources) besides the connectivity / tag records (complete before the first chunk)
// and VxDir / VyDir on Dirichlet faces (sent first); ke and delta are not read by it at all and follow the last chunk,
// under the tail of the downloads.  Only page-locked (or registered) host memory can be copied asynchronously, so this path
// needs every array to be that; the caller falls back to the plain sequence otherwise.
template <int FORM>
int compute_fcfv_chunks(fcfv_t* h, bool mesh, const double* Omega, const double* nx, const double* ny, const double* Gamma,
                        const double* ke, const double* delta, const double* taue, const double* tau_in, double A, const double* sex,
                        const double* sey, const double* VxDir, const double* VyDir, const double* const sig[4], double* alpha,
                        double* beta, double* Z, double* tau) {
    const long long nel = h->nel, nf = h->nf;
    const size_t be = sizeof(double) * (size_t)nel;
    if (!h->stream_out) {
        CU(cudaStreamCreateWithFlags(&h->stream_out, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
    }
    if (tau_in && h->has_orphans) TRY(upload_cached(h, h->tau, tau_in, sizeof(double) * nf));   // faces without elements keep the caller's tau
    TRY(fcfv_set_face_data(h, VxDir, VyDir, nullptr, nullptr, nullptr, nullptr));
    struct In { DBuf* b; const double* src; int ncol; bool skip; } ins[7] = {
        {&h->Om, mesh ? Omega : nullptr, 1, false}, {&h->nx, mesh ? nx : nullptr, 3, false}, {&h->ny, mesh ? ny : nullptr, 3, false},
        {&h->G, mesh ? Gamma : nullptr, 3, false}, {&h->taue, taue, 1, false}, {&h->sex, sex, 1, false}, {&h->sey, sey, 1, false}};
    for (In& in : ins) {
        TRY(ensure(h, *in.b, be * in.ncol));
        if (!in.src) { in.skip = true; continue; }
        in.skip = cache_hit(h, *in.b, in.src, be * in.ncol);     // resident mesh: an array the device already mirrors stays
        if (in.skip) { h->cache_hits++; h->cache_hit_bytes += (long long)(be * in.ncol); }
        else in.b->hm.epoch = -1;
    }
    if (mesh && !taue) { h->taue.hm.epoch = -1; CU(cudaMemsetAsync(h->taue.p, 0, be, h->stream)); }
    h->alpha.hm.epoch = h->beta.hm.epoch = h->Z.hm.epoch = h->tau.hm.epoch = -1;      // rewritten on the device
    h->coef_variant = -1;
    const int nchunks = nel >= (1LL << 24) ? 2 * kLocalChunks : kLocalChunks;     // large meshes: shorter fill and drain of the pipeline
    long long chunk = (nel + nchunks - 1) / nchunks;
    chunk = std::max<long long>(kBlock, (chunk + kBlock - 1) / kBlock * kBlock);
    struct Out { double* dst; const DBuf* b; int ncol; } outs[3] = {{alpha, &h->alpha, 1}, {beta, &h->beta, 2}, {Z, &h->Z, 4}};
    for (long long e0 = 0; e0 < nel; e0 += chunk) {
        const long long e1 = std::min(nel, e0 + chunk);
        const size_t off = sizeof(double) * (size_t)e0, len = sizeof(double) * (size_t)(e1 - e0);
        for (const In& in : ins) {
            if (in.skip) continue;
            for (int c = 0; c < in.ncol; c++)
                TRY(copy(h, (char*)in.b->p + c * be + off, (const char*)in.src + c * be + off, len));
        }
        launch_local<FORM, -1>(h, A, e0, e1);
        CU(cudaEventRecord(h->ev_out, h->stream));
        CU(cudaStreamWaitEvent(h->stream_out, h->ev_out, 0));
        for (const Out& o : outs) {
            if (!o.dst) continue;
            for (int c = 0; c < o.ncol; c++)
                CU(cudaMemcpyAsync((char*)o.dst + c * be + off, (const char*)o.b->p + c * be + off, len, cudaMemcpyDeviceToHost, h->stream_out));
            h->d2h_bytes += (long long)(len * o.ncol);
        }
    }
    CU(cudaGetLastError());
    if (tau && nf) {          // complete once every chunk has run (a face's tau comes from its highest element)
        CU(cudaMemcpyAsync(tau, h->tau.p, sizeof(double) * (size_t)nf, cudaMemcpyDeviceToHost, h->stream_out));
        h->d2h_bytes += (long long)(sizeof(double) * nf);
    }
    // what the element kernel does not read goes up under the tail of the downloads
    if (ke) TRY(upload_cached(h, h->ke, ke, be));
    if (delta) {
        const long long hits = h->cache_hits;
        TRY(upload_cached(h, h->dl, delta, be));
        if (h->cache_hits == hits) h->delta_dirty = true;
    }
    if (h->cache_on) {
        for (const In& in : ins)       // the chunks together are the whole array: it is a mirror from now on
            if (!in.skip && in.b != &h->Om && in.b != &h->nx && in.b != &h->ny && in.b != &h->G)
                in.b->hm = HostMirror{in.src, be * in.ncol, fingerprint(in.src, be * in.ncol), h->epoch};
        if (sig[0] || sig[1] || sig[2] || sig[3]) TRY(fcfv_set_face_data(h, nullptr, nullptr, sig[0], sig[1], sig[2], sig[3]));
    }
    h->have_sources = true;
    h->have_local = true;
    return FCFV_OK;
}
}  // namespace

// set_mesh (e2f != NULL) or field refresh (e2f == NULL: the mesh is resident) + ComputeFCFV in one call: what the shims issue
// for ComputeFCFV(mesh, ...).  Same results as fcfv_set_mesh / fcfv_update_fields followed by fcfv_compute_local_ex.
int fcfv_compute_fcfv(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, const int64_t* bc, const double* Omega, const double* nx,
                      const double* ny, const double* Gamma, const double* ke, const double* delta, const double* taue, int64_t ntaue,
                      const double* tau_in, int formulation, double A, const double* sex, const double* sey, const double* VxDir,
                      const double* VyDir, const double* sxxNeu, const double* syyNeu, const double* sxyNeu, const double* syxNeu,
                      double* alpha, double* beta, double* Z, double* tau) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(check_form(h, formulation));
    const bool mesh = e2f != nullptr;
    if (mesh && (!bc || !Omega || !nx || !ny || !Gamma || !ke || !delta)) return fail(h, FCFV_ERR_INVALID, "fcfv_compute_fcfv: null mesh array");
    if (!mesh) { TRY(need(h, h->have_mesh, "fcfv_compute_fcfv: no mesh resident and none given")); nel = h->nel; nf = h->nf; }
    if (taue && ntaue < nel) return fail(h, FCFV_ERR_INVALID, "taue has %lld entries, need at least nel=%lld", (long long)ntaue, (long long)nel);
    if (!sex || !sey || !VxDir || !VyDir) return fail(h, FCFV_ERR_INVALID, "fcfv_compute_fcfv: null source / Dirichlet array");
    const bool chunks = mesh ? chunk_pipeline_ok(h, nel, {Omega, nx, ny, Gamma, ke, delta, taue, sex, sey, alpha, beta, Z, tau})
                             : chunk_pipeline_ok(h, nel, {ke, delta, taue, sex, sey, alpha, beta, Z, tau});
    if (!chunks) {
        if (mesh) {
            TRY(fcfv_set_mesh(h, nel, nf, e2f, bc, Omega, nx, ny, Gamma, ke, delta, taue, ntaue));
            if (tau_in && !h->multi && h->has_orphans) TRY(fcfv_update_fields(h, nullptr, nullptr, nullptr, 0, tau_in));
            if (tau_in && h->multi && multi_has_orphans(h)) TRY(fcfv_update_fields(h, nullptr, nullptr, nullptr, 0, tau_in));
        } else {
            const bool orph = h->multi ? multi_has_orphans(h) != 0 : h->has_orphans;
            TRY(fcfv_update_fields(h, ke, delta, taue, ntaue, orph ? tau_in : nullptr));
        }
        return fcfv_compute_local_ex(h, formulation, A, sex, sey, VxDir, VyDir, sxxNeu, syyNeu, sxyNeu, syxNeu, alpha, beta, Z, tau);
    }
    if (mesh) TRY(set_mesh_connectivity(h, nel, nf, e2f, bc));
    CU(cudaSetDevice(h->device));
    h->pipelined_calls++;
    const bool was_async = h->async;
    h->async = true;
    h->have_mesh = true;            // the entries used below ask for it; withdrawn again if the fields do not arrive
    const double* const sig[4] = {sxxNeu, syyNeu, sxyNeu, syxNeu};
    int rc = formulation == FCFV_GRADIENT
                 ? compute_fcfv_chunks<kGradient>(h, mesh, Omega, nx, ny, Gamma, ke, delta, taue, tau_in, A, sex, sey, VxDir, VyDir, sig, alpha, beta, Z, tau)
                 : compute_fcfv_chunks<kSymmetricGradient>(h, mesh, Omega, nx, ny, Gamma, ke, delta, taue, tau_in, A, sex, sey, VxDir, VyDir, sig, alpha, beta, Z, tau);
    h->async = was_async;
    const int rc2 = local_downloaded(h, alpha, beta, Z, tau);
    if (rc && mesh) h->have_mesh = false;
    return rc ? rc : rc2;
}

// -------------------------------------------------------------------------------------------------
namespace {

CoefView coef_view(fcfv_t* h, bool schur) {
    double* c = P<double>(h->coef);
    double* x = P<double>(h->coefx);
    const size_t nel = (size_t)h->nel;
    return CoefView{c, c + nel, h->aniso ? x : nullptr, h->aniso ? x + nel : nullptr, h->aniso ? x + 2 * nel : nullptr,
                    schur ? P<double>(h->coefs) : nullptr};
}

template <int VARIANT, int FORM, int EXTRA, bool ANISO>
int launch_assemble(fcfv_t* h, double A, double gamma) {
    const MeshView M = mesh_view(h);
    const LocalView L = local_view(h);
    const FaceDataView D = face_view(h);
    const CoefView Cf = coef_view(h, EXTRA == kExtraSchur);
    const int nblk = nblocks(h->nf, kAsmFaces);
    const int nchunks = nblocks(nblk, kScanChunk);
    int* sums = P<int>(h->sums);
    int* local = P<int>(h->bases);
    long long* chunk_tot = P<long long>(h->chunks);
    long long* chunk_base = chunk_tot + (size_t)kNumSums * nchunks;
    // fp was zeroed by the caller; only Dirichlet faces contribute: walk their list when the mesh has one
    if (h->sparse_faces) {
        if (!h->dir_list.empty())
            TLAUNCH(h, T_OTHER, k_asm_fp_listed, nblocks((long long)h->dir_list.size(), 256), 256, M, D, P<int>(h->dir_idx), (int)h->dir_list.size(),
                    P<double>(h->fp));
    } else {
        TLAUNCH(h, T_OTHER, k_asm_fp, nblocks((h->nf + 3) / 4, 256), 256, M, D, P<double>(h->fp));
    }
    if (!ANISO && EXTRA != kExtraSchur && h->coef_variant == VARIANT) {
        // 1/alpha and c11 were written by the element kernel of this pass (fcfv_run_pipeline)
    } else if (!ANISO && EXTRA != kExtraSchur && (h->nel & 1) == 0 && h->nel > 0)
        TLAUNCH(h, T_COEF, k_elem_coef2<VARIANT>, nblocks(h->nel / 2, 256), 256, (int)(h->nel / 2), (const double2*)L.alpha,
                (const double2*)M.ke, (const double2*)M.Om, (double2*)Cf.ia, (double2*)Cf.c11);
    else
        TLAUNCH(h, T_COEF, k_elem_coef<VARIANT>, nblocks(h->nel, 256), 256, M.nel, L.alpha, M.ke, M.dl, M.Om, gamma, (double*)Cf.ia,
                (double*)Cf.c11, (double*)Cf.c12, (double*)Cf.c21, (double*)Cf.c22, (double*)Cf.cs);
    h->coef_variant = (!ANISO && EXTRA != kExtraSchur) ? VARIANT : -1;      // what `coef` now holds for the resident alpha / ke
    TLAUNCH(h, T_COUNT, (k_asm_count<VARIANT, FORM, EXTRA, ANISO>), nblk, kAsmBlock, M, L, Cf, D, A, sums, nblk);
    TLAUNCH(h, T_SCAN, k_scan_local, dim3(nchunks, kNumSums), 1024, sums, local, nblk, nchunks, chunk_tot);
    Csr& X = EXTRA == kExtraSchur ? h->Kuusc : h->Muu;   // third block
    CsrOut ouu{h->Kuu.rowptr.p, P<int>(h->Kuu.col), P<double>(h->Kuu.val), (long long)(h->Kuu.val.cap / sizeof(double))};
    CsrOut oup{h->Kup.rowptr.p, P<int>(h->Kup.col), P<double>(h->Kup.val), (long long)(h->Kup.val.cap / sizeof(double))};
    CsrOut oex{X.rowptr.p, P<int>(X.col), P<double>(X.val), (long long)(X.val.cap / sizeof(double))};
    if (h->rp64) {
        TLAUNCH(h, T_SCAN, k_scan_chunks<long long>, 1, 32, chunk_tot, chunk_base, nchunks, P<long long>(h->totals),
                (long long*)ouu.rowptr, (long long*)oup.rowptr, (long long*)oex.rowptr, M.nf, (int)(EXTRA != kExtraNone), EXTRA == kExtraSchur ? 9 : 8);
        TLAUNCH(h, T_FILL, (k_asm_fill<VARIANT, FORM, EXTRA, ANISO, long long>), nblk, kAsmBlock, M, L, Cf, D, A, local, chunk_base, nblk,
                nchunks, ouu, oup, oex, P<double>(h->fu));
    } else {
        TLAUNCH(h, T_SCAN, k_scan_chunks<int>, 1, 32, chunk_tot, chunk_base, nchunks, P<long long>(h->totals), (int*)ouu.rowptr,
                (int*)oup.rowptr, (int*)oex.rowptr, M.nf, (int)(EXTRA != kExtraNone), EXTRA == kExtraSchur ? 9 : 8);
        TLAUNCH(h, T_FILL, (k_asm_fill<VARIANT, FORM, EXTRA, ANISO, int>), nblk, kAsmBlock, M, L, Cf, D, A, local, chunk_base, nblk, nchunks,
                ouu, oup, oex, P<double>(h->fu));
    }
    return FCFV_OK;
}

template <int VARIANT, int FORM>
int launch_assemble_x(fcfv_t* h, double A, int extra, double gamma) {
    const bool an = VARIANT == kLoopNew && h->aniso;
    switch (extra) {
        case kExtraMuu:
            return an ? launch_assemble<VARIANT, FORM, kExtraMuu, true>(h, A, gamma) : launch_assemble<VARIANT, FORM, kExtraMuu, false>(h, A, gamma);
        case kExtraSchur:
            return an ? launch_assemble<VARIANT, FORM, kExtraSchur, true>(h, A, gamma) : launch_assemble<VARIANT, FORM, kExtraSchur, false>(h, A, gamma);
        default:
            return an ? launch_assemble<VARIANT, FORM, kExtraNone, true>(h, A, gamma) : launch_assemble<VARIANT, FORM, kExtraNone, false>(h, A, gamma);
    }
}

int fetch_totals(fcfv_t* h) {
    if (h->multi) return multi_fetch_totals(h);
    if (h->totals_valid) return FCFV_OK;
    CU(cudaMemcpyAsync(h->totals_host, h->totals.p, sizeof(long long) * 10, cudaMemcpyDeviceToHost, h->stream));
    TRY(sync(h));
    h->Kuu.nnz = h->totals_host[6];
    h->Kup.nnz = h->totals_host[7];
    if (h->Muu.nnz == -2) h->Muu.nnz = h->totals_host[8];
    if (h->Kuusc.nnz == -2) h->Kuusc.nnz = h->totals_host[9];
    h->totals_valid = true;
    return FCFV_OK;
}

// extra: kExtraNone / kExtraMuu / kExtraSchur (gamma = Powell-Hestenes penalty, only read for the latter)
int run_assemble_impl(fcfv_t* h, int variant, int formulation, double A, int extra, double gamma) {
    TRY(check_form(h, formulation));
    if (variant != FCFV_LOOP_OLD && variant != FCFV_LOOP_NEW)
        return fail(h, FCFV_ERR_INVALID, "variant must be FCFV_LOOP_OLD or FCFV_LOOP_NEW (got %d)", variant);
    if (h->multi) return multi_run_assemble(h, variant, formulation, A, extra, gamma);
    TRY(need(h, h->have_mesh, "fcfv_run_assemble: no mesh set"));
    TRY(need(h, h->have_local, "fcfv_run_assemble: local coefficients missing (fcfv_compute_local / fcfv_set_local)"));
    TRY(need(h, h->have_face_data, "fcfv_run_assemble: face data not set"));
    CU(cudaSetDevice(h->device));
    const long long nel = h->nel, nf = h->nf;
    const size_t rpw = h->rp64 ? 8 : 4;
    const int nblk = nblocks(nf, kAsmFaces);
    const int nchunks = nblocks(nblk, kScanChunk);
    // capacity = upper bounds (10 / 2 / 5 entries per row), so no host round trip between the passes
    TRY(ensure(h, h->Kuu.rowptr, rpw * (2 * nf + 1)));
    TRY(ensure(h, h->Kuu.col, sizeof(int) * 20 * nf));
    TRY(ensure(h, h->Kuu.val, sizeof(double) * 20 * nf));
    TRY(ensure(h, h->Kup.rowptr, rpw * (2 * nf + 1)));
    TRY(ensure(h, h->Kup.col, sizeof(int) * 4 * nf));
    TRY(ensure(h, h->Kup.val, sizeof(double) * 4 * nf));
    if (extra == kExtraMuu) {
        TRY(ensure(h, h->Muu.rowptr, rpw * (2 * nf + 1)));
        TRY(ensure(h, h->Muu.col, sizeof(int) * 10 * nf));
        TRY(ensure(h, h->Muu.val, sizeof(double) * 10 * nf));
    }
    if (extra == kExtraSchur) {
        TRY(ensure(h, h->Kuusc.rowptr, rpw * (2 * nf + 1)));
        TRY(ensure(h, h->Kuusc.col, sizeof(int) * 20 * nf));
        TRY(ensure(h, h->Kuusc.val, sizeof(double) * 20 * nf));
        TRY(ensure(h, h->coefs, sizeof(double) * nel));
    }
    TRY(ensure(h, h->fu, sizeof(double) * 2 * nf));
    TRY(ensure(h, h->fp, sizeof(double) * nel));
    if (h->delta_dirty) {   // does any element have delta != 1?  (only after delta was uploaded)
        CU(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
        LAUNCH(h, k_delta_flag, nblocks(nel, 256), 256, P<double>(h->dl), (int)nel, P<int>(h->err_flag));
        int flag = 0;
        TRY(read_err_flag(h, &flag));
        CU(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
        h->aniso = flag != 0;
        h->delta_dirty = false;
    }
    TRY(ensure(h, h->coef, sizeof(double) * 2 * nel));
    if (h->aniso) TRY(ensure(h, h->coefx, sizeof(double) * 3 * nel));
    TRY(ensure(h, h->sums, sizeof(int) * kNumSums * (size_t)nblk));
    TRY(ensure(h, h->bases, sizeof(int) * kNumSums * (size_t)nblk));
    TRY(ensure(h, h->chunks, sizeof(long long) * 2 * kNumSums * (size_t)nchunks));
    h->Kuu.nrows = 2 * nf; h->Kuu.ncols = 2 * nf;
    h->Kup.nrows = 2 * nf; h->Kup.ncols = nel;
    h->Muu.nrows = 2 * nf; h->Muu.ncols = 2 * nf;
    h->Kuusc.nrows = 2 * nf; h->Kuusc.ncols = 2 * nf;
    h->Kuu.nnz = h->Kup.nnz = -2;            // assembled, count not fetched yet
    if (extra != kExtraSchur) h->Muu.nnz = extra == kExtraMuu ? -2 : -1;   // fcfv_schur leaves an assembled Muu in place
    h->Kuusc.nnz = extra == kExtraSchur ? -2 : -1;
    h->totals_valid = false;
    h->last_variant = variant; h->last_form = formulation; h->last_A = A;
    // only elements with a Dirichlet face get a value from k_asm_fp, everything else keeps +0.0
    CU(cudaMemsetAsync(h->fp.p, 0, sizeof(double) * nel, h->stream));
    CU(cudaEventRecord(h->ev0, h->stream));
    int rc;
    if (variant == FCFV_LOOP_OLD)
        rc = formulation == FCFV_GRADIENT ? launch_assemble_x<kLoopOld, kGradient>(h, A, extra, gamma)
                                          : launch_assemble_x<kLoopOld, kSymmetricGradient>(h, A, extra, gamma);
    else
        rc = formulation == FCFV_GRADIENT ? launch_assemble_x<kLoopNew, kGradient>(h, A, extra, gamma)
                                          : launch_assemble_x<kLoopNew, kSymmetricGradient>(h, A, extra, gamma);
    TRY(rc);
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaGetLastError());
    return finish(h);
}

}  // namespace

int fcfv_run_assemble(fcfv_t* h, int variant, int formulation, double A, int want_muu) {
    if (!h) return FCFV_ERR_INVALID;
    return run_assemble_impl(h, variant, formulation, A, want_muu ? kExtraMuu : kExtraNone, 0.0);
}

int fcfv_run_assemble_schur(fcfv_t* h, int variant, int formulation, double A, double gamma) {
    if (!h) return FCFV_ERR_INVALID;
    return run_assemble_impl(h, variant, formulation, A, kExtraSchur, gamma);
}

int fcfv_schur(fcfv_t* h, double gamma, int64_t* nnz_sc) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->last_variant >= 0 && h->Kuu.nnz != -1, "fcfv_schur: call fcfv_assemble first"));
    const bool was_async = h->async;
    h->async = true;
    const int rc = run_assemble_impl(h, h->last_variant, h->last_form, h->last_A, kExtraSchur, gamma);
    h->async = was_async;
    TRY(rc);
    TRY(fetch_totals(h));
    if (nnz_sc) *nnz_sc = h->Kuusc.nnz;
    return FCFV_OK;
}

int fcfv_get_nnz(fcfv_t* h, int64_t* nnz_uu, int64_t* nnz_up, int64_t* nnz_muu) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_get_nnz: nothing assembled"));
    if (!h->multi) CU(cudaSetDevice(h->device));
    TRY(fetch_totals(h));
    if (nnz_uu) *nnz_uu = h->Kuu.nnz;
    if (nnz_up) *nnz_up = h->Kup.nnz;
    if (nnz_muu) *nnz_muu = h->Muu.nnz < 0 ? 0 : h->Muu.nnz;
    return FCFV_OK;
}

int fcfv_assemble(fcfv_t* h, int variant, int formulation, double A, const double* alpha, const double* beta,
                  const double* Z, const double* VxDir, const double* VyDir, const double* sxxNeu, const double* syyNeu,
                  const double* sxyNeu, const double* syxNeu, int want_muu, int64_t* nnz_uu, int64_t* nnz_up,
                  int64_t* nnz_muu, double* seconds) {
    if (!h) return FCFV_ERR_INVALID;
    const bool was_async = h->async;
    h->async = true;
    int rc = FCFV_OK;
    if (alpha || beta || Z) rc = fcfv_set_local(h, alpha, beta, Z);
    if (!rc) rc = fcfv_set_face_data(h, VxDir, VyDir, sxxNeu, syyNeu, sxyNeu, syxNeu);
    if (!rc) rc = fcfv_run_assemble(h, variant, formulation, A, want_muu);
    h->async = was_async;
    if (rc) return rc;
    TRY(fcfv_get_nnz(h, nnz_uu, nnz_up, nnz_muu));   // synchronises
    if (seconds && h->multi) {
        *seconds = multi_asm_seconds(h);       // host clock around the launches on all devices and their completion
    } else if (seconds) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *seconds = 1e-3 * ms;
    }
    return FCFV_OK;
}

int fcfv_index_width(fcfv_t* h) { return h && h->rp64 ? 8 : 4; }

namespace {
Csr* block_of(fcfv_t* h, int block) {
    switch (block) {
        case FCFV_KUU: return &h->Kuu;
        case FCFV_KUP: return &h->Kup;
        case FCFV_MUU: return &h->Muu;
        case FCFV_KUUSC: return &h->Kuusc;
        default: return nullptr;
    }
}
}  // namespace

int fcfv_block_size(fcfv_t* h, int block, int64_t* nrows, int64_t* ncols, int64_t* nnz) {
    if (!h) return FCFV_ERR_INVALID;
    Csr* c = block_of(h, block);
    if (!c) return fail(h, FCFV_ERR_INVALID, "unknown block %d", block);
    TRY(need(h, c->nnz != -1, "fcfv_block_size: block not assembled"));
    if (!h->multi) CU(cudaSetDevice(h->device));
    TRY(fetch_totals(h));
    if (nrows) *nrows = c->nrows;
    if (ncols) *ncols = c->ncols;
    if (nnz) *nnz = c->nnz;
    return FCFV_OK;
}

int fcfv_export(fcfv_t* h, int block, int layout, void* ptr, void* idx, double* val) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_export(h, block, layout, ptr, idx, val);
    Csr* c = block_of(h, block);
    if (!c) return fail(h, FCFV_ERR_INVALID, "unknown block %d", block);
    TRY(need(h, c->nnz != -1, "fcfv_export: block not assembled"));
    CU(cudaSetDevice(h->device));
    TRY(fetch_totals(h));
    const size_t rpw = h->rp64 ? 8 : 4;
    const long long nnz = c->nnz, nrows = c->nrows, ncols = c->ncols;
    if (layout == FCFV_CSR0) {
        if (ptr) TRY(copy(h, ptr, c->rowptr.p, rpw * (nrows + 1)));
        if (idx) TRY(copy(h, idx, c->col.p, sizeof(int) * nnz));
        if (val) TRY(copy(h, val, c->val.p, sizeof(double) * nnz));
        return sync(h);
    }
    if (layout != FCFV_CSC1_INT64) return fail(h, FCFV_ERR_INVALID, "unknown layout %d", layout);
    if (!ptr || !idx || !val) return fail(h, FCFV_ERR_INVALID, "fcfv_export: CSC export needs ptr, idx and val");
    // CSR -> Julia CSC on the device (k_csc_*): count, scan, scatter, per-column sort by row
    if (nrows >= (1LL << 31) || ncols >= (1LL << 31)) return fail(h, FCFV_ERR_INVALID, "fcfv_export: block too large");
    DBuf d_cp, d_rv, d_nz, d_cur, d_chunks;
    auto body = [&]() -> int {
        const int nchunks = (int)((ncols + 4095) / 4096);
        TRY(ensure(h, d_cp, sizeof(long long) * (ncols + 1)));
        TRY(ensure(h, d_rv, sizeof(long long) * std::max<long long>(nnz, 1)));
        TRY(ensure(h, d_nz, sizeof(double) * std::max<long long>(nnz, 1)));
        TRY(ensure(h, d_cur, sizeof(unsigned int) * std::max<long long>(ncols, 1)));
        TRY(ensure(h, d_chunks, sizeof(unsigned long long) * (nchunks + 1)));
        CU(cudaMemsetAsync(d_cp.p, 0, sizeof(long long) * (ncols + 1), h->stream));
        CU(cudaMemsetAsync(d_cur.p, 0, sizeof(unsigned int) * std::max<long long>(ncols, 1), h->stream));
        unsigned long long* cnt = P<unsigned long long>(d_cp);
        if (h->rp64) LAUNCH(h, k_csc_count<long long>, nblocks(nrows, 256), 256, (int)nrows, P<long long>(c->rowptr), P<int>(c->col), cnt);
        else LAUNCH(h, k_csc_count<int>, nblocks(nrows, 256), 256, (int)nrows, P<int>(c->rowptr), P<int>(c->col), cnt);
        LAUNCH(h, k_scan64_local, nchunks, 1024, cnt, ncols, P<unsigned long long>(d_chunks));
        LAUNCH(h, k_scan64_chunks, 1, 1, P<unsigned long long>(d_chunks), nchunks);
        LAUNCH(h, k_scan64_finish, nblocks(ncols + 1, 256), 256, cnt, ncols, P<unsigned long long>(d_chunks), nchunks);
        if (h->rp64)
            LAUNCH(h, k_csc_scatter<long long>, nblocks(nrows, 256), 256, (int)nrows, P<long long>(c->rowptr), P<int>(c->col), P<double>(c->val),
                   P<long long>(d_cp), P<unsigned int>(d_cur), P<long long>(d_rv), P<double>(d_nz));
        else
            LAUNCH(h, k_csc_scatter<int>, nblocks(nrows, 256), 256, (int)nrows, P<int>(c->rowptr), P<int>(c->col), P<double>(c->val),
                   P<long long>(d_cp), P<unsigned int>(d_cur), P<long long>(d_rv), P<double>(d_nz));
        LAUNCH(h, k_csc_sort, nblocks(ncols, 256), 256, (int)ncols, P<long long>(d_cp), P<long long>(d_rv), P<double>(d_nz));
        CU(cudaGetLastError());
        TRY(copy(h, ptr, d_cp.p, sizeof(long long) * (ncols + 1)));
        TRY(copy(h, idx, d_rv.p, sizeof(long long) * nnz));
        TRY(copy(h, val, d_nz.p, sizeof(double) * nnz));
        return sync(h);
    };
    const int rc = body();
    cudaStreamSynchronize(h->stream);
    for (DBuf* bf : {&d_cp, &d_rv, &d_nz, &d_cur, &d_chunks}) release(h, *bf);
    return rc;
}

int fcfv_export_rhs(fcfv_t* h, double* fu, double* fp) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_export_rhs(h, fu, fp);
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_export_rhs: nothing assembled"));
    CU(cudaSetDevice(h->device));
    if (fu) TRY(copy(h, fu, h->fu.p, sizeof(double) * 2 * h->nf));
    if (fp) TRY(copy(h, fp, h->fp.p, sizeof(double) * h->nel));
    return sync(h);
}

int fcfv_device_csr(fcfv_t* h, int block, void** rowptr, void** colidx, void** val) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_device_csr (the rows live on several devices)");
    Csr* c = block_of(h, block);
    if (!c) return fail(h, FCFV_ERR_INVALID, "unknown block %d", block);
    TRY(need(h, c->nnz != -1, "fcfv_device_csr: block not assembled"));
    if (rowptr) *rowptr = c->rowptr.p;
    if (colidx) *colidx = c->col.p;
    if (val) *val = c->val.p;
    return FCFV_OK;
}

int fcfv_device_rhs(fcfv_t* h, void** fu, void** fp) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_device_rhs");
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_device_rhs: nothing assembled"));
    if (fu) *fu = h->fu.p;
    if (fp) *fp = h->fp.p;
    return FCFV_OK;
}

// -------------------------------------------------------------------------------------------------
int fcfv_run_residuals(fcfv_t* h, int formulation, int tau_mode, double* sumsq, double* norms) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_run_residuals(h, formulation, tau_mode, sumsq, norms);
    TRY(check_form(h, formulation));
    if (tau_mode != FCFV_TAU_REFERENCE && tau_mode != FCFV_TAU_FACE)
        return fail(h, FCFV_ERR_INVALID, "tau_mode must be FCFV_TAU_REFERENCE or FCFV_TAU_FACE (got %d)", tau_mode);
    TRY(need(h, h->have_mesh, "fcfv_run_residuals: no mesh set"));
    TRY(need(h, h->have_local, "fcfv_run_residuals: local coefficients missing"));
    TRY(need(h, h->have_sources && h->have_face_data && h->have_solution, "fcfv_run_residuals: sources / face data / solution not set"));
    if (tau_mode == FCFV_TAU_REFERENCE && !h->have_tau_ref && h->nf < h->nel)
        return fail(h, FCFV_ERR_INVALID, "FCFV_TAU_REFERENCE reads tau[e] for e <= nel but nf < nel");
    CU(cudaSetDevice(h->device));
    const long long nel = h->nel, nf = h->nf;
    // element blocks (which also form F_nodes of the faces inside their tile) + blocks of the late-face kernel (4 faces per thread)
    const int nbe = nblocks(nel), nbl = nblocks((nf + kLateFaces - 1) / kLateFaces, kLateBlock);
    const int stride = nbe + nbl;
    TRY(ensure(h, h->Fg, sizeof(double) * 4 * nf));        // [face][side][component], only cut faces are touched
    TRY(ensure(h, h->partial, sizeof(double) * 6 * stride));
    const MeshView M = mesh_view(h);
    double* part = P<double>(h->partial);
    if (formulation == FCFV_GRADIENT)
        TLAUNCH(h, T_RES_ELEM, (k_res_elem<kGradient, false>), nbe, kBlock, M, local_view(h), face_view(h), sol_view(h), P<double>(h->sex),
               P<double>(h->sey), tau_mode, P<double>(h->Fg), part, stride, (double*)nullptr, (double*)nullptr, (double*)nullptr);
    else
        TLAUNCH(h, T_RES_ELEM, (k_res_elem<kSymmetricGradient, false>), nbe, kBlock, M, local_view(h), face_view(h), sol_view(h), P<double>(h->sex),
               P<double>(h->sey), tau_mode, P<double>(h->Fg), part, stride, (double*)nullptr, (double*)nullptr, (double*)nullptr);
    TLAUNCH(h, T_RES_FACE, k_res_late, nbl, kLateBlock, M, P<double>(h->Fg), part, stride, nbe, (double*)nullptr, (double*)nullptr);
    // sumsq order = print order: F_eq1, F_eq2, F_eq3, F_nodes_x, F_nodes_y, F_glob2
    double* ss = P<double>(h->sumsq);
    TLAUNCH(h, T_REDUCE, k_reduce_residuals, dim3(kRedSlices, 6), 256, part, nbe, nbe + nbl, stride, P<double>(h->red_scratch),
            P<unsigned int>(h->red_tickets), ss);
    CU(cudaGetLastError());
    if (h->comm) TRY(fcfv_comm_allreduce_sum(h, ss, 6));
    if (sumsq || norms) {
        double hs[6];
        CU(cudaMemcpyAsync(hs, ss, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
        TRY(sync(h));
        if (sumsq) memcpy(sumsq, hs, sizeof hs);
        if (norms) {
            const double ne = (double)h->nel_global, nfg = (double)h->nf_global;
            const double len[6] = {4.0 * ne, 2.0 * ne, ne, nfg, nfg, ne};
            for (int k = 0; k < 6; k++) norms[k] = len[k] > 0 ? sqrt(hs[k]) / sqrt(len[k]) : 0.0;
        }
        return FCFV_OK;
    }
    return finish(h);
}

// One pass of the resident pipeline (local coefficients -> assembly -> residual sums) as a CUDA graph: for small
// meshes the pass is bound by launch latency (ten kernels of a few microseconds), a replayed graph removes it.
// The graph is rebuilt when its key changes: loop / formulation / tau mode / A, the mesh size, or any device
// buffer was reallocated (alloc_epoch); ownership / ghost flags and the anisotropy switch are re-checked through
// the direct pass that precedes every capture.  The NCCL all-reduce of a partitioned mesh is captured with the kernels;
// with per-kernel timing the pass runs directly.
int fcfv_run_pipeline(fcfv_t* h, int variant, int formulation, double A, int tau_mode) {
    if (!h) return FCFV_ERR_INVALID;
    auto direct = [&]() -> int {
        const bool was_async = h->async;
        h->async = true;
        int rc = run_local_impl(h, formulation, A, variant);
        if (!rc) { h->alpha.hm.epoch = h->beta.hm.epoch = h->Z.hm.epoch = h->tau.hm.epoch = -1; h->have_local = true; }
        if (!rc) rc = fcfv_run_assemble(h, variant, formulation, A, 0);
        if (!rc) rc = fcfv_run_residuals(h, formulation, tau_mode, nullptr, nullptr);
        h->async = was_async;
        return rc;
    };
    if (h->multi) { TRY(direct()); return multi_sync(h) ? FCFV_ERR_CUDA : FCFV_OK; }
    if (h->timing || h->graph_disabled) { TRY(direct()); return finish(h); }
    long long abits;
    memcpy(&abits, &A, sizeof abits);
    const long long flags = (h->have_elem_owned ? 1 : 0) | (h->have_face_owned ? 2 : 0) | (h->have_tau_ref ? 4 : 0) | (h->aniso ? 8 : 0);
    const long long key[7] = {variant, formulation, tau_mode + 16 * flags, h->alloc_epoch, h->nel, h->nf, abits};
    if (h->graph_exec && !h->delta_dirty && memcmp(key, h->graph_key, sizeof key) == 0) {
        CU(cudaSetDevice(h->device));
        CU(cudaGraphLaunch(h->graph_exec, h->stream));
        h->launches += h->graph_launches;
        h->coef_variant = h->aniso ? -1 : variant;
        h->alpha.hm.epoch = h->beta.hm.epoch = h->Z.hm.epoch = h->tau.hm.epoch = -1;
        h->have_local = true;
        h->Kuu.nnz = h->Kup.nnz = -2; h->Muu.nnz = -1; h->Kuusc.nnz = -1;
        h->totals_valid = false;
        return finish(h);
    }
    // direct pass first: validates the arguments, sizes every buffer and settles the lazy state
    TRY(direct());
    TRY(sync(h));
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    const long long flags2 = (h->have_elem_owned ? 1 : 0) | (h->have_face_owned ? 2 : 0) | (h->have_tau_ref ? 4 : 0) | (h->aniso ? 8 : 0);
    const long long key2[7] = {variant, formulation, tau_mode + 16 * flags2, h->alloc_epoch, h->nel, h->nf, abits};
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const long long launches0 = h->launches;
    const int rc = direct();
    h->graph_launches = h->launches - launches0;   // kernels one replay launches
    h->launches = launches0;                       // captured, not launched
    cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (h->comm) {      // a collective that cannot be captured (old NCCL): passes run directly from now on; this call's
            h->graph_disabled = true;      // result was produced by the direct pass above
            return finish(h);
        }
        if (rc) return rc;
        return fail(h, FCFV_ERR_CUDA, "fcfv_run_pipeline: stream capture failed: %s", cudaGetErrorString(ce));
    }
    ce = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
        h->graph_exec = nullptr;
        cudaGetLastError();
        if (h->comm) { h->graph_disabled = true; return finish(h); }
        return fail(h, FCFV_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    }
    memcpy(h->graph_key, key2, sizeof key2);
    return finish(h);   // the direct pass above already produced this call's result
}

int fcfv_residuals(fcfv_t* h, int formulation, int tau_mode, const double* Vxh, const double* Vyh, const double* Pe,
                   const double* alpha, const double* beta, const double* Z, const double* sex, const double* sey,
                   const double* VxDir, const double* VyDir, const double* SxxNeu, const double* SyyNeu,
                   const double* SxyNeu, const double* SyxNeu, double* norms, double* F_nodes_x, double* F_nodes_y,
                   double* F_glob2) {
    if (!h) return FCFV_ERR_INVALID;
    const bool was_async = h->async;
    h->async = true;
    int rc = fcfv_set_solution(h, Vxh, Vyh, Pe);
    if (!rc && (alpha || beta || Z)) rc = fcfv_set_local(h, alpha, beta, Z);
    if (!rc) rc = fcfv_set_sources(h, sex, sey);
    if (!rc) rc = fcfv_set_face_data(h, VxDir, VyDir, SxxNeu, SyyNeu, SxyNeu, SyxNeu);
    h->async = was_async;
    if (rc) return rc;
    double dummy[6];
    TRY(fcfv_run_residuals(h, formulation, tau_mode, nullptr, norms ? norms : dummy));
    if ((F_nodes_x || F_nodes_y || F_glob2) && h->multi) return multi_unsupported(h, "fcfv_residuals with F_nodes / F_glob2 outputs");
    if (F_nodes_x || F_nodes_y || F_glob2) {
        // optional arrays: rerun the two kernels with output pointers (device scratch), then copy out
        const long long nel = h->nel, nf = h->nf;
        TRY(ensure(h, h->tmp, sizeof(double) * (2 * nf + nel)));
        double* t = P<double>(h->tmp);
        const int nbe = nblocks(nel), nbl = nblocks((nf + kLateFaces - 1) / kLateFaces, kLateBlock), stride = nbe + nbl;
        const MeshView M = mesh_view(h);
        double* part = P<double>(h->partial);
        if (formulation == FCFV_GRADIENT)
            TLAUNCH(h, T_RES_ELEM, (k_res_elem<kGradient, true>), nbe, kBlock, M, local_view(h), face_view(h), sol_view(h), P<double>(h->sex),
                   P<double>(h->sey), tau_mode, P<double>(h->Fg), part, stride, t + 2 * nf, t, t + nf);
        else
            TLAUNCH(h, T_RES_ELEM, (k_res_elem<kSymmetricGradient, true>), nbe, kBlock, M, local_view(h), face_view(h), sol_view(h),
                   P<double>(h->sex), P<double>(h->sey), tau_mode, P<double>(h->Fg), part, stride, t + 2 * nf, t, t + nf);
        TLAUNCH(h, T_RES_FACE, k_res_late, nbl, kLateBlock, M, P<double>(h->Fg), part, stride, nbe, t, t + nf);
        CU(cudaGetLastError());
        if (F_nodes_x) TRY(copy(h, F_nodes_x, t, sizeof(double) * nf));
        if (F_nodes_y) TRY(copy(h, F_nodes_y, t + nf, sizeof(double) * nf));
        if (F_glob2) TRY(copy(h, F_glob2, t + 2 * nf, sizeof(double) * nel));
        TRY(sync(h));
    }
    return FCFV_OK;
}

// -------------------------------------------------------------------------------------------------
// The three calls of a driver (ComputeFCFV -> ElementAssemblyLoop[NEW] -> ComputeResidualsFCFV_Stokes_o1) as one:
// every host array crosses PCIe once (the separate calls upload VxDir, VyDir three times, the Neumann arrays and the
// sources twice) and the download of alpha / beta / Z / tau runs on a second stream while the remaining inputs are
// still going up (page-locked or registered destinations; pageable ones are staged and complete before the uploads
// continue).
int fcfv_pass(fcfv_t* h, int variant, int formulation, double A, int tau_mode, const double* sex, const double* sey,
              const double* VxDir, const double* VyDir, const double* sxxNeu, const double* syyNeu, const double* sxyNeu,
              const double* syxNeu, const double* Vxh, const double* Vyh, const double* Pe, double* alpha, double* beta,
              double* Z, double* tau, int want_muu, int64_t* nnz_uu, int64_t* nnz_up, int64_t* nnz_muu, double* norms) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_pass: no mesh set"));
    if (!VxDir || !VyDir || !Vxh || !Vyh || !Pe) return fail(h, FCFV_ERR_INVALID, "fcfv_pass: null input array");
    if (h->multi) {      // the three calls one after the other (each array is scattered once per call)
        TRY(fcfv_compute_local(h, formulation, A, sex, sey, VxDir, VyDir, alpha, beta, Z, tau));
        TRY(fcfv_assemble(h, variant, formulation, A, nullptr, nullptr, nullptr, VxDir, VyDir, sxxNeu, syyNeu, sxyNeu, syxNeu, want_muu, nnz_uu,
                          nnz_up, nnz_muu, nullptr));
        return fcfv_residuals(h, formulation, tau_mode, Vxh, Vyh, Pe, nullptr, nullptr, nullptr, sex, sey, VxDir, VyDir, sxxNeu, syyNeu, sxyNeu,
                              syxNeu, norms, nullptr, nullptr, nullptr);
    }
    TRY(check_form(h, formulation));
    CU(cudaSetDevice(h->device));
    const bool was_async = h->async;
    h->async = true;
    auto body = [&]() -> int {
        if (sex && sey && chunk_pipeline_ok(h, h->nel, {sex, sey, alpha, beta, Z, tau})) {
            // sources up | element kernel | alpha, beta, Z down, chunk by chunk (fcfv_compute_fcfv); the downloads then keep
            // running under the assembly and the upload of the solution
            const double* const none[4] = {nullptr, nullptr, nullptr, nullptr};
            h->pipelined_calls++;
            TRY(formulation == FCFV_GRADIENT
                    ? compute_fcfv_chunks<kGradient>(h, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, A, sex, sey, VxDir,
                                                     VyDir, none, alpha, beta, Z, tau)
                    : compute_fcfv_chunks<kSymmetricGradient>(h, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, A, sex,
                                                              sey, VxDir, VyDir, none, alpha, beta, Z, tau));
        } else {
            TRY(fcfv_set_sources(h, sex, sey));
            TRY(fcfv_set_face_data(h, VxDir, VyDir, nullptr, nullptr, nullptr, nullptr));
            TRY(fcfv_run_local(h, formulation, A));
            // outputs of the element kernel go down while the rest goes up
            TRY(download_local_async(h, alpha, beta, Z, tau));
        }
        TRY(fcfv_set_face_data(h, nullptr, nullptr, sxxNeu, syyNeu, sxyNeu, syxNeu));   // NULL = keep what is resident
        TRY(fcfv_run_assemble(h, variant, formulation, A, want_muu));
        TRY(fcfv_set_solution(h, Vxh, Vyh, Pe));
        return FCFV_OK;
    };
    int rc = body();
    h->async = was_async;
    if (rc == FCFV_OK) {
        double dummy[6];
        rc = fcfv_run_residuals(h, formulation, tau_mode, nullptr, norms ? norms : dummy);   // synchronises the main stream
    }
    if (rc == FCFV_OK) rc = fcfv_get_nnz(h, nnz_uu, nnz_up, nnz_muu);
    if (h->stream_out) {
        const int rc2 = local_downloaded(h, alpha, beta, Z, tau);
        if (rc == FCFV_OK) rc = rc2;
    }
    return rc;
}

// -------------------------------------------------------------------------------------------------
int fcfv_element_values(fcfv_t* h, const double* Vxh, const double* Vyh, const double* alpha, const double* beta,
                        const double* Z, double* Vxe, double* Vye, double* Txxe, double* Tyye, double* Txye) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_element_values(h, Vxh, Vyh, alpha, beta, Z, Vxe, Vye, Txxe, Tyye, Txye);
    TRY(need(h, h->have_mesh, "fcfv_element_values: no mesh set"));
    CU(cudaSetDevice(h->device));
    const long long nel = h->nel, nf = h->nf;
    if ((Vxh == nullptr) != (Vyh == nullptr)) return fail(h, FCFV_ERR_INVALID, "fcfv_element_values: pass both Vxh and Vyh or neither");
    if (Vxh) {
        TRY(upload(h, h->Vxh, Vxh, sizeof(double) * nf));
        TRY(upload(h, h->Vyh, Vyh, sizeof(double) * nf));
        if (!h->have_solution) { TRY(ensure(h, h->Pe, sizeof(double) * nel)); h->Pe.hm.epoch = -1; CU(cudaMemsetAsync(h->Pe.p, 0, sizeof(double) * nel, h->stream)); }
        h->have_solution = true;
    }
    if (alpha || beta || Z) {
        const bool was_async = h->async;
        h->async = true;
        const int rc = fcfv_set_local(h, alpha, beta, Z);
        h->async = was_async;
        TRY(rc);
    }
    TRY(need(h, h->have_solution, "fcfv_element_values: no face velocities (fcfv_set_solution)"));
    TRY(need(h, h->have_local, "fcfv_element_values: local coefficients missing"));
    TRY(ensure(h, h->tmp2, sizeof(double) * 5 * nel));
    double* out = P<double>(h->tmp2);
    LAUNCH(h, k_elem_values, nblocks(nel), kBlock, mesh_view(h), local_view(h), sol_view(h), out);
    CU(cudaGetLastError());
    double* dst[5] = {Vxe, Vye, Txxe, Tyye, Txye};
    for (int k = 0; k < 5; k++)
        if (dst[k]) TRY(copy(h, dst[k], out + (size_t)k * nel, sizeof(double) * nel));
    return sync(h);
}

// -------------------------------------------------------------------------------------------------
namespace {
int ph_inputs(fcfv_t* h, const double* u, const double* p, double** du, double** dp) {
    // u (2nf) and p (nel) staged in tmp: [u | p]
    TRY(ensure(h, h->tmp, sizeof(double) * (2 * h->nf + h->nel)));
    double* t = P<double>(h->tmp);
    if (u) TRY(copy(h, t, u, sizeof(double) * 2 * h->nf));
    if (p) TRY(copy(h, t + 2 * h->nf, p, sizeof(double) * h->nel));
    *du = t; *dp = t + 2 * h->nf;
    return FCFV_OK;
}
}  // namespace

int fcfv_ph_residual(fcfv_t* h, const double* u, const double* p, double* ru, double* rp, double* nrm) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_ph_residual(h, u, p, ru, rp, nrm);
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_ph_residual: nothing assembled"));
    if (!u || !p) return fail(h, FCFV_ERR_INVALID, "fcfv_ph_residual: null u or p");
    CU(cudaSetDevice(h->device));
    const long long nel = h->nel, nf = h->nf;
    double *du, *dp;
    TRY(ph_inputs(h, u, p, &du, &dp));
    TRY(ensure(h, h->tmp2, sizeof(double) * (2 * nf + nel)));
    double* dru = P<double>(h->tmp2);
    double* drp = dru + 2 * nf;
    const int nbr = nblocks(2 * nf), nbe = nblocks(nel), stride = std::max(nbr, nbe);
    TRY(ensure(h, h->partial, sizeof(double) * 6 * std::max<long long>(stride, nblocks(nf))));
    double* part = P<double>(h->partial);
    const MeshView M = mesh_view(h);
    if (h->rp64)
        LAUNCH(h, k_ph_ru<long long>, nbr, kBlock, (int)(2 * nf), P<long long>(h->Kuu.rowptr), P<int>(h->Kuu.col), P<double>(h->Kuu.val),
               P<long long>(h->Kup.rowptr), P<int>(h->Kup.col), P<double>(h->Kup.val), P<double>(h->fu), du, dp, dru, part);
    else
        LAUNCH(h, k_ph_ru<int>, nbr, kBlock, (int)(2 * nf), P<int>(h->Kuu.rowptr), P<int>(h->Kuu.col), P<double>(h->Kuu.val),
               P<int>(h->Kup.rowptr), P<int>(h->Kup.col), P<double>(h->Kup.val), P<double>(h->fu), du, dp, dru, part);
    LAUNCH(h, k_ph_rp, nbe, kBlock, M, P<double>(h->fp), du, 0, 0.0, drp, (double*)nullptr, part + stride);
    double* ss = P<double>(h->sumsq);
    TLAUNCH(h, T_REDUCE, k_reduce_partials, 1, 1024, part, nbr, stride, ss + 6);
    TLAUNCH(h, T_REDUCE, k_reduce_partials, 1, 1024, part + stride, nbe, stride, ss + 7);
    CU(cudaGetLastError());
    if (h->comm) TRY(fcfv_comm_allreduce_sum(h, ss + 6, 2));
    double hs[2];
    CU(cudaMemcpyAsync(hs, ss + 6, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
    if (ru) TRY(copy(h, ru, dru, sizeof(double) * 2 * nf));
    if (rp) TRY(copy(h, rp, drp, sizeof(double) * nel));
    TRY(sync(h));
    if (nrm) { nrm[0] = sqrt(hs[0]); nrm[1] = sqrt(hs[1]); }
    return FCFV_OK;
}

int fcfv_ph_update_p(fcfv_t* h, double gamma, const double* u, double* p) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_ph_update_p(h, gamma, u, p);
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_ph_update_p: nothing assembled"));
    if (!u || !p) return fail(h, FCFV_ERR_INVALID, "fcfv_ph_update_p: null u or p");
    CU(cudaSetDevice(h->device));
    double *du, *dp;
    TRY(ph_inputs(h, u, p, &du, &dp));
    const MeshView M = mesh_view(h);
    LAUNCH(h, k_ph_rp, nblocks(h->nel), kBlock, M, P<double>(h->fp), du, 1, gamma, (double*)nullptr, dp, (double*)nullptr);
    CU(cudaGetLastError());
    TRY(copy(h, p, dp, sizeof(double) * h->nel));
    return sync(h);
}

int fcfv_center_pressure(fcfv_t* h, double* Pe, double* mean) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_center_pressure(h, Pe, mean);
    TRY(need(h, h->have_mesh, "fcfv_center_pressure: no mesh set"));
    if (!Pe) return fail(h, FCFV_ERR_INVALID, "fcfv_center_pressure: null Pe");
    CU(cudaSetDevice(h->device));
    const long long nel = h->nel;
    const int nbe = nblocks(nel);
    TRY(ensure(h, h->tmp, sizeof(double) * nel));
    TRY(ensure(h, h->partial, sizeof(double) * 6 * std::max<long long>(nbe, nblocks(h->nf))));
    double* dp = P<double>(h->tmp);
    double* ss = P<double>(h->sumsq);
    TRY(copy(h, dp, Pe, sizeof(double) * nel));
    LAUNCH(h, k_sum_owned, nbe, kBlock, (int)nel, dp, P<int>(h->ebc), P<double>(h->partial));
    LAUNCH(h, k_reduce_partials, 1, 1024, P<double>(h->partial), nbe, nbe, ss + 7);
    CU(cudaGetLastError());
    if (h->comm) TRY(fcfv_comm_allreduce_sum(h, ss + 7, 1));
    LAUNCH(h, k_sub_mean, nblocks(nel, 256), 256, (int)nel, dp, ss + 7, (double)h->nel_global);
    CU(cudaGetLastError());
    double hs = 0.0;
    CU(cudaMemcpyAsync(&hs, ss + 7, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
    TRY(copy(h, Pe, dp, sizeof(double) * nel));
    TRY(sync(h));
    if (mean) *mean = hs / (double)h->nel_global;
    return FCFV_OK;
}

int fcfv_ph_fusc(fcfv_t* h, double gamma, const double* p, double* fusc) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_ph_fusc(h, gamma, p, fusc);
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_ph_fusc: nothing assembled"));
    if (!p || !fusc) return fail(h, FCFV_ERR_INVALID, "fcfv_ph_fusc: null p or fusc");
    CU(cudaSetDevice(h->device));
    const long long nel = h->nel, nf = h->nf;
    double *du, *dp;
    TRY(ph_inputs(h, nullptr, p, &du, &dp));
    TRY(ensure(h, h->phw, sizeof(double) * nel));
    TRY(ensure(h, h->tmp2, sizeof(double) * (2 * nf + nel)));
    double* out = P<double>(h->tmp2);
    LAUNCH(h, k_ph_w, nblocks(nel, 256), 256, (int)nel, P<double>(h->ke), P<double>(h->Om), P<double>(h->fp), dp, gamma, P<double>(h->phw));
    if (h->rp64)
        LAUNCH(h, k_ph_fusc<long long>, nblocks(2 * nf, 256), 256, (int)(2 * nf), P<long long>(h->Kup.rowptr), P<int>(h->Kup.col),
               P<double>(h->Kup.val), P<double>(h->fu), P<double>(h->phw), out);
    else
        LAUNCH(h, k_ph_fusc<int>, nblocks(2 * nf, 256), 256, (int)(2 * nf), P<int>(h->Kup.rowptr), P<int>(h->Kup.col),
               P<double>(h->Kup.val), P<double>(h->fu), P<double>(h->phw), out);
    CU(cudaGetLastError());
    TRY(copy(h, fusc, out, sizeof(double) * 2 * nf));
    return sync(h);
}

// -------------------------------------------------------------------------------------------------
int fcfv_has_orphan_faces(fcfv_t* h, int* flag) {
    if (!h || !flag) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_has_orphan_faces: no mesh set"));
    *flag = h->multi ? multi_has_orphans(h) : (h->has_orphans ? 1 : 0);
    return FCFV_OK;
}

int fcfv_mesh_size(fcfv_t* h, int64_t* nel, int64_t* nf) {
    if (!h) return FCFV_ERR_INVALID;
    TRY(need(h, h->have_mesh, "fcfv_mesh_size: no mesh set"));
    if (nel) *nel = h->nel;
    if (nf) *nf = h->nf;
    return FCFV_OK;
}

int fcfv_get_mesh(fcfv_t* h, int64_t* e2f, int64_t* bc, double* Omega, double* nx, double* ny, double* Gamma,
                  double* ke, double* delta, double* taue, double* tau) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_get_mesh");
    TRY(need(h, h->have_mesh, "fcfv_get_mesh: no mesh set"));
    CU(cudaSetDevice(h->device));
    const long long nel = h->nel, nf = h->nf;
    if (e2f) {
        TRY(ensure(h, h->tmp, sizeof(int64_t) * 3 * nel));
        LAUNCH(h, k_widen_e2f, nblocks(3 * nel, 256), 256, P<int>(h->e2f), P<long long>(h->tmp), (long long)(3 * nel));
        TRY(copy(h, e2f, h->tmp.p, sizeof(int64_t) * 3 * nel));
        TRY(sync(h));
    }
    if (bc) {
        TRY(ensure(h, h->tmp, sizeof(int64_t) * nf));
        LAUNCH(h, k_widen_bc, nblocks(nf, 256), 256, P<int8_t>(h->bc), P<long long>(h->tmp), (long long)nf);
        TRY(copy(h, bc, h->tmp.p, sizeof(int64_t) * nf));
        TRY(sync(h));
    }
    if (Omega) TRY(copy(h, Omega, h->Om.p, sizeof(double) * nel));
    if (nx) TRY(copy(h, nx, h->nx.p, sizeof(double) * 3 * nel));
    if (ny) TRY(copy(h, ny, h->ny.p, sizeof(double) * 3 * nel));
    if (Gamma) TRY(copy(h, Gamma, h->G.p, sizeof(double) * 3 * nel));
    if (ke) TRY(copy(h, ke, h->ke.p, sizeof(double) * nel));
    if (delta) TRY(copy(h, delta, h->dl.p, sizeof(double) * nel));
    if (taue) TRY(copy(h, taue, h->taue.p, sizeof(double) * nel));
    if (tau) TRY(copy(h, tau, h->tau.p, sizeof(double) * nf));
    return sync(h);
}

int fcfv_numbering_locality(fcfv_t* h, double* element_span, double* face_offset) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_numbering_locality");
    TRY(need(h, h->have_mesh, "fcfv_numbering_locality: no mesh set"));
    CU(cudaSetDevice(h->device));
    const int nblk = nblocks(h->nf);
    DBuf part;
    std::vector<double> host(3 * (size_t)nblk);
    auto body = [&]() -> int {
        TRY(ensure(h, part, sizeof(double) * 3 * nblk));
        LAUNCH(h, k_locality, nblk, kBlock, P<int2>(h->f2e), (int)h->nf, (int)std::max<long long>(h->nel, 1), P<double>(part), nblk);
        TRY(copy(h, host.data(), part.p, sizeof(double) * 3 * nblk));
        return sync(h);
    };
    const int rc = body();
    cudaStreamSynchronize(h->stream);
    release(h, part);
    if (rc != FCFV_OK) return rc;
    double s[3] = {0.0, 0.0, 0.0};
    for (int q = 0; q < 3; q++)
        for (int b = 0; b < nblk; b++) s[q] += host[(size_t)q * nblk + b];
    if (element_span) *element_span = s[2] > 0.0 ? s[0] / s[2] : 0.0;
    if (face_offset) *face_offset = s[2] > 0.0 ? s[1] / s[2] : 0.0;
    return FCFV_OK;
}

int fcfv_get_ownership(fcfv_t* h, uint8_t* elem_owned, uint8_t* face_owned, int64_t* nel_global, int64_t* nf_global) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_get_ownership");
    TRY(need(h, h->have_mesh, "fcfv_get_ownership: no mesh set"));
    CU(cudaSetDevice(h->device));
    if (elem_owned) {
        if (h->have_elem_owned) TRY(copy(h, elem_owned, h->elem_owned.p, (size_t)h->nel));
        else memset(elem_owned, 1, (size_t)h->nel);
    }
    if (face_owned) {
        if (h->have_face_owned) TRY(copy(h, face_owned, h->face_owned.p, (size_t)h->nf));
        else memset(face_owned, 1, (size_t)h->nf);
    }
    if (nel_global) *nel_global = h->nel_global;
    if (nf_global) *nf_global = h->nf_global;
    return sync(h);
}

int fcfv_get_data(fcfv_t* h, double* sex, double* sey, double* VxDir, double* VyDir, double* sxxNeu, double* syyNeu,
                  double* sxyNeu, double* syxNeu, double* Vxh, double* Vyh, double* Pe) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_get_data");
    TRY(need(h, h->have_mesh && h->have_sources && h->have_face_data && h->have_solution, "fcfv_get_data: data not set"));
    CU(cudaSetDevice(h->device));
    const size_t be = sizeof(double) * h->nel, bf = sizeof(double) * h->nf;
    if (sex) TRY(copy(h, sex, h->sex.p, be));
    if (sey) TRY(copy(h, sey, h->sey.p, be));
    if (VxDir) TRY(copy(h, VxDir, h->VxDir.p, bf));
    if (VyDir) TRY(copy(h, VyDir, h->VyDir.p, bf));
    if (sxxNeu) TRY(copy(h, sxxNeu, h->sxx.p, bf));
    if (syyNeu) TRY(copy(h, syyNeu, h->syy.p, bf));
    if (sxyNeu) TRY(copy(h, sxyNeu, h->sxy.p, bf));
    if (syxNeu) TRY(copy(h, syxNeu, h->syx.p, bf));
    if (Vxh) TRY(copy(h, Vxh, h->Vxh.p, bf));
    if (Vyh) TRY(copy(h, Vyh, h->Vyh.p, bf));
    if (Pe) TRY(copy(h, Pe, h->Pe.p, be));
    return sync(h);
}

// -------------------------------------------------------------------------------------------------
int fcfv_generate_structured(fcfv_t* h, int64_t n, int64_t row0, int64_t row1, int ghost_below, int setting,
                             double tau_r, double eta_incl, int neumann_south) {
    if (!h) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_generate_structured");
    CU(cudaSetDevice(h->device));
    if (n < 1 || row0 < 0 || row1 > n || row0 >= row1) return fail(h, FCFV_ERR_INVALID, "fcfv_generate_structured: bad row range");
    if (setting != 0 && setting != 1) return fail(h, FCFV_ERR_INVALID, "fcfv_generate_structured: setting must be 0 or 1");
    (void)ghost_below;   // ghost rows are added automatically where the strip has neighbours
    const StripLayout S = make_strip_layout(n, row0, row1);
    if (S.nel >= (1LL << 29) || S.nf >= (1LL << 30)) return fail(h, FCFV_ERR_INVALID, "fcfv_generate_structured: strip too large");
    h->have_mesh = false; h->have_local = false;
    h->delta_dirty = true;
    h->Kuu.nnz = h->Kup.nnz = h->Muu.nnz = h->Kuusc.nnz = -1;
    const long long nel = S.nel, nf = S.nf;
    h->nel = nel; h->nf = nf;
    h->nel_global = 4 * n * n; h->nf_global = 6 * n * n + 2 * n;
    TRY(ensure(h, h->e2f, sizeof(int) * 3 * nel));
    TRY(ensure(h, h->bc, (size_t)nf));
    new_epoch(h);        // every data buffer is rewritten on the device
    h->coef_variant = -1;
    for (DBuf* b : {&h->Om, &h->ke, &h->dl, &h->taue, &h->sex, &h->sey, &h->Pe}) TRY(ensure(h, *b, sizeof(double) * nel));
    for (DBuf* b : {&h->nx, &h->ny, &h->G}) TRY(ensure(h, *b, sizeof(double) * 3 * nel));
    for (DBuf* b : {&h->VxDir, &h->VyDir, &h->sxx, &h->syy, &h->sxy, &h->syx, &h->Vxh, &h->Vyh}) TRY(ensure(h, *b, sizeof(double) * nf));
    TRY(ensure(h, h->elem_owned, (size_t)nel));
    TRY(ensure(h, h->face_owned, (size_t)nf));
    TRY(alloc_mesh_dependent(h));
    CU(cudaMemsetAsync(h->tau.p, 0, sizeof(double) * nf, h->stream));
    GenOut G;
    G.e2f = P<int>(h->e2f); G.bc = P<int8_t>(h->bc);
    G.Om = P<double>(h->Om); G.nx = P<double>(h->nx); G.ny = P<double>(h->ny); G.G = P<double>(h->G);
    G.ke = P<double>(h->ke); G.dl = P<double>(h->dl); G.taue = P<double>(h->taue);
    G.sex = P<double>(h->sex); G.sey = P<double>(h->sey); G.Pe = P<double>(h->Pe);
    G.VxDir = P<double>(h->VxDir); G.VyDir = P<double>(h->VyDir); G.sxx = P<double>(h->sxx); G.syy = P<double>(h->syy);
    G.sxy = P<double>(h->sxy); G.syx = P<double>(h->syx); G.Vxh = P<double>(h->Vxh); G.Vyh = P<double>(h->Vyh);
    G.elem_owned = P<uint8_t>(h->elem_owned); G.face_owned = P<uint8_t>(h->face_owned);
    LAUNCH(h, k_gen_elements, nblocks(nel), kBlock, S, G, setting, tau_r, eta_incl);
    LAUNCH(h, k_gen_faces, nblocks(S.nquads), kBlock, S, G, setting, neumann_south);
    CU(cudaGetLastError());
    const bool partitioned = S.rlo != 0 || S.rhi != n;
    h->have_elem_owned = h->have_face_owned = partitioned;
    h->have_tau_ref = false;
    CU(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
    TRY(build_f2e(h));
    if (setting == 1) {   // interface faces: the two adjacent elements have different viscosity
        LAUNCH(h, k_gen_interface, nblocks(nf, 256), 256, P<int2>(h->f2e), P<double>(h->ke), P<int8_t>(h->bc), (int)nf);
        CU(cudaGetLastError());
    }
    int flag = 0;
    TRY(read_err_flag(h, &flag));
    h->has_orphans = (flag & 8) != 0;
    flag &= 7;
    if (flag) return fail(h, FCFV_ERR_STATE, "fcfv_generate_structured: generated connectivity failed validation (%d)", flag);
    TRY(build_fconn(h));
    TRY(build_face_lists(h));
    if (partitioned) {
        LAUNCH(h, k_pack_owned, nblocks(nel, 256), 256, (int)nel, P<uint8_t>(h->elem_owned), kOwnedElem, P<int>(h->ebc));
        LAUNCH(h, k_pack_owned, nblocks(nf, 256), 256, (int)nf, P<uint8_t>(h->face_owned), kOwnedFace, P<int>(h->fbc));
        LAUNCH(h, k_pack_face_owned, nblocks(nel, 256), 256, (int)nel, P<int>(h->e2f), P<uint8_t>(h->face_owned), P<int>(h->eadj));
        CU(cudaGetLastError());
    }
    TRY(sync(h));
    h->have_mesh = h->have_sources = h->have_face_data = h->have_solution = true;
    h->delta_dirty = false; h->aniso = false;   // the generator sets delta = 1 everywhere
    return FCFV_OK;
}

// -------------------------------------------------------------------------------------------------
namespace {
int load_nccl(fcfv_t* h) {
    if (h->nccl.lib) return FCFV_OK;
    const char* env = getenv("FCFV_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nme : names) {
        if (!nme) continue;
        lib = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(h, FCFV_ERR_NCCL, "cannot load libnccl.so.2 (set FCFV_NCCL_LIB): %s", dlerror());
    h->nccl.lib = lib;
    h->nccl.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    h->nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(lib, "ncclCommInitRank");
    h->nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    h->nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    h->nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!h->nccl.GetUniqueId || !h->nccl.CommInitRank || !h->nccl.AllReduce || !h->nccl.CommDestroy)
        return fail(h, FCFV_ERR_NCCL, "libnccl is missing a required symbol");
    return FCFV_OK;
}
int nccl_fail(fcfv_t* h, const char* what, int rc) {
    return fail(h, FCFV_ERR_NCCL, "%s failed: %s", what, h->nccl.GetErrorString ? h->nccl.GetErrorString(rc) : "?");
}
}  // namespace

int fcfv_comm_unique_id(fcfv_t* h, void* id128) {
    if (!h || !id128) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_comm_unique_id (the handle owns its communicators)");
    TRY(load_nccl(h));
    const int rc = h->nccl.GetUniqueId(id128);
    if (rc) return nccl_fail(h, "ncclGetUniqueId", rc);
    return FCFV_OK;
}

int fcfv_comm_init(fcfv_t* h, const void* id128, int rank, int nranks) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_comm_init");
    CU(cudaSetDevice(h->device));
    TRY(load_nccl(h));
    Id128 id;
    memcpy(id.b, id128, 128);
    const int rc = h->nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (rc) { h->comm = nullptr; return nccl_fail(h, "ncclCommInitRank", rc); }
    h->rank = rank; h->nranks = nranks;
    return FCFV_OK;
}

int fcfv_comm_allreduce_sum(fcfv_t* h, double* data, int n) {
    if (!h || !data || n < 0) return FCFV_ERR_INVALID;
    if (h->multi) return multi_unsupported(h, "fcfv_comm_allreduce_sum");
    if (!h->comm) return FCFV_OK;   // single rank: identity
    CU(cudaSetDevice(h->device));
    cudaPointerAttributes at;
    const bool on_device = cudaPointerGetAttributes(&at, data) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    double* d = data;
    if (!on_device) {
        if (n > 8) return fail(h, FCFV_ERR_INVALID, "host all-reduce supports at most 8 values");
        TRY(ensure(h, h->phw, sizeof(double) * 8));
        d = P<double>(h->phw);
        TRY(copy(h, d, data, sizeof(double) * n));
    }
    const int rc = h->nccl.AllReduce(d, d, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, h->comm, h->stream);
    if (rc) return nccl_fail(h, "ncclAllReduce", rc);
    if (!on_device) {
        TRY(copy(h, data, d, sizeof(double) * n));
        return sync(h);
    }
    return FCFV_OK;
}

// =================================================================================================
#include "fcfv_multi.inl"
