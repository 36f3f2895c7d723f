// fcfv_multi.inl -- several GPUs behind ONE handle (included at the end of fcfv_api.cu).
//
// The reference's drivers are single-process scripts (examples/ViscousInclusionFCFV_v2.jl:60-154,
// examples/SolKzFCFV.jl:50-215), so the multi-GPU path has to live inside the handle: fcfv_create_multi(h, ngpus)
// returns a handle whose entry points take and return the GLOBAL arrays of the unpartitioned mesh, exactly like a
// single-GPU handle.  Inside, the mesh is split by element (C++ partitioner below: contiguous element ranges, face
// owner = part of the highest adjacent element -- the element whose tau_e the reference's last-writer rule puts
// into mesh.tau, DiscretisationFCFV_Stokes.jl:40 --, one layer of ghost elements replicated), every part is an
// ordinary single-GPU handle on its own device and stream, and each part assembles the rows of the faces it owns
// with NO data-path communication.  Local ids ascend with the global ids, so column order, summation order and
// therefore every bit of an owned row are those of the unpartitioned run.  The only exchange is the all-reduce of
// the six residual sums of squares: one NCCL communicator per device (ncclCommInitAll), one grouped
// ncclAllReduce per evaluation.  Host work (partitioning, gather / scatter of the caller's arrays) runs on one
// thread per part.

namespace {

struct Part {
    fcfv_t* h = nullptr;
    std::vector<int64_t> elems, faces;      // local -> global (0-based), ascending
    std::vector<int64_t> own_e, own_f;      // local ids of the owned elements / faces, ascending
    std::vector<uint8_t> eown, fown;
    std::vector<std::pair<int64_t, int64_t>> erun, frun;   // (first local id, length) of every run of consecutive global ids
    std::vector<std::pair<int64_t, int64_t>> oerun, ofrun; // the same for the OWNED elements / faces only
};

}  // namespace

struct fcfv_multi {
    std::vector<Part> parts;
    std::vector<void*> comms;
    NcclApi nccl;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    bool tau_ref_stale = true;       // tau changed on the devices since tau_ref was last set (FCFV_TAU_REFERENCE)
    double asm_seconds = 0.0;
    double partition_seconds = 0.0;
    // faces no element references belong to no part; their tau is whatever the caller uploaded last (0 after set_mesh)
    std::vector<int64_t> unref;
    std::vector<double> unref_tau;
};

namespace {

// fn(p, part) on one thread per part; the first failure wins and its message is copied to the top handle
template <typename F>
int for_parts(fcfv_t* h, F&& fn) {
    fcfv_multi& M = *h->multi;
    const int n = (int)M.parts.size();
    std::vector<int> rc(n, FCFV_OK);
    std::vector<std::thread> th;
    th.reserve(n);
    for (int p = 0; p < n; p++) {
        auto work = [&, p]() {
            cudaSetDevice(M.parts[p].h->device);
            rc[p] = fn(p, M.parts[p]);
        };
        try {
            th.emplace_back(work);
        } catch (...) {
            work();
        }
    }
    for (auto& t : th)
        if (t.joinable()) t.join();
    for (int p = 0; p < n; p++)
        if (rc[p] != FCFV_OK) {
            h->err = "part " + std::to_string(p) + ": " + M.parts[p].h->err;
            return rc[p];
        }
    return FCFV_OK;
}

int multi_sync(fcfv_t* h) {
    for (Part& P : h->multi->parts) {
        if (cudaSetDevice(P.h->device) != cudaSuccess || cudaStreamSynchronize(P.h->stream) != cudaSuccess)
            return fail(h, FCFV_ERR_CUDA, "synchronising device %d failed", P.h->device);
    }
    return FCFV_OK;
}

int multi_finish(fcfv_t* h) { return h->async ? FCFV_OK : multi_sync(h); }

// out[c * n + i] = g[c * ld + idx[i]]  (columns of a column-major global array restricted to local ids)
// temporaries of the gather / scatter: a vector that does not zero its storage (every entry is written before it is read)
template <typename T>
struct no_init_alloc : std::allocator<T> {
    template <typename U> struct rebind { using other = no_init_alloc<U>; };
    template <typename U, typename... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
template <typename T> using uvec = std::vector<T, no_init_alloc<T>>;

// out[c * n + i] = g[c * ld + idx[i]].  Local ids ascend with the global ids and a part is a few contiguous id ranges (its
// own range plus ghost layers), so the gather runs as one memcpy per run of consecutive ids.
template <typename T, typename V>
void take(const T* g, int64_t ld, const std::vector<int64_t>& idx, int ncol, V& out) {
    const size_t n = idx.size();
    out.resize(n * (size_t)ncol);
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && idx[j] == idx[j - 1] + 1) j++;
        for (int c = 0; c < ncol; c++) memcpy(out.data() + (size_t)c * n + i, g + (size_t)c * ld + idx[i], (j - i) * sizeof(T));
        i = j;
    }
}

void make_runs(const std::vector<int64_t>& idx, std::vector<std::pair<int64_t, int64_t>>& runs) {
    runs.clear();
    const size_t n = idx.size();
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && idx[j] == idx[j - 1] + 1) j++;
        runs.emplace_back((int64_t)i, (int64_t)(j - i));
        i = j;
    }
}

// A part of a range partition is a handful of runs (its own id range plus ghost layers): its restriction of a global array
// then goes straight from the caller's memory into the part's device buffer, one copy per run and column -- no gather into
// a temporary, no extra pass over host memory (a copy costs a few microseconds).  Beyond this many runs (scattered ghosts of an unordered mesh) the gather
// into a temporary is cheaper than that many small copies.
size_t max_direct_runs() {      // FCFV_MULTI_MAX_RUNS overrides (0 forces the gather path: test hook)
    const char* e = getenv("FCFV_MULTI_MAX_RUNS");
    return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)16384;
}

template <typename T>
int upload_runs(fcfv_t* q, DBuf& b, const T* g, int64_t ld, const std::vector<int64_t>& idx,
                const std::vector<std::pair<int64_t, int64_t>>& runs, int ncol) {
    const size_t n = idx.size();
    TRY(ensure(q, b, sizeof(T) * n * (size_t)ncol));
    b.hm.epoch = -1;
    fcfv_t* h = q;      // for the CU / TRY macros
    for (const auto& r : runs)
        for (int c = 0; c < ncol; c++)
            TRY(copy(h, static_cast<T*>(b.p) + (size_t)c * n + r.first, g + (size_t)c * ld + idx[(size_t)r.first], (size_t)r.second * sizeof(T)));
    return FCFV_OK;
}

// owned entries of a part's device array straight into the caller's global array, one copy per run and column
template <typename T>
int download_runs(fcfv_t* q, const T* dev, T* g, int64_t ld, const std::vector<int64_t>& idx,
                  const std::vector<std::pair<int64_t, int64_t>>& runs, int ncol) {
    const size_t n = idx.size();
    fcfv_t* h = q;
    for (const auto& r : runs)
        for (int c = 0; c < ncol; c++)
            TRY(copy(h, g + (size_t)c * ld + idx[(size_t)r.first], dev + (size_t)c * n + r.first, (size_t)r.second * sizeof(T)));
    return FCFV_OK;
}

// g[c * ld + idx[own[k]]] = loc[c * n + own[k]]  (owned entries of a local array back into the global one), by runs
template <typename T>
void put_owned(T* g, int64_t ld, const std::vector<int64_t>& idx, const std::vector<int64_t>& own, int ncol, const T* loc) {
    const size_t n = idx.size(), m = own.size();
    for (size_t a = 0; a < m;) {
        size_t b = a + 1;
        while (b < m && own[b] == own[b - 1] + 1 && idx[own[b]] == idx[own[b - 1]] + 1) b++;
        for (int c = 0; c < ncol; c++) memcpy(g + (size_t)c * ld + idx[own[a]], loc + (size_t)c * n + own[a], (b - a) * sizeof(T));
        a = b;
    }
}

// ---- the partitioner (C++ port of what one rank of a torchrun job does in fcfv_nme23_b200/partition.py) ------------
// lo / hi: lowest / highest adjacent element of every face (-1: none).  Returns false on an out-of-range face id.
// 32-bit ids (a multi-GPU handle takes nel < 2^31) in storage that is not zeroed on allocation: the worker threads that
// fill it with -1 are the first to touch it, so the page faults of the two nf-sized arrays run in parallel.
// The walks themselves are plain loops: in ascending element order the first visitor of a face is its lowest element, in
// descending order its highest -- two threads, no atomics (a compare-and-swap per visit on many threads was measured 8x
// slower than one plain walk: the locked operations serialise the memory pipeline).
using AdjVec = uvec<int32_t>;
bool face_adjacency(int64_t nel, int64_t nf, const int64_t* e2f, AdjVec& lo, AdjVec& hi) {
    lo.resize((size_t)nf);
    hi.resize((size_t)nf);
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)(hw ? hw : 1u), 16, nf / 262144 + 1}));
    auto fill = [&](int t) {
        const int64_t f0 = nf * t / nt, f1 = nf * (t + 1) / nt;
        for (int64_t f = f0; f < f1; f++) { lo[f] = -1; hi[f] = -1; }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) {
            try { th.emplace_back(fill, t); } catch (...) { fill(t); }
        }
        fill(0);
        for (auto& x : th) x.join();
    }
    bool bad_lo = false, bad_hi = false;
    auto walk_lo = [&]() {
        for (int64_t e = 0; e < nel; e++)
            for (int i = 0; i < 3; i++) {
                const int64_t f = e2f[(size_t)i * nel + e] - 1;
                if (f < 0 || f >= nf) { bad_lo = true; return; }
                if (lo[f] < 0) lo[f] = (int32_t)e;
            }
    };
    auto walk_hi = [&]() {
        for (int64_t e = nel - 1; e >= 0; e--)
            for (int i = 0; i < 3; i++) {
                const int64_t f = e2f[(size_t)i * nel + e] - 1;
                if (f < 0 || f >= nf) { bad_hi = true; return; }
                if (hi[f] < 0) hi[f] = (int32_t)e;
            }
    };
    try {
        std::thread other(walk_hi);
        walk_lo();
        other.join();
    } catch (...) {      // no second thread to be had
        walk_lo();
        walk_hi();
    }
    return !(bad_lo || bad_hi);
}

// ascending ids with a mark in [lo, hi]
void collect_marked(const std::vector<uint8_t>& mark, int64_t lo, int64_t hi, std::vector<int64_t>& out) {
    size_t n = 0;
    for (int64_t k = lo; k <= hi; k++) n += mark[k];
    out.clear();
    out.reserve(n);
    for (int64_t k = lo; k <= hi; k++)
        if (mark[k]) out.push_back(k);
}

// Part `p` of `np`: owned elements [e0, e1); (A) the lower neighbours of its owned faces; (B) the tau donors: the
// highest neighbour of every face of an owned / A element, so that the local last-writer rule reproduces the
// global mesh.tau on every column the owned rows touch.  Sets are kept as byte marks over the id range they span
// (no sorting: the ids come out ascending from a scan).
void build_part(int64_t nel, int64_t nf, const int64_t* e2f, const AdjVec& lo, const AdjVec& hi,
                int p, int np, Part& P) {
    const int64_t e0 = (int64_t)p * nel / np, e1 = (int64_t)(p + 1) * nel / np;
    auto in_range = [&](int64_t e) { return e >= e0 && e < e1; };
    std::vector<uint8_t> em((size_t)nel, 0);
    int64_t emin = e0, emax = e1 - 1;
    std::vector<int64_t> ghostA;
    // owned faces are found among the faces of the owned elements (hi[f] in range implies f is a face of hi[f])
    for (int64_t e = e0; e < e1; e++) {
        em[e] = 1;
        for (int i = 0; i < 3; i++) {
            const int64_t f = e2f[(size_t)i * nel + e] - 1;
            const int64_t l = lo[f];
            if (hi[f] == e && !in_range(l) && !em[l]) { em[l] = 1; ghostA.push_back(l); }       // (A)
        }
    }
    auto donors = [&](int64_t e) {
        for (int i = 0; i < 3; i++) {
            const int64_t d = hi[e2f[(size_t)i * nel + e] - 1];
            if (!in_range(d)) em[d] |= 2;                                    // (B)
        }
    };
    for (int64_t e = e0; e < e1; e++) donors(e);
    for (int64_t e : ghostA) donors(e);
    for (int64_t e : ghostA) { emin = std::min(emin, e); emax = std::max(emax, e); }
    // B elements can lie anywhere: find their span while collecting
    if (e1 > e0) {
        for (int64_t e = 0; e < e0; e++) if (em[e]) { emin = std::min(emin, e); break; }
        for (int64_t e = nel - 1; e >= e1; e--) if (em[e]) { emax = std::max(emax, e); break; }
    } else {
        emin = 0; emax = nel - 1;
    }
    std::vector<int64_t> elems;
    if (nel > 0) collect_marked(em, std::max<int64_t>(emin, 0), std::min(emax, nel - 1), elems);
    std::vector<uint8_t>().swap(em);
    std::vector<uint8_t> fm((size_t)nf, 0);
    int64_t fmin = nf, fmax = -1;
    for (int64_t e : elems)
        for (int i = 0; i < 3; i++) {
            const int64_t f = e2f[(size_t)i * nel + e] - 1;
            fm[f] = 1;
            fmin = std::min(fmin, f); fmax = std::max(fmax, f);
        }
    std::vector<int64_t> faces;
    if (fmax >= fmin) collect_marked(fm, fmin, fmax, faces);
    P.elems.swap(elems);
    P.faces.swap(faces);
    P.eown.resize(P.elems.size());
    P.fown.resize(P.faces.size());
    P.own_e.clear(); P.own_f.clear();
    P.own_e.reserve((size_t)std::max<int64_t>(0, e1 - e0));      // exactly the owned elements; most faces of a part are owned
    P.own_f.reserve(P.faces.size());
    for (size_t k = 0; k < P.elems.size(); k++) {
        P.eown[k] = in_range(P.elems[k]) ? 1 : 0;
        if (P.eown[k]) P.own_e.push_back((int64_t)k);
    }
    for (size_t k = 0; k < P.faces.size(); k++) {
        P.fown[k] = in_range(hi[P.faces[k]]) ? 1 : 0;
        if (P.fown[k]) P.own_f.push_back((int64_t)k);
    }
    make_runs(P.elems, P.erun);
    make_runs(P.faces, P.frun);
    auto owned_runs = [](const std::vector<int64_t>& idx, const std::vector<int64_t>& own, std::vector<std::pair<int64_t, int64_t>>& runs) {
        runs.clear();
        const size_t m = own.size();
        for (size_t a = 0; a < m;) {
            size_t b = a + 1;
            while (b < m && own[b] == own[b - 1] + 1 && idx[own[b]] == idx[own[b - 1]] + 1) b++;
            runs.emplace_back(own[a], (int64_t)(b - a));
            a = b;
        }
    };
    owned_runs(P.elems, P.own_e, P.oerun);
    owned_runs(P.faces, P.own_f, P.ofrun);
}

bool multi_hit(fcfv_t* h, const DBuf& rec, const void* src, size_t bytes);
void multi_note(fcfv_t* h, DBuf& rec, const void* src, size_t bytes);

int multi_load_nccl(fcfv_t* h) {
    fcfv_multi& M = *h->multi;
    if (M.nccl.lib) return FCFV_OK;
    fcfv_t* h0 = M.parts[0].h;
    const int rc = load_nccl(h0);
    if (rc) { h->err = h0->err; return rc; }
    M.nccl = h0->nccl;
    M.CommInitAll = (int (*)(void**, int, const int*))dlsym(M.nccl.lib, "ncclCommInitAll");
    M.GroupStart = (int (*)())dlsym(M.nccl.lib, "ncclGroupStart");
    M.GroupEnd = (int (*)())dlsym(M.nccl.lib, "ncclGroupEnd");
    if (!M.CommInitAll || !M.GroupStart || !M.GroupEnd) return fail(h, FCFV_ERR_NCCL, "libnccl is missing ncclCommInitAll / ncclGroupStart / ncclGroupEnd");
    return FCFV_OK;
}

// sum over the parts of n doubles at `field` of every part (device memory), result on every device
int multi_allreduce(fcfv_t* h, double* (*where)(fcfv_t*), int n) {
    fcfv_multi& M = *h->multi;
    if (M.parts.size() == 1) return FCFV_OK;
    if (M.comms.empty()) {
        // test hook (FCFV_MULTI_OVERSUBSCRIBE=1: several parts share a device, which NCCL does not allow): the same
        // sum taken on the host in part order and written back to every part
        std::vector<double> acc(n, 0.0), tmp(n);
        for (Part& Pt : M.parts) {
            CU(cudaSetDevice(Pt.h->device));
            CU(cudaMemcpyAsync(tmp.data(), where(Pt.h), sizeof(double) * n, cudaMemcpyDeviceToHost, Pt.h->stream));
            CU(cudaStreamSynchronize(Pt.h->stream));
            for (int k = 0; k < n; k++) acc[k] += tmp[k];
        }
        for (Part& Pt : M.parts) {
            CU(cudaSetDevice(Pt.h->device));
            CU(cudaMemcpyAsync(where(Pt.h), acc.data(), sizeof(double) * n, cudaMemcpyHostToDevice, Pt.h->stream));
            CU(cudaStreamSynchronize(Pt.h->stream));
        }
        return FCFV_OK;
    }
    int rc = M.GroupStart();
    if (rc) return fail(h, FCFV_ERR_NCCL, "ncclGroupStart failed: %s", M.nccl.GetErrorString ? M.nccl.GetErrorString(rc) : "?");
    for (size_t p = 0; p < M.parts.size() && !rc; p++) {
        fcfv_t* q = M.parts[p].h;
        cudaSetDevice(q->device);
        double* d = where(q);
        rc = M.nccl.AllReduce(d, d, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, M.comms[p], q->stream);
    }
    const int rc2 = M.GroupEnd();
    if (rc || rc2) return fail(h, FCFV_ERR_NCCL, "grouped ncclAllReduce failed: %s", M.nccl.GetErrorString ? M.nccl.GetErrorString(rc ? rc : rc2) : "?");
    return FCFV_OK;
}

double* sumsq_of(fcfv_t* q) { return P<double>(q->sumsq); }

int multi_set_mesh(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, const int64_t* bc, const double* Omega, const double* nx,
                   const double* ny, const double* Gamma, const double* ke, const double* delta, const double* taue, int64_t ntaue) {
    fcfv_multi& M = *h->multi;
    const int np = (int)M.parts.size();
    if (nel < 0 || nf < 0 || nel >= (1LL << 31) || nf >= (1LL << 30))
        return fail(h, FCFV_ERR_INVALID, "mesh size out of range (nel=%lld, nf=%lld)", (long long)nel, (long long)nf);
    if (!e2f || !bc || !Omega || !nx || !ny || !Gamma || !ke || !delta) return fail(h, FCFV_ERR_INVALID, "fcfv_set_mesh: null mesh array");
    if (taue && ntaue < nel) return fail(h, FCFV_ERR_INVALID, "taue has %lld entries, need at least nel=%lld", (long long)ntaue, (long long)nel);
    for (const void* p : {(const void*)e2f, (const void*)bc, (const void*)Omega, (const void*)nx, (const void*)ny, (const void*)Gamma,
                          (const void*)ke, (const void*)delta, (const void*)taue})
        if (p && is_device(p)) return fail(h, FCFV_ERR_INVALID, "a multi-GPU handle takes host arrays (it scatters them over its devices)");
    h->have_mesh = false; h->have_local = false;
    new_epoch(h);
    if (nel != h->nel || nf != h->nf) h->have_sources = h->have_face_data = h->have_solution = false;
    h->nel = nel; h->nf = nf; h->nel_global = nel; h->nf_global = nf;
    h->rp64 = 20 * nf >= (1LL << 31);
    h->Kuu.nnz = h->Kup.nnz = h->Muu.nnz = h->Kuusc.nnz = -1;
    const auto t0 = std::chrono::steady_clock::now();
    AdjVec lo, hi;
    if (!face_adjacency(nel, nf, e2f, lo, hi)) return fail(h, FCFV_ERR_INVALID, "e2f contains a face id outside 1..nf");
    TRY(for_parts(h, [&](int p, Part& Pt) -> int {
        build_part(nel, nf, e2f, lo, hi, p, np, Pt);
        return FCFV_OK;
    }));
    M.unref.clear();
    for (int64_t f = 0; f < nf; f++)
        if (hi[f] < 0) M.unref.push_back(f);
    M.unref_tau.assign(M.unref.size(), 0.0);
    M.partition_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    TRY(for_parts(h, [&](int, Part& Pt) -> int {
        const int64_t ne = (int64_t)Pt.elems.size(), nfl = (int64_t)Pt.faces.size();
        uvec<int64_t> le2f((size_t)3 * ne), lbc;
        {   // global -> local face ids through a dense map over the span of the part's faces
            const int64_t fmin = nfl ? Pt.faces.front() : 0, fspan = nfl ? Pt.faces.back() - fmin + 1 : 0;
            uvec<int32_t> fl((size_t)fspan);
            for (int64_t k = 0; k < nfl; k++) fl[(size_t)(Pt.faces[k] - fmin)] = (int32_t)(k + 1);
            for (int i = 0; i < 3; i++)
                for (int64_t k = 0; k < ne; k++) le2f[(size_t)i * ne + k] = fl[(size_t)(e2f[(size_t)i * nel + Pt.elems[k]] - 1 - fmin)];
        }
        take(bc, nf, Pt.faces, 1, lbc);
        fcfv_t* q = Pt.h;
        if (Pt.erun.size() <= max_direct_runs()) {
            // the floating-point fields go from the caller's arrays into the part's own buffers; fcfv_set_mesh then finds
            // them in place (device pointers of its own buffers: nothing is copied twice)
            TRY(upload_runs(q, q->Om, Omega, nel, Pt.elems, Pt.erun, 1)); TRY(upload_runs(q, q->nx, nx, nel, Pt.elems, Pt.erun, 3));
            TRY(upload_runs(q, q->ny, ny, nel, Pt.elems, Pt.erun, 3)); TRY(upload_runs(q, q->G, Gamma, nel, Pt.elems, Pt.erun, 3));
            TRY(upload_runs(q, q->ke, ke, nel, Pt.elems, Pt.erun, 1)); TRY(upload_runs(q, q->dl, delta, nel, Pt.elems, Pt.erun, 1));
            if (taue) TRY(upload_runs(q, q->taue, taue, ntaue, Pt.elems, Pt.erun, 1));
            TRY(fcfv_set_mesh(q, ne, nfl, le2f.data(), lbc.data(), P<double>(q->Om), P<double>(q->nx), P<double>(q->ny), P<double>(q->G),
                              P<double>(q->ke), P<double>(q->dl), taue ? P<double>(q->taue) : nullptr, ne));
        } else {
            uvec<double> lOm, lnx, lny, lG, lke, ldl, lte;
            take(Omega, nel, Pt.elems, 1, lOm); take(nx, nel, Pt.elems, 3, lnx); take(ny, nel, Pt.elems, 3, lny);
            take(Gamma, nel, Pt.elems, 3, lG); take(ke, nel, Pt.elems, 1, lke); take(delta, nel, Pt.elems, 1, ldl);
            if (taue) take(taue, ntaue, Pt.elems, 1, lte);
            TRY(fcfv_set_mesh(q, ne, nfl, le2f.data(), lbc.data(), lOm.data(), lnx.data(), lny.data(), lG.data(), lke.data(), ldl.data(),
                              taue ? lte.data() : nullptr, ne));
        }
        return fcfv_set_ownership(Pt.h, Pt.eown.data(), Pt.fown.data(), nel, nf, nullptr);
    }));
    M.tau_ref_stale = true;
    h->have_mesh = true;
    // the field refresh of the first ComputeFCFV on this mesh (same epoch) does not scatter them a second time
    multi_note(h, h->ke, ke, sizeof(double) * (size_t)nel);
    multi_note(h, h->dl, delta, sizeof(double) * (size_t)nel);
    multi_note(h, h->taue, taue, sizeof(double) * (size_t)nel);
    return FCFV_OK;
}

// one call of a part-level setter with arrays restricted to the part: kind 'e' (per element) / 'f' (per face)
// rec: the top-level handle's (otherwise unused) buffer record of the same name -- the upload cache of a multi-GPU handle
// works on the GLOBAL arrays: when every array of a call is one the parts already mirror (same address, size, epoch and
// fingerprint), the whole gather / scatter is skipped
struct Field { const double* g; char kind; int ncol; DBuf* rec; };

bool multi_hit(fcfv_t* h, const DBuf& rec, const void* src, size_t bytes) {
    return h->cache_on && rec.hm.epoch == h->epoch && rec.hm.ptr == src && rec.hm.bytes == bytes && fingerprint(src, bytes) == rec.hm.fp;
}
void multi_note(fcfv_t* h, DBuf& rec, const void* src, size_t bytes) {
    if (h->cache_on && src) rec.hm = HostMirror{src, bytes, fingerprint(src, bytes), h->epoch};
    else rec.hm.epoch = -1;
}

template <typename Call>
int multi_scatter(fcfv_t* h, std::initializer_list<Field> fields, Call&& call) {
    const std::vector<Field> fl(fields);
    for (const Field& f : fl)
        if (f.g && is_device(f.g)) return fail(h, FCFV_ERR_INVALID, "a multi-GPU handle takes host arrays");
    auto bytes_of = [&](const Field& f) { return sizeof(double) * (size_t)f.ncol * (size_t)(f.kind == 'e' ? h->nel : h->nf); };
    if (h->cache_on) {
        bool all = true, any = false;
        for (const Field& f : fl)
            if (f.g) { any = true; all = all && f.rec && multi_hit(h, *f.rec, f.g, bytes_of(f)); }
        if (any && all) {
            for (const Field& f : fl)
                if (f.g) { h->cache_hits++; h->cache_hit_bytes += (long long)bytes_of(f); }
            return FCFV_OK;
        }
    }
    for (const Field& f : fl)
        if (f.rec) f.rec->hm.epoch = -1;
    TRY(for_parts(h, [&](int, Part& Pt) -> int {
        std::vector<uvec<double>> loc(fl.size());
        std::vector<const double*> ptr(fl.size(), nullptr);
        for (size_t k = 0; k < fl.size(); k++) {
            if (!fl[k].g) continue;
            const bool el = fl[k].kind == 'e';
            const auto& runs = el ? Pt.erun : Pt.frun;
            if (fl[k].rec && runs.size() <= max_direct_runs()) {
                // straight into the part's buffer of the same name; the setter below is handed that device pointer
                DBuf& pb = *reinterpret_cast<DBuf*>(reinterpret_cast<char*>(Pt.h) + (reinterpret_cast<char*>(fl[k].rec) - reinterpret_cast<char*>(h)));
                TRY(upload_runs(Pt.h, pb, fl[k].g, el ? h->nel : h->nf, el ? Pt.elems : Pt.faces, runs, fl[k].ncol));
                ptr[k] = P<double>(pb);
                continue;
            }
            if (el) take(fl[k].g, h->nel, Pt.elems, fl[k].ncol, loc[k]);
            else take(fl[k].g, h->nf, Pt.faces, fl[k].ncol, loc[k]);
            ptr[k] = loc[k].data();
        }
        const bool was = Pt.h->async;
        Pt.h->async = false;               // the temporaries die at the end of this scope
        const int rc = call(Pt.h, ptr);
        Pt.h->async = was;
        return rc;
    }));
    for (const Field& f : fl)
        if (f.g && f.rec) multi_note(h, *f.rec, f.g, bytes_of(f));
    return FCFV_OK;
}

int multi_update_fields(fcfv_t* h, const double* ke, const double* delta, const double* taue, int64_t ntaue, const double* tau) {
    if (taue && ntaue < h->nel) return fail(h, FCFV_ERR_INVALID, "taue has %lld entries, need at least nel", (long long)ntaue);
    if (ke) TRY(multi_scatter(h, {{ke, 'e', 1, &h->ke}}, [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_update_fields(q, a[0], nullptr, nullptr, 0, nullptr); }));
    if (delta) TRY(multi_scatter(h, {{delta, 'e', 1, &h->dl}}, [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_update_fields(q, nullptr, a[0], nullptr, 0, nullptr); }));
    const bool tau_hit = tau && multi_hit(h, h->tau, tau, sizeof(double) * (size_t)h->nf);
    if (tau) TRY(multi_scatter(h, {{tau, 'f', 1, &h->tau}}, [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_update_fields(q, nullptr, nullptr, nullptr, 0, a[0]); }));
    if (tau_hit) tau = nullptr;       // the parts already hold it: nothing below has anything new to do
    if (taue && multi_hit(h, h->taue, taue, sizeof(double) * (size_t)h->nel)) {
        h->cache_hits++; h->cache_hit_bytes += (long long)(sizeof(double) * (size_t)h->nel);
        taue = nullptr;
    }
    if (taue) {     // may have more than nel entries (ViscousInclusionFCFV_v2.jl:127): only [0, nel) is read
        TRY(multi_scatter(h, {{taue, 'e', 1, &h->taue}}, [&](fcfv_t* q, const std::vector<const double*>& a) {
            return fcfv_update_fields(q, nullptr, nullptr, a[0], q->nel, nullptr);
        }));
    }
    if (tau)
        for (size_t k = 0; k < h->multi->unref.size(); k++) h->multi->unref_tau[k] = tau[h->multi->unref[k]];
    if (tau && h->nf >= h->nel) {
        // FCFV_TAU_REFERENCE reads the FACE array at the ELEMENT's global id (Residuals:17,54): tau_ref[e] = tau[global id of e]
        TRY(for_parts(h, [&](int, Part& Pt) -> int {
            std::vector<double> ref(Pt.elems.size());
            for (size_t k = 0; k < Pt.elems.size(); k++) ref[k] = tau[Pt.elems[k]];
            return fcfv_set_ownership(Pt.h, Pt.eown.data(), Pt.fown.data(), h->nel, h->nf, ref.data());
        }));
        h->multi->tau_ref_stale = false;
    }
    return FCFV_OK;
}

int multi_set_sources(fcfv_t* h, const double* sex, const double* sey) {
    if (!sex || !sey) return fail(h, FCFV_ERR_INVALID, "fcfv_set_sources: null array");
    TRY(multi_scatter(h, {{sex, 'e', 1, &h->sex}, {sey, 'e', 1, &h->sey}}, [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_set_sources(q, a[0], a[1]); }));
    h->have_sources = true;
    return FCFV_OK;
}

int multi_set_face_data(fcfv_t* h, const double* VxDir, const double* VyDir, const double* sxx, const double* syy, const double* sxy,
                        const double* syx) {
    // every part lists its Dirichlet / Neumann faces: the values on those faces go from the global arrays to the parts and
    // nothing else (fcfv_sparse_face_data); no nf-sized array is scattered
    bool listed = true;
    for (const Part& Pt : h->multi->parts) listed = listed && Pt.h->sparse_faces;
    const double* const src[6] = {VxDir, VyDir, sxx, syy, sxy, syx};
    for (const double* p : src)
        if (p && is_device(p)) return fail(h, FCFV_ERR_INVALID, "a multi-GPU handle takes host arrays");
    if (listed) {
        for (DBuf* rec : {&h->VxDir, &h->VyDir, &h->sxx, &h->syy, &h->sxy, &h->syx}) rec->hm.epoch = -1;
        TRY(for_parts(h, [&](int, Part& Pt) -> int { return set_face_data_listed(Pt.h, src, Pt.faces.data()); }));
        h->have_face_data = true;
        return FCFV_OK;
    }
    // NULL = keep what is resident, so the Dirichlet pair and the Neumann quadruple are scattered (or skipped) separately
    const bool first = !h->have_face_data;
    if (first || VxDir || VyDir)
        TRY(multi_scatter(h, {{VxDir, 'f', 1, &h->VxDir}, {VyDir, 'f', 1, &h->VyDir}},
                          [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_set_face_data(q, a[0], a[1], nullptr, nullptr, nullptr, nullptr); }));
    if (sxx || syy || sxy || syx)
        TRY(multi_scatter(h, {{sxx, 'f', 1, &h->sxx}, {syy, 'f', 1, &h->syy}, {sxy, 'f', 1, &h->sxy}, {syx, 'f', 1, &h->syx}},
                          [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_set_face_data(q, nullptr, nullptr, a[0], a[1], a[2], a[3]); }));
    h->have_face_data = true;
    return FCFV_OK;
}

int multi_set_solution(fcfv_t* h, const double* Vxh, const double* Vyh, const double* Pe) {
    if (!Vxh || !Vyh || !Pe) return fail(h, FCFV_ERR_INVALID, "fcfv_set_solution: null array");
    TRY(multi_scatter(h, {{Vxh, 'f', 1, &h->Vxh}, {Vyh, 'f', 1, &h->Vyh}, {Pe, 'e', 1, &h->Pe}},
                      [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_set_solution(q, a[0], a[1], a[2]); }));
    h->have_solution = true;
    return FCFV_OK;
}

int multi_set_local(fcfv_t* h, const double* alpha, const double* beta, const double* Z) {
    if (!alpha || !beta || !Z) return fail(h, FCFV_ERR_INVALID, "fcfv_set_local: null array");
    TRY(multi_scatter(h, {{alpha, 'e', 1, &h->alpha}, {beta, 'e', 2, &h->beta}, {Z, 'e', 4, &h->Z}},
                      [&](fcfv_t* q, const std::vector<const double*>& a) { return fcfv_set_local(q, a[0], a[1], a[2]); }));
    h->have_local = true;
    return FCFV_OK;
}

// part-level run_* on every device (asynchronous), then one synchronisation unless the handle is in async mode
template <typename Call>
int multi_run(fcfv_t* h, Call&& call) {
    for (Part& Pt : h->multi->parts) {
        const bool was = Pt.h->async;
        Pt.h->async = true;
        const int rc = call(Pt.h);
        Pt.h->async = was;
        if (rc) { h->err = Pt.h->err; return rc; }
    }
    return FCFV_OK;
}

int multi_get_local_parts(fcfv_t* h, double* alpha, double* beta, double* Z, double* tau);

int multi_run_local(fcfv_t* h, int formulation, double A) {
    TRY(need(h, h->have_mesh, "fcfv_run_local: no mesh set"));
    TRY(need(h, h->have_sources && h->have_face_data, "fcfv_run_local: sources / face data not set"));
    TRY(multi_run(h, [&](fcfv_t* q) { return fcfv_run_local(q, formulation, A); }));
    h->alpha.hm.epoch = h->beta.hm.epoch = h->Z.hm.epoch = h->tau.hm.epoch = -1;      // rewritten on the devices
    h->have_local = true;
    h->multi->tau_ref_stale = true;
    return multi_finish(h);
}

int multi_get_local(fcfv_t* h, double* alpha, double* beta, double* Z, double* tau) {
    TRY(need(h, h->have_local, "fcfv_get_local: local coefficients not computed"));
    if (tau)
        for (size_t k = 0; k < h->multi->unref.size(); k++) tau[h->multi->unref[k]] = h->multi->unref_tau[k];
    TRY(multi_get_local_parts(h, alpha, beta, Z, tau));
    // the caller's arrays now mirror what the parts hold
    multi_note(h, h->alpha, alpha, sizeof(double) * (size_t)h->nel);
    multi_note(h, h->beta, beta, sizeof(double) * 2 * (size_t)h->nel);
    multi_note(h, h->Z, Z, sizeof(double) * 4 * (size_t)h->nel);
    if (h->multi->unref.empty()) multi_note(h, h->tau, tau, sizeof(double) * (size_t)h->nf);
    return FCFV_OK;
}

int multi_get_local_parts(fcfv_t* h, double* alpha, double* beta, double* Z, double* tau) {
    return for_parts(h, [&](int, Part& Pt) -> int {
        const size_t ne = Pt.elems.size(), nfl = Pt.faces.size();
        if (Pt.oerun.size() <= max_direct_runs() && Pt.ofrun.size() <= max_direct_runs()) {
            fcfv_t* q = Pt.h;
            TRY(need(q, q->have_local, "fcfv_get_local: local coefficients not computed"));
            if (alpha) TRY(download_runs(q, P<double>(q->alpha), alpha, h->nel, Pt.elems, Pt.oerun, 1));
            if (beta) TRY(download_runs(q, P<double>(q->beta), beta, h->nel, Pt.elems, Pt.oerun, 2));
            if (Z) TRY(download_runs(q, P<double>(q->Z), Z, h->nel, Pt.elems, Pt.oerun, 4));
            if (tau) TRY(download_runs(q, P<double>(q->tau), tau, h->nf, Pt.faces, Pt.ofrun, 1));
            return sync(q);
        }
        uvec<double> a(alpha ? ne : 0), b(beta ? 2 * ne : 0), z(Z ? 4 * ne : 0), t(tau ? nfl : 0);
        TRY(fcfv_get_local(Pt.h, alpha ? a.data() : nullptr, beta ? b.data() : nullptr, Z ? z.data() : nullptr, tau ? t.data() : nullptr));
        if (alpha) put_owned(alpha, h->nel, Pt.elems, Pt.own_e, 1, a.data());
        if (beta) put_owned(beta, h->nel, Pt.elems, Pt.own_e, 2, b.data());
        if (Z) put_owned(Z, h->nel, Pt.elems, Pt.own_e, 4, z.data());
        if (tau) put_owned(tau, h->nf, Pt.faces, Pt.own_f, 1, t.data());
        return FCFV_OK;
    });
}

int multi_run_assemble(fcfv_t* h, int variant, int formulation, double A, int extra, double gamma) {
    TRY(need(h, h->have_mesh, "fcfv_run_assemble: no mesh set"));
    TRY(need(h, h->have_local, "fcfv_run_assemble: local coefficients missing (fcfv_compute_local / fcfv_set_local)"));
    TRY(need(h, h->have_face_data, "fcfv_run_assemble: face data not set"));
    const auto t0 = std::chrono::steady_clock::now();
    TRY(multi_run(h, [&](fcfv_t* q) { return run_assemble_impl(q, variant, formulation, A, extra, gamma); }));
    h->Kuu.nnz = h->Kup.nnz = -2;
    if (extra != kExtraSchur) h->Muu.nnz = extra == kExtraMuu ? -2 : -1;
    h->Kuusc.nnz = extra == kExtraSchur ? -2 : -1;
    h->Kuu.nrows = h->Kuu.ncols = h->Kup.nrows = h->Muu.nrows = h->Muu.ncols = h->Kuusc.nrows = h->Kuusc.ncols = 2 * h->nf;
    h->Kup.ncols = h->nel;
    h->totals_valid = false;
    h->last_variant = variant; h->last_form = formulation; h->last_A = A;
    TRY(multi_finish(h));
    h->multi->asm_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return FCFV_OK;
}

// every part's rows outside its ownership are empty, so the global counts are the sums of the parts' counts
int multi_fetch_totals(fcfv_t* h) {
    if (h->totals_valid) return FCFV_OK;
    long long uu = 0, up = 0, mu = 0, sc = 0;
    for (Part& Pt : h->multi->parts) {
        cudaSetDevice(Pt.h->device);
        const int rc = fetch_totals(Pt.h);
        if (rc) { h->err = Pt.h->err; return rc; }
        uu += Pt.h->Kuu.nnz; up += Pt.h->Kup.nnz;
        if (Pt.h->Muu.nnz > 0) mu += Pt.h->Muu.nnz;
        if (Pt.h->Kuusc.nnz > 0) sc += Pt.h->Kuusc.nnz;
    }
    h->Kuu.nnz = uu; h->Kup.nnz = up;
    if (h->Muu.nnz == -2) h->Muu.nnz = mu;
    if (h->Kuusc.nnz == -2) h->Kuusc.nnz = sc;
    h->totals_valid = true;
    return FCFV_OK;
}

// The global block in the reference's CSR order, built on the host from the owned rows of every part (columns
// mapped local -> global).  rowptr: int64.
int multi_gather_csr(fcfv_t* h, int block, std::vector<int64_t>& rowptr, std::vector<int>& col, std::vector<double>& val, bool entries) {
    fcfv_multi& M = *h->multi;
    const int64_t nf = h->nf, nrows = 2 * nf;
    const bool elem_cols = block == FCFV_KUP;
    const int np = (int)M.parts.size();
    std::vector<std::vector<int64_t>> lrp(np);
    std::vector<std::vector<int>> lci(np);
    std::vector<std::vector<double>> lv(np);
    rowptr.assign(nrows + 1, 0);
    TRY(for_parts(h, [&](int p, Part& Pt) -> int {
        int64_t lr, lc, lnnz;
        TRY(fcfv_block_size(Pt.h, block, &lr, &lc, &lnnz));
        const int w = fcfv_index_width(Pt.h);
        std::vector<char> raw((size_t)w * (lr + 1));
        if (entries) { lci[p].resize(lnnz); lv[p].resize(lnnz); }
        TRY(fcfv_export(Pt.h, block, FCFV_CSR0, raw.data(), entries ? lci[p].data() : nullptr, entries ? lv[p].data() : nullptr));
        lrp[p].resize(lr + 1);
        for (int64_t r = 0; r <= lr; r++) lrp[p][r] = w == 8 ? ((const int64_t*)raw.data())[r] : (int64_t)((const int*)raw.data())[r];
        const int64_t nfl = (int64_t)Pt.faces.size();
        for (int64_t j : Pt.own_f)                       // owned rows of different parts are disjoint: no race
            for (int c = 0; c < 2; c++)
                rowptr[(size_t)c * nf + Pt.faces[j] + 1] = lrp[p][(size_t)c * nfl + j + 1] - lrp[p][(size_t)c * nfl + j];
        return FCFV_OK;
    }));
    for (int64_t r = 0; r < nrows; r++) rowptr[r + 1] += rowptr[r];
    if (!entries) return FCFV_OK;
    col.resize((size_t)rowptr[nrows]);
    val.resize((size_t)rowptr[nrows]);
    return for_parts(h, [&](int p, Part& Pt) -> int {
        const int64_t nfl = (int64_t)Pt.faces.size();
        for (int64_t j : Pt.own_f)
            for (int c = 0; c < 2; c++) {
                const int64_t l0 = lrp[p][(size_t)c * nfl + j], l1 = lrp[p][(size_t)c * nfl + j + 1];
                int64_t o = rowptr[(size_t)c * nf + Pt.faces[j]];
                for (int64_t k = l0; k < l1; k++, o++) {
                    const int64_t lc = lci[p][k];
                    col[o] = elem_cols ? (int)Pt.elems[lc] : (lc >= nfl ? (int)(nf + Pt.faces[lc - nfl]) : (int)Pt.faces[lc]);
                    val[o] = lv[p][k];
                }
            }
        return FCFV_OK;
    });
}

int multi_export(fcfv_t* h, int block, int layout, void* ptr, void* idx, double* val) {
    Csr* c = block_of(h, block);
    if (!c) return fail(h, FCFV_ERR_INVALID, "unknown block %d", block);
    TRY(need(h, c->nnz != -1, "fcfv_export: block not assembled"));
    TRY(multi_fetch_totals(h));
    std::vector<int64_t> rp;
    std::vector<int> ci;
    std::vector<double> v;
    const int64_t nrows = 2 * h->nf, ncols = block == FCFV_KUP ? h->nel : 2 * h->nf;
    if (layout == FCFV_CSR0) {
        TRY(multi_gather_csr(h, block, rp, ci, v, idx != nullptr || val != nullptr));
        if (ptr) {
            if (h->rp64) memcpy(ptr, rp.data(), sizeof(int64_t) * (nrows + 1));
            else for (int64_t r = 0; r <= nrows; r++) ((int*)ptr)[r] = (int)rp[r];
        }
        if (idx) memcpy(idx, ci.data(), sizeof(int) * ci.size());
        if (val) memcpy(val, v.data(), sizeof(double) * v.size());
        return FCFV_OK;
    }
    if (layout != FCFV_CSC1_INT64) return fail(h, FCFV_ERR_INVALID, "unknown layout %d", layout);
    if (!ptr || !idx || !val) return fail(h, FCFV_ERR_INVALID, "fcfv_export: CSC export needs ptr, idx and val");
    TRY(multi_gather_csr(h, block, rp, ci, v, true));
    // CSR -> Julia CSC on the host: rows are visited in ascending order, so every column comes out sorted by row
    int64_t* colptr = (int64_t*)ptr;
    int64_t* rowval = (int64_t*)idx;
    std::vector<int64_t> cur(ncols + 1, 0);
    for (int c2 : ci) cur[(size_t)c2 + 1]++;
    colptr[0] = 1;
    for (int64_t j = 0; j < ncols; j++) { colptr[j + 1] = colptr[j] + cur[j + 1]; cur[j] = colptr[j] - 1; }
    for (int64_t r = 0; r < nrows; r++)
        for (int64_t k = rp[r]; k < rp[r + 1]; k++) {
            const int64_t o = cur[ci[k]]++;
            rowval[o] = r + 1;
            val[o] = v[k];
        }
    return FCFV_OK;
}

int multi_export_rhs(fcfv_t* h, double* fu, double* fp) {
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_export_rhs: nothing assembled"));
    if (fu)      // rows of faces no element references: empty, right-hand side +0.0 (as on one GPU)
        for (int64_t f : h->multi->unref) { fu[f] = 0.0; fu[(size_t)h->nf + f] = 0.0; }
    return for_parts(h, [&](int, Part& Pt) -> int {
        const size_t ne = Pt.elems.size(), nfl = Pt.faces.size();
        std::vector<double> lfu(fu ? 2 * nfl : 0), lfp(fp ? ne : 0);
        TRY(fcfv_export_rhs(Pt.h, fu ? lfu.data() : nullptr, fp ? lfp.data() : nullptr));
        if (fu) put_owned(fu, h->nf, Pt.faces, Pt.own_f, 2, lfu.data());      // fu is 2nf x 1: the y block is a second column of length nf
        if (fp) put_owned(fp, h->nel, Pt.elems, Pt.own_e, 1, lfp.data());
        return FCFV_OK;
    });
}

// FCFV_TAU_REFERENCE on a partition needs tau at the element's GLOBAL id: collect the face array, hand every part its slice
int multi_refresh_tau_ref(fcfv_t* h) {
    if (!h->multi->tau_ref_stale) return FCFV_OK;
    if (h->nf < h->nel) return fail(h, FCFV_ERR_INVALID, "FCFV_TAU_REFERENCE reads tau[e] for e <= nel but nf < nel");
    std::vector<double> tau((size_t)h->nf, 0.0);
    TRY(for_parts(h, [&](int, Part& Pt) -> int {
        std::vector<double> t(Pt.faces.size());
        TRY(fcfv_get_local(Pt.h, nullptr, nullptr, nullptr, t.data()));
        put_owned(tau.data(), h->nf, Pt.faces, Pt.own_f, 1, t.data());
        return FCFV_OK;
    }));
    TRY(for_parts(h, [&](int, Part& Pt) -> int {
        std::vector<double> ref(Pt.elems.size());
        for (size_t k = 0; k < Pt.elems.size(); k++) ref[k] = tau[Pt.elems[k]];
        return fcfv_set_ownership(Pt.h, Pt.eown.data(), Pt.fown.data(), h->nel, h->nf, ref.data());
    }));
    h->multi->tau_ref_stale = false;
    return FCFV_OK;
}

int multi_run_residuals(fcfv_t* h, int formulation, int tau_mode, double* sumsq, double* norms) {
    TRY(check_form(h, formulation));
    if (tau_mode != FCFV_TAU_REFERENCE && tau_mode != FCFV_TAU_FACE) return fail(h, FCFV_ERR_INVALID, "tau_mode must be FCFV_TAU_REFERENCE or FCFV_TAU_FACE (got %d)", tau_mode);
    TRY(need(h, h->have_mesh && h->have_local, "fcfv_run_residuals: mesh / local coefficients missing"));
    TRY(need(h, h->have_sources && h->have_face_data && h->have_solution, "fcfv_run_residuals: sources / face data / solution not set"));
    if (tau_mode == FCFV_TAU_REFERENCE) TRY(multi_refresh_tau_ref(h));
    // kernels on every device (the parts' own all-reduce is switched off: one thread drives all devices, so the
    // collective has to be issued as a group)
    TRY(multi_run(h, [&](fcfv_t* q) { return fcfv_run_residuals(q, formulation, tau_mode, nullptr, nullptr); }));
    TRY(multi_allreduce(h, sumsq_of, 6));
    if (sumsq || norms) {
        fcfv_t* q = h->multi->parts[0].h;
        double hs[6];
        CU(cudaSetDevice(q->device));
        CU(cudaMemcpyAsync(hs, P<double>(q->sumsq), sizeof hs, cudaMemcpyDeviceToHost, q->stream));
        TRY(multi_sync(h));
        if (sumsq) memcpy(sumsq, hs, sizeof hs);
        if (norms) {
            const double ne = (double)h->nel, nfg = (double)h->nf;
            const double len[6] = {4.0 * ne, 2.0 * ne, ne, nfg, nfg, ne};
            for (int k = 0; k < 6; k++) norms[k] = len[k] > 0 ? sqrt(hs[k]) / sqrt(len[k]) : 0.0;
        }
        return FCFV_OK;
    }
    return multi_finish(h);
}

// ---- element values and the Powell-Hestenes lines on a multi-GPU handle ----------------------------------------------
// Global arrays in and out: the inputs are restricted to every part (owned + ghost entries), the part-level entry runs on
// its device, and each part returns the entries it owns.  Nothing here is on the hot path (post-processing / the outer
// iteration around a host factorisation), so the outputs come back through part-local host temporaries.
Part& part_of(fcfv_t* h, fcfv_t* q) {
    for (Part& Pt : h->multi->parts)
        if (Pt.h == q) return Pt;
    return h->multi->parts[0];
}

int multi_element_values(fcfv_t* h, const double* Vxh, const double* Vyh, const double* alpha, const double* beta, const double* Z,
                         double* Vxe, double* Vye, double* Txxe, double* Tyye, double* Txye) {
    TRY(need(h, h->have_mesh, "fcfv_element_values: no mesh set"));
    if ((Vxh == nullptr) != (Vyh == nullptr)) return fail(h, FCFV_ERR_INVALID, "fcfv_element_values: pass both Vxh and Vyh or neither");
    if (alpha || beta || Z) {
        if (!alpha || !beta || !Z) return fail(h, FCFV_ERR_INVALID, "fcfv_element_values: pass alpha, beta and Z together (or none of them)");
        TRY(multi_set_local(h, alpha, beta, Z));
    }
    TRY(need(h, h->have_local, "fcfv_element_values: local coefficients missing"));
    TRY(need(h, Vxh != nullptr || h->have_solution, "fcfv_element_values: no face velocities (fcfv_set_solution)"));
    double* out[5] = {Vxe, Vye, Txxe, Tyye, Txye};
    auto run = [&](fcfv_t* q, const double* vx, const double* vy) -> int {
        Part& Pt = part_of(h, q);
        const size_t ne = Pt.elems.size();
        uvec<double> loc[5];
        double* lp[5];
        for (int k = 0; k < 5; k++) { if (out[k]) loc[k].resize(ne); lp[k] = out[k] ? loc[k].data() : nullptr; }
        TRY(fcfv_element_values(q, vx, vy, nullptr, nullptr, nullptr, lp[0], lp[1], lp[2], lp[3], lp[4]));
        for (int k = 0; k < 5; k++)
            if (out[k]) put_owned(out[k], h->nel, Pt.elems, Pt.own_e, 1, lp[k]);
        return FCFV_OK;
    };
    if (Vxh) {
        TRY(multi_scatter(h, {{Vxh, 'f', 1, nullptr}, {Vyh, 'f', 1, nullptr}},
                          [&](fcfv_t* q, const std::vector<const double*>& a) { return run(q, a[0], a[1]); }));
        h->Vxh.hm.epoch = h->Vyh.hm.epoch = -1;      // the parts' face velocities were replaced behind the cache's back
        h->have_solution = true;
        return FCFV_OK;
    }
    return for_parts(h, [&](int, Part& Pt) -> int { return run(Pt.h, nullptr, nullptr); });
}

int multi_ph_residual(fcfv_t* h, const double* u, const double* p, double* ru, double* rp, double* nrm) {
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_ph_residual: nothing assembled"));
    if (!u || !p) return fail(h, FCFV_ERR_INVALID, "fcfv_ph_residual: null u or p");
    if (ru)      // rows of faces no element references: empty, fu = 0
        for (int64_t f : h->multi->unref) { ru[f] = 0.0; ru[(size_t)h->nf + f] = 0.0; }
    TRY(multi_scatter(h, {{u, 'f', 2, nullptr}, {p, 'e', 1, nullptr}}, [&](fcfv_t* q, const std::vector<const double*>& a) -> int {
        Part& Pt = part_of(h, q);
        uvec<double> lru(ru ? 2 * Pt.faces.size() : 0), lrp(rp ? Pt.elems.size() : 0);
        TRY(fcfv_ph_residual(q, a[0], a[1], ru ? lru.data() : nullptr, rp ? lrp.data() : nullptr, nullptr));
        if (ru) put_owned(ru, h->nf, Pt.faces, Pt.own_f, 2, lru.data());
        if (rp) put_owned(rp, h->nel, Pt.elems, Pt.own_e, 1, lrp.data());
        return FCFV_OK;
    }));
    if (nrm) {
        // every part has left the sums of squares over ITS rows / elements at sumsq[6..7]
        TRY(multi_allreduce(h, [](fcfv_t* q) -> double* { return P<double>(q->sumsq) + 6; }, 2));
        fcfv_t* q = h->multi->parts[0].h;
        double hs[2];
        CU(cudaSetDevice(q->device));
        CU(cudaMemcpyAsync(hs, P<double>(q->sumsq) + 6, sizeof hs, cudaMemcpyDeviceToHost, q->stream));
        TRY(multi_sync(h));
        nrm[0] = sqrt(hs[0]); nrm[1] = sqrt(hs[1]);
    }
    return FCFV_OK;
}

int multi_ph_update_p(fcfv_t* h, double gamma, const double* u, double* p) {
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_ph_update_p: nothing assembled"));
    if (!u || !p) return fail(h, FCFV_ERR_INVALID, "fcfv_ph_update_p: null u or p");
    return multi_scatter(h, {{u, 'f', 2, nullptr}, {p, 'e', 1, nullptr}}, [&](fcfv_t* q, const std::vector<const double*>& a) -> int {
        Part& Pt = part_of(h, q);
        uvec<double> lp(Pt.elems.size());
        TRY(copy(q, lp.data(), a[1], sizeof(double) * lp.size()));     // the part's slice of p (host or device) ...
        TRY(sync(q));
        TRY(fcfv_ph_update_p(q, gamma, a[0], lp.data()));              // ... updated in place
        put_owned(p, h->nel, Pt.elems, Pt.own_e, 1, lp.data());
        return FCFV_OK;
    });
}

int multi_ph_fusc(fcfv_t* h, double gamma, const double* p, double* fusc) {
    TRY(need(h, h->Kuu.nnz != -1, "fcfv_ph_fusc: nothing assembled"));
    if (!p || !fusc) return fail(h, FCFV_ERR_INVALID, "fcfv_ph_fusc: null p or fusc");
    // w = (gamma ke / Omega) fp - p is needed on the ghost elements of a part as well, and fp of an element is only formed
    // by the part that owns it: collect it and hand every part its slice (owned entries are rewritten with their own values)
    uvec<double> fpg((size_t)h->nel);
    TRY(multi_export_rhs(h, nullptr, fpg.data()));
    TRY(multi_scatter(h, {{fpg.data(), 'e', 1, nullptr}}, [&](fcfv_t* q, const std::vector<const double*>& a) -> int {
        TRY(copy(q, q->fp.p, a[0], sizeof(double) * (size_t)q->nel));
        return sync(q);
    }));
    for (int64_t f : h->multi->unref) { fusc[f] = 0.0; fusc[(size_t)h->nf + f] = 0.0; }
    return multi_scatter(h, {{p, 'e', 1, nullptr}}, [&](fcfv_t* q, const std::vector<const double*>& a) -> int {
        Part& Pt = part_of(h, q);
        uvec<double> lf(2 * Pt.faces.size());
        TRY(fcfv_ph_fusc(q, gamma, a[0], lf.data()));
        put_owned(fusc, h->nf, Pt.faces, Pt.own_f, 2, lf.data());
        return FCFV_OK;
    });
}

// Pe .-= mean(Pe) (SolversFCFV_Stokes.jl:20) on the host: the array is the caller's and nothing of it is resident
int multi_center_pressure(fcfv_t* h, double* Pe, double* mean) {
    TRY(need(h, h->have_mesh, "fcfv_center_pressure: no mesh set"));
    if (!Pe) return fail(h, FCFV_ERR_INVALID, "fcfv_center_pressure: null Pe");
    if (is_device(Pe)) return fail(h, FCFV_ERR_INVALID, "a multi-GPU handle takes host arrays");
    const int64_t n = h->nel;
    // pairwise sums over fixed blocks (deterministic, close to what one GPU forms)
    double total = 0.0;
    for (int64_t b = 0; b < n; b += 4096) {
        double s = 0.0;
        for (int64_t e = b; e < std::min(n, b + 4096); e++) s += Pe[e];
        total += s;
    }
    const double m = n > 0 ? total / (double)n : 0.0;
    for (int64_t e = 0; e < n; e++) Pe[e] = Pe[e] - m;
    if (mean) *mean = m;
    return FCFV_OK;
}

double multi_asm_seconds(fcfv_t* h) { return h->multi->asm_seconds; }
int multi_has_orphans(fcfv_t* h) { return h->multi->unref.empty() ? 0 : 1; }
fcfv_t* multi_part0(fcfv_t* h) { return h->multi->parts[0].h; }
template <typename F>
void multi_each(fcfv_t* h, F&& fn) {
    for (Part& Pt : h->multi->parts) fn(Pt.h);
}

int multi_unsupported(fcfv_t* h, const char* what) {
    return fail(h, FCFV_ERR_STATE, "%s is not available on a multi-GPU handle (fcfv_create_multi); use a single-GPU handle", what);
}

void multi_destroy(fcfv_t* h) {
    fcfv_multi* M = h->multi;
    if (!M) return;
    for (size_t p = 0; p < M->parts.size(); p++) {
        if (!M->parts[p].h) continue;
        M->parts[p].h->comm = nullptr;                 // owned here
        if (p < M->comms.size() && M->comms[p] && M->nccl.CommDestroy) {
            cudaSetDevice(M->parts[p].h->device);
            cudaStreamSynchronize(M->parts[p].h->stream);
            M->nccl.CommDestroy(M->comms[p]);
        }
        fcfv_destroy(M->parts[p].h);
    }
    delete M;
    h->multi = nullptr;
}

}  // namespace

int fcfv_create_multi(fcfv_t** out, int ngpus) {
    if (!out) return FCFV_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return FCFV_ERR_CUDA;   // no CPU fallback
    // test hook: FCFV_MULTI_OVERSUBSCRIBE=1 lets the parts share the devices that exist (part p on device p % count), so
    // that partitioning, scatter and export can be exercised on a box with fewer GPUs; the all-reduce is then a host sum
    const char* ov = getenv("FCFV_MULTI_OVERSUBSCRIBE");
    const bool oversub = ov && ov[0] == '1' && ngpus > count;
    if (ngpus < 1 || ngpus > 64 || (ngpus > count && !oversub)) return FCFV_ERR_INVALID;
    if (ngpus == 1) return fcfv_create(out, 0);
    fcfv_t* h = new (std::nothrow) fcfv_handle();
    if (!h) return FCFV_ERR_ALLOC;
    h->multi = new (std::nothrow) fcfv_multi();
    if (!h->multi) { delete h; return FCFV_ERR_ALLOC; }
    h->multi->parts.resize(ngpus);
    int rc = FCFV_OK;
    for (int p = 0; p < ngpus && !rc; p++) {
        rc = fcfv_create(&h->multi->parts[p].h, p % count);
        if (!rc) h->multi->parts[p].h->stage_share = ngpus;      // the parts copy concurrently (for_parts): they share the host's threads
    }
    if (!rc && !oversub) rc = multi_load_nccl(h);
    if (!rc && !oversub) {
        std::vector<int> devs(ngpus);
        for (int p = 0; p < ngpus; p++) devs[p] = p;
        h->multi->comms.assign(ngpus, nullptr);
        const int nrc = h->multi->CommInitAll(h->multi->comms.data(), ngpus, devs.data());
        if (nrc) rc = FCFV_ERR_NCCL;
    }
    if (rc) { multi_destroy(h); delete h; return rc; }
    *out = h;
    return FCFV_OK;
}

int fcfv_num_devices(fcfv_t* h) { return !h ? 0 : (h->multi ? (int)h->multi->parts.size() : 1); }

int fcfv_partition_info(fcfv_t* h, int part, int64_t* nel_local, int64_t* nf_local, int64_t* nel_owned, int64_t* nf_owned, double* seconds) {
    if (!h) return FCFV_ERR_INVALID;
    if (!h->multi) return fail(h, FCFV_ERR_STATE, "fcfv_partition_info: not a multi-GPU handle");
    TRY(need(h, h->have_mesh, "fcfv_partition_info: no mesh set"));
    if (part < 0 || part >= (int)h->multi->parts.size()) return fail(h, FCFV_ERR_INVALID, "fcfv_partition_info: no such part");
    const Part& Pt = h->multi->parts[part];
    if (nel_local) *nel_local = (int64_t)Pt.elems.size();
    if (nf_local) *nf_local = (int64_t)Pt.faces.size();
    if (nel_owned) *nel_owned = (int64_t)Pt.own_e.size();
    if (nf_owned) *nf_owned = (int64_t)Pt.own_f.size();
    if (seconds) *seconds = h->multi->partition_seconds;
    return FCFV_OK;
}

// Host-only view of the partition (no GPU involved): what part `part` of `nparts` holds.  Call with the output arrays
// NULL to learn the sizes, then again with arrays of nel_local / nf_local entries.  Ids are 1-based and ascending.
int fcfv_partition_part(int64_t nel, int64_t nf, const int64_t* e2f, int nparts, int part, int64_t* nel_local, int64_t* nf_local,
                        int64_t* elem_global, int64_t* face_global, uint8_t* elem_owned, uint8_t* face_owned) {
    if (!e2f || nel < 0 || nf < 0 || nparts < 1 || part < 0 || part >= nparts) return FCFV_ERR_INVALID;
    if (nel >= (1LL << 31)) return FCFV_ERR_INVALID;
    AdjVec lo, hi;
    if (!face_adjacency(nel, nf, e2f, lo, hi)) return FCFV_ERR_INVALID;
    Part Pt;
    build_part(nel, nf, e2f, lo, hi, part, nparts, Pt);
    if (nel_local) *nel_local = (int64_t)Pt.elems.size();
    if (nf_local) *nf_local = (int64_t)Pt.faces.size();
    if (elem_global) for (size_t k = 0; k < Pt.elems.size(); k++) elem_global[k] = Pt.elems[k] + 1;
    if (face_global) for (size_t k = 0; k < Pt.faces.size(); k++) face_global[k] = Pt.faces[k] + 1;
    if (elem_owned) memcpy(elem_owned, Pt.eown.data(), Pt.eown.size());
    if (face_owned) memcpy(face_owned, Pt.fown.data(), Pt.fown.size());
    return FCFV_OK;
}
