/*
 * fcfv_b200.h -- C ABI of libfcfv_b200.so: the B200 (sm_100a, FP64) implementation of the
 * data-parallel hot path of tduretz/FCFV_NME23.
 *
 * This header is the drop-in boundary.  The reference is pure Julia and has no FFI of its
 * own; each entry point below replaces the body of one exported Julia function (cited as
 * file:line of the reference checkout) and is what a Julia `ccall` shim binds
 * (julia/FCFV_B200.jl).  tests/ and bench.py bind the same symbols through
 * Python ctypes.
 *
 * Conventions
 *  - All functions return an int status (FCFV_OK == 0).  No exceptions, no printing, no
 *    signal handlers.  fcfv_last_error(h) returns a message owned by the handle.
 *  - Array arguments use the reference's (Julia) memory layout: Float64 / Int64,
 *    column-major, 1-based connectivity.  X[e,i] of an nel x 3 matrix is at X[(i-1)*nel+(e-1)],
 *    Z[e,a,b] of nel x 2 x 2 at Z[(e-1) + nel*(a-1) + 2*nel*(b-1)].
 *  - Every array pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer of the
 *    handle's GPU; the library detects which (cudaPointerGetAttributes).  Pointers are borrowed
 *    for the duration of the call only.  Output arrays are caller-allocated; NULL skips them.
 *  - A handle owns one GPU (fcfv_create) or several (fcfv_create_multi), one stream per GPU and all device-resident
 *    state.  It is not thread-safe.
 *    Calls return after the stream has been synchronised unless stated otherwise.
 *  - Assembled blocks stay on the device as CSR: 0-based, int32 column indices, row pointers
 *    int32 when 20*nf < 2^31 and int64 otherwise (fcfv_index_width).  Within a row columns are
 *    ascending; entries whose summed value is exactly +-0.0 are dropped, like
 *    dropzeros(sparse(...)) in the reference (DiscretisationFCFV_Stokes.jl:355-364).
 */
#ifndef FCFV_B200_H
#define FCFV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fcfv_handle fcfv_t;

enum { FCFV_OK = 0, FCFV_ERR_INVALID = 1, FCFV_ERR_CUDA = 2, FCFV_ERR_STATE = 3, FCFV_ERR_ALLOC = 4, FCFV_ERR_NCCL = 5 };
enum { FCFV_GRADIENT = 0, FCFV_SYMMETRIC_GRADIENT = 1 };  /* Formulation == :Gradient / :SymmetricGradient */
enum { FCFV_LOOP_OLD = 0, FCFV_LOOP_NEW = 1 };             /* ElementAssemblyLoop / ElementAssemblyLoopNEW  */
enum { FCFV_TAU_REFERENCE = 0, FCFV_TAU_FACE = 1 };        /* residual: mesh.τ[e] (as written) / mesh.τ[e2f[e,i]] */
enum { FCFV_KUU = 0, FCFV_KUP = 1, FCFV_MUU = 2, FCFV_KUUSC = 3 };   /* KUUSC: Schur complement, see fcfv_schur */
enum { FCFV_CSR0 = 0, FCFV_CSC1_INT64 = 1 };               /* export layouts */

/* ---- life cycle ---------------------------------------------------------------------- */
int fcfv_create(fcfv_t** h, int device);
/* Several GPUs behind ONE handle, for the reference's single-process drivers (examples/ViscousInclusionFCFV_v2.jl:60-154,
 * examples/SolKzFCFV.jl:50-215): devices 0..ngpus-1 of this process.  Every entry point keeps its meaning and takes / returns
 * the GLOBAL arrays of the unpartitioned mesh (host pointers only).  fcfv_set_mesh splits the mesh by element (contiguous
 * ranges; a face belongs to the part of its highest adjacent element, the element whose tau_e DiscretisationFCFV_Stokes.jl:40
 * leaves in mesh.tau; one layer of ghost elements is replicated) and scatters the arrays; each part assembles the rows of its
 * faces with no data-path communication, bit-identical to the single-GPU rows; fcfv_export concatenates them into the global
 * CSR / Julia CSC.  One NCCL communicator per device (ncclCommInitAll), used for one grouped all-reduce of the six residual
 * sums per evaluation.  Supported: set_mesh, update_fields, set_* / run_* / get_local / get_nnz, compute_local, assemble,
 * compute_fcfv, schur, export (CSR0, CSC1_INT64), export_rhs, residuals (norms), run_pipeline, pass, element_values, ph_residual /
 * ph_update_p / ph_fusc / center_pressure, geometry / first_seen_faces / hilbert_elements (first device).  What hands out
 * device pointers or part-local state (device_csr, device_rhs, get_mesh, get_data, generate_structured, set_ownership, comm_*)
 * returns FCFV_ERR_STATE.  ngpus == 1 gives a plain single-GPU handle. */
int fcfv_create_multi(fcfv_t** h, int ngpus);
int fcfv_num_devices(fcfv_t* h);
/* What part `part` of a multi-GPU handle holds after fcfv_set_mesh: local / owned element and face counts (local - owned =
 * replicated ghosts) and the host seconds the partitioner took. */
int fcfv_partition_info(fcfv_t* h, int part, int64_t* nel_local, int64_t* nf_local, int64_t* nel_owned, int64_t* nf_owned,
                        double* seconds);
/* The same partition computed on the host alone (no GPU, no handle): global ids (1-based, ascending = the part's local
 * numbering) and ownership masks of part `part` of `nparts`.  Call with the arrays NULL for the sizes first. */
int fcfv_partition_part(int64_t nel, int64_t nf, const int64_t* e2f, int nparts, int part, int64_t* nel_local, int64_t* nf_local,
                        int64_t* elem_global, int64_t* face_global, uint8_t* elem_owned, uint8_t* face_owned);
int fcfv_destroy(fcfv_t* h);
const char* fcfv_last_error(fcfv_t* h);
int fcfv_version(void);

/* ---- mesh: the fields of FCFV_Mesh the path reads (src/CreateMeshFCFV.jl:10-44) --------
 * e2f nel x 3 (1-based), bc nf, Omega nel, nx/ny/Gamma nel x 3, ke/delta nel,
 * taue ntaue >= nel entries (examples/ViscousInclusionFCFV_v2.jl:127 passes nf of them).
 * Copies everything to the GPU (int32 connectivity, int8 tags) and builds the face->element
 * map.  A face referenced by more than two (element, local face) pairs is rejected. */
int fcfv_set_mesh(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, const int64_t* bc,
                  const double* Omega, const double* nx, const double* ny, const double* Gamma,
                  const double* ke, const double* delta, const double* taue, int64_t ntaue);

/* The drivers mutate mesh.ke, mesh.δ, mesh.τe, mesh.τ after mesh creation
 * (ViscousInclusionFCFV_v2.jl:44,127; SolKzFCFV.jl:25,149).  NULL leaves a field as it is. */
int fcfv_update_fields(fcfv_t* h, const double* ke, const double* delta, const double* taue,
                       int64_t ntaue, const double* tau);

/* Element/face ownership of a partitioned mesh (uint8 masks, NULL = everything owned), the
 * global sizes used to normalise the residual norms and, for FCFV_TAU_REFERENCE on a
 * partition, tau_ref[e] = global mesh.τ[global id of e] (NULL: use local τ[e]). */
int fcfv_set_ownership(fcfv_t* h, const uint8_t* elem_owned, const uint8_t* face_owned,
                       int64_t nel_global, int64_t nf_global, const double* tau_ref);

/* ---- a2: Ω, centroids, outward normals, face lengths ------------------------------------
 * src/CreateMeshFCFV.jl:142-156,180-211 (numbering 0: nodeA=[2 3 1], nodeB=[3 1 2]) and
 * :300-314,337-367 (numbering 1: nodeA=[3 2 1], nodeB=[1 3 2]).  Stand-alone: does not need
 * fcfv_set_mesh.  tau (nf) receives tau_r on every referenced face (:209,365); may be NULL. */
int fcfv_geometry(fcfv_t* h, int64_t nel, int64_t nf, int64_t nv, const int64_t* e2v,
                  const int64_t* e2f, const double* xv, const double* yv, const double* xf,
                  const double* yf, int numbering, double tau_r, double* Omega, double* xc,
                  double* yc, double* nx, double* ny, double* Gamma, double* tau);

/* ---- mesh preparation: first-seen face numbering (meshes/ReadMeshesDat.jl:36-88) -----------------
 * The face kernels gather the records of the elements next to consecutive faces; they are fast when face ids
 * follow the element order, which is how the reference numbers the faces of its own meshes (a face gets the next
 * id when the element loop first meets it).  For a mesh numbered otherwise this returns that numbering, computed
 * from e2f alone: new_of_old[f] (nf, 1-based) and e2f_new (nel x 3); faces no element references keep the ids after
 * all the others, in their old order.  The caller permutes its face arrays once (x_new[new_of_old[f]] = x_old[f]);
 * stand-alone, does not need fcfv_set_mesh.  Either output may be NULL. */
int fcfv_first_seen_faces(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, int64_t* new_of_old, int64_t* e2f_new);

/* ---- mesh preparation: space-filling-curve ELEMENT order -----------------------------------------
 * For a mesh whose elements come in an order without spatial coherence (the passes are ~3x slower on it):
 * old_of_new[k] (nel, 1-based) = the element that comes k-th when the elements are sorted along a Hilbert curve
 * through their centroids (xc, yc: FCFV_Mesh fields, src/CreateMeshFCFV.jl:23-24; 2^20 cells per direction over
 * the bounding box, elements in one cell keep their order).  The caller permutes its element arrays
 * (x_new[k] = x_old[old_of_new[k]]), remaps f2e / e2e, and then calls fcfv_first_seen_faces on the permuted
 * e2f.  Stand-alone, does not need fcfv_set_mesh.  Meshes from Triangle and the generators here do not need it. */
int fcfv_hilbert_elements(fcfv_t* h, int64_t nel, const double* xc, const double* yc, int64_t* old_of_new);

/* Diagnostic for the two preparation calls above, on the mesh that is set: element_span = mean over the faces of
 * |e_hi - e_lo| / nel (distance in memory between the two elements of a face; 0 for boundary faces), face_offset =
 * mean of |f/nf - e_lo/nel| (how far the face ids stray from the element order).  Meshes from Triangle, the .dat
 * fixtures and the generators give <= 0.01 for both; random numbering gives 1/3 and passes 3-7x slower. */
int fcfv_numbering_locality(fcfv_t* h, double* element_span, double* face_offset);

/* ---- a4: ComputeFCFV (src/DiscretisationFCFV_Stokes.jl:8-44) ---------------------------
 * Outputs alpha nel, beta nel x 2, Z nel x 2 x 2, tau nf (mesh.τ after the call: face f gets
 * τe of its highest-numbered adjacent element, faces without elements keep their value). */
int fcfv_compute_local(fcfv_t* h, int formulation, double A, const double* sex, const double* sey,
                       const double* VxDir, const double* VyDir, double* alpha, double* beta,
                       double* Z, double* tau);

/* The same with the reference's full argument list: ComputeFCFV (DiscretisationFCFV_Stokes.jl:8) also receives the four
 * Neumann arrays (and never reads them).  Any that are non-NULL are uploaded while alpha / beta / Z / tau come down --
 * the two directions of the link run concurrently -- so that fcfv_assemble / fcfv_residuals find them resident when
 * the upload cache is on.  Results identical to fcfv_compute_local. */
int fcfv_compute_local_ex(fcfv_t* h, int formulation, double A, const double* sex, const double* sey,
                          const double* VxDir, const double* VyDir, const double* sxxNeu, const double* syyNeu,
                          const double* sxyNeu, const double* syxNeu, double* alpha, double* beta, double* Z,
                          double* tau);

/* ComputeFCFV(mesh, ...) as ONE call (DiscretisationFCFV_Stokes.jl:8-44 on the mesh of CreateMeshFCFV.jl:10-44): what the
 * shims issue for it.
 *   e2f != NULL: the mesh goes up first (= fcfv_set_mesh(nel, nf, e2f, bc, Omega, nx, ny, Gamma, ke, delta, taue, ntaue));
 *   e2f == NULL: the mesh is resident, nel / nf and the geometry arguments are ignored and ke / delta / taue (each may be
 *                NULL = unchanged) are refreshed (= fcfv_update_fields);
 *   tau_in     : mesh.tau as it stands, or NULL; only read when the mesh has faces no element references (every other entry
 *                is overwritten, Discretisation:40);
 * then = fcfv_compute_local_ex.  Results are those of the separate calls, bit for bit.  When every array is page-locked (or
 * registered) host memory and nel >= 2^20 (FCFV_PIPELINE_MIN_NEL overrides), the call runs as a pipeline over 8 (16 from 2^24 elements) element
 * chunks: connectivity and tags first (the records are built on the device), VxDir / VyDir on the Dirichlet faces, then per
 * chunk geometry + tau_e + sources up | element kernel | alpha, beta, Z down on a second stream, and ke / delta (which the
 * element kernel does not read) and the Neumann arrays under the tail of the downloads -- both directions of PCIe stay busy.
 * Otherwise (pageable or device memory, small meshes, multi-GPU handles) it issues the separate calls. */
int fcfv_compute_fcfv(fcfv_t* h, int64_t nel, int64_t nf, const int64_t* e2f, const int64_t* bc, const double* Omega,
                      const double* nx, const double* ny, const double* Gamma, const double* ke, const double* delta,
                      const double* taue, int64_t ntaue, const double* tau_in, int formulation, double A,
                      const double* sex, const double* sey, const double* VxDir, const double* VyDir, const double* sxxNeu,
                      const double* syyNeu, const double* sxyNeu, const double* syxNeu, double* alpha, double* beta,
                      double* Z, double* tau);

/* ---- upload cache ---------------------------------------------------------------------------------
 * A driver hands the same host arrays to ComputeFCFV, ElementAssemblyLoop[NEW] and ComputeResidualsFCFV_Stokes_o1
 * (examples/SolKzFCFV.jl:150,155,205: VxDir / VyDir three times, the Neumann arrays and the sources twice, and
 * alpha / beta / Z go back in after they came out).  With the cache on, every device buffer remembers the host array
 * it mirrors -- uploaded from, or downloaded to (alpha, beta, Z, tau) --; a later upload of the same address and size
 * in the same epoch is skipped if a sampled fingerprint of the host contents (768 words at both ends + 4096 spread
 * over the array) still matches.  The caller vouches that arrays are not modified in place inside an epoch;
 * fcfv_new_epoch starts a new one (the shims do so at the top of every ComputeFCFV, i.e. once per driver pass),
 * fcfv_set_mesh and fcfv_generate_structured too.  fcfv_set_mesh itself always uploads.  Off by default. */
/* *flag = 1 when the resident mesh has faces no element references.  Without such faces ComputeFCFV overwrites every
 * entry of mesh.tau (DiscretisationFCFV_Stokes.jl:40), so the shims do not upload mesh.tau in front of it. */
int fcfv_has_orphan_faces(fcfv_t* h, int* flag);
int fcfv_upload_cache(fcfv_t* h, int on);
int fcfv_new_epoch(fcfv_t* h);
/* Bytes this handle has moved host -> device / device -> host so far and bytes the upload cache saved; reset != 0
 * zeroes the counters after reading.  Any pointer may be NULL. */
int fcfv_transfer_stats(fcfv_t* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* cache_hit_bytes, int reset);

/* ---- per-face boundary data through face lists ------------------------------------------------------
 * The reference reads VxDir / VyDir only behind the Bool factor (bc==1) (DiscretisationFCFV_Stokes.jl:27-37,191-194,
 * 321-325; ComputeResidualsFCFV_Stokes.jl:62-66) and the four Neumann arrays only through t*Xi with Xi = (bc==2)
 * (Discretisation:167-169,177-178; Residuals:57-61,69): of the nf values of each array only those on the Dirichlet /
 * Neumann faces enter the result.  fcfv_set_mesh therefore lists these faces (device pass over bc, sorted ids), and every
 * HOST array of per-face boundary data (fcfv_set_face_data and all calls that take VxDir ... SyxNeu) crosses PCIe as its
 * values on the listed faces (gather into a page-locked slot, one small copy, scatter on the device) instead of nf
 * doubles.  mode 0: off (whole arrays), 1 (default): lists when both are short (<= nf/16 and <= 2^18 faces -- every mesh
 * of the reference: boundary faces are O(sqrt(nf))), 2: always.  Entries of the device arrays on other faces keep what
 * they held (zeros after allocation); no kernel reads them, fcfv_get_data returns them as they are.  Device-pointer
 * arguments are copied whole as before.  Changing the mode on a handle with a mesh rebuilds the lists. */
int fcfv_sparse_face_data(fcfv_t* h, int mode);
/* *active = 1 when lists are in use for the resident mesh; their lengths (-1 when inactive); number of list uploads so
 * far.  Any pointer may be NULL. */
int fcfv_sparse_face_info(fcfv_t* h, int* active, int64_t* n_dirichlet, int64_t* n_neumann, int64_t* uploads);

/* ---- a5/a6/a7: ElementAssemblyLoop[NEW] + Sparsify (Discretisation:102-202,210-347,355-364)
 * alpha/beta/Z: NULL = use the device-resident result of the last fcfv_compute_local.
 * Uses mesh.τ as resident (fcfv_compute_local / fcfv_update_fields).  Result: device CSR of
 * Kuu (2nf x 2nf), Kup (2nf x nel), optional Muu, plus fu (2nf) and fp (nel).
 * seconds receives the device time of the build (the reference's `tsparse`). */
int fcfv_assemble(fcfv_t* h, int variant, int formulation, double A, const double* alpha,
                  const double* beta, const double* Z, const double* VxDir, const double* VyDir,
                  const double* sxxNeu, const double* syyNeu, const double* sxyNeu,
                  const double* syxNeu, int want_muu, int64_t* nnz_uu, int64_t* nnz_up,
                  int64_t* nnz_muu, double* seconds);

/* ---- one pass through host buffers: fcfv_compute_local + fcfv_assemble + fcfv_residuals in one call -------------
 * For drivers that can replace their three calls (ViscousInclusionFCFV_v2.jl:128,132,146) by one: every array crosses
 * PCIe once (the separate calls take VxDir / VyDir three times, the Neumann arrays and the sources twice) and
 * alpha / beta / Z / tau come down on a second stream while the remaining inputs go up.  Same results as the three
 * calls.  Outputs may be NULL; the Neumann arrays may be NULL (zeros on first use, else what is resident); norms: 6. */
int fcfv_pass(fcfv_t* h, int variant, int formulation, double A, int tau_mode, const double* sex, const double* sey,
              const double* VxDir, const double* VyDir, const double* sxxNeu, const double* syyNeu,
              const double* sxyNeu, const double* syxNeu, const double* Vxh, const double* Vyh, const double* Pe,
              double* alpha, double* beta, double* Z, double* tau, int want_muu, int64_t* nnz_uu, int64_t* nnz_up,
              int64_t* nnz_muu, double* norms);

/* ---- Powell-Hestenes Schur complement, fused into the assembly (src/SolversFCFV_Stokes.jl:23-25)
 * Kuusc = Kuu .- Kup*(Kppi*Kpu) with Kppi = spdiagm(gamma .* ke ./ Ω), Kpu = -Kup': the same face-to-face
 * rows as Kuu with one extra term per adjacent element, so the fill pass writes it next to Kuu / Kup
 * instead of a host-side sparse product.  Entries whose value is exactly +-0.0 are dropped (what
 * Julia's sparse broadcast `.-` does).  Block id FCFV_KUUSC for fcfv_export / fcfv_device_csr.
 * fcfv_schur re-runs the last fcfv_assemble (same loop, formulation, A) on the resident data. */
int fcfv_schur(fcfv_t* h, double gamma, int64_t* nnz_sc);
int fcfv_run_assemble_schur(fcfv_t* h, int variant, int formulation, double A, double gamma);

/* Bytes per row-pointer entry of the device CSR (4 or 8). */
int fcfv_index_width(fcfv_t* h);
/* Sizes of an assembled block. */
int fcfv_block_size(fcfv_t* h, int block, int64_t* nrows, int64_t* ncols, int64_t* nnz);
/* Copy a block out.  FCFV_CSR0: ptr has nrows+1 entries of fcfv_index_width bytes, idx int32,
 * 0-based.  FCFV_CSC1_INT64: ptr ncols+1, idx nnz, both int64 and 1-based, rows ascending in a
 * column -- the fields of Julia's SparseMatrixCSC{Float64,Int64}.  The transposition runs on the device
 * (count / scan / scatter / per-column sort by row: deterministic), the host only receives the result. */
int fcfv_export(fcfv_t* h, int block, int layout, void* ptr, void* idx, double* val);
int fcfv_export_rhs(fcfv_t* h, double* fu, double* fp);
/* Device pointers of the resident CSR for GPU consumers (valid until the next assemble). */
int fcfv_device_csr(fcfv_t* h, int block, void** rowptr, void** colidx, void** val);
int fcfv_device_rhs(fcfv_t* h, void** fu, void** fp);

/* ---- a8: ComputeResidualsFCFV_Stokes_o1 (src/ComputeResidualsFCFV_Stokes.jl:5-108) ------
 * norms[6] = norm(F)/sqrt(length(F)) for F_eq1, F_eq2, F_eq3, F_nodes_x, F_nodes_y, F_glob2
 * (print order :101-106).  alpha/beta/Z NULL = resident.  Optional outputs F_nodes_x/y (nf),
 * F_glob2 (nel). */
int fcfv_residuals(fcfv_t* h, int formulation, int tau_mode, const double* Vxh, const double* Vyh,
                   const double* Pe, const double* alpha, const double* beta, const double* Z,
                   const double* sex, const double* sey, const double* VxDir, const double* VyDir,
                   const double* SxxNeu, const double* SyyNeu, const double* SxyNeu,
                   const double* SyxNeu, double* norms, double* F_nodes_x, double* F_nodes_y,
                   double* F_glob2);

/* ---- ComputeElementValues (src/DiscretisationFCFV_Stokes.jl:52-94) --------------------------
 * Element velocities and deviatoric stresses from the face velocities.  Vxh/Vyh NULL = resident
 * solution (fcfv_set_solution), alpha/beta/Z NULL = resident local coefficients; uses mesh.τ as
 * resident.  Outputs nel doubles each (NULL skips).  The reference's Pe, VxDir, VyDir and
 * Formulation arguments are unused there (:52) and have no counterpart here. */
int fcfv_element_values(fcfv_t* h, const double* Vxh, const double* Vyh, const double* alpha, const double* beta,
                        const double* Z, double* Vxe, double* Vye, double* Txxe, double* Tyye, double* Txye);

/* ---- a9: Powell-Hestenes residual / updates on the resident CSR --------------------------
 * src/SolversFCFV_Stokes.jl:40-42: ru = fu - Kuu*u - Kup*p, rp = fp - Kpu*u (Kpu = -Kup'),
 * nrm = {norm(ru), norm(rp)};  :23,49: p += (gamma*ke/Ω) .* (fp - Kpu*u);
 * :47: fusc = fu - Kup*((gamma*ke/Ω).*fp + p).   u has 2nf entries, p nel. */
int fcfv_ph_residual(fcfv_t* h, const double* u, const double* p, double* ru, double* rp, double* nrm);
int fcfv_ph_update_p(fcfv_t* h, double gamma, const double* u, double* p);
int fcfv_ph_fusc(fcfv_t* h, double gamma, const double* p, double* fusc);
/* src/SolversFCFV_Stokes.jl:20 (:CoupledBackslash): Pe .= Pe .- mean(Pe); the sum runs over owned elements
 * and is all-reduced over the communicator if one is attached (mean of the global pressure). */
int fcfv_center_pressure(fcfv_t* h, double* Pe, double* mean);

/* ---- device-resident pipeline (what bench.py times as `value`) ---------------------------
 * Upload once, then run the kernels on resident data without host traffic.  The host-buffer
 * entry points above are compositions of these. */
int fcfv_set_sources(fcfv_t* h, const double* sex, const double* sey);
int fcfv_set_face_data(fcfv_t* h, const double* VxDir, const double* VyDir, const double* sxxNeu,
                       const double* syyNeu, const double* sxyNeu, const double* syxNeu);
int fcfv_set_solution(fcfv_t* h, const double* Vxh, const double* Vyh, const double* Pe);
int fcfv_set_local(fcfv_t* h, const double* alpha, const double* beta, const double* Z);
int fcfv_run_local(fcfv_t* h, int formulation, double A);
int fcfv_run_assemble(fcfv_t* h, int variant, int formulation, double A, int want_muu);
/* sumsq[6]: sums of squares over owned elements/faces (device-reduced, then all-reduced over
 * the communicator if one is attached); norms[6] as fcfv_residuals.  Either may be NULL. */
int fcfv_run_residuals(fcfv_t* h, int formulation, int tau_mode, double* sumsq, double* norms);
/* fcfv_run_local + fcfv_run_assemble (no Muu) + fcfv_run_residuals as one CUDA graph, captured on first use and
 * replayed while loop / formulation / A / tau mode / mesh / buffers are unchanged: removes the launch latency that
 * bounds a pass on small meshes (BASELINE configs 1-2).  Results are read with the fcfv_get_* / fcfv_export calls. */
int fcfv_run_pipeline(fcfv_t* h, int variant, int formulation, double A, int tau_mode);
int fcfv_get_local(fcfv_t* h, double* alpha, double* beta, double* Z, double* tau);
int fcfv_get_nnz(fcfv_t* h, int64_t* nnz_uu, int64_t* nnz_up, int64_t* nnz_muu);
/* run_* calls enqueue on the handle's stream and return without synchronising when async != 0. */
int fcfv_set_async(fcfv_t* h, int async);
int fcfv_sync(fcfv_t* h);
void* fcfv_stream(fcfv_t* h);               /* cudaStream_t */
/* Optional: page-locks host arrays the caller passes again and again (mesh fields, boundary data, outputs), so
 * that the copies of the host-buffer entries run at PCIe speed instead of through the driver's staging buffer
 * (Julia / NumPy arrays are pageable).  The range must stay allocated until fcfv_host_unregister. */
int fcfv_host_register(fcfv_t* h, void* ptr, int64_t bytes);
int fcfv_host_unregister(fcfv_t* h, void* ptr);
/* Per-kernel device timing: when on, every hot-path kernel launch is bracketed by CUDA events on the
 * handle's stream; fcfv_get_timers returns the accumulated milliseconds and launch counts per
 * kernel (arrays of fcfv_num_timers() entries, names via fcfv_timer_name).  set_timing resets. */
int fcfv_set_timing(fcfv_t* h, int on);
int fcfv_num_timers(void);
const char* fcfv_timer_name(int k);
int fcfv_get_timers(fcfv_t* h, double* ms, int64_t* calls);
int64_t fcfv_kernel_launches(fcfv_t* h);    /* kernels launched by this handle so far */
/* Host arrays that are neither page-locked nor registered (plain Julia / NumPy arrays) are copied by up to sixteen worker
 * threads through page-locked staging buffers when a transfer is 8 MiB or larger (FCFV_STAGED_MIN_BYTES overrides the
 * threshold, 0 disables); the call still returns with the host side of the copy complete.  Number of such copies: */
int64_t fcfv_staged_copies(fcfv_t* h);
int64_t fcfv_pipelined_calls(fcfv_t* h);    /* fcfv_compute_fcfv calls that ran as a chunk pipeline */
/* Bounds-checked build of the library (make -C fcfv_nme23_b200/csrc variant NAME=bounds DEFS=-DFCFV_BOUNDS_CHECK=1): number of
 * out-of-range shared-memory staging indices, flushed CSR segments, gathered element / face ids and flux slots seen so far on
 * all devices of the process; -1 in a normal build (the checks are compiled out). */
int64_t fcfv_debug_bounds_violations(void);
int64_t fcfv_device_bytes(fcfv_t* h);       /* device memory currently held by the handle */

/* ---- synthetic structured mesh S(n) generated on the device (bench / large tests) --------
 * Quad rows row0 <= qy < row1 of the n x n criss-cross triangulation of the unit square with
 * the closed-form first-seen face numbering (meshes/ReadMeshesDat.jl:36-88), geometry from the
 * a2 kernel, then one of the closed-form field settings of SURVEY.md §8(d):
 * setting 0 = SolKz (ke = exp(13.8 yc), sey = sin(1.6 pi yc) cos(3 pi xc), τe = tau_r*max(ke,1),
 * all boundary Dirichlet), 1 = inclusion (ke = eta_incl inside radius 1/6, interface faces
 * tagged -1, τe = tau_r, south boundary Neumann when neumann_south != 0).
 * With ghost_below != 0 the quad row row0-1 is included as non-owned ghost elements.
 * Installs the mesh, ownership, sources, face data and a smooth (û, p) in the handle. */
int fcfv_generate_structured(fcfv_t* h, int64_t n, int64_t row0, int64_t row1, int ghost_below,
                             int setting, double tau_r, double eta_incl, int neumann_south);
/* Copies the resident mesh / data back (any pointer may be NULL); sizes via fcfv_mesh_size. */
int fcfv_mesh_size(fcfv_t* h, int64_t* nel, int64_t* nf);
int fcfv_get_mesh(fcfv_t* h, int64_t* e2f, int64_t* bc, double* Omega, double* nx, double* ny,
                  double* Gamma, double* ke, double* delta, double* taue, double* tau);
int fcfv_get_ownership(fcfv_t* h, uint8_t* elem_owned, uint8_t* face_owned, int64_t* nel_global, int64_t* nf_global);
int fcfv_get_data(fcfv_t* h, double* sex, double* sey, double* VxDir, double* VyDir, double* sxxNeu,
                  double* syyNeu, double* sxyNeu, double* syxNeu, double* Vxh, double* Vyh, double* Pe);

/* ---- multi-GPU: one process per GPU, NCCL only for the norm reductions --------------------
 * fcfv_comm_unique_id fills 128 bytes (ncclUniqueId) on rank 0; broadcast them by any means,
 * then every rank calls fcfv_comm_init.  libnccl.so.2 is loaded at run time (dlopen). */
int fcfv_comm_unique_id(fcfv_t* h, void* id128);
int fcfv_comm_init(fcfv_t* h, const void* id128, int rank, int nranks);
/* In-place sum over ranks of n doubles (host or device pointer). */
int fcfv_comm_allreduce_sum(fcfv_t* h, double* data, int n);

#ifdef __cplusplus
}
#endif
#endif /* FCFV_B200_H */
