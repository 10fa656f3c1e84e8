/*
 * libqrochet_b200.so -- C-ABI of the B200-native backend for Qrochet.jl's tensor-network hot path.
 *
 * The reference (bsc-quantic/Qrochet.jl v0.1.2) has NO FFI of its own: its only device hook is
 * `ext/QrochetAdaptExt.jl:7-9` (Adapt.adapt_structure for Quantum/Product/Chain); after `adapt`
 * every numeric call dispatches on the array type inside `Tenet.Tensor{T,N,A}`.  The entry points
 * below are therefore the calls a Julia package extension (julia/ext/QrochetB200Ext.jl, modelled on
 * QrochetAdaptExt) binds with `ccall`; each one cites the reference call site(s) it replaces.
 * INTEGRATION.md shows the Julia-side stubs.
 *
 * Conventions
 *   - every function returns int32: 0 = ok, <0 = error class (QB200_E_*); never throws across the
 *     boundary; message via qb200_last_error(ctx).
 *   - tensors are dense, COLUMN-MAJOR (first mode fastest, as Julia arrays), extents are int64,
 *     modes are caller-chosen int32 labels (the Symbol <-> int32 map lives on the Julia side).
 *   - dtype: QB200_C128 (ComplexF64, primary; FP64 tensor-core DMMA arithmetic) and QB200_C64 (ComplexF32: native
 *     float2 storage; qb200_contract runs on tcgen05 (TF32, 3xTF32 split, TMEM accumulators: FP32-level accuracy), the
 *     HBM-bound helpers run natively, qb200_qr / qb200_svd factorise in FP64 and narrow the factors).  The operands
 *     of one call must share the complex type.  QB200_F64 holds Schmidt vectors; QB200_F32 vectors are accepted and
 *     held widened to FP64 (upload / download convert).  The fused qb200_mps_* chains store ComplexF64 or ComplexF32 sites
 *     (qb200_mps_create_typed) and compute in FP64; qb200_tn_* is ComplexF64.
 *   - one context = one device + one stream; calls on a context are serialised by the caller.
 *     Kernels are asynchronous on that stream; only *_download, scalar-returning calls and calls
 *     with `kept` outputs synchronise.
 *   - handles are owned by the caller (Julia attaches a finalizer calling *_free).
 */
#ifndef QROCHET_B200_H
#define QROCHET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qb200_ctx qb200_ctx;
typedef struct qb200_tensor qb200_tensor;
typedef struct qb200_mps qb200_mps;
typedef struct qb200_tnplan qb200_tnplan;

enum { QB200_C128 = 0, QB200_C64 = 1, QB200_F64 = 2, QB200_F32 = 3 };

enum {
    QB200_OK = 0,
    QB200_E_INVALID = -1,     /* bad argument (ArgumentError on the Julia side) */
    QB200_E_CUDA = -2,        /* CUDA runtime failure */
    QB200_E_UNSUPPORTED = -3, /* dtype / shape not implemented */
    QB200_E_NOSPECTRUM = -4,  /* MissingSchmidtCoefficientsException (src/Ansatz.jl:91-99) */
    QB200_E_NOCONVERGE = -5,  /* Jacobi SVD hit its sweep limit */
    QB200_E_COMM = -6         /* NCCL failure */
};

#define QB200_MAX_RANK 64
/* largest physical dimension of a site accepted by the fused chain entry points (gates are staged through one pinned page) */
#define QB200_MAX_PHYS 8

/* ---- context ------------------------------------------------------------------------------ */
int32_t qb200_create(int32_t device, qb200_ctx** ctx);
int32_t qb200_destroy(qb200_ctx* ctx);
const char* qb200_last_error(qb200_ctx* ctx);
/* run on an externally owned cudaStream_t (e.g. torch's current stream); NULL = own stream */
int32_t qb200_set_stream(qb200_ctx* ctx, void* cuda_stream);
int32_t qb200_synchronize(qb200_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t qb200_launch_count(qb200_ctx* ctx);
/* CUDA-event timer on the context's stream: begin(); ...; end() returns milliseconds (synchronises) */
int32_t qb200_timer_begin(qb200_ctx* ctx);
int32_t qb200_timer_end(qb200_ctx* ctx, double* ms);

/* phase profiler (CUDA events around the library's internal phases, used by bench.py for the roofline):
 * phases: 0 theta GEMM, 1 gate, 2 SVD (whole), 3 Jacobi gram, 4 Jacobi evd, 5 Jacobi update, 6 SVD emit,
 * 7 mode scale, 8 QR, 9 sliced-TN GEMM, 10 / 11 / 12 low-precision stage of the mixed-precision Jacobi SVD (Gram, update,
 * FP64 orthonormalisation + application of the accumulated rotations).  read() synchronises, returns per-phase launch count, summed
 * milliseconds and summed algorithmic work (flops, or bytes for HBM-bound phases), and resets. */
int32_t qb200_prof_enable(qb200_ctx* ctx, int32_t on);
int32_t qb200_prof_read(qb200_ctx* ctx, int32_t nphases, int64_t* counts, double* ms, double* work);

/* ---- tensors: the device array type behind Tenet.Tensor{T,N,B200Array} ---------------------- */
/* replaces Adapt.adapt_storage(Array -> device) reached from ext/QrochetAdaptExt.jl:7-9 */
int32_t qb200_tensor_alloc(qb200_ctx* ctx, int32_t dtype, int32_t rank, const int64_t* extents, qb200_tensor** out);
/* borrow device memory owned by someone else (torch / CUDA.jl); never freed by the library */
int32_t qb200_tensor_wrap(qb200_ctx* ctx, int32_t dtype, int32_t rank, const int64_t* extents, void* device_ptr,
                          qb200_tensor** out);
int32_t qb200_tensor_free(qb200_ctx* ctx, qb200_tensor* t);
int32_t qb200_tensor_upload(qb200_ctx* ctx, qb200_tensor* t, const void* host);   /* host -> device */
int32_t qb200_tensor_download(qb200_ctx* ctx, const qb200_tensor* t, void* host); /* device -> host, syncs */
int32_t qb200_tensor_rank(const qb200_tensor* t);
int32_t qb200_tensor_dtype(const qb200_tensor* t);
int64_t qb200_tensor_extent(const qb200_tensor* t, int32_t mode_pos);
void* qb200_tensor_data(const qb200_tensor* t);
/* `copy` / `reshape` (src/Quantum.jl:87-90, Chain.jl:431,449) */
int32_t qb200_tensor_copy(qb200_ctx* ctx, const qb200_tensor* src, qb200_tensor* dst);
int32_t qb200_tensor_reshape(qb200_ctx* ctx, qb200_tensor* t, int32_t rank, const int64_t* extents);

/* ---- K1: pairwise contraction  (Tenet.contract(a,b;dims); call sites Chain.jl:372,602,616,636,
 *      682,734,747 and every node of an EinExprs path, examples/distributed.jl:89) --------------
 * C[modesC] = alpha * sum_{k} op(A)[modesA] * op(B)[modesB] + beta * C.
 * A mode in A, B and C is a batch (kept, `dims=()`-style) mode; in A and B only: summed; in exactly
 * one of A/B and in C: free.  The index permutation is fused into the tile loads (no TTGT copy). */
int32_t qb200_contract(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* modesA, int32_t conjA,
                       const qb200_tensor* B, const int32_t* modesB, int32_t conjB, qb200_tensor* C,
                       const int32_t* modesC, const double alpha[2], const double beta[2]);

/* ---- K2/K3: mode scale  (contract(x, Λ; dims=()) and pinv(Diagonal(λ); atol), Chain.jl:322,325,
 *      484,491,710,713).  out = A .* f(vec) along mode_pos; f(v) = v, or (|v|>atol ? 1/v : 0) when
 *      inverse != 0.  out may alias A. vec is QB200_F64 of length extent(mode_pos). */
int32_t qb200_scale_mode(qb200_ctx* ctx, const qb200_tensor* A, int32_t mode_pos, const qb200_tensor* vec,
                         int32_t inverse, double atol, qb200_tensor* out);
/* ---- K6: slice!(tn, ind, 1:count) (Chain.jl:419): keep the first `count` entries of a mode ---- */
int32_t qb200_slice_mode(qb200_ctx* ctx, const qb200_tensor* A, int32_t mode_pos, int64_t count, qb200_tensor* out);
/* view(tn, ind => value): drop a mode at a fixed position (examples/distributed.jl:72,82) */
int32_t qb200_select_mode(qb200_ctx* ctx, const qb200_tensor* A, int32_t mode_pos, int64_t index, qb200_tensor* out);
/* ---- K7: conj (src/Quantum.jl:105) and permutedims ---------------------------------------- */
int32_t qb200_conj(qb200_ctx* ctx, const qb200_tensor* A, qb200_tensor* out);
int32_t qb200_permute(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* perm, qb200_tensor* out);
/* ---- K8: norm / normalize (Chain.jl:534,654; Product.jl:59,73): result = sqrt(sum |a|^2), syncs */
int32_t qb200_norm2(qb200_ctx* ctx, const qb200_tensor* A, double* result);
int32_t qb200_scale(qb200_ctx* ctx, qb200_tensor* A, const double factor[2]);

/* ---- K4: thin QR of the (left | right) matricisation (LinearAlgebra.qr(::Tensor; left_inds,
 *      right_inds, virtualind), Chain.jl:367).  A has `rank` modes; the first `nleft` entries of
 *      `order` are the positions (in A) of the left modes, the rest the right modes.  Q gets extents
 *      [left..., k], R gets [k, right...], k = min(rows, cols). */
int32_t qb200_qr(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* order, int32_t nleft, qb200_tensor* Q,
                 qb200_tensor* R);

/* ---- K5: thin SVD of the same matricisation (LinearAlgebra.svd(::Tensor; ...), Chain.jl:365,645,
 *      705) with the truncation rule of truncate! (Chain.jl:404-417) applied on the spot:
 *      kept = #{ i < min(k, maxdim) : s[i] > threshold }, sigma sorted descending.
 *      maxdim <= 0 means no limit; threshold < 0 means keep everything.
 *      U [left..., kept], S (F64) [kept], Vc = conj(V) [right..., kept]; the caller allocates U, S, Vc
 *      with k = min(rows, cols) columns and the library shrinks their last extent to `kept`.
 *      discarded_weight = sum_{i >= kept} s[i]^2.  Synchronises (kept is returned). */
int32_t qb200_svd(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* order, int32_t nleft, int64_t maxdim,
                  double threshold, qb200_tensor* U, qb200_tensor* S, qb200_tensor* Vc, int64_t* kept,
                  double* discarded_weight);
/* statistics of the last qb200_svd on this context: Jacobi sweeps executed */
int32_t qb200_svd_last_sweeps(qb200_ctx* ctx);
/* totals over the context and its worker streams since creation: SVDs factorised, Jacobi sweeps executed */
int32_t qb200_svd_totals(qb200_ctx* ctx, int64_t* calls, int64_t* sweeps);

/* ---- fused MPS path (Chain.jl:460-752 on a device-resident open-boundary MPS) -------------------
 * A qb200_mps owns n site tensors in the private layout (l, o, r) column-major plus the Schmidt
 * vectors (device + host mirror).  form: 0 = plain (no Λ), 1 = Vidal Γ/Λ (after canonize!). */
int32_t qb200_mps_create(qb200_ctx* ctx, int32_t nsites, qb200_mps** out);
/* the same with the storage type of the site tensors chosen: QB200_C128, or QB200_C64 = ComplexF32 sites in HBM (half
 * the footprint; `rand(...; eltype = ComplexF32)`, Chain.jl:226-227).  The fused chains compute in FP64 by explicit
 * choice (the Jacobi SVD is 95 % of the work and needs FP64 for its 1e-12 bar): a ComplexF32 chain is widened for the
 * duration of a call and its results are rounded to ComplexF32 once, at the end (parity bar 1e-5, north star). */
int32_t qb200_mps_create_typed(qb200_ctx* ctx, int32_t nsites, int32_t dtype, qb200_mps** out);
int32_t qb200_mps_dtype(const qb200_mps* mps);
int32_t qb200_mps_free(qb200_ctx* ctx, qb200_mps* mps);
int32_t qb200_mps_copy(qb200_ctx* ctx, const qb200_mps* src, qb200_mps** out);
/* site (0-based) from a host array with extents (chi_l, p, chi_r) column-major */
int32_t qb200_mps_set_site(qb200_ctx* ctx, qb200_mps* mps, int32_t site, int64_t chil, int64_t p, int64_t chir,
                           const void* host_c128);
/* the same from host data of any element type of the reference's `eltype` keyword: QB200_C128 / C64 / F64 / F32 (real
 * chains -- the reference's default eltype is Float64 -- are accepted and held as complex); converted on the device */
int32_t qb200_mps_set_site_typed(qb200_ctx* ctx, qb200_mps* mps, int32_t site, int32_t host_dtype, int64_t chil,
                                 int64_t p, int64_t chir, const void* host);
int32_t qb200_mps_site_dims(const qb200_mps* mps, int32_t site, int64_t dims[3]);
int32_t qb200_mps_get_site(qb200_ctx* ctx, const qb200_mps* mps, int32_t site, void* host_c128);
/* host_dtype = QB200_C128 or QB200_C64, whatever the storage type */
int32_t qb200_mps_get_site_typed(qb200_ctx* ctx, const qb200_mps* mps, int32_t site, int32_t host_dtype, void* host);
int32_t qb200_mps_set_lambda(qb200_ctx* ctx, qb200_mps* mps, int32_t bond, int64_t n, const double* host);
/* returns length; host may be NULL to query; -1 (as length 0 + QB200_E_NOSPECTRUM) if absent */
int32_t qb200_mps_get_lambda(qb200_ctx* ctx, const qb200_mps* mps, int32_t bond, double* host, int64_t* n);
int32_t qb200_mps_form(const qb200_mps* mps);
/* declare the form of arrays set by hand (adapt of an already canonical Chain): 0 plain, 1 Vidal, 2 mixed */
int32_t qb200_mps_set_form(qb200_mps* mps, int32_t form);
/* canonize! (Chain.jl:469-497): QR sweep <-, SVD sweep ->, Γ = A Λ^-1 (pinv atol 1e-64) */
int32_t qb200_mps_canonize(qb200_ctx* ctx, qb200_mps* mps);
/* mixed_canonize! (Chain.jl:509-524): center is 0-based, Λ left on bond (center-1, center) */
int32_t qb200_mps_mixed_canonize(qb200_ctx* ctx, qb200_mps* mps, int32_t center);
/* truncate! (Chain.jl:390-422) on a bond holding Λ */
int32_t qb200_mps_truncate(qb200_ctx* ctx, qb200_mps* mps, int32_t bond, int64_t maxdim, double threshold,
                           int64_t* kept);
/* evolve!(ψ, gate; threshold, maxdim, iscanonical, renormalize) for a 2-site gate on sites (bond, bond+1)
 * (evolve_2site!, Chain.jl:606-661): gate = (p1 p2)^2 c128 numbers (16 for qubits), array dims (o1,o2,i1,i2)
 * column-major (Dense.jl:21-34).  maxdim <= 0 / threshold < 0 = the reference's `nothing`.
 *   iscanonical != 0: contract_2sitewf! / unpack_2sitewf! (Chain.jl:669-722) as ONE fused chain: Λ-scale -> θ GEMM ->
 *     gate -> Jacobi SVD -> truncate! -> Λ^-1 scale (pinv atol 1e-32); the neighbouring Schmidt vectors are used where
 *     they exist; renormalize normalises the new Schmidt vector (Chain.jl:653-654).
 *   iscanonical == 0 (the reference's default): contract!(tn, bond) of the two sites and the Schmidt vector on the bond
 *     if there is one, svd! leaves U, s, V^H (Chain.jl:615,645); renormalize = normalize!(ψ, sitel) =
 *     mixed_canonize! + normalisation (Chain.jl:655-656,532-536; QB200_E_INVALID on the first bond, as
 *     "Cannot right-canonize left-most tensor", Chain.jl:344). */
int32_t qb200_mps_evolve2(qb200_ctx* ctx, qb200_mps* mps, int32_t bond, const void* gate_c128, int64_t maxdim,
                          double threshold, int32_t renormalize, int32_t iscanonical, int64_t* kept,
                          double* discarded_weight);
/* One TEBD layer: nb two-site gates (concatenated) on pairwise non-adjacent bonds -- the user-level
 * loop `for bond in odd_bonds evolve!(psi, G[bond]; ...)`.  The bond updates are independent
 * units and run concurrently on worker streams; results are identical to nb calls of qb200_mps_evolve2.
 * kept / discarded_weight: arrays of nb (may be NULL). */
int32_t qb200_mps_evolve2_layer(qb200_ctx* ctx, qb200_mps* mps, int32_t nb, const int32_t* bonds, const void* gates_c128,
                                int64_t maxdim, double threshold, int32_t renormalize, int32_t iscanonical,
                                int64_t* kept, double* discarded_weight);
/* A circuit of nearest-neighbour two-site gates in program order -- the user loop
 * `for (G, bond) in circuit evolve!(psi, G; ...)` (Chain.jl:543-584 called repeatedly; the Quac/Yao front-ends of
 * SURVEY.md §8 f2 produce such lists).  Bonds may repeat and touch: an update starts as soon as the earlier updates
 * on its two sites are complete, so consecutive TEBD layers overlap.  Results are identical to nops calls of
 * qb200_mps_evolve2 in the given order.  kept / discarded_weight: arrays of nops (may be NULL). */
int32_t qb200_mps_evolve2_circuit(qb200_ctx* ctx, qb200_mps* mps, int32_t nops, const int32_t* bonds,
                                  const void* gates_c128, int64_t maxdim, double threshold, int32_t renormalize,
                                  int32_t iscanonical, int64_t* kept, double* discarded_weight);
/* evolve_1site! (Chain.jl:586-603): gate = p*p c128 numbers (o, i) column-major */
int32_t qb200_mps_evolve1(qb200_ctx* ctx, qb200_mps* mps, int32_t site, const void* gate_c128);
/* ---- MPO x MPS (no function exists in the reference: composed from MPO(arrays) Chain.jl:133-172,
 *      merge Quantum.jl:330-348 and contract).  The MPO is passed as host arrays, one per site, in the reference's
 *      default order (o, i, l, r) column-major (Chain.jl:34), concatenated; dl/dr = bond dimensions per site
 *      (dl[0] = dr[n-1] = 1). */
/* B_s[(la,lw), o, (ra,rw)] = sum_i W_s[o,i,lw,rw] A_s[la,i,ra]; result is a plain chain with bonds chi*D */
int32_t qb200_mps_apply_mpo(qb200_ctx* ctx, qb200_mps* mps, const int64_t* dl, const int64_t* dr,
                            const void* sites_c128);
/* canonize! with truncate!(maxdim, threshold) applied to each bond right after its SVD (Vidal form result) */
int32_t qb200_mps_compress(qb200_ctx* ctx, qb200_mps* mps, int64_t maxdim, double threshold);
/* <psi|H|psi> = contract(merge(psi, H, psi')), un-normalised; left-environment sweep resident in HBM */
int32_t qb200_mps_expect_mpo(qb200_ctx* ctx, const qb200_mps* mps, const int64_t* dl, const int64_t* dr,
                             const void* sites_c128, double result[2]);
/* overlap(a,b) = <b|a> (Chain.jl:737-748) by a left-environment sweep resident in HBM */
int32_t qb200_mps_overlap(qb200_ctx* ctx, const qb200_mps* a, const qb200_mps* b, double result[2]);
/* expect(ψ, [O]) for a batch of single-site observables (Chain.jl:724-735), un-normalised;
 * sites[nobs] 0-based, ops = nobs * p*p c128, results = nobs complex.  Left/right environments are
 * built once and reused by every observable. */
int32_t qb200_mps_expect1_batch(qb200_ctx* ctx, const qb200_mps* mps, int32_t nobs, const int32_t* sites,
                                const void* ops_c128, double* results);
/* expect(ψ, observables) with the reference's exact composition (Chain.jl:724-735): ϕ = copy(ψ); evolve!(ϕ, O) for
 * every observable in order (default keywords: no truncation, iscanonical = false); result = contract(merge(ϕ, ψ')) =
 * <ψ|O_k ... O_1|ψ>, un-normalised.  nlanes[i] is 1 or 2; sites[i] = 0-based (left) site of observable i; ops = the
 * operator arrays concatenated, dims (o, i) resp. (o1,o2,i1,i2) column-major. */
int32_t qb200_mps_expect(qb200_ctx* ctx, const qb200_mps* mps, int32_t nobs, const int32_t* nlanes,
                         const int32_t* sites, const void* ops_c128, double result[2]);

/* ---- sliced contraction of a general tensor network (examples/distributed.jl:46-101) -------------
 * The network is given as `ntensors` leaves (rank, modes, extents concatenated).  The planner is deterministic and
 * replaces the reference's `transform!(tn, ContractSimplification())` + `einexpr(tn; optimizer = HyPar(...))` +
 * `findslices(SizeScorer(), path; size)` (examples/distributed.jl:29-46): simplification, multi-start greedy,
 * sub-tree reconfiguration (local search) and slicing until the largest intermediate has at most `max_elements`
 * entries (rules in csrc/tn_plan.cu; mirrored bit-exactly by oracle/circuit.py::plan). */
int32_t qb200_tn_plan(qb200_ctx* ctx, int32_t ntensors, const int32_t* ranks, const int32_t* modes,
                      const int64_t* extents, int64_t max_elements, qb200_tnplan** out);
/* the same with the optimiser chosen: 0 = one greedy tree, no simplification, no local search (the round-1 planner, kept
 * as the "before" column of the benchmark), 1 = the full planner (what qb200_tn_plan uses) */
int32_t qb200_tn_plan_opt(qb200_ctx* ctx, int32_t ntensors, const int32_t* ranks, const int32_t* modes,
                          const int64_t* extents, int64_t max_elements, int32_t optimizer, qb200_tnplan** out);
int32_t qb200_tn_plan_free(qb200_ctx* ctx, qb200_tnplan* plan);
/* queries: number of slices, sliced modes (array may be NULL to get the count), flops per slice (8 * complex MACs, the
 * EinExprs `flops` figure x 8, of the tree nodes that depend on a cut index), flops of the slice-invariant nodes
 * (executed once per contract call), largest intermediate (elements) */
int64_t qb200_tn_plan_nslices(const qb200_tnplan* plan);
int32_t qb200_tn_plan_sliced_modes(const qb200_tnplan* plan, int32_t* modes_out);
double qb200_tn_plan_flops_per_slice(const qb200_tnplan* plan);
double qb200_tn_plan_flops_invariant(const qb200_tnplan* plan);
int64_t qb200_tn_plan_max_intermediate(const qb200_tnplan* plan);
/* contraction order as pairs (i, j) -> new id ntensors + step; returns number of steps */
int32_t qb200_tn_plan_path(const qb200_tnplan* plan, int32_t* pairs_out);
/* contract slices first_slice, first_slice+stride, ... (< nslices), first cut index fastest, and ADD
 * the scalar results into acc (device-side accumulation, one download at the end). */
int32_t qb200_tn_contract_sliced(qb200_ctx* ctx, qb200_tnplan* plan, qb200_tensor* const* leaves,
                                 int64_t first_slice, int64_t stride, double acc[2]);

/* ---- NCCL sum of the per-rank partial results (examples/distributed.jl:101) ------------------- */
int32_t qb200_comm_unique_id(char id_out[128]);
int32_t qb200_comm_init(qb200_ctx* ctx, int32_t nranks, int32_t rank, const char id[128]);
int32_t qb200_comm_allreduce_sum(qb200_ctx* ctx, double* host_values, int32_t count);
/* one-to-all replication over NVLink (ncclBroadcast, in place): a tensor (the `@everywhere` broadcast of the leaf
 * tensors, examples/distributed.jl:58-64) ... */
int32_t qb200_comm_broadcast(qb200_ctx* ctx, qb200_tensor* t, int32_t root);
/* ... and a whole device-resident MPS (batched independent expectation values: the state is replicated once, the
 * observables are dealt i mod W, one sum gathers the values).  Root passes its chain in *inout, every other rank
 * passes NULL and receives a new handle. */
int32_t qb200_mps_broadcast(qb200_ctx* ctx, qb200_mps** inout, int32_t root);
int32_t qb200_comm_destroy(qb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
