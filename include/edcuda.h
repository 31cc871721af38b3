/*
 * edcuda.h -- C ABI of libedcuda.so, the B200 (sm_100a) engine behind
 * ExactDiagonalization.jl's Hamiltonian-application path.
 *
 * The reference (v0.14.3, /root/reference) has no FFI of its own: its boundary is the
 * Julia method surface.  Each entry point below names the reference method(s) it
 * replaces (file:line under /root/reference/src).  The Julia shim that binds them by
 * `ccall` is exactdiagonalization.jl_b200/julia/EDCuda.jl; the Python mirror used by
 * the tests is exactdiagonalization.jl_b200/edcuda/.
 *
 * Conventions
 *   - every call returns an int status (ED_OK == 0); ed_last_error() gives the message
 *     of the calling thread's last failure.  Status codes map the reference's exception
 *     types one-to-one (ArgumentError, DimensionMismatch, BoundsError, KeyError).
 *   - indices crossing the ABI are 1-based Int64 exactly as in the reference
 *     (-1 = "not in the basis"); basis words are uint64 (BR = UInt64 storage; the
 *     `br_bits` argument only reproduces the reference's width check).
 *   - vectors are Float64 (ED_F64) or ComplexF64 (ED_C128, interleaved re,im).
 *   - `void*` vector arguments may be HOST or DEVICE pointers; the library detects which
 *     (cudaPointerGetAttributes).  Host buffers are staged through device memory inside
 *     the call; device buffers are used in place on the library's stream.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with ED_ERR_CUDA.
 *   - handles are opaque, created by *_create / *_generate / *_reduce and released by the
 *     matching *_destroy.  A handle keeps what it needs from its inputs alive itself.
 *   - one in-flight call per handle; calls are stream-ordered on the library stream
 *     (ed_set_stream) and synchronise before returning unless stated otherwise.
 */
#ifndef EDCUDA_H
#define EDCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------ */
#define ED_OK                     0
#define ED_ERR_ARGUMENT           1  /* Julia ArgumentError     */
#define ED_ERR_DIMENSION_MISMATCH 2  /* Julia DimensionMismatch */
#define ED_ERR_BOUNDS             3  /* Julia BoundsError       */
#define ED_ERR_KEY                4  /* Julia KeyError          */
#define ED_ERR_CUDA               5  /* CUDA runtime failure / no device */
#define ED_ERR_UNSUPPORTED        6  /* valid in the reference, outside this engine (e.g. BR wider than 64 bit) */
#define ED_ERR_INTERNAL           7

/* ---- scalar types / flags ---------------------------------------------------------- */
#define ED_F64   0
#define ED_C128  1

#define ED_SIDE_LEFT  0  /* out (+)= H * x   : row walk,    apply!(out, opr, state) */
#define ED_SIDE_RIGHT 1  /* out (+)= x * H   : column walk, apply!(out, state, opr) */

/* how a basis answers "word -> index" */
#define ED_BASIS_LIST        0  /* sorted array + binary search  (FrozenSortedArrayIndex) */
#define ED_BASIS_FULL        1  /* whole space, every site a power of two: index = word    */
#define ED_BASIS_COMBINADIC  2  /* 1-bit sites, one U(1) value: combinatorial number system */
#define ED_BASIS_DPRANK      3  /* any HilbertSpaceSector: per-site DP rank tables          */

typedef struct ed_space    ed_space;
typedef struct ed_basis    ed_basis;
typedef struct ed_operator ed_operator;
typedef struct ed_symmetry ed_symmetry;
typedef struct ed_rbasis   ed_rbasis;
typedef struct ed_oprep    ed_oprep;
typedef struct ed_ctx      ed_ctx;      /* multi-GPU communicator (one rank per GPU of the node) */
typedef struct ed_sharded  ed_sharded;  /* operator representation with its rows distributed over the ranks */
typedef struct ed_dvec     ed_dvec;     /* distributed vector: every rank holds its rows */
typedef struct ed_lanczos_state ed_lanczos_state;   /* resumable Lanczos run */

/* ---- library ----------------------------------------------------------------------- */
const char* ed_last_error(void);
const char* ed_version(void);
/* number of CUDA devices visible (0 when none: every compute call will fail). */
int  ed_device_count(void);
/* select the device used by the calling thread's subsequent calls (cudaSetDevice). */
int  ed_set_device(int device);
/* enable != 0: run this thread's library work on `cuda_stream` (a cudaStream_t; NULL = the legacy default
 * stream) instead of the library's own non-blocking stream; enable == 0 restores the library stream. */
int  ed_set_stream(void* cuda_stream, int32_t enable);
/* host vectors are staged through device buffers that the library keeps between calls (per host thread, the two
 * largest); this frees them. */
int  ed_release_staging(void);
/* total number of kernel launches issued by the library in this process (for bench accounting). */
int64_t ed_kernel_launch_count(void);

/* plain device buffers owned by the caller (e.g. a host language without its own CUDA binding). */
int ed_device_malloc(int64_t bytes, void** ptr);
int ed_device_free(void* ptr);

/* ---- HilbertSpace  (HilbertSpace/hilbert_space.jl:25-41, site.jl:69-93) ------------- */
/* n_states[i] local states on site i; qn is [sum_i n_states[i]][n_qn] row-major: the quantum
 * number tuple of every local state in site order.  Site 0 occupies the least significant
 * bits; width(site) = ceil(log2(n_states)). */
int ed_space_create(int32_t n_sites, const int32_t* n_states, const int64_t* qn, int32_t n_qn,
                    ed_space** out);
int ed_space_destroy(ed_space* space);
int ed_space_bitwidth(const ed_space* space, int32_t* bitwidth);   /* hilbert_space.jl:96 */

/* ---- HilbertSpaceRepresentation  (Representation/hilbert_space_representation.jl) -- */
/* represent(hs, BR) / represent(HilbertSpaceSector(hs, allowed), BR)   :215-230, basis :109-206.
 * allowed_qn is [n_allowed][n_qn]; n_allowed < 0 means "whole space" (represent(hs)).
 * Allowed values that the space cannot reach are ignored (hilbert_space_sector.jl:27-63).
 * br_bits = 8*sizeof(BR) of the caller: ArgumentError when br_bits <= bitwidth (:63-69, :110, :129);
 * br_bits > 64 -> ED_ERR_UNSUPPORTED.  The basis is generated ON DEVICE, ascending. */
int ed_basis_generate(const ed_space* space, const int64_t* allowed_qn, int64_t n_allowed,
                      int32_t br_bits, ed_basis** out);
/* represent(hs, basis_list) :242-256 -- sorts if unsorted; duplicates -> ED_ERR_ARGUMENT
 * (frozensortedarray.jl:15-19). words is a host pointer. */
int ed_basis_from_list(const ed_space* space, const uint64_t* words, int64_t n, int32_t br_bits,
                       ed_basis** out);
int ed_basis_destroy(ed_basis* basis);
int ed_basis_dim(const ed_basis* basis, int64_t* dim);              /* dimension() :90 */
int ed_basis_kind(const ed_basis* basis, int32_t* kind);
/* copy basis_list[lo+1 .. lo+n] (0-based offset lo) to a host buffer. */
int ed_basis_download(ed_basis* basis, int64_t lo, int64_t n, uint64_t* words_out);
/* get(basis_lookup, key, -1) for n host keys (frozensortedarray.jl:29-48): 1-based index or -1. */
int ed_basis_lookup(ed_basis* basis, const uint64_t* keys, int64_t n, int64_t* index_out);
/* device pointer to the (materialised) ascending word array, for zero-copy consumers. */
int ed_basis_device_words(ed_basis* basis, const uint64_t** dev_words);

/* ---- SumOperator  (Operator/pure_operator.jl:25-52, sum_operator.jl:13-23) ---------- */
/* terms in the reference's order; amplitude is n_terms doubles, or n_terms (re,im) pairs when
 * is_complex.  ArgumentError if bitrow/bitcol has a bit outside bitmask (pure_operator.jl:32-36). */
int ed_operator_create(int64_t n_terms, const uint64_t* bitmask, const uint64_t* bitrow,
                       const uint64_t* bitcol, const double* amplitude, int32_t is_complex,
                       ed_operator** out);
int ed_operator_destroy(ed_operator* op);

/* ---- symmetry  (Symmetry/symmetry_apply.jl:65-92, bitflipsymmetry.jl:23-35) --------- */
/* n_ops group elements; perm[g*n_sites + i] = j sends site i to site j (0-based; the
 * SitePermutation.permutation.map of the reference minus one).  flip[g] != 0 composes the
 * element with GlobalBitFlip(true) (NULL = none).  chi[2g], chi[2g+1] = (re, im) of the
 * amplitude handed to symmetry_reduce(hsr, symops_and_amplitudes).  Element 0 must be the
 * identity (symmetry_reduce_generic.jl:56).  ArgumentError unless every |chi| ~ 1 (:27-29). */
int ed_symmetry_create(int32_t n_ops, int32_t n_sites, const int32_t* perm, const uint8_t* flip,
                       const double* chi, ed_symmetry** out);
int ed_symmetry_destroy(ed_symmetry* sym);
/* symmetry_apply(hs, op_g, word) for n host words: images_out[k] = g(words[k]). */
int ed_symmetry_apply(const ed_space* space, const ed_symmetry* sym, int32_t g,
                      const uint64_t* words, int64_t n, uint64_t* images_out);

/* isinvariant(hs, symop, op) for every element of `sym` (Symmetry/symmetry_apply.jl:110-135): *invariant = 1 iff
 * <g b'| op |g b> == <b'| op |b> within tol (tol < 0: sqrt(eps)) for every element g on the sampled words b (every word of
 * spaces up to 10 bits, 48 random words otherwise); *first_bad_element (may be NULL) = the first violating element or -1.
 * Host only, no GPU needed.  ed_oprep_create_reduced runs it and fails with ED_ERR_ARGUMENT on a non-invariant operator. */
int ed_operator_isinvariant(const ed_space* space, const ed_symmetry* sym, const ed_operator* op, double tol,
                            int32_t* invariant, int32_t* first_bad_element);

/* ---- ReducedHilbertSpaceRepresentation  (Symmetry/symmetry_reduce_generic.jl:22-255) */
/* symmetry_reduce(hsr, symops_and_amplitudes; tol).  Representatives (orbit minima whose
 * stabiliser characters are ~1 within tol) are filtered and compacted on device without
 * the per-parent-state arrays of the reference; those are served on demand below. */
int ed_symmetry_reduce(ed_basis* parent, const ed_symmetry* sym, double tol, ed_rbasis** out);
int ed_rbasis_destroy(ed_rbasis* rbasis);
int ed_rbasis_dim(const ed_rbasis* rbasis, int64_t* dim);
int ed_rbasis_download(ed_rbasis* rbasis, int64_t lo, int64_t n, uint64_t* words_out);
/* orbit sizes N_r (number of distinct images) of representatives lo .. lo+n-1. */
int ed_rbasis_orbit_sizes(ed_rbasis* rbasis, int64_t lo, int64_t n, int32_t* sizes_out);
/* (basis_mapping_index, basis_mapping_amplitude) of n parent WORDS (host):
 * index 1-based or -1; amplitude (re,im) = conj(chi_g)/sqrt(N_r) with word = g(r), last g wins
 * (symmetry_reduce_generic.jl:74-101).  Words outside the parent basis give -1. */
int ed_rbasis_mapping(ed_rbasis* rbasis, const uint64_t* parent_words, int64_t n,
                      int64_t* index_out, double* amplitude_out);
/* the reference's two length-dim(parent) arrays for parent rows lo .. lo+n-1. */
int ed_rbasis_mapping_rows(ed_rbasis* rbasis, int64_t lo, int64_t n,
                           int64_t* index_out, double* amplitude_out);
/* symmetry_reduce(rhsr, large_vector) / symmetry_reduce!(out, ...)  (symmetry_reduce.jl:38-153):
 * small[idx[p]] += conj(amp[p]) * large[p].  accumulate=0 zero-fills first.  small is ED_C128;
 * large has dtype `large_dtype`.  Lengths are checked -> ED_ERR_DIMENSION_MISMATCH. */
int ed_vector_reduce(ed_rbasis* rbasis, void* small_out, int64_t n_small, const void* large,
                     int64_t n_large, int32_t large_dtype, int32_t accumulate);
/* symmetry_unreduce(rhsr, small_vector)  (symmetry_reduce.jl:208-225): large[p] = amp[p]*small[idx[p]]. */
int ed_vector_unreduce(ed_rbasis* rbasis, void* large_out, int64_t n_large, const void* small,
                       int64_t n_small, int32_t small_dtype);

/* ---- OperatorRepresentation / ReducedOperatorRepresentation ------------------------- */
/* represent(hsr, op)  (Representation/operator_representation.jl:34-36) */
int ed_oprep_create(ed_basis* basis, const ed_operator* op, ed_oprep** out);
/* represent(rhsr, op) (Symmetry/reduced_operator_representation.jl:40-42); scalar type is always ComplexF64 (:26) */
int ed_oprep_create_reduced(ed_rbasis* rbasis, const ed_operator* op, ed_oprep** out);
int ed_oprep_destroy(ed_oprep* oprep);
int ed_oprep_dim(const ed_oprep* oprep, int64_t* dim);
/* ED_F64 or ED_C128: valtype of the representation (complex operator or reduced space -> ED_C128). */
int ed_oprep_dtype(const ed_oprep* oprep, int32_t* dtype);
/* Row-shard the representation (multi-GPU): this process owns output rows [row_lo, row_hi)
 * (0-based, half open).  x keeps the full length; out has row_hi-row_lo elements. Default: all rows. */
int ed_oprep_set_rows(ed_oprep* oprep, int64_t row_lo, int64_t row_hi);
/* Row range rank `rank` of `world` should own: the reference's balanced splitrange (src/util.jl:102-121) with the
 * boundaries snapped to the fast kernel's tile boundaries, so that segmented inputs (below) are tile aligned. */
int ed_oprep_suggest_rows(ed_oprep* oprep, int32_t dtype, int32_t world, int32_t rank, int64_t* row_lo, int64_t* row_hi);
/* Choose the kernel: 0 = automatic (fastest exact path), 1 = force the generic term-walk kernel. */
int ed_oprep_set_kernel(ed_oprep* oprep, int32_t which);

/* apply!(out, opr, x) / apply!(out, x, opr) / mul!(out, opr, x)
 * (Representation/abstract_operator_representation.jl:110-118, 260-409):
 *   accumulate=1: out += H*x (apply!);  accumulate=0: out = H*x (mul!).
 *   n_out / n_x are the callers' vector lengths: checked against the dimension
 *   (n_out against the owned row count) -> ED_ERR_DIMENSION_MISMATCH (:303-307).
 *   dtype is the element type of BOTH out and x; a complex representation needs ED_C128. */
int ed_apply(ed_oprep* oprep, void* out, int64_t n_out, const void* x, int64_t n_x,
             int32_t dtype, int32_t side, int32_t accumulate);
/* same, device pointers only, asynchronous on the library stream (no sync before return);
 * `alpha_dot` (device double[2] or NULL) receives sum_i conj(x[row_i]) * out[i] over the owned rows
 * (the Lanczos alpha partial, fused in the epilogue). */
int ed_apply_async(ed_oprep* oprep, void* out, const void* x, int32_t dtype, int32_t side,
                   int32_t accumulate, double* alpha_dot);

/* get_row_iterator(opr, i) / get_column_iterator(opr, i) (operator_representation.jl:66-103,
 * reduced_operator_representation.jl:57-116): up to `cap` (index, amplitude) pairs in term
 * order, misses as -1; *n_out = number of pairs.  i is 1-based; BoundsError outside 1..dim.
 * amplitude_out holds (re,im) pairs when the representation is complex, else doubles. */
int ed_oprep_row_iterator(ed_oprep* oprep, int64_t i, int32_t side, int64_t cap,
                          int64_t* index_out, double* amplitude_out, int64_t* n_out);
/* get_element(opr, i, j) (operator_representation.jl:109-119, reduced :120-138). value_out = (re, im). */
int ed_oprep_get_element(ed_oprep* oprep, int64_t i, int64_t j, double* value_out);

/* sparse(opr; tol) (abstract_operator_representation.jl:136-204, util.jl:88-94):
 * CSC, 1-based Int64 colptr/rowval, rows ascending inside a column, entries with |v| < tol removed.
 * Two calls: ed_sparse_count assembles on device and returns nnz; ed_sparse_fetch copies the three
 * arrays (colptr has dim+1 entries; nzval is nnz doubles or nnz (re,im) pairs) and frees the
 * device copy.  tol < 0 selects the reference default sqrt(eps(Float64)). */
int ed_sparse_count(ed_oprep* oprep, double tol, int64_t* nnz_out);
int ed_sparse_fetch(ed_oprep* oprep, int64_t* colptr, int64_t* rowval, void* nzval);
/* Keep the assembled rows of the representation on device (owned row range, side LEFT = rows of H, RIGHT = rows of
 * H^T; nothing chopped) so that later ed_apply / ed_apply_async / ed_lanczos calls run as a bandwidth-bound SpMV
 * instead of redoing the term walk -- the practical path for symmetry-reduced representations, whose matrix-free
 * apply is instruction-bound (|G| images per hit).  ed_oprep_set_kernel(oprep, 1) bypasses the cache; rebuild after
 * ed_oprep_set_rows.  The reference's get_row/get_column/Matrix users are served by the iterators above. */
int ed_oprep_cache_matrix(ed_oprep* oprep, int32_t side, int64_t* nnz_out);
int ed_oprep_drop_cache(ed_oprep* oprep);
/* Matrix(opr) (:121-132): dense column-major dim x dim, no chop.  out is a host buffer of the
 * representation's dtype. */
int ed_dense(ed_oprep* oprep, void* out);

/* ---- Lanczos driver (not in the reference: it hands mul! to Arpack, docs/src/examples/spinhalf.md:26) */
/* n_steps of three-term Lanczos on device from v0 (host or device, dim elements of `dtype`; NULL =
 * Philox-seeded normal vector from `seed`).  alpha[n_steps], beta[n_steps] (host) receive the
 * tridiagonal; ritz[n_ritz] the lowest Ritz values.  Single device; the sharded multi-GPU loop
 * lives in the host layer and uses the ed_lanczos_* kernels below with NCCL collectives between them.
 * steps_done (may be NULL) = number of valid (alpha, beta) pairs (smaller than n_steps on breakdown). */
int ed_lanczos(ed_oprep* oprep, int32_t n_steps, const void* v0, int32_t dtype, uint64_t seed,
               double* alpha, double* beta, double* ritz, int32_t n_ritz, int32_t* steps_done);
/* fused vector update on the owned rows (device pointers, async).  Krylov vectors are kept unnormalised:
 *   u_next = (w - alpha * u_cur) / n_cur - (n_cur / n_prev) * u_prev ,  alpha = dot[0] / n_cur^2,
 * written over u_prev;  norm2_out[0] (device double[2]) = sum |u_next|^2 over the owned rows.
 * dot = device double[2] from ed_apply_async (already summed over shards), norm2_cur / norm2_prev = device
 * doubles holding |u_cur|^2 and |u_prev|^2 (norm2_prev NULL on the first step). */
int ed_lanczos_update_async(void* u_prev_inout, const void* w, const void* u_cur, int64_t n, int32_t dtype,
                            const double* dot, const double* norm2_cur, const double* norm2_prev,
                            double* norm2_out);
/* norm2_out[0] (device double[2]) = sum |v|^2 over n elements (async). */
int ed_vector_norm2_async(const void* v, int64_t n, int32_t dtype, double* norm2_out);
/* v[i] = normal(0,1) from Philox keyed by (seed, global_row_offset + i): shard-count independent. */
int ed_vector_randn_async(void* v, int64_t n, int32_t dtype, uint64_t seed, int64_t global_row_offset);
/* v *= a (async). */
int ed_vector_scale_async(void* v, int64_t n, int32_t dtype, double a);
/* lowest eigenvalues of the k x k symmetric tridiagonal (alpha, beta[0..k-2]) -- host helper. */
int ed_tridiag_eigvals(const double* alpha, const double* beta, int32_t k, double* eig_out);

/* ---- checkpoints (raw binary files; not in the reference, which has no persistence) ----------------------------
 * A sector basis stores how it was generated (its rank tables are rebuilt on load), a user list its words; both are
 * bound to the Hilbert space by a hash.  A reduced basis stores representatives, orbit sizes and stabiliser marks and is
 * bound to the symmetry operations, characters and tolerance it was made with: loading skips the filter pass over the
 * parent space.  Mismatching files fail with ED_ERR_ARGUMENT. */
int ed_basis_save(ed_basis* basis, const char* path);
int ed_basis_load(const ed_space* space, const char* path, ed_basis** out);
int ed_rbasis_save(ed_rbasis* rbasis, const char* path);
int ed_rbasis_load(ed_basis* parent, const ed_symmetry* sym, double tol, const char* path, ed_rbasis** out);
/* ed_lanczos as a resumable object: create (v0 NULL = Philox vector from seed), run n_steps more steps as often as
 * wanted, read (alpha, beta, lowest Ritz values) of all steps taken so far, save / load the full state (both Krylov
 * vectors and the scalars).  A loaded state continues bit for bit like an uninterrupted run. */
int ed_lanczos_state_create(ed_oprep* oprep, int32_t dtype, const void* v0, uint64_t seed, ed_lanczos_state** out);
int ed_lanczos_state_destroy(ed_lanczos_state* state);
int ed_lanczos_state_step(ed_lanczos_state* state, int32_t n_steps);
int ed_lanczos_state_result(ed_lanczos_state* state, int32_t capacity, double* alpha, double* beta, double* ritz,
                            int32_t n_ritz, int32_t* steps_done);
int ed_lanczos_state_save(ed_lanczos_state* state, const char* path);
int ed_lanczos_state_load(ed_oprep* oprep, const char* path, ed_lanczos_state** out);

/* ---- multi-GPU: rows sharded over the GPUs of one node ------------------------------------------------
 * The reference parallelises INSIDE apply! (Threads.@threads over statically split rows,
 * Representation/abstract_operator_representation.jl:260-267, 358-378; splitrange, util.jl:102-121): callers never
 * partition anything.  The same holds here: a context owns the streams and the NCCL communicator(s), a sharded
 * representation owns the row partition and the exchange, and a distributed vector is addressed through the API
 * (ascending-basis) order.  Two ways to create a context:
 *   ed_ctx_create       ONE process drives n_gpus devices (ncclCommInitAll; peer access between the devices).  All
 *                       device ids equal = "loopback": the ranks share one GPU and the collectives are emulated with
 *                       stream-ordered kernels (no NCCL) -- for testing the world > 1 logic on a single GPU.
 *   ed_ctx_create_rank  one process per GPU (torchrun / mpirun): rank 0 calls ed_ctx_unique_id, the launcher broadcasts
 *                       the 128 bytes, every process joins with its rank (ncclCommInitRank; buffers through CUDA IPC).
 * Every function below that takes a context-bound handle is COLLECTIVE: all processes call it in the same order.
 * Per-rank arguments are arrays over the LOCAL ranks of the calling process (n_gpus entries, or 1). */
int ed_ctx_unique_id(uint8_t* uid128);
int ed_ctx_create(int32_t n_gpus, const int32_t* device_ids /* NULL: 0..n_gpus-1 */, ed_ctx** out);
int ed_ctx_create_rank(int32_t world, int32_t rank, int32_t device, const uint8_t* uid128, ed_ctx** out);
int ed_ctx_destroy(ed_ctx* ctx);
/* any output may be NULL; nccl_version = 0 when no NCCL communicator exists (world 1 or loopback). */
int ed_ctx_info(const ed_ctx* ctx, int32_t* world, int32_t* n_local, int32_t* first_rank, int32_t* nccl_version);
/* device and cudaStream_t of local rank `local_index`: all asynchronous work of the context runs on that stream. */
int ed_ctx_device_stream(const ed_ctx* ctx, int32_t local_index, int32_t* device, void** cuda_stream);
int ed_ctx_sync(ed_ctx* ctx);       /* host waits for this process's ranks */
int ed_ctx_barrier(ed_ctx* ctx);    /* stream-ordered barrier over all ranks, then ed_ctx_sync */
/* CUDA-event timing on the ranks' streams: record into slot 0..61, then elapsed = MAX over all ranks (milliseconds). */
int ed_ctx_timer_record(ed_ctx* ctx, int32_t slot);
int ed_ctx_timer_elapsed(ed_ctx* ctx, int32_t slot_a, int32_t slot_b, double* ms_max);
/* in-place all-reduce of up to 4 host doubles over the PROCESSES of the context (op 0 = sum, 1 = max). */
int ed_ctx_allreduce_host(ed_ctx* ctx, double* values, int32_t count, int32_t op);

/* opreps[i] = the representation created on local rank i's device (same operator and basis everywhere).
 * exchange: 0 = automatic, 1 = NCCL all-gather of x per matvec (contiguous row ranges; any representation),
 *           2, 3, 4 = halo exchange (tiled U(1) kernel only): every rank owns whole kernel tiles; the peer tiles it reads
 *               land in a compact halo buffer, piece by piece in the order its n_chunks launch chunks need them, each
 *               chunk waiting only for its own piece.  2: the owner packs a send buffer and the READER's copy engines pull
 *               it; 3: the OWNER writes the tiles with remote stores from a small persistent kernel and bumps a per-chunk
 *               arrival counter (remote atomic); 4: the owner packs and its copy engines push.  Automatic = halo where
 *               supported (transport: EDCUDA_SHARD_TRANSPORT = pull | push | cepush, default pull).  n_chunks <= 0: 8. */
int ed_sharded_create(ed_ctx* ctx, ed_oprep* const* opreps, int32_t dtype, int32_t exchange, int32_t n_chunks, ed_sharded** out);
int ed_sharded_destroy(ed_sharded* sh);
/* rows owned by local rank `local_index`, elements it copies from peers per matvec, number of its global row ranges,
 * peer copies (pull) or pushed pieces (push) and launch chunks per matvec, and the exchange in use: 0 all-gather,
 * 1 halo by pulls, 2 halo by SM pushes, 3 halo by copy-engine pushes.  Outputs may be NULL. */
int ed_sharded_info(const ed_sharded* sh, int32_t local_index, int64_t* n_local, int64_t* n_halo, int32_t* n_ranges,
                    int32_t* n_pulls, int32_t* n_chunks, int32_t* halo_exchange);
/* the global row ranges [row_lo[k], row_hi[k]) of that rank, ascending; its local vectors store them back to back. */
int ed_sharded_ranges(const ed_sharded* sh, int32_t local_index, int64_t* row_lo, int64_t* row_hi);

int ed_dvec_create(ed_sharded* sh, ed_dvec** out);     /* zero-filled */
int ed_dvec_destroy(ed_dvec* v);                        /* ranks unmap their peers' buffers before the owners free them */
int ed_dvec_local(ed_dvec* v, int32_t local_index, void** dev_ptr, int64_t* n_local);
/* v[row] = scale * normal(0,1) from Philox keyed by (seed, global row): independent of the number of ranks. */
int ed_dvec_randn(ed_dvec* v, uint64_t seed, double scale);
/* host_full: full-length vector in basis order.  upload: every process copies in the rows its ranks own;
 * download: every process writes the rows its ranks own and leaves the others untouched. */
int ed_dvec_upload(ed_dvec* v, const void* host_full);
int ed_dvec_download(ed_dvec* v, void* host_full);

/* y = H x over all ranks (mul!), asynchronous on the ranks' streams unless dot_out (host double[2]) is given, which then
 * receives <x, Hx> summed over all rows.  no_fence != 0 skips the stream-ordered barrier that publishes x to the peers
 * (allowed when a collective already ran on every rank's stream after x was last written). */
int ed_apply_sharded(ed_sharded* sh, ed_dvec* y, ed_dvec* x, int32_t no_fence, double* dot_out);
/* One matvec taken apart (halo exchange only; phases that overlap in ed_apply_sharded run back to back here):
 * ms4 = {owner-side pack, fence, all peer copies alone, all kernel chunks with the halo in place}, each max over ranks. */
int ed_sharded_profile(ed_sharded* sh, ed_dvec* y, ed_dvec* x, double* ms4);
/* ed_lanczos over the shards: Krylov vectors stay distributed and device resident, the two scalars of every step are
 * all-reduced (NCCL), no host synchronisation inside the loop.  v0 NULL = Philox vector from seed.  ms_per_step (may
 * be NULL) = device time per step, max over ranks. */
int ed_lanczos_sharded(ed_sharded* sh, int32_t n_steps, uint64_t seed, ed_dvec* v0, double* alpha, double* beta,
                       double* ritz, int32_t n_ritz, int32_t* steps_done, double* ms_per_step);

/* Host-only description (no GPU needed) of rank `rank`'s share of the halo-exchange layout for operator `op` on the
 * n_set-particle sector of n_bits spin-1/2 sites (policy 0 = planner's choice, 1 = ascending row ranges, 2 = wrap-aware
 * ranges, 3 = best popcount ordering).
 * counts[12] = {rows, halo rows, ranges, tiles, pulls, reads, chunks, dim, packs, send rows, pushes, 0}.
 * Optional outputs (NULL to skip): ranges[2*ranges] = (lo, hi);
 * tiles[4*tiles] = (global first row, rows, local offset, launch chunk), in launch order;
 * pulls[5*pulls] = (peer, chunk, offset in the PEER'S SEND BUFFER, offset in the halo, rows);
 * packs[3*packs] = (offset in this rank's vector, offset in its send buffer, rows);
 * reads[4*reads] = (tile index, global first row of the tile it reads, rows, where: offset | 1<<62 if in the halo);
 * pushes[5*pushes] = (receiver, its launch chunk, offset in this rank's vector, offset in the RECEIVER'S halo, rows). */
int ed_shard_plan_describe(const ed_operator* op, int32_t n_bits, int32_t n_set, int32_t dtype, int32_t world, int32_t rank,
                           int32_t n_chunks, int32_t policy, int64_t* counts, int64_t* ranges, int64_t* tiles,
                           int64_t* pulls, int64_t* packs, int64_t* reads, int64_t* pushes);

#ifdef __cplusplus
}
#endif
#endif /* EDCUDA_H */
