// Internal declarations shared by the translation units of libedcuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <complex>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/edcuda.h"

// ------------------------------------------------------------------ errors
struct EdError : public std::runtime_error {
  int code;
  EdError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void ed_set_error(const std::string& msg);

#define ED_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      throw EdError(ED_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e__) +    \
                                     " at " + __FILE__ + ":" + std::to_string(__LINE__));   \
  } while (0)

#define ED_REQUIRE(cond, code, msg)            \
  do {                                         \
    if (!(cond)) throw EdError((code), (msg)); \
  } while (0)

// Every exported function body is wrapped with these.
#define ED_TRY try {
#define ED_CATCH                                        \
  }                                                     \
  catch (const EdError& e) {                            \
    ed_set_error(e.what());                             \
    return e.code;                                      \
  }                                                     \
  catch (const std::bad_alloc&) {                       \
    ed_set_error("host allocation failed");             \
    return ED_ERR_INTERNAL;                             \
  }                                                     \
  catch (const std::exception& e) {                     \
    ed_set_error(std::string("internal: ") + e.what()); \
    return ED_ERR_INTERNAL;                             \
  }                                                     \
  return ED_OK;

// ------------------------------------------------------------------ runtime context
extern std::atomic<int64_t> g_launch_count;
cudaStream_t ed_stream();          // stream all library work is issued on
void ed_require_device();          // throws ED_ERR_CUDA when no usable device
bool ed_is_device_pointer(const void* p);
int ed_sm_count();

#define ED_LAUNCH(kernel, grid, block, smem, ...)                         \
  do {                                                                    \
    kernel<<<(grid), (block), (smem), ed_stream()>>>(__VA_ARGS__);        \
    g_launch_count.fetch_add(1, std::memory_order_relaxed);               \
    ED_CUDA(cudaGetLastError());                                          \
  } while (0)

int ed_current_device();            // runtime.cu

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) ED_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void upload(const T* h, size_t count) {
    if (count > n) alloc(count);
    if (count) ED_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, ed_stream()));
  }
  void upload(const std::vector<T>& h) { upload(h.data(), h.size()); }
  void download(T* h, size_t count, size_t offset = 0) const {
    if (count) {
      ED_CUDA(cudaMemcpyAsync(h, p + offset, count * sizeof(T), cudaMemcpyDeviceToHost, ed_stream()));
      ED_CUDA(cudaStreamSynchronize(ed_stream()));
    }
  }
};

// Scratch buffers that persist between calls are kept per host thread AND per device (one process may drive several
// GPUs through an ed_ctx): TAG distinguishes the call sites.
template <typename T, int TAG>
DevBuf<T>& ed_scratch() {
  static thread_local DevBuf<T> buf[16];
  return buf[ed_current_device() & 15];
}

// Stages a caller buffer (host or device) as a device pointer for the duration of a call.
struct Staged {
  void* dev = nullptr;
  void* host = nullptr;
  size_t bytes = 0;
  bool owned = false;
  bool writeback = false;
  Staged(const void* p, size_t nbytes, bool copy_in, bool copy_out);
  ~Staged();
  void finish();  // copy back (if requested) and free
};

// ------------------------------------------------------------------ Hilbert space
struct ed_space {
  int n_sites = 0;
  int n_qn = 0;
  std::vector<int> n_states, width, offset;  // offset has n_sites+1 entries
  std::vector<int64_t> qn;                    // [site][state][n_qn] flattened in site order
  std::vector<int> qn_base;                   // index of site's first state in qn (in tuples)
  int bits = 0;
  bool one_bit_sites = true;  // every site has exactly 2 states
  bool all_pow2 = true;       // every site has 2^w states
  const int64_t* qn_of(int site, int state) const { return &qn[(size_t)(qn_base[site] + state) * n_qn]; }
  uint64_t fullmask() const { return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull); }
};

// ------------------------------------------------------------------ word -> index descriptors
#define ED_MAX_SITES 64

// Passed by value to kernels.  `kind` selects which members are meaningful.
struct LookupDesc {
  int kind;
  int64_t dim;
  // LIST
  const uint64_t* words;  // full ascending list (device)
  // FULL / COMBINADIC
  int n_bits;
  int n_set;
  int n_chunks;
  const uint64_t* comb_lut;  // [(chunk*(n_set+1)+below)*256 + byte]
  const uint64_t* binom;     // [n*65 + k], n,k in 0..64
  // DPRANK
  int n_sites;
  int max_q;
  int max_states;
  int root_q;
  const uint64_t* dp_prefix;  // [(site*max_q+q)*(max_states+1) + v]  = sum_{v'<v} count(child)
  const int32_t* dp_next;     // [(site*max_q+q)*max_states + v]      = child q id or -1
  const uint8_t* dp_accept;   // [q] at level 0
  const uint8_t* site_off;    // [n_sites]
  const uint8_t* site_w;      // [n_sites]
  const uint8_t* site_ns;     // [n_sites]
};

struct ed_basis {
  ed_space space;
  int kind = ED_BASIS_LIST;
  int64_t dim = 0;
  int br_bits = 64;
  // materialised ascending words (device); always the full list once present
  DevBuf<uint64_t> words;
  bool words_ready = false;
  // COMBINADIC
  int n_set = 0;
  DevBuf<uint64_t> comb_lut, binom;
  std::vector<uint64_t> h_binom;  // host copy [65*65]
  // DPRANK
  int max_q = 0, max_states = 0, root_q = 0;
  DevBuf<uint64_t> dp_prefix;
  DevBuf<int32_t> dp_next;
  DevBuf<uint8_t> dp_accept, site_off, site_w, site_ns;
  // how the basis was made (checkpoints regenerate sector bases from this instead of storing words)
  int64_t gen_n_allowed = -2;            // -2: user list, -1: whole space, >= 0: sector with these allowed tuples
  std::vector<int64_t> gen_allowed;
  LookupDesc desc() const;
  void materialize();  // generate `words` on device if not present (K1)
};

// ------------------------------------------------------------------ operators
struct ed_operator {
  int64_t n_terms = 0;
  bool is_complex = false;
  std::vector<uint64_t> mask, row, col;
  std::vector<double> amp;  // n_terms or 2*n_terms
};

// Device-side term table (SoA).  amp is double or double2 depending on `is_complex`.
struct TermsDev {
  int n_terms = 0;
  bool is_complex = false;
  DevBuf<uint64_t> mask, match, target;  // match = bitrow (left) or bitcol (right); target likewise swapped
  DevBuf<double> amp;
};

// ------------------------------------------------------------------ symmetry
struct ed_symmetry {
  int n_ops = 0;
  int n_sites = 0;
  std::vector<int32_t> perm;   // [n_ops][n_sites]
  std::vector<uint8_t> flip;   // [n_ops]
  std::vector<double> chi;     // [n_ops][2]
  bool is_group = false;       // closed under inverse & product
  std::vector<int32_t> inverse;  // [n_ops] index of inverse element (when is_group)
};

// Device form of a symmetry bound to a space: byte-chunk permutation LUTs.
struct SymDev {
  int n_ops = 0;
  int n_chunks = 0;        // ceil(bits / 8)
  uint64_t fullmask = 0;
  DevBuf<uint64_t> lut;    // [(g*n_chunks + c)*256 + byte] -> image bits (flip already folded in chunk 0.. see symmetry.cu)
  int n_chunks6 = 0;       // ceil(bits / 6)
  DevBuf<uint64_t> lut6;   // [(g*n_chunks6 + c)*64 + v]: the same action in 6-bit chunks (shared-memory streaming, reduced_staged.cu)
  DevBuf<double> chi;      // [n_ops][2]
  DevBuf<int32_t> inverse; // [n_ops]
  DevBuf<uint8_t> chi_is_one;  // [n_ops]  |chi-1| <= tol
  // Translation factorisation (1-bit sites, site = x + tr_n1 * y): when the group contains the n1 x n2 lattice
  // translations T, G is swept as cosets T p_j -- p_j through its 6-bit LUT, the translations as rotate-within-field /
  // rotate-word ALU steps (reduced_staged.cu: k6b_canonicalize_tr).
  bool tr_on = false;
  int tr_n1 = 0, tr_n2 = 0, tr_ncos = 0;
  DevBuf<uint64_t> tr_lut6;    // [(j*n_chunks6 + c)*64 + v] for the coset representatives p_j
  DevBuf<int32_t> tr_inv;      // [(j*n2 + b)*n1 + a] = inverse index of the element Tx^a Ty^b p_j
};

struct SymDesc {
  int n_ops;
  int n_chunks;
  const uint64_t* lut;
  const double* chi;
  const int32_t* inverse;
  const uint8_t* chi_is_one;
};

struct ed_rbasis {
  ed_basis* parent = nullptr;     // borrowed; the caller keeps it alive (mirrors rhsr.parent)
  ed_symmetry sym;                // copy
  double tol = 0;
  int64_t dim = 0;
  SymDev symdev;
  DevBuf<uint64_t> words;         // representatives ascending
  DevBuf<uint16_t> orbit_size;    // N_r
  DevBuf<uint16_t> last_stab;     // last group element index stabilising r (its chi sets amp[r])
  // bucket index over the top bits of the words for the reduced lookup
  int bucket_shift = 0;
  DevBuf<uint32_t> bucket_start;  // [n_buckets+1]
  int64_t n_buckets = 0;
  // hash index word -> reduced index (open addressing, one 8-byte slot = word << idx_bits | index): one memory access per
  // look-up where the bucketed search needs ~log2(bucket) dependent ones (orbit minima crowd into few top-bit buckets).
  // Built when word and index fit one slot together; otherwise the bucketed search stays the only index.
  DevBuf<unsigned long long> hash;
  int hash_shift = 0, idx_bits = 0;
  SymDesc symdesc() const;
  struct RLookupDesc rdesc() const;
};

struct RLookupDesc {
  const uint64_t* words;
  const uint16_t* orbit_size;
  const uint16_t* last_stab;
  const uint32_t* bucket_start;
  int bucket_shift;
  int64_t n_buckets;
  int64_t dim;
  const unsigned long long* hash;   // nullptr: bucketed binary search only
  int hash_shift;                   // slot = (word * golden) >> hash_shift
  int idx_bits;
};

// ------------------------------------------------------------------ operator representation
struct FastU1Plan;  // apply_u1.cu

// assembled rows kept on device for repeated matvecs (sparse.cu)
struct CsrCache {
  int side = 0;
  int64_t row_lo = 0, row_hi = 0, nnz = 0;
  bool val_complex = false;
  DevBuf<int64_t> rowptr;   // 0-based, row_hi-row_lo+1 entries (released when the matrix is column-blocked)
  DevBuf<int32_t> col;      // 0-based
  DevBuf<double> val;       // nnz doubles or nnz (re,im) pairs (released when the values are dictionary coded)
  bool coded = false;       // real values stored as 16-bit codes into `lut` (sparse.cu: csr_code_values)
  int n_values = 0;
  DevBuf<uint16_t> code;
  DevBuf<double> lut;
  // column blocking (x larger than the L2): the entries regrouped block-major -- block b holds, row by row, the entries
  // whose column lies in [b * block_cols, (b + 1) * block_cols) -- so that one pass per block gathers x from an L2-sized window
  int n_blocks = 1;
  int64_t block_cols = 0;
  DevBuf<uint32_t> blk_rowptr;        // [n_blocks][n_rows + 1], relative to blk_base[b]
  std::vector<int64_t> blk_base;      // [n_blocks + 1] first entry of block b in col / val
};

struct ed_oprep;
int64_t ed_csr_window_cols(const ed_oprep* o);                     // sparse.cu: columns per pass of the column-blocked SpMV
void ed_csr_set_column_gate(const cudaEvent_t* ev, int n);         // sparse.cu: pass b waits for ev[b]; (nullptr, 0) clears

struct ed_oprep {
  ed_basis* basis = nullptr;    // plain representation: its basis; reduced: the parent basis
  ed_rbasis* rbasis = nullptr;  // non-null for ReducedOperatorRepresentation
  ed_operator op;               // host copy (term order preserved)
  bool is_complex = false;      // valtype complex?
  int64_t dim = 0;
  int64_t row_lo = 0, row_hi = 0;
  int kernel_choice = 0;
  TermsDev terms_left, terms_right;
  bool terms_ready = false;
  std::shared_ptr<FastU1Plan> u1plan, u1plan_c;  // f64 / c128 vectors
  std::shared_ptr<CsrCache> csr[2];               // per side, built by ed_oprep_cache_matrix
  std::vector<int64_t> k6_hits[2];                // staged reduced matvec: emitted column words per row batch, per side
  int64_t k6_hits_lo = -1, k6_hits_hi = -1, k6_hits_batch = -1;   // ... valid for this row range and batch size
  // sparse() result kept between ed_sparse_count and ed_sparse_fetch
  DevBuf<int64_t> sp_colptr, sp_rowval;
  DevBuf<double> sp_nzval;
  int64_t sp_nnz = -1;
};

void ed_upload_terms(ed_oprep* o);

// kernels / drivers implemented in the other translation units
void ed_basis_generate_words(ed_basis* b, int64_t lo, int64_t n, uint64_t* dev_out);          // basis.cu (K1)
void ed_apply_generic(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate,
                      double* alpha_dot);                                                       // apply.cu (K2)
bool ed_apply_u1_supported(ed_oprep* o, int dtype, int side);                                   // apply_u1.cu
void ed_apply_u1(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate,
                 double* alpha_dot);
void ed_u1_suggest_rows(ed_oprep* o, int dtype, int world, int rank, int64_t* lo, int64_t* hi);
// ---- sharded tiled matvec (apply_u1.cu; driven by ctx.cu) -----------------------------------------------------------
struct U1Pull { int peer; int chunk; int64_t src_off, dst_off, len; };   // elements: peer's SEND buffer -> this rank's halo
struct U1Pack { int64_t src_off, dst_off, len; };                        // elements: this rank's x -> its send buffer
struct U1Push { int recv; int chunk; int64_t src_off, dst_off, len; };   // elements: this rank's x -> rank recv's halo
struct U1ShardLayout {
  int world = 1, rank = 0, n_chunks = 1;
  int64_t n_local = 0, n_halo = 0, n_send = 0; // elements
  std::string key_name;                        // the tile ordering the planner chose
  std::vector<int64_t> range_lo, range_hi;     // global row ranges owned by this rank, ascending; local order = this order
  std::vector<uint32_t> tile_H;                // this rank's tiles in launch order
  std::vector<int64_t> tile_off;               // their offsets in the local vectors (storage is ascending in H)
  std::vector<U1Pack> packs;                   // what this rank gathers into its send buffer for its peers (pull exchange)
  std::vector<U1Push> pushes;                  // the same tiles written straight into the peers' halos (push exchange), by chunk
  std::vector<U1Push> piece_pushes;            // one contiguous piece per (receiver, chunk): src_off is in this rank's SEND buffer
  std::vector<int64_t> chunk_halo_rows;        // [n_chunks] rows of this rank's halo that launch chunk c waits for
  std::vector<int> chunk_first;                // [n_chunks + 1] positions in tile_H
  std::vector<int64_t> dir;                    // [2^hb] see U1Params::dir
  std::vector<U1Pull> pulls;                   // ordered by chunk, one per (chunk, peer)
  std::vector<int64_t> rows_of_rank;           // [world]
};
struct U1ShardLaunch {
  const uint32_t* tile_H;      // device copy of U1ShardLayout::tile_H
  int first, count;            // the chunk launched by this call
  const int64_t* dir;          // device copy of U1ShardLayout::dir
  const void* x_local;
  const void* x_halo;
  void* y_local;
  int stream_mode, accumulate;
  double* partials;            // device, 2 doubles per tile of this rank, or nullptr
};
FastU1Plan* ed_u1_plan(ed_oprep* o, int dtype);                                                  // nullptr: not supported
std::shared_ptr<FastU1Plan> ed_u1_host_plan(const ed_operator& op, int n_bits, int n_set, int dtype);   // no device needed
void ed_u1_shard_layout(const FastU1Plan* plan, int world, int rank, int n_chunks, int policy, U1ShardLayout* out);
void ed_apply_u1_sharded(ed_oprep* o, int dtype, const U1ShardLaunch& L);
void ed_reduce_pairs(const double* partials, int n, double* out2);                              // apply.cu
int ed_lanczos_finish(const double* hd, const double* hn, int n_steps, double* alpha, double* beta, double* ritz, int n_ritz);   // lanczos.cu
void ed_rbasis_finish_index(ed_rbasis* r);                                                      // symmetry.cu: bucket index over the words
void ed_push_stream(cudaStream_t s);    // runtime.cu: run this thread's library work on s until ed_pop_stream()
void ed_pop_stream();
void ed_apply_reduced(ed_oprep* o, void* out, const void* x, int side, int accumulate,
                      double* alpha_dot);
bool ed_reduced_fill_raw_staged(ed_oprep* o, int side, int64_t line0, int64_t n, const int64_t* raw_offs, int64_t* raw_row, void* raw_val);  // reduced_staged.cu
bool ed_apply_reduced_staged_supported(ed_oprep* o);                                            // reduced_staged.cu
void ed_apply_reduced_staged(ed_oprep* o, void* out, const void* x, int side, int accumulate, double* alpha_dot);                                                       // reduced.cu (K6)
void ed_apply_csr(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot);
void ed_sparse_assemble(ed_oprep* o, double tol);                                               // sparse.cu (K3/K4)
void ed_symdev_build(const ed_space& space, const ed_symmetry& sym, double tol, SymDev* out);   // symmetry.cu
