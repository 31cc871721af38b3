// HilbertSpace / HilbertSpaceRepresentation: rank tables built on the host, basis words
// generated on the device (kernel K1), word -> index lookups.
//
// Replaces (reference, /root/reference/src):
//   Representation/hilbert_space_representation.jl:109-119  hs_get_basis_list(::HilbertSpace)
//   Representation/hilbert_space_representation.jl:127-206  hs_get_basis_list(::HilbertSpaceSector)
//   frozensortedarray.jl:11-48                              FrozenSortedArrayIndex
// The reference builds the sector basis by a serial site-by-site DP that materialises every
// partial list and k-way merges them.  Here the same DP is run on COUNTS only (host, tiny),
// which gives a closed-form unrank(index) -> word evaluated independently per thread, so the
// ascending basis is written in one coalesced pass and any row range can be generated on any GPU.
#include <algorithm>
#include <set>

#include "ed_device.cuh"

// ------------------------------------------------------------------ host: space
static ed_space make_space(int32_t n_sites, const int32_t* n_states, const int64_t* qn, int32_t n_qn) {
  ED_REQUIRE(n_sites >= 0 && n_sites <= ED_MAX_SITES, ED_ERR_ARGUMENT, "n_sites must be in 0..64");
  ED_REQUIRE(n_qn >= 0 && n_qn <= 8, ED_ERR_ARGUMENT, "n_qn must be in 0..8");
  ed_space s;
  s.n_sites = n_sites;
  s.n_qn = n_qn;
  s.offset.push_back(0);
  int total_states = 0;
  for (int i = 0; i < n_sites; ++i) {
    int ns = n_states[i];
    ED_REQUIRE(ns >= 1 && ns <= 255, ED_ERR_ARGUMENT, "every site needs 1..255 states");
    int w = 0;
    while ((1 << w) < ns) ++w;  // ceil(log2(ns)), site.jl:93
    s.n_states.push_back(ns);
    s.width.push_back(w);
    s.offset.push_back(s.offset.back() + w);
    s.qn_base.push_back(total_states);
    total_states += ns;
    if (ns != 2) s.one_bit_sites = false;
    if ((1 << w) != ns) s.all_pow2 = false;
  }
  s.bits = s.offset.back();
  s.qn.assign(qn, qn + (size_t)total_states * n_qn);
  return s;
}

typedef std::vector<int64_t> QN;

struct DpTables {
  int max_q = 0, max_states = 0, root_q = 0;
  std::vector<uint64_t> prefix;
  std::vector<int32_t> next;
  std::vector<uint8_t> accept;
  uint64_t dim = 0;
};

// Count-only version of the DP in hilbert_space_representation.jl:146-195, run from the most
// significant site down so that prefix sums over the local state give the ascending rank.
static DpTables build_dp(const ed_space& sp, const std::set<QN>* allowed /* null = everything */) {
  const int n = sp.n_sites;
  std::vector<std::map<QN, int>> ids(n + 1);  // ids[i]: partial sums of sites i..n-1
  QN zero(sp.n_qn, 0);
  ids[n][zero] = 0;
  for (int i = n - 1; i >= 0; --i) {
    for (auto& kv : ids[i + 1]) {
      for (int v = 0; v < sp.n_states[i]; ++v) {
        QN q = kv.first;
        const int64_t* dq = sp.qn_of(i, v);
        for (int k = 0; k < sp.n_qn; ++k) q[k] += dq[k];
        if (!ids[i].count(q)) {
          int id = (int)ids[i].size();
          ids[i][q] = id;
        }
      }
    }
    ED_REQUIRE(ids[i].size() < (1u << 20), ED_ERR_UNSUPPORTED, "too many distinct partial quantum numbers");
  }
  DpTables T;
  for (int i = 0; i <= n; ++i) T.max_q = std::max(T.max_q, (int)ids[i].size());
  for (int i = 0; i < n; ++i) T.max_states = std::max(T.max_states, sp.n_states[i]);
  if (T.max_states == 0) T.max_states = 1;
  // F[i][q]: number of completions of sites 0..i-1
  std::vector<std::vector<uint64_t>> F(n + 1);
  F[0].assign(ids[0].size(), 0);
  T.accept.assign(std::max<size_t>(ids[0].size(), 1), 0);
  for (auto& kv : ids[0]) {
    bool ok = allowed ? allowed->count(kv.first) > 0 : true;
    F[0][kv.second] = ok ? 1 : 0;
    T.accept[kv.second] = ok ? 1 : 0;
  }
  T.prefix.assign((size_t)std::max(n, 1) * T.max_q * (T.max_states + 1), 0);
  T.next.assign((size_t)std::max(n, 1) * T.max_q * T.max_states, -1);
  for (int i = 1; i <= n; ++i) {
    int site = i - 1;
    F[i].assign(ids[i].size(), 0);
    for (auto& kv : ids[i]) {
      uint64_t acc = 0;
      size_t base = (size_t)site * T.max_q + kv.second;
      for (int v = 0; v < sp.n_states[site]; ++v) {
        QN q = kv.first;
        const int64_t* dq = sp.qn_of(site, v);
        for (int k = 0; k < sp.n_qn; ++k) q[k] += dq[k];
        int child = ids[site].at(q);
        uint64_t c = F[site][child];
        T.prefix[base * (T.max_states + 1) + v] = acc;
        T.next[base * T.max_states + v] = c ? child : -1;
        ED_REQUIRE(acc + c >= acc, ED_ERR_UNSUPPORTED, "basis dimension overflows 64 bits");
        acc += c;
      }
      for (int v = sp.n_states[site]; v <= T.max_states; ++v) T.prefix[base * (T.max_states + 1) + v] = acc;
      F[i][kv.second] = acc;
    }
  }
  T.root_q = 0;
  T.dim = F[n][0];
  return T;
}

static void build_binom(std::vector<uint64_t>& b) {
  b.assign(65 * 65, 0);
  for (int n = 0; n <= 64; ++n) {
    b[n * 65 + 0] = 1;
    for (int k = 1; k <= n; ++k) {
      unsigned __int128 v = (unsigned __int128)b[(n - 1) * 65 + k - 1] + (n - 1 >= k ? b[(n - 1) * 65 + k] : 0);
      b[n * 65 + k] = v > (unsigned __int128)0xFFFFFFFFFFFFFFFFull ? 0xFFFFFFFFFFFFFFFFull : (uint64_t)v;
    }
  }
}

static void setup_combinadic(ed_basis* b, int n_set) {
  b->kind = ED_BASIS_COMBINADIC;
  b->n_set = n_set;
  build_binom(b->h_binom);
  const int nb = b->space.bits;
  auto C = [&](int n, int k) -> uint64_t { return (k < 0 || k > n || n > 64) ? 0 : b->h_binom[n * 65 + k]; };
  b->dim = (int64_t)C(nb, n_set);
  const int n_chunks = (nb + 7) / 8;
  std::vector<uint64_t> lut((size_t)std::max(n_chunks, 1) * (n_set + 1) * 256, 0);
  for (int c = 0; c < n_chunks; ++c)
    for (int below = 0; below <= n_set; ++below)
      for (int byte = 0; byte < 256; ++byte) {
        uint64_t acc = 0;
        int i = 0;
        for (int q = 0; q < 8; ++q)
          if (byte >> q & 1) {
            acc += C(8 * c + q, below + i + 1);
            ++i;
          }
        lut[((size_t)c * (n_set + 1) + below) * 256 + byte] = acc;
      }
  b->comb_lut.upload(lut);
  b->binom.upload(b->h_binom);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
}

static void setup_dprank(ed_basis* b, const DpTables& T) {
  b->kind = ED_BASIS_DPRANK;
  b->dim = (int64_t)T.dim;
  b->max_q = T.max_q;
  b->max_states = T.max_states;
  b->root_q = T.root_q;
  b->dp_prefix.upload(T.prefix);
  b->dp_next.upload(T.next);
  b->dp_accept.upload(T.accept);
  std::vector<uint8_t> off, w, ns;
  for (int i = 0; i < b->space.n_sites; ++i) {
    off.push_back((uint8_t)b->space.offset[i]);
    w.push_back((uint8_t)b->space.width[i]);
    ns.push_back((uint8_t)b->space.n_states[i]);
  }
  if (off.empty()) { off.push_back(0); w.push_back(0); ns.push_back(1); }
  b->site_off.upload(off);
  b->site_w.upload(w);
  b->site_ns.upload(ns);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
}

LookupDesc ed_basis::desc() const {
  LookupDesc L;
  memset(&L, 0, sizeof(L));
  L.kind = kind;
  L.dim = dim;
  L.words = words.p;
  L.n_bits = space.bits;
  L.n_set = n_set;
  L.n_chunks = (space.bits + 7) / 8;
  L.comb_lut = comb_lut.p;
  L.binom = binom.p;
  L.n_sites = space.n_sites;
  L.max_q = max_q;
  L.max_states = max_states;
  L.root_q = root_q;
  L.dp_prefix = dp_prefix.p;
  L.dp_next = dp_next.p;
  L.dp_accept = dp_accept.p;
  L.site_off = site_off.p;
  L.site_w = site_w.p;
  L.site_ns = site_ns.p;
  return L;
}

// ------------------------------------------------------------------ K1: words of rows [lo, lo+n)
template <int KIND>
__global__ void __launch_bounds__(256) k1_generate_words(LookupDesc L, int64_t lo, int64_t n, uint64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t r = (uint64_t)(lo + i);
    uint64_t s;
    if (KIND == ED_BASIS_FULL) s = r;
    else if (KIND == ED_BASIS_COMBINADIC) s = unrank_combinadic(L, r);
    else s = unrank_dprank(L, r);
    out[i] = s;
  }
}

__global__ void __launch_bounds__(256) k_copy_words(const uint64_t* __restrict__ src, int64_t n, uint64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = src[i];
}

static int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)ed_sm_count() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

void ed_basis_generate_words(ed_basis* b, int64_t lo, int64_t n, uint64_t* dev_out) {
  if (n <= 0) return;
  LookupDesc L = b->desc();
  int grid = grid_for(n, 256);
  switch (b->kind) {
    case ED_BASIS_LIST: ED_LAUNCH(k_copy_words, grid, 256, 0, b->words.p + lo, n, dev_out); break;
    case ED_BASIS_FULL: ED_LAUNCH(k1_generate_words<ED_BASIS_FULL>, grid, 256, 0, L, lo, n, dev_out); break;
    case ED_BASIS_COMBINADIC: ED_LAUNCH(k1_generate_words<ED_BASIS_COMBINADIC>, grid, 256, 0, L, lo, n, dev_out); break;
    default: ED_LAUNCH(k1_generate_words<ED_BASIS_DPRANK>, grid, 256, 0, L, lo, n, dev_out); break;
  }
}

void ed_basis::materialize() {
  if (words_ready) return;
  ed_require_device();
  words.alloc((size_t)std::max<int64_t>(dim, 1));
  ed_basis_generate_words(this, 0, dim, words.p);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  words_ready = true;
}

__global__ void __launch_bounds__(256) k_lookup(LookupDesc L, const uint64_t* __restrict__ keys, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = rank_word_dyn(L, keys[i]);
    out[i] = r >= 0 ? r + 1 : -1;
  }
}

// ------------------------------------------------------------------ C ABI
static void check_br(const ed_space& s, int br_bits) {
  ED_REQUIRE(br_bits > 0, ED_ERR_ARGUMENT, "br_bits must be positive");
  // hilbert_space_representation.jl:63-69/:110/:129: sizeof(BR)*8 <= bitwidth -> ArgumentError
  ED_REQUIRE(br_bits > s.bits, ED_ERR_ARGUMENT,
             "type UInt" + std::to_string(br_bits) + " not enough to represent the hilbert space (need " +
                 std::to_string(s.bits) + " bits)");
  ED_REQUIRE(br_bits <= 64, ED_ERR_UNSUPPORTED, "binary representations wider than 64 bits are not supported by the engine");
}

extern "C" {

int ed_space_create(int32_t n_sites, const int32_t* n_states, const int64_t* qn, int32_t n_qn, ed_space** out) {
  ED_TRY
  ED_REQUIRE(out != nullptr, ED_ERR_ARGUMENT, "out is null");
  ED_REQUIRE(n_sites == 0 || (n_states != nullptr), ED_ERR_ARGUMENT, "n_states is null");
  *out = new ed_space(make_space(n_sites, n_states, qn, n_qn));
  ED_CATCH
}

int ed_space_destroy(ed_space* space) {
  delete space;
  return ED_OK;
}

int ed_space_bitwidth(const ed_space* space, int32_t* bitwidth) {
  ED_TRY
  ED_REQUIRE(space && bitwidth, ED_ERR_ARGUMENT, "null argument");
  *bitwidth = space->bits;
  ED_CATCH
}

int ed_basis_generate(const ed_space* space, const int64_t* allowed_qn, int64_t n_allowed, int32_t br_bits,
                      ed_basis** out) {
  ED_TRY
  ED_REQUIRE(space && out, ED_ERR_ARGUMENT, "null argument");
  check_br(*space, br_bits);
  ed_require_device();
  std::unique_ptr<ed_basis> b(new ed_basis());
  b->space = *space;
  b->br_bits = br_bits;
  const ed_space& sp = b->space;
  b->gen_n_allowed = n_allowed < 0 ? -1 : n_allowed;
  if (n_allowed > 0) b->gen_allowed.assign(allowed_qn, allowed_qn + n_allowed * sp.n_qn);
  if (n_allowed < 0) {
    if (sp.all_pow2) {
      ED_REQUIRE(sp.bits <= 62, ED_ERR_UNSUPPORTED, "full space too large");
      b->kind = ED_BASIS_FULL;
      b->dim = (int64_t)1 << sp.bits;
    } else {
      setup_dprank(b.get(), build_dp(sp, nullptr));
    }
  } else {
    std::set<QN> allowed;
    for (int64_t a = 0; a < n_allowed; ++a) allowed.insert(QN(allowed_qn + a * sp.n_qn, allowed_qn + (a + 1) * sp.n_qn));
    // combinadic fast case: spin-1/2-like sites sharing one U(1) pair, exactly one reachable target
    bool comb = sp.one_bit_sites && sp.n_qn == 1 && sp.n_sites >= 1;
    int64_t q0 = 0, q1 = 0;
    if (comb) {
      q0 = sp.qn_of(0, 0)[0];
      q1 = sp.qn_of(0, 1)[0];
      if (q0 == q1) comb = false;
      for (int i = 1; comb && i < sp.n_sites; ++i)
        if (sp.qn_of(i, 0)[0] != q0 || sp.qn_of(i, 1)[0] != q1) comb = false;
    }
    if (comb) {
      int n_reach = 0, n_set = -1;
      for (auto& q : allowed) {
        int64_t num = q[0] - (int64_t)sp.n_sites * q0;
        int64_t den = q1 - q0;
        if (num % den != 0) continue;
        int64_t k = num / den;
        if (k < 0 || k > sp.n_sites) continue;
        ++n_reach;
        n_set = (int)k;
      }
      if (n_reach == 1) {
        setup_combinadic(b.get(), n_set);
      } else if (n_reach == 0) {
        b->kind = ED_BASIS_LIST;
        b->dim = 0;
        b->words_ready = true;
      } else {
        comb = false;
      }
    }
    if (!comb) setup_dprank(b.get(), build_dp(sp, &allowed));
    if (b->kind == ED_BASIS_DPRANK && b->dim == 0) {
      b->kind = ED_BASIS_LIST;
      b->words_ready = true;
    }
  }
  *out = b.release();
  ED_CATCH
}

int ed_basis_from_list(const ed_space* space, const uint64_t* words, int64_t n, int32_t br_bits, ed_basis** out) {
  ED_TRY
  ED_REQUIRE(space && out && (n == 0 || words), ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(n >= 0, ED_ERR_ARGUMENT, "negative length");
  check_br(*space, br_bits);
  ed_require_device();
  std::vector<uint64_t> w(words, words + n);
  if (!std::is_sorted(w.begin(), w.end())) std::sort(w.begin(), w.end());  // :251-253
  for (int64_t i = 1; i < n; ++i)
    ED_REQUIRE(w[i - 1] != w[i], ED_ERR_ARGUMENT, "vals contains duplicates " + std::to_string(w[i]));  // frozensortedarray.jl:16-20
  std::unique_ptr<ed_basis> b(new ed_basis());
  b->space = *space;
  b->br_bits = br_bits;
  b->kind = ED_BASIS_LIST;
  b->dim = n;
  b->words.alloc((size_t)std::max<int64_t>(n, 1));
  b->words.upload(w.data(), (size_t)n);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  b->words_ready = true;
  *out = b.release();
  ED_CATCH
}

int ed_basis_destroy(ed_basis* basis) {
  delete basis;
  return ED_OK;
}

int ed_basis_dim(const ed_basis* basis, int64_t* dim) {
  ED_TRY
  ED_REQUIRE(basis && dim, ED_ERR_ARGUMENT, "null argument");
  *dim = basis->dim;
  ED_CATCH
}

int ed_basis_kind(const ed_basis* basis, int32_t* kind) {
  ED_TRY
  ED_REQUIRE(basis && kind, ED_ERR_ARGUMENT, "null argument");
  *kind = basis->kind;
  ED_CATCH
}

int ed_basis_download(ed_basis* basis, int64_t lo, int64_t n, uint64_t* words_out) {
  ED_TRY
  ED_REQUIRE(basis && (n == 0 || words_out), ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(lo >= 0 && n >= 0 && lo + n <= basis->dim, ED_ERR_BOUNDS, "range outside the basis");
  if (n == 0) return ED_OK;
  ed_require_device();
  if (basis->words_ready) {
    basis->words.download(words_out, (size_t)n, (size_t)lo);
  } else {
    DevBuf<uint64_t> tmp((size_t)n);
    ed_basis_generate_words(basis, lo, n, tmp.p);
    tmp.download(words_out, (size_t)n);
  }
  ED_CATCH
}

int ed_basis_lookup(ed_basis* basis, const uint64_t* keys, int64_t n, int64_t* index_out) {
  ED_TRY
  ED_REQUIRE(basis && (n == 0 || (keys && index_out)), ED_ERR_ARGUMENT, "null argument");
  if (n == 0) return ED_OK;
  ed_require_device();
  if (basis->kind == ED_BASIS_LIST) basis->materialize();
  DevBuf<uint64_t> dk((size_t)n);
  DevBuf<int64_t> di((size_t)n);
  dk.upload(keys, (size_t)n);
  ED_LAUNCH(k_lookup, grid_for(n, 256), 256, 0, basis->desc(), dk.p, n, di.p);
  di.download(index_out, (size_t)n);
  ED_CATCH
}

int ed_basis_device_words(ed_basis* basis, const uint64_t** dev_words) {
  ED_TRY
  ED_REQUIRE(basis && dev_words, ED_ERR_ARGUMENT, "null argument");
  basis->materialize();
  *dev_words = basis->words.p;
  ED_CATCH
}

}  // extern "C"
