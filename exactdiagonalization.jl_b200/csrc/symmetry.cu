// Symmetry: group action as byte-chunk bit shuffles, K5 (representative filter + compaction),
// on-demand basis_mapping_index / basis_mapping_amplitude, K8 vector reduce / unreduce.
//
// Replaces (reference, /root/reference/src):
//   Symmetry/symmetry_apply.jl:65-92, Symmetry/bitflipsymmetry.jl:23-35     symmetry_apply on bit words
//   Symmetry/symmetry_reduce_generic.jl:22-255                               symmetry_reduce_serial/parallel
//   Symmetry/reduced_hilbert_space_representation.jl:13-22                   ReducedHilbertSpaceRepresentation
//   Symmetry/symmetry_reduce.jl:38-153, 208-225                              vector symmetry_reduce / symmetry_unreduce
// The reference stores 32 bytes per PARENT state (index + amplitude + visited).  Here only the
// representatives, their orbit size and their last stabilising element are kept (12 bytes per REDUCED
// state); index/amplitude of any parent word are recomputed in registers by an orbit-minimum search.
#include <algorithm>
#include <cmath>
#include <cub/cub.cuh>

#include "ed_device.cuh"

static const double kRtolDefault = 1.4901161193847656e-08;  // Base.rtoldefault(Float64)

// ------------------------------------------------------------------ host: group bookkeeping
static void analyse_group(ed_symmetry* s) {
  const int G = s->n_ops, n = s->n_sites;
  std::map<std::vector<int32_t>, int> index;
  bool distinct = true;
  for (int g = 0; g < G; ++g) {
    std::vector<int32_t> key(s->perm.begin() + (size_t)g * n, s->perm.begin() + (size_t)(g + 1) * n);
    key.push_back(s->flip[g]);
    if (!index.emplace(key, g).second) distinct = false;
  }
  s->is_group = false;
  s->inverse.assign(G, -1);
  if (!distinct) return;
  // identity first?
  for (int i = 0; i < n; ++i)
    if (s->perm[i] != i) return;
  if (s->flip[0]) return;
  // closure: (a*b)(i) = a.map[b.map[i]], flips add
  std::vector<int32_t> key(n + 1);
  for (int a = 0; a < G; ++a)
    for (int b = 0; b < G; ++b) {
      for (int i = 0; i < n; ++i) key[i] = s->perm[(size_t)a * n + s->perm[(size_t)b * n + i]];
      key[n] = s->flip[a] ^ s->flip[b];
      auto it = index.find(key);
      if (it == index.end()) return;
      if (it->second == 0) s->inverse[a] = b;
    }
  for (int a = 0; a < G; ++a)
    if (s->inverse[a] < 0) return;
  s->is_group = true;
}

void ed_symdev_build(const ed_space& space, const ed_symmetry& sym, double tol, SymDev* out) {
  ED_REQUIRE(sym.n_sites == space.n_sites, ED_ERR_ARGUMENT, "symmetry acts on a different number of sites than the space has");
  const int G = sym.n_ops, n = space.n_sites;
  const int n_chunks = std::max(1, (space.bits + 7) / 8);
  std::vector<uint64_t> lut((size_t)G * n_chunks * 256, 0);
  // target position of every bit under g: bit (offset[i] + k) -> offset[map[i]] + k   (symmetry_apply.jl:88-90)
  std::vector<int> tgt(64);
  for (int g = 0; g < G; ++g) {
    for (int i = 0; i < n; ++i) {
      int j = sym.perm[(size_t)g * n + i];
      ED_REQUIRE(j >= 0 && j < n, ED_ERR_ARGUMENT, "permutation entry out of range");
      ED_REQUIRE(space.width[i] == space.width[j], ED_ERR_ARGUMENT, "permutation maps sites of different bit width onto each other");
      for (int k = 0; k < space.width[i]; ++k) tgt[space.offset[i] + k] = space.offset[j] + k;
    }
    for (int c = 0; c < n_chunks; ++c) {
      uint64_t chunk_image_all = 0;
      for (int q = 0; q < 8; ++q)
        if (8 * c + q < space.bits) chunk_image_all |= 1ull << tgt[8 * c + q];
      for (int byte = 0; byte < 256; ++byte) {
        uint64_t im = 0;
        for (int q = 0; q < 8; ++q)
          if ((byte >> q & 1) && 8 * c + q < space.bits) im |= 1ull << tgt[8 * c + q];
        // GlobalBitFlip(true): mask & ~b (bitflipsymmetry.jl:31-32) -- folded per chunk, images are disjoint
        if (sym.flip[g]) im ^= chunk_image_all;
        lut[((size_t)g * n_chunks + c) * 256 + byte] = im;
      }
    }
  }
  // the same action in 6-bit chunks for the word-parallel orbit sweep (k6b_canonicalize)
  int n_chunks6 = std::max(1, (space.bits + 5) / 6);
  // padded to the kernel's compile-time chunk counts (4, 6, 8, 11); padding chunks only ever see the value 0 -> image 0
  n_chunks6 = n_chunks6 <= 4 ? 4 : n_chunks6 <= 6 ? 6 : n_chunks6 <= 8 ? 8 : 11;
  std::vector<uint64_t> lut6((size_t)G * n_chunks6 * 64, 0);
  for (int g = 0; g < G; ++g) {
    for (int i = 0; i < n; ++i) {
      int j = sym.perm[(size_t)g * n + i];
      for (int k = 0; k < space.width[i]; ++k) tgt[space.offset[i] + k] = space.offset[j] + k;
    }
    for (int c = 0; c < n_chunks6; ++c) {
      uint64_t all = 0;
      for (int q = 0; q < 6; ++q)
        if (6 * c + q < space.bits) all |= 1ull << tgt[6 * c + q];
      for (int v = 0; v < 64; ++v) {
        uint64_t im = 0;
        for (int q = 0; q < 6; ++q)
          if ((v >> q & 1) && 6 * c + q < space.bits) im |= 1ull << tgt[6 * c + q];
        if (sym.flip[g]) im ^= all;
        lut6[((size_t)g * n_chunks6 + c) * 64 + v] = im;
      }
    }
  }
  // ---- translation factorisation: G = union of cosets T p_j with T the n1 x n2 lattice translations ------------------
  out->tr_on = false;
  if (sym.is_group && space.bits == n && n >= 4 && n <= 60 && G >= 8 && !getenv("EDCUDA_K6_NOTR")) {
    std::map<std::vector<int32_t>, int> index;
    std::vector<int32_t> key(n + 1);
    for (int g = 0; g < G; ++g) {
      for (int i = 0; i < n; ++i) key[i] = sym.perm[(size_t)g * n + i];
      key[n] = sym.flip[g];
      index.emplace(key, g);
    }
    auto tmap = [&](int n1, int n2, int a, int b, int i) { const int x = i % n1, y = i / n1; return (x + a) % n1 + n1 * ((y + b) % n2); };
    int best_n1 = 0;
    for (int n1 = n; n1 >= 2 && !best_n1; --n1) {
      if (n % n1) continue;
      const int n2 = n / n1;
      bool ok = true;
      for (int gen = 0; gen < 2 && ok; ++gen) {
        if (gen == 1 && n2 == 1) break;
        for (int i = 0; i < n; ++i) key[i] = tmap(n1, n2, gen == 0 ? 1 : 0, gen == 1 ? 1 : 0, i);
        key[n] = 0;
        ok = index.count(key) != 0;
      }
      if (ok) best_n1 = n1;
    }
    if (best_n1 && G % n == 0) {
      const int n1 = best_n1, n2 = n / n1, nt = n, ncos = G / nt;
      std::vector<int> coset_of(G, -1);
      std::vector<int> reps;
      std::vector<int32_t> tinv((size_t)ncos * nt, 0);
      bool ok = true;
      for (int g = 0; g < G && ok; ++g) {
        if (coset_of[g] >= 0) continue;
        const int j = (int)reps.size();
        if (j >= ncos) { ok = false; break; }
        reps.push_back(g);
        for (int b = 0; b < n2 && ok; ++b)
          for (int a = 0; a < n1 && ok; ++a) {
            for (int i = 0; i < n; ++i) key[i] = tmap(n1, n2, a, b, sym.perm[(size_t)g * n + i]);   // (t p)(i) = t.map[p.map[i]]
            key[n] = sym.flip[g];
            auto it = index.find(key);
            if (it == index.end() || coset_of[it->second] >= 0) { ok = false; break; }
            coset_of[it->second] = j;
            tinv[((size_t)j * n2 + b) * n1 + a] = sym.inverse[it->second];
          }
      }
      if (ok && (int)reps.size() == ncos) {
        std::vector<uint64_t> tl((size_t)ncos * n_chunks6 * 64);
        for (int j = 0; j < ncos; ++j)
          std::copy(lut6.begin() + (size_t)reps[j] * n_chunks6 * 64, lut6.begin() + (size_t)(reps[j] + 1) * n_chunks6 * 64,
                    tl.begin() + (size_t)j * n_chunks6 * 64);
        out->tr_lut6.upload(tl);
        out->tr_inv.upload(tinv);
        out->tr_on = true;
        out->tr_n1 = n1; out->tr_n2 = n2; out->tr_ncos = ncos;
      }
    }
  }
  out->n_chunks6 = n_chunks6;
  out->lut6.upload(lut6);
  out->n_ops = G;
  out->n_chunks = n_chunks;
  out->fullmask = space.fullmask();
  out->lut.upload(lut);
  {
    // the reference never reads the first element's amplitude: basis_phases starts as ones and its loop begins at the
    // second element (symmetry_reduce_generic.jl:47, 56-78), so the identity always contributes phase 1
    std::vector<double> chi = sym.chi;
    if (chi.size() >= 2) { chi[0] = 1.0; chi[1] = 0.0; }
    out->chi.upload(chi);
  }
  std::vector<int32_t> inv = sym.inverse;
  if ((int)inv.size() != G) inv.assign(G, 0);
  out->inverse.upload(inv);
  std::vector<uint8_t> one(G);
  for (int g = 0; g < G; ++g) {
    // isapprox(ampl*sgn, one; atol=tol): |chi - 1| <= tol   (symmetry_reduce_generic.jl:62)
    double dr = sym.chi[2 * g] - 1.0, di = sym.chi[2 * g + 1];
    if (g == 0) dr = di = 0.0;      // element 0 (identity): amplitude never read by the reference
    one[g] = std::sqrt(dr * dr + di * di) <= tol ? 1 : 0;
  }
  out->chi_is_one.upload(one);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
}

SymDesc ed_rbasis::symdesc() const {
  SymDesc S;
  S.n_ops = symdev.n_ops;
  S.n_chunks = symdev.n_chunks;
  S.lut = symdev.lut.p;
  S.chi = symdev.chi.p;
  S.inverse = symdev.inverse.p;
  S.chi_is_one = symdev.chi_is_one.p;
  return S;
}

RLookupDesc ed_rbasis::rdesc() const {
  RLookupDesc R;
  R.words = words.p;
  R.orbit_size = orbit_size.p;
  R.last_stab = last_stab.p;
  R.bucket_start = bucket_start.p;
  R.bucket_shift = bucket_shift;
  R.n_buckets = n_buckets;
  R.dim = dim;
  R.hash = hash.n ? hash.p : nullptr;
  R.hash_shift = hash_shift;
  R.idx_bits = idx_bits;
  return R;
}
static RLookupDesc rdesc(const ed_rbasis* r) { return r->rdesc(); }

// ------------------------------------------------------------------ K5: representative filter
// One thread per parent word of the chunk [lo, lo+n).  A word survives iff no group element maps it
// to a smaller word and every stabilising element has chi ~ 1 (symmetry_reduce_generic.jl:56-69).
// For survivors the stabiliser is counted (orbit size = |G| / |Stab|, the number of distinct images
// the reference counts with a Dict at :74-79) and the last stabilising index recorded.
__global__ void __launch_bounds__(256)
k5_filter(LookupDesc L, const uint64_t* __restrict__ parent_words, int64_t lo, int64_t n, SymDesc S,
          uint64_t* __restrict__ cand, uint8_t* __restrict__ flag, uint16_t* __restrict__ orbit,
          uint16_t* __restrict__ last_stab, int* __restrict__ key_error) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t s;
    if (parent_words) s = __ldg(parent_words + lo + i);
    else if (L.kind == ED_BASIS_FULL) s = (uint64_t)(lo + i);
    else if (L.kind == ED_BASIS_COMBINADIC) s = unrank_combinadic(L, (uint64_t)(lo + i));
    else s = unrank_dprank(L, (uint64_t)(lo + i));
    bool ok = true;
    int stab = 1, last = 0;
#pragma unroll 1
    for (int g = 1; g < S.n_ops; ++g) {
      uint64_t im = sym_apply(S, g, s);
      if (im < s) { ok = false; break; }
      if (im == s) {
        if (!S.chi_is_one[g]) { ok = false; break; }
        ++stab;
        last = g;
      }
    }
    bool outside = false;
    if (ok && L.kind != ED_BASIS_FULL) {
      // the reference looks every image of a representative up in the parent and throws KeyError when absent (:81)
      for (int g = 1; g < S.n_ops; ++g)
        if (!in_basis_dyn(L, sym_apply(S, g, s))) outside = true;
    }
    cand[i] = s;
    flag[i] = ok ? 1 : 0;
    orbit[i] = (uint16_t)(S.n_ops / stab);
    last_stab[i] = (uint16_t)last;
    if (ok && outside) atomicExch(key_error, 1);
  }
}

__global__ void __launch_bounds__(256)
k_bucket_start(const uint64_t* __restrict__ words, int64_t dim, int shift, int64_t n_buckets, uint32_t* __restrict__ start) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b <= n_buckets; b += (int64_t)gridDim.x * blockDim.x) {
    // first index whose (word >> shift) >= b
    int64_t lo = 0, hi = dim;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((words[mid] >> shift) < (uint64_t)b) lo = mid + 1; else hi = mid;
    }
    start[b] = (uint32_t)lo;
  }
}

template <typename T>
static void append_selected(DevBuf<T>& dst, int64_t& cap, int64_t used, const T* src, const uint8_t* flags, int64_t n,
                            DevBuf<unsigned char>& tmp, int* d_count) {
  // grow
  if (used + n > cap) {
    int64_t ncap = std::max<int64_t>(cap * 2, used + n);
    DevBuf<T> bigger((size_t)ncap);
    if (used) ED_CUDA(cudaMemcpyAsync(bigger.p, dst.p, (size_t)used * sizeof(T), cudaMemcpyDeviceToDevice, ed_stream()));
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    dst = std::move(bigger);
    cap = ncap;
  }
  size_t bytes = 0;
  cub::DeviceSelect::Flagged(nullptr, bytes, src, flags, dst.p + used, d_count, n, ed_stream());
  if (tmp.n < bytes) tmp.alloc(bytes);
  cub::DeviceSelect::Flagged(tmp.p, bytes, src, flags, dst.p + used, d_count, n, ed_stream());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
}

static void build_rbasis(ed_rbasis* r) {
  ed_basis* parent = r->parent;
  const int64_t Dp = parent->dim;
  SymDesc S = r->symdesc();
  LookupDesc L = parent->desc();
  const uint64_t* pw = nullptr;
  if (parent->kind == ED_BASIS_LIST || parent->words_ready) {
    parent->materialize();
    pw = parent->words.p;
    L = parent->desc();
  }
  const int64_t CH = 1ll << 24;
  const int64_t chunk = std::min<int64_t>(std::max<int64_t>(Dp, 1), CH);
  DevBuf<uint64_t> cand((size_t)chunk);
  DevBuf<uint8_t> flag((size_t)chunk);
  DevBuf<uint16_t> orb((size_t)chunk), lst((size_t)chunk);
  DevBuf<int> d_count(1), d_err(1);
  DevBuf<unsigned char> tmp;
  ED_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), ed_stream()));
  int64_t cap_w = 0, cap_o = 0, cap_l = 0, used = 0;
  int64_t est = Dp / std::max(1, r->sym.n_ops - 1) + 1024;
  r->words.alloc((size_t)est); cap_w = est;
  r->orbit_size.alloc((size_t)est); cap_o = est;
  r->last_stab.alloc((size_t)est); cap_l = est;
  for (int64_t lo = 0; lo < Dp; lo += chunk) {
    const int64_t n = std::min(chunk, Dp - lo);
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ed_sm_count() * 16);
    ED_LAUNCH(k5_filter, grid, 256, 0, L, pw, lo, n, S, cand.p, flag.p, orb.p, lst.p, d_err.p);
    append_selected(r->words, cap_w, used, cand.p, flag.p, n, tmp, d_count.p);
    append_selected(r->orbit_size, cap_o, used, orb.p, flag.p, n, tmp, d_count.p);
    append_selected(r->last_stab, cap_l, used, lst.p, flag.p, n, tmp, d_count.p);
    int c = 0;
    d_count.download(&c, 1);
    used += c;
  }
  int err = 0;
  d_err.download(&err, 1);
  ED_REQUIRE(err == 0, ED_ERR_KEY, "a symmetry image of a representative is not in the parent basis (KeyError in the reference)");
  r->dim = used;
  ed_rbasis_finish_index(r);
}

__global__ void __launch_bounds__(256)
k_rhash_insert(const uint64_t* __restrict__ words, int64_t dim, unsigned long long* __restrict__ hash, int hash_shift, int idx_bits) {
  const uint64_t slot_mask = ~0ull >> hash_shift;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < dim; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = words[i];
    const unsigned long long e = (key << idx_bits) | (unsigned long long)i;
    uint64_t s = (key * 0x9E3779B97F4A7C15ull) >> hash_shift;
    while (atomicCAS(hash + s, ~0ull, e) != ~0ull) s = (s + 1) & slot_mask;
  }
}

// indices over the (ascending) representatives for the reduced lookup: buckets over the top bits (narrow the binary search)
// and, when a word and its index fit 64 bits together, the hash table rank_reduced prefers (EDCUDA_RBASIS_HASH=0: none,
// saves 16 bytes per representative)
void ed_rbasis_finish_index(ed_rbasis* r) {
  r->hash.release();
  r->hash_shift = r->idx_bits = 0;
  {
    const int bits_w = r->parent->space.bits;
    const char* e = getenv("EDCUDA_RBASIS_HASH");
    if (r->dim > 0 && bits_w >= 1 && bits_w <= 63 && !(e && atoi(e) == 0)) {
      const int idx_bits = 64 - bits_w;
      const bool fits = idx_bits >= 63 || (uint64_t)r->dim <= (1ull << idx_bits) - 1;     // the index field is never all ones
      if (fits) {
        int lg = 4;
        while ((1ll << lg) < 2 * r->dim) ++lg;
        r->hash.alloc((size_t)1 << lg);
        ED_CUDA(cudaMemsetAsync(r->hash.p, 0xFF, sizeof(unsigned long long) << lg, ed_stream()));
        r->hash_shift = 64 - lg;
        r->idx_bits = idx_bits;
        const int grid = (int)std::min<int64_t>((r->dim + 255) / 256, (int64_t)ed_sm_count() * 16);
        ED_LAUNCH(k_rhash_insert, grid, 256, 0, r->words.p, r->dim, r->hash.p, r->hash_shift, r->idx_bits);
      }
    }
  }
  const int64_t used = r->dim;
  ED_REQUIRE(used < (1ll << 32), ED_ERR_UNSUPPORTED, "reduced dimension exceeds 2^32");
  const int bits = r->parent->space.bits;
  int bb = std::min(bits, 22);
  while (bb > 0 && (1ll << bb) > std::max<int64_t>(4 * used, 16)) --bb;
  r->bucket_shift = bits - bb;
  r->n_buckets = 1ll << bb;
  r->bucket_start.alloc((size_t)r->n_buckets + 2);
  if (used == 0) {
    ED_CUDA(cudaMemsetAsync(r->bucket_start.p, 0, (size_t)(r->n_buckets + 2) * sizeof(uint32_t), ed_stream()));
  } else {
    int grid = (int)std::min<int64_t>((r->n_buckets + 256) / 256, (int64_t)ed_sm_count() * 16);
    ED_LAUNCH(k_bucket_start, grid, 256, 0, r->words.p, used, r->bucket_shift, r->n_buckets, r->bucket_start.p);
  }
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
}

// ------------------------------------------------------------------ mapping kernels
__global__ void __launch_bounds__(256)
k_map_words(LookupDesc L, SymDesc S, RLookupDesc R, const uint64_t* __restrict__ words, int64_t lo, int64_t n,
            int64_t* __restrict__ idx_out, double* __restrict__ amp_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t s;
    if (words) s = words[i];
    else if (L.kind == ED_BASIS_FULL) s = (uint64_t)(lo + i);
    else if (L.kind == ED_BASIS_COMBINADIC) s = unrank_combinadic(L, (uint64_t)(lo + i));
    else s = unrank_dprank(L, (uint64_t)(lo + i));
    int64_t j = -1;
    c128 a = make_c128(0.0, 0.0);
    if (rank_word_dyn(L, s) >= 0) j = reduced_map_word(S, R, s, &a);
    if (j < 0) a = make_c128(0.0, 0.0);
    idx_out[i] = j >= 0 ? j + 1 : -1;
    amp_out[2 * i] = a.re;
    amp_out[2 * i + 1] = a.im;
  }
}

// K8: small[idx[p]] += conj(amp[p]) * large[p]   (symmetry_reduce.jl:38-56)
template <typename VecT>
__global__ void __launch_bounds__(256)
k8_vector_reduce(LookupDesc L, SymDesc S, RLookupDesc R, const uint64_t* __restrict__ parent_words, int64_t n,
                 const VecT* __restrict__ large, double* __restrict__ small) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    c128 a;
    int64_t j = reduced_map_word(S, R, parent_words[i], &a);
    if (j < 0) continue;
    c128 p = cmul(cconj(a), to_c128(large[i]));
    atomicAdd(small + 2 * j, p.re);
    atomicAdd(small + 2 * j + 1, p.im);
  }
}

// large[p] = amp[p] * small[idx[p]]   (symmetry_reduce.jl:208-225)
template <typename VecT>
__global__ void __launch_bounds__(256)
k8_vector_unreduce(LookupDesc L, SymDesc S, RLookupDesc R, const uint64_t* __restrict__ parent_words, int64_t n,
                   const VecT* __restrict__ small, c128* __restrict__ large) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    c128 a;
    int64_t j = reduced_map_word(S, R, parent_words[i], &a);
    c128 out = make_c128(0.0, 0.0);
    if (j >= 0) {
      out = cmul(a, to_c128(small[j]));
    }
    st_val(large + i, out);
  }
}

static int grid_rows(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ed_sm_count() * 16));
}

// ------------------------------------------------------------------ C ABI
extern "C" {

int ed_symmetry_create(int32_t n_ops, int32_t n_sites, const int32_t* perm, const uint8_t* flip, const double* chi,
                       ed_symmetry** out) {
  ED_TRY
  ED_REQUIRE(out && perm && chi, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(n_ops >= 1 && n_ops < 65535, ED_ERR_ARGUMENT, "n_ops must be in 1..65534");
  ED_REQUIRE(n_sites >= 1 && n_sites <= ED_MAX_SITES, ED_ERR_ARGUMENT, "n_sites must be in 1..64");
  std::unique_ptr<ed_symmetry> s(new ed_symmetry());
  s->n_ops = n_ops;
  s->n_sites = n_sites;
  s->perm.assign(perm, perm + (size_t)n_ops * n_sites);
  s->flip.assign(n_ops, 0);
  if (flip)
    for (int g = 0; g < n_ops; ++g) s->flip[g] = flip[g] ? 1 : 0;
  s->chi.assign(chi, chi + (size_t)2 * n_ops);
  for (int g = 0; g < n_ops; ++g) {
    std::vector<char> seen(n_sites, 0);
    for (int i = 0; i < n_sites; ++i) {
      int j = perm[(size_t)g * n_sites + i];
      ED_REQUIRE(j >= 0 && j < n_sites && !seen[j], ED_ERR_ARGUMENT, "element " + std::to_string(g) + " is not a permutation");
      seen[j] = 1;
    }
    // isapprox(abs(y), one(abs(y))) with the default rtol (symmetry_reduce_generic.jl:27-29)
    double a = std::hypot(chi[2 * g], chi[2 * g + 1]);
    ED_REQUIRE(std::fabs(a - 1.0) <= kRtolDefault * std::max(a, 1.0), ED_ERR_ARGUMENT, "all amplitudes need to have norm 1");
  }
  analyse_group(s.get());
  *out = s.release();
  ED_CATCH
}

int ed_symmetry_destroy(ed_symmetry* sym) {
  delete sym;
  return ED_OK;
}

__global__ void k_sym_apply(SymDesc S, int g, const uint64_t* __restrict__ in, int64_t n, uint64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = sym_apply(S, g, in[i]);
}

int ed_symmetry_apply(const ed_space* space, const ed_symmetry* sym, int32_t g, const uint64_t* words, int64_t n,
                      uint64_t* images_out) {
  ED_TRY
  ED_REQUIRE(space && sym && (n == 0 || (words && images_out)), ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(g >= 0 && g < sym->n_ops, ED_ERR_BOUNDS, "group element index out of range");
  if (n == 0) return ED_OK;
  ed_require_device();
  SymDev sd;
  ed_symdev_build(*space, *sym, kRtolDefault, &sd);
  SymDesc S;
  S.n_ops = sd.n_ops; S.n_chunks = sd.n_chunks; S.lut = sd.lut.p; S.chi = sd.chi.p; S.inverse = sd.inverse.p; S.chi_is_one = sd.chi_is_one.p;
  DevBuf<uint64_t> din((size_t)n), dout((size_t)n);
  din.upload(words, (size_t)n);
  ED_LAUNCH(k_sym_apply, grid_rows(n), 256, 0, S, g, din.p, n, dout.p);
  dout.download(images_out, (size_t)n);
  ED_CATCH
}

int ed_symmetry_reduce(ed_basis* parent, const ed_symmetry* sym, double tol, ed_rbasis** out) {
  ED_TRY
  ED_REQUIRE(parent && sym && out, ED_ERR_ARGUMENT, "null argument");
  if (tol < 0) tol = kRtolDefault;
  // element 1 of the reference's list is never applied (loop from 2, :56): it has to be the identity
  bool id0 = sym->flip[0] == 0;
  for (int i = 0; i < sym->n_sites; ++i) id0 = id0 && sym->perm[i] == i;
  ED_REQUIRE(id0, ED_ERR_ARGUMENT, "the first symmetry operation must be the identity");
  ED_REQUIRE(sym->is_group, ED_ERR_UNSUPPORTED,
             "the symmetry operations must be distinct and closed under composition (a group): the engine "
             "recomputes basis_mapping_* by orbit-minimum search instead of storing it per parent state");
  ed_require_device();
  std::unique_ptr<ed_rbasis> r(new ed_rbasis());
  r->parent = parent;
  r->sym = *sym;
  r->tol = tol;
  ed_symdev_build(parent->space, *sym, tol, &r->symdev);
  build_rbasis(r.get());
  *out = r.release();
  ED_CATCH
}

int ed_rbasis_destroy(ed_rbasis* rbasis) {
  delete rbasis;
  return ED_OK;
}

int ed_rbasis_dim(const ed_rbasis* rbasis, int64_t* dim) {
  ED_TRY
  ED_REQUIRE(rbasis && dim, ED_ERR_ARGUMENT, "null argument");
  *dim = rbasis->dim;
  ED_CATCH
}

int ed_rbasis_download(ed_rbasis* rbasis, int64_t lo, int64_t n, uint64_t* words_out) {
  ED_TRY
  ED_REQUIRE(rbasis && (n == 0 || words_out), ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(lo >= 0 && n >= 0 && lo + n <= rbasis->dim, ED_ERR_BOUNDS, "range outside the reduced basis");
  rbasis->words.download(words_out, (size_t)n, (size_t)lo);
  ED_CATCH
}

int ed_rbasis_orbit_sizes(ed_rbasis* rbasis, int64_t lo, int64_t n, int32_t* sizes_out) {
  ED_TRY
  ED_REQUIRE(rbasis && (n == 0 || sizes_out), ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(lo >= 0 && n >= 0 && lo + n <= rbasis->dim, ED_ERR_BOUNDS, "range outside the reduced basis");
  std::vector<uint16_t> tmp((size_t)n);
  rbasis->orbit_size.download(tmp.data(), (size_t)n, (size_t)lo);
  for (int64_t i = 0; i < n; ++i) sizes_out[i] = tmp[i];
  ED_CATCH
}

int ed_rbasis_mapping(ed_rbasis* rbasis, const uint64_t* parent_words, int64_t n, int64_t* index_out, double* amplitude_out) {
  ED_TRY
  ED_REQUIRE(rbasis && (n == 0 || (parent_words && index_out && amplitude_out)), ED_ERR_ARGUMENT, "null argument");
  if (n == 0) return ED_OK;
  ed_require_device();
  if (rbasis->parent->kind == ED_BASIS_LIST) rbasis->parent->materialize();
  DevBuf<uint64_t> dw((size_t)n);
  DevBuf<int64_t> di((size_t)n);
  DevBuf<double> da((size_t)2 * n);
  dw.upload(parent_words, (size_t)n);
  ED_LAUNCH(k_map_words, grid_rows(n), 256, 0, rbasis->parent->desc(), rbasis->symdesc(), rdesc(rbasis), dw.p, (int64_t)0, n, di.p, da.p);
  di.download(index_out, (size_t)n);
  da.download(amplitude_out, (size_t)2 * n);
  ED_CATCH
}

int ed_rbasis_mapping_rows(ed_rbasis* rbasis, int64_t lo, int64_t n, int64_t* index_out, double* amplitude_out) {
  ED_TRY
  ED_REQUIRE(rbasis && (n == 0 || (index_out && amplitude_out)), ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(lo >= 0 && n >= 0 && lo + n <= rbasis->parent->dim, ED_ERR_BOUNDS, "range outside the parent basis");
  if (n == 0) return ED_OK;
  ed_require_device();
  ed_basis* p = rbasis->parent;
  const uint64_t* pw = nullptr;
  if (p->kind == ED_BASIS_LIST || p->words_ready) { p->materialize(); pw = p->words.p + lo; }
  DevBuf<int64_t> di((size_t)n);
  DevBuf<double> da((size_t)2 * n);
  ED_LAUNCH(k_map_words, grid_rows(n), 256, 0, p->desc(), rbasis->symdesc(), rdesc(rbasis), pw, lo, n, di.p, da.p);
  di.download(index_out, (size_t)n);
  da.download(amplitude_out, (size_t)2 * n);
  ED_CATCH
}

int ed_vector_reduce(ed_rbasis* rbasis, void* small_out, int64_t n_small, const void* large, int64_t n_large,
                     int32_t large_dtype, int32_t accumulate) {
  ED_TRY
  ED_REQUIRE(rbasis, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(n_large == rbasis->parent->dim, ED_ERR_DIMENSION_MISMATCH,
             "Dimension of the input vector should match the larger representation");
  ED_REQUIRE(n_small == rbasis->dim, ED_ERR_DIMENSION_MISMATCH,
             "Dimension of the output vector should match the smaller representation");
  ED_REQUIRE(large_dtype == ED_F64 || large_dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  if (n_small == 0) return ED_OK;
  ED_REQUIRE(small_out && large, ED_ERR_ARGUMENT, "null vector");
  ed_require_device();
  ed_basis* p = rbasis->parent;
  p->materialize();
  Staged sl(large, (size_t)n_large * (large_dtype == ED_C128 ? 16 : 8), true, false);
  Staged ss(small_out, (size_t)n_small * 16, accumulate != 0, true);
  if (!accumulate) ED_CUDA(cudaMemsetAsync(ss.dev, 0, (size_t)n_small * 16, ed_stream()));
  if (large_dtype == ED_C128)
    ED_LAUNCH(k8_vector_reduce<c128>, grid_rows(n_large), 256, 0, p->desc(), rbasis->symdesc(), rdesc(rbasis), p->words.p, n_large,
              reinterpret_cast<const c128*>(sl.dev), reinterpret_cast<double*>(ss.dev));
  else
    ED_LAUNCH(k8_vector_reduce<double>, grid_rows(n_large), 256, 0, p->desc(), rbasis->symdesc(), rdesc(rbasis), p->words.p, n_large,
              reinterpret_cast<const double*>(sl.dev), reinterpret_cast<double*>(ss.dev));
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  ss.finish();
  sl.finish();
  ED_CATCH
}

int ed_vector_unreduce(ed_rbasis* rbasis, void* large_out, int64_t n_large, const void* small, int64_t n_small,
                       int32_t small_dtype) {
  ED_TRY
  ED_REQUIRE(rbasis, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(n_small == rbasis->dim, ED_ERR_DIMENSION_MISMATCH,
             "Dimension of the input vector should match the reduced representation");
  ED_REQUIRE(n_large == rbasis->parent->dim, ED_ERR_DIMENSION_MISMATCH,
             "Dimension of the output vector should match the larger representation");
  ED_REQUIRE(small_dtype == ED_F64 || small_dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  if (n_large == 0) return ED_OK;
  ED_REQUIRE(large_out && (small || n_small == 0), ED_ERR_ARGUMENT, "null vector");
  ed_require_device();
  ed_basis* p = rbasis->parent;
  p->materialize();
  Staged ss(small, (size_t)n_small * (small_dtype == ED_C128 ? 16 : 8), true, false);
  Staged sl(large_out, (size_t)n_large * 16, false, true);
  if (small_dtype == ED_C128)
    ED_LAUNCH(k8_vector_unreduce<c128>, grid_rows(n_large), 256, 0, p->desc(), rbasis->symdesc(), rdesc(rbasis), p->words.p, n_large,
              reinterpret_cast<const c128*>(ss.dev), reinterpret_cast<c128*>(sl.dev));
  else
    ED_LAUNCH(k8_vector_unreduce<double>, grid_rows(n_large), 256, 0, p->desc(), rbasis->symdesc(), rdesc(rbasis), p->words.p, n_large,
              reinterpret_cast<const double*>(ss.dev), reinterpret_cast<c128*>(sl.dev));
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  sl.finish();
  ss.finish();
  ED_CATCH
}

}  // extern "C"
