// Device-side building blocks: word -> index ranking, group action, reduced-space mapping.
#pragma once
#include "ed_internal.cuh"

// ------------------------------------------------------------------ scalar helpers
struct c128 {
  double re, im;
};
__host__ __device__ __forceinline__ c128 make_c128(double r, double i) { c128 z; z.re = r; z.im = i; return z; }
__device__ __forceinline__ c128 cmul(c128 a, c128 b) {
  return make_c128(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
__device__ __forceinline__ c128 cadd(c128 a, c128 b) { return make_c128(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ c128 cconj(c128 a) { return make_c128(a.re, -a.im); }
__device__ __forceinline__ c128 cscale(c128 a, double s) { return make_c128(a.re * s, a.im * s); }
__device__ __forceinline__ c128 cinv(c128 a) {
  double d = a.re * a.re + a.im * a.im;
  return make_c128(a.re / d, -a.im / d);
}

// VecT/AmpT arithmetic used by the templated kernels: acc += amp * x
__device__ __forceinline__ void fma_acc(double& acc, double a, double x) { acc += a * x; }
__device__ __forceinline__ void fma_acc(c128& acc, double a, c128 x) { acc.re += a * x.re; acc.im += a * x.im; }
__device__ __forceinline__ void fma_acc(c128& acc, c128 a, c128 x) {
  acc.re += a.re * x.re - a.im * x.im;
  acc.im += a.re * x.im + a.im * x.re;
}
__device__ __forceinline__ c128 to_c128(double v) { return make_c128(v, 0.0); }
__device__ __forceinline__ c128 to_c128(c128 v) { return v; }
__device__ __forceinline__ double vzero(double*) { return 0.0; }
__device__ __forceinline__ c128 vzero(c128*) { return make_c128(0.0, 0.0); }

__device__ __forceinline__ c128 ldg_c128(const c128* p) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return make_c128(v.x, v.y);
}
__device__ __forceinline__ double ldg_val(const double* p) { return __ldg(p); }
__device__ __forceinline__ c128 ldg_val(const c128* p) { return ldg_c128(p); }
__device__ __forceinline__ void st_val(double* p, double v) { *p = v; }
__device__ __forceinline__ void st_val(c128* p, c128 v) { *reinterpret_cast<double2*>(p) = make_double2(v.re, v.im); }

// conj(x) * y accumulated as (re, im)
__device__ __forceinline__ void dot_acc(double& re, double& im, double x, double y) { re += x * y; }
__device__ __forceinline__ void dot_acc(double& re, double& im, c128 x, c128 y) {
  re += x.re * y.re + x.im * y.im;
  im += x.re * y.im - x.im * y.re;
}

// ------------------------------------------------------------------ ranking (word -> 0-based index, -1 = miss)
// FrozenSortedArrayIndex.get (src/frozensortedarray.jl:29-48): searchsortedfirst + equality test.
__device__ __forceinline__ int64_t rank_list(const uint64_t* __restrict__ w, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(w + mid) < key) lo = mid + 1; else hi = mid;
  }
  return (lo < n && __ldg(w + lo) == key) ? lo : -1;
}

// Combinatorial number system: index = sum_j C(p_j, j+1) over set-bit positions p_0 < p_1 < ...
// evaluated 8 bits at a time from LUT[chunk][#set bits below chunk][byte].
__device__ __forceinline__ int64_t rank_combinadic(const LookupDesc& L, uint64_t s) {
  if (L.n_bits < 64 && (s >> L.n_bits)) return -1;
  if (__popcll(s) != L.n_set) return -1;
  uint64_t r = 0;
  int below = 0;
  const int stride = L.n_set + 1;
#pragma unroll 1
  for (int c = 0; c < L.n_chunks; ++c) {
    unsigned b = (unsigned)(s >> (8 * c)) & 255u;
    r += __ldg(L.comb_lut + ((size_t)(c * stride + below) * 256u + b));
    below += __popc(b);
  }
  return (int64_t)r;
}

__device__ __forceinline__ uint64_t unrank_combinadic(const LookupDesc& L, uint64_t r) {
  uint64_t s = 0;
  int p = L.n_bits;  // search positions below p
  for (int j = L.n_set; j >= 1; --j) {
    // largest p' < p with C(p', j) <= r
    int q = p - 1;
    while (__ldg(L.binom + q * 65 + j) > r) --q;
    s |= 1ull << q;
    r -= __ldg(L.binom + q * 65 + j);
    p = q;
  }
  return s;
}

// Generic HilbertSpaceSector: walk sites from the most significant one, summing the number of
// basis words that share the higher fields and have a smaller field here.
__device__ __forceinline__ int64_t rank_dprank(const LookupDesc& L, uint64_t s) {
  if (L.n_bits < 64 && (s >> L.n_bits)) return -1;
  uint64_t r = 0;
  int q = L.root_q;
  for (int i = L.n_sites - 1; i >= 0; --i) {
    unsigned w = L.site_w[i];
    unsigned f = (unsigned)((s >> L.site_off[i]) & ((1ull << w) - 1ull));
    if (f >= L.site_ns[i]) return -1;
    size_t base = (size_t)i * L.max_q + q;
    r += __ldg(L.dp_prefix + base * (L.max_states + 1) + f);
    q = __ldg(L.dp_next + base * L.max_states + f);
    if (q < 0) return -1;
  }
  return L.dp_accept[q] ? (int64_t)r : -1;
}

__device__ __forceinline__ uint64_t unrank_dprank(const LookupDesc& L, uint64_t r) {
  uint64_t s = 0;
  int q = L.root_q;
  for (int i = L.n_sites - 1; i >= 0; --i) {
    size_t base = (size_t)i * L.max_q + q;
    const uint64_t* pre = L.dp_prefix + base * (L.max_states + 1);
    int ns = L.site_ns[i];
    int v = 0;
    while (v + 1 < ns && __ldg(pre + v + 1) <= r) ++v;  // pre[v] <= r < pre[v+1]; empty children are skipped
    r -= __ldg(pre + v);
    s |= (uint64_t)v << L.site_off[i];
    q = __ldg(L.dp_next + base * L.max_states + v);
  }
  return s;
}

template <int KIND>
__device__ __forceinline__ int64_t rank_word(const LookupDesc& L, uint64_t s) {
  if (KIND == ED_BASIS_LIST) return rank_list(L.words, L.dim, s);
  if (KIND == ED_BASIS_FULL) return (s < (uint64_t)L.dim) ? (int64_t)s : -1;
  if (KIND == ED_BASIS_COMBINADIC) return rank_combinadic(L, s);
  return rank_dprank(L, s);
}

__device__ __forceinline__ int64_t rank_word_dyn(const LookupDesc& L, uint64_t s) {
  switch (L.kind) {
    case ED_BASIS_LIST: return rank_list(L.words, L.dim, s);
    case ED_BASIS_FULL: return (s < (uint64_t)L.dim) ? (int64_t)s : -1;
    case ED_BASIS_COMBINADIC: return rank_combinadic(L, s);
    default: return rank_dprank(L, s);
  }
}

// membership only ("get(basis_lookup, word, -1) > 0"): no rank needed where the index is not used
__device__ __forceinline__ bool in_basis_dyn(const LookupDesc& L, uint64_t s) {
  switch (L.kind) {
    case ED_BASIS_LIST: return rank_list(L.words, L.dim, s) >= 0;
    case ED_BASIS_FULL: return s < (uint64_t)L.dim;
    case ED_BASIS_COMBINADIC: return !(L.n_bits < 64 && (s >> L.n_bits)) && __popcll(s) == L.n_set;
    default: return rank_dprank(L, s) >= 0;
  }
}

// ------------------------------------------------------------------ group action
// symmetry_apply (src/Symmetry/symmetry_apply.jl:82-92, bitflipsymmetry.jl:23-35) as byte-chunk
// LUTs: image = OR_c lut[g][c][byte_c]; a GlobalBitFlip is folded into the tables.
__device__ __forceinline__ uint64_t sym_apply(const SymDesc& S, int g, uint64_t s) {
  const uint64_t* t = S.lut + (size_t)g * S.n_chunks * 256;
  uint64_t out = 0;
#pragma unroll 1
  for (int c = 0; c < S.n_chunks; ++c) {
    out |= __ldg(t + c * 256 + ((unsigned)(s >> (8 * c)) & 255u));
  }
  return out;
}

// index of `key` in the ascending representative list, bucketed by the top bits.
__device__ __forceinline__ int64_t rank_reduced(const RLookupDesc& R, uint64_t key) {
  if (R.hash) {
    const uint64_t slot_mask = ~0ull >> R.hash_shift;
    uint64_t s = (key * 0x9E3779B97F4A7C15ull) >> R.hash_shift;
    for (;;) {
      const unsigned long long e = __ldg(R.hash + s);
      if (e == ~0ull) return -1;                                        // load factor <= 1/2: an empty slot ends every probe
      if ((e >> R.idx_bits) == key) return (int64_t)(e & ((1ull << R.idx_bits) - 1));
      s = (s + 1) & slot_mask;
    }
  }
  uint64_t b = key >> R.bucket_shift;
  if (b >= (uint64_t)R.n_buckets) return -1;
  int64_t lo = __ldg(R.bucket_start + b), hi = __ldg(R.bucket_start + b + 1);
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(R.words + mid) < key) lo = mid + 1; else hi = mid;
  }
  return (lo < R.dim && __ldg(R.words + lo) == key) ? lo : -1;
}

// (basis_mapping_index, basis_mapping_amplitude) of a parent word, on the fly
// (src/Symmetry/symmetry_reduce_generic.jl:51-101):  word = g_i(r)  =>  index(r), conj(chi_i)/sqrt(N_r),
// with the LAST such i (the Dict overwrite at :74-78).  Requires the elements to form a group:
// r = min_j g_j(word) and g_j(word) = r  <=>  word = g_inv(j)(r).
// Returns 0-based reduced index or -1; *amp is only written on success.
__device__ __forceinline__ int64_t reduced_map_word(const SymDesc& S, const RLookupDesc& R, uint64_t word,
                                                    c128* amp) {
  uint64_t best = word;
  int best_i = __ldg(S.inverse + 0);  // identity
#pragma unroll 1
  for (int j = 1; j < S.n_ops; ++j) {
    uint64_t im = sym_apply(S, j, word);
    int inv = __ldg(S.inverse + j);
    if (im < best) { best = im; best_i = inv; }
    else if (im == best && inv > best_i) best_i = inv;
  }
  int64_t idx = rank_reduced(R, best);
  if (idx < 0) return -1;
  double inv_norm = 1.0 / sqrt((double)__ldg(R.orbit_size + idx));  // inv(sqrt(float(n))), :41
  double cr = __ldg(S.chi + 2 * best_i), ci = __ldg(S.chi + 2 * best_i + 1);
  *amp = make_c128(cr * inv_norm, -ci * inv_norm);
  return idx;
}

// amplitude stored for the representative itself: conj(chi_last_stab)/sqrt(N_r)
__device__ __forceinline__ c128 reduced_rep_amp(const SymDesc& S, const RLookupDesc& R, int64_t idx) {
  double inv_norm = 1.0 / sqrt((double)__ldg(R.orbit_size + idx));
  int i = __ldg(R.last_stab + idx);
  return make_c128(__ldg(S.chi + 2 * i) * inv_norm, -__ldg(S.chi + 2 * i + 1) * inv_norm);
}

// ------------------------------------------------------------------ block reduction of (re, im) pairs
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
