// K2 fast path: matrix-free  out (+)= H * x  for spin-1/2 (1-bit) sites in one U(1) sector.
//
// Replaces the same reference path as apply.cu (apply_serial!/apply_parallel!,
// Representation/abstract_operator_representation.jl:296-409 + operator_representation.jl:66-103 +
// operator_iterator.jl:48-63 + frozensortedarray.jl:29-48) for operators that lower to
//     H = const + sum_i h_i n_i + sum_{bonds (p,q)} [ J_pq (|01><10| + |10><01|)_pq + K_pq n_p n_q ]
// i.e. any XXZ / Heisenberg / J1-J2 / field Hamiltonian on any lattice (n_i = bit i of the basis word).
// Everything else takes the generic kernel.
//
// Design (B200: 148 SMs, 227 KB smem/CTA, 126 MB L2, HBM-bound path):
//   * the ascending Sz-sector basis is the combinatorial number system, so the rows that share their high
//     (n_bits - k) bits H form a CONTIGUOUS tile of C(k, p_low) rows whose low k bits enumerate all k-bit words
//     of popcount p_low in ascending order.  One CTA owns one tile:
//       - x of the tile is staged once in shared memory (coalesced);
//       - bonds inside the low k bits gather from shared memory, the column found by two byte-LUT lookups
//         (no binary search, no basis array);
//       - bonds inside the high bits are evaluated ONCE per tile (they depend on H only): each firing bond is a
//         shifted, fully coalesced stream x[base(H') + i];
//       - the one or two bonds straddling bit k gather from a neighbouring tile;
//       - the diagonal is a handful of popcounts (no per-term walk).
//   * basis words are never read from HBM: per row the traffic is x (8 B) + y (8 B) + a 2-byte low-word table
//     that lives in L2.
//   * row-owner writes: deterministic, no atomics; the Lanczos <x, Hx> partial is fused in the epilogue.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "ed_device.cuh"

void ed_reduce_pairs(const double* partials, int n, double* out2);  // apply.cu

#define U1_MAX_CLASSES 12
#define U1_MAX_HH 192
#define U1_MAX_MX 64

struct U1Params {
  int n_bits, n_set, k;
  int t1_stride;                // 2^(max(k-8,0))
  uint32_t tile_cap;            // largest tile (rows)
  const uint32_t* tile_H;       // [n_tiles] non-empty tiles in ascending H
  const uint64_t* tile_base;    // [2^(n_bits-k)] rank of the first row of tile H
  const uint16_t* lowword;      // [2^k] k-bit words sorted by (popcount, value)
  const uint32_t* lowofs;       // [k+2]
  const uint16_t* T0;           // [256]
  const uint16_t* T1;           // [9 * t1_stride]
  double dconst;
  int n_lin;  uint64_t lin_mask[U1_MAX_CLASSES];  double lin_coef[U1_MAX_CLASSES];
  int n_quad; int quad_d[U1_MAX_CLASSES]; uint64_t quad_mask[U1_MAX_CLASSES]; double quad_coef[U1_MAX_CLASSES];
  int n_ll;   int ll_d[U1_MAX_CLASSES];   uint32_t ll_mask[U1_MAX_CLASSES];   double ll_amp[U1_MAX_CLASSES];
  int n_hh;   const uint8_t* hh_p; const uint8_t* hh_q; const double* hh_amp;   // positions inside H
  int n_mx;   const uint8_t* mx_p; const uint8_t* mx_q; const double* mx_amp;   // p: low bit, q: bit inside H
  int64_t row_lo, row_hi;
  int accumulate;
  int tile_first;               // first tile of the launch (row shards launch only the tiles they overlap)
};

struct FastU1Plan {
  bool supported = false;
  U1Params P;
  int n_tiles = 0;
  size_t smem_bytes = 0;
  int vec_bytes = 8;
  DevBuf<uint32_t> tile_H, lowofs;
  DevBuf<uint64_t> tile_base;
  DevBuf<uint16_t> lowword, T0, T1;
  DevBuf<uint8_t> hh_p, hh_q, mx_p, mx_q;
  DevBuf<double> hh_amp, mx_amp;
  DevBuf<double> partials;
  std::vector<uint64_t> h_base, h_size;  // per non-empty tile, ascending
};

template <typename T>
__device__ __forceinline__ T vec_scale(double a, T v);
template <>
__device__ __forceinline__ double vec_scale<double>(double a, double v) { return a * v; }
template <>
__device__ __forceinline__ c128 vec_scale<c128>(double a, c128 v) { return make_c128(a * v.re, a * v.im); }

__device__ __forceinline__ void vec_fma(double& acc, double a, double v) { acc = fma(a, v, acc); }
__device__ __forceinline__ void vec_fma(c128& acc, double a, c128 v) { acc.re = fma(a, v.re, acc.re); acc.im = fma(a, v.im, acc.im); }
__device__ __forceinline__ double vec_add(double a, double b) { return a + b; }
__device__ __forceinline__ c128 vec_add(c128 a, c128 b) { return cadd(a, b); }

template <typename WordT> __device__ __forceinline__ int popc_w(WordT v);
template <> __device__ __forceinline__ int popc_w<uint32_t>(uint32_t v) { return __popc(v); }
template <> __device__ __forceinline__ int popc_w<uint64_t>(uint64_t v) { return __popcll(v); }

template <typename VecT, typename WordT, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
k2_apply_u1(const U1Params P, const VecT* __restrict__ x, VecT* __restrict__ y, double* __restrict__ dot_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VecT* xs = reinterpret_cast<VecT*>(smem_raw);
  uint64_t* hh_base = reinterpret_cast<uint64_t*>(xs + P.tile_cap);
  double* hh_amp = reinterpret_cast<double*>(hh_base + U1_MAX_HH);
  uint64_t* mx_base = reinterpret_cast<uint64_t*>(hh_amp + U1_MAX_HH);
  double* mx_amp = reinterpret_cast<double*>(mx_base + U1_MAX_MX);
  uint32_t* mx_info = reinterpret_cast<uint32_t*>(mx_amp + U1_MAX_MX);   // low bit position | (H bit value << 8)
  uint16_t* sT0 = reinterpret_cast<uint16_t*>(mx_info + U1_MAX_MX);
  uint16_t* sT1 = sT0 + 256;
  __shared__ int s_counts[2];

  const int tid = threadIdx.x;
  const uint32_t H = P.tile_H[P.tile_first + blockIdx.x];
  const int p_low = P.n_set - __popc(H);
  const uint32_t lofs = P.lowofs[p_low];
  const uint32_t size = P.lowofs[p_low + 1] - lofs;
  const uint64_t base = P.tile_base[H];
  const int k = P.k;

  // ---- prologue: per-tile bond lists (deterministic ballot compaction), rank LUTs, x tile ---------------
  if (tid < 32) {
    int n = 0;
    for (int b0 = 0; b0 < P.n_hh; b0 += 32) {
      const int b = b0 + tid;
      bool fire = false;
      uint32_t H2 = 0;
      if (b < P.n_hh) {
        const int p = P.hh_p[b], q = P.hh_q[b];
        fire = (((H >> p) ^ (H >> q)) & 1u) != 0;
        H2 = H ^ ((1u << p) | (1u << q));
      }
      const unsigned m = __ballot_sync(0xffffffffu, fire);
      if (fire) {
        const int slot = n + __popc(m & ((1u << tid) - 1u));
        hh_base[slot] = P.tile_base[H2];
        hh_amp[slot] = P.hh_amp[b];
      }
      n += __popc(m);
    }
    if (tid == 0) s_counts[0] = n;
  } else if (tid < 64) {
    const int lane = tid - 32;
    for (int b = lane; b < P.n_mx; b += 32) {
      const int q = P.mx_q[b];
      const uint32_t hbit = (H >> q) & 1u;
      mx_base[b] = P.tile_base[H ^ (1u << q)];
      mx_amp[b] = P.mx_amp[b];
      mx_info[b] = (uint32_t)P.mx_p[b] | (hbit << 8);
    }
  }
  for (int i = tid; i < 256; i += THREADS) sT0[i] = P.T0[i];
  for (int i = tid; i < 9 * P.t1_stride; i += THREADS) sT1[i] = P.T1[i];
  for (uint32_t i = tid; i < size; i += THREADS) xs[i] = ldg_val(x + base + i);
  __syncthreads();
  const int n_hh = s_counts[0];
  const int n_mx = P.n_mx;
  const int t1s = P.t1_stride;

  double dre = 0.0, dim_ = 0.0;
  for (uint32_t i = tid; i < size; i += THREADS) {
    const int64_t row = (int64_t)(base + i);
    if (row < P.row_lo || row >= P.row_hi) continue;
    const uint32_t low = __ldg(P.lowword + lofs + i);
    const WordT s = ((WordT)H << k) | (WordT)low;
    // diagonal: const + sum coef * popcount(...)
    double d = P.dconst;
#pragma unroll 1
    for (int c = 0; c < P.n_lin; ++c) d = fma(P.lin_coef[c], (double)popc_w<WordT>(s & (WordT)P.lin_mask[c]), d);
#pragma unroll 1
    for (int c = 0; c < P.n_quad; ++c)
      d = fma(P.quad_coef[c], (double)popc_w<WordT>(s & (s >> P.quad_d[c]) & (WordT)P.quad_mask[c]), d);
    const VecT xi = xs[i];
    VecT acc = vec_scale<VecT>(d, xi);
    // exchange bonds inside the low k bits: shared-memory gathers
#pragma unroll 1
    for (int c = 0; c < P.n_ll; ++c) {
      const int dd = P.ll_d[c];
      uint32_t t = (low ^ (low >> dd)) & P.ll_mask[c];
      const double a = P.ll_amp[c];
      const uint32_t pair = 1u | (1u << dd);
      while (t) {
        const int q = __ffs(t) - 1;
        t &= t - 1;
        const uint32_t low2 = low ^ (pair << q);
        const uint32_t b0 = low2 & 255u;
        const uint32_t j = (uint32_t)sT0[b0] + (uint32_t)sT1[__popc(b0) * t1s + (low2 >> 8)];
        vec_fma(acc, a, xs[j]);
      }
    }
    // bonds straddling bit k: gather from the neighbouring tile
#pragma unroll 1
    for (int e = 0; e < n_mx; ++e) {
      const uint32_t info = mx_info[e];
      const int p = info & 255u;
      if (((low >> p) & 1u) != (info >> 8)) {
        const uint32_t low2 = low ^ (1u << p);
        const uint32_t b0 = low2 & 255u;
        const uint32_t j = (uint32_t)sT0[b0] + (uint32_t)sT1[__popc(b0) * t1s + (low2 >> 8)];
        vec_fma(acc, mx_amp[e], ldg_val(x + mx_base[e] + j));
      }
    }
    // bonds inside the high bits: same local index in another tile, coalesced
#pragma unroll 4
    for (int e = 0; e < n_hh; ++e) vec_fma(acc, hh_amp[e], ldg_val(x + hh_base[e] + i));
    VecT* dst = y + (row - P.row_lo);
    if (P.accumulate) acc = vec_add(acc, *dst);
    st_val(dst, acc);
    if (dot_partials) dot_acc(dre, dim_, xi, acc);
  }
  if (dot_partials) {
    __shared__ double s_red[2][THREADS / 32];
    dre = warp_sum(dre);
    dim_ = warp_sum(dim_);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = dre; s_red[1][tid >> 5] = dim_; }
    __syncthreads();
    if (tid == 0) {
      double a = 0, c = 0;
      for (int w = 0; w < THREADS / 32; ++w) { a += s_red[0][w]; c += s_red[1][w]; }
      dot_partials[2 * blockIdx.x] = a;
      dot_partials[2 * blockIdx.x + 1] = c;
    }
  }
}

// ------------------------------------------------------------------ host: lowering
namespace {

struct Bond {
  double e01 = 0, e10 = 0;          // exchange amplitudes (row pattern b_p=1,b_q=0 -> 0,1) and reverse
  bool has_e01 = false, has_e10 = false;
  double a[4] = {0, 0, 0, 0};       // diagonal, index = b_p | (b_q << 1)
};

struct Lowered {
  bool ok = false;
  double dconst = 0;
  std::vector<std::pair<double, uint64_t>> lin;                  // coef, mask
  struct Cls { int d; double v; uint64_t mask; };
  std::vector<Cls> quad, exch;
};

Lowered lower_operator(const ed_operator& op, int n_bits, int n_set) {
  Lowered L;
  if (op.is_complex) return L;
  std::map<std::pair<int, int>, Bond> bonds;
  std::vector<double> lin(n_bits, 0.0);
  double dconst = 0;
  for (int64_t t = 0; t < op.n_terms; ++t) {
    const uint64_t m = op.mask[t], r = op.row[t], c = op.col[t];
    const double a = op.amp[t];
    if (n_bits < 64 && (m >> n_bits)) return L;
    const int pc = __builtin_popcountll(m);
    if (pc == 0) { dconst += a; continue; }
    if (pc == 1) {
      if (r != c) return L;                      // single-site flips leave the sector
      const int p = __builtin_ctzll(m);
      if (r) lin[p] += a; else { dconst += a; lin[p] -= a; }
      continue;
    }
    if (pc != 2) return L;
    const int p = __builtin_ctzll(m), q = 63 - __builtin_clzll(m);
    Bond& b = bonds[{p, q}];
    if (r == c) {
      const int idx = (int)((r >> p) & 1) | ((int)((r >> q) & 1) << 1);
      b.a[idx] += a;
    } else {
      if ((r ^ c) != m || __builtin_popcountll(r) != 1) return L;   // must move one particle along the bond
      if ((r >> p) & 1) { b.e01 += a; b.has_e01 = true; } else { b.e10 += a; b.has_e10 = true; }
    }
  }
  std::map<std::pair<int, double>, uint64_t> quad, exch;
  for (auto& kv : bonds) {
    const int p = kv.first.first, q = kv.first.second;
    const Bond& b = kv.second;
    if (b.has_e01 != b.has_e10 || b.e01 != b.e10) return L;       // symmetric exchange only
    dconst += b.a[0];
    lin[p] += b.a[1] - b.a[0];
    lin[q] += b.a[2] - b.a[0];
    const double c3 = b.a[3] - b.a[1] - b.a[2] + b.a[0];
    if (c3 != 0.0) quad[{q - p, c3}] |= 1ull << p;
    if (b.has_e01 && b.e01 != 0.0) exch[{q - p, b.e01}] |= 1ull << p;
  }
  // linear part: group equal coefficients; a coefficient on every site is a constant in a fixed-popcount sector
  std::map<double, uint64_t> lin_groups;
  for (int p = 0; p < n_bits; ++p)
    if (lin[p] != 0.0) lin_groups[lin[p]] |= 1ull << p;
  const uint64_t all = n_bits >= 64 ? ~0ull : ((1ull << n_bits) - 1ull);
  for (auto& kv : lin_groups) {
    if (kv.second == all) dconst += kv.first * n_set;
    else L.lin.push_back({kv.first, kv.second});
  }
  for (auto& kv : quad) L.quad.push_back({kv.first.first, kv.first.second, kv.second});
  for (auto& kv : exch) L.exch.push_back({kv.first.first, kv.first.second, kv.second});
  L.dconst = dconst;
  L.ok = true;
  return L;
}

uint64_t binom_u64(int n, int k) {
  if (k < 0 || k > n) return 0;
  unsigned __int128 r = 1;
  for (int i = 1; i <= k; ++i) r = r * (unsigned)(n - k + i) / (unsigned)i;
  return (uint64_t)r;
}

int choose_k(int n_bits, int vec_bytes) {
  int k = std::min(16, std::max(4, n_bits - 11));
  k = std::min(k, n_bits);
  if (const char* e = getenv("EDCUDA_U1_K")) {
    int v = atoi(e);
    if (v >= 1 && v <= 16) k = std::min(v, n_bits);
  }
  // x tile must leave room for two CTAs per SM: <= 104 KB
  while (k > 1 && binom_u64(k, k / 2) * (uint64_t)vec_bytes > 104 * 1024) --k;
  return k;
}

}  // namespace

static std::shared_ptr<FastU1Plan> build_plan(ed_oprep* o, int vec_bytes) {
  auto plan = std::make_shared<FastU1Plan>();
  ed_basis* b = o->basis;
  plan->vec_bytes = vec_bytes;
  if (b->kind != ED_BASIS_COMBINADIC || b->dim <= 0) return plan;
  const int n_bits = b->space.bits, n_set = b->n_set;
  if (n_bits < 1 || n_bits > 48) return plan;
  Lowered L = lower_operator(o->op, n_bits, n_set);
  if (!L.ok) return plan;
  if ((int)L.lin.size() > U1_MAX_CLASSES || (int)L.quad.size() > U1_MAX_CLASSES) return plan;
  const int k = choose_k(n_bits, vec_bytes);
  const int hb = n_bits - k;
  if (hb > 26) return plan;  // tile tables of 2^hb entries
  U1Params& P = plan->P;
  memset(&P, 0, sizeof(P));
  P.n_bits = n_bits; P.n_set = n_set; P.k = k;
  P.t1_stride = 1 << std::max(k - 8, 0);
  P.dconst = L.dconst;
  P.n_lin = (int)L.lin.size();
  for (int c = 0; c < P.n_lin; ++c) { P.lin_coef[c] = L.lin[c].first; P.lin_mask[c] = L.lin[c].second; }
  P.n_quad = (int)L.quad.size();
  for (int c = 0; c < P.n_quad; ++c) { P.quad_d[c] = L.quad[c].d; P.quad_coef[c] = L.quad[c].v; P.quad_mask[c] = L.quad[c].mask; }
  // split the exchange bonds at bit k
  std::vector<uint8_t> hh_p, hh_q, mx_p, mx_q;
  std::vector<double> hh_amp, mx_amp;
  const uint64_t lowmask = (1ull << k) - 1ull;
  for (auto& c : L.exch) {
    uint64_t ll = 0;
    for (int p = 0; p < n_bits; ++p) {
      if (!(c.mask >> p & 1)) continue;
      const int q = p + c.d;
      if (q < k) ll |= 1ull << p;
      else if (p >= k) { hh_p.push_back((uint8_t)(p - k)); hh_q.push_back((uint8_t)(q - k)); hh_amp.push_back(c.v); }
      else { mx_p.push_back((uint8_t)p); mx_q.push_back((uint8_t)(q - k)); mx_amp.push_back(c.v); }
    }
    if (ll) {
      if (P.n_ll >= U1_MAX_CLASSES) return plan;
      P.ll_d[P.n_ll] = c.d; P.ll_mask[P.n_ll] = (uint32_t)(ll & lowmask); P.ll_amp[P.n_ll] = c.v;
      ++P.n_ll;
    }
  }
  if ((int)hh_p.size() > U1_MAX_HH || (int)mx_p.size() > U1_MAX_MX) return plan;
  P.n_hh = (int)hh_p.size();
  P.n_mx = (int)mx_p.size();
  // low-word tables
  std::vector<uint32_t> lowofs(k + 2, 0);
  for (int p = 0; p <= k; ++p) lowofs[p + 1] = lowofs[p] + (uint32_t)binom_u64(k, p);
  std::vector<uint16_t> lowword((size_t)1 << k);
  {
    std::vector<uint32_t> at(lowofs.begin(), lowofs.end());
    for (uint32_t w = 0; w < (1u << k); ++w) lowword[at[__builtin_popcount(w)]++] = (uint16_t)w;
  }
  std::vector<uint16_t> T0(256, 0), T1((size_t)9 * P.t1_stride, 0);
  for (int byte = 0; byte < 256; ++byte) {
    uint64_t acc = 0; int i = 0;
    for (int q = 0; q < 8; ++q) if (byte >> q & 1) { acc += binom_u64(q, i + 1); ++i; }
    T0[byte] = (uint16_t)acc;
  }
  for (int pc0 = 0; pc0 <= 8; ++pc0)
    for (int hi = 0; hi < P.t1_stride; ++hi) {
      uint64_t acc = 0; int i = 0;
      for (int q = 0; q < 8; ++q) if (hi >> q & 1) { acc += binom_u64(8 + q, pc0 + i + 1); ++i; }
      T1[(size_t)pc0 * P.t1_stride + hi] = (uint16_t)acc;
    }
  // tiles
  const uint32_t nH = 1u << hb;
  std::vector<uint64_t> tile_base(nH, 0);
  std::vector<uint32_t> tile_H;
  uint32_t tile_cap = 1;
  {
    // rank of (H << k | lowest word with p_low bits) = sum over set bits of H of C(k + pos, p_low + idx + 1)
    for (uint32_t H = 0; H < nH; ++H) {
      const int p_low = n_set - __builtin_popcount(H);
      if (p_low < 0 || p_low > k) continue;
      uint64_t acc = 0; int i = 0;
      for (int q = 0; q < hb; ++q) if (H >> q & 1) { acc += binom_u64(k + q, p_low + i + 1); ++i; }
      tile_base[H] = acc;
      tile_H.push_back(H);
      plan->h_base.push_back(acc);
      plan->h_size.push_back(binom_u64(k, p_low));
      tile_cap = std::max<uint32_t>(tile_cap, (uint32_t)binom_u64(k, p_low));
    }
  }
  P.tile_cap = tile_cap;
  plan->n_tiles = (int)tile_H.size();
  plan->tile_H.upload(tile_H); plan->tile_base.upload(tile_base);
  plan->lowword.upload(lowword); plan->lowofs.upload(lowofs);
  plan->T0.upload(T0); plan->T1.upload(T1);
  if (hh_p.empty()) { hh_p.push_back(0); hh_q.push_back(0); hh_amp.push_back(0); }
  if (mx_p.empty()) { mx_p.push_back(0); mx_q.push_back(0); mx_amp.push_back(0); }
  plan->hh_p.upload(hh_p); plan->hh_q.upload(hh_q); plan->hh_amp.upload(hh_amp);
  plan->mx_p.upload(mx_p); plan->mx_q.upload(mx_q); plan->mx_amp.upload(mx_amp);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  P.tile_H = plan->tile_H.p; P.tile_base = plan->tile_base.p; P.lowword = plan->lowword.p; P.lowofs = plan->lowofs.p;
  P.T0 = plan->T0.p; P.T1 = plan->T1.p;
  P.hh_p = plan->hh_p.p; P.hh_q = plan->hh_q.p; P.hh_amp = plan->hh_amp.p;
  P.mx_p = plan->mx_p.p; P.mx_q = plan->mx_q.p; P.mx_amp = plan->mx_amp.p;
  plan->smem_bytes = (size_t)tile_cap * vec_bytes + U1_MAX_HH * 16 + U1_MAX_MX * (16 + 4) + 256 * 2 + (size_t)9 * P.t1_stride * 2;
  plan->smem_bytes = (plan->smem_bytes + 15) & ~(size_t)15;
  plan->supported = true;
  return plan;
}

static FastU1Plan* get_plan(ed_oprep* o, int dtype) {
  // one plan per vector type (the tile size depends on the element size), cached on the representation
  std::shared_ptr<FastU1Plan>& slot = dtype == ED_C128 ? o->u1plan_c : o->u1plan;
  if (!slot) slot = build_plan(o, dtype == ED_C128 ? 16 : 8);
  return slot.get();
}

bool ed_apply_u1_supported(ed_oprep* o, int dtype, int side) {
  (void)side;  // symmetric real H: x*H == H*x
  if (o->rbasis) return false;
  if (o->op.is_complex) return false;
  if (o->basis->kind != ED_BASIS_COMBINADIC) return false;
  return get_plan(o, dtype)->supported;
}

template <typename VecT, typename WordT>
static void launch_u1(FastU1Plan* plan, const U1Params& P, int n_launch, const void* x, void* out, double* partials) {
  constexpr int THREADS = 512;
  auto kern = k2_apply_u1<VecT, WordT, THREADS>;
  static thread_local size_t configured = 0;
  if (plan->smem_bytes > 48 * 1024 && configured < plan->smem_bytes) {
    ED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes));
    configured = plan->smem_bytes;
  }
  ED_LAUNCH(kern, n_launch, THREADS, plan->smem_bytes, P, reinterpret_cast<const VecT*>(x), reinterpret_cast<VecT*>(out), partials);
}

void ed_apply_u1(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot) {
  (void)side;
  FastU1Plan* plan = get_plan(o, dtype);
  ED_REQUIRE(plan->supported, ED_ERR_INTERNAL, "u1 fast path requested for an unsupported representation");
  U1Params P = plan->P;
  P.row_lo = o->row_lo;
  P.row_hi = o->row_hi;
  P.accumulate = accumulate;
  // tiles overlapping the owned rows [row_lo, row_hi)
  int first = (int)(std::upper_bound(plan->h_base.begin(), plan->h_base.end(), (uint64_t)std::max<int64_t>(o->row_lo, 0)) - plan->h_base.begin()) - 1;
  first = std::max(first, 0);
  int last = (int)(std::lower_bound(plan->h_base.begin(), plan->h_base.end(), (uint64_t)o->row_hi) - plan->h_base.begin());
  const int n_launch = last - first;
  P.tile_first = first;
  if (n_launch <= 0 || o->row_hi <= o->row_lo) {
    if (alpha_dot) ED_CUDA(cudaMemsetAsync(alpha_dot, 0, 2 * sizeof(double), ed_stream()));
    return;
  }
  double* partials = nullptr;
  if (alpha_dot) {
    if (plan->partials.n < (size_t)2 * plan->n_tiles) plan->partials.alloc((size_t)2 * plan->n_tiles);
    partials = plan->partials.p;
  }
  const bool w32 = P.n_bits <= 32;
  if (dtype == ED_F64) {
    if (w32) launch_u1<double, uint32_t>(plan, P, n_launch, x, out, partials); else launch_u1<double, uint64_t>(plan, P, n_launch, x, out, partials);
  } else {
    if (w32) launch_u1<c128, uint32_t>(plan, P, n_launch, x, out, partials); else launch_u1<c128, uint64_t>(plan, P, n_launch, x, out, partials);
  }
  if (alpha_dot) ed_reduce_pairs(partials, n_launch, alpha_dot);
}
