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

#define U1_MAX_CLASSES 8
#define U1_MAX_HH 192
#define U1_MAX_MX 64
#define U1_MAX_MQ 64
#define U1_MAX_MS 16
#define ED_MAX_SEG 16

// Everything that depends only on the LOW k bits of a row is tabulated once per plan (tables live in L2):
//   dcode / dval   diagonal of the low-bit terms (u8 code -> value)
//   ell            for every exchange class, the local column indices of the firing low-bit bonds in ELL form
//                  (slot-major, padded with a dummy index that points at a zero in shared memory)
//   mx_tab         local column index inside the neighbouring tile for every bond straddling bit k
struct U1Params {
  int n_bits, n_set, k;
  uint32_t tile_cap;            // largest tile (rows)
  const uint32_t* tile_H;       // [n_tiles] non-empty tiles in ascending H
  const uint64_t* tile_base;    // [2^(n_bits-k)] rank of the first row of tile H
  const double* tile_diag;      // [2^(n_bits-k)] constant + diagonal terms that depend on H only
  const uint16_t* lowword;      // [2^k] k-bit words sorted by (popcount, value)
  const uint32_t* lowofs;       // [k+2] first row of popcount p in the tables
  const uint32_t* grpofs;       // [k+2] first 32-row group of popcount p
  int diag_mode;                // 0: no low-bit diagonal, 1: u8 code + dval, 2: f64 table
  const uint8_t* dcode;         // [2^k]
  const double* dval;           // [256]
  const double* dlow;           // [2^k]
  int n_mq; const uint8_t* mq_p; const uint8_t* mq_q; const double* mq_coef;   // n_p n_q terms straddling bit k
  // mq_folded: the u8 code also enumerates the low bits the straddling terms look at (dpat = those bits per code,
  // mq_pidx = which pattern bit a term reads), so the per-tile value table absorbs them and rows need no extra work
  int mq_folded; const uint8_t* dpat; const uint8_t* mq_pidx;
  int n_ll; double ll_amp[U1_MAX_CLASSES];
  const uint32_t* ell[U1_MAX_CLASSES];     // class tables: two consecutive slots per 32-bit word
  const uint8_t* ell_cnt[U1_MAX_CLASSES];  // slot PAIRS used per 32-row group
  const uint32_t* ell_ofs;                 // [n_ll * (k+1)] start of popcount p inside ell[c]
  // nearest-neighbour exchange bonds inside the low bits (q, q+1): the column of row i is i +- C(q, r_q) with r_q the
  // number of set bits below q, so no per-row index table is needed at all.  The signed offsets come from two tiny
  // shared-memory tables: lut_lo[low 8 bits] -> 7 x i8 (bonds 0..6), lut_up[class][word >> 7] -> 8 x i16 (bonds 7..14).
  int cap_hh, cap_mx, cap_mq, cap_ms, cap_codes;   // shared-memory list capacities of this plan (>= 1, <= U1_MAX_*)
  int lut_on;
  const uint64_t* lut_lo;       // [256]
  const uint4* lut_up;          // [(k + 1) << (k - 7)]
  double nn_amp[16];
  int n_hh;   const uint8_t* hh_p; const uint8_t* hh_q; const double* hh_amp;   // exchange bonds inside H
  int n_mx;   const uint8_t* mx_q; const double* mx_amp; const uint16_t* mx_tab; // straddling exchange bonds (gathered)
  // straddling bonds whose low site is bit k-1: the firing rows are a contiguous block of the tile and so are their
  // columns in the neighbour tile -> a shifted coalesced stream like the high-bit bonds
  int n_ms;   const uint8_t* ms_q; const double* ms_amp;
  uint32_t ck1[20];             // C(k-1, p)
  int64_t row_lo, row_hi;
  int accumulate;
  int tile_first;               // first tile of the launch (row shards launch only the tiles they overlap)
  const uint32_t* tile_order;   // optional launch order (whole-basis launches): position -> index into tile_H
  // group kernel (k2_apply_u1g): H = (hi << gm) | mid; a CTA owns every tile of one (hi, popcount(mid)) class
  int gm, gcap, n_codes;
  const uint32_t* grp;          // [n_groups] (hi << 3) | popcount(mid), ascending
  int grp_first;
  uint8_t gcount[8];            // tiles per group of mid popcount c
  uint8_t gmid[8][8];           // the mid patterns of popcount c, ascending
  // x as up to ED_MAX_SEG contiguous, tile-aligned segments (local memory or peer GPUs' memory mapped over NVLink)
  int n_seg;
  int64_t seg_lo[ED_MAX_SEG + 1];
  const void* seg_ptr[ED_MAX_SEG];
};

template <typename VecT>
__device__ __forceinline__ const VecT* u1_seg_resolve(const U1Params& P, uint64_t idx) {
  int s = 0;
  while (s + 1 < P.n_seg && (int64_t)idx >= P.seg_lo[s + 1]) ++s;
  return reinterpret_cast<const VecT*>(P.seg_ptr[s]) + ((int64_t)idx - P.seg_lo[s]);
}

struct FastU1Plan {
  bool supported = false;
  U1Params P;
  int n_tiles = 0;
  size_t smem_bytes = 0;
  int vec_bytes = 8;
  bool idx32 = true;
  DevBuf<uint32_t> tile_H, lowofs, grpofs, ell_ofs;
  DevBuf<uint64_t> tile_base;
  DevBuf<uint16_t> lowword, mx_tab;
  DevBuf<uint8_t> dcode, hh_p, hh_q, mx_q, mq_p, mq_q, ms_q, dpat, mq_pidx;
  DevBuf<uint32_t> tile_order;
  DevBuf<uint64_t> lut_lo;
  DevBuf<uint4> lut_up;
  DevBuf<double> tile_diag, dval, dlow, hh_amp, mx_amp, mq_coef, ms_amp;
  std::vector<DevBuf<uint32_t>> ell;
  std::vector<DevBuf<uint8_t>> ell_cnt;
  DevBuf<double> partials;
  std::vector<uint64_t> h_base, h_size;  // per non-empty tile, ascending
  bool order_on = false;
  // group kernel configuration (k2_apply_u1g)
  bool g_on = false;
  int g_threads = 512, g_rs = 2;
  size_t g_smem = 0;
  DevBuf<uint32_t> grp;
  std::vector<uint64_t> g_lo, g_hi;      // per group: first row of its first tile, end row of its last tile
  std::vector<uint64_t> blk_base;        // first row of every non-empty hi block (shard boundaries in group mode)
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
// y is written once and never re-read by this kernel: streaming (evict-first) stores keep L2 for x
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(c128* p, c128 v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.re, v.im)); }
__device__ __forceinline__ c128 vec_add(c128 a, c128 b) { return cadd(a, b); }

// Per-tile state handed from the prologue to the slab body.
template <typename VecT>
struct U1Tile {
  VecT* xs;
  const double* hh_amp; const VecT* const* hh_ptr; int n_hh;      // neighbour tiles as resolved pointers
  const double* mx_amp; const VecT* const* mx_ptr; const uint32_t* mx_toff; int n_mx;
  const double* ms_amp; const VecT* const* ms_ptr; const uint32_t* ms_lo; const uint32_t* ms_len; int n_ms;
  const double* mq_coef; const uint32_t* mq_bit; int n_mq;
  const double* s_dval;
  const uint64_t* s_lut_lo; const uint4* s_lut_up;
  uint32_t lofs, gofs, size;
  int p_low;
  int64_t base;
  double d_tile;
};

// Slab body.  Slab r holds rows i = tid + r*THREADS.  NF = number of COMPLETE slabs of this tile (compile time:
// constant offsets, no predicates, NF+1 independent loads in flight per thread and bond).  The last, partial slab is
// addressed through the clamped per-thread index `it`, so every load stays in bounds; only its store is predicated.
template <typename VecT, int THREADS, int NF>
__device__ __forceinline__ void u1_tile_body(const U1Params& P, const U1Tile<VecT>& T,
                                             VecT* __restrict__ y, bool want_dot, double& dre, double& dim_, int slab0) {
  constexpr int CH = sizeof(VecT) == 8 ? 6 : 3;   // loads issued back to back before their FMAs (register budget)
  const int tid = threadIdx.x + slab0 * THREADS;   // local row of slab 0 of this pass
  VecT* xs = T.xs;
  const uint32_t size = T.size;
  const uint32_t i_tail = tid + NF * THREADS;
  const bool tail_ok = i_tail < size;
  const uint32_t it = tail_ok ? i_tail : size - 1;
  VecT acc[NF > 0 ? NF : 1];
  VecT acc_t;
  // diagonal
  {
    const uint8_t* dc = P.dcode + T.lofs;
    const double* dl = P.dlow + T.lofs;
    double dd[NF > 0 ? NF : 1];
    double dt = T.d_tile;
    if (P.diag_mode == 1) {
      uint32_t code[NF > 0 ? NF : 1];
#pragma unroll
      for (int r = 0; r < NF; ++r) code[r] = __ldg(dc + tid + r * THREADS);
      const uint32_t ct = __ldg(dc + it);
#pragma unroll
      for (int r = 0; r < NF; ++r) dd[r] = T.d_tile + T.s_dval[code[r]];
      dt += T.s_dval[ct];
    } else if (P.diag_mode == 2) {
#pragma unroll
      for (int r = 0; r < NF; ++r) dd[r] = T.d_tile + __ldg(dl + tid + r * THREADS);
      dt += __ldg(dl + it);
    } else {
#pragma unroll
      for (int r = 0; r < NF; ++r) dd[r] = T.d_tile;
    }
    if (T.n_mq) {
      const uint16_t* lw = P.lowword + T.lofs;
#pragma unroll
      for (int r = 0; r <= NF; ++r) {
        const uint32_t low = __ldg(lw + (r < NF ? tid + r * THREADS : it));
        double d = 0.0;
        for (int e = 0; e < T.n_mq; ++e) d += ((low >> T.mq_bit[e]) & 1u) ? T.mq_coef[e] : 0.0;
        if (r < NF) dd[r] += d; else dt += d;
      }
    }
#pragma unroll
    for (int r = 0; r < NF; ++r) acc[r] = vec_scale<VecT>(dd[r], xs[tid + r * THREADS]);
    acc_t = vec_scale<VecT>(dt, xs[it]);
  }
  // bonds inside the high bits: same local index in another tile -> coalesced streams
#pragma unroll 1
  for (int e = 0; e < T.n_hh; ++e) {
    const double a = T.hh_amp[e];
    const VecT* xe = T.hh_ptr[e];
    const VecT* xt = xe + tid;
    const VecT vt = ldg_val(xe + it);
#pragma unroll
    for (int r0 = 0; r0 < NF; r0 += CH) {
      VecT v[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (r0 + c < NF) v[c] = ldg_val(xt + (r0 + c) * THREADS);
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (r0 + c < NF) vec_fma(acc[r0 + c], a, v[c]);
    }
    vec_fma(acc_t, a, vt);
  }
  // straddling bonds on bit k-1: a contiguous block of rows reads a shifted stream of the neighbour tile
#pragma unroll 1
  for (int e = 0; e < T.n_ms; ++e) {
    const double a = T.ms_amp[e];
    const VecT* xe = T.ms_ptr[e];            // already offset: column = row index
    const uint32_t lo = T.ms_lo[e], len = T.ms_len[e];
    const VecT* xt = xe + tid;
    const VecT vt = (it - lo < len) ? ldg_val(xe + it) : vzero((VecT*)nullptr);
#pragma unroll
    for (int r0 = 0; r0 < NF; r0 += CH) {
      VecT v[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (r0 + c < NF) v[c] = ((uint32_t)(tid + (r0 + c) * THREADS) - lo < len) ? ldg_val(xt + (r0 + c) * THREADS) : vzero((VecT*)nullptr);
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (r0 + c < NF) vec_fma(acc[r0 + c], a, v[c]);
    }
    vec_fma(acc_t, a, vt);
  }
  // nearest-neighbour bonds inside the low bits: signed column offsets from the two shared-memory LUTs
  if (P.lut_on) {
    const uint16_t* lw = P.lowword + T.lofs;
    uint32_t wv[NF > 0 ? NF : 1];
#pragma unroll
    for (int r = 0; r < NF; ++r) wv[r] = __ldg(lw + tid + r * THREADS);
    const uint32_t wt = __ldg(lw + it);
#pragma unroll
    for (int r = 0; r <= NF; ++r) {
      const uint32_t w = r < NF ? wv[r < NF ? r : 0] : wt;
      const VecT* xi = xs + (r < NF ? (uint32_t)(tid + r * THREADS) : it);
      const uint64_t dl = T.s_lut_lo[w & 0xFFu];
      const uint4 du = T.s_lut_up[w >> 7];
      const uint32_t dl0 = (uint32_t)dl, dl1 = (uint32_t)(dl >> 32);
      VecT a_ = vzero((VecT*)nullptr);
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int d = (int)(int8_t)((j < 4 ? dl0 : dl1) >> (8 * (j & 3)));
        if (d != 0) vec_fma(a_, P.nn_amp[j], xi[d]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t wd = j < 2 ? du.x : j < 4 ? du.y : j < 6 ? du.z : du.w;
        const int d = (int)(int16_t)(wd >> (16 * (j & 1)));
        if (d != 0) vec_fma(a_, P.nn_amp[7 + j], xi[d]);
      }
      if (r < NF) acc[r < NF ? r : 0] = vec_add(acc[r < NF ? r : 0], a_); else acc_t = vec_add(acc_t, a_);
    }
  }
  // exchange bonds inside the low k bits: ELL table of local columns (top bond first), shared-memory gathers
#pragma unroll 1
  for (int c = 0; c < P.n_ll; ++c) {
    const uint8_t* cnt = P.ell_cnt[c] + T.gofs;
    int nmax = (int)__ldg(cnt + (it >> 5));
#pragma unroll
    for (int r = 0; r < NF; ++r) nmax = max(nmax, (int)__ldg(cnt + ((tid + r * THREADS) >> 5)));   // warp-uniform
    const uint32_t* e = P.ell[c] + P.ell_ofs[c * (P.k + 1) + T.p_low];
    const double a = P.ll_amp[c];
#pragma unroll 1
    for (int sl = 0; sl < nmax; ++sl) {     // one iteration = a PAIR of slots packed in one 32-bit word
      const uint32_t* et = e + tid;
      const uint32_t jt = __ldg(e + it);
#pragma unroll
      for (int r0 = 0; r0 < NF; r0 += CH) {
        uint32_t j[CH];
#pragma unroll
        for (int cc = 0; cc < CH; ++cc)
          if (r0 + cc < NF) j[cc] = __ldg(et + (r0 + cc) * THREADS);
#pragma unroll
        for (int cc = 0; cc < CH; ++cc)
          if (r0 + cc < NF) { vec_fma(acc[r0 + cc], a, xs[j[cc] & 0xFFFFu]); vec_fma(acc[r0 + cc], a, xs[j[cc] >> 16]); }
      }
      vec_fma(acc_t, a, xs[jt & 0xFFFFu]);
      vec_fma(acc_t, a, xs[jt >> 16]);
      e += size;
    }
  }
  // bonds straddling bit k: tabulated local column inside the neighbouring tile (0xFFFF = does not fire)
#pragma unroll 1
  for (int e = 0; e < T.n_mx; ++e) {
    const uint16_t* tab = P.mx_tab + T.mx_toff[e];
    const uint16_t* tt = tab + tid;
    const VecT* xe = T.mx_ptr[e];
    const double a = T.mx_amp[e];
    const uint32_t jt = __ldg(tab + it);
    const VecT vt = jt != 0xFFFFu ? ldg_val(xe + jt) : vzero((VecT*)nullptr);
#pragma unroll
    for (int r0 = 0; r0 < NF; r0 += CH) {
      uint32_t j[CH];
      VecT v[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (r0 + c < NF) j[c] = __ldg(tt + (r0 + c) * THREADS);
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (r0 + c < NF) v[c] = j[c] != 0xFFFFu ? ldg_val(xe + j[c]) : vzero((VecT*)nullptr);
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (r0 + c < NF) vec_fma(acc[r0 + c], a, v[c]);
    }
    vec_fma(acc_t, a, vt);
  }
  // store (row-owner writes)
  const int64_t row0 = (int64_t)T.base + tid;
  VecT* yt = y + (row0 - P.row_lo);
  const bool whole = (int64_t)T.base >= P.row_lo && (int64_t)T.base + size <= P.row_hi;   // tile fully owned (uniform)
#pragma unroll
  for (int r = 0; r < NF; ++r) {
    const int64_t row = row0 + r * THREADS;
    if (!whole && (row < P.row_lo || row >= P.row_hi)) continue;
    VecT out = acc[r];
    if (P.accumulate) out = vec_add(out, yt[r * THREADS]);
    st_stream(yt + r * THREADS, out);
    if (want_dot) dot_acc(dre, dim_, xs[tid + r * THREADS], out);
  }
  const int64_t row = (int64_t)T.base + i_tail;
  if (tail_ok && (whole || (row >= P.row_lo && row < P.row_hi))) {
    VecT* dst = y + (row - P.row_lo);
    VecT out = acc_t;
    if (P.accumulate) out = vec_add(out, *dst);
    st_stream(dst, out);
    if (want_dot) dot_acc(dre, dim_, xs[i_tail], out);
  }
}

template <typename VecT, int THREADS, int R, int NF>
struct U1Dispatch {
  static __device__ __forceinline__ void run(int nfull, const U1Params& P, const U1Tile<VecT>& T, VecT* y,
                                             bool want_dot, double& dre, double& dim_, int slab0) {
    if (nfull == NF) u1_tile_body<VecT, THREADS, NF>(P, T, y, want_dot, dre, dim_, slab0);
    else U1Dispatch<VecT, THREADS, R, NF - 1>::run(nfull, P, T, y, want_dot, dre, dim_, slab0);
  }
};
template <typename VecT, int THREADS, int R>
struct U1Dispatch<VecT, THREADS, R, -1> {
  static __device__ __forceinline__ void run(int, const U1Params&, const U1Tile<VecT>&, VecT*, bool, double&, double&, int) {}
};

// One CTA = one tile of C(k, p_low) contiguous rows; R = ceil(tile_cap / THREADS) bounds the slabs per thread.
template <typename VecT, int THREADS, int R>
__global__ void __launch_bounds__(THREADS, sizeof(VecT) != 8 ? 2 : R <= 5 ? 4 : R <= 7 ? 3 : 2)
k2_apply_u1(const U1Params P, VecT* __restrict__ y, double* __restrict__ dot_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VecT* xs = reinterpret_cast<VecT*>(smem_raw);                       // tile_cap + 1 (last = 0: ELL padding target)
  double* hh_amp = reinterpret_cast<double*>(xs + P.tile_cap + 1);
  double* mx_amp = hh_amp + P.cap_hh;
  double* mq_coef = mx_amp + P.cap_mx;
  double* s_dval = mq_coef + P.cap_mq;
  const VecT** hh_ptr = reinterpret_cast<const VecT**>(s_dval + P.cap_codes);
  const VecT** mx_ptr = hh_ptr + P.cap_hh;
  double* ms_amp = reinterpret_cast<double*>(mx_ptr + P.cap_mx);
  const VecT** ms_ptr = reinterpret_cast<const VecT**>(ms_amp + P.cap_ms);
  uint32_t* mx_toff = reinterpret_cast<uint32_t*>(ms_ptr + P.cap_ms);
  uint32_t* mq_bit = mx_toff + P.cap_mx;
  uint32_t* ms_lo = mq_bit + P.cap_mq;
  uint32_t* ms_len = ms_lo + P.cap_ms;
  uint4* s_lut_up = reinterpret_cast<uint4*>(smem_raw + ((reinterpret_cast<unsigned char*>(ms_len + P.cap_ms) - smem_raw + 15) & ~(size_t)15));
  uint64_t* s_lut_lo = reinterpret_cast<uint64_t*>(s_lut_up + (P.lut_on ? (1u << (P.k - 7)) : 0u));
  __shared__ int s_counts[2];

  const int tid = threadIdx.x;
  const uint32_t H = P.tile_H[P.tile_order ? P.tile_order[blockIdx.x] : P.tile_first + blockIdx.x];
  const int p_low = P.n_set - __popc(H);
  const uint32_t lofs = P.lowofs[p_low];
  const uint32_t size = P.lowofs[p_low + 1] - lofs;
  const uint64_t base = P.tile_base[H];
  const int k = P.k;

  // ---- prologue: per-tile bond lists (deterministic ballot compaction), x tile --------------------------
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
        hh_ptr[slot] = u1_seg_resolve<VecT>(P, P.tile_base[H2]);
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
      mx_ptr[b] = u1_seg_resolve<VecT>(P, P.tile_base[H ^ (1u << q)]);
      mx_amp[b] = P.mx_amp[b];
      mx_toff[b] = ((uint32_t)(2 * b + hbit) << k) + lofs;
    }
    for (int b = lane; b < P.n_ms; b += 32) {
      const int q = P.ms_q[b];
      const uint32_t hbit = (H >> q) & 1u;
      const VecT* xn = u1_seg_resolve<VecT>(P, P.tile_base[H ^ (1u << q)]);
      const uint32_t n0 = P.ck1[p_low];                  // rows of this tile whose bit k-1 is clear (they come first)
      if (hbit == 0) {                                   // rows with bit k-1 set -> first rows of the neighbour (p_low - 1)
        ms_lo[b] = n0; ms_len[b] = size - n0; ms_ptr[b] = xn - n0;
      } else {                                           // rows with bit k-1 clear -> last rows of the neighbour (p_low + 1)
        ms_lo[b] = 0; ms_len[b] = n0; ms_ptr[b] = xn + P.ck1[p_low + 1];
      }
      ms_amp[b] = P.ms_amp[b];
    }
  } else if (tid < 96) {
    const int lane = tid - 64;
    int n = 0;
    for (int b0 = 0; b0 < P.n_mq && !P.mq_folded; b0 += 32) {
      const int b = b0 + lane;
      const bool on = b < P.n_mq && ((H >> P.mq_q[b]) & 1u);
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (on) {
        const int slot = n + __popc(m & ((1u << lane) - 1u));
        mq_bit[slot] = P.mq_p[b];
        mq_coef[slot] = P.mq_coef[b];
      }
      n += __popc(m);
    }
    if (lane == 0) s_counts[1] = n;
  }
  if (P.diag_mode == 1) {
    for (int i = tid; i < P.cap_codes; i += THREADS) {
      double v = P.dval[i];
      if (P.mq_folded) {
        const uint32_t pat = P.dpat[i];
        for (int e = 0; e < P.n_mq; ++e)
          if (((H >> P.mq_q[e]) & 1u) && ((pat >> P.mq_pidx[e]) & 1u)) v += P.mq_coef[e];
      }
      s_dval[i] = v;
    }
  }
  {
    const VecT* xo = u1_seg_resolve<VecT>(P, base);     // segments are tile aligned: the whole tile is in one segment
    for (uint32_t i = tid; i < size; i += THREADS) xs[i] = ldg_val(xo + i);
  }
  if (tid == 0) xs[size] = vzero((VecT*)nullptr);
  if (P.lut_on) {
    const uint32_t nu = 1u << (k - 7);
    const uint4* src = P.lut_up + (size_t)p_low * nu;
    for (uint32_t i = tid; i < nu; i += THREADS) s_lut_up[i] = __ldg(src + i);
    for (uint32_t i = tid; i < 256; i += THREADS) s_lut_lo[i] = __ldg(P.lut_lo + i);
  }
  __syncthreads();

  U1Tile<VecT> T;
  T.xs = xs;
  T.s_lut_lo = s_lut_lo; T.s_lut_up = s_lut_up;
  T.hh_amp = hh_amp; T.hh_ptr = hh_ptr; T.n_hh = s_counts[0];
  T.mx_amp = mx_amp; T.mx_ptr = mx_ptr; T.mx_toff = mx_toff; T.n_mx = P.n_mx;
  T.ms_amp = ms_amp; T.ms_ptr = ms_ptr; T.ms_lo = ms_lo; T.ms_len = ms_len; T.n_ms = P.n_ms;
  T.mq_coef = mq_coef; T.mq_bit = mq_bit; T.n_mq = s_counts[1];
  T.s_dval = s_dval;
  T.lofs = lofs; T.gofs = P.grpofs[p_low]; T.size = size; T.p_low = p_low; T.base = (int64_t)base;
  T.d_tile = P.tile_diag[H];
  double dre = 0.0, dim_ = 0.0;
  // passes of at most R slabs (R accumulators per thread stay in registers); all slabs but the very last are complete
  const int n_slab = (int)((size + THREADS - 1) / THREADS);
#pragma unroll 1
  for (int s0 = 0; s0 < n_slab; s0 += R) {
    const int nfull = min(R, n_slab - s0) - 1;
    U1Dispatch<VecT, THREADS, R, R - 1>::run(nfull, P, T, y, dot_partials != nullptr, dre, dim_, s0);
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

// ------------------------------------------------------------------------------------------------------------------
// Group kernel.  The plain kernel above is bound by L2 -> SM bandwidth (126 B per row: every x element is fetched ~9
// times as a neighbour stream and every CTA re-reads the 2-byte ELL column indices of its popcount class).  Here a CTA
// owns ALL tiles that share their high bits `hi` and the popcount c of the gm "mid" bits just above the low k:
// C(gm, c) tiles of identical shape.  The ELL indices, diagonal codes and straddler tables are loaded once and applied
// to every tile of the group (index traffic / C(gm,c)), and the bonds inside the mid bits become conflict-free
// shared-memory streams between the group's own tiles instead of L2 streams.
struct U1GLayout {
  size_t xs, dval, hh_ptr, hh_amp, int_amp, int_src, ms_ptr, ms_amp, ms_lo, ms_len, mx_ptr, mx_amp, mx_toff, dt, base, cnt, total;
};
__host__ __device__ inline U1GLayout u1g_layout(int G, uint32_t tile_cap, int vec_bytes, int n_codes, int n_hh, int n_ms, int n_mx) {
  U1GLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
  L.xs = take((size_t)G * (tile_cap + 1) * vec_bytes);
  L.dval = take((size_t)G * n_codes * 8);
  L.hh_ptr = take((size_t)G * n_hh * 8);
  L.hh_amp = take((size_t)G * n_hh * 8);
  L.int_amp = take((size_t)G * n_hh * 8);
  L.int_src = take((size_t)G * n_hh * 4);
  L.ms_ptr = take((size_t)G * n_ms * 8);
  L.ms_amp = take((size_t)G * n_ms * 8);
  L.ms_lo = take((size_t)G * n_ms * 4);
  L.ms_len = take((size_t)G * n_ms * 4);
  L.mx_ptr = take((size_t)G * n_mx * 8);
  L.mx_amp = take((size_t)G * n_mx * 8);
  L.mx_toff = take((size_t)G * n_mx * 4);
  L.dt = take((size_t)G * 8);
  L.base = take((size_t)G * 8);
  L.cnt = take((size_t)G * 2 * 4);
  L.total = o;
  return L;
}

template <typename VecT>
struct U1GView {
  VecT* xs; uint32_t xstride;
  const double* dval; int n_codes;
  const VecT* const* hh_ptr; const double* hh_amp; const double* int_amp; const uint32_t* int_src;
  const VecT* const* ms_ptr; const double* ms_amp; const uint32_t* ms_lo; const uint32_t* ms_len;
  const VecT* const* mx_ptr; const double* mx_amp; const uint32_t* mx_toff;
  const double* dt; const int64_t* base; const int* cnt;
  int nh, nms, nmx;
  uint32_t lofs, gofs, size;
  int p_low;
};

// Rows i = tid + slab*THREADS of all G tiles; RS slabs per pass -> G*RS accumulators in registers.
template <typename VecT, int THREADS, int G, int RS>
__device__ __forceinline__ void u1g_body(const U1Params& P, const U1GView<VecT>& V, VecT* __restrict__ y, bool want_dot,
                                         double& dre, double& dim_) {
  constexpr int CH = (G * RS >= 12) ? 2 : (RS >= 4 ? 2 : 3);   // neighbour streams loaded back to back (x RS rows each)
  const uint32_t S = V.size;
  const int n_slab = (int)((S + THREADS - 1) / THREADS);
#pragma unroll 1
  for (int s0 = 0; s0 < n_slab; s0 += RS) {
    if ((threadIdx.x & ~31u) + (uint32_t)s0 * THREADS >= S) break;   // this warp has no rows left (warp-uniform)
    uint32_t ic[RS];
    bool ok[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) {
      const uint32_t i = threadIdx.x + (uint32_t)(s0 + r) * THREADS;
      ok[r] = i < S;
      ic[r] = ok[r] ? i : S - 1;      // clamped: every load stays in bounds, only the store is predicated
    }
    VecT acc[G][RS];
    // diagonal: one u8 code per row (shared by the group), per-tile value tables
    {
      uint32_t code[RS];
#pragma unroll
      for (int r = 0; r < RS; ++r) code[r] = P.diag_mode == 1 ? (uint32_t)__ldg(P.dcode + V.lofs + ic[r]) : 0u;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const VecT* xg = V.xs + g * V.xstride;
        const double dtg = V.dt[g];
        const double* dv = V.dval + g * V.n_codes;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
          const double d = P.diag_mode == 1 ? dtg + dv[code[r]] : dtg;
          acc[g][r] = vec_scale<VecT>(d, xg[ic[r]]);
        }
      }
    }
    // exchange bonds inside the low k bits: one index load serves 2 slots x G tiles
#pragma unroll 1
    for (int c = 0; c < P.n_ll; ++c) {
      const uint8_t* cnt = P.ell_cnt[c] + V.gofs;
      int nmax = 0;
#pragma unroll
      for (int r = 0; r < RS; ++r) nmax = max(nmax, (int)__ldg(cnt + (ic[r] >> 5)));
      const uint32_t* e = P.ell[c] + P.ell_ofs[c * (P.k + 1) + V.p_low];
      const double a = P.ll_amp[c];
      uint32_t jn[RS];
#pragma unroll
      for (int r = 0; r < RS; ++r) jn[r] = nmax > 0 ? __ldg(e + ic[r]) : 0u;
#pragma unroll 1
      for (int sl = 0; sl < nmax; ++sl) {
        uint32_t j[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) j[r] = jn[r];
        e += S;
        if (sl + 1 < nmax) {
#pragma unroll
          for (int r = 0; r < RS; ++r) jn[r] = __ldg(e + ic[r]);     // next slot pair in flight during the gathers
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const VecT* xg = V.xs + g * V.xstride;
#pragma unroll
          for (int r = 0; r < RS; ++r) {
            vec_fma(acc[g][r], a, xg[j[r] & 0xFFFFu]);
            vec_fma(acc[g][r], a, xg[j[r] >> 16]);
          }
        }
      }
    }
    // bonds inside the mid bits: the partner tile is in this CTA's shared memory, same local index
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int ni = V.cnt[2 * g + 1];
#pragma unroll 1
      for (int t = 0; t < ni; ++t) {
        const VecT* src = V.xs + V.int_src[g * V.nh + t] * V.xstride;
        const double a = V.int_amp[g * V.nh + t];
#pragma unroll
        for (int r = 0; r < RS; ++r) vec_fma(acc[g][r], a, src[ic[r]]);
      }
    }
    // straddling bonds on bit k-1: contiguous block of rows <-> shifted stream of the neighbour tile
#pragma unroll 1
    for (int e = 0; e < P.n_ms; ++e) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const double a = V.ms_amp[g * V.nms + e];
        const VecT* xe = V.ms_ptr[g * V.nms + e];
        const uint32_t lo = V.ms_lo[g * V.nms + e], len = V.ms_len[g * V.nms + e];
        VecT v[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) v[r] = (ic[r] - lo < len) ? ldg_val(xe + ic[r]) : vzero((VecT*)nullptr);
#pragma unroll
        for (int r = 0; r < RS; ++r) vec_fma(acc[g][r], a, v[r]);
      }
    }
    // other bonds straddling bit k: tabulated local column inside the neighbouring tile (0xFFFF = does not fire)
#pragma unroll 1
    for (int e = 0; e < P.n_mx; ++e) {
      uint32_t j[G][RS];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const uint16_t* tab = P.mx_tab + V.mx_toff[g * V.nmx + e];
#pragma unroll
        for (int r = 0; r < RS; ++r) j[g][r] = __ldg(tab + ic[r]);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const VecT* xe = V.mx_ptr[g * V.nmx + e];
        const double a = V.mx_amp[g * V.nmx + e];
        VecT v[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) v[r] = j[g][r] != 0xFFFFu ? ldg_val(xe + j[g][r]) : vzero((VecT*)nullptr);
#pragma unroll
        for (int r = 0; r < RS; ++r) vec_fma(acc[g][r], a, v[r]);
      }
    }
    // bonds with at least one site above the mid bits: same local index in a tile of another group (L2 / HBM streams)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int ne = V.cnt[2 * g];
      const VecT* const* pp = V.hh_ptr + g * V.nh;
      const double* aa = V.hh_amp + g * V.nh;
#pragma unroll 1
      for (int e = 0; e < ne; e += CH) {
        VecT v[CH][RS];
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (e + c < ne) {
            const VecT* xe = pp[e + c];
#pragma unroll
            for (int r = 0; r < RS; ++r) v[c][r] = ldg_val(xe + ic[r]);
          }
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (e + c < ne) {
            const double a = aa[e + c];
#pragma unroll
            for (int r = 0; r < RS; ++r) vec_fma(acc[g][r], a, v[c][r]);
          }
      }
    }
    // store (row-owner writes)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int64_t b = V.base[g];
      const VecT* xg = V.xs + g * V.xstride;
#pragma unroll
      for (int r = 0; r < RS; ++r) {
        const int64_t row = b + ic[r];
        if (!ok[r] || row < P.row_lo || row >= P.row_hi) continue;
        VecT* dst = y + (row - P.row_lo);
        VecT out = acc[g][r];
        if (P.accumulate) out = vec_add(out, *dst);
        st_stream(dst, out);
        if (want_dot) dot_acc(dre, dim_, xg[ic[r]], out);
      }
    }
  }
}

template <typename VecT, int THREADS, int RS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k2_apply_u1g(const U1Params P, VecT* __restrict__ y, double* __restrict__ dot_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nh = max(P.n_hh, 1), nms = max(P.n_ms, 1), nmx = max(P.n_mx, 1);
  const U1GLayout L = u1g_layout(P.gcap, P.tile_cap, (int)sizeof(VecT), P.n_codes, nh, nms, nmx);
  VecT* xs = reinterpret_cast<VecT*>(smem_raw + L.xs);
  double* s_dval = reinterpret_cast<double*>(smem_raw + L.dval);
  const VecT** hh_ptr = reinterpret_cast<const VecT**>(smem_raw + L.hh_ptr);
  double* hh_amp = reinterpret_cast<double*>(smem_raw + L.hh_amp);
  double* int_amp = reinterpret_cast<double*>(smem_raw + L.int_amp);
  uint32_t* int_src = reinterpret_cast<uint32_t*>(smem_raw + L.int_src);
  const VecT** ms_ptr = reinterpret_cast<const VecT**>(smem_raw + L.ms_ptr);
  double* ms_amp = reinterpret_cast<double*>(smem_raw + L.ms_amp);
  uint32_t* ms_lo = reinterpret_cast<uint32_t*>(smem_raw + L.ms_lo);
  uint32_t* ms_len = reinterpret_cast<uint32_t*>(smem_raw + L.ms_len);
  const VecT** mx_ptr = reinterpret_cast<const VecT**>(smem_raw + L.mx_ptr);
  double* mx_amp = reinterpret_cast<double*>(smem_raw + L.mx_amp);
  uint32_t* mx_toff = reinterpret_cast<uint32_t*>(smem_raw + L.mx_toff);
  double* s_dt = reinterpret_cast<double*>(smem_raw + L.dt);
  int64_t* s_base = reinterpret_cast<int64_t*>(smem_raw + L.base);
  int* s_cnt = reinterpret_cast<int*>(smem_raw + L.cnt);

  const int tid = threadIdx.x;
  const uint32_t gd = P.grp[P.grp_first + blockIdx.x];
  const uint32_t hi = gd >> 3;
  const int c = (int)(gd & 7u);
  const int G = P.gcount[c];
  const int k = P.k, gm = P.gm;
  const uint32_t midmask = (1u << gm) - 1u;
  const int p_low = P.n_set - __popc(hi) - c;
  const uint32_t lofs = P.lowofs[p_low];
  const uint32_t size = P.lowofs[p_low + 1] - lofs;
  const uint32_t xstride = P.tile_cap + 1;

  // ---- prologue: warp g resolves the bond lists of tile g (deterministic ballot compaction) --------------------
  const int wid = tid >> 5, lane = tid & 31;
  if (wid < G) {
    const int g = wid;
    const uint32_t H = (hi << gm) | P.gmid[c][g];
    int n_ext = 0, n_int = 0;
    for (int b0 = 0; b0 < P.n_hh; b0 += 32) {
      const int b = b0 + lane;
      bool fire = false, internal = false;
      uint32_t H2 = 0;
      if (b < P.n_hh) {
        const int p = P.hh_p[b], q = P.hh_q[b];
        fire = (((H >> p) ^ (H >> q)) & 1u) != 0;
        H2 = H ^ ((1u << p) | (1u << q));
        internal = fire && p < gm && q < gm;          // both sites in the mid bits: the partner is in this group
      }
      const bool ext = fire && !internal;
      const unsigned me = __ballot_sync(0xffffffffu, ext), mi = __ballot_sync(0xffffffffu, internal);
      if (ext) {
        const int slot = n_ext + __popc(me & ((1u << lane) - 1u));
        hh_ptr[g * nh + slot] = u1_seg_resolve<VecT>(P, P.tile_base[H2]);
        hh_amp[g * nh + slot] = P.hh_amp[b];
      }
      if (internal) {
        const int slot = n_int + __popc(mi & ((1u << lane) - 1u));
        const uint32_t mid2 = H2 & midmask;
        uint32_t g2 = 0;
        for (int t = 0; t < G; ++t) if (P.gmid[c][t] == mid2) g2 = (uint32_t)t;
        int_src[g * nh + slot] = g2;
        int_amp[g * nh + slot] = P.hh_amp[b];
      }
      n_ext += __popc(me);
      n_int += __popc(mi);
    }
    for (int b = lane; b < P.n_mx; b += 32) {
      const int q = P.mx_q[b];
      const uint32_t hbit = (H >> q) & 1u;
      mx_ptr[g * nmx + b] = u1_seg_resolve<VecT>(P, P.tile_base[H ^ (1u << q)]);
      mx_amp[g * nmx + b] = P.mx_amp[b];
      mx_toff[g * nmx + b] = ((uint32_t)(2 * b + hbit) << k) + lofs;
    }
    for (int b = lane; b < P.n_ms; b += 32) {
      const int q = P.ms_q[b];
      const uint32_t hbit = (H >> q) & 1u;
      const VecT* xn = u1_seg_resolve<VecT>(P, P.tile_base[H ^ (1u << q)]);
      const uint32_t n0 = P.ck1[p_low];                  // rows of this tile whose bit k-1 is clear (they come first)
      if (hbit == 0) { ms_lo[g * nms + b] = n0; ms_len[g * nms + b] = size - n0; ms_ptr[g * nms + b] = xn - n0; }
      else { ms_lo[g * nms + b] = 0; ms_len[g * nms + b] = n0; ms_ptr[g * nms + b] = xn + P.ck1[p_low + 1]; }
      ms_amp[g * nms + b] = P.ms_amp[b];
    }
    if (lane == 0) {
      s_cnt[2 * g] = n_ext;
      s_cnt[2 * g + 1] = n_int;
      s_dt[g] = P.tile_diag[H];
      s_base[g] = (int64_t)P.tile_base[H];
    }
  }
  if (P.diag_mode == 1) {
    for (int i = tid; i < G * P.n_codes; i += THREADS) {
      const int g = i / P.n_codes, code = i - g * P.n_codes;
      const uint32_t H = (hi << gm) | P.gmid[c][g];
      double v = P.dval[code];
      if (P.mq_folded) {
        const uint32_t pat = P.dpat[code];
        for (int e = 0; e < P.n_mq; ++e)
          if (((H >> P.mq_q[e]) & 1u) && ((pat >> P.mq_pidx[e]) & 1u)) v += P.mq_coef[e];
      }
      s_dval[i] = v;
    }
  }
  for (int g = 0; g < G; ++g) {
    const uint32_t H = (hi << gm) | P.gmid[c][g];
    const VecT* xo = u1_seg_resolve<VecT>(P, P.tile_base[H]);   // shards are aligned to hi blocks: one segment per tile
    VecT* xg = xs + g * xstride;
    for (uint32_t i = tid; i < size; i += THREADS) xg[i] = ldg_val(xo + i);
    if (tid == 0) xg[size] = vzero((VecT*)nullptr);
  }
  __syncthreads();

  U1GView<VecT> V;
  V.xs = xs; V.xstride = xstride; V.dval = s_dval; V.n_codes = P.n_codes;
  V.hh_ptr = hh_ptr; V.hh_amp = hh_amp; V.int_amp = int_amp; V.int_src = int_src;
  V.ms_ptr = ms_ptr; V.ms_amp = ms_amp; V.ms_lo = ms_lo; V.ms_len = ms_len;
  V.mx_ptr = mx_ptr; V.mx_amp = mx_amp; V.mx_toff = mx_toff;
  V.dt = s_dt; V.base = s_base; V.cnt = s_cnt;
  V.nh = nh; V.nms = nms; V.nmx = nmx;
  V.lofs = lofs; V.gofs = P.grpofs[p_low]; V.size = size; V.p_low = p_low;
  double dre = 0.0, dim_ = 0.0;
  const bool want_dot = dot_partials != nullptr;
  switch (G) {
    case 1: u1g_body<VecT, THREADS, 1, RS>(P, V, y, want_dot, dre, dim_); break;
    case 2: u1g_body<VecT, THREADS, 2, RS>(P, V, y, want_dot, dre, dim_); break;
    case 3: u1g_body<VecT, THREADS, 3, RS>(P, V, y, want_dot, dre, dim_); break;
    case 4: u1g_body<VecT, THREADS, 4, RS>(P, V, y, want_dot, dre, dim_); break;
    case 6: u1g_body<VecT, THREADS, 6, RS>(P, V, y, want_dot, dre, dim_); break;
    default: break;
  }
  if (dot_partials) {
    __shared__ double s_red[2][THREADS / 32];
    dre = warp_sum(dre);
    dim_ = warp_sum(dim_);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = dre; s_red[1][tid >> 5] = dim_; }
    __syncthreads();
    if (tid == 0) {
      double a = 0, cc = 0;
      for (int w = 0; w < THREADS / 32; ++w) { a += s_red[0][w]; cc += s_red[1][w]; }
      dot_partials[2 * blockIdx.x] = a;
      dot_partials[2 * blockIdx.x + 1] = cc;
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

struct GroupCfg { bool on = false; int k = 14, m = 4, threads = 1024, rs = 2; };

// EDCUDA_U1_G="k,m,threads,rows_per_pass" selects the group kernel; "0" forces the single-tile kernel
GroupCfg group_cfg(int n_bits, int vec_bytes) {
  GroupCfg g;
  if (const char* e = getenv("EDCUDA_U1_G")) {
    int k, m, t, r;
    if (sscanf(e, "%d,%d,%d,%d", &k, &m, &t, &r) == 4 && k >= 2 && k <= 16 && m >= 0 && m <= 4 && k + m <= n_bits) {
      g.on = true; g.k = k; g.m = m; g.threads = t; g.rs = r;
    }
  }
  (void)vec_bytes;
  return g;
}

int choose_k(int n_bits, int vec_bytes) {
  int k = std::min(15, std::max(4, n_bits - 10));   // k = 15 measured best on B200 (L=32: k=14 11.4 ms, k=15 9.1 ms, k=16 12.2 ms)
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

static std::shared_ptr<FastU1Plan> build_plan(ed_oprep* o, int vec_bytes, bool allow_group = true) {
  auto plan = std::make_shared<FastU1Plan>();
  ed_basis* b = o->basis;
  plan->vec_bytes = vec_bytes;
  if (b->kind != ED_BASIS_COMBINADIC || b->dim <= 0) return plan;
  const int n_bits = b->space.bits, n_set = b->n_set;
  if (n_bits < 1 || n_bits > 42) return plan;
  Lowered L = lower_operator(o->op, n_bits, n_set);
  if (!L.ok) return plan;
  GroupCfg gcfg = group_cfg(n_bits, vec_bytes);
  if (!allow_group) gcfg.on = false;
  const int k = gcfg.on ? gcfg.k : choose_k(n_bits, vec_bytes);
  const int hb = n_bits - k;
  if (hb > 26) return plan;  // tile tables of 2^hb entries
  U1Params& P = plan->P;
  memset(&P, 0, sizeof(P));
  P.n_bits = n_bits; P.n_set = n_set; P.k = k;
  plan->idx32 = b->dim < (1ll << 32);
  const uint32_t nlow = 1u << k;
  const uint32_t nH = 1u << hb;
  const uint64_t lowmask = (1ull << k) - 1ull;

  // ---- low-word enumeration ------------------------------------------------------------------------------
  std::vector<uint32_t> lowofs(k + 2, 0), grpofs(k + 2, 0);
  for (int p = 0; p <= k; ++p) {
    lowofs[p + 1] = lowofs[p] + (uint32_t)binom_u64(k, p);
    grpofs[p + 1] = grpofs[p] + (uint32_t)((binom_u64(k, p) + 31) / 32);
  }
  std::vector<uint16_t> lowword(nlow);
  std::vector<uint32_t> lowrank(nlow);   // index inside its popcount class
  {
    std::vector<uint32_t> at(k + 1, 0);
    for (uint32_t w = 0; w < nlow; ++w) {
      const int p = __builtin_popcount(w);
      lowrank[w] = at[p];
      lowword[lowofs[p] + at[p]++] = (uint16_t)w;
    }
  }

  // ---- diagonal: split into H-only (per tile), low-only (tabulated) and straddling n_p n_q terms ----------
  std::vector<double> tile_diag(nH, 0.0), dlow(nlow, 0.0);
  std::vector<uint8_t> mq_p, mq_q;
  std::vector<double> mq_coef;
  for (uint32_t H = 0; H < nH; ++H) {
    double d = L.dconst;
    const uint64_t s = (uint64_t)H << k;
    for (auto& l : L.lin) d += l.first * __builtin_popcountll(s & l.second & ~lowmask);
    for (auto& q : L.quad) {
      const uint64_t hm = q.mask & ~lowmask;   // lower site already inside H
      d += q.v * __builtin_popcountll(s & (s >> q.d) & hm);
    }
    tile_diag[H] = d;
  }
  bool any_low_diag = false;
  for (uint32_t w = 0; w < nlow; ++w) {
    double d = 0;
    for (auto& l : L.lin) d += l.first * __builtin_popcountll((uint64_t)w & l.second & lowmask);
    for (auto& q : L.quad) {
      uint64_t m = 0;   // bonds with both sites inside the low bits
      for (int p = 0; p + q.d < k; ++p) if (q.mask >> p & 1) m |= 1ull << p;
      d += q.v * __builtin_popcountll((uint64_t)w & ((uint64_t)w >> q.d) & m);
    }
    dlow[lowofs[__builtin_popcount(w)] + lowrank[w]] = d;
    if (d != 0.0) any_low_diag = true;
  }
  for (auto& q : L.quad)
    for (int p = 0; p < k; ++p)
      if ((q.mask >> p & 1) && p + q.d >= k) { mq_p.push_back((uint8_t)p); mq_q.push_back((uint8_t)(p + q.d - k)); mq_coef.push_back(q.v); }
  if ((int)mq_p.size() > U1_MAX_MQ) return plan;
  P.n_mq = (int)mq_p.size();
  std::vector<double> dval;
  std::vector<uint8_t> dcode(nlow, 0), dpat(256, 0), mq_pidx;
  P.mq_folded = 0;
  {
    // low sites the straddling n_p n_q terms look at; when (low diagonal value, those bits) takes <= 256 distinct
    // values the u8 code enumerates the pairs and the kernel folds the straddling terms into its per-tile value table
    std::vector<int> pm;
    for (uint8_t pp : mq_p) if (std::find(pm.begin(), pm.end(), (int)pp) == pm.end()) pm.push_back(pp);
    bool folded = false;
    if (!mq_p.empty() && pm.size() <= 8 && !getenv("EDCUDA_U1_NOFOLD")) {
      std::map<std::pair<double, uint32_t>, int> codes;
      bool fits = true;
      for (uint32_t i = 0; i < nlow && fits; ++i) {
        const uint32_t w = lowword[i];
        uint32_t pat = 0;
        for (size_t j = 0; j < pm.size(); ++j) pat |= ((w >> pm[j]) & 1u) << j;
        auto key = std::make_pair(dlow[i], pat);
        auto it = codes.find(key);
        if (it == codes.end()) {
          if (codes.size() >= 256) { fits = false; break; }
          it = codes.emplace(key, (int)codes.size()).first;
        }
        dcode[i] = (uint8_t)it->second;
      }
      if (fits) {
        folded = true;
        P.diag_mode = 1;
        P.mq_folded = 1;
        dval.assign(256, 0.0);
        for (auto& kv : codes) { dval[kv.second] = kv.first.first; dpat[kv.second] = (uint8_t)kv.first.second; }
        for (uint8_t pp : mq_p) mq_pidx.push_back((uint8_t)(std::find(pm.begin(), pm.end(), (int)pp) - pm.begin()));
      }
    }
    if (!folded) {
      std::fill(dcode.begin(), dcode.end(), 0);
      if (!any_low_diag) P.diag_mode = 0;
      else {
        std::map<double, int> codes;
        bool fits = true;
        for (uint32_t i = 0; i < nlow && fits; ++i) {
          auto it = codes.find(dlow[i]);
          if (it == codes.end()) {
            if (codes.size() >= 256) { fits = false; break; }
            it = codes.emplace(dlow[i], (int)codes.size()).first;
          }
          dcode[i] = (uint8_t)it->second;
        }
        if (fits) {
          P.diag_mode = 1;
          dval.assign(256, 0.0);
          for (auto& kv : codes) dval[kv.second] = kv.first;
        } else P.diag_mode = 2;
      }
    }
  }

  // ---- exchange bonds: split at bit k --------------------------------------------------------------------
  std::vector<uint8_t> hh_p, hh_q, mx_p, mx_q, ms_q;
  std::vector<double> hh_amp, mx_amp, ms_amp;
  struct LL { int d; uint32_t mask; double amp; };
  std::vector<LL> lls;
  for (auto& c : L.exch) {
    uint64_t ll = 0;
    for (int p = 0; p < n_bits; ++p) {
      if (!(c.mask >> p & 1)) continue;
      const int q = p + c.d;
      if (q < k) ll |= 1ull << p;
      else if (p >= k) { hh_p.push_back((uint8_t)(p - k)); hh_q.push_back((uint8_t)(q - k)); hh_amp.push_back(c.v); }
      else if (p == k - 1) { ms_q.push_back((uint8_t)(q - k)); ms_amp.push_back(c.v); }   // contiguous block form
      else { mx_p.push_back((uint8_t)p); mx_q.push_back((uint8_t)(q - k)); mx_amp.push_back(c.v); }
    }
    if (ll) lls.push_back({c.d, (uint32_t)(ll & lowmask), c.v});
  }
  // nearest-neighbour low bonds go to the delta LUTs (needs bit 7 inside the low word); the ELL keeps the rest
  P.lut_on = 0;
  for (int j = 0; j < 16; ++j) P.nn_amp[j] = 0.0;
  uint32_t nn_mask = 0;
  if (k >= 8 && !gcfg.on && !getenv("EDCUDA_U1_NOLUT")) {
    std::vector<LL> rest;
    for (auto& l : lls) {
      if (l.d != 1) { rest.push_back(l); continue; }
      for (int p = 0; p + 1 < k; ++p)
        if (l.mask >> p & 1u) { P.nn_amp[p] += l.amp; nn_mask |= 1u << p; }
    }
    if (nn_mask) { P.lut_on = 1; lls.swap(rest); }
  }
  if (P.lut_on) {
    std::vector<uint64_t> lut_lo(256, 0);
    for (uint32_t b = 0; b < 256; ++b) {
      uint64_t e = 0;
      for (int j = 0; j < 7; ++j) {
        if (!(nn_mask >> j & 1u) || ((b >> j) & 1u) == ((b >> (j + 1)) & 1u)) continue;
        const int r = __builtin_popcount(b & ((1u << j) - 1u));
        const int64_t dlt = (int64_t)binom_u64(j, r);
        const int8_t sd = (int8_t)(((b >> j) & 1u) ? dlt : -dlt);      // particle on j moves up: the word (and its rank) grows
        e |= (uint64_t)(uint8_t)sd << (8 * j);
      }
      lut_lo[b] = e;
    }
    const uint32_t nu = 1u << (k - 7);
    std::vector<uint4> lut_up((size_t)(k + 1) * nu, make_uint4(0, 0, 0, 0));
    for (int p = 0; p <= k; ++p)
      for (uint32_t u = 0; u < nu; ++u) {
        uint16_t ent[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < 8; ++j) {
          const int q = 7 + j;
          if (q + 1 >= k || !(nn_mask >> q & 1u)) continue;
          const uint32_t bq = (u >> j) & 1u, bq1 = (u >> (j + 1)) & 1u;
          if (bq == bq1) continue;
          const int r = p - __builtin_popcount(u >> j);               // set bits below q
          if (r < 0 || r > q) continue;                               // word not in this popcount class
          const int64_t dlt = (int64_t)binom_u64(q, r);
          ent[j] = (uint16_t)(int16_t)(bq ? dlt : -dlt);
        }
        lut_up[(size_t)p * nu + u] = make_uint4(ent[0] | ((uint32_t)ent[1] << 16), ent[2] | ((uint32_t)ent[3] << 16),
                                                ent[4] | ((uint32_t)ent[5] << 16), ent[6] | ((uint32_t)ent[7] << 16));
      }
    plan->lut_lo.upload(lut_lo);
    plan->lut_up.upload(lut_up);
    P.lut_lo = plan->lut_lo.p;
    P.lut_up = plan->lut_up.p;
  }
  // classes with the same amplitude share one ELL table
  std::map<double, std::vector<LL>> by_amp;
  for (auto& l : lls) by_amp[l.amp].push_back(l);
  if ((int)by_amp.size() > U1_MAX_CLASSES) return plan;
  if ((int)hh_p.size() > U1_MAX_HH || (int)mx_p.size() > U1_MAX_MX || (int)ms_q.size() > U1_MAX_MS) return plan;
  P.n_ms = (int)ms_q.size();
  for (int p = 0; p < 20; ++p) P.ck1[p] = (uint32_t)binom_u64(k - 1, p);
  P.n_hh = (int)hh_p.size();
  P.n_mx = (int)mx_p.size();
  P.n_ll = (int)by_amp.size();
  std::vector<uint32_t> ell_ofs((size_t)std::max(P.n_ll, 1) * (k + 1), 0);
  plan->ell.resize(P.n_ll);
  plan->ell_cnt.resize(P.n_ll);
  {
    int c = 0;
    std::vector<uint16_t> nb;
    for (auto& kv : by_amp) {
      P.ll_amp[c] = kv.first;
      std::vector<uint32_t> table;      // two consecutive slots per word
      std::vector<uint8_t> cnt(grpofs[k + 1], 0);
      for (int p = 0; p <= k; ++p) {
        const uint32_t size = lowofs[p + 1] - lowofs[p];
        // neighbour lists of every row of this popcount
        std::vector<std::vector<uint16_t>> lists(size);
        uint32_t maxslot = 0;
        for (uint32_t i = 0; i < size; ++i) {
          const uint32_t w = lowword[lowofs[p] + i];
          for (auto& l : kv.second) {
            // highest bond first: the upper bits are shared by the rows of a warp, so the first slots are
            // warp-uniform shifts (contiguous, conflict-free shared-memory reads)
            uint32_t t = (w ^ (w >> l.d)) & l.mask;
            while (t) {
              const int q = 31 - __builtin_clz(t);
              t &= ~(1u << q);
              lists[i].push_back((uint16_t)lowrank[w ^ ((1u | (1u << l.d)) << q)]);
            }
          }
          maxslot = std::max<uint32_t>(maxslot, (uint32_t)lists[i].size());
          uint8_t& g = cnt[grpofs[p] + i / 32];
          g = std::max<uint8_t>(g, (uint8_t)((lists[i].size() + 1) / 2));
        }
        const uint32_t maxpair = (maxslot + 1) / 2;
        ell_ofs[(size_t)c * (k + 1) + p] = (uint32_t)table.size();
        const size_t start = table.size();
        table.resize(start + (size_t)maxpair * size, (uint32_t)size | ((uint32_t)size << 16));   // padding -> xs[size] == 0
        for (uint32_t i = 0; i < size; ++i)
          for (size_t sl = 0; sl < lists[i].size(); ++sl) {
            uint32_t& wd = table[start + (sl / 2) * size + i];
            if (sl & 1) wd = (wd & 0x0000FFFFu) | ((uint32_t)lists[i][sl] << 16);
            else wd = (wd & 0xFFFF0000u) | lists[i][sl];
          }
      }
      if (table.empty()) table.push_back(0);
      plan->ell[c].upload(table);
      plan->ell_cnt[c].upload(cnt);
      ++c;
    }
  }
  // straddling bonds: local column in the neighbouring tile, per (bond, value of the H bit)
  std::vector<uint16_t> mx_tab((size_t)std::max(P.n_mx, 1) * 2 * nlow, 0xFFFF);
  for (int e = 0; e < P.n_mx; ++e)
    for (uint32_t hbit = 0; hbit < 2; ++hbit)
      for (uint32_t w = 0; w < nlow; ++w) {
        if (((w >> mx_p[e]) & 1u) == hbit) continue;                 // fires only when the two bits differ
        const uint32_t w2 = w ^ (1u << mx_p[e]);
        mx_tab[((size_t)(2 * e + hbit) << k) + lowofs[__builtin_popcount(w)] + lowrank[w]] = (uint16_t)lowrank[w2];
      }

  // ---- tiles ---------------------------------------------------------------------------------------------
  std::vector<uint64_t> tile_base(nH, 0);
  std::vector<uint32_t> tile_H;
  uint32_t tile_cap = 1;
  for (uint32_t H = 0; H < nH; ++H) {
    const int p_low = n_set - __builtin_popcount(H);
    if (p_low < 0 || p_low > k) continue;
    // rank of (H << k | lowest word with p_low bits) = sum over set bits of H of C(k + pos, p_low + idx + 1)
    uint64_t acc = 0; int i = 0;
    for (int q = 0; q < hb; ++q) if (H >> q & 1) { acc += binom_u64(k + q, p_low + i + 1); ++i; }
    tile_base[H] = acc;
    tile_H.push_back(H);
    plan->h_base.push_back(acc);
    plan->h_size.push_back(binom_u64(k, p_low));
    tile_cap = std::max<uint32_t>(tile_cap, (uint32_t)binom_u64(k, p_low));
  }
  P.tile_cap = tile_cap;
  plan->n_tiles = (int)tile_H.size();
  // optional launch order for whole-basis launches: tiles grouped by a window of "slow" H bits, so that the tiles
  // running at the same time are closed under the bonds on the remaining (fast) bits -- including the periodic bond,
  // whose H bit is the top one -- and find each other's x in L2.  EDCUDA_U1_ORDER="first_slow_bit,n_slow_bits".
  plan->order_on = false;
  if (const char* e = getenv("EDCUDA_U1_ORDER")) {
    int s0 = -1, ns = 0;
    if (sscanf(e, "%d,%d", &s0, &ns) == 2 && s0 >= 0 && ns > 0 && s0 + ns <= hb) {
      std::vector<uint32_t> order(tile_H.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
      const uint32_t smask = ((1u << ns) - 1u) << s0;
      std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (tile_H[a] & smask) < (tile_H[b] & smask); });
      plan->tile_order.upload(order);
      plan->order_on = true;
    }
  }

  // ---- groups (k2_apply_u1g): every (hi, popcount(mid)) class with a valid low popcount --------------------------
  P.gm = 0; P.gcap = 1; P.n_codes = 1; P.grp = nullptr; P.grp_first = 0;
  if (P.diag_mode == 1) { int mc = 0; for (uint8_t cde : dcode) mc = std::max(mc, (int)cde); P.n_codes = mc + 1; }
  if (gcfg.on && P.diag_mode != 2 && (P.n_mq == 0 || P.mq_folded) && gcfg.m <= hb) {
    const int gm = gcfg.m;
    P.gm = gm;
    memset(P.gcount, 0, sizeof(P.gcount));
    memset(P.gmid, 0, sizeof(P.gmid));
    for (uint32_t mid = 0; mid < (1u << gm); ++mid) {
      const int cc = __builtin_popcount(mid);
      P.gmid[cc][P.gcount[cc]++] = (uint8_t)mid;
    }
    for (int cc = 0; cc <= gm; ++cc) P.gcap = std::max<int>(P.gcap, P.gcount[cc]);
    std::vector<uint32_t> grp;
    for (uint32_t hi = 0; hi < (nH >> gm); ++hi) {
      bool any = false;
      for (int cc = 0; cc <= gm; ++cc) {
        const int p_low = n_set - __builtin_popcount(hi) - cc;
        if (p_low < 0 || p_low > k) continue;
        grp.push_back((hi << 3) | (uint32_t)cc);
        uint64_t lo = ~0ull, hiend = 0;
        for (int g = 0; g < P.gcount[cc]; ++g) {
          const uint32_t H = (hi << gm) | P.gmid[cc][g];
          lo = std::min(lo, tile_base[H]);
          hiend = std::max(hiend, tile_base[H] + binom_u64(k, p_low));
        }
        plan->g_lo.push_back(lo);
        plan->g_hi.push_back(hiend);
        if (!any) { plan->blk_base.push_back(lo); any = true; }
        else plan->blk_base.back() = std::min(plan->blk_base.back(), lo);
      }
    }
    const U1GLayout GL = u1g_layout(P.gcap, tile_cap, vec_bytes, P.n_codes, std::max(P.n_hh, 1), std::max(P.n_ms, 1), std::max(P.n_mx, 1));
    if (!grp.empty() && GL.total <= 227 * 1024) {
      plan->grp.upload(grp);
      P.grp = plan->grp.p;
      plan->g_on = true;
      plan->g_threads = gcfg.threads;
      plan->g_rs = gcfg.rs;
      plan->g_smem = GL.total;
    }
  }

  auto nonempty8 = [](std::vector<uint8_t>& v) { if (v.empty()) v.push_back(0); };
  auto nonemptyd = [](std::vector<double>& v) { if (v.empty()) v.push_back(0.0); };
  nonempty8(hh_p); nonempty8(hh_q); nonempty8(mx_q); nonempty8(mq_p); nonempty8(mq_q); nonempty8(ms_q); nonempty8(mq_pidx);
  nonemptyd(hh_amp); nonemptyd(mx_amp); nonemptyd(mq_coef); nonemptyd(dval); nonemptyd(ms_amp);
  plan->tile_H.upload(tile_H); plan->tile_base.upload(tile_base); plan->tile_diag.upload(tile_diag);
  plan->lowword.upload(lowword); plan->lowofs.upload(lowofs); plan->grpofs.upload(grpofs);
  plan->dcode.upload(dcode); plan->dval.upload(dval); plan->dlow.upload(dlow);
  plan->mq_p.upload(mq_p); plan->mq_q.upload(mq_q); plan->mq_coef.upload(mq_coef);
  plan->dpat.upload(dpat); plan->mq_pidx.upload(mq_pidx);
  plan->ell_ofs.upload(ell_ofs);
  plan->hh_p.upload(hh_p); plan->hh_q.upload(hh_q); plan->hh_amp.upload(hh_amp);
  plan->mx_q.upload(mx_q); plan->mx_amp.upload(mx_amp); plan->mx_tab.upload(mx_tab);
  plan->ms_q.upload(ms_q); plan->ms_amp.upload(ms_amp);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  P.tile_H = plan->tile_H.p; P.tile_base = plan->tile_base.p; P.tile_diag = plan->tile_diag.p;
  P.lowword = plan->lowword.p; P.lowofs = plan->lowofs.p; P.grpofs = plan->grpofs.p;
  P.dcode = plan->dcode.p; P.dval = plan->dval.p; P.dlow = plan->dlow.p;
  P.mq_p = plan->mq_p.p; P.mq_q = plan->mq_q.p; P.mq_coef = plan->mq_coef.p;
  P.dpat = plan->dpat.p; P.mq_pidx = plan->mq_pidx.p;
  P.ell_ofs = plan->ell_ofs.p;
  for (int c = 0; c < P.n_ll; ++c) { P.ell[c] = plan->ell[c].p; P.ell_cnt[c] = plan->ell_cnt[c].p; }
  P.hh_p = plan->hh_p.p; P.hh_q = plan->hh_q.p; P.hh_amp = plan->hh_amp.p;
  P.mx_q = plan->mx_q.p; P.mx_amp = plan->mx_amp.p; P.mx_tab = plan->mx_tab.p;
  P.ms_q = plan->ms_q.p; P.ms_amp = plan->ms_amp.p;
  P.cap_hh = std::max(P.n_hh, 1); P.cap_mx = std::max(P.n_mx, 1); P.cap_ms = std::max(P.n_ms, 1);
  P.cap_mq = P.mq_folded ? 1 : std::max(P.n_mq, 1);
  P.cap_codes = P.diag_mode == 1 ? P.n_codes : 1;
  plan->smem_bytes = (size_t)(tile_cap + 1) * vec_bytes + (size_t)(P.cap_hh + P.cap_mx + P.cap_mq + P.cap_codes) * 8 +
                     (size_t)(P.cap_hh + P.cap_mx) * 8 + (size_t)P.cap_ms * 16 + (size_t)(P.cap_mx + P.cap_mq + 2 * P.cap_ms) * 4;
  if (P.lut_on) plan->smem_bytes += 16 + ((size_t)16 << (k - 7)) + 256 * 8;
  plan->smem_bytes = (plan->smem_bytes + 15) & ~(size_t)15;
  if (gcfg.on && !plan->g_on) return build_plan(o, vec_bytes, false);   // group layout does not fit: single-tile kernel, its own k
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

constexpr int U1_THREADS = 512;

template <typename VecT, int R>
static void launch_u1(FastU1Plan* plan, const U1Params& P, int n_launch, void* out, double* partials) {
  auto kern = k2_apply_u1<VecT, U1_THREADS, R>;
  static thread_local size_t configured = 0;
  if (plan->smem_bytes > 48 * 1024 && configured < plan->smem_bytes) {
    ED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes));
    configured = plan->smem_bytes;
  }
  ED_LAUNCH(kern, n_launch, U1_THREADS, plan->smem_bytes, P, reinterpret_cast<VecT*>(out), partials);
}

template <typename VecT, int THREADS, int RS, int MINB>
static void launch_u1g_inst(FastU1Plan* plan, const U1Params& P, int n_launch, void* out, double* partials) {
  auto kern = k2_apply_u1g<VecT, THREADS, RS, MINB>;
  static thread_local size_t configured = 0;
  if (plan->g_smem > 48 * 1024 && configured < plan->g_smem) {
    ED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->g_smem));
    configured = plan->g_smem;
  }
  ED_LAUNCH(kern, n_launch, THREADS, plan->g_smem, P, reinterpret_cast<VecT*>(out), partials);
}

template <typename VecT>
static void launch_u1g(FastU1Plan* plan, const U1Params& P, int n_launch, void* out, double* partials) {
  const bool two = plan->g_smem <= 112 * 1024;     // two CTAs per SM fit
  const int t = plan->g_threads, r = plan->g_rs;
  if (t == 1024 && r == 2) launch_u1g_inst<VecT, 1024, 2, 1>(plan, P, n_launch, out, partials);
  else if (t == 896 && r == 2) launch_u1g_inst<VecT, 896, 2, 1>(plan, P, n_launch, out, partials);
  else if (t == 768 && r == 2) launch_u1g_inst<VecT, 768, 2, 1>(plan, P, n_launch, out, partials);
  else if (t == 512 && r == 4 && !two) launch_u1g_inst<VecT, 512, 4, 1>(plan, P, n_launch, out, partials);
  else if (t == 512 && r == 4) launch_u1g_inst<VecT, 512, 4, 2>(plan, P, n_launch, out, partials);
  else if (t == 512 && r == 2 && two) launch_u1g_inst<VecT, 512, 2, 2>(plan, P, n_launch, out, partials);
  else if (t == 448 && r == 2 && two) launch_u1g_inst<VecT, 448, 2, 2>(plan, P, n_launch, out, partials);
  else if (t == 448 && r == 4 && two) launch_u1g_inst<VecT, 448, 4, 2>(plan, P, n_launch, out, partials);
  else ED_REQUIRE(false, ED_ERR_ARGUMENT, "EDCUDA_U1_G: no kernel instance for this (threads, rows per pass, shared memory) combination");
}

// contiguous, count-balanced row ranges whose boundaries fall on tile boundaries (so every tile, and therefore every
// neighbour stream, lives in exactly one x segment)
void ed_u1_suggest_rows(ed_oprep* o, int dtype, int world, int rank, int64_t* lo, int64_t* hi) {
  FastU1Plan* plan = get_plan(o, dtype);
  const int64_t dim = o->dim;
  auto snap = [&](int64_t target) -> int64_t {
    if (target <= 0) return 0;
    if (target >= dim) return dim;
    if (!plan->supported) return target;
    const std::vector<uint64_t>& bases = plan->g_on ? plan->blk_base : plan->h_base;
    auto it = std::lower_bound(bases.begin(), bases.end(), (uint64_t)target);
    int64_t up = it == bases.end() ? dim : (int64_t)*it;
    int64_t down = it == bases.begin() ? 0 : (int64_t)*(it - 1);
    return (up - target <= target - down) ? up : down;
  };
  *lo = snap(dim / world * rank + std::min<int64_t>(rank, dim % world));
  *hi = snap(dim / world * (rank + 1) + std::min<int64_t>(rank + 1, dim % world));
}

void ed_apply_u1(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot) {
  (void)side;
  FastU1Plan* plan = get_plan(o, dtype);
  ED_REQUIRE(plan->supported, ED_ERR_INTERNAL, "u1 fast path requested for an unsupported representation");
  U1Params P = plan->P;
  P.row_lo = o->row_lo;
  P.row_hi = o->row_hi;
  P.accumulate = accumulate;
  // profiling knob (results are WRONG when set): drop parts of the kernel to measure what each costs
  static const int ablate = getenv("EDCUDA_U1_ABLATE") ? atoi(getenv("EDCUDA_U1_ABLATE")) : 0;
  if (ablate & 1) P.n_ll = 0;
  if (ablate & 2) P.n_hh = 0;
  if (ablate & 4) P.n_mx = 0;
  if (ablate & 8) P.n_ms = 0;
  if (ablate & 16) P.n_mq = 0;
  if (ablate & 32) P.diag_mode = 0;
  if (o->x_seg_ptr.empty()) {
    ED_REQUIRE(x != nullptr, ED_ERR_ARGUMENT, "null input vector");
    P.n_seg = 1;
    P.seg_lo[0] = 0; P.seg_lo[1] = o->dim;
    P.seg_ptr[0] = x;
  } else {
    P.n_seg = (int)o->x_seg_ptr.size();
    for (int s = 0; s <= P.n_seg; ++s) P.seg_lo[s] = o->x_seg_lo[s];
    for (int s = 0; s < P.n_seg; ++s) {
      P.seg_ptr[s] = o->x_seg_ptr[s];
      if (s > 0) {
        const uint64_t b = (uint64_t)o->x_seg_lo[s];
        const std::vector<uint64_t>& bases = plan->g_on ? plan->blk_base : plan->h_base;
        ED_REQUIRE(b == (uint64_t)o->dim || std::binary_search(bases.begin(), bases.end(), b), ED_ERR_ARGUMENT,
                   "x segment boundaries must fall on tile boundaries (use ed_oprep_suggest_rows)");
      }
    }
  }
  if (plan->g_on) {
    // groups overlapping the owned rows (a group's tiles interleave with other groups of its hi block)
    static thread_local FastU1Plan* c_plan = nullptr;
    static thread_local int64_t c_lo = -1, c_hi = -1;
    static thread_local int c_first = 0, c_n = 0;
    if (c_plan != plan || c_lo != o->row_lo || c_hi != o->row_hi) {
      int gf = -1, gl = -1;
      for (size_t g = 0; g < plan->g_lo.size(); ++g)
        if ((int64_t)plan->g_hi[g] > o->row_lo && (int64_t)plan->g_lo[g] < o->row_hi) { if (gf < 0) gf = (int)g; gl = (int)g; }
      c_plan = plan; c_lo = o->row_lo; c_hi = o->row_hi;
      c_first = std::max(gf, 0); c_n = gf < 0 ? 0 : gl - gf + 1;
    }
    P.grp_first = c_first;
    const int n_launch = c_n;
    if (n_launch <= 0 || o->row_hi <= o->row_lo) {
      if (alpha_dot) ED_CUDA(cudaMemsetAsync(alpha_dot, 0, 2 * sizeof(double), ed_stream()));
      return;
    }
    double* partials = nullptr;
    if (alpha_dot) {
      if (plan->partials.n < (size_t)2 * plan->g_lo.size()) plan->partials.alloc((size_t)2 * plan->g_lo.size());
      partials = plan->partials.p;
    }
    if (dtype == ED_F64) launch_u1g<double>(plan, P, n_launch, out, partials);
    else launch_u1g<c128>(plan, P, n_launch, out, partials);
    if (alpha_dot) ed_reduce_pairs(partials, n_launch, alpha_dot);
    return;
  }
  // tiles overlapping the owned rows [row_lo, row_hi)
  int first = (int)(std::upper_bound(plan->h_base.begin(), plan->h_base.end(), (uint64_t)std::max<int64_t>(o->row_lo, 0)) - plan->h_base.begin()) - 1;
  first = std::max(first, 0);
  int last = (int)(std::lower_bound(plan->h_base.begin(), plan->h_base.end(), (uint64_t)o->row_hi) - plan->h_base.begin());
  const int n_launch = last - first;
  P.tile_first = first;
  P.tile_order = (plan->order_on && first == 0 && n_launch == plan->n_tiles) ? plan->tile_order.p : nullptr;
  if (n_launch <= 0 || o->row_hi <= o->row_lo) {
    if (alpha_dot) ED_CUDA(cudaMemsetAsync(alpha_dot, 0, 2 * sizeof(double), ed_stream()));
    return;
  }
  double* partials = nullptr;
  if (alpha_dot) {
    if (plan->partials.n < (size_t)2 * plan->n_tiles) plan->partials.alloc((size_t)2 * plan->n_tiles);
    partials = plan->partials.p;
  }
  // rows per thread and pass: 7 (40 registers, 3 CTAs/SM) measured 8.94 ms vs 9.11 ms for 13 (64 registers, 2 CTAs/SM)
  static const int r_f64 = getenv("EDCUDA_U1_R") ? atoi(getenv("EDCUDA_U1_R")) : 7;
  if (dtype == ED_F64 && r_f64 == 7) launch_u1<double, 7>(plan, P, n_launch, out, partials);
  else if (dtype == ED_F64 && r_f64 == 5) launch_u1<double, 5>(plan, P, n_launch, out, partials);
  else if (dtype == ED_F64 && r_f64 == 4) launch_u1<double, 4>(plan, P, n_launch, out, partials);
  else if (dtype == ED_F64 && r_f64 == 6) launch_u1<double, 6>(plan, P, n_launch, out, partials);
  else if (dtype == ED_F64) launch_u1<double, 13>(plan, P, n_launch, out, partials);
  else launch_u1<c128, 7>(plan, P, n_launch, out, partials);
  if (alpha_dot) ed_reduce_pairs(partials, n_launch, alpha_dot);
}
