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
  const uint32_t* ell_ofs;                 // [n_ll * (k+1)] start of popcount p inside ell[c] (multiple of 32 words)
  int n_hh;   const uint8_t* hh_p; const uint8_t* hh_q; const double* hh_amp;   // exchange bonds inside H
  int n_mx;   const uint8_t* mx_q; const double* mx_amp; const uint16_t* mx_tab; // straddling exchange bonds (gathered)
  // straddling bonds whose low site is bit k-1: the firing rows are a contiguous block of the tile and so are their
  // columns in the neighbour tile -> a shifted coalesced stream like the high-bit bonds
  int n_ms;   const uint8_t* ms_q; const double* ms_amp;
  uint32_t ck1[20];             // C(k-1, p)
  int64_t row_lo, row_hi;
  int accumulate;
  int tile_first;               // first tile of the launch (row shards launch only the tiles they overlap)
  // split exchange (multi-GPU): stream_mode 1 = LOCAL pass (neighbour streams whose tile lives in a peer segment are
  // skipped), 2 = REMOTE pass (only those streams, read from `mirror` -- a full-length local vector whose needed remote
  // rows were filled by copy engines meanwhile -- and added to y); 0 = everything in one pass through the segments
  int stream_mode;
  uint32_t local_seg_mask;      // bit s: segment s is this rank's own memory
  const void* mirror;
  int far_bit;                  // bonds whose upper H bit is >= far_bit read their neighbour tile with evict-first loads:
                                // its re-use distance (2^(bit+1) tiles) exceeds the L2, so the line is a one-shot

  const uint32_t* tile_order;   // optional launch order (whole-basis launches): position -> index into tile_H
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

template <typename VecT>
__device__ __forceinline__ const VecT* u1_seg_resolve(const U1Params& P, uint64_t idx, int& seg) {
  int s = 0;
  while (s + 1 < P.n_seg && (int64_t)idx >= P.seg_lo[s + 1]) ++s;
  seg = s;
  return reinterpret_cast<const VecT*>(P.seg_ptr[s]) + ((int64_t)idx - P.seg_lo[s]);
}

// neighbour tile starting at global row `idx` under the exchange mode: false = this pass does not read it
template <typename VecT>
__device__ __forceinline__ bool u1_neighbour(const U1Params& P, uint64_t idx, const VecT*& ptr) {
  int seg;
  ptr = u1_seg_resolve<VecT>(P, idx, seg);
  if (P.stream_mode == 0) return true;
  const bool local = (P.local_seg_mask >> seg) & 1u;
  if (P.stream_mode == 1) return local;
  ptr = reinterpret_cast<const VecT*>(P.mirror) + idx;
  return !local;
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
  DevBuf<double> tile_diag, dval, dlow, hh_amp, mx_amp, mq_coef, ms_amp;
  std::vector<DevBuf<uint32_t>> ell;
  std::vector<DevBuf<uint8_t>> ell_cnt;
  DevBuf<double> partials;
  std::vector<uint64_t> h_base, h_size;  // per non-empty tile, ascending
  bool order_on = false;
  bool wraps = false;       // a tabulated straddler flips the top site (the periodic bond of a ring)
  int hb = 0;               // bits of H
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
// one-shot reads of far tiles: streaming (evict-first) loads so that they do not push the re-used near tiles out of L2
__device__ __forceinline__ double ldcs_val(const double* p) { return __ldcs(p); }
__device__ __forceinline__ c128 ldcs_val(const c128* p) { const double2 v = __ldcs(reinterpret_cast<const double2*>(p)); return make_c128(v.x, v.y); }
// the same load instruction for both kinds of tile, the L2 eviction priority in a (warp-uniform) policy register
__device__ __forceinline__ uint64_t l2_policy(int kind) {   // 0 normal, 1 evict first, 2 evict last
  uint64_t pol;
  if (kind == 1) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ld_hint(const double* p, uint64_t pol) {
  double v;
  asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ c128 ld_hint(const c128* p, uint64_t pol) {
  c128 v;
  asm("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.re), "=d"(v.im) : "l"(p), "l"(pol));
  return v;
}
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
  VecT acc_t = vzero((VecT*)nullptr);
#pragma unroll
  for (int r = 0; r < NF; ++r) acc[r] = vzero((VecT*)nullptr);

  // diagonal (reads the x tile in shared memory)
  auto part_diag = [&]() {
    const uint8_t* dc = P.dcode + T.lofs;
    const double* dl = P.dlow + T.lofs;
    double dd[NF > 0 ? NF : 1];
    double dt = T.d_tile;
    if (P.stream_mode == 2) {                    // remote pass: only the contributions of the peer tiles
#pragma unroll
      for (int r = 0; r < NF; ++r) dd[r] = 0.0;
      dt = 0.0;
    } else if (P.diag_mode == 1) {
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
    if (T.n_mq && P.stream_mode != 2) {
      const uint16_t* lw = P.lowword + T.lofs;
#pragma unroll
      for (int r = 0; r <= NF; ++r) {
        const uint32_t low = __ldg(lw + (r < NF ? tid + r * THREADS : it));
        double d = 0.0;
        for (int e = 0; e < T.n_mq; ++e) d += ((low >> T.mq_bit[e]) & 1u) ? T.mq_coef[e] : 0.0;
        if (r < NF) dd[r] += d; else dt += d;
      }
    }
    // the accumulators are still zero on the default path (fma(d, x, 0) == d * x exactly)
#pragma unroll
    for (int r = 0; r < NF; ++r) vec_fma(acc[r], dd[r], xs[tid + r * THREADS]);
    vec_fma(acc_t, dt, xs[it]);
  };
  // bonds inside the high bits: same local index in another tile -> coalesced streams
  auto part_high = [&]() {
#pragma unroll 1
    for (int e = 0; e < T.n_hh; ++e) {
      const double a = T.hh_amp[e];
      const uintptr_t tagged = reinterpret_cast<uintptr_t>(T.hh_ptr[e]);
      const bool far = tagged & 1u;                        // warp-uniform: one-shot read of a far tile
      const VecT* xe = reinterpret_cast<const VecT*>(tagged & ~(uintptr_t)1u);
      const VecT* xt = xe + tid;
      const uint64_t pol = l2_policy(far ? 1 : 0);
      const VecT vt = ld_hint(xe + it, pol);
#pragma unroll
      for (int r0 = 0; r0 < NF; r0 += CH) {
        VecT v[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (r0 + c < NF) v[c] = ld_hint(xt + (r0 + c) * THREADS, pol);
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (r0 + c < NF) vec_fma(acc[r0 + c], a, v[c]);
      }
      vec_fma(acc_t, a, vt);
    }
  };
  // straddling bonds on bit k-1: a contiguous block of rows reads a shifted stream of the neighbour tile
  auto part_block = [&]() {
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
  };
  // exchange bonds inside the low k bits: ELL table of local columns (top bond first), shared-memory gathers.
  //  * a table word holds the BYTE offsets (8 * column) of two slots, so a gather is mask/shift + LDS [offset + base];
  //  * every slot row starts on a 128-byte boundary (row stride = size rounded up to 32 words): one line per warp load.
  // Measured and dropped: stopping every slab at the slot count of its own 32-row group (16 % fewer slot pairs, but the
  // per-slab predicates become branches: 8.14 ms vs 7.90 ms).
  auto part_low = [&]() {
    constexpr int SH = sizeof(VecT) == 16 ? 1 : 0;
    const char* xsb = reinterpret_cast<const char*>(xs);
    const uint32_t size_pad = (size + 31u) & ~31u;
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
            if (r0 + cc < NF) {
              vec_fma(acc[r0 + cc], a, *reinterpret_cast<const VecT*>(xsb + ((j[cc] & 0xFFFFu) << SH)));
              vec_fma(acc[r0 + cc], a, *reinterpret_cast<const VecT*>(xsb + ((j[cc] >> 16) << SH)));
            }
        }
        vec_fma(acc_t, a, *reinterpret_cast<const VecT*>(xsb + ((jt & 0xFFFFu) << SH)));
        vec_fma(acc_t, a, *reinterpret_cast<const VecT*>(xsb + ((jt >> 16) << SH)));
        e += size_pad;
      }
    }
  };
  // bonds straddling bit k: tabulated local column inside the neighbouring tile (0xFFFF = does not fire)
  auto part_straddle = [&]() {
#pragma unroll 1
    for (int e = 0; e < T.n_mx; ++e) {
      const uint16_t* tab = P.mx_tab + T.mx_toff[e];
      const uint16_t* tt = tab + tid;
      const uintptr_t tagged = reinterpret_cast<uintptr_t>(T.mx_ptr[e]);
      const VecT* xe = reinterpret_cast<const VecT*>(tagged & ~(uintptr_t)1u);
      const uint64_t pol = l2_policy((tagged & 1u) ? 1 : 0);
      const double a = T.mx_amp[e];
      const uint32_t jt = __ldg(tab + it);
      const VecT vt = jt != 0xFFFFu ? ld_hint(xe + jt, pol) : vzero((VecT*)nullptr);
#pragma unroll
      for (int r0 = 0; r0 < NF; r0 += CH) {
        uint32_t j[CH];
        VecT v[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (r0 + c < NF) j[c] = __ldg(tt + (r0 + c) * THREADS);
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (r0 + c < NF) v[c] = j[c] != 0xFFFFu ? ld_hint(xe + j[c], pol) : vzero((VecT*)nullptr);
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (r0 + c < NF) vec_fma(acc[r0 + c], a, v[c]);
      }
      vec_fma(acc_t, a, vt);
    }
  };

  part_diag();
  part_high();
  part_block();
  part_low();
  part_straddle();

  // store (row-owner writes)
  const int64_t row0 = (int64_t)T.base + tid;
  VecT* yt = y + (row0 - P.row_lo);
  const bool whole = (int64_t)T.base >= P.row_lo && (int64_t)T.base + size <= P.row_hi;   // tile fully owned (uniform)
#pragma unroll
  for (int r = 0; r < NF; ++r) {
    const int64_t row = row0 + r * THREADS;
    if (!whole && (row < P.row_lo || row >= P.row_hi)) continue;
    VecT out = acc[r];
    if (want_dot && P.stream_mode == 2) dot_acc(dre, dim_, xs[tid + r * THREADS], out);
    if (P.accumulate) out = vec_add(out, yt[r * THREADS]);
    st_stream(yt + r * THREADS, out);
    if (want_dot && P.stream_mode != 2) dot_acc(dre, dim_, xs[tid + r * THREADS], out);
  }
  const int64_t row = (int64_t)T.base + i_tail;
  if (tail_ok && (whole || (row >= P.row_lo && row < P.row_hi))) {
    VecT* dst = y + (row - P.row_lo);
    VecT out = acc_t;
    if (want_dot && P.stream_mode == 2) dot_acc(dre, dim_, xs[i_tail], out);
    if (P.accumulate) out = vec_add(out, *dst);
    st_stream(dst, out);
    if (want_dot && P.stream_mode != 2) dot_acc(dre, dim_, xs[i_tail], out);
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
__global__ void __launch_bounds__(THREADS, (R <= 7 && sizeof(VecT) == 8) ? 3 : 2)
k2_apply_u1(const U1Params P, VecT* __restrict__ y, double* __restrict__ dot_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VecT* xs = reinterpret_cast<VecT*>(smem_raw);                       // tile_cap + 1 (last = 0: ELL padding target)
  double* hh_amp = reinterpret_cast<double*>(xs + P.tile_cap + 1);
  double* mx_amp = hh_amp + U1_MAX_HH;
  double* mq_coef = mx_amp + U1_MAX_MX;
  double* s_dval = mq_coef + U1_MAX_MQ;                                // 256
  const VecT** hh_ptr = reinterpret_cast<const VecT**>(s_dval + 256);
  const VecT** mx_ptr = hh_ptr + U1_MAX_HH;
  uint32_t* mx_toff = reinterpret_cast<uint32_t*>(mx_ptr + U1_MAX_MX);
  uint32_t* mq_bit = mx_toff + U1_MAX_MX;
  uint32_t* ms_lo = mq_bit + U1_MAX_MQ;
  uint32_t* ms_len = ms_lo + U1_MAX_MS;
  double* ms_amp = reinterpret_cast<double*>(ms_len + U1_MAX_MS);
  const VecT** ms_ptr = reinterpret_cast<const VecT**>(ms_amp + U1_MAX_MS);
  __shared__ int s_counts[4];

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
      const VecT* xn = nullptr;
      if (fire) fire = u1_neighbour<VecT>(P, P.tile_base[H2], xn);
      const unsigned m = __ballot_sync(0xffffffffu, fire);
      if (fire) {
        const int slot = n + __popc(m & ((1u << tid) - 1u));
        hh_ptr[slot] = reinterpret_cast<const VecT*>(reinterpret_cast<uintptr_t>(xn) | ((int)P.hh_q[b] >= P.far_bit ? 1u : 0u));
        hh_amp[slot] = P.hh_amp[b];
      }
      n += __popc(m);
    }
    if (tid == 0) s_counts[0] = n;
  } else if (tid < 64) {
    const int lane = tid - 32;
    int n = 0;
    for (int b0 = 0; b0 < P.n_mx; b0 += 32) {
      const int b = b0 + lane;
      bool on = b < P.n_mx;
      const VecT* xn = nullptr;
      int q = 0;
      if (on) { q = P.mx_q[b]; on = u1_neighbour<VecT>(P, P.tile_base[H ^ (1u << q)], xn); }
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (on) {
        const int slot = n + __popc(m & ((1u << lane) - 1u));
        const uint32_t hbit = (H >> q) & 1u;
        mx_ptr[slot] = reinterpret_cast<const VecT*>(reinterpret_cast<uintptr_t>(xn) | (q >= P.far_bit ? 1u : 0u));
        mx_amp[slot] = P.mx_amp[b];
        mx_toff[slot] = ((uint32_t)(2 * b + hbit) << k) + lofs;
      }
      n += __popc(m);
    }
    if (lane == 0) s_counts[2] = n;
    n = 0;
    for (int b0 = 0; b0 < P.n_ms; b0 += 32) {
      const int b = b0 + lane;
      bool on = b < P.n_ms;
      const VecT* xn = nullptr;
      int q = 0;
      if (on) { q = P.ms_q[b]; on = u1_neighbour<VecT>(P, P.tile_base[H ^ (1u << q)], xn); }
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (on) {
        const int slot = n + __popc(m & ((1u << lane) - 1u));
        const uint32_t hbit = (H >> q) & 1u;
        const uint32_t n0 = P.ck1[p_low];                  // rows of this tile whose bit k-1 is clear (they come first)
        if (hbit == 0) {                                   // rows with bit k-1 set -> first rows of the neighbour (p_low - 1)
          ms_lo[slot] = n0; ms_len[slot] = size - n0; ms_ptr[slot] = xn - n0;
        } else {                                           // rows with bit k-1 clear -> last rows of the neighbour (p_low + 1)
          ms_lo[slot] = 0; ms_len[slot] = n0; ms_ptr[slot] = xn + P.ck1[p_low + 1];
        }
        ms_amp[slot] = P.ms_amp[b];
      }
      n += __popc(m);
    }
    if (lane == 0) s_counts[3] = n;
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
    for (int i = tid; i < 256; i += THREADS) {
      double v = P.dval[i];
      if (P.mq_folded) {
        const uint32_t pat = P.dpat[i];
        for (int e = 0; e < P.n_mq; ++e)
          if (((H >> P.mq_q[e]) & 1u) && ((pat >> P.mq_pidx[e]) & 1u)) v += P.mq_coef[e];
      }
      s_dval[i] = v;
    }
  }
  if (P.stream_mode == 2) {                             // remote pass: most tiles have nothing to add
    __syncthreads();
    if (s_counts[0] + s_counts[2] + s_counts[3] == 0) {
      if (dot_partials && tid == 0) { dot_partials[2 * blockIdx.x] = 0.0; dot_partials[2 * blockIdx.x + 1] = 0.0; }
      return;
    }
  }
  {
    const VecT* xo = u1_seg_resolve<VecT>(P, base);     // segments are tile aligned: the whole tile is in one segment
    for (uint32_t i = tid; i < size; i += THREADS) xs[i] = ldg_val(xo + i);
  }
  if (tid == 0) xs[size] = vzero((VecT*)nullptr);
  __syncthreads();

  U1Tile<VecT> T;
  T.xs = xs;
  T.hh_amp = hh_amp; T.hh_ptr = hh_ptr; T.n_hh = s_counts[0];
  T.mx_amp = mx_amp; T.mx_ptr = mx_ptr; T.mx_toff = mx_toff; T.n_mx = s_counts[2];
  T.ms_amp = ms_amp; T.ms_ptr = ms_ptr; T.ms_lo = ms_lo; T.ms_len = ms_len; T.n_ms = s_counts[3];
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
  int k = std::min(15, std::max(4, n_bits - 10));   // k = 15 measured best on B200 (L=32: k=14 11.4 ms, k=15 9.1 ms, k=16 12.2 ms)
  k = std::min(k, n_bits);
  if (const char* e = getenv("EDCUDA_U1_K")) {
    int v = atoi(e);
    if (v >= 1 && v <= 16) k = std::min(v, n_bits);
  }
  // x tile must leave room for two CTAs per SM: <= 104 KB; the ELL table stores 8 * column in 16 bits
  while (k > 1 && (binom_u64(k, k / 2) * (uint64_t)vec_bytes > 104 * 1024 || binom_u64(k, k / 2) * 8 > 65535)) --k;
  return k;
}

}  // namespace

static std::shared_ptr<FastU1Plan> build_plan(ed_oprep* o, int vec_bytes) {
  auto plan = std::make_shared<FastU1Plan>();
  ed_basis* b = o->basis;
  plan->vec_bytes = vec_bytes;
  if (b->kind != ED_BASIS_COMBINADIC || b->dim <= 0) return plan;
  const int n_bits = b->space.bits, n_set = b->n_set;
  if (n_bits < 1 || n_bits > 42) return plan;
  Lowered L = lower_operator(o->op, n_bits, n_set);
  if (!L.ok) return plan;
  const int k = choose_k(n_bits, vec_bytes);
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
  // classes with the same amplitude share one ELL table (at most 60 bonds per class keeps the u8 pair counts small)
  std::vector<std::pair<double, std::vector<LL>>> by_amp;
  {
    std::map<double, std::vector<LL>> grouped;
    for (auto& l : lls) grouped[l.amp].push_back(l);
    for (auto& kv : grouped) {
      std::vector<LL> cur;
      int n_cur = 0;
      for (auto& l : kv.second) {
        uint32_t m = l.mask;
        while (m) {
          const int room = 60 - n_cur;
          uint32_t take = 0;
          for (int t = 0; t < room && m; ++t) { const uint32_t low = m & (0u - m); take |= low; m &= ~low; }
          cur.push_back({l.d, take, l.amp});
          n_cur += __builtin_popcount(take);
          if (n_cur == 60) { by_amp.push_back({kv.first, cur}); cur.clear(); n_cur = 0; }
        }
      }
      if (!cur.empty()) by_amp.push_back({kv.first, cur});
    }
  }
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
  uint32_t max_pairs = 0;
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
        max_pairs = std::max(max_pairs, maxpair);
        // every slot row starts on a 128-byte boundary: row stride = size rounded up to 32 words, class start likewise
        const uint32_t size_pad = (size + 31u) & ~31u;
        ell_ofs[(size_t)c * (k + 1) + p] = (uint32_t)table.size();
        const size_t start = table.size();
        // entries are byte offsets 8 * column (c128 tiles shift once more in the kernel); padding -> xs[size] == 0
        table.resize(start + (size_t)maxpair * size_pad, (uint32_t)(8 * size) | ((uint32_t)(8 * size) << 16));
        for (uint32_t i = 0; i < size; ++i)
          for (size_t sl = 0; sl < lists[i].size(); ++sl) {
            uint32_t& wd = table[start + (sl / 2) * size_pad + i];
            const uint32_t off = 8u * lists[i][sl];
            if (sl & 1) wd = (wd & 0x0000FFFFu) | (off << 16);
            else wd = (wd & 0xFFFF0000u) | off;
          }
      }
      if (table.empty()) table.push_back(0);
      plan->ell[c].upload(table);
      plan->ell_cnt[c].upload(cnt);
      ++c;
    }
  }
  if (max_pairs > 127) return plan;  // u8 slot-pair counts
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
  {
    int dev = 0, l2 = 0;
    ED_CUDA(cudaGetDevice(&dev));
    ED_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    const double tile_bytes = (double)tile_cap * vec_bytes;
    P.far_bit = 0;
    while (P.far_bit < hb && std::ldexp(tile_bytes, P.far_bit + 1) <= (double)l2) ++P.far_bit;
  }
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

  plan->hb = hb;
  plan->wraps = false;
  for (int e = 0; e < P.n_mx; ++e) plan->wraps |= (int)mx_q[e] == hb - 1;

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
  plan->smem_bytes = (size_t)(tile_cap + 1) * vec_bytes + (U1_MAX_HH + U1_MAX_MX + U1_MAX_MQ + 256) * 8 +
                     (U1_MAX_HH + U1_MAX_MX) * 8 + (U1_MAX_MX + U1_MAX_MQ) * 4 + U1_MAX_MS * (4 + 4 + 8 + 8);
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

// contiguous, count-balanced row ranges whose boundaries fall on tile boundaries (so every tile, and therefore every
// neighbour stream, lives in exactly one x segment)
void ed_u1_suggest_rows(ed_oprep* o, int dtype, int world, int rank, int64_t* lo, int64_t* hi) {
  FastU1Plan* plan = get_plan(o, dtype);
  const int64_t dim = o->dim;
  auto snap = [&](int64_t target) -> int64_t {
    if (target <= 0) return 0;
    if (target >= dim) return dim;
    if (!plan->supported) return target;
    auto it = std::lower_bound(plan->h_base.begin(), plan->h_base.end(), (uint64_t)target);
    int64_t up = it == plan->h_base.end() ? dim : (int64_t)*it;
    int64_t down = it == plan->h_base.begin() ? 0 : (int64_t)*(it - 1);
    return (up - target <= target - down) ? up : down;
  };
  *lo = snap(dim / world * rank + std::min<int64_t>(rank, dim % world));
  *hi = snap(dim / world * (rank + 1) + std::min<int64_t>(rank + 1, dim % world));
}

// Wrap-aware shards.  With contiguous single ranges the bond that wraps around the top site (periodic bond of a ring)
// always lands on another rank, as an 8-byte gather that wastes half of every NVLink sector.  Giving every rank the
// SAME range of the remaining high bits in both halves of the basis (top site empty / occupied) keeps that bond local:
// two tile-aligned row ranges per rank, balanced by their combined row count.  Returns 2 ranges, or 1 (the plain
// split) when the operator has no such bond or the world is too large for the tile grid.
int ed_u1_suggest_rows2(ed_oprep* o, int dtype, int world, int rank, int64_t* lo, int64_t* hi) {
  FastU1Plan* plan = get_plan(o, dtype);
  const int hb = plan->hb;
  if (!plan->supported || !plan->wraps || hb < 2 || world < 2 || (1 << (hb - 1)) < 4 * world || getenv("EDCUDA_U1_NOWRAPSHARD")) {
    ed_u1_suggest_rows(o, dtype, world, rank, lo, hi);
    return 1;
  }
  const U1Params& P = plan->P;
  const uint32_t top = 1u << (hb - 1);
  auto tile_rows = [&](uint32_t H) -> uint64_t {
    const int pl = P.n_set - __builtin_popcount(H);
    return (pl < 0 || pl > P.k) ? 0ull : binom_u64(P.k, pl);
  };
  // prefix[h] = rows of the tiles h' < h in the lower half, and the same for the upper half
  std::vector<uint64_t> pre_lo(top + 1, 0), pre_hi(top + 1, 0);
  for (uint32_t h = 0; h < top; ++h) {
    pre_lo[h + 1] = pre_lo[h] + tile_rows(h);
    pre_hi[h + 1] = pre_hi[h] + tile_rows(h | top);
  }
  const uint64_t total = pre_lo[top] + pre_hi[top];
  auto cut = [&](int r) -> uint32_t {   // first h whose combined prefix reaches r/world of all rows
    if (r <= 0) return 0;
    if (r >= world) return top;
    const uint64_t target = total / (uint64_t)world * (uint64_t)r;
    uint32_t a = 0, b = top;
    while (a < b) { const uint32_t m = (a + b) / 2; if (pre_lo[m] + pre_hi[m] < target) a = m + 1; else b = m; }
    return a;
  };
  const uint32_t h0 = cut(rank), h1 = cut(rank + 1);
  lo[0] = (int64_t)pre_lo[h0]; hi[0] = (int64_t)pre_lo[h1];
  lo[1] = (int64_t)(pre_lo[top] + pre_hi[h0]); hi[1] = (int64_t)(pre_lo[top] + pre_hi[h1]);
  return 2;
}

// Rows of x outside the given local ranges that the fast kernel reads for the tiles inside them: whole neighbour tiles,
// merged into ascending disjoint ranges (what the copy engines must bring into the mirror vector before the remote pass).
void ed_u1_remote_rows(ed_oprep* o, int dtype, int n_ranges, const int64_t* lo, const int64_t* hi, std::vector<int64_t>& out_lo,
                       std::vector<int64_t>& out_hi) {
  FastU1Plan* plan = get_plan(o, dtype);
  ED_REQUIRE(plan->supported, ED_ERR_UNSUPPORTED, "remote rows are defined for the U(1) fast-path kernel only");
  const int hb = plan->hb;
  const U1Params& P = plan->P;
  // host copies of the bond lists
  std::vector<uint8_t> hh_p(std::max(P.n_hh, 1)), hh_q(std::max(P.n_hh, 1)), mx_q(std::max(P.n_mx, 1)), ms_q(std::max(P.n_ms, 1));
  plan->hh_p.download(hh_p.data(), hh_p.size()); plan->hh_q.download(hh_q.data(), hh_q.size());
  plan->mx_q.download(mx_q.data(), mx_q.size()); plan->ms_q.download(ms_q.data(), ms_q.size());
  const size_t nt = plan->h_base.size();
  std::vector<uint32_t> tile_H(nt);
  plan->tile_H.download(tile_H.data(), nt);
  std::vector<int32_t> index_of((size_t)1 << hb, -1);
  for (size_t t = 0; t < nt; ++t) index_of[tile_H[t]] = (int32_t)t;
  auto is_local = [&](size_t t) {
    const int64_t b = (int64_t)plan->h_base[t];
    for (int r = 0; r < n_ranges; ++r) if (b >= lo[r] && b < hi[r]) return true;
    return false;
  };
  std::vector<uint8_t> need(nt, 0);
  for (size_t t = 0; t < nt; ++t) {
    if (!is_local(t)) continue;
    const uint32_t H = tile_H[t];
    auto touch = [&](uint32_t H2) { const int32_t j = index_of[H2]; if (j >= 0 && !is_local((size_t)j)) need[j] = 1; };
    for (int b = 0; b < P.n_hh; ++b)
      if (((H >> hh_p[b]) ^ (H >> hh_q[b])) & 1u) touch(H ^ ((1u << hh_p[b]) | (1u << hh_q[b])));
    for (int b = 0; b < P.n_mx; ++b) touch(H ^ (1u << mx_q[b]));
    for (int b = 0; b < P.n_ms; ++b) touch(H ^ (1u << ms_q[b]));
  }
  out_lo.clear(); out_hi.clear();
  for (size_t t = 0; t < nt; ++t) {
    if (!need[t]) continue;
    const int64_t b = (int64_t)plan->h_base[t], e = b + (int64_t)plan->h_size[t];
    if (!out_hi.empty() && out_hi.back() == b) out_hi.back() = e;
    else { out_lo.push_back(b); out_hi.push_back(e); }
  }
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
  static const int far_bit = getenv("EDCUDA_U1_FARBIT") ? atoi(getenv("EDCUDA_U1_FARBIT")) : -1;   // override of the plan's choice
  if (far_bit >= 0) P.far_bit = far_bit;
  P.stream_mode = o->x_seg_ptr.empty() ? 0 : o->exchange_mode;
  P.local_seg_mask = o->local_seg_mask;
  P.mirror = o->mirror;
  if (P.stream_mode == 2) {                 // remote pass: neighbour streams only, added to what the local pass wrote
    ED_REQUIRE(o->mirror != nullptr, ED_ERR_ARGUMENT, "remote pass without a mirror vector (ed_oprep_set_exchange)");
    P.n_ll = 0;
    P.accumulate = 1;
  }
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
        ED_REQUIRE(b == (uint64_t)o->dim || std::binary_search(plan->h_base.begin(), plan->h_base.end(), b), ED_ERR_ARGUMENT,
                   "x segment boundaries must fall on tile boundaries (use ed_oprep_suggest_rows)");
      }
    }
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
  // and 8.26 vs 10.39 ms for 5 (4 CTAs/SM); only R = 7 is instantiated
  if (dtype == ED_F64) launch_u1<double, 7>(plan, P, n_launch, out, partials);
  else launch_u1<c128, 7>(plan, P, n_launch, out, partials);
  if (alpha_dot) ed_reduce_pairs(partials, n_launch, alpha_dot);
}
