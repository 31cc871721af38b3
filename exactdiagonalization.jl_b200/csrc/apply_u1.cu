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
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <random>
#include <thread>

#include "ed_device.cuh"

void ed_reduce_pairs(const double* partials, int n, double* out2);  // apply.cu

#define U1_MAX_CLASSES 8
#define U1_MAX_HH 192
#define U1_MAX_MX 64
#define U1_MAX_MQ 64
#define U1_MAX_MS 16

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
  int partial_first;            // slot of the launch's first tile in the <x,Hx> partials (chunked sharded launches)
  // Sharded (multi-GPU) launches.  stream_mode 0 = everything in one pass; 1 = LOCAL pass (neighbour tiles that live in
  // the halo buffer are skipped); 2 = REMOTE pass (only those, added to y).
  int stream_mode;
  int far_bit;                  // bonds whose upper H bit is >= far_bit read their neighbour tile with evict-first loads:
                                // its re-use distance (2^(bit+1) tiles) exceeds the L2, so the line is a one-shot

  const uint32_t* tile_order;   // optional launch order (whole-basis launches): position -> index into tile_H
  // Where x lives.  Plain: x_local is the full vector, tile H starts at tile_base[H].  Sharded: `dir[H]` is the element
  // offset of tile H inside x_local (this rank's own tiles, concatenated) or, with U1_DIR_HALO set, inside x_halo (a
  // compact buffer holding copies of the peer tiles this rank reads, filled by copy engines over NVLink); -1 = never read.
  const int64_t* dir;
  const void* x_local;
  const void* x_halo;
  int x_bulk;                   // x_local is 16-byte aligned: the tile is staged with one TMA bulk copy (cp.async.bulk) per CTA
};

#define U1_DIR_HALO (1ll << 62)
#define U1_XS_EXTRA 4           // slots of the staged x tile beyond tile_cap: alignment shift, zero padding target, rounding

// the staged-tile bulk copy needs a 16-byte aligned vector (EDCUDA_U1_NOBULK=1: per-thread loads, same results)
static inline int u1_bulk_ok(const void* x) {
  static const bool off = getenv("EDCUDA_U1_NOBULK") != nullptr;
  return (!off && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier (sm_90+: UBLKCP / SYNCS in SASS)
__device__ __forceinline__ uint32_t u1_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void u1_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
  const uint32_t b = u1_smem_addr(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(u1_smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void u1_bulk_wait(unsigned long long* bar) {
  const uint32_t b = u1_smem_addr(bar);
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(b), "r"(0) : "memory");
  } while (!ok);
}

// neighbour tile H2 under the exchange mode: false = this pass does not read it
template <typename VecT>
__device__ __forceinline__ bool u1_neighbour(const U1Params& P, uint32_t H2, const VecT*& ptr) {
  if (!P.dir) {
    ptr = reinterpret_cast<const VecT*>(P.x_local) + P.tile_base[H2];
    return true;
  }
  const int64_t e = P.dir[H2];
  if (e < 0) return false;
  const bool halo = (e & U1_DIR_HALO) != 0;
  if (P.stream_mode == 1 && halo) return false;
  if (P.stream_mode == 2 && !halo) return false;
  ptr = reinterpret_cast<const VecT*>(halo ? P.x_halo : P.x_local) + (e & (U1_DIR_HALO - 1));
  return true;
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
  // host copies of what the shard planner needs (ed_u1_shard_layout works without a device)
  std::vector<uint32_t> h_tile_H;
  std::vector<uint8_t> h_hh_p, h_hh_q, h_mx_q, h_ms_q;
  int n_hh = 0, n_mx = 0, n_ms = 0, k = 0;
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
  VecT* xs_store = reinterpret_cast<VecT*>(smem_raw);                 // tile_cap + U1_XS_EXTRA slots, 16-byte aligned
  double* hh_amp = reinterpret_cast<double*>(xs_store + P.tile_cap + U1_XS_EXTRA);
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
  __shared__ __align__(8) unsigned long long s_bar;

  const int tid = threadIdx.x;
  const uint32_t H = P.tile_H[P.tile_order ? P.tile_order[blockIdx.x] : P.tile_first + blockIdx.x];
  const int p_low = P.n_set - __popc(H);
  const uint32_t lofs = P.lowofs[p_low];
  const uint32_t size = P.lowofs[p_low + 1] - lofs;
  // position of the tile in x_local (and, minus row_lo, in y): its global rank, or its slot in this rank's shard
  const int64_t base = P.dir ? P.dir[H] : (int64_t)P.tile_base[H];
  const int k = P.k;
  // x tile in shared memory: xs[0 .. size) = x[base .. base + size), xs[size] = 0 (the ELL padding target).  With a 16-byte
  // aligned vector the tile comes in through ONE bulk copy issued by thread 0 while the other warps build the bond lists:
  // from the 16-byte boundary at or below the tile (Float64: xs is shifted by one slot for odd `base`), whole 16-byte units
  // only -- a trailing half unit and the zero slot are written by thread 0 itself, so the copy never touches them and never
  // reads past the tile.
  const uint32_t mis = (sizeof(VecT) == 8 && P.x_bulk) ? (uint32_t)(base & 1) : 0u;
  VecT* xs = xs_store + mis;
  const uint32_t bulk_elems = P.x_bulk ? (sizeof(VecT) == 8 ? ((mis + size) & ~1u) : size) : 0u;   // in VecT units from xs_store
  const bool bulk = bulk_elems > 0 && P.stream_mode != 2;
  if (bulk && tid == 0) {
    const VecT* src = reinterpret_cast<const VecT*>(P.x_local) + (base - mis);
    u1_bulk_load(xs_store, src, bulk_elems * (uint32_t)sizeof(VecT), &s_bar);
    if (bulk_elems < mis + size) xs_store[bulk_elems] = ldg_val(src + bulk_elems);   // odd tail
    xs[size] = vzero((VecT*)nullptr);
  }

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
      if (fire) fire = u1_neighbour<VecT>(P, H2, xn);
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
      if (on) { q = P.mx_q[b]; on = u1_neighbour<VecT>(P, H ^ (1u << q), xn); }
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
      if (on) { q = P.ms_q[b]; on = u1_neighbour<VecT>(P, H ^ (1u << q), xn); }
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
      if (dot_partials && tid == 0) { dot_partials[2 * (P.partial_first + blockIdx.x)] = 0.0; dot_partials[2 * (P.partial_first + blockIdx.x) + 1] = 0.0; }
      return;
    }
  }
  if (!bulk) {
    const VecT* xo = reinterpret_cast<const VecT*>(P.x_local) + base;
    for (uint32_t i = tid; i < size; i += THREADS) xs[i] = ldg_val(xo + i);
    if (tid == 0) xs[size] = vzero((VecT*)nullptr);
  }
  __syncthreads();
  if (bulk) u1_bulk_wait(&s_bar);

  U1Tile<VecT> T;
  T.xs = xs;
  T.hh_amp = hh_amp; T.hh_ptr = hh_ptr; T.n_hh = s_counts[0];
  T.mx_amp = mx_amp; T.mx_ptr = mx_ptr; T.mx_toff = mx_toff; T.n_mx = s_counts[2];
  T.ms_amp = ms_amp; T.ms_ptr = ms_ptr; T.ms_lo = ms_lo; T.ms_len = ms_len; T.n_ms = s_counts[3];
  T.mq_coef = mq_coef; T.mq_bit = mq_bit; T.n_mq = s_counts[1];
  T.s_dval = s_dval;
  T.lofs = lofs; T.gofs = P.grpofs[p_low]; T.size = size; T.p_low = p_low; T.base = base;
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
      dot_partials[2 * (P.partial_first + blockIdx.x)] = a;
      dot_partials[2 * (P.partial_first + blockIdx.x) + 1] = c;
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

// host_only: everything except the device uploads (shard planning and its CPU tests need no GPU)
static std::shared_ptr<FastU1Plan> build_plan(const ed_operator& op, int n_bits, int n_set, int vec_bytes, bool host_only) {
  auto plan = std::make_shared<FastU1Plan>();
  plan->vec_bytes = vec_bytes;
  if (n_bits < 1 || n_bits > 42 || n_set < 0 || n_set > n_bits) return plan;
  Lowered L = lower_operator(op, n_bits, n_set);
  if (!L.ok) return plan;
  const int k = choose_k(n_bits, vec_bytes);
  const int hb = n_bits - k;
  if (hb > 26) return plan;  // tile tables of 2^hb entries
  U1Params& P = plan->P;
  memset(&P, 0, sizeof(P));
  P.n_bits = n_bits; P.n_set = n_set; P.k = k;
  plan->idx32 = binom_u64(n_bits, n_set) < (1ull << 32);
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
      if (!host_only) {
        plan->ell[c].upload(table);
        plan->ell_cnt[c].upload(cnt);
      }
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
  plan->k = k; plan->n_hh = P.n_hh; plan->n_mx = P.n_mx; plan->n_ms = P.n_ms;
  plan->h_tile_H = tile_H;
  plan->h_hh_p = hh_p; plan->h_hh_q = hh_q; plan->h_mx_q = mx_q; plan->h_ms_q = ms_q;
  plan->hb = hb;
  plan->wraps = false;
  for (int e = 0; e < P.n_mx; ++e) plan->wraps |= (int)mx_q[e] == hb - 1;
  if (host_only) {
    plan->supported = true;
    return plan;
  }
  {
    int dev = 0, l2 = 0;
    ED_CUDA(cudaGetDevice(&dev));
    ED_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    const double tile_bytes = (double)tile_cap * vec_bytes;
    P.far_bit = 0;
    while (P.far_bit < hb && std::ldexp(tile_bytes, P.far_bit + 1) <= (double)l2) ++P.far_bit;
  }
  plan->order_on = false;

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
  plan->smem_bytes = (size_t)(tile_cap + U1_XS_EXTRA) * vec_bytes + (U1_MAX_HH + U1_MAX_MX + U1_MAX_MQ + 256) * 8 +
                     (U1_MAX_HH + U1_MAX_MX) * 8 + (U1_MAX_MX + U1_MAX_MQ) * 4 + U1_MAX_MS * (4 + 4 + 8 + 8);
  plan->smem_bytes = (plan->smem_bytes + 15) & ~(size_t)15;
  plan->supported = true;
  return plan;
}

static void u1_apply_env_order(FastU1Plan* plan);

static FastU1Plan* get_plan(ed_oprep* o, int dtype) {
  // one plan per vector type (the tile size depends on the element size), cached on the representation
  std::shared_ptr<FastU1Plan>& slot = dtype == ED_C128 ? o->u1plan_c : o->u1plan;
  if (!slot) {
    ed_basis* b = o->basis;
    if (b->kind != ED_BASIS_COMBINADIC || b->dim <= 0) slot = std::make_shared<FastU1Plan>();
    else slot = build_plan(o->op, b->space.bits, b->n_set, dtype == ED_C128 ? 16 : 8, false);
    if (slot->supported) u1_apply_env_order(slot.get());
  }
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
  static size_t configured[64] = {0};   // per device: the opt-in shared-memory size is a per-device function attribute
  int dev = 0;
  ED_CUDA(cudaGetDevice(&dev));
  if (plan->smem_bytes > 48 * 1024 && configured[dev & 63] < plan->smem_bytes) {
    ED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes));
    configured[dev & 63] = plan->smem_bytes;
  }
  ED_LAUNCH(kern, n_launch, U1_THREADS, plan->smem_bytes, P, reinterpret_cast<VecT*>(out), partials);
}

FastU1Plan* ed_u1_plan(ed_oprep* o, int dtype) {
  FastU1Plan* plan = get_plan(o, dtype);
  return plan->supported ? plan : nullptr;
}

std::shared_ptr<FastU1Plan> ed_u1_host_plan(const ed_operator& op, int n_bits, int n_set, int dtype) {
  return build_plan(op, n_bits, n_set, dtype == ED_C128 ? 16 : 8, true);
}

// contiguous, count-balanced row ranges whose boundaries fall on tile boundaries
static int64_t u1_snap(const FastU1Plan* plan, int64_t dim, int64_t target) {
  if (target <= 0) return 0;
  if (target >= dim) return dim;
  if (!plan->supported) return target;
  auto it = std::lower_bound(plan->h_base.begin(), plan->h_base.end(), (uint64_t)target);
  int64_t up = it == plan->h_base.end() ? dim : (int64_t)*it;
  int64_t down = it == plan->h_base.begin() ? 0 : (int64_t)*(it - 1);
  return (up - target <= target - down) ? up : down;
}

void ed_u1_suggest_rows(ed_oprep* o, int dtype, int world, int rank, int64_t* lo, int64_t* hi) {
  FastU1Plan* plan = get_plan(o, dtype);
  const int64_t dim = o->dim;
  *lo = u1_snap(plan, dim, dim / world * rank + std::min<int64_t>(rank, dim % world));
  *hi = u1_snap(plan, dim, dim / world * (rank + 1) + std::min<int64_t>(rank + 1, dim % world));
}

// ------------------------------------------------------------------ shard layout (host only)
// Which rank owns which tile, and what moves between ranks per matvec.
//
// A rank reads, besides its own tiles, every tile H' = bond(H) of its tiles H (about 10 per tile at L=32).  Tiles are
// assigned by cutting an ORDERING of all tiles into `world` pieces of equal row count, so the volume a rank must fetch
// ("halo") is the boundary of its piece.  Candidate orderings:
//   * ascending H -- contiguous row ranges, the reference's splitrange (util.jl:102-121) snapped to tiles; when a bond
//     wraps around the most significant site, with the top bit as the LEAST significant key (every rank owns the same
//     range of the remaining bits in both halves of the basis, which keeps the wrapping bond local);
//   * popcount keys: the high bits are split into one or two contiguous blocks and tiles are ordered by the number of
//     particles in each block (boustrophedon), then by the block bits.  A bond inside a block keeps the popcounts, so
//     only the few bonds at the block boundaries leave a class: at L=32 the largest halo drops from 2.3e8 to 1.0e8 rows
//     at 8 ranks and from 2.3e8 to 0.8e8 rows at 2 ranks.
// The planner evaluates every candidate on the actual bond lists (any lattice) and keeps the one with the smallest
// maximum halo (policy 0).  policy 1 = plain ascending ranges; 2 = wrap-aware ranges; 3 = best popcount key.
namespace {

struct U1Global {
  int world = 1, n_chunks = 1;
  std::string key_name;
  std::vector<int> owner;                       // [tile]
  std::vector<int64_t> off;                     // [tile] offset in its owner's local vectors (ascending-H storage)
  std::vector<int64_t> rows_of_rank;
  std::vector<std::vector<uint32_t>> launch;    // [rank] tile indices in launch order (the key's order)
  std::vector<std::vector<int>> chunk_first;    // [rank][n_chunks + 1]
  // halo of rank r: tile indices in halo order = (chunk, owner, tile) ascending
  std::vector<std::vector<uint32_t>> halo_tiles;
  std::vector<std::vector<int>> halo_chunk;     // chunk of every halo tile
  std::vector<int64_t> exch;                    // [(p * world + r) * n_chunks + c] rows sent p -> r for chunk c
  int64_t send_off(int p, int r, int c) const {  // offset of that segment in p's send buffer: receiver-major, then chunk
    int64_t o = 0;
    for (int r2 = 0; r2 < r; ++r2) for (int c2 = 0; c2 < n_chunks; ++c2) o += exch[((size_t)p * world + r2) * n_chunks + c2];
    for (int c2 = 0; c2 < c; ++c2) o += exch[((size_t)p * world + r) * n_chunks + c2];
    return o;
  }
};

struct TileGraph {
  size_t nt = 0;
  int hb = 0;
  std::vector<int32_t> index_of;                // [2^hb]
  std::vector<int32_t> nbr;                     // [nt * deg] neighbour tile index or -1
  int deg = 0;
};

TileGraph u1_tile_graph(const FastU1Plan* plan) {
  TileGraph G;
  G.nt = plan->h_tile_H.size();
  G.hb = plan->hb;
  G.index_of.assign((size_t)1 << G.hb, -1);
  for (size_t t = 0; t < G.nt; ++t) G.index_of[plan->h_tile_H[t]] = (int32_t)t;
  G.deg = plan->n_hh + plan->n_mx + plan->n_ms;
  G.nbr.assign(G.nt * (size_t)std::max(G.deg, 1), -1);
  for (size_t t = 0; t < G.nt; ++t) {
    const uint32_t H = plan->h_tile_H[t];
    int32_t* row = G.nbr.data() + t * G.deg;
    int d = 0;
    for (int b = 0; b < plan->n_hh; ++b, ++d)
      if (((H >> plan->h_hh_p[b]) ^ (H >> plan->h_hh_q[b])) & 1u) row[d] = G.index_of[H ^ ((1u << plan->h_hh_p[b]) | (1u << plan->h_hh_q[b]))];
    for (int b = 0; b < plan->n_mx; ++b, ++d) row[d] = G.index_of[H ^ (1u << plan->h_mx_q[b])];
    for (int b = 0; b < plan->n_ms; ++b, ++d) row[d] = G.index_of[H ^ (1u << plan->h_ms_q[b])];
  }
  return G;
}

// cut an ordering of the tiles into `world` pieces of (nearly) equal row count
void u1_cut(const FastU1Plan* plan, const std::vector<uint32_t>& order, int world, std::vector<int>& owner) {
  uint64_t total = 0;
  for (uint64_t v : plan->h_size) total += v;
  owner.assign(order.size(), 0);
  uint64_t acc = 0;
  for (uint32_t t : order) {
    // piece of the tile's midpoint: tiles never straddle, the pieces differ by at most one tile
    const uint64_t mid = acc + plan->h_size[t] / 2;
    owner[t] = (int)std::min<uint64_t>((unsigned __int128)mid * (uint64_t)world / std::max<uint64_t>(total, 1), (uint64_t)world - 1);
    acc += plan->h_size[t];
  }
}

// largest halo (rows of distinct peer tiles read) over the ranks
uint64_t u1_max_halo(const FastU1Plan* plan, const TileGraph& G, const std::vector<int>& owner, int world) {
  std::vector<std::vector<uint32_t>> by_rank(world);
  for (size_t t = 0; t < G.nt; ++t) by_rank[owner[t]].push_back((uint32_t)t);
  std::vector<int32_t> stamp(G.nt, -1);         // last rank that counted the tile: ranks are swept one after the other
  uint64_t worst = 0;
  for (int r = 0; r < world; ++r) {
    uint64_t halo = 0;
    for (uint32_t t : by_rank[r]) {
      const int32_t* row = G.nbr.data() + (size_t)t * G.deg;
      for (int d = 0; d < G.deg; ++d) {
        const int32_t j = row[d];
        if (j < 0 || owner[j] == r || stamp[j] == r) continue;
        stamp[j] = r;
        halo += plan->h_size[j];
      }
    }
    worst = std::max(worst, halo);
  }
  return worst;
}

struct KeySpec { std::string name; int kind; int a0, s, b1; };   // kind 0 plain, 1 wrap ranges, 2 popcount blocks [a0,s)[s,b1) (s<0: one block)

void u1_order_by_key(const FastU1Plan* plan, const KeySpec& K, std::vector<uint32_t>& order) {
  const size_t nt = plan->h_tile_H.size();
  const int hb = plan->hb;
  std::vector<uint64_t> key(nt);
  auto maskof = [](int a, int b) -> uint32_t { return b <= a ? 0u : (uint32_t)((((uint64_t)1 << b) - 1) & ~(((uint64_t)1 << a) - 1)); };
  for (size_t t = 0; t < nt; ++t) {
    const uint32_t H = plan->h_tile_H[t];
    if (K.kind == 0) key[t] = H;
    else if (K.kind == 1) key[t] = ((uint64_t)(H & ((1u << (hb - 1)) - 1u)) << 1) | (H >> (hb - 1));
    else {
      const uint32_t m_lo = K.s < 0 ? 0u : maskof(K.a0, K.s), m_hi = K.s < 0 ? maskof(K.a0, K.b1) : maskof(K.s, K.b1);
      const uint32_t q_hi = (uint32_t)__builtin_popcount(H & m_hi), q_lo = (uint32_t)__builtin_popcount(H & m_lo);
      const uint32_t w_lo = (uint32_t)__builtin_popcount(m_lo);
      const uint32_t snake = (q_hi & 1u) ? w_lo - q_lo : q_lo;         // boustrophedon through the lower block
      key[t] = ((uint64_t)q_hi << 56) | ((uint64_t)snake << 48) | ((uint64_t)(H & (m_lo | m_hi)) << 20) | (uint64_t)(H & ~(m_lo | m_hi) & 0xFFFFFu);
    }
  }
  order.resize(nt);
  for (size_t t = 0; t < nt; ++t) order[t] = (uint32_t)t;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return key[x] < key[y]; });
}

std::shared_ptr<U1Global> u1_global_layout(const FastU1Plan* plan, int world, int n_chunks, int policy) {
  auto Gp = std::make_shared<U1Global>();
  U1Global& G = *Gp;
  const size_t nt = plan->h_tile_H.size();
  const int hb = plan->hb;
  const TileGraph TG = u1_tile_graph(plan);
  G.world = world;
  // ---- choose the ordering
  std::vector<KeySpec> cands;
  const bool wrap_ok = plan->wraps && hb >= 2 && (1 << (hb - 1)) >= 4 * world;
  if (policy == 0 || policy == 1) cands.push_back({"ascending ranges", 0, 0, -1, 0});
  if ((policy == 0 || policy == 2) && wrap_ok) cands.push_back({"wrap-aware ranges", 1, 0, -1, 0});
  if (policy == 2 && !wrap_ok) cands.push_back({"ascending ranges", 0, 0, -1, 0});
  if ((policy == 0 || policy == 3) && world > 1 && hb >= 4) {
    for (int a0 = 0; a0 <= std::min(2, hb - 3); ++a0)
      for (int b1 = std::max(a0 + 3, hb - 2); b1 <= hb; ++b1) {
        cands.push_back({"popcount[" + std::to_string(a0) + "," + std::to_string(b1) + ")", 2, a0, -1, b1});
        for (int s = a0 + 2; s <= b1 - 2; ++s)
          cands.push_back({"popcount[" + std::to_string(a0) + "," + std::to_string(s) + ")[" + std::to_string(s) + "," + std::to_string(b1) + ")", 2, a0, s, b1});
      }
  }
  if (cands.empty()) cands.push_back({"ascending ranges", 0, 0, -1, 0});
  std::vector<uint32_t> order, best_order;
  std::vector<int> owner;
  uint64_t best = ~0ull;
  std::vector<std::pair<uint64_t, size_t>> ranked;           // (largest halo of the plain cut, candidate)
  for (size_t ci = 0; ci < cands.size(); ++ci) {
    u1_order_by_key(plan, cands[ci], order);
    u1_cut(plan, order, world, owner);
    const uint64_t h = world > 1 ? u1_max_halo(plan, TG, owner, world) : 0;
    ranked.push_back({h, ci});
    if (h < best) { best = h; best_order = order; G.owner = owner; G.key_name = cands[ci].name; }
    if (world == 1) break;
  }
  // ---- local refinement of the cut (policy 0 only): move single tiles between ranks while that lowers
  //      max over the two ranks of max(rows, rho * halo rows)   [or keeps it and lowers the halo sum],
  // rho = 1.5 = measured cost of one halo row (8 B at the ~400 GB/s the copy engines reach over NVLink in an 8-rank
  // exchange) over one kernel row (7.86 ms / 6.0e8).  Where the transfer dominates (4 and 8 ranks at L=32) this trades a
  // few per cent of row balance for less halo; where the kernel dominates (2 ranks) the row term keeps the balance.
  // The descent stops in a local optimum that depends on where it starts, so it is run from the EDCUDA_SHARD_STARTS (16)
  // best plain cuts, on host threads, and the partition with the smallest max over ranks of max(rows, rho * halo) wins
  // (L=32, 8 ranks: largest halo 0.99e8 plain -> 0.62e8 refined from the best plain cut -> 0.55e8 from the best of 16).
  // Deterministic (fixed-seed shuffles, ties to the better plain cut): every rank computes the same partition.
  // EDCUDA_SHARD_REFINE=0 switches it off.
  if (world > 1 && policy == 0 && !(getenv("EDCUDA_SHARD_REFINE") && atoi(getenv("EDCUDA_SHARD_REFINE")) == 0)) {
    const double rho = 1.5;
    std::stable_sort(ranked.begin(), ranked.end());
    const int want = getenv("EDCUDA_SHARD_STARTS") ? std::max(1, atoi(getenv("EDCUDA_SHARD_STARTS"))) : 16;
    const size_t n_starts = std::min<size_t>((size_t)want, ranked.size());
    struct Start { std::vector<uint32_t> order; std::vector<int> owner; double objective = 0, halo_max = 0, halo_sum = 0; };
    std::vector<Start> starts(n_starts);
    auto refine = [&](Start& S, const KeySpec& K) {
      u1_order_by_key(plan, K, S.order);
      u1_cut(plan, S.order, world, S.owner);
      std::vector<int>& owner_r = S.owner;
      std::vector<std::vector<int32_t>> cnt(world, std::vector<int32_t>(nt, 0));     // cnt[r][n]: tiles of r that read n
      std::vector<double> rows(world, 0.0), halo(world, 0.0);
      for (size_t t = 0; t < nt; ++t) {
        rows[owner_r[t]] += (double)plan->h_size[t];
        const int32_t* row = TG.nbr.data() + t * TG.deg;
        for (int d = 0; d < TG.deg; ++d) if (row[d] >= 0) ++cnt[owner_r[t]][row[d]];
      }
      for (int r = 0; r < world; ++r)
        for (size_t n = 0; n < nt; ++n) if (cnt[r][n] > 0 && owner_r[n] != r) halo[r] += (double)plan->h_size[n];
      auto obj = [&](double rw, double hl) { return std::max(rw, rho * hl); };
      double total_rows = 0;
      for (int r = 0; r < world; ++r) total_rows += rows[r];
      const double row_cap = 1.25 * total_rows / world;        // no shard grows beyond +25 % whatever the halo says
      std::vector<uint32_t> perm(nt);
      for (int pass = 0; pass < 8; ++pass) {
        for (size_t t = 0; t < nt; ++t) perm[t] = (uint32_t)t;
        std::mt19937 rng(12345u + (unsigned)pass);
        for (size_t i = nt; i > 1; --i) std::swap(perm[i - 1], perm[rng() % i]);      // own Fisher-Yates: same on every rank
        size_t moved = 0;
        for (uint32_t t : perm) {
          const int a = owner_r[t];
          const double sz = (double)plan->h_size[t];
          const int32_t* row = TG.nbr.data() + (size_t)t * TG.deg;
          int best_b = -1;
          double best_new = 0, best_sum = 0, best_da = 0, best_db = 0;
          for (int d0 = 0; d0 < TG.deg; ++d0) {
            if (row[d0] < 0) continue;
            const int b = owner_r[row[d0]];
            if (b == a || rows[b] + sz > row_cap) continue;
            bool seen_b = false;
            for (int d1 = 0; d1 < d0 && !seen_b; ++d1) seen_b = row[d1] >= 0 && owner_r[row[d1]] == b;
            if (seen_b) continue;
            double da = 0, db = 0;
            for (int d = 0; d < TG.deg; ++d) {
              const int32_t n = row[d];
              if (n < 0) continue;
              if (cnt[a][n] == 1 && owner_r[n] != a) da -= (double)plan->h_size[n];
              if (cnt[b][n] == 0 && owner_r[n] != b) db += (double)plan->h_size[n];
            }
            if (cnt[b][t] > 0) db -= sz;
            if (cnt[a][t] > 0) da += sz;
            const double old_v = std::max(obj(rows[a], halo[a]), obj(rows[b], halo[b]));
            const double new_v = std::max(obj(rows[a] - sz, halo[a] + da), obj(rows[b] + sz, halo[b] + db));
            const bool better = new_v < old_v || (new_v == old_v && da + db < 0);
            if (better && (best_b < 0 || new_v < best_new || (new_v == best_new && da + db < best_sum))) {
              best_b = b; best_new = new_v; best_sum = da + db; best_da = da; best_db = db;
            }
          }
          if (best_b < 0) continue;
          for (int d = 0; d < TG.deg; ++d) if (row[d] >= 0) { --cnt[a][row[d]]; ++cnt[best_b][row[d]]; }
          owner_r[t] = best_b;
          rows[a] -= sz; rows[best_b] += sz;
          halo[a] += best_da; halo[best_b] += best_db;
          ++moved;
        }
        if (moved == 0) break;
      }
      S.objective = S.halo_max = S.halo_sum = 0;
      for (int r = 0; r < world; ++r) {
        S.objective = std::max(S.objective, obj(rows[r], halo[r]));
        S.halo_max = std::max(S.halo_max, halo[r]);
        S.halo_sum += halo[r];
      }
    };
    {
      std::atomic<size_t> next{0};
      auto worker = [&]() {
        for (size_t i = next.fetch_add(1); i < n_starts; i = next.fetch_add(1)) refine(starts[i], cands[ranked[i].second]);
      };
      const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
      const size_t n_threads = std::min<size_t>(n_starts, std::min<unsigned>(hw, 8u));
      std::vector<std::thread> pool;
      for (size_t i = 1; i < n_threads; ++i) pool.emplace_back(worker);
      worker();
      for (auto& th : pool) th.join();
    }
    size_t win = 0;
    for (size_t i = 1; i < n_starts; ++i) {       // smallest objective, then largest halo, then total halo; ties stay with the better plain cut
      const Start &A = starts[i], &B = starts[win];
      if (A.objective < B.objective || (A.objective == B.objective && (A.halo_max < B.halo_max || (A.halo_max == B.halo_max && A.halo_sum < B.halo_sum)))) win = i;
    }
    G.owner = std::move(starts[win].owner);
    best_order = std::move(starts[win].order);
    G.key_name = cands[ranked[win].second].name + " + local refinement";
  }
  // ---- storage: every rank keeps its tiles in ascending H (fewest global row ranges); launch order = the key's order
  G.rows_of_rank.assign(world, 0);
  G.off.assign(nt, 0);
  for (size_t t = 0; t < nt; ++t) { G.off[t] = G.rows_of_rank[G.owner[t]]; G.rows_of_rank[G.owner[t]] += (int64_t)plan->h_size[t]; }
  // launch order of a rank: first its INTERIOR tiles (every tile they read is its own: they need no transfer and run
  // while the copy engines fetch the halo), then the boundary tiles; both in the key's order
  G.launch.assign(world, {});
  std::vector<std::vector<uint32_t>> boundary(world);
  for (uint32_t t : best_order) {
    const int r = G.owner[t];
    bool interior = true;
    const int32_t* row = TG.nbr.data() + (size_t)t * TG.deg;
    for (int d = 0; d < TG.deg && interior; ++d) interior = row[d] < 0 || G.owner[row[d]] == r;
    (interior || world == 1 ? G.launch[r] : boundary[r]).push_back(t);
  }
  size_t most = 1;
  for (auto& l : boundary) most = std::max(most, l.size());
  // chunk 0 = the interior tiles; chunks 1 .. n_chunks-1 = the boundary tiles in equal row counts
  n_chunks = world > 1 ? (int)std::max<size_t>(2, std::min<size_t>((size_t)n_chunks, most + 1)) : 1;
  G.n_chunks = n_chunks;
  G.chunk_first.assign(world, std::vector<int>(n_chunks + 1, 0));
  G.halo_tiles.assign(world, {});
  G.halo_chunk.assign(world, {});
  G.exch.assign((size_t)world * world * n_chunks, 0);
  for (int r = 0; r < world; ++r) {
    auto& L = G.launch[r];
    auto& cf = G.chunk_first[r];
    const int n_int = (int)L.size();
    int64_t rows_b = 0;
    for (uint32_t t : boundary[r]) rows_b += (int64_t)plan->h_size[t];
    L.insert(L.end(), boundary[r].begin(), boundary[r].end());
    for (int c = 0; c <= n_chunks; ++c) cf[c] = (int)L.size();
    cf[0] = 0;
    if (world > 1) {
      cf[1] = n_int;
      int c = 2;
      int64_t acc = 0;
      const int nb = n_chunks - 1;
      for (size_t i = (size_t)n_int; i < L.size() && c < n_chunks; ++i) {
        acc += (int64_t)plan->h_size[L[i]];
        if (acc >= rows_b / nb * (c - 1)) cf[c++] = (int)i + 1;
      }
    }
    std::vector<uint8_t> seen(nt, 0);
    for (int ch = 0; ch < n_chunks; ++ch) {
      std::vector<uint32_t> fresh;
      for (int i = cf[ch]; i < cf[ch + 1]; ++i) {
        const int32_t* row = TG.nbr.data() + (size_t)L[i] * TG.deg;
        for (int d = 0; d < TG.deg; ++d) {
          const int32_t j = row[d];
          if (j >= 0 && G.owner[j] != r && !seen[j]) { seen[j] = 1; fresh.push_back((uint32_t)j); }
        }
      }
      // peers in ROTATED order (r+1, r+2, ... mod world): at any moment every source is read by one reader -- with all
      // ranks walking their peers in the same ascending order the seven readers of an 8-rank run would queue up on one
      // source's NVLink egress at a time
      auto turn = [&](uint32_t t) { return (G.owner[t] - r - 1 + world) % world; };
      std::sort(fresh.begin(), fresh.end(), [&](uint32_t x, uint32_t y) { return turn(x) != turn(y) ? turn(x) < turn(y) : x < y; });
      for (uint32_t t : fresh) {
        G.halo_tiles[r].push_back(t);
        G.halo_chunk[r].push_back(ch);
        G.exch[((size_t)G.owner[t] * world + r) * n_chunks + ch] += (int64_t)plan->h_size[t];
      }
    }
  }
  return Gp;
}

}  // namespace

// Everything rank `rank` needs to run its share of the tiled matvec: its tiles (launch order), their offsets in the local
// vectors (ascending-H storage), the directory that resolves any tile it reads to "local vector + offset" or "halo
// buffer + offset", what it PACKS for its peers (the tiles they read, gathered into one send buffer, receiver-major then
// by the receiver's launch chunk) and what it PULLS (one contiguous piece per peer and chunk).  The chunks are launched
// in order, each waiting only for its own pieces: the transfer of chunk c+1 overlaps the kernel of chunk c.
void ed_u1_shard_layout(const FastU1Plan* plan, int world, int rank, int n_chunks, int policy, U1ShardLayout* out) {
  ED_REQUIRE(plan && plan->supported, ED_ERR_UNSUPPORTED, "the shard layout is defined for the tiled U(1) kernel only");
  ED_REQUIRE(world >= 1 && rank >= 0 && rank < world && n_chunks >= 1, ED_ERR_ARGUMENT, "bad rank / world / chunks");
  // the global layout is the same for every rank: cached per (plan, world, chunks, policy) so that a process driving
  // several GPUs computes it once
  // keyed by CONTENT (everything u1_global_layout reads: tiles, bond lists, world, chunks, policy, the refinement switch),
  // never by the plan's address -- a freed plan's address can come back with another operator behind it
  uint64_t sig = 1469598103934665603ull;
  auto mix = [&](const void* data, size_t bytes) {
    const unsigned char* b = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < bytes; ++i) { sig ^= b[i]; sig *= 1099511628211ull; }
  };
  {
    const char* re = getenv("EDCUDA_SHARD_REFINE");
    const int64_t head[10] = {world, n_chunks, policy, plan->hb, plan->k, plan->wraps ? 1 : 0, plan->n_hh, plan->n_mx, plan->n_ms,
                              (re && atoi(re) == 0) ? 0 : 1};
    mix(head, sizeof(head));
    mix(plan->h_tile_H.data(), plan->h_tile_H.size() * sizeof(uint32_t));
    mix(plan->h_size.data(), plan->h_size.size() * sizeof(uint64_t));
    mix(plan->h_hh_p.data(), plan->h_hh_p.size());
    mix(plan->h_hh_q.data(), plan->h_hh_q.size());
    mix(plan->h_mx_q.data(), plan->h_mx_q.size());
    mix(plan->h_ms_q.data(), plan->h_ms_q.size());
  }
  static thread_local struct { uint64_t sig = 0; size_t nt = 0; std::shared_ptr<U1Global> G; } cache;
  if (!(cache.G && cache.sig == sig && cache.nt == plan->h_tile_H.size())) {
    cache.G = u1_global_layout(plan, world, n_chunks, policy);
    cache.sig = sig; cache.nt = plan->h_tile_H.size();
  }
  const U1Global& G = *cache.G;
  const size_t nt = plan->h_tile_H.size();
  const int hb = plan->hb;
  U1ShardLayout& S = *out;
  S = U1ShardLayout();
  S.world = world; S.rank = rank; S.n_chunks = G.n_chunks;
  S.key_name = G.key_name;
  S.rows_of_rank = G.rows_of_rank;
  S.n_local = G.rows_of_rank[rank];
  S.dir.assign((size_t)1 << hb, -1);
  for (size_t t = 0; t < nt; ++t)
    if (G.owner[t] == rank) {
      S.dir[plan->h_tile_H[t]] = G.off[t];
      const int64_t b = (int64_t)plan->h_base[t], e = b + (int64_t)plan->h_size[t];
      if (!S.range_hi.empty() && S.range_hi.back() == b) S.range_hi.back() = e;
      else { S.range_lo.push_back(b); S.range_hi.push_back(e); }
    }
  for (uint32_t t : G.launch[rank]) { S.tile_H.push_back(plan->h_tile_H[t]); S.tile_off.push_back(G.off[t]); }
  S.chunk_first = G.chunk_first[rank];
  // halo + pulls
  int64_t halo = 0;
  const auto& ht = G.halo_tiles[rank];
  const auto& hc = G.halo_chunk[rank];
  for (size_t i = 0; i < ht.size(); ++i) {
    const uint32_t t = ht[i];
    const int p = G.owner[t];
    const int64_t len = (int64_t)plan->h_size[t];
    S.dir[plan->h_tile_H[t]] = halo | U1_DIR_HALO;
    if (!S.pulls.empty() && S.pulls.back().peer == p && S.pulls.back().chunk == hc[i]) S.pulls.back().len += len;
    else S.pulls.push_back({p, hc[i], G.send_off(p, rank, hc[i]), halo, len});
    halo += len;
  }
  S.n_halo = halo;
  S.chunk_halo_rows.assign(G.n_chunks, 0);
  for (size_t i = 0; i < ht.size(); ++i) S.chunk_halo_rows[hc[i]] += (int64_t)plan->h_size[ht[i]];
  // pushes: the receivers' halo tiles owned by this rank with their offsets in the RECEIVER's halo, earliest chunk first
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    const auto& rt = G.halo_tiles[r];
    int64_t at = 0;
    for (size_t i = 0; i < rt.size(); ++i) {
      const uint32_t t = rt[i];
      const int64_t len = (int64_t)plan->h_size[t];
      if (G.owner[t] == rank) {
        const int ch = G.halo_chunk[r][i];
        if (!S.pushes.empty() && S.pushes.back().recv == r && S.pushes.back().chunk == ch && S.pushes.back().src_off + S.pushes.back().len == G.off[t] &&
            S.pushes.back().dst_off + S.pushes.back().len == at)
          S.pushes.back().len += len;
        else S.pushes.push_back({r, ch, G.off[t], at, len});
      }
      at += len;
    }
  }
  // the same as contiguous pieces of the packed send buffer (copy-engine push): one per (receiver, chunk)
  for (const U1Push& p : S.pushes) {
    if (!S.piece_pushes.empty() && S.piece_pushes.back().recv == p.recv && S.piece_pushes.back().chunk == p.chunk) S.piece_pushes.back().len += p.len;
    else S.piece_pushes.push_back({p.recv, p.chunk, G.send_off(rank, p.recv, p.chunk), p.dst_off, p.len});
  }
  // by chunk, receivers in rotated order (rank+1, rank+2, ...): no two owners write the same receiver at the same time
  auto push_less = [&](const U1Push& a, const U1Push& b) {
    if (a.chunk != b.chunk) return a.chunk < b.chunk;
    return (a.recv - rank - 1 + world) % world < (b.recv - rank - 1 + world) % world;
  };
  std::stable_sort(S.pushes.begin(), S.pushes.end(), push_less);
  std::stable_sort(S.piece_pushes.begin(), S.piece_pushes.end(), push_less);
  // packs: for every receiver (ascending) and chunk, its halo tiles owned by this rank, in its halo order
  int64_t send = 0;
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    const auto& rt = G.halo_tiles[r];
    for (size_t i = 0; i < rt.size(); ++i) {
      const uint32_t t = rt[i];
      if (G.owner[t] != rank) continue;
      const int64_t len = (int64_t)plan->h_size[t];
      if (!S.packs.empty() && S.packs.back().src_off + S.packs.back().len == G.off[t] && S.packs.back().dst_off + S.packs.back().len == send)
        S.packs.back().len += len;
      else S.packs.push_back({G.off[t], send, len});
      send += len;
    }
  }
  S.n_send = send;
}

// Launch order of whole-basis launches (EDCUDA_U1_ORDER): "class" = tiles grouped by popcount(H) -- the tiles running at
// the same time then share one ELL table and their high-bit neighbours (same popcount) are 4-5x closer in launch order, so
// more of the far bonds hit the L2 -- or "key:a,s,b" = the shard planner's popcount-block ordering [a,s)[s,b) (s = -1:
// one block).  Unset: ascending H.
static void u1_apply_env_order(FastU1Plan* plan) {
  const char* e = getenv("EDCUDA_U1_ORDER");
  if (!e || !*e) return;
  std::vector<uint32_t> order;
  int a = 0, sp = -1, b = 0;
  if (!strcmp(e, "class")) u1_order_by_key(plan, {"class", 2, 0, -1, plan->hb}, order);
  else if (sscanf(e, "key:%d,%d,%d", &a, &sp, &b) == 3 && a >= 0 && b <= plan->hb && a < b) u1_order_by_key(plan, {"key", 2, a, sp, b}, order);
  else return;
  plan->tile_order.upload(order);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  plan->order_on = true;
}

static void u1_fill_params(ed_oprep* o, U1Params& P) {
  // profiling knobs, compiled in only with -DEDCUDA_PROFILING (EDCUDA_U1_ABLATE drops parts of the Hamiltonian: results are WRONG)
#ifdef EDCUDA_PROFILING
  static const int ablate = getenv("EDCUDA_U1_ABLATE") ? atoi(getenv("EDCUDA_U1_ABLATE")) : 0;
  static const int far_bit = getenv("EDCUDA_U1_FARBIT") ? atoi(getenv("EDCUDA_U1_FARBIT")) : -1;   // override of the plan's choice
  if (far_bit >= 0) P.far_bit = far_bit;
  if (ablate & 1) P.n_ll = 0;
  if (ablate & 2) P.n_hh = 0;
  if (ablate & 4) P.n_mx = 0;
  if (ablate & 8) P.n_ms = 0;
  if (ablate & 16) P.n_mq = 0;
  if (ablate & 32) P.diag_mode = 0;
#endif
  (void)o;
}

void ed_apply_u1(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot) {
  (void)side;
  FastU1Plan* plan = get_plan(o, dtype);
  ED_REQUIRE(plan->supported, ED_ERR_INTERNAL, "u1 fast path requested for an unsupported representation");
  ED_REQUIRE(x != nullptr, ED_ERR_ARGUMENT, "null input vector");
  U1Params P = plan->P;
  P.row_lo = o->row_lo;
  P.row_hi = o->row_hi;
  P.accumulate = accumulate;
  P.stream_mode = 0;
  P.dir = nullptr;
  P.x_local = x;
  P.x_bulk = u1_bulk_ok(x);
  P.x_halo = nullptr;
  P.partial_first = 0;
  u1_fill_params(o, P);
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

// One chunk of a rank's tiles (sharded matvec, ctx.cu): x through the directory, y and the dot partials by tile slot.
void ed_apply_u1_sharded(ed_oprep* o, int dtype, const U1ShardLaunch& L) {
  FastU1Plan* plan = get_plan(o, dtype);
  ED_REQUIRE(plan->supported, ED_ERR_INTERNAL, "u1 fast path requested for an unsupported representation");
  if (L.count <= 0) return;
  U1Params P = plan->P;
  P.row_lo = 0;
  P.row_hi = INT64_MAX;
  P.accumulate = L.stream_mode == 2 ? 1 : L.accumulate;
  P.stream_mode = L.stream_mode;
  if (P.stream_mode == 2) P.n_ll = 0;      // remote pass: neighbour streams only, added to what the local pass wrote
  P.dir = L.dir;
  P.x_local = L.x_local;
  P.x_bulk = u1_bulk_ok(L.x_local);
  P.x_halo = L.x_halo;
  P.tile_H = L.tile_H;
  P.tile_first = L.first;
  P.partial_first = L.first;
  P.tile_order = nullptr;
  u1_fill_params(o, P);
  if (dtype == ED_F64) launch_u1<double, 7>(plan, P, L.count, L.y_local, L.partials);
  else launch_u1<c128, 7>(plan, P, L.count, L.y_local, L.partials);
}

// Host-only description of the shard layout (no device needed): what tests/test_shard_plan.py replays on CPU with gloo.
extern "C" int ed_shard_plan_describe(const ed_operator* op, int32_t n_bits, int32_t n_set, int32_t dtype, int32_t world, int32_t rank,
                                      int32_t n_chunks, int32_t policy, int64_t* counts, int64_t* ranges, int64_t* tiles,
                                      int64_t* pulls, int64_t* packs, int64_t* reads, int64_t* pushes) {
  ED_TRY
  ED_REQUIRE(op && counts, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  std::shared_ptr<FastU1Plan> plan = ed_u1_host_plan(*op, n_bits, n_set, dtype);
  ED_REQUIRE(plan->supported, ED_ERR_UNSUPPORTED, "the operator does not lower to the tiled U(1) kernel");
  U1ShardLayout S;
  ed_u1_shard_layout(plan.get(), world, rank, n_chunks, policy, &S);
  std::vector<int32_t> index_of((size_t)1 << plan->hb, -1);
  for (size_t t = 0; t < plan->h_tile_H.size(); ++t) index_of[plan->h_tile_H[t]] = (int32_t)t;
  int64_t n_reads = 0;
  auto each_read = [&](auto&& f) {
    for (size_t i = 0; i < S.tile_H.size(); ++i) {
      const uint32_t H = S.tile_H[i];
      auto touch = [&](uint32_t H2) { if (index_of[H2] >= 0) f((int64_t)i, index_of[H2], H2); };
      for (int b = 0; b < plan->n_hh; ++b)
        if (((H >> plan->h_hh_p[b]) ^ (H >> plan->h_hh_q[b])) & 1u) touch(H ^ ((1u << plan->h_hh_p[b]) | (1u << plan->h_hh_q[b])));
      for (int b = 0; b < plan->n_mx; ++b) touch(H ^ (1u << plan->h_mx_q[b]));
      for (int b = 0; b < plan->n_ms; ++b) touch(H ^ (1u << plan->h_ms_q[b]));
    }
  };
  each_read([&](int64_t, int32_t, uint32_t) { ++n_reads; });
  uint64_t dim = 0;
  for (uint64_t v : plan->h_size) dim += v;
  counts[0] = S.n_local; counts[1] = S.n_halo; counts[2] = (int64_t)S.range_lo.size(); counts[3] = (int64_t)S.tile_H.size();
  counts[4] = (int64_t)S.pulls.size(); counts[5] = n_reads; counts[6] = S.n_chunks; counts[7] = (int64_t)dim;
  counts[8] = (int64_t)S.packs.size(); counts[9] = S.n_send; counts[10] = (int64_t)S.pushes.size(); counts[11] = 0;
  if (pushes)
    for (size_t i = 0; i < S.pushes.size(); ++i) {
      pushes[5 * i] = S.pushes[i].recv; pushes[5 * i + 1] = S.pushes[i].chunk; pushes[5 * i + 2] = S.pushes[i].src_off;
      pushes[5 * i + 3] = S.pushes[i].dst_off; pushes[5 * i + 4] = S.pushes[i].len;
    }
  if (packs)
    for (size_t i = 0; i < S.packs.size(); ++i) { packs[3 * i] = S.packs[i].src_off; packs[3 * i + 1] = S.packs[i].dst_off; packs[3 * i + 2] = S.packs[i].len; }
  if (ranges) for (size_t k = 0; k < S.range_lo.size(); ++k) { ranges[2 * k] = S.range_lo[k]; ranges[2 * k + 1] = S.range_hi[k]; }
  if (tiles)
    for (size_t i = 0; i < S.tile_H.size(); ++i) {
      const int32_t t = index_of[S.tile_H[i]];
      tiles[4 * i] = (int64_t)plan->h_base[t]; tiles[4 * i + 1] = (int64_t)plan->h_size[t]; tiles[4 * i + 2] = S.tile_off[i];
      int c = 0;
      while (c + 1 < S.n_chunks && (int)i >= S.chunk_first[c + 1]) ++c;
      tiles[4 * i + 3] = c;
    }
  if (pulls)
    for (size_t i = 0; i < S.pulls.size(); ++i) {
      pulls[5 * i] = S.pulls[i].peer; pulls[5 * i + 1] = S.pulls[i].chunk; pulls[5 * i + 2] = S.pulls[i].src_off;
      pulls[5 * i + 3] = S.pulls[i].dst_off; pulls[5 * i + 4] = S.pulls[i].len;
    }
  if (reads) {
    int64_t k = 0;
    each_read([&](int64_t i, int32_t t2, uint32_t H2) {
      reads[4 * k] = i; reads[4 * k + 1] = (int64_t)plan->h_base[t2]; reads[4 * k + 2] = (int64_t)plan->h_size[t2]; reads[4 * k + 3] = S.dir[H2];
      ++k;
    });
  }
  ED_CATCH
}
