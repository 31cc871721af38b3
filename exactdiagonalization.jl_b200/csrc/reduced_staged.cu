// K6 (staged): matrix-free apply for a ReducedOperatorRepresentation, reorganised for the GPU.
//
// Same arithmetic as k6_apply_reduced in reduced.cu (reference: Symmetry/reduced_operator_representation.jl:57-116,
// Symmetry/symmetry_reduce_generic.jl:51-101), but the expensive part -- reducing every off-diagonal column word to
// its orbit minimum over the |G| group elements -- is pulled out of the per-row loop and run WORD-parallel:
//   A  k6a_count / k6a_emit  : per row, the off-diagonal column words of the term walk, compacted in (row, term) order
//   B  k6b_canonicalize      : every thread holds 4 words in registers and sweeps the group; the 6-bit-chunk
//                              permutation LUT of each element is streamed through shared memory once per CTA pass
//                              (1024 words), so the sweep is pure LDS + ALU with no divergence
//   C  k6c_combine           : per row, term walk again in the reference's order; each hit reads its (minimum word,
//                              element) pair, finds the representative (bucketed search), forms
//                              a * conj(chi_g)/sqrt(N_col) / amp_row and accumulates -- row-owner, deterministic
// Rows are processed in batches so the scratch (10 bytes per hit) stays bounded.
#include <algorithm>
#include <cstdlib>
#include <cub/cub.cuh>

#include "ed_device.cuh"

void ed_reduce_pairs(const double* partials, int n, double* out2);  // apply.cu

struct K6Terms {
  int n_terms;
  const uint64_t* mask;
  const uint64_t* match;
  const uint64_t* target;
  const double* amp;
  int amp_complex;
};

// the term table staged in shared memory (broadcast LDS instead of uniform global loads in the per-row term walks)
struct K6TermsS {
  const uint64_t* mask; const uint64_t* match; const uint64_t* target;
};
__device__ __forceinline__ K6TermsS k6_stage_terms(const K6Terms& T, unsigned char* smem) {
  uint64_t* m = reinterpret_cast<uint64_t*>(smem);
  for (int t = threadIdx.x; t < T.n_terms; t += blockDim.x) {
    m[t] = __ldg(T.mask + t);
    m[T.n_terms + t] = __ldg(T.match + t);
    m[2 * T.n_terms + t] = __ldg(T.target + t);
  }
  __syncthreads();
  return K6TermsS{m, m + T.n_terms, m + 2 * T.n_terms};
}

__global__ void __launch_bounds__(256)
k6a_count(K6Terms T, const uint64_t* __restrict__ rwords, int64_t row0, int64_t n, int64_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char k6_smem[];
  const K6TermsS TS = k6_stage_terms(T, k6_smem);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t b = __ldg(rwords + row0 + i);
    int c = 0;
    for (int t = 0; t < T.n_terms; ++t) {
      const uint64_t m = TS.mask[t];
      if ((b & m) == TS.match[t] && ((b & ~m) | TS.target[t]) != b) ++c;
    }
    counts[i] = c;
  }
}

__global__ void __launch_bounds__(256)
k6a_emit(K6Terms T, const uint64_t* __restrict__ rwords, int64_t row0, int64_t n, const int64_t* __restrict__ offs,
         uint64_t* __restrict__ hit_words) {
  extern __shared__ __align__(16) unsigned char k6_smem[];
  const K6TermsS TS = k6_stage_terms(T, k6_smem);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t b = __ldg(rwords + row0 + i);
    int64_t at = offs[i];
    for (int t = 0; t < T.n_terms; ++t) {
      const uint64_t m = TS.mask[t];
      if ((b & m) != TS.match[t]) continue;
      const uint64_t b2 = (b & ~m) | TS.target[t];
      if (b2 != b) hit_words[at++] = b2;
    }
  }
}

// B: words[i] <- min_g g(words[i]); garg[i] <- the index i* with words[i] = g_{i*}(min) that the reference's Dict keeps
// (largest such index, see reduced_map_word).  lut6 layout: [g][chunk][64] uint64.
template <int NCH>
__global__ void __launch_bounds__(256)
k6b_canonicalize(int n_ops, const uint64_t* __restrict__ lut6, const int32_t* __restrict__ inverse, int64_t n_words,
                 uint64_t* __restrict__ words, uint16_t* __restrict__ garg) {
  constexpr int W = 4;          // words per thread
  constexpr int GB = 4;         // group elements staged per barrier (2 x GB x NCH x 512 B of shared memory)
  __shared__ __align__(16) uint64_t s_lut[2][GB * NCH * 64];
  __shared__ int s_inv[2][GB];
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * (256 * W);
  uint64_t w[W], best[W];
  int besti[W];
  uint32_t off[W][NCH];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int64_t i = base + tid + k * 256;
    w[k] = i < n_words ? words[i] : 0ull;
    best[k] = w[k];
    besti[k] = 0;   // identity (element 0, its own inverse)
#pragma unroll
    for (int c = 0; c < NCH; ++c) off[k][c] = (uint32_t)((w[k] >> (6 * c)) & 63ull) + c * 64;
  }
  const int n_batches = (n_ops + GB - 1) / GB;
  auto stage = [&](int batch, int buf) {
    const int g0 = batch * GB;
    const int ng = min(GB, n_ops - g0);
    const uint64_t* src = lut6 + (size_t)g0 * NCH * 64;
    for (int i = tid; i < ng * NCH * 64; i += 256) s_lut[buf][i] = __ldg(src + i);
    if (tid < ng) s_inv[buf][tid] = __ldg(inverse + g0 + tid);
  };
  stage(0, 0);
  __syncthreads();
  for (int batch = 0; batch < n_batches; ++batch) {
    const int buf = batch & 1;
    if (batch + 1 < n_batches) stage(batch + 1, buf ^ 1);
    const int g0 = batch * GB;
    const int ng = min(GB, n_ops - g0);
    for (int gi = (batch == 0 ? 1 : 0); gi < ng; ++gi) {   // element 0 is the identity
      const uint64_t* L = s_lut[buf] + gi * NCH * 64;
      const int inv = s_inv[buf][gi];
#pragma unroll
      for (int k = 0; k < W; ++k) {
        uint64_t im = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) im |= L[off[k][c]];
        if (im < best[k]) { best[k] = im; besti[k] = inv; }
        else if (im == best[k] && inv > besti[k]) besti[k] = inv;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int64_t i = base + tid + k * 256;
    if (i < n_words) { words[i] = best[k]; garg[i] = (uint16_t)besti[k]; }
  }
}

// B (translation-factorised): the same minimum / element selection as k6b_canonicalize, but the group is swept as
// cosets T p_j of the lattice translations T = {Tx^a Ty^b}: one 6-bit-LUT image per coset representative p_j, then
// the n1 x n2 translations as ALU steps on the image (Tx = rotate every n1-bit field by one, Ty = rotate the word by
// n1).  |G| / (n1 n2) LUT images per word instead of |G|: the sweep moves from the shared-memory pipe (bank
// conflicts of the LUT gathers) to the integer pipes.  tinv[(j*n2 + b)*n1 + a] = inverse index of Tx^a Ty^b p_j.
// N1 x N2 > 0: lattice shape known at compile time (6x6, 4x4): both translation loops are fully unrolled, the masks are
// immediates and the rotation that would only restore the starting image (last Tx of a row sweep, last Ty) is skipped.
template <int NCH, int N1, int N2>
__global__ void __launch_bounds__(256)
k6b_canonicalize_tr(int n_cos, int n1_rt, int n2_rt, int n_bits_rt, const uint64_t* __restrict__ lut6c, const int32_t* __restrict__ tinv,
                    int64_t n_words, uint64_t* __restrict__ words, uint16_t* __restrict__ garg) {
  constexpr int W = 4;          // words per thread
  constexpr int CB = 4;         // cosets staged per barrier
  constexpr bool FIXED = N1 > 0;
  const int n1 = FIXED ? N1 : n1_rt, n2 = FIXED ? N2 : n2_rt, n_bits = FIXED ? N1 * N2 : n_bits_rt;
  extern __shared__ __align__(16) unsigned char k6_smem[];
  uint64_t* s_lut = reinterpret_cast<uint64_t*>(k6_smem);                   // [2][CB * NCH * 64]
  int32_t* s_inv = reinterpret_cast<int32_t*>(s_lut + 2 * CB * NCH * 64);   // [2][CB * n1 * n2]
  const int nt = n1 * n2;
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * (256 * W);
  const uint64_t full = n_bits >= 64 ? ~0ull : ((1ull << n_bits) - 1ull);
  uint64_t m_lo = 0;
#pragma unroll
  for (int y = 0; y < (FIXED ? N2 : 64); ++y)
    if (y < n2) m_lo |= 1ull << (n1 * y);
  const uint64_t m_hi = full & ~m_lo;
  uint64_t w[W], best[W];
  int besti[W];
  uint32_t off[W][NCH];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int64_t i = base + tid + k * 256;
    w[k] = i < n_words ? words[i] : 0ull;
    best[k] = w[k];
    besti[k] = 0;   // identity (element 0, its own inverse)
#pragma unroll
    for (int c = 0; c < NCH; ++c) off[k][c] = (uint32_t)((w[k] >> (6 * c)) & 63ull) + c * 64;
  }
  const int n_batches = (n_cos + CB - 1) / CB;
  auto stage = [&](int batch, int buf) {
    const int j0 = batch * CB;
    const int nj = min(CB, n_cos - j0);
    const uint64_t* src = lut6c + (size_t)j0 * NCH * 64;
    uint64_t* dst = s_lut + (size_t)buf * CB * NCH * 64;
    for (int i = tid; i < nj * NCH * 64; i += 256) dst[i] = __ldg(src + i);
    int32_t* di = s_inv + (size_t)buf * CB * nt;
    for (int i = tid; i < nj * nt; i += 256) di[i] = __ldg(tinv + (size_t)j0 * nt + i);
  };
  stage(0, 0);
  __syncthreads();
  for (int batch = 0; batch < n_batches; ++batch) {
    const int buf = batch & 1;
    if (batch + 1 < n_batches) stage(batch + 1, buf ^ 1);
    const int nj = min(CB, n_cos - batch * CB);
    for (int ji = 0; ji < nj; ++ji) {
      const uint64_t* L = s_lut + ((size_t)buf * CB + ji) * NCH * 64;
      const int32_t* inv_tab = s_inv + ((size_t)buf * CB + ji) * nt;
      uint64_t im[W];
#pragma unroll
      for (int k = 0; k < W; ++k) {
        uint64_t v = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) v |= L[off[k][c]];
        im[k] = v;
      }
#pragma unroll(FIXED ? N2 : 1)
      for (int b = 0; b < n2; ++b) {
        uint64_t u[W];
#pragma unroll
        for (int k = 0; k < W; ++k) u[k] = im[k];
#pragma unroll(FIXED ? N1 : 1)
        for (int a = 0; a < n1; ++a) {
          bool hit = false;
#pragma unroll
          for (int k = 0; k < W; ++k) hit |= u[k] <= best[k];
          if (__any_sync(0xffffffffu, hit)) {            // a real (warp-uniform) branch: new minima and ties are rare
            const int inv = inv_tab[b * n1 + a];
#pragma unroll
            for (int k = 0; k < W; ++k) {
              if (u[k] < best[k]) { best[k] = u[k]; besti[k] = inv; }
              else if (u[k] == best[k] && inv > besti[k]) besti[k] = inv;
            }
          }
          if (!FIXED || a + 1 < N1) {
#pragma unroll
            for (int k = 0; k < W; ++k) u[k] = ((u[k] << 1) & m_hi) | ((u[k] >> (n1 - 1)) & m_lo);        // Tx
          }
        }
        if (!FIXED || b + 1 < N2) {
#pragma unroll
          for (int k = 0; k < W; ++k) im[k] = ((im[k] << n1) | (im[k] >> (n_bits - n1))) & full;   // Ty
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int64_t i = base + tid + k * 256;
    if (i < n_words) { words[i] = best[k]; garg[i] = (uint16_t)besti[k]; }
  }
}

// B (necklace): the same minimum / element selection again, but the n1 x n2 translations of a coset image are not swept
// at all.  The minimum over Tx^a Ty^b of a word has, as its most significant row, the smallest ROTATION of any of the
// word's rows: a 2^n1-entry table gives for every row value its minimal rotation and the set of shifts a that reach it.
// Per coset: one table look-up per row, the row minimum m, and -- only when m does not exceed the top row of the best word
// so far -- a full comparison for the few (row, shift) pairs that reach m (usually one).  Every element attaining the
// global minimum is among those candidates (it must minimise the top row inside its coset), so (best, besti) come out
// exactly as in the full sweep: ~45 instead of ~245 instructions per (word, coset).
template <int NCH, int N1, int N2>
__global__ void __launch_bounds__(256)
k6b_canonicalize_nk(int n_cos, int n1_rt, int n2_rt, const uint64_t* __restrict__ lut6c, const int32_t* __restrict__ tinv,
                    int64_t n_words, uint64_t* __restrict__ words, uint16_t* __restrict__ garg) {
  constexpr int W = 2;          // words per thread
  constexpr int CB = 4;         // cosets staged per barrier
  constexpr bool FIXED = N1 > 0;
  const int n1 = FIXED ? N1 : n1_rt, n2 = FIXED ? N2 : n2_rt, n_bits = n1 * n2;
  extern __shared__ __align__(16) unsigned char k6_smem[];
  uint64_t* s_lut = reinterpret_cast<uint64_t*>(k6_smem);                   // [2][CB * NCH * 64]
  int32_t* s_inv = reinterpret_cast<int32_t*>(s_lut + 2 * CB * NCH * 64);   // [2][CB * n1 * n2]
  uint64_t* s_mlo = reinterpret_cast<uint64_t*>(s_inv + 2 * CB * n1 * n2 + ((2 * CB * n1 * n2) & 1));   // [8] columns x < a of every row
  uint16_t* s_row = reinterpret_cast<uint16_t*>(s_mlo + 8);                 // [2^n1] (minimal rotation << 8) | shifts reaching it
  // rows taken two at a time where that fits (2 n1 <= 12 bits, even n2): [2^(2 n1)] min of the two minimal rotations --
  // half the look-ups of the scan that every (word, coset) pays; the rare candidate path still uses s_row
  const bool pairs = (2 * n1 <= 12) && (n2 % 2 == 0);
  uint16_t* s_pair = s_row + (1 << n1);                                     // (min of the two << 8) | which of the two rows attain it
  const int nt = n1 * n2;
  const int tid = threadIdx.x;
  const uint64_t full = n_bits >= 64 ? ~0ull : ((1ull << n_bits) - 1ull);
  const uint32_t rmask = (1u << n1) - 1u;
  for (int v = tid; v < (1 << n1); v += 256) {
    uint32_t best = (uint32_t)v, arg = 1u, r = (uint32_t)v;
    for (int a = 1; a < n1; ++a) {
      r = ((r << 1) | (r >> (n1 - 1))) & rmask;                // rotate left by one: Tx on one row
      if (r < best) { best = r; arg = 1u << a; }
      else if (r == best) arg |= 1u << a;
    }
    s_row[v] = (uint16_t)((best << 8) | arg);
  }
  if (tid < 8) {
    uint64_t m = 0;
    for (int y = 0; y < n2; ++y) m |= (uint64_t)((1u << tid) - 1u) << (n1 * y);
    s_mlo[tid] = tid <= n1 ? (m & full) : 0ull;
  }
  if (pairs) {
    __syncthreads();
    for (int v = tid; v < (1 << (2 * n1)); v += 256) {
      const uint32_t lo = (uint32_t)s_row[v & rmask] >> 8, hi = (uint32_t)s_row[(v >> n1) & rmask] >> 8;
      const uint32_t mn = min(lo, hi);
      s_pair[v] = (uint16_t)((mn << 8) | (lo == mn ? 1u : 0u) | (hi == mn ? 2u : 0u));
    }
  }
  const int64_t base = (int64_t)blockIdx.x * (256 * W);
  uint64_t w[W], best[W];
  int besti[W];
  uint32_t off[W][NCH];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int64_t i = base + tid + k * 256;
    w[k] = i < n_words ? words[i] : 0ull;
    best[k] = w[k];
    besti[k] = 0;   // identity (element 0, its own inverse)
#pragma unroll
    for (int c = 0; c < NCH; ++c) off[k][c] = (uint32_t)((w[k] >> (6 * c)) & 63ull) + c * 64;
  }
  const int n_batches = (n_cos + CB - 1) / CB;
  auto stage = [&](int batch, int buf) {
    const int j0 = batch * CB;
    const int nj = min(CB, n_cos - j0);
    const uint64_t* src = lut6c + (size_t)j0 * NCH * 64;
    uint64_t* dst = s_lut + (size_t)buf * CB * NCH * 64;
    for (int i = tid; i < nj * NCH * 64; i += 256) dst[i] = __ldg(src + i);
    int32_t* di = s_inv + (size_t)buf * CB * nt;
    for (int i = tid; i < nj * nt; i += 256) di[i] = __ldg(tinv + (size_t)j0 * nt + i);
  };
  stage(0, 0);
  __syncthreads();
  const int top_shift = n_bits - n1;
  for (int batch = 0; batch < n_batches; ++batch) {
    const int buf = batch & 1;
    if (batch + 1 < n_batches) stage(batch + 1, buf ^ 1);
    const int nj = min(CB, n_cos - batch * CB);
    for (int ji = 0; ji < nj; ++ji) {
      const uint64_t* L = s_lut + ((size_t)buf * CB + ji) * NCH * 64;
      const int32_t* inv_tab = s_inv + ((size_t)buf * CB + ji) * nt;
      uint64_t im[W];
      uint32_t m[W], rows[W];     // row minimum of the image and the set of rows that attain it
      bool cand = false;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        uint64_t v = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) v |= L[off[k][c]];
        im[k] = v;
        uint32_t mm = 0xFFu, rr = 0u;
        if (pairs) {
          const uint32_t pmask = (1u << (2 * n1)) - 1u;
#pragma unroll(FIXED ? N2 / 2 : 1)
          for (int y = 0; y < n2; y += 2) {
            const uint32_t t = s_pair[(uint32_t)(v >> (n1 * y)) & pmask];
            const uint32_t tm = t >> 8;
            rr = tm < mm ? ((t & 3u) << y) : (tm == mm ? (rr | ((t & 3u) << y)) : rr);
            mm = min(mm, tm);
          }
        } else {
#pragma unroll(FIXED ? N2 : 1)
          for (int y = 0; y < n2; ++y) {
            const uint32_t tm = (uint32_t)s_row[(uint32_t)(v >> (n1 * y)) & rmask] >> 8;
            rr = tm < mm ? (1u << y) : (tm == mm ? (rr | (1u << y)) : rr);
            mm = min(mm, tm);
          }
        }
        m[k] = mm;
        rows[k] = rr;
        cand |= mm <= (uint32_t)(best[k] >> top_shift);
      }
      if (__any_sync(0xffffffffu, cand)) {            // warp-uniform branch: most cosets cannot beat the best word so far
#pragma unroll
        for (int k = 0; k < W; ++k) {
          uint32_t rr = m[k] <= (uint32_t)(best[k] >> top_shift) ? rows[k] : 0u;
          while (rr) {                                // usually one row
            const int y = __ffs(rr) - 1;
            rr &= rr - 1u;
            const int b = n2 - 1 - y;                  // Ty^b brings row y to the top
            const uint64_t ub = b ? (((im[k] << (n1 * b)) | (im[k] >> (n_bits - n1 * b))) & full) : im[k];
            uint32_t shifts = (uint32_t)s_row[(uint32_t)(im[k] >> (n1 * y)) & rmask] & 0xFFu;
            while (shifts) {                          // usually one shift
              const int a = __ffs(shifts) - 1;
              shifts &= shifts - 1u;
              const uint64_t mlo = s_mlo[a];
              const uint64_t u = a ? (((ub << a) & full & ~mlo) | ((ub >> (n1 - a)) & mlo)) : ub;     // Tx^a
              const int inv = inv_tab[b * n1 + a];
              if (u < best[k]) { best[k] = u; besti[k] = inv; }
              else if (u == best[k] && inv > besti[k]) besti[k] = inv;
            }
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int64_t i = base + tid + k * 256;
    if (i < n_words) { words[i] = best[k]; garg[i] = (uint16_t)besti[k]; }
  }
}

// B (necklace, two-row keys, queued).  What limits k6b_canonicalize_nk (ncu source page): its running-minimum filter on ONE
// row lets most cosets through -- the four point-group elements that keep a line direction produce the same rows, rows with at
// most two particles equal their mirror necklace, and an empty row is minimal under every shift: 10 (coset, row, shift)
// candidates per word on the 6x6 lattice -- and one candidate lane sends its whole warp down the comparison path (42 % of the
// instructions at 18 of 32 lanes).  Here
//   * the key is the top TWO rows: a 2^(2 n1)-entry table gives, for (row y, row y-1), the smallest pair under a common shift
//     and the shifts reaching it, so ties are broken inside the table (about 1.5 candidates per word);
//   * pass A (uniform, branch-free) finds m* = the smallest key over ALL cosets of a word and the cosets attaining it;
//   * pass B (warp-cooperative): the (word, coset) pairs of a warp are compacted into a shared-memory queue and handed out one
//     per lane; a pair rebuilds its image, walks the (row, shift) candidates that reach m* and folds word || element into the
//     word's slot with a 64-bit shared-memory atomicMin.
// The minimum over the group has top rows m*, so every element attaining it is among the candidates: (best, besti) as in the
// full sweep.  key = word << 16 | (0xFFFF - inverse index): smallest word first, then the LARGEST inverse index (the Dict
// overwrite).  CTAs are persistent (tables built once per CTA, all coset tables resident).
template <int NCH, int N1, int N2>
__global__ void __launch_bounds__(256)
k6b_canonicalize_nkq(int n_cos, int n1_rt, int n2_rt, const uint64_t* __restrict__ lut6c, const int32_t* __restrict__ tinv,
                     int64_t n_words, uint64_t* __restrict__ words, uint16_t* __restrict__ garg) {
  constexpr int W = 2;          // words per thread and block
  constexpr bool FIXED = N1 > 0;
  const int n1 = FIXED ? N1 : n1_rt, n2 = FIXED ? N2 : n2_rt, n_bits = n1 * n2;
  const int nt = n1 * n2;
  extern __shared__ __align__(16) unsigned char k6_smem[];
  uint64_t* s_lut = reinterpret_cast<uint64_t*>(k6_smem);                   // [n_cos][NCH * 64]
  uint64_t* s_mlo = s_lut + (size_t)n_cos * NCH * 64;                       // [8] columns x < a of every row
  uint64_t* s_w = s_mlo + 8;                                                // [8 warps][32] words of the half in flight
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(s_w + 256);   // [8][32]
  int32_t* s_inv = reinterpret_cast<int32_t*>(s_best + 256);                // [n_cos][n1 * n2]
  uint32_t* s_ms = reinterpret_cast<uint32_t*>(s_inv + (size_t)n_cos * nt); // [8][32] m* of those words
  uint16_t* s_q = reinterpret_cast<uint16_t*>(s_ms + 256);                  // [8][32 * n_cos] queue entries lane << 8 | coset
  uint16_t* s_key = s_q + (size_t)8 * 32 * n_cos;                           // [2^(2 n1)] smallest (top, next) pair under a common shift
  uint8_t* s_sh = reinterpret_cast<uint8_t*>(s_key + ((size_t)1 << (2 * n1)));   // [2^(2 n1)] the shifts reaching it
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t full = n_bits >= 64 ? ~0ull : ((1ull << n_bits) - 1ull);
  const uint32_t rmask = (1u << n1) - 1u;
  const uint32_t pmask = (1u << (2 * n1)) - 1u;
  for (int i = tid; i < n_cos * NCH * 64; i += 256) s_lut[i] = __ldg(lut6c + i);
  for (int i = tid; i < n_cos * nt; i += 256) s_inv[i] = __ldg(tinv + i);
  for (int v = tid; v < (1 << (2 * n1)); v += 256) {
    uint32_t top = ((uint32_t)v >> n1) & rmask, nxt = (uint32_t)v & rmask;
    uint32_t best = (uint32_t)v, arg = 1u;
    for (int a = 1; a < n1; ++a) {
      top = ((top << 1) | (top >> (n1 - 1))) & rmask;          // rotate both rows left by one: Tx
      nxt = ((nxt << 1) | (nxt >> (n1 - 1))) & rmask;
      const uint32_t r = (top << n1) | nxt;
      if (r < best) { best = r; arg = 1u << a; }
      else if (r == best) arg |= 1u << a;
    }
    s_key[v] = (uint16_t)best;
    s_sh[v] = (uint8_t)arg;
  }
  if (tid < 8) {
    uint64_t m = 0;
    for (int y = 0; y < n2; ++y) m |= (uint64_t)((1u << tid) - 1u) << (n1 * y);
    s_mlo[tid] = tid <= n1 ? (m & full) : 0ull;
  }
  __syncthreads();
  uint64_t* my_w = s_w + warp * 32;
  unsigned long long* my_best = s_best + warp * 32;
  uint32_t* my_ms = s_ms + warp * 32;
  uint16_t* my_q = s_q + (size_t)warp * 32 * n_cos;
  const int wrap_shift = n_bits - n1;                                       // the top row, wrapped below row 0 for y = 0
  for (int64_t base = (int64_t)blockIdx.x * (256 * W); base < n_words; base += (int64_t)gridDim.x * (256 * W)) {
    uint64_t w[W];
    uint32_t off[W][NCH];
    uint32_t mstar[W], cosets[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const int64_t i = base + tid + k * 256;
      w[k] = i < n_words ? words[i] : 0ull;
#pragma unroll
      for (int c = 0; c < NCH; ++c) off[k][c] = (uint32_t)((w[k] >> (6 * c)) & 63ull) + c * 64;
      mstar[k] = 0xFFFFu;
      cosets[k] = 0u;
    }
    // ---- pass A
#pragma unroll 2
    for (int j = 0; j < n_cos; ++j) {
      const uint64_t* L = s_lut + (size_t)j * NCH * 64;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        uint64_t v = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) v |= L[off[k][c]];
        // (row 0, row n2-1) first, then (row y, row y-1) as contiguous 2 n1-bit windows
        uint32_t mm = s_key[(((uint32_t)v & rmask) << n1) | (uint32_t)(v >> wrap_shift)];
#pragma unroll(FIXED ? N2 - 1 : 1)
        for (int y = 1; y < n2; ++y) mm = min(mm, (uint32_t)s_key[(uint32_t)(v >> (n1 * (y - 1))) & pmask]);
        cosets[k] = mm < mstar[k] ? (1u << j) : (mm == mstar[k] ? (cosets[k] | (1u << j)) : cosets[k]);
        mstar[k] = min(mstar[k], mm);
      }
    }
    // ---- pass B, one half (the k-th word of every lane) at a time
#pragma unroll
    for (int k = 0; k < W; ++k) {
      my_w[lane] = w[k];
      my_best[lane] = ~0ull;
      my_ms[lane] = mstar[k];
      uint32_t cs = (base + tid + k * 256) < n_words ? cosets[k] : 0u;      // padding lanes queue nothing
      const int cnt = __popc(cs);
      int pre = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
      }
      const int total = __shfl_sync(0xffffffffu, pre, 31);
      int pos = pre - cnt;
      while (cs) {
        const int j = __ffs(cs) - 1;
        cs &= cs - 1u;
        my_q[pos++] = (uint16_t)((lane << 8) | j);
      }
      __syncwarp();
#pragma unroll 1
      for (int it = lane; it < total; it += 32) {
        const uint32_t e = my_q[it];
        const int wl = (int)(e >> 8), j = (int)(e & 255u);
        const uint64_t ww = my_w[wl];
        const uint32_t ms = my_ms[wl];
        const uint64_t* L = s_lut + (size_t)j * NCH * 64;
        const int32_t* inv_tab = s_inv + (size_t)j * nt;
        uint64_t im = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) im |= L[(uint32_t)((ww >> (6 * c)) & 63ull) + c * 64];
        unsigned long long lb = ~0ull;
#pragma unroll 1
        for (int y = 0; y < n2; ++y) {
          const uint32_t idx = y ? ((uint32_t)(im >> (n1 * (y - 1))) & pmask) : ((((uint32_t)im & rmask) << n1) | (uint32_t)(im >> wrap_shift));
          if ((uint32_t)s_key[idx] != ms) continue;
          const int b = n2 - 1 - y;                  // Ty^b brings row y to the top (and row y-1 below it)
          const uint64_t ub = b ? (((im << (n1 * b)) | (im >> (n_bits - n1 * b))) & full) : im;
          uint32_t shifts = s_sh[idx];
          while (shifts) {                          // usually one shift
            const int a = __ffs(shifts) - 1;
            shifts &= shifts - 1u;
            const uint64_t mlo = s_mlo[a];
            const uint64_t u = a ? (((ub << a) & full & ~mlo) | ((ub >> (n1 - a)) & mlo)) : ub;     // Tx^a
            const unsigned long long key = (u << 16) | (unsigned long long)(0xFFFF - inv_tab[b * n1 + a]);
            lb = key < lb ? key : lb;
          }
        }
        atomicMin(my_best + wl, lb);
      }
      __syncwarp();
      const unsigned long long key = my_best[lane];
      const int64_t i = base + tid + k * 256;
      if (i < n_words) { words[i] = key >> 16; garg[i] = (uint16_t)(0xFFFFu - (uint32_t)(key & 0xFFFFull)); }
      __syncwarp();
    }
  }
}

// B': representative index and orbit size of every canonical word, word-parallel (independent searches: the memory
// latency of the bucketed binary search is hidden by parallelism instead of sitting inside the per-row combine loop)
__global__ void __launch_bounds__(256)
k6b_lookup(RLookupDesc R, int64_t n_words, const uint64_t* __restrict__ words, int32_t* __restrict__ hit_j, uint16_t* __restrict__ hit_orbit) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = rank_reduced(R, __ldg(words + i));
    hit_j[i] = (int32_t)j;
    hit_orbit[i] = j >= 0 ? __ldg(R.orbit_size + j) : (uint16_t)1;
  }
}

__global__ void __launch_bounds__(128)
k6c_combine(K6Terms T, LookupDesc L, SymDesc S, RLookupDesc R, int64_t row0, int64_t n, int64_t out_row0,
            const int64_t* __restrict__ offs, const int32_t* __restrict__ hit_j, const uint16_t* __restrict__ hit_orbit,
            const uint16_t* __restrict__ garg, int conj_side, const c128* __restrict__ x, c128* __restrict__ out, int accumulate,
            double* __restrict__ dot_partials, int partial_slot0) {
  extern __shared__ __align__(16) unsigned char k6_smem[];
  const K6TermsS TS = k6_stage_terms(T, k6_smem);
  double dre = 0.0, dim_ = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row0 + i;
    const uint64_t b = __ldg(R.words + r);
    c128 a_self = reduced_rep_amp(S, R, r);
    if (conj_side) a_self = cconj(a_self);
    const c128 inv_self = cinv(a_self);
    c128 acc = accumulate ? out[out_row0 + i] : make_c128(0.0, 0.0);
    int64_t at = offs[i];
    for (int t = 0; t < T.n_terms; ++t) {
      const uint64_t m = TS.mask[t];
      if ((b & m) != TS.match[t]) continue;
      const uint64_t b2 = (b & ~m) | TS.target[t];
      const c128 a = T.amp_complex ? make_c128(__ldg(T.amp + 2 * t), __ldg(T.amp + 2 * t + 1)) : make_c128(__ldg(T.amp + t), 0.0);
      int64_t j;
      c128 a2;
      if (b2 == b) {
        j = r;
        a2 = a_self;
      } else {
        j = hit_j[at];
        const int gi = garg[at];
        const double norm2 = (double)hit_orbit[at];
        ++at;
        if (j < 0) continue;                                     // orbit not in this irrep
        if (!in_basis_dyn(L, b2)) continue;                     // not in the parent basis
        const double inv_norm = 1.0 / sqrt(norm2);
        a2 = make_c128(__ldg(S.chi + 2 * gi) * inv_norm, -__ldg(S.chi + 2 * gi + 1) * inv_norm);
        if (conj_side) a2 = cconj(a2);
      }
      fma_acc(acc, cmul(cmul(a, a2), inv_self), ldg_c128(x + j));
    }
    st_val(out + out_row0 + i, acc);
    if (dot_partials) dot_acc(dre, dim_, ldg_c128(x + r), acc);
  }
  if (dot_partials) {
    __shared__ double s_red[2][4];
    dre = warp_sum(dre);
    dim_ = warp_sum(dim_);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][wid] = dre; s_red[1][wid] = dim_; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, c = 0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) { a += s_red[0][w]; c += s_red[1][w]; }
      dot_partials[2 * (partial_slot0 + blockIdx.x)] = a;
      dot_partials[2 * (partial_slot0 + blockIdx.x) + 1] = c;
    }
  }
}

// sparse()/cached-CSR assembly for a reduced representation: the raw (index, amplitude) entries of k4_fill_raw
// (sparse.cu: every matching term in term order, -1 for misses) with the orbit minima taken from the staged sweep
// instead of one row-per-thread orbit search per hit.  Same arithmetic as k6c_combine / walk_line.
__global__ void __launch_bounds__(128)
k6c_fill_raw(K6Terms T, LookupDesc L, SymDesc S, RLookupDesc R, int64_t row0, int64_t n, const int64_t* __restrict__ hit_offs,
             const int32_t* __restrict__ hit_j, const uint16_t* __restrict__ hit_orbit, const uint16_t* __restrict__ garg, int conj_side,
             const int64_t* __restrict__ raw_offs, int64_t* __restrict__ raw_row, c128* __restrict__ raw_val) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row0 + i;
    const uint64_t b = __ldg(R.words + r);
    c128 a_self = reduced_rep_amp(S, R, r);
    if (conj_side) a_self = cconj(a_self);
    const c128 inv_self = cinv(a_self);
    int64_t at = hit_offs[i];
    int64_t out = raw_offs[i];
    for (int t = 0; t < T.n_terms; ++t) {
      const uint64_t m = __ldg(T.mask + t);
      if ((b & m) != __ldg(T.match + t)) continue;
      const uint64_t b2 = (b & ~m) | __ldg(T.target + t);
      const c128 a = T.amp_complex ? make_c128(__ldg(T.amp + 2 * t), __ldg(T.amp + 2 * t + 1)) : make_c128(__ldg(T.amp + t), 0.0);
      int64_t j = -1;
      c128 v = a;
      if (b2 == b) {
        j = r;
        v = cmul(cmul(a, a_self), inv_self);
      } else {
        const int64_t jj = hit_j[at];
        const int gi = garg[at];
        const double norm2 = (double)hit_orbit[at];
        ++at;
        if (in_basis_dyn(L, b2)) {
          j = jj;
          if (j >= 0) {
            const double inv_norm = 1.0 / sqrt(norm2);
            c128 a2 = make_c128(__ldg(S.chi + 2 * gi) * inv_norm, -__ldg(S.chi + 2 * gi + 1) * inv_norm);
            if (conj_side) a2 = cconj(a2);
            v = cmul(cmul(a, a2), inv_self);
          }
        }
      }
      raw_row[out] = j;
      raw_val[out] = v;
      ++out;
    }
  }
}

struct K6Scratch {
  DevBuf<int64_t> counts, offs;
  DevBuf<uint64_t> words;
  DevBuf<uint16_t> garg, horb;
  DevBuf<int32_t> hitj;
  DevBuf<unsigned char> tmp;
  DevBuf<double> partials;
};

static K6Scratch& scratch() {
  static thread_local K6Scratch s[16];     // per device (one process may drive several GPUs)
  return s[ed_current_device() & 15];
}

// Phases A + B for rows [row0, row0 + nb): sc.offs = exclusive offsets of the off-diagonal hits per row, sc.words = the
// orbit minimum of every hit, sc.garg = the element the reference's Dict keeps for it.  Returns the number of hits.
// *known_hits >= 0: the number of column words this batch emits is known from an earlier matvec of the same representation --
// no count read-back, no host synchronisation; < 0: read it back and store it there
static int64_t k6_stage_batch(ed_oprep* o, const K6Terms& T, int64_t row0, int64_t nb, K6Scratch& sc, cudaEvent_t ev_mid,
                              int64_t* known_hits = nullptr) {
  ed_rbasis* rb = o->rbasis;
  ed_basis* parent = rb->parent;
  const SymDesc S = rb->symdesc();
  const int sm = ed_sm_count();
  if (sc.counts.n < (size_t)nb + 1) { sc.counts.alloc((size_t)nb + 1); sc.offs.alloc((size_t)nb + 1); }
  ED_CUDA(cudaMemsetAsync(sc.counts.p + nb, 0, sizeof(int64_t), ed_stream()));
  const int grid_a = (int)std::max<int64_t>(1, std::min<int64_t>((nb + 255) / 256, (int64_t)sm * 16));
  const size_t smem_t = (size_t)3 * T.n_terms * sizeof(uint64_t);
  ED_LAUNCH(k6a_count, grid_a, 256, smem_t, T, rb->words.p, row0, nb, sc.counts.p);
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, sc.counts.p, sc.offs.p, nb + 1, ed_stream());
  if (sc.tmp.n < bytes) sc.tmp.alloc(bytes);
  cub::DeviceScan::ExclusiveSum(sc.tmp.p, bytes, sc.counts.p, sc.offs.p, nb + 1, ed_stream());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  int64_t n_hits = 0;
  if (known_hits && *known_hits >= 0) {
    n_hits = *known_hits;
  } else {
    ED_CUDA(cudaMemcpyAsync(&n_hits, sc.offs.p + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, ed_stream()));
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    if (known_hits) *known_hits = n_hits;
  }
  if (sc.words.n < (size_t)std::max<int64_t>(n_hits, 1)) {
    sc.words.alloc((size_t)std::max<int64_t>(n_hits, 1));
    sc.garg.alloc((size_t)std::max<int64_t>(n_hits, 1));
    sc.horb.alloc((size_t)std::max<int64_t>(n_hits, 1));
    sc.hitj.alloc((size_t)std::max<int64_t>(n_hits, 1));
  }
  if (n_hits > 0) {
    ED_LAUNCH(k6a_emit, grid_a, 256, smem_t, T, rb->words.p, row0, nb, sc.offs.p, sc.words.p);
    if (ev_mid) ED_CUDA(cudaEventRecord(ev_mid, ed_stream()));
    const int grid_b = (int)((n_hits + 1023) / 1024);
    const int nch = rb->symdev.n_chunks6;
    const uint64_t* lut6 = rb->symdev.lut6.p;
    const int32_t* inv = rb->symdev.inverse.p;
    const SymDev& sd = rb->symdev;
    static const bool no_necklace = getenv("EDCUDA_K6_NONECKLACE") != nullptr;
    if (sd.tr_on && nch <= 8 && sd.tr_n1 <= 8 && sd.tr_n1 >= 2 && parent->space.bits == sd.tr_n1 * sd.tr_n2 && !no_necklace) {
      const int nt_ = sd.tr_n1 * sd.tr_n2;
      const int grid_n = (int)((n_hits + 511) / 512);
      // two-row keys + queued comparisons when n1 <= 6 (2^(2 n1)-entry table), the tables of all cosets fit shared memory, the
      // coset set fits a 32-bit mask and word || element fits 64 bits (EDCUDA_K6_NK1=1: the one-pass kernel)
      static const bool one_pass = getenv("EDCUDA_K6_NK1") != nullptr;
      const int nch_t = (sd.tr_n1 == 6 && sd.tr_n2 == 6 && nch == 6) ? 6 : (nch <= 4 ? 4 : (nch <= 6 ? 6 : 8));
      const size_t smem_q = (size_t)sd.tr_ncos * nch_t * 64 * 8 + 8 * 8 + 256 * 8 + 256 * 8 + (size_t)sd.tr_ncos * nt_ * 4 + 256 * 4 +
                            (size_t)8 * 32 * sd.tr_ncos * 2 + ((size_t)3 << (2 * sd.tr_n1));
      if (!one_pass && sd.tr_n1 <= 6 && sd.tr_n2 >= 2 && sd.tr_ncos <= 32 && nt_ <= 48 && smem_q <= 72 * 1024) {
        const int grid_q = (int)std::min<int64_t>((n_hits + 511) / 512, (int64_t)sm * 3);
#define ED_K6NKQ(NCH_, N1_, N2_)                                                                                             \
        do {                                                                                                                 \
          ED_CUDA(cudaFuncSetAttribute(k6b_canonicalize_nkq<NCH_, N1_, N2_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q)); \
          ED_LAUNCH((k6b_canonicalize_nkq<NCH_, N1_, N2_>), grid_q, 256, smem_q, sd.tr_ncos, sd.tr_n1, sd.tr_n2, sd.tr_lut6.p, \
                    sd.tr_inv.p, n_hits, sc.words.p, sc.garg.p);                                                             \
        } while (0)
        if (sd.tr_n1 == 6 && sd.tr_n2 == 6 && nch == 6) ED_K6NKQ(6, 6, 6);
        else if (sd.tr_n1 == 4 && sd.tr_n2 == 4 && nch <= 4) ED_K6NKQ(4, 4, 4);
        else if (nch <= 4) ED_K6NKQ(4, 0, 0);
        else if (nch <= 6) ED_K6NKQ(6, 0, 0);
        else ED_K6NKQ(8, 0, 0);
#undef ED_K6NKQ
      } else {
#define ED_K6NK(NCH_, N1_, N2_)                                                                                              \
      do {                                                                                                                   \
        const size_t smem_n = (size_t)2 * 4 * NCH_ * 64 * 8 + (size_t)(2 * 4 * nt_ + ((2 * 4 * nt_) & 1)) * 4 + 8 * 8 + ((size_t)2 << sd.tr_n1) + \
                              ((2 * sd.tr_n1 <= 12 && sd.tr_n2 % 2 == 0) ? ((size_t)2 << (2 * sd.tr_n1)) : 0); \
        ED_LAUNCH((k6b_canonicalize_nk<NCH_, N1_, N2_>), grid_n, 256, smem_n, sd.tr_ncos, sd.tr_n1, sd.tr_n2, sd.tr_lut6.p, \
                  sd.tr_inv.p, n_hits, sc.words.p, sc.garg.p);                                                               \
      } while (0)
      if (sd.tr_n1 == 6 && sd.tr_n2 == 6 && nch == 6) ED_K6NK(6, 6, 6);
      else if (sd.tr_n1 == 4 && sd.tr_n2 == 4 && nch <= 4) ED_K6NK(4, 4, 4);
      else if (nch <= 4) ED_K6NK(4, 0, 0);
      else if (nch <= 6) ED_K6NK(6, 0, 0);
      else ED_K6NK(8, 0, 0);
#undef ED_K6NK
      }
    }
    else if (sd.tr_on && nch <= 8) {
      const size_t smem = (size_t)2 * 4 * nch * 64 * 8 + (size_t)2 * 4 * sd.tr_n1 * sd.tr_n2 * 4;
      const int nb_ = parent->space.bits;
#define ED_K6TR(NCH_, N1_, N2_)                                                                                               \
      ED_LAUNCH((k6b_canonicalize_tr<NCH_, N1_, N2_>), grid_b, 256, smem, sd.tr_ncos, sd.tr_n1, sd.tr_n2, nb_, sd.tr_lut6.p, \
                sd.tr_inv.p, n_hits, sc.words.p, sc.garg.p)
      const bool unrolled = !getenv("EDCUDA_K6_NOUNROLL") && nb_ == sd.tr_n1 * sd.tr_n2;
      if (unrolled && sd.tr_n1 == 6 && sd.tr_n2 == 6 && nch == 6) ED_K6TR(6, 6, 6);
      else if (unrolled && sd.tr_n1 == 4 && sd.tr_n2 == 4 && nch <= 4) ED_K6TR(4, 4, 4);
      else if (nch <= 4) ED_K6TR(4, 0, 0);
      else if (nch <= 6) ED_K6TR(6, 0, 0);
      else ED_K6TR(8, 0, 0);
#undef ED_K6TR
    }
    else if (nch <= 4) ED_LAUNCH(k6b_canonicalize<4>, grid_b, 256, 0, S.n_ops, lut6, inv, n_hits, sc.words.p, sc.garg.p);
    else if (nch <= 6) ED_LAUNCH(k6b_canonicalize<6>, grid_b, 256, 0, S.n_ops, lut6, inv, n_hits, sc.words.p, sc.garg.p);
    else if (nch <= 8) ED_LAUNCH(k6b_canonicalize<8>, grid_b, 256, 0, S.n_ops, lut6, inv, n_hits, sc.words.p, sc.garg.p);
    else ED_LAUNCH(k6b_canonicalize<11>, grid_b, 256, 0, S.n_ops, lut6, inv, n_hits, sc.words.p, sc.garg.p);
    const RLookupDesc R = rb->rdesc();
    const int grid_l = (int)std::max<int64_t>(1, std::min<int64_t>((n_hits + 255) / 256, (int64_t)sm * 32));
    ED_LAUNCH(k6b_lookup, grid_l, 256, 0, R, n_hits, sc.words.p, sc.hitj.p, sc.horb.p);
  }
  return n_hits;
}

bool ed_apply_reduced_staged_supported(ed_oprep* o) {
  const ed_rbasis* rb = o->rbasis;
  const char* e = getenv("EDCUDA_K6_MIN_ROWS");   // rows below which the simple row-per-thread kernel is used
  const int64_t min_rows = e ? atoll(e) : 2048;
  return rb && rb->symdev.lut6.n > 0 && rb->symdev.n_chunks6 <= 11 && (o->row_hi - o->row_lo) >= min_rows &&
         o->op.n_terms <= 1900;        // the term table is staged in (static-limit) shared memory: 24 B per term
}

void ed_apply_reduced_staged(ed_oprep* o, void* out, const void* x, int side, int accumulate, double* alpha_dot) {
  ed_upload_terms(o);
  ed_rbasis* rb = o->rbasis;
  ed_basis* parent = rb->parent;
  if (parent->kind == ED_BASIS_LIST) parent->materialize();
  const TermsDev& TD = side == ED_SIDE_LEFT ? o->terms_left : o->terms_right;
  K6Terms T{TD.n_terms, TD.mask.p, TD.match.p, TD.target.p, TD.amp.p, TD.is_complex ? 1 : 0};
  const int64_t n_rows = o->row_hi - o->row_lo;
  const RLookupDesc R = rb->rdesc();
  const SymDesc S = rb->symdesc();
  const LookupDesc L = parent->desc();
  K6Scratch& sc = scratch();
  // batch rows so that a batch emits at most ~2^27 hits (upper bound: every term fires)
  const int64_t cap_hits = 1ll << 27;
  const int64_t batch_rows = std::max<int64_t>(4096, cap_hits / std::max(1, T.n_terms / 2));
  const int n_batches = (int)((n_rows + batch_rows - 1) / batch_rows);
  const int sm = ed_sm_count();
  const int grid_c_max = sm * 16;
  if (alpha_dot && sc.partials.n < (size_t)2 * grid_c_max * n_batches) sc.partials.alloc((size_t)2 * grid_c_max * n_batches);
  int slots_used = 0;
  // hit counts per batch of this (side, row range, batching): filled by the first matvec, after which the loop never waits
  // for the device
  std::vector<int64_t>& hits_cache = o->k6_hits[side == ED_SIDE_RIGHT ? 1 : 0];
  if (o->k6_hits_lo != o->row_lo || o->k6_hits_hi != o->row_hi || o->k6_hits_batch != batch_rows) {
    o->k6_hits[0].clear(); o->k6_hits[1].clear();
    o->k6_hits_lo = o->row_lo; o->k6_hits_hi = o->row_hi; o->k6_hits_batch = batch_rows;
  }
  if (hits_cache.size() != (size_t)n_batches) hits_cache.assign((size_t)n_batches, -1);
  // EDCUDA_K6_TIMING=1: per-phase device times of this call on stderr (profiling aid; adds event records only)
  static const bool timing = getenv("EDCUDA_K6_TIMING") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float t_phase[3] = {0.f, 0.f, 0.f};
  if (timing) for (auto& e : ev) ED_CUDA(cudaEventCreate(&e));
  for (int64_t b0 = 0; b0 < n_rows; b0 += batch_rows) {
    if (timing) ED_CUDA(cudaEventRecord(ev[0], ed_stream()));
    const int64_t nb = std::min(batch_rows, n_rows - b0);
    const int64_t row0 = o->row_lo + b0;
    const int64_t n_hits = k6_stage_batch(o, T, row0, nb, sc, timing ? ev[1] : nullptr, &hits_cache[(size_t)(b0 / batch_rows)]);
    if (timing) ED_CUDA(cudaEventRecord(ev[2], ed_stream()));
    const int grid_c = (int)std::max<int64_t>(1, std::min<int64_t>((nb + 127) / 128, (int64_t)grid_c_max));
    ED_LAUNCH(k6c_combine, grid_c, 128, (size_t)3 * T.n_terms * sizeof(uint64_t), T, L, S, R, row0, nb, b0, sc.offs.p, sc.hitj.p, sc.horb.p, sc.garg.p,
              side == ED_SIDE_RIGHT ? 1 : 0, reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), accumulate,
              alpha_dot ? sc.partials.p : nullptr, slots_used);
    slots_used += grid_c;
    if (timing && n_hits > 0) {
      ED_CUDA(cudaEventRecord(ev[3], ed_stream()));
      ED_CUDA(cudaEventSynchronize(ev[3]));
      for (int q = 0; q < 3; ++q) { float ms = 0; ED_CUDA(cudaEventElapsedTime(&ms, ev[q], ev[q + 1])); t_phase[q] += ms; }
    }
  }
  if (timing) {
    fprintf(stderr, "[edcuda] K6 staged: emit %.2f ms, canonicalize %.2f ms, combine %.2f ms (%d batches)\n", t_phase[0], t_phase[1], t_phase[2], n_batches);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  if (alpha_dot) ed_reduce_pairs(sc.partials.p, slots_used, alpha_dot);
}

// raw entries of lines [line0, line0 + n) of a reduced representation for the sparse assembly (sparse.cu)
bool ed_reduced_fill_raw_staged(ed_oprep* o, int side, int64_t line0, int64_t n, const int64_t* raw_offs, int64_t* raw_row, void* raw_val) {
  const ed_rbasis* rbc = o->rbasis;
  if (!rbc || rbc->symdev.lut6.n == 0 || rbc->symdev.n_chunks6 > 11 || o->op.n_terms > 1900 || getenv("EDCUDA_K6_SIMPLE")) return false;
  ed_upload_terms(o);
  ed_rbasis* rb = o->rbasis;
  ed_basis* parent = rb->parent;
  if (parent->kind == ED_BASIS_LIST) parent->materialize();
  const TermsDev& TD = side == ED_SIDE_LEFT ? o->terms_left : o->terms_right;
  K6Terms T{TD.n_terms, TD.mask.p, TD.match.p, TD.target.p, TD.amp.p, TD.is_complex ? 1 : 0};
  const RLookupDesc R = rb->rdesc();
  const SymDesc S = rb->symdesc();
  const LookupDesc L = parent->desc();
  K6Scratch& sc = scratch();
  k6_stage_batch(o, T, line0, n, sc, nullptr);
  const int grid_c = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, (int64_t)ed_sm_count() * 16));
  ED_LAUNCH(k6c_fill_raw, grid_c, 128, 0, T, L, S, R, line0, n, sc.offs.p, sc.hitj.p, sc.horb.p, sc.garg.p, side == ED_SIDE_RIGHT ? 1 : 0,
            raw_offs, raw_row, reinterpret_cast<c128*>(raw_val));
  return true;
}
