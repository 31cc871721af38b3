// K6 (linear): matrix-free apply for a ReducedOperatorRepresentation, one WARP per reduced row.
//
// Same arithmetic as k6_apply_reduced (reduced.cu; reference: Symmetry/reduced_operator_representation.jl:57-116 and
// Symmetry/symmetry_reduce_generic.jl:51-101), organised around one observation: a site permutation acts LINEARLY on
// bit words, g(b xor f) = g(b) xor g(f).  Every off-diagonal term sends the representative b to b xor f_t with a
// fixed flip mask f_t (two bits for a bond), so for each group element the image of the row word g(b) is formed ONCE
// per row (warp-cooperatively: each lane moves the bits it owns through a byte table of target positions that lives in
// shared memory, one REDUX.OR combines them) and every hit of the row only adds xor (1 << tgt[p]) | (1 << tgt[q]).
// The orbit-minimum search for all ~55 hits of a row therefore costs |G| * (1 image + 55 two-bit updates) instead of
// |G| * 55 images, with no scratch memory and no table streaming (the whole |G| x 64-byte table is smem resident).
// Lanes own hits (slot s of lane l = hit number 32 s + l in term order); contributions are summed in hit order by one
// lane, so the result is deterministic.  Row-owner writes.
#include <algorithm>
#include <cstdlib>

#include "ed_device.cuh"

void ed_reduce_pairs(const double* partials, int n, double* out2);  // apply.cu

#define K6L_WARPS 8
#define K6L_SLOTS 4
#define K6L_HCAP 256   // off-diagonal hits of one row that fit the per-warp list

struct K6LParams {
  int n_terms;
  const uint64_t* mask;
  const uint64_t* match;
  const uint64_t* target;
  const double* amp;
  int amp_complex;
  int n_ops, n_bits;
  const uint8_t* tgt_bit;     // [n_ops][64]
  const uint64_t* flipmask;   // [n_ops]
  const int32_t* inverse;     // [n_ops]
  int64_t row_lo, n_rows;
  int conj_side, accumulate;
};

__global__ void __launch_bounds__(K6L_WARPS * 32)
k6l_apply(K6LParams P, LookupDesc L, SymDesc S, RLookupDesc R, const c128* __restrict__ x, c128* __restrict__ out,
          double* __restrict__ dot_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* s_tgt = smem_raw;                                                     // n_ops * 64
  uint64_t* s_flip = reinterpret_cast<uint64_t*>(s_tgt + (size_t)P.n_ops * 64);   // n_ops
  int* s_inv = reinterpret_cast<int*>(s_flip + P.n_ops);                          // n_ops
  c128* s_contrib = reinterpret_cast<c128*>(s_inv + ((P.n_ops + 3) & ~3));        // warps * 128
  uint16_t* s_hits = reinterpret_cast<uint16_t*>(s_contrib + K6L_WARPS * K6L_SLOTS * 32);   // warps * HCAP
  for (int i = threadIdx.x; i < P.n_ops * 64; i += blockDim.x) s_tgt[i] = P.tgt_bit[i];
  for (int i = threadIdx.x; i < P.n_ops; i += blockDim.x) { s_flip[i] = P.flipmask[i]; s_inv[i] = P.inverse[i]; }
  __syncthreads();

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  c128* contrib = s_contrib + wid * (K6L_SLOTS * 32);
  uint16_t* hits = s_hits + wid * K6L_HCAP;
  double dre = 0.0, dim_ = 0.0;
  const int64_t warp0 = (int64_t)blockIdx.x * K6L_WARPS + wid;
  const int64_t nwarps = (int64_t)gridDim.x * K6L_WARPS;
  for (int64_t i = warp0; i < P.n_rows; i += nwarps) {
    const int64_t r = P.row_lo + i;
    const uint64_t b = __ldg(R.words + r);
    c128 a_self = reduced_rep_amp(S, R, r);
    if (P.conj_side) a_self = cconj(a_self);
    const c128 inv_self = cinv(a_self);
    // ---- phase 0: term walk, 32 terms at a time; diagonal amplitudes summed, off-diagonal hits listed in term order
    int nh = 0;
    c128 dsum = make_c128(0.0, 0.0);
    for (int t0 = 0; t0 < P.n_terms; t0 += 32) {
      const int t = t0 + lane;
      bool od = false;
      if (t < P.n_terms) {
        const uint64_t m = __ldg(P.mask + t);
        if ((b & m) == __ldg(P.match + t)) {
          const uint64_t b2 = (b & ~m) | __ldg(P.target + t);
          if (b2 == b) {
            if (P.amp_complex) { dsum.re += __ldg(P.amp + 2 * t); dsum.im += __ldg(P.amp + 2 * t + 1); }
            else dsum.re += __ldg(P.amp + t);
          } else od = true;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, od);
      if (od) {
        const int pos = nh + __popc(bal & lt);
        if (pos < K6L_HCAP) hits[pos] = (uint16_t)t;
      }
      nh += __popc(bal);
    }
    dsum.re = warp_sum(dsum.re);
    dsum.im = warp_sum(dsum.im);
    dsum.re = __shfl_sync(0xffffffffu, dsum.re, 0);
    dsum.im = __shfl_sync(0xffffffffu, dsum.im, 0);
    __syncwarp();
    // diagonal: (a * amp_row) * (1/amp_row) like the reference's row iterator
    c128 acc = cmul(cmul(dsum, a_self), inv_self);
    const c128 xr = ldg_c128(x + r);
    acc = cmul(acc, xr);
    // ---- phase 1: orbit minima of the hits, 128 at a time
    for (int hb = 0; hb < nh; hb += K6L_SLOTS * 32) {
      const int nb = min(K6L_SLOTS * 32, nh - hb);
      const int ns = (nb + 31) >> 5;
      uint64_t best[K6L_SLOTS];
      uint32_t pq[K6L_SLOTS];     // the two flipped bit positions p | q << 8 (0xFFFF..: generic mask, see below)
      uint64_t fmask[K6L_SLOTS];
      int arg[K6L_SLOTS];
      int tid_[K6L_SLOTS];
#pragma unroll
      for (int s = 0; s < K6L_SLOTS; ++s) {
        const int h = hb + s * 32 + lane;
        tid_[s] = (s < ns && h < nh) ? (int)hits[h] : -1;
        fmask[s] = 0; best[s] = ~0ull; arg[s] = 0; pq[s] = 0;
        if (tid_[s] >= 0) {
          fmask[s] = __ldg(P.match + tid_[s]) ^ __ldg(P.target + tid_[s]);
          best[s] = b ^ fmask[s];      // identity element
          const int p = __ffsll((long long)fmask[s]) - 1;
          const int q = 63 - __clzll((long long)fmask[s]);
          pq[s] = (__popcll(fmask[s]) == 2) ? (uint32_t)(p | (q << 8)) : 0xFFFFFFFFu;
        }
      }
#pragma unroll 1
      for (int g = 1; g < P.n_ops; ++g) {
        const uint8_t* tg = s_tgt + g * 64;
        // image of the row word: every lane moves the bits it owns, one REDUX.OR per half
        uint32_t lo = 0, hi = 0;
        {
          const uint32_t tb0 = tg[lane];
          if ((b >> lane) & 1ull) { if (tb0 < 32) lo = 1u << tb0; else hi = 1u << (tb0 - 32); }
          if (P.n_bits > 32) {
            const uint32_t tb1 = tg[lane + 32];
            if ((b >> (lane + 32)) & 1ull) { if (tb1 < 32) lo |= 1u << tb1; else hi |= 1u << (tb1 - 32); }
          }
        }
        lo = __reduce_or_sync(0xffffffffu, lo);
        hi = __reduce_or_sync(0xffffffffu, hi);
        const uint64_t u = (((uint64_t)hi << 32) | lo) ^ s_flip[g];
        const int inv = s_inv[g];
#pragma unroll
        for (int s = 0; s < K6L_SLOTS; ++s) {
          if (s < ns && tid_[s] >= 0) {
            uint64_t pf;
            if (pq[s] != 0xFFFFFFFFu) {
              pf = (1ull << tg[pq[s] & 255u]) | (1ull << tg[(pq[s] >> 8) & 255u]);
            } else {
              pf = 0;
              uint64_t m = fmask[s];
              while (m) { const int bit = __ffsll((long long)m) - 1; m &= m - 1; pf |= 1ull << tg[bit]; }
            }
            const uint64_t cand = u ^ pf;
            if (cand < best[s]) { best[s] = cand; arg[s] = inv; }
            else if (cand == best[s] && inv > arg[s]) arg[s] = inv;
          }
        }
      }
      // ---- phase 2: representative index, coefficient, contribution
#pragma unroll
      for (int s = 0; s < K6L_SLOTS; ++s) {
        c128 cv = make_c128(0.0, 0.0);
        if (s < ns && tid_[s] >= 0) {
          const int t = tid_[s];
          const uint64_t b2 = b ^ fmask[s];
          if (rank_word_dyn(L, b2) >= 0) {                                   // column word inside the parent basis
            const int64_t j = rank_reduced(R, best[s]);
            if (j >= 0) {                                                    // orbit belongs to this irrep
              const double inv_norm = 1.0 / sqrt((double)__ldg(R.orbit_size + j));
              c128 a2 = make_c128(__ldg(S.chi + 2 * arg[s]) * inv_norm, -__ldg(S.chi + 2 * arg[s] + 1) * inv_norm);
              if (P.conj_side) a2 = cconj(a2);
              const c128 a = P.amp_complex ? make_c128(__ldg(P.amp + 2 * t), __ldg(P.amp + 2 * t + 1)) : make_c128(__ldg(P.amp + t), 0.0);
              cv = cmul(cmul(cmul(a, a2), inv_self), ldg_c128(x + j));
            }
          }
        }
        if (s < ns) contrib[s * 32 + lane] = cv;
      }
      __syncwarp();
      if (lane == 0)
        for (int h = 0; h < nb; ++h) acc = cadd(acc, contrib[h]);       // fixed (term) order
      __syncwarp();
    }
    if (lane == 0) {
      if (P.accumulate) acc = cadd(acc, out[i]);
      st_val(out + i, acc);
      if (dot_partials) dot_acc(dre, dim_, xr, acc);
    }
  }
  if (dot_partials) {
    __shared__ double s_red[2][K6L_WARPS];
    if (lane == 0) { s_red[0][wid] = dre; s_red[1][wid] = dim_; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, c = 0;
      for (int w = 0; w < K6L_WARPS; ++w) { a += s_red[0][w]; c += s_red[1][w]; }
      dot_partials[2 * blockIdx.x] = a;
      dot_partials[2 * blockIdx.x + 1] = c;
    }
  }
}

static size_t k6l_smem(int n_ops) {
  return (size_t)n_ops * 64 + (size_t)n_ops * 8 + (size_t)((n_ops + 3) & ~3) * 4 + (size_t)K6L_WARPS * K6L_SLOTS * 32 * 16 +
         (size_t)K6L_WARPS * K6L_HCAP * 2;
}

bool ed_apply_reduced_linear_supported(ed_oprep* o) {
  const ed_rbasis* rb = o->rbasis;
  if (!rb || rb->symdev.tgt_bit.n == 0) return false;
  // experimental: measured 2.1 s per config-4 matvec against 1.4 s for the staged sweep (111 registers, long
  // dependent chains per group element); opt in with EDCUDA_K6_LINEAR=1
  if (!getenv("EDCUDA_K6_LINEAR")) return false;
  if (k6l_smem(rb->symdev.n_ops) > 200 * 1024) return false;
  // every row's off-diagonal hits must fit the per-warp list
  int n_od = 0;
  for (int64_t t = 0; t < o->op.n_terms; ++t) n_od += o->op.row[t] != o->op.col[t];
  return n_od <= K6L_HCAP && o->op.n_terms < 65536;
}

void ed_apply_reduced_linear(ed_oprep* o, void* out, const void* x, int side, int accumulate, double* alpha_dot) {
  ed_upload_terms(o);
  ed_rbasis* rb = o->rbasis;
  ed_basis* parent = rb->parent;
  if (parent->kind == ED_BASIS_LIST) parent->materialize();
  const TermsDev& TD = side == ED_SIDE_LEFT ? o->terms_left : o->terms_right;
  const int64_t n_rows = o->row_hi - o->row_lo;
  K6LParams P;
  P.n_terms = TD.n_terms; P.mask = TD.mask.p; P.match = TD.match.p; P.target = TD.target.p; P.amp = TD.amp.p;
  P.amp_complex = TD.is_complex ? 1 : 0;
  P.n_ops = rb->symdev.n_ops; P.n_bits = parent->space.bits;
  P.tgt_bit = rb->symdev.tgt_bit.p; P.flipmask = rb->symdev.flipmask.p; P.inverse = rb->symdev.inverse.p;
  P.row_lo = o->row_lo; P.n_rows = n_rows; P.conj_side = side == ED_SIDE_RIGHT ? 1 : 0; P.accumulate = accumulate;
  RLookupDesc R;
  R.words = rb->words.p; R.orbit_size = rb->orbit_size.p; R.last_stab = rb->last_stab.p;
  R.bucket_start = rb->bucket_start.p; R.bucket_shift = rb->bucket_shift; R.n_buckets = rb->n_buckets; R.dim = rb->dim;
  const size_t smem = k6l_smem(P.n_ops);
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && configured < smem) {
    ED_CUDA(cudaFuncSetAttribute(k6l_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_rows + K6L_WARPS - 1) / K6L_WARPS, (int64_t)ed_sm_count() * 8));
  static thread_local DevBuf<double> pbuf;
  double* partials = nullptr;
  if (alpha_dot) {
    if (pbuf.n < (size_t)2 * grid) pbuf.alloc((size_t)2 * grid);
    partials = pbuf.p;
  }
  ED_LAUNCH(k6l_apply, grid, K6L_WARPS * 32, smem, P, parent->desc(), rb->symdesc(), R, reinterpret_cast<const c128*>(x),
            reinterpret_cast<c128*>(out), partials);
  if (alpha_dot) ed_reduce_pairs(partials, grid, alpha_dot);
}
