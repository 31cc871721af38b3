// Term walk over one row / column of a plain or reduced operator representation (shared by the
// iterator, Matrix() and sparse() kernels).  Follows operator_representation.jl:66-103 and
// reduced_operator_representation.jl:57-116 of the reference.
#pragma once
#include "ed_device.cuh"

// ------------------------------------------------------------------ generic walker
// Context shared by the iterator / dense / sparse kernels: everything needed to enumerate the
// (index, amplitude) pairs of one row or column of a plain or reduced representation in term order.
struct WalkCtx {
  LookupDesc L;          // parent (or plain) basis
  const uint64_t* words; // plain: basis words; reduced: representative words
  int64_t dim;
  int reduced;
  SymDesc S;
  RLookupDesc R;
  int n_terms;
  const uint64_t* mask;
  const uint64_t* match;
  const uint64_t* target;
  const double* amp;
  int amp_complex;
  int side;  // ED_SIDE_LEFT: row iterator, ED_SIDE_RIGHT: column iterator
};

__device__ __forceinline__ c128 load_amp(const WalkCtx& W, int t) {
  if (W.amp_complex) return make_c128(W.amp[2 * t], W.amp[2 * t + 1]);
  return make_c128(W.amp[t], 0.0);
}

// emit(index0 or -1, amplitude) for every matching term of row/column i (0-based), in term order.
template <typename F>
__device__ __forceinline__ void walk_line(const WalkCtx& W, int64_t i, F&& emit) {
  const uint64_t b = W.words[i];
  c128 inv_self = make_c128(1.0, 0.0);
  c128 a_self = make_c128(1.0, 0.0);
  if (W.reduced) {
    a_self = reduced_rep_amp(W.S, W.R, i);
    if (W.side == ED_SIDE_RIGHT) a_self = cconj(a_self);
    // row: one(S)/ampl_row (reduced_operator_representation.jl:66); column: one(S)/conj(ampl_col) (:98)
    inv_self = cinv(a_self);
  }
  for (int t = 0; t < W.n_terms; ++t) {
    const uint64_t m = W.mask[t];
    if ((b & m) != W.match[t]) continue;
    const uint64_t b2 = (b & ~m) | W.target[t];
    const c128 a = load_amp(W, t);
    if (!W.reduced) {
      emit(rank_word_dyn(W.L, b2), a);
    } else if (b2 == b) {
      emit(i, cmul(cmul(a, a_self), inv_self));   // diagonal hit: the word is the representative itself, no orbit scan
    } else {
      if (rank_word_dyn(W.L, b2) < 0) { emit((int64_t)-1, a); continue; }  // not in the parent basis (:75-76)
      c128 a2;
      int64_t j = reduced_map_word(W.S, W.R, b2, &a2);
      if (j < 0) { emit((int64_t)-1, a); continue; }                        // orbit not in this irrep (:77-78)
      if (W.side == ED_SIDE_RIGHT) a2 = cconj(a2);
      emit(j, cmul(cmul(a, a2), inv_self));                                   // ampl * ampl_col * inv_ampl_row (:80)
    }
  }
}


WalkCtx ed_make_walk_ctx(ed_oprep* o, int side);
