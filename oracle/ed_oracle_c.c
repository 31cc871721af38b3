/* CPU ORACLE (C twin) -- test infrastructure and timed CPU baseline only, never the product path.
 *
 * Plain C + OpenMP restatement of the reference's matrix-free apply (ExactDiagonalization.jl v0.14.3,
 * /root/reference/src), kept algorithmically identical so that its timing stands in for Julia's
 * apply_parallel! (Julia itself is not installed in this image):
 *   Representation/abstract_operator_representation.jl:358-378  apply_parallel!(out, opr, state)
 *       Threads.@threads over rows (static contiguous partition)  ->  #pragma omp for schedule(static)
 *   Representation/abstract_operator_representation.jl:389-409  apply_parallel!(out, state, opr)
 *   Representation/operator_representation.jl:66-103            get_row_iterator / get_column_iterator
 *   Operator/operator_iterator.jl:48-63                         term walk, in term order
 *   frozensortedarray.jl:29-48                                  fs_index = searchsortedfirst + equality
 * Parity status: pinned -- tests/test_oracle_c.py checks it against oracle/ed_oracle.py, which is pinned
 * against the reference's golden vectors.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* searchsortedfirst(keys, key): first index with keys[i] >= key (0-based; n if none) */
static inline int64_t searchsortedfirst(const uint64_t* keys, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = lo + ((hi - lo) >> 1);
    if (keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* fs_index (frozensortedarray.jl:29-32): 1-based index or 0 */
static inline int64_t fs_index(const uint64_t* keys, int64_t n, uint64_t key) {
  int64_t idx = searchsortedfirst(keys, n, key);
  return (idx < n && keys[idx] == key) ? idx + 1 : 0;
}

int oc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Ascending list of all n-bit words with n_set bits set (the Sz sector of a spin-1/2 space:
 * result of hs_get_basis_list, hilbert_space_representation.jl:127-206; generated here by
 * next-combination stepping -- setup for the timed apply, not itself timed).  Returns the count. */
int64_t oc_basis_fixed_popcount(int n_bits, int n_set, uint64_t* out, int64_t cap) {
  if (n_set < 0 || n_set > n_bits || n_bits > 63) return 0;
  int64_t k = 0;
  if (n_set == 0) { if (cap > 0) out[0] = 0; return 1; }
  uint64_t v = (1ull << n_set) - 1ull;
  const uint64_t limit = 1ull << n_bits;
  while (v < limit) {
    if (k < cap) out[k] = v;
    ++k;
    uint64_t t = v | (v - 1);
    uint64_t w = (t + 1) | (((~t & -~t) - 1) >> (__builtin_ctzll(v) + 1));
    if (w <= v) break;
    v = w;
  }
  return k;
}

/* sector basis by the reference's own DP (hilbert_space_representation.jl:146-206) restricted to one
 * integer quantum number per state: per-site lists keyed by the partial sum, then concatenation+sort of
 * the single requested sector (the reference merges sorted lists; with one sector that is the list itself).
 * qn[site*max_states + state]; returns count, words ascending. Small sizes only (validation). */
int64_t oc_sector_basis_dp(int n_sites, const int* n_states, const int* bitoffset, const int64_t* qn, int max_states,
                           int64_t target, uint64_t* out, int64_t cap) {
  /* partial sums range */
  int64_t lo = 0, hi = 0;
  for (int i = 0; i < n_sites; ++i) {
    int64_t mn = qn[i * max_states], mx = mn;
    for (int s = 1; s < n_states[i]; ++s) { int64_t q = qn[i * max_states + s]; if (q < mn) mn = q; if (q > mx) mx = q; }
    lo += mn; hi += mx;
  }
  if (lo > 0) lo = 0;
  if (hi < 0) hi = 0;
  const int64_t nq = hi - lo + 1;
  uint64_t** lists = (uint64_t**)calloc((size_t)nq, sizeof(uint64_t*));
  int64_t* counts = (int64_t*)calloc((size_t)nq, sizeof(int64_t));
  lists[0 - lo] = (uint64_t*)malloc(sizeof(uint64_t));
  lists[0 - lo][0] = 0; counts[0 - lo] = 1;
  for (int i = 0; i < n_sites; ++i) {
    uint64_t** nl = (uint64_t**)calloc((size_t)nq, sizeof(uint64_t*));
    int64_t* nc = (int64_t*)calloc((size_t)nq, sizeof(int64_t));
    for (int64_t q = lo; q <= hi; ++q) {
      int64_t total = 0;
      for (int s = 0; s < n_states[i]; ++s) {
        int64_t qp = q - qn[i * max_states + s];
        if (qp >= lo && qp <= hi) total += counts[qp - lo];
      }
      if (!total) continue;
      nl[q - lo] = (uint64_t*)malloc((size_t)total * sizeof(uint64_t));
      int64_t at = 0;
      for (int s = 0; s < n_states[i]; ++s) {        /* states in order: appended lists stay ascending (:182-189) */
        int64_t qp = q - qn[i * max_states + s];
        if (qp < lo || qp > hi) continue;
        for (int64_t k = 0; k < counts[qp - lo]; ++k) nl[q - lo][at++] = lists[qp - lo][k] | ((uint64_t)s << bitoffset[i]);
      }
      nc[q - lo] = total;
    }
    for (int64_t q = 0; q < nq; ++q) free(lists[q]);
    free(lists); free(counts);
    lists = nl; counts = nc;
  }
  int64_t n = 0;
  if (target >= lo && target <= hi) {
    n = counts[target - lo];
    for (int64_t k = 0; k < n && k < cap; ++k) out[k] = lists[target - lo][k];
  }
  for (int64_t q = 0; q < nq; ++q) free(lists[q]);
  free(lists); free(counts);
  return n;
}

/* apply_parallel!: out[i - row_lo] += sum_hits amp * x[col]   for rows i in [row_lo, row_hi) (0-based).
 * side = 0: row walk (match bitrow, emit bitcol); side = 1: column walk (match bitcol, emit bitrow).
 * vec_complex / amp_complex select Float64 or ComplexF64 (interleaved) storage. */
void oc_apply(const uint64_t* basis, int64_t dim, int64_t n_terms, const uint64_t* mask, const uint64_t* bitrow,
              const uint64_t* bitcol, const double* amp, int amp_complex, const double* x, double* out,
              int vec_complex, int64_t row_lo, int64_t row_hi, int side) {
  const uint64_t* match = side == 0 ? bitrow : bitcol;
  const uint64_t* target = side == 0 ? bitcol : bitrow;
#pragma omp parallel for schedule(static)
  for (int64_t irow = row_lo; irow < row_hi; ++irow) {
    const uint64_t brow = basis[irow];
    for (int64_t t = 0; t < n_terms; ++t) {
      if ((brow & mask[t]) != match[t]) continue;
      const uint64_t bcol = (brow & ~mask[t]) | target[t];
      const int64_t icol = fs_index(basis, dim, bcol);   /* get(basis_lookup, bcol, -1) */
      if (icol < 1) continue;                            /* 1 <= icol <= ncols */
      if (!vec_complex) {
        out[irow - row_lo] += amp[t] * x[icol - 1];
      } else {
        const double xr = x[2 * (icol - 1)], xi = x[2 * (icol - 1) + 1];
        const double ar = amp_complex ? amp[2 * t] : amp[t], ai = amp_complex ? amp[2 * t + 1] : 0.0;
        out[2 * (irow - row_lo)] += ar * xr - ai * xi;
        out[2 * (irow - row_lo) + 1] += ar * xi + ai * xr;
      }
    }
  }
}

/* number of (valid) hits of rows [row_lo,row_hi): the nnz_eff counter used by the bench */
int64_t oc_count_hits(const uint64_t* basis, int64_t dim, int64_t n_terms, const uint64_t* mask, const uint64_t* bitrow,
                      const uint64_t* bitcol, int64_t row_lo, int64_t row_hi) {
  int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int64_t irow = row_lo; irow < row_hi; ++irow) {
    const uint64_t brow = basis[irow];
    for (int64_t t = 0; t < n_terms; ++t) {
      if ((brow & mask[t]) != bitrow[t]) continue;
      if (fs_index(basis, dim, (brow & ~mask[t]) | bitcol[t]) >= 1) ++total;
    }
  }
  return total;
}

/* symmetry_apply for a site permutation on 1-bit sites (symmetry_apply.jl:82-92) */
static inline uint64_t perm_apply(const int* map, int n_sites, uint64_t b) {
  uint64_t out = 0;
  for (int i = 0; i < n_sites; ++i) out |= ((b >> i) & 1ull) << map[i];
  return out;
}

/* Reduced apply on the fly (CPU stand-in for config 4, where the reference's per-parent-state arrays do not
 * fit): rows [row_lo,row_hi) of out += H_r x following reduced_operator_representation.jl:57-85 with
 * amp[g(r)] = conj(chi_g)/sqrt(N_r) recomputed by an orbit scan (symmetry_reduce_generic.jl:51-86).
 * rbasis ascending representatives, orbit[r] = N_r.  1-bit sites, no bit flips.  x/out complex interleaved. */
void oc_apply_reduced_onthefly(const uint64_t* rbasis, const int32_t* orbit, int64_t rdim, int n_sites, int n_ops,
                               const int* perms, const double* chi, int64_t n_terms, const uint64_t* mask,
                               const uint64_t* bitrow, const uint64_t* bitcol, const double* amp, const double* x,
                               double* out, int64_t row_lo, int64_t row_hi) {
#pragma omp parallel for schedule(static)
  for (int64_t ir = row_lo; ir < row_hi; ++ir) {
    const uint64_t brow = rbasis[ir];
    double accr = 0.0, acci = 0.0;
    for (int64_t t = 0; t < n_terms; ++t) {
      if ((brow & mask[t]) != bitrow[t]) continue;
      const uint64_t bcol = (brow & ~mask[t]) | bitcol[t];
      uint64_t best = bcol; int bestg = 0;
      for (int g = 1; g < n_ops; ++g) {
        uint64_t im = perm_apply(perms + (size_t)g * n_sites, n_sites, bcol);
        if (im < best) { best = im; bestg = g; }
      }
      const int64_t jc = fs_index(rbasis, rdim, best);
      if (jc < 1) continue;
      /* bcol = g^-1(best): amplitude conj(chi_{g^-1}) = chi_g for unitary one-dimensional characters */
      const double s = __builtin_sqrt((double)orbit[ir] / (double)orbit[jc - 1]);
      const double cr = chi[2 * bestg] * s * amp[t], ci = chi[2 * bestg + 1] * s * amp[t];
      const double xr = x[2 * (jc - 1)], xi = x[2 * (jc - 1) + 1];
      accr += cr * xr - ci * xi;
      acci += cr * xi + ci * xr;
    }
    out[2 * (ir - row_lo)] += accr;
    out[2 * (ir - row_lo) + 1] += acci;
  }
}
