// K6: matrix-free apply!/mul! for a ReducedOperatorRepresentation.
//
// Replaces (reference, /root/reference/src):
//   Symmetry/reduced_operator_representation.jl:57-85    get_row_iterator  (H_r[r,c] += a * amp[col_p] / amp[row_p])
//   Symmetry/reduced_operator_representation.jl:88-116   get_column_iterator (conjugated form)
//   Representation/abstract_operator_representation.jl:296-409  apply loops
// The reference reads basis_mapping_index / basis_mapping_amplitude (two arrays of parent dimension).
// Here every generated column word is reduced on the fly: orbit-minimum search over the group in
// registers, character of the minimising element, orbit size of the representative (see
// reduced_map_word in ed_device.cuh).  Diagonal hits (b' == b) skip the search.
#include <algorithm>
#include <cstdlib>

#include "ed_device.cuh"

void ed_reduce_pairs(const double* partials, int n, double* out2);  // apply.cu

template <typename AmpT>
__global__ void __launch_bounds__(128)
k6_apply_reduced(LookupDesc L, SymDesc S, RLookupDesc R, int64_t row_lo, int64_t n_rows, int n_terms,
                 const uint64_t* __restrict__ g_mask, const uint64_t* __restrict__ g_match,
                 const uint64_t* __restrict__ g_target, const AmpT* __restrict__ g_amp, int conj_side,
                 const c128* __restrict__ x, c128* __restrict__ out, int accumulate, double* __restrict__ dot_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_mask = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_match = s_mask + n_terms;
  uint64_t* s_target = s_match + n_terms;
  AmpT* s_amp = reinterpret_cast<AmpT*>(s_target + n_terms);
  for (int t = threadIdx.x; t < n_terms; t += blockDim.x) {
    s_mask[t] = g_mask[t];
    s_match[t] = g_match[t];
    s_target[t] = g_target[t];
    s_amp[t] = g_amp[t];
  }
  __syncthreads();
  double dre = 0.0, dim_ = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row_lo + i;
    const uint64_t b = __ldg(R.words + r);
    c128 a_self = reduced_rep_amp(S, R, r);
    if (conj_side) a_self = cconj(a_self);
    const c128 inv_self = cinv(a_self);
    c128 acc = accumulate ? out[i] : make_c128(0.0, 0.0);
#pragma unroll 1
    for (int t = 0; t < n_terms; ++t) {
      const uint64_t m = s_mask[t];
      if ((b & m) != s_match[t]) continue;
      const uint64_t b2 = (b & ~m) | s_target[t];
      int64_t j;
      c128 a2;
      if (b2 == b) {
        j = r;
        a2 = a_self;  // already conjugated when conj_side
      } else {
        if (!in_basis_dyn(L, b2)) continue;
        j = reduced_map_word(S, R, b2, &a2);
        if (j < 0) continue;
        if (conj_side) a2 = cconj(a2);
      }
      const c128 coef = cmul(cmul(to_c128(s_amp[t]), a2), inv_self);
      fma_acc(acc, coef, ldg_c128(x + j));
    }
    st_val(out + i, acc);
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
      dot_partials[2 * blockIdx.x] = a;
      dot_partials[2 * blockIdx.x + 1] = c;
    }
  }
}

void ed_apply_reduced(ed_oprep* o, void* out, const void* x, int side, int accumulate, double* alpha_dot) {
  // large row ranges: word-parallel orbit sweep (reduced_staged.cu); small ones: the simple row-per-thread kernel below
  static const bool force_simple = getenv("EDCUDA_K6_SIMPLE") != nullptr;
  if (!force_simple && ed_apply_reduced_staged_supported(o)) {
    ed_apply_reduced_staged(o, out, x, side, accumulate, alpha_dot);
    return;
  }
  ed_upload_terms(o);
  ed_rbasis* rb = o->rbasis;
  ed_basis* parent = rb->parent;
  if (parent->kind == ED_BASIS_LIST) parent->materialize();
  const TermsDev& T = side == ED_SIDE_LEFT ? o->terms_left : o->terms_right;
  const int64_t n_rows = o->row_hi - o->row_lo;
  if (n_rows <= 0 || T.n_terms == 0) {
    if (!accumulate && n_rows > 0) ED_CUDA(cudaMemsetAsync(out, 0, (size_t)n_rows * 16, ed_stream()));
    if (alpha_dot) ED_CUDA(cudaMemsetAsync(alpha_dot, 0, 2 * sizeof(double), ed_stream()));
    return;
  }
  const RLookupDesc R = rb->rdesc();
  const int block = 128;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_rows + block - 1) / block, (int64_t)ed_sm_count() * 16));
  DevBuf<double>& partial_buf = ed_scratch<double, 3>();
  double* partials = nullptr;
  if (alpha_dot) {
    if (partial_buf.n < (size_t)2 * grid) partial_buf.alloc((size_t)2 * grid);
    partials = partial_buf.p;
  }
  const size_t smem = (size_t)T.n_terms * (3 * sizeof(uint64_t) + (T.is_complex ? 16 : 8));
  ED_REQUIRE(smem <= 200 * 1024, ED_ERR_UNSUPPORTED, "too many terms for the reduced apply kernel");
  if (T.is_complex) {
    if (smem > 48 * 1024)
      ED_CUDA(cudaFuncSetAttribute(k6_apply_reduced<c128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ED_LAUNCH(k6_apply_reduced<c128>, grid, block, smem, parent->desc(), rb->symdesc(), R, o->row_lo, n_rows, T.n_terms,
              T.mask.p, T.match.p, T.target.p, reinterpret_cast<const c128*>(T.amp.p), side == ED_SIDE_RIGHT ? 1 : 0,
              reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), accumulate, partials);
  } else {
    if (smem > 48 * 1024)
      ED_CUDA(cudaFuncSetAttribute(k6_apply_reduced<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ED_LAUNCH(k6_apply_reduced<double>, grid, block, smem, parent->desc(), rb->symdesc(), R, o->row_lo, n_rows, T.n_terms,
              T.mask.p, T.match.p, T.target.p, T.amp.p, side == ED_SIDE_RIGHT ? 1 : 0,
              reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), accumulate, partials);
  }
  if (alpha_dot) ed_reduce_pairs(partials, grid, alpha_dot);
}
