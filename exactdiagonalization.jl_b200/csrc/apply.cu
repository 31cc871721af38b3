// K2 (generic): matrix-free  out (+)= H * x  /  out (+)= x * H  for any SumOperator on any basis kind.
//
// Replaces (reference, /root/reference/src):
//   Representation/abstract_operator_representation.jl:296-316, 358-378   apply_serial!/apply_parallel!(out, opr, state)
//   Representation/abstract_operator_representation.jl:327-347, 389-409   apply_serial!/apply_parallel!(out, state, opr)
//   Representation/operator_representation.jl:66-103                      get_row_iterator / get_column_iterator
//   Operator/operator_iterator.jl:48-63                                   SumOperator term walk
//   frozensortedarray.jl:29-48                                            basis lookup
// One thread owns one output row (row-owner writes: no atomics, deterministic), walks the term
// table staged in shared memory (all lanes read the same term -> broadcast), and accumulates the
// hits in the reference's term order.  The column is found by the basis' ranking function instead
// of a per-hit binary search whenever the basis is a sector (combinadic / DP rank); a user-supplied
// list keeps the binary search.  This is the general path; the spin-1/2 U(1) fast path is apply_u1.cu.
#include "ed_device.cuh"

struct TermChunkSmem {
  // layout inside dynamic shared memory: mask[n] | match[n] | target[n] | amp[n or 2n]
};

template <typename VecT, typename AmpT, int KIND>
__global__ void __launch_bounds__(256)
k2_apply_generic(LookupDesc L, const uint64_t* __restrict__ row_words, int64_t n_rows, int64_t row_lo,
                 int n_terms, int chunk, const uint64_t* __restrict__ g_mask, const uint64_t* __restrict__ g_match,
                 const uint64_t* __restrict__ g_target, const AmpT* __restrict__ g_amp,
                 const VecT* __restrict__ x, VecT* __restrict__ out, int accumulate, double* __restrict__ dot_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_mask = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_match = s_mask + chunk;
  uint64_t* s_target = s_match + chunk;
  AmpT* s_amp = reinterpret_cast<AmpT*>(s_target + chunk);

  double dre = 0.0, dim_ = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // every thread of the block runs the same number of outer iterations (barriers inside)
  const int64_t n_iter = (n_rows + stride - 1) / stride;
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t i = it * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = i < n_rows;
    uint64_t b = live ? __ldg(row_words + i) : 0ull;
    VecT acc = vzero((VecT*)nullptr);
    if (live && accumulate) acc = out[i];
    for (int t0 = 0; t0 < n_terms; t0 += chunk) {
      const int nt = min(chunk, n_terms - t0);
      if (n_terms > chunk || it == 0) {
        __syncthreads();
        for (int t = threadIdx.x; t < nt; t += blockDim.x) {
          s_mask[t] = g_mask[t0 + t];
          s_match[t] = g_match[t0 + t];
          s_target[t] = g_target[t0 + t];
          s_amp[t] = g_amp[t0 + t];
        }
        __syncthreads();
      }
      if (live) {
#pragma unroll 1
        for (int t = 0; t < nt; ++t) {
          const uint64_t m = s_mask[t];
          if ((b & m) == s_match[t]) {
            const uint64_t b2 = (b & ~m) | s_target[t];
            const int64_t j = rank_word<KIND>(L, b2);
            if (j >= 0) fma_acc(acc, s_amp[t], ldg_val(x + j));
          }
        }
      }
    }
    if (live) {
      st_val(out + i, acc);
      if (dot_partials) dot_acc(dre, dim_, ldg_val(x + row_lo + i), acc);
    }
  }
  if (dot_partials) {
    __shared__ double s_red[2][8];
    dre = warp_sum(dre);
    dim_ = warp_sum(dim_);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
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

// deterministic second stage: one block sums the per-block partials in a fixed order
__global__ void __launch_bounds__(256) k_reduce_pairs(const double* __restrict__ partials, int n, double* __restrict__ out2) {
  __shared__ double s[2][256];
  double a = 0, c = 0;
  for (int i = threadIdx.x; i < n; i += 256) { a += partials[2 * i]; c += partials[2 * i + 1]; }
  s[0][threadIdx.x] = a;
  s[1][threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s[0][threadIdx.x] += s[0][threadIdx.x + o]; s[1][threadIdx.x] += s[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = s[0][0]; out2[1] = s[1][0]; }
}

void ed_reduce_pairs(const double* partials, int n, double* out2) {
  ED_LAUNCH(k_reduce_pairs, 1, 256, 0, partials, n, out2);
}

void ed_upload_terms(ed_oprep* o) {
  if (o->terms_ready) return;
  const ed_operator& op = o->op;
  auto fill = [&](TermsDev& T, bool right) {
    T.n_terms = (int)op.n_terms;
    T.is_complex = op.is_complex;
    T.mask.upload(op.mask);
    T.match.upload(right ? op.col : op.row);
    T.target.upload(right ? op.row : op.col);
    T.amp.upload(op.amp);
  };
  fill(o->terms_left, false);
  fill(o->terms_right, true);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  o->terms_ready = true;
}

static DevBuf<double>& dot_scratch(int n_blocks) {
  DevBuf<double>& buf = ed_scratch<double, 1>();
  if (buf.n < (size_t)2 * n_blocks) buf.alloc((size_t)2 * n_blocks);
  return buf;
}

template <typename VecT, typename AmpT>
static void launch_generic(ed_oprep* o, const TermsDev& T, void* out, const void* x, int accumulate, double* alpha_dot) {
  ed_basis* b = o->basis;
  b->materialize();
  const int64_t n_rows = o->row_hi - o->row_lo;
  if (n_rows <= 0) {
    if (alpha_dot) ED_CUDA(cudaMemsetAsync(alpha_dot, 0, 2 * sizeof(double), ed_stream()));
    return;
  }
  const int block = 256;
  int64_t g = (n_rows + block - 1) / block;
  const int64_t cap = (int64_t)ed_sm_count() * 8;
  const int grid = (int)std::max<int64_t>(1, std::min(g, cap));
  const int chunk = std::max(1, std::min(T.n_terms, 1024));
  const size_t smem = (size_t)chunk * (3 * sizeof(uint64_t) + sizeof(AmpT));
  double* partials = alpha_dot ? dot_scratch(grid).p : nullptr;
  LookupDesc L = b->desc();
  const uint64_t* rw = b->words.p + o->row_lo;
#define ED_GO(KIND)                                                                                              \
  ED_LAUNCH((k2_apply_generic<VecT, AmpT, KIND>), grid, block, smem, L, rw, n_rows, o->row_lo, T.n_terms, chunk, \
            T.mask.p, T.match.p, T.target.p, reinterpret_cast<const AmpT*>(T.amp.p),                             \
            reinterpret_cast<const VecT*>(x), reinterpret_cast<VecT*>(out), accumulate, partials)
  switch (b->kind) {
    case ED_BASIS_LIST: ED_GO(ED_BASIS_LIST); break;
    case ED_BASIS_FULL: ED_GO(ED_BASIS_FULL); break;
    case ED_BASIS_COMBINADIC: ED_GO(ED_BASIS_COMBINADIC); break;
    default: ED_GO(ED_BASIS_DPRANK); break;
  }
#undef ED_GO
  if (alpha_dot) ed_reduce_pairs(partials, grid, alpha_dot);
}

void ed_apply_generic(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot) {
  ed_upload_terms(o);
  const TermsDev& T = side == ED_SIDE_LEFT ? o->terms_left : o->terms_right;
  if (T.n_terms == 0) {
    const int64_t n_rows = o->row_hi - o->row_lo;
    if (!accumulate && n_rows > 0)
      ED_CUDA(cudaMemsetAsync(out, 0, (size_t)n_rows * (dtype == ED_C128 ? 16 : 8), ed_stream()));
    if (alpha_dot) ED_CUDA(cudaMemsetAsync(alpha_dot, 0, 2 * sizeof(double), ed_stream()));
    return;
  }
  if (dtype == ED_F64) {
    ED_REQUIRE(!T.is_complex, ED_ERR_ARGUMENT, "a complex operator representation needs ComplexF64 vectors");
    launch_generic<double, double>(o, T, out, x, accumulate, alpha_dot);
  } else if (!T.is_complex) {
    launch_generic<c128, double>(o, T, out, x, accumulate, alpha_dot);
  } else {
    launch_generic<c128, c128>(o, T, out, x, accumulate, alpha_dot);
  }
}
