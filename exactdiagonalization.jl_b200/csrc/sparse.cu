// K3/K4: sparse(opr; tol) as a count-then-fill CSC assembly on device.
//
// Replaces (reference, /root/reference/src):
//   Representation/abstract_operator_representation.jl:136-142   sparse (tol default sqrt(eps))
//   Representation/abstract_operator_representation.jl:145-169   sparse_serial
//   Representation/abstract_operator_representation.jl:172-204   sparse_parallel
//   util.jl:88-94                                                 choptol! (strict |v| < tol)
// Reference: per column a Dict accumulates duplicates in term order, entries with |v| < tol are deleted,
// the rest sorted by row.  Here:
//   pass A  (count)  : per column, number of matching terms  -> exclusive scan -> scratch offsets
//   pass B  (fill)   : per column, (row, amplitude) of every hit in term order into scratch
//   pass C  (merge)  : per column, stable insertion sort by row, sum equal rows in term order
//                      (same association as the Dict), chop, count survivors -> scan -> colptr
//   pass D  (gather) : survivors to rowval / nzval
// Output is CSC with 1-based Int64 colptr/rowval like SparseMatrixCSC{S,Int}.
#include <algorithm>
#include <cub/cub.cuh>

#include "ed_walk.cuh"

__global__ void __launch_bounds__(128)
k3_count_matches(WalkCtx W, int64_t col_lo, int64_t n_cols, int64_t* __restrict__ counts) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_cols; k += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t b = W.words[col_lo + k];
    int64_t c = 0;
    for (int t = 0; t < W.n_terms; ++t) c += ((b & W.mask[t]) == W.match[t]) ? 1 : 0;
    counts[k] = c;
  }
}

__global__ void __launch_bounds__(128)
k4_fill_raw(WalkCtx W, int64_t col_lo, int64_t n_cols, const int64_t* __restrict__ offs,
            int64_t* __restrict__ raw_row, c128* __restrict__ raw_val) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_cols; k += (int64_t)gridDim.x * blockDim.x) {
    int64_t at = offs[k];
    walk_line(W, col_lo + k, [&](int64_t i, c128 a) {
      raw_row[at] = i;  // -1 for misses, dropped in the merge pass
      raw_val[at] = a;
      ++at;
    });
  }
}

// per column: stable insertion sort by row (misses last), merge, chop; survivors compacted to the
// front of the column's scratch segment; kept[k] = number of survivors.
__global__ void __launch_bounds__(128)
k4_merge_chop(int64_t n_cols, const int64_t* __restrict__ offs, int64_t* __restrict__ raw_row,
              c128* __restrict__ raw_val, double tol, int is_complex, int64_t* __restrict__ kept) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_cols; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = offs[k], hi = offs[k + 1];
    const uint64_t BIG = ~0ull;
    for (int64_t p = lo + 1; p < hi; ++p) {
      int64_t r = raw_row[p];
      c128 v = raw_val[p];
      uint64_t key = r < 0 ? BIG : (uint64_t)r;
      int64_t q = p - 1;
      while (q >= lo) {
        int64_t rq = raw_row[q];
        uint64_t kq = rq < 0 ? BIG : (uint64_t)rq;
        if (kq <= key) break;  // stable: equal rows keep term order
        raw_row[q + 1] = rq;
        raw_val[q + 1] = raw_val[q];
        --q;
      }
      raw_row[q + 1] = r;
      raw_val[q + 1] = v;
    }
    int64_t w = lo, p = lo;
    while (p < hi && raw_row[p] >= 0) {
      const int64_t r = raw_row[p];
      c128 s = raw_val[p];
      ++p;
      while (p < hi && raw_row[p] == r) {
        s = cadd(s, raw_val[p]);  // colvec[irow] = get(colvec, irow, 0) + ampl, in term order
        ++p;
      }
      const double mag = is_complex ? hypot(s.re, s.im) : fabs(s.re);
      if (!(mag < tol)) {  // choptol! deletes abs(v) < tol
        raw_row[w] = r;
        raw_val[w] = s;
        ++w;
      }
    }
    kept[k] = w - lo;
  }
}

// The same merge with one WARP per column: the column's raw entries are ranked in shared memory (stable: equal rows keep
// their term order), every run of equal rows is summed front to back by the lane that holds its head -- the same addition
// order as the Dict of the reference and as k4_merge_chop, so the values are bit-identical -- chopped and compacted with
// warp ballots.  O(n^2 / 32) comparisons per lane instead of a thread-serial insertion sort through global memory;
// columns with more than K4_MAXN raw entries take the serial path on lane 0.
#define K4_MAXN 256
__global__ void __launch_bounds__(128)
k4_merge_chop_warp(int64_t n_cols, const int64_t* __restrict__ offs, int64_t* __restrict__ raw_row,
                   c128* __restrict__ raw_val, double tol, int is_complex, int64_t* __restrict__ kept) {
  extern __shared__ __align__(16) unsigned char k4_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t* skey = reinterpret_cast<uint64_t*>(k4_smem) + (size_t)wid * 2 * K4_MAXN;
  uint64_t* okey = skey + K4_MAXN;
  c128* sval = reinterpret_cast<c128*>(reinterpret_cast<uint64_t*>(k4_smem) + (size_t)(blockDim.x >> 5) * 2 * K4_MAXN) + (size_t)wid * 2 * K4_MAXN;
  c128* oval = sval + K4_MAXN;
  const uint64_t BIG = ~0ull;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp0; k < n_cols; k += nwarps) {
    const int64_t lo = offs[k], hi = offs[k + 1];
    const int n = (int)(hi - lo);
    if (n > K4_MAXN) {                     // rare: serial path (same algorithm as k4_merge_chop)
      if (lane == 0) {
        for (int64_t p = lo + 1; p < hi; ++p) {
          int64_t r = raw_row[p];
          c128 v = raw_val[p];
          uint64_t key = r < 0 ? BIG : (uint64_t)r;
          int64_t q = p - 1;
          while (q >= lo) {
            int64_t rq = raw_row[q];
            uint64_t kq = rq < 0 ? BIG : (uint64_t)rq;
            if (kq <= key) break;
            raw_row[q + 1] = rq;
            raw_val[q + 1] = raw_val[q];
            --q;
          }
          raw_row[q + 1] = r;
          raw_val[q + 1] = v;
        }
        int64_t w = lo, p = lo;
        while (p < hi && raw_row[p] >= 0) {
          const int64_t r = raw_row[p];
          c128 s = raw_val[p];
          ++p;
          while (p < hi && raw_row[p] == r) { s = cadd(s, raw_val[p]); ++p; }
          const double mag = is_complex ? hypot(s.re, s.im) : fabs(s.re);
          if (!(mag < tol)) { raw_row[w] = r; raw_val[w] = s; ++w; }
        }
        kept[k] = w - lo;
      }
      __syncwarp();
      continue;
    }
    for (int i = lane; i < n; i += 32) {
      const int64_t r = raw_row[lo + i];
      skey[i] = r < 0 ? BIG : (uint64_t)r;
      sval[i] = raw_val[lo + i];
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {       // stable rank
      const uint64_t key = skey[i];
      int r = 0;
      for (int j = 0; j < n; ++j) {
        const uint64_t kj = skey[j];
        r += (kj < key || (kj == key && j < i)) ? 1 : 0;
      }
      okey[r] = key;
      oval[r] = sval[i];
    }
    __syncwarp();
    int w = 0;                                  // survivors written so far (warp-uniform)
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      bool keep = false;
      uint64_t key = BIG;
      c128 sum = make_c128(0.0, 0.0);
      if (i < n) {
        key = okey[i];
        if (key != BIG && (i == 0 || okey[i - 1] != key)) {     // head of a run of equal rows
          sum = oval[i];
          for (int j = i + 1; j < n && okey[j] == key; ++j) sum = cadd(sum, oval[j]);    // term order
          const double mag = is_complex ? hypot(sum.re, sum.im) : fabs(sum.re);
          keep = !(mag < tol);                                    // choptol! deletes abs(v) < tol
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int at = w + __popc(m & ((1u << lane) - 1u));
        raw_row[lo + at] = (int64_t)key;
        raw_val[lo + at] = sum;
      }
      w += __popc(m);
    }
    if (lane == 0) kept[k] = w;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128)
k4_gather(int64_t n_cols, const int64_t* __restrict__ offs, const int64_t* __restrict__ raw_row,
          const c128* __restrict__ raw_val, const int64_t* __restrict__ out_offs /* 0-based, relative */,
          int64_t out_base, int64_t* __restrict__ rowval, double* __restrict__ nzval, int is_complex) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_cols; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = out_offs[k + 1] - out_offs[k];
    const int64_t src = offs[k];
    const int64_t dst = out_base + out_offs[k];
    for (int64_t e = 0; e < n; ++e) {
      rowval[dst + e] = raw_row[src + e] + 1;
      if (is_complex) { nzval[2 * (dst + e)] = raw_val[src + e].re; nzval[2 * (dst + e) + 1] = raw_val[src + e].im; }
      else nzval[dst + e] = raw_val[src + e].re;
    }
  }
}

__global__ void k_colptr(int64_t n_cols, const int64_t* __restrict__ out_offs, int64_t out_base, int64_t* __restrict__ colptr) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k <= n_cols; k += (int64_t)gridDim.x * blockDim.x)
    colptr[k] = out_base + out_offs[k] + 1;
}

static void exclusive_scan(const int64_t* in, int64_t* out, int64_t n, DevBuf<unsigned char>& tmp) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, ed_stream());
  if (tmp.n < bytes) tmp.alloc(bytes);
  cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, ed_stream());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
}

// Assemble lines [lo, hi) of the representation (columns for side = RIGHT, rows for side = LEFT): merged, chopped,
// index-sorted entries.  ptr has hi-lo+1 entries (1-based positions like colptr), idx is 1-based.
static void assemble_lines(ed_oprep* o, double tol, int side, int64_t lo, int64_t hi, DevBuf<int64_t>& ptr,
                           DevBuf<int64_t>& idx, DevBuf<double>& val, int64_t& nnz_out) {
  const int64_t nlines = hi - lo;
  const int cplx = o->is_complex ? 1 : 0;
  WalkCtx W = ed_make_walk_ctx(o, side);
  ptr.alloc((size_t)nlines + 1);
  // lines are processed in batches so the scratch stays bounded (~2^28 raw entries)
  int64_t BATCH = 1ll << 22;
  if (W.n_terms > 64) BATCH = std::max<int64_t>(1 << 16, (1ll << 28) / W.n_terms);
  DevBuf<unsigned char> tmp;
  DevBuf<int64_t> counts, offs, kept, out_offs;
  DevBuf<int64_t> raw_row;
  DevBuf<c128> raw_val;
  std::vector<DevBuf<int64_t>> rows_parts;
  std::vector<DevBuf<double>> vals_parts;
  std::vector<int64_t> part_nnz;
  int64_t nnz_total = 0;
  for (int64_t c0 = 0; c0 < nlines; c0 += BATCH) {
    const int64_t nc = std::min(BATCH, nlines - c0);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nc + 127) / 128, (int64_t)ed_sm_count() * 16));
    counts.alloc((size_t)nc + 1);
    offs.alloc((size_t)nc + 1);
    kept.alloc((size_t)nc + 1);
    out_offs.alloc((size_t)nc + 1);
    ED_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)(nc + 1) * sizeof(int64_t), ed_stream()));
    ED_CUDA(cudaMemsetAsync(kept.p, 0, (size_t)(nc + 1) * sizeof(int64_t), ed_stream()));
    ED_LAUNCH(k3_count_matches, grid, 128, 0, W, lo + c0, nc, counts.p);
    exclusive_scan(counts.p, offs.p, nc + 1, tmp);
    int64_t raw_total = 0;
    ED_CUDA(cudaMemcpyAsync(&raw_total, offs.p + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, ed_stream()));
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    raw_row.alloc((size_t)std::max<int64_t>(raw_total, 1));
    raw_val.alloc((size_t)std::max<int64_t>(raw_total, 1));
    // reduced representations with enough lines: orbit minima from the word-parallel staged sweep (reduced_staged.cu)
    const int64_t staged_min = getenv("EDCUDA_K6_MIN_ROWS") ? atoll(getenv("EDCUDA_K6_MIN_ROWS")) : 2048;
    if (!(o->rbasis && nc >= staged_min && ed_reduced_fill_raw_staged(o, side, lo + c0, nc, offs.p, raw_row.p, raw_val.p)))
      ED_LAUNCH(k4_fill_raw, grid, 128, 0, W, lo + c0, nc, offs.p, raw_row.p, raw_val.p);
    if (getenv("EDCUDA_SPARSE_SERIAL_MERGE")) {
      ED_LAUNCH(k4_merge_chop, grid, 128, 0, nc, offs.p, raw_row.p, raw_val.p, tol, cplx, kept.p);
    } else {       // warp per column, 4 warps per CTA, 12 KB of shared memory per warp
      const int grid_w = (int)std::max<int64_t>(1, std::min<int64_t>((nc + 3) / 4, (int64_t)ed_sm_count() * 16));
      ED_LAUNCH(k4_merge_chop_warp, grid_w, 128, (size_t)4 * 2 * K4_MAXN * (sizeof(uint64_t) + sizeof(c128)), nc, offs.p, raw_row.p, raw_val.p, tol, cplx, kept.p);
    }
    exclusive_scan(kept.p, out_offs.p, nc + 1, tmp);
    int64_t nnz_b = 0;
    ED_CUDA(cudaMemcpyAsync(&nnz_b, out_offs.p + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, ed_stream()));
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    rows_parts.emplace_back((size_t)std::max<int64_t>(nnz_b, 1));
    vals_parts.emplace_back((size_t)std::max<int64_t>(nnz_b, 1) * (cplx ? 2 : 1));
    ED_LAUNCH(k4_gather, grid, 128, 0, nc, offs.p, raw_row.p, raw_val.p, out_offs.p, (int64_t)0, rows_parts.back().p,
              vals_parts.back().p, cplx);
    ED_LAUNCH(k_colptr, grid, 128, 0, nc, out_offs.p, nnz_total, ptr.p + c0);
    part_nnz.push_back(nnz_b);
    nnz_total += nnz_b;
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
  }
  if (nlines == 0) {
    int64_t one = 1;
    ptr.upload(&one, 1);
  }
  raw_row.release();
  raw_val.release();
  idx.alloc((size_t)std::max<int64_t>(nnz_total, 1));
  val.alloc((size_t)std::max<int64_t>(nnz_total, 1) * (cplx ? 2 : 1));
  int64_t at = 0;
  for (size_t p = 0; p < part_nnz.size(); ++p) {
    if (part_nnz[p]) {
      ED_CUDA(cudaMemcpyAsync(idx.p + at, rows_parts[p].p, (size_t)part_nnz[p] * sizeof(int64_t), cudaMemcpyDeviceToDevice, ed_stream()));
      ED_CUDA(cudaMemcpyAsync(val.p + at * (cplx ? 2 : 1), vals_parts[p].p, (size_t)part_nnz[p] * (cplx ? 16 : 8), cudaMemcpyDeviceToDevice, ed_stream()));
    }
    at += part_nnz[p];
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    rows_parts[p].release();
    vals_parts[p].release();
  }
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  nnz_out = nnz_total;
}

void ed_sparse_assemble(ed_oprep* o, double tol) {
  assemble_lines(o, tol, ED_SIDE_RIGHT, 0, o->dim, o->sp_colptr, o->sp_rowval, o->sp_nzval, o->sp_nnz);  // column iterator
}

// ------------------------------------------------------------------ cached CSR + SpMV
// The assembled rows [row_lo, row_hi) of the representation kept on device in a compact form (0-based int32 columns,
// real values when every imaginary part is exactly zero); repeated matvecs then run as a bandwidth-bound SpMV
// instead of redoing the term walk / orbit searches.  Entries are NOT chopped (tol = 0), so the cached apply agrees
// with the matrix-free one to rounding.
__global__ void __launch_bounds__(256)
k_csr_compact(int64_t n_lines, int64_t nnz, const int64_t* __restrict__ ptr1, const int64_t* __restrict__ idx1,
              const double* __restrict__ val, int val_complex, int64_t* __restrict__ rowptr0, int32_t* __restrict__ col0,
              int* __restrict__ any_imag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int flag = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += stride) {
    col0[i] = (int32_t)(idx1[i] - 1);
    if (val_complex && val[2 * i + 1] != 0.0) flag = 1;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n_lines; i += stride) rowptr0[i] = ptr1[i] - 1;
  if (flag) atomicExch(any_imag, 1);
}

__global__ void __launch_bounds__(256) k_take_real(int64_t nnz, const double* __restrict__ cval, double* __restrict__ rval) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) rval[i] = cval[2 * i];
}

// matrix values: ComplexF64, Float64, or a 16-bit code into a dictionary of the distinct values (CsrCache::coded)
__device__ __forceinline__ c128 ld_mat(const c128* v, int64_t i, const double*) { return ldg_c128(v + i); }
__device__ __forceinline__ double ld_mat(const double* v, int64_t i, const double*) { return __ldg(v + i); }
__device__ __forceinline__ double ld_mat(const uint16_t* v, int64_t i, const double* lut) { return __ldg(lut + __ldg(v + i)); }
__device__ __forceinline__ void mat_fma(double& acc, double a, double x) { acc = fma(a, x, acc); }
__device__ __forceinline__ void mat_fma(c128& acc, double a, c128 x) { acc.re = fma(a, x.re, acc.re); acc.im = fma(a, x.im, acc.im); }
__device__ __forceinline__ void mat_fma(c128& acc, c128 a, c128 x) { fma_acc(acc, a, x); }

// one warp per row: lanes stride the row's entries, shuffle reduction (fixed tree -> deterministic)
template <typename VecT, typename ValT>
__global__ void __launch_bounds__(256)
k_spmv_csr(int64_t n_rows, int64_t row_lo, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
           const ValT* __restrict__ val, const double* __restrict__ lut, const VecT* __restrict__ x, VecT* __restrict__ y, int accumulate,
           double* __restrict__ dot_partials) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dre = 0.0, dim_ = 0.0;
  for (int64_t r = warp0; r < n_rows; r += nwarps) {
    const int64_t b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    VecT acc = vzero((VecT*)nullptr);
    for (int64_t p = b + lane; p < e; p += 32) mat_fma(acc, ld_mat(val, p, lut), ldg_val(x + __ldg(col + p)));
    if (sizeof(VecT) == 16) {
      c128* a = reinterpret_cast<c128*>(&acc);
      a->re = warp_sum(a->re);
      a->im = warp_sum(a->im);
    } else {
      double* a = reinterpret_cast<double*>(&acc);
      *a = warp_sum(*a);
    }
    if (lane == 0) {
      if (accumulate) acc = (sizeof(VecT) == 16) ? acc : acc;
      VecT out = acc;
      if (accumulate) {
        VecT old = y[r];
        fma_acc(out, 1.0, old);
      }
      st_val(y + r, out);
      if (dot_partials) dot_acc(dre, dim_, ldg_val(x + row_lo + r), out);
    }
  }
  if (dot_partials) {
    __shared__ double s_red[2][8];
    dre = warp_sum(dre);
    dim_ = warp_sum(dim_);
    if (lane == 0) { s_red[0][threadIdx.x >> 5] = dre; s_red[1][threadIdx.x >> 5] = dim_; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, c = 0;
      for (int w = 0; w < 8; ++w) { a += s_red[0][w]; c += s_red[1][w]; }
      dot_partials[2 * blockIdx.x] = a;
      dot_partials[2 * blockIdx.x + 1] = c;
    }
  }
}

// ---- column-blocked SpMV -----------------------------------------------------------------------------------------
// A warp-per-row SpMV gathers x from all over the vector; once x is larger than the L2 (6x6 triangular: 336 MB against
// 126 MB) nearly every gather is a 32-byte DRAM sector and the kernel is bound by random DRAM accesses.  The cached
// matrix is therefore regrouped into column blocks sized to a fraction of the L2: pass b streams block b of the matrix
// (evict-first loads) and gathers only from x[b * W, (b + 1) * W), which stays L2-resident for the whole pass; y is
// accumulated across the passes.  Rows are short inside a block, so LANES (4) lanes share a row.  The gain is modest
// (9.64 -> 8.37 ms): ncu still sees a 39 % L2 hit rate and 39 GB of DRAM reads per matvec (14.3 GB of matrix), i.e. the
// gather stream keeps about one L2 partition's worth of x, and smaller windows cost more in per-pass overhead.
__global__ void __launch_bounds__(256)
k_blk_count(int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n_blocks, int64_t block_cols,
            uint32_t* __restrict__ cnt) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = rowptr[r], e = rowptr[r + 1];
    int64_t prev = b;
    for (int k = 0; k < n_blocks; ++k) {
      // first entry of the row whose column is >= (k + 1) * block_cols (columns ascend within a row)
      const int64_t bound = (int64_t)(k + 1) * block_cols;
      int64_t lo = prev, hi = e;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)col[mid] < bound) lo = mid + 1; else hi = mid; }
      cnt[(size_t)k * (n_rows + 1) + r] = (uint32_t)(lo - prev);
      prev = lo;
    }
  }
}

template <typename ValT>
__global__ void __launch_bounds__(256)
k_blk_fill(int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const ValT* __restrict__ val,
           int n_blocks, int64_t block_cols, const uint32_t* __restrict__ blk_rowptr, const int64_t* __restrict__ blk_base,
           int32_t* __restrict__ col_out, ValT* __restrict__ val_out) {
  // one warp per row: every lane moves the entries p = b + lane, b + lane + 32, ...
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += nwarps) {
    const int64_t b = rowptr[r], e = rowptr[r + 1];
    int64_t first_in_row = b;      // position (in the row-major matrix) of the first entry of the current block
    for (int k = 0; k < n_blocks; ++k) {
      const uint32_t s = blk_rowptr[(size_t)k * (n_rows + 1) + r];
      const uint32_t n_k = blk_rowptr[(size_t)k * (n_rows + 1) + r + 1] - s;
      const int64_t dst = blk_base[k] + s;
      for (uint32_t i = lane; i < n_k; i += 32) {
        col_out[dst + i] = col[first_in_row + i];
        val_out[dst + i] = val[first_in_row + i];
      }
      first_in_row += n_k;
    }
    (void)e;
  }
}

__device__ __forceinline__ double ld_mat_cs(const double* v, int64_t i, const double*) { return __ldcs(v + i); }
__device__ __forceinline__ c128 ld_mat_cs(const c128* v, int64_t i, const double*) {
  const double2 t = __ldcs(reinterpret_cast<const double2*>(v) + i);
  return make_c128(t.x, t.y);
}
__device__ __forceinline__ double ld_mat_cs(const uint16_t* v, int64_t i, const double* lut) { return __ldg(lut + __ldcs(v + i)); }

// ---- value dictionary ------------------------------------------------------------------------------------------------
// The entries of a (reduced) Hamiltonian take few distinct values -- a * sqrt(N_r / N_r') * chi with a handful of
// amplitudes, orbit sizes and characters -- so real-valued cached matrices store a 16-bit code per entry and a small
// table instead of 8-byte doubles: 6 instead of 12 bytes per entry to stream.  Values are keyed by bit pattern (exact);
// more than 65,535 distinct values leave the matrix uncoded.
#define ED_DICT_SLOTS (1u << 18)
#define ED_DICT_EMPTY 0x7FF8DEADBEEF0001ull
__device__ __forceinline__ uint32_t dict_hash(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
  return (uint32_t)k & (ED_DICT_SLOTS - 1u);
}
__global__ void __launch_bounds__(256) k_dict_insert(int64_t nnz, const double* __restrict__ val, unsigned long long* __restrict__ table, int* __restrict__ count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(val[i]);
    uint32_t slot = dict_hash(key);
    for (int probe = 0; probe < (int)ED_DICT_SLOTS; ++probe) {
      unsigned long long cur = table[slot];
      if (cur == key) break;
      if (cur == ED_DICT_EMPTY) {
        if (*count > 70000) return;                        // hopeless: stop filling the table
        cur = atomicCAS(table + slot, (unsigned long long)ED_DICT_EMPTY, key);
        if (cur == ED_DICT_EMPTY) { atomicAdd(count, 1); break; }
        if (cur == key) break;
      }
      slot = (slot + 1u) & (ED_DICT_SLOTS - 1u);
    }
  }
}
__global__ void __launch_bounds__(256) k_dict_encode(int64_t nnz, const double* __restrict__ val, const unsigned long long* __restrict__ table,
                                                     const uint16_t* __restrict__ slot_code, uint16_t* __restrict__ code) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(__ldcs(val + i));
    uint32_t slot = dict_hash(key);
    while (__ldg(table + slot) != key) slot = (slot + 1u) & (ED_DICT_SLOTS - 1u);
    code[i] = __ldg(slot_code + slot);
  }
}

// mode 0: y = partial (first block of mul!), 1: y += partial.  dot_partials only on the last block (y is final there).
template <typename VecT, typename ValT, int LANES>
__global__ void __launch_bounds__(256)
k_spmv_csr_blk(int64_t n_rows, int64_t row_lo, const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               const ValT* __restrict__ val, const double* __restrict__ lut, const VecT* __restrict__ x, VecT* __restrict__ y, int mode,
               double* __restrict__ dot_partials) {
  const int sub = threadIdx.x & (LANES - 1);
  const int64_t grp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
  const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / LANES;
  double dre = 0.0, dim_ = 0.0;
  // every group of a warp runs the same number of iterations (shuffles need the full warp)
  const int64_t n_iter = (n_rows + ngrp - 1) / ngrp;
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t r = grp0 + it * ngrp;
    const bool live = r < n_rows;
    uint32_t b = 0, e = 0;
    if (live) { b = __ldg(rowptr + r); e = __ldg(rowptr + r + 1); }
    VecT acc = vzero((VecT*)nullptr);
    uint32_t p = b + sub;
    // four independent gathers in flight per lane
    for (; p + 3 * LANES < e; p += 4 * LANES) {
      int32_t j[4];
      VecT xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) j[u] = __ldcs(col + p + u * LANES);
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = ldg_val(x + j[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) mat_fma(acc, ld_mat_cs(val, p + u * LANES, lut), xv[u]);
    }
    for (; p < e; p += LANES) mat_fma(acc, ld_mat_cs(val, p, lut), ldg_val(x + __ldcs(col + p)));
    if (LANES > 1) {
      if (sizeof(VecT) == 16) {
        c128* a = reinterpret_cast<c128*>(&acc);
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) { a->re += __shfl_down_sync(0xffffffffu, a->re, o, LANES); a->im += __shfl_down_sync(0xffffffffu, a->im, o, LANES); }
      } else {
        double* a = reinterpret_cast<double*>(&acc);
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) *a += __shfl_down_sync(0xffffffffu, *a, o, LANES);
      }
    }
    if (live && sub == 0) {
      if (mode == 0) {
        st_val(y + r, acc);
        if (dot_partials) dot_acc(dre, dim_, ldg_val(x + row_lo + r), acc);
      } else if (e != b || dot_partials) {
        VecT out = y[r];
        fma_acc(out, 1.0, acc);
        if (e != b) st_val(y + r, out);
        if (dot_partials) dot_acc(dre, dim_, ldg_val(x + row_lo + r), out);
      }
    }
  }
  if (dot_partials) {
    __shared__ double s_red[2][8];
    dre = warp_sum(dre);
    dim_ = warp_sum(dim_);
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = dre; s_red[1][threadIdx.x >> 5] = dim_; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, c = 0;
      for (int w = 0; w < 8; ++w) { a += s_red[0][w]; c += s_red[1][w]; }
      dot_partials[2 * blockIdx.x] = a;
      dot_partials[2 * blockIdx.x + 1] = c;
    }
  }
}

// columns per pass of a column-blocked cached matrix = the x window kept in L2.  Depends on the operator and the device
// model only, so every rank of a sharded operator gets the same windows (ctx.cu gathers x window by window).
int64_t ed_csr_window_cols(const ed_oprep* o) {
  int dev = 0, l2 = 0;
  ED_CUDA(cudaGetDevice(&dev));
  ED_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
  const int64_t elem = (o->is_complex || o->rbasis) ? 16 : 8;
  // x window of a pass: two thirds of the L2 (the matrix streams through with evict-first loads).  Measured on the 6x6
  // triangular sector (336 MB of x, 126 MB L2, ms per SpMV): unblocked 9.64; 3 / 4 / 6 blocks with 4 lanes per row
  // 8.65 / 8.37 / 8.60 -- every extra pass re-reads the row pointers and y, so the window is as large as the L2 allows
  int64_t W = std::max<int64_t>(4096, (int64_t)(0.67 * (double)l2) / elem);
  if (const char* e = getenv("EDCUDA_CSR_BLOCK_COLS")) { const long long v = atoll(e); if (v > 0) W = v; }
  int64_t B = (o->dim + W - 1) / W;
  if (B > 32) { B = 32; W = (o->dim + B - 1) / B; }
  return W;
}

// per-thread gate of the column-blocked SpMV: pass b first waits for ev[b] (x of window b has arrived)
static thread_local const cudaEvent_t* t_gate_ev = nullptr;
static thread_local int t_gate_n = 0;
void ed_csr_set_column_gate(const cudaEvent_t* ev, int n) { t_gate_ev = ev; t_gate_n = n; }

// regroup the row-major cached matrix into column blocks (see above); leaves it untouched when one block suffices
static void csr_block_columns(ed_oprep* o, CsrCache* c) {
  const int64_t n = c->row_hi - c->row_lo;
  if (n <= 0 || c->nnz <= 0) return;
  const int64_t W0 = ed_csr_window_cols(o);
  if (W0 >= o->dim) return;
  int64_t W = W0;
  const int64_t B = (o->dim + W - 1) / W;
  DevBuf<uint32_t> cnt((size_t)B * (n + 1));
  ED_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)B * (n + 1) * sizeof(uint32_t), ed_stream()));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ed_sm_count() * 16));
  ED_LAUNCH(k_blk_count, grid, 256, 0, n, c->rowptr.p, c->col.p, (int)B, W, cnt.p);
  c->blk_rowptr.alloc((size_t)B * (n + 1));
  DevBuf<unsigned char> tmp;
  std::vector<int64_t> base((size_t)B + 1, 0);
  for (int64_t k = 0; k < B; ++k) {
    size_t bytes = 0;
    const uint32_t* in = cnt.p + (size_t)k * (n + 1);
    uint32_t* out = c->blk_rowptr.p + (size_t)k * (n + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n + 1, ed_stream());
    if (tmp.n < bytes) tmp.alloc(bytes);
    cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n + 1, ed_stream());
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    uint32_t total = 0;
    ED_CUDA(cudaMemcpyAsync(&total, out + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ed_stream()));
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    base[k + 1] = base[k] + (int64_t)total;
  }
  ED_REQUIRE(base[B] == c->nnz, ED_ERR_INTERNAL, "column blocking lost entries (columns not ascending within a row?)");
  cnt.release();
  DevBuf<int64_t> d_base;
  d_base.upload(base);
  DevBuf<int32_t> col2((size_t)c->nnz);
  DevBuf<double> val2((size_t)c->nnz * (c->val_complex ? 2 : 1));
  const int grid_w = (int)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)ed_sm_count() * 16));
  if (c->val_complex)
    ED_LAUNCH(k_blk_fill<c128>, grid_w, 256, 0, n, c->rowptr.p, c->col.p, reinterpret_cast<const c128*>(c->val.p), (int)B, W,
              c->blk_rowptr.p, d_base.p, col2.p, reinterpret_cast<c128*>(val2.p));
  else
    ED_LAUNCH(k_blk_fill<double>, grid_w, 256, 0, n, c->rowptr.p, c->col.p, c->val.p, (int)B, W, c->blk_rowptr.p, d_base.p, col2.p, val2.p);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  c->col = std::move(col2);
  c->val = std::move(val2);
  c->rowptr.release();
  c->n_blocks = (int)B;
  c->block_cols = W;
  c->blk_base = base;
}

void ed_reduce_pairs(const double* partials, int n, double* out2);  // apply.cu

// replace the Float64 values of a cached matrix by 16-bit codes + dictionary when it has at most 65,535 distinct values
static void csr_code_values(CsrCache* c) {
  const int64_t nnz = c->nnz;
  DevBuf<unsigned long long> table((size_t)ED_DICT_SLOTS);
  DevBuf<int> count(1);
  std::vector<unsigned long long> h_table((size_t)ED_DICT_SLOTS, ED_DICT_EMPTY);
  table.upload(h_table);
  ED_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int), ed_stream()));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nnz + 255) / 256, (int64_t)ed_sm_count() * 16));
  ED_LAUNCH(k_dict_insert, grid, 256, 0, nnz, c->val.p, table.p, count.p);
  int n_distinct = 0;
  count.download(&n_distinct, 1);
  if (n_distinct <= 0 || n_distinct > 65535) return;
  table.download(h_table.data(), h_table.size());
  std::vector<std::pair<double, uint32_t>> vals;     // (value, slot)
  for (uint32_t s = 0; s < ED_DICT_SLOTS; ++s)
    if (h_table[s] != ED_DICT_EMPTY) {
      double v;
      memcpy(&v, &h_table[s], sizeof(double));
      vals.push_back({v, s});
    }
  if ((int)vals.size() != n_distinct) return;
  std::sort(vals.begin(), vals.end(), [](const std::pair<double, uint32_t>& a, const std::pair<double, uint32_t>& b) {
    if (a.first != b.first) return a.first < b.first;
    return a.second < b.second;
  });
  std::vector<double> lut(vals.size());
  std::vector<uint16_t> slot_code((size_t)ED_DICT_SLOTS, 0);
  for (size_t k = 0; k < vals.size(); ++k) { lut[k] = vals[k].first; slot_code[vals[k].second] = (uint16_t)k; }
  DevBuf<uint16_t> d_slot_code;
  d_slot_code.upload(slot_code);
  c->lut.upload(lut);
  c->code.alloc((size_t)nnz);
  ED_LAUNCH(k_dict_encode, grid, 256, 0, nnz, c->val.p, table.p, d_slot_code.p, c->code.p);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  c->val.release();
  c->coded = true;
  c->n_values = (int)vals.size();
}

void ed_csr_cache_build(ed_oprep* o, int side) {
  auto cache = std::make_shared<CsrCache>();
  cache->side = side;
  cache->row_lo = o->row_lo;
  cache->row_hi = o->row_hi;
  const int64_t n = o->row_hi - o->row_lo;
  ED_REQUIRE(o->dim < (1ll << 31), ED_ERR_UNSUPPORTED, "cached CSR needs dimension < 2^31");
  DevBuf<int64_t> ptr1, idx1;
  DevBuf<double> val;
  int64_t nnz = 0;
  assemble_lines(o, 0.0, side, o->row_lo, o->row_hi, ptr1, idx1, val, nnz);
  cache->nnz = nnz;
  cache->rowptr.alloc((size_t)n + 1);
  cache->col.alloc((size_t)std::max<int64_t>(nnz, 1));
  DevBuf<int> flag(1);
  ED_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), ed_stream()));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((std::max(nnz, n + 1) + 255) / 256, (int64_t)ed_sm_count() * 16));
  ED_LAUNCH(k_csr_compact, grid, 256, 0, n, nnz, ptr1.p, idx1.p, val.p, o->is_complex ? 1 : 0, cache->rowptr.p, cache->col.p, flag.p);
  int any_imag = 0;
  flag.download(&any_imag, 1);
  idx1.release();
  ptr1.release();
  if (o->is_complex && !any_imag) {
    cache->val_complex = false;
    cache->val.alloc((size_t)std::max<int64_t>(nnz, 1));
    ED_LAUNCH(k_take_real, grid, 256, 0, nnz, val.p, cache->val.p);
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
  } else {
    cache->val_complex = o->is_complex;
    cache->val = std::move(val);
  }
  if (!getenv("EDCUDA_CSR_NOBLOCK")) csr_block_columns(o, cache.get());
  if (!cache->val_complex && nnz > 0 && !getenv("EDCUDA_CSR_NOCODE")) csr_code_values(cache.get());
  o->csr[side] = cache;
}

void ed_apply_csr(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot) {
  CsrCache* c = o->csr[side].get();
  ED_REQUIRE(c && c->row_lo == o->row_lo && c->row_hi == o->row_hi, ED_ERR_ARGUMENT,
             "the cached matrix was built for a different row range; rebuild it after ed_oprep_set_rows");
  const int64_t n = c->row_hi - c->row_lo;
  if (n <= 0) {
    if (alpha_dot) ED_CUDA(cudaMemsetAsync(alpha_dot, 0, 2 * sizeof(double), ed_stream()));
    return;
  }
  ED_REQUIRE(!(c->val_complex && dtype == ED_F64), ED_ERR_ARGUMENT, "a complex operator representation needs ComplexF64 vectors");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)ed_sm_count() * 16));
  DevBuf<double>& pbuf = ed_scratch<double, 4>();
  double* partials = nullptr;
  if (alpha_dot) {
    if (pbuf.n < (size_t)2 * grid) pbuf.alloc((size_t)2 * grid);
    partials = pbuf.p;
  }
  if (c->n_blocks > 1) {
    // one pass per column block; rows are short inside a block: 8 lanes per row
    const int grid_b = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ed_sm_count() * 16));
    if (alpha_dot && pbuf.n < (size_t)2 * grid_b) { pbuf.alloc((size_t)2 * grid_b); partials = pbuf.p; }
    // lanes per row: 4 measured best (8: 9.26 ms, 4: 8.37, 2: 11.9, 1: 15.8 at 4 blocks; rows hold ~14 entries per block)
    const int lanes = getenv("EDCUDA_CSR_LANES") ? atoi(getenv("EDCUDA_CSR_LANES")) : 4;
    for (int b = 0; b < c->n_blocks; ++b) {
      if (t_gate_ev && b < t_gate_n) ED_CUDA(cudaStreamWaitEvent(ed_stream(), t_gate_ev[b], 0));
      const uint32_t* rp = c->blk_rowptr.p + (size_t)b * (n + 1);
      const int32_t* cb = c->col.p + c->blk_base[b];
      const int mode = (accumulate || b > 0) ? 1 : 0;
      double* dp = (b == c->n_blocks - 1) ? partials : nullptr;
#define ED_BLK_LAUNCH(L)                                                                                                        \
      do {                                                                                                                      \
        if (c->coded && dtype == ED_F64)                                                                                        \
          ED_LAUNCH((k_spmv_csr_blk<double, uint16_t, L>), grid_b, 256, 0, n, c->row_lo, rp, cb, c->code.p + c->blk_base[b],    \
                    c->lut.p, reinterpret_cast<const double*>(x), reinterpret_cast<double*>(out), mode, dp);                    \
        else if (c->coded)                                                                                                      \
          ED_LAUNCH((k_spmv_csr_blk<c128, uint16_t, L>), grid_b, 256, 0, n, c->row_lo, rp, cb, c->code.p + c->blk_base[b],      \
                    c->lut.p, reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), mode, dp);                        \
        else if (dtype == ED_F64)                                                                                               \
          ED_LAUNCH((k_spmv_csr_blk<double, double, L>), grid_b, 256, 0, n, c->row_lo, rp, cb, c->val.p + c->blk_base[b],       \
                    nullptr, reinterpret_cast<const double*>(x), reinterpret_cast<double*>(out), mode, dp);                     \
        else if (!c->val_complex)                                                                                               \
          ED_LAUNCH((k_spmv_csr_blk<c128, double, L>), grid_b, 256, 0, n, c->row_lo, rp, cb, c->val.p + c->blk_base[b],         \
                    nullptr, reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), mode, dp);                         \
        else                                                                                                                    \
          ED_LAUNCH((k_spmv_csr_blk<c128, c128, L>), grid_b, 256, 0, n, c->row_lo, rp, cb,                                      \
                    reinterpret_cast<const c128*>(c->val.p) + c->blk_base[b], nullptr, reinterpret_cast<const c128*>(x),        \
                    reinterpret_cast<c128*>(out), mode, dp);                                                                    \
      } while (0)
      if (lanes == 1) ED_BLK_LAUNCH(1);
      else if (lanes == 2) ED_BLK_LAUNCH(2);
      else if (lanes == 4) ED_BLK_LAUNCH(4);
      else ED_BLK_LAUNCH(8);
#undef ED_BLK_LAUNCH
    }
    if (alpha_dot) ed_reduce_pairs(partials, grid_b, alpha_dot);
    return;
  }
  if (c->coded && dtype == ED_F64)
    ED_LAUNCH((k_spmv_csr<double, uint16_t>), grid, 256, 0, n, c->row_lo, c->rowptr.p, c->col.p, c->code.p, c->lut.p,
              reinterpret_cast<const double*>(x), reinterpret_cast<double*>(out), accumulate, partials);
  else if (c->coded)
    ED_LAUNCH((k_spmv_csr<c128, uint16_t>), grid, 256, 0, n, c->row_lo, c->rowptr.p, c->col.p, c->code.p, c->lut.p,
              reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), accumulate, partials);
  else if (dtype == ED_F64)
    ED_LAUNCH((k_spmv_csr<double, double>), grid, 256, 0, n, c->row_lo, c->rowptr.p, c->col.p, c->val.p, nullptr,
              reinterpret_cast<const double*>(x), reinterpret_cast<double*>(out), accumulate, partials);
  else if (!c->val_complex)
    ED_LAUNCH((k_spmv_csr<c128, double>), grid, 256, 0, n, c->row_lo, c->rowptr.p, c->col.p, c->val.p, nullptr,
              reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), accumulate, partials);
  else
    ED_LAUNCH((k_spmv_csr<c128, c128>), grid, 256, 0, n, c->row_lo, c->rowptr.p, c->col.p, reinterpret_cast<const c128*>(c->val.p), nullptr,
              reinterpret_cast<const c128*>(x), reinterpret_cast<c128*>(out), accumulate, partials);
  if (alpha_dot) ed_reduce_pairs(partials, grid, alpha_dot);
}

extern "C" {

int ed_sparse_count(ed_oprep* oprep, double tol, int64_t* nnz_out) {
  ED_TRY
  ED_REQUIRE(oprep && nnz_out, ED_ERR_ARGUMENT, "null argument");
  if (tol < 0) tol = 1.4901161193847656e-08;  // Base.rtoldefault(Float64)
  ed_require_device();
  ed_sparse_assemble(oprep, tol);
  *nnz_out = oprep->sp_nnz;
  ED_CATCH
}

int ed_oprep_cache_matrix(ed_oprep* oprep, int32_t side, int64_t* nnz_out) {
  ED_TRY
  ED_REQUIRE(oprep, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(side == ED_SIDE_LEFT || side == ED_SIDE_RIGHT, ED_ERR_ARGUMENT, "bad side");
  ed_require_device();
  ed_csr_cache_build(oprep, side);
  if (nnz_out) *nnz_out = oprep->csr[side]->nnz;
  ED_CATCH
}

int ed_oprep_drop_cache(ed_oprep* oprep) {
  ED_TRY
  ED_REQUIRE(oprep, ED_ERR_ARGUMENT, "null argument");
  oprep->csr[0].reset();
  oprep->csr[1].reset();
  ED_CATCH
}

int ed_sparse_fetch(ed_oprep* oprep, int64_t* colptr, int64_t* rowval, void* nzval) {
  ED_TRY
  ED_REQUIRE(oprep && colptr, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(oprep->sp_nnz >= 0, ED_ERR_ARGUMENT, "ed_sparse_count has to be called first");
  const int64_t nnz = oprep->sp_nnz;
  oprep->sp_colptr.download(colptr, (size_t)oprep->dim + 1);
  if (nnz > 0) {
    ED_REQUIRE(rowval && nzval, ED_ERR_ARGUMENT, "null output");
    oprep->sp_rowval.download(rowval, (size_t)nnz);
    oprep->sp_nzval.download(reinterpret_cast<double*>(nzval), (size_t)nnz * (oprep->is_complex ? 2 : 1));
  }
  oprep->sp_colptr.release();
  oprep->sp_rowval.release();
  oprep->sp_nzval.release();
  oprep->sp_nnz = -1;
  ED_CATCH
}

}  // extern "C"
