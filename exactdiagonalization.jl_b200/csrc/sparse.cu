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

void ed_sparse_assemble(ed_oprep* o, double tol) {
  const int64_t dim = o->dim;
  const int cplx = o->is_complex ? 1 : 0;
  WalkCtx W = ed_make_walk_ctx(o, ED_SIDE_RIGHT);  // column iterator
  o->sp_colptr.alloc((size_t)dim + 1);
  // columns are processed in batches so the scratch stays bounded
  const int64_t BATCH = 1ll << 22;
  DevBuf<unsigned char> tmp;
  DevBuf<int64_t> counts, offs, kept, out_offs;
  DevBuf<int64_t> raw_row;
  DevBuf<c128> raw_val;
  std::vector<DevBuf<int64_t>> rows_parts;
  std::vector<DevBuf<double>> vals_parts;
  std::vector<int64_t> part_nnz;
  int64_t nnz_total = 0;
  for (int64_t c0 = 0; c0 < dim; c0 += BATCH) {
    const int64_t nc = std::min(BATCH, dim - c0);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nc + 127) / 128, (int64_t)ed_sm_count() * 16));
    counts.alloc((size_t)nc + 1);
    offs.alloc((size_t)nc + 1);
    kept.alloc((size_t)nc + 1);
    out_offs.alloc((size_t)nc + 1);
    ED_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)(nc + 1) * sizeof(int64_t), ed_stream()));
    ED_CUDA(cudaMemsetAsync(kept.p, 0, (size_t)(nc + 1) * sizeof(int64_t), ed_stream()));
    ED_LAUNCH(k3_count_matches, grid, 128, 0, W, c0, nc, counts.p);
    exclusive_scan(counts.p, offs.p, nc + 1, tmp);
    int64_t raw_total = 0;
    ED_CUDA(cudaMemcpyAsync(&raw_total, offs.p + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, ed_stream()));
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    raw_row.alloc((size_t)std::max<int64_t>(raw_total, 1));
    raw_val.alloc((size_t)std::max<int64_t>(raw_total, 1));
    ED_LAUNCH(k4_fill_raw, grid, 128, 0, W, c0, nc, offs.p, raw_row.p, raw_val.p);
    ED_LAUNCH(k4_merge_chop, grid, 128, 0, nc, offs.p, raw_row.p, raw_val.p, tol, cplx, kept.p);
    exclusive_scan(kept.p, out_offs.p, nc + 1, tmp);
    int64_t nnz_b = 0;
    ED_CUDA(cudaMemcpyAsync(&nnz_b, out_offs.p + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, ed_stream()));
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
    rows_parts.emplace_back((size_t)std::max<int64_t>(nnz_b, 1));
    vals_parts.emplace_back((size_t)std::max<int64_t>(nnz_b, 1) * (cplx ? 2 : 1));
    ED_LAUNCH(k4_gather, grid, 128, 0, nc, offs.p, raw_row.p, raw_val.p, out_offs.p, (int64_t)0, rows_parts.back().p,
              vals_parts.back().p, cplx);
    ED_LAUNCH(k_colptr, grid, 128, 0, c0 + nc == dim ? nc : nc - 1, out_offs.p, nnz_total, o->sp_colptr.p + c0);
    part_nnz.push_back(nnz_b);
    nnz_total += nnz_b;
    ED_CUDA(cudaStreamSynchronize(ed_stream()));
  }
  if (dim == 0) {
    int64_t one = 1;
    o->sp_colptr.upload(&one, 1);
  }
  o->sp_rowval.alloc((size_t)std::max<int64_t>(nnz_total, 1));
  o->sp_nzval.alloc((size_t)std::max<int64_t>(nnz_total, 1) * (cplx ? 2 : 1));
  int64_t at = 0;
  for (size_t p = 0; p < part_nnz.size(); ++p) {
    if (part_nnz[p]) {
      ED_CUDA(cudaMemcpyAsync(o->sp_rowval.p + at, rows_parts[p].p, (size_t)part_nnz[p] * sizeof(int64_t), cudaMemcpyDeviceToDevice, ed_stream()));
      ED_CUDA(cudaMemcpyAsync(o->sp_nzval.p + at * (cplx ? 2 : 1), vals_parts[p].p, (size_t)part_nnz[p] * (cplx ? 16 : 8), cudaMemcpyDeviceToDevice, ed_stream()));
    }
    at += part_nnz[p];
  }
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  o->sp_nnz = nnz_total;
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
