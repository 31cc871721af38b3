// SumOperator upload, OperatorRepresentation / ReducedOperatorRepresentation handles,
// apply!/mul! dispatch, row/column iterators, get_element, Matrix().
//
// Replaces (reference, /root/reference/src):
//   Operator/pure_operator.jl:25-52, Operator/sum_operator.jl:13-23        term containers
//   Representation/operator_representation.jl:13-36, 66-139                  OperatorRepresentation
//   Symmetry/reduced_operator_representation.jl:16-138                       ReducedOperatorRepresentation
//   Representation/abstract_operator_representation.jl:110-132, 260-285      mul!, Matrix, apply! dispatch
#include <algorithm>

#include "ed_walk.cuh"

static WalkCtx make_walk_ctx(ed_oprep* o, int side) {
  ed_upload_terms(o);
  WalkCtx W;
  memset(&W, 0, sizeof(W));
  o->basis->materialize();
  W.L = o->basis->desc();
  W.dim = o->dim;
  W.side = side;
  const TermsDev& T = side == ED_SIDE_LEFT ? o->terms_left : o->terms_right;
  W.n_terms = T.n_terms;
  W.mask = T.mask.p;
  W.match = T.match.p;
  W.target = T.target.p;
  W.amp = T.amp.p;
  W.amp_complex = T.is_complex ? 1 : 0;
  if (o->rbasis) {
    W.reduced = 1;
    W.words = o->rbasis->words.p;
    W.S = o->rbasis->symdesc();
    const RLookupDesc R = o->rbasis->rdesc();
    W.R = R;
  } else {
    W.words = o->basis->words.p;
  }
  return W;
}

WalkCtx ed_make_walk_ctx(ed_oprep* o, int side) { return make_walk_ctx(o, side); }

__global__ void k_line_iterator(WalkCtx W, int64_t i, int64_t cap, int64_t* __restrict__ idx_out,
                                double* __restrict__ amp_out, int64_t* __restrict__ n_out, int complex_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t n = 0;
  walk_line(W, i, [&](int64_t j, c128 a) {
    if (n < cap) {
      idx_out[n] = j >= 0 ? j + 1 : -1;
      if (complex_out) { amp_out[2 * n] = a.re; amp_out[2 * n + 1] = a.im; }
      else amp_out[n] = a.re;
    }
    ++n;
  });
  *n_out = n;
}

// Matrix(opr) (abstract_operator_representation.jl:121-132): thread per column, no chop.
__global__ void __launch_bounds__(128) k_dense(WalkCtx W, double* __restrict__ out, int complex_out) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < W.dim; j += (int64_t)gridDim.x * blockDim.x) {
    walk_line(W, j, [&](int64_t i, c128 a) {
      if (i < 0) return;
      size_t at = (size_t)j * W.dim + i;
      if (complex_out) { out[2 * at] += a.re; out[2 * at + 1] += a.im; }
      else out[at] += a.re;
    });
  }
}

// ------------------------------------------------------------------ C ABI
extern "C" {

int ed_operator_create(int64_t n_terms, const uint64_t* bitmask, const uint64_t* bitrow, const uint64_t* bitcol,
                       const double* amplitude, int32_t is_complex, ed_operator** out) {
  ED_TRY
  ED_REQUIRE(out && n_terms >= 0, ED_ERR_ARGUMENT, "bad arguments");
  ED_REQUIRE(n_terms == 0 || (bitmask && bitrow && bitcol && amplitude), ED_ERR_ARGUMENT, "null term arrays");
  ED_REQUIRE(n_terms < (1ll << 24), ED_ERR_UNSUPPORTED, "too many terms");
  std::unique_ptr<ed_operator> op(new ed_operator());
  op->n_terms = n_terms;
  op->is_complex = is_complex != 0;
  for (int64_t t = 0; t < n_terms; ++t) {
    // pure_operator.jl:32-36
    ED_REQUIRE((~bitmask[t] & bitrow[t]) == 0, ED_ERR_ARGUMENT, "every bit of bitrow not in bitmask should be set to zero");
    ED_REQUIRE((~bitmask[t] & bitcol[t]) == 0, ED_ERR_ARGUMENT, "every bit of bitcol not in bitmask should be set to zero");
  }
  op->mask.assign(bitmask, bitmask + n_terms);
  op->row.assign(bitrow, bitrow + n_terms);
  op->col.assign(bitcol, bitcol + n_terms);
  op->amp.assign(amplitude, amplitude + n_terms * (is_complex ? 2 : 1));
  *out = op.release();
  ED_CATCH
}

int ed_operator_destroy(ed_operator* op) {
  delete op;
  return ED_OK;
}

int ed_oprep_create(ed_basis* basis, const ed_operator* op, ed_oprep** out) {
  ED_TRY
  ED_REQUIRE(basis && op && out, ED_ERR_ARGUMENT, "null argument");
  std::unique_ptr<ed_oprep> o(new ed_oprep());
  o->basis = basis;
  o->op = *op;
  o->is_complex = op->is_complex;
  o->dim = basis->dim;
  o->row_lo = 0;
  o->row_hi = basis->dim;
  *out = o.release();
  ED_CATCH
}

int ed_oprep_create_reduced(ed_rbasis* rbasis, const ed_operator* op, ed_oprep** out) {
  ED_TRY
  ED_REQUIRE(rbasis && op && out, ED_ERR_ARGUMENT, "null argument");
  // The reduced matrix elements a * amp[col] / amp[row] (reduced_operator_representation.jl:57-116) are those of the
  // projected operator only if the operator commutes with every symmetry element; the reference leaves that check to the
  // caller (isinvariant, Symmetry/symmetry_apply.jl:110-135) and silently returns a wrong matrix otherwise.  Checked here,
  // on the host, before any kernel is launched (EDCUDA_SKIP_INVARIANCE_CHECK=1 restores the reference's behaviour).
  if (!getenv("EDCUDA_SKIP_INVARIANCE_CHECK")) {
    int32_t inv = 1, bad = -1;
    const int rc = ed_operator_isinvariant(&rbasis->parent->space, &rbasis->sym, op, rbasis->tol, &inv, &bad);
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
    ED_REQUIRE(inv, ED_ERR_ARGUMENT, "the operator is not invariant under symmetry element " + std::to_string(bad) +
                                         " (isinvariant, Symmetry/symmetry_apply.jl:110-135): its reduced representation is not defined");
  }
  std::unique_ptr<ed_oprep> o(new ed_oprep());
  o->basis = rbasis->parent;
  o->rbasis = rbasis;
  o->op = *op;
  o->is_complex = true;  // promote_type(ComplexF64, valtype(op)), reduced_operator_representation.jl:26
  o->dim = rbasis->dim;
  o->row_lo = 0;
  o->row_hi = rbasis->dim;
  *out = o.release();
  ED_CATCH
}

int ed_oprep_destroy(ed_oprep* oprep) {
  delete oprep;
  return ED_OK;
}

int ed_oprep_dim(const ed_oprep* oprep, int64_t* dim) {
  ED_TRY
  ED_REQUIRE(oprep && dim, ED_ERR_ARGUMENT, "null argument");
  *dim = oprep->dim;
  ED_CATCH
}

int ed_oprep_dtype(const ed_oprep* oprep, int32_t* dtype) {
  ED_TRY
  ED_REQUIRE(oprep && dtype, ED_ERR_ARGUMENT, "null argument");
  *dtype = oprep->is_complex ? ED_C128 : ED_F64;
  ED_CATCH
}

int ed_oprep_set_rows(ed_oprep* oprep, int64_t row_lo, int64_t row_hi) {
  ED_TRY
  ED_REQUIRE(oprep, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= oprep->dim, ED_ERR_BOUNDS, "row range outside 0..dim");
  oprep->row_lo = row_lo;
  oprep->row_hi = row_hi;
  ED_CATCH
}

int ed_oprep_suggest_rows(ed_oprep* oprep, int32_t dtype, int32_t world, int32_t rank, int64_t* row_lo, int64_t* row_hi) {
  ED_TRY
  ED_REQUIRE(oprep && row_lo && row_hi, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(world >= 1 && rank >= 0 && rank < world, ED_ERR_ARGUMENT, "bad rank / world");
  const int64_t dim = oprep->dim;
  if (!oprep->rbasis && oprep->kernel_choice == 0 && ed_device_count() > 0 && ed_apply_u1_supported(oprep, dtype, ED_SIDE_LEFT)) {
    ed_u1_suggest_rows(oprep, dtype, world, rank, row_lo, row_hi);
  } else {  // splitrange (src/util.jl:102-121)
    *row_lo = dim / world * rank + std::min<int64_t>(rank, dim % world);
    *row_hi = dim / world * (rank + 1) + std::min<int64_t>(rank + 1, dim % world);
  }
  ED_CATCH
}

int ed_oprep_set_kernel(ed_oprep* oprep, int32_t which) {
  ED_TRY
  ED_REQUIRE(oprep, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(which == 0 || which == 1, ED_ERR_ARGUMENT, "kernel choice must be 0 (auto) or 1 (generic)");
  oprep->kernel_choice = which;
  ED_CATCH
}

static void apply_dispatch(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot) {
  if (o->csr[side] && o->kernel_choice != 1) {   // cached matrix: bandwidth-bound SpMV
    if (o->rbasis) ED_REQUIRE(dtype == ED_C128, ED_ERR_ARGUMENT, "a reduced operator representation is ComplexF64: vectors must be ED_C128");
    ed_apply_csr(o, out, x, dtype, side, accumulate, alpha_dot);
    return;
  }
  const bool fast = !o->rbasis && o->kernel_choice == 0 && ed_apply_u1_supported(o, dtype, side);
  if (o->rbasis) {
    ED_REQUIRE(dtype == ED_C128, ED_ERR_ARGUMENT, "a reduced operator representation is ComplexF64: vectors must be ED_C128");
    ed_apply_reduced(o, out, x, side, accumulate, alpha_dot);
  } else if (fast) {
    ed_apply_u1(o, out, x, dtype, side, accumulate, alpha_dot);
  } else {
    ed_apply_generic(o, out, x, dtype, side, accumulate, alpha_dot);
  }
}

int ed_apply(ed_oprep* oprep, void* out, int64_t n_out, const void* x, int64_t n_x, int32_t dtype, int32_t side,
             int32_t accumulate) {
  ED_TRY
  ED_REQUIRE(oprep, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "dtype must be ED_F64 or ED_C128");
  ED_REQUIRE(side == ED_SIDE_LEFT || side == ED_SIDE_RIGHT, ED_ERR_ARGUMENT, "bad side");
  const int64_t n_rows = oprep->row_hi - oprep->row_lo;
  // abstract_operator_representation.jl:303-307 / :334-338
  ED_REQUIRE(n_out == n_rows, ED_ERR_DIMENSION_MISMATCH,
             "out has length " + std::to_string(n_out) + " != dimension " + std::to_string(n_rows));
  ED_REQUIRE(n_x == oprep->dim, ED_ERR_DIMENSION_MISMATCH,
             "state has length " + std::to_string(n_x) + " != dimension " + std::to_string(oprep->dim));
  ED_REQUIRE(!(oprep->is_complex && dtype == ED_F64), ED_ERR_ARGUMENT,
             "a complex operator representation needs ComplexF64 vectors");
  if (n_rows == 0) return ED_OK;
  ED_REQUIRE(out && x, ED_ERR_ARGUMENT, "null vector");
  ed_require_device();
  const size_t es = dtype == ED_C128 ? 16 : 8;
  Staged sx(x, (size_t)n_x * es, true, false);
  Staged so(out, (size_t)n_out * es, accumulate != 0, true);
  apply_dispatch(oprep, so.dev, sx.dev, dtype, side, accumulate, nullptr);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  so.finish();
  sx.finish();
  ED_CATCH
}

int ed_apply_async(ed_oprep* oprep, void* out, const void* x, int32_t dtype, int32_t side, int32_t accumulate,
                   double* alpha_dot) {
  ED_TRY
  ED_REQUIRE(oprep && out && x, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "dtype must be ED_F64 or ED_C128");
  ED_REQUIRE(!(oprep->is_complex && dtype == ED_F64), ED_ERR_ARGUMENT,
             "a complex operator representation needs ComplexF64 vectors");
  ed_require_device();
  apply_dispatch(oprep, out, x, dtype, side, accumulate, alpha_dot);
  ED_CATCH
}

int ed_oprep_row_iterator(ed_oprep* oprep, int64_t i, int32_t side, int64_t cap, int64_t* index_out,
                          double* amplitude_out, int64_t* n_out) {
  ED_TRY
  ED_REQUIRE(oprep && n_out, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(i >= 1 && i <= oprep->dim, ED_ERR_BOUNDS, "index " + std::to_string(i) + " outside 1.." + std::to_string(oprep->dim));
  ed_require_device();
  WalkCtx W = make_walk_ctx(oprep, side);
  const int cplx = oprep->is_complex ? 1 : 0;
  const int64_t c = std::max<int64_t>(cap, 1);
  DevBuf<int64_t> di((size_t)c), dn(1);
  DevBuf<double> da((size_t)c * 2);
  ED_LAUNCH(k_line_iterator, 1, 32, 0, W, i - 1, cap, di.p, da.p, dn.p, cplx);
  int64_t n = 0;
  dn.download(&n, 1);
  *n_out = n;
  int64_t ncopy = std::min(n, cap);
  if (ncopy > 0) {
    ED_REQUIRE(index_out && amplitude_out, ED_ERR_ARGUMENT, "null output");
    di.download(index_out, (size_t)ncopy);
    da.download(amplitude_out, (size_t)ncopy * (cplx ? 2 : 1));
  }
  ED_CATCH
}

int ed_oprep_get_element(ed_oprep* oprep, int64_t i, int64_t j, double* value_out) {
  ED_TRY
  ED_REQUIRE(oprep && value_out, ED_ERR_ARGUMENT, "null argument");
  const int64_t dim = oprep->dim;
  // operator_representation.jl:111-115 / reduced_operator_representation.jl:125-130
  ED_REQUIRE(i >= 1 && i <= dim && j >= 1 && j <= dim, ED_ERR_BOUNDS, "index outside 1..dim");
  ed_require_device();
  double re = 0, im = 0;
  if (!oprep->rbasis) {
    uint64_t w[2];
    ED_REQUIRE(ed_basis_download(oprep->basis, i - 1, 1, &w[0]) == ED_OK, ED_ERR_INTERNAL, ed_last_error());
    ED_REQUIRE(ed_basis_download(oprep->basis, j - 1, 1, &w[1]) == ED_OK, ED_ERR_INTERNAL, ed_last_error());
    const ed_operator& op = oprep->op;
    for (int64_t t = 0; t < op.n_terms; ++t) {  // operator_iterator.jl:71-83
      if ((w[0] & op.mask[t]) == op.row[t] && ((w[0] & ~op.mask[t]) | op.col[t]) == w[1]) {
        if (op.is_complex) { re += op.amp[2 * t]; im += op.amp[2 * t + 1]; }
        else re += op.amp[t];
      }
    }
  } else {
    // sum of the column iterator's entries that land on row i (:131-137)
    int64_t cap = oprep->op.n_terms, n = 0;
    std::vector<int64_t> idx((size_t)std::max<int64_t>(cap, 1));
    std::vector<double> amp((size_t)std::max<int64_t>(cap, 1) * 2);
    int rc = ed_oprep_row_iterator(oprep, j, ED_SIDE_RIGHT, cap, idx.data(), amp.data(), &n);
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
    for (int64_t k = 0; k < n; ++k)
      if (idx[k] == i) { re += amp[2 * k]; im += amp[2 * k + 1]; }
  }
  value_out[0] = re;
  value_out[1] = im;
  ED_CATCH
}

int ed_dense(ed_oprep* oprep, void* out) {
  ED_TRY
  ED_REQUIRE(oprep && (out || oprep->dim == 0), ED_ERR_ARGUMENT, "null argument");
  if (oprep->dim == 0) return ED_OK;
  ed_require_device();
  const int cplx = oprep->is_complex ? 1 : 0;
  const size_t bytes = (size_t)oprep->dim * oprep->dim * (cplx ? 16 : 8);
  ED_REQUIRE(oprep->dim <= 65536, ED_ERR_UNSUPPORTED, "dense matrix too large");
  WalkCtx W = make_walk_ctx(oprep, ED_SIDE_RIGHT);  // column iterator, as Matrix() does
  Staged so(out, bytes, false, true);
  ED_CUDA(cudaMemsetAsync(so.dev, 0, bytes, ed_stream()));
  int grid = (int)std::min<int64_t>((oprep->dim + 127) / 128, (int64_t)ed_sm_count() * 8);
  ED_LAUNCH(k_dense, std::max(grid, 1), 128, 0, W, reinterpret_cast<double*>(so.dev), cplx);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  so.finish();
  ED_CATCH
}

}  // extern "C"
