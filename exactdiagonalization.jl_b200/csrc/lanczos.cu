// K7: device-resident three-term Lanczos around the matrix-free apply.
//
// Not in the reference: it hands `mul!` to Arpack (docs/src/examples/spinhalf.md:26).  The engine
// keeps the Krylov vectors UNNORMALISED on device (u_j, n_j = |u_j|, v_j = u_j / n_j) so that no pass
// is spent on scaling; all scalars stay in device memory and the loop never synchronises with the host:
//     w        = H u_j                         (K2/K6, epilogue gives d_j = <u_j, w>)
//     alpha_j  = d_j / n_j^2
//     u_{j+1}  = (w - alpha_j u_j) / n_j - (n_j / n_{j-1}) u_{j-1}     (written over u_{j-1}; |u_{j+1}|^2 fused)
//     beta_j   = n_{j+1}
// In the sharded multi-GPU loop (host layer) d_j and |u_{j+1}|^2 are all-reduced over NCCL between the
// two kernels; the kernels only ever see device pointers to those scalars.
#include <algorithm>
#include <cmath>

#include "ed_device.cuh"

void ed_reduce_pairs(const double* partials, int n, double* out2);  // apply.cu

// Lanczos coefficients from the device-resident scalars (see the header comment)
__device__ __forceinline__ void k7_coefs(const double* dot, const double* norm2_cur, const double* norm2_prev,
                                         double& c1, double& c2, double& c3) {
  const double n2c = *norm2_cur;
  const double nc = sqrt(n2c);
  const double alpha = dot[0] / n2c;
  c1 = 1.0 / nc;
  c2 = alpha / nc;
  c3 = 0.0;
  if (norm2_prev) {
    const double n2p = *norm2_prev;
    c3 = n2p > 0.0 ? nc / sqrt(n2p) : 0.0;
  }
}

__device__ __forceinline__ void k7_block_reduce(double acc, double* __restrict__ partials) {
  __shared__ double s_red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) a += s_red[k];
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = 0.0;
  }
}

// u_{j+1} = c1 w - c2 u_j - c3 u_{j-1} over a flat array of doubles (a ComplexF64 vector is 2n doubles: the update is
// component-wise and |r|^2 = re^2 + im^2), 32 B/row of traffic.  Pure streaming: every thread keeps U 16-byte loads per
// array in flight (Little's law on HBM3e needs > 40 KB in flight per SM); w is dead afterwards -> evict-first loads.
template <int U>
__global__ void __launch_bounds__(256)
k7_lanczos_update_vec(double2* __restrict__ u_prev_inout, const double2* __restrict__ w, const double2* __restrict__ u_cur,
                      int64_t n2, double* __restrict__ tail_prev, const double* __restrict__ tail_w, const double* __restrict__ tail_cur,
                      const double* __restrict__ dot, const double* __restrict__ norm2_cur,
                      const double* __restrict__ norm2_prev, double* __restrict__ partials) {
  double c1, c2, c3;
  k7_coefs(dot, norm2_cur, norm2_prev, c1, c2, c3);
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n2; i += U * stride) {
    double2 wv[U], uc[U], up[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { wv[k] = __ldcs(w + i + k * stride); uc[k] = __ldg(u_cur + i + k * stride); up[k] = __ldcs(u_prev_inout + i + k * stride); }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      double2 r;
      r.x = c1 * wv[k].x - c2 * uc[k].x - c3 * up[k].x;
      r.y = c1 * wv[k].y - c2 * uc[k].y - c3 * up[k].y;
      u_prev_inout[i + k * stride] = r;
      acc = fma(r.x, r.x, fma(r.y, r.y, acc));
    }
  }
  for (; i < n2; i += stride) {
    const double2 wv = __ldcs(w + i), uc = __ldg(u_cur + i), up = __ldcs(u_prev_inout + i);
    double2 r;
    r.x = c1 * wv.x - c2 * uc.x - c3 * up.x;
    r.y = c1 * wv.y - c2 * uc.y - c3 * up.y;
    u_prev_inout[i] = r;
    acc = fma(r.x, r.x, fma(r.y, r.y, acc));
  }
  if (tail_prev && blockIdx.x == 0 && threadIdx.x == 0) {   // odd number of doubles
    const double r = c1 * *tail_w - c2 * *tail_cur - c3 * *tail_prev;
    *tail_prev = r;
    acc = fma(r, r, acc);
  }
  k7_block_reduce(acc, partials);
}

// scalar fallback for vectors that are not 16-byte aligned (views into larger buffers)
__global__ void __launch_bounds__(256)
k7_lanczos_update_scalar(double* __restrict__ u_prev_inout, const double* __restrict__ w, const double* __restrict__ u_cur, int64_t n,
                         const double* __restrict__ dot, const double* __restrict__ norm2_cur,
                         const double* __restrict__ norm2_prev, double* __restrict__ partials) {
  double c1, c2, c3;
  k7_coefs(dot, norm2_cur, norm2_prev, c1, c2, c3);
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double r = c1 * w[i] - c2 * u_cur[i] - c3 * u_prev_inout[i];
    u_prev_inout[i] = r;
    acc = fma(r, r, acc);
  }
  k7_block_reduce(acc, partials);
}

template <typename VecT>
__global__ void __launch_bounds__(256) k_norm2(const VecT* __restrict__ v, int64_t n, double* __restrict__ partials) {
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    c128 a = to_c128(v[i]);
    acc += a.re * a.re + a.im * a.im;
  }
  __shared__ double s_red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) a += s_red[k];
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = 0.0;
  }
}

// Philox4x32-10 counter-based generator keyed by (seed, global index): shard-count independent.
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

template <typename VecT>
__global__ void __launch_bounds__(256) k_randn(VecT* __restrict__ v, int64_t n, uint64_t seed, int64_t offset) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t g = (uint64_t)(offset + i);
    uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), 0u, 0u};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // two Box-Muller pairs from 4x32 bits -> 53-bit uniforms in (0,1]
    const double u1 = ((double)(((uint64_t)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0) * 0.99999999999999989;
    const double u2 = (double)(((uint64_t)c[2] << 21) ^ (c[3] >> 11)) * (1.0 / 9007199254740992.0);
    const double rr = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    if (sizeof(VecT) == 16) st_val(reinterpret_cast<c128*>(v) + i, make_c128(rr * cs, rr * sn));
    else reinterpret_cast<double*>(v)[i] = rr * cs;
  }
}

static int grid_rows(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ed_sm_count() * 8));
}

static DevBuf<double>& partial_scratch(int n_blocks) {
  DevBuf<double>& buf = ed_scratch<double, 2>();
  if (buf.n < (size_t)2 * n_blocks) buf.alloc((size_t)2 * n_blocks);
  return buf;
}

static void lanczos_update(void* u_prev_inout, const void* w, const void* u_cur, int64_t n, int dtype, const double* dot,
                           const double* norm2_cur, const double* norm2_prev, double* norm2_out) {
  const int64_t nd = dtype == ED_C128 ? 2 * n : n;   // doubles
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nd / 2 + 256 * 4 - 1) / (256 * 4), (int64_t)ed_sm_count() * 8));
  double* partials = partial_scratch(grid).p;
  const bool aligned = (((uintptr_t)u_prev_inout | (uintptr_t)w | (uintptr_t)u_cur) & 15u) == 0;
  if (aligned) {
    double* up = reinterpret_cast<double*>(u_prev_inout);
    const double* wp = reinterpret_cast<const double*>(w);
    const double* uc = reinterpret_cast<const double*>(u_cur);
    const bool odd = (nd & 1) != 0;
    ED_LAUNCH(k7_lanczos_update_vec<4>, grid, 256, 0, reinterpret_cast<double2*>(up), reinterpret_cast<const double2*>(wp),
              reinterpret_cast<const double2*>(uc), nd / 2, odd ? up + nd - 1 : nullptr, odd ? wp + nd - 1 : nullptr,
              odd ? uc + nd - 1 : nullptr, dot, norm2_cur, norm2_prev, partials);
  } else {
    ED_LAUNCH(k7_lanczos_update_scalar, grid, 256, 0, reinterpret_cast<double*>(u_prev_inout), reinterpret_cast<const double*>(w),
              reinterpret_cast<const double*>(u_cur), nd, dot, norm2_cur, norm2_prev, partials);
  }
  ed_reduce_pairs(partials, grid, norm2_out);
}

static void vector_norm2(const void* v, int64_t n, int dtype, double* norm2_out) {
  const int grid = grid_rows(n);
  double* partials = partial_scratch(grid).p;
  if (dtype == ED_C128) ED_LAUNCH(k_norm2<c128>, grid, 256, 0, reinterpret_cast<const c128*>(v), n, partials);
  else ED_LAUNCH(k_norm2<double>, grid, 256, 0, reinterpret_cast<const double*>(v), n, partials);
  ed_reduce_pairs(partials, grid, norm2_out);
}

// eigenvalues of a symmetric tridiagonal matrix by the implicit QL algorithm (no vectors)
static void tridiag_eig(std::vector<double> d, std::vector<double> e, std::vector<double>& out) {
  const int n = (int)d.size();
  e.resize(n, 0.0);
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        ED_REQUIRE(iter++ < 200, ED_ERR_INTERNAL, "tridiagonal QL did not converge");
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i], b = c * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0.0;
      }
    } while (m != l);
  }
  std::sort(d.begin(), d.end());
  out = d;
}

// (alpha, beta, Ritz values) from the device-resident scalars d_j = <u_j, H u_j> and |u_j|^2 (interleaved pairs as the
// kernels write them).  Stops at an invariant subspace: beta_j below roundoff of the tridiagonal's scale -- a running
// estimate of ||T|| = max(|alpha| + beta) times 1e-10 -- means the next Krylov vector is noise divided by its own norm,
// and every later Ritz value would be a ghost.  Returns the number of valid (alpha, beta) pairs.
int ed_lanczos_finish(const double* hd, const double* hn, int n_steps, double* alpha, double* beta, double* ritz, int n_ritz) {
  int done = 0;
  double t_norm = 0.0;
  for (int j = 0; j < n_steps; ++j) {
    if (!(hn[2 * j] > 0.0) || !std::isfinite(hn[2 * j])) break;
    alpha[j] = hd[2 * j] / hn[2 * j];
    beta[j] = std::sqrt(hn[2 * (j + 1)]);
    if (!std::isfinite(alpha[j]) || !std::isfinite(beta[j])) break;
    ++done;
    t_norm = std::max(t_norm, std::fabs(alpha[j]) + beta[j]);
    if (!(beta[j] > 1e-10 * t_norm + 1e-300)) break;
  }
  if (ritz && n_ritz > 0) {
    for (int i = 0; i < n_ritz; ++i) ritz[i] = NAN;
    if (done > 0) {
      std::vector<double> d(alpha, alpha + done), e(done, 0.0), out;
      for (int i = 0; i + 1 < done; ++i) e[i] = beta[i];
      tridiag_eig(d, e, out);
      for (int i = 0; i < n_ritz && i < done; ++i) ritz[i] = out[i];
    }
  }
  return done;
}

__global__ void __launch_bounds__(256) k_scale(double* __restrict__ v, int64_t n, double a) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] *= a;
}

extern "C" {

int ed_vector_scale_async(void* v, int64_t n, int32_t dtype, double a) {
  ED_TRY
  ED_REQUIRE(v || n == 0, ED_ERR_ARGUMENT, "null vector");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  if (n == 0) return ED_OK;
  ed_require_device();
  const int64_t nd = dtype == ED_C128 ? 2 * n : n;
  ED_LAUNCH(k_scale, grid_rows(nd), 256, 0, reinterpret_cast<double*>(v), nd, a);
  ED_CATCH
}

int ed_tridiag_eigvals(const double* alpha, const double* beta, int32_t k, double* eig_out) {
  ED_TRY
  ED_REQUIRE(k >= 0 && (k == 0 || (alpha && eig_out)), ED_ERR_ARGUMENT, "bad arguments");
  if (k == 0) return ED_OK;
  std::vector<double> d(alpha, alpha + k), e(k, 0.0), out;
  for (int i = 0; i + 1 < k; ++i) e[i] = beta[i];
  tridiag_eig(d, e, out);
  for (int i = 0; i < k; ++i) eig_out[i] = out[i];
  ED_CATCH
}

int ed_vector_randn_async(void* v, int64_t n, int32_t dtype, uint64_t seed, int64_t global_row_offset) {
  ED_TRY
  ED_REQUIRE(v || n == 0, ED_ERR_ARGUMENT, "null vector");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  if (n == 0) return ED_OK;
  ed_require_device();
  if (dtype == ED_C128) ED_LAUNCH(k_randn<c128>, grid_rows(n), 256, 0, reinterpret_cast<c128*>(v), n, seed, global_row_offset);
  else ED_LAUNCH(k_randn<double>, grid_rows(n), 256, 0, reinterpret_cast<double*>(v), n, seed, global_row_offset);
  ED_CATCH
}

int ed_vector_norm2_async(const void* v, int64_t n, int32_t dtype, double* norm2_out) {
  ED_TRY
  ED_REQUIRE(norm2_out && (v || n == 0), ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  ed_require_device();
  if (n == 0) { ED_CUDA(cudaMemsetAsync(norm2_out, 0, 2 * sizeof(double), ed_stream())); return ED_OK; }
  vector_norm2(v, n, dtype, norm2_out);
  ED_CATCH
}

int ed_lanczos_update_async(void* u_prev_inout, const void* w, const void* u_cur, int64_t n, int32_t dtype,
                            const double* dot, const double* norm2_cur, const double* norm2_prev, double* norm2_out) {
  ED_TRY
  ED_REQUIRE(norm2_out && dot && norm2_cur, ED_ERR_ARGUMENT, "null scalar pointer");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  ed_require_device();
  if (n == 0) { ED_CUDA(cudaMemsetAsync(norm2_out, 0, 2 * sizeof(double), ed_stream())); return ED_OK; }
  ED_REQUIRE(u_prev_inout && w && u_cur, ED_ERR_ARGUMENT, "null vector");
  lanczos_update(u_prev_inout, w, u_cur, n, dtype, dot, norm2_cur, norm2_prev, norm2_out);
  ED_CATCH
}

int ed_lanczos(ed_oprep* oprep, int32_t n_steps, const void* v0, int32_t dtype, uint64_t seed, double* alpha,
               double* beta, double* ritz, int32_t n_ritz, int32_t* steps_done) {
  ED_TRY
  ED_REQUIRE(oprep && alpha && beta, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(n_steps >= 1, ED_ERR_ARGUMENT, "n_steps must be positive");
  // the resumable state object (checkpoint.cu) run in one go
  ed_lanczos_state* st = nullptr;
  int rc = ed_lanczos_state_create(oprep, dtype, v0, seed, &st);
  ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
  struct Guard { ed_lanczos_state* s; ~Guard() { ed_lanczos_state_destroy(s); } } guard{st};
  rc = ed_lanczos_state_step(st, n_steps);
  ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
  rc = ed_lanczos_state_result(st, n_steps, alpha, beta, ritz, n_ritz, steps_done);
  ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
  ED_CATCH
}

}  // extern "C"
