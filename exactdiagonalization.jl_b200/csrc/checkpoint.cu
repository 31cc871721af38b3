// Checkpoints (raw binary files) and the operator-invariance gate of reduced representations.
//
// Not in the reference (it has no persistence; SURVEY section 8(f)-4 asks for basis / Lanczos-state dumps):
//   ed_basis_save / ed_basis_load       a HilbertSpaceRepresentation: sector bases store how they were generated (the rank
//                                       tables are rebuilt in microseconds), user lists store their words
//   ed_rbasis_save / ed_rbasis_load     a ReducedHilbertSpaceRepresentation: representatives, orbit sizes, stabiliser marks
//                                       (skips the filter pass over the parent space: 1.3 s for the 9.08e9-state parent
//                                       of the 6x6 triangular sector); bound to the symmetry it was made with by a hash
//   ed_lanczos_state_*                  the K7 loop as a resumable object: run k steps, save, load, continue -- the
//                                       continued (alpha, beta) equal those of an uninterrupted run bit for bit
//   ed_operator_isinvariant             isinvariant(hs, symop, op) (Symmetry/symmetry_apply.jl:110-135) for every element
//                                       of a symmetry; ed_oprep_create_reduced calls it so that a non-invariant operator
//                                       is rejected (ArgumentError) instead of silently giving a wrong reduced matrix
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <unordered_map>

#include "ed_device.cuh"

namespace {

struct File {
  FILE* f = nullptr;
  File(const char* path, const char* mode) { f = path ? fopen(path, mode) : nullptr; }
  ~File() { if (f) fclose(f); }
  void write(const void* p, size_t n) { ED_REQUIRE(fwrite(p, 1, n, f) == n, ED_ERR_INTERNAL, "short write to checkpoint file"); }
  void read(void* p, size_t n) { ED_REQUIRE(fread(p, 1, n, f) == n, ED_ERR_ARGUMENT, "checkpoint file is truncated"); }
  template <typename T> void put(const T& v) { write(&v, sizeof(T)); }
  template <typename T> T get() { T v; read(&v, sizeof(T)); return v; }
};

uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

uint64_t space_hash(const ed_space& sp) {
  uint64_t h = fnv1a(&sp.n_sites, sizeof(int));
  h = fnv1a(&sp.n_qn, sizeof(int), h);
  h = fnv1a(sp.n_states.data(), sp.n_states.size() * sizeof(int), h);
  h = fnv1a(sp.qn.data(), sp.qn.size() * sizeof(int64_t), h);
  return h;
}

uint64_t symmetry_hash(const ed_symmetry& s, double tol) {
  uint64_t h = fnv1a(&s.n_ops, sizeof(int));
  h = fnv1a(&s.n_sites, sizeof(int), h);
  h = fnv1a(s.perm.data(), s.perm.size() * sizeof(int32_t), h);
  h = fnv1a(s.flip.data(), s.flip.size(), h);
  h = fnv1a(s.chi.data(), s.chi.size() * sizeof(double), h);
  h = fnv1a(&tol, sizeof(double), h);
  return h;
}

const uint64_t MAGIC_BASIS = 0x3153424443454445ull;    // "EDECDBS1"
const uint64_t MAGIC_RBASIS = 0x3142524443454445ull;   // "EDECDRB1"
const uint64_t MAGIC_LANCZOS = 0x315a4c4443454445ull;  // "EDECDLZ1"

// host-side symmetry_apply (symmetry_apply.jl:82-92, bitflipsymmetry.jl:23-35)
uint64_t host_sym_apply(const ed_space& sp, const ed_symmetry& s, int g, uint64_t b) {
  uint64_t out = 0;
  for (int i = 0; i < sp.n_sites; ++i) {
    const int j = s.perm[(size_t)g * s.n_sites + i];
    const uint64_t field = (b >> sp.offset[i]) & ((1ull << sp.width[i]) - 1ull);
    out |= field << sp.offset[j];
  }
  if (s.flip[g]) out = sp.fullmask() & ~out;
  return out;
}

}  // namespace

// ------------------------------------------------------------------ Lanczos state
struct ed_lanczos_state {
  ed_oprep* op = nullptr;
  int dtype = ED_F64;
  int64_t n = 0;
  int steps = 0;                       // completed steps
  DevBuf<unsigned char> u_cur, u_prev, w;
  std::vector<double> dots, norms;     // host copies of the finished steps: dots[2*j], norms[2*j] (norms has steps + 1 pairs)
};

static void lanczos_advance(ed_lanczos_state* st, int n_more) {
  const size_t es = st->dtype == ED_C128 ? 16 : 8;
  (void)es;
  DevBuf<double> dots((size_t)2 * n_more), norms((size_t)2 * (n_more + 2));
  // norms layout on device: [prev, cur, next_1, next_2, ...]
  double seed_norms[4] = {st->steps > 0 ? st->norms[2 * (st->steps - 1)] : 0.0, 0.0, st->norms[2 * st->steps], 0.0};
  ED_CUDA(cudaMemcpyAsync(norms.p, seed_norms, 4 * sizeof(double), cudaMemcpyHostToDevice, ed_stream()));
  for (int j = 0; j < n_more; ++j) {
    int rc = ed_apply_async(st->op, st->w.p, st->u_cur.p, st->dtype, ED_SIDE_LEFT, 0, dots.p + 2 * j);
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
    const bool has_prev = st->steps + j > 0;
    rc = ed_lanczos_update_async(st->u_prev.p, st->w.p, st->u_cur.p, st->n, st->dtype, dots.p + 2 * j, norms.p + 2 * (j + 1),
                                 has_prev ? norms.p + 2 * j : nullptr, norms.p + 2 * (j + 2));
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
    std::swap(st->u_cur, st->u_prev);
  }
  std::vector<double> hd((size_t)2 * n_more), hn((size_t)2 * (n_more + 2));
  dots.download(hd.data(), hd.size());
  norms.download(hn.data(), hn.size());
  st->dots.insert(st->dots.end(), hd.begin(), hd.end());
  st->norms.insert(st->norms.end(), hn.begin() + 4, hn.end());
  st->steps += n_more;
}

extern "C" {

// ---------------------------------------------------------------- isinvariant
int ed_operator_isinvariant(const ed_space* space, const ed_symmetry* sym, const ed_operator* op, double tol, int32_t* invariant,
                            int32_t* first_bad_element) {
  ED_TRY
  ED_REQUIRE(space && sym && op && invariant, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(sym->n_sites == space->n_sites, ED_ERR_ARGUMENT, "symmetry and Hilbert space have different numbers of sites");
  if (tol < 0) tol = 1.4901161193847656e-08;
  *invariant = 1;
  if (first_bad_element) *first_bad_element = -1;
  // <g b'| O |g b> == <b'| O |b> for every element g, tested on sample words b: every valid word of small spaces, random
  // words (every field a valid local state) otherwise.  The row of O at b is merged over duplicate targets first, so the
  // test compares matrix elements, not term lists (two different term lists can be the same operator).
  std::vector<uint64_t> samples;
  const int bits = space->bits;
  std::mt19937_64 rng(0x9e3779b97f4a7c15ull);
  auto random_word = [&]() {
    uint64_t b = 0;
    for (int i = 0; i < space->n_sites; ++i) b |= (uint64_t)(rng() % (uint64_t)space->n_states[i]) << space->offset[i];
    return b;
  };
  if (bits <= 10) {
    for (uint64_t b = 0; b < (1ull << bits); ++b) {
      bool ok = true;
      for (int i = 0; i < space->n_sites; ++i) ok = ok && (int)((b >> space->offset[i]) & ((1ull << space->width[i]) - 1ull)) < space->n_states[i];
      if (ok) samples.push_back(b);
    }
  } else {
    for (int k = 0; k < 48; ++k) samples.push_back(random_word());
  }
  auto row_of = [&](uint64_t b, std::unordered_map<uint64_t, std::pair<double, double>>& row) {
    row.clear();
    for (int64_t t = 0; t < op->n_terms; ++t) {
      if ((b & op->mask[t]) != op->row[t]) continue;
      auto& v = row[(b & ~op->mask[t]) | op->col[t]];
      if (op->is_complex) { v.first += op->amp[2 * t]; v.second += op->amp[2 * t + 1]; }
      else v.first += op->amp[t];
    }
  };
  std::unordered_map<uint64_t, std::pair<double, double>> row_b, row_gb;
  for (uint64_t b : samples) {
    row_of(b, row_b);
    for (int g = 1; g < sym->n_ops; ++g) {
      row_of(host_sym_apply(*space, *sym, g, b), row_gb);
      bool same = true;
      size_t matched = 0;
      for (auto& kv : row_b) {
        auto it = row_gb.find(host_sym_apply(*space, *sym, g, kv.first));
        const double re = it == row_gb.end() ? 0.0 : it->second.first, im = it == row_gb.end() ? 0.0 : it->second.second;
        if (it != row_gb.end()) ++matched;
        if (std::hypot(re - kv.second.first, im - kv.second.second) > tol) { same = false; break; }
      }
      if (same && matched != row_gb.size())      // targets of g(b) that are no image of a target of b must vanish
        for (auto& kv : row_gb) {
          if (std::hypot(kv.second.first, kv.second.second) <= tol) continue;
          bool found = false;
          for (auto& kb : row_b) if (host_sym_apply(*space, *sym, g, kb.first) == kv.first) { found = true; break; }
          if (!found) { same = false; break; }
        }
      if (!same) {
        *invariant = 0;
        if (first_bad_element) *first_bad_element = g;
        return ED_OK;
      }
    }
  }
  ED_CATCH
}

// ---------------------------------------------------------------- basis checkpoints
int ed_basis_save(ed_basis* basis, const char* path) {
  ED_TRY
  ED_REQUIRE(basis && path, ED_ERR_ARGUMENT, "null argument");
  File f(path, "wb");
  ED_REQUIRE(f.f, ED_ERR_ARGUMENT, std::string("cannot open ") + path + " for writing");
  f.put(MAGIC_BASIS);
  f.put(space_hash(basis->space));
  f.put((int64_t)basis->dim);
  f.put((int32_t)basis->kind);
  f.put((int32_t)basis->br_bits);
  f.put((int64_t)basis->gen_n_allowed);
  f.put((int64_t)basis->gen_allowed.size());
  if (!basis->gen_allowed.empty()) f.write(basis->gen_allowed.data(), basis->gen_allowed.size() * sizeof(int64_t));
  if (basis->gen_n_allowed == -2) {       // user list: the words themselves
    ed_require_device();
    std::vector<uint64_t> w((size_t)basis->dim);
    basis->materialize();
    basis->words.download(w.data(), w.size());
    if (!w.empty()) f.write(w.data(), w.size() * sizeof(uint64_t));
  }
  ED_CATCH
}

int ed_basis_load(const ed_space* space, const char* path, ed_basis** out) {
  ED_TRY
  ED_REQUIRE(space && path && out, ED_ERR_ARGUMENT, "null argument");
  File f(path, "rb");
  ED_REQUIRE(f.f, ED_ERR_ARGUMENT, std::string("cannot open ") + path);
  ED_REQUIRE(f.get<uint64_t>() == MAGIC_BASIS, ED_ERR_ARGUMENT, "not a basis checkpoint");
  ED_REQUIRE(f.get<uint64_t>() == space_hash(*space), ED_ERR_ARGUMENT, "the checkpoint was written for a different Hilbert space");
  const int64_t dim = f.get<int64_t>();
  const int32_t kind = f.get<int32_t>();
  const int32_t br_bits = f.get<int32_t>();
  const int64_t n_allowed = f.get<int64_t>();
  const int64_t n_vals = f.get<int64_t>();
  ED_REQUIRE(n_vals >= 0 && n_vals < (1ll << 30) && dim >= 0, ED_ERR_ARGUMENT, "corrupt basis checkpoint");
  std::vector<int64_t> allowed((size_t)n_vals);
  if (n_vals) f.read(allowed.data(), allowed.size() * sizeof(int64_t));
  int rc;
  if (n_allowed == -2) {
    std::vector<uint64_t> w((size_t)dim);
    if (dim) f.read(w.data(), w.size() * sizeof(uint64_t));
    rc = ed_basis_from_list(space, w.data(), dim, br_bits, out);
  } else {
    rc = ed_basis_generate(space, allowed.empty() ? nullptr : allowed.data(), n_allowed, br_bits, out);
  }
  ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
  if ((*out)->dim != dim || (*out)->kind != kind) {
    ed_basis_destroy(*out);
    *out = nullptr;
    throw EdError(ED_ERR_ARGUMENT, "the regenerated basis differs from the checkpoint (dimension or kind)");
  }
  ED_CATCH
}

int ed_rbasis_save(ed_rbasis* rbasis, const char* path) {
  ED_TRY
  ED_REQUIRE(rbasis && path, ED_ERR_ARGUMENT, "null argument");
  ed_require_device();
  File f(path, "wb");
  ED_REQUIRE(f.f, ED_ERR_ARGUMENT, std::string("cannot open ") + path + " for writing");
  f.put(MAGIC_RBASIS);
  f.put(space_hash(rbasis->parent->space));
  f.put(symmetry_hash(rbasis->sym, rbasis->tol));
  f.put((int64_t)rbasis->parent->dim);
  f.put((int64_t)rbasis->dim);
  const size_t n = (size_t)rbasis->dim;
  std::vector<uint64_t> w(n);
  std::vector<uint16_t> o(n), l(n);
  rbasis->words.download(w.data(), n);
  rbasis->orbit_size.download(o.data(), n);
  rbasis->last_stab.download(l.data(), n);
  if (n) { f.write(w.data(), n * 8); f.write(o.data(), n * 2); f.write(l.data(), n * 2); }
  const uint64_t check = fnv1a(w.data(), n * 8);
  f.put(check);
  ED_CATCH
}

int ed_rbasis_load(ed_basis* parent, const ed_symmetry* sym, double tol, const char* path, ed_rbasis** out) {
  ED_TRY
  ED_REQUIRE(parent && sym && path && out, ED_ERR_ARGUMENT, "null argument");
  if (tol < 0) tol = 1.4901161193847656e-08;
  ed_require_device();
  File f(path, "rb");
  ED_REQUIRE(f.f, ED_ERR_ARGUMENT, std::string("cannot open ") + path);
  ED_REQUIRE(f.get<uint64_t>() == MAGIC_RBASIS, ED_ERR_ARGUMENT, "not a reduced-basis checkpoint");
  ED_REQUIRE(f.get<uint64_t>() == space_hash(parent->space), ED_ERR_ARGUMENT, "the checkpoint was written for a different Hilbert space");
  ED_REQUIRE(f.get<uint64_t>() == symmetry_hash(*sym, tol), ED_ERR_ARGUMENT, "the checkpoint was written for different symmetry operations / characters / tolerance");
  ED_REQUIRE(f.get<int64_t>() == parent->dim, ED_ERR_ARGUMENT, "the checkpoint was written for a different parent basis");
  const int64_t dim = f.get<int64_t>();
  ED_REQUIRE(dim >= 0 && dim <= parent->dim, ED_ERR_ARGUMENT, "corrupt reduced-basis checkpoint");
  ED_REQUIRE(sym->is_group, ED_ERR_UNSUPPORTED, "the symmetry operations must form a group");
  const size_t n = (size_t)dim;
  std::vector<uint64_t> w(n);
  std::vector<uint16_t> o(n), l(n);
  if (n) { f.read(w.data(), n * 8); f.read(o.data(), n * 2); f.read(l.data(), n * 2); }
  ED_REQUIRE(f.get<uint64_t>() == fnv1a(w.data(), n * 8), ED_ERR_ARGUMENT, "reduced-basis checkpoint fails its checksum");
  std::unique_ptr<ed_rbasis> r(new ed_rbasis());
  r->parent = parent;
  r->sym = *sym;
  r->tol = tol;
  ed_symdev_build(parent->space, *sym, tol, &r->symdev);
  r->words.alloc(std::max<size_t>(n, 1)); r->orbit_size.alloc(std::max<size_t>(n, 1)); r->last_stab.alloc(std::max<size_t>(n, 1));
  r->words.upload(w.data(), n); r->orbit_size.upload(o.data(), n); r->last_stab.upload(l.data(), n);
  r->dim = dim;
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  ed_rbasis_finish_index(r.get());
  *out = r.release();
  ED_CATCH
}

// ---------------------------------------------------------------- resumable Lanczos
int ed_lanczos_state_create(ed_oprep* oprep, int32_t dtype, const void* v0, uint64_t seed, ed_lanczos_state** out) {
  ED_TRY
  ED_REQUIRE(oprep && out, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  ED_REQUIRE(!(oprep->is_complex && dtype == ED_F64), ED_ERR_ARGUMENT, "a complex operator representation needs ComplexF64 vectors");
  ED_REQUIRE(oprep->row_lo == 0 && oprep->row_hi == oprep->dim, ED_ERR_ARGUMENT, "ed_lanczos drives an unsharded representation; use ed_lanczos_sharded for row shards");
  ED_REQUIRE(oprep->dim >= 1, ED_ERR_ARGUMENT, "empty representation");
  ed_require_device();
  std::unique_ptr<ed_lanczos_state> st(new ed_lanczos_state());
  st->op = oprep; st->dtype = dtype; st->n = oprep->dim;
  const size_t bytes = (size_t)st->n * (dtype == ED_C128 ? 16 : 8);
  st->u_cur.alloc(bytes); st->u_prev.alloc(bytes); st->w.alloc(bytes);
  ED_CUDA(cudaMemsetAsync(st->u_prev.p, 0, bytes, ed_stream()));
  if (v0) ED_CUDA(cudaMemcpyAsync(st->u_cur.p, v0, bytes, ed_is_device_pointer(v0) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ed_stream()));
  else ED_REQUIRE(ed_vector_randn_async(st->u_cur.p, st->n, dtype, seed, 0) == ED_OK, ED_ERR_INTERNAL, ed_last_error());
  DevBuf<double> n0(2);
  ED_REQUIRE(ed_vector_norm2_async(st->u_cur.p, st->n, dtype, n0.p) == ED_OK, ED_ERR_INTERNAL, ed_last_error());
  st->norms.resize(2);
  n0.download(st->norms.data(), 2);
  *out = st.release();
  ED_CATCH
}

int ed_lanczos_state_destroy(ed_lanczos_state* st) {
  delete st;
  return ED_OK;
}

int ed_lanczos_state_step(ed_lanczos_state* st, int32_t n_steps) {
  ED_TRY
  ED_REQUIRE(st && n_steps >= 1, ED_ERR_ARGUMENT, "bad arguments");
  ed_require_device();
  lanczos_advance(st, n_steps);
  ED_CATCH
}

int ed_lanczos_state_result(ed_lanczos_state* st, int32_t capacity, double* alpha, double* beta, double* ritz, int32_t n_ritz,
                            int32_t* steps_done) {
  ED_TRY
  ED_REQUIRE(st && alpha && beta, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(capacity >= st->steps, ED_ERR_ARGUMENT, "alpha / beta capacity smaller than the number of steps taken");
  const int done = ed_lanczos_finish(st->dots.data(), st->norms.data(), st->steps, alpha, beta, ritz, n_ritz);
  if (steps_done) *steps_done = done;
  ED_CATCH
}

int ed_lanczos_state_save(ed_lanczos_state* st, const char* path) {
  ED_TRY
  ED_REQUIRE(st && path, ED_ERR_ARGUMENT, "null argument");
  ed_require_device();
  File f(path, "wb");
  ED_REQUIRE(f.f, ED_ERR_ARGUMENT, std::string("cannot open ") + path + " for writing");
  f.put(MAGIC_LANCZOS);
  f.put((int64_t)st->n);
  f.put((int32_t)st->dtype);
  f.put((int32_t)st->steps);
  f.write(st->dots.data(), st->dots.size() * sizeof(double));
  f.write(st->norms.data(), st->norms.size() * sizeof(double));
  const size_t bytes = (size_t)st->n * (st->dtype == ED_C128 ? 16 : 8);
  std::vector<unsigned char> host(bytes);
  st->u_cur.download(host.data(), bytes);
  f.write(host.data(), bytes);
  st->u_prev.download(host.data(), bytes);
  f.write(host.data(), bytes);
  ED_CATCH
}

int ed_lanczos_state_load(ed_oprep* oprep, const char* path, ed_lanczos_state** out) {
  ED_TRY
  ED_REQUIRE(oprep && path && out, ED_ERR_ARGUMENT, "null argument");
  ed_require_device();
  File f(path, "rb");
  ED_REQUIRE(f.f, ED_ERR_ARGUMENT, std::string("cannot open ") + path);
  ED_REQUIRE(f.get<uint64_t>() == MAGIC_LANCZOS, ED_ERR_ARGUMENT, "not a Lanczos checkpoint");
  std::unique_ptr<ed_lanczos_state> st(new ed_lanczos_state());
  st->op = oprep;
  st->n = f.get<int64_t>();
  st->dtype = f.get<int32_t>();
  st->steps = f.get<int32_t>();
  ED_REQUIRE(st->n == oprep->dim, ED_ERR_DIMENSION_MISMATCH, "the Lanczos checkpoint has a different dimension than the representation");
  ED_REQUIRE((st->dtype == ED_F64 || st->dtype == ED_C128) && st->steps >= 0 && st->steps < (1 << 24), ED_ERR_ARGUMENT, "corrupt Lanczos checkpoint");
  st->dots.resize((size_t)2 * st->steps);
  st->norms.resize((size_t)2 * (st->steps + 1));
  f.read(st->dots.data(), st->dots.size() * sizeof(double));
  f.read(st->norms.data(), st->norms.size() * sizeof(double));
  const size_t bytes = (size_t)st->n * (st->dtype == ED_C128 ? 16 : 8);
  std::vector<unsigned char> host(bytes);
  st->u_cur.alloc(bytes); st->u_prev.alloc(bytes); st->w.alloc(bytes);
  f.read(host.data(), bytes);
  st->u_cur.upload(host.data(), bytes);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  f.read(host.data(), bytes);
  st->u_prev.upload(host.data(), bytes);
  ED_CUDA(cudaStreamSynchronize(ed_stream()));
  *out = st.release();
  ED_CATCH
}

}  // extern "C"
