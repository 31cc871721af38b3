// Multi-GPU context behind the C ABI: row-sharded matvec and Lanczos over the GPUs of one node.
//
// The reference picks its parallelism INSIDE apply! (Threads.@threads over rows,
// Representation/abstract_operator_representation.jl:260-267, 358-378; splitrange, util.jl:102-121); a caller never
// sees it.  Here the rows are split over GPUs the same way -- inside the library:
//   ed_ctx      a communicator of `world` ranks, one rank per GPU.  Either ONE process drives all GPUs
//               (ed_ctx_create: ncclCommInitAll, peer access between the devices), or there is one process per GPU
//               (ed_ctx_create_rank: ncclCommInitRank from a unique id the launcher broadcast; peer buffers through CUDA IPC).
//   ed_sharded  an operator representation whose rows are distributed over the ranks.
//                 * tiled U(1) kernel (apply_u1.cu): every rank owns whole tiles (ed_u1_shard_layout chooses the
//                   partition with the smallest halo).  Per matvec every rank PACKS the tiles its peers read into one
//                   send buffer (k_pack), the peers' copy engines PULL their pieces over NVLink into a compact halo
//                   buffer -- one contiguous copy per peer and launch chunk -- and the kernel chunks are launched in
//                   order, each waiting only for its own pieces: the transfer of chunk c+1 overlaps the kernel of
//                   chunk c.  No collective moves vector data; one tiny all-reduce per matvec is the fence;
//                 * everything else (generic kernel, reduced representations, cached CSR): contiguous row ranges and an
//                   NCCL all-gather of x per matvec (what north_star prescribes).
//   ed_dvec     a distributed vector: each rank holds its rows.
//   ed_lanczos_sharded  the K7 loop with the two scalars per step all-reduced over NCCL.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch already loaded, or the system one), so the library
// itself has no link-time dependency on it; a context whose ranks all sit on ONE device of one process ("loopback", used
// by the single-GPU tests of the world > 1 logic) replaces the collectives by stream-ordered kernels and needs no NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>

#include "ed_device.cuh"

// ------------------------------------------------------------------ NCCL, bound at run time
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi ready;
  if (ready.handle) return ready;
  NcclApi api;                         // published only when every symbol resolved
  const char* names[] = {getenv("EDCUDA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    const char* why = dlerror();
    throw EdError(ED_ERR_UNSUPPORTED, std::string("cannot load NCCL (libnccl.so.2): ") + (why ? why : ""));
  }
#define ED_NCCL_SYM(name)                                                                  \
  api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.handle, "nccl" #name));        \
  ED_REQUIRE(api.name, ED_ERR_UNSUPPORTED, "libnccl lacks nccl" #name)
  ED_NCCL_SYM(GetVersion); ED_NCCL_SYM(GetUniqueId); ED_NCCL_SYM(CommInitRank); ED_NCCL_SYM(CommInitAll);
  ED_NCCL_SYM(CommDestroy); ED_NCCL_SYM(AllReduce); ED_NCCL_SYM(AllGather); ED_NCCL_SYM(Broadcast);
  ED_NCCL_SYM(GroupStart); ED_NCCL_SYM(GroupEnd); ED_NCCL_SYM(GetErrorString); ED_NCCL_SYM(Send); ED_NCCL_SYM(Recv);
#undef ED_NCCL_SYM
  ready = api;
  return ready;
}

#define ED_NCCL(call)                                                                                     \
  do {                                                                                                    \
    ncclResult_t r__ = (call);                                                                            \
    if (r__ != ncclSuccess)                                                                               \
      throw EdError(ED_ERR_CUDA, std::string("NCCL error: ") + nccl().GetErrorString(r__) + " at " +      \
                                     __FILE__ + ":" + std::to_string(__LINE__));                          \
  } while (0)
}  // namespace

// ------------------------------------------------------------------ context
#define ED_CTX_PULL_STREAMS 4
#define ED_CTX_TIMER_SLOTS 64

struct CtxRank {
  int rank = 0, device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy[ED_CTX_PULL_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t push = nullptr;                   // high priority: the push kernel must get SM slots beside the compute kernels
  ncclComm_t comm = nullptr;
  cudaEvent_t ev = nullptr;                      // scratch event (loopback joins, fences)
  cudaEvent_t timer[ED_CTX_TIMER_SLOTS] = {nullptr};
  double* scalars = nullptr;                     // 8 device doubles of scratch for the small collectives
};

struct ed_ctx {
  int world = 1;
  bool loopback = false;        // every rank on the same device of this process: collectives emulated, no NCCL
  bool multi_process = false;   // one rank in this process; peers' buffers come through CUDA IPC
  std::vector<CtxRank> local;
  int n_pull_streams = 1;       // one: the pieces arrive in the order the launch chunks need them (measured at N=2: 4.55 vs 4.91 ms with two)
};

namespace {

struct DeviceGuard {
  int saved = 0;
  DeviceGuard() { cudaGetDevice(&saved); }
  ~DeviceGuard() { cudaSetDevice(saved); }
};

// runs the calling thread's library work (ED_LAUNCH, DevBuf uploads) on a rank's device and stream
struct RankScope {
  explicit RankScope(const CtxRank& r) {
    ED_CUDA(cudaSetDevice(r.device));
    ed_push_stream(r.stream);
  }
  ~RankScope() { ed_pop_stream(); }
};

__global__ void k_loop_allreduce(double* const* bufs, int n_ranks, int count, int op) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double acc = bufs[0][i];
    for (int r = 1; r < n_ranks; ++r) acc = op == 0 ? acc + bufs[r][i] : fmax(acc, bufs[r][i]);
    for (int r = 0; r < n_ranks; ++r) bufs[r][i] = acc;
  }
}

// loopback: make every rank's stream wait for everything queued so far on all the others
void loop_join(ed_ctx* c) {
  for (auto& r : c->local) ED_CUDA(cudaEventRecord(r.ev, r.stream));
  for (auto& r : c->local)
    for (auto& o : c->local)
      if (&o != &r) ED_CUDA(cudaStreamWaitEvent(r.stream, o.ev, 0));
}

// loopback only: the ranks share one device and one host thread, hence also the library's per-thread, per-device scratch
// buffers (reduction partials, K6 staging): rank i+1's work must not start before rank i's is done
void loop_chain(ed_ctx* c, size_t i) {
  if (!c->loopback || i + 1 >= c->local.size()) return;
  ED_CUDA(cudaEventRecord(c->local[i].ev, c->local[i].stream));
  ED_CUDA(cudaStreamWaitEvent(c->local[i + 1].stream, c->local[i].ev, 0));
}

// in-place all-reduce of `count` doubles held at bufs[i] on local rank i (op 0 = sum, 1 = max), stream ordered
void ctx_allreduce(ed_ctx* c, const std::vector<double*>& bufs, int count, int op) {
  if (c->world == 1) return;
  if (c->loopback) {
    loop_join(c);
    DevBuf<double*> ptrs;
    {
      RankScope s(c->local[0]);
      static thread_local std::vector<double*> keep;
      keep = bufs;
      ptrs.upload(keep.data(), keep.size());
      ED_LAUNCH(k_loop_allreduce, 1, 64, 0, ptrs.p, (int)bufs.size(), count, op);
      ED_CUDA(cudaStreamSynchronize(c->local[0].stream));   // ptrs is freed on return (test path only)
    }
    loop_join(c);
    return;
  }
  NcclApi& N = nccl();
  DeviceGuard g;
  ED_NCCL(N.GroupStart());
  for (size_t i = 0; i < c->local.size(); ++i) {
    ED_CUDA(cudaSetDevice(c->local[i].device));
    ED_NCCL(N.AllReduce(bufs[i], bufs[i], (size_t)count, ncclDouble, op == 0 ? ncclSum : ncclMax, c->local[i].comm, c->local[i].stream));
  }
  ED_NCCL(N.GroupEnd());
}

// ragged all-gather: rank r contributes bytes[r] bytes from send[local r]; every local rank i receives them at
// recv[i] + offset[r].  NCCL: one broadcast per root inside a group (no padding to the largest shard).
void ctx_allgatherv(ed_ctx* c, const std::vector<const void*>& send, const std::vector<void*>& recv,
                    const std::vector<int64_t>& bytes, const std::vector<int64_t>& offset) {
  if (c->loopback || c->world == 1) {
    if (c->world > 1) loop_join(c);
    for (size_t i = 0; i < c->local.size(); ++i)
      for (size_t r = 0; r < c->local.size(); ++r)
        if (bytes[c->local[r].rank] > 0)
          ED_CUDA(cudaMemcpyAsync(static_cast<char*>(recv[i]) + offset[c->local[r].rank], send[r], (size_t)bytes[c->local[r].rank],
                                  cudaMemcpyDeviceToDevice, c->local[i].stream));
    if (c->world > 1) loop_join(c);
    return;
  }
  NcclApi& N = nccl();
  DeviceGuard g;
  ED_NCCL(N.GroupStart());
  for (size_t i = 0; i < c->local.size(); ++i) {
    ED_CUDA(cudaSetDevice(c->local[i].device));
    for (int root = 0; root < c->world; ++root) {
      if (bytes[root] <= 0) continue;
      const void* src = c->local[i].rank == root ? send[i] : nullptr;
      void* dst = static_cast<char*>(recv[i]) + offset[root];
      ED_NCCL(N.Broadcast(src ? src : dst, dst, (size_t)bytes[root], ncclChar, root, c->local[i].comm, c->local[i].stream));
    }
  }
  ED_NCCL(N.GroupEnd());
}

void ctx_sync(ed_ctx* c) {
  DeviceGuard g;
  for (auto& r : c->local) {
    ED_CUDA(cudaSetDevice(r.device));
    for (int s = 0; s < ED_CTX_PULL_STREAMS; ++s) ED_CUDA(cudaStreamSynchronize(r.copy[s]));
    ED_CUDA(cudaStreamSynchronize(r.push));
    ED_CUDA(cudaStreamSynchronize(r.stream));
  }
}

// stream-ordered barrier across all ranks
void ctx_fence(ed_ctx* c) {
  if (c->world == 1) return;
  std::vector<double*> bufs;
  for (auto& r : c->local) bufs.push_back(r.scalars + 7);
  ctx_allreduce(c, bufs, 1, 0);
}

void ctx_init_rank(ed_ctx* c, CtxRank& r) {
  ED_CUDA(cudaSetDevice(r.device));
  ED_CUDA(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
  for (int s = 0; s < ED_CTX_PULL_STREAMS; ++s) ED_CUDA(cudaStreamCreateWithFlags(&r.copy[s], cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    ED_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    ED_CUDA(cudaStreamCreateWithPriority(&r.push, cudaStreamNonBlocking, hi));
  }
  ED_CUDA(cudaEventCreateWithFlags(&r.ev, cudaEventDisableTiming));
  for (int t = 0; t < ED_CTX_TIMER_SLOTS; ++t) ED_CUDA(cudaEventCreate(&r.timer[t]));
  ED_CUDA(cudaMalloc(&r.scalars, 8 * sizeof(double)));
  ED_CUDA(cudaMemset(r.scalars, 0, 8 * sizeof(double)));
  (void)c;
}

}  // namespace

// ------------------------------------------------------------------ sharded representation / vectors
struct ShardRank {
  ed_oprep* op = nullptr;
  int64_t n_local = 0;
  // halo exchange (tiled U(1) kernel)
  U1ShardLayout L;
  DevBuf<uint32_t> d_tile_H;
  DevBuf<int64_t> d_dir;
  DevBuf<unsigned char> halo;
  void* send[2] = {nullptr, nullptr};            // packed copies of the tiles the peers read (double buffered)
  std::vector<const void*> peer_send[2];         // [world] the peers' send buffers as seen from this rank
  std::vector<void*> opened;                     // IPC mappings to close
  DevBuf<int64_t> d_packs;                       // [3 * n_pack_items] (src, dst, len), items of at most PACK_ITEM elements
  int n_pack_items = 0;
  // push exchange: the halo (+ one arrival counter per launch chunk behind it) is what the peers map and write
  void* halo_mem = nullptr;
  unsigned long long* counters = nullptr;        // [n_chunks] rows that have arrived for each chunk, cumulative over matvecs
  std::vector<void*> peer_halo;                  // [world]
  std::vector<unsigned long long*> peer_counters; // [world] the peers' arrival counters (behind their halos)
  DevBuf<int64_t> d_push;                        // [4 * n_push_items] (src offset, destination address, rows, counter address)
  int n_push_items = 0;
  unsigned long long seq = 0;                    // matvecs pushed so far
  DevBuf<int> wait_error;                        // set by a wait that timed out
  DevBuf<double> partials;
  std::vector<cudaEvent_t> ev_pull;      // [n_chunks * n_pull_streams]
  cudaEvent_t ev_side = nullptr;         // pack + fence done on the side stream: the copy engines may pull
  std::vector<cudaEvent_t> ev_win;       // windowed all-gather: x of column window w has arrived
  // all-gather exchange
  int64_t row_lo = 0, row_hi = 0;
  DevBuf<unsigned char> x_full;
  DevBuf<double> dot;                    // 2 doubles: this rank's <x, Hx> partial, all-reduced in place
  std::vector<int64_t> range_lo, range_hi;   // global rows of this rank, local order
};

struct ed_sharded {
  ed_ctx* ctx = nullptr;
  int dtype = ED_F64;
  bool halo = false;
  bool push = false;                         // halo exchange by owner-side pushes (arrival counters) instead of reader-side pulls
  bool ce_push = false;                      // ... pushed by the owner's copy engines from a packed send buffer instead of by SM stores
  bool nccl_p2p = false;                     // halo pieces moved by grouped ncclSend / ncclRecv (one group per launch chunk)
  bool side = false;                         // pull transport: pack + fence run on the side stream, beside the interior kernels
  int push_ctas = 64;                        // grid of the (persistent) push kernel: a few SMs' worth, the rest keep computing
  int64_t dim = 0;
  int64_t win_cols = 0;                      // all-gather exchange: x travels in windows of this many rows (0 = in one piece)
  size_t es = 8;
  std::vector<ShardRank> r;                  // per local rank
  std::vector<int64_t> rows_of_rank;         // [world]
  std::vector<int64_t> row_offset;           // all-gather exchange: first global row of every rank (+ dim)
  int parity = 0;                            // send buffer in use
};

struct ed_dvec {
  ed_sharded* sh = nullptr;
  std::vector<void*> local;                        // per local rank: cudaMalloc'ed rows of this rank
};

#define PACK_ITEM 4096

// PUSH exchange: the owner writes the tiles a peer reads straight into that peer's halo buffer over NVLink (remote
// stores are posted: no round trip per access, unlike peer loads or copy-engine reads), one CTA per item of <= PACK_ITEM
// rows, items ordered by the receiver's launch chunk.  After its stores every thread fences at system scope; then one
// thread adds the item's row count to the receiver's arrival counter of that chunk (a remote atomic).  The receiver's
// stream waits (k_wait_rows) until the counter has reached seq * (rows the chunk needs): everything this kernel and the
// kernels behind the wait need is ordered by fence + atomic on the sender and by the kernel boundary on the receiver.
template <typename VecT>
__global__ void __launch_bounds__(256) k_push(const VecT* __restrict__ x, const int64_t* __restrict__ items, int n_items) {
  // persistent: a small grid walks the item list (earliest chunk first) so that the compute kernels keep most SMs
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t* it = items + 4 * (int64_t)item;
    const VecT* src = x + it[0];
    VecT* dst = reinterpret_cast<VecT*>(it[1]);
    const int len = (int)it[2];
    int i = threadIdx.x;
    for (; i + 7 * 256 < len; i += 8 * 256) {          // eight independent loads in flight per thread
      VecT v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[i + u * 256];
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[i + u * 256] = v[u];
    }
    for (; i < len; i += 256) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd_system(reinterpret_cast<unsigned long long*>(it[3]), (unsigned long long)len);
  }
}

// copy-engine push: the arrival counter is bumped by a one-thread kernel queued behind the copy
__global__ void k_signal_rows(unsigned long long* counter, unsigned long long rows) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { __threadfence_system(); atomicAdd_system(counter, rows); }
}

// one thread polls the arrival counter of a launch chunk; gives up after ~20 s (sets *err) instead of hanging the GPU
__global__ void k_wait_rows(const unsigned long long* counter, unsigned long long target, int* err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const volatile unsigned long long* c = counter;
  unsigned long long t0 = 0, now = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*c < target) {
    __nanosleep(500);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - t0 > 20000000000ull) { *err = 1; break; }
  }
  __threadfence_system();
}

// owner-side packing: the tiles the peers read, gathered from x into the send buffer (one CTA per item of <= PACK_ITEM elements)
template <typename VecT>
__global__ void __launch_bounds__(256) k_pack(const VecT* __restrict__ x, VecT* __restrict__ send, const int64_t* __restrict__ items) {
  const int64_t src = items[3 * (int64_t)blockIdx.x], dst = items[3 * (int64_t)blockIdx.x + 1];
  const int len = (int)items[3 * (int64_t)blockIdx.x + 2];
  for (int i = threadIdx.x; i < len; i += 256) send[dst + i] = x[src + i];
}

namespace {

// gather the tiles the peers read from x into this rank's (other) send buffer
void sharded_pack(ed_sharded* S, ed_dvec* x) {
  if (!S->halo || S->ctx->world == 1) return;
  ed_ctx* c = S->ctx;
  if (S->nccl_p2p) {
    // pack on the high-priority stream, then one ncclSend/ncclRecv group per launch chunk: the sends carry the pieces the
    // peers' chunk c reads, the receives fill this rank's chunk-c part of the halo; an event per chunk releases the kernels.
    // No fence: send/recv pairs synchronise the two ranks involved, the halo is guarded by the event recorded below.
    NcclApi& N = nccl();
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = S->r[i];
      ED_CUDA(cudaSetDevice(R.device));
      ED_CUDA(cudaEventRecord(R.ev, R.stream));          // x final; this rank's kernels of the previous matvec done with the halo
      ED_CUDA(cudaStreamWaitEvent(R.push, R.ev, 0));
      if (Q.n_pack_items) {
        ed_push_stream(R.push);
        if (S->dtype == ED_F64) ED_LAUNCH(k_pack<double>, Q.n_pack_items, 256, 0, reinterpret_cast<const double*>(x->local[i]), reinterpret_cast<double*>(Q.send[0]), Q.d_packs.p);
        else ED_LAUNCH(k_pack<double2>, Q.n_pack_items, 256, 0, reinterpret_cast<const double2*>(x->local[i]), reinterpret_cast<double2*>(Q.send[0]), Q.d_packs.p);
        ed_pop_stream();
      }
    }
    const size_t per = S->es / 8;      // doubles per element
    const int n_chunks = S->r[0].L.n_chunks;
    for (int ch = 0; ch < n_chunks; ++ch) {
      ED_NCCL(N.GroupStart());
      for (size_t i = 0; i < c->local.size(); ++i) {
        CtxRank& R = c->local[i];
        ShardRank& Q = S->r[i];
        ED_CUDA(cudaSetDevice(R.device));
        for (const U1Push& p : Q.L.piece_pushes)
          if (p.chunk == ch) ED_NCCL(N.Send(static_cast<const char*>(Q.send[0]) + (size_t)p.src_off * S->es, (size_t)p.len * per, ncclDouble, p.recv, R.comm, R.push));
        for (const U1Pull& p : Q.L.pulls)
          if (p.chunk == ch) ED_NCCL(N.Recv(Q.halo.p + (size_t)p.dst_off * S->es, (size_t)p.len * per, ncclDouble, p.peer, R.comm, R.push));
      }
      ED_NCCL(N.GroupEnd());
      for (size_t i = 0; i < c->local.size(); ++i) {
        ED_CUDA(cudaSetDevice(c->local[i].device));
        ED_CUDA(cudaEventRecord(S->r[i].ev_pull[(size_t)ch * ED_CTX_PULL_STREAMS], c->local[i].push));
      }
    }
    return;
  }
  if (S->push) {
    // x of every rank is final at this point of its main stream, and every peer is done with its previous halo (a
    // collective ran after its last kernel): start filling the peers' halos on the high-priority stream
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = S->r[i];
      ++Q.seq;
      if (Q.L.pushes.empty()) continue;
      ED_CUDA(cudaSetDevice(R.device));
      ED_CUDA(cudaEventRecord(R.ev, R.stream));
      ED_CUDA(cudaStreamWaitEvent(R.push, R.ev, 0));
      ed_push_stream(R.push);
      if (!S->ce_push) {
        const int grid = std::min(Q.n_push_items, S->push_ctas);
        if (S->dtype == ED_F64) ED_LAUNCH(k_push<double>, grid, 256, 0, reinterpret_cast<const double*>(x->local[i]), Q.d_push.p, Q.n_push_items);
        else ED_LAUNCH(k_push<double2>, grid, 256, 0, reinterpret_cast<const double2*>(x->local[i]), Q.d_push.p, Q.n_push_items);
      } else {
        if (S->dtype == ED_F64) ED_LAUNCH(k_pack<double>, Q.n_pack_items, 256, 0, reinterpret_cast<const double*>(x->local[i]), reinterpret_cast<double*>(Q.send[0]), Q.d_packs.p);
        else ED_LAUNCH(k_pack<double2>, Q.n_pack_items, 256, 0, reinterpret_cast<const double2*>(x->local[i]), reinterpret_cast<double2*>(Q.send[0]), Q.d_packs.p);
        for (const U1Push& p : Q.L.piece_pushes) {
          ED_CUDA(cudaMemcpyAsync(static_cast<char*>(Q.peer_halo[p.recv]) + (size_t)p.dst_off * S->es, static_cast<const char*>(Q.send[0]) + (size_t)p.src_off * S->es,
                                  (size_t)p.len * S->es, cudaMemcpyDeviceToDevice, R.push));
          ED_LAUNCH(k_signal_rows, 1, 32, 0, Q.peer_counters[p.recv] + p.chunk, (unsigned long long)p.len);
        }
      }
      ed_pop_stream();
    }
    return;
  }
  S->parity ^= 1;
  if (S->side) {
    // pack and fence on the high-priority side stream: the main stream goes straight on to the interior tiles, which
    // need neither; the copy engines start when the side stream's fence is through (ev_side)
    NcclApi& N = nccl();
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = S->r[i];
      ED_CUDA(cudaSetDevice(R.device));
      ED_CUDA(cudaEventRecord(R.ev, R.stream));           // x final, previous kernels (halo readers) queued before this point
      ED_CUDA(cudaStreamWaitEvent(R.push, R.ev, 0));
      if (Q.n_pack_items) {
        ed_push_stream(R.push);
        if (S->dtype == ED_F64) ED_LAUNCH(k_pack<double>, Q.n_pack_items, 256, 0, reinterpret_cast<const double*>(x->local[i]), reinterpret_cast<double*>(Q.send[S->parity]), Q.d_packs.p);
        else ED_LAUNCH(k_pack<double2>, Q.n_pack_items, 256, 0, reinterpret_cast<const double2*>(x->local[i]), reinterpret_cast<double2*>(Q.send[S->parity]), Q.d_packs.p);
        ed_pop_stream();
      }
    }
    ED_NCCL(N.GroupStart());
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ED_CUDA(cudaSetDevice(R.device));
      ED_NCCL(N.AllReduce(R.scalars + 7, R.scalars + 7, 1, ncclDouble, ncclSum, R.comm, R.push));
    }
    ED_NCCL(N.GroupEnd());
    for (size_t i = 0; i < c->local.size(); ++i) {
      ED_CUDA(cudaSetDevice(c->local[i].device));
      ED_CUDA(cudaEventRecord(S->r[i].ev_side, c->local[i].push));
    }
    return;
  }
  for (size_t i = 0; i < c->local.size(); ++i) {
    ShardRank& Q = S->r[i];
    if (!Q.n_pack_items) continue;
    RankScope scope(c->local[i]);
    if (S->dtype == ED_F64) ED_LAUNCH(k_pack<double>, Q.n_pack_items, 256, 0, reinterpret_cast<const double*>(x->local[i]), reinterpret_cast<double*>(Q.send[S->parity]), Q.d_packs.p);
    else ED_LAUNCH(k_pack<double2>, Q.n_pack_items, 256, 0, reinterpret_cast<const double2*>(x->local[i]), reinterpret_cast<double2*>(Q.send[S->parity]), Q.d_packs.p);
  }
}

// y = H x.  packed = the caller already ran sharded_pack(x) and a collective on every rank's stream after it.
void sharded_apply(ed_sharded* S, ed_dvec* y, ed_dvec* x, bool packed, bool want_dot) {
  ed_ctx* c = S->ctx;
  DeviceGuard g;
  if (S->halo && S->nccl_p2p) {
    if (!packed) sharded_pack(S, x);
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = S->r[i];
      RankScope scope(R);
      for (int ch = 0; ch < Q.L.n_chunks; ++ch) {
        if (c->world > 1 && ch > 0) ED_CUDA(cudaStreamWaitEvent(R.stream, Q.ev_pull[(size_t)ch * ED_CTX_PULL_STREAMS], 0));
        U1ShardLaunch A;
        A.tile_H = Q.d_tile_H.p;
        A.first = Q.L.chunk_first[ch];
        A.count = Q.L.chunk_first[ch + 1] - Q.L.chunk_first[ch];
        A.dir = Q.d_dir.p;
        A.x_local = x->local[i];
        A.x_halo = Q.halo.p;
        A.y_local = y->local[i];
        A.stream_mode = 0;
        A.accumulate = 0;
        A.partials = want_dot ? Q.partials.p : nullptr;
        ed_apply_u1_sharded(Q.op, S->dtype, A);
      }
      if (want_dot) {
        if (Q.L.tile_H.empty()) ED_CUDA(cudaMemsetAsync(Q.dot.p, 0, 2 * sizeof(double), R.stream));
        else ed_reduce_pairs(Q.partials.p, (int)Q.L.tile_H.size(), Q.dot.p);
      }
    }
  } else if (S->halo && S->push) {
    if (!packed) {
      ctx_fence(c);          // every rank is done with its previous halo and its x is final
      sharded_pack(S, x);
    }
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = S->r[i];
      RankScope scope(R);
      for (int ch = 0; ch < Q.L.n_chunks; ++ch) {
        if (c->world > 1 && Q.L.chunk_halo_rows[ch] > 0)
          ED_LAUNCH(k_wait_rows, 1, 32, 0, Q.counters + ch, Q.seq * (unsigned long long)Q.L.chunk_halo_rows[ch], Q.wait_error.p);
        U1ShardLaunch A;
        A.tile_H = Q.d_tile_H.p;
        A.first = Q.L.chunk_first[ch];
        A.count = Q.L.chunk_first[ch + 1] - Q.L.chunk_first[ch];
        A.dir = Q.d_dir.p;
        A.x_local = x->local[i];
        A.x_halo = Q.halo_mem;
        A.y_local = y->local[i];
        A.stream_mode = 0;
        A.accumulate = 0;
        A.partials = want_dot ? Q.partials.p : nullptr;
        ed_apply_u1_sharded(Q.op, S->dtype, A);
      }
      if (want_dot) {
        if (Q.L.tile_H.empty()) ED_CUDA(cudaMemsetAsync(Q.dot.p, 0, 2 * sizeof(double), R.stream));
        else ed_reduce_pairs(Q.partials.p, (int)Q.L.tile_H.size(), Q.dot.p);
      }
      // the next push may overwrite x's successor only after this rank's push kernel has read x: order it behind
      ED_CUDA(cudaEventRecord(R.ev, R.push));
      ED_CUDA(cudaStreamWaitEvent(R.stream, R.ev, 0));
    }
  } else if (S->halo) {
    if (!packed) {
      sharded_pack(S, x);
      if (!S->side) ctx_fence(c);          // every rank's send buffer is complete, and every rank is done with the previous halo
    }
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = S->r[i];
      RankScope scope(R);
      const int np = c->n_pull_streams;
      if (S->side) {
        // the copy engines wait for the side stream (pack + fence); the main stream does not
        for (int s = 0; s < np; ++s) ED_CUDA(cudaStreamWaitEvent(R.copy[s], Q.ev_side, 0));
      } else {
        // release the copy engines at this point of the (fenced) main stream
        ED_CUDA(cudaEventRecord(R.ev, R.stream));
        for (int s = 0; s < np; ++s) ED_CUDA(cudaStreamWaitEvent(R.copy[s], R.ev, 0));
      }
      size_t ip = 0;
      int k = 0;
      for (int ch = 0; ch < Q.L.n_chunks; ++ch) {
        bool used[ED_CTX_PULL_STREAMS] = {false, false, false, false};
        for (; ip < Q.L.pulls.size() && Q.L.pulls[ip].chunk == ch; ++ip, ++k) {
          const U1Pull& p = Q.L.pulls[ip];
          const int s = k % np;
          used[s] = true;
          ED_CUDA(cudaMemcpyAsync(Q.halo.p + (size_t)p.dst_off * S->es, static_cast<const char*>(Q.peer_send[S->parity][p.peer]) + (size_t)p.src_off * S->es,
                                  (size_t)p.len * S->es, cudaMemcpyDeviceToDevice, R.copy[s]));
        }
        for (int s = 0; s < np; ++s)
          if (used[s]) {
            cudaEvent_t e = Q.ev_pull[(size_t)ch * ED_CTX_PULL_STREAMS + s];
            ED_CUDA(cudaEventRecord(e, R.copy[s]));
            ED_CUDA(cudaStreamWaitEvent(R.stream, e, 0));
          }
        U1ShardLaunch A;
        A.tile_H = Q.d_tile_H.p;
        A.first = Q.L.chunk_first[ch];
        A.count = Q.L.chunk_first[ch + 1] - Q.L.chunk_first[ch];
        A.dir = Q.d_dir.p;
        A.x_local = x->local[i];
        A.x_halo = Q.halo.p;
        A.y_local = y->local[i];
        A.stream_mode = 0;
        A.accumulate = 0;
        A.partials = want_dot ? Q.partials.p : nullptr;
        ed_apply_u1_sharded(Q.op, S->dtype, A);
      }
      if (want_dot) {
        if (Q.L.tile_H.empty()) ED_CUDA(cudaMemsetAsync(Q.dot.p, 0, 2 * sizeof(double), R.stream));
        else ed_reduce_pairs(Q.partials.p, (int)Q.L.tile_H.size(), Q.dot.p);
      }
    }
  } else {
    std::vector<const void*> send;
    std::vector<void*> recv;
    std::vector<int64_t> bytes(c->world), offs(c->world);
    for (int r = 0; r < c->world; ++r) { bytes[r] = S->rows_of_rank[r] * (int64_t)S->es; offs[r] = S->row_offset[r] * (int64_t)S->es; }
    for (size_t i = 0; i < c->local.size(); ++i) { send.push_back(x->local[i]); recv.push_back(S->r[i].x_full.p); }
    // NCCL contexts gather x window by window on the side stream (the windows of the column-blocked cached SpMV, the same
    // on every rank): pass b of a cached matrix starts when window b has arrived, the later windows travel under it.
    // Every other kernel waits for the last window.
    const int n_win = (c->world > 1 && !c->loopback && S->win_cols > 0) ? (int)((S->dim + S->win_cols - 1) / S->win_cols) : 0;
    if (n_win > 1) {
      NcclApi& N = nccl();
      for (size_t i = 0; i < c->local.size(); ++i) {
        CtxRank& R = c->local[i];
        ShardRank& Q = S->r[i];
        ED_CUDA(cudaSetDevice(R.device));
        while ((int)Q.ev_win.size() < n_win) {
          cudaEvent_t e;
          ED_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
          Q.ev_win.push_back(e);
        }
        ED_CUDA(cudaEventRecord(R.ev, R.stream));          // x final; the previous readers of x_full are queued before this
        ED_CUDA(cudaStreamWaitEvent(R.push, R.ev, 0));
      }
      for (int w = 0; w < n_win; ++w) {
        const int64_t w_lo = (int64_t)w * S->win_cols, w_hi = std::min<int64_t>(S->dim, w_lo + S->win_cols);
        ED_NCCL(N.GroupStart());
        for (size_t i = 0; i < c->local.size(); ++i) {
          CtxRank& R = c->local[i];
          ED_CUDA(cudaSetDevice(R.device));
          for (int root = 0; root < c->world; ++root) {
            const int64_t lo = std::max<int64_t>(w_lo, S->row_offset[root]);
            const int64_t hi = std::min<int64_t>(w_hi, S->row_offset[root] + S->rows_of_rank[root]);
            if (hi <= lo) continue;
            char* dst = static_cast<char*>(recv[i]) + (size_t)lo * S->es;
            const char* src = R.rank == root ? static_cast<const char*>(send[i]) + (size_t)(lo - S->row_offset[root]) * S->es : dst;
            ED_NCCL(N.Broadcast(src, dst, (size_t)(hi - lo) * S->es, ncclChar, root, R.comm, R.push));
          }
        }
        ED_NCCL(N.GroupEnd());
        for (size_t i = 0; i < c->local.size(); ++i) {
          ED_CUDA(cudaSetDevice(c->local[i].device));
          ED_CUDA(cudaEventRecord(S->r[i].ev_win[w], c->local[i].push));
        }
      }
    } else if (c->world > 1) {
      ctx_allgatherv(c, send, recv, bytes, offs);
    }
    for (size_t i = 0; i < c->local.size(); ++i) {
      RankScope scope(c->local[i]);
      ShardRank& Q = S->r[i];
      const void* xin = c->world > 1 ? (const void*)Q.x_full.p : (const void*)x->local[i];
      bool gated = false;
      if (n_win > 1) {
        const CsrCache* cc = Q.op->csr[ED_SIDE_LEFT].get();
        gated = cc && Q.op->kernel_choice != 1 && cc->n_blocks == n_win && cc->block_cols == S->win_cols;
        if (gated) ed_csr_set_column_gate(Q.ev_win.data(), n_win);
        else ED_CUDA(cudaStreamWaitEvent(c->local[i].stream, Q.ev_win[n_win - 1], 0));
      }
      if (Q.row_hi > Q.row_lo) {
        const int rc = ed_apply_async(Q.op, y->local[i], xin, S->dtype, ED_SIDE_LEFT, 0, want_dot ? Q.dot.p : nullptr);
        if (gated) {
          ed_csr_set_column_gate(nullptr, 0);
          // the later kernels of this stream must not overtake the windows either (a pass may have been skipped)
          ED_CUDA(cudaStreamWaitEvent(c->local[i].stream, Q.ev_win[n_win - 1], 0));
        }
        ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
      } else if (want_dot) {
        ED_CUDA(cudaMemsetAsync(Q.dot.p, 0, 2 * sizeof(double), c->local[i].stream));
      }
      loop_chain(c, i);
    }
  }
  if (want_dot) {
    std::vector<double*> bufs;
    for (auto& Q : S->r) bufs.push_back(Q.dot.p);
    ctx_allreduce(c, bufs, 2, 0);
  }
}

// a wait that timed out (a peer never delivered its tiles) is reported at the next host synchronisation
void check_wait_error(ed_sharded* S) {
  if (!S->halo || !S->push) return;
  for (size_t i = 0; i < S->r.size(); ++i) {
    RankScope scope(S->ctx->local[i]);
    int e = 0;
    S->r[i].wait_error.download(&e, 1);
    ED_REQUIRE(e == 0, ED_ERR_CUDA, "halo exchange timed out: a peer rank did not deliver its tiles within 20 s");
  }
}

}  // namespace

extern "C" {

int ed_ctx_unique_id(uint8_t* uid128) {
  ED_TRY
  ED_REQUIRE(uid128, ED_ERR_ARGUMENT, "null argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ED_NCCL(nccl().GetUniqueId(&id));
  memcpy(uid128, &id, 128);
  ED_CATCH
}

int ed_ctx_create(int32_t n_gpus, const int32_t* device_ids, ed_ctx** out) {
  ED_TRY
  ED_REQUIRE(out && n_gpus >= 1 && n_gpus <= 64, ED_ERR_ARGUMENT, "bad arguments");
  ed_require_device();
  DeviceGuard g;
  std::unique_ptr<ed_ctx> c(new ed_ctx());
  c->world = n_gpus;
  c->local.resize(n_gpus);
  bool same = true;
  for (int r = 0; r < n_gpus; ++r) {
    c->local[r].rank = r;
    c->local[r].device = device_ids ? device_ids[r] : r;
    same &= c->local[r].device == c->local[0].device;
  }
  c->loopback = n_gpus > 1 && same;
  ED_REQUIRE(c->loopback || n_gpus == 1 || [&] { for (int a = 0; a < n_gpus; ++a) for (int b = a + 1; b < n_gpus; ++b) if (c->local[a].device == c->local[b].device) return false; return true; }(),
             ED_ERR_ARGUMENT, "device ids must be all distinct (NCCL) or all equal (loopback test context)");
  for (auto& r : c->local) ctx_init_rank(c.get(), r);
  if (!c->loopback && n_gpus > 1) {
    for (auto& a : c->local)
      for (auto& b : c->local) {
        if (a.device == b.device) continue;
        int can = 0;
        ED_CUDA(cudaDeviceCanAccessPeer(&can, a.device, b.device));
        ED_REQUIRE(can, ED_ERR_UNSUPPORTED, "devices without peer access");
        ED_CUDA(cudaSetDevice(a.device));
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else ED_CUDA(e);
      }
    std::vector<ncclComm_t> comms(n_gpus);
    std::vector<int> devs(n_gpus);
    for (int r = 0; r < n_gpus; ++r) devs[r] = c->local[r].device;
    ED_NCCL(nccl().CommInitAll(comms.data(), n_gpus, devs.data()));
    for (int r = 0; r < n_gpus; ++r) c->local[r].comm = comms[r];
  }
  if (const char* e = getenv("EDCUDA_PULL_STREAMS")) c->n_pull_streams = std::max(1, std::min(ED_CTX_PULL_STREAMS, atoi(e)));
  *out = c.release();
  ED_CATCH
}

int ed_ctx_create_rank(int32_t world, int32_t rank, int32_t device, const uint8_t* uid128, ed_ctx** out) {
  ED_TRY
  ED_REQUIRE(out && world >= 1 && rank >= 0 && rank < world, ED_ERR_ARGUMENT, "bad arguments");
  ED_REQUIRE(world == 1 || uid128, ED_ERR_ARGUMENT, "a unique id (ed_ctx_unique_id on rank 0, broadcast by the launcher) is required");
  ed_require_device();
  std::unique_ptr<ed_ctx> c(new ed_ctx());
  c->world = world;
  c->multi_process = world > 1;
  c->local.resize(1);
  c->local[0].rank = rank;
  c->local[0].device = device;
  ctx_init_rank(c.get(), c->local[0]);
  if (world > 1) {
    ncclUniqueId id;
    memcpy(&id, uid128, 128);
    ED_NCCL(nccl().CommInitRank(&c->local[0].comm, world, id, rank));
  }
  if (const char* e = getenv("EDCUDA_PULL_STREAMS")) c->n_pull_streams = std::max(1, std::min(ED_CTX_PULL_STREAMS, atoi(e)));
  *out = c.release();
  ED_CATCH
}

int ed_ctx_destroy(ed_ctx* ctx) {
  ED_TRY
  if (!ctx) return ED_OK;
  DeviceGuard g;
  for (auto& r : ctx->local) {
    cudaSetDevice(r.device);
    cudaStreamSynchronize(r.stream);
    if (r.comm) nccl().CommDestroy(r.comm);
    for (int s = 0; s < ED_CTX_PULL_STREAMS; ++s) if (r.copy[s]) cudaStreamDestroy(r.copy[s]);
    if (r.push) cudaStreamDestroy(r.push);
    if (r.stream) cudaStreamDestroy(r.stream);
    if (r.ev) cudaEventDestroy(r.ev);
    for (int t = 0; t < ED_CTX_TIMER_SLOTS; ++t) if (r.timer[t]) cudaEventDestroy(r.timer[t]);
    if (r.scalars) cudaFree(r.scalars);
  }
  delete ctx;
  ED_CATCH
}

int ed_ctx_info(const ed_ctx* ctx, int32_t* world, int32_t* n_local, int32_t* first_rank, int32_t* nccl_version) {
  ED_TRY
  ED_REQUIRE(ctx, ED_ERR_ARGUMENT, "null argument");
  if (world) *world = ctx->world;
  if (n_local) *n_local = (int32_t)ctx->local.size();
  if (first_rank) *first_rank = ctx->local[0].rank;
  if (nccl_version) {
    *nccl_version = 0;
    if (!ctx->loopback && ctx->world > 1) { int v = 0; ED_NCCL(nccl().GetVersion(&v)); *nccl_version = v; }
  }
  ED_CATCH
}

int ed_ctx_device_stream(const ed_ctx* ctx, int32_t local_index, int32_t* device, void** cuda_stream) {
  ED_TRY
  ED_REQUIRE(ctx && local_index >= 0 && local_index < (int)ctx->local.size(), ED_ERR_ARGUMENT, "bad arguments");
  if (device) *device = ctx->local[local_index].device;
  if (cuda_stream) *cuda_stream = ctx->local[local_index].stream;
  ED_CATCH
}

int ed_ctx_sync(ed_ctx* ctx) {
  ED_TRY
  ED_REQUIRE(ctx, ED_ERR_ARGUMENT, "null argument");
  ctx_sync(ctx);
  ED_CATCH
}

int ed_ctx_barrier(ed_ctx* ctx) {
  ED_TRY
  ED_REQUIRE(ctx, ED_ERR_ARGUMENT, "null argument");
  ctx_fence(ctx);
  ctx_sync(ctx);
  ED_CATCH
}

int ed_ctx_timer_record(ed_ctx* ctx, int32_t slot) {
  ED_TRY
  ED_REQUIRE(ctx && slot >= 0 && slot < ED_CTX_TIMER_SLOTS, ED_ERR_ARGUMENT, "bad arguments");
  DeviceGuard g;
  for (auto& r : ctx->local) {
    ED_CUDA(cudaSetDevice(r.device));
    ED_CUDA(cudaEventRecord(r.timer[slot], r.stream));
  }
  ED_CATCH
}

int ed_ctx_timer_elapsed(ed_ctx* ctx, int32_t slot_a, int32_t slot_b, double* ms_max) {
  ED_TRY
  ED_REQUIRE(ctx && ms_max && slot_a >= 0 && slot_a < ED_CTX_TIMER_SLOTS && slot_b >= 0 && slot_b < ED_CTX_TIMER_SLOTS, ED_ERR_ARGUMENT, "bad arguments");
  DeviceGuard g;
  double worst = 0.0;
  for (auto& r : ctx->local) {
    ED_CUDA(cudaSetDevice(r.device));
    ED_CUDA(cudaEventSynchronize(r.timer[slot_b]));
    float ms = 0.f;
    ED_CUDA(cudaEventElapsedTime(&ms, r.timer[slot_a], r.timer[slot_b]));
    worst = std::max(worst, (double)ms);
  }
  if (ctx->multi_process) {      // max over the per-GPU processes: one more tiny all-reduce
    CtxRank& r = ctx->local[0];
    ED_CUDA(cudaSetDevice(r.device));
    ED_CUDA(cudaMemcpyAsync(r.scalars + 6, &worst, sizeof(double), cudaMemcpyHostToDevice, r.stream));
    ctx_allreduce(ctx, {r.scalars + 6}, 1, 1);
    ED_CUDA(cudaMemcpyAsync(&worst, r.scalars + 6, sizeof(double), cudaMemcpyDeviceToHost, r.stream));
    ED_CUDA(cudaStreamSynchronize(r.stream));
  }
  *ms_max = worst;
  ED_CATCH
}

int ed_ctx_allreduce_host(ed_ctx* ctx, double* values, int32_t count, int32_t op) {
  ED_TRY
  ED_REQUIRE(ctx && values && count >= 1 && count <= 4 && (op == 0 || op == 1), ED_ERR_ARGUMENT, "bad arguments (count <= 4; op 0 = sum, 1 = max)");
  if (!ctx->multi_process) return ED_OK;     // one process holds every rank: the caller's values are already global
  CtxRank& r = ctx->local[0];
  DeviceGuard g;
  ED_CUDA(cudaSetDevice(r.device));
  ED_CUDA(cudaMemcpyAsync(r.scalars, values, count * sizeof(double), cudaMemcpyHostToDevice, r.stream));
  ctx_allreduce(ctx, {r.scalars}, count, op);
  ED_CUDA(cudaMemcpyAsync(values, r.scalars, count * sizeof(double), cudaMemcpyDeviceToHost, r.stream));
  ED_CUDA(cudaStreamSynchronize(r.stream));
  ED_CATCH
}

// ---------------------------------------------------------------- sharded representation
int ed_sharded_create(ed_ctx* ctx, ed_oprep* const* opreps, int32_t dtype, int32_t exchange, int32_t n_chunks, ed_sharded** out) {
  ED_TRY
  ED_REQUIRE(ctx && opreps && out, ED_ERR_ARGUMENT, "null argument");
  ED_REQUIRE(dtype == ED_F64 || dtype == ED_C128, ED_ERR_ARGUMENT, "bad dtype");
  ED_REQUIRE(exchange >= 0 && exchange <= 5, ED_ERR_ARGUMENT, "exchange: 0 automatic, 1 NCCL all-gather, 2 halo by copy-engine pulls, 3 halo by owner SM pushes, 4 halo by owner copy-engine pushes, 5 halo by grouped ncclSend/ncclRecv");
  DeviceGuard g;
  std::unique_ptr<ed_sharded> S(new ed_sharded());
  S->ctx = ctx;
  S->dtype = dtype;
  S->es = dtype == ED_C128 ? 16 : 8;
  const int nl = (int)ctx->local.size();
  S->r.resize(nl);
  for (int i = 0; i < nl; ++i) {
    ED_REQUIRE(opreps[i], ED_ERR_ARGUMENT, "null representation");
    ED_REQUIRE(!(opreps[i]->is_complex && dtype == ED_F64), ED_ERR_ARGUMENT, "a complex operator representation needs ComplexF64 vectors");
    S->r[i].op = opreps[i];
  }
  S->dim = opreps[0]->dim;
  if (n_chunks <= 0) n_chunks = getenv("EDCUDA_SHARD_CHUNKS") ? std::max(1, atoi(getenv("EDCUDA_SHARD_CHUNKS"))) : 8;
  bool fast = true;
  for (int i = 0; i < nl; ++i) {
    ED_CUDA(cudaSetDevice(ctx->local[i].device));
    ED_REQUIRE(opreps[i]->dim == S->dim, ED_ERR_DIMENSION_MISMATCH, "the per-rank representations differ");
    fast = fast && !opreps[i]->rbasis && opreps[i]->kernel_choice == 0 && !opreps[i]->csr[0] && ed_apply_u1_supported(opreps[i], dtype, ED_SIDE_LEFT);
  }
  ED_REQUIRE(exchange < 2 || fast, ED_ERR_UNSUPPORTED, "the halo exchange is defined for the tiled U(1) kernel only");
  S->halo = fast && exchange != 1;
  // transport of the halo: 2 = reader-side copy-engine pulls, 3 = owner-side SM pushes, 4 = owner-side copy-engine pushes;
  // automatic: EDCUDA_SHARD_TRANSPORT (pull | push | cepush), default pull
  int transport = exchange >= 2 ? exchange : 2;
  if (exchange == 0)
    if (const char* e = getenv("EDCUDA_SHARD_TRANSPORT")) transport = !strcmp(e, "push") ? 3 : !strcmp(e, "cepush") ? 4 : !strcmp(e, "nccl") ? 5 : 2;
  if (transport == 5 && (ctx->loopback || ctx->world == 1)) transport = 2;      // no NCCL communicator: copy-engine pulls
  S->push = S->halo && (transport == 3 || transport == 4);
  S->ce_push = S->halo && transport == 4;
  S->nccl_p2p = S->halo && transport == 5;
  S->side = S->halo && transport == 2 && !ctx->loopback && ctx->world > 1 && !(getenv("EDCUDA_SHARD_SIDE") && atoi(getenv("EDCUDA_SHARD_SIDE")) == 0);
  if (const char* e = getenv("EDCUDA_PUSH_CTAS")) S->push_ctas = std::max(1, atoi(e));
  S->rows_of_rank.assign(ctx->world, 0);
  for (int i = 0; i < nl; ++i) {
    CtxRank& R = ctx->local[i];
    ShardRank& Q = S->r[i];
    RankScope scope(R);
    Q.dot.alloc(2);
    if (S->halo) {
      FastU1Plan* plan = ed_u1_plan(Q.op, dtype);
      const int policy = getenv("EDCUDA_SHARD_POLICY") ? atoi(getenv("EDCUDA_SHARD_POLICY")) : 0;
      ed_u1_shard_layout(plan, ctx->world, R.rank, ctx->world > 1 ? n_chunks : 1, policy, &Q.L);
      Q.n_local = Q.L.n_local;
      Q.range_lo = Q.L.range_lo; Q.range_hi = Q.L.range_hi;
      S->rows_of_rank = Q.L.rows_of_rank;
      Q.d_tile_H.upload(Q.L.tile_H);
      Q.d_dir.upload(Q.L.dir);
      Q.wait_error.alloc(1);
      ED_CUDA(cudaMemsetAsync(Q.wait_error.p, 0, sizeof(int), R.stream));
      if (S->push) {
        // halo + arrival counters in ONE allocation: one IPC mapping per peer covers both
        const size_t halo_bytes = (((size_t)std::max<int64_t>(Q.L.n_halo, 1) * S->es) + 255) & ~(size_t)255;
        ED_CUDA(cudaMalloc(&Q.halo_mem, halo_bytes + (size_t)Q.L.n_chunks * sizeof(unsigned long long)));
        Q.counters = reinterpret_cast<unsigned long long*>(static_cast<char*>(Q.halo_mem) + halo_bytes);
        ED_CUDA(cudaMemsetAsync(Q.counters, 0, (size_t)Q.L.n_chunks * sizeof(unsigned long long), R.stream));
        if (S->ce_push) ED_CUDA(cudaMalloc(&Q.send[0], (size_t)std::max<int64_t>(Q.L.n_send, 1) * S->es));
      } else {
        Q.halo.alloc((size_t)std::max<int64_t>(Q.L.n_halo, 1) * S->es);
        for (int b = 0; b < (S->nccl_p2p ? 1 : 2); ++b) ED_CUDA(cudaMalloc(&Q.send[b], (size_t)std::max<int64_t>(Q.L.n_send, 1) * S->es));
      }
      if (!S->push || S->ce_push) {
        std::vector<int64_t> items;
        for (const U1Pack& p : Q.L.packs)
          for (int64_t o = 0; o < p.len; o += PACK_ITEM) { items.push_back(p.src_off + o); items.push_back(p.dst_off + o); items.push_back(std::min<int64_t>(PACK_ITEM, p.len - o)); }
        Q.n_pack_items = (int)(items.size() / 3);
        if (items.empty()) items.assign(3, 0);
        Q.d_packs.upload(items);
      }
      Q.partials.alloc((size_t)2 * std::max<size_t>(Q.L.tile_H.size(), 1));
      Q.ev_pull.resize((size_t)Q.L.n_chunks * ED_CTX_PULL_STREAMS);
      for (auto& e : Q.ev_pull) ED_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ED_CUDA(cudaEventCreateWithFlags(&Q.ev_side, cudaEventDisableTiming));
      ED_CUDA(cudaStreamSynchronize(R.stream));
    } else {
      for (int r = 0; r < ctx->world; ++r) {
        int64_t lo = 0, hi = 0;
        const int rc = ed_oprep_suggest_rows(Q.op, dtype, ctx->world, r, &lo, &hi);
        ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
        S->rows_of_rank[r] = hi - lo;
        if (r == R.rank) { Q.row_lo = lo; Q.row_hi = hi; }
      }
      Q.n_local = Q.row_hi - Q.row_lo;
      Q.range_lo = {Q.row_lo}; Q.range_hi = {Q.row_hi};
      const int rc = ed_oprep_set_rows(Q.op, Q.row_lo, Q.row_hi);
      ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
      if (ctx->world > 1) Q.x_full.alloc((size_t)S->dim * S->es);
      // windowed gather (see sharded_apply): measured on config 4's cached SpMV, 2 / 4 / 8 ranks: 4.65 -> 4.45 ms, 2.65 -> 2.90,
      // 1.74 -> 2.11 -- with more ranks a window holds few roots' pieces and the per-group NCCL latency outweighs the overlap,
      // so it is the default for two ranks only (EDCUDA_SHARD_WINDOWS=1 / 0 forces it on / off)
      const char* we = getenv("EDCUDA_SHARD_WINDOWS");
      if (we ? atoi(we) != 0 : ctx->world == 2) S->win_cols = ed_csr_window_cols(Q.op);
    }
  }
  S->row_offset.assign(ctx->world + 1, 0);
  for (int r = 0; r < ctx->world; ++r) S->row_offset[r + 1] = S->row_offset[r] + S->rows_of_rank[r];
  ED_REQUIRE(S->row_offset[ctx->world] == S->dim, ED_ERR_INTERNAL, "the shards do not cover the basis");
  // every rank sees the buffers its peers read (pull: send buffers) or write (push: halos): directly when one process drives
  // all GPUs, through CUDA IPC with one process per GPU
  auto share = [&](const std::vector<void*>& mine /* per local rank */, std::vector<std::vector<void*>>& seen /* [local][world] */) {
    seen.assign(nl, std::vector<void*>(ctx->world, nullptr));
    if (!ctx->multi_process) {
      for (int i = 0; i < nl; ++i)
        for (int j = 0; j < nl; ++j) seen[i][ctx->local[j].rank] = mine[j];
      return;
    }
    CtxRank& R = ctx->local[0];
    ED_CUDA(cudaSetDevice(R.device));
    if (ctx->world == 1) { seen[0][0] = mine[0]; return; }
    cudaIpcMemHandle_t h_mine;
    ED_CUDA(cudaIpcGetMemHandle(&h_mine, mine[0]));
    DevBuf<unsigned char> sbuf(64), rbuf((size_t)64 * ctx->world);
    ED_CUDA(cudaMemcpyAsync(sbuf.p, &h_mine, 64, cudaMemcpyHostToDevice, R.stream));
    ED_NCCL(nccl().AllGather(sbuf.p, rbuf.p, 64, ncclChar, R.comm, R.stream));
    std::vector<unsigned char> all((size_t)64 * ctx->world);
    ED_CUDA(cudaMemcpyAsync(all.data(), rbuf.p, all.size(), cudaMemcpyDeviceToHost, R.stream));
    ED_CUDA(cudaStreamSynchronize(R.stream));
    for (int r = 0; r < ctx->world; ++r) {
      if (r == R.rank) { seen[0][r] = mine[0]; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, all.data() + (size_t)64 * r, 64);
      void* p = nullptr;
      ED_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      S->r[0].opened.push_back(p);
      seen[0][r] = p;
    }
  };
  if (S->halo && S->push) {
    std::vector<void*> mine;
    for (int i = 0; i < nl; ++i) mine.push_back(S->r[i].halo_mem);
    std::vector<std::vector<void*>> seen;
    share(mine, seen);
    for (int i = 0; i < nl; ++i) {
      ShardRank& Q = S->r[i];
      RankScope scope(ctx->local[i]);
      Q.peer_halo = seen[i];
      // every rank's counters sit behind its halo: the offset follows from that rank's halo size (known to the planner)
      std::vector<int64_t> items;
      std::vector<size_t> halo_bytes_of(ctx->world, 0);
      {
        // halo sizes of the peers: rows pushed to r by all senders = r's n_halo; recompute from the layouts
        FastU1Plan* plan = ed_u1_plan(Q.op, dtype);
        const int policy = getenv("EDCUDA_SHARD_POLICY") ? atoi(getenv("EDCUDA_SHARD_POLICY")) : 0;
        for (int r = 0; r < ctx->world; ++r) {
          U1ShardLayout Lr;
          ed_u1_shard_layout(plan, ctx->world, r, ctx->world > 1 ? n_chunks : 1, policy, &Lr);
          halo_bytes_of[r] = (((size_t)std::max<int64_t>(Lr.n_halo, 1) * S->es) + 255) & ~(size_t)255;
        }
      }
      for (const U1Push& p : Q.L.pushes)
        for (int64_t o = 0; o < p.len; o += PACK_ITEM) {
          items.push_back(p.src_off + o);
          items.push_back((int64_t)(reinterpret_cast<uintptr_t>(Q.peer_halo[p.recv]) + (size_t)(p.dst_off + o) * S->es));
          items.push_back(std::min<int64_t>(PACK_ITEM, p.len - o));
          items.push_back((int64_t)(reinterpret_cast<uintptr_t>(Q.peer_halo[p.recv]) + halo_bytes_of[p.recv] + (size_t)p.chunk * sizeof(unsigned long long)));
        }
      Q.n_push_items = (int)(items.size() / 4);
      if (items.empty()) items.assign(4, 0);
      Q.d_push.upload(items);
      Q.peer_counters.assign(ctx->world, nullptr);
      for (int r = 0; r < ctx->world; ++r)
        Q.peer_counters[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(Q.peer_halo[r]) + halo_bytes_of[r]);
      ED_CUDA(cudaStreamSynchronize(ctx->local[i].stream));
    }
    ctx_fence(ctx);        // every rank's counters are zeroed before anyone pushes
    ctx_sync(ctx);
  } else if (S->halo && !S->nccl_p2p) {
    for (int b = 0; b < 2; ++b) {
      std::vector<void*> mine;
      for (int i = 0; i < nl; ++i) mine.push_back(S->r[i].send[b]);
      std::vector<std::vector<void*>> seen;
      share(mine, seen);
      for (int i = 0; i < nl; ++i) {
        S->r[i].peer_send[b].assign(ctx->world, nullptr);
        for (int r = 0; r < ctx->world; ++r) S->r[i].peer_send[b][r] = seen[i][r];
      }
    }
  }
  *out = S.release();
  ED_CATCH
}

/* collective: every rank stops using the buffers, unmaps its peers' send buffers, and only then the owners free them
 * (an owner that frees a buffer its peers still map and later exports a new one out of the same allocation makes the
 * peers' next cudaIpcOpenMemHandle fail with "resource already mapped") */
int ed_sharded_destroy(ed_sharded* sh) {
  ED_TRY
  if (!sh) return ED_OK;
  DeviceGuard g;
  ed_ctx* c = sh->ctx;
  if (sh->halo && c->world > 1) ctx_fence(c);
  ctx_sync(c);
  for (auto& Q : sh->r)
    for (void* p : Q.opened) cudaIpcCloseMemHandle(p);
  if (sh->halo && c->world > 1) { ctx_fence(c); ctx_sync(c); }
  for (size_t i = 0; i < sh->r.size(); ++i) {
    cudaSetDevice(c->local[i].device);
    ShardRank& Q = sh->r[i];
    for (auto& e : Q.ev_pull) cudaEventDestroy(e);
    if (Q.ev_side) cudaEventDestroy(Q.ev_side);
    for (auto& e : Q.ev_win) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) if (Q.send[b]) cudaFree(Q.send[b]);
    if (Q.halo_mem) cudaFree(Q.halo_mem);
    Q.d_push.release(); Q.wait_error.release();
    Q.d_tile_H.release(); Q.d_dir.release(); Q.halo.release(); Q.partials.release(); Q.x_full.release(); Q.dot.release(); Q.d_packs.release();
  }
  delete sh;
  ED_CATCH
}

int ed_sharded_info(const ed_sharded* sh, int32_t local_index, int64_t* n_local, int64_t* n_halo, int32_t* n_ranges, int32_t* n_pulls,
                    int32_t* n_chunks, int32_t* halo_exchange) {
  ED_TRY
  ED_REQUIRE(sh && local_index >= 0 && local_index < (int)sh->r.size(), ED_ERR_ARGUMENT, "bad arguments");
  const ShardRank& Q = sh->r[local_index];
  if (n_local) *n_local = Q.n_local;
  if (n_halo) *n_halo = sh->halo ? Q.L.n_halo : (sh->ctx->world > 1 ? sh->dim - Q.n_local : 0);
  if (n_ranges) *n_ranges = (int32_t)Q.range_lo.size();
  if (n_pulls) *n_pulls = sh->halo ? (int32_t)(sh->push ? Q.L.pushes.size() : Q.L.pulls.size()) : 0;
  if (n_chunks) *n_chunks = sh->halo ? Q.L.n_chunks : 1;
  if (halo_exchange) *halo_exchange = sh->halo ? (sh->nccl_p2p ? 4 : sh->ce_push ? 3 : sh->push ? 2 : 1) : 0;
  ED_CATCH
}

int ed_sharded_ranges(const ed_sharded* sh, int32_t local_index, int64_t* row_lo, int64_t* row_hi) {
  ED_TRY
  ED_REQUIRE(sh && row_lo && row_hi && local_index >= 0 && local_index < (int)sh->r.size(), ED_ERR_ARGUMENT, "bad arguments");
  const ShardRank& Q = sh->r[local_index];
  for (size_t k = 0; k < Q.range_lo.size(); ++k) { row_lo[k] = Q.range_lo[k]; row_hi[k] = Q.range_hi[k]; }
  ED_CATCH
}

// ---------------------------------------------------------------- distributed vectors
int ed_dvec_create(ed_sharded* sh, ed_dvec** out) {
  ED_TRY
  ED_REQUIRE(sh && out, ED_ERR_ARGUMENT, "null argument");
  ed_ctx* c = sh->ctx;
  DeviceGuard g;
  std::unique_ptr<ed_dvec> v(new ed_dvec());
  v->sh = sh;
  const int nl = (int)c->local.size();
  v->local.assign(nl, nullptr);
  for (int i = 0; i < nl; ++i) {
    ED_CUDA(cudaSetDevice(c->local[i].device));
    const size_t bytes = (size_t)std::max<int64_t>(sh->r[i].n_local, 1) * sh->es;
    ED_CUDA(cudaMalloc(&v->local[i], bytes));
    ED_CUDA(cudaMemsetAsync(v->local[i], 0, bytes, c->local[i].stream));
  }
  ctx_sync(c);
  *out = v.release();
  ED_CATCH
}

int ed_dvec_destroy(ed_dvec* v) {
  ED_TRY
  if (!v) return ED_OK;
  ed_ctx* c = v->sh->ctx;
  DeviceGuard g;
  ctx_sync(c);
  for (size_t i = 0; i < v->local.size(); ++i) {
    cudaSetDevice(c->local[i].device);
    if (v->local[i]) cudaFree(v->local[i]);
  }
  delete v;
  ED_CATCH
}

int ed_dvec_local(ed_dvec* v, int32_t local_index, void** ptr, int64_t* n_local) {
  ED_TRY
  ED_REQUIRE(v && local_index >= 0 && local_index < (int)v->local.size(), ED_ERR_ARGUMENT, "bad arguments");
  if (ptr) *ptr = v->local[local_index];
  if (n_local) *n_local = v->sh->r[local_index].n_local;
  ED_CATCH
}

int ed_dvec_randn(ed_dvec* v, uint64_t seed, double scale) {
  ED_TRY
  ED_REQUIRE(v, ED_ERR_ARGUMENT, "null argument");
  ed_sharded* S = v->sh;
  DeviceGuard g;
  for (size_t i = 0; i < v->local.size(); ++i) {
    RankScope scope(S->ctx->local[i]);
    const ShardRank& Q = S->r[i];
    int64_t off = 0;
    for (size_t k = 0; k < Q.range_lo.size(); ++k) {
      const int64_t n = Q.range_hi[k] - Q.range_lo[k];
      if (n > 0) {
        const int rc = ed_vector_randn_async(static_cast<char*>(v->local[i]) + (size_t)off * S->es, n, S->dtype, seed, Q.range_lo[k]);
        ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
      }
      off += n;
    }
    if (scale != 1.0 && Q.n_local > 0) {
      const int rc = ed_vector_scale_async(v->local[i], Q.n_local, S->dtype, scale);
      ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
    }
  }
  ED_CATCH
}

/* host_full: full-length vector in the API (ascending basis) order; every process copies the rows its ranks own */
int ed_dvec_upload(ed_dvec* v, const void* host_full) {
  ED_TRY
  ED_REQUIRE(v && host_full, ED_ERR_ARGUMENT, "null argument");
  ed_sharded* S = v->sh;
  DeviceGuard g;
  for (size_t i = 0; i < v->local.size(); ++i) {
    CtxRank& R = S->ctx->local[i];
    ED_CUDA(cudaSetDevice(R.device));
    const ShardRank& Q = S->r[i];
    int64_t off = 0;
    for (size_t k = 0; k < Q.range_lo.size(); ++k) {
      const int64_t n = Q.range_hi[k] - Q.range_lo[k];
      if (n > 0) ED_CUDA(cudaMemcpyAsync(static_cast<char*>(v->local[i]) + (size_t)off * S->es, static_cast<const char*>(host_full) + (size_t)Q.range_lo[k] * S->es,
                                         (size_t)n * S->es, cudaMemcpyHostToDevice, R.stream));
      off += n;
    }
  }
  ctx_sync(S->ctx);
  ED_CATCH
}

/* the rows owned by this process's ranks are written into host_full (API order); the other rows are left untouched */
int ed_dvec_download(ed_dvec* v, void* host_full) {
  ED_TRY
  ED_REQUIRE(v && host_full, ED_ERR_ARGUMENT, "null argument");
  ed_sharded* S = v->sh;
  DeviceGuard g;
  for (size_t i = 0; i < v->local.size(); ++i) {
    CtxRank& R = S->ctx->local[i];
    ED_CUDA(cudaSetDevice(R.device));
    const ShardRank& Q = S->r[i];
    int64_t off = 0;
    for (size_t k = 0; k < Q.range_lo.size(); ++k) {
      const int64_t n = Q.range_hi[k] - Q.range_lo[k];
      if (n > 0) ED_CUDA(cudaMemcpyAsync(static_cast<char*>(host_full) + (size_t)Q.range_lo[k] * S->es, static_cast<const char*>(v->local[i]) + (size_t)off * S->es,
                                         (size_t)n * S->es, cudaMemcpyDeviceToHost, R.stream));
      off += n;
    }
  }
  ctx_sync(S->ctx);
  ED_CATCH
}

// ---------------------------------------------------------------- matvec / Lanczos
int ed_apply_sharded(ed_sharded* sh, ed_dvec* y, ed_dvec* x, int32_t no_fence, double* dot_out) {
  ED_TRY
  ED_REQUIRE(sh && y && x && y->sh == sh && x->sh == sh && y != x, ED_ERR_ARGUMENT, "bad arguments");
  (void)no_fence;
  sharded_apply(sh, y, x, false, dot_out != nullptr);
  if (dot_out) {
    CtxRank& R = sh->ctx->local[0];
    DeviceGuard g;
    ED_CUDA(cudaSetDevice(R.device));
    ED_CUDA(cudaMemcpyAsync(dot_out, sh->r[0].dot.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, R.stream));
    ED_CUDA(cudaStreamSynchronize(R.stream));
    check_wait_error(sh);
  }
  ED_CATCH
}

/* One matvec taken apart (halo exchange only): ms[0] = owner-side pack, ms[1] = fence (tiny all-reduce), ms[2] = all peer
 * copies with nothing else running, ms[3] = all kernel chunks with the halo already in place; each the max over ranks.
 * The phases overlap in ed_apply_sharded; here they run back to back to show what bounds it. */
int ed_sharded_profile(ed_sharded* sh, ed_dvec* y, ed_dvec* x, double* ms4) {
  ED_TRY
  ED_REQUIRE(sh && y && x && ms4 && y->sh == sh && x->sh == sh && y != x, ED_ERR_ARGUMENT, "bad arguments");
  ED_REQUIRE(sh->halo, ED_ERR_UNSUPPORTED, "only the halo exchange has separate phases");
  ed_ctx* c = sh->ctx;
  DeviceGuard g;
  auto kernels = [&]() {
    for (size_t i = 0; i < c->local.size(); ++i) {
      ShardRank& Q = sh->r[i];
      RankScope scope(c->local[i]);
      for (int ch = 0; ch < Q.L.n_chunks; ++ch) {
        U1ShardLaunch A;
        A.tile_H = Q.d_tile_H.p; A.first = Q.L.chunk_first[ch]; A.count = Q.L.chunk_first[ch + 1] - Q.L.chunk_first[ch];
        A.dir = Q.d_dir.p; A.x_local = x->local[i]; A.x_halo = sh->push ? Q.halo_mem : (void*)Q.halo.p; A.y_local = y->local[i];
        A.stream_mode = 0; A.accumulate = 0; A.partials = nullptr;
        ed_apply_u1_sharded(Q.op, sh->dtype, A);
      }
    }
  };
  ctx_fence(c);
  ed_ctx_timer_record(c, 50);
  if (sh->nccl_p2p) {
    ed_ctx_timer_record(c, 51);
    ed_ctx_timer_record(c, 52);
    sharded_pack(sh, x);                 // pack + the send/recv groups
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ED_CUDA(cudaSetDevice(R.device));
      ED_CUDA(cudaEventRecord(R.ev, R.push));
      ED_CUDA(cudaStreamWaitEvent(R.stream, R.ev, 0));
    }
  } else if (sh->push) {
    ed_ctx_timer_record(c, 51);          // no pack pass
    ctx_fence(c);
    ed_ctx_timer_record(c, 52);
    sharded_pack(sh, x);                 // the push kernels
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = sh->r[i];
      RankScope scope(R);
      for (int ch = 0; ch < Q.L.n_chunks; ++ch)
        if (c->world > 1 && Q.L.chunk_halo_rows[ch] > 0)
          ED_LAUNCH(k_wait_rows, 1, 32, 0, Q.counters + ch, Q.seq * (unsigned long long)Q.L.chunk_halo_rows[ch], Q.wait_error.p);
      ED_CUDA(cudaEventRecord(R.ev, R.push));
      ED_CUDA(cudaStreamWaitEvent(R.stream, R.ev, 0));
    }
  } else {
    const bool side = sh->side;          // the phases are timed one after another on the main stream
    sh->side = false;
    sharded_pack(sh, x);
    sh->side = side;
    ed_ctx_timer_record(c, 51);
    ctx_fence(c);
    ed_ctx_timer_record(c, 52);
    for (size_t i = 0; i < c->local.size(); ++i) {
      CtxRank& R = c->local[i];
      ShardRank& Q = sh->r[i];
      ED_CUDA(cudaSetDevice(R.device));
      ED_CUDA(cudaEventRecord(R.ev, R.stream));
      ED_CUDA(cudaStreamWaitEvent(R.copy[0], R.ev, 0));
      for (const U1Pull& p : Q.L.pulls)
        ED_CUDA(cudaMemcpyAsync(Q.halo.p + (size_t)p.dst_off * sh->es, static_cast<const char*>(Q.peer_send[sh->parity][p.peer]) + (size_t)p.src_off * sh->es,
                                (size_t)p.len * sh->es, cudaMemcpyDeviceToDevice, R.copy[0]));
      ED_CUDA(cudaEventRecord(R.ev, R.copy[0]));
      ED_CUDA(cudaStreamWaitEvent(R.stream, R.ev, 0));
    }
  }
  ed_ctx_timer_record(c, 53);
  kernels();
  ed_ctx_timer_record(c, 54);
  for (int k = 0; k < 4; ++k) {
    const int rc = ed_ctx_timer_elapsed(c, 50 + k, 51 + k, ms4 + k);
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
  }
  check_wait_error(sh);
  ED_CATCH
}

int ed_lanczos_sharded(ed_sharded* sh, int32_t n_steps, uint64_t seed, ed_dvec* v0, double* alpha, double* beta, double* ritz,
                       int32_t n_ritz, int32_t* steps_done, double* ms_per_step) {
  ED_TRY
  ED_REQUIRE(sh && alpha && beta && n_steps >= 1, ED_ERR_ARGUMENT, "bad arguments");
  ed_ctx* c = sh->ctx;
  DeviceGuard g;
  const int nl = (int)c->local.size();
  ctx_fence(c);       // every rank has left its previous matvec: halos and send buffers may be written again
  ed_dvec *u_cur = nullptr, *u_prev = nullptr, *w = nullptr;
  auto mk = [&](ed_dvec** v) { const int rc = ed_dvec_create(sh, v); ED_REQUIRE(rc == ED_OK, rc, ed_last_error()); };
  mk(&u_cur); mk(&u_prev); mk(&w);
  struct Cleanup { ed_dvec *a, *b, *c; ~Cleanup() { ed_dvec_destroy(a); ed_dvec_destroy(b); ed_dvec_destroy(c); } } cleanup{u_cur, u_prev, w};
  std::vector<DevBuf<double>> dots(nl), norms(nl);
  for (int i = 0; i < nl; ++i) {
    RankScope scope(c->local[i]);
    dots[i].alloc((size_t)2 * n_steps);
    norms[i].alloc((size_t)2 * (n_steps + 1));
  }
  if (v0) {
    ED_REQUIRE(v0->sh == sh, ED_ERR_ARGUMENT, "start vector of another representation");
    for (int i = 0; i < nl; ++i) {
      ED_CUDA(cudaSetDevice(c->local[i].device));
      ED_CUDA(cudaMemcpyAsync(u_cur->local[i], v0->local[i], (size_t)sh->r[i].n_local * sh->es, cudaMemcpyDeviceToDevice, c->local[i].stream));
    }
  } else {
    const int rc = ed_dvec_randn(u_cur, seed, 1.0);
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
  }
  auto allreduce_at = [&](std::vector<DevBuf<double>>& arr, size_t at) {
    std::vector<double*> bufs;
    for (int i = 0; i < nl; ++i) bufs.push_back(arr[i].p + at);
    ctx_allreduce(c, bufs, 2, 0);
  };
  for (int i = 0; i < nl; ++i) {
    RankScope scope(c->local[i]);
    const int rc = ed_vector_norm2_async(u_cur->local[i], sh->r[i].n_local, sh->dtype, norms[i].p);
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
    loop_chain(c, i);
  }
  // pull / push transports: the exchange is started BEFORE the all-reduce, which doubles as its fence; grouped send/recv
  // shares the communicator with the all-reduce and NCCL runs a communicator's operations in issue order, so there the
  // all-reduce goes first (it would otherwise wait for the whole halo transfer, and the interior kernels behind it too)
  const bool pack_after = sh->nccl_p2p || sh->side;      // these issue their own collective on the side stream
  if (!pack_after) sharded_pack(sh, u_cur);
  allreduce_at(norms, 0);
  if (pack_after) sharded_pack(sh, u_cur);
  ed_ctx_timer_record(c, ED_CTX_TIMER_SLOTS - 2);
  for (int j = 0; j < n_steps; ++j) {
    // the peers' send buffers hold their tiles of u_cur: packed before the all-reduce of norms[j] (stream ordered)
    sharded_apply(sh, w, u_cur, true, true);
    for (int i = 0; i < nl; ++i) {
      ED_CUDA(cudaSetDevice(c->local[i].device));
      ED_CUDA(cudaMemcpyAsync(dots[i].p + 2 * j, sh->r[i].dot.p, 2 * sizeof(double), cudaMemcpyDeviceToDevice, c->local[i].stream));
    }
    for (int i = 0; i < nl; ++i) {
      RankScope scope(c->local[i]);
      const int rc = ed_lanczos_update_async(u_prev->local[i], w->local[i], u_cur->local[i], sh->r[i].n_local, sh->dtype, dots[i].p + 2 * j,
                                             norms[i].p + 2 * j, j > 0 ? norms[i].p + 2 * (j - 1) : nullptr, norms[i].p + 2 * (j + 1));
      ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
      loop_chain(c, i);
    }
    std::swap(u_cur, u_prev);
    if (j + 1 < n_steps && !pack_after) sharded_pack(sh, u_cur);
    allreduce_at(norms, (size_t)2 * (j + 1));
    if (j + 1 < n_steps && pack_after) sharded_pack(sh, u_cur);
  }
  ed_ctx_timer_record(c, ED_CTX_TIMER_SLOTS - 1);
  if (ms_per_step) {
    double ms = 0;
    const int rc = ed_ctx_timer_elapsed(c, ED_CTX_TIMER_SLOTS - 2, ED_CTX_TIMER_SLOTS - 1, &ms);
    ED_REQUIRE(rc == ED_OK, rc, ed_last_error());
    *ms_per_step = ms / n_steps;
  }
  std::vector<double> hd((size_t)2 * n_steps), hn((size_t)2 * (n_steps + 1));
  {
    RankScope scope(c->local[0]);
    dots[0].download(hd.data(), hd.size());
    norms[0].download(hn.data(), hn.size());
  }
  ctx_sync(c);
  check_wait_error(sh);
  const int done = ed_lanczos_finish(hd.data(), hn.data(), n_steps, alpha, beta, ritz, n_ritz);
  if (steps_done) *steps_done = done;
  ED_CATCH
}

}  // extern "C"
