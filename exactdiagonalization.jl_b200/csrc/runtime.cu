// Library runtime: error state, device/stream selection, host<->device staging.
#include "ed_internal.cuh"

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launch_count{0};

static thread_local cudaStream_t t_user_stream = nullptr;
static thread_local bool t_use_user_stream = false;
static cudaStream_t g_own_stream[64] = {nullptr};

void ed_set_error(const std::string& msg) { t_last_error = msg; }

// internal, nestable by one level: the multi-GPU context runs each rank's work on that rank's stream
static thread_local cudaStream_t t_saved_stream = nullptr;
static thread_local bool t_saved_use = false;
void ed_push_stream(cudaStream_t s) {
  t_saved_stream = t_user_stream; t_saved_use = t_use_user_stream;
  t_user_stream = s; t_use_user_stream = true;
}
void ed_pop_stream() { t_user_stream = t_saved_stream; t_use_user_stream = t_saved_use; }

void ed_require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    throw EdError(ED_ERR_CUDA, "no CUDA device available: libedcuda has no CPU fallback");
  }
}

cudaStream_t ed_stream() {
  if (t_use_user_stream) return t_user_stream;
  int dev = 0;
  ED_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) throw EdError(ED_ERR_CUDA, "device index out of range");
  if (!g_own_stream[dev]) ED_CUDA(cudaStreamCreateWithFlags(&g_own_stream[dev], cudaStreamNonBlocking));
  return g_own_stream[dev];
}

bool ed_is_device_pointer(const void* p) {
  if (!p) return false;
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int ed_current_device() {
  int dev = 0;
  ED_CUDA(cudaGetDevice(&dev));
  return dev;
}

int ed_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  ED_CUDA(cudaGetDevice(&dev));
  if (!cached[dev]) ED_CUDA(cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev));
  return cached[dev];
}

// Staging buffers for host vectors are cached per thread (the two largest are kept): a Krylov solver calling
// mul! with host arrays would otherwise pay cudaMalloc/cudaFree of gigabytes on every call.
namespace {
struct StageSlot { void* p = nullptr; size_t cap = 0; bool busy = false; };
thread_local StageSlot t_stage[2];
void* stage_acquire(size_t nbytes) {
  for (auto& s : t_stage)
    if (!s.busy && s.cap >= nbytes) { s.busy = true; return s.p; }
  for (auto& s : t_stage)
    if (!s.busy) {
      if (s.p) cudaFree(s.p);
      s.p = nullptr; s.cap = 0;
      ED_CUDA(cudaMalloc(&s.p, nbytes));
      s.cap = nbytes; s.busy = true;
      return s.p;
    }
  return nullptr;
}
bool stage_release(void* p) {
  for (auto& s : t_stage)
    if (s.p == p && s.busy) { s.busy = false; return true; }
  return false;
}
}  // namespace

Staged::Staged(const void* p, size_t nbytes, bool copy_in, bool copy_out) : bytes(nbytes), writeback(copy_out) {
  if (nbytes == 0 || p == nullptr) { dev = const_cast<void*>(p); return; }
  if (ed_is_device_pointer(p)) {
    dev = const_cast<void*>(p);
    owned = false;
    writeback = false;
    return;
  }
  host = const_cast<void*>(p);
  owned = true;
  dev = stage_acquire(nbytes);
  if (!dev) ED_CUDA(cudaMalloc(&dev, nbytes));
  if (copy_in) ED_CUDA(cudaMemcpyAsync(dev, host, nbytes, cudaMemcpyHostToDevice, ed_stream()));
}

void Staged::finish() {
  if (owned && dev) {
    if (writeback) {
      ED_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ed_stream()));
      ED_CUDA(cudaStreamSynchronize(ed_stream()));
    }
    if (!stage_release(dev)) cudaFree(dev);
    dev = nullptr;
    owned = false;
  }
}

Staged::~Staged() {
  if (owned && dev) {
    cudaStreamSynchronize(ed_stream());
    if (!stage_release(dev)) cudaFree(dev);
  }
}

extern "C" {

const char* ed_last_error(void) { return t_last_error.c_str(); }
const char* ed_version(void) { return "edcuda 0.1.0 (sm_100a)"; }

int ed_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ed_set_device(int device) {
  ED_TRY
  ed_require_device();
  ED_CUDA(cudaSetDevice(device));
  ED_CATCH
}

int ed_set_stream(void* cuda_stream, int32_t enable) {
  ED_TRY
  t_use_user_stream = enable != 0;
  t_user_stream = enable ? reinterpret_cast<cudaStream_t>(cuda_stream) : nullptr;
  ED_CATCH
}

int64_t ed_kernel_launch_count(void) { return g_launch_count.load(); }

// ---- plain device buffers owned by the caller ----
int ed_device_malloc(int64_t bytes, void** ptr) {
  ED_TRY
  ED_REQUIRE(ptr && bytes >= 0, ED_ERR_ARGUMENT, "bad arguments");
  ed_require_device();
  *ptr = nullptr;
  if (bytes > 0) ED_CUDA(cudaMalloc(ptr, (size_t)bytes));
  ED_CATCH
}

int ed_device_free(void* ptr) {
  ED_TRY
  if (ptr) ED_CUDA(cudaFree(ptr));
  ED_CATCH
}

int ed_release_staging(void) {
  ED_TRY
  for (auto& s : t_stage)
    if (!s.busy && s.p) { cudaFree(s.p); s.p = nullptr; s.cap = 0; }
  ED_CATCH
}

}  // extern "C"
