// Library-wide plumbing: version, thread-local error string, device info.
#include "mq_common.cuh"
#include "mq_tc.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace mq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return MQ_ERR_CUDA;
}

int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

int ensure_dyn_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;  // (kernel, device) -> bytes granted
  int dev = 0;
  MQ_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int& have = done[std::make_pair(kernel, dev)];
  if (have >= bytes) return MQ_OK;
  MQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  have = bytes;
  return MQ_OK;
}

static std::atomic<int> g_force_hadamard{0};
int forced_hadamard_kernel() { return g_force_hadamard.load(std::memory_order_relaxed); }

#ifdef MQ_PROBES
namespace gemm {
unsigned probe_mask() {
  const char* e = getenv("MQ_GEMM_DEBUG");
  return e ? (unsigned)atoi(e) : 0u;
}
}  // namespace gemm
#endif

namespace tc {
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}
}  // namespace tc

}  // namespace mq

extern "C" int mq_version(void) { return MQ_VERSION; }

extern "C" int mq_debug_force_hadamard_kernel(int kind) {
  return mq::g_force_hadamard.exchange(kind >= 1 && kind <= 4 ? kind : 0);
}

extern "C" const char* mq_last_error(void) { return mq::g_err; }

extern "C" int mq_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MQ_CUDA(cudaGetDevice(&dev));
  int n = 0, ma = 0, mi = 0;
  MQ_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  MQ_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  MQ_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  return MQ_OK;
}
