// Library-level entry points: ABI version, per-thread error text, device properties.
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace smaq {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  // one immutable value per device ordinal; racing writers store the same number
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    return -1;
  }
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

static int dependent_launch_level() {
  static const int level = [] { const char* e = getenv("SMAQ_DEPENDENT_LAUNCH"); return e ? atoi(e) : 2; }();
  return level;
}
bool dependent_launch_enabled() { return dependent_launch_level() != 0; }

void set_first_kernel_dependent(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr) {
  attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr->val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dependent_launch_level() >= 2 ? 1 : 0;
}

void set_dependent_launch(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr) {
  attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr->val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dependent_launch_enabled() ? 1 : 0;
}

__global__ void counter_add_kernel(unsigned long long* counter, unsigned long long delta) { *counter += delta; }

}  // namespace smaq

extern "C" {
int smaq_counter_add(uint64_t* counter, uint64_t delta, smaq_stream_t stream) {
  if (!counter) return smaq::fail(SMAQ_ERR_ARG, "counter_add: null pointer");
  smaq::counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)counter, (unsigned long long)delta);
  SMAQ_LAUNCH_OK();
  return SMAQ_OK;
}
int smaq_b200_abi_version(void) { return SMAQ_B200_ABI_VERSION; }
const char* smaq_b200_last_error(void) { return smaq::last_error_buf(); }
int smaq_b200_sm_count(void) { return smaq::sm_count(); }
}
