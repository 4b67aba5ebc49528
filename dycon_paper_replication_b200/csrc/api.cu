// Error reporting, version and device checks of the C ABI (include/dycon_b200.h).
#include "common.cuh"

#include <atomic>
#include <cstring>

namespace dycon {

char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t err, const char* what) {
  snprintf(error_buffer(), 512, "CUDA error %d (%s) at %s", (int)err, cudaGetErrorString(err), what);
  return (int)err;
}

static std::atomic<uint64_t> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
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

}  // namespace dycon

extern "C" {

int dycon_abi_version(void) { return DYCON_ABI_VERSION; }

const char* dycon_last_error(void) { return dycon::error_buffer(); }

uint64_t dycon_launch_count(void) { return dycon::launches(); }

int dycon_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  DYCON_CUDA(cudaGetDevice(&dev));
  DYCON_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DYCON_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  DYCON_REQUIRE(major == 10, DYCON_ERR_DEVICE,
                "device %d has compute capability %d.%d; this library is built for sm_100a (B200) only",
                dev, major, minor);
  return DYCON_OK;
}

}  // extern "C"
