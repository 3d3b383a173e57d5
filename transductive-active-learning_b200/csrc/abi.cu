// Library-level entry points of include/tal_b200.h: version, error text,
// launch accounting, device check.
#include "common.cuh"
#include <atomic>
#include <cstdarg>

namespace tal {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace tal

extern "C" {

int tal_abi_version(void) { return TAL_ABI_VERSION; }

const char* tal_last_error_string(void) { return tal::g_err; }

long long tal_launch_count(void) { return tal::g_launches.load(std::memory_order_relaxed); }

void tal_reset_launch_count(void) { tal::g_launches.store(0, std::memory_order_relaxed); }

int tal_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    tal::set_error("tal_device_ok: no CUDA device");
    cudaGetLastError();
    return 0;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (major != 10) {
    tal::set_error("tal_device_ok: built for sm_100a, device is compute capability %d.x", major);
    return 0;
  }
  return 1;
}

}  // extern "C"
