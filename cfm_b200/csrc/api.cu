// Library-level entry points and the error/devinfo plumbing shared by every kernel file.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cfm {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct DevInfo { int dev = -1, sms = 0, cc = 0; };
static thread_local DevInfo g_dev;
static void refresh_dev() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  if (dev == g_dev.dev) return;
  int sms = 0, maj = 0, min = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
  g_dev.dev = dev; g_dev.sms = sms; g_dev.cc = maj * 10 + min;
}
int sm_count() { refresh_dev(); return g_dev.sms > 0 ? g_dev.sms : 148; }
int cc_major_minor() { refresh_dev(); return g_dev.cc; }

}  // namespace cfm

extern "C" long long cfm_launch_count(void) { return cfm::g_launches.load(); }
extern "C" int cfm_abi_version(void) { return CFM_ABI_VERSION; }
extern "C" const char* cfm_last_error(void) { return cfm::g_err; }
extern "C" int cfm_device_info(int* sm_count, int* cc) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cfm::set_error("cfm_device_info: no CUDA device (%s)", cudaGetErrorString(e));
    return CFM_ERR_CUDA;
  }
  if (sm_count) *sm_count = cfm::sm_count();
  if (cc) *cc = cfm::cc_major_minor();
  if (cfm::cc_major_minor() / 10 != 10) {
    cfm::set_error("cfm_b200 is built for sm_100a only; device reports sm_%d", cfm::cc_major_minor());
    return CFM_ERR_ARCH;
  }
  return CFM_OK;
}
