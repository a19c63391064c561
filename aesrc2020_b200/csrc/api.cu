// api.cu -- error channel and library identification for libsarnet_sm100.so
#include <stdlib.h>
#include <mutex>
#include <utility>
#include <vector>
#include "common.cuh"

namespace sar {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int allow_max_smem_impl(const void* fn, const char* what) {
  static std::mutex mu;
  static std::vector<std::pair<int, const void*>> done;      // (device, function) pairs already raised
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  for (const auto& d : done)
    if (d.first == dev && d.second == fn) return SAR_OK;
  // the opt-in limit covers static + dynamic shared memory: leave room for the kernel's static allocations
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, fn);
  int optin = 0;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) {
    long long lim = (long long)optin - (long long)fa.sharedSizeBytes;
    if (lim > (long long)SAR_MAX_DYN_SMEM) lim = (long long)SAR_MAX_DYN_SMEM;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim);
  }
  if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(max dynamic shared memory): %s", what, cudaGetErrorString(e)); return (int)e; }
  done.emplace_back(dev, fn);
  return SAR_OK;
}
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SAR_NO_PDL"); on = (e && e[0] && e[0] != '0') ? 0 : 1; }
  return on != 0;
}
}  // namespace sar

extern "C" {

int sar_version(void) { return 100; }   // 0.1.0

const char* sar_last_error(void) { return sar::g_err; }

int sar_compiled_arch(void) {
#ifdef SAR_COMPILED_ARCH
  return SAR_COMPILED_ARCH;
#else
  return 100;
#endif
}

}  // extern "C"
