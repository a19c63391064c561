// api.cu -- error channel and library identification for libsarnet_sm100.so
#include <stdlib.h>
#include "common.cuh"

namespace sar {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
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
