// Error reporting and ABI versioning for libgeomae_b200.so.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void gm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* geomae_last_error(void) { return g_err; }
extern "C" int geomae_abi_version(void) { return 2; }

static thread_local bool g_weights_stable = false;
void gm_set_weights_stable(bool stable) { g_weights_stable = stable; }
bool gm_weights_stable() { return g_weights_stable; }
