// Error channel + version of the C-ABI library.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

int tatt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

extern "C" {
const char* tatt_last_error(void) { return g_err; }
int tatt_version(void) { return 100; }
int tatt_arch(void) {
#if defined(TATT_SM100A)
  return 1;
#else
  return 0;
#endif
}
}
