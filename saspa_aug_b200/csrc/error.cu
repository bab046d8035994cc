// Error string + version entry points of the C ABI (thread-local last error, never throws).
#include "common.cuh"
#include "../../include/saspa_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void saspa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* saspa_last_error_string(void) { return g_err; }
extern "C" int saspa_version(void) { return SASPA_B200_VERSION; }
