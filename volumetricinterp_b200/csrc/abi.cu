// vi_version / vi_last_error and the error sink shared by all translation units.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void vi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* vi_last_error(void) { return g_err; }
extern "C" const char* vi_version(void) { return "volinterp_b200 0.1.0 (sm_100a)"; }
