// Error reporting for the C ABI (include/cirs_b200.h): thread-local message + ABI version.
#include <string.h>
#include "common.cuh"
#include "../../include/cirs_b200.h"

static thread_local char g_err[512] = "";

void cirs_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

extern "C" const char* cirs_last_error(void) { return g_err; }
extern "C" int cirs_abi_version(void) { return CIRS_ABI_VERSION; }
