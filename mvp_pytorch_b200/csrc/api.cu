// Error plumbing and ABI version of libmvptr_b200.so.
#include <stdarg.h>

#include "common.cuh"

namespace mvptr {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mvptr

extern "C" int mvptr_abi_version(void) { return MVPTR_ABI_VERSION; }
extern "C" const char* mvptr_last_error(void) { return mvptr::g_err; }
