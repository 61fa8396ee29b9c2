// Library-level entry points: version, thread-local error message.
#include "common.cuh"

namespace fnst {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace fnst

extern "C" int fnst_version(void) { return 100; }
extern "C" const char* fnst_last_error(void) { return fnst::g_err; }
