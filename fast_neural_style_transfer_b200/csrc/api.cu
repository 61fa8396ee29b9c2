// Library-level entry points: version, thread-local error message.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace fnst {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
Tuning& tuning() {
  static Tuning t = [] {
    Tuning v;
    if (const char* e = getenv("FNST_CONV_BLOCK_N")) v.conv_block_n = atoi(e);
    if (const char* e = getenv("FNST_WGRAD_WAVES_X2")) v.wgrad_waves_x2 = atoi(e);
    if (const char* e = getenv("FNST_WGRAD_BN")) v.wgrad_bn = atoi(e);
    if (const char* e = getenv("FNST_PDL")) v.pdl = atoi(e);
    if (const char* e = getenv("FNST_CONV_PAIR")) v.conv_pair = atoi(e);
    if (const char* e = getenv("FNST_CONV_STAGE_OUT")) v.conv_stage_out = atoi(e);
    if (const char* e = getenv("FNST_CONV_ROWSTREAM")) v.conv_rowstream = atoi(e);
    if (const char* e = getenv("FNST_INORM_BWD_TMA")) v.inorm_bwd_tma = atoi(e);
    if (const char* e = getenv("FNST_INORM_BWD_BLOCKS")) v.inorm_bwd_blocks = atoi(e);
    if (const char* e = getenv("FNST_RESIZE_STAGED")) v.resize_staged = atoi(e);
    return v;
  }();
  return t;
}
bool pdl_enabled() { return tuning().pdl != 0; }
}  // namespace fnst

extern "C" int fnst_version(void) { return 101; }

extern "C" int fnst_set_debug_buffer(void* device_ptr) {
  fnst::tuning().debug_buf = reinterpret_cast<unsigned long long*>(device_ptr);
  return 0;
}

extern "C" int fnst_set_tuning(const char* name, int value) {
  fnst::Tuning& t = fnst::tuning();
  if (!strcmp(name, "conv_block_n")) t.conv_block_n = value;
  else if (!strcmp(name, "wgrad_waves_x2")) t.wgrad_waves_x2 = value;
  else if (!strcmp(name, "wgrad_bn")) t.wgrad_bn = value;
  else if (!strcmp(name, "pdl")) t.pdl = value;
  else if (!strcmp(name, "conv_pair")) t.conv_pair = value;
  else if (!strcmp(name, "conv_stage_out")) t.conv_stage_out = value;
  else if (!strcmp(name, "conv_rowstream")) t.conv_rowstream = value;
  else if (!strcmp(name, "inorm_bwd_tma")) t.inorm_bwd_tma = value;
  else if (!strcmp(name, "inorm_bwd_blocks")) t.inorm_bwd_blocks = value;
  else if (!strcmp(name, "resize_staged")) t.resize_staged = value;
  else if (!strcmp(name, "dbg_mode")) t.dbg_mode = value;
  else { fnst::set_error("fnst_set_tuning: unknown knob '%s'", name); return -1; }
  return 0;
}
extern "C" const char* fnst_last_error(void) { return fnst::g_err; }
