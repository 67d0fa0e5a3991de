// api.cu — error reporting and ABI version of libgsplat_b200.so
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gs {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace gs

extern "C" {

int gs_abi_version(void) { return 5; }   // 5: *_counted backward / stage / capped emit, multimem all-reduce

const char* gs_last_error_string(void) { return gs::g_error; }

}  // extern "C"
