// api.cu -- error reporting and version of libumhs_sm100.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace umhs {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

}  // namespace umhs

extern "C" const char *umhs_last_error(void) { return umhs::g_error; }

extern "C" int umhs_version(void) { return 100; }
