// Library-level C ABI: version, error string, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sgta {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace sgta

extern "C" int sgta_abi_version(void) { return 1; }
extern "C" const char* sgta_last_error(void) { return sgta::g_err; }
extern "C" int64_t sgta_launch_count(void) { return (int64_t)sgta::g_launches.load(); }
