// Library-level entry points: version and the per-thread error text.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace dg {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace dg

extern "C" int dg_version(void) { return 100; }
extern "C" const char* dg_last_error_string(void) { return dg::g_err; }
extern "C" unsigned long long dg_kernel_launches(void) { return dg::g_launches.load(std::memory_order_relaxed); }
