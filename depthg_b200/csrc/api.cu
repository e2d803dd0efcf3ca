// Library-level entry points: version and the per-thread error text.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>

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

// ---- optional per-kernel event timing (diagnostics for bench.py; single-threaded use) ----
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static cudaEvent_t g_pending = nullptr;
static cudaStream_t g_pending_stream = nullptr;

void profile_pre(cudaStream_t st) {
  if (!g_prof) return;
  cudaEventCreate(&g_pending);
  cudaEventRecord(g_pending, st);
  g_pending_stream = st;
}
void profile_post(const char* name) {
  if (!g_prof || !g_pending) return;
  cudaEvent_t b;
  cudaEventCreate(&b);
  cudaEventRecord(b, g_pending_stream);
  g_recs.push_back({name, g_pending, b});
  g_pending = nullptr;
}
}  // namespace dg

extern "C" int dg_version(void) { return 100; }
extern "C" const char* dg_last_error_string(void) { return dg::g_err; }
extern "C" unsigned long long dg_kernel_launches(void) { return dg::g_launches.load(std::memory_order_relaxed); }

extern "C" int dg_profile_enable(int on) {
  for (auto& r : dg::g_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  dg::g_recs.clear();
  dg::g_prof = on != 0;
  return DG_OK;
}

// Synchronises, then writes "name<TAB>launches<TAB>total_us\n" per kernel into buf; returns the bytes needed.
extern "C" size_t dg_profile_collect(char* buf, size_t buf_bytes) {
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : dg::g_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += (double)ms * 1e3;
    }
  }
  std::string out;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof line, "%s\t%d\t%.3f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && buf_bytes > 0) {
    size_t n = out.size() < buf_bytes - 1 ? out.size() : buf_bytes - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}

namespace dg {
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DEPTHG_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
}  // namespace dg

namespace dg { void set_clock_buffer(long long* p); }
// Debug: when non-NULL, the tcgen05 correlation kernel writes [grid][16] globaltimer stamps of its phases there.
extern "C" int dg_debug_set_clock_buffer(long long* dev_ptr) {
  dg::set_clock_buffer(dev_ptr);
  return DG_OK;
}

// A batch of device-to-device copies, copy i on streams[i]: what the sharded KNN build uses to pull its peers' panel
// rows out of their symmetric memory over NVLink.  cudaMemcpyAsync between device pointers runs on the copy engines -
// no SMs - and one call enqueues the lot (the per-copy host cost of a framework copy_ was what paced the exchange).
extern "C" int dg_memcpy_batch(int n, void* const* dst, const void* const* src, const size_t* bytes,
                               const dg_stream_t* streams) {
  using namespace dg;
  DG_REQUIRE(n >= 0 && (n == 0 || (dst && src && bytes && streams)), DG_ERR_INVALID, "dg_memcpy_batch: bad arguments");
  for (int i = 0; i < n; ++i) {
    if (bytes[i] == 0) continue;
    DG_REQUIRE(dst[i] && src[i], DG_ERR_INVALID, "dg_memcpy_batch: null pointer in copy %d", i);
    DG_CUDA_OK(cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(streams[i])));
  }
  return DG_OK;
}
