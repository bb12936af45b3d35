// Error string, version and launch accounting for the vlfm_b200 C-ABI.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vlfm {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VLFM_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
}  // namespace vlfm

extern "C" const char* vlfm_last_error(void) { return vlfm::g_err; }
extern "C" int vlfm_version(void) { return 100; }
extern "C" unsigned long long vlfm_launch_count(void) {
  return vlfm::g_launches.load(std::memory_order_relaxed);
}
