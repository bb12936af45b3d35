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

#ifdef VLFM_DEV_PROBES   // development builds only (scripts/pdl_probe.py): not part of the shipped C-ABI
// ---- development probe: does programmatic dependent launch overlap kernels (in streams / in graphs)?
namespace vlfm {
__global__ void pdl_probe_kernel(int pre_ns, int post_ns, int* sink) {
  pdl_trigger();
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)pre_ns);
  pdl_wait();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)post_ns);
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sink, 1);
}
}  // namespace vlfm
extern "C" int vlfm_pdl_probe(int blocks, int smem_bytes, int pre_ns, int post_ns, int* d_sink, void* stream) {
  static int cfg = 0;
  if (smem_bytes > 48 * 1024 && smem_bytes > cfg) {
    int rc = vlfm::check_cuda(cudaFuncSetAttribute(vlfm::pdl_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), "attr(pdl_probe)");
    if (rc) return rc; cfg = smem_bytes;
  }
  return vlfm::check_cuda(vlfm::launch_pdl(vlfm::pdl_probe_kernel, dim3(blocks), dim3(128), (size_t)smem_bytes, (cudaStream_t)stream, pre_ns, post_ns, d_sink), "pdl_probe_kernel");
}
#endif
