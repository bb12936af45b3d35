// Shared helpers for the vlfm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vlfm_b200.h"

namespace vlfm {

void set_error(const char* fmt, ...);
void count_launch(unsigned n = 1);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VLFM_E_CUDA;
  }
  return VLFM_OK;
}

#define VLFM_CHECK_LAUNCH(what)                                   \
  do {                                                            \
    int _rc = ::vlfm::check_cuda(cudaGetLastError(), what);       \
    if (_rc != VLFM_OK) return _rc;                               \
  } while (0)

// Programmatic dependent launch (PDL): every kernel of the per-step sequence lets its successor start
// its prologue early (launch_dependents) and waits for its predecessors' memory before touching global
// memory (wait).  Hides launch latency + prologue of the ~400 small kernels of a batch-1 forward.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// LayerNorm of one row by one warp, reading the row from L2 (ld.cg) three times instead of holding it in
// registers (used after a grid-wide barrier inside the residual GEMM, where register pressure matters).
__device__ __forceinline__ void ln_row_warp(const float* __restrict__ xr, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, __half* o16, float* o32, int D, float eps, int lane) {
  const int D4 = D >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(xr);
  float s = 0.f;
  for (int j = lane; j < D4; j += 32) { const float4 v = __ldcg(x4 + j); s += (v.x + v.y) + (v.z + v.w); }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.f;
  for (int j = lane; j < D4; j += 32) {
    const float4 v = __ldcg(x4 + j);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)D + eps);
  for (int j = lane; j < D4; j += 32) {
    const float4 v = __ldcg(x4 + j);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j), bt = __ldg(reinterpret_cast<const float4*>(beta) + j);
    float4 y;
    y.x = (v.x - mean) * rstd * g.x + bt.x; y.y = (v.y - mean) * rstd * g.y + bt.y;
    y.z = (v.z - mean) * rstd * g.z + bt.z; y.w = (v.w - mean) * rstd * g.w + bt.w;
    if (o16) {
      __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      *reinterpret_cast<uint2*>(o16 + 4 * j) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
    if (o32) *reinterpret_cast<float4*>(o32 + 4 * j) = y;
  }
}

__device__ __forceinline__ float ld_cg_f32(const float* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) { return __ldcg(p); }

}  // namespace vlfm
