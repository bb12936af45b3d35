// Shared helpers for the vlfm_b200 CUDA library (sm_100a only).
#pragma once
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

__device__ __forceinline__ float ld_cg_f32(const float* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) { return __ldcg(p); }

}  // namespace vlfm
