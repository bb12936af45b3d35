// fp16 x fp16 -> fp32 GEMM on Blackwell 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   out[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias[N])      A, W row-major (K-major)
//
// Replaces the nn.Linear layers of the BLIP-2 ViT-g / Q-Former forward that
// vlfm/vlm/blip2itm.py:52 runs through lavis (fp16 autocast in the reference).
//
// Structure (one 128 x BN output tile per CTA, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of the A (128x64) and W
//               (BNx64) K-slices into a STAGES-deep 128B-swizzled smem ring,
//               mbarrier complete_tx signalling.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (kind::f16,
//               M=128, N=BN, K=16, fp32 accumulators in TMEM); tcgen05.commit
//               releases smem stages and finally signals the epilogue.
//   warps 2-9   epilogue (two warps per TMEM lane quarter, interleaved column chunks):
//               tcgen05.ld (32 lanes x 32 columns per instruction), fused
//               bias / GELU(erf) / fp32-residual-add, vectorised global stores.
// M / N / K tails are handled by TMA out-of-bounds zero fill + store guards.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace vlfm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 bytes = one swizzle atom row
constexpr int GEMM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();  // a protocol bug must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(1)<<16 | SBO(1024B>>4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// optional per-CTA timeline (development aid): vlfm_gemm_debug_timeline(ptr) makes CTA (0,0,0) of every
// launch record clock64() at its phase boundaries into ptr[0..7].
static unsigned long long* g_gemm_dbg = nullptr;

struct GemmArgs {
  const float* bias;
  void* out;
  int M, N, K, ldo, epi;
  int kb_per_split;   // K-blocks per grid.z slice (split-K, residual epilogue only)
  unsigned long long* dbg;
  // VLFM_EPI_PARTIAL_F32 (deterministic split-K): split z stores its tile at out + z * split_stride (plain stores, no atomics);
  // the consumer (layernorm_reduce_kernel) adds the partial sums to the residual stream in a fixed order
  long long split_stride;
  // "tail rows": when M = 128*q + r with 1 <= r <= GEMM_TAIL_MAX (ViT: 257 tokens), only q row tiles are launched and the CTAs
  // of the last one also compute the r extra rows on CUDA cores from the W tiles already staged for the tensor core
  const __half* a_tail; int lda, tail_rows, tail_row0;
  void* out_lo;       // VLFM_EPI_BIAS_GELU_F16X2: the x2 residual of the fp16 output (same ldo)
  // rows of the A tile the TMA box carries (32 / 64 / 128): with M <= 32 (Q-Former queries, text tokens) a 128-row box spends 3/4 of
  // every stage's TMA time on out-of-bounds zero fill.  Rows of the smem tile beyond the box keep whatever they held: row i of the
  // accumulator depends on row i of A only, and rows >= M are never stored.
  int a_box_rows;
};
constexpr int GEMM_TAIL_MAX = 2;
constexpr int GEMM_TAIL_KMAX = 6144;   // K elements of one CTA's slice that fit the tail-row staging buffer

__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// Epilogue of one 128-row accumulator slab: TMEM -> registers (32 lanes x 32 columns per tcgen05.ld),
// fused bias / GELU / residual, vectorised global stores.  `q` = TMEM lane quarter of this warp.
template <int BN, bool X2 = false>
__device__ __forceinline__ void epilogue_slab(uint32_t tmem_base, int q, int row, int n_blk, const GemmArgs& g, bool split,
                                              int c_first, const float* sbias = nullptr, uint32_t tmem_corr = 0) {
  // the two warps sharing a lane quarter interleave the 32-column chunks
#pragma unroll 1
  for (int c = c_first; c < BN / 32; c += 2) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
    if (X2) {   // x2 GEMM: main accumulator (hi.hi) + correction accumulator (lo.hi + hi.lo, scaled by 2048)
      uint32_t r2[32];
      tmem_ld32(tmem_corr + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r2);
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(r2[j]), 1.f / X2_SCALE, __uint_as_float(r[j])));
    }
    const int n0 = n_blk * BN + c * 32;
    if (row >= g.M || n0 >= g.N) continue;
    float v[32];
    const bool fullw = (n0 + 32 <= g.N);
    if (sbias) {   // bias tile staged in shared memory before the accumulator wait (zero-filled past N / without bias)
      const bool addb = blockIdx.z == 0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = addb ? *reinterpret_cast<const float4*>(sbias + c * 32 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[j] = __uint_as_float(r[j]) + b4.x; v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
        v[j + 2] = __uint_as_float(r[j + 2]) + b4.z; v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
      }
    } else if (g.bias && blockIdx.z == 0 && fullw) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + j));
        v[j] = __uint_as_float(r[j]) + b4.x; v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
        v[j + 2] = __uint_as_float(r[j + 2]) + b4.z; v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float b = (g.bias && blockIdx.z == 0 && n0 + j < g.N) ? __ldg(g.bias + n0 + j) : 0.f;
        v[j] = __uint_as_float(r[j]) + b;
      }
    }
    if (g.epi == VLFM_EPI_BIAS_F16 || g.epi == VLFM_EPI_BIAS_GELU_F16 || g.epi == VLFM_EPI_BIAS_RELU_F16 || g.epi == VLFM_EPI_BIAS_GELU_F16X2) {
      if (g.epi == VLFM_EPI_BIAS_GELU_F16 || g.epi == VLFM_EPI_BIAS_GELU_F16X2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      } else if (g.epi == VLFM_EPI_BIAS_RELU_F16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      __half* o = reinterpret_cast<__half*>(g.out) + (size_t)row * g.ldo + n0;
      if (fullw) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          __half2 h0 = __floats2half2_rn(v[j], v[j + 1]), h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
          __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]), h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(o + j) = pk;
          if (g.epi == VLFM_EPI_BIAS_GELU_F16X2) {
            const __half2 hh[4] = {h0, h1, h2, h3};
            uint32_t pl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(hh[e]);
              __half2 l = __floats2half2_rn((v[j + 2 * e] - f.x) * X2_SCALE, (v[j + 2 * e + 1] - f.y) * X2_SCALE);
              pl[e] = *reinterpret_cast<uint32_t*>(&l);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(g.out_lo) + (size_t)row * g.ldo + n0 + j) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (n0 + j < g.N) {   // static indices: keeps v[] in registers
          const __half hi = __float2half_rn(v[j]);
          o[j] = hi;
          if (g.epi == VLFM_EPI_BIAS_GELU_F16X2) reinterpret_cast<__half*>(g.out_lo)[(size_t)row * g.ldo + n0 + j] = __float2half_rn((v[j] - __half2float(hi)) * X2_SCALE);
        }
      }
    } else {
      const bool partial = (g.epi == VLFM_EPI_PARTIAL_F32);
      float* o = reinterpret_cast<float*>(g.out) + (partial ? (size_t)blockIdx.z * (size_t)g.split_stride : 0) + (size_t)row * g.ldo + n0;
      const bool add = (g.epi == VLFM_EPI_BIAS_RESID_F32);
      if (fullw) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 t = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (split && !partial) { red_add_f32x4(o + j, t.x, t.y, t.z, t.w); continue; }
          if (add) {
            float4 old = *reinterpret_cast<const float4*>(o + j);
            t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
          }
          *reinterpret_cast<float4*>(o + j) = t;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (n0 + j < g.N) {
            if (split && !partial) atomicAdd(o + j, v[j]);
            else o[j] = add ? o[j] + v[j] : v[j];
          }
        }
      }
    }
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, STAGES <= 4 ? 2 : 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));   // allocation granularity: powers of two
  // instruction descriptor: D=f32 (1<<4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accbar = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~(uintptr_t)15);   // [BN]

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const bool dbg = g.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  const unsigned cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (g.dbg && threadIdx.x == 0 && cta_lin < 2048) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g.dbg[8 + 2 * cta_lin] = t; }
  if (dbg && threadIdx.x == 0) { g.dbg[0] = clock64(); }
  const int kb_begin = blockIdx.z * g.kb_per_split;
  const int num_k = min((g.K + BK - 1) / BK - kb_begin, g.kb_per_split);   // K-blocks of this split
  const bool split = gridDim.z > 1;

  const bool tail_cta = g.tail_rows > 0 && m_blk == (int)gridDim.y - 1;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    // a stage is released by the MMA commit and, in tail CTAs, by the eight epilogue warps that also read its W tile
    for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, tail_cta ? 9 : 1); }
    mbar_init(accbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (dbg && threadIdx.x == 0) g.dbg[1] = clock64();

  if (warp == 0) {
    if (lane == 0) {
      // Weights do not depend on the predecessor kernel: start streaming the first STAGES W tiles BEFORE the
      // programmatic-dependency wait (hides the HBM/L2 latency of the first loads behind the predecessor's tail);
      // the matching A tiles (activations) are issued right after the wait and complete the same barriers.
      const int pre = num_k < STAGES ? num_k : STAGES;
      const uint32_t a_tx = (uint32_t)g.a_box_rows * BK * 2;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(full0 + 8 * kb, a_tx + B_BYTES);
        tma_load_2d(smem_u32(sB + kb * B_BYTES), &tmB, (kb_begin + kb) * BK, n_blk * BN, full0 + 8 * kb);
      }
      pdl_wait();
      if (dbg) g.dbg[2] = clock64();
      for (int kb = 0; kb < pre; ++kb)
        tma_load_2d(smem_u32(sA + kb * A_BYTES), &tmA, (kb_begin + kb) * BK, m_blk * BM, full0 + 8 * kb);
      int s = pre == STAGES ? 0 : pre; uint32_t ph = pre == STAGES ? 1 : 0;
      for (int kb = pre; kb < num_k; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, a_tx + B_BYTES);
        tma_load_2d(smem_u32(sA + s * A_BYTES), &tmA, (kb_begin + kb) * BK, m_blk * BM, full0 + 8 * s);
        tma_load_2d(smem_u32(sB + s * B_BYTES), &tmB, (kb_begin + kb) * BK, n_blk * BN, full0 + 8 * s);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        if (dbg && kb == 0) g.dbg[3] = clock64();
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          tc_mma_f16(tmem_base, umma_desc_k128(a0 + k * 32), umma_desc_k128(b0 + k * 32), IDESC,
                     (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty0 + 8 * s);   // smem stage reusable once these MMAs retire
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      tc_commit(accbar);             // accumulator complete
      if (dbg) g.dbg[4] = clock64();
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
    const int q = warp & 3;
    const int row = m_blk * BM + q * 32 + lane;
    // bias is a weight (no dependency on the predecessor): stage this tile's slice while the main loop runs
    const int et = threadIdx.x - 64;
    if (et < BN) { const int n = n_blk * BN + et; sbias[et] = (g.bias && n < g.N) ? __ldg(g.bias + n) : 0.f; }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
    pdl_wait();          // residual stream / output buffers of the predecessor are visible
    if (tail_cta) {
      // ---- tail rows on CUDA cores: thread = (feature f of this tile, K half hk); W from the 128B-swizzled stage
      const int f = et >> 1, hk = et & 1;
      float tacc[GEMM_TAIL_MAX];
#pragma unroll
      for (int r = 0; r < GEMM_TAIL_MAX; ++r) tacc[r] = 0.f;
      // stage this CTA's K slice of the tail rows once (global latency must not sit between "stage full" and "stage released")
      __half* sx = reinterpret_cast<__half*>(sbias + BN);
      const int kslice = num_k * BK, kbase = kb_begin * BK;
      for (int i = et; i < g.tail_rows * (kslice >> 3); i += 256) {
        const int r = i / (kslice >> 3), c8 = (i - r * (kslice >> 3)) << 3;
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (kbase + c8 < g.K) xv = __ldg(reinterpret_cast<const uint4*>(g.a_tail + (size_t)r * g.lda + kbase + c8));
        *reinterpret_cast<uint4*>(sx + (size_t)r * kslice + c8) = xv;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        if (f < BN) {
          const uint8_t* wrow = sB + (size_t)s * B_BYTES + (size_t)f * 128;
          const int k0 = kb * BK + hk * 32;                    // offset inside the staged slice
          uint4 wv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const uint4*>(wrow + (((hk * 4 + c) ^ (f & 7)) << 4));
#pragma unroll
          for (int r = 0; r < GEMM_TAIL_MAX; ++r) {
            if (r < g.tail_rows) {
              const __half* xr = sx + (size_t)r * kslice + k0;
              float acc = 0.f;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                {                                              // chunks past K were staged as zeros (and W is zero-filled by TMA)
                  const uint4 xv = *reinterpret_cast<const uint4*>(xr + c * 8);
                  const __half2* xh = reinterpret_cast<const __half2*>(&xv);
                  const __half2* wh = reinterpret_cast<const __half2*>(&wv[c]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 xf = __half22float2(xh[e]), wf = __half22float2(wh[e]);
                    acc = fmaf(xf.x, wf.x, acc); acc = fmaf(xf.y, wf.y, acc);
                  }
                }
              }
              tacc[r] += acc;
            }
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
#pragma unroll
      for (int r = 0; r < GEMM_TAIL_MAX; ++r) tacc[r] += __shfl_xor_sync(0xffffffffu, tacc[r], 1);
      const int n = n_blk * BN + f;
      if (hk == 0 && f < BN && n < g.N) {
#pragma unroll
        for (int r = 0; r < GEMM_TAIL_MAX; ++r) {
          if (r < g.tail_rows) {
            float v = tacc[r] + (blockIdx.z == 0 ? sbias[f] : 0.f);
            const size_t o = (size_t)(g.tail_row0 + r) * g.ldo + n;
            if (g.epi == VLFM_EPI_BIAS_F16 || g.epi == VLFM_EPI_BIAS_GELU_F16 || g.epi == VLFM_EPI_BIAS_RELU_F16) {
              if (g.epi == VLFM_EPI_BIAS_GELU_F16) v = gelu_erf(v);
              else if (g.epi == VLFM_EPI_BIAS_RELU_F16) v = fmaxf(v, 0.f);
              reinterpret_cast<__half*>(g.out)[o] = __float2half_rn(v);
            } else {
              const bool partial = (g.epi == VLFM_EPI_PARTIAL_F32);
              float* po = reinterpret_cast<float*>(g.out) + (partial ? (size_t)blockIdx.z * (size_t)g.split_stride : 0) + o;
              if (split && !partial) atomicAdd(po, v);
              else *po = (g.epi == VLFM_EPI_BIAS_RESID_F32) ? *po + v : v;
            }
          }
        }
      }
    }
    mbar_wait(accbar, 0);
    if (dbg && threadIdx.x == 64) g.dbg[5] = clock64();
    tc_fence_after();
    epilogue_slab<BN>(tmem_base, q, row, n_blk, g, split, (warp - 2) >> 2, sbias);
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) g.dbg[6] = clock64();
  if (g.dbg && threadIdx.x == 0 && cta_lin < 2048) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g.dbg[9 + 2 * cta_lin] = t; }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}


// ================================================================================================
// "x2" GEMM: fp32-grade product on the fp16 tensor path (the Q-Former, which the reference runs in float32).
//   A = A_hi + A_lo / 2048,  W = W_hi + W_lo / 2048   (fp16 pairs, see split_x2 in common.cuh)
//   out = A_hi.W_hi  +  (A_lo.W_hi + A_hi.W_lo) / 2048          (the lo.lo term is ~2^-22 relative: dropped)
// Same structure as gemm_f16_tcgen05_kernel; a stage holds FOUR tiles, the issuer thread emits three tcgen05.mma per K step into
// TWO TMEM accumulators (main, correction) and the epilogue merges them.  At 32 query rows the tensor pipe is idle anyway: the
// cost is the second weight tile per stage (fp32-sized weight traffic).
// ================================================================================================
template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16x2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
                          const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBl, GemmArgs g) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : 256);
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sAl = sA + STAGES * A_BYTES;
  uint8_t* sB = sAl + STAGES * A_BYTES;
  uint8_t* sBl = sB + STAGES * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBl + STAGES * B_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accbar = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~(uintptr_t)15);   // [BN]

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const int kb_begin = blockIdx.z * g.kb_per_split;
  const int num_k = min((g.K + BK - 1) / BK - kb_begin, g.kb_per_split);
  const bool split = gridDim.z > 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(accbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // weights first (independent of the predecessor kernel), activations after the dependency wait
      const int pre = num_k < STAGES ? num_k : STAGES;
      const uint32_t a_tx = (uint32_t)g.a_box_rows * BK * 2;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(full0 + 8 * kb, 2 * (a_tx + B_BYTES));
        tma_load_2d(smem_u32(sB + kb * B_BYTES), &tmB, (kb_begin + kb) * BK, n_blk * BN, full0 + 8 * kb);
        tma_load_2d(smem_u32(sBl + kb * B_BYTES), &tmBl, (kb_begin + kb) * BK, n_blk * BN, full0 + 8 * kb);
      }
      pdl_wait();
      for (int kb = 0; kb < pre; ++kb) {
        tma_load_2d(smem_u32(sA + kb * A_BYTES), &tmA, (kb_begin + kb) * BK, m_blk * BM, full0 + 8 * kb);
        tma_load_2d(smem_u32(sAl + kb * A_BYTES), &tmAl, (kb_begin + kb) * BK, m_blk * BM, full0 + 8 * kb);
      }
      int s = pre == STAGES ? 0 : pre; uint32_t ph = pre == STAGES ? 1 : 0;
      for (int kb = pre; kb < num_k; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, 2 * (a_tx + B_BYTES));
        tma_load_2d(smem_u32(sA + s * A_BYTES), &tmA, (kb_begin + kb) * BK, m_blk * BM, full0 + 8 * s);
        tma_load_2d(smem_u32(sAl + s * A_BYTES), &tmAl, (kb_begin + kb) * BK, m_blk * BM, full0 + 8 * s);
        tma_load_2d(smem_u32(sB + s * B_BYTES), &tmB, (kb_begin + kb) * BK, n_blk * BN, full0 + 8 * s);
        tma_load_2d(smem_u32(sBl + s * B_BYTES), &tmBl, (kb_begin + kb) * BK, n_blk * BN, full0 + 8 * s);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t acc0 = tmem_base, acc1 = tmem_base + (uint32_t)BN;
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * A_BYTES), al0 = smem_u32(sAl + s * A_BYTES);
        const uint32_t b0 = smem_u32(sB + s * B_BYTES), bl0 = smem_u32(sBl + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = umma_desc_k128(a0 + k * 32), dal = umma_desc_k128(al0 + k * 32);
          const uint64_t db = umma_desc_k128(b0 + k * 32), dbl = umma_desc_k128(bl0 + k * 32);
          const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
          tc_mma_f16(acc0, da, db, IDESC, first);
          tc_mma_f16(acc1, dal, db, IDESC, first);
          tc_mma_f16(acc1, da, dbl, IDESC, 1u);
        }
        tc_commit(empty0 + 8 * s);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      tc_commit(accbar);
    }
  } else {
    const int q = warp & 3;
    const int row = m_blk * BM + q * 32 + lane;
    const int et = threadIdx.x - 64;
    if (et < BN) { const int n = n_blk * BN + et; sbias[et] = (g.bias && n < g.N) ? __ldg(g.bias + n) : 0.f; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    pdl_wait();
    mbar_wait(accbar, 0);
    tc_fence_after();
    epilogue_slab<BN, true>(tmem_base, q, row, n_blk, g, split, (warp - 2) >> 2, sbias, tmem_base + (uint32_t)BN);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}


// ================================================================================================
// 2-CTA variant (cta_group::2): a thread-block cluster of two CTAs on one TPC computes a 256 x BN
// output tile with UMMA M=256.  Each CTA stages its own 128 rows of A and HALF of the B tile
// (BN/2 rows); the tensor core reads both halves, so every SM ingests (128 + BN/2) x 128 B per
// K-block for 128 x BN outputs -- half the operand bytes per FLOP of the 1-CTA kernel, which is
// SM-ingest bound (~60 B/clk/SM measured).  Leader CTA (cluster rank 0) issues the MMAs; both
// CTAs' TMA loads complete on the leader's full barrier; tcgen05.commit multicasts stage-release
// and accumulator-ready to both CTAs; each CTA drains its own 128 TMEM lanes.
// ================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> even CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t leader_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
  constexpr uint32_t A_BYTES = BM * BK * 2, BH_BYTES = (BN / 2) * BK * 2;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);   // M = 256

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * BH_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accbar = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;     // this CTA's own 128-row slab; the pair is (2p, 2p+1)
  const int kb_begin = blockIdx.z * g.kb_per_split;
  const int num_k = min((g.K + BK - 1) / BK - kb_begin, g.kb_per_split);
  const bool split = gridDim.z > 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 2); mbar_init(empty0 + 8 * i, 1); }   // full: leader expect_tx + peer arrive
    mbar_init(accbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();        // both CTAs' barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t lead_full = (full0 + 8 * s) & PEER_MASK;
        if (leader) mbar_expect_tx(full0 + 8 * s, 2 * (A_BYTES + BH_BYTES));
        else mbar_arrive_leader(lead_full);
        tma_load_2d_2sm(smem_u32(sA + s * A_BYTES), &tmA, (kb_begin + kb) * BK, m_blk * BM, lead_full);
        tma_load_2d_2sm(smem_u32(sB + s * BH_BYTES), &tmB, (kb_begin + kb) * BK, n_blk * BN + (int)rank * (BN / 2), lead_full);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * BH_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          tc_mma_f16_2sm(tmem_base, umma_desc_k128(a0 + k * 32), umma_desc_k128(b0 + k * 32), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit_2sm(empty0 + 8 * s);     // frees stage s in BOTH CTAs
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      tc_commit_2sm(accbar);               // accumulator ready in BOTH CTAs
    }
  } else {
    const int q = warp & 3;
    const int row = m_blk * BM + q * 32 + lane;
    mbar_wait(accbar, 0);
    tc_fence_after();
    epilogue_slab<BN>(tmem_base, q, row, n_blk, g, split, (warp - 2) >> 2);
  }
  tc_fence_before();
  cluster_sync_all();        // the peer may still be reading our smem / signalling our barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ================================================================================================
// Persistent 2-CTA variant: one cluster (CTA pair) per TPC loops over 256 x 256 output tiles
// (static round-robin, M fastest so concurrently processed tiles share weight tiles in L2).
// TMEM holds TWO 256-column accumulators (all 512 columns): the epilogue of tile i (tcgen05.ld,
// bias/GELU/residual, stores) overlaps the TMA + MMA main loop of tile i+1.  Pipelines:
//   smem ring   full[s]  (leader, 2 arrivals + tx)   / empty[s]  (per CTA, multicast commit)
//   accumulators tfull[a] (per CTA, multicast commit) / tempty[a] (leader, 16 epilogue-warp arrivals)
// ================================================================================================
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_tcgen05_2cta_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
  static_assert(2 * BN <= 512, "two accumulators must fit TMEM");
  constexpr uint32_t A_BYTES = BM * BK * 2, BH_BYTES = (BN / 2) * BK * 2;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * BH_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_k = (g.K + BK - 1) / BK;
  const int pairsM = (g.M + 2 * BM - 1) / (2 * BM), tilesN = (g.N + BN - 1) / BN;
  const int total = pairsM * tilesN;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 2); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = cluster_id; tile < total; tile += num_clusters) {
        const int m_blk = 2 * (tile % pairsM) + (int)rank, n_blk = tile / pairsM;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t lead_full = (full0 + 8 * s) & PEER_MASK;
          if (leader) mbar_expect_tx(full0 + 8 * s, 2 * (A_BYTES + BH_BYTES));
          else mbar_arrive_leader(lead_full);
          tma_load_2d_2sm(smem_u32(sA + s * A_BYTES), &tmA, kb * BK, m_blk * BM, lead_full);
          tma_load_2d_2sm(smem_u32(sB + s * BH_BYTES), &tmB, kb * BK, n_blk * BN + (int)rank * (BN / 2), lead_full);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total; tile += num_clusters, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty0 + 8 * acc, ((it >> 1) & 1) ^ 1);     // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * BH_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_f16_2sm(d_tmem, umma_desc_k128(a0 + k * 32), umma_desc_k128(b0 + k * 32), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit_2sm(empty0 + 8 * s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit_2sm(tfull0 + 8 * acc);
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int tile = cluster_id; tile < total; tile += num_clusters, ++it) {
      const int acc = it & 1;
      const int m_blk = 2 * (tile % pairsM) + (int)rank, n_blk = tile / pairsM;
      const int row = m_blk * BM + q * 32 + lane;
      mbar_wait(tfull0 + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      epilogue_slab<BN>(tmem_base + (uint32_t)(acc * BN), q, row, n_blk, g, false, (warp - 2) >> 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader((tempty0 + 8 * acc) & PEER_MASK);   // this warp no longer reads accumulator `acc`
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---------------------------------------------------------------- host side ------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D fp16 row-major [rows, cols] (cols contiguous, leading dimension ld elements), box {64, boxRows}
static int make_map(CUtensorMap* m, const void* base, int rows, int cols, int ld, int boxRows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VLFM_E_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld); return VLFM_E_DRIVER; }
  return VLFM_OK;
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& ta, const void* W, int ldw, const GemmArgs& g, cudaStream_t st) {
  CUtensorMap tb;
  int rc = make_map(&tb, W, g.N, g.K, ldw, BN);
  if (rc) return rc;
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 2) * 8 + 1024 + BN * 4 + 32 +
                          (size_t)GEMM_TAIL_MAX * GEMM_TAIL_KMAX * 2;
  static bool configured = false;
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_f16_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(gemm)");
    if (rc) return rc;
    configured = true;
  }
  const int num_k = (g.K + BK - 1) / BK;
  dim3 grid((g.N + BN - 1) / BN, g.tail_rows > 0 ? g.M / BM : (g.M + BM - 1) / BM, (num_k + g.kb_per_split - 1) / g.kb_per_split);
  rc = check_cuda(launch_pdl(gemm_f16_tcgen05_kernel<BN, STAGES>, grid, dim3(GEMM_THREADS), smem, st, ta, tb, g), "gemm_f16_tcgen05_kernel");
  if (rc) return rc;
  count_launch();
  return VLFM_OK;
}


template <int BN, int STAGES>
static int launch_gemm_2cta(const CUtensorMap& ta, const void* W, int ldw, const GemmArgs& g, cudaStream_t st) {
  CUtensorMap tb;
  int rc = make_map(&tb, W, g.N, g.K, ldw, BN / 2);
  if (rc) return rc;
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + (BN / 2) * BK * 2) + (2 * STAGES + 2) * 8 + 1024;
  static bool configured = false;
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_f16_tcgen05_2cta_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(gemm 2cta)");
    if (rc) return rc;
    configured = true;
  }
  const int num_k = (g.K + BK - 1) / BK;
  const int pairs = (g.M + 2 * BM - 1) / (2 * BM);
  dim3 grid(2 * pairs, (g.N + BN - 1) / BN, (num_k + g.kb_per_split - 1) / g.kb_per_split);
  rc = check_cuda(launch_pdl(gemm_f16_tcgen05_2cta_kernel<BN, STAGES>, grid, dim3(GEMM_THREADS), smem, st, ta, tb, g), "gemm_f16_tcgen05_2cta_kernel");
  if (rc) return rc;
  count_launch();
  return VLFM_OK;
}

template <int BN, int STAGES>
static int launch_gemm_2cta_persistent(const CUtensorMap& ta, const void* W, int ldw, const GemmArgs& g, cudaStream_t st) {
  CUtensorMap tb;
  int rc = make_map(&tb, W, g.N, g.K, ldw, BN / 2);
  if (rc) return rc;
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + (BN / 2) * BK * 2) + (2 * STAGES + 6) * 8 + 1024;
  static bool configured = false;
  static int clusters_max = 74;
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_f16_tcgen05_2cta_persistent_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(gemm 2cta persistent)");
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    clusters_max = sms / 2;
    configured = true;
  }
  const int pairs = (g.M + 2 * BM - 1) / (2 * BM), tiles = pairs * ((g.N + BN - 1) / BN);
  const int clusters = tiles < clusters_max ? tiles : clusters_max;
  rc = check_cuda(launch_pdl(gemm_f16_tcgen05_2cta_persistent_kernel<BN, STAGES>, dim3(2 * clusters), dim3(GEMM_THREADS), smem, st, ta, tb, g),
                  "gemm_f16_tcgen05_2cta_persistent_kernel");
  if (rc) return rc;
  count_launch();
  return VLFM_OK;
}

template <int BN, int STAGES>
static int launch_gemm_x2(const CUtensorMap& ta, const CUtensorMap& tal, const void* W, const void* Wl, int ldw, const GemmArgs& g, cudaStream_t st) {
  CUtensorMap tb, tbl;
  int rc = make_map(&tb, W, g.N, g.K, ldw, BN);
  if (!rc) rc = make_map(&tbl, Wl, g.N, g.K, ldw, BN);
  if (rc) return rc;
  constexpr size_t smem = (size_t)STAGES * 2 * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 2) * 8 + 1024 + BN * 4 + 32;
  static_assert(smem <= 227 * 1024, "x2 GEMM stage ring exceeds shared memory");
  static bool configured = false;
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_f16x2_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(gemm x2)");
    if (rc) return rc;
    configured = true;
  }
  const int num_k = (g.K + BK - 1) / BK;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, (num_k + g.kb_per_split - 1) / g.kb_per_split);
  rc = check_cuda(launch_pdl(gemm_f16x2_tcgen05_kernel<BN, STAGES>, grid, dim3(GEMM_THREADS), smem, st, ta, tal, tb, tbl, g), "gemm_f16x2_tcgen05_kernel");
  if (rc) return rc;
  count_launch();
  return VLFM_OK;
}

}  // namespace vlfm

using namespace vlfm;

static int gemm_dispatch(const void* d_A, const void* d_W, int M, int N, int K, int lda, int ldw, GemmArgs g, void* stream, float* d_partials,
                         size_t partial_bytes, int* splits_out);

#ifdef VLFM_DEV_PROBES   // development builds only (scripts/gemm_timeline.py): not part of the shipped C-ABI
extern "C" void vlfm_gemm_debug_timeline(unsigned long long* d_buf8) { vlfm::g_gemm_dbg = d_buf8; }
#endif

extern "C" int vlfm_gemm_f16(const void* d_A, const void* d_W, const float* d_bias, void* d_out, int M, int N,
                             int K, int lda, int ldw, int ldo, int epilogue, void* stream) {
  if (!d_A || !d_W || !d_out || M < 1 || N < 1 || K < 1) { set_error("vlfm_gemm_f16: bad argument"); return VLFM_E_INVALID; }
  if ((K & 7) || (lda & 7) || (ldw & 7) || (ldo & 7) || ((uintptr_t)d_A & 15) || ((uintptr_t)d_W & 15) || ((uintptr_t)d_out & 15)) {
    set_error("vlfm_gemm_f16: K, lda, ldw, ldo must be multiples of 8 and pointers 16-byte aligned"); return VLFM_E_INVALID; }
  if (epilogue < 0 || epilogue > 4) { set_error("vlfm_gemm_f16: unknown epilogue %d", epilogue); return VLFM_E_INVALID; }
  GemmArgs g{d_bias, d_out, M, N, K, ldo, epilogue, (K + BK - 1) / BK, g_gemm_dbg, 0, nullptr, 0, 0, 0, nullptr, BM};
  return gemm_dispatch(d_A, d_W, M, N, K, lda, ldw, g, stream, nullptr, 0, nullptr);
}

static int gemm_dispatch(const void* d_A, const void* d_W, int M, int N, int K, int lda, int ldw, GemmArgs g, void* stream, float* d_partials,
                         size_t partial_bytes, int* splits_out) {
  if (splits_out) *splits_out = 1;
  const int epilogue = g.epi;
  CUtensorMap ta;
  g.a_box_rows = M <= 32 ? 32 : (M <= 64 ? 64 : BM);
  int rc = make_map(&ta, d_A, M, K, lda, g.a_box_rows);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // tile width: fill >= ~120 of the 148 SMs when M is small, widest tile otherwise
  // Tile / split-K plan from a small cost model (us): waves x (fixed + bytes a CTA must pull / its share of
  // the L2->SM bandwidth).  Never spill into a second wave for a handful of CTAs; split K only for the
  // fp32 residual epilogue (partials are reduced with red.add into the residual stream).
  const int mt = (M + BM - 1) / BM, num_k = (K + BK - 1) / BK;
  {
    // Large problems (>= ~2 waves of 128x128 tiles): CTA pairs, 256 x 256 tiles, half the operand bytes per FLOP.
    static int two = -1;
    if (two < 0) { const char* e = getenv("VLFM_GEMM_2CTA"); two = (e && e[0] == '0') ? 0 : 1; }
    const long tiles128 = (long)mt * ((N + 127) / 128);
    if (two && tiles128 >= 296 && N >= 256) {
      g.kb_per_split = num_k;
      static int persist = -1;
      if (persist < 0) { const char* e = getenv("VLFM_GEMM_PERSIST"); persist = (e && e[0] == '0') ? 0 : 1; }
      if (persist) return launch_gemm_2cta_persistent<256, 6>(ta, d_W, ldw, g, st);
      return launch_gemm_2cta<256, 6>(ta, d_W, ldw, g, st);
    }
  }
  // M = 128*q + r with a tiny remainder (ViT: 257 tokens = 2*128 + 1): launch q row tiles only; the last tile's CTAs compute
  // the r tail rows on CUDA cores from the W tiles they stage anyway.  A third of the CTAs (and of the W traffic) disappears,
  // which buys narrower tiles / deeper split-K inside one wave.
  static int tail_on = -1;
  if (tail_on < 0) { const char* e = getenv("VLFM_GEMM_TAIL"); tail_on = (e && e[0] == '0') ? 0 : 1; }
  const int rem = M % BM;
  const bool tail = tail_on && M > BM && rem >= 1 && rem <= GEMM_TAIL_MAX && num_k * BK <= GEMM_TAIL_KMAX;
  int best_bn = 128, best_s = 1, force_shallow = 0;
  double best_t = 1e30;
  bool best_tail = false;
  // Cost model fitted to in-graph measurements on B200 (scripts/gemm_m256_probe.py, profiles/r01_gemm_tail_probe.txt):
  //   t = waves * (c0 + K-blocks per CTA * per_kb),  per_kb = max(0.33 us pipeline floor, CTAs * KB per K-block / 12.6 TB/s chip L2->SM),
  //   c0 = prologue + epilogue: 2.6 / 3.0 / 3.8 us for 32 / 64 / 128-wide tiles, + split-K atomics, + the tail-row work.
  const int bns[4] = {128, 96, 64, 32};
  for (int tm = 0; tm < (tail ? 2 : 1); ++tm) {
    const bool use_tail = tail && tm == 0;
    const int mte = use_tail ? M / BM : mt;
    for (int bi = 0; bi < 4; ++bi) {
      const int bn = bns[bi], nt = (N + bn - 1) / bn;
      static int bn96 = -1;
      if (bn96 < 0) { const char* e = getenv("VLFM_GEMM_BN96"); bn96 = (e && e[0] == '1') ? 1 : 0; }
      // 96-wide tiles for the two-row-tile (tail) plans: FC1 at 257 tokens is 0.6 us faster in isolation (L2-resident weights) but
      // 1 % slower inside the forward (weights from HBM): opt-in only
      if (bn == 96 && (!use_tail || !bn96)) continue;
      int smax = (epilogue == VLFM_EPI_BIAS_RESID_F32) ? num_k / 4 : 1;
      static int smax_cap = -1;
      if (smax_cap < 0) { const char* e = getenv("VLFM_GEMM_SMAX"); smax_cap = e ? atoi(e) : 8; if (smax_cap < 1 || smax_cap > 8) smax_cap = 8; }
      if (smax < 1) smax = 1;
      if (smax > smax_cap) smax = smax_cap;
      for (int sp = 1; sp <= smax; ++sp) {
        const double ctas = (double)mte * nt * sp;
        const int kb = (num_k + sp - 1) / sp;
        const double waves = (double)(long)((ctas + 147) / 148);
        const double active = ctas < 148 ? ctas : 148;
        const double kb_kbytes = (double)(128 + bn) * 128 / 1024.0;
        double per_kb = active * kb_kbytes / 12600.0;
        if (per_kb < 0.33) per_kb = 0.33;
        double c0 = bn == 128 ? 3.8 + 0.22 * sp : (bn == 96 ? 3.4 + 0.16 * sp : (bn == 64 ? 3.0 : 2.6) + (sp > 1 ? 0.1 * sp : 0.0));
        if (use_tail) c0 += bn == 128 ? 0.9 : (bn == 96 ? 0.6 : 0.3);
        const double t = waves * (c0 + kb * per_kb);
        if (t < best_t) { best_t = t; best_bn = bn; best_s = sp; best_tail = use_tail; }
      }
    }
  }
  if (const char* f = getenv("VLFM_GEMM_FORCE")) {   // development sweep: "bn:splits"
    int fb = 0, fs = 0, fsh = 0;
    if (sscanf(f, "%d:%d:%d", &fb, &fs, &fsh) >= 2 && (fb == 128 || fb == 64 || fb == 32 || (fb == 96 && tail)) && fs >= 1) {
      if (fb == 96) best_tail = true;
      best_bn = fb;
      best_s = (epilogue == VLFM_EPI_BIAS_RESID_F32) ? (fs > num_k ? num_k : fs) : 1;
      force_shallow = fsh;
    }
  }
  g.kb_per_split = (num_k + best_s - 1) / best_s;
  // Small grids (<= one wave) use the shallow-pipeline variants (<= 100 KB smem, 2 CTAs/SM) so that, with
  // programmatic dependent launch, the next GEMM's CTAs are already resident when this one drains.
  static int shallow = -1;
  if (shallow < 0) { const char* e = getenv("VLFM_GEMM_SHALLOW"); shallow = (e && e[0] == '1') ? 1 : 0; }  // measured slower on B200 (3 stages cannot cover the latency): off
  const bool one_wave = (long)(best_tail ? M / BM : mt) * ((N + best_bn - 1) / best_bn) * best_s <= 148;
  if (best_tail) { g.a_tail = (const __half*)d_A + (size_t)(M / BM) * BM * lda; g.lda = lda; g.tail_rows = rem; g.tail_row0 = (M / BM) * BM; }
  (void)one_wave;
  // deterministic split-K: the splits store their partial sums side by side and the caller reduces them in a fixed order
  const int launched_splits = (num_k + g.kb_per_split - 1) / g.kb_per_split;      // grid.z (can be below best_s when num_k is small)
  if (d_partials && launched_splits > 1 && epilogue == VLFM_EPI_BIAS_RESID_F32 && (size_t)launched_splits * (size_t)M * (size_t)N * 4 <= partial_bytes) {
    g.epi = VLFM_EPI_PARTIAL_F32; g.out = d_partials; g.ldo = N; g.split_stride = (long long)M * N;
    if (splits_out) *splits_out = launched_splits;
  }
  if (((shallow && one_wave) || force_shallow) && best_bn != 96) {
    if (best_bn == 128) return launch_gemm<128, 3>(ta, d_W, ldw, g, st);
    if (best_bn == 64) return launch_gemm<64, 4>(ta, d_W, ldw, g, st);
    return launch_gemm<32, 4>(ta, d_W, ldw, g, st);
  }
  if (best_bn == 128) return launch_gemm<128, 6>(ta, d_W, ldw, g, st);
  if (best_bn == 96) return launch_gemm<96, 7>(ta, d_W, ldw, g, st);
  if (best_bn == 64) return launch_gemm<64, 8>(ta, d_W, ldw, g, st);
  return launch_gemm<32, 8>(ta, d_W, ldw, g, st);
}

extern "C" int vlfm_layernorm(const float* d_x, const float* d_gamma, const float* d_beta, void* d_out16, float* d_out32,
                   int rows, int D, int ldx, int ldo16, int ldo32, float eps, void* stream);
extern "C" int vlfm_layernorm_reduce(float* d_x, const float* d_partials, int splits, long long split_stride, const float* d_gamma,
                                     const float* d_beta, void* d_out16, float* d_out32, int rows, int D, int ldx, int ldo16, int ldo32,
                                     float eps, void* stream);

// x += A @ W^T + bias ; out = LayerNorm(x) -- bitwise reproducible: when the plan splits K, the splits store their partial sums
// in d_partials (no atomics) and the LayerNorm kernel adds them to x in split order before normalising; an unsplit GEMM adds
// into x directly (one writer per element).  (Round 1 reduced the splits with red.global.add: the order of arrival varied from run
// to run and the 39-layer residual stream amplified the last-bit differences to ~6e-5 on the cosine.)
extern "C" int vlfm_gemm_f16_resid_ln(const void* d_A, const void* d_W, const float* d_bias, float* d_x, int M, int N, int K,
                                      int lda, int ldw, int ldx, const float* d_gamma, const float* d_beta, void* d_out16,
                                      int ld16, float* d_out32, int ld32, float eps, float* d_partials, size_t partial_bytes, void* stream) {
  if (!d_A || !d_W || !d_x || !d_gamma || !d_beta || (!d_out16 && !d_out32) || M < 1 || N < 1 || K < 1) {
    set_error("vlfm_gemm_f16_resid_ln: bad argument"); return VLFM_E_INVALID; }
  if ((K & 7) || (lda & 7) || (ldw & 7) || (ldx & 7) || (N & 3) || (ld16 & 3) || (ld32 & 3) || ((uintptr_t)d_A & 15) || ((uintptr_t)d_W & 15) ||
      ((uintptr_t)d_x & 15) || ((uintptr_t)d_partials & 15)) { set_error("vlfm_gemm_f16_resid_ln: alignment (K, strides %% 8; N %% 4; 16-byte pointers)"); return VLFM_E_INVALID; }
  GemmArgs g{d_bias, d_x, M, N, K, ldx, VLFM_EPI_BIAS_RESID_F32, (K + BK - 1) / BK, g_gemm_dbg, 0, nullptr, 0, 0, 0, nullptr, BM};
  int splits = 1;
  int rc = gemm_dispatch(d_A, d_W, M, N, K, lda, ldw, g, stream, d_partials, partial_bytes, &splits);
  if (rc) return rc;
  if (splits > 1) return vlfm_layernorm_reduce(d_x, d_partials, splits, (long long)M * N, d_gamma, d_beta, d_out16, d_out32, M, N, ldx, ld16, ld32, eps, stream);
  return vlfm_layernorm(d_x, d_gamma, d_beta, d_out16, d_out32, M, N, ldx, ld16, ld32, eps, stream);
}

// ---- x2 GEMMs (fp32-grade, see gemm_f16x2_tcgen05_kernel) ----
static int gemm_x2_dispatch(const void* d_A_hi, const void* d_A_lo, const void* d_W_hi, const void* d_W_lo, int M, int N, int K, int lda, int ldw,
                            GemmArgs g, void* stream, float* d_partials, size_t partial_bytes, int* splits_out) {
  if (splits_out) *splits_out = 1;
  // M = 128 q + r with a small remainder (257 image tokens): the r rows would cost a whole extra row of tiles -- run them as their
  // own skinny launch (32-row A box) after the q full row tiles
  const int rem = M % BM;
  if (M > BM && rem >= 1 && rem <= 32 && (g.epi == VLFM_EPI_BIAS_F32 || g.epi == VLFM_EPI_BIAS_GELU_F16X2)) {
    const int m0 = M - rem;
    GemmArgs g0 = g; g0.M = m0;
    int rc0 = gemm_x2_dispatch(d_A_hi, d_A_lo, d_W_hi, d_W_lo, m0, N, K, lda, ldw, g0, stream, nullptr, 0, nullptr);
    if (rc0) return rc0;
    GemmArgs g1 = g; g1.M = rem;
    const size_t esz = g.epi == VLFM_EPI_BIAS_F32 ? 4 : 2;
    g1.out = (uint8_t*)g.out + (size_t)m0 * g.ldo * esz;
    if (g.out_lo) g1.out_lo = (uint8_t*)g.out_lo + (size_t)m0 * g.ldo * 2;
    return gemm_x2_dispatch((const __half*)d_A_hi + (size_t)m0 * lda, (const __half*)d_A_lo + (size_t)m0 * lda, d_W_hi, d_W_lo, rem, N, K, lda, ldw, g1,
                            stream, nullptr, 0, nullptr);
  }
  CUtensorMap ta, tal;
  g.a_box_rows = M <= 32 ? 32 : (M <= 64 ? 64 : BM);
  int rc = make_map(&ta, d_A_hi, M, K, lda, g.a_box_rows);
  if (!rc) rc = make_map(&tal, d_A_lo, M, K, lda, g.a_box_rows);
  if (rc) return rc;
  const int mt = (M + BM - 1) / BM, num_k = (K + BK - 1) / BK;
  // tile width: wide tiles for the one big problem (cross-attention K/V of all layers: 257 x 9216 x 1408), else enough CTAs to
  // cover the machine; split K (deterministic partial sums) only for the residual epilogue -- at 32 rows a partial slab is 100 KB
  int bn = 32;
  if (mt >= 2 && N >= 4096) bn = 128;
  else if ((long)mt * ((N + 63) / 64) >= 96) bn = 64;
  const int tiles = mt * ((N + bn - 1) / bn);
  int sp = 1;
  if (g.epi == VLFM_EPI_BIAS_RESID_F32 && d_partials) {
    sp = 148 / tiles; if (sp > num_k / 3) sp = num_k / 3; if (sp > 8) sp = 8; if (sp < 1) sp = 1;
  }
  g.kb_per_split = (num_k + sp - 1) / sp;
  const int launched = (num_k + g.kb_per_split - 1) / g.kb_per_split;
  if (launched > 1) {
    if ((size_t)launched * (size_t)M * (size_t)N * 4 > partial_bytes) { g.kb_per_split = num_k; }
    else { g.epi = VLFM_EPI_PARTIAL_F32; g.out = d_partials; g.ldo = N; g.split_stride = (long long)M * N; if (splits_out) *splits_out = launched; }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 128) return launch_gemm_x2<128, 3>(ta, tal, d_W_hi, d_W_lo, ldw, g, st);
  if (bn == 64) return launch_gemm_x2<64, 4>(ta, tal, d_W_hi, d_W_lo, ldw, g, st);
  return launch_gemm_x2<32, 5>(ta, tal, d_W_hi, d_W_lo, ldw, g, st);
}

static int x2_args_ok(const char* who, const void* a, const void* al, const void* w, const void* wl, const void* out, int M, int N, int K, int lda, int ldw, int ldo) {
  if (!a || !al || !w || !wl || !out || M < 1 || N < 1 || K < 1) { set_error("%s: bad argument", who); return VLFM_E_INVALID; }
  if ((K & 7) || (lda & 7) || (ldw & 7) || (ldo & 7) || ((uintptr_t)a & 15) || ((uintptr_t)al & 15) || ((uintptr_t)w & 15) || ((uintptr_t)wl & 15) || ((uintptr_t)out & 15)) {
    set_error("%s: K, lda, ldw, ldo must be multiples of 8 and pointers 16-byte aligned", who); return VLFM_E_INVALID; }
  return VLFM_OK;
}

// out = epilogue((A_hi + A_lo/2048) @ (W_hi + W_lo/2048)^T + bias).  epilogue: VLFM_EPI_BIAS_F32 (fp32 out), VLFM_EPI_BIAS_RESID_F32
// (fp32 out += ...), VLFM_EPI_BIAS_GELU_F16X2 (GELU, then fp16 x2 operands into d_out / d_out_lo).
extern "C" int vlfm_gemm_f16x2(const void* d_A_hi, const void* d_A_lo, const void* d_W_hi, const void* d_W_lo, const float* d_bias, void* d_out,
                               void* d_out_lo, int M, int N, int K, int lda, int ldw, int ldo, int epilogue, void* stream) {
  int rc = x2_args_ok("vlfm_gemm_f16x2", d_A_hi, d_A_lo, d_W_hi, d_W_lo, d_out, M, N, K, lda, ldw, ldo);
  if (rc) return rc;
  if (epilogue != VLFM_EPI_BIAS_F32 && epilogue != VLFM_EPI_BIAS_RESID_F32 && epilogue != VLFM_EPI_BIAS_GELU_F16X2) {
    set_error("vlfm_gemm_f16x2: epilogue %d unsupported (fp32, fp32 residual, GELU x2)", epilogue); return VLFM_E_INVALID; }
  if (epilogue == VLFM_EPI_BIAS_GELU_F16X2 && (!d_out_lo || ((uintptr_t)d_out_lo & 15))) { set_error("vlfm_gemm_f16x2: d_out_lo missing / unaligned"); return VLFM_E_INVALID; }
  GemmArgs g{d_bias, d_out, M, N, K, ldo, epilogue, (K + BK - 1) / BK, nullptr, 0, nullptr, 0, 0, 0, d_out_lo, BM};
  return gemm_x2_dispatch(d_A_hi, d_A_lo, d_W_hi, d_W_lo, M, N, K, lda, ldw, g, stream, nullptr, 0, nullptr);
}

extern "C" int vlfm_layernorm_x2(const float* d_x, const float* d_gamma, const float* d_beta, void* d_out_hi, void* d_out_lo, float* d_out32,
                                 int rows, int D, int ldx, int ldo16, int ldo32, float eps, void* stream);
extern "C" int vlfm_layernorm_reduce_x2(float* d_x, const float* d_partials, int splits, long long split_stride, const float* d_gamma,
                                        const float* d_beta, void* d_out_hi, void* d_out_lo, float* d_out32, int rows, int D, int ldx, int ldo16,
                                        int ldo32, float eps, void* stream);

// x += (x2 product) + bias ; LayerNorm(x) -> x2 operands (hi, lo) and/or fp32.  Bitwise reproducible like vlfm_gemm_f16_resid_ln.
extern "C" int vlfm_gemm_f16x2_resid_ln(const void* d_A_hi, const void* d_A_lo, const void* d_W_hi, const void* d_W_lo, const float* d_bias,
                                        float* d_x, int M, int N, int K, int lda, int ldw, int ldx, const float* d_gamma, const float* d_beta,
                                        void* d_out_hi, void* d_out_lo, int ld16, float* d_out32, int ld32, float eps, float* d_partials,
                                        size_t partial_bytes, void* stream) {
  int rc = x2_args_ok("vlfm_gemm_f16x2_resid_ln", d_A_hi, d_A_lo, d_W_hi, d_W_lo, d_x, M, N, K, lda, ldw, ldx);
  if (rc) return rc;
  if (!d_gamma || !d_beta || !d_out_hi || !d_out_lo || (N & 3) || (ld16 & 3) || (ld32 & 3) || ((uintptr_t)d_partials & 15)) {
    set_error("vlfm_gemm_f16x2_resid_ln: bad argument / alignment"); return VLFM_E_INVALID; }
  GemmArgs g{d_bias, d_x, M, N, K, ldx, VLFM_EPI_BIAS_RESID_F32, (K + BK - 1) / BK, nullptr, 0, nullptr, 0, 0, 0, nullptr, BM};
  int splits = 1;
  rc = gemm_x2_dispatch(d_A_hi, d_A_lo, d_W_hi, d_W_lo, M, N, K, lda, ldw, g, stream, d_partials, partial_bytes, &splits);
  if (rc) return rc;
  if (splits > 1) return vlfm_layernorm_reduce_x2(d_x, d_partials, splits, (long long)M * N, d_gamma, d_beta, d_out_hi, d_out_lo, d_out32, M, N, ldx, ld16, ld32, eps, stream);
  return vlfm_layernorm_x2(d_x, d_gamma, d_beta, d_out_hi, d_out_lo, d_out32, M, N, ldx, ld16, ld32, eps, stream);
}
