// GroundingDINO feature-enhancer / decoder kernels that are not GEMMs.
//
// Replaces, on the path vlfm/vlm/grounding_dino.py:61-67 -> groundingdino ... MSDeformAttn, the third-party
// `ms_deform_attn_cuda.cu` gather: multi-scale deformable attention
//   out[b,q,h,:] = sum_{l,p} w[b,q,h,l,p] * bilinear(value_l[b,:,h,:], loc[b,q,h,l,p])      (zeros padding,
// align_corners=False, exactly torch.nn.functional.grid_sample's unnormalisation).
#include "common.cuh"

namespace vlfm {

constexpr int MSDA_MAX_LEVELS = 8;
struct MsdaArgs {
  int B, S, Q, heads, hd, levels, points;
  int H[MSDA_MAX_LEVELS], W[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS];
};

template <typename T> __device__ __forceinline__ float ld_val(const T* p);
template <> __device__ __forceinline__ float ld_val<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_val<__half>(const __half* p) { return __half2float(__ldg(p)); }

// one warp per (b, q, head); lane = channel (hd <= 32 per pass).  Every tap is one coalesced hd*sizeof(T) read.
template <typename T>
__global__ void __launch_bounds__(256)
msda_kernel(const T* __restrict__ value, const float* __restrict__ loc, const float* __restrict__ attw, float* __restrict__ out, MsdaArgs a) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long total = (long)a.B * a.Q * a.heads;
  if (warp >= total) return;
  const int h = (int)(warp % a.heads);
  const long bq = warp / a.heads;
  const int b = (int)(bq / a.Q);
  const int LP = a.levels * a.points;
  const float* lp = loc + warp * LP * 2;
  const float* wp = attw + warp * LP;
  const T* vb = value + (size_t)b * a.S * a.heads * a.hd + (size_t)h * a.hd;
  const size_t vstride = (size_t)a.heads * a.hd;
  for (int c0 = 0; c0 < a.hd; c0 += 32) {
    const int c = c0 + lane;
    const bool cv = c < a.hd;
    float acc = 0.f;
    for (int j0 = 0; j0 < LP; j0 += 16) {                      // 16 (level, point) pairs per coalesced 128-byte load
      const int nj = min(16, LP - j0);
      const float mine = lane < 2 * nj ? __ldg(lp + 2 * j0 + lane) : 0.f;
      const float wmine = lane < nj ? __ldg(wp + j0 + lane) : 0.f;
      for (int jj = 0; jj < nj; ++jj) {
        const float x = __shfl_sync(0xffffffffu, mine, 2 * jj), y = __shfl_sync(0xffffffffu, mine, 2 * jj + 1);
        const float w = __shfl_sync(0xffffffffu, wmine, jj);
        const int l = (j0 + jj) / a.points;
        const int Wl = a.W[l], Hl = a.H[l];
        // HF: grid = 2*loc - 1 ; grid_sample: ((grid + 1) / 2) * size - 0.5   (no FMA contraction)
        const float gx = __fsub_rn(__fmul_rn(2.f, x), 1.f), gy = __fsub_rn(__fmul_rn(2.f, y), 1.f);
        const float ix = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)Wl), 0.5f);
        const float iy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)Hl), 0.5f);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const float tx = ix - fx0, ty = iy - fy0;
        const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
        const bool xin0 = x0 >= 0 && x0 < Wl, xin1 = x0 + 1 >= 0 && x0 + 1 < Wl, yin0 = y0 >= 0 && y0 < Hl, yin1 = y0 + 1 >= 0 && y0 + 1 < Hl;
        if (!cv || !((xin0 || xin1) && (yin0 || yin1))) continue;
        const T* base = vb + (size_t)a.start[l] * vstride + c;
        float s = 0.f;
        if (yin0 && xin0) s += w00 * ld_val<T>(base + (size_t)(y0 * Wl + x0) * vstride);
        if (yin0 && xin1) s += w01 * ld_val<T>(base + (size_t)(y0 * Wl + x0 + 1) * vstride);
        if (yin1 && xin0) s += w10 * ld_val<T>(base + (size_t)((y0 + 1) * Wl + x0) * vstride);
        if (yin1 && xin1) s += w11 * ld_val<T>(base + (size_t)((y0 + 1) * Wl + x0 + 1) * vstride);
        acc += w * s;
      }
    }
    if (cv) out[(size_t)warp * a.hd + c] = acc;
  }
}

// Fused deformable attention for head_dim 32 / fp16 values.  One warp per (b, q, head).  Lanes 0-15 own one (level, point)
// pair each for the softmax and the location arithmetic; the gather runs with lane = (tap = lane >> 3, channel quad =
// lane & 7): all four bilinear taps of a point are ONE 8-byte load per lane, the taps are summed by two shuffles at the end.
template <int REF_DIM, int LEVELS, int POINTS>      // LEVELS == 0: runtime levels / points
__global__ void __launch_bounds__(256)
msda_fused_kernel(const __half* __restrict__ value, const float* __restrict__ offlog, int ld, int logit_col, const float* __restrict__ ref,
                  __half* __restrict__ out, MsdaArgs a) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long total = (long)a.B * a.Q * a.heads;
  if (warp >= total) return;
  const int h = (int)(warp % a.heads);
  const long bq = warp / a.heads;
  const int b = (int)(bq / a.Q);
  const int points = LEVELS ? POINTS : a.points;
  const int LP = LEVELS ? LEVELS * POINTS : a.levels * a.points;
  const float* row = offlog + (size_t)bq * ld;
  // ---- per-point quantities on lanes 0..LP-1: softmax weight, integer tap origin, bilinear fractions, level base
  float logit = -INFINITY, ax = 0.f, ay = 0.f;
  int x0 = -4, y0 = -4, Wl = 1, Hl = 1, lbase = 0;
  if (lane < LP) {
    const float2 off = __ldg(reinterpret_cast<const float2*>(row + (size_t)h * LP * 2) + lane);
    logit = __ldg(row + logit_col + h * LP + lane);
    const int l = lane / points;
    Wl = a.W[l]; Hl = a.H[l]; lbase = a.start[l];
    const float* rp = ref + ((size_t)bq * (LEVELS ? LEVELS : a.levels) + l) * REF_DIM;
    float lx, ly;
    if (REF_DIM == 2) {
      lx = __fadd_rn(__ldg(rp), __fdiv_rn(off.x, (float)Wl));
      ly = __fadd_rn(__ldg(rp + 1), __fdiv_rn(off.y, (float)Hl));
    } else {
      lx = __fadd_rn(__ldg(rp), __fmul_rn(__fmul_rn(__fdiv_rn(off.x, (float)points), __ldg(rp + 2)), 0.5f));
      ly = __fadd_rn(__ldg(rp + 1), __fmul_rn(__fmul_rn(__fdiv_rn(off.y, (float)points), __ldg(rp + 3)), 0.5f));
    }
    // HF: grid = 2*loc - 1 ; grid_sample: ((grid + 1) / 2) * size - 0.5   (no FMA contraction; /2 is exact)
    const float gx = __fsub_rn(__fmul_rn(2.f, lx), 1.f), gy = __fsub_rn(__fmul_rn(2.f, ly), 1.f);
    const float ix = __fsub_rn(__fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)Wl), 0.5f);
    const float iy = __fsub_rn(__fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)Hl), 0.5f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    ax = ix - fx0; ay = iy - fy0;
    // clamp far-away samples so the int conversion is defined; anything <= -2 or >= size contributes nothing
    x0 = (int)fminf(fmaxf(fx0, -4.f), 16384.f); y0 = (int)fminf(fmaxf(fy0, -4.f), 16384.f);
  }
  float mx = logit;
#pragma unroll
  for (int o = 8; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  mx = __shfl_sync(0xffffffffu, mx, 0);                 // lanes 16-31 hold -inf: take the low half's maximum
  const float e = lane < LP ? expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 8; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  sum = __shfl_sync(0xffffffffu, sum, 0);
  const float wgt = e / sum;
  // ---- gather: lane = (tap, channel quad)
  const int tap = lane >> 3, quad = lane & 7;
  const int tx = tap & 1, ty = tap >> 1;
  const __half* vb = value + (size_t)b * a.S * a.heads * 32 + (size_t)h * 32 + quad * 4;
  const size_t vstride = (size_t)a.heads * 32;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
  for (int j = 0; j < (LEVELS ? LEVELS * POINTS : 16); ++j) {
    if (!LEVELS && j >= LP) break;
    const float fx = __shfl_sync(0xffffffffu, ax, j), fy = __shfl_sync(0xffffffffu, ay, j), w = __shfl_sync(0xffffffffu, wgt, j);
    const int px = __shfl_sync(0xffffffffu, x0, j) + tx, py = __shfl_sync(0xffffffffu, y0, j) + ty;
    const int pw = __shfl_sync(0xffffffffu, Wl, j), ph = __shfl_sync(0xffffffffu, Hl, j), pb = __shfl_sync(0xffffffffu, lbase, j);
    const float bw = (tx ? fx : 1.f - fx) * (ty ? fy : 1.f - fy) * w;
    if (px >= 0 && px < pw && py >= 0 && py < ph) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(vb + (size_t)(pb + py * pw + px) * vstride));
      const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), f23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
      acc0 += bw * f01.x; acc1 += bw * f01.y; acc2 += bw * f23.x; acc3 += bw * f23.y;
    }
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, o); acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
    acc2 += __shfl_xor_sync(0xffffffffu, acc2, o); acc3 += __shfl_xor_sync(0xffffffffu, acc3, o);
  }
  if (lane < 8) {
    const __half2 o01 = __floats2half2_rn(acc0, acc1), o23 = __floats2half2_rn(acc2, acc3);
    uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&o01); pk.y = *reinterpret_cast<const uint32_t*>(&o23);
    *reinterpret_cast<uint2*>(out + (size_t)warp * 32 + quad * 4) = pk;
  }
}

__global__ void cast_addpos_kernel(const float4* __restrict__ x, const float4* __restrict__ pos, uint2* __restrict__ ox, uint2* __restrict__ oxp, long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    if (ox) {
      const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
      uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b); ox[i] = o;
    }
    if (oxp) {
      const float4 p = pos ? __ldg(pos + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const __half2 a = __floats2half2_rn(v.x + p.x, v.y + p.y), b = __floats2half2_rn(v.z + p.z, v.w + p.w);
      uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b); oxp[i] = o;
    }
  }
}

__global__ void cast_f32_f16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long n4, const float* __restrict__ in_tail,
                                    __half* __restrict__ out_tail, int tail) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + i);
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
    out[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < tail) out_tail[threadIdx.x] = __float2half_rn(in_tail[threadIdx.x]);
}

}  // namespace vlfm

using namespace vlfm;

extern "C" int vlfm_msda_forward(const void* d_value, int value_is_f16, const float* d_loc, const float* d_attw, float* d_out, int B, int S,
                                 int Q, int heads, int hd, int levels, int points, const int32_t* h_shapes_hw, void* stream) {
  if (!d_value || !d_loc || !d_attw || !d_out || !h_shapes_hw || B < 1 || S < 1 || Q < 1 || heads < 1 || hd < 1 || points < 1 ||
      levels < 1 || levels > MSDA_MAX_LEVELS) {
    set_error("vlfm_msda_forward: bad argument"); return VLFM_E_INVALID;
  }
  MsdaArgs a{};
  a.B = B; a.S = S; a.Q = Q; a.heads = heads; a.hd = hd; a.levels = levels; a.points = points;
  int acc = 0;
  for (int l = 0; l < levels; ++l) { a.H[l] = h_shapes_hw[2 * l]; a.W[l] = h_shapes_hw[2 * l + 1]; a.start[l] = acc; acc += a.H[l] * a.W[l]; }
  if (acc != S) { set_error("vlfm_msda_forward: spatial shapes sum to %d, value has %d positions", acc, S); return VLFM_E_INVALID; }
  const long warps = (long)B * Q * heads;
  const long blocks = (warps + 7) / 8;
  if (blocks > 0x7fffffffL) { set_error("vlfm_msda_forward: too many queries"); return VLFM_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (value_is_f16) msda_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)d_value, d_loc, d_attw, d_out, a);
  else msda_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)d_value, d_loc, d_attw, d_out, a);
  VLFM_CHECK_LAUNCH("vlfm_msda_forward");
  count_launch();
  return VLFM_OK;
}

extern "C" int vlfm_cast_f32_f16(const float* d_in, void* d_out16, long n, void* stream) {
  if (!d_in || !d_out16 || n < 0) { set_error("vlfm_cast_f32_f16: bad argument"); return VLFM_E_INVALID; }
  if (n == 0) return VLFM_OK;
  if (((uintptr_t)d_in & 15) || ((uintptr_t)d_out16 & 7)) { set_error("vlfm_cast_f32_f16: pointers must be 16 / 8 byte aligned"); return VLFM_E_INVALID; }
  const long n4 = n >> 2; const int tail = (int)(n & 3);
  long blocks = (n4 + 255) / 256; if (blocks < 1) blocks = 1; if (blocks > 148 * 16) blocks = 148 * 16;
  cast_f32_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)d_in, (uint2*)d_out16, n4, d_in + 4 * n4,
                                                                          (__half*)d_out16 + 4 * n4, tail);
  VLFM_CHECK_LAUNCH("vlfm_cast_f32_f16");
  count_launch();
  return VLFM_OK;
}

extern "C" int vlfm_msda_fused(const void* d_value16, const float* d_offlog, int ld, int logit_col, const float* d_ref, int ref_dim,
                               void* d_out16, int B, int S, int Q, int heads, int levels, int points, const int32_t* h_shapes_hw,
                               void* stream) {
  if (!d_value16 || !d_offlog || !d_ref || !d_out16 || !h_shapes_hw || B < 1 || S < 1 || Q < 1 || heads < 1 || points < 1 || levels < 1 ||
      levels > MSDA_MAX_LEVELS || levels * points > 16 || (ref_dim != 2 && ref_dim != 4) || (ld & 1) || ((heads * levels * points * 2) & 1)) {
    set_error("vlfm_msda_fused: bad argument"); return VLFM_E_INVALID;
  }
  MsdaArgs a{};
  a.B = B; a.S = S; a.Q = Q; a.heads = heads; a.hd = 32; a.levels = levels; a.points = points;
  int acc = 0;
  for (int l = 0; l < levels; ++l) { a.H[l] = h_shapes_hw[2 * l]; a.W[l] = h_shapes_hw[2 * l + 1]; a.start[l] = acc; acc += a.H[l] * a.W[l]; }
  if (acc != S) { set_error("vlfm_msda_fused: spatial shapes sum to %d, value has %d positions", acc, S); return VLFM_E_INVALID; }
  const long blocks = ((long)B * Q * heads + 7) / 8;
  if (blocks > 0x7fffffffL) { set_error("vlfm_msda_fused: too many queries"); return VLFM_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const __half* v = (const __half*)d_value16; __half* o = (__half*)d_out16;
  const bool std44 = levels == 4 && points == 4;      // GroundingDINO's configuration: fully unrolled gather
  if (ref_dim == 2) {
    if (std44) msda_fused_kernel<2, 4, 4><<<(unsigned)blocks, 256, 0, st>>>(v, d_offlog, ld, logit_col, d_ref, o, a);
    else msda_fused_kernel<2, 0, 0><<<(unsigned)blocks, 256, 0, st>>>(v, d_offlog, ld, logit_col, d_ref, o, a);
  } else {
    if (std44) msda_fused_kernel<4, 4, 4><<<(unsigned)blocks, 256, 0, st>>>(v, d_offlog, ld, logit_col, d_ref, o, a);
    else msda_fused_kernel<4, 0, 0><<<(unsigned)blocks, 256, 0, st>>>(v, d_offlog, ld, logit_col, d_ref, o, a);
  }
  VLFM_CHECK_LAUNCH("vlfm_msda_fused");
  count_launch();
  return VLFM_OK;
}

extern "C" int vlfm_cast_addpos_f16(const float* d_x, const float* d_pos, void* d_out_x16, void* d_out_xp16, long n, void* stream) {
  if (!d_x || (!d_out_x16 && !d_out_xp16) || n < 0 || (n & 3)) { set_error("vlfm_cast_addpos_f16: bad argument"); return VLFM_E_INVALID; }
  if (n == 0) return VLFM_OK;
  const long n4 = n >> 2;
  long blocks = (n4 + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16;
  cast_addpos_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)d_x, (const float4*)d_pos, (uint2*)d_out_x16, (uint2*)d_out_xp16, n4);
  VLFM_CHECK_LAUNCH("vlfm_cast_addpos_f16");
  count_launch();
  return VLFM_OK;
}
