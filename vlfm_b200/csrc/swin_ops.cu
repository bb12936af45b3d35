// Non-GEMM operators of the GroundingDINO Swin-T backbone (sm_100a).
// Reference call site: vlfm/vlm/grounding_dino.py:52-67 (to_tensor + ImageNet normalise, no resize,
// then groundingdino's Swin-T backbone inside predict()).
//   - swin_patch_im2col_kernel: uint8 HWC -> (x/255 - mean)/std (float32) -> fp16 rows of the 4x4/4
//     patch-embedding GEMM (zero padding to a multiple of 4 AFTER normalisation, as the conv sees it);
//   - swin_window_attention_kernel: (shifted-)window multi-head attention, 7x7 windows, hd=32:
//     gathers q/k/v through the cyclic shift and the pad-to-7 (padded tokens are LayerNorm zeros, so
//     their q/k/v equal the projection bias), adds the relative-position bias and the shift mask
//     (-100 across regions), softmax, PV, scatters back (reverse shift + crop);
//   - swin_patch_merge_kernel: 2x2 neighbourhood gather [x00,x10,x01,x11] with zero padding.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace vlfm {

__global__ void swin_patch_im2col_kernel(const uint8_t* __restrict__ img, __half* __restrict__ out, int B, int H, int W,
                                         int Hp, int Wp, float m0, float m1, float m2, float s0, float s1, float s2) {
  // out [B*Hp*Wp, 48], col = c*16 + ky*4 + kx
  const size_t n = (size_t)B * Hp * Wp * 48;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % 48);
    const size_t r = i / 48;
    const int px = (int)(r % Wp), py = (int)((r / Wp) % Hp), b = (int)(r / ((size_t)Wp * Hp));
    const int c = col / 16, ky = (col % 16) / 4, kx = col % 4;
    const int y = py * 4 + ky, x = px * 4 + kx;
    float v = 0.f;
    if (y < H && x < W) {
      const float p = (float)img[(((size_t)b * H + y) * W + x) * 3 + c];
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
      v = __fdiv_rn(__fsub_rn(__fdiv_rn(p, 255.f), mean), sd);
    }
    out[i] = __float2half_rn(v);
  }
}

struct WinAttnArgs {
  const __half* qkv;      // [B*H*W, 3C]
  const float* qkv_bias;  // [3C]
  const float* rel_bias;  // [169, heads]
  __half* out;            // [B*H*W, C]
  int H, W, C, heads, shift;
};

constexpr int WS = 7, WT = 49, WHD = 32;

__global__ void __launch_bounds__(128)
swin_window_attention_kernel(WinAttnArgs a) {
  __shared__ float sq[WT][WHD + 1], sk[WT][WHD + 1], sv[WT][WHD + 1];
  __shared__ float sS[WT][WT + 1];
  __shared__ int srow[WT], sreg[WT];
  pdl_trigger();
  pdl_wait();
  const int Hp = (a.H + WS - 1) / WS * WS, Wp = (a.W + WS - 1) / WS * WS;
  const int nwx = Wp / WS, nwy = Hp / WS;
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / nwx, wx = win % nwx;
  const int tid = threadIdx.x;
  if (tid < WT) {
    const int ty = tid / WS, tx = tid % WS;
    const int y = wy * WS + ty, x = wx * WS + tx;            // coordinates in the shifted, padded frame
    const int ys = (y + a.shift) % Hp, xs = (x + a.shift) % Wp;  // source token (torch.roll by -shift)
    srow[tid] = (ys < a.H && xs < a.W) ? (b * a.H + ys) * a.W + xs : -1;
    int ry = 0, rx = 0;
    if (a.shift > 0) {
      ry = y < Hp - WS ? 0 : (y < Hp - a.shift ? 1 : 2);
      rx = x < Wp - WS ? 0 : (x < Wp - a.shift ? 1 : 2);
    }
    sreg[tid] = ry * 3 + rx;
  }
  __syncthreads();
  const int C = a.C;
  for (int i = tid; i < WT * WHD; i += 128) {
    const int t = i / WHD, d = i % WHD;
    const int row = srow[t];
    const int cq = head * WHD + d;
    float q, k, v;
    if (row >= 0) {
      const __half* p = a.qkv + (size_t)row * 3 * C;
      q = __half2float(p[cq]); k = __half2float(p[C + cq]); v = __half2float(p[2 * C + cq]);
    } else {  // padded token: LayerNorm output was padded with zeros -> projection = bias (rounded like the GEMM output)
      q = __half2float(__float2half_rn(a.qkv_bias[cq]));
      k = __half2float(__float2half_rn(a.qkv_bias[C + cq]));
      v = __half2float(__float2half_rn(a.qkv_bias[2 * C + cq]));
    }
    sq[t][d] = q * 0.17677669529663687f;  // 1/sqrt(32)
    sk[t][d] = k; sv[t][d] = v;
  }
  __syncthreads();
  for (int i = tid; i < WT * WT; i += 128) {
    const int qi = i / WT, kj = i % WT;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < WHD; ++d) s += sq[qi][d] * sk[kj][d];
    const int dy = qi / WS - kj / WS + WS - 1, dx = qi % WS - kj % WS + WS - 1;
    s += a.rel_bias[(dy * (2 * WS - 1) + dx) * a.heads + head];
    if (sreg[qi] != sreg[kj]) s += -100.f;
    sS[qi][kj] = s;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < WT; r += 4) {
    float v0 = lane < WT ? sS[r][lane] : -INFINITY, v1 = lane + 32 < WT ? sS[r][lane + 32] : -INFINITY;
    float m = fmaxf(v0, v1);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float e0 = lane < WT ? __expf(v0 - m) : 0.f, e1 = lane + 32 < WT ? __expf(v1 - m) : 0.f;
    float s = e0 + e1;
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.f / s;
    if (lane < WT) sS[r][lane] = e0 * inv;
    if (lane + 32 < WT) sS[r][lane + 32] = e1 * inv;
  }
  __syncthreads();
  for (int i = tid; i < WT * WHD; i += 128) {
    const int t = i / WHD, d = i % WHD;
    const int row = srow[t];
    if (row < 0) continue;
    float o = 0.f;
#pragma unroll 7
    for (int j = 0; j < WT; ++j) o += sS[t][j] * sv[j][d];
    a.out[(size_t)row * C + head * WHD + d] = __float2half_rn(o);
  }
}

__global__ void swin_patch_merge_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H, int W, int C) {
  // x [B,H,W,C] -> out [B, H2*W2, 4C]; channel blocks: (0,0), (1,0), (0,1), (1,1) (row offset, col offset)
  pdl_trigger();
  pdl_wait();
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const size_t n = (size_t)B * H2 * W2 * 4 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (4 * C));
    const size_t r = i / (4 * C);
    const int x2 = (int)(r % W2), y2 = (int)((r / W2) % H2), b = (int)(r / ((size_t)W2 * H2));
    const int blk = c4 / C, c = c4 % C;
    const int y = 2 * y2 + (blk & 1), xx = 2 * x2 + (blk >> 1);
    out[i] = (y < H && xx < W) ? x[(((size_t)b * H + y) * W + xx) * C + c] : 0.f;
  }
}

}  // namespace vlfm

using namespace vlfm;

extern "C" int vlfm_swin_patch_im2col(const uint8_t* d_img, void* d_out, int B, int H, int W, const float* h_mean3,
                                      const float* h_std3, void* stream) {
  if (!d_img || !d_out || !h_mean3 || !h_std3 || B < 1) { set_error("vlfm_swin_patch_im2col: bad argument"); return VLFM_E_INVALID; }
  const int Hp = (H + 3) / 4, Wp = (W + 3) / 4;
  size_t n = (size_t)B * Hp * Wp * 48;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  swin_patch_im2col_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_img, (__half*)d_out, B, H, W, Hp, Wp, h_mean3[0], h_mean3[1],
                                                                     h_mean3[2], h_std3[0], h_std3[1], h_std3[2]);
  VLFM_CHECK_LAUNCH("swin_patch_im2col_kernel");
  count_launch();
  return VLFM_OK;
}

extern "C" int vlfm_swin_window_attention(const void* d_qkv, const float* d_qkv_bias, const float* d_rel_bias, void* d_out, int B,
                                          int H, int W, int C, int heads, int shift, void* stream) {
  if (!d_qkv || !d_qkv_bias || !d_rel_bias || !d_out || B < 1 || heads < 1 || C != heads * WHD || shift < 0 || shift >= WS) {
    set_error("vlfm_swin_window_attention: bad argument (head_dim must be 32, window 7)"); return VLFM_E_INVALID; }
  WinAttnArgs a{(const __half*)d_qkv, d_qkv_bias, d_rel_bias, (__half*)d_out, H, W, C, heads, shift};
  const int nw = ((H + WS - 1) / WS) * ((W + WS - 1) / WS);
  int rc = check_cuda(launch_pdl(swin_window_attention_kernel, dim3(nw, heads, B), dim3(128), 0, (cudaStream_t)stream, a),
                      "swin_window_attention_kernel");
  if (rc) return rc;
  count_launch();
  return VLFM_OK;
}

extern "C" int vlfm_swin_patch_merge(const float* d_x, float* d_out, int B, int H, int W, int C, void* stream) {
  if (!d_x || !d_out || B < 1) { set_error("vlfm_swin_patch_merge: bad argument"); return VLFM_E_INVALID; }
  size_t n = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * 4 * C;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  int rc = check_cuda(launch_pdl(swin_patch_merge_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, d_x, d_out, B, H, W, C),
                      "swin_patch_merge_kernel");
  if (rc) return rc;
  count_launch();
  return VLFM_OK;
}
