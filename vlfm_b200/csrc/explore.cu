// Explore half of the obstacle map on the GPU (sm_100a): fog-of-war, explored-area component selection,
// frontier waypoints.
//
// Reference: vlfm/mapping/obstacle_map.py:114-169 and the two `frontier_exploration` functions it calls
// (reveal_fog_of_war, detect_frontier_waypoints; third-party, absent from the reference tree).  The SPEC these
// kernels follow step by step is oracle/explore_oracle.py (numpy backend) with oracle/contours.py,
// oracle/cv_draw.py and oracle/cv_prims.py -- restatements of the OpenCV primitives pinned against cv2.
//
// Building blocks
//   ccl_*            label-equivalence connected components (union-find, atomicMin): 8-connected foreground
//                    and 4-connected background; root = raster-first pixel of the component.
//   collect_roots    cv2.findContours(RETR_EXTERNAL): outer borders of the components whose west background
//                    region is the outer background, in REVERSE raster order of their first pixels.
//   trace_kernel     Suzuki-Abe border following (one thread per contour; borders are short), CHAIN_APPROX_NONE
//                    chain + bounding box; CHAIN_APPROX_SIMPLE vertices = direction changes of the chain.
//   fill_chain_*     cv2.drawContours(..., -1) of a traced chain: outline + even-odd scan conversion as XOR toggles
//                    + per-row prefix XOR (every chain edge is a unit step, so intercepts are exact).
//   sector_fill      cv2.ellipse filled sector: 16.16 polygon from the host (integer-degree ellipse2Poly), same
//                    scan conversion with fractional columns.
//   rays / thick     occlusion rays: cv2.polylines thickness 2 = FillConvexPoly rectangle (Line2 outline + two-edge
//                    scan) + radius-1 discs.
//   frontier_kernel  contour split at cells whose 3x3 blurred unexplored mask is 0, arc-length midpoints.
#include <math.h>

#include "common.cuh"

namespace vlfm {

constexpr int EX_MAXC = 8192;      // contours per image
constexpr int XYS = 16;
constexpr long long XYONE = 1ll << XYS;

struct Contour { int start, off, len, x0, y0, x1, y1, ed; };   // start pixel, chain offset/length, bbox, entry direction (0 = west: outer border, 4 = east: hole border)

// device-side bookkeeping of one explore step
struct ExState {
  int n_cont;          // contours of the image being processed
  int cursor;          // chain buffer cursor
  int n_rays;
  int skip_fog;        // reveal_fog_of_war returned the (empty) input mask
  int chosen;          // selected contour index
  int n_front;
  int overflow;
  int pad;
};

// ------------------------------------------------------------------------------------------- CCL ----
__device__ __forceinline__ int uf_find(int* L, int i) {
  while (true) {
    int p = *reinterpret_cast<volatile int*>(&L[i]);
    if (p == i) return i;
    i = p;
  }
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a); b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// init: every pixel points at the first pixel of its horizontal run, for the foreground (8-connected) and the background
// (4-connected) label arrays at once, so that the union phase only has to stitch runs of adjacent rows.  One warp per
// row, 32 cells per ballot.  Block 0 also resets
// the per-image bookkeeping.
__global__ void __launch_bounds__(256)
ccl_init2_kernel(const uint8_t* __restrict__ img, int* __restrict__ Lfg, int* __restrict__ Lbg, int W, int H, ExState* st, int keep_fog) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->n_cont = 0; st->cursor = 0; st->n_rays = 0; st->chosen = -1;
    if (!keep_fog) { st->skip_fog = 0; st->overflow = 0; st->n_front = 0; }
  }
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < H; y += gridDim.x * wpb) {
    const uint8_t* row = img + (size_t)y * W;
    const int base = y * W;
    int carry_fg = -1, carry_bg = -1;                 // column where a run that reaches the chunk boundary started
    const unsigned below = (1u << lane) - 1;
    for (int x0 = 0; x0 < W; x0 += 32) {
      const int x = x0 + lane;
      const bool valid = x < W;
      const bool fg = valid && row[x] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, fg);
      const unsigned zf = ~m & below, zb = m & below;  // cells below this lane that end a fg / bg run
      const int sf = zf ? x0 + 32 - __clz(zf) : (carry_fg >= 0 ? carry_fg : x0);
      const int sb = zb ? x0 + 32 - __clz(zb) : (carry_bg >= 0 ? carry_bg : x0);
      if (valid) { Lfg[base + x] = fg ? base + sf : -1; Lbg[base + x] = fg ? -1 : base + sb; }
      const int sf31 = __shfl_sync(0xffffffffu, sf, 31), sb31 = __shfl_sync(0xffffffffu, sb, 31);
      carry_fg = (m >> 31) ? sf31 : -1;
      carry_bg = (m >> 31) ? -1 : sb31;
    }
  }
}
// stitch: a pixel unions with the row above only where a NEW overlap between runs begins; also clears the per-label flags
__global__ void ccl_merge2_kernel(const uint8_t* __restrict__ img, int* __restrict__ Lfg, int* __restrict__ Lbg, uint8_t* __restrict__ outer,
                                  uint8_t* __restrict__ hashole, uint8_t* __restrict__ nbm, int W, int H) {
  const int n = W * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    outer[i] = 0;
    if (hashole) hashole[i] = 0;
    const int y = i / W, x = i - y * W;
    const bool c = img[i] != 0;
    if (c) {   // 8-neighbourhood mask for the border tracer, bit d = neighbour in direction d (clockwise from west)
      const bool l = x > 0, r = x + 1 < W, u = y > 0, d = y + 1 < H;
      unsigned m = 0;
      if (l && img[i - 1]) m |= 1u;
      if (l && u && img[i - W - 1]) m |= 2u;
      if (u && img[i - W]) m |= 4u;
      if (r && u && img[i - W + 1]) m |= 8u;
      if (r && img[i + 1]) m |= 16u;
      if (r && d && img[i + W + 1]) m |= 32u;
      if (d && img[i + W]) m |= 64u;
      if (l && d && img[i + W - 1]) m |= 128u;
      nbm[i] = (uint8_t)m;
    }
    if (y == 0) continue;
    const bool west = x > 0 && ((img[i - 1] != 0) == c);
    const bool north = (img[i - W] != 0) == c;
    const bool nwest = x > 0 && ((img[i - W - 1] != 0) == c);
    int* L = c ? Lfg : Lbg;
    if (north && (!west || !nwest)) uf_union(L, i, i - W);
    if (c) {   // foreground is 8-connected: diagonal contacts not already implied by a north contact
      const bool neast = x + 1 < W && img[i - W + 1] != 0;
      if (nwest && !north && !west) uf_union(L, i, i - W - 1);
      if (neast && !north) uf_union(L, i, i - W + 1);
    }
  }
}
__global__ void ccl_flatten2_kernel(int* __restrict__ Lfg, int* __restrict__ Lbg, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (Lfg[i] >= 0) Lfg[i] = uf_find(Lfg, i); else Lbg[i] = uf_find(Lbg, i);
  }
}
// background components touching the image frame are the "outer" background (the frame is background for Suzuki)
__global__ void bg_outer_kernel(const int* __restrict__ Lbg, uint8_t* __restrict__ outer, int W, int H) {
  const int per = 2 * (W + H);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per; t += gridDim.x * blockDim.x) {
    int x, y;
    if (t < W) { x = t; y = 0; } else if (t < 2 * W) { x = t - W; y = H - 1; }
    else if (t < 2 * W + H) { x = 0; y = t - 2 * W; } else { x = W - 1; y = t - 2 * W - H; }
    const int l = Lbg[y * W + x];
    if (l >= 0) outer[l] = 1;
  }
}
__global__ void collect_roots_kernel(const int* __restrict__ Lfg, const int* __restrict__ Lbg, const uint8_t* __restrict__ outer,
                                     int W, int H, Contour* __restrict__ cont, ExState* st) {
  const int n = W * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (Lfg[i] != i) continue;                      // roots only (= raster-first pixel of the component)
    const int y = i / W, x = i - y * W;
    bool top = true;
    if (x > 0) { const int lb = Lbg[i - 1]; top = lb >= 0 && outer[lb]; }
    if (!top) continue;                             // nested inside a hole of another component: not external
    const int k = atomicAdd(&st->n_cont, 1);
    if (k < EX_MAXC) { cont[k].start = i; cont[k].ed = 0; } else st->overflow = 1;
  }
}
// a foreground component that directly encloses a background region (a hole): flag its root
__global__ void mark_holes_kernel(const int* __restrict__ Lfg, const int* __restrict__ Lbg, const uint8_t* __restrict__ outer,
                                  uint8_t* __restrict__ hashole, int W, int H) {
  const int n = W * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (Lbg[i] != i || outer[i]) continue;            // raster-first pixel of an enclosed background region
    const int x = i % W;
    if (x > 0 && Lfg[i - 1] >= 0) hashole[Lfg[i - 1]] = 1;
  }
}
// reverse raster order (cv2 returns the last-found contour first); one block, bitonic sort in shared memory
__global__ void __launch_bounds__(1024) sort_roots_kernel(Contour* __restrict__ cont, ExState* st) {
  __shared__ int keys[EX_MAXC];
  int n = st->n_cont;
  if (n > EX_MAXC) n = EX_MAXC;
  __syncthreads();
  if (threadIdx.x == 0) st->n_cont = n;
  int P = 2;
  while (P < n) P <<= 1;                              // bitonic network over the next power of two only
  for (int i = threadIdx.x; i < P; i += 1024) keys[i] = i < n ? cont[i].start : -1;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += 1024) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int a = keys[i], b = keys[ixj];
          const bool desc = (i & k) == 0;           // descending overall
          if ((a < b) == desc) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += 1024) { cont[i].start = keys[i]; cont[i].ed = 0; }
}

// ----------------------------------------------------------------------------------------- tracing ----
// directions 0..7: clockwise from west (image coordinates, y down): dx = {-1,-1,0,1,1,1,0,-1}, dy = {0,-1,-1,-1,0,1,1,1}

// direction tables packed into registers (a dynamically indexed __constant__ array costs a dependent LDC per probe)
__device__ __forceinline__ int dir_dx(int d) { return (int)((0x01222100u >> (d * 4)) & 0xFu) - 1; }
__device__ __forceinline__ int dir_dy(int d) { return (int)((0x22210001u >> (d * 4)) & 0xFu) - 1; }
__device__ __forceinline__ unsigned rotr8(unsigned v, int r) { return ((v >> r) | (v << (8 - r))) & 0xFFu; }
// Suzuki-Abe steps 3.1-3.5 (oracle/contours.py::_trace) on the per-pixel neighbour masks written by ccl_merge2_kernel: one
// byte load per step; the clockwise / counter-clockwise probe loops are a byte rotation + ffs / clz.  WRITE=false only counts.
template <bool WRITE>
__device__ int trace_border(const uint8_t* __restrict__ nbm, int W, int x0, int y0, int2* out, Contour* c, int ed = 0) {
  int n = 0, minx = x0, maxx = x0, miny = y0, maxy = y0;
  // 3.1 clockwise from the (zero) entry pixel (west for outer, east for hole borders): first set bit among ed+1 .. ed+7
  const unsigned r0 = rotr8(nbm[y0 * W + x0], ed) & 0xFEu;
  if (!r0) {
    if (WRITE) out[0] = make_int2(x0, y0);
    n = 1;
  } else {
    const int df = (ed + __ffs(r0) - 1) & 7;
    const int fx = x0 + dir_dx(df), fy = y0 + dir_dy(df);
    int x3 = x0, y3 = y0, d0 = df;                   // d0: direction from (x3,y3) to the previously examined pixel (x2,y2)
    while (true) {
      // 3.3 counter-clockwise, starting after (x2,y2): probes d0-1, ..., d0-8 are bits 7..0 after rotating right by d0, so the
      // first hit is the highest set bit (bit 0 = (x2,y2) itself is always set)
      const unsigned r = rotr8(nbm[y3 * W + x3], d0);
      const int di = (d0 + (31 - __clz(r))) & 7;
      const int x4 = x3 + dir_dx(di), y4 = y3 + dir_dy(di);
      if (WRITE) { out[n] = make_int2(x3, y3); minx = min(minx, x3); maxx = max(maxx, x3); miny = min(miny, y3); maxy = max(maxy, y3); }
      ++n;
      if (x4 == x0 && y4 == y0 && x3 == fx && y3 == fy) break;   // 3.5
      x3 = x4; y3 = y4; d0 = (di + 4) & 7;          // seen from the new pixel, the old one lies in the opposite direction
      if (n > (1 << 22)) break;                      // safety
    }
  }
  if (WRITE) { c->x0 = minx; c->x1 = maxx; c->y0 = miny; c->y1 = maxy; }
  return n;
}
// mode 0: trace every contour; 1: only when more than one contour exists (component selection); 2: skip components
// that enclose a hole (their filled polygon contains a zero cell, so F1 can never absorb them)
__global__ void trace_kernel(const uint8_t* __restrict__ nbm, int W, int H, Contour* __restrict__ cont, int2* __restrict__ chain,
                             int cap, ExState* st, int mode, const uint8_t* __restrict__ hashole) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= st->n_cont || c >= gridDim.x * blockDim.x) return;
  const int s = cont[c].start, y0 = s / W, x0 = s - y0 * W;
  if ((mode == 1 && st->n_cont <= 1) || (mode == 2 && hashole[s])) { cont[c].off = 0; cont[c].len = 0; return; }
  const int n = trace_border<false>(nbm, W, x0, y0, nullptr, nullptr, cont[c].ed);
  const int off = atomicAdd(&st->cursor, n);
  if (off + n > cap) { st->overflow = 1; cont[c].off = 0; cont[c].len = 0; return; }
  cont[c].off = off; cont[c].len = n;
  trace_border<true>(nbm, W, x0, y0, chain + off, &cont[c], cont[c].ed);
}

// CHAIN_APPROX_SIMPLE: point i of a chain is kept iff the step into it differs from the step out of it
__device__ __forceinline__ bool simple_vertex(const int2* p, int n, int i) {
  if (n <= 2) return true;
  const int2 a = p[i == 0 ? n - 1 : i - 1], b = p[i], c = p[i + 1 == n ? 0 : i + 1];
  return (b.x - a.x != c.x - b.x) || (b.y - a.y != c.y - b.y);
}

// cv2.pointPolygonTest(cnt, pt, True) on the SIMPLE vertices (oracle/contours.py::point_polygon_distance)
__device__ double ppt_distance(const int2* p, int n, int ptx, int pty) {
  if (n == 0) return -1.7976931348623157e308;
  const float px = (float)ptx, py = (float)pty;
  int last = -1;
  for (int i = n - 1; i >= 0; --i) if (simple_vertex(p, n, i)) { last = i; break; }
  if (last < 0) last = n - 1;
  float vx = (float)p[last].x, vy = (float)p[last].y;
  double min_num = 3.4028234663852886e38, min_den = 1.0;
  int counter = 0;
  for (int i = 0; i < n; ++i) {
    if (!simple_vertex(p, n, i)) continue;
    const float v0x = vx, v0y = vy;
    vx = (float)p[i].x; vy = (float)p[i].y;
    const double dx = vx - v0x, dy = vy - v0y, dx1 = px - v0x, dy1 = py - v0y, dx2 = px - vx, dy2 = py - vy;
    double num, den = 1.0;
    if (dx1 * dx + dy1 * dy <= 0) num = dx1 * dx1 + dy1 * dy1;
    else if (dx2 * dx + dy2 * dy >= 0) num = dx2 * dx2 + dy2 * dy2;
    else { num = dy1 * dx - dx1 * dy; num *= num; den = dx * dx + dy * dy; }
    if (num * min_den < min_num * den) { min_num = num; min_den = den; if (min_num == 0) break; }
    if ((v0y <= py && vy <= py) || (v0y > py && vy > py)) continue;
    double cr = dy1 * dx - dx1 * dy;
    if (dy < 0) cr = -cr;
    counter += cr > 0;
  }
  const double r = sqrt(min_num / min_den);
  return (counter & 1) ? r : -r;
}

// ------------------------------------------------------------------------------ scan conversion ----
// planes: tog / orb, `pw` 32-bit words per row, rows [0, H)
__global__ void zero_planes_kernel(uint32_t* tog, uint32_t* orb, int words) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) { tog[i] = 0; orb[i] = 0; }
}
// chain polygon: every edge is a unit step; outline = the chain points
__global__ void chain_edges_kernel(const Contour* __restrict__ cont, const int* __restrict__ which, const int2* __restrict__ chain,
                                   uint32_t* tog, uint32_t* orb, int W, int H, int pw, const ExState* st) {
  const int ci = which ? *which : blockIdx.y;
  if (ci < 0 || ci >= st->n_cont) return;
  const Contour c = cont[ci];
  const int2* p = chain + c.off;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.len; i += gridDim.x * blockDim.x) {
    const int2 a = p[i], b = p[i + 1 == c.len ? 0 : i + 1];
    atomicOr(&orb[a.y * pw + (a.x >> 5)], 1u << (a.x & 31));
    if (a.y != b.y) {
      const int xa = a.y < b.y ? a.x : b.x, ya = min(a.y, b.y);     // active on row ya only; intercept exactly xa
      const int t = xa + 1;
      if (t < W) atomicXor(&tog[ya * pw + (t >> 5)], 1u << (t & 31));
      atomicOr(&orb[ya * pw + (xa >> 5)], 1u << (xa & 31));
    }
  }
}
// rows -> image: img[cell] = value where filled; optionally everything else := 0 (clear_rest)
__global__ void planes_to_image_kernel(const uint32_t* __restrict__ tog, const uint32_t* __restrict__ orb, uint8_t* __restrict__ img, int W,
                                       int H, int pw, int value, int clear_rest, const int* __restrict__ enable) {
  if (enable && *enable < 0) return;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < H; r += gridDim.x * blockDim.x) {
    uint32_t carry = 0;
    for (int w = 0; w < pw; ++w) {
      const uint32_t t = tog[r * pw + w];
      uint32_t x = t;
      x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
      x ^= carry;
      if (__popc(t) & 1) carry = ~carry;
      const uint32_t f = x | orb[r * pw + w];
      for (int bq = 0; bq < 32; ++bq) {
        const int col = w * 32 + bq;
        if (col >= W) break;
        if ((f >> bq) & 1u) img[r * W + col] = (uint8_t)value;
        else if (clear_rest) img[r * W + col] = 0;
      }
    }
  }
}

// cv::clipLine(Size2l(W, H), pt1, pt2) (oracle/cv_prims.py::clip_line): Cohen-Sutherland, intersections in double, truncated
// toward zero.  cv2 clips every line to the image before walking it, so a line that leaves the grid is the walk of the CLIPPED
// segment.  The end points are modified even when the function returns false (as in OpenCV).
__device__ __forceinline__ long long clip_isect(long long a, long long b, long long c) {   // (int64)((double)a * b / c)
  return (long long)__ddiv_rn(__dmul_rn((double)a, (double)b), (double)c);
}
__device__ bool clip_line(long long W, long long H, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = W - 1, bottom = H - 1;
  if (W <= 0 || H <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += clip_isect(a - y1, x2 - x1, y2 - y1); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
    if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += clip_isect(a - y2, x2 - x1, y2 - y1); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) { a = c1 == 1 ? 0 : right; y1 += clip_isect(a - x1, y2 - y1, x2 - x1); x1 = a; c1 = 0; }
      if (c2) { a = c2 == 1 ? 0 : right; y2 += clip_isect(a - x2, y2 - y1, x2 - x1); x2 = a; c2 = 0; }
    }
  }
  return (c1 | c2) == 0;
}

// cv2.ellipse filled sector: polygon (x, y in 16.16, last vertex = centre) from the host; CollectPolyEdges +
// FillEdgeCollection (oracle/cv_draw.py::fill_poly_fixed / poly_edge).  One block; the image is the GW x GH grid, of which the
// W x H window at (ox, oy) is rasterised (the window may hang over the grid edge: those cells are masked by the caller).
// 8-connected line between two pixels of the grid (already clipped), plotted into the window's outline plane
__device__ __forceinline__ void plot_line8(uint32_t* orb, int pw, int W, int H, int ox, int oy, int ax, int ay, int bx, int by) {
  if (bx < ax) { int t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
  const int dx = bx - ax, dy = by - ay, sy = dy >= 0 ? 1 : -1, ady = dy >= 0 ? dy : -dy;
  const bool ymaj = ady > dx;
  const int major = ymaj ? ady : dx, minor = ymaj ? dx : ady;
  int x = ax - ox, y = ay - oy, err = major - 2 * minor;
  for (int k = 0; k <= major; ++k) {
    if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) atomicOr(&orb[y * pw + (x >> 5)], 1u << (x & 31));
    const bool m = err < 0;
    err += -2 * minor + (m ? 2 * major : 0);
    if (ymaj) { y += sy; x += m ? 1 : 0; } else { x += 1; y += m ? sy : 0; }
  }
}
__global__ void sector_edges_kernel(const long long* __restrict__ v, int nv, int ox, int oy, int GW, int GH, uint32_t* tog, uint32_t* orb,
                                    int W, int H, int pw) {
  for (int e = threadIdx.x; e < nv; e += blockDim.x) {
    const int e0 = e == 0 ? nv - 1 : e - 1;
    const long long x0 = v[2 * e0], x1 = v[2 * e];                                                 // 16.16 columns
    const long long y0 = (v[2 * e0 + 1] + (XYONE >> 1)) >> XYS, y1 = (v[2 * e + 1] + (XYONE >> 1)) >> XYS;   // rounded rows
    long long ax = (x0 + (XYONE >> 1)) >> XYS, ay = y0, bx = (x1 + (XYONE >> 1)) >> XYS, by = y1;        // pixel end points
    const bool outside = (unsigned long long)ax >= (unsigned long long)GW || (unsigned long long)bx >= (unsigned long long)GW ||
                         (unsigned long long)ay >= (unsigned long long)GH || (unsigned long long)by >= (unsigned long long)GH;
    long long cx0 = x0, cy0 = y0, cx1 = x1, cy1 = y1;
    bool vis = true;
    if (outside) {
      vis = clip_line(GW, GH, ax, ay, bx, by);
      // PolyEdge from the clipped columns (always) and the clipped rows (when they differ)
      cx0 = ax << XYS; cx1 = bx << XYS;
      if (ay != by) { cy0 = ay; cy1 = by; }
    }
    if (vis) plot_line8(orb, pw, W, H, ox, oy, (int)ax, (int)ay, (int)bx, (int)by);
    if (y0 == y1) continue;
    const long long dxe = (cx1 - cx0) / (cy1 - cy0);          // C truncating division
    long long ya, yb, xs;
    if (y0 < y1) { ya = y0; yb = y1; xs = cx0 + (y0 - cy0) * dxe; } else { ya = y1; yb = y0; xs = cx1 + (y1 - cy1) * dxe; }
    long long r0 = ya > 0 ? ya : 0, r1 = yb < GH ? yb : GH;
    if (r0 < oy) r0 = oy;
    if (r1 > oy + H) r1 = oy + H;
    for (long long r = r0; r < r1; ++r) {
      const long long X = xs + dxe * (r - ya);
      const int wr = (int)(r - oy);
      long long t = (X >> XYS) + 1 - ox;
      if (t < 0) t = 0;
      if (t < W) atomicXor(&tog[wr * pw + (int)(t >> 5)], 1u << ((int)t & 31));
      if ((X & (XYONE - 1)) == 0) { const long long c = (X >> XYS) - ox; if (c >= 0 && c < W) atomicOr(&orb[wr * pw + (int)(c >> 5)], 1u << ((int)c & 31)); }
    }
  }
}

// ValueMap confidence cone (value_map.py:321-355): sector(0/1 byte image) x remap(cos^2(remap(atan2(|dc|,|dr|), 0, fov/2, 0, pi/2)),
// 0, 1, min_conf, 1) in float64 with numpy's operation order, cast to float32.
__global__ void cone_template_kernel(const uint8_t* __restrict__ sector, float* __restrict__ out, int R, double fov, double min_conf) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * R; i += gridDim.x * blockDim.x) {
    const int r = i / R, c = i - r * R;
    const double dr = fabs((double)(r - R / 2)), dc = fabs((double)(c - R / 2));
    const double ang = __ddiv_rn(__dmul_rn(atan2(dc, dr), 3.14159265358979323846 / 2), fov / 2);
    const double cs = cos(ang);
    const double conf = __dadd_rn(__dmul_rn(__dmul_rn(cs, cs), 1.0 - min_conf), min_conf);
    out[i] = sector[i] ? (float)conf : 0.f;
  }
}

// --------------------------------------------------------------------------------- window images ----
// blocked = cone & !nav ; visible = cone & nav   (window W0 x W0 at grid origin (ox, oy))
__global__ void fog_masks_kernel(const uint8_t* __restrict__ cone, const uint8_t* __restrict__ nav, int G, int ox, int oy, int W0,
                                 uint8_t* __restrict__ blocked, uint8_t* __restrict__ visible) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W0 * W0; i += gridDim.x * blockDim.x) {
    const int y = i / W0, x = i - y * W0;
    const int gx = ox + x, gy = oy + y;
    if ((unsigned)gx >= (unsigned)G || (unsigned)gy >= (unsigned)G) { blocked[i] = 0; visible[i] = 0; continue; }   // cv2 clips at the grid edge
    const uint8_t c = cone[i], nv = nav[(size_t)gy * G + gx];
    blocked[i] = c && !nv; visible[i] = c && nv;
  }
}

// CHAIN_APPROX_SIMPLE vertex list of every contour, compacted in order (one warp per contour)
__global__ void simple_vertices_kernel(const Contour* __restrict__ cont, const int2* __restrict__ chain, int2* __restrict__ sv,
                                       int* __restrict__ nsv, const ExState* st) {
  const int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ci >= st->n_cont) return;
  const Contour c = cont[ci];
  const int2* p = chain + c.off;
  int2* o = sv + c.off;
  int cnt = 0;
  for (int b = 0; b < c.len; b += 32) {
    const int i = b + lane;
    const bool keep = i < c.len && simple_vertex(p, c.len, i);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) o[cnt + __popc(m & ((1u << lane) - 1))] = p[i];
    cnt += __popc(m);
  }
  if (lane == 0) nsv[ci] = cnt;
}
// R3/R4: obstacle contours -> ray list (x0,y0,x1,y1 in window coordinates); one warp per contour
__global__ void rays_kernel(const Contour* __restrict__ cont, const int2* __restrict__ sv, const int* __restrict__ nsv, int4* __restrict__ rays,
                            int cap, int sx, int sy, int ox, int oy, double heading_deg, double ray_len, ExState* st) {
  const int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ci >= st->n_cont) return;
  const int2* v = sv + cont[ci].off;
  const int nv = nsv[ci];
  // cv2.isContourConvex on the SIMPLE vertices (oracle/contours.py::is_convex): every turn has the same strict sign
  int orient = 0;
  for (int j = lane; j < nv; j += 32) {
    const int2 a = v[(j + 2 * nv - 2) % nv], b = v[(j + nv - 1) % nv], c = v[j];
    const long long dx0 = b.x - a.x, dy0 = b.y - a.y, dx = c.x - b.x, dy = c.y - b.y;
    const long long dxdy0 = dx * dy0, dydx0 = dy * dx0;
    orient |= dydx0 > dxdy0 ? 1 : (dydx0 < dxdy0 ? 2 : 3);
  }
  orient = __reduce_or_sync(0xffffffffu, orient);
  const bool convex = nv > 0 && orient != 3;
  auto emit = [&](int qx, int qy) {
    const double ang = atan2((double)(qy - sy), (double)(qx - sx));
    // astype(np.int32) truncates toward zero in GRID coordinates (the window origin is subtracted afterwards)
    const int ex = (int)((double)(qx + ox) + ray_len * cos(ang)) - ox, ey = (int)((double)(qy + oy) + ray_len * sin(ang)) - oy;
    const int k = atomicAdd(&st->n_rays, 1);
    if (k < cap) rays[k] = make_int4(qx, qy, ex, ey); else st->overflow = 1;
  };
  if (convex) {
    // _extreme_bearing_points: the heading in DEGREES is used as radians, as in the restated package; np.argmin /
    // np.argmax return the FIRST extreme vertex
    const double cs = cos(-heading_deg), sn = sin(-heading_deg);
    double amin = 1e300, amax = -1e300; int imin = 0x7fffffff, imax = 0x7fffffff;
    for (int j = lane; j < nv; j += 32) {
      const double qx = (double)(v[j].x - sx), qy = (double)(v[j].y - sy);
      const double rx = qx * cs + qy * sn, ry = qx * (-sn) + qy * cs;
      const double a = atan2(ry, rx);
      if (a < amin) { amin = a; imin = j; }
      if (a > amax) { amax = a; imax = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double a1 = __shfl_xor_sync(0xffffffffu, amin, o), a2 = __shfl_xor_sync(0xffffffffu, amax, o);
      const int i1 = __shfl_xor_sync(0xffffffffu, imin, o), i2 = __shfl_xor_sync(0xffffffffu, imax, o);
      if (a1 < amin || (a1 == amin && i1 < imin)) { amin = a1; imin = i1; }
      if (a2 > amax || (a2 == amax && i2 < imax)) { amax = a2; imax = i2; }
    }
    if (lane == 0) { emit(v[imin].x, v[imin].y); emit(v[imax].x, v[imax].y); }
  } else {
    for (int j = lane; j < nv; j += 32) emit(v[j].x, v[j].y);
  }
}

// cv2 thickness-2 line into the byte image `cut` (oracle/cv_draw.py::thick_line2).  All geometry is in GRID coordinates (the
// clipping rules refer to the grid); `cut` is the window at (ox, oy), pixels outside it are skipped.
struct CutWin { uint8_t* img; int W, H, ox, oy; };
__device__ __forceinline__ void put_px(const CutWin& c, long long x, long long y) {
  x -= c.ox; y -= c.oy;
  if (x >= 0 && x < c.W && y >= 0 && y < c.H) c.img[y * c.W + x] = 1;
}
__device__ __forceinline__ long long cdiv(long long a, long long b) {   // C truncating division (b > 0)
  return a / b;
}
// drawing.cpp Line2: clipLine against the image scaled to 16.16, then a DDA between the clipped end points
__device__ void line2_fixed(const CutWin& c, int G, long long x1, long long y1, long long x2, long long y2) {
  if (!clip_line((long long)G << XYS, (long long)G << XYS, x1, y1, x2, y2)) return;
  long long dx = x2 - x1, dy = y2 - y1;
  const long long ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
  long long x_step, y_step, ecount;
  if (ax > ay) {
    if (dx < 0) { long long t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; dy = -dy; }
    x_step = XYONE; y_step = cdiv(dy << XYS, ax | 1); ecount = (x2 - x1) >> XYS;
  } else {
    if (dy < 0) { long long t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; dx = -dx; }
    x_step = cdiv(dx << XYS, ay | 1); y_step = XYONE; ecount = (y2 - y1) >> XYS;
  }
  x1 += XYONE >> 1; y1 += XYONE >> 1;
  put_px(c, (x2 + (XYONE >> 1)) >> XYS, (y2 + (XYONE >> 1)) >> XYS);
  if (ax > ay) {
    long long x = x1 >> XYS, y = y1;
    while (ecount >= 0) { put_px(c, x, y >> XYS); ++x; y += y_step; --ecount; }
  } else {
    long long y = y1 >> XYS, x = x1;
    while (ecount >= 0) { put_px(c, x >> XYS, y); x += x_step; ++y; --ecount; }
  }
}
__device__ __forceinline__ long long pick4(const long long (&a)[4], int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3])); }
__global__ void thick_rays_kernel(const int4* __restrict__ rays, uint8_t* __restrict__ cut, int W, int H, int ox, int oy, int G, const ExState* st) {
  const int ri = blockIdx.x * blockDim.x + threadIdx.x;
  if (ri >= st->n_rays) return;
  const int4 r = rays[ri];
  const CutWin cw{cut, W, H, ox, oy};
  // ThickLine (cv2 4.13): the integer centre line is first clipped to the image grown by the thickness on every side
  long long px0 = r.x + ox + 2, py0 = r.y + oy + 2, px1 = r.z + ox + 2, py1 = r.w + oy + 2;
  if (!clip_line((long long)G + 4, (long long)G + 4, px0, py0, px1, py1)) return;
  px0 -= 2; py0 -= 2; px1 -= 2; py1 -= 2;
  const long long x0 = px0 << XYS, y0 = py0 << XYS, x1 = px1 << XYS, y1 = py1 << XYS;
  const double dx = (double)(x0 - x1) / 65536.0, dy = (double)(y1 - y0) / 65536.0;
  double rr = dx * dx + dy * dy;
  if (fabs(rr) > 2.220446049250313e-16) {
    rr = 65536.0 / sqrt(rr);                                      // thickness 2 -> half width one pixel (16.16)
    const long long dpx = (long long)rint(dy * rr), dpy = (long long)rint(dx * rr);
    long long vx[4] = {x0 + dpx, x0 - dpx, x1 - dpx, x1 + dpx}, vy[4] = {y0 + dpy, y0 - dpy, y1 - dpy, y1 + dpy};
    // FillConvexPoly (shift = 16): Line2 outline ...
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int j = (i + 3) & 3; line2_fixed(cw, G, vx[j], vy[j], vx[i], vy[i]); }
    // ... + two-edge scan
    const long long delta = XYONE >> 1;
    int imin = 0;
    long long ymin_f = vy[0], ymax_f = vy[0], xmin_f = vx[0], xmax_f = vx[0];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (vy[i] < ymin_f) { ymin_f = vy[i]; imin = i; }
      ymax_f = vy[i] > ymax_f ? vy[i] : ymax_f; xmax_f = vx[i] > xmax_f ? vx[i] : xmax_f; xmin_f = vx[i] < xmin_f ? vx[i] : xmin_f;
    }
    long long ymin = (ymin_f + delta) >> XYS, ymax = (ymax_f + delta) >> XYS;
    const long long xmin = (xmin_f + delta) >> XYS, xmax = (xmax_f + delta) >> XYS;
    if (!(xmax < 0 || ymax < 0 || xmin >= G || ymin >= G)) {       // OpenCV's early-out refers to the grid
      if (ymax > G - 1) ymax = G - 1;
      struct { int idx, di; long long x, dx; long long ye; } e[2];
      e[0].idx = e[1].idx = imin; e[0].ye = e[1].ye = ymin; e[0].di = 1; e[1].di = 3;
      e[0].x = e[1].x = -XYONE; e[0].dx = e[1].dx = 0;
      int edges = 4;
      long long y = ymin;
      do {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (y >= e[i].ye) {
            int idx0 = e[i].idx, di = e[i].di, idx = (idx0 + di) & 3;
            for (; edges-- > 0;) {
              const long long ty = (pick4(vy, idx) + delta) >> XYS;
              if (ty > y) {
                const long long xs = pick4(vx, idx0), xe = pick4(vx, idx);
                e[i].ye = ty; e[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y)); e[i].x = xs; e[i].idx = idx;
                break;
              }
              idx0 = idx; idx = (idx + di) & 3;
            }
          }
        }
        if (edges < 0) break;
        {
          const bool sw = e[0].x > e[1].x;
          long long xx1 = ((sw ? e[1].x : e[0].x) + delta) >> XYS, xx2 = ((sw ? e[0].x : e[1].x) + delta) >> XYS;
          const long long wy = y - oy;
          if (y >= 0 && wy >= 0 && wy < H) {
            xx1 -= ox; xx2 -= ox;
            for (long long x = xx1 < 0 ? 0 : xx1; x <= xx2 && x < W; ++x) cut[wy * W + x] = 1;
          }
        }
        e[0].x += e[0].dx; e[1].x += e[1].dx;
      } while (++y <= ymax);
    }
  }
  // Circle(center, 1, filled) at both (clipped) ends
  const long long cxs[2] = {px0, px1}, cys[2] = {py0, py1};
  for (int k = 0; k < 2; ++k) {
    put_px(cw, cxs[k], cys[k]); put_px(cw, cxs[k] - 1, cys[k]); put_px(cw, cxs[k] + 1, cys[k]);
    put_px(cw, cxs[k], cys[k] - 1); put_px(cw, cxs[k], cys[k] + 1);
  }
}
__global__ void apply_cut_kernel(uint8_t* __restrict__ visible, const uint8_t* __restrict__ cut, int n, const ExState* st) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) if (cut[i]) visible[i] = 0;
}
// no obstacle contour in the cone -> reveal_fog_of_war returns the (all-zero) input mask
__global__ void fog_gate_kernel(ExState* st) { if (st->n_cont == 0) st->skip_fog = 1; }

// R5: pick the contour with the smallest |pointPolygonTest| to the agent; > 3 px -> nothing revealed
__global__ void pick_nearest_kernel(const Contour* __restrict__ cont, const int2* __restrict__ chain, double* __restrict__ dist, int sx, int sy,
                                    ExState* st, int phase) {
  if (phase == 0) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= st->n_cont) return;
    dist[ci] = ppt_distance(chain + cont[ci].off, cont[ci].len, sx, sy);
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    int best = -1; double bd = INFINITY;
    for (int i = 0; i < st->n_cont; ++i) { const double d = fabs(dist[i]); if (d < bd) { bd = d; best = i; } }
    st->chosen = (st->skip_fog || bd > 3.0) ? -1 : best;
    if (st->chosen < 0) st->skip_fog = 1;
  }
}

// new = dilate3(newexp); explored |= new (window); then explored &= nav over the obstacle-update window
__global__ void explored_update_kernel(const uint8_t* __restrict__ newexp, int W0, int ox, int oy, uint8_t* __restrict__ explored,
                                       const uint8_t* __restrict__ nav, int G, int rx0, int ry0, int rw, int rh, const ExState* st) {
  const int n = rw * rh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int gy = ry0 + i / rw, gx = rx0 + i % rw;
    uint8_t e = explored[(size_t)gy * G + gx];
    if (!st->skip_fog) {
      const int wx = gx - ox, wy = gy - oy;
      bool hit = false;
      for (int dy = -1; dy <= 1 && !hit; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = wx + dx, y = wy + dy;
          if ((unsigned)x < (unsigned)W0 && (unsigned)y < (unsigned)W0 && newexp[y * W0 + x]) { hit = true; break; }
        }
      if (hit) e = 1;
    }
    if (nav[(size_t)gy * G + gx] == 0) e = 0;
    explored[(size_t)gy * G + gx] = e;
  }
}

// obstacle_map.py:133-146: more than one external contour -> keep the first (cv2 order) that contains the agent
// (dist >= 0), else the nearest; the kept one is redrawn FILLED
__global__ void select_component_kernel(const Contour* __restrict__ cont, const int2* __restrict__ chain, double* __restrict__ dist, int ax, int ay,
                                        ExState* st, int* __restrict__ which, int phase) {
  if (phase == 0) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= st->n_cont || st->n_cont <= 1) return;
    dist[ci] = ppt_distance(chain + cont[ci].off, cont[ci].len, ax, ay);
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    int best = -1;
    if (st->n_cont > 1) {
      double md = INFINITY; best = 0;
      for (int i = 0; i < st->n_cont; ++i) {
        const double d = dist[i];
        if (d >= 0) { best = i; break; }
        if (fabs(d) < md) { md = fabs(d); best = i; }
      }
    }
    *which = best;      // -1: a single contour, explored area stays as it is
  }
}

// k x k box dilation of a 0/1 byte image (whole image; used for the 5x5 growth of the explored area)
__global__ void dilate_full_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H, int k) {
  const int r = k / 2, n = W * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    uint8_t m = 0;
    for (int dy = -r; dy <= r && !m; ++dy) {
      const int yy = y + dy;
      if ((unsigned)yy >= (unsigned)H) continue;
      for (int dx = -r; dx <= r; ++dx) { const int xx = x + dx; if ((unsigned)xx < (unsigned)W && src[yy * W + xx]) { m = 1; break; } }
    }
    dst[i] = m;
  }
}
// unexplored = nav & !grown
__global__ void unexplored_kernel(const uint8_t* __restrict__ nav, const uint8_t* __restrict__ grown, uint8_t* __restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (nav[i] && !grown[i]) ? 1 : 0;
}

// F1: small unexplored pockets (contourArea < thresh, filled mask only covers unexplored cells) are absorbed
// into the explored mask as 255.  One block per contour; cells of the bounding box are tested against the chain polygon.
__global__ void absorb_small_kernel(const Contour* __restrict__ cont, const int2* __restrict__ chain, const uint8_t* __restrict__ unexplored,
                                    uint8_t* __restrict__ explored2, int W, int H, double area_thresh, const ExState* st) {
  const int ci = blockIdx.x;
  if (ci >= st->n_cont) return;
  const Contour c = cont[ci];
  if (c.len == 0) return;                       // not traced: the component encloses a hole
  const int2* p = chain + c.off;
  __shared__ long long s_a2;
  __shared__ int s_bad;
  if (threadIdx.x == 0) { s_a2 = 0; s_bad = 0; }
  __syncthreads();
  long long acc = 0;                                             // shoelace (twice the signed area), exact in integers
  for (int i = threadIdx.x; i < c.len; i += blockDim.x) {
    const int2 q = p[i == 0 ? c.len - 1 : i - 1], b = p[i];
    acc += (long long)q.x * b.y - (long long)b.x * q.y;
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(&s_a2), (unsigned long long)acc);
  __syncthreads();
  const double area = fabs((double)s_a2 * 0.5);
  if (!(area < area_thresh)) return;
  const int bw = c.x1 - c.x0 + 1, bh = c.y1 - c.y0 + 1;
  // pass 1: every cell drawContours would fill must be an unexplored (== 1) cell
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
      const int x = c.x0 + i % bw, y = c.y0 + i / bw;
      // even-odd with the exact-hit rule (unit edges: intercept is the lower endpoint's column) + outline
      int less = 0; bool exact = false;
      for (int e = 0; e < c.len; ++e) {
        const int2 a = p[e], b = p[e + 1 == c.len ? 0 : e + 1];
        if (a.x == x && a.y == y) exact = true;
        if (a.y == b.y) continue;
        const int ya = min(a.y, b.y), xa = a.y < b.y ? a.x : b.x;
        if (ya != y) continue;
        if (xa < x) ++less; else if (xa == x) exact = true;
      }
      if (exact || (less & 1)) {
        if (pass == 0) { if (unexplored[y * W + x] != 1) s_bad = 1; }
        else explored2[y * W + x] = 255;
      }
    }
    __syncthreads();
    if (s_bad) return;
  }
}

// F2-F4 (oracle/explore_oracle.py::_interpolate/_split/_midpoint): one thread walks every external contour of the
// grown+absorbed explored mask in cv2 order and emits the frontier midpoints.
__device__ __forceinline__ bool blur_zero(const uint8_t* nav, const uint8_t* ex2, int W, int H, int x, int y) {
  // cv2.blur 3x3 of 255*(nav & !explored2) is 0 iff all nine (BORDER_REFLECT_101) cells are 0
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      int xx = x + dx, yy = y + dy;
      if (xx < 0) xx = -xx; if (xx >= W) xx = 2 * W - 2 - xx;
      if (yy < 0) yy = -yy; if (yy >= H) yy = 2 * H - 2 - yy;
      if (nav[yy * W + xx] && !ex2[yy * W + xx]) return false;
    }
  return true;
}
__device__ void emit_midpoint(const int2* p, int n, int a, int b, int a2, int b2, double* out, int maxf, ExState* st) {
  // the frontier is q[a..b) followed by q[a2..b2) (second range empty unless merged); q[k] = p[(k+1)/2 mod n]
  auto Q = [&](int k) { return p[((k + 1) >> 1) % n]; };
  const int len1 = b - a, len2 = b2 - a2, len = len1 + len2;
  if (len < 2) return;
  auto at = [&](int i) { return i < len1 ? Q(a + i) : Q(a2 + (i - len1)); };
  double total = 0.0;
  for (int i = 0; i + 1 < len; ++i) { const int2 u = at(i), v = at(i + 1); total += sqrt((double)((u.x - v.x) * (u.x - v.x) + (u.y - v.y) * (u.y - v.y))); }
  const double half = total / 2;
  double cum = 0.0, before = 0.0; int seg = 0;
  bool found = false;
  for (int i = 0; i + 1 < len; ++i) {
    const int2 u = at(i), v = at(i + 1);
    const double l = sqrt((double)((u.x - v.x) * (u.x - v.x) + (u.y - v.y) * (u.y - v.y)));
    if (cum + l > half) { seg = i; before = cum; found = true; break; }
    cum += l;
  }
  if (!found) { seg = 0; before = 0.0; }          // np.argmax of an all-False array is 0
  const int2 u = at(seg), v = at(seg + 1);
  const double l = sqrt((double)((u.x - v.x) * (u.x - v.x) + (u.y - v.y) * (u.y - v.y)));
  const double t = (half - before) / l;
  const int k = st->n_front;
  if (k < maxf) { out[2 * k] = u.x + t * (v.x - u.x); out[2 * k + 1] = u.y + t * (v.y - u.y); st->n_front = k + 1; } else st->overflow = 1;
}
// F3 "bad" test of every traced border point at once (the serial walk below only reads the flags)
__global__ void bad_flags_kernel(const int2* __restrict__ chain, const uint8_t* __restrict__ nav, const uint8_t* __restrict__ ex2, int W, int H,
                                 uint8_t* __restrict__ flags, const ExState* st) {
  const int n = st->cursor;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    flags[i] = blur_zero(nav, ex2, W, H, chain[i].x, chain[i].y) ? 1 : 0;
}
__global__ void frontier_kernel(const Contour* __restrict__ cont, const int2* __restrict__ chain, const uint8_t* __restrict__ flags,
                                double* __restrict__ out, int maxf, ExState* st) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  st->n_front = 0;
  for (int ci = 0; ci < st->n_cont; ++ci) {
    const int2* p = chain + cont[ci].off;
    const uint8_t* fl = flags + cont[ci].off;
    const int n = cont[ci].len, m = 2 * n;              // interpolated sequence q has 2n entries
    if (n == 0) continue;
    auto bad = [&](int k) { return fl[((k + 1) >> 1) % n] != 0; };
    // bad indices split q; piece 0 = [0, b0), piece j = [b_{j-1}, b_j) minus its first element, last = [b_last, m)
    int nbad = 0, first_bad = -1, last_bad = -1;
    for (int k = 0; k < m; ++k) if (bad(k)) { if (first_bad < 0) first_bad = k; last_bad = k; ++nbad; }
    const bool wrap = nbad > 0 && first_bad != 0 && last_bad < m - 2;
    if (nbad == 0) { if (m > 2) emit_midpoint(p, n, 0, m, 0, 0, out, maxf, st); continue; }   // a single piece is kept iff len > 2
    // collect kept pieces in order; with wrap the LAST kept piece is prepended to the FIRST kept piece
    // pass 1: find first kept piece and last kept piece
    int fk_a = -1, fk_b = -1, lk_a = -1, lk_b = -1, nkept = 0;
    {
      int prev = 0, idx = 0;
      for (int k = 0; k <= m; ++k) {
        if (k == m || bad(k)) {
          const int a = prev, b = k, len = b - a;
          const bool keep = (len > 2) || (idx == 0 && wrap);
          if (keep) {
            const int ka = idx == 0 ? a : a + 1;
            if (nkept == 0) { fk_a = ka; fk_b = b; }
            lk_a = ka; lk_b = b; ++nkept;
          }
          prev = k; ++idx;
        }
      }
    }
    if (nkept == 0) continue;
    const bool merge = nkept > 1 && wrap;
    // pass 2: emit in order
    {
      int prev = 0, idx = 0, seen = 0;
      for (int k = 0; k <= m; ++k) {
        if (k == m || bad(k)) {
          const int a = prev, b = k, len = b - a;
          const bool keep = (len > 2) || (idx == 0 && wrap);
          if (keep) {
            const int ka = idx == 0 ? a : a + 1;
            ++seen;
            if (merge && seen == 1) { if ((lk_b - lk_a) + (fk_b - fk_a) >= 2) emit_midpoint(p, n, lk_a, lk_b, fk_a, fk_b, out, maxf, st); }
            else if (merge && seen == nkept) { /* consumed by the merge */ }
            else if (b - ka >= 2) emit_midpoint(p, n, ka, b, 0, 0, out, maxf, st);
          }
          prev = k; ++idx;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- fill_small_holes (img_utils.py:361-390) ----
__global__ void zero_mask_kernel(const float* __restrict__ depth, uint8_t* __restrict__ mask, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mask[i] = depth[i] == 0.f ? 1 : 0;
}
// RETR_TREE: the outer border of EVERY component and the border of every hole
__global__ void collect_all_kernel(const int* __restrict__ Lfg, const int* __restrict__ Lbg, const uint8_t* __restrict__ outer, int W, int H,
                                   Contour* __restrict__ cont, int maxc, ExState* st) {
  const int n = W * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int start = -1, ed = 0;
    if (Lfg[i] == i) { start = i; ed = 0; }
    else if (Lbg[i] == i && !outer[i]) { start = i - 1; ed = 4; }     // hole: the pixel west of its raster-first cell, entered from the east
    if (start < 0) continue;
    const int k = atomicAdd(&st->n_cont, 1);
    if (k < maxc) { cont[k].start = start; cont[k].ed = ed; } else st->overflow = 1;
  }
}
// one block per contour: contourArea < thresh -> drawContours(filled, [cnt], 0, 1, -1); bounding-box scan conversion in
// shared memory, processed in row bands when the box is tall
__global__ void __launch_bounds__(256)
fill_small_contours_kernel(const Contour* __restrict__ cont, const int2* __restrict__ chain, uint8_t* __restrict__ filled, int W, int H,
                           double area_thresh, int smem_words, int maxc, const ExState* st) {
  extern __shared__ uint32_t fs_smem[];
  __shared__ long long s_a2;
  const int nc = min(st->n_cont, maxc);
  for (int ci = blockIdx.x; ci < nc; ci += gridDim.x) {
    const Contour c = cont[ci];
    if (c.len == 0) continue;
    const int2* p = chain + c.off;
    __syncthreads();
    if (threadIdx.x == 0) s_a2 = 0;
    __syncthreads();
    long long acc = 0;
    for (int i = threadIdx.x; i < c.len; i += blockDim.x) {
      const int2 q = p[i == 0 ? c.len - 1 : i - 1], b = p[i];
      acc += (long long)q.x * b.y - (long long)b.x * q.y;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_a2), (unsigned long long)acc);
    __syncthreads();
    if (!(fabs((double)s_a2 * 0.5) < area_thresh)) continue;
    if (c.len == 1) { if (threadIdx.x == 0) filled[(size_t)c.y0 * W + c.x0] = 1; continue; }   // speckle: the common case
    const int bw = c.x1 - c.x0 + 2, pw = (bw + 31) / 32;           // +1 column for the toggle right of the last cell
    const int band = smem_words / (2 * pw);
    for (int r0 = c.y0; r0 <= c.y1; r0 += band) {
      const int rows = min(band, c.y1 - r0 + 1);
      uint32_t* tog = fs_smem; uint32_t* orb = fs_smem + rows * pw;
      for (int i = threadIdx.x; i < 2 * rows * pw; i += blockDim.x) fs_smem[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < c.len; i += blockDim.x) {
        const int2 a = p[i], b = p[i + 1 == c.len ? 0 : i + 1];
        if (a.y >= r0 && a.y < r0 + rows) atomicOr(&orb[(a.y - r0) * pw + ((a.x - c.x0) >> 5)], 1u << ((a.x - c.x0) & 31));
        if (a.y != b.y) {
          const int xa = (a.y < b.y ? a.x : b.x) - c.x0, ya = min(a.y, b.y);
          if (ya >= r0 && ya < r0 + rows) {
            atomicXor(&tog[(ya - r0) * pw + ((xa + 1) >> 5)], 1u << ((xa + 1) & 31));
            atomicOr(&orb[(ya - r0) * pw + (xa >> 5)], 1u << (xa & 31));
          }
        }
      }
      __syncthreads();
      for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        uint32_t carry = 0;
        for (int w = 0; w < pw; ++w) {
          const uint32_t t = tog[r * pw + w];
          uint32_t x = t;
          x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
          x ^= carry;
          if (__popc(t) & 1) carry = ~carry;
          uint32_t f = x | orb[r * pw + w];
          while (f) {
            const int bq = __ffs(f) - 1; f &= f - 1;
            const int col = c.x0 + w * 32 + bq;
            if (col <= c.x1) filled[(size_t)(r0 + r) * W + col] = 1;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void sticky_status_kernel(const ExState* st, int32_t* status) { if (st->overflow) *status |= 1; }
__global__ void reset_state_kernel(ExState* st, int keep_fog) {
  st->n_cont = 0; st->cursor = 0; st->n_rays = 0; st->chosen = -1;
  if (!keep_fog) { st->skip_fog = 0; st->overflow = 0; st->n_front = 0; }
}
__global__ void clear_bytes_kernel(uint8_t* p, int n) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0; }
__global__ void copy_bytes_kernel(const uint8_t* s, uint8_t* d, int n) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = s[i]; }
__global__ void fog_fill_gate_kernel(const ExState* st, int* which) { *which = st->skip_fog ? -1 : st->chosen; }

}  // namespace vlfm

using namespace vlfm;

// -------------------------------------------------------------------------------------- host side ----
namespace {

struct Ws {       // carved from the caller's workspace
  uint8_t *cone, *blocked, *visible, *cut, *newexp, *outer, *grown, *unexp, *ex2, *hashole, *nbm;
  int *Lfg, *Lbg, *which;
  Contour* cont; int2 *chain, *sv; int4* rays; double* dist; uint32_t *tog, *orb; long long* verts; ExState* st;
  int* nsv; uint8_t* flags;
  int chain_cap, rays_cap;
};
constexpr int WIN_MAX = 512;
constexpr int CHAIN_CAP = 1 << 20, RAYS_CAP = 1 << 16, MAXF = 4096;

constexpr int HOLES_MAXC = 1 << 16;   // fill_small_holes sees sensor speckle: many more (tiny) contours than a map does
size_t carve(Ws* w, uint8_t* base, int G, int maxc = EX_MAXC) {
  size_t o = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + o : nullptr; o += (bytes + 255) & ~(size_t)255; return p; };
  const size_t n = (size_t)G * G, wn = (size_t)WIN_MAX * WIN_MAX;
  uint8_t* p;
  p = take(wn); if (w) w->cone = p;
  p = take(wn); if (w) w->blocked = p;
  p = take(wn); if (w) w->visible = p;
  p = take(wn); if (w) w->cut = p;
  p = take(wn); if (w) w->newexp = p;
  p = take(n); if (w) w->outer = p;
  p = take(n); if (w) w->grown = p;
  p = take(n); if (w) w->unexp = p;
  p = take(n); if (w) w->ex2 = p;
  p = take(n); if (w) w->hashole = p;
  p = take(n); if (w) w->nbm = p;
  p = take(n * 4); if (w) w->Lfg = (int*)p;
  p = take(n * 4); if (w) w->Lbg = (int*)p;
  p = take(sizeof(Contour) * (size_t)maxc); if (w) w->cont = (Contour*)p;
  p = take(sizeof(int2) * (size_t)CHAIN_CAP); if (w) w->chain = (int2*)p;
  p = take(sizeof(int2) * (size_t)CHAIN_CAP); if (w) w->sv = (int2*)p;
  p = take((size_t)CHAIN_CAP); if (w) w->flags = p;
  p = take(sizeof(int) * (size_t)maxc); if (w) w->nsv = (int*)p;
  p = take(sizeof(int4) * (size_t)RAYS_CAP); if (w) w->rays = (int4*)p;
  p = take(sizeof(double) * EX_MAXC); if (w) w->dist = (double*)p;
  const size_t pw = ((size_t)G + 31) / 32;
  p = take(pw * G * 4); if (w) w->tog = (uint32_t*)p;
  p = take(pw * G * 4); if (w) w->orb = (uint32_t*)p;
  p = take(64 * 2 * 8); if (w) w->verts = (long long*)p;
  p = take(sizeof(ExState)); if (w) w->st = (ExState*)p;
  p = take(64); if (w) w->which = (int*)p;
  if (w) { w->chain_cap = CHAIN_CAP; w->rays_cap = RAYS_CAP; }
  return o;
}

inline int nblk(long n, int t = 256) { long b = (n + t - 1) / t; return (int)(b < 1 ? 1 : (b > 2368 ? 2368 : b)); }

// external contours of `img` (W x H): CCL fg/bg, top-level roots in cv2 order, traced chains
void contours(const Ws& w, const uint8_t* img, int W, int H, cudaStream_t st, int keep_fog, int mode = 0) {
  const int n = W * H;
  ccl_init2_kernel<<<nblk(H, 8), 256, 0, st>>>(img, w.Lfg, w.Lbg, W, H, w.st, keep_fog);
  ccl_merge2_kernel<<<nblk(n), 256, 0, st>>>(img, w.Lfg, w.Lbg, w.outer, mode == 2 ? w.hashole : nullptr, w.nbm, W, H);
  ccl_flatten2_kernel<<<nblk(n), 256, 0, st>>>(w.Lfg, w.Lbg, n);
  bg_outer_kernel<<<nblk(2 * (W + H)), 256, 0, st>>>(w.Lbg, w.outer, W, H);
  collect_roots_kernel<<<nblk(n), 256, 0, st>>>(w.Lfg, w.Lbg, w.outer, W, H, w.cont, w.st);
  sort_roots_kernel<<<1, 1024, 0, st>>>(w.cont, w.st);
  if (mode == 2) mark_holes_kernel<<<nblk(n), 256, 0, st>>>(w.Lfg, w.Lbg, w.outer, w.hashole, W, H);
  trace_kernel<<<EX_MAXC / 64, 64, 0, st>>>(w.nbm, W, H, w.cont, w.chain, w.chain_cap, w.st, mode, w.hashole);
  count_launch(mode == 2 ? 8 : 7);
}

// cv2.ellipse sector polygon (oracle/cv_draw.py::ellipse_sector), vertices in grid coordinates, 16.16
int sector_polygon(int cx, int cy, int radius, double start_deg, double end_deg, long long* v) {
  static float sintab[451];
  static bool init = false;
  if (!init) {
    for (int a = 0; a <= 450; ++a) { double s = sin(a * 3.14159265358979323846 / 180.0); sintab[a] = (float)(nearbyint(s * 1e7) / 1e7); }
    init = true;
  }
  auto cvr = [](double x) { return (long long)nearbyint(x); };
  int a0 = (int)cvr(start_deg), a1 = (int)cvr(end_deg);
  const long long CX = (long long)cx << XYS, CY = (long long)cy << XYS, AX = (long long)abs(radius) << XYS;
  long long d = (AX + (XYONE >> 1)) >> XYS;
  const int delta = d < 3 ? 90 : d < 10 ? 30 : d < 15 ? 18 : 5;
  if (a0 > a1) { int t = a0; a0 = a1; a1 = t; }
  while (a0 < 0) { a0 += 360; a1 += 360; }
  while (a1 > 360) { a1 -= 360; a0 -= 360; }
  if (a1 - a0 > 360) { a0 = 0; a1 = 360; }
  int nv = 0;
  long long px = 0, py = 0; bool have = false;
  for (int i = a0; i < a1 + delta; i += delta) {
    int ang = i > a1 ? a1 : i;
    if (ang < 0) ang += 360;
    const double x = (double)AX * (double)sintab[450 - ang], y = (double)AX * (double)sintab[ang];
    const double fx = (double)CX + x, fy = (double)CY + y;
    long long qx = cvr(fx / 65536.0) << XYS, qy = cvr(fy / 65536.0) << XYS;
    qx += cvr(fx - (double)qx); qy += cvr(fy - (double)qy);
    if (!have || qx != px || qy != py) { if (nv < 62) { v[2 * nv] = qx; v[2 * nv + 1] = qy; ++nv; } px = qx; py = qy; have = true; }
  }
  if (nv <= 1) { v[0] = CX; v[1] = CY; v[2] = CX; v[3] = CY; nv = 2; }
  v[2 * nv] = CX; v[2 * nv + 1] = CY; ++nv;
  return nv;
}

}  // namespace

extern "C" int vlfm_explore_workspace_bytes(int G, size_t* bytes) {
  if (!bytes || G < 1) { set_error("vlfm_explore_workspace_bytes: bad argument"); return VLFM_E_INVALID; }
  *bytes = carve(nullptr, nullptr, G);
  return VLFM_OK;
}

extern "C" int vlfm_explore_update(int G, uint8_t* d_explored, const uint8_t* d_nav, int agent_col, int agent_row, double heading_deg,
                                   double fov_deg, double max_line_len, double area_thresh_px, int nav_half, double* d_frontiers,
                                   int32_t* d_count, void* d_workspace, int32_t* d_status, void* stream) {
  if (!d_explored || !d_nav || !d_frontiers || !d_count || !d_workspace || !d_status || G < 8) { set_error("vlfm_explore_update: bad argument"); return VLFM_E_INVALID; }
  const int L = (int)max_line_len, W0 = 2 * L + 9;
  if (W0 > WIN_MAX) { set_error("vlfm_explore_update: max_line_len %d too large (window %d > %d)", L, W0, WIN_MAX); return VLFM_E_UNSUPPORTED; }
  // the window is centred on the agent and may hang over the grid edge: cv2 clips the cone and the rays there
  // (clip_line rules above), cells outside the grid are masked out of the window images
  const int ox = agent_col - L - 4, oy = agent_row - L - 4;
  cudaStream_t st = (cudaStream_t)stream;
  Ws w;
  carve(&w, (uint8_t*)d_workspace, G);
  const int n = G * G, wn = W0 * W0, pw0 = (W0 + 31) / 32, pwG = (G + 31) / 32;
  const int sx = agent_col - ox, sy = agent_row - oy;

  // ---- R1: cone sector (window)
  long long hv[128];
  const int nv = sector_polygon(agent_col, agent_row, L, heading_deg - fov_deg / 2, heading_deg + fov_deg / 2, hv);
  int rc = check_cuda(cudaMemcpyAsync(w.verts, hv, sizeof(long long) * 2 * nv, cudaMemcpyHostToDevice, st), "explore: vertex upload");
  if (rc) return rc;
  zero_planes_kernel<<<nblk(pw0 * W0), 256, 0, st>>>(w.tog, w.orb, pw0 * W0);
  sector_edges_kernel<<<1, 64, 0, st>>>(w.verts, nv, ox, oy, G, G, w.tog, w.orb, W0, W0, pw0);
  planes_to_image_kernel<<<nblk(W0, 64), 64, 0, st>>>(w.tog, w.orb, w.cone, W0, W0, pw0, 1, 1, nullptr);
  fog_masks_kernel<<<nblk(wn), 256, 0, st>>>(w.cone, d_nav, G, ox, oy, W0, w.blocked, w.visible);
  // ---- R2/R3/R4: obstacle contours -> rays -> cut
  contours(w, w.blocked, W0, W0, st, 0);
  fog_gate_kernel<<<1, 1, 0, st>>>(w.st);
  simple_vertices_kernel<<<EX_MAXC / 8, 256, 0, st>>>(w.cont, w.chain, w.sv, w.nsv, w.st);
  rays_kernel<<<EX_MAXC / 8, 256, 0, st>>>(w.cont, w.sv, w.nsv, w.rays, w.rays_cap, sx, sy, ox, oy, heading_deg, max_line_len * 1.05, w.st);
  clear_bytes_kernel<<<nblk(wn), 256, 0, st>>>(w.cut, wn);
  thick_rays_kernel<<<RAYS_CAP / 64, 64, 0, st>>>(w.rays, w.cut, W0, W0, ox, oy, G, w.st);
  apply_cut_kernel<<<nblk(wn), 256, 0, st>>>(w.visible, w.cut, wn, w.st);
  // ---- R5: contours of the visible area, nearest to the agent, filled
  contours(w, w.visible, W0, W0, st, 1);
  pick_nearest_kernel<<<EX_MAXC / 64, 64, 0, st>>>(w.cont, w.chain, w.dist, sx, sy, w.st, 0);
  pick_nearest_kernel<<<1, 1, 0, st>>>(w.cont, w.chain, w.dist, sx, sy, w.st, 1);
  fog_fill_gate_kernel<<<1, 1, 0, st>>>(w.st, w.which);
  clear_bytes_kernel<<<nblk(wn), 256, 0, st>>>(w.newexp, wn);
  zero_planes_kernel<<<nblk(pw0 * W0), 256, 0, st>>>(w.tog, w.orb, pw0 * W0);
  chain_edges_kernel<<<dim3(32, 1), 256, 0, st>>>(w.cont, w.which, w.chain, w.tog, w.orb, W0, W0, pw0, w.st);
  planes_to_image_kernel<<<nblk(W0, 64), 64, 0, st>>>(w.tog, w.orb, w.newexp, W0, W0, pw0, 1, 0, w.which + 0);
  // ---- explored |= dilate3(new); explored[nav == 0] = 0
  int half = nav_half > L + 5 ? nav_half : L + 5;
  int rx0 = agent_col - half, ry0 = agent_row - half, rx1 = agent_col + half + 1, ry1 = agent_row + half + 1;
  if (rx0 < 0) rx0 = 0; if (ry0 < 0) ry0 = 0; if (rx1 > G) rx1 = G; if (ry1 > G) ry1 = G;
  explored_update_kernel<<<nblk((long)(rx1 - rx0) * (ry1 - ry0)), 256, 0, st>>>(w.newexp, W0, ox, oy, d_explored, d_nav, G, rx0, ry0, rx1 - rx0,
                                                                              ry1 - ry0, w.st);
  // ---- component selection on the whole explored map (obstacle_map.py:128-146)
  contours(w, d_explored, G, G, st, 1, 1);
  select_component_kernel<<<EX_MAXC / 64, 64, 0, st>>>(w.cont, w.chain, w.dist, agent_col, agent_row, w.st, w.which, 0);
  select_component_kernel<<<1, 1, 0, st>>>(w.cont, w.chain, w.dist, agent_col, agent_row, w.st, w.which, 1);
  zero_planes_kernel<<<nblk(pwG * G), 256, 0, st>>>(w.tog, w.orb, pwG * G);
  chain_edges_kernel<<<dim3(64, 1), 256, 0, st>>>(w.cont, w.which, w.chain, w.tog, w.orb, G, G, pwG, w.st);
  planes_to_image_kernel<<<nblk(G, 64), 64, 0, st>>>(w.tog, w.orb, d_explored, G, G, pwG, 1, 1, w.which);   // no-op when a single contour exists (which == -1)
  // ---- frontiers (obstacle_map.py:155-169 -> detect_frontier_waypoints)
  dilate_full_kernel<<<nblk(n), 256, 0, st>>>(d_explored, w.grown, G, G, 5);
  unexplored_kernel<<<nblk(n), 256, 0, st>>>(d_nav, w.grown, w.unexp, n);
  copy_bytes_kernel<<<nblk(n), 256, 0, st>>>(w.grown, w.ex2, n);
  contours(w, w.unexp, G, G, st, 1, 2);
  absorb_small_kernel<<<EX_MAXC, 128, 0, st>>>(w.cont, w.chain, w.unexp, w.ex2, G, G, area_thresh_px, w.st);
  contours(w, w.ex2, G, G, st, 1);
  bad_flags_kernel<<<nblk(CHAIN_CAP / 64), 256, 0, st>>>(w.chain, d_nav, w.ex2, G, G, w.flags, w.st);
  frontier_kernel<<<1, 32, 0, st>>>(w.cont, w.chain, w.flags, d_frontiers, MAXF, w.st);
  rc = check_cuda(cudaMemcpyAsync(d_count, &w.st->n_front, sizeof(int), cudaMemcpyDeviceToDevice, st), "explore: count");
  if (rc) return rc;
  rc = check_cuda(cudaMemcpyAsync(d_status, &w.st->overflow, sizeof(int), cudaMemcpyDeviceToDevice, st), "explore: status");
  if (rc) return rc;
  VLFM_CHECK_LAUNCH("vlfm_explore_update");
  count_launch(40);
  return VLFM_OK;
}

extern "C" int vlfm_holes_workspace_bytes(int H, int W, size_t* bytes) {
  if (!bytes || H < 1 || W < 1) { set_error("vlfm_holes_workspace_bytes: bad argument"); return VLFM_E_INVALID; }
  int G = H > W ? H : W;
  *bytes = carve(nullptr, nullptr, G, HOLES_MAXC);
  return VLFM_OK;
}

// fill_small_holes (vlfm/utils/img_utils.py:361-390): d_filled[H,W] := 1 where the reference would write depth 1.0
extern "C" int vlfm_fill_small_holes(const float* d_depth, int H, int W, double area_thresh, uint8_t* d_filled, void* d_workspace,
                                     int32_t* d_status, void* stream) {
  if (!d_depth || !d_filled || !d_workspace || !d_status || H < 1 || W < 1) { set_error("vlfm_fill_small_holes: bad argument"); return VLFM_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  Ws w;
  carve(&w, (uint8_t*)d_workspace, H > W ? H : W, HOLES_MAXC);
  const int n = H * W;
  uint8_t* mask = w.unexp;                       // scratch planes of the explore workspace layout
  zero_mask_kernel<<<nblk(n), 256, 0, st>>>(d_depth, mask, n);
  clear_bytes_kernel<<<nblk(n), 256, 0, st>>>(d_filled, n);
  ccl_init2_kernel<<<nblk(H, 8), 256, 0, st>>>(mask, w.Lfg, w.Lbg, W, H, w.st, 0);
  ccl_merge2_kernel<<<nblk(n), 256, 0, st>>>(mask, w.Lfg, w.Lbg, w.outer, nullptr, w.nbm, W, H);
  ccl_flatten2_kernel<<<nblk(n), 256, 0, st>>>(w.Lfg, w.Lbg, n);
  bg_outer_kernel<<<nblk(2 * (W + H)), 256, 0, st>>>(w.Lbg, w.outer, W, H);
  collect_all_kernel<<<nblk(n), 256, 0, st>>>(w.Lfg, w.Lbg, w.outer, W, H, w.cont, HOLES_MAXC, w.st);
  trace_kernel<<<HOLES_MAXC / 64, 64, 0, st>>>(w.nbm, W, H, w.cont, w.chain, w.chain_cap, w.st, 0, w.hashole);
  static bool cfg = false;
  const int smem_words = 24 * 1024;             // 96 KB of toggle / outline bit planes per block
  if (!cfg) {
    int rc = check_cuda(cudaFuncSetAttribute(fill_small_contours_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_words * 4), "attr(fill_small_contours)");
    if (rc) return rc; cfg = true;
  }
  fill_small_contours_kernel<<<148 * 2, 256, smem_words * 4, st>>>(w.cont, w.chain, d_filled, W, H, area_thresh, smem_words, HOLES_MAXC, w.st);
  sticky_status_kernel<<<1, 1, 0, st>>>(w.st, d_status);      // sticky: the host may poll it many steps later
  VLFM_CHECK_LAUNCH("vlfm_fill_small_holes");
  count_launch(11);
  return VLFM_OK;
}

// Confidence-cone template of ValueMap (vlfm/mapping/value_map.py:321-355 `_get_confidence_mask` / `_get_blank_cone_mask`):
// cv2.ellipse filled sector (+-fov/2 about +row) x cos^2 falloff.  d_out [R,R] float32, R = 2*int(max_depth*ppm)+1;
// d_scratch: R*R bytes + 2*R*ceil(R/32) uint32 + 2 KB.
extern "C" int vlfm_value_cone_template(double fov, double max_depth, int ppm, double min_conf, float* d_out, void* d_scratch,
                                        size_t scratch_bytes, void* stream) {
  const int half = (int)(max_depth * ppm), R = 2 * half + 1, pw = (R + 31) / 32;
  const size_t need = (((size_t)R * R + 255) & ~(size_t)255) + (size_t)2 * R * pw * 4 + 2048;
  if (!d_out || !d_scratch || half < 1 || scratch_bytes < need) { set_error("vlfm_value_cone_template: bad argument (scratch %zu < %zu)", scratch_bytes, need); return VLFM_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* sector = (uint8_t*)d_scratch;
  uint32_t* tog = (uint32_t*)(sector + (((size_t)R * R + 255) & ~(size_t)255));
  uint32_t* orb = tog + (size_t)R * pw;
  long long* verts = (long long*)(orb + (size_t)R * pw);
  const double deg = fov * 180.0 / 3.14159265358979323846;       // np.rad2deg
  long long hv[128];
  const int nv = sector_polygon(half, half, half, -deg / 2 + 90, deg / 2 + 90, hv);
  int rc = check_cuda(cudaMemcpyAsync(verts, hv, sizeof(long long) * 2 * nv, cudaMemcpyHostToDevice, st), "cone template: vertex upload");
  if (rc) return rc;
  zero_planes_kernel<<<nblk((long)pw * R), 256, 0, st>>>(tog, orb, pw * R);
  sector_edges_kernel<<<1, 64, 0, st>>>(verts, nv, 0, 0, R, R, tog, orb, R, R, pw);
  planes_to_image_kernel<<<nblk(R, 64), 64, 0, st>>>(tog, orb, sector, R, R, pw, 1, 1, nullptr);
  cone_template_kernel<<<nblk((long)R * R), 256, 0, st>>>(sector, d_out, R, fov, min_conf);
  rc = check_cuda(cudaStreamSynchronize(st), "cone template");   // hv is a stack buffer: the upload must finish before returning
  if (rc) return rc;
  VLFM_CHECK_LAUNCH("vlfm_value_cone_template");
  count_launch(4);
  return VLFM_OK;
}
