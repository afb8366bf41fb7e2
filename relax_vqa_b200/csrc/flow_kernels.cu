// A5/A6/A7: Farneback dense optical flow (cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15,
// 3, 5, 1.2, 0)), flow colouring (flow_to_rgb) and the flow-fragment gather + merge.
// Algorithm restated from OpenCV's optflowgf.cpp as pinned by oracle/farneback.py.
// Data layout in HBM (per level, SoA so that every access is a coalesced float stream):
//   I   [2][B][h][w]      f32  blurred + resized images
//   R   [2][B][5][h][w]   f32  polynomial expansion coefficients (y, x, yy, xx, xy)
//   M                     structure-tensor entries: shared memory only (fused iteration kernel)
//   flow[B][h][w][2]      f32  (dx, dy), ping-pong between levels
#include <math.h>
#include <vector>
#include "context.h"

namespace b200vqa {

struct Taps { float t[19]; int ksize; };
struct PolyConsts { float g[6], xg[6], xxg[6]; float ig11, ig03, ig33, ig55; };   // index k = 0..5 (symmetric)

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// ---- (a+b) fused pyramid level: GaussianBlur(full-res uint8, ksize, sigma) sampled by cv::resize
// INTER_LINEAR at the level's pixel centres.  One block = 32 x 8 outputs; the source window is staged in
// shared memory (uint8), blurred horizontally only at the <= 64 columns the bilinear taps need (f32, smem),
// then blurred vertically at the <= 2 rows each output needs.  Row filter before column filter, like OpenCV.
constexpr int PY_TX = 32, PY_TY = 8;
__device__ __forceinline__ void lin_split(int d, double scale, int n_in, bool same, int& i0, int& i1, float& a) {
  if (same) { i0 = i1 = d; a = 0.f; return; }
  const double s = (d + 0.5) * scale - 0.5;
  int f = (int)floor(s);
  a = (float)(s - f);
  if (f < 0) { f = 0; a = 0.f; }
  i0 = f; i1 = f + 1;
  if (f >= n_in - 1) { i0 = i1 = n_in - 1; a = 0.f; }
}

__global__ void __launch_bounds__(256)
k4_pyr_level(const uint8_t* __restrict__ gray0, const uint8_t* __restrict__ gray1, int B, int H, int W, Taps taps, int h, int w,
             double sx, double sy, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t py_smem[];
  const int r = taps.ksize >> 1;
  const bool same = (h == H && w == W);
  const int z = blockIdx.z;
  const uint8_t* img = (z < B ? gray0 + (size_t)z * H * W : gray1 + (size_t)(z - B) * H * W);
  const int xo0 = blockIdx.x * PY_TX, yo0 = blockIdx.y * PY_TY;
  const int xo_last = min(xo0 + PY_TX, w) - 1, yo_last = min(yo0 + PY_TY, h) - 1;
  int i0, i1; float a;
  lin_split(xo0, sx, W, same, i0, i1, a);      const int x_lo = i0 - r;
  lin_split(xo_last, sx, W, same, i0, i1, a);  const int x_hi = i1 + r;
  lin_split(yo0, sy, H, same, i0, i1, a);      const int y_lo = i0 - r;
  lin_split(yo_last, sy, H, same, i0, i1, a);  const int y_hi = i1 + r;
  const int sw = x_hi - x_lo + 1, sh = y_hi - y_lo + 1;
  const int sw_p = (sw + 3) & ~3;
  uint8_t* src = py_smem;                                         // [sh][sw_p] uint8
  float* hs = reinterpret_cast<float*>(py_smem + (((size_t)sh * sw_p + 15) & ~(size_t)15));   // [sh][2*PY_TX] f32
  const int tid = threadIdx.x;
  for (int i = tid; i < sh * sw; i += 256) {
    const int ty = i / sw, tx = i - ty * sw;
    src[ty * sw_p + tx] = img[(size_t)reflect101(y_lo + ty, H) * W + reflect101(x_lo + tx, W)];
  }
  __syncthreads();
  // horizontal pass at the needed columns: column slot 2j / 2j+1 = left / right tap of output xo0 + j
  for (int i = tid; i < sh * 2 * PY_TX; i += 256) {
    const int ty = i / (2 * PY_TX), slot = i - ty * (2 * PY_TX);
    const int xo = min(xo0 + (slot >> 1), w - 1);
    lin_split(xo, sx, W, same, i0, i1, a);
    const int xc = ((slot & 1) ? i1 : i0) - r - x_lo;              // window start inside the tile
    const uint8_t* row = src + ty * sw_p + xc;
    float acc = 0.f;
    for (int k = 0; k < taps.ksize; ++k) acc += taps.t[k] * (float)row[k];
    hs[i] = acc;
  }
  __syncthreads();
  const int tx = tid & (PY_TX - 1), ty = tid / PY_TX;
  const int xo = xo0 + tx, yo = yo0 + ty;
  if (xo < w && yo < h) {
    float ax, ay; int xa, xb, ya, yb;
    lin_split(xo, sx, W, same, xa, xb, ax);
    lin_split(yo, sy, H, same, ya, yb, ay);
    const float* c0 = hs + (size_t)(ya - r - y_lo) * (2 * PY_TX) + 2 * tx;
    float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
    for (int k = 0; k < taps.ksize; ++k) { const float t = taps.t[k]; b00 += t * c0[k * 2 * PY_TX]; b01 += t * c0[k * 2 * PY_TX + 1]; }
    if (yb != ya) {
      const float* c1 = hs + (size_t)(yb - r - y_lo) * (2 * PY_TX) + 2 * tx;
      for (int k = 0; k < taps.ksize; ++k) { const float t = taps.t[k]; b10 += t * c1[k * 2 * PY_TX]; b11 += t * c1[k * 2 * PY_TX + 1]; }
    } else { b10 = b00; b11 = b01; }
    const float top = b00 * (1.f - ax) + b01 * ax, bot = b10 * (1.f - ax) + b11 * ax;
    out[((size_t)z * h + yo) * w + xo] = top * (1.f - ay) + bot * ay;
  }
}

// ---- (c) polynomial expansion: I [h][w] -> R planes.  Tile 32 x 16, halo 5, smem staged.
constexpr int PE_TW = 32, PE_TH = 16, PE_N = 5;
__global__ void __launch_bounds__(256)
k4_polyexp(const float* __restrict__ I, int h, int w, PolyConsts pc, float* __restrict__ R) {
  __shared__ float tile[PE_TH + 2 * PE_N][PE_TW + 2 * PE_N + 1];
  __shared__ float v0[PE_TH][PE_TW + 2 * PE_N + 1], v1[PE_TH][PE_TW + 2 * PE_N + 1], v2[PE_TH][PE_TW + 2 * PE_N + 1];
  const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH;
  const float* img = I + (size_t)blockIdx.z * h * w;
  const int tid = threadIdx.x;
  for (int i = tid; i < (PE_TH + 2 * PE_N) * (PE_TW + 2 * PE_N); i += 256) {
    const int ty = i / (PE_TW + 2 * PE_N), tx = i % (PE_TW + 2 * PE_N);
    const int gy = min(max(y0 + ty - PE_N, 0), h - 1), gx = min(max(x0 + tx - PE_N, 0), w - 1);
    tile[ty][tx] = img[(size_t)gy * w + gx];
  }
  __syncthreads();
  // vertical pass (f32): rows clamped -> the tile already holds clamped rows.  NB: row clamping of the
  // *source* index equals OpenCV's max(y-k,0)/min(y+k,h-1) because tile rows were clamped on load.
  for (int i = tid; i < PE_TH * (PE_TW + 2 * PE_N); i += 256) {
    const int ty = i / (PE_TW + 2 * PE_N), tx = i % (PE_TW + 2 * PE_N);
    const int c = ty + PE_N;
    float r0 = tile[c][tx] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 1; k <= PE_N; ++k) {
      const float up = tile[c - k][tx], dn = tile[c + k][tx];
      const float p = up + dn;
      r0 = r0 + pc.g[k] * p;
      r1 = r1 + pc.xg[k] * (dn - up);
      r2 = r2 + pc.xxg[k] * p;
    }
    v0[ty][tx] = r0; v1[ty][tx] = r1; v2[ty][tx] = r2;
  }
  __syncthreads();
  // horizontal pass.  OpenCV accumulates this pass in f64; f32 changes the flow by < 1e-5 px (measured, see
  // DESIGN.md) and avoids the quarter-rate f32->f64 conversions.  Columns beyond the image replicate the edge
  // column of the *vertical* result - which is what the clamped tile load produced.
  const size_t plane = (size_t)h * w;
  float* Rb = R + (size_t)blockIdx.z * 5 * plane;
  for (int i = tid; i < PE_TH * PE_TW; i += 256) {
    const int ty = i / PE_TW, tx = i % PE_TW;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= w || gy >= h) continue;
    const int c = tx + PE_N;
    float b1 = v0[ty][c] * pc.g[0], b2 = 0.f, b3 = v1[ty][c] * pc.g[0], b4 = 0.f, b5 = v2[ty][c] * pc.g[0], b6 = 0.f;
#pragma unroll
    for (int k = 1; k <= PE_N; ++k) {
      const float p0 = v0[ty][c + k], m0 = v0[ty][c - k], p1 = v1[ty][c + k], m1 = v1[ty][c - k], p2 = v2[ty][c + k], m2 = v2[ty][c - k];
      const float tg = p0 + m0;
      b1 += tg * pc.g[k];
      b4 += tg * pc.xxg[k];
      b2 += (p0 - m0) * pc.xg[k];
      b3 += (p1 + m1) * pc.g[k];
      b6 += (p1 - m1) * pc.xg[k];
      b5 += (p2 + m2) * pc.g[k];
    }
    const size_t o = (size_t)gy * w + gx;
    Rb[o] = b3 * pc.ig11;
    Rb[plane + o] = b2 * pc.ig11;
    Rb[2 * plane + o] = b1 * pc.ig03 + b5 * pc.ig33;
    Rb[3 * plane + o] = b1 * pc.ig03 + b4 * pc.ig33;
    Rb[4 * plane + o] = b6 * pc.ig55;
  }
}

// ---- (d+e) one Farneback iteration, fused: FarnebackUpdateMatrices on the 46 x 46 halo of a 32 x 32 tile
// (M never touches HBM), 15 x 15 box sums with replicate border, 2 x 2 solve.  flow_in -> flow_out.
__device__ __forceinline__ float border_w(int i, int n) {
  float s = 1.f;
  if (i < 5) s *= (i < 2 ? 0.14f : 0.4472f);
  const int j = n - 1 - i;
  if (j < 5) s *= (j < 2 ? 0.14f : 0.4472f);
  return s;
}

constexpr int BX_T = 32, BX_M = 7, BX_IN = BX_T + 2 * BX_M, BX_LD = BX_IN + 1;     // 46, 47
constexpr int BX_SMEM = (5 * BX_IN * BX_LD + 5 * BX_T * BX_LD) * 4;
__global__ void __launch_bounds__(256)
k4_flow_iter(const float* __restrict__ R0, const float* __restrict__ R1, const float* __restrict__ flow_in, int h, int w,
             float* __restrict__ flow_out) {
  extern __shared__ float bx_smem[];
  float* Ms = bx_smem;                               // [5][46][47]
  float* Vs = bx_smem + 5 * BX_IN * BX_LD;           // [5][32][47]
  const int x0 = blockIdx.x * BX_T, y0 = blockIdx.y * BX_T;
  const size_t plane = (size_t)h * w;
  const float* r0 = R0 + (size_t)blockIdx.z * 5 * plane;
  const float* r1 = R1 + (size_t)blockIdx.z * 5 * plane;
  const float2* fin = reinterpret_cast<const float2*>(flow_in) + (size_t)blockIdx.z * plane;
  const int tid = threadIdx.x;
  // phase A: structure-tensor entries at the (replicate-clamped) halo pixels
  for (int i = tid; i < BX_IN * BX_IN; i += 256) {
    const int ty = i / BX_IN, tx = i - ty * BX_IN;
    const int x = min(max(x0 + tx - BX_M, 0), w - 1), y = min(max(y0 + ty - BX_M, 0), h - 1);
    const size_t o = (size_t)y * w + x;
    const float2 d = fin[o];
    const float dx = d.x, dy = d.y;
    float fx = (float)x + dx, fy = (float)y + dy;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1; fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    const float a4 = r0[2 * plane + o], a5 = r0[3 * plane + o], a6 = r0[4 * plane + o];
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
      const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
      const size_t q = (size_t)y1 * w + x1;
#define BIL(c) (a00 * r1[(c) * plane + q] + a01 * r1[(c) * plane + q + 1] + a10 * r1[(c) * plane + q + w] + a11 * r1[(c) * plane + q + w + 1])
      r2 = BIL(0); r3 = BIL(1); r4 = BIL(2); r5 = BIL(3); r6 = BIL(4);
#undef BIL
      r4 = (a4 + r4) * 0.5f; r5 = (a5 + r5) * 0.5f; r6 = (a6 + r6) * 0.25f;
    } else {
      r2 = r3 = 0.f; r4 = a4; r5 = a5; r6 = a6 * 0.5f;
    }
    r2 = (r0[o] - r2) * 0.5f;
    r3 = (r0[plane + o] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
      const float s = border_w(y, h) * border_w(x, w);
      r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
    }
    float* m = Ms + ty * BX_LD + tx;
    m[0] = r4 * r4 + r6 * r6;
    m[BX_IN * BX_LD] = (r4 + r5) * r6;
    m[2 * BX_IN * BX_LD] = r5 * r5 + r6 * r6;
    m[3 * BX_IN * BX_LD] = r4 * r2 + r6 * r3;
    m[4 * BX_IN * BX_LD] = r6 * r2 + r5 * r3;
  }
  __syncthreads();
  // phase B: vertical sliding sums (f32; the walk is only 32 rows long, error ~ that of a direct 15-tap sum)
  if (tid < 5 * BX_IN) {
    const int c = tid / BX_IN, j = tid - c * BX_IN;
    const float* col = Ms + (size_t)c * BX_IN * BX_LD + j;
    float* dst = Vs + (size_t)c * BX_T * BX_LD + j;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 15; ++i) s += col[i * BX_LD];
    dst[0] = s;
    for (int rr = 1; rr < BX_T; ++rr) {
      s += col[(rr + 14) * BX_LD] - col[(rr - 1) * BX_LD];
      dst[rr * BX_LD] = s;
    }
  }
  __syncthreads();
  // phase C: horizontal sums over 4-column segments + solve (f64, as OpenCV): task = (row, segment)
  {
    const int rr = tid >> 3, seg = tid & 7;
    const int gy = y0 + rr;
    float g[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const float* row = Vs + ((size_t)c * BX_T + rr) * BX_LD + seg * 4;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 15; ++i) s += row[i];
      g[c] = s;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gx = x0 + seg * 4 + k;
      if (gx < w && gy < h) {
        const double sc = 1.0 / 225.0;
        const double g11 = g[0] * sc, g12 = g[1] * sc, g22 = g[2] * sc, h1 = g[3] * sc, h2 = g[4] * sc;
        const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
        float2 f;
        f.x = (float)((g11 * h2 - g12 * h1) * idet);
        f.y = (float)((g22 * h1 - g12 * h2) * idet);
        reinterpret_cast<float2*>(flow_out)[(size_t)blockIdx.z * plane + (size_t)gy * w + gx] = f;
      }
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const float* row = Vs + ((size_t)c * BX_T + rr) * BX_LD + seg * 4 + k;
        g[c] += row[15] - row[0];
      }
    }
  }
}

// ---- (f) flow upsample between levels: resize(prev, (w,h), INTER_LINEAR) * 2
__global__ void __launch_bounds__(256)
k4_flow_upsample(const float* __restrict__ prev, int hp, int wp, int h, int w, double sx, double sy, float* __restrict__ out) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y;
  if (xo >= w) return;
  const float2* p = reinterpret_cast<const float2*>(prev) + (size_t)blockIdx.z * hp * wp;
  double s = (xo + 0.5) * sx - 0.5; int f = (int)floor(s); float ax = (float)(s - f);
  int x0, x1, y0, y1;
  if (f < 0) { f = 0; ax = 0.f; } x0 = f; x1 = f + 1; if (f >= wp - 1) { x0 = x1 = wp - 1; ax = 0.f; }
  s = (yo + 0.5) * sy - 0.5; f = (int)floor(s); float ay = (float)(s - f);
  if (f < 0) { f = 0; ay = 0.f; } y0 = f; y1 = f + 1; if (f >= hp - 1) { y0 = y1 = hp - 1; ay = 0.f; }
  const float2 p00 = p[(size_t)y0 * wp + x0], p01 = p[(size_t)y0 * wp + x1], p10 = p[(size_t)y1 * wp + x0], p11 = p[(size_t)y1 * wp + x1];
  float2 o;
  o.x = ((p00.x * (1.f - ax) + p01.x * ax) * (1.f - ay) + (p10.x * (1.f - ax) + p11.x * ax) * ay) * 2.f;
  o.y = ((p00.y * (1.f - ax) + p01.y * ax) * (1.f - ay) + (p10.y * (1.f - ax) + p11.y * ax) * ay) * 2.f;
  reinterpret_cast<float2*>(out)[((size_t)blockIdx.z * h + yo) * w + xo] = o;
}

// =========================================================================== flow colouring
__device__ __forceinline__ float magnitude(float dx, float dy) {
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

__global__ void k5_minmax_init(float* minmax, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { minmax[2 * i] = __int_as_float(0x7f800000); minmax[2 * i + 1] = 0.f; }
}

__global__ void __launch_bounds__(256)
k5_mag_minmax(const float* __restrict__ flow, size_t npix, float* __restrict__ minmax) {
  const float2* f = reinterpret_cast<const float2*>(flow) + (size_t)blockIdx.y * npix;
  float mn = __int_as_float(0x7f800000), mx = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const float2 d = f[i];
    const float m = magnitude(d.x, d.y);
    mn = fminf(mn, m); mx = fmaxf(mx, m);
  }
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  __shared__ float smn[8], smx[8];
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    // magnitudes are >= 0, so the IEEE bit patterns order like unsigned integers
    atomicMin(reinterpret_cast<unsigned int*>(minmax + 2 * blockIdx.y), __float_as_uint(mn));
    atomicMax(reinterpret_cast<unsigned int*>(minmax + 2 * blockIdx.y + 1), __float_as_uint(mx));
  }
}

// cv::fastAtan2 polynomial (degrees), without FMA contraction, as pinned by the oracle
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = 57.28362274169922f, p3 = -18.66744613647461f, p5 = 8.914000511169434f, p7 = -2.539724588394165f;
  const float ax = fabsf(x), ay = fabsf(y);
  const float eps = 2.22044605e-16f;
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

// flow vector -> BGR uint8 triple, exactly flow_to_rgb (src/main_fragment_layerstack.py:162-175)
__device__ __forceinline__ uchar3 flow_colour(float dx, float dy, double nscale, double nshift) {
  const float mag = magnitude(dx, dy);
  const float ang = __fmul_rn(fast_atan2_deg(dy, dx), 0.01745329238474369f);             // radians (float32(pi/180))
  const float magn = (float)__dadd_rn(__dmul_rn((double)mag, nscale), nshift);    // cv2.normalize (double scale/shift)
  const float hue = __fmul_rn(__fdiv_rn(__fmul_rn(ang, 180.f), 3.1415927410125732f), 0.5f);
  const int H8 = (int)hue & 255, V8 = (int)fminf(fmaxf(magn, 0.f), 255.f);
  // cv2 HSV2BGR (8-bit, hrange 180): S = 255 -> s = 1
  const float hh = __fmul_rn((float)H8, 0.03333333507180214f);                           // float32(6/180)
  const float s = __fmul_rn(255.f, 0.003921568859368563f), v = __fmul_rn((float)V8, 0.003921568859368563f);   // float32(1/255)
  int sector = (int)floorf(hh);
  const float f = __fsub_rn(hh, (float)sector);
  sector %= 6;
  float tab[4];
  tab[0] = v;
  tab[1] = __fmul_rn(v, __fsub_rn(1.f, s));
  tab[2] = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, f)));
  tab[3] = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f))));
  const int ib[6] = {1, 1, 3, 0, 0, 2}, ig[6] = {3, 0, 0, 2, 1, 1}, ir[6] = {0, 2, 1, 1, 3, 0};
  const float b = __fmul_rn(tab[ib[sector]], 255.f), g = __fmul_rn(tab[ig[sector]], 255.f), r = __fmul_rn(tab[ir[sector]], 255.f);
  return make_uchar3((unsigned char)fminf(fmaxf(b, 0.f), 255.f), (unsigned char)fminf(fmaxf(g, 0.f), 255.f),
                     (unsigned char)fminf(fmaxf(r, 0.f), 255.f));
}

__device__ __forceinline__ void norm_consts(const float* minmax, double& sc, double& sh) {
  const double mn = minmax[0], mx = minmax[1];
  sc = (mx - mn > 2.220446049250313e-16) ? 255.0 / (mx - mn) : 0.0;
  sh = 0.0 - mn * sc;
}

// colour + 16x16 patch sums (+ optional image store).  block = 64 px x 16 rows (4 patches).
__global__ void __launch_bounds__(1024)
k5_flow_rgb_patchsum(const float* __restrict__ flow, int H, int W, const float* __restrict__ minmax, uint8_t* __restrict__ rgb,
                     uint32_t* __restrict__ sums) {
  __shared__ uint32_t part[4];
  const int gw = W >> 4, gh = H >> 4;
  if (threadIdx.y == 0 && threadIdx.x < 4) part[threadIdx.x] = 0;
  __syncthreads();
  double sc, sh;
  norm_consts(minmax + 2 * blockIdx.z, sc, sh);
  const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 16 + threadIdx.y;
  uint32_t s = 0;
  if (x < W && y < H) {
    const size_t p = ((size_t)blockIdx.z * H + y) * W + x;
    const float2 d = reinterpret_cast<const float2*>(flow)[p];
    const uchar3 c = flow_colour(d.x, d.y, sc, sh);
    if (rgb) { rgb[p * 3] = c.x; rgb[p * 3 + 1] = c.y; rgb[p * 3 + 2] = c.z; }
    s = (uint32_t)c.x + c.y + c.z;
  }
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);     // 16-lane groups = one patch row
  if (sums && (threadIdx.x & 15) == 0 && (x >> 4) < gw && blockIdx.y < gh && x < W && y < H) atomicAdd(&part[threadIdx.x >> 4], s);
  __syncthreads();
  if (sums && threadIdx.y == 0 && threadIdx.x < 4) {
    const int px = blockIdx.x * 4 + threadIdx.x;
    if (px < gw && blockIdx.y < gh) sums[((size_t)blockIdx.z * gh + blockIdx.y) * gw + px] = part[threadIdx.x];
  }
}

// gather flow-colour patches (recomputed from the flow) and merge with the diff fragment
__global__ void __launch_bounds__(256)
k5_flow_fragment_merge(const float* __restrict__ flow, const float* __restrict__ minmax, int H, int W,
                       const int32_t* __restrict__ pos, const int32_t* __restrict__ count, int top_n,
                       const uint8_t* __restrict__ diff_frag, uint8_t* __restrict__ flow_frag, uint8_t* __restrict__ merged) {
  const int j = blockIdx.x, b = blockIdx.y;
  const int cy = j / 14, cx = j % 14;
  const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
  uchar3 col = make_uchar3(0, 0, 0);
  if (j < count[b]) {
    const int py = pos[((size_t)b * top_n + j) * 2], px = pos[((size_t)b * top_n + j) * 2 + 1];
    double sc, sh;
    norm_consts(minmax + 2 * b, sc, sh);
    const float2 d = reinterpret_cast<const float2*>(flow)[((size_t)b * H + py * 16 + r) * W + px * 16 + c];
    col = flow_colour(d.x, d.y, sc, sh);
  }
  const size_t o = (((size_t)b * 224 + cy * 16 + r) * 224 + cx * 16 + c) * 3;
  if (flow_frag) { flow_frag[o] = col.x; flow_frag[o + 1] = col.y; flow_frag[o + 2] = col.z; }
  if (merged) {
    const uint32_t v[3] = {col.x, col.y, col.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint32_t s = (uint32_t)diff_frag[o + k] + v[k], hf = s >> 1;
      merged[o + k] = (uint8_t)(hf + ((s & 1) & (hf & 1)));
    }
  }
}

// ------------------------------------------------------------------------------ host side
static int cv_round(double x) { return (int)nearbyint(x); }   // default rounding mode = half to even

static Taps gaussian_taps(int ksize, double sigma) {
  Taps t{};
  t.ksize = ksize;
  if (sigma <= 0) { t.t[0] = 0.25f; t.t[1] = 0.5f; t.t[2] = 0.25f; return t; }
  double k[19], sum = 0;
  for (int i = 0; i < ksize; ++i) { double x = i - (ksize - 1) * 0.5; k[i] = exp(-(x * x) / (2.0 * sigma * sigma)); sum += k[i]; }
  for (int i = 0; i < ksize; ++i) t.t[i] = (float)(k[i] / sum);
  return t;
}

static bool invert6(double a[6][6], double inv[6][6]) {
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) inv[i][j] = i == j;
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (fabs(a[piv][c]) < 1e-300) return false;
    for (int j = 0; j < 6; ++j) { std::swap(a[c][j], a[piv][j]); std::swap(inv[c][j], inv[piv][j]); }
    const double d = a[c][c];
    for (int j = 0; j < 6; ++j) { a[c][j] /= d; inv[c][j] /= d; }
    for (int r = 0; r < 6; ++r) if (r != c) {
      const double f = a[r][c];
      for (int j = 0; j < 6; ++j) { a[r][j] -= f * a[c][j]; inv[r][j] -= f * inv[c][j]; }
    }
  }
  return true;
}

static PolyConsts poly_consts() {           // FarnebackPrepareGaussian(n = 5, sigma = 1.2)
  const int n = 5; const double sigma = 1.2;
  float g[11], xg[11], xxg[11];
  double s = 0;
  for (int x = -n; x <= n; ++x) { g[x + n] = (float)exp(-x * x / (2 * sigma * sigma)); s += g[x + n]; }
  s = 1. / s;
  for (int x = -n; x <= n; ++x) { g[x + n] = (float)(g[x + n] * s); xg[x + n] = (float)(x * g[x + n]); xxg[x + n] = (float)(x * x * g[x + n]); }
  double G[6][6] = {}, inv[6][6];
  for (int y = -n; y <= n; ++y) for (int x = -n; x <= n; ++x) {
    const double w = (double)g[y + n] * g[x + n];
    G[0][0] += w; G[1][1] += w * x * x; G[3][3] += w * x * x * x * x; G[5][5] += w * x * x * y * y;
  }
  G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
  G[4][4] = G[3][3];
  G[3][4] = G[4][3] = G[5][5];
  invert6(G, inv);
  PolyConsts pc;
  for (int k = 0; k <= n; ++k) { pc.g[k] = g[n + k]; pc.xg[k] = xg[n + k]; pc.xxg[k] = xxg[n + k]; }
  pc.ig11 = (float)inv[1][1]; pc.ig03 = (float)inv[0][3]; pc.ig33 = (float)inv[3][3]; pc.ig55 = (float)inv[5][5];
  return pc;
}

struct Level { double scale, sigma; int ksize, w, h; };

static std::vector<Level> pyramid_plan(int H, int W) {
  int levels = 0; double scale = 1.0;
  while (levels < 3) { scale *= 0.5; if (W * scale < 32 || H * scale < 32) break; ++levels; }
  std::vector<Level> plan;
  for (int k = levels; k >= 0; --k) {
    Level L; L.scale = pow(0.5, k); L.sigma = (1.0 / L.scale - 1.0) * 0.5;
    L.ksize = cv_round(L.sigma * 5) | 1; if (L.ksize < 3) L.ksize = 3;
    L.w = cv_round(W * L.scale); L.h = cv_round(H * L.scale);
    plan.push_back(L);
  }
  return plan;
}

}  // namespace b200vqa

using namespace b200vqa;

extern "C" int b200vqa_farneback(b200vqa_t* h, const uint8_t* gray0, const uint8_t* gray1, int B, int H, int W, float* flow,
                                 void* stream) {
  if (!h || !gray0 || !gray1 || !flow || B <= 0 || H < 16 || W < 16) return B200VQA_EINVAL;
  CtxScope scope(h);
  cudaStream_t st = as_stream(stream);
  const size_t P = (size_t)H * W;
  // workspace (float): I [2][B][P] | R [2][B][5][P] | flowA, flowB [B][P][2]
  const std::vector<Level> plan = pyramid_plan(H, W);
  const size_t floats = (size_t)B * P * (2 + 10 + 4);
  int rc = h->ws_flow.reserve(floats * sizeof(float));
  if (rc) return rc;
  float* I = static_cast<float*>(h->ws_flow.ptr);
  float* R = I + 2 * (size_t)B * P;
  float* flowA = R + 10 * (size_t)B * P;
  float* flowB = flowA + 2 * (size_t)B * P;
  static const PolyConsts pc = poly_consts();
  static bool attr_done = false;
  if (!attr_done) {
    VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, BX_SMEM));
    VQA_CUDA(cudaFuncSetAttribute(k4_pyr_level, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_done = true;
  }
  float* prev = nullptr;         // flow of the previous (coarser) level
  int ph = 0, pw = 0;
  for (size_t li = 0; li < plan.size(); ++li) {
    const Level& L = plan[li];
    const bool last = li + 1 == plan.size();
    const Taps taps = gaussian_taps(L.ksize, L.sigma);
    const size_t lp = (size_t)L.h * L.w;
    float* R0 = R; float* R1 = R + (size_t)B * 5 * lp;
    {
      const double sx = (double)W / L.w, sy = (double)H / L.h;
      // shared-memory window: uint8 source tile + f32 horizontally blurred columns
      const int sw = (int)(PY_TX * sx) + L.ksize + 4, sh = (int)(PY_TY * sy) + L.ksize + 4;
      const size_t smem = (((size_t)sh * ((sw + 3) & ~3) + 15) & ~(size_t)15) + (size_t)sh * 2 * PY_TX * sizeof(float);
      if (smem > 100 * 1024) return B200VQA_EINVAL;
      k4_pyr_level<<<dim3(cdiv(L.w, PY_TX), cdiv(L.h, PY_TY), 2 * B), 256, smem, st>>>(gray0, gray1, B, H, W, taps, L.h, L.w, sx, sy, I);
      VQA_LAUNCH_CHECK();
      // I holds [2B][h][w]; R0 = expansion of images 0..B-1, R1 of B..2B-1 (contiguous)
      k4_polyexp<<<dim3(cdiv(L.w, PE_TW), cdiv(L.h, PE_TH), 2 * B), 256, 0, st>>>(I, L.h, L.w, pc, R0);
      VQA_LAUNCH_CHECK();
    }
    float* fin = flowA;
    if (!prev) {
      VQA_CUDA(cudaMemsetAsync(fin, 0, (size_t)B * lp * 2 * sizeof(float), st));
      count_launch();
    } else {
      fin = (prev == flowA) ? flowB : flowA;
      k4_flow_upsample<<<dim3(cdiv(L.w, 256), L.h, B), 256, 0, st>>>(prev, ph, pw, L.h, L.w, (double)pw / L.w, (double)ph / L.h, fin);
      VQA_LAUNCH_CHECK();
    }
    const dim3 gbox(cdiv(L.w, BX_T), cdiv(L.h, BX_T), B);
    float* fout = nullptr;
    for (int it = 0; it < 3; ++it) {
      fout = (last && it == 2) ? flow : (fin == flowA ? flowB : flowA);
      k4_flow_iter<<<gbox, 256, BX_SMEM, st>>>(R0, R1, fin, L.h, L.w, fout);
      VQA_LAUNCH_CHECK();
      fin = fout;
    }
    prev = fout; ph = L.h; pw = L.w;
  }
  return B200VQA_OK;
}

extern "C" int b200vqa_flow_to_rgb(const float* flow, int B, int H, int W, uint8_t* rgb, uint32_t* sums, float* minmax,
                                   void* stream) {
  if (!flow || !minmax || B <= 0 || H <= 0 || W <= 0) return B200VQA_EINVAL;
  cudaStream_t st = as_stream(stream);
  k5_minmax_init<<<cdiv(B, 128), 128, 0, st>>>(minmax, B);
  VQA_LAUNCH_CHECK();
  const size_t npix = (size_t)H * W;
  int gx = (int)((npix + 256 * 8 - 1) / (256 * 8));
  if (gx > 592) gx = 592;
  k5_mag_minmax<<<dim3(gx, B), 256, 0, st>>>(flow, npix, minmax);
  VQA_LAUNCH_CHECK();
  if (rgb || sums) {
    k5_flow_rgb_patchsum<<<dim3(cdiv(W, 64), cdiv(H, 16), B), dim3(64, 16), 0, st>>>(flow, H, W, minmax, rgb, sums);
    VQA_LAUNCH_CHECK();
  }
  return B200VQA_OK;
}

extern "C" int b200vqa_flow_fragment_merge(const float* flow, const float* minmax, int B, int H, int W, const int32_t* pos,
                                           const int32_t* count, int top_n, const uint8_t* diff_frag, uint8_t* flow_frag,
                                           uint8_t* merged_frag, void* stream) {
  if (!flow || !minmax || !pos || !count || B <= 0 || top_n != B200VQA_TOPN || (merged_frag && !diff_frag)) return B200VQA_EINVAL;
  k5_flow_fragment_merge<<<dim3(top_n, B), 256, 0, as_stream(stream)>>>(flow, minmax, H, W, pos, count, top_n, diff_frag, flow_frag, merged_frag);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}
