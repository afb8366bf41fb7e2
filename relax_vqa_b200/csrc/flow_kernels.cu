// A5/A6/A7: Farneback dense optical flow (cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15,
// 3, 5, 1.2, 0)), flow colouring (flow_to_rgb) and the flow-fragment gather + merge.
// Algorithm restated from OpenCV's optflowgf.cpp as pinned by oracle/farneback.py.
// Data layout in HBM (per level, SoA so that every access is a coalesced float stream):
//   I   [2][B][h][w]      f32  blurred + resized images
//   RA  [2B][h][w] float4 / RB [2B][h][w] float : polynomial expansion coefficients (y, x, yy, xx | xy)
//   M                     structure-tensor entries: shared memory only (fused iteration kernel)
//   flow[B][h][w][2]      f32  (dx, dy), ping-pong between levels
#include <math.h>
#include <stdlib.h>
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
constexpr int PY_TX = 32;
// exact uint8 -> float without the quarter-rate I2F unit: 2^23 + b has b in its low mantissa bits
__device__ __forceinline__ float u8_to_float(uint8_t b) { return __uint_as_float(0x4B000000u | (uint32_t)b) - 8388608.0f; }
__device__ __forceinline__ void lin_split(int d, double scale, int n_in, bool same, int& i0, int& i1, float& a) {
  if (same) { i0 = i1 = d; a = 0.f; return; }
  const double s = (d + 0.5) * scale - 0.5;
  int f = (int)floor(s);
  a = (float)(s - f);
  if (f < 0) { f = 0; a = 0.f; }
  i0 = f; i1 = f + 1;
  if (f >= n_in - 1) { i0 = i1 = n_in - 1; a = 0.f; }
}

// KS = Gaussian kernel size known at compile time (3, 9, 19 for pyramid levels 1..3) so the tap loops unroll with
// the taps in registers; KS = 0 is the generic run-time-size fallback (unusual resolutions).  TY = output rows per
// block (32 / 16 / 8 for the three levels: the finer the level, the smaller its source window per output, so more
// outputs per block amortise the per-block setup).  The bilinear source coordinates of the tile's 32 columns and TY
// rows are computed once (in double, like cv::resize) into shared tables.
template <int KS, int TY>
__global__ void __launch_bounds__(256)
k4_pyr_level(const uint8_t* __restrict__ gray0, const uint8_t* __restrict__ gray1, int B, int H, int W, Taps taps, int h, int w,
             double sx, double sy, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t py_smem[];
  __shared__ int s_xa[PY_TX], s_xb[PY_TX], s_ya[TY], s_yb[TY];
  __shared__ float s_ax[PY_TX], s_ay[TY];
  const int ks = KS ? KS : taps.ksize;
  const int r = ks >> 1;
  float tp[KS ? KS : 1];
  if (KS) {
#pragma unroll
    for (int k = 0; k < (KS ? KS : 1); ++k) tp[k] = taps.t[k];
  }
  const bool same = (h == H && w == W);
  const int z = blockIdx.z;
  const uint8_t* img = (z < B ? gray0 + (size_t)z * H * W : gray1 + (size_t)(z - B) * H * W);
  const int xo0 = blockIdx.x * PY_TX, yo0 = blockIdx.y * TY;
  const int tid = threadIdx.x;
  const int wrp = tid >> 5, lane = tid & 31;
  if (tid < PY_TX) { int i0, i1; float a; lin_split(min(xo0 + tid, w - 1), sx, W, same, i0, i1, a); s_xa[tid] = i0; s_xb[tid] = i1; s_ax[tid] = a; }
  else if (tid < PY_TX + TY) { const int j = tid - PY_TX; int i0, i1; float a; lin_split(min(yo0 + j, h - 1), sy, H, same, i0, i1, a); s_ya[j] = i0; s_yb[j] = i1; s_ay[j] = a; }
  __syncthreads();
  int x_lo = s_xa[0] - r;
  const int x_hi = s_xb[PY_TX - 1] + r, y_lo = s_ya[0] - r, y_hi = s_yb[TY - 1] + r;
  // interior tiles whose rows are 4-byte addressable are staged with 32-bit loads (window start aligned down)
  const bool vec = (W & 3) == 0 && (((uintptr_t)img & 3) == 0) && y_lo >= 0 && y_hi < H && (x_lo & ~3) >= 0 &&
                   (x_lo & ~3) + ((x_hi - (x_lo & ~3) + 4) & ~3) <= W;
  if (vec) x_lo &= ~3;
  const int sw = x_hi - x_lo + 1, sh = y_hi - y_lo + 1;
  const int sw_p = (sw + 3) & ~3;
  uint8_t* src = py_smem;                                         // [sh][sw_p] uint8
  float* hs = reinterpret_cast<float*>(py_smem + (((size_t)sh * sw_p + 15) & ~(size_t)15));   // [sh][2*PY_TX] f32
  if (vec) {
    const int wpr = sw_p >> 2;                                     // 32-bit words per row
    for (int i = tid; i < sh * wpr; i += 256) {
      const int ty = i / wpr, tx = i - ty * wpr;
      reinterpret_cast<uint32_t*>(src + ty * sw_p)[tx] = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)(y_lo + ty) * W + x_lo) + tx);
    }
  } else {
    for (int ty = wrp; ty < sh; ty += 8) {
      const uint8_t* grow = img + (size_t)reflect101(y_lo + ty, H) * W;
      for (int tx = lane; tx < sw; tx += 32) src[ty * sw_p + tx] = grow[reflect101(x_lo + tx, W)];
    }
  }
  __syncthreads();
  // horizontal pass at the needed columns: column slot 2j / 2j+1 = left / right tap of output xo0 + j
  int xc2[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int slot = lane + 32 * u;
    xc2[u] = ((slot & 1) ? s_xb[slot >> 1] : s_xa[slot >> 1]) - r - x_lo;      // window start inside the tile
  }
  for (int ty = wrp; ty < sh; ty += 8) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint8_t* row = src + ty * sw_p + xc2[u];
      float acc = 0.f;
      if (KS) {
#pragma unroll
        for (int k = 0; k < (KS ? KS : 1); ++k) acc += tp[k] * u8_to_float(row[k]);
      } else {
        for (int k = 0; k < ks; ++k) acc += taps.t[k] * u8_to_float(row[k]);
      }
      hs[ty * (2 * PY_TX) + lane + 32 * u] = acc;
    }
  }
  __syncthreads();
  const int tx = lane;
  const int xo = xo0 + tx;
  const float ax = s_ax[tx];
  for (int ty = wrp; ty < TY; ty += 8) {
    const int yo = yo0 + ty;
    if (xo < w && yo < h) {
      const float ay = s_ay[ty];
      const float* c0 = hs + (size_t)(s_ya[ty] - r - y_lo) * (2 * PY_TX) + 2 * tx;
      const float* c1 = hs + (size_t)(s_yb[ty] - r - y_lo) * (2 * PY_TX) + 2 * tx;
      float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
      if (KS) {
#pragma unroll
        for (int k = 0; k < (KS ? KS : 1); ++k) {
          const float2 p0 = *reinterpret_cast<const float2*>(c0 + k * 2 * PY_TX), p1 = *reinterpret_cast<const float2*>(c1 + k * 2 * PY_TX);
          b00 += tp[k] * p0.x; b01 += tp[k] * p0.y; b10 += tp[k] * p1.x; b11 += tp[k] * p1.y;
        }
      } else {
        for (int k = 0; k < ks; ++k) {
          const float t = taps.t[k];
          b00 += t * c0[k * 2 * PY_TX]; b01 += t * c0[k * 2 * PY_TX + 1]; b10 += t * c1[k * 2 * PY_TX]; b11 += t * c1[k * 2 * PY_TX + 1];
        }
      }
      const float top = b00 * (1.f - ax) + b01 * ax, bot = b10 * (1.f - ax) + b11 * ax;
      out[((size_t)z * h + yo) * w + xo] = top * (1.f - ay) + bot * ay;
    }
  }
}

// ---- (a+b), fused streaming form for the exact power-of-two levels (W % 8 == 0 and H, W divisible by the level's scale
// S = 2, 4, 8: 1080p, 2160p, 720p ...).  There the bilinear taps of cv::resize sit at columns / rows S x + S/2 - 1 and + 1
// with weights 0.5 / 0.5.  One thread owns 8 source columns (= 4 / 2 / 1 output columns of levels 1 / 2 / 3) and marches
// down the source rows of a segment: per row it loads its 24-byte window once (3 x LDG.64, coalesced across the warp),
// runs the horizontal Gaussian at the left and right tap column of each of its outputs, for all three levels, and adds
// the two values into the (at most three) pending output rows per level whose vertical windows contain the row, with
// compile-time tap indices.  The arithmetic is k4_pyr_level's, operation for operation (taps in ascending order from a
// zero accumulator, the four Gaussian values of an output kept apart until the bilinear combination), so the two kernels
// agree bit for bit - which matters: on static scenes the flow near the image border is decided by the SIGN of a ~1e-6 px
// displacement (in / out of bounds in UpdateMatrices), and a variant with folded taps (half the arithmetic, values equal
// to 3e-7 relative) moved single border pixels by 0.06 px away from OpenCV.
// The full-resolution frames are read once for the three levels; the tile kernel reads them once per level, stages them
// through shared memory byte by byte and runs at a tenth of its HBM roofline (461 us per 22 pairs of 1080p; this: see
// profiles/).
struct PyrFusedTaps { float t1[3], t2[9], t3[19]; };
constexpr int PF_NT = 128;                         // threads per block: 4 warps x 32 column groups x 8 source pixels

// pending output rows of one level.  S = stride, KS = Gaussian size, NO = 8 / S output columns per thread.
// v[o][slot][.] = Gaussian at (upper row, left col), (upper, right), (lower, left), (lower, right) of the bilinear cell
template <int S, int KS, int NO>
struct PyrAcc {
  float v[NO][3][4];                               // slots: output rows m - 1, m, m + 1 (m = current period)
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
      for (int sl = 0; sl < 3; ++sl) v[o][sl][0] = v[o][sl][1] = v[o][sl][2] = v[o][sl][3] = 0.f;
  }
  // source row with phase I (= row index mod S, compile-time after unrolling); hl / hr = horizontal Gaussian at the left /
  // right tap column of each output
  __device__ __forceinline__ void add(int I, const float (&hl)[NO], const float (&hr)[NO], const float* T) {
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
      const int k0 = I - (S / 2 - 1) + (KS >> 1) - S * d, k1 = k0 - 1;       // tap index in the upper / lower row's window
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        if (k0 >= 0 && k0 < KS) { v[o][d + 1][0] = fmaf(T[k0], hl[o], v[o][d + 1][0]); v[o][d + 1][1] = fmaf(T[k0], hr[o], v[o][d + 1][1]); }
        if (k1 >= 0 && k1 < KS) { v[o][d + 1][2] = fmaf(T[k1], hl[o], v[o][d + 1][2]); v[o][d + 1][3] = fmaf(T[k1], hr[o], v[o][d + 1][3]); }
      }
    }
  }
  static __device__ __forceinline__ bool completes(int I) { return I - (S / 2 - 1) + (KS >> 1) + S == KS; }   // row m - 1 got its last tap
  __device__ __forceinline__ float result(int o) const {                   // cv::resize with both weights 0.5
    const float top = fmaf(v[o][0][0], 0.5f, v[o][0][1] * 0.5f), bot = fmaf(v[o][0][2], 0.5f, v[o][0][3] * 0.5f);
    return fmaf(top, 0.5f, bot * 0.5f);
  }
  __device__ __forceinline__ void rotate() {
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
      for (int c = 0; c < 4; ++c) { v[o][0][c] = v[o][1][c]; v[o][1][c] = v[o][2][c]; v[o][2][c] = 0.f; }
  }
};

__device__ __forceinline__ uint32_t u2_byte(const uint2& v, int j) { return ((j < 4 ? v.x : v.y) >> (8 * (j & 3))) & 0xffu; }

// horizontal Gaussian of KS taps starting at window position p0 (k4_pyr_level's order: ascending, from a zero accumulator)
template <int KS>
__device__ __forceinline__ float pyr_hsum(const float* T, const float (&f)[24], int p0) {
  float acc = T[0] * f[p0];
#pragma unroll
  for (int k = 1; k < KS; ++k) acc = fmaf(T[k], f[p0 + k], acc);
  return acc;
}

template <int MASK>                                 // bit l - 1 set: level l is produced by this instantiation
__global__ void __launch_bounds__(PF_NT)
k4_pyr_fused(const uint8_t* __restrict__ gray0, const uint8_t* __restrict__ gray1, int B, int H, int W, PyrFusedTaps tp, int os3,
             float* __restrict__ I1, float* __restrict__ I2, float* __restrict__ I3) {
  const int G = W >> 3;                                              // column groups of 8 source pixels
  const int g = blockIdx.x * PF_NT + threadIdx.x;
  if (g >= G) return;
  const int z = blockIdx.z;
  const uint8_t* img = (z < B ? gray0 + (size_t)z * H * W : gray1 + (size_t)(z - B) * H * W);
  const int a3 = blockIdx.y * os3, b3 = min(a3 + os3, (H + 7) >> 3);   // this segment's output rows in level-3 units (8 source rows)
  const int gA = max(g - 1, 0), gC = min(g + 1, G - 1);
  PyrAcc<2, 3, 4> L1; PyrAcc<4, 9, 2> L2; PyrAcc<8, 19, 1> L3;
  L1.clear(); L2.clear(); L3.clear();
  const int h1 = H >> 1, w1 = W >> 1, h2 = H >> 2, w2 = W >> 2, h3 = H >> 3, w3 = W >> 3;
  for (int M = a3 - 1; M <= b3; ++M) {                               // periods of 8 source rows; M - 1 = the level-3 row that completes in it
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int y = reflect101(8 * M + i, H);
      const uint2* row = reinterpret_cast<const uint2*>(img + (size_t)y * W);
      uint2 wa = __ldg(row + gA), wb = __ldg(row + g), wc = __ldg(row + gC);
      if (g == 0) {                                                  // x = -j -> j (REFLECT_101): positions 0..7 of the window
        uint2 n = make_uint2(0u, 0u);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t bb = q == 0 ? u2_byte(wc, 0) : u2_byte(wb, 8 - q);
          if (q < 4) n.x |= bb << (8 * q); else n.y |= bb << (8 * (q - 4));
        }
        wa = n;
      }
      if (g == G - 1) {                                              // x = W + j -> W - 2 - j: positions 16..23
        uint2 n = make_uint2(0u, 0u);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t bb = q == 7 ? u2_byte(wa, 7) : u2_byte(wb, 6 - q);
          if (q < 4) n.x |= bb << (8 * q); else n.y |= bb << (8 * (q - 4));
        }
        wc = n;
      }
      // bytes 2..22 of the window as floats (2^23 + b through one byte permute, minus 2^23)
      float f[24];
      const uint32_t wd[6] = {wa.x, wa.y, wb.x, wb.y, wc.x, wc.y};
#pragma unroll
      for (int pz = 2; pz <= 22; ++pz)
        f[pz] = __uint_as_float(__byte_perm(wd[pz >> 2], 0x4B000000u, 0x7650u + (pz & 3))) - 8388608.0f;
      if (MASK & 1) {                                                // outputs 4 g + o: left tap column 8 g + 2 o, window from 2 o - 1
        float hl[4], hr[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) { hl[o] = pyr_hsum<3>(tp.t1, f, 7 + 2 * o); hr[o] = pyr_hsum<3>(tp.t1, f, 8 + 2 * o); }
        L1.add(i & 1, hl, hr, tp.t1);
        if (PyrAcc<2, 3, 4>::completes(i & 1)) {
          const int yo = (4 * M + (i >> 1)) - 1;
          if (yo >= 4 * a3 && yo < min(4 * b3, h1))
            *reinterpret_cast<float4*>(I1 + ((size_t)z * h1 + yo) * w1 + 4 * g) = make_float4(L1.result(0), L1.result(1), L1.result(2), L1.result(3));
        }
        if ((i & 1) == 1) L1.rotate();
      }
      if (MASK & 2) {                                                // outputs 2 g + o: left tap column 8 g + 4 o + 1, window from 4 o - 3
        float hl[2], hr[2];
#pragma unroll
        for (int o = 0; o < 2; ++o) { hl[o] = pyr_hsum<9>(tp.t2, f, 5 + 4 * o); hr[o] = pyr_hsum<9>(tp.t2, f, 6 + 4 * o); }
        L2.add(i & 3, hl, hr, tp.t2);
        if (PyrAcc<4, 9, 2>::completes(i & 3)) {
          const int yo = (2 * M + (i >> 2)) - 1;
          if (yo >= 2 * a3 && yo < min(2 * b3, h2))
            *reinterpret_cast<float2*>(I2 + ((size_t)z * h2 + yo) * w2 + 2 * g) = make_float2(L2.result(0), L2.result(1));
        }
        if ((i & 3) == 3) L2.rotate();
      }
      if (MASK & 4) {                                                // output g: left tap column 8 g + 3, window from -6
        float hl[1], hr[1];
        hl[0] = pyr_hsum<19>(tp.t3, f, 2); hr[0] = pyr_hsum<19>(tp.t3, f, 3);
        L3.add(i, hl, hr, tp.t3);
        if (PyrAcc<8, 19, 1>::completes(i)) {
          const int yo = M - 1;
          if (yo >= a3 && yo < min(b3, h3)) I3[((size_t)z * h3 + yo) * w3 + g] = L3.result(0);
        }
        if (i == 7) L3.rotate();
      }
    }
  }
}

// ---- (c) polynomial expansion: I -> R.  Tile 64 x 16 outputs, halo 5; each thread produces 4 consecutive
// values per pass from a 14-value register window (3x fewer shared-memory loads than one value per thread).
// R layout: RA [img][h][w] float4 = (y, x, yy, xx) coefficients, RB [img][h][w] float = xy coefficient.
// kFromGray (finest level): I = GaussianBlur(gray, 3x3, taps .25 .5 .25, REFLECT_101) is computed on the fly
// from the uint8 image, so the full-resolution f32 image never exists in HBM.
constexpr int PE_TW = 64, PE_TH = 16, PE_N = 5, PE_LW = PE_TW + 2 * PE_N, PE_LH = PE_TH + 2 * PE_N;   // 74, 26
template <bool kFromGray>
__global__ void __launch_bounds__(256)
k4_polyexp(const float* __restrict__ I, const uint8_t* __restrict__ gray0, const uint8_t* __restrict__ gray1, int B, int h, int w,
           PolyConsts pc, float4* __restrict__ RA, float* __restrict__ RB) {
  __shared__ float tile[PE_LH][PE_LW + 1];
  __shared__ __align__(16) float v0[PE_TH][PE_LW + 2], v1[PE_TH][PE_LW + 2], v2[PE_TH][PE_LW + 2];   // stride 76: float4 rows
  __shared__ int hrow[kFromGray ? PE_LH + 2 : 1][kFromGray ? PE_LW + 3 : 1];   // 4 x row-filtered gray, exact integers
  const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH, z = blockIdx.z;
  const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;
  if (kFromGray) {
    const uint8_t* g = (z < B ? gray0 + (size_t)z * h * w : gray1 + (size_t)(z - B) * h * w);
    // row-filtered gray at raw coordinates (y0-6+ty, x0-5+tx): rows 0..27, cols 0..73
    const bool interior = x0 - PE_N - 1 >= 0 && x0 + PE_TW + PE_N + 1 <= w && y0 - 6 >= 0 && y0 + PE_TH + 6 <= h;
    if (interior && (w & 3) == 0 && (((uintptr_t)g & 3) == 0)) {      // no clamping / reflection needed
      // stage the 28 x 80-byte gray window with independent 32-bit loads (all in flight at once), then filter from smem
      __shared__ uint32_t graw[PE_LH + 2][20];
      const uint8_t* g0 = g + (size_t)(y0 - 6) * w + (x0 - 8);         // x0 is a multiple of 64 -> 4-byte aligned
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int i = tid + it * 256;
        if (i < (PE_LH + 2) * 20) { const int r = i / 20, c = i - r * 20; graw[r][c] = __ldg(reinterpret_cast<const uint32_t*>(g0 + (size_t)r * w) + c); }
      }
      __syncthreads();
      for (int ty = wrp; ty < PE_LH + 2; ty += 8) {
        const uint8_t* grow = reinterpret_cast<const uint8_t*>(graw[ty]) + 2;      // byte of global column x0 - 6
#pragma unroll
        for (int it = 0; it < 3; ++it) {
          const int tx = lane + 32 * it;
          if (tx < PE_LW) hrow[ty][tx] = (int)grow[tx] + 2 * (int)grow[tx + 1] + (int)grow[tx + 2];
        }
      }
    } else {
      for (int ty = wrp; ty < PE_LH + 2; ty += 8) {
        const uint8_t* grow = g + (size_t)reflect101(min(max(y0 - 6 + ty, -1), h), h) * w;
        for (int tx = lane; tx < PE_LW; tx += 32) {
          const int cx = min(max(x0 - PE_N + tx, 0), w - 1);          // I is replicated outside the image
          hrow[ty][tx] = (int)grow[reflect101(cx - 1, w)] + 2 * (int)grow[cx] + (int)grow[reflect101(cx + 1, w)];
        }
      }
    }
    __syncthreads();
    for (int ty = wrp; ty < PE_LH; ty += 8) {
      const int cy = min(max(y0 - PE_N + ty, 0), h - 1);
      const int r = cy - (y0 - 6);                                      // hrow row of image row cy
      // (.25 .5 .25) x (.25 .5 .25) of 8-bit integers is exact in fp32, so integer sums / 16 are bit-identical to OpenCV's float passes
      for (int tx = lane; tx < PE_LW; tx += 32) tile[ty][tx] = (float)(hrow[r - 1][tx] + 2 * hrow[r][tx] + hrow[r + 1][tx]) * 0.0625f;
    }
  } else {
    const float* img = I + (size_t)z * h * w;
#pragma unroll
    for (int it = 0; it < (PE_LH * PE_LW + 255) / 256; ++it) {        // independent loads, all in flight at once
      const int i = tid + it * 256;
      if (i < PE_LH * PE_LW) {
        const int ty = i / PE_LW, tx = i - ty * PE_LW;
        tile[ty][tx] = img[(size_t)min(max(y0 + ty - PE_N, 0), h - 1) * w + min(max(x0 + tx - PE_N, 0), w - 1)];
      }
    }
  }
  __syncthreads();
  // vertical pass: task = (column, group of 4 rows); 74 * 4 = 296 tasks
  for (int task = tid; task < PE_LW * (PE_TH / 4); task += 256) {
    const int tx = task % PE_LW, rg = task / PE_LW;
    float win[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) win[i] = tile[rg * 4 + i][tx];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float r0 = win[o + 5] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int k = 1; k <= PE_N; ++k) {
        const float up = win[o + 5 - k], dn = win[o + 5 + k], p = up + dn;
        r0 = r0 + pc.g[k] * p;
        r1 = r1 + pc.xg[k] * (dn - up);
        r2 = r2 + pc.xxg[k] * p;
      }
      v0[rg * 4 + o][tx] = r0; v1[rg * 4 + o][tx] = r1; v2[rg * 4 + o][tx] = r2;
    }
  }
  __syncthreads();
  // horizontal pass (f32; OpenCV uses f64 accumulators here - the flow changes by < 1e-5 px):
  // task = (row, group of 4 columns) = 16 * 16 = 256
  {
    const int ty = tid >> 4, cg = tid & 15;
    const int gy = y0 + ty;
    float a0[16], a1[16], a2[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 q0 = *reinterpret_cast<const float4*>(&v0[ty][cg * 4 + 4 * i]);
      const float4 q1 = *reinterpret_cast<const float4*>(&v1[ty][cg * 4 + 4 * i]);
      const float4 q2 = *reinterpret_cast<const float4*>(&v2[ty][cg * 4 + 4 * i]);
      a0[4 * i] = q0.x; a0[4 * i + 1] = q0.y; a0[4 * i + 2] = q0.z; a0[4 * i + 3] = q0.w;
      a1[4 * i] = q1.x; a1[4 * i + 1] = q1.y; a1[4 * i + 2] = q1.z; a1[4 * i + 3] = q1.w;
      a2[4 * i] = q2.x; a2[4 * i + 1] = q2.y; a2[4 * i + 2] = q2.z; a2[4 * i + 3] = q2.w;
    }
    const size_t plane = (size_t)h * w;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int gx = x0 + cg * 4 + o;
      const int c = o + 5;
      float b1 = a0[c] * pc.g[0], b2 = 0.f, b3 = a1[c] * pc.g[0], b4 = 0.f, b5 = a2[c] * pc.g[0], b6 = 0.f;
#pragma unroll
      for (int k = 1; k <= PE_N; ++k) {
        const float tg = a0[c + k] + a0[c - k];
        b1 += tg * pc.g[k];
        b4 += tg * pc.xxg[k];
        b2 += (a0[c + k] - a0[c - k]) * pc.xg[k];
        b3 += (a1[c + k] + a1[c - k]) * pc.g[k];
        b6 += (a1[c + k] - a1[c - k]) * pc.xg[k];
        b5 += (a2[c + k] + a2[c - k]) * pc.g[k];
      }
      if (gx < w && gy < h) {
        const size_t oo = (size_t)z * plane + (size_t)gy * w + gx;
        RA[oo] = make_float4(b3 * pc.ig11, b2 * pc.ig11, b1 * pc.ig03 + b5 * pc.ig33, b1 * pc.ig03 + b4 * pc.ig33);
        RB[oo] = b6 * pc.ig55;
      }
    }
  }
}

// ---- (c), streaming form: one block owns a strip of PM_SX output columns (+ 5 halo columns each side) and marches
// down a segment of rows, one thread per column with the 11 most recent rows of I in registers, so the vertical
// 11-tap pass reads every input exactly once (the 64 x 16 tile kernel above re-reads 26/16 rows and re-filters
// 74/64 columns).  Every PM_RB rows the three vertically filtered rows go through shared memory to the horizontal
// pass (4 adjacent outputs per thread from aligned 128-bit windows).  Same arithmetic, in the same order, as
// k4_polyexp: results are bit-identical.
constexpr int PM_NT = 256, PM_SX = 244, PM_RB = 4;
template <bool kFromGray>
__global__ void __launch_bounds__(PM_NT, 3)
k4_polyexp_march(const float* __restrict__ I, const uint8_t* __restrict__ gray0, const uint8_t* __restrict__ gray1, int B, int h, int w,
                 int rows_per_seg, PolyConsts pc, float4* __restrict__ RA, float* __restrict__ RB) {
  __shared__ __align__(16) float2 vb[PM_RB / 2][3][PM_NT];            // (row 2j, row 2j + 1) interleaved per column
  const int tx = threadIdx.x, z = blockIdx.z;
  const int x0 = blockIdx.x * PM_SX;
  const int ya = blockIdx.y * rows_per_seg, yb = min(ya + rows_per_seg, h);
  const int cx = min(max(x0 - PE_N + tx, 0), w - 1);                 // I is replicated outside the image
  // input stage, fetched one iteration (two rows) ahead of its use.  Gray path: the 3x3 pre-blur (.25 .5 .25)^2 with
  // REFLECT_101 is kept as exact integers: hrow(r) = g[r][x-1] + 2 g[r][x] + g[r][x+1]; moving down one row shifts
  // (ha, hb, hc) = hrow(reflect(cy-1)), hrow(cy), hrow(reflect(cy+1)) and needs only the new third row.
  const float* img = nullptr;
  const uint8_t *gl = nullptr, *gc = nullptr, *gr = nullptr;
  if (kFromGray) {
    const uint8_t* g = (z < B ? gray0 + (size_t)z * h * w : gray1 + (size_t)(z - B) * h * w);
    gl = g + reflect101(cx - 1, w); gc = g + cx; gr = g + reflect101(cx + 1, w);
  } else {
    img = I + (size_t)z * h * w + cx;
  }
  auto hrow = [&](int r) -> int { const int o = r * w; return (int)gl[o] + 2 * (int)gc[o] + (int)gr[o]; };
  // look-ahead rows are kept as the three raw bytes: summing them where they are loaded made every load wait for itself
  // (ncu: 34 % of the stall samples on that line); they are summed one iteration later, BEFORE the next loads are issued
  // (ptxas tracks all global loads of the kernel on one scoreboard, so a consumer placed after newer loads waits for those too)
  auto hraw = [&](int r, uint8_t (&b)[3]) { const int o = r * w; b[0] = gl[o]; b[1] = gc[o]; b[2] = gr[o]; };
  auto crow = [&](int k) -> int { return min(max(ya - PE_N + k, 0), h - 1); };        // image row of march step k
  float win[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) win[i] = 0.f;
  int ha = 0, hb = 0, hc = 0, prev_cy = crow(0);
  uint8_t hraw2[2][3] = {{0, 0, 0}, {0, 0, 0}};   // raw inputs of march steps k, k + 1 (loaded one iteration ahead)
  float fn[4] = {0.f, 0.f, 0.f, 0.f};
  if (kFromGray) {
    ha = hrow(reflect101(prev_cy - 1, h)); hb = hrow(prev_cy); hc = hrow(reflect101(prev_cy + 1, h));
#pragma unroll
    for (int u = 0; u < 2; ++u) hraw(reflect101(crow(u) + 1, h), hraw2[u]);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) fn[u] = img[(size_t)crow(u) * w];
  }
  const int nk = 10 + ((yb - ya + PM_RB - 1) & ~(PM_RB - 1));
  for (int k = 0; k < nk; k += 2) {
    float in2[2];
    int hcur[2] = {0, 0};
    if (kFromGray) {
#pragma unroll
      for (int u = 0; u < 2; ++u) hcur[u] = (int)hraw2[u][0] + 2 * (int)hraw2[u][1] + (int)hraw2[u][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) hraw(reflect101(crow(k + 2 + u) + 1, h), hraw2[u]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int cy = crow(k + u);
        if (cy != prev_cy) { ha = hb; hb = hc; hc = hcur[u]; prev_cy = cy; }
        in2[u] = (float)(ha + 2 * hb + hc) * 0.0625f;     // integer sums / 16 are exactly OpenCV's two float passes
      }
    } else {
#pragma unroll
      for (int u = 0; u < 2; ++u) { in2[u] = fn[u]; fn[u] = fn[2 + u]; fn[2 + u] = img[(size_t)crow(k + 4 + u) * w]; }
    }
    float vr[2][3];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int i = 0; i < 10; ++i) win[i] = win[i + 1];
      win[10] = in2[u];
      float r0 = win[5] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int t = 1; t <= PE_N; ++t) {
        const float up = win[5 - t], dn = win[5 + t], p = up + dn;
        r0 = r0 + pc.g[t] * p;
        r1 = r1 + pc.xg[t] * (dn - up);
        r2 = r2 + pc.xxg[t] * p;
      }
      vr[u][0] = r0; vr[u][1] = r1; vr[u][2] = r2;
    }
    if (k >= 10) {                                  // k is even: output rows k - 10 and k - 9 form one interleaved pair
      const int jp = ((k - 10) >> 1) & (PM_RB / 2 - 1);
#pragma unroll
      for (int a = 0; a < 3; ++a) vb[jp][a][tx] = make_float2(vr[0][a], vr[1][a]);
    }
    if (k >= 12 && (k & 3) == 0) {
      // output rows ya + k - 12 .. ya + k - 9 are filtered vertically: horizontal pass, task = (row, 4 columns)
      __syncthreads();
      if (tx < (PM_RB / 2) * (PM_SX / 2)) {
        // task = (row pair, column pair): both rows run through the packed fp32x2 pipe (FFMA2 / FADD2 on sm_100) with the
        // tap as the broadcast operand; per lane this is the scalar sequence of k4_polyexp, so results are unchanged
        const int jp = tx / (PM_SX / 2), cp = tx - jp * (PM_SX / 2);
        const int gy = ya + k - 12 + 2 * jp;
        const float2 neg1 = make_float2(-1.f, -1.f), zero2 = make_float2(0.f, 0.f);
        float2 b1[2], b2[2], b3[2], b4[2], b5[2], b6[2];
        {
          float2 a[12];
#pragma unroll
          for (int i = 0; i < 6; ++i) { const float4 q = *reinterpret_cast<const float4*>(&vb[jp][0][cp * 2 + 2 * i]); a[2 * i] = make_float2(q.x, q.y); a[2 * i + 1] = make_float2(q.z, q.w); }
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const int c = o + 5;
            b1[o] = __fmul2_rn(a[c], make_float2(pc.g[0], pc.g[0])); b2[o] = zero2; b4[o] = zero2;
#pragma unroll
            for (int t = 1; t <= PE_N; ++t) {
              const float2 tg = __fadd2_rn(a[c + t], a[c - t]);
              b1[o] = __ffma2_rn(tg, make_float2(pc.g[t], pc.g[t]), b1[o]);
              b4[o] = __ffma2_rn(tg, make_float2(pc.xxg[t], pc.xxg[t]), b4[o]);
              b2[o] = __ffma2_rn(__ffma2_rn(a[c - t], neg1, a[c + t]), make_float2(pc.xg[t], pc.xg[t]), b2[o]);
            }
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) { const float4 q = *reinterpret_cast<const float4*>(&vb[jp][1][cp * 2 + 2 * i]); a[2 * i] = make_float2(q.x, q.y); a[2 * i + 1] = make_float2(q.z, q.w); }
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const int c = o + 5;
            b3[o] = __fmul2_rn(a[c], make_float2(pc.g[0], pc.g[0])); b6[o] = zero2;
#pragma unroll
            for (int t = 1; t <= PE_N; ++t) {
              b3[o] = __ffma2_rn(__fadd2_rn(a[c + t], a[c - t]), make_float2(pc.g[t], pc.g[t]), b3[o]);
              b6[o] = __ffma2_rn(__ffma2_rn(a[c - t], neg1, a[c + t]), make_float2(pc.xg[t], pc.xg[t]), b6[o]);
            }
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) { const float4 q = *reinterpret_cast<const float4*>(&vb[jp][2][cp * 2 + 2 * i]); a[2 * i] = make_float2(q.x, q.y); a[2 * i + 1] = make_float2(q.z, q.w); }
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const int c = o + 5;
            b5[o] = __fmul2_rn(a[c], make_float2(pc.g[0], pc.g[0]));
#pragma unroll
            for (int t = 1; t <= PE_N; ++t) b5[o] = __ffma2_rn(__fadd2_rn(a[c + t], a[c - t]), make_float2(pc.g[t], pc.g[t]), b5[o]);
          }
        }
        const int gx0 = x0 + cp * 2;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          if (gy + rr < yb) {
            const size_t row = (size_t)z * h * w + (size_t)(gy + rr) * w;
            float xy[2];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              const float v1 = rr ? b1[o].y : b1[o].x, v2 = rr ? b2[o].y : b2[o].x, v3 = rr ? b3[o].y : b3[o].x;
              const float v4 = rr ? b4[o].y : b4[o].x, v5 = rr ? b5[o].y : b5[o].x, v6 = rr ? b6[o].y : b6[o].x;
              xy[o] = v6 * pc.ig55;
              if (gx0 + o < w) RA[row + gx0 + o] = make_float4(v3 * pc.ig11, v2 * pc.ig11, v1 * pc.ig03 + v5 * pc.ig33, v1 * pc.ig03 + v4 * pc.ig33);
            }
            if (gx0 + 1 < w && ((row + gx0) & 1) == 0 && (reinterpret_cast<uintptr_t>(RB) & 7) == 0) {
              *reinterpret_cast<float2*>(RB + row + gx0) = make_float2(xy[0], xy[1]);
            } else {
#pragma unroll
              for (int o = 0; o < 2; ++o) if (gx0 + o < w) RB[row + gx0 + o] = xy[o];
            }
          }
        }
      }
      __syncthreads();
    }
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

// 48 x 32 outputs per block: the 62-pixel halo row is covered by exactly two warp-wide passes (64 lanes)
constexpr int BX_TX = 48, BX_TY = 32, BX_M = 7, BX_W = BX_TX + 2 * BX_M, BX_H = BX_TY + 2 * BX_M, BX_LD = 65;   // 62, 46
constexpr int BX_SEG = BX_TX / 8;                                                                            // columns per warp in phase C
constexpr int BX_SMEM = 5 * BX_H * BX_LD * 4;
__global__ void __launch_bounds__(256, 3)
k4_flow_iter(const float4* __restrict__ RA0, const float* __restrict__ RB0, const float4* __restrict__ RA1,
             const float* __restrict__ RB1, const float* __restrict__ flow_in, int h, int w, float* __restrict__ flow_out) {
  extern __shared__ float Ms[];                      // [5][46][65]; rows 0..31 become the vertical sums in place
  const int x0 = blockIdx.x * BX_TX, y0 = blockIdx.y * BX_TY;
  const size_t plane = (size_t)h * w, zo = (size_t)blockIdx.z * plane;
  const float2* fin = reinterpret_cast<const float2*>(flow_in) + zo;
  const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;
  // phase A: structure-tensor entries at the (replicate-clamped) halo pixels; one warp per halo row, two pixels
  // per lane processed together so that their dependent gathers (flow -> address -> R1 taps) overlap.  Branch-free:
  // out-of-image taps are clamped and blended out, lanes 62/63 of the second pass recompute column 61.
  const float4* ra0 = RA0 + zo; const float* rb0 = RB0 + zo;
  const float4* ra1 = RA1 + zo; const float* rb1 = RB1 + zo;
  // the flow vector heads the dependent chain (flow -> tap address -> R1 taps): fetch it one row ahead
  int xs[2], txs[2];
  float2 dn[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    txs[u] = min(lane + 32 * u, BX_W - 1);
    xs[u] = min(max(x0 + txs[u] - BX_M, 0), w - 1);
    dn[u] = fin[min(max(y0 + wrp - BX_M, 0), h - 1) * w + xs[u]];
  }
  for (int ty = wrp; ty < BX_H; ty += 8) {
    const int y = min(max(y0 + ty - BX_M, 0), h - 1);
    const int yo = y * w;
    const int yno = min(max(y0 + min(ty + 8, BX_H - 1) - BX_M, 0), h - 1) * w;
    float2 d[2]; float4 c0[2]; float c0xy[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int o = yo + xs[u];
      d[u] = dn[u]; c0[u] = ra0[o]; c0xy[u] = rb0[o];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) dn[u] = fin[yno + xs[u]];
    float fx[2], fy[2], inside[2];
    float4 p00[2], p01[2], p10[2], p11[2]; float s00[2], s01[2], s10[2], s11[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float gx = (float)xs[u] + d[u].x, gy = (float)y + d[u].y;
      const float flx = floorf(gx), fly = floorf(gy);
      fx[u] = gx - flx; fy[u] = gy - fly;
      const int x1 = (int)flx, y1 = (int)fly;
      inside[u] = ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) ? 1.f : 0.f;
      const int q = min(max(y1, 0), h - 2) * w + min(max(x1, 0), w - 2);
      p00[u] = ra1[q]; p01[u] = ra1[q + 1]; p10[u] = ra1[q + w]; p11[u] = ra1[q + w + 1];
      s00[u] = rb1[q]; s01[u] = rb1[q + 1]; s10[u] = rb1[q + w]; s11[u] = rb1[q + w + 1];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int x = xs[u];
      const float dx = d[u].x, dy = d[u].y;
      const float a00 = (1.f - fx[u]) * (1.f - fy[u]), a01 = fx[u] * (1.f - fy[u]), a10 = (1.f - fx[u]) * fy[u], a11 = fx[u] * fy[u];
      float r2 = a00 * p00[u].x + a01 * p01[u].x + a10 * p10[u].x + a11 * p11[u].x;
      float r3 = a00 * p00[u].y + a01 * p01[u].y + a10 * p10[u].y + a11 * p11[u].y;
      float r4 = a00 * p00[u].z + a01 * p01[u].z + a10 * p10[u].z + a11 * p11[u].z;
      float r5 = a00 * p00[u].w + a01 * p01[u].w + a10 * p10[u].w + a11 * p11[u].w;
      float r6 = a00 * s00[u] + a01 * s01[u] + a10 * s10[u] + a11 * s11[u];
      const bool in = inside[u] != 0.f;
      // inside: r4 = (R0.yy + r4)/2, r5 = (R0.xx + r5)/2, r6 = (R0.xy + r6)/4; outside: r2 = r3 = 0, r4 = R0.yy, r5 = R0.xx, r6 = R0.xy/2
      r2 = in ? r2 : 0.f; r3 = in ? r3 : 0.f;
      r4 = in ? (c0[u].z + r4) * 0.5f : c0[u].z;
      r5 = in ? (c0[u].w + r5) * 0.5f : c0[u].w;
      r6 = in ? (c0xy[u] + r6) * 0.25f : c0xy[u] * 0.5f;
      r2 = (c0[u].x - r2) * 0.5f;
      r3 = (c0[u].y - r3) * 0.5f;
      r2 += r4 * dy + r6 * dx;
      r3 += r6 * dy + r5 * dx;
      if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float s = border_w(y, h) * border_w(x, w);
        r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
      }
      float* m = Ms + ty * BX_LD + txs[u];
      m[0] = r4 * r4 + r6 * r6;
      m[BX_H * BX_LD] = (r4 + r5) * r6;
      m[2 * BX_H * BX_LD] = r5 * r5 + r6 * r6;
      m[3 * BX_H * BX_LD] = r4 * r2 + r6 * r3;
      m[4 * BX_H * BX_LD] = r6 * r2 + r5 * r3;
    }
  }
  __syncthreads();
  // phase B: vertical sliding sums, in place (f32; the walk is 32 rows, error ~ a direct 15-tap sum).
  // The sum for output row r is written over input row r one step late, after that row's last use.
  for (int task = tid; task < 5 * BX_W; task += 256) {
    const int c = task / BX_W, j = task - c * BX_W;
    float* col = Ms + (size_t)c * BX_H * BX_LD + j;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 15; ++i) s += col[i * BX_LD];
    float prev = s;
    for (int rr = 1; rr < BX_TY; ++rr) {
      const float add = col[(rr + 14) * BX_LD], sub = col[(rr - 1) * BX_LD];
      col[(rr - 1) * BX_LD] = prev;
      s += add - sub;
      prev = s;
    }
    col[(BX_TY - 1) * BX_LD] = prev;
  }
  __syncthreads();
  // phase C: horizontal sums over BX_SEG-column segments + solve (f64 solve, as OpenCV): warp = segment, lane = row
  {
    const int rr = lane, seg = wrp;
    const int gy = y0 + rr;
    float g[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const float* row = Ms + ((size_t)c * BX_H + rr) * BX_LD + seg * BX_SEG;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 15; ++i) s += row[i];
      g[c] = s;
    }
#pragma unroll
    for (int k = 0; k < BX_SEG; ++k) {
      const int gx = x0 + seg * BX_SEG + k;
      if (gx < w && gy < h) {
        const double sc = 1.0 / 225.0;
        const double g11 = g[0] * sc, g12 = g[1] * sc, g22 = g[2] * sc, h1 = g[3] * sc, h2 = g[4] * sc;
        const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
        float2 f;
        f.x = (float)((g11 * h2 - g12 * h1) * idet);
        f.y = (float)((g22 * h1 - g12 * h2) * idet);
        reinterpret_cast<float2*>(flow_out)[zo + (size_t)gy * w + gx] = f;
      }
      if (k < BX_SEG - 1) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const float* row = Ms + ((size_t)c * BX_H + rr) * BX_LD + seg * BX_SEG + k;
          g[c] += row[15] - row[0];
        }
      }
    }
  }
}

// ---- (d+e), streaming form: one block owns a strip of MS_SX output columns (+ 8 halo columns each side, so the
// strip starts on a 128-byte boundary of the float4 planes) and marches down a segment of rows, one thread per
// column.  Per row a thread evaluates FarnebackUpdateMatrices at its pixel, slides the 15-row vertical box sum
// (running sums in registers, the 15 previous rows of M in a shared-memory ring that only this thread touches), and
// every MS_RB rows the block turns the vertical sums into 15-column sums with aligned 128-bit shared-memory windows
// (4 adjacent outputs per thread) and solves.  Halo recompute: 256/240 columns x (rows + 14)/rows instead of the
// 1.92x of the 48 x 32 tile kernel; moving down a column the upper two bilinear taps of a row are the lower two of
// the previous row whenever the displacement is locally smooth, so they are reused from registers.
constexpr int MS_HALO = 8, MS_RB = 4;
// NT threads = NT - 16 output columns per strip.  256 threads x 2 CTAs / SM (97,280 B each) or 192 x 3 (72,960 B each: 18 warps / SM
// instead of 16, strip efficiency 176/192 instead of 240/256)
constexpr int ms_sx(int nt) { return nt - 2 * MS_HALO; }
constexpr int ms_smem(int nt) { return (15 * 5 * nt + MS_RB * 5 * nt) * 4; }          // ring + row-batch buffer
struct Tap2 { float4 a0, a1; float b0, b1; };                              // columns x1, x1 + 1 of one R1 row

__device__ __forceinline__ Tap2 ld_tap2(const float4* __restrict__ ra, const float* __restrict__ rb, int q) {
  Tap2 t; t.a0 = ra[q]; t.a1 = ra[q + 1]; t.b0 = rb[q]; t.b1 = rb[q + 1]; return t;
}

// cv2.cartToPolar's magnitude: sqrt(fma(x, x, fl(y*y))) (bit-identical to cv2 4.13: oracle/farneback.py::cv_magnitude)
__device__ __forceinline__ float magnitude(float dx, float dy) {
  return __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// cv::resize(prev, (w, h), INTER_LINEAR) * 2 between pyramid levels: source cell and weight of one output coordinate (computed
// in double, as OpenCV does) and the bilinear combination, in the association k4_flow_upsample has always compiled to.
struct UpCoord { int i0, i1; float a; };
__device__ __forceinline__ UpCoord up_coord(int o, double scale, int n_in) {
  const double s = (o + 0.5) * scale - 0.5;
  int f = (int)floor(s);
  UpCoord c; c.a = (float)(s - f);
  if (f < 0) { f = 0; c.a = 0.f; }
  c.i0 = f; c.i1 = f + 1;
  if (f >= n_in - 1) { c.i0 = c.i1 = n_in - 1; c.a = 0.f; }
  return c;
}
__device__ __forceinline__ float up_mix(float p00, float p01, float p10, float p11, float ax, float ay) {
  const float omx = __fsub_rn(1.f, ax);
  const float h0 = __fmaf_rn(p00, omx, __fmul_rn(p01, ax)), h1 = __fmaf_rn(p10, omx, __fmul_rn(p11, ax));
  const float v = __fmaf_rn(h0, __fsub_rn(1.f, ay), __fmul_rn(ay, h1));
  return __fadd_rn(v, v);
}
__device__ __forceinline__ float2 up_sample(const float2* __restrict__ p, int wp, const UpCoord& cy, const UpCoord& cx) {
  const float2 p00 = p[(size_t)cy.i0 * wp + cx.i0], p01 = p[(size_t)cy.i0 * wp + cx.i1];
  const float2 p10 = p[(size_t)cy.i1 * wp + cx.i0], p11 = p[(size_t)cy.i1 * wp + cx.i1];
  return make_float2(up_mix(p00.x, p01.x, p10.x, p11.x, cx.a, cy.a), up_mix(p00.y, p01.y, p10.y, p11.y, cx.a, cy.a));
}
struct UpSrc { const float* prev; int hp, wp; double sx, sy; };     // the coarser level's flow (kUp)

// kMinMax: the launch that writes the final flow also reduces min / max of its magnitude per image into minmax[z][2]
// (what flow_to_rgb's first normalisation needs): the colouring then skips its own pass over the flow.
// kUp: the first iteration of a level reads its input flow straight from the coarser level (the bilinear x 2 upsample of
// k4_flow_upsample evaluated where the flow is consumed: the level's upsampled field is never written to or read from HBM);
// the rows' source coordinates sit in a shared-memory table built once per block, the column's in registers.
template <bool kDoubleSums, bool kPrefetch, int MS_NT, bool kMinMax = false, bool kUp = false>
__global__ void __launch_bounds__(MS_NT, MS_NT == 256 ? 2 : 3)
k4_flow_iter_march(const float4* __restrict__ RA0, const float* __restrict__ RB0, const float4* __restrict__ RA1,
                   const float* __restrict__ RB1, const float* __restrict__ flow_in, int h, int w, int rows_per_seg,
                   float* __restrict__ flow_out, float* __restrict__ minmax = nullptr, UpSrc up = UpSrc{}) {
  constexpr int MS_SX = MS_NT - 2 * MS_HALO;
  float mag_mn = __int_as_float(0x7f800000), mag_mx = 0.f;
  extern __shared__ __align__(16) float ms_smem_buf[];
  float* ring = ms_smem_buf;                               // [15][5][MS_NT]
  float* hb = ms_smem_buf + 15 * 5 * MS_NT;                // [MS_RB][5][MS_NT] vertical sums of the current row batch
  UpCoord* uprow = reinterpret_cast<UpCoord*>(hb + MS_RB * 5 * MS_NT);      // kUp: [nk + 4] source rows of march steps 0 ..
  const int tx = threadIdx.x;
  const int x0 = blockIdx.x * MS_SX;
  const int ya = blockIdx.y * rows_per_seg, yb = min(ya + rows_per_seg, h);
  const size_t plane = (size_t)h * w, zo = (size_t)blockIdx.z * plane;
  const float2* fin = reinterpret_cast<const float2*>(flow_in) + zo;
  const float4* ra0 = RA0 + zo; const float* rb0 = RB0 + zo;
  const float4* ra1 = RA1 + zo; const float* rb1 = RB1 + zo;
  const int x = min(max(x0 - MS_HALO + tx, 0), w - 1);                      // replicate border: M at the clamped pixel
  const bool xedge = (unsigned)(x - 5) >= (unsigned)(w - 10);
  const float bwx = border_w(x, w);
  const int nk = 14 + ((yb - ya + MS_RB - 1) & ~(MS_RB - 1));               // rows of M this block walks (even)
  // vertical running sums.  kDoubleSums: f64 accumulators, as OpenCV's own sliding sums (FarnebackUpdateFlow_Blur) - a strong
  // edge that passed through the 15-row window leaves no residue in the sums of the flat region below it, whatever the
  // dynamic range (tests/test_flow_streaming_model.py); same instruction count as the compensated fp32 sums (vs + comp),
  // the conversions and adds run on the conversion / fp64 pipes beside the fp32 work.
  double vd[5];
  float vs[5], comp[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) { vd[c] = 0.0; vs[c] = 0.f; comp[c] = 0.f; }
  int slot = 0;
  // flow vectors head the dependent chain flow -> tap address -> taps: they are fetched two iterations ahead, and one
  // iteration ahead their tap lines (and the R0 lines) are requested into L2, so the demand loads of an iteration find
  // an L2 hit instead of a DRAM access (kPrefetch)
  UpCoord upx{0, 0, 0.f};
  const float2* upp = nullptr;
  if (kUp) {
    for (int i = tx; i < nk + 4; i += MS_NT) uprow[i] = up_coord(min(max(ya - 7 + i, 0), h - 1), up.sy, up.hp);
    upx = up_coord(x, up.sx, up.wp);
    upp = reinterpret_cast<const float2*>(up.prev) + (size_t)blockIdx.z * up.hp * up.wp;
    __syncthreads();
  }
  // input flow of march step kidx (row clamp(ya - 7 + kidx)) at this thread's column
  auto flow_at = [&](int kidx) -> float2 {
    if (kUp) return up_sample(upp, up.wp, uprow[kidx], upx);
    return fin[min(max(ya - 7 + kidx, 0), h - 1) * w + x];
  };
  float2 dn[2], dnn[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    dn[u] = flow_at(u);
    dnn[u] = flow_at(2 + u);
  }
  Tap2 pbot; pbot.a0 = pbot.a1 = make_float4(0.f, 0.f, 0.f, 0.f); pbot.b0 = pbot.b1 = 0.f;
  int pq = -1;                                                              // address of the row held in pbot
  for (int k = 0; k < nk; k += 2) {
    int y[2], q[2]; float2 d[2]; float4 c0[2]; float c0xy[2]; float fx[2], fy[2]; bool in[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      y[u] = min(max(ya - 7 + k + u, 0), h - 1);
      const int o = y[u] * w + x;
      d[u] = dn[u]; c0[u] = ra0[o]; c0xy[u] = rb0[o];
      const float gxf = (float)x + d[u].x, gyf = (float)y[u] + d[u].y;
      const float flx = floorf(gxf), fly = floorf(gyf);
      fx[u] = gxf - flx; fy[u] = gyf - fly;
      const int x1 = (int)flx, y1 = (int)fly;
      in[u] = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
      q[u] = min(max(y1, 0), h - 2) * w + min(max(x1, 0), w - 2);
    }
    // lower tap rows are always fetched; upper rows only where they are not the previous row's lower taps
    Tap2 bot[2], topl[2]; bool need[2];
    need[0] = q[0] != pq;
    need[1] = q[1] != q[0] + w;
#pragma unroll
    for (int u = 0; u < 2; ++u) bot[u] = ld_tap2(ra1, rb1, q[u] + w);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      topl[u].a0 = topl[u].a1 = make_float4(0.f, 0.f, 0.f, 0.f); topl[u].b0 = topl[u].b1 = 0.f;
      if (need[u]) topl[u] = ld_tap2(ra1, rb1, q[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      dn[u] = dnn[u];
      const int yn = min(max(ya - 7 + k + 2 + u, 0), h - 1);
      dnn[u] = flow_at(k + 4 + u);
      if (kPrefetch) {
        const int on = yn * w + x;
        const int x1 = (int)floorf((float)x + dn[u].x), y1 = (int)floorf((float)yn + dn[u].y);
        const int qn = min(max(y1, 0), h - 2) * w + min(max(x1, 0), w - 2) + w;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ra0 + on));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rb0 + on));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ra1 + qn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rb1 + qn));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const Tap2 prev = u == 0 ? pbot : bot[0];
      Tap2 top;
      top.a0.x = need[u] ? topl[u].a0.x : prev.a0.x; top.a0.y = need[u] ? topl[u].a0.y : prev.a0.y;
      top.a0.z = need[u] ? topl[u].a0.z : prev.a0.z; top.a0.w = need[u] ? topl[u].a0.w : prev.a0.w;
      top.a1.x = need[u] ? topl[u].a1.x : prev.a1.x; top.a1.y = need[u] ? topl[u].a1.y : prev.a1.y;
      top.a1.z = need[u] ? topl[u].a1.z : prev.a1.z; top.a1.w = need[u] ? topl[u].a1.w : prev.a1.w;
      top.b0 = need[u] ? topl[u].b0 : prev.b0; top.b1 = need[u] ? topl[u].b1 : prev.b1;
      const float dx = d[u].x, dy = d[u].y;
      const float a00 = (1.f - fx[u]) * (1.f - fy[u]), a01 = fx[u] * (1.f - fy[u]), a10 = (1.f - fx[u]) * fy[u], a11 = fx[u] * fy[u];
      float r2 = a00 * top.a0.x + a01 * top.a1.x + a10 * bot[u].a0.x + a11 * bot[u].a1.x;
      float r3 = a00 * top.a0.y + a01 * top.a1.y + a10 * bot[u].a0.y + a11 * bot[u].a1.y;
      float r4 = a00 * top.a0.z + a01 * top.a1.z + a10 * bot[u].a0.z + a11 * bot[u].a1.z;
      float r5 = a00 * top.a0.w + a01 * top.a1.w + a10 * bot[u].a0.w + a11 * bot[u].a1.w;
      float r6 = a00 * top.b0 + a01 * top.b1 + a10 * bot[u].b0 + a11 * bot[u].b1;
      r2 = in[u] ? r2 : 0.f; r3 = in[u] ? r3 : 0.f;
      r4 = in[u] ? (c0[u].z + r4) * 0.5f : c0[u].z;
      r5 = in[u] ? (c0[u].w + r5) * 0.5f : c0[u].w;
      r6 = in[u] ? (c0xy[u] + r6) * 0.25f : c0xy[u] * 0.5f;
      r2 = (c0[u].x - r2) * 0.5f;
      r3 = (c0[u].y - r3) * 0.5f;
      r2 += r4 * dy + r6 * dx;
      r3 += r6 * dy + r5 * dx;
      if (xedge || (unsigned)(y[u] - 5) >= (unsigned)(h - 10)) {
        const float s = border_w(y[u], h) * bwx;
        r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
      }
      float m[5];
      m[0] = r4 * r4 + r6 * r6;
      m[1] = (r4 + r5) * r6;
      m[2] = r5 * r5 + r6 * r6;
      m[3] = r4 * r2 + r6 * r3;
      m[4] = r6 * r2 + r5 * r3;
      // vertical running sums (compensated: a segment slides over a few hundred rows) and the ring of the last 15 rows
      const int kk = k + u;
      float* rp = ring + slot * (5 * MS_NT) + tx;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (kDoubleSums) {
          double inc = (double)m[c];
          if (kk >= 15) inc -= (double)rp[c * MS_NT];      // exact in f64
          rp[c * MS_NT] = m[c];
          vd[c] += inc;
          vs[c] = (float)vd[c];
        } else {
          float inc = m[c];
          if (kk >= 15) inc -= rp[c * MS_NT];
          rp[c * MS_NT] = m[c];
          const float yk = inc - comp[c], t = vs[c] + yk;
          comp[c] = (t - vs[c]) - yk;
          vs[c] = t;
        }
      }
      slot = slot == 14 ? 0 : slot + 1;
      if (kk >= 14) {
        float* hp = hb + ((kk - 14) & (MS_RB - 1)) * (5 * MS_NT) + tx;
#pragma unroll
        for (int c = 0; c < 5; ++c) hp[c * MS_NT] = vs[c];
      }
    }
    pbot = bot[1]; pq = q[1] + w;
    if (k >= 16 && (k & 3) == 0) {
      // rows ya + k - 16 .. ya + k - 13 are complete: horizontal 15-sums (slots o+1 .. o+15 for output o) + solve
      __syncthreads();
      if (tx < MS_RB * (MS_SX / 4)) {
        const int j = tx / (MS_SX / 4), qd = tx - j * (MS_SX / 4);
        const int gy = ya + k - 16 + j, gx0 = x0 + 4 * qd;
        float g[5][4];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const float4* p = reinterpret_cast<const float4*>(hb + (j * 5 + c) * MS_NT + 4 * qd);
          float wv[20];
#pragma unroll
          for (int i = 0; i < 5; ++i) { const float4 v = p[i]; wv[4 * i] = v.x; wv[4 * i + 1] = v.y; wv[4 * i + 2] = v.z; wv[4 * i + 3] = v.w; }
          float s = 0.f;
#pragma unroll
          for (int i = 1; i <= 15; ++i) s += wv[i];
          g[c][0] = s;
#pragma unroll
          for (int o = 1; o < 4; ++o) { s += wv[o + 15] - wv[o]; g[c][o] = s; }
        }
        if (gy < yb && gx0 < w) {
          float2 f[4];
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            // f64 solve as OpenCV (an fp32 version with FMA-compensated determinants was as accurate but slower:
            // the fp64 pipe runs beside the fp32 one, profiles/r1_flow_experiments.md)
            const double sc = 1.0 / 225.0;
            const double g11 = g[0][o] * sc, g12 = g[1][o] * sc, g22 = g[2][o] * sc, h1 = g[3][o] * sc, h2 = g[4][o] * sc;
            const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
            f[o].x = (float)((g11 * h2 - g12 * h1) * idet);
            f[o].y = (float)((g22 * h1 - g12 * h2) * idet);
            if (kMinMax && gx0 + o < w) {
              const float m = magnitude(f[o].x, f[o].y);
              mag_mn = fminf(mag_mn, m); mag_mx = fmaxf(mag_mx, m);
            }
          }
          const size_t oi = zo + (size_t)gy * w + gx0;
          float2* dst = reinterpret_cast<float2*>(flow_out) + oi;
          if (gx0 + 3 < w && (oi & 1) == 0 && (reinterpret_cast<uintptr_t>(flow_out) & 15) == 0) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
            reinterpret_cast<float4*>(dst)[1] = make_float4(f[2].x, f[2].y, f[3].x, f[3].y);
          } else {
#pragma unroll
            for (int o = 0; o < 4; ++o) if (gx0 + o < w) dst[o] = f[o];
          }
        }
      }
      __syncthreads();
    }
  }
  if (kMinMax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mag_mn = fminf(mag_mn, __shfl_xor_sync(0xffffffffu, mag_mn, o)); mag_mx = fmaxf(mag_mx, __shfl_xor_sync(0xffffffffu, mag_mx, o));
    }
    if ((tx & 31) == 0) {                                   // magnitudes are >= 0: their bit patterns order like unsigned integers
      atomicMin(reinterpret_cast<unsigned int*>(minmax + 2 * blockIdx.z), __float_as_uint(mag_mn));
      atomicMax(reinterpret_cast<unsigned int*>(minmax + 2 * blockIdx.z + 1), __float_as_uint(mag_mx));
    }
  }
}

// ---- (d+e), streaming form, second version (A/B variant, NOT the default: same speed; profiles/r2_flow_experiments.md).
// k4_flow_iter_march3 keeps the dataflow of k4_flow_iter_march and cuts the instruction count per pixel from ~340 to ~290
// (an intermediate version without the software pipeline reached ~250):
//  * every global address is ONE 64-bit multiply-add from a 32-bit pixel index against a per-image base pointer held in
//    registers (the first version spends 4 integer instructions per address on 11 addresses per pixel);
//  * the per-thread L2 prefetches (4 addresses per pixel, each with its own floor / clamp / address chain) became one
//    prefetch instruction per row issued by 14 lanes of the warp (details at the kernel);
//  * the bilinear taps and half of UpdateMatrices run on the packed fp32x2 pipe (the float4 coefficient layout puts
//    (y, x) and (yy, xx) in aligned register pairs);
//  * the row-batch buffer holds the five vertical sums as three float2 planes, so the horizontal 15-sums are packed
//    adds (60 instead of 100 per 4 outputs) on the same 128-bit shared-memory windows, conflict-free with a row pitch
//    of 16 B mod 128 B and two rows per quarter-warp;
//  * the 2x2 solve keeps OpenCV's fp64 determinants (products of fp32 values are exact in fp64) with the 1/225^2 scale
//    folded into the regulariser, and divides in fp32 (relative 2e-7: < 1e-5 px) instead of a 12-instruction fp64
//    reciprocal.  kExactSolve keeps the first version's solve: with it the two kernels agree bit for bit
//    (tools/flow_ab.py), which is how this one was checked.
// Measured: the instruction diet alone changed nothing (issue slots 58 % -> 41 % busy, same 750 us per level-0 launch):
// the kernel is bound by dependent-instruction latency at 4 warps per scheduler (ring + registers cap the SM at 16 warps).
constexpr int M2_NT = 256, M2_SX = M2_NT - 2 * MS_HALO, M2_P2 = M2_NT + 2;              // hb row pitch: 258 float2 = 2064 B
constexpr int m2_smem() { return 15 * 5 * M2_NT * 4 + 3 * MS_RB * M2_P2 * 8; }         // ring + row-batch buffer = 101,568 B

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
// a00 t00 + a01 t01 + a10 t10 + a11 t11, in the association the first version compiled to
__device__ __forceinline__ float2 bil2(float2 w00, float2 w01, float2 w10, float2 w11, float2 t00, float2 t01, float2 t10, float2 t11) {
  float2 r = __fmul2_rn(w01, t01);
  r = __ffma2_rn(w00, t00, r);
  r = __ffma2_rn(w10, t10, r);
  return __ffma2_rn(w11, t11, r);
}

// Software pipeline: the loads of row k + 1 (R0, flow-displaced R1 taps) are issued BEFORE the arithmetic of row k, so a
// row's worth of instructions (x 4 warps per scheduler) covers their latency.  Two register sets (even / odd rows) hold
// the in-flight row; the lower taps of the row just finished are copied to `T` (the upper taps of the next row when the
// displacement is locally smooth) before their set is reloaded.  Flow vectors run four rows ahead; the row prefetcher
// (one instruction per row, lanes 0-13) requests the R0 / R1 lines of row k + 3 (R1 rows and columns shifted by the
// lane's own displacement one row earlier) and the flow lines of row k + 8 into L2.
// Same arithmetic in the same order as k4_flow_iter_march: kExactSolve gives bit-identical flow.
struct M3Row { float4 c0; float c0xy; Tap2 bot; float2 d; float fx, fy; int q, yv; bool in; };

template <bool kExactSolve>
__global__ void __launch_bounds__(M2_NT, 2)
k4_flow_iter_march3(const float4* __restrict__ RA0, const float* __restrict__ RB0, const float4* __restrict__ RA1,
                    const float* __restrict__ RB1, const float* __restrict__ flow_in, int h, int w, int rows_per_seg,
                    float* __restrict__ flow_out) {
  extern __shared__ __align__(16) float ms_smem_buf[];
  float* ring = ms_smem_buf;                                                       // [15][5][M2_NT]
  float2* hb = reinterpret_cast<float2*>(ms_smem_buf + 15 * 5 * M2_NT);            // [3][MS_RB][M2_P2]: (m0,m1) (m2,m3) (m4,-)
  const int tx = threadIdx.x;
  const int x0 = blockIdx.x * M2_SX;
  const int ya = blockIdx.y * rows_per_seg, yb = min(ya + rows_per_seg, h);
  const size_t zo = (size_t)blockIdx.z * h * w;
  const float2* fin = reinterpret_cast<const float2*>(flow_in) + zo;
  const float4* ra0 = RA0 + zo; const float* rb0 = RB0 + zo;
  const float4* ra1 = RA1 + zo; const float* rb1 = RB1 + zo;
  asm volatile("" : "+l"(fin), "+l"(ra0), "+l"(rb0), "+l"(ra1), "+l"(rb1));        // opaque: addresses = base + 32-bit index
  __builtin_assume(__isGlobal(fin)); __builtin_assume(__isGlobal(ra0)); __builtin_assume(__isGlobal(rb0));
  __builtin_assume(__isGlobal(ra1)); __builtin_assume(__isGlobal(rb1));
  const int x = min(max(x0 - MS_HALO + tx, 0), w - 1);                             // replicate border: M at the clamped pixel
  const bool xedge = (unsigned)(x - 5) >= (unsigned)(w - 10);
  const float bwx = border_w(x, w), xf = (float)x;
  const int nk = 14 + ((yb - ya + MS_RB - 1) & ~(MS_RB - 1));                      // rows of M this block walks (even)
  // row prefetcher of this lane: base pointer (at the line's first pixel), row pitch in bytes, row / column follow the flow?
  const int lane = tx & 31;
  const char* pf_base; int pf_pitch, pf_follow, pf_esz, pf_ahead, pf_px;
  {
    const int xw = x0 - MS_HALO + (tx & ~31);
    pf_px = lane < 8 ? xw + 8 * (lane & 3) : (lane < 12 ? xw + 31 * (lane & 1) : xw + 16 * (lane & 1));
    if (lane < 4) { pf_base = reinterpret_cast<const char*>(ra0); pf_esz = 16; pf_follow = 0; pf_ahead = 3; }
    else if (lane < 8) { pf_base = reinterpret_cast<const char*>(ra1); pf_esz = 16; pf_follow = 1; pf_ahead = 3; }
    else if (lane < 10) { pf_base = reinterpret_cast<const char*>(rb0); pf_esz = 4; pf_follow = 0; pf_ahead = 3; }
    else if (lane < 12) { pf_base = reinterpret_cast<const char*>(rb1); pf_esz = 4; pf_follow = 1; pf_ahead = 3; }
    else { pf_base = reinterpret_cast<const char*>(fin); pf_esz = 8; pf_follow = 0; pf_ahead = 8; }
    pf_pitch = pf_esz * w;
  }
  const bool pf_on = lane < 14;
  double vd[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) vd[c] = 0.0;
  int slot = 0;
  const float2 neg1 = make_float2(-1.f, -1.f), one2 = make_float2(1.f, 1.f), half2 = make_float2(0.5f, 0.5f);
  auto row_y = [&](int kk) { return min(max(ya - 7 + kk, 0), h - 1); };
  // start the loads of march row kk (flow vector d): first-image coefficients at the pixel, lower taps of the second image
  auto issue = [&](M3Row& R, int kk, float2 d) {
    R.yv = row_y(kk);
    const int o = R.yv * w + x;
    R.d = d; R.c0 = ra0[o]; R.c0xy = rb0[o];
    const float gxf = xf + d.x, gyf = (float)R.yv + d.y;
    const int x1 = __float2int_rd(gxf), y1 = __float2int_rd(gyf);
    R.fx = gxf - (float)x1; R.fy = gyf - (float)y1;
    R.in = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
    R.q = R.in ? y1 * w + x1 : 0;                                                  // outside: any valid address, result unused
    R.bot = ld_tap2(ra1, rb1, R.q + w);
  };
  // UpdateMatrices at the row's pixel + vertical running sums + row-batch buffer
  auto compute = [&](const M3Row& R, const Tap2& top, int kk) {
    const float dx = R.d.x, dy = R.d.y;
    const float2 fx2 = make_float2(R.fx, R.fx), fy2 = make_float2(R.fy, R.fy);
    const float2 ox2 = __ffma2_rn(fx2, neg1, one2), oy2 = __ffma2_rn(fy2, neg1, one2);
    const float2 w00 = __fmul2_rn(ox2, oy2), w01 = __fmul2_rn(fx2, oy2), w10 = __fmul2_rn(ox2, fy2), w11 = __fmul2_rn(fx2, fy2);
    float2 r23 = bil2(w00, w01, w10, w11, lo2(top.a0), lo2(top.a1), lo2(R.bot.a0), lo2(R.bot.a1));
    float2 r45 = bil2(w00, w01, w10, w11, hi2(top.a0), hi2(top.a1), hi2(R.bot.a0), hi2(R.bot.a1));
    float r6 = w01.x * top.b1;
    r6 = fmaf(w00.x, top.b0, r6); r6 = fmaf(w10.x, R.bot.b0, r6); r6 = fmaf(w11.x, R.bot.b1, r6);
    // outside the image: r2 = r3 = 0 and the second image's quadratic terms are replaced by the first image's
    r23.x = R.in ? r23.x : 0.f; r23.y = R.in ? r23.y : 0.f;
    r45.x = R.in ? r45.x : R.c0.z; r45.y = R.in ? r45.y : R.c0.w;
    r6 = R.in ? r6 : R.c0xy;
    r45 = __fmul2_rn(__fadd2_rn(hi2(R.c0), r45), half2);
    r6 = (R.c0xy + r6) * 0.25f;
    r23 = __ffma2_rn(r23, neg1, lo2(R.c0));
    float r2 = fmaf(r23.x, 0.5f, fmaf(r45.x, dy, r6 * dx));
    float r3 = fmaf(r23.y, 0.5f, fmaf(r45.y, dx, r6 * dy));
    float r4 = r45.x, r5 = r45.y;
    if (xedge || (unsigned)(R.yv - 5) >= (unsigned)(h - 10)) {
      const float s = border_w(R.yv, h) * bwx;
      r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
    }
    float m[5];
    const float r66 = r6 * r6;
    m[0] = fmaf(r4, r4, r66);
    m[1] = (r4 + r5) * r6;
    m[2] = fmaf(r5, r5, r66);
    m[3] = fmaf(r2, r4, r3 * r6);
    m[4] = fmaf(r3, r5, r2 * r6);
    float* rp = ring + slot * (5 * M2_NT) + tx;
    float vs[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      double inc = (double)m[c];
      if (kk >= 15) inc -= (double)rp[c * M2_NT];                                  // exact in f64
      rp[c * M2_NT] = m[c];
      vd[c] += inc;
      vs[c] = (float)vd[c];
    }
    slot = slot == 14 ? 0 : slot + 1;
    if (kk >= 14) {
      float2* hp = hb + ((kk - 14) & (MS_RB - 1)) * M2_P2 + tx;
      hp[0] = make_float2(vs[0], vs[1]);
      hp[MS_RB * M2_P2] = make_float2(vs[2], vs[3]);
      hp[2 * MS_RB * M2_P2] = make_float2(vs[4], 0.f);
    }
  };
  // flow vectors of march rows kk + 1 .. kk + 3 (f2, f3, f4); row kk's is in its register set
  M3Row X, Y;
  Tap2 T; T.a0 = T.a1 = make_float4(0.f, 0.f, 0.f, 0.f); T.b0 = T.b1 = 0.f;
  int tq = -1;                                                                     // pixel index of the row held in T
  float2 f2 = fin[row_y(1) * w + x], f3 = fin[row_y(2) * w + x], f4 = fin[row_y(3) * w + x];
  issue(X, 0, fin[row_y(0) * w + x]);
  // one pipeline step: prefetch for row kk + 3, flow of row kk + 4, loads of row kk + 1 into N, arithmetic of row kk from C
  const unsigned ones = h > 0 ? 0xffffffffu : 0u;                                  // all-ones the compiler cannot fold
  auto step = [&](M3Row& C, M3Row& N, int kk) {
    // Row kk's loads must have landed before row kk + 1's are issued: ptxas tracks every global load of this kernel on
    // ONE scoreboard counter, so a consumer of row kk issued after row kk + 1's loads would wait for those as well
    // (that is what march3 did without this line: tools/sbdecode.py).  A warp barrier whose mask depends on a loaded
    // value waits for the counter here, and keeps ptxas from hoisting the next loads above it.
    __syncwarp(__float_as_uint(C.c0xy) | ones);
    {
      const int prow = min(max(row_y(kk + pf_ahead) + pf_follow * (__float2int_rd(f3.y) + 1), 0), h - 1);
      const int pcol = min(max(pf_px + pf_follow * __float2int_rd(f3.x), 0), w - 1);
      if (pf_on) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_base + (long long)prow * pf_pitch + pcol * pf_esz));
    }
    const float2 fnew = fin[row_y(kk + 4) * w + x];
    issue(N, kk + 1, f2);
    if (C.q != tq) T = ld_tap2(ra1, rb1, C.q);                                     // upper taps = the previous row's lower taps, usually
    compute(C, T, kk);
    T = C.bot; tq = C.q + w;
    f2 = f3; f3 = f4; f4 = fnew;
  };
  for (int k = 0; k < nk; k += 2) {
    step(X, Y, k);
    step(Y, X, k + 1);
    if (k >= 16 && (k & 3) == 0) {
      // rows ya + k - 16 .. ya + k - 13 are complete: horizontal 15-sums (slots o+1 .. o+15 for output o) + solve.
      // threads 2 qd + (j & 1) + 120 (j >> 1): a quarter-warp reads 4 windows x 2 rows = 8 distinct 16-byte bank groups
      __syncthreads();
      if (tx < MS_RB * (M2_SX / 4)) {
        const int hf = tx >= 2 * (M2_SX / 4) ? 1 : 0, tt = tx - hf * 2 * (M2_SX / 4);
        const int j = 2 * hf + (tt & 1), qd = tt >> 1;
        const int gy = ya + k - 16 + j, gx0 = x0 + 4 * qd;
        float2 g[3][4];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float4* p = reinterpret_cast<const float4*>(hb + (a * MS_RB + j) * M2_P2 + 4 * qd);
          float2 wv[20];
#pragma unroll
          for (int i = 0; i < 10; ++i) { const float4 v = p[i]; wv[2 * i] = make_float2(v.x, v.y); wv[2 * i + 1] = make_float2(v.z, v.w); }
          float2 s = wv[1];
#pragma unroll
          for (int i = 2; i <= 15; ++i) s = __fadd2_rn(s, wv[i]);
          g[a][0] = s;
#pragma unroll
          for (int o = 1; o < 4; ++o) { s = __fadd2_rn(s, __ffma2_rn(wv[o], neg1, wv[o + 15])); g[a][o] = s; }
        }
        if (gy < yb && gx0 < w) {
          float2 f[4];
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            if (kExactSolve) {
              const double sc = 1.0 / 225.0;
              const double g11 = g[0][o].x * sc, g12 = g[0][o].y * sc, g22 = g[1][o].x * sc, h1 = g[1][o].y * sc, h2 = g[2][o].x * sc;
              const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
              f[o].x = (float)((g11 * h2 - g12 * h1) * idet);
              f[o].y = (float)((g22 * h1 - g12 * h2) * idet);
            } else {
              // flow = (G11 H2 - G12 H1, G22 H1 - G12 H2) / (G11 G22 - G12^2 + 1e-3 * 225^2) on the un-scaled sums
              const double g11 = g[0][o].x, g12 = g[0][o].y, g22 = g[1][o].x, h1 = g[1][o].y, h2 = g[2][o].x;
              const float det = (float)fma(g11, g22, fma(-g12, g12, 1e-3 * 225.0 * 225.0));
              const float n1 = (float)fma(g11, h2, -(g12 * h1)), n2 = (float)fma(g22, h1, -(g12 * h2));
              float idet;
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(idet) : "f"(det));            // det >= 50.6: one MUFU, 1 ulp
              f[o].x = n1 * idet; f[o].y = n2 * idet;
            }
          }
          const size_t oi = zo + (size_t)gy * w + gx0;
          float2* dst = reinterpret_cast<float2*>(flow_out) + oi;
          if (gx0 + 3 < w && (oi & 1) == 0 && (reinterpret_cast<uintptr_t>(flow_out) & 15) == 0) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
            reinterpret_cast<float4*>(dst)[1] = make_float4(f[2].x, f[2].y, f[3].x, f[3].y);
          } else {
#pragma unroll
            for (int o = 0; o < 4; ++o) if (gx0 + o < w) dst[o] = f[o];
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- (f) flow upsample between levels: resize(prev, (w,h), INTER_LINEAR) * 2.  One thread per column walks UP_ROWS
// rows: the column's source coordinates (double, as cv::resize computes them) are evaluated once, the rows' once per
// block into shared memory.
constexpr int UP_ROWS = 8;
__global__ void __launch_bounds__(256)
k4_flow_upsample(const float* __restrict__ prev, int hp, int wp, int h, int w, double sx, double sy, float* __restrict__ out) {
  __shared__ UpCoord s_y[UP_ROWS];
  const int xo = blockIdx.x * blockDim.x + threadIdx.x, yb = blockIdx.y * UP_ROWS;
  if (threadIdx.x < UP_ROWS) s_y[threadIdx.x] = up_coord(yb + threadIdx.x, sy, hp);
  __syncthreads();
  if (xo >= w) return;
  const float2* p = reinterpret_cast<const float2*>(prev) + (size_t)blockIdx.z * hp * wp;
  const UpCoord cx = up_coord(xo, sx, wp);
  float2* o = reinterpret_cast<float2*>(out) + ((size_t)blockIdx.z * h + yb) * w + xo;
#pragma unroll
  for (int r = 0; r < UP_ROWS; ++r) {
    if (yb + r < h) o[(size_t)r * w] = up_sample(p, wp, s_y[r], cx);
  }
}

// =========================================================================== flow colouring
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
// The reference normalises the magnitude twice (mag = normalize(mag); hsv[..., 2] = normalize(mag)); cv2.normalize on
// float32 is dst = fma(src, scale, shift) with scale = (float)(255 * (1 / (max - min))) (double division), shift = -(min * scale).
struct NormConsts { float a1, b1, a2, b2; };
__device__ __forceinline__ uchar3 flow_colour(float dx, float dy, const NormConsts nc) {
  const float mag = magnitude(dx, dy);
  const float ang = __fmul_rn(fast_atan2_deg(dy, dx), 0.01745329238474369f);             // radians (float32(pi/180))
  const float magn = __fmaf_rn(__fmaf_rn(mag, nc.a1, nc.b1), nc.a2, nc.b2);
  const float hue = __fmul_rn(__fdiv_rn(__fmul_rn(ang, 180.f), 3.1415927410125732f), 0.5f);
  const int H8 = (int)hue & 255, V8 = (int)fminf(fmaxf(magn, 0.f), 255.f);
  // cv2 HSV2BGR (8-bit, hrange 180): S = 255 -> s = 1
  const float hh = __fmul_rn((float)H8, 0.03333333507180214f);                           // float32(6/180)
  const float s = __fmul_rn(255.f, 0.003921568859368563f), v = __fmul_rn((float)V8, 0.003921568859368563f);   // float32(1/255)
  int sector = (int)floorf(hh);
  const float f = __fsub_rn(hh, (float)sector);
  sector %= 6;
  const float t0 = v;
  const float t1 = __fmul_rn(v, __fsub_rn(1.f, s));
  const float t2 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, f)));
  const float t3 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f))));
  // sector -> (b, g, r) table {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0} as selects (no local memory)
  const float bsel = sector < 2 ? t1 : (sector == 2 ? t3 : (sector < 5 ? t0 : t2));
  const float gsel = sector == 0 ? t3 : (sector < 3 ? t0 : (sector == 3 ? t2 : t1));
  const float rsel = sector == 0 ? t0 : (sector == 1 ? t2 : (sector < 4 ? t1 : (sector == 4 ? t3 : t0)));
  const float b = __fmul_rn(bsel, 255.f), g = __fmul_rn(gsel, 255.f), r = __fmul_rn(rsel, 255.f);
  return make_uchar3((unsigned char)fminf(fmaxf(b, 0.f), 255.f), (unsigned char)fminf(fmaxf(g, 0.f), 255.f),
                     (unsigned char)fminf(fmaxf(r, 0.f), 255.f));
}

__device__ __forceinline__ void norm_pass(float mn, float mx, float& a, float& b) {
  const double dmn = mn, dmx = mx;
  const double sc = (dmx - dmn > 2.220446049250313e-16) ? 255.0 * (1.0 / (dmx - dmn)) : 0.0;
  a = (float)sc;
  b = -__fmul_rn(mn, a);                 // cv2 4.13: the shift uses the float32-rounded scale (probed: oracle/farneback.py)
}
__device__ __forceinline__ NormConsts norm_consts(const float* minmax) {
  NormConsts nc;
  norm_pass(minmax[0], minmax[1], nc.a1, nc.b1);
  // the first pass is monotone, so the extrema of its output are the images of the extrema
  norm_pass(__fmaf_rn(minmax[0], nc.a1, nc.b1), __fmaf_rn(minmax[1], nc.a1, nc.b1), nc.a2, nc.b2);
  return nc;
}

// colour + 16x16 patch sums (+ optional image store).  block = 64 px x 16 rows (4 patches).
__global__ void __launch_bounds__(1024)
k5_flow_rgb_patchsum(const float* __restrict__ flow, int H, int W, const float* __restrict__ minmax, uint8_t* __restrict__ rgb,
                     uint32_t* __restrict__ sums) {
  __shared__ uint32_t part[4];
  __shared__ NormConsts s_norm;
  const int gw = W >> 4, gh = H >> 4;
  if (threadIdx.y == 0 && threadIdx.x < 4) part[threadIdx.x] = 0;
  if (threadIdx.y == 1 && threadIdx.x == 0) s_norm = norm_consts(minmax + 2 * blockIdx.z);      // two f64 divisions per block
  __syncthreads();
  const NormConsts nc = s_norm;
  const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 16 + threadIdx.y;
  uint32_t s = 0;
  if (x < W && y < H) {
    const size_t p = ((size_t)blockIdx.z * H + y) * W + x;
    const float2 d = reinterpret_cast<const float2*>(flow)[p];
    const uchar3 c = flow_colour(d.x, d.y, nc);
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

// Same outputs, warp-per-band form (default): one warp colours a band of 128 pixels x 16 rows (8 patches), each lane 4
// consecutive pixels per row (two 16-byte loads), and keeps its part of the patch sums in a register over the 16 rows:
// two shuffles per patch instead of four per pixel row, no shared memory, no atomics, no block barrier (the 64 x 16 block
// form above stalls all 1024 threads on the two f64 divisions of the normalisation constants: 372 us per 22 pairs of
// 1080p against a 57 us read of the flow).  The constants are computed by lane 0 of every warp and broadcast.
constexpr int RG_WARPS = 8;
__global__ void __launch_bounds__(RG_WARPS * 32)
k5_flow_rgb_patchsum_band(const float* __restrict__ flow, int H, int W, const float* __restrict__ minmax, uint8_t* __restrict__ rgb,
                          uint32_t* __restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const int strip = blockIdx.x * RG_WARPS + (threadIdx.x >> 5);
  const int x0 = strip * 128 + lane * 4, py = blockIdx.y, b = blockIdx.z;
  if (strip * 128 >= W) return;                               // whole warp
  NormConsts nc;
  if (lane == 0) nc = norm_consts(minmax + 2 * b);
  nc.a1 = __shfl_sync(0xffffffffu, nc.a1, 0); nc.b1 = __shfl_sync(0xffffffffu, nc.b1, 0);
  nc.a2 = __shfl_sync(0xffffffffu, nc.a2, 0); nc.b2 = __shfl_sync(0xffffffffu, nc.b2, 0);
  const int gw = W >> 4, gh = H >> 4;
  const bool vec = (W & 3) == 0 && x0 + 3 < W && (reinterpret_cast<uintptr_t>(flow) & 15) == 0;
  uint32_t s = 0;
#pragma unroll 4
  for (int r = 0; r < 16; ++r) {
    const int y = py * 16 + r;
    if (y >= H) break;
    const size_t p = ((size_t)b * H + y) * W + x0;
    float2 d[4];
    if (vec) {
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(flow + 2 * p)), q1 = __ldg(reinterpret_cast<const float4*>(flow + 2 * p) + 1);
      d[0] = make_float2(q0.x, q0.y); d[1] = make_float2(q0.z, q0.w); d[2] = make_float2(q1.x, q1.y); d[3] = make_float2(q1.z, q1.w);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = x0 + j < W ? reinterpret_cast<const float2*>(flow)[p + j] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (x0 + j < W) {
        const uchar3 c = flow_colour(d[j].x, d[j].y, nc);
        if (rgb) { rgb[(p + j) * 3] = c.x; rgb[(p + j) * 3 + 1] = c.y; rgb[(p + j) * 3 + 2] = c.z; }
        s += (uint32_t)c.x + c.y + c.z;
      }
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);                     // 4 lanes = the 16 columns of one patch
  if (sums && (lane & 3) == 0 && (x0 >> 4) < gw && py < gh) sums[((size_t)b * gh + py) * gw + (x0 >> 4)] = s;
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
    const NormConsts nc = norm_consts(minmax + 2 * b);
    const float2 d = reinterpret_cast<const float2*>(flow)[((size_t)b * H + py * 16 + r) * W + px * 16 + c];
    col = flow_colour(d.x, d.y, nc);
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

// rows per segment of the streaming expansion kernel (no running sums: the split does not change any result bit)
static int poly_rows_per_seg(int h, int w, int images, int sm_count) {
  const long strips = cdiv(w, PM_SX), resident = 3L * sm_count;
  double best = 1e30;
  int best_rs = (h + 3) & ~3;
  for (int nseg = 1; nseg <= cdiv(h, 8); ++nseg) {
    const int rs = (cdiv(h, nseg) + 3) & ~3;
    const long ctas = strips * cdiv(h, rs) * images;
    const long waves = (ctas + resident - 1) / resident;
    const double cost = (double)waves * (rs + 14);
    if (cost < best) { best = cost; best_rs = rs; }
  }
  return best_rs;
}

// rows per segment of the streaming iteration kernel: whole waves of the 2 x SM-count resident blocks, each block
// paying 14 warm-up rows (+ ~6 rows' worth of fixed cost).  The running sums make the last bits depend on where a
// segment starts, so the split is a function of (h, w) alone (sized for a nominal batch of 22 pairs = one 10 s clip):
// results stay bit-identical for any batch size / GPU count.
static int march_rows_per_seg(int h, int w, int sm_count, int MS_SX = 240, int per_sm = 2) {
  const int B = 22;
  const long strips = cdiv(w, MS_SX), resident = (long)per_sm * sm_count;
  double best = 1e30;
  int best_rs = (h + 3) & ~3;
  for (int nseg = 1; nseg <= cdiv(h, 8); ++nseg) {
    const int rs = (cdiv(h, nseg) + 3) & ~3;
    const long ctas = strips * cdiv(h, rs) * B;
    const long waves = (ctas + resident - 1) / resident;
    const double cost = (double)waves * (rs + 20);
    if (cost < best) { best = cost; best_rs = rs; }
  }
  return best_rs;
}

// Function attributes are per device: called by b200vqa_create for the context's device (not behind a process-wide flag).
int flow_init_device_attrs() {
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, BX_SMEM));
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter_march<true, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms_smem(256)));
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter_march<true, true, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms_smem(256)));
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter_march<true, true, 256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms_smem(256) + 16 * 1024));
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter_march<false, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms_smem(256)));
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter_march<true, true, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms_smem(192)));
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter_march3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, m2_smem()));
  VQA_CUDA(cudaFuncSetAttribute(k4_flow_iter_march3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, m2_smem()));
  VQA_CUDA(cudaFuncSetAttribute(k4_pyr_level<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  VQA_CUDA(cudaFuncSetAttribute(k4_pyr_level<3, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  VQA_CUDA(cudaFuncSetAttribute(k4_pyr_level<9, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  VQA_CUDA(cudaFuncSetAttribute(k4_pyr_level<19, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  return B200VQA_OK;
}

}  // namespace b200vqa

using namespace b200vqa;

// minmax != nullptr: [B][2] min / max of the final flow's magnitude, reduced by the launch that writes the flow
// (default streaming kernel only; otherwise a separate pass)
static int farneback_impl(b200vqa_t* h, const uint8_t* gray0, const uint8_t* gray1, int B, int H, int W, float* flow, float* minmax,
                          cudaStream_t st) {
  if (minmax) {
    k5_minmax_init<<<cdiv(B, 128), 128, 0, st>>>(minmax, B);
    VQA_LAUNCH_CHECK();
  }
  bool minmax_done = false;
  const size_t P = (size_t)H * W;
  // workspace (float units): RA [2B][P] float4 | RB [2B][P] | I [2B][P/4] (levels >= 1 only) | flowA, flowB [B][P][2]
  const std::vector<Level> plan = pyramid_plan(H, W);
  // I: one image pair per coarser level (the fused pyramid kernel writes all of them before the first level runs)
  std::vector<size_t> ioff(plan.size(), 0);
  size_t Pq = 16;
  for (size_t li = 0; li + 1 < plan.size(); ++li) { ioff[li] = Pq; Pq += (((size_t)plan[li].h * plan[li].w * 2 * B) + 15) & ~(size_t)15; }
  const size_t floats = (size_t)B * P * (8 + 2 + 4) + Pq + 64;
  int rc = h->ws_flow.reserve(floats * sizeof(float));
  if (rc) return rc;
  float4* RA = static_cast<float4*>(h->ws_flow.ptr);
  float* RB = reinterpret_cast<float*>(RA + 2 * (size_t)B * P);
  float* flowA = RB + 2 * (size_t)B * P;
  float* flowB = flowA + 2 * (size_t)B * P;
  float* Ibase = flowB + 2 * (size_t)B * P;
  // levels with an exact power-of-two scale are filtered by ONE pass of k4_pyr_fused over the full-resolution frames
  bool fused[4] = {false, false, false, false};                   // index = log2(scale)
  static const bool pyr_tile = getenv("B200VQA_PYR_TILE") != nullptr;        // A/B: tile kernel for every level
  if (!pyr_tile && h->flow_impl != 2 && (W & 7) == 0 && ((reinterpret_cast<uintptr_t>(gray0) | reinterpret_cast<uintptr_t>(gray1)) & 7) == 0) {
    float* Il[4] = {nullptr, nullptr, nullptr, nullptr};
    PyrFusedTaps ft{};
    for (size_t li = 0; li + 1 < plan.size(); ++li) {
      const Level& L = plan[li];
      const int lg = (int)(plan.size() - 1 - li);                 // log2(1 / scale)
      const int S = 1 << lg;
      if (lg < 1 || lg > 3 || H % S || W % S || L.h != H / S || L.w != W / S || L.ksize != (lg == 1 ? 3 : (lg == 2 ? 9 : 19))) continue;
      const Taps t = gaussian_taps(L.ksize, L.sigma);
      float* dst = lg == 1 ? ft.t1 : (lg == 2 ? ft.t2 : ft.t3);
      for (int k = 0; k < L.ksize; ++k) dst[k] = t.t[k];
      fused[lg] = true;
      Il[lg] = Ibase + ioff[li];
    }
    // too few threads for a streaming kernel (short clips of small frames): the tile kernel is as fast there
    const int G = W >> 3, p3 = (H + 7) >> 3;
    // segment height (in level-3 rows = 8 source rows; each segment re-reads 16 rows of window overlap): whole waves of the
    // 4 x SM-count resident blocks at the least rows per block.  Every output row is computed inside one segment, so the
    // split changes no result bit.
    int os3 = 0;
    if ((long)2 * B * G * cdiv(p3, 4) >= 40000) {
      const long resident = 4L * h->sm_count, per_seg = (long)cdiv(G, PF_NT) * 2 * B;
      long best = -1;
      for (int cand = 4; cand <= p3; ++cand) {
        const long blocks = per_seg * cdiv(p3, cand), waves = (blocks + resident - 1) / resident, cost = waves * (cand + 2);
        if (best < 0 || cost < best) { best = cost; os3 = cand; }
      }
    }
    if (const char* e = getenv("B200VQA_PYR_OS3")) os3 = atoi(e);
    if (!os3) fused[1] = fused[2] = fused[3] = false;
    const int mask = (fused[1] ? 1 : 0) | (fused[2] ? 2 : 0) | (fused[3] ? 4 : 0);
    static const bool pyr_split = getenv("B200VQA_PYR_SPLIT") != nullptr;      // A/B: level 1 in its own launch (fewer registers each)
    const dim3 gf(cdiv(G, PF_NT), cdiv(p3, os3 ? os3 : 1), 2 * B);
#define PYR_FUSED_LAUNCH(M) k4_pyr_fused<M><<<gf, PF_NT, 0, st>>>(gray0, gray1, B, H, W, ft, os3, Il[1], Il[2], Il[3])
    if (mask == 7 && pyr_split) { PYR_FUSED_LAUNCH(1); VQA_LAUNCH_CHECK(); PYR_FUSED_LAUNCH(6); VQA_LAUNCH_CHECK(); }
    else if (mask) {
      switch (mask) {
        case 1: PYR_FUSED_LAUNCH(1); break; case 2: PYR_FUSED_LAUNCH(2); break; case 3: PYR_FUSED_LAUNCH(3); break;
        case 4: PYR_FUSED_LAUNCH(4); break; case 5: PYR_FUSED_LAUNCH(5); break; case 6: PYR_FUSED_LAUNCH(6); break;
        default: PYR_FUSED_LAUNCH(7); break;
      }
      VQA_LAUNCH_CHECK();
    }
#undef PYR_FUSED_LAUNCH
  }
  static const PolyConsts pc = poly_consts();
  static const bool poly_tile = getenv("B200VQA_POLY_TILE") != nullptr;      // A/B: 64 x 16 tile expansion kernel
  float* prev = nullptr;         // flow of the previous (coarser) level
  int ph = 0, pw = 0;
  for (size_t li = 0; li < plan.size(); ++li) {
    const Level& L = plan[li];
    const bool last = li + 1 == plan.size();
    const Taps taps = gaussian_taps(L.ksize, L.sigma);
    const size_t lp = (size_t)L.h * L.w;
    const float4* RA0 = RA; const float4* RA1 = RA + (size_t)B * lp;
    const float* RB0 = RB; const float* RB1 = RB + (size_t)B * lp;
    const int poly_rows = poly_rows_per_seg(L.h, L.w, 2 * B, h->sm_count);
    const dim3 gpoly(cdiv(L.w, PM_SX), cdiv(L.h, poly_rows), 2 * B);
    if (L.h == H && L.w == W) {
      // finest level: 3x3 blur fused into the expansion (no f32 image in HBM)
      if (h->flow_impl == 2 || poly_tile) k4_polyexp<true><<<dim3(cdiv(L.w, PE_TW), cdiv(L.h, PE_TH), 2 * B), 256, 0, st>>>(nullptr, gray0, gray1, B, L.h, L.w, pc, RA, RB);
      else k4_polyexp_march<true><<<gpoly, PM_NT, 0, st>>>(nullptr, gray0, gray1, B, L.h, L.w, poly_rows, pc, RA, RB);
      VQA_LAUNCH_CHECK();
    } else {
      float* I = Ibase + ioff[li];
      const int lg = (int)(plan.size() - 1 - li);
      if (!(lg >= 1 && lg <= 3 && fused[lg])) {
      const double sx = (double)W / L.w, sy = (double)H / L.h;
      // shared-memory window: uint8 source tile + f32 horizontally blurred columns
      const int TY = L.ksize == 3 ? 32 : (L.ksize == 9 ? 16 : 8);
      const int sw = (int)(PY_TX * sx) + L.ksize + 8, sh = (int)(TY * sy) + L.ksize + 4;
      const size_t smem = (((size_t)sh * ((sw + 3) & ~3) + 15) & ~(size_t)15) + (size_t)sh * 2 * PY_TX * sizeof(float);
      if (smem > 100 * 1024) return B200VQA_EINVAL;
      const dim3 gpyr(cdiv(L.w, PY_TX), cdiv(L.h, TY), 2 * B);
      if (L.ksize == 3) k4_pyr_level<3, 32><<<gpyr, 256, smem, st>>>(gray0, gray1, B, H, W, taps, L.h, L.w, sx, sy, I);
      else if (L.ksize == 9) k4_pyr_level<9, 16><<<gpyr, 256, smem, st>>>(gray0, gray1, B, H, W, taps, L.h, L.w, sx, sy, I);
      else if (L.ksize == 19) k4_pyr_level<19, 8><<<gpyr, 256, smem, st>>>(gray0, gray1, B, H, W, taps, L.h, L.w, sx, sy, I);
      else k4_pyr_level<0, 8><<<gpyr, 256, smem, st>>>(gray0, gray1, B, H, W, taps, L.h, L.w, sx, sy, I);
      VQA_LAUNCH_CHECK();
      }
      // I holds [2B][h][w]; expansion of images 0..B-1 then B..2B-1 (contiguous)
      if (h->flow_impl == 2 || poly_tile) k4_polyexp<false><<<dim3(cdiv(L.w, PE_TW), cdiv(L.h, PE_TH), 2 * B), 256, 0, st>>>(I, nullptr, nullptr, B, L.h, L.w, pc, RA, RB);
      else k4_polyexp_march<false><<<gpoly, PM_NT, 0, st>>>(I, nullptr, nullptr, B, L.h, L.w, poly_rows, pc, RA, RB);
      VQA_LAUNCH_CHECK();
    }
    float* fin = flowA;
    if (!prev) {
      VQA_CUDA(cudaMemsetAsync(fin, 0, (size_t)B * lp * 2 * sizeof(float), st));
      count_launch();
    }
    const dim3 gbox(cdiv(L.w, BX_TX), cdiv(L.h, BX_TY), B);
    const int nt = h->flow_impl == 3 ? 192 : 256;
    const int rows_per_seg = march_rows_per_seg(L.h, L.w, 148, ms_sx(nt), nt == 256 ? 2 : 3);      // fixed SM count: the split must not vary between devices
    const dim3 gmarch(cdiv(L.w, ms_sx(nt)), cdiv(L.h, rows_per_seg), B);
    // A/B (B200VQA_UPSAMPLE_FUSED): the first iteration upsamples the coarser level's flow where it reads it instead of a
    // k4_flow_upsample pass.  Bit-identical; MEASURED 4.383 vs 4.417 ms per 22 pairs of 1080p, 5.632 vs 5.620 at 2160p x 6 -
    // the four dependent loads at the head of the flow -> tap chain cost the iteration what the pass saved.  Off by default.
    static const bool up_opt_in = getenv("B200VQA_UPSAMPLE_FUSED") != nullptr;
    const size_t up_table = (((size_t)rows_per_seg + 24) * sizeof(UpCoord) + 15) & ~(size_t)15;      // must leave room for two blocks per SM
    const bool up_fused = prev && h->flow_impl == 0 && up_opt_in && up_table <= 16 * 1024;
    if (prev) {
      if (up_fused) fin = prev;        // read through UpSrc; the first iteration writes the other buffer
      else {
        fin = (prev == flowA) ? flowB : flowA;
        k4_flow_upsample<<<dim3(cdiv(L.w, 256), cdiv(L.h, UP_ROWS), B), 256, 0, st>>>(prev, ph, pw, L.h, L.w, (double)pw / L.w, (double)ph / L.h, fin);
        VQA_LAUNCH_CHECK();
      }
    }
    float* fout = nullptr;
    for (int it = 0; it < 3; ++it) {
      fout = (last && it == 2) ? flow : (fin == flowA ? flowB : flowA);
      std::pair<cudaEvent_t, cudaEvent_t> ev{};
      if (h->profiling) {
        if (!h->prof_pool.empty()) { ev = h->prof_pool.back(); h->prof_pool.pop_back(); }
        else { VQA_CUDA(cudaEventCreate(&ev.first)); VQA_CUDA(cudaEventCreate(&ev.second)); }
        VQA_CUDA(cudaEventRecord(ev.first, st));
      }
      if (up_fused && it == 0) {
        const UpSrc up{prev, ph, pw, (double)pw / L.w, (double)ph / L.h};
        k4_flow_iter_march<true, true, 256, false, true><<<gmarch, 256, ms_smem(256) + up_table, st>>>(RA0, RB0, RA1, RB1, nullptr, L.h, L.w, rows_per_seg, fout, nullptr, up);
      }
      else if (h->flow_impl == 0 && minmax && last && it == 2) {
        k4_flow_iter_march<true, true, 256, true><<<gmarch, 256, ms_smem(256), st>>>(RA0, RB0, RA1, RB1, fin, L.h, L.w, rows_per_seg, fout, minmax);
        minmax_done = true;
      }
      else if (h->flow_impl == 4) k4_flow_iter_march3<false><<<gmarch, M2_NT, m2_smem(), st>>>(RA0, RB0, RA1, RB1, fin, L.h, L.w, rows_per_seg, fout);
      else if (h->flow_impl == 5) k4_flow_iter_march3<true><<<gmarch, M2_NT, m2_smem(), st>>>(RA0, RB0, RA1, RB1, fin, L.h, L.w, rows_per_seg, fout);
      else if (h->flow_impl == 0) k4_flow_iter_march<true, true, 256><<<gmarch, 256, ms_smem(256), st>>>(RA0, RB0, RA1, RB1, fin, L.h, L.w, rows_per_seg, fout);
      else if (h->flow_impl == 3) k4_flow_iter_march<true, true, 192><<<gmarch, 192, ms_smem(192), st>>>(RA0, RB0, RA1, RB1, fin, L.h, L.w, rows_per_seg, fout);
      else if (h->flow_impl == 1) k4_flow_iter_march<false, true, 256><<<gmarch, 256, ms_smem(256), st>>>(RA0, RB0, RA1, RB1, fin, L.h, L.w, rows_per_seg, fout);
      else k4_flow_iter<<<gbox, 256, BX_SMEM, st>>>(RA0, RB0, RA1, RB1, fin, L.h, L.w, fout);
      if (h->profiling) {
        VQA_CUDA(cudaEventRecord(ev.second, st));
        h->prof_events_flow.push_back(ev);
        h->prof_bytes_flow += 56.0 * (double)B * lp;       // R0 20 + R1 20 + flow 8 read, flow 8 written, per pixel
      }
      VQA_LAUNCH_CHECK();
      fin = fout;
    }
    prev = fout; ph = L.h; pw = L.w;
  }
  if (minmax && !minmax_done) {
    const size_t npix = (size_t)H * W;
    int gx = (int)((npix + 256 * 8 - 1) / (256 * 8));
    if (gx > 592) gx = 592;
    k5_mag_minmax<<<dim3(gx, B), 256, 0, st>>>(flow, npix, minmax);
    VQA_LAUNCH_CHECK();
  }
  return B200VQA_OK;
}

static int launch_flow_rgb_sums(const float* flow, int B, int H, int W, const float* minmax, uint8_t* rgb, uint32_t* sums, cudaStream_t st) {
  static const bool rgb_block = getenv("B200VQA_RGB_BLOCK") != nullptr;      // A/B: the 64 x 16 block form
  static const bool rgb_band = getenv("B200VQA_RGB_BAND") != nullptr;        // force the band form on small batches (sanitizer pass)
  // a band = one warp: short clips of small frames do not fill the GPU with bands (273 x 481 x 3: 216 warps), the block form does
  const long bands = (long)B * cdiv(W, 128) * cdiv(H, 16);
  if (rgb_block || (bands < 2400 && !rgb_band)) k5_flow_rgb_patchsum<<<dim3(cdiv(W, 64), cdiv(H, 16), B), dim3(64, 16), 0, st>>>(flow, H, W, minmax, rgb, sums);
  else k5_flow_rgb_patchsum_band<<<dim3(cdiv(cdiv(W, 128), RG_WARPS), cdiv(H, 16), B), RG_WARPS * 32, 0, st>>>(flow, H, W, minmax, rgb, sums);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

extern "C" int b200vqa_farneback(b200vqa_t* h, const uint8_t* gray0, const uint8_t* gray1, int B, int H, int W, float* flow,
                                 void* stream) {
  if (!h || !gray0 || !gray1 || !flow || B <= 0 || H < 16 || W < 16) return B200VQA_EINVAL;
  CtxScope scope(h);
  return farneback_impl(h, gray0, gray1, B, H, W, flow, nullptr, as_stream(stream));
}

extern "C" int b200vqa_farneback_flow_sums(b200vqa_t* h, const uint8_t* gray0, const uint8_t* gray1, int B, int H, int W, float* flow,
                                           uint32_t* sums, float* minmax, void* stream) {
  if (!h || !gray0 || !gray1 || !flow || !sums || !minmax || B <= 0 || H < 16 || W < 16) return B200VQA_EINVAL;
  CtxScope scope(h);
  cudaStream_t st = as_stream(stream);
  int rc = farneback_impl(h, gray0, gray1, B, H, W, flow, minmax, st);
  return rc ? rc : launch_flow_rgb_sums(flow, B, H, W, minmax, nullptr, sums, st);
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
  if (rgb || sums) return launch_flow_rgb_sums(flow, B, H, W, minmax, rgb, sums, st);
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
