// A9: Pillow's 8-bit two-pass resampler (BILINEAR-antialiased and LANCZOS), bit-exact.
// Coefficients are computed on the host exactly as Pillow's precompute_coeffs /
// normalize_coeffs_8bpc do (double math, 22-bit fixed point) and cached per context.
#include <math.h>
#include <vector>
#include "context.h"

namespace b200vqa {

static const int kPrecisionBits = 32 - 8 - 2;

static double filt_bilinear(double x) { x = x < 0 ? -x : x; return x < 1.0 ? 1.0 - x : 0.0; }
static double sinc_filter(double x) { if (x == 0.0) return 1.0; x *= M_PI; return sin(x) / x; }
static double filt_lanczos(double x) { return (-3.0 <= x && x < 3.0) ? sinc_filter(x) * sinc_filter(x / 3.0) : 0.0; }

static int build_table(int in_size, int out_size, int filter, ResizeTable* t) {
  double (*fn)(double) = filter == B200VQA_FILTER_LANCZOS ? filt_lanczos : filt_bilinear;
  const double fsup = filter == B200VQA_FILTER_LANCZOS ? 3.0 : 1.0;
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = fsup * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  std::vector<int> bounds(out_size * 2), kk((size_t)out_size * ksize, 0);
  std::vector<double> w(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (xx + 0.5) * scale, ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) { w[x] = fn((x + xmin - center + 0.5) * ss); ww += w[x]; }
    for (int x = 0; x < xmax; ++x) {
      double v = ww != 0.0 ? w[x] / ww : w[x];
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[xx * 2] = xmin; bounds[xx * 2 + 1] = xmax;
  }
  t->ksize = ksize;
  VQA_CUDA(cudaMalloc(&t->d_bounds, bounds.size() * sizeof(int)));
  VQA_CUDA(cudaMalloc(&t->d_kk, kk.size() * sizeof(int)));
  VQA_CUDA(cudaMemcpy(t->d_bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
  VQA_CUDA(cudaMemcpy(t->d_kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
  std::vector<int> kkT((size_t)out_size * ksize);
  for (int xx = 0; xx < out_size; ++xx)
    for (int i = 0; i < ksize; ++i) kkT[(size_t)i * out_size + xx] = kk[(size_t)xx * ksize + i];
  VQA_CUDA(cudaMalloc(&t->d_kkT, kkT.size() * sizeof(int)));
  VQA_CUDA(cudaMemcpy(t->d_kkT, kkT.data(), kkT.size() * sizeof(int), cudaMemcpyHostToDevice));
  return B200VQA_OK;
}

__device__ __forceinline__ uint8_t clip8(int acc) {
  int v = acc >> kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: one block per input row.  The row is staged in shared memory (coalesced 16-byte loads) and
// repacked to one 32-bit word per pixel, so a tap costs one shared-memory load for all three channels and lanes
// (whose windows start ~scale pixels apart) collide on at most ~2 banks instead of ~6 with packed 3-byte pixels.
// kTwo: a second filter from the same staged row (the engine resizes every frame with BILINEAR and with LANCZOS).
// (A packed-fp32 variant - row as float4 pixels, coefficients split into exact hi / lo halves, three FFMA2 per tap - was
// bit-identical and 1.8x SLOWER, 611 vs 345 us per 22 frames of 1080p: 16-byte shared-memory reads per tap and lane.)
struct ResizeH { const int* bounds; const int* kkT; uint8_t* dst; int swap_rb; };

__device__ __forceinline__ void resample_h_taps(const uint32_t* __restrict__ px, const ResizeH& t, int xx, size_t out_row) {
  const int xmin = t.bounds[xx * 2], cnt = t.bounds[xx * 2 + 1];
  const int* k = t.kkT + xx;                          // tap-major table: kk[i * 224 + xx]
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  const uint32_t* p = px + xmin;
#pragma unroll 4
  for (int i = 0; i < cnt; ++i) {
    const int c = __ldg(k + i * 224);
    const uint32_t v = p[i];
    a0 += (int)(v & 0xff) * c; a1 += (int)((v >> 8) & 0xff) * c; a2 += (int)(v >> 16) * c;
  }
  uint8_t* d = t.dst + (out_row * 224 + xx) * 3;
  if (t.swap_rb) { d[0] = clip8(a2); d[1] = clip8(a1); d[2] = clip8(a0); }
  else { d[0] = clip8(a0); d[1] = clip8(a1); d[2] = clip8(a2); }
}

template <bool kTwo>
__global__ void __launch_bounds__(224)
k7_resample_h(const uint8_t* __restrict__ src, int H, int W, ResizeH ta, ResizeH tb) {
  extern __shared__ __align__(16) uint8_t row[];
  const int y = blockIdx.x, b = blockIdx.y;
  const uint8_t* s = src + ((size_t)b * H + y) * W * 3;
  const int nbytes = W * 3;
  uint32_t* px = reinterpret_cast<uint32_t*>(row + ((nbytes + 15) & ~15));
  if ((nbytes & 15) == 0 && (((uintptr_t)s & 15) == 0)) {
    for (int i = threadIdx.x; i < nbytes / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(row)[i] = __ldg(reinterpret_cast<const uint4*>(s) + i);
  } else {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) row[i] = s[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W; i += blockDim.x)
    px[i] = (uint32_t)row[3 * i] | ((uint32_t)row[3 * i + 1] << 8) | ((uint32_t)row[3 * i + 2] << 16);
  __syncthreads();
  const size_t out_row = (size_t)b * H + y;
  resample_h_taps(px, ta, threadIdx.x, out_row);
  if (kTwo) resample_h_taps(px, tb, threadIdx.x, out_row);
}

// vertical pass over a [H][224][3] image: one block per output row, one thread per (x, c).
__global__ void __launch_bounds__(672)
k7_resample_v(const uint8_t* __restrict__ src, int H, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
              uint8_t* __restrict__ dst, int swap_rb) {
  const int yy = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int ymin = bounds[yy * 2], cnt = bounds[yy * 2 + 1];
  const int* k = kk + (size_t)yy * ksize;
  const uint8_t* s = src + ((size_t)b * H + ymin) * 672 + t;
  int acc = 1 << (kPrecisionBits - 1);
  for (int i = 0; i < cnt; ++i) acc += s[(size_t)i * 672] * __ldg(k + i);
  const int x = t / 3, c = t % 3;
  dst[((size_t)b * 224 + yy) * 672 + x * 3 + (swap_rb ? 2 - c : c)] = clip8(acc);
}

__global__ void k7_copy_swap(const uint8_t* __restrict__ src, size_t npix, uint8_t* __restrict__ dst, int swap_rb) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  uint8_t a = src[i * 3], b = src[i * 3 + 1], c = src[i * 3 + 2];
  dst[i * 3] = swap_rb ? c : a; dst[i * 3 + 1] = b; dst[i * 3 + 2] = swap_rb ? a : c;
}

static int get_table(b200vqa_ctx* h, int in_size, int filter, ResizeTable** out) {
  auto key = std::make_pair(in_size, filter);
  auto it = h->resize_tables.find(key);
  if (it == h->resize_tables.end()) {
    ResizeTable t;
    int rc = build_table(in_size, 224, filter, &t);
    if (rc) return rc;
    it = h->resize_tables.emplace(key, t).first;
  }
  *out = &it->second;
  return B200VQA_OK;
}

}  // namespace b200vqa

using namespace b200vqa;

// horizontal pass of one or two filters over the same source (tb == nullptr: one)
static int launch_resample_h(const uint8_t* src, int B, int H, int W, const ResizeTable* ta, uint8_t* da, int swap_a,
                             const ResizeTable* tb, uint8_t* db, int swap_b, cudaStream_t st) {
  const size_t smem = (((size_t)W * 3 + 15) & ~(size_t)15) + (size_t)W * 4;
  const ResizeH a{ta->d_bounds, ta->d_kkT, da, swap_a};
  if (tb) {
    const ResizeH b2{tb->d_bounds, tb->d_kkT, db, swap_b};
    if (smem > 48 * 1024) VQA_CUDA(cudaFuncSetAttribute(k7_resample_h<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k7_resample_h<true><<<dim3(H, B), 224, smem, st>>>(src, H, W, a, b2);
  } else {
    if (smem > 48 * 1024) VQA_CUDA(cudaFuncSetAttribute(k7_resample_h<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k7_resample_h<false><<<dim3(H, B), 224, smem, st>>>(src, H, W, a, a);
  }
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

extern "C" int b200vqa_resize_pil(b200vqa_t* h, const uint8_t* src, int B, int H, int W, int filter, int swap_rb,
                                  uint8_t* dst, void* stream) {
  if (!h || !src || !dst || B <= 0 || H <= 0 || W <= 0) return B200VQA_EINVAL;
  if (filter != B200VQA_FILTER_BILINEAR && filter != B200VQA_FILTER_LANCZOS) return B200VQA_EINVAL;
  CtxScope scope(h);
  cudaStream_t st = as_stream(stream);
  const bool need_h = W != 224, need_v = H != 224;
  if (!need_h && !need_v) {
    size_t npix = (size_t)B * 224 * 224;
    k7_copy_swap<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(src, npix, dst, swap_rb);
    VQA_LAUNCH_CHECK();
    return B200VQA_OK;
  }
  ResizeTable *th = nullptr, *tv = nullptr;
  int rc;
  if (need_h && (rc = get_table(h, W, filter, &th))) return rc;
  if (need_v && (rc = get_table(h, H, filter, &tv))) return rc;
  const uint8_t* vsrc = src;
  if (need_h) {
    uint8_t* hdst = dst;
    if (need_v) {
      if ((rc = h->ws_resize.reserve((size_t)B * H * 672))) return rc;
      hdst = static_cast<uint8_t*>(h->ws_resize.ptr);
    }
    if ((rc = launch_resample_h(src, B, H, W, th, hdst, need_v ? 0 : swap_rb, nullptr, nullptr, 0, st))) return rc;
    vsrc = hdst;
  }
  if (need_v) {
    k7_resample_v<<<dim3(224, B), 672, 0, st>>>(vsrc, H, tv->d_bounds, tv->d_kk, tv->ksize, dst, swap_rb);
    VQA_LAUNCH_CHECK();
  }
  return B200VQA_OK;
}

extern "C" int b200vqa_resize_pil_pair(b200vqa_t* h, const uint8_t* src, int B, int H, int W, int swap_rb, uint8_t* dst_bilinear,
                                       uint8_t* dst_lanczos, void* stream) {
  if (!h || !src || !dst_bilinear || !dst_lanczos || B <= 0 || H <= 0 || W <= 0) return B200VQA_EINVAL;
  if (W == 224) {            // no horizontal pass to share
    int rc = b200vqa_resize_pil(h, src, B, H, W, B200VQA_FILTER_BILINEAR, swap_rb, dst_bilinear, stream);
    return rc ? rc : b200vqa_resize_pil(h, src, B, H, W, B200VQA_FILTER_LANCZOS, swap_rb, dst_lanczos, stream);
  }
  CtxScope scope(h);
  cudaStream_t st = as_stream(stream);
  const bool need_v = H != 224;
  ResizeTable *hb = nullptr, *hl = nullptr, *vb = nullptr, *vl = nullptr;
  int rc;
  if ((rc = get_table(h, W, B200VQA_FILTER_BILINEAR, &hb)) || (rc = get_table(h, W, B200VQA_FILTER_LANCZOS, &hl))) return rc;
  if (need_v && ((rc = get_table(h, H, B200VQA_FILTER_BILINEAR, &vb)) || (rc = get_table(h, H, B200VQA_FILTER_LANCZOS, &vl)))) return rc;
  uint8_t *tb = dst_bilinear, *tl = dst_lanczos;
  if (need_v) {
    const size_t one = (((size_t)B * H * 672) + 255) & ~(size_t)255;
    if ((rc = h->ws_resize.reserve(2 * one))) return rc;
    tb = static_cast<uint8_t*>(h->ws_resize.ptr); tl = tb + one;
  }
  if ((rc = launch_resample_h(src, B, H, W, hb, tb, need_v ? 0 : swap_rb, hl, tl, need_v ? 0 : swap_rb, st))) return rc;
  if (need_v) {
    k7_resample_v<<<dim3(224, B), 672, 0, st>>>(tb, H, vb->d_bounds, vb->d_kk, vb->ksize, dst_bilinear, swap_rb);
    VQA_LAUNCH_CHECK();
    k7_resample_v<<<dim3(224, B), 672, 0, st>>>(tl, H, vl->d_bounds, vl->d_kk, vl->ksize, dst_lanczos, swap_rb);
    VQA_LAUNCH_CHECK();
  }
  return B200VQA_OK;
}
