// A10 + A12 + A13: torchvision ResNet-50 (v1.5) forward on 224x224 images with the 15-hook
// layer-stack average pooling fused into the epilogue of the producing convolution.
//   reference: src/extractor/visualise_resnet.py:62-109 (15 hooks, one full forward each),
//              src/extractor/visualise_resnet_layer.py:62-102 (avgpool hook),
//              src/main_fragment_layerstack.py:131-149 (pooling).
// Activations: fp16 NHWC.  Weights: fp16 [Cout][R][S][Cin] (K-major).  BatchNorm (eval) stays in
// fp32 as a per-channel scale/shift applied to the fp32 accumulator.  Every convolution except the
// 7x7 stem is an implicit GEMM fed by 4-D TMA boxes (no im2col in HBM); the stem uses an explicit
// [pixels][192] patch matrix because Cin = 3.
#include <map>
#include <string>
#include <vector>
#include <stdlib.h>
#include "context.h"
#include "gemm_tcgen05.cuh"

namespace b200vqa {

struct ConvW {
  int Cin = 0, Cout = 0, R = 1, S = 1, stride = 1, pad = 0, K = 0;   // K = padded GEMM depth
  __half* w = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
  CUtensorMap map_a;
};

struct Bottleneck { ConvW c1, c2, c3, ds; bool has_ds = false; };

struct ResNetWeights {
  ConvW stem;
  __half* eye = nullptr;          // [128][128] one-hot A operand for the tensor-core residual add
  float* ones64 = nullptr;        // scale / shift of the raw conv1 hook (b200vqa_resnet50_maps)
  float* zeros64 = nullptr;
  CUtensorMap map_eye;
  std::vector<Bottleneck> blocks;
  std::vector<void*> allocs;
};

void free_resnet(ResNetWeights* r) {
  if (!r) return;
  for (void* p : r->allocs) cudaFree(p);
  delete r;
}

// ------------------------------------------------------------------------------- kernels
// Stem patch matrix: uint8 HWC image -> fp16 [B*112*112][192].  K layout: k = r*24 + s*3 + c for the 7 kernel
// rows (21 taps + 3 zero pads each, so every row is three 16-byte stores), rows 7 (k = 168..191) are zero.
// ToTensor + Normalize applied here; zero padding is applied after normalisation, as in conv2d.
__global__ void __launch_bounds__(256)
k8_stem_im2col(const uint8_t* __restrict__ img, int is_bgr, __half* __restrict__ out, size_t npix) {
  // per-block table of the 3 x 256 possible normalised values (same float expression as ToTensor + Normalize, rounded to
  // fp16 once): the 42 IEEE divisions per thread become 3
  __shared__ uint16_t lut[3][256];
  {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const __half hv = __float2half_rn(((float)threadIdx.x / 255.0f - mean[c]) / stdv[c]);
      lut[c][threadIdx.x] = *reinterpret_cast<const uint16_t*>(&hv);
    }
  }
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;     // one thread = one kernel row of one output pixel
  const size_t pix = idx >> 3;
  if (pix >= npix) return;
  const int r = (int)(idx & 7);
  const int x = pix % 112, y = (pix / 112) % 112;
  const size_t n = pix / (112 * 112);
  uint32_t hv[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) hv[i] = 0u;
  const int iy = y * 2 + r - 3;
  if (r < 7 && iy >= 0 && iy < 224) {
    const uint8_t* row = img + (n * 224 + iy) * (224 * 3);
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int ix = x * 2 + s - 3;
      if (ix >= 0 && ix < 224) {
#pragma unroll
        for (int c = 0; c < 3; ++c) hv[s * 3 + c] = lut[c][row[ix * 3 + (is_bgr ? 2 - c : c)]];
      }
    }
  }
  uint32_t pk[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) pk[i] = hv[2 * i] | (hv[2 * i + 1] << 16);
  uint4* o = reinterpret_cast<uint4*>(out + pix * 192 + r * 24);
  o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  o[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
}

// 3x3 / stride 2 / pad 1 max pooling, NHWC fp16, 8 channels (16 B) per thread
__global__ void __launch_bounds__(256)
k8_maxpool(const __half* __restrict__ in, __half* __restrict__ out, int Nimg, int Hin, int Win, int C) {
  const int Hout = Hin / 2, Wout = Win / 2, cv = C / 8;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)Nimg * Hout * Wout * cv;
  if (idx >= total) return;
  const int c8 = idx % cv;
  size_t p = idx / cv;
  const int x = p % Wout; p /= Wout;
  const int y = p % Hout;
  const size_t n = p / Hout;
  __half2 m[4];
  const __half2 ninf = __half2half2(__ushort_as_half(0xFC00));
  for (int i = 0; i < 4; ++i) m[i] = ninf;
  for (int r = 0; r < 3; ++r) {
    const int iy = y * 2 + r - 1;
    if (iy < 0 || iy >= Hin) continue;
    for (int s = 0; s < 3; ++s) {
      const int ix = x * 2 + s - 1;
      if (ix < 0 || ix >= Win) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(in + ((n * Hin + iy) * Win + ix) * C + c8 * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
      for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], h[i]);
    }
  }
  *reinterpret_cast<uint4*>(out + ((n * Hout + y) * Wout + x) * C + c8 * 8) = *reinterpret_cast<uint4*>(m);
}

// boundary-fidelity path (b200vqa_resnet50_maps): one hooked activation, fp16 NHWC -> fp32 CHW at `dst` of each image's record
__global__ void __launch_bounds__(256)
k8_dump_map(const __half* __restrict__ act, int HW, int C, float* __restrict__ maps, size_t per_img, size_t offset) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (p < HW && c < C) ? __half2float(act[((size_t)n * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < C && p < HW) maps[(size_t)n * per_img + offset + (size_t)c * HW + p] = tile[tx][r];
  }
}

struct HookDesc { const float* partial; int tiles_y, C, offset; float inv_hw; };
struct HookTable { HookDesc h[15]; };

// layer-stack finish: out[n][offset + c] = (sum over tiles of partial[n][t][c]) / (H*W), fixed order
__global__ void __launch_bounds__(256)
k8_gap_finish(HookTable tab, float* __restrict__ stack, int stack_ld, int Nimg) {
  const HookDesc d = tab.h[blockIdx.y];
  const int n = blockIdx.z;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d.C; c += gridDim.x * blockDim.x) {
    const float* p = d.partial + (size_t)n * d.tiles_y * d.C + c;
    float s = 0.f;
    for (int t = 0; t < d.tiles_y; ++t) s += p[(size_t)t * d.C];
    stack[(size_t)n * stack_ld + d.offset + c] = s * d.inv_hw;
  }
}

// A13: avgpool vector (2048) + [mean, max, std(ddof=0)] over its channels -> [B][2051]
__global__ void __launch_bounds__(256)
k8_pool_stats(const float* __restrict__ partial, int slots, float inv_hw, float* __restrict__ pool) {
  __shared__ float red[256];
  __shared__ float s_mean;
  const int n = blockIdx.x, t = threadIdx.x;
  float v[8], sum = 0.f, mx = -INFINITY;
  for (int i = 0; i < 8; ++i) {
    float a = 0.f;
    for (int k = 0; k < slots; ++k) a += partial[((size_t)n * slots + k) * 2048 + t + i * 256];
    v[i] = a * inv_hw;
    pool[(size_t)n * 2051 + t + i * 256] = v[i];
    sum += v[i]; mx = fmaxf(mx, v[i]);
  }
  red[t] = sum; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
  if (t == 0) s_mean = red[0] / 2048.f;
  __syncthreads();
  const float mean = s_mean;
  __syncthreads();
  red[t] = mx; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) red[t] = fmaxf(red[t], red[t + o]); __syncthreads(); }
  const float gmax = red[0];
  __syncthreads();
  float sq = 0.f;
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; sq += d * d; }
  red[t] = sq; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
  if (t == 0) {
    pool[(size_t)n * 2051 + 2048] = mean;
    pool[(size_t)n * 2051 + 2049] = gmax;
    pool[(size_t)n * 2051 + 2050] = sqrtf(red[0] / 2048.f);
  }
}

// ------------------------------------------------------------------------- weight loading
static int upload(ResNetWeights* rw, const void* host, size_t bytes, void** dev) {
  VQA_CUDA(cudaMalloc(dev, bytes));
  rw->allocs.push_back(*dev);
  VQA_CUDA(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
  return B200VQA_OK;
}

typedef std::map<std::string, std::pair<const float*, int64_t>> TensorMap;

static int load_conv(ResNetWeights* rw, const TensorMap& t, const std::string& conv, const std::string& bn, int Cin, int Cout,
                     int R, int stride, int pad, ConvW* out) {
  auto wi = t.find(conv + ".weight");
  auto g = t.find(bn + ".weight"), b = t.find(bn + ".bias"), m = t.find(bn + ".running_mean"), v = t.find(bn + ".running_var");
  if (wi == t.end() || g == t.end() || b == t.end() || m == t.end() || v == t.end()) return B200VQA_EINVAL;
  if (wi->second.second != (int64_t)Cout * Cin * R * R || g->second.second != Cout) return B200VQA_EINVAL;
  const bool stem = (Cin == 3);
  const int K = stem ? 192 : R * R * Cin;
  std::vector<float> sc(Cout), sh(Cout);
  for (int o = 0; o < Cout; ++o) {
    const float inv = 1.0f / sqrtf(v->second.first[o] + 1e-5f);
    sc[o] = g->second.first[o] * inv;
    sh[o] = b->second.first[o] - m->second.first[o] * sc[o];
  }
  // BN scale is folded into the fp16 weights (in fp32, one rounding) so that the residual identity can be
  // accumulated by the tensor core next to the convolution.  The stem keeps its scale in the epilogue
  // because its hook is the raw, pre-BN convolution output.
  std::vector<__half> w((size_t)Cout * K, __float2half(0.f));
  const float* src = wi->second.first;                     // [Cout][Cin][R][S]
  for (int o = 0; o < Cout; ++o) {
    const float fold = stem ? 1.0f : sc[o];
    for (int c = 0; c < Cin; ++c)
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < R; ++s)
          w[(size_t)o * K + (stem ? (size_t)r * 24 + s * 3 + c : ((size_t)r * R + s) * Cin + c)] =
              __float2half_rn(src[(((size_t)o * Cin + c) * R + r) * R + s] * fold);
    if (!stem) sc[o] = 1.0f;
  }
  out->Cin = Cin; out->Cout = Cout; out->R = out->S = R; out->stride = stride; out->pad = pad; out->K = K;
  int rc;
  if ((rc = upload(rw, w.data(), w.size() * sizeof(__half), (void**)&out->w))) return rc;
  if ((rc = upload(rw, sc.data(), Cout * sizeof(float), (void**)&out->scale))) return rc;
  if ((rc = upload(rw, sh.data(), Cout * sizeof(float), (void**)&out->shift))) return rc;
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)Cout}, strides[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {GEMM_BK, GEMM_BM};
  return make_tmap_f16(&out->map_a, out->w, 2, dims, strides, box, nullptr);
}

// ----------------------------------------------------------------------------- conv launch
struct Geo { int tw, th, tn; };
static Geo geo_for(int Wout) {
  switch (Wout) {
    case 112: return {112, 2, 1};
    case 56: return {56, 4, 1};
    case 28: return {28, 4, 1};
    case 14: return {14, 16, 1};
    default: return {8, 8, 4};      // 7x7: boxes overhang by one row / column (masked in the epilogue)
  }
}

// in: NHWC fp16 [N][Hin][Win][Cin] (or the stem patch matrix); out: [N][Hout][Wout][Cout]
static int run_conv(b200vqa_ctx* h, const ConvW& cw, const __half* in, int Nimg, int Hin, int Win, __half* out,
                    const __half* identity, int relu, float* gap_partial, int gap_raw, cudaStream_t st) {
  const ResNetWeights& rw = *h->resnet;
  const bool stem = (cw.Cin == 3);
  const int Hout = stem ? 112 : (Hin + 2 * cw.pad - cw.R) / cw.stride + 1;
  const int Wout = stem ? 112 : (Win + 2 * cw.pad - cw.S) / cw.stride + 1;
  if (h->gemm_impl == 1) {      // SIMT check path; the stem patch matrix is a 1x1 conv over 192 "channels"
    if (stem) return launch_ref_conv_nhwc(in, cw.w, cw.scale, cw.shift, identity, out, gap_partial, gap_raw, Nimg, Hout, Wout, 192,
                                          Hout, Wout, cw.Cout, 1, 1, 1, 0, relu ? ACT_RELU : ACT_NONE, st);
    return launch_ref_conv_nhwc(in, cw.w, cw.scale, cw.shift, identity, out, gap_partial, gap_raw, Nimg, Hin, Win, cw.Cin, Hout, Wout,
                                cw.Cout, cw.R, cw.S, cw.stride, cw.pad, relu ? ACT_RELU : ACT_NONE, st);
  }
  Geo g = geo_for(Wout);
  GemmParams p{};
  // 3x3 / stride 1 convolutions on maps of 14 pixels and wider (11 of the 16 conv2 layers): the haloed activation tile is
  // fetched ONCE per channel block and the nine taps read it as shifted shared-memory views; the per-tap boxes of the generic
  // path fetch it nine times through L2.  MEASURED at 264 images: layer1 conv2 133 vs 174 us, layer2 76 vs 112, layer3 70 vs 80
  // (ResNet pass 4.80 vs 5.12 ms); 7 x 7 maps stay on the generic path (B200VQA_CONV_NOHALO=1 switches all of them back).
  static const bool no_halo = getenv("B200VQA_CONV_NOHALO") != nullptr;
  static const int halo_min_w = getenv("B200VQA_CONV_HALO_MINW") ? atoi(getenv("B200VQA_CONV_HALO_MINW")) : 14;
  const bool halo = !no_halo && !h->conv_no_halo && !stem && cw.R == 3 && cw.S == 3 && cw.stride == 1 && cw.pad == 1 && !identity && !gap_partial && Wout >= halo_min_w && Wout <= 62 &&
                    Hout == Wout && cw.Cin % GEMM_BK == 0;
  if (halo) {
    g.tw = Wout; g.tn = 1;
    g.th = 256 / (Wout + 2);                        // rows per tile: th * (Wout + 2) accumulator columns
    if (g.th > Hout) g.th = Hout;
    p.halo = 1;
    p.halo_pitch = Wout + 2;
    p.halo_rows_img = (g.th + 2) * p.halo_pitch;
    p.halo_bytes = (int)(((size_t)p.halo_rows_img * 128 + 18 * 128 + 1023) & ~(size_t)1023);      // + the over-read of the rounded-up N
  }
  p.block_n = halo ? ((g.th * p.halo_pitch + 15) & ~15) : g.tw * g.th * g.tn;
  p.m_tiles = cdiv(cw.Cout, GEMM_BM);
  p.tiles_y = cdiv(Hout, g.th);
  p.n_tiles = p.tiles_y * cdiv(Nimg, g.tn);
  p.taps_r = stem ? 1 : cw.R; p.taps_s = stem ? 1 : cw.S;
  p.k_blocks_per_tap = stem ? 3 : cw.Cin / GEMM_BK;
  p.stages = pick_stages(p.block_n, false);       // EPI_CONV: no transposition tiles in shared memory
  if (halo) {
    p.stages = (int)((227 * 1024 - gemm_smem_fixed(false) - 2 * (size_t)p.halo_bytes) / (GEMM_BM * GEMM_BK * 2));
    if (p.stages > GEMM_MAX_STAGES) p.stages = GEMM_MAX_STAGES;
  }
  p.b_is_conv = stem ? 0 : 1;
  p.conv_stride = cw.stride; p.conv_pad = cw.pad;
  p.tw = g.tw; p.th = g.th; p.tn = g.tn;
  p.Hout = Hout; p.Wout = Wout; p.Nimg = Nimg;
  p.epi = EPI_CONV; p.act = relu ? ACT_RELU : ACT_NONE; p.M = cw.Cout;
  p.out = out; p.scale = cw.scale; p.shift = cw.shift; p.idt_blocks = identity ? 2 : 0;
  p.gap_partial = gap_partial; p.gap_raw = gap_raw;
  CUtensorMap mb;
  int rc;
  if (stem) {
    uint64_t dims[2] = {192, (uint64_t)Nimg * 112 * 112}, strides[1] = {192 * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)p.block_n};
    rc = make_tmap_f16(&mb, in, 2, dims, strides, box, nullptr);
  } else {
    uint64_t dims[4] = {(uint64_t)cw.Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)Nimg};
    uint64_t strides[3] = {(uint64_t)cw.Cin * 2, (uint64_t)Win * cw.Cin * 2, (uint64_t)Hin * Win * cw.Cin * 2};
    uint32_t box[4] = {GEMM_BK, (uint32_t)(g.tw * cw.stride), (uint32_t)(g.th * cw.stride), (uint32_t)g.tn};
    uint32_t es[4] = {1, (uint32_t)cw.stride, (uint32_t)cw.stride, 1};
    if (halo) { box[1] = (uint32_t)p.halo_pitch; box[2] = (uint32_t)(g.th + 2); }
    rc = make_tmap_f16(&mb, in, 4, dims, strides, box, es);
  }
  if (rc) return rc;
  if (identity) {
    if (cw.Cout % GEMM_BM) return B200VQA_EINVAL;
    CUtensorMap mi;
    uint64_t dims[4] = {(uint64_t)cw.Cout, (uint64_t)Wout, (uint64_t)Hout, (uint64_t)Nimg};
    uint64_t strides[3] = {(uint64_t)cw.Cout * 2, (uint64_t)Wout * cw.Cout * 2, (uint64_t)Hout * Wout * cw.Cout * 2};
    uint32_t box[4] = {GEMM_BK, (uint32_t)g.tw, (uint32_t)g.th, (uint32_t)g.tn};
    if ((rc = make_tmap_f16(&mi, identity, 4, dims, strides, box, nullptr))) return rc;
    return launch_gemm(cw.map_a, mb, p, gemm_grid_sms(h), st, &rw.map_eye, &mi);
  }
  return launch_gemm(cw.map_a, mb, p, gemm_grid_sms(h), st);
}

static const int kStagePlanes[4] = {64, 128, 256, 512};
static const int kStageBlocks[4] = {3, 4, 6, 3};
static const int kStageHooks[4] = {3, 4, 4, 3};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace b200vqa

using namespace b200vqa;

extern "C" int b200vqa_load_resnet50(b200vqa_t* h, int n, const char* const* names, const float* const* h_ptrs,
                                     const int64_t* numels) {
  if (!h || n <= 0 || !names || !h_ptrs || !numels) return B200VQA_EINVAL;
  VQA_CUDA(cudaSetDevice(h->device));
  TensorMap t;
  for (int i = 0; i < n; ++i) t[names[i]] = std::make_pair(h_ptrs[i], numels[i]);
  ResNetWeights* rw = new ResNetWeights();
  int rc = load_conv(rw, t, "conv1", "bn1", 3, 64, 7, 2, 3, &rw->stem);
  int inplanes = 64;
  for (int s = 0; s < 4 && !rc; ++s) {
    const int planes = kStagePlanes[s];
    for (int b = 0; b < kStageBlocks[s] && !rc; ++b) {
      const std::string p = "layer" + std::to_string(s + 1) + "." + std::to_string(b);
      const int stride = (b == 0 && s > 0) ? 2 : 1;
      Bottleneck bk;
      rc = load_conv(rw, t, p + ".conv1", p + ".bn1", inplanes, planes, 1, 1, 0, &bk.c1);
      if (!rc) rc = load_conv(rw, t, p + ".conv2", p + ".bn2", planes, planes, 3, stride, 1, &bk.c2);
      if (!rc) rc = load_conv(rw, t, p + ".conv3", p + ".bn3", planes, planes * 4, 1, 1, 0, &bk.c3);
      if (!rc && b == 0) {
        bk.has_ds = true;
        rc = load_conv(rw, t, p + ".downsample.0", p + ".downsample.1", inplanes, planes * 4, 1, stride, 0, &bk.ds);
      }
      rw->blocks.push_back(bk);
      inplanes = planes * 4;
    }
  }
  if (!rc) {
    std::vector<__half> eye(128 * 128, __float2half(0.f));
    for (int i = 0; i < 128; ++i) eye[i * 128 + i] = __float2half(1.0f);
    rc = upload(rw, eye.data(), eye.size() * sizeof(__half), (void**)&rw->eye);
    uint64_t dims[2] = {128, 128}, strides[1] = {128 * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    if (!rc) rc = make_tmap_f16(&rw->map_eye, rw->eye, 2, dims, strides, box, nullptr);
    const std::vector<float> one(64, 1.0f), zero(64, 0.0f);
    if (!rc) rc = upload(rw, one.data(), 64 * sizeof(float), (void**)&rw->ones64);
    if (!rc) rc = upload(rw, zero.data(), 64 * sizeof(float), (void**)&rw->zeros64);
  }
  if (rc) { free_resnet(rw); return rc; }
  free_resnet(h->resnet);
  h->resnet = rw;
  return B200VQA_OK;
}

// maps: [B][B200VQA_RESNET_MAP_FLOATS] fp32, the 15 hooked activations of each image as (C,H,W) arrays in hook order, or null
static int resnet_forward(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* stack, float* pool, float* maps,
                          void* stream) {
  if (!h || !img || B <= 0 || (!stack && !pool && !maps)) return B200VQA_EINVAL;
  if (!h->resnet) return B200VQA_ENOTLOADED;
  CtxScope scope(h);
  // The raw-map boundary path (every activation compared element by element with the fp32 reference, max over ~200 K fp16 values
  // per map at the 1e-2 bound) keeps the convolution order it was validated with; the fused path's pooled features have 3-4x margin.
  struct HaloGuard { b200vqa_ctx* c; int old; ~HaloGuard() { c->conv_no_halo = old; } } halo_guard{h, h->conv_no_halo};
  h->conv_no_halo = maps != nullptr;
  cudaStream_t st = as_stream(stream);
  const ResNetWeights& rw = *h->resnet;
  const int CH = 512;                                  // images per pass (bounds the workspace at ~4.6 GB): one pass per step in practice,
                                                       // so the ~10 us fixed cost of each of the 53 launches is paid once
  const int passes = cdiv(B, CH);
  const int nb = cdiv(B, passes);                      // balanced passes: no tiny, latency-bound remainder pass
  // workspace carve-up (bytes), all fp16 NHWC unless noted
  const size_t sz_col = (size_t)nb * 12544 * 192 * 2, sz_c1 = (size_t)nb * 12544 * 64 * 2, sz_big = (size_t)nb * 3136 * 256 * 2;
  const size_t sz_mid = (size_t)nb * 3136 * 128 * 2;
  const size_t partial_per_img = (size_t)GEMM_EPI_GROUPS * (56 * 64 + 3 * 14 * 256 + 4 * 7 * 512 + 4 * 1024 + 3 * 2048);
  const size_t sz_part = align_up((size_t)nb * partial_per_img * sizeof(float), 256);
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t o_col = carve(sz_col), o_c1 = carve(sz_c1), o_x = carve(sz_big), o_y = carve(sz_big), o_idt = carve(sz_big);
  const size_t o_a = carve(sz_mid), o_b = carve(sz_mid), o_part = carve(sz_part);
  int rc = h->ws_resnet.reserve(off);
  if (rc) return rc;
  uint8_t* ws = static_cast<uint8_t*>(h->ws_resnet.ptr);
  __half* col = (__half*)(ws + o_col); __half* c1 = (__half*)(ws + o_c1);
  __half* bx = (__half*)(ws + o_x); __half* by = (__half*)(ws + o_y); __half* bidt = (__half*)(ws + o_idt);
  __half* ba = (__half*)(ws + o_a); __half* bb = (__half*)(ws + o_b);
  float* part = (float*)(ws + o_part);

  for (int b0 = 0; b0 < B; b0 += nb) {
    const int n = (B - b0) < nb ? (B - b0) : nb;
    const uint8_t* im = img + (size_t)b0 * 224 * 224 * 3;
    HookTable tab{};
    int nh = 0, feat_off = 0;
    float* pp = part;
    auto add_hook = [&](int tiles_y, int C, int hw) {
      tab.h[nh] = HookDesc{pp, tiles_y, C, feat_off, 1.0f / (float)hw};
      float* r = pp; pp += (size_t)n * tiles_y * C; feat_off += C; ++nh; return r;
    };
    // stem: im2col -> conv1 (hooked raw, pre-BN) -> BN+ReLU -> maxpool
    const size_t npix = (size_t)n * 12544;
    k8_stem_im2col<<<(unsigned)((npix * 8 + 255) / 256), 256, 0, st>>>(im, is_bgr, col, npix);
    VQA_LAUNCH_CHECK();
    float* gp = add_hook(h->gemm_impl == 1 ? 1 : 56 * GEMM_EPI_GROUPS, 64, 12544);
    size_t map_off = 0;
    const size_t per_img = B200VQA_RESNET_MAP_FLOATS;
    auto dump = [&](const __half* act, int HW, int C) {
      k8_dump_map<<<dim3(cdiv(HW, 32), cdiv(C, 32), n), 256, 0, st>>>(act, HW, C, maps + (size_t)b0 * per_img, per_img, map_off);
      count_launch();
      map_off += (size_t)HW * C;
    };
    if (maps) {
      // the conv1 hook is the raw convolution output (before bn1 / relu): one extra stem launch with scale 1, shift 0, no ReLU
      ConvW raw = rw.stem;
      raw.scale = rw.ones64; raw.shift = rw.zeros64;
      if ((rc = run_conv(h, raw, col, n, 224, 224, c1, nullptr, 0, nullptr, 0, st))) return rc;
      dump(c1, 12544, 64);
    }
    if ((rc = run_conv(h, rw.stem, col, n, 224, 224, c1, nullptr, 1, gp, 1, st))) return rc;
    {
      const size_t total = (size_t)n * 56 * 56 * 8;
      k8_maxpool<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c1, bx, n, 112, 112, 64);
      VQA_LAUNCH_CHECK();
    }
    __half* x = bx; __half* y = by;
    int Hc = 56, bi = 0;
    for (int s = 0; s < 4; ++s) {
      for (int b = 0; b < kStageBlocks[s]; ++b, ++bi) {
        const Bottleneck& bk = rw.blocks[bi];
        const int stride = bk.c2.stride, Ho = Hc / stride;
        if ((rc = run_conv(h, bk.c1, x, n, Hc, Hc, ba, nullptr, 1, nullptr, 0, st))) return rc;
        if ((rc = run_conv(h, bk.c2, ba, n, Hc, Hc, bb, nullptr, 1, nullptr, 0, st))) return rc;
        const __half* idt = x;
        if (bk.has_ds) {
          if ((rc = run_conv(h, bk.ds, x, n, Hc, Hc, bidt, nullptr, 0, nullptr, 0, st))) return rc;
          idt = bidt;
        }
        float* g = nullptr;
        if (b < kStageHooks[s]) {
          const Geo ge = geo_for(Ho);
          g = add_hook(h->gemm_impl == 1 ? 1 : cdiv(Ho, ge.th) * GEMM_EPI_GROUPS, bk.c3.Cout, Ho * Ho);
        }
        if ((rc = run_conv(h, bk.c3, bb, n, Ho, Ho, y, idt, 1, g, 0, st))) return rc;
        if (maps && b < kStageHooks[s]) dump(y, Ho * Ho, bk.c3.Cout);
        __half* t = x; x = y; y = t;
        Hc = Ho;
      }
    }
    if (stack) {
      k8_gap_finish<<<dim3(8, 15, n), 256, 0, st>>>(tab, stack + (size_t)b0 * B200VQA_RESNET_STACK, B200VQA_RESNET_STACK, n);
      VQA_LAUNCH_CHECK();
    }
    if (pool) {
      // layer4[2] is the last hook; its partial slots summed in order are the avgpool numerators
      k8_pool_stats<<<n, 256, 0, st>>>(tab.h[14].partial, tab.h[14].tiles_y, tab.h[14].inv_hw, pool + (size_t)b0 * B200VQA_RESNET_POOL);
      VQA_LAUNCH_CHECK();
    }
  }
  return B200VQA_OK;
}

extern "C" int b200vqa_resnet50_features(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* stack, float* pool,
                                         void* stream) {
  if (!stack && !pool) return B200VQA_EINVAL;
  return resnet_forward(h, img, B, is_bgr, stack, pool, nullptr, stream);
}

extern "C" int b200vqa_resnet50_maps(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* maps, void* stream) {
  if (!maps) return B200VQA_EINVAL;
  return resnet_forward(h, img, B, is_bgr, nullptr, nullptr, maps, stream);
}
