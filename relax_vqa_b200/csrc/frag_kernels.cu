// Integer stages of the fragment path: A1-A4, A7 (sums), A8 of SURVEY.md 8(a).
// All results are bit-exact with the reference; HBM-bound uint8 kernels with 16-byte
// vector accesses when W % 16 == 0 (row pitch 3W is then a multiple of 16 bytes).
#include "common.cuh"

namespace b200vqa {

// cv2 BGR2GRAY fixed point: (3735 B + 19235 G + 9798 R + 2^14) >> 15
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
  return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

__device__ __forceinline__ void gray16(const uint4& a, const uint4& b, const uint4& c, uint4& out) {
  // 48 bytes = 16 BGR pixels -> 16 gray bytes
  uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  uint32_t g[4] = {0, 0, 0, 0};
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    int o = 3 * p;
    uint32_t bb = (w[o >> 2] >> (8 * (o & 3))) & 0xff;
    uint32_t gg = (w[(o + 1) >> 2] >> (8 * ((o + 1) & 3))) & 0xff;
    uint32_t rr = (w[(o + 2) >> 2] >> (8 * ((o + 2) & 3))) & 0xff;
    g[p >> 2] |= gray_of(bb, gg, rr) << (8 * (p & 3));
  }
  out = make_uint4(g[0], g[1], g[2], g[3]);
}

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc) { return __vsadu4(a, b) + acc; }

// ---------------------------------------------------------------------------------------
// K1 fast path (W % 16 == 0): thread = (patch column px, row r in the 16-row band).
// block = 32 patch columns x 16 rows.  Each thread owns 48 contiguous bytes (16 pixels).
template <bool kResidual, bool kGray, bool kPair>
__global__ void __launch_bounds__(512)
k1_absdiff_patchsum_vec(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ next, int H, int W,
                        uint8_t* __restrict__ residual, uint32_t* __restrict__ sums,
                        uint8_t* __restrict__ gray0, uint8_t* __restrict__ gray1) {
  __shared__ uint32_t part[16][33];
  const int gw = W >> 4, gh = H >> 4;
  const int px = blockIdx.x * 32 + threadIdx.x;
  const int r = threadIdx.y;
  const int y = blockIdx.y * 16 + r;
  const size_t img = (size_t)blockIdx.z * H * W;
  uint32_t s = 0;
  if (px < gw && y < H) {
    const size_t off = (img + (size_t)y * W) * 3 + (size_t)px * 48;
    const uint4* pa = reinterpret_cast<const uint4*>(frame + off);
    uint4 a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
    if (kPair) {
      const uint4* pb = reinterpret_cast<const uint4*>(next + off);
      uint4 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
      s = sad4(a0.x, b0.x, s); s = sad4(a0.y, b0.y, s); s = sad4(a0.z, b0.z, s); s = sad4(a0.w, b0.w, s);
      s = sad4(a1.x, b1.x, s); s = sad4(a1.y, b1.y, s); s = sad4(a1.z, b1.z, s); s = sad4(a1.w, b1.w, s);
      s = sad4(a2.x, b2.x, s); s = sad4(a2.y, b2.y, s); s = sad4(a2.z, b2.z, s); s = sad4(a2.w, b2.w, s);
      if (kResidual) {
        uint4* pr = reinterpret_cast<uint4*>(residual + off);
        pr[0] = make_uint4(__vabsdiffu4(a0.x, b0.x), __vabsdiffu4(a0.y, b0.y), __vabsdiffu4(a0.z, b0.z), __vabsdiffu4(a0.w, b0.w));
        pr[1] = make_uint4(__vabsdiffu4(a1.x, b1.x), __vabsdiffu4(a1.y, b1.y), __vabsdiffu4(a1.z, b1.z), __vabsdiffu4(a1.w, b1.w));
        pr[2] = make_uint4(__vabsdiffu4(a2.x, b2.x), __vabsdiffu4(a2.y, b2.y), __vabsdiffu4(a2.z, b2.z), __vabsdiffu4(a2.w, b2.w));
      }
      if (kGray) {
        uint4 g;
        const size_t goff = img + (size_t)y * W + (size_t)px * 16;
        gray16(a0, a1, a2, g); *reinterpret_cast<uint4*>(gray0 + goff) = g;
        gray16(b0, b1, b2, g); *reinterpret_cast<uint4*>(gray1 + goff) = g;
      }
    } else {   // plain patch sums of one image (flow colour image)
      s = sad4(a0.x, 0, s); s = sad4(a0.y, 0, s); s = sad4(a0.z, 0, s); s = sad4(a0.w, 0, s);
      s = sad4(a1.x, 0, s); s = sad4(a1.y, 0, s); s = sad4(a1.z, 0, s); s = sad4(a1.w, 0, s);
      s = sad4(a2.x, 0, s); s = sad4(a2.y, 0, s); s = sad4(a2.z, 0, s); s = sad4(a2.w, 0, s);
    }
  }
  part[r][threadIdx.x] = s;
  __syncthreads();
  if (r == 0 && px < gw && blockIdx.y < gh) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += part[i][threadIdx.x];
    sums[((size_t)blockIdx.z * gh + blockIdx.y) * gw + px] = t;
  }
}

// K1 generic path (any W): one thread per pixel, patch sums by shared-memory atomics.
// block = 16 rows x 64 pixels (4 patch columns).
template <bool kPair>
__global__ void __launch_bounds__(1024)
k1_absdiff_patchsum_generic(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ next, int H, int W,
                            uint8_t* __restrict__ residual, uint32_t* __restrict__ sums,
                            uint8_t* __restrict__ gray0, uint8_t* __restrict__ gray1) {
  __shared__ uint32_t part[4];
  const int gw = W >> 4, gh = H >> 4;
  if (threadIdx.y == 0 && threadIdx.x < 4) part[threadIdx.x] = 0;
  __syncthreads();
  const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 16 + threadIdx.y;
  if (x < W && y < H) {
    const size_t p = ((size_t)blockIdx.z * H + y) * W + x;
    uint32_t b0 = frame[p * 3], g0 = frame[p * 3 + 1], r0 = frame[p * 3 + 2];
    uint32_t s;
    if (kPair) {
      uint32_t b1 = next[p * 3], g1 = next[p * 3 + 1], r1 = next[p * 3 + 2];
      uint32_t db = b0 > b1 ? b0 - b1 : b1 - b0, dg = g0 > g1 ? g0 - g1 : g1 - g0, dr = r0 > r1 ? r0 - r1 : r1 - r0;
      s = db + dg + dr;
      if (residual) { residual[p * 3] = db; residual[p * 3 + 1] = dg; residual[p * 3 + 2] = dr; }
      if (gray0) gray0[p] = gray_of(b0, g0, r0);
      if (gray1) gray1[p] = gray_of(b1, g1, r1);
    } else {
      s = b0 + g0 + r0;
    }
    if ((x >> 4) < gw && blockIdx.y < gh) atomicAdd(&part[threadIdx.x >> 4], s);
  }
  __syncthreads();
  if (threadIdx.y == 0 && threadIdx.x < 4) {
    int px = blockIdx.x * 4 + threadIdx.x;
    if (px < gw && blockIdx.y < gh) sums[((size_t)blockIdx.z * gh + blockIdx.y) * gw + px] = part[threadIdx.x];
  }
}

// ---------------------------------------------------------------------------------------
// K2: exact top-k selection, value descending / flat index ascending, raster-ordered output.
// One 1024-thread block per image.  Values are < 2^18 (16*16*3*255), so a bit-wise binary
// search of the threshold needs at most 32 counting passes over <= 32400 values (L1-resident).
__global__ void __launch_bounds__(1024)
k2_topk(const uint32_t* __restrict__ sums, int G, int gw, int top_n, int32_t* __restrict__ pos, int32_t* __restrict__ count) {
  __shared__ uint32_t s_cnt[32];
  __shared__ uint32_t s_scan_gt[1024], s_scan_eq[1024];
  __shared__ uint32_t s_total;
  const uint32_t* v = sums + (size_t)blockIdx.x * G;
  int32_t* out = pos + (size_t)blockIdx.x * top_n * 2;
  const int tid = threadIdx.x;
  const int n_sel = G < top_n ? G : top_n;
  if (tid == 0) count[blockIdx.x] = n_sel;
  for (int i = tid; i < top_n * 2; i += 1024) out[i] = -1;
  __syncthreads();
  // T = largest value such that count(v >= T) >= n_sel  (the n_sel-th largest value)
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    uint32_t c = 0;
    for (int i = tid; i < G; i += 1024) c += (v[i] >= cand);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = c;
    __syncthreads();
    if (tid < 32) {
      uint32_t t = s_cnt[tid];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (tid == 0) s_total = t;
    }
    __syncthreads();
    if (s_total >= (uint32_t)n_sel) T = cand;
    __syncthreads();
  }
  // ordered compaction: contiguous chunk per thread
  const int chunk = (G + 1023) / 1024;
  const int lo = tid * chunk, hi = min(G, lo + chunk);
  uint32_t cgt = 0, ceq = 0;
  for (int i = lo; i < hi; ++i) { uint32_t x = v[i]; cgt += (x > T); ceq += (x == T); }
  s_scan_gt[tid] = cgt; s_scan_eq[tid] = ceq;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {       // Hillis-Steele inclusive scan
    uint32_t a = 0, b = 0;
    if (tid >= off) { a = s_scan_gt[tid - off]; b = s_scan_eq[tid - off]; }
    __syncthreads();
    s_scan_gt[tid] += a; s_scan_eq[tid] += b;
    __syncthreads();
  }
  const uint32_t total_gt = s_scan_gt[1023];
  const uint32_t need_eq = (uint32_t)n_sel - total_gt;          // how many ties to take, in index order
  uint32_t rgt = s_scan_gt[tid] - cgt, req = s_scan_eq[tid] - ceq;   // exclusive prefixes
  for (int i = lo; i < hi; ++i) {
    uint32_t x = v[i];
    bool sel = (x > T) || (x == T && req < need_eq);
    if (sel) {
      uint32_t o = rgt + (req < need_eq ? req : need_eq);
      out[2 * o] = i / gw;
      out[2 * o + 1] = i % gw;
    }
    rgt += (x > T); req += (x == T);
  }
}

// ---------------------------------------------------------------------------------------
// K3: gather 16x16x3 patches into the 14x14 canvas.  grid = (top_n, B); 48 threads x 16B per
// row when vectorisable, byte path otherwise.
template <bool kVec>
__global__ void __launch_bounds__(256)
k3_gather(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ next, int H, int W,
          const int32_t* __restrict__ pos, const int32_t* __restrict__ count, int top_n,
          uint8_t* __restrict__ ori_frag, uint8_t* __restrict__ diff_frag) {
  const int j = blockIdx.x, b = blockIdx.y;
  const int cy = j / 14, cx = j % 14;
  const bool live = j < count[b];
  const int py = live ? pos[((size_t)b * top_n + j) * 2] : 0;
  const int px = live ? pos[((size_t)b * top_n + j) * 2 + 1] : 0;
  const size_t src0 = ((size_t)b * H + (size_t)py * 16) * W * 3 + (size_t)px * 48;
  const size_t dst0 = ((size_t)b * 224 + (size_t)cy * 16) * 224 * 3 + (size_t)cx * 48;
  if (kVec) {
    const int t = threadIdx.x;            // 48 threads: row = t / 3, 16B chunk = t % 3
    if (t >= 48) return;
    const int r = t / 3, c = t % 3;
    const size_t so = src0 + (size_t)r * W * 3 + c * 16, d = dst0 + (size_t)r * 224 * 3 + c * 16;
    uint4 a = make_uint4(0, 0, 0, 0), df = a;
    if (live) {
      a = __ldg(reinterpret_cast<const uint4*>(frame + so));
      if (diff_frag) {
        uint4 n = __ldg(reinterpret_cast<const uint4*>(next + so));
        df = make_uint4(__vabsdiffu4(a.x, n.x), __vabsdiffu4(a.y, n.y), __vabsdiffu4(a.z, n.z), __vabsdiffu4(a.w, n.w));
      }
    }
    if (ori_frag) *reinterpret_cast<uint4*>(ori_frag + d) = a;
    if (diff_frag) *reinterpret_cast<uint4*>(diff_frag + d) = df;
  } else {
    for (int i = threadIdx.x; i < 16 * 48; i += blockDim.x) {
      const int r = i / 48, c = i % 48;
      const size_t so = src0 + (size_t)r * W * 3 + c, d = dst0 + (size_t)r * 224 * 3 + c;
      uint8_t a = 0, df = 0;
      if (live) {
        a = frame[so];
        if (diff_frag) { uint8_t n = next[so]; df = a > n ? a - n : n - a; }
      }
      if (ori_frag) ori_frag[d] = a;
      if (diff_frag) diff_frag[d] = df;
    }
  }
}

// K6: (a + b) / 2 rounded half to even, 16 bytes per thread (tail: bytes)
__device__ __forceinline__ uint32_t avg_rne4(uint32_t a, uint32_t b) {
  // floor average per byte, then +1 where the sum is odd and the floor is odd
  uint32_t fl = (a & b) + (((a ^ b) & 0xfefefefeu) >> 1);
  uint32_t odd = (a ^ b) & 0x01010101u;
  return fl + (odd & fl);
}
__global__ void k6_merge(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, uint8_t* __restrict__ out) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= n && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0)) {
    uint4 x = *reinterpret_cast<const uint4*>(a + i), y = *reinterpret_cast<const uint4*>(b + i);
    *reinterpret_cast<uint4*>(out + i) = make_uint4(avg_rne4(x.x, y.x), avg_rne4(x.y, y.y), avg_rne4(x.z, y.z), avg_rne4(x.w, y.w));
  } else {
    for (size_t k = i; k < n && k < i + 16; ++k) {
      uint32_t s = (uint32_t)a[k] + b[k], h = s >> 1;
      out[k] = (uint8_t)(h + ((s & 1) & (h & 1)));
    }
  }
}

}  // namespace b200vqa

using namespace b200vqa;

extern "C" int b200vqa_absdiff_patchsum_u8(const uint8_t* frame, const uint8_t* next, int B, int H, int W,
                                           uint8_t* residual, uint32_t* sums, uint8_t* gray_frame,
                                           uint8_t* gray_next, void* stream) {
  if (!frame || !next || !sums || B <= 0 || H <= 0 || W <= 0) return B200VQA_EINVAL;
  if ((gray_frame == nullptr) != (gray_next == nullptr)) return B200VQA_EINVAL;
  cudaStream_t st = as_stream(stream);
  const bool aligned = (W % 16 == 0) && ((((uintptr_t)frame | (uintptr_t)next | (uintptr_t)residual |
                                           (uintptr_t)gray_frame | (uintptr_t)gray_next) & 15) == 0);
  if (aligned) {
    dim3 grid(cdiv(W / 16, 32), cdiv(H, 16), B), block(32, 16);
    if (residual && gray_frame) k1_absdiff_patchsum_vec<true, true, true><<<grid, block, 0, st>>>(frame, next, H, W, residual, sums, gray_frame, gray_next);
    else if (residual) k1_absdiff_patchsum_vec<true, false, true><<<grid, block, 0, st>>>(frame, next, H, W, residual, sums, nullptr, nullptr);
    else if (gray_frame) k1_absdiff_patchsum_vec<false, true, true><<<grid, block, 0, st>>>(frame, next, H, W, nullptr, sums, gray_frame, gray_next);
    else k1_absdiff_patchsum_vec<false, false, true><<<grid, block, 0, st>>>(frame, next, H, W, nullptr, sums, nullptr, nullptr);
  } else {
    dim3 grid(cdiv(W, 64), cdiv(H, 16), B), block(64, 16);
    k1_absdiff_patchsum_generic<true><<<grid, block, 0, st>>>(frame, next, H, W, residual, sums, gray_frame, gray_next);
  }
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

extern "C" int b200vqa_patchsum_u8(const uint8_t* img, int B, int H, int W, uint32_t* sums, void* stream) {
  if (!img || !sums || B <= 0 || H <= 0 || W <= 0) return B200VQA_EINVAL;
  cudaStream_t st = as_stream(stream);
  if (W % 16 == 0 && (((uintptr_t)img & 15) == 0)) {
    dim3 grid(cdiv(W / 16, 32), cdiv(H, 16), B), block(32, 16);
    k1_absdiff_patchsum_vec<false, false, false><<<grid, block, 0, st>>>(img, nullptr, H, W, nullptr, sums, nullptr, nullptr);
  } else {
    dim3 grid(cdiv(W, 64), cdiv(H, 16), B), block(64, 16);
    k1_absdiff_patchsum_generic<false><<<grid, block, 0, st>>>(img, nullptr, H, W, nullptr, sums, nullptr, nullptr);
  }
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

extern "C" int b200vqa_topk_patches(const uint32_t* sums, int B, int gh, int gw, int top_n, int32_t* pos,
                                    int32_t* count, void* stream) {
  if (!sums || !pos || !count || B <= 0 || gh < 0 || gw < 0 || top_n <= 0) return B200VQA_EINVAL;
  k2_topk<<<B, 1024, 0, as_stream(stream)>>>(sums, gh * gw, gw > 0 ? gw : 1, top_n, pos, count);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

extern "C" int b200vqa_gather_fragments(const uint8_t* frame, const uint8_t* next, int B, int H, int W,
                                        const int32_t* pos, const int32_t* count, int top_n,
                                        uint8_t* ori_frag, uint8_t* diff_frag, void* stream) {
  if (!frame || !pos || !count || B <= 0 || top_n != B200VQA_TOPN || (diff_frag && !next)) return B200VQA_EINVAL;
  if (!ori_frag && !diff_frag) return B200VQA_OK;
  const bool vec = (W % 16 == 0) && ((((uintptr_t)frame | (uintptr_t)next | (uintptr_t)ori_frag | (uintptr_t)diff_frag) & 15) == 0);
  dim3 grid(top_n, B);
  if (vec) k3_gather<true><<<grid, 64, 0, as_stream(stream)>>>(frame, next, H, W, pos, count, top_n, ori_frag, diff_frag);
  else k3_gather<false><<<grid, 256, 0, as_stream(stream)>>>(frame, next, H, W, pos, count, top_n, ori_frag, diff_frag);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

extern "C" int b200vqa_merge_fragments(const uint8_t* a, const uint8_t* b, size_t nbytes, uint8_t* out, void* stream) {
  if (!a || !b || !out) return B200VQA_EINVAL;
  if (nbytes == 0) return B200VQA_OK;
  size_t threads = (nbytes + 15) / 16;
  k6_merge<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(a, b, nbytes, out);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}
