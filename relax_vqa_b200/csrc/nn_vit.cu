// A11 + A14 + A15: DINO ViT-B/16 forward and [mean, max, std] token pooling.
//   reference: src/extractor/visualise_vit_layer.py:81-260 (model), :447-500 (preprocess + forward),
//              src/main_fragment_pool.py:124-132 (pooling).
// Residual stream fp32; GEMM operands fp16 (tcgen05, fp32 accumulate); LayerNorm, softmax and
// pooling in fp32.  Attention is a fused per-(image, head) kernel (N = 197 fits on chip).
#include <map>
#include <string>
#include <vector>
#include "context.h"
#include "gemm_tcgen05.cuh"

namespace b200vqa {

constexpr int VD = 768, VT = 197, VP = 196, VH = 12, VHD = 64, VDEPTH = 12, VMLP = 3072;

struct Linear { int N = 0, K = 0; __half* w = nullptr; float* b = nullptr; CUtensorMap map_b, map_b128; };   // 256- / 128-row boxes
struct VitBlock { float *ln1_w, *ln1_b, *ln2_w, *ln2_b; Linear qkv, proj, fc1, fc2; };
struct ViTWeights {
  Linear patch;
  float *cls = nullptr, *pos = nullptr, *norm_w = nullptr, *norm_b = nullptr;
  VitBlock blocks[VDEPTH];
  std::vector<void*> allocs;
};

void free_vit(ViTWeights* v) {
  if (!v) return;
  for (void* p : v->allocs) cudaFree(p);
  delete v;
}

// ------------------------------------------------------------------------------- kernels
// ToTensor (x/255, no mean/std) + 16x16 patch extraction: [B][224][224][3] u8 -> fp16 [B*196][768],
// k = c*256 + i*16 + j (the flattening of the conv weight [768][3][16][16]).
__global__ void __launch_bounds__(256)
k9_patchify(const uint8_t* __restrict__ img, int is_bgr, __half* __restrict__ out) {
  const int p = blockIdx.x, b = blockIdx.y;              // patch index 0..195
  const int py = p / 14, px = p % 14;
  __half* o = out + ((size_t)b * VP + p) * VD;
  for (int k = threadIdx.x; k < VD; k += 256) {
    const int c = k >> 8, i = (k >> 4) & 15, j = k & 15;
    const uint8_t u = img[(((size_t)b * 224 + py * 16 + i) * 224 + px * 16 + j) * 3 + (is_bgr ? 2 - c : c)];
    o[k] = __float2half_rn((float)u / 255.0f);
  }
}

// tokens: x[b][0] = cls + pos[0]; x[b][1+p] = embed[b*196+p] + pos[1+p]   (fp32)
__global__ void __launch_bounds__(256)
k9_assemble_tokens(const float* __restrict__ embed, const float* __restrict__ cls, const float* __restrict__ pos,
                   float* __restrict__ x) {
  const int t = blockIdx.x, b = blockIdx.y;
  for (int c = threadIdx.x; c < VD; c += 256) {
    const float v = (t == 0) ? cls[c] : embed[((size_t)b * VP + t - 1) * VD + c];
    x[((size_t)b * VT + t) * VD + c] = v + pos[(size_t)t * VD + c];
  }
}

// LayerNorm (eps 1e-6) over 768 channels, one warp per token, fp32 in -> fp16 out
__global__ void __launch_bounds__(256)
k11_layernorm(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, __half* __restrict__ out, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * VD);
  float4 v[6];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) { v[i] = xr[lane + 32 * i]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / VD;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / VD + 1e-6f);
  __half* orow = out + (size_t)row * VD;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 ww = *reinterpret_cast<const float4*>(w + c), bb = *reinterpret_cast<const float4*>(bias + c);
    __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y);
    __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(orow + c) = pk;
  }
}

// ---- fused attention for one (image, head): S = QK^T * scale, softmax, O = PV.  fp16 operands on
// the warp-level tensor path (197 keys fit in registers/smem), fp32 softmax.
constexpr int AT_NP = 208, AT_LD = 72;
constexpr int AT_WARPS = 8;
constexpr int AT_SMEM = (2 * AT_NP + AT_WARPS * 16) * AT_LD * 2;
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

__global__ void __launch_bounds__(AT_WARPS * 32, 2)
k10_attention(const __half* __restrict__ qkv, __half* __restrict__ out, float scale) {
  extern __shared__ __align__(16) uint8_t at_smem[];
  typedef __half (*RowPtr)[AT_LD];
  RowPtr sK = reinterpret_cast<RowPtr>(at_smem);
  RowPtr sV = sK + AT_NP;
  typedef __half (*QPtr)[16][AT_LD];
  QPtr sQ = reinterpret_cast<QPtr>(sV + AT_NP);
  const int hh = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* base = qkv + (size_t)b * VT * (3 * VD) + hh * VHD;
  {
    constexpr int kIters = (AT_NP * 8 + AT_WARPS * 32 - 1) / (AT_WARPS * 32);      // 7: all loads issued before the first store
    uint4 kv[kIters], vv[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int i = tid + it * AT_WARPS * 32;
      const int row = i >> 3, ch = i & 7;
      kv[it] = make_uint4(0, 0, 0, 0); vv[it] = kv[it];
      if (i < AT_NP * 8 && row < VT) {
        kv[it] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)row * (3 * VD) + VD + ch * 8));
        vv[it] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)row * (3 * VD) + 2 * VD + ch * 8));
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int i = tid + it * AT_WARPS * 32;
      if (i < AT_NP * 8) {
        const int row = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(&sK[row][ch * 8]) = kv[it];
        *reinterpret_cast<uint4*>(&sV[row][ch * 8]) = vv[it];
      }
    }
  }
  __syncthreads();
  const int g = lane >> 2, t4 = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;             // scale * log2(e)
  for (int qb = warp; qb < AT_NP / 16; qb += AT_WARPS) {
    for (int i = lane; i < 16 * 8; i += 32) {
      const int r = i >> 3, ch = i & 7, row = qb * 16 + r;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (row < VT) q = *reinterpret_cast<const uint4*>(base + (size_t)row * (3 * VD) + ch * 8);
      *reinterpret_cast<uint4*>(&sQ[warp][r][ch * 8]) = q;
    }
    __syncwarp();
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qf[ks], &sQ[warp][lane & 15][ks * 16 + (lane >> 4) * 8]);
    // two key halves (112 + 96) with an online softmax: half the score registers, so two CTAs fit per SM
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int key0 = half * 112;
      constexpr int NT_MAX = 14;
      const int ntn = half == 0 ? 14 : 12;                // n-tiles (8 keys) in this half
      float s[NT_MAX][4];
#pragma unroll
      for (int nt = 0; nt < NT_MAX; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        if (nt < ntn) {
#pragma unroll
          for (int kp = 0; kp < 2; ++kp) {
            uint32_t kb[4];
            ldsm_x4(kb, &sK[key0 + nt * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
            mma16816(s[nt], qf[2 * kp], kb[0], kb[1]);
            mma16816(s[nt], qf[2 * kp + 1], kb[2], kb[3]);
          }
        }
      }
      // this thread owns rows g (regs 0,1) and g+8 (regs 2,3); keys >= 197 are masked
      float hm0 = -INFINITY, hm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT_MAX; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = key0 + nt * 8 + t4 * 2 + (e & 1);
          s[nt][e] = (nt < ntn && col < VT) ? s[nt][e] : -INFINITY;      // raw scores; scale > 0 is folded into the exponent
        }
        hm0 = fmaxf(hm0, fmaxf(s[nt][0], s[nt][1]));
        hm1 = fmaxf(hm1, fmaxf(s[nt][2], s[nt][3]));
      }
      hm0 = fmaxf(hm0, __shfl_xor_sync(0xffffffffu, hm0, 1)); hm0 = fmaxf(hm0, __shfl_xor_sync(0xffffffffu, hm0, 2));
      hm1 = fmaxf(hm1, __shfl_xor_sync(0xffffffffu, hm1, 1)); hm1 = fmaxf(hm1, __shfl_xor_sync(0xffffffffu, hm1, 2));
      const float n0 = fmaxf(m0, hm0), n1 = fmaxf(m1, hm1);
      const float c0 = exp2f((m0 - n0) * sl2), c1 = exp2f((m1 - n1) * sl2);     // rescale of the running sums (0 on the first half)
      m0 = n0; m1 = n1;
      const float mk0 = m0 * sl2, mk1 = m1 * sl2;
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      float h0 = 0.f, h1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT_MAX; ++nt) {
        s[nt][0] = exp2f(fmaf(s[nt][0], sl2, -mk0)); s[nt][1] = exp2f(fmaf(s[nt][1], sl2, -mk0));   // exp(scale * (s - m))
        s[nt][2] = exp2f(fmaf(s[nt][2], sl2, -mk1)); s[nt][3] = exp2f(fmaf(s[nt][3], sl2, -mk1));
        h0 += s[nt][0] + s[nt][1]; h1 += s[nt][2] + s[nt][3];
      }
      l0 += h0; l1 += h1;                                          // per-thread partial row sums; reduced across the quad at the end
#pragma unroll
      for (int kk = 0; kk < NT_MAX / 2; ++kk) {
        if (2 * kk < ntn) {
          uint32_t pf[4];
          pf[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
          pf[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
          pf[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pf[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t vb[4];
            ldsm_x4_t(vb, &sV[key0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]);
            mma16816(o[2 * np], pf, vb[0], vb[1]);
            mma16816(o[2 * np + 1], pf, vb[2], vb[3]);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float r0 = 1.f / l0, r1 = 1.f / l1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] *= r0; o[i][1] *= r0; o[i][2] *= r1; o[i][3] *= r1; }
    const int row0 = qb * 16 + g, row1 = row0 + 8;
    __half* ob = out + (size_t)b * VT * VD + hh * VHD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + t4 * 2;
      if (row0 < VT) *reinterpret_cast<uint32_t*>(ob + (size_t)row0 * VD + col) = pack_h2(o[nt][0], o[nt][1]);
      if (row1 < VT) *reinterpret_cast<uint32_t*>(ob + (size_t)row1 * VD + col) = pack_h2(o[nt][2], o[nt][3]);
    }
    __syncwarp();
  }
}

// ---- fused attention on the 5th-generation tensor cores (tcgen05 + TMEM), the default path.
//   reference op: Attention.forward, src/extractor/visualise_vit_layer.py:93-106 (q k^T * scale, softmax, @ v).
// One work item = one (image, head): Q, K, V are 197 x 64 fp16 slices of the qkv buffer, fetched by TMA (SWIZZLE_128B)
// as two 128-row query tiles and 208-row K / V tiles (rows past the image belong to the next image or are zero-filled
// past the end of the buffer: their scores are masked, their outputs never stored).  Per query tile:
//   S = Q K^T      tcgen05.mma kind::f16, M = 128, N = 208, K = 64: A and B K-major from shared memory -> TMEM cols [0, 208)
//   softmax        two threads per row (TMEM lane): warp set A takes keys [0, 112), set B keys [112, 208); two tcgen05.ld sweeps
//                  (row max; exp2 / row sum) with the next chunk's load in flight, partial max / sum exchanged through shared
//                  memory.  P goes back to TENSOR MEMORY as packed fp16: set A over the S columns it has consumed
//                  (cols [0, 56)), set B into the free columns [208, 256)
//   O = P V        tcgen05.mma with the A operand in tensor memory and V as an MN-major shared-memory operand (the
//                  [key][dim] tile TMA delivers IS the canonical MN-major SW128 layout), M = 128, N = 64, K = 208 -> cols [128, 192)
//   O / rowsum     tcgen05.ld, fp16, 32-byte vector stores (each warp set writes 64 bytes of the row).
// A CTA owns 256 TMEM columns and one shared-memory stage (84 KB; Q / K are refilled as soon as the second S has retired,
// V when the second P V has): two CTAs per SM alternate, one in its softmax (XU-bound: 197 x 208 exponentials per item)
// while the other multiplies.
constexpr int ATC_SM_WARPS = 8;                            // softmax / epilogue warps: two per TMEM lane quadrant
constexpr int ATC_THREADS = 64 + 32 * ATC_SM_WARPS;        // warp 0: TMA producer, warp 1: TMEM allocator + MMA issuer
constexpr int ATC_KV_ROWS = 208, ATC_SPLIT = 112;          // keys [0, 112) -> warp set A, [112, 208) -> set B
constexpr uint32_t ATC_Q_BYTES = 128 * VHD * 2, ATC_KV_BYTES = ATC_KV_ROWS * VHD * 2;
constexpr uint32_t ATC_STAGE_BYTES = 2 * ATC_Q_BYTES + 2 * ATC_KV_BYTES;          // 86,016
constexpr uint32_t ATC_TMEM_COLS = 256, ATC_O_COL = 128, ATC_PB_COL = 208;
constexpr int ATC_SMEM = ATC_STAGE_BYTES + 2 * 2 * 128 * 4 + 8 * 8 + 16 + 1024;

__device__ __forceinline__ void tcgen05_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(ATC_THREADS, 2)
k10_attention_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, __half* __restrict__ out,
                 int nimg, float scale) {
  extern __shared__ __align__(1024) uint8_t atc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ0 = smem; uint8_t* sQ1 = smem + ATC_Q_BYTES; uint8_t* sK = smem + 2 * ATC_Q_BYTES; uint8_t* sV = sK + ATC_KV_BYTES;
  float* s_max = reinterpret_cast<float*>(smem + ATC_STAGE_BYTES);            // [2][128] partial row maxima of the two warp sets
  float* s_sum = s_max + 256;                                                  // [2][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + 256);
  uint64_t* full_qk = bars; uint64_t* full_v = bars + 1; uint64_t* free_qk = bars + 2; uint64_t* free_v = bars + 3;
  uint64_t* s_full = bars + 4; uint64_t* p_full = bars + 5; uint64_t* o_full = bars + 6; uint64_t* s_free = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    mbar_init(full_qk, 1); mbar_init(full_v, 1); mbar_init(free_qk, 1); mbar_init(free_v, 1);
    mbar_init(s_full, 1); mbar_init(p_full, ATC_SM_WARPS); mbar_init(o_full, 1); mbar_init(s_free, ATC_SM_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ATC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int items = nimg * VH;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int b = item / VH, hh = item - b * VH;
        const uint32_t par = (uint32_t)(it & 1);
        mbar_wait(free_qk, par ^ 1u);                                         // both S products of the previous item have retired
        mbar_expect_tx(full_qk, 2 * ATC_Q_BYTES + ATC_KV_BYTES);
        tma_load_2d(&map_q, full_qk, sQ0, hh * VHD, b * VT);
        tma_load_2d(&map_q, full_qk, sQ1, hh * VHD, b * VT + 128);
        tma_load_2d(&map_kv, full_qk, sK, VD + hh * VHD, b * VT);
        mbar_wait(free_v, par ^ 1u);                                          // ... and its second P V
        mbar_expect_tx(full_v, ATC_KV_BYTES);
        tma_load_2d(&map_kv, full_v, sV, 2 * VD + hh * VHD, b * VT);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(ATC_KV_ROWS);                      // A, B K-major
      const uint32_t idesc_o = make_idesc(VHD) | (1u << 16);                  // B (V) MN-major
      const uint64_t dK = make_smem_desc(smem_u32(sK)), dV = make_smem_desc(smem_u32(sV));
      const uint64_t dQ0 = make_smem_desc(smem_u32(sQ0)), dQ1 = make_smem_desc(smem_u32(sQ1));
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        mbar_wait(full_qk, par);
        tcgen05_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t n = (uint32_t)(2 * it + t);
          const uint64_t dQ = t ? dQ1 : dQ0;
          mbar_wait(s_free, (n & 1u) ^ 1u);                                   // the previous tile's O has been read out of the region
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < VHD / 16; ++k)                                  // +32 B per K = 16 step inside the swizzled row
            tcgen05_mma_f16(tmem_base, dQ + (uint64_t)(2 * k), dK + (uint64_t)(2 * k), idesc_s, k != 0);
          tcgen05_commit(s_full);
          if (t == 1) tcgen05_commit(free_qk);                                // Q / K of the next item may land
          mbar_wait(p_full, n & 1u);
          if (t == 0) mbar_wait(full_v, par);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < ATC_KV_ROWS / 16; ++k) {                        // 16 keys = 8 TMEM columns of packed fp16 = 2048 B of V
            const uint32_t pcol = k < ATC_SPLIT / 16 ? (uint32_t)(8 * k) : ATC_PB_COL + (uint32_t)(8 * (k - ATC_SPLIT / 16));
            tcgen05_mma_f16_ts(tmem_base + ATC_O_COL, tmem_base + pcol, dV + (uint64_t)(128 * k), idesc_o, k != 0);
          }
          tcgen05_commit(o_full);
          if (t == 1) tcgen05_commit(free_v);
        }
      }
    }
  } else {
    // ================================ softmax / epilogue ==========================
    const int q = warp & 3;                                                   // TMEM lane quadrant of this warp
    const int half = (warp - 2) >> 2;                                         // 0: keys [0, 112) (set A), 1: keys [112, 208) (set B)
    const int rl = q * 32 + lane;                                             // row inside the tile = TMEM lane
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int col0 = half ? ATC_SPLIT : 0, nchunk = half ? (ATC_KV_ROWS - ATC_SPLIT) / 16 : ATC_SPLIT / 16;
    const uint32_t pcol0 = half ? ATC_PB_COL : 0u;
    const float sl2 = scale * 1.4426950408889634f;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = item / VH, hh = item - b * VH;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const uint32_t n = (uint32_t)(2 * it + t);
        const int row = t * 128 + rl;
        const bool warp_valid = t * 128 + q * 32 < VT;                        // warp-uniform
        mbar_wait(s_full, n & 1u);
        tcgen05_fence_after();
        float m = -INFINITY;
        if (warp_valid) {
          uint32_t v[2][16];
          tmem_ld16(taddr + col0, v[0]);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            if (c < nchunk) {
              tmem_ld_wait();
              if (c + 1 < nchunk) tmem_ld16(taddr + col0 + 16 * (c + 1), v[(c + 1) & 1]);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + 16 * c + j < VT) m = fmaxf(m, __uint_as_float(v[c & 1][j]));
            }
          }
          s_max[half * 128 + rl] = m;
        }
        named_bar_sync(1, 32 * ATC_SM_WARPS);
        float inv_l = 0.f;
        if (warp_valid) {
          m = fmaxf(m, s_max[(half ^ 1) * 128 + rl]);
          const float mk = m * sl2;
          float l = 0.f;
          uint32_t v[2][16];
          tmem_ld16(taddr + col0, v[0]);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            if (c < nchunk) {
              tmem_ld_wait();
              if (c + 1 < nchunk) tmem_ld16(taddr + col0 + 16 * (c + 1), v[(c + 1) & 1]);
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int k0 = col0 + 16 * c + 2 * j;
                const float p0 = k0 < VT ? ex2_approx(fmaf(__uint_as_float(v[c & 1][2 * j]), sl2, -mk)) : 0.f;
                const float p1 = k0 + 1 < VT ? ex2_approx(fmaf(__uint_as_float(v[c & 1][2 * j + 1]), sl2, -mk)) : 0.f;
                l += p0 + p1;
                pk[j] = pack_h2(p0, p1);
              }
              tmem_st8(taddr + pcol0 + 8 * c, pk);                           // set A: over S columns it has already consumed
            }
          }
          tmem_st_wait();
          s_sum[half * 128 + rl] = l;
        }
        tcgen05_fence_before();
        named_bar_sync(2, 32 * ATC_SM_WARPS);
        if (lane == 0) mbar_arrive(p_full);
        if (warp_valid) inv_l = 1.0f / (s_sum[rl] + s_sum[128 + rl]);
        mbar_wait(o_full, n & 1u);
        tcgen05_fence_after();
        if (warp_valid) {
          uint32_t o[32];
          tmem_ld32(taddr + ATC_O_COL + 32 * half, o);
          tmem_ld_wait();
          if (row < VT) {
            __half* op = out + ((size_t)b * VT + row) * VD + hh * VHD + 32 * half;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t h8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                h8[j] = pack_h2(__uint_as_float(o[16 * i + 2 * j]) * inv_l, __uint_as_float(o[16 * i + 2 * j + 1]) * inv_l);
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op + 16 * i), "r"(h8[0]), "r"(h8[1]), "r"(h8[2]),
                           "r"(h8[3]), "r"(h8[4]), "r"(h8[5]), "r"(h8[6]), "r"(h8[7]) : "memory");
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ATC_TMEM_COLS) : "memory");
  }
}

// ---- the same attention, software-pipelined inside one CTA per SM (A/B variant, attn_impl 2; MEASURED 96 us per 264-image
// launch vs 91 us for the two-CTA version above - with two tiles in flight either way, throughput is two tiles per
// tile-chain latency (~4.4 us: TMEM round trips of the two sweeps, barrier hops, P V, drain), and the single issuer thread
// adds ordering stalls; not the default): all 512 TMEM columns as TWO score / output
// regions and two shared-memory stages (one (image, head) item each).  The MMA issuer runs one query tile ahead: S of tile
// n + 1 goes into the other region while the softmax of tile n is in progress, so a softmax group (8 warps per region) finds
// its next scores ready when it has stored its outputs, and the exponentials of one region (XU-bound) overlap the P V, the
// output drain and the barrier round trips of the other.  Tiles alternate regions; the two query tiles of an item swap
// order on odd items, so both groups see the same mix of 128-row and 69-row tiles.
constexpr int AT2_GROUP_WARPS = 8, AT2_GROUPS = 2;
constexpr int AT2_THREADS = 64 + 32 * AT2_GROUP_WARPS * AT2_GROUPS;         // 576: warp 0 TMA, warp 1 TMEM alloc + MMA, 16 softmax warps
constexpr uint32_t AT2_TMEM_COLS = 512, AT2_REGION = 256;
constexpr int AT2_SMEM = 2 * ATC_STAGE_BYTES + AT2_GROUPS * 2 * 2 * 128 * 4 + 16 * 8 + 16 + 1024;

__global__ void __launch_bounds__(AT2_THREADS, 1)
k10_attention_tc2(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, __half* __restrict__ out,
                  int nimg, float scale) {
  extern __shared__ __align__(1024) uint8_t atc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~(uintptr_t)1023);
  float* s_stat = reinterpret_cast<float*>(smem + 2 * ATC_STAGE_BYTES);       // [group][max | sum][half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + AT2_GROUPS * 2 * 2 * 128);
  uint64_t* full_qk = bars; uint64_t* full_v = bars + 2; uint64_t* free_qk = bars + 4; uint64_t* free_v = bars + 6;      // per stage
  uint64_t* s_full = bars + 8; uint64_t* p_full = bars + 10; uint64_t* o_full = bars + 12; uint64_t* s_free = bars + 14;  // per region
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_qk[i], 1); mbar_init(&full_v[i], 1); mbar_init(&free_qk[i], 1); mbar_init(&free_v[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], AT2_GROUP_WARPS); mbar_init(&o_full[i], 1); mbar_init(&s_free[i], AT2_GROUP_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(AT2_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int items = nimg * VH;
  const int my_items = items > (int)blockIdx.x ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int ntiles = 2 * my_items;
  // tile n of this CTA: item it = n / 2 (global item blockIdx.x + it * gridDim.x), query tile t, region n & 1, stage it & 1
  auto tile_t = [](int n) { return ((n >> 1) & 1) ? 1 - (n & 1) : (n & 1); };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (int it = 0; it < my_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int b = item / VH, hh = item - b * VH;
        const int st = it & 1;
        const uint32_t par = (uint32_t)((it >> 1) & 1);
        uint8_t* base = smem + (size_t)st * ATC_STAGE_BYTES;
        mbar_wait(&free_qk[st], par ^ 1u);
        mbar_expect_tx(&full_qk[st], 2 * ATC_Q_BYTES + ATC_KV_BYTES);
        tma_load_2d(&map_q, &full_qk[st], base, hh * VHD, b * VT);
        tma_load_2d(&map_q, &full_qk[st], base + ATC_Q_BYTES, hh * VHD, b * VT + 128);
        tma_load_2d(&map_kv, &full_qk[st], base + 2 * ATC_Q_BYTES, VD + hh * VHD, b * VT);
        mbar_wait(&free_v[st], par ^ 1u);
        mbar_expect_tx(&full_v[st], ATC_KV_BYTES);
        tma_load_2d(&map_kv, &full_v[st], base + 2 * ATC_Q_BYTES + ATC_KV_BYTES, 2 * VD + hh * VHD, b * VT);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0 && ntiles > 0) {
      const uint32_t idesc_s = make_idesc(ATC_KV_ROWS);                      // A, B K-major
      const uint32_t idesc_o = make_idesc(VHD) | (1u << 16);                  // B (V) MN-major
      auto issue_s = [&](int n) {
        const int it = n >> 1, st = it & 1, r = n & 1, t = tile_t(n);
        if ((n & 1) == 0) { mbar_wait(&full_qk[st], (uint32_t)((it >> 1) & 1)); }
        mbar_wait(&s_free[r], (uint32_t)((n >> 1) & 1) ^ 1u);                 // O of the region's previous tile has been read out
        tcgen05_fence_after();
        const uint32_t sb = smem_u32(smem + (size_t)st * ATC_STAGE_BYTES);
        const uint64_t dQ = make_smem_desc(sb + (uint32_t)t * ATC_Q_BYTES), dK = make_smem_desc(sb + 2 * ATC_Q_BYTES);
#pragma unroll
        for (int k = 0; k < VHD / 16; ++k)
          tcgen05_mma_f16(tmem_base + (uint32_t)r * AT2_REGION, dQ + (uint64_t)(2 * k), dK + (uint64_t)(2 * k), idesc_s, k != 0);
        tcgen05_commit(&s_full[r]);
        if (n & 1) tcgen05_commit(&free_qk[st]);                              // both S products of the item are issued
      };
      issue_s(0);
      for (int n = 0; n < ntiles; ++n) {
        if (n + 1 < ntiles) issue_s(n + 1);                                   // one tile ahead, into the other region
        const int it = n >> 1, st = it & 1, r = n & 1;
        mbar_wait(&p_full[r], (uint32_t)((n >> 1) & 1));
        if ((n & 1) == 0) mbar_wait(&full_v[st], (uint32_t)((it >> 1) & 1));
        tcgen05_fence_after();
        const uint64_t dV = make_smem_desc(smem_u32(smem + (size_t)st * ATC_STAGE_BYTES) + 2 * ATC_Q_BYTES + ATC_KV_BYTES);
        const uint32_t rb = tmem_base + (uint32_t)r * AT2_REGION;
#pragma unroll
        for (int k = 0; k < ATC_KV_ROWS / 16; ++k) {
          const uint32_t pcol = k < ATC_SPLIT / 16 ? (uint32_t)(8 * k) : ATC_PB_COL + (uint32_t)(8 * (k - ATC_SPLIT / 16));
          tcgen05_mma_f16_ts(rb + ATC_O_COL, rb + pcol, dV + (uint64_t)(128 * k), idesc_o, k != 0);
        }
        tcgen05_commit(&o_full[r]);
        if (n & 1) tcgen05_commit(&free_v[st]);
      }
    }
  } else {
    // ================================ softmax / epilogue: group g owns region g ======================
    const int sw = warp - 2;                                                  // 0..15
    const int g = sw >> 3, half = (sw >> 2) & 1, q = warp & 3;                // region, key half (A / B), TMEM lane quadrant
    const int rl = q * 32 + lane;
    const uint32_t taddr = tmem_base + (uint32_t)g * AT2_REGION + ((uint32_t)(q * 32) << 16);
    const int col0 = half ? ATC_SPLIT : 0, nchunk = half ? (ATC_KV_ROWS - ATC_SPLIT) / 16 : ATC_SPLIT / 16;
    const uint32_t pcol0 = half ? ATC_PB_COL : 0u;
    float* s_max = s_stat + g * 512; float* s_sum = s_max + 256;
    const float sl2 = scale * 1.4426950408889634f;
    for (int n = g; n < ntiles; n += 2) {
      const int it = n >> 1, t = tile_t(n);
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / VH, hh = item - b * VH;
      const uint32_t par = (uint32_t)((n >> 1) & 1);
      const int row = t * 128 + rl;
      const bool warp_valid = t * 128 + q * 32 < VT;                          // warp-uniform
      mbar_wait(&s_full[g], par);
      tcgen05_fence_after();
      float m = -INFINITY;
      if (warp_valid) {
        uint32_t v[2][16];
        tmem_ld16(taddr + col0, v[0]);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          if (c < nchunk) {
            tmem_ld_wait();
            if (c + 1 < nchunk) tmem_ld16(taddr + col0 + 16 * (c + 1), v[(c + 1) & 1]);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + 16 * c + j < VT) m = fmaxf(m, __uint_as_float(v[c & 1][j]));
          }
        }
        s_max[half * 128 + rl] = m;
      }
      named_bar_sync(1 + 2 * g, 32 * AT2_GROUP_WARPS);
      float inv_l = 0.f;
      if (warp_valid) {
        m = fmaxf(m, s_max[(half ^ 1) * 128 + rl]);
        const float mk = m * sl2;
        float l = 0.f;
        uint32_t v[2][16];
        tmem_ld16(taddr + col0, v[0]);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          if (c < nchunk) {
            tmem_ld_wait();
            if (c + 1 < nchunk) tmem_ld16(taddr + col0 + 16 * (c + 1), v[(c + 1) & 1]);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k0 = col0 + 16 * c + 2 * j;
              const float p0 = k0 < VT ? ex2_approx(fmaf(__uint_as_float(v[c & 1][2 * j]), sl2, -mk)) : 0.f;
              const float p1 = k0 + 1 < VT ? ex2_approx(fmaf(__uint_as_float(v[c & 1][2 * j + 1]), sl2, -mk)) : 0.f;
              l += p0 + p1;
              pk[j] = pack_h2(p0, p1);
            }
            tmem_st8(taddr + pcol0 + 8 * c, pk);
          }
        }
        tmem_st_wait();
        s_sum[half * 128 + rl] = l;
      }
      tcgen05_fence_before();
      named_bar_sync(2 + 2 * g, 32 * AT2_GROUP_WARPS);
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (warp_valid) inv_l = 1.0f / (s_sum[rl] + s_sum[128 + rl]);
      mbar_wait(&o_full[g], par);
      tcgen05_fence_after();
      if (warp_valid) {
        uint32_t o[32];
        tmem_ld32(taddr + ATC_O_COL + 32 * half, o);
        tmem_ld_wait();
        if (row < VT) {
          __half* op = out + ((size_t)b * VT + row) * VD + hh * VHD + 32 * half;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t h8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              h8[j] = pack_h2(__uint_as_float(o[16 * i + 2 * j]) * inv_l, __uint_as_float(o[16 * i + 2 * j + 1]) * inv_l);
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op + 16 * i), "r"(h8[0]), "r"(h8[1]), "r"(h8[2]),
                         "r"(h8[3]), "r"(h8[4]), "r"(h8[5]), "r"(h8[6]), "r"(h8[7]) : "memory");
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[g]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(AT2_TMEM_COLS) : "memory");
  }
}

// SIMT check version of the attention (one thread per (query, head-dim) output)
__global__ void ref_attention(const __half* __restrict__ qkv, __half* __restrict__ out, float scale) {
  const int hh = blockIdx.x, b = blockIdx.y, q = blockIdx.z;
  __shared__ float p[VT];
  __shared__ float red[64];
  const __half* base = qkv + (size_t)b * VT * (3 * VD) + hh * VHD;
  const int t = threadIdx.x;      // 64 threads
  for (int k = t; k < VT; k += 64) {
    float acc = 0.f;
    for (int d = 0; d < VHD; ++d) acc = fmaf(__half2float(base[(size_t)q * 3 * VD + d]), __half2float(base[(size_t)k * 3 * VD + VD + d]), acc);
    p[k] = acc * scale;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int k = t; k < VT; k += 64) m = fmaxf(m, p[k]);
  red[t] = m; __syncthreads();
  for (int o = 32; o > 0; o >>= 1) { if (t < o) red[t] = fmaxf(red[t], red[t + o]); __syncthreads(); }
  m = red[0]; __syncthreads();
  float l = 0.f;
  for (int k = t; k < VT; k += 64) { p[k] = expf(p[k] - m); l += p[k]; }
  red[t] = l; __syncthreads();
  for (int o = 32; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
  l = red[0];
  float acc = 0.f;
  for (int k = 0; k < VT; ++k) acc = fmaf(p[k] / l, __half2float(base[(size_t)k * 3 * VD + 2 * VD + t]), acc);
  out[((size_t)b * VT + q) * VD + hh * VHD + t] = __float2half_rn(acc);
}

// final LayerNorm + pooling over the 196 patch tokens: out[b] = [mean | max | std(ddof=0)] (3 x 768)
__global__ void __launch_bounds__(768)
k11_final_norm_pool(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out) {
  __shared__ float s_mean[VP], s_rstd[VP];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xb = x + ((size_t)b * VT + 1) * VD;          // skip CLS
  for (int t = warp; t < VP; t += 24) {
    const float* xr = xb + (size_t)t * VD;
    float v[24], s = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / VD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) { s_mean[t] = mean; s_rstd[t] = rsqrtf(q / VD + 1e-6f); }
  }
  __syncthreads();
  const int c = tid;
  const float wc = w[c], bc = bias[c];
  float sum = 0.f, mx = -INFINITY;
  for (int t = 0; t < VP; ++t) {
    const float y = (xb[(size_t)t * VD + c] - s_mean[t]) * s_rstd[t] * wc + bc;
    sum += y; mx = fmaxf(mx, y);
  }
  const float mean = sum / VP;
  float sq = 0.f;
  for (int t = 0; t < VP; ++t) {
    const float y = (xb[(size_t)t * VD + c] - s_mean[t]) * s_rstd[t] * wc + bc;
    const float d = y - mean;
    sq += d * d;
  }
  float* o = out + (size_t)b * (3 * VD);
  o[c] = mean; o[VD + c] = mx; o[2 * VD + c] = sqrtf(sq / VP);
}

// final LayerNorm of the 196 patch tokens, un-pooled: what the reference's extract_features returns, norm(x)[:, 1:]
// (visualise_vit_layer.py:234-239, :492-500).  Debug / boundary-fidelity path only (b200vqa_vitb16_tokens).
__global__ void __launch_bounds__(256)
k11_final_norm_tokens(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int nimg) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nimg * VP) return;
  const int b = row / VP, t = row - b * VP;
  const float* xr = x + ((size_t)b * VT + 1 + t) * VD;
  float v[24], s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / VD;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; q += d * d; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / VD + 1e-6f);
  float* orow = out + (size_t)row * VD;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const int c = lane + 32 * i; orow[c] = (v[i] - mean) * rstd * w[c] + bias[c]; }
}

int vit_init_device_attrs() {
  VQA_CUDA(cudaFuncSetAttribute(k10_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
  VQA_CUDA(cudaFuncSetAttribute(k10_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
  VQA_CUDA(cudaFuncSetAttribute(k10_attention_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
  return B200VQA_OK;
}

// ------------------------------------------------------------------------- weight loading
typedef std::map<std::string, std::pair<const float*, int64_t>> TensorMap;

static int upload_f32(ViTWeights* vw, const TensorMap& t, const std::string& name, int64_t numel, float** dev) {
  auto it = t.find(name);
  if (it == t.end() || it->second.second != numel) return B200VQA_EINVAL;
  VQA_CUDA(cudaMalloc((void**)dev, numel * sizeof(float)));
  vw->allocs.push_back(*dev);
  VQA_CUDA(cudaMemcpy(*dev, it->second.first, numel * sizeof(float), cudaMemcpyHostToDevice));
  return B200VQA_OK;
}

static int load_linear(ViTWeights* vw, const TensorMap& t, const std::string& name, int N, int K, Linear* lin) {
  auto it = t.find(name + ".weight");
  if (it == t.end() || it->second.second != (int64_t)N * K) return B200VQA_EINVAL;
  std::vector<__half> w((size_t)N * K);
  for (size_t i = 0; i < w.size(); ++i) w[i] = __float2half_rn(it->second.first[i]);
  VQA_CUDA(cudaMalloc((void**)&lin->w, w.size() * sizeof(__half)));
  vw->allocs.push_back(lin->w);
  VQA_CUDA(cudaMemcpy(lin->w, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
  int rc = upload_f32(vw, t, name + ".bias", N, &lin->b);
  if (rc) return rc;
  lin->N = N; lin->K = K;
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, strides[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {GEMM_BK, 256}, box128[2] = {GEMM_BK, 128};
  rc = make_tmap_f16(&lin->map_b128, lin->w, 2, dims, strides, box128, nullptr);
  if (rc) return rc;
  return make_tmap_f16(&lin->map_b, lin->w, 2, dims, strides, box, nullptr);
}

// out[M][N] = act(A[M][K] W^T + b) (+ residual)
static int run_linear(b200vqa_ctx* h, const Linear& lin, const __half* A, int M, void* out, int out_is_f32, int act,
                      const float* residual, cudaStream_t st) {
  if (h->gemm_impl == 1) {
    return launch_ref_gemm_rowmajor(A, lin.w, lin.b, residual, out, M, lin.N, lin.K, lin.N, act, out_is_f32, st);
  }
  CUtensorMap ma;
  uint64_t dims[2] = {(uint64_t)lin.K, (uint64_t)M}, strides[1] = {(uint64_t)lin.K * 2};
  uint32_t box[2] = {GEMM_BK, GEMM_BM};
  int rc = make_tmap_f16(&ma, A, 2, dims, strides, box, nullptr);
  if (rc) return rc;
  if (h->gemm_impl == 0 && lin.N % 256 == 0 && lin.K % GEMM_BK == 0)      // default: SM-pair kernel (cta_group::2)
    return launch_gemm_2cta(ma, lin.map_b128, M, lin.N, lin.K, lin.b, residual, out, out_is_f32, act, gemm_grid_sms(h), st);
  GemmParams p{};
  p.block_n = 256;
  p.m_tiles = cdiv(M, GEMM_BM); p.n_tiles = cdiv(lin.N, 256);
  p.k_blocks_per_tap = lin.K / GEMM_BK; p.taps_r = p.taps_s = 1;
  p.stages = pick_stages(256);
  p.epi = EPI_ROW; p.act = act; p.M = M; p.N = lin.N; p.ldo = lin.N; p.out_is_f32 = out_is_f32;
  p.bias = lin.b; p.residual = residual; p.out = out;
  return launch_gemm(ma, lin.map_b, p, gemm_grid_sms(h), st);
}

}  // namespace b200vqa

using namespace b200vqa;

extern "C" int b200vqa_load_vitb16(b200vqa_t* h, int n, const char* const* names, const float* const* h_ptrs,
                                   const int64_t* numels) {
  if (!h || n <= 0 || !names || !h_ptrs || !numels) return B200VQA_EINVAL;
  VQA_CUDA(cudaSetDevice(h->device));
  TensorMap t;
  for (int i = 0; i < n; ++i) t[names[i]] = std::make_pair(h_ptrs[i], numels[i]);
  ViTWeights* vw = new ViTWeights();
  int rc = load_linear(vw, t, "patch_embed.proj", VD, VD, &vw->patch);
  if (!rc) rc = upload_f32(vw, t, "cls_token", VD, &vw->cls);
  if (!rc) rc = upload_f32(vw, t, "pos_embed", (int64_t)VT * VD, &vw->pos);
  if (!rc) rc = upload_f32(vw, t, "norm.weight", VD, &vw->norm_w);
  if (!rc) rc = upload_f32(vw, t, "norm.bias", VD, &vw->norm_b);
  for (int i = 0; i < VDEPTH && !rc; ++i) {
    const std::string p = "blocks." + std::to_string(i);
    VitBlock& bk = vw->blocks[i];
    rc = upload_f32(vw, t, p + ".norm1.weight", VD, &bk.ln1_w);
    if (!rc) rc = upload_f32(vw, t, p + ".norm1.bias", VD, &bk.ln1_b);
    if (!rc) rc = upload_f32(vw, t, p + ".norm2.weight", VD, &bk.ln2_w);
    if (!rc) rc = upload_f32(vw, t, p + ".norm2.bias", VD, &bk.ln2_b);
    if (!rc) rc = load_linear(vw, t, p + ".attn.qkv", 3 * VD, VD, &bk.qkv);
    if (!rc) rc = load_linear(vw, t, p + ".attn.proj", VD, VD, &bk.proj);
    if (!rc) rc = load_linear(vw, t, p + ".mlp.fc1", VMLP, VD, &bk.fc1);
    if (!rc) rc = load_linear(vw, t, p + ".mlp.fc2", VD, VMLP, &bk.fc2);
  }
  if (rc) { free_vit(vw); return rc; }
  free_vit(h->vit);
  h->vit = vw;
  return B200VQA_OK;
}

// out: pooled [B][2304] or null; tokens: un-pooled final-LayerNorm patch tokens [B][196][768] or null
static int vit_forward(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* out, float* tokens, void* stream) {
  if (!h || !img || (!out && !tokens) || B <= 0) return B200VQA_EINVAL;
  if (!h->vit) return B200VQA_ENOTLOADED;
  CtxScope scope(h);
  cudaStream_t st = as_stream(stream);
  const ViTWeights& vw = *h->vit;
  // images per pass: as many as the workspace allows (512 images = 1.3 GB).  Each of the 49 tcgen05 launches of a pass
  // costs ~10 us of prologue / ramp / tail (tools/gemm_waves.py), and with >= 8 waves per launch the partial last wave
  // matters less than that fixed cost (a 264-image step: 9 / 25 / 34 waves for the 3 / 9 / 12 column tiles).
  const int CH = 512;
  const int passes = cdiv(B, CH);
  const int nb = cdiv(B, passes);                      // balanced passes
  const size_t M = (size_t)nb * VT;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  const size_t o_x = carve(M * VD * 4), o_h = carve(M * VD * 2), o_big = carve(M * VMLP * 2), o_att = carve(M * VD * 2);
  const size_t o_emb = carve((size_t)nb * VP * VD * 4);
  int rc = h->ws_vit.reserve(off);
  if (rc) return rc;
  uint8_t* ws = static_cast<uint8_t*>(h->ws_vit.ptr);
  float* x = (float*)(ws + o_x); __half* hbuf = (__half*)(ws + o_h); __half* big = (__half*)(ws + o_big);
  __half* att = (__half*)(ws + o_att); float* emb = (float*)(ws + o_emb);
  for (int b0 = 0; b0 < B; b0 += nb) {
    const int n = (B - b0) < nb ? (B - b0) : nb;
    const int m = n * VT;
    k9_patchify<<<dim3(VP, n), 256, 0, st>>>(img + (size_t)b0 * 224 * 224 * 3, is_bgr, hbuf);
    VQA_LAUNCH_CHECK();
    if ((rc = run_linear(h, vw.patch, hbuf, n * VP, emb, 1, ACT_NONE, nullptr, st))) return rc;
    k9_assemble_tokens<<<dim3(VT, n), 256, 0, st>>>(emb, vw.cls, vw.pos, x);
    VQA_LAUNCH_CHECK();
    for (int l = 0; l < VDEPTH; ++l) {
      const VitBlock& bk = vw.blocks[l];
      k11_layernorm<<<cdiv(m, 8), 256, 0, st>>>(x, bk.ln1_w, bk.ln1_b, hbuf, m);
      VQA_LAUNCH_CHECK();
      if ((rc = run_linear(h, bk.qkv, hbuf, m, big, 0, ACT_NONE, nullptr, st))) return rc;
      if (h->gemm_impl == 1) ref_attention<<<dim3(VH, n, VT), 64, 0, st>>>(big, att, 0.125f);
      else if (h->attn_impl == 1) k10_attention<<<dim3(VH, n), AT_WARPS * 32, AT_SMEM, st>>>(big, att, 0.125f);
      else {
        // Q / K / V are column slices of the qkv buffer [m][2304]: one descriptor with 128-row boxes, one with 208-row boxes
        CUtensorMap mq, mkv;
        uint64_t dims[2] = {(uint64_t)(3 * VD), (uint64_t)m}, strides[1] = {(uint64_t)(3 * VD) * 2};
        uint32_t boxq[2] = {VHD, 128}, boxkv[2] = {VHD, ATC_KV_ROWS};
        if ((rc = make_tmap_f16(&mq, big, 2, dims, strides, boxq, nullptr))) return rc;
        if ((rc = make_tmap_f16(&mkv, big, 2, dims, strides, boxkv, nullptr))) return rc;
        const int items = n * VH;
        if (h->attn_impl == 2) {          // A/B: one CTA per SM, two TMEM regions, S issued one tile ahead
          k10_attention_tc2<<<items < h->sm_count ? items : h->sm_count, AT2_THREADS, AT2_SMEM, st>>>(mq, mkv, att, n, 0.125f);
        } else {                          // default: two independent CTAs per SM
          k10_attention_tc<<<items < 2 * h->sm_count ? items : 2 * h->sm_count, ATC_THREADS, ATC_SMEM, st>>>(mq, mkv, att, n, 0.125f);
        }
      }
      VQA_LAUNCH_CHECK();
      if ((rc = run_linear(h, bk.proj, att, m, x, 1, ACT_NONE, x, st))) return rc;
      k11_layernorm<<<cdiv(m, 8), 256, 0, st>>>(x, bk.ln2_w, bk.ln2_b, hbuf, m);
      VQA_LAUNCH_CHECK();
      if ((rc = run_linear(h, bk.fc1, hbuf, m, big, 0, ACT_GELU, nullptr, st))) return rc;
      if ((rc = run_linear(h, bk.fc2, big, m, x, 1, ACT_NONE, x, st))) return rc;
    }
    if (out) {
      k11_final_norm_pool<<<n, 768, 0, st>>>(x, vw.norm_w, vw.norm_b, out + (size_t)b0 * B200VQA_VIT_POOL);
      VQA_LAUNCH_CHECK();
    }
    if (tokens) {
      k11_final_norm_tokens<<<cdiv(n * VP, 8), 256, 0, st>>>(x, vw.norm_w, vw.norm_b, tokens + (size_t)b0 * VP * VD, n);
      VQA_LAUNCH_CHECK();
    }
  }
  return B200VQA_OK;
}

extern "C" int b200vqa_vitb16_features(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* out, void* stream) {
  if (!out) return B200VQA_EINVAL;
  return vit_forward(h, img, B, is_bgr, out, nullptr, stream);
}

extern "C" int b200vqa_vitb16_tokens(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* tokens, void* stream) {
  if (!tokens) return B200VQA_EINVAL;
  return vit_forward(h, img, B, is_bgr, nullptr, tokens, stream);
}
