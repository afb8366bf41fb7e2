// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 64] (fp16, smem, K-major, SW128) * B[BN x 64]^T
//
// Roles (512 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-15 = epilogue (TMEM -> registers -> global; three warps per
// TMEM lane quadrant, each taking a share of the accumulator columns).
// Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty double buffer
// (MMA <-> epilogue), static round-robin tile scheduler (tile = blockIdx.x + i*gridDim.x).
//
// Operand A is always a 2-D K-major matrix [rowsA][K_total].  Operand B is either a 2-D
// K-major matrix (linear layers) or an NHWC activation tensor addressed through a 4-D TMA
// box (C, W, H, N): for a convolution tap (r, s) the box is fetched at (c0, x0*stride+s-pad,
// y0*stride+r-pad, n0); out-of-bounds elements are zero-filled by TMA, which implements the
// padding, and `elementStrides` implements stride 2.  The K loop runs over taps x (Cin/64).
//
// Epilogues:
//   EPI_ROW  (linear layers; A rows = tokens, B rows = output features):
//       out[m][n] = act(acc + bias[n]) (+ residual[m][n]), fp16 or fp32 row-major
//   EPI_CONV (convolutions; A rows = output channels, B rows = output pixels):
//       v = acc*scale[c] + shift[c]; ReLU; out[pix][c] fp16 NHWC.  The bottleneck's residual identity is
//       added by the tensor core itself: two extra K-blocks whose A tile is a one-hot (identity) matrix
//       and whose B tile is the identity tensor fetched by TMA (BN scale is folded into the weights, so
//       acc = scale*conv + identity exactly); the epilogue therefore issues no global loads.
//       optional per-(image, tile, channel) partial sums of v (or of the raw acc) for the
//       fused layer-stack global average pooling - written, not atomically added, so the
//       reduction order is fixed and results are bit-reproducible.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace b200vqa {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 12;              // three warps per TMEM lane quadrant
constexpr int GEMM_EPI_GROUPS = GEMM_EPI_WARPS / 4;
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_MAX_STAGES = 8;
constexpr int EPI_LD = 20;                       // staging row stride (floats): 16 columns + 4 pad, keeps float4 alignment
constexpr uint32_t GEMM_TMEM_COLS = 512;      // two accumulator stages of up to 256 columns

enum { EPI_ROW = 0, EPI_CONV = 1 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct GemmParams {
  // tiling
  int m_tiles, n_tiles;          // tiles along A rows / B rows
  int block_n;                   // B rows per tile (multiple of 16, <= 256)
  int k_blocks_per_tap;          // Cin / 64 (or K / 64 for plain GEMM)
  int taps_r, taps_s;            // 1x1 or 3x3
  int stages;
  // B addressing
  int b_is_conv;                 // 0: 2-D [rows][K]; 1: 4-D NHWC box
  int conv_stride, conv_pad;
  int tw, th, tn;                // output-pixel box of one tile (block_n == tw*th*tn)
  int tiles_y;                   // tiles per image along y (ceil(Hout/th)); x is never split
  int Hout, Wout, Nimg;
  // epilogue
  int epi;                       // EPI_ROW / EPI_CONV
  int act;
  int M, N;                      // EPI_ROW: valid rows of A / B.  EPI_CONV: M = Cout
  int ldo;                       // EPI_ROW: output leading dimension (elements)
  int out_is_f32;                // EPI_ROW
  const float* bias;             // EPI_ROW: [N] or null
  const float* residual;         // EPI_ROW: fp32 [M][ldo] or null (may alias out)
  void* out;
  const float* scale;            // EPI_CONV: [Cout]
  const float* shift;            // EPI_CONV: [Cout]
  int idt_blocks;                // EPI_CONV: 0, or 2 = residual identity added by the tensor core (see kernel)
  float* gap_partial;            // EPI_CONV: [Nimg][tiles_y * GEMM_EPI_GROUPS][Cout] or null
  int gap_raw;                   // pool the raw accumulator (conv1 hook is pre-BN)
  int dbg_skip_epilogue;         // profiling experiments only (B200VQA_GEMM_NOEPI=1): drain nothing, store nothing
  // halo mode (3x3, stride 1, pad 1): the activation tile is fetched ONCE per channel block with its one-pixel halo and the nine
  // taps read it as shifted views (a K-major SWIZZLE_128B operand may start at any 128-byte row: tools/probes/desc_probe.py);
  // block_n = th * halo_pitch rounded up to 16 accumulator columns, column y * halo_pitch + x = output pixel (y, x)
  int halo;                      // 0 / 1
  int halo_pitch;                // pixels per haloed row (Wout + 2)
  int halo_rows_img;             // (th + 2) * halo_pitch: haloed pixels per image of the tile
  int halo_bytes;                // one halo buffer (box + over-read padding, multiple of 1024)
  int dbg_bshift, dbg_bbase;     // descriptor probe (tools/probes/desc_probe.py): B operand read from row dbg_bshift of the tile, base-offset field dbg_bbase
};

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (sticky launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { printf("b200vqa gemm: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: 8-row groups are 1024 B apart (SBO), LBO unused (1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (ignored for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp16 A/B (K-major), fp32 accumulate, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
}

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the
// fp16 rounding of the output): ~16 instructions instead of ~40 for erff(), which made the fc1 epilogue the
// bottleneck of the ViT MLP (see profiles/).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);        // erf(|x| / sqrt 2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// Two GELUs per instruction on the packed fp32x2 pipe (FFMA2 / FMUL2 / FADD2 with broadcast constants): per lane the
// operation sequence of gelu_erf() with the polynomial's sign folded into its coefficients, i.e. the same values.  The
// fc1 epilogue is issue-bound on this function (ncu: 150 M warp instructions in the 264-image fc1 launch, issue slots
// 50 % busy with the tensor pipe 40 % active), so halving its FP32 instruction count shortens the kernel.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = __fmul2_rn(ax, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  float2 t;                                            // = __fdividef(1, d) for d in [1, 2^126): the bare reciprocal
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
  float2 pn = __ffma2_rn(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));    // -poly
  pn = __ffma2_rn(pn, t, make_float2(-1.421413741f, -1.421413741f));
  pn = __ffma2_rn(pn, t, make_float2(0.284496736f, 0.284496736f));
  pn = __ffma2_rn(pn, t, make_float2(-0.254829592f, -0.254829592f));
  const float2 zz = __fmul2_rn(z, z);
  const float2 arg = __fmul2_rn(zz, make_float2(-1.4426950408889634f, -1.4426950408889634f));       // -z^2 log2(e)
  float2 ex;                                                                                         // = __expf(-z^2) of the scalar version
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(arg.y));
  const float2 e = __ffma2_rn(__fmul2_rn(pn, t), ex, make_float2(1.0f, 1.0f));                        // erf(|x| / sqrt 2)
  const float2 s = __fadd2_rn(make_float2(copysignf(e.x, x.x), copysignf(e.y, x.y)), make_float2(1.0f, 1.0f));
  return __fmul2_rn(__fmul2_rn(x, make_float2(0.5f, 0.5f)), s);
}

// Row-mode epilogue of one accumulator tile (shared by the 1-CTA and 2-CTA kernels).  The accumulator arrives one
// row per thread; a warp-private 32 x 16 fp32 staging tile in shared memory transposes it so that every global
// access covers whole 64-byte row segments (8 rows per instruction) instead of 32 different cache lines.
// 16-column chunks alternate between the warps of a quadrant.  Bias / activation / residual are applied after
// the transposition; the residual lines are requested before the TMEM wait.
__device__ __forceinline__ void epi_row_fast(const float* __restrict__ bias, const float* residual, void* out, int M, int ldo,
                                             int out_is_f32, int act, int block_n, uint32_t taddr, float* stg, int m_base, int n0,
                                             int grp, int lane) {
  const int tr = lane >> 2, tc = (lane & 3) * 4;      // transposed ownership: row i*8 + tr, columns tc..tc+3
  for (int c0 = grp * 16; c0 < block_n; c0 += 16 * GEMM_EPI_GROUPS) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    const int n = n0 + c0 + tc;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
    float4 r4[4];
    if (residual) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mm = m_base + i * 8 + tr;
        r4[i] = mm < M ? *reinterpret_cast<const float4*>(residual + (size_t)mm * ldo + n) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * j) =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 8 + tr;
      const int mm = m_base + r;
      float4 x = *reinterpret_cast<const float4*>(stg + r * EPI_LD + tc);
      if (mm < M) {
        x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
        if (act == ACT_GELU) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
        else if (act == ACT_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
        const size_t o = (size_t)mm * ldo + n;
        if (residual) { x.x += r4[i].x; x.y += r4[i].y; x.z += r4[i].z; x.w += r4[i].w; }
        if (out_is_f32) {
          *reinterpret_cast<float4*>(static_cast<float*>(out) + o) = x;
        } else {
          __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
          *reinterpret_cast<uint2*>(static_cast<__half*>(out) + o) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
      }
    }
    __syncwarp();
  }
}

// fp16-output row epilogue without shared memory (qkv / fc1 of the ViT: 7/12 of its FLOPs, K = 768).  At full tensor
// rate the operand traffic of a K-block (TMA fill + MMA read) already saturates the 128 B/clk shared-memory port, so
// the 2 x 128 KB the staged transposition moves per 128 x 256 tile is paid for in tensor throughput (12 K-blocks x
// 64 KB of operands vs 256 KB of staging: -25 %).  Here each thread converts 32 consecutive columns of its own
// accumulator row and writes them as one 64-byte segment (two full sectors) straight from registers.
__device__ __forceinline__ void epi_row_direct_f16(const float* __restrict__ bias, __half* __restrict__ out, int M, int ldo, int act,
                                                   int block_n, uint32_t taddr, int m, int n0, int grp) {
  for (int c0 = grp * 32; c0 < block_n; c0 += 32 * GEMM_EPI_GROUPS) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    float4 b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + c0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    tmem_ld_wait();
    if (m < M) {
      uint32_t hv[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 xa = __fadd2_rn(make_float2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), make_float2(b[i].x, b[i].y));
        float2 xb = __fadd2_rn(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), make_float2(b[i].z, b[i].w));
        if (act == ACT_GELU) { xa = gelu_erf2(xa); xb = gelu_erf2(xb); }
        else if (act == ACT_RELU) { xa.x = fmaxf(xa.x, 0.f); xa.y = fmaxf(xa.y, 0.f); xb.x = fmaxf(xb.x, 0.f); xb.y = fmaxf(xb.y, 0.f); }
        const __half2 h0 = __floats2half2_rn(xa.x, xa.y), h1 = __floats2half2_rn(xb.x, xb.y);
        hv[2 * i] = *reinterpret_cast<const uint32_t*>(&h0); hv[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      }
      // two 256-bit stores (STG.256, sm_100): one full 32-byte sector per lane and instruction
      __half* op = out + (size_t)m * ldo + n0 + c0;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op + 16 * i), "r"(hv[8 * i]), "r"(hv[8 * i + 1]), "r"(hv[8 * i + 2]),
                     "r"(hv[8 * i + 3]), "r"(hv[8 * i + 4]), "r"(hv[8 * i + 5]), "r"(hv[8 * i + 6]), "r"(hv[8 * i + 7]) : "memory");
    }
  }
}

// fp32-output row epilogue without shared memory (proj / fc2 of the ViT: out = acc + bias + residual, fp32 residual stream).
// Same idea as epi_row_direct_f16: a thread owns 32 consecutive columns of its accumulator row = 128 contiguous bytes, read
// (residual) and written as 32-byte vectors, one full sector per lane and instruction; the 2 x 128 KB per 128 x 256 tile
// that the staged transposition moves through the shared-memory port (27 % on top of the operand traffic of a K = 768
// tile) disappear.  MEASURED (round 2, 264 images): ViT pass 13.84 ms with this path vs 13.64 ms with the staged one - the
// linears are bound by L2 -> SM operand traffic (32 KB per K-block and SM against the ~42 B/clk/SM L2 cap), not by the
// epilogue's shared-memory port share, and 128-byte-per-thread rows cost more LSU transactions.  Kept for A/B only.
__device__ __forceinline__ void epi_row_direct_f32(const float* __restrict__ bias, const float* residual, float* out, int M, int ldo,
                                                   int act, int block_n, uint32_t taddr, int m, int n0, int grp) {
  for (int c0 = grp * 32; c0 < block_n; c0 += 32 * GEMM_EPI_GROUPS) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    float r[32];
    const size_t o = (size_t)m * ldo + n0 + c0;
    if (residual && m < M) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r[8 * i]), "=f"(r[8 * i + 1]), "=f"(r[8 * i + 2]), "=f"(r[8 * i + 3]), "=f"(r[8 * i + 4]), "=f"(r[8 * i + 5]),
                       "=f"(r[8 * i + 6]), "=f"(r[8 * i + 7]) : "l"(residual + o + 8 * i) : "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = 0.f;
    }
    tmem_ld_wait();
    if (m < M) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b0 = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + c0) + 2 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b1 = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + c0) + 2 * i + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float bj8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float bj = bj8[j];
          float t = __uint_as_float(v[8 * i + j]) + bj;
          if (act == ACT_GELU) t = gelu_erf(t); else if (act == ACT_RELU) t = fmaxf(t, 0.f);
          x[j] = t + r[8 * i + j];
        }
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + o + 8 * i), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]),
                     "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]) : "memory");
      }
    }
  }
}

// ------------------------------------------------------------------------------ the kernel
#ifdef B200VQA_GEMM_KERNEL_TU      // defined by gemm_host.cu only (one definition per library)
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_eye, const __grid_constant__ CUtensorMap map_idt,
                    const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = GEMM_BM * GEMM_BK * 2;
  const uint32_t b_bytes = (uint32_t)p.block_n * GEMM_BK * 2;
  const uint32_t stage_bytes = p.halo ? a_bytes : a_bytes + b_bytes;       // halo mode: the ring holds weight K-blocks only
  uint8_t* halo_buf = smem + (size_t)p.stages * stage_bytes;               // halo mode: two haloed activation tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(halo_buf + (p.halo ? 2 * (size_t)p.halo_bytes : 0));
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* halo_full = tmem_empty + 2;
  uint64_t* halo_empty = halo_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(halo_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(tmem_slot + 4);        // [GEMM_EPI_WARPS][32][EPI_LD] fp32, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int conv_blocks = p.taps_r * p.taps_s * p.k_blocks_per_tap;
  const int k_blocks = conv_blocks + p.idt_blocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], GEMM_EPI_WARPS); mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(GEMM_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int hb = 0; uint32_t hphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
        int bx = 0, by = 0, bn = 0;
        if (p.b_is_conv) {
          const int ty = nt % p.tiles_y, tg = nt / p.tiles_y;
          by = ty * p.th * p.conv_stride - p.conv_pad;
          bx = -p.conv_pad;
          bn = tg * p.tn;
        }
        if (p.halo) {
          const uint32_t box_bytes = (uint32_t)p.halo_rows_img * (uint32_t)p.tn * 128u;
          for (int cb = 0; cb < p.k_blocks_per_tap; ++cb) {
            mbar_wait(&halo_empty[hb], hphase ^ 1);
            mbar_expect_tx(&halo_full[hb], box_bytes);
            tma_load_4d(&map_b, &halo_full[hb], halo_buf + (size_t)hb * p.halo_bytes, cb * GEMM_BK, bx, by, bn);
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], a_bytes);
              tma_load_2d(&map_a, &full_bar[stage], smem + (size_t)stage * stage_bytes, (tap * p.k_blocks_per_tap + cb) * GEMM_BK, mt * GEMM_BM);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (++hb == 2) { hb = 0; hphase ^= 1; }
          }
          continue;
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if (p.dbg_skip_epilogue & 2) {          // experiment: no loads, just hand the slot over
            mbar_arrive(&full_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          if (kb >= conv_blocks) {
            // residual identity: D[c][pix] += sum_k onehot[c][k] * identity[pix][mt*128 + 64*j + k]
            const int j = kb - conv_blocks;
            tma_load_2d(&map_eye, &full_bar[stage], sa, j * GEMM_BK, 0);
            tma_load_4d(&map_idt, &full_bar[stage], sb, mt * GEMM_BM + j * GEMM_BK, 0, by + p.conv_pad, bn);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          tma_load_2d(&map_a, &full_bar[stage], sa, kb * GEMM_BK, mt * GEMM_BM);
          if (p.b_is_conv) {
            const int tap = kb / p.k_blocks_per_tap, cb = kb % p.k_blocks_per_tap;
            const int r = tap / p.taps_s, s = tap % p.taps_s;
            tma_load_4d(&map_b, &full_bar[stage], sb, cb * GEMM_BK, bx + s, by + r, bn);
          } else {
            tma_load_2d(&map_b, &full_bar[stage], sb, kb * GEMM_BK, nt * p.block_n);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = make_idesc(p.block_n);
    int stage = 0; uint32_t phase = 0;
    int hb = 0; uint32_t hphase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if (lane == 0) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      __syncwarp();
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      if (p.halo) {
        // One MMA per tap and K step, as in the per-tap-box path: for tap (r, s) the th output rows of the tile are th * pitch
        // CONSECUTIVE rows of the haloed tile starting at row r * pitch + s (the two halo columns between image rows become two
        // masked accumulator columns per row).  N = block_n = th * pitch rounded up to 16.  (Per-row MMAs of N = 64 / 32 / 16 were
        // 2-6x slower: a tcgen05.mma costs ~120 clk whatever its N.)
        if (lane == 0) {
          for (int cb = 0; cb < p.k_blocks_per_tap; ++cb) {
            mbar_wait(&halo_full[hb], hphase);
            const uint64_t hdesc = make_smem_desc(smem_u32(halo_buf + (size_t)hb * p.halo_bytes));
            for (int tap = 0; tap < 9; ++tap) {
              const int r = tap / 3, sft = tap - 3 * r;
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
              const uint64_t adesc = make_smem_desc(smem_u32(smem + (size_t)stage * stage_bytes));
              const uint64_t bdesc = hdesc + (uint64_t)((r * p.halo_pitch + sft) * 8);      // 128-byte rows = 8 descriptor units
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k)
                tcgen05_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (cb | tap | k) != 0);
              tcgen05_commit(&empty_bar[stage]);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            tcgen05_commit(&halo_empty[hb]);            // the haloed tile may be overwritten when these MMAs retire
            if (++hb == 2) { hb = 0; hphase ^= 1; }
          }
          tcgen05_commit(&tmem_full[acc]);
        }
        __syncwarp();
        // every lane tracks the ring positions
        if (lane != 0) {
          const int adv = 9 * p.k_blocks_per_tap;
          for (int i = 0; i < adv; ++i) if (++stage == p.stages) { stage = 0; phase ^= 1; }
          for (int i = 0; i < p.k_blocks_per_tap; ++i) if (++hb == 2) { hb = 0; hphase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      for (int kb = 0; kb < k_blocks; ++kb) {
        if (lane == 0) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + a_bytes + 128u * (uint32_t)p.dbg_bshift) | ((uint64_t)(p.dbg_bbase & 7) << 49);
          if (!(p.dbg_skip_epilogue & 4)) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)      // +32 B per K=16 step inside the swizzled row
              tcgen05_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          tcgen05_commit(&empty_bar[stage]);          // frees the smem slot when the MMAs retire
          if (kb == k_blocks - 1) tcgen05_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int q = warp & 3;                           // TMEM lane quadrant of this warp
    const int grp = (warp - 4) >> 2;                  // which share of the columns this warp drains
    const int row = q * 32 + lane;                    // accumulator row owned by this thread
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
      if (p.dbg_skip_epilogue & 1) {
      } else if (p.epi == EPI_ROW) {
        const int m = mt * GEMM_BM + row;
        const int n0 = nt * p.block_n;
        if ((p.block_n & 31) == 0 && n0 + p.block_n <= p.N && !p.out_is_f32 && !p.residual && (p.ldo & 15) == 0 && !(p.dbg_skip_epilogue & 8)) {
          epi_row_direct_f16(p.bias, static_cast<__half*>(p.out), p.M, p.ldo, p.act, p.block_n, taddr, m, n0, grp);
        } else if ((p.block_n & 31) == 0 && n0 + p.block_n <= p.N) {
          epi_row_fast(p.bias, p.residual, p.out, p.M, p.ldo, p.out_is_f32, p.act, p.block_n, taddr, epi_stage + (warp - 4) * (32 * EPI_LD),
                       mt * GEMM_BM + q * 32, n0, grp, lane);
        } else {
          // generic path (ragged N / odd tile widths): 16-column chunks, scalar tail
          for (int c0 = grp * 16; c0 < p.block_n; c0 += 16 * GEMM_EPI_GROUPS) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
            if (m < p.M) {
              const int n = n0 + c0;
              for (int j = 0; j < 16 && n + j < p.N; ++j) {
                float x = __uint_as_float(v[j]) + (p.bias ? p.bias[n + j] : 0.f);
                if (p.act == ACT_GELU) x = gelu_erf(x); else if (p.act == ACT_RELU) x = fmaxf(x, 0.f);
                const size_t o = (size_t)m * p.ldo + n + j;
                if (p.residual) x += p.residual[o];
                if (p.out_is_f32) static_cast<float*>(p.out)[o] = x; else static_cast<__half*>(p.out)[o] = __float2half_rn(x);
              }
            }
          }
        }
      } else {
        // EPI_CONV: this thread owns output channel c; columns are output pixels of the box.
        // The two warps of a quadrant split the pixel columns into contiguous halves (16-px chunks).
        const int c = mt * GEMM_BM + row;
        const int ty = nt % p.tiles_y, tg = nt / p.tiles_y;
        if (mt * GEMM_BM + q * 32 < p.M) {              // warp-uniform: Cout is a multiple of 32
          const float sc = p.scale[c], sh = p.shift[c];
          const float lo = p.act == ACT_RELU ? 0.f : -INFINITY;
          __half* __restrict__ out = static_cast<__half*>(p.out);
          if (p.halo) {
            // halo mode: accumulator columns y * pitch + x, x < Wout valid; the tile's rows are dealt to the warp groups
            // (no pooling hook on these layers, so no reduction order to preserve)
            const int cpr = (p.Wout + 15) >> 4;                                   // 16-column chunks per row
            for (int ci = grp; ci < p.th * cpr; ci += GEMM_EPI_GROUPS) {
              const int y = ci / cpr, x0 = (ci - y * cpr) << 4;
              const int oy = ty * p.th + y;
              if (oy >= p.Hout) break;
              __half* op = out + ((size_t)(tg * p.Hout + oy) * p.Wout + x0) * p.M + c;
              uint32_t v[16];
              tmem_ld16(taddr + y * p.halo_pitch + x0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (x0 + j < p.Wout) op[(size_t)j * p.M] = __float2half_rn(fmaxf(fmaf(__uint_as_float(v[j]), sc, sh), lo));
              }
            }
          } else if (p.tw == p.Wout && p.tn == 1) {
            // full-width boxes: the tile's valid pixels are consecutive NHWC rows -> linear addressing
            const int y0 = ty * p.th;
            const int valid = min(p.block_n, (p.Hout - y0) * p.Wout);
            const int nchunks = (valid + 15) >> 4;
            const int half = (nchunks + GEMM_EPI_GROUPS - 1) / GEMM_EPI_GROUPS;
            const int ch_lo = grp * half, ch_hi = min(nchunks, ch_lo + half);
            __half* obase = out + ((size_t)(tg * p.Hout + y0) * p.Wout) * p.M + c;
            const size_t stride = (size_t)p.M;
            float gsum = 0.f;
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
              const int pix0 = ch << 4;
              uint32_t v[16];
              tmem_ld16(taddr + pix0, v);
              __half* op = obase + (size_t)pix0 * stride;
              tmem_ld_wait();
              if (pix0 + 16 <= valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float raw = __uint_as_float(v[j]);
                  const float val = fmaxf(fmaf(raw, sc, sh), lo);
                  gsum += p.gap_raw ? raw : val;
                  op[(size_t)j * stride] = __float2half_rn(val);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (pix0 + j < valid) {
                    const float raw = __uint_as_float(v[j]);
                    const float val = fmaxf(fmaf(raw, sc, sh), lo);
                    gsum += p.gap_raw ? raw : val;
                    op[(size_t)j * stride] = __float2half_rn(val);
                  }
                }
              }
            }
            if (p.gap_partial)
              p.gap_partial[((size_t)tg * p.tiles_y * GEMM_EPI_GROUPS + ty * GEMM_EPI_GROUPS + grp) * p.M + c] = gsum;
          } else {
            // overhanging boxes (7x7 maps in 8x8x4 boxes): per-pixel coordinates and masks.  Whole images are
            // assigned to the warps of a quadrant (image i -> warp group i % GROUPS) so that the pooling sum of an
            // image never depends on its position in the batch (bit-exact batch invariance).
            const int per_img = p.tw * p.th;                 // multiple of 16 (checked on the host)
            const int chunks_per_img = per_img >> 4;
            float gs[4] = {0.f, 0.f, 0.f, 0.f};             // pooling partials per image of the tile (tn <= 4)
            for (int img = grp; img < p.tn; img += GEMM_EPI_GROUPS) {
              const int n = tg * p.tn + img;
              if (n >= p.Nimg) break;
              float gsum = 0.f;
              for (int ci = 0; ci < chunks_per_img; ++ci) {
                const int pix0 = img * per_img + (ci << 4);
                uint32_t v[16];
                tmem_ld16(taddr + pix0, v);
                const int rem = ci << 4;
                int yrel = rem / p.tw, x = rem - yrel * p.tw;
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int y = ty * p.th + yrel;
                  if (x < p.Wout && y < p.Hout) {
                    const float raw = __uint_as_float(v[j]);
                    const float val = fmaxf(fmaf(raw, sc, sh), lo);
                    gsum += p.gap_raw ? raw : val;
                    out[((size_t)(n * p.Hout + y) * p.Wout + x) * p.M + c] = __float2half_rn(val);
                  }
                  if (++x == p.tw) { x = 0; ++yrel; }
                }
              }
              gs[img] = gsum;
            }
            if (p.gap_partial) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int n = tg * p.tn + i;
                if (i < p.tn && n < p.Nimg)
                  p.gap_partial[((size_t)n * p.tiles_y * GEMM_EPI_GROUPS + ty * GEMM_EPI_GROUPS + grp) * p.M + c] = gs[i];
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GEMM_TMEM_COLS) : "memory");
  }
}


// =====================================================================================================
// 2-CTA variant for the linear layers: a cluster of two CTAs (one SM pair) computes a 256 x 256 tile with
// tcgen05.mma.cta_group::2.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256
// weight rows), so per SM the shared-memory traffic per K-block drops from 96 KB (48 written by TMA + 48 read by
// the MMA) to 64 KB - the 1-CTA kernel is capped at ~2/3 of the tensor rate by the 128 B/clk shared-memory port.
// Protocol: both producers signal the LEADER's full barrier (cta_group::2 TMA); the leader issues the MMAs and
// multicasts its commits to both CTAs' empty / tmem_full barriers; both epilogues arrive on the leader's
// tmem_empty barrier (remote mbarrier arrive).
struct Gemm2Params {
  int m2_tiles, n_tiles, k_blocks, stages;
  int act, M, N, ldo, out_is_f32;
  const float* bias; const float* residual; void* out;
  int dbg_skip_epilogue;
  int raster;                    // 0: row-tile index fastest; G > 0: groups of G row tiles, walked row-tile fastest inside a group
};
constexpr int G2_BN = 256;
// tile -> (row-tile pair m2, column tile nt).  With groups of G row tiles the 74 clusters of a wave share G A panels and
// ~74/G B panels instead of streaming 74 different A panels against one B panel.
__device__ __forceinline__ void g2_tile_coords(const Gemm2Params& p, int tile, int& m2, int& nt) {
  if (p.raster <= 0) { m2 = tile % p.m2_tiles; nt = tile / p.m2_tiles; return; }
  const int per_group = p.raster * p.n_tiles;
  const int g = tile / per_group, r = tile - g * per_group;
  const int m_first = g * p.raster;
  const int gsize = min(p.raster, p.m2_tiles - m_first);
  nt = r / gsize;
  m2 = m_first + (r - nt * gsize);
}
constexpr uint32_t G2_STAGE_BYTES = 2 * GEMM_BM * GEMM_BK * 2;       // A 128x64 + B-half 128x64 (fp16) per CTA

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm_mc(uint64_t* bar) {      // arrive on this barrier in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {   // arrive on `bar` of CTA `cta`
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2cta_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Gemm2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = GEMM_BM * GEMM_BK * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * G2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = p.m2_tiles * p.n_tiles;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * GEMM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(GEMM_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();         // orders the allocator's write of tmem_slot for this CTA's readers (the cluster barrier below does too, but
                           // compute-sanitizer racecheck only models CTA barriers: round 1 reported 98 hazards on this read)
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own A rows + own half of the B rows =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m2, nt; g2_tile_coords(p, tile, m2, nt);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * G2_STAGE_BYTES;
          uint8_t* sb = sa + a_bytes;
          if (p.dbg_skip_epilogue & 2) {          // experiment: no loads
            if (rank == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);      // bytes of both CTAs land on the leader's barrier
          const uint32_t leader_bar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
          tma_load_2d_2sm(&map_a, leader_bar, sa, kb * GEMM_BK, m2 * 256 + (int)rank * GEMM_BM);
          tma_load_2d_2sm(&map_b, leader_bar, sb, kb * GEMM_BK, nt * G2_BN + (int)rank * 128);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(G2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        if (lane == 0) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        __syncwarp();
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (lane == 0) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * G2_STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + a_bytes);
            if (!(p.dbg_skip_epilogue & 4)) {
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k)
                tcgen05_mma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            }
            tcgen05_commit_2sm_mc(&empty_bar[stage]);
            if (kb == p.k_blocks - 1) tcgen05_commit_2sm_mc(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (both CTAs): own 128 rows x 256 columns =================
    const int q = warp & 3, grp = (warp - 4) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m2, nt; g2_tile_coords(p, tile, m2, nt);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
      if (p.dbg_skip_epilogue & 1) {
      } else if (!p.out_is_f32 && !p.residual && (p.ldo & 15) == 0 && !(p.dbg_skip_epilogue & 8)) {
        epi_row_direct_f16(p.bias, static_cast<__half*>(p.out), p.M, p.ldo, p.act, G2_BN, taddr,
                           m2 * 256 + (int)rank * GEMM_BM + q * 32 + lane, nt * G2_BN, grp);
      } else if (p.out_is_f32 && (p.ldo & 7) == 0 && (p.dbg_skip_epilogue & 16)) {      // A/B only (B200VQA_GEMM_NOEPI=16): measured slower, see below
        epi_row_direct_f32(p.bias, p.residual, static_cast<float*>(p.out), p.M, p.ldo, p.act, G2_BN, taddr,
                           m2 * 256 + (int)rank * GEMM_BM + q * 32 + lane, nt * G2_BN, grp);
      } else {
        epi_row_fast(p.bias, p.residual, p.out, p.M, p.ldo, p.out_is_f32, p.act, G2_BN, taddr, epi_stage + (warp - 4) * (32 * EPI_LD),
                     m2 * 256 + (int)rank * GEMM_BM + q * 32, nt * G2_BN, grp, lane);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                      // neither CTA may exit (or free TMEM) while its peer can still touch it
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GEMM_TMEM_COLS) : "memory");
  }
}


// =====================================================================================================
// 4-CTA variant (two SM pairs per cluster, TMA multicast of the weight tile) - A/B ONLY, measured slower (below).
// Hypothesis tested: the linears of the ViT are bound by the L2 -> SM operand stream (32 KB per K-block and SM).  Here the
// two pairs of a cluster compute two vertically adjacent 256 x 256 tiles (row-tile pairs 2 m4 and 2 m4 + 1, same 256
// weight rows): a CTA (pair p, position r) loads its own 128 rows of A and only a QUARTER of the weight tile (64 rows),
// multicast to the CTA with the same position in the other pair - 24 KB instead of 32 KB requested from L2 per K-block.
// Protocol on top of the 2-CTA kernel's: every multicast lands on the full barrier of the destination's pair leader
// (cta_group::2 TMA, peer bit cleared), whose expected byte count is unchanged (2 x 16 KB of A + 4 x 8 KB of B); a
// producer may only refill a stage when BOTH pairs have retired the MMAs that read it (it writes into the other pair's
// shared memory too), so the empty barriers count two arrivals and each leader's tcgen05.commit is multicast to all
// four CTAs.  The pairs therefore advance in lockstep.  Results are bit-identical to the 2-CTA kernel's.
// MEASURED (round 2, tools/gemm_mc_ab.py, 264 images, L2 flushed): a B200 keeps only 33 such clusters resident (132 of
// 148 SMs: tools/probes/cluster_probe.cu) and the kernel runs at exactly that ratio - 8192^3: 1115 vs 1252 TFLOP/s (0.89);
// qkv 774 vs 809, proj 595 vs 606, fc1 801 vs 830, fc2 905 vs 1054.  So the operand stream from L2 is NOT the bound (the
// L2 already serves both pairs' requests for the same lines once: multicast saves nothing below 8 CTAs); what the SM
// pays per K-block is unchanged - 32 KB written into and 32 KB read out of its shared memory in the 512 clk of the MMAs.
constexpr uint32_t G4_BQ_BYTES = 64 * GEMM_BK * 2;                    // one multicast quarter of the B tile: 64 rows x 64 fp16
__device__ __forceinline__ void tma_load_2d_2sm_mc(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm_mask(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm4cta_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b64, const Gemm2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = GEMM_BM * GEMM_BK * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * G2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                 // 0..3
  const uint32_t pair = rank >> 1, pos = rank & 1;         // pair of the cluster, position in the pair (0 = leader)
  const int m4_tiles = (p.m2_tiles + 1) >> 1;
  const int num_tiles = m4_tiles * p.n_tiles;
  const int cluster_id = blockIdx.x >> 2, num_clusters = gridDim.x >> 2;
  Gemm2Params q = p; q.m2_tiles = m4_tiles; q.raster = p.raster > 1 ? p.raster >> 1 : p.raster;   // raster groups in units of two row-tile pairs

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b64) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * GEMM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(GEMM_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (all four CTAs): own A rows + one multicast quarter of the B rows =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint16_t bmask = (uint16_t)((1u << pos) | (1u << (2 + pos)));       // the CTAs at my position in both pairs
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m4, nt; g2_tile_coords(q, tile, m4, nt);
        const int m2 = 2 * m4 + (int)pair;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * G2_STAGE_BYTES;
          uint8_t* sb = sa + a_bytes;
          if (pos == 0) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);      // bytes of both CTAs of the pair land on its leader's barrier
          const uint32_t leader_bar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
          tma_load_2d_2sm(&map_a, leader_bar, sa, kb * GEMM_BK, m2 * 256 + (int)pos * GEMM_BM);
          tma_load_2d_2sm_mc(&map_b64, leader_bar, sb + pair * G4_BQ_BYTES, kb * GEMM_BK, nt * G2_BN + (int)pos * 128 + (int)pair * 64, bmask);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (the leader of each pair) =================
    if (pos == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(G2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        if (lane == 0) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        __syncwarp();
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (lane == 0) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * G2_STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + a_bytes);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              tcgen05_mma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            tcgen05_commit_2sm_mask(&empty_bar[stage], (uint16_t)0xF);      // the stage is refilled by CTAs of both pairs
            if (kb == p.k_blocks - 1) tcgen05_commit_2sm_mask(&tmem_full[acc], pair_mask);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (all four CTAs): own 128 rows x 256 columns =================
    const int qd = warp & 3, grp = (warp - 4) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m4, nt; g2_tile_coords(q, tile, m4, nt);
      const int m2 = 2 * m4 + (int)pair;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(qd * 32) << 16);
      if (!p.out_is_f32 && !p.residual && (p.ldo & 15) == 0) {
        epi_row_direct_f16(p.bias, static_cast<__half*>(p.out), p.M, p.ldo, p.act, G2_BN, taddr,
                           m2 * 256 + (int)pos * GEMM_BM + qd * 32 + lane, nt * G2_BN, grp);
      } else {
        epi_row_fast(p.bias, p.residual, p.out, p.M, p.ldo, p.out_is_f32, p.act, G2_BN, taddr, epi_stage + (warp - 4) * (32 * EPI_LD),
                     m2 * 256 + (int)pos * GEMM_BM + qd * 32, nt * G2_BN, grp, lane);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], rank & ~1u);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                      // no CTA may exit (or free TMEM) while a peer can still write into it or signal it
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GEMM_TMEM_COLS) : "memory");
  }
}

#endif  // B200VQA_GEMM_KERNEL_TU

// barriers + TMEM slot + alignment slack, plus the row epilogue's transposition tiles (the convolution epilogue stores straight
// from registers and needs none: its share goes to the operand ring)
inline size_t gemm_smem_fixed(bool row_epilogue = true) {
  return (2 * GEMM_MAX_STAGES + 8) * 8 + 16 + (row_epilogue ? (size_t)GEMM_EPI_WARPS * 32 * EPI_LD * 4 : 0) + 1024;
}
inline size_t gemm_smem_bytes(int block_n, int stages, bool row_epilogue = true) {
  return (size_t)stages * (GEMM_BM * GEMM_BK * 2 + (size_t)block_n * GEMM_BK * 2) + gemm_smem_fixed(row_epilogue);
}
inline size_t gemm_smem_bytes_halo(int halo_bytes, int stages) {
  return (size_t)stages * (GEMM_BM * GEMM_BK * 2) + 2 * (size_t)halo_bytes + gemm_smem_fixed(false);
}

// ---------------------------------------------------------------- host: TMA descriptors
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// fp16 tensor, innermost dimension contiguous; dims/strides innermost-first; strides in bytes for dims 1..rank-1
int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides);

// map_eye / map_idt are only read when p.idt_blocks != 0 (pass nullptr otherwise)
int launch_gemm(const CUtensorMap& map_a, const CUtensorMap& map_b, const GemmParams& p, int sm_count, cudaStream_t st,
                const CUtensorMap* map_eye = nullptr, const CUtensorMap* map_idt = nullptr);
int pick_stages(int block_n, bool row_epilogue = true);

// 2-CTA (cta_group::2) linear layer: out[M][N] = act(A[M][K] W[N][K]^T + bias) (+ residual); N % 256 == 0, K % 64 == 0.
// map_a / map_b must be built with 128-row boxes.
int launch_gemm_2cta(const CUtensorMap& map_a, const CUtensorMap& map_b, int M, int N, int K, const float* bias, const float* residual,
                     void* out, int out_is_f32, int act, int sm_count, cudaStream_t st);
// 4-CTA clusters with the weight tile multicast between two SM pairs; map_b64 must be built with 64-row boxes.
int launch_gemm_4cta(const CUtensorMap& map_a, const CUtensorMap& map_b64, int M, int N, int K, const float* bias, const float* residual,
                     void* out, int out_is_f32, int act, int sm_count, cudaStream_t st);

// SIMT check kernels (gemm_ref.cuh), launched from other translation units through these wrappers
int launch_ref_gemm_rowmajor(const __half* A, const __half* B, const float* bias, const float* residual, void* out, int M, int N,
                             int K, int ldo, int act, int out_is_f32, cudaStream_t st);
int launch_ref_conv_nhwc(const __half* in, const __half* w, const float* scale, const float* shift, const __half* identity,
                         __half* out, float* gap_sum, int gap_raw, int Nimg, int Hin, int Win, int Cin, int Hout, int Wout,
                         int Cout, int R, int S, int stride, int pad, int act, cudaStream_t st);

}  // namespace b200vqa
