// Host side of the tcgen05 GEMM: TMA descriptor construction (driver entry point resolved at
// run time, so the library has no link-time dependency on libcuda), launch, and the SIMT
// check kernels used to validate the tensor-core path on the device.
#include <mutex>
#include <stdlib.h>
#define B200VQA_GEMM_KERNEL_TU
#include "context.h"
#include "gemm_tcgen05.cuh"
#include "gemm_ref.cuh"

namespace b200vqa {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point", cudaErrorNotSupported); return B200VQA_ECUDA; }
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled rank %d failed with CUresult %d", rank, (int)r);
    set_last_error(msg, cudaErrorInvalidValue);
    return B200VQA_ECUDA;
  }
  return B200VQA_OK;
}

int launch_gemm(const CUtensorMap& map_a, const CUtensorMap& map_b, const GemmParams& p, int sm_count, cudaStream_t st,
                const CUtensorMap* map_eye, const CUtensorMap* map_idt) {
  if (p.idt_blocks && (!map_eye || !map_idt || p.idt_blocks != 2)) return B200VQA_EINVAL;
  if (p.epi == EPI_CONV && p.tn > 1 && ((p.tw * p.th) % 16 != 0 || p.tn > 4)) return B200VQA_EINVAL;
  if (p.block_n % 16 || p.block_n < 16 || p.block_n > 256 || p.stages < 2 || p.stages > GEMM_MAX_STAGES) return B200VQA_EINVAL;
  if (p.halo && (p.taps_r != 3 || p.taps_s != 3 || p.conv_stride != 1 || p.conv_pad != 1 || p.idt_blocks || !p.b_is_conv || p.epi != EPI_CONV || p.gap_partial || p.tn != 1 ||
                 p.block_n < p.th * p.halo_pitch || (p.halo_bytes & 1023))) return B200VQA_EINVAL;
  const bool row_epi = p.epi == EPI_ROW;
  size_t smem = p.halo ? gemm_smem_bytes_halo(p.halo_bytes, p.stages) : gemm_smem_bytes(p.block_n, p.stages, row_epi);
  if (smem > 227 * 1024) return B200VQA_EINVAL;
  GemmParams q = p;
  if (const char* e = getenv("B200VQA_GEMM_STAGES")) { int v = atoi(e); if (!p.halo && v >= 2 && v <= GEMM_MAX_STAGES && gemm_smem_bytes(p.block_n, v, row_epi) <= 227 * 1024) q.stages = v; }
  if (const char* e = getenv("B200VQA_GEMM_NOEPI")) q.dbg_skip_epilogue = atoi(e);
  if (const char* e = getenv("B200VQA_GEMM_BSHIFT")) q.dbg_bshift = atoi(e);
  if (const char* e = getenv("B200VQA_GEMM_BBASE")) q.dbg_bbase = atoi(e);
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < sm_count ? tiles : sm_count;
  b200vqa_ctx* ctx = g_ctx;
  std::pair<cudaEvent_t, cudaEvent_t> ev{};
  const bool prof = ctx && ctx->profiling;
  if (prof) {
    if (!ctx->prof_pool.empty()) { ev = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    else { VQA_CUDA(cudaEventCreate(&ev.first)); VQA_CUDA(cudaEventCreate(&ev.second)); }
    VQA_CUDA(cudaEventRecord(ev.first, st));
  }
  smem = q.halo ? gemm_smem_bytes_halo(q.halo_bytes, q.stages) : gemm_smem_bytes(q.block_n, q.stages, row_epi);
  gemm_tcgen05_kernel<<<grid, GEMM_THREADS, smem, st>>>(map_a, map_b, map_eye ? *map_eye : map_a, map_idt ? *map_idt : map_b, q);
  if (prof) {
    VQA_CUDA(cudaEventRecord(ev.second, st));
    ctx->prof_events.push_back(ev);
    // algorithmic FLOPs of the un-padded problem
    const double kdepth = (double)p.taps_r * p.taps_s * p.k_blocks_per_tap * GEMM_BK;   // identity K-blocks are not counted
    if (p.epi == EPI_ROW) ctx->prof_flops += 2.0 * p.M * p.N * kdepth;
    else ctx->prof_flops += 2.0 * p.M * ((double)p.Nimg * p.Hout * p.Wout) * (p.b_is_conv ? kdepth : 147.0);
  }
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

int launch_gemm_2cta(const CUtensorMap& map_a, const CUtensorMap& map_b, int M, int N, int K, const float* bias, const float* residual,
                     void* out, int out_is_f32, int act, int sm_count, cudaStream_t st) {
  if (N % G2_BN || K % GEMM_BK || M <= 0) return B200VQA_EINVAL;
  Gemm2Params p{};
  p.m2_tiles = cdiv(M, 256); p.n_tiles = N / G2_BN; p.k_blocks = K / GEMM_BK;
  const size_t fixed = (2 * GEMM_MAX_STAGES + 4) * 8 + 16 + (size_t)GEMM_EPI_WARPS * 32 * EPI_LD * 4 + 1024;
  p.stages = (int)((227 * 1024 - fixed) / G2_STAGE_BYTES);
  if (p.stages > GEMM_MAX_STAGES) p.stages = GEMM_MAX_STAGES;
  if (const char* e = getenv("B200VQA_GEMM_STAGES")) { int v = atoi(e); if (v >= 2 && v <= p.stages) p.stages = v; }
  if (const char* e = getenv("B200VQA_GEMM_NOEPI")) p.dbg_skip_epilogue = atoi(e);
  p.raster = 8;            // groups of 8 row-tile pairs: +4 % on the ViT linears over row-tile-fastest (profiles/r1_gemm_experiments.md)
  if (const char* e = getenv("B200VQA_GEMM_RASTER")) p.raster = atoi(e);
  p.act = act; p.M = M; p.N = N; p.ldo = N; p.out_is_f32 = out_is_f32; p.bias = bias; p.residual = residual; p.out = out;
  const size_t smem = (size_t)p.stages * G2_STAGE_BYTES + fixed;
  const int tiles = p.m2_tiles * p.n_tiles;
  int clusters = sm_count / 2;
  if (clusters > tiles) clusters = tiles;
  b200vqa_ctx* ctx = g_ctx;
  std::pair<cudaEvent_t, cudaEvent_t> ev{};
  const bool prof = ctx && ctx->profiling;
  if (prof) {
    if (!ctx->prof_pool.empty()) { ev = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    else { VQA_CUDA(cudaEventCreate(&ev.first)); VQA_CUDA(cudaEventCreate(&ev.second)); }
    VQA_CUDA(cudaEventRecord(ev.first, st));
  }
  gemm2cta_tcgen05_kernel<<<2 * clusters, GEMM_THREADS, smem, st>>>(map_a, map_b, p);
  if (prof) {
    VQA_CUDA(cudaEventRecord(ev.second, st));
    ctx->prof_events.push_back(ev);
    ctx->prof_flops += 2.0 * M * (double)N * K;
  }
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

int launch_gemm_4cta(const CUtensorMap& map_a, const CUtensorMap& map_b64, int M, int N, int K, const float* bias, const float* residual,
                     void* out, int out_is_f32, int act, int sm_count, cudaStream_t st) {
  if (N % G2_BN || K % GEMM_BK || M <= 0) return B200VQA_EINVAL;
  Gemm2Params p{};
  p.m2_tiles = cdiv(M, 256); p.n_tiles = N / G2_BN; p.k_blocks = K / GEMM_BK;
  const size_t fixed = (2 * GEMM_MAX_STAGES + 4) * 8 + 16 + (size_t)GEMM_EPI_WARPS * 32 * EPI_LD * 4 + 1024;
  p.stages = (int)((227 * 1024 - fixed) / G2_STAGE_BYTES);
  if (p.stages > GEMM_MAX_STAGES) p.stages = GEMM_MAX_STAGES;
  if (const char* e = getenv("B200VQA_GEMM_STAGES")) { int v = atoi(e); if (v >= 2 && v <= p.stages) p.stages = v; }
  p.raster = 8;
  if (const char* e = getenv("B200VQA_GEMM_RASTER")) p.raster = atoi(e);
  p.act = act; p.M = M; p.N = N; p.ldo = N; p.out_is_f32 = out_is_f32; p.bias = bias; p.residual = residual; p.out = out;
  const size_t smem = (size_t)p.stages * G2_STAGE_BYTES + fixed;
  const int tiles = ((p.m2_tiles + 1) / 2) * p.n_tiles;
  // resident 4-CTA clusters of this device (33 on a 148-SM B200), optionally capped by the context's SM budget
  static int max_clusters[64] = {};
  int dev = 0; VQA_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && max_clusters[dev] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    VQA_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm4cta_tcgen05_kernel, &cfg));
    max_clusters[dev] = n > 0 ? n : 1;
  }
  int clusters = dev < 64 ? max_clusters[dev] : sm_count / 4;
  if (clusters > sm_count / 4) clusters = sm_count / 4;
  if (clusters > tiles) clusters = tiles;
  if (clusters < 1) clusters = 1;
  b200vqa_ctx* ctx = g_ctx;
  std::pair<cudaEvent_t, cudaEvent_t> ev{};
  const bool prof = ctx && ctx->profiling;
  if (prof) {
    if (!ctx->prof_pool.empty()) { ev = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    else { VQA_CUDA(cudaEventCreate(&ev.first)); VQA_CUDA(cudaEventCreate(&ev.second)); }
    VQA_CUDA(cudaEventRecord(ev.first, st));
  }
  gemm4cta_tcgen05_kernel<<<4 * clusters, GEMM_THREADS, smem, st>>>(map_a, map_b64, p);
  if (prof) {
    VQA_CUDA(cudaEventRecord(ev.second, st));
    ctx->prof_events.push_back(ev);
    ctx->prof_flops += 2.0 * M * (double)N * K;
  }
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

int gemm_init_device_attrs() {
  VQA_CUDA(cudaFuncSetAttribute(gemm4cta_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  VQA_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  VQA_CUDA(cudaFuncSetAttribute(gemm2cta_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return B200VQA_OK;
}

int pick_stages(int block_n, bool row_epilogue) {
  const size_t per = GEMM_BM * GEMM_BK * 2 + (size_t)block_n * GEMM_BK * 2;
  int s = (int)((227 * 1024 - gemm_smem_fixed(row_epilogue)) / per);
  if (s > 6) s = 6;
  return s;
}

int launch_ref_gemm_rowmajor(const __half* A, const __half* B, const float* bias, const float* residual, void* out, int M, int N,
                             int K, int ldo, int act, int out_is_f32, cudaStream_t st) {
  dim3 grid(cdiv(N, 16), cdiv(M, 16)), block(16, 16);
  ref_gemm_rowmajor<<<grid, block, 0, st>>>(A, B, bias, residual, out, M, N, K, ldo, act, out_is_f32);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

int launch_ref_conv_nhwc(const __half* in, const __half* w, const float* scale, const float* shift, const __half* identity,
                         __half* out, float* gap_sum, int gap_raw, int Nimg, int Hin, int Win, int Cin, int Hout, int Wout,
                         int Cout, int R, int S, int stride, int pad, int act, cudaStream_t st) {
  const size_t total = (size_t)Nimg * Hout * Wout * Cout;
  if (gap_sum) VQA_CUDA(cudaMemsetAsync(gap_sum, 0, (size_t)Nimg * Cout * sizeof(float), st));
  ref_conv_nhwc<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, w, scale, shift, identity, out, gap_sum, gap_raw, Nimg, Hin, Win,
                                                                  Cin, Hout, Wout, Cout, R, S, stride, pad, act);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

}  // namespace b200vqa

using namespace b200vqa;

// Plain GEMM entry (tests / profiling): D[M][N] = A[M][K] B[N][K]^T + bias
extern "C" int b200vqa_gemm_f16(b200vqa_t* h, const void* A, const void* B, const float* bias, float* D, int M, int N,
                                int K, int impl, void* stream) {
  if (!h || !A || !B || !D || M <= 0 || N <= 0 || K <= 0 || K % 8) return B200VQA_EINVAL;
  CtxScope scope(h);
  cudaStream_t st = as_stream(stream);
  if (impl == 1) {
    return launch_ref_gemm_rowmajor(static_cast<const __half*>(A), static_cast<const __half*>(B), bias, nullptr, D, M, N, K, N,
                                    ACT_NONE, 1, st);
  }
  if (impl == 2) {
    CUtensorMap ma2, mb2;
    uint64_t da2[2] = {(uint64_t)K, (uint64_t)M}, db2[2] = {(uint64_t)K, (uint64_t)N}, sa2[1] = {(uint64_t)K * 2};
    uint32_t box2[2] = {GEMM_BK, GEMM_BM};
    int rc2;
    if ((rc2 = make_tmap_f16(&ma2, A, 2, da2, sa2, box2, nullptr))) return rc2;
    if ((rc2 = make_tmap_f16(&mb2, B, 2, db2, sa2, box2, nullptr))) return rc2;
    return launch_gemm_2cta(ma2, mb2, M, N, K, bias, nullptr, D, 1, ACT_NONE, gemm_grid_sms(h), st);
  }
  if (impl == 3) {
    CUtensorMap ma4, mb4;
    uint64_t da4[2] = {(uint64_t)K, (uint64_t)M}, db4[2] = {(uint64_t)K, (uint64_t)N}, sa4[1] = {(uint64_t)K * 2};
    uint32_t boxa4[2] = {GEMM_BK, GEMM_BM}, boxb4[2] = {GEMM_BK, 64};
    int rc4;
    if ((rc4 = make_tmap_f16(&ma4, A, 2, da4, sa4, boxa4, nullptr))) return rc4;
    if ((rc4 = make_tmap_f16(&mb4, B, 2, db4, sa4, boxb4, nullptr))) return rc4;
    return launch_gemm_4cta(ma4, mb4, M, N, K, bias, nullptr, D, 1, ACT_NONE, gemm_grid_sms(h), st);
  }
  int bn = N >= 256 ? 256 : ((N + 15) / 16) * 16;
  if (const char* e = getenv("B200VQA_GEMM_BN")) { int v = atoi(e); if (v >= 16 && v <= 256 && v % 16 == 0) bn = v; }
  CUtensorMap ma, mb;
  uint64_t da[2] = {(uint64_t)K, (uint64_t)M}, db[2] = {(uint64_t)K, (uint64_t)N}, sa[1] = {(uint64_t)K * 2};
  uint32_t boxa[2] = {GEMM_BK, GEMM_BM}, boxb[2] = {GEMM_BK, (uint32_t)bn};
  int rc;
  if ((rc = make_tmap_f16(&ma, A, 2, da, sa, boxa, nullptr))) return rc;
  if ((rc = make_tmap_f16(&mb, B, 2, db, sa, boxb, nullptr))) return rc;
  GemmParams p{};
  p.m_tiles = cdiv(M, GEMM_BM); p.n_tiles = cdiv(N, bn); p.block_n = bn;
  p.k_blocks_per_tap = cdiv(K, GEMM_BK); p.taps_r = p.taps_s = 1; p.stages = pick_stages(bn);
  p.epi = EPI_ROW; p.act = ACT_NONE; p.M = M; p.N = N; p.ldo = N; p.out_is_f32 = 1; p.bias = bias; p.out = D;
  return launch_gemm(ma, mb, p, gemm_grid_sms(h), st);
}
