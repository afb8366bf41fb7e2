// A16-A18: temporal mean + concat into the 35,203-dim video vector, then imputer + scaler + Mlp head.
//   reference: src/demo_test.py:171-219, src/model_regression.py:37-58,
//              src/data_processing/extract_npy2mat.py:121-126 (temporal mean).
// All fp32 (fp64 for the imputer/scaler step, like sklearn); negligible cost, fused to two kernels.
#include <vector>
#include "context.h"

namespace b200vqa {

constexpr int HF = B200VQA_FEATURES, HH1 = 256, HH2 = 128;

struct HeadWeights {
  float *fc1_w, *fc1_b, *bn_scale, *bn_shift, *fc2_w, *fc2_b, *fc3_w, *fc3_b;
  double *imp_mean, *sc_scale, *sc_min;
  std::vector<void*> allocs;
};

void free_head(HeadWeights* hw) {
  if (!hw) return;
  for (void* p : hw->allocs) cudaFree(p);
  delete hw;
}

// features[v] = [mean_t full_stack | mean_t full_vit | mean_p frag_stack | mean_p frag_pool |
//                mean_p frag_vit_ori | mean_p frag_vit_mer]; sequential fp32 accumulation in frame
// order then one division, like np.mean(axis=0) on a (T, D) float32 array.
__global__ void __launch_bounds__(256)
k12_temporal_mean_concat(const float* __restrict__ full_stack, const float* __restrict__ full_vit,
                         const float* __restrict__ frag_stack, const float* __restrict__ frag_pool,
                         const float* __restrict__ frag_vit_ori, const float* __restrict__ frag_vit_mer,
                         const int32_t* __restrict__ full_off, const int32_t* __restrict__ pair_off, float* __restrict__ features) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (col >= HF) return;
  const float* src; int ld, c, r0, r1;
  if (col < 13120) { src = full_stack; ld = 13120; c = col; r0 = full_off[v]; r1 = full_off[v + 1]; }
  else if (col < 15424) { src = full_vit; ld = 2304; c = col - 13120; r0 = full_off[v]; r1 = full_off[v + 1]; }
  else if (col < 28544) { src = frag_stack; ld = 13120; c = col - 15424; r0 = pair_off[v]; r1 = pair_off[v + 1]; }
  else if (col < 30595) { src = frag_pool; ld = 2051; c = col - 28544; r0 = pair_off[v]; r1 = pair_off[v + 1]; }
  else if (col < 32899) { src = frag_vit_ori; ld = 2304; c = col - 30595; r0 = pair_off[v]; r1 = pair_off[v + 1]; }
  else { src = frag_vit_mer; ld = 2304; c = col - 32899; r0 = pair_off[v]; r1 = pair_off[v + 1]; }
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += src[(size_t)r * ld + c];
  features[(size_t)v * HF + col] = s / (float)(r1 - r0);      // 0 rows -> NaN, which the imputer replaces
}

// fc1 over up to 8 videos per block: one block per hidden unit
__global__ void __launch_bounds__(256)
k12_head_fc1(const float* __restrict__ feat, int V, const double* __restrict__ imp, const double* __restrict__ scl,
             const double* __restrict__ mn, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ h1) {
  __shared__ float red[8][256];
  const int j = blockIdx.x, v0 = blockIdx.y * 8, t = threadIdx.x;
  const int nv = min(8, V - v0);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float* wr = w + (size_t)j * HF;
  for (int k = t; k < HF; k += 256) {
    const float wk = wr[k];
    const double im = imp[k], sc = scl[k], m0 = mn[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nv) {
        const float xf = feat[(size_t)(v0 + i) * HF + k];
        const double xd = isnan(xf) ? im : (double)xf;
        acc[i] = fmaf((float)(xd * sc + m0), wk, acc[i]);
      }
    }
  }
  for (int i = 0; i < 8; ++i) red[i][t] = acc[i];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) for (int i = 0; i < 8; ++i) red[i][t] += red[i][t + o];
    __syncthreads();
  }
  if (t < nv) h1[(size_t)(v0 + t) * HH1 + j] = red[t][0] + bias[j];
}

__device__ __forceinline__ float gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// BN(eval) + GELU + fc2 + GELU + fc3: one block (128 threads) per video
__global__ void __launch_bounds__(128)
k12_head_tail(const float* __restrict__ h1, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
              const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ b3,
              float* __restrict__ score) {
  __shared__ float a[HH1];
  __shared__ float red[HH2];
  const int v = blockIdx.x, t = threadIdx.x;
  for (int i = t; i < HH1; i += 128) a[i] = gelu(h1[(size_t)v * HH1 + i] * bn_scale[i] + bn_shift[i]);
  __syncthreads();
  float acc = b2[t];
  for (int k = 0; k < HH1; ++k) acc = fmaf(w2[(size_t)t * HH1 + k], a[k], acc);
  red[t] = gelu(acc) * w3[t];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
  if (t == 0) score[v] = red[0] + b3[0];
}

template <typename T>
static int up(HeadWeights* hw, const T* host, size_t n, T** dev) {
  VQA_CUDA(cudaMalloc((void**)dev, n * sizeof(T)));
  hw->allocs.push_back(*dev);
  VQA_CUDA(cudaMemcpy(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return B200VQA_OK;
}

}  // namespace b200vqa

using namespace b200vqa;

extern "C" int b200vqa_load_head(b200vqa_t* h, int in_features, const float* fc1_w, const float* fc1_b, const float* bn_w,
                                 const float* bn_b, const float* bn_mean, const float* bn_var, const float* fc2_w,
                                 const float* fc2_b, const float* fc3_w, const float* fc3_b, const double* imp_mean,
                                 const double* sc_scale, const double* sc_min) {
  if (!h || in_features != HF || !fc1_w || !fc1_b || !bn_w || !bn_b || !bn_mean || !bn_var || !fc2_w || !fc2_b || !fc3_w ||
      !fc3_b || !imp_mean || !sc_scale || !sc_min)
    return B200VQA_EINVAL;
  VQA_CUDA(cudaSetDevice(h->device));
  HeadWeights* hw = new HeadWeights();
  std::vector<float> s(HH1), sh(HH1);
  for (int i = 0; i < HH1; ++i) {         // BatchNorm1d eval: (x - mean) / sqrt(var + eps) * w + b
    s[i] = bn_w[i] / sqrtf(bn_var[i] + 1e-5f);
    sh[i] = bn_b[i] - bn_mean[i] * s[i];
  }
  int rc = up(hw, fc1_w, (size_t)HH1 * HF, &hw->fc1_w);
  if (!rc) rc = up(hw, fc1_b, HH1, &hw->fc1_b);
  if (!rc) rc = up(hw, s.data(), HH1, &hw->bn_scale);
  if (!rc) rc = up(hw, sh.data(), HH1, &hw->bn_shift);
  if (!rc) rc = up(hw, fc2_w, (size_t)HH2 * HH1, &hw->fc2_w);
  if (!rc) rc = up(hw, fc2_b, HH2, &hw->fc2_b);
  if (!rc) rc = up(hw, fc3_w, HH2, &hw->fc3_w);
  if (!rc) rc = up(hw, fc3_b, 1, &hw->fc3_b);
  if (!rc) rc = up(hw, imp_mean, HF, &hw->imp_mean);
  if (!rc) rc = up(hw, sc_scale, HF, &hw->sc_scale);
  if (!rc) rc = up(hw, sc_min, HF, &hw->sc_min);
  if (rc) { free_head(hw); return rc; }
  free_head(h->head);
  h->head = hw;
  return B200VQA_OK;
}

extern "C" int b200vqa_temporal_mean_concat(const float* full_stack, const float* full_vit, const float* frag_stack,
                                            const float* frag_pool, const float* frag_vit_ori, const float* frag_vit_mer,
                                            const int32_t* full_off, const int32_t* pair_off, int V, float* features,
                                            void* stream) {
  if (!full_stack || !full_vit || !frag_stack || !frag_pool || !frag_vit_ori || !frag_vit_mer || !full_off || !pair_off ||
      !features || V <= 0)
    return B200VQA_EINVAL;
  k12_temporal_mean_concat<<<dim3(cdiv(HF, 256), V), 256, 0, as_stream(stream)>>>(full_stack, full_vit, frag_stack, frag_pool,
                                                                                 frag_vit_ori, frag_vit_mer, full_off, pair_off, features);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

extern "C" int b200vqa_head_forward(b200vqa_t* h, const float* features, int V, float* score, void* stream) {
  if (!h || !features || !score || V <= 0) return B200VQA_EINVAL;
  if (!h->head) return B200VQA_ENOTLOADED;
  CtxScope scope(h);
  cudaStream_t st = as_stream(stream);
  int rc = h->ws_head.reserve((size_t)V * HH1 * sizeof(float));
  if (rc) return rc;
  float* h1 = static_cast<float*>(h->ws_head.ptr);
  const HeadWeights& w = *h->head;
  k12_head_fc1<<<dim3(HH1, cdiv(V, 8)), 256, 0, st>>>(features, V, w.imp_mean, w.sc_scale, w.sc_min, w.fc1_w, w.fc1_b, h1);
  VQA_LAUNCH_CHECK();
  k12_head_tail<<<V, 128, 0, st>>>(h1, w.bn_scale, w.bn_shift, w.fc2_w, w.fc2_b, w.fc3_w, w.fc3_b, score);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}
