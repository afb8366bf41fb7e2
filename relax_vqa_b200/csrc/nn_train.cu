// training of the Mlp quality head on the device (fp32, like the reference's PyTorch loop).
//   reference: src/model_regression.py:37-58 (Mlp: fc1 -> BatchNorm1d -> GELU -> Dropout -> fc2 -> GELU -> Dropout -> fc3),
//              :61-89 (MAEAndRankLoss), :292-306 (train_one_epoch: zero_grad / forward / loss / backward / SGD step),
//              :375-390 (SGD momentum 0.9 + weight decay, AveragedModel), :399-405 (swa update), :454-459 (update_bn);
//              src/fine_tune.py:130-193 runs the same step on a pre-trained head.
// One optimisation step = b200vqa_trainer_step: forward in train mode (batch statistics, running-stat update with
// momentum 0.1 and the unbiased variance), the loss and its gradient, the backward pass and the SGD update, all on the
// caller's stream; the host (relax_vqa_b200/model_regression.py) keeps only the schedule (cosine LR, SWA start, k-fold,
// early stopping).  The three contractions are hand-written shared-memory-tiled fp32 SIMT kernels with a deterministic
// split-K (a head step is ~9 GFLOP and 0.3 GB of weight / gradient / momentum traffic: fp32 SIMT keeps bit-level
// reproducibility and parity with the reference's fp32 arithmetic; the tensor cores are not needed here).
#include <vector>
#include "context.h"

namespace b200vqa {

constexpr int TR_H2_DIV = 2;      // hidden2 = hidden / 2 (Mlp)

struct Trainer {
  int device = 0, in = 0, h1 = 0, h2 = 0, cap_b = 0;
  int64_t n_params = 0;
  // parameters / gradients / momentum / SWA average: one flat buffer each, same layout:
  // fc1_w [h1][in] | fc1_b [h1] | bn_w [h1] | bn_b [h1] | fc2_w [h2][h1] | fc2_b [h2] | fc3_w [h2] | fc3_b [1]
  float *p = nullptr, *g = nullptr, *mom = nullptr, *swa = nullptr;
  float *bn_mean = nullptr, *bn_var = nullptr, *swa_bn_mean = nullptr, *swa_bn_var = nullptr;   // running statistics (buffers)
  int64_t steps = 0, n_averaged = 0;
  int64_t o_fc1w, o_fc1b, o_bnw, o_bnb, o_fc2w, o_fc2b, o_fc3w, o_fc3b;
  // activations (grow with the batch)
  float *z1 = nullptr, *xhat = nullptr, *a1 = nullptr, *z2 = nullptr, *a2 = nullptr, *pred = nullptr;
  float *d2 = nullptr, *d1 = nullptr, *dpred = nullptr, *bmean = nullptr, *brstd = nullptr, *partial = nullptr, *loss_rows = nullptr;
  size_t partial_floats = 0;
  std::vector<void*> allocs;
};

// ---------------------------------------------------------------------------------------- tiled fp32 contraction
// C[i][j] (+)= sum_k A(i, k) * B(k, j), i < M, j < N, k in this z-slice of [0, K); arbitrary element strides, so the same
// kernel serves X W^T (forward), dY^T X (weight gradients) and dY W (input gradients).  64 x 64 tile, 16-deep slabs, 256
// threads x 4 x 4 outputs.  gridDim.z > 1: split-K partials [z][M][N] (reduced in a fixed order by tr_reduce_splits).
constexpr int TG_T = 64, TG_K = 16;
__global__ void __launch_bounds__(256)
tr_gemm(const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const float* __restrict__ B, int64_t b_rs, int64_t b_cs,
        float* __restrict__ C, int M, int N, int K, int k_per_split) {
  __shared__ float sA[TG_K][TG_T + 4], sB[TG_K][TG_T + 4];
  const int i0 = blockIdx.y * TG_T, j0 = blockIdx.x * TG_T;
  const int k_lo = blockIdx.z * k_per_split, k_hi = min(K, k_lo + k_per_split);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // loaders: the fastest-varying thread index follows the unit-stride dimension of each operand
  const bool a_kfast = a_cs == 1, b_kfast = b_rs == 1;
  for (int k0 = k_lo; k0 < k_hi; k0 += TG_K) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = t + 256 * r;                       // 1024 elements per operand slab
      const int kk = a_kfast ? (e & 15) : (e >> 6), ii = a_kfast ? (e >> 4) : (e & 63);
      const int gi = i0 + ii, gk = k0 + kk;
      sA[kk][ii] = (gi < M && gk < k_hi) ? A[gi * a_rs + gk * a_cs] : 0.f;
      const int kb = b_kfast ? (e & 15) : (e >> 6), jj = b_kfast ? (e >> 4) : (e & 63);
      const int gj = j0 + jj, gkb = k0 + kb;
      sB[kb][jj] = (gj < N && gkb < k_hi) ? B[gkb * b_rs + gj * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TG_K; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* Cz = C + (size_t)blockIdx.z * M * N;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = i0 + ty * 4 + a;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + tx * 4 + b;
      if (gj < N) Cz[(size_t)gi * N + gj] = acc[a][b];
    }
  }
}

// out[i][j] = sum_z partial[z][i][j] (+ bias[j]) in ascending z: deterministic split-K
__global__ void __launch_bounds__(256)
tr_reduce_splits(const float* __restrict__ partial, int splits, size_t mn, int N, const float* __restrict__ bias, float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= mn) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(size_t)z * mn + idx];
  out[idx] = s + (bias ? bias[idx % N] : 0.f);
}

// ---------------------------------------------------------------------------------------- element-wise pieces
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// BatchNorm1d, train mode: one block per feature.  xhat and the batch statistics are kept for the backward pass; running
// statistics: mean <- (1 - m) mean + m batch_mean, var <- (1 - m) var + m * unbiased batch var (m = 0.1; update_bn passes
// m = 1 / (batches seen) for the cumulative average of torch.optim.swa_utils.update_bn).  Then GELU and dropout
// (mask = 1 keeps; kept values are scaled by 1 / (1 - p)).
__global__ void __launch_bounds__(256)
tr_bn_gelu_drop_fwd(const float* __restrict__ z1, int B, int H, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ run_mean, float* __restrict__ run_var, float momentum, int update_running, int apply,
                    float* __restrict__ xhat, float* __restrict__ bmean, float* __restrict__ brstd, const uint8_t* __restrict__ mask,
                    float keep_scale, float* __restrict__ a1) {
  __shared__ double red[256];
  const int j = blockIdx.x, t = threadIdx.x;
  double s = 0.0;
  for (int i = t; i < B; i += 256) s += z1[(size_t)i * H + j];
  red[t] = s; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
  const float mean = (float)(red[0] / B);
  __syncthreads();
  double q = 0.0;
  for (int i = t; i < B; i += 256) { const double d = (double)z1[(size_t)i * H + j] - mean; q += d * d; }
  red[t] = q; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
  const float var = (float)(red[0] / B);
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  if (t == 0) {
    bmean[j] = mean; brstd[j] = rstd;
    if (update_running) {
      const float unbiased = B > 1 ? (float)(red[0] / (B - 1)) : var;
      run_mean[j] = (1.f - momentum) * run_mean[j] + momentum * mean;
      run_var[j] = (1.f - momentum) * run_var[j] + momentum * unbiased;
    }
  }
  if (!apply) return;
  const float g = gamma[j], b = beta[j];
  for (int i = t; i < B; i += 256) {
    const size_t o = (size_t)i * H + j;
    const float xh = (z1[o] - mean) * rstd;
    xhat[o] = xh;
    float a = gelu_f(fmaf(g, xh, b));
    if (mask) a = mask[o] ? a * keep_scale : 0.f;
    a1[o] = a;
  }
}

// eval mode: running statistics (Mlp.eval())
__global__ void __launch_bounds__(256)
tr_bn_gelu_eval(const float* __restrict__ z1, size_t n, int H, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ run_mean, const float* __restrict__ run_var, float* __restrict__ a1) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  const int j = idx % H;
  a1[idx] = gelu_f(fmaf(gamma[j], (z1[idx] - run_mean[j]) / sqrtf(run_var[j] + 1e-5f), beta[j]));
}

__global__ void __launch_bounds__(256)
tr_gelu_drop_fwd(const float* __restrict__ z, size_t n, const uint8_t* __restrict__ mask, float keep_scale, float* __restrict__ a) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  float v = gelu_f(z[idx]);
  if (mask) v = mask[idx] ? v * keep_scale : 0.f;
  a[idx] = v;
}

// d(z) = d(a) * dropout * gelu'(z)   (in place on d)
__global__ void __launch_bounds__(256)
tr_gelu_drop_bwd(float* __restrict__ d, const float* __restrict__ z, size_t n, const uint8_t* __restrict__ mask, float keep_scale) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  float v = d[idx];
  if (mask) v = mask[idx] ? v * keep_scale : 0.f;
  d[idx] = v * gelu_grad(z[idx]);
}

// backward of dropout -> GELU -> BatchNorm (train mode) for one feature per block:
//   dy = da * mask * gelu'(gamma xhat + beta); dgamma = sum dy xhat; dbeta = sum dy;
//   dz = gamma rstd / B * (B dy - dbeta - xhat dgamma)          (in place on d1)
__global__ void __launch_bounds__(256)
tr_bn_gelu_drop_bwd(float* __restrict__ d1, const float* __restrict__ xhat, int B, int H, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ brstd, const uint8_t* __restrict__ mask, float keep_scale,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double r1[256], r2[256];
  const int j = blockIdx.x, t = threadIdx.x;
  const float g = gamma[j], b = beta[j];
  double sg = 0.0, sb = 0.0;
  for (int i = t; i < B; i += 256) {
    const size_t o = (size_t)i * H + j;
    float v = d1[o];
    if (mask) v = mask[o] ? v * keep_scale : 0.f;
    const float xh = xhat[o];
    const float dy = v * gelu_grad(fmaf(g, xh, b));
    d1[o] = dy;
    sg += (double)dy * xh; sb += dy;
  }
  r1[t] = sg; r2[t] = sb; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) { r1[t] += r1[t + o]; r2[t] += r2[t + o]; } __syncthreads(); }
  const float dg = (float)r1[0], db = (float)r2[0];
  if (t == 0) { dgamma[j] = dg; dbeta[j] = db; }
  const float k = g * brstd[j] / (float)B;
  for (int i = t; i < B; i += 256) {
    const size_t o = (size_t)i * H + j;
    d1[o] = k * ((float)B * d1[o] - db - xhat[o] * dg);
  }
}

// column sums of d [B][N] -> bias gradients
__global__ void __launch_bounds__(256)
tr_colsum(const float* __restrict__ d, int B, int N, float* __restrict__ out) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= N) return;
  double s = 0.0;
  for (int i = 0; i < B; ++i) s += d[(size_t)i * N + j];
  out[j] = (float)s;
}

// MAEAndRankLoss (:61-89, use_margin = False) and its gradient with respect to the predictions; one block per row k:
//   L = l1_w mean|p - y| + rank_w sum_ij relu(t_ij - m_ij (p_i - p_j)) / (n (n - 1)),  t_ij = y_i - y_j, m_ij = sign(t_ij)
//   dL/dp_k = l1_w sign(p_k - y_k) / n + rank_w [sum_i 1_ik m_ik - sum_j 1_kj m_kj] / (n (n - 1))     (relu'(0) = 0, sign(0) = 0)
__global__ void __launch_bounds__(256)
tr_loss_grad(const float* __restrict__ pred, const float* __restrict__ y, int n, float l1_w, float rank_w, float* __restrict__ dpred,
             float* __restrict__ loss_rows) {
  __shared__ double r1[256], r2[256];
  const int k = blockIdx.x, t = threadIdx.x;
  const float pk = pred[k], yk = y[k];
  double rank_sum = 0.0, grad = 0.0;
  for (int j = t; j < n; j += 256) {
    const float pj = pred[j], yj = y[j];
    const float tkj = yk - yj, mkj = (tkj > 0.f) - (tkj < 0.f);
    const float v = tkj - mkj * (pk - pj);                 // pair (k, j)
    if (v > 0.f) { rank_sum += v; grad -= mkj; }
    const float tjk = -tkj, mjk = -mkj;
    const float u = tjk - mjk * (pj - pk);                 // pair (j, k): contributes +m_jk to dL/dp_k
    if (u > 0.f) grad += mjk;
  }
  r1[t] = rank_sum; r2[t] = grad; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) { r1[t] += r1[t + o]; r2[t] += r2[t + o]; } __syncthreads(); }
  if (t == 0) {
    const double nn = n > 1 ? (double)n * (n - 1) : 1.0;
    const float e = pk - yk;
    dpred[k] = l1_w * (float)((e > 0.f) - (e < 0.f)) / n + rank_w * (float)(r2[0] / nn);
    loss_rows[k] = l1_w * fabsf(e) / n + rank_w * (float)(r1[0] / nn);
  }
}

__global__ void tr_sum_rows(const float* __restrict__ rows, int n, float* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += rows[i];
    *out = (float)s;
  }
}

// torch.optim.SGD: g += wd p; buf = g (first step) or mu buf + g; p -= lr buf
__global__ void __launch_bounds__(256)
tr_sgd(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, int64_t n, float lr, float mu, float wd, int first) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  const float gr = fmaf(wd, p[idx], g[idx]);
  const float b = first ? gr : fmaf(mu, mom[idx], gr);
  mom[idx] = b;
  p[idx] = fmaf(-lr, b, p[idx]);
}

// AveragedModel.update_parameters: avg = p (first) or avg + (p - avg) / (n_averaged + 1)
__global__ void __launch_bounds__(256)
tr_swa(float* __restrict__ avg, const float* __restrict__ p, int64_t n, float inv_np1, int first) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  avg[idx] = first ? p[idx] : avg[idx] + (p[idx] - avg[idx]) * inv_np1;
}

static int tr_alloc(Trainer* t, float** ptr, size_t floats) {
  VQA_CUDA(cudaMalloc((void**)ptr, floats * sizeof(float)));
  t->allocs.push_back(*ptr);
  return B200VQA_OK;
}

// C = A op B with split-K when the output alone cannot fill the GPU (fc1 forward: 4 x 4 tiles of K = 35,203)
static int tr_contract(Trainer* t, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs, float* C,
                       int M, int N, int K, const float* bias, cudaStream_t st) {
  const int tiles = cdiv(M, TG_T) * cdiv(N, TG_T);
  int splits = 1;
  if (tiles < 296 && K >= 1024) splits = min(32, min(cdiv(K, 512), max(1, 592 / tiles)));      // partials: <= 32 x B x h1 floats
  int kps = cdiv(cdiv(K, splits), TG_K) * TG_K;
  splits = cdiv(K, kps);
  const dim3 grid(cdiv(N, TG_T), cdiv(M, TG_T), splits);
  if (splits == 1 && !bias) {
    tr_gemm<<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, M, N, K, kps);
    VQA_LAUNCH_CHECK();
    return B200VQA_OK;
  }
  const size_t need = (size_t)splits * M * N;
  if (need > t->partial_floats) return B200VQA_ENOMEM;
  tr_gemm<<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, t->partial, M, N, K, kps);
  VQA_LAUNCH_CHECK();
  const size_t mn = (size_t)M * N;
  tr_reduce_splits<<<(unsigned)((mn + 255) / 256), 256, 0, st>>>(t->partial, splits, mn, N, bias, C);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}

static int tr_reserve_batch(Trainer* t, int B) {
  if (B <= t->cap_b) return B200VQA_OK;
  // activations are small (B x 256): allocate for the new capacity; old buffers stay in `allocs` until destroy
  const int cap = B + B / 4 + 8;
  int rc;
  if ((rc = tr_alloc(t, &t->z1, (size_t)cap * t->h1)) || (rc = tr_alloc(t, &t->xhat, (size_t)cap * t->h1)) ||
      (rc = tr_alloc(t, &t->a1, (size_t)cap * t->h1)) || (rc = tr_alloc(t, &t->d1, (size_t)cap * t->h1)) ||
      (rc = tr_alloc(t, &t->z2, (size_t)cap * t->h2)) || (rc = tr_alloc(t, &t->a2, (size_t)cap * t->h2)) ||
      (rc = tr_alloc(t, &t->d2, (size_t)cap * t->h2)) || (rc = tr_alloc(t, &t->pred, cap)) || (rc = tr_alloc(t, &t->dpred, cap)) ||
      (rc = tr_alloc(t, &t->loss_rows, cap)))
    return rc;
  // split-K partials: fc1 forward needs splits x B x h1, the weight gradient never splits (h1 x in outputs)
  const size_t pf = (size_t)40 * cap * t->h1 + 1024;
  if ((rc = tr_alloc(t, &t->partial, pf))) return rc;
  t->partial_floats = pf;
  t->cap_b = cap;
  return B200VQA_OK;
}

// forward up to the predictions; train: batch statistics (+ running update), dropout masks; eval: running statistics
static int tr_forward(Trainer* t, const float* params, float* run_mean, float* run_var, const float* X, int B, int train, float bn_momentum,
                      const uint8_t* drop1, const uint8_t* drop2, float keep_scale, cudaStream_t st) {
  const float* fc1w = params + t->o_fc1w; const float* fc1b = params + t->o_fc1b;
  const float* bnw = params + t->o_bnw; const float* bnb = params + t->o_bnb;
  const float* fc2w = params + t->o_fc2w; const float* fc2b = params + t->o_fc2b;
  const float* fc3w = params + t->o_fc3w; const float* fc3b = params + t->o_fc3b;
  int rc;
  if ((rc = tr_contract(t, X, t->in, 1, fc1w, 1, t->in, t->z1, B, t->h1, t->in, fc1b, st))) return rc;         // z1 = X W1^T + b1
  const size_t n1 = (size_t)B * t->h1, n2 = (size_t)B * t->h2;
  if (train) {
    tr_bn_gelu_drop_fwd<<<t->h1, 256, 0, st>>>(t->z1, B, t->h1, bnw, bnb, run_mean, run_var, bn_momentum, 1, 1, t->xhat, t->bmean,
                                                 t->brstd, drop1, keep_scale, t->a1);
  } else {
    tr_bn_gelu_eval<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(t->z1, n1, t->h1, bnw, bnb, run_mean, run_var, t->a1);
  }
  VQA_LAUNCH_CHECK();
  if ((rc = tr_contract(t, t->a1, t->h1, 1, fc2w, 1, t->h1, t->z2, B, t->h2, t->h1, fc2b, st))) return rc;       // z2 = a1 W2^T + b2
  tr_gelu_drop_fwd<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(t->z2, n2, train ? drop2 : nullptr, keep_scale, t->a2);
  VQA_LAUNCH_CHECK();
  return tr_contract(t, t->a2, t->h2, 1, fc3w, 1, t->h2, t->pred, B, 1, t->h2, fc3b, st);                        // pred = a2 w3 + b3
}

}  // namespace b200vqa

using namespace b200vqa;

struct b200vqa_trainer { Trainer t; b200vqa_ctx* ctx; };

extern "C" int b200vqa_trainer_create(b200vqa_t* h, int in_features, int hidden, b200vqa_trainer_t** out) {
  if (!h || !out || in_features <= 0 || hidden <= 0 || hidden % TR_H2_DIV) return B200VQA_EINVAL;
  CtxScope scope(h);
  b200vqa_trainer* tr = new b200vqa_trainer();
  tr->ctx = h;
  Trainer& t = tr->t;
  t.device = h->device; t.in = in_features; t.h1 = hidden; t.h2 = hidden / TR_H2_DIV;
  int64_t o = 0;
  t.o_fc1w = o; o += (int64_t)t.h1 * t.in;
  t.o_fc1b = o; o += t.h1;
  t.o_bnw = o; o += t.h1;
  t.o_bnb = o; o += t.h1;
  t.o_fc2w = o; o += (int64_t)t.h2 * t.h1;
  t.o_fc2b = o; o += t.h2;
  t.o_fc3w = o; o += t.h2;
  t.o_fc3b = o; o += 1;
  t.n_params = o;
  int rc = 0;
  float** bufs[] = {&t.p, &t.g, &t.mom, &t.swa};
  for (float** b : bufs)
    if (!rc) rc = tr_alloc(&t, b, (size_t)o);
  float** stats[] = {&t.bn_mean, &t.bn_var, &t.swa_bn_mean, &t.swa_bn_var, &t.bmean, &t.brstd};
  for (float** b : stats)
    if (!rc) rc = tr_alloc(&t, b, (size_t)t.h1);
  if (!rc && cudaMemset(t.mom, 0, o * sizeof(float)) != cudaSuccess) rc = B200VQA_ECUDA;
  if (!rc && cudaMemset(t.swa, 0, o * sizeof(float)) != cudaSuccess) rc = B200VQA_ECUDA;
  if (rc) { for (void* p : t.allocs) cudaFree(p); delete tr; return rc; }
  *out = tr;
  return B200VQA_OK;
}

extern "C" int b200vqa_trainer_destroy(b200vqa_trainer_t* tr) {
  if (!tr) return B200VQA_EINVAL;
  cudaSetDevice(tr->t.device);
  cudaDeviceSynchronize();
  for (void* p : tr->t.allocs) cudaFree(p);
  delete tr;
  return B200VQA_OK;
}

// which: 0 = the model being trained, 1 = the SWA average.  Host float32 arrays in state-dict layout.
static int tr_copy_params(b200vqa_trainer* tr, int which, int to_device, float* fc1_w, float* fc1_b, float* bn_w, float* bn_b, float* bn_mean,
                          float* bn_var, float* fc2_w, float* fc2_b, float* fc3_w, float* fc3_b) {
  Trainer& t = tr->t;
  VQA_CUDA(cudaSetDevice(t.device));
  VQA_CUDA(cudaDeviceSynchronize());
  float* base = which ? t.swa : t.p;
  struct Item { float* host; float* dev; size_t n; } items[] = {
      {fc1_w, base + t.o_fc1w, (size_t)t.h1 * t.in}, {fc1_b, base + t.o_fc1b, (size_t)t.h1}, {bn_w, base + t.o_bnw, (size_t)t.h1},
      {bn_b, base + t.o_bnb, (size_t)t.h1}, {bn_mean, which ? t.swa_bn_mean : t.bn_mean, (size_t)t.h1},
      {bn_var, which ? t.swa_bn_var : t.bn_var, (size_t)t.h1}, {fc2_w, base + t.o_fc2w, (size_t)t.h2 * t.h1},
      {fc2_b, base + t.o_fc2b, (size_t)t.h2}, {fc3_w, base + t.o_fc3w, (size_t)t.h2}, {fc3_b, base + t.o_fc3b, 1}};
  for (const Item& it : items) {
    if (!it.host) return B200VQA_EINVAL;
    if (to_device) VQA_CUDA(cudaMemcpy(it.dev, it.host, it.n * sizeof(float), cudaMemcpyHostToDevice));
    else VQA_CUDA(cudaMemcpy(it.host, it.dev, it.n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return B200VQA_OK;
}

extern "C" int b200vqa_trainer_set_params(b200vqa_trainer_t* tr, const float* h_fc1_w, const float* h_fc1_b, const float* h_bn_w,
                                          const float* h_bn_b, const float* h_bn_mean, const float* h_bn_var, const float* h_fc2_w,
                                          const float* h_fc2_b, const float* h_fc3_w, const float* h_fc3_b) {
  if (!tr) return B200VQA_EINVAL;
  int rc = tr_copy_params(tr, 0, 1, (float*)h_fc1_w, (float*)h_fc1_b, (float*)h_bn_w, (float*)h_bn_b, (float*)h_bn_mean, (float*)h_bn_var,
                          (float*)h_fc2_w, (float*)h_fc2_b, (float*)h_fc3_w, (float*)h_fc3_b);
  if (rc) return rc;
  Trainer& t = tr->t;
  t.steps = 0; t.n_averaged = 0;
  VQA_CUDA(cudaMemset(t.mom, 0, t.n_params * sizeof(float)));
  // AveragedModel(model) deep-copies the model: parameters and buffers (src/model_regression.py:388)
  VQA_CUDA(cudaMemcpy(t.swa, t.p, t.n_params * sizeof(float), cudaMemcpyDeviceToDevice));
  VQA_CUDA(cudaMemcpy(t.swa_bn_mean, t.bn_mean, t.h1 * sizeof(float), cudaMemcpyDeviceToDevice));
  VQA_CUDA(cudaMemcpy(t.swa_bn_var, t.bn_var, t.h1 * sizeof(float), cudaMemcpyDeviceToDevice));
  return B200VQA_OK;
}

extern "C" int b200vqa_trainer_get_params(b200vqa_trainer_t* tr, int which, float* h_fc1_w, float* h_fc1_b, float* h_bn_w, float* h_bn_b,
                                          float* h_bn_mean, float* h_bn_var, float* h_fc2_w, float* h_fc2_b, float* h_fc3_w, float* h_fc3_b) {
  if (!tr || which < 0 || which > 1) return B200VQA_EINVAL;
  return tr_copy_params(tr, which, 0, h_fc1_w, h_fc1_b, h_bn_w, h_bn_b, h_bn_mean, h_bn_var, h_fc2_w, h_fc2_b, h_fc3_w, h_fc3_b);
}

extern "C" int b200vqa_trainer_step(b200vqa_trainer_t* tr, const float* X, const float* y, int B, const uint8_t* drop1, const uint8_t* drop2,
                                    float drop_rate, float lr, float momentum, float weight_decay, float l1_w, float rank_w, float* loss_out,
                                    void* stream) {
  if (!tr || !X || !y || B < 2 || drop_rate < 0.f || drop_rate >= 1.f) return B200VQA_EINVAL;
  Trainer& t = tr->t;
  CtxScope scope(tr->ctx);
  cudaStream_t st = as_stream(stream);
  int rc = tr_reserve_batch(&t, B);
  if (rc) return rc;
  const float keep_scale = 1.f / (1.f - drop_rate);
  if (drop_rate == 0.f) drop1 = drop2 = nullptr;
  if ((rc = tr_forward(&t, t.p, t.bn_mean, t.bn_var, X, B, 1, 0.1f, drop1, drop2, keep_scale, st))) return rc;
  tr_loss_grad<<<B, 256, 0, st>>>(t.pred, y, B, l1_w, rank_w, t.dpred, t.loss_rows);
  VQA_LAUNCH_CHECK();
  if (loss_out) { tr_sum_rows<<<1, 1, 0, st>>>(t.loss_rows, B, loss_out); VQA_LAUNCH_CHECK(); }
  float* g = t.g;
  const size_t n1 = (size_t)B * t.h1, n2 = (size_t)B * t.h2;
  // fc3: dW3 = dpred^T a2, db3 = sum dpred, d(a2) = dpred w3
  if ((rc = tr_contract(&t, t.dpred, 1, 1, t.a2, t.h2, 1, g + t.o_fc3w, 1, t.h2, B, nullptr, st))) return rc;
  tr_colsum<<<1, 256, 0, st>>>(t.dpred, B, 1, g + t.o_fc3b); VQA_LAUNCH_CHECK();
  if ((rc = tr_contract(&t, t.dpred, 1, 1, t.p + t.o_fc3w, t.h2, 1, t.d2, B, t.h2, 1, nullptr, st))) return rc;
  tr_gelu_drop_bwd<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(t.d2, t.z2, n2, drop2, keep_scale); VQA_LAUNCH_CHECK();
  // fc2: dW2 = d2^T a1, db2 = colsum d2, d(a1) = d2 W2
  if ((rc = tr_contract(&t, t.d2, 1, t.h2, t.a1, t.h1, 1, g + t.o_fc2w, t.h2, t.h1, B, nullptr, st))) return rc;
  tr_colsum<<<cdiv(t.h2, 256), 256, 0, st>>>(t.d2, B, t.h2, g + t.o_fc2b); VQA_LAUNCH_CHECK();
  if ((rc = tr_contract(&t, t.d2, t.h2, 1, t.p + t.o_fc2w, t.h1, 1, t.d1, B, t.h1, t.h2, nullptr, st))) return rc;
  tr_bn_gelu_drop_bwd<<<t.h1, 256, 0, st>>>(t.d1, t.xhat, B, t.h1, t.p + t.o_bnw, t.p + t.o_bnb, t.brstd, drop1, keep_scale, g + t.o_bnw,
                                             g + t.o_bnb);
  VQA_LAUNCH_CHECK();
  // fc1: dW1 = d1^T X (the big one: h1 x in outputs, reduction over the batch), db1 = colsum d1
  if ((rc = tr_contract(&t, t.d1, 1, t.h1, X, t.in, 1, g + t.o_fc1w, t.h1, t.in, B, nullptr, st))) return rc;
  tr_colsum<<<cdiv(t.h1, 256), 256, 0, st>>>(t.d1, B, t.h1, g + t.o_fc1b); VQA_LAUNCH_CHECK();
  tr_sgd<<<(unsigned)((t.n_params + 255) / 256), 256, 0, st>>>(t.p, g, t.mom, t.n_params, lr, momentum, weight_decay, t.steps == 0);
  VQA_LAUNCH_CHECK();
  ++t.steps;
  (void)n1;
  return B200VQA_OK;
}

extern "C" int b200vqa_trainer_swa_update(b200vqa_trainer_t* tr, void* stream) {
  if (!tr) return B200VQA_EINVAL;
  Trainer& t = tr->t;
  CtxScope scope(tr->ctx);
  tr_swa<<<(unsigned)((t.n_params + 255) / 256), 256, 0, as_stream(stream)>>>(t.swa, t.p, t.n_params, 1.f / (float)(t.n_averaged + 1),
                                                                              t.n_averaged == 0);
  VQA_LAUNCH_CHECK();
  ++t.n_averaged;
  return B200VQA_OK;
}

extern "C" int b200vqa_trainer_predict(b200vqa_trainer_t* tr, int which, const float* X, int B, float* pred, void* stream) {
  if (!tr || !X || !pred || B <= 0 || which < 0 || which > 1) return B200VQA_EINVAL;
  Trainer& t = tr->t;
  CtxScope scope(tr->ctx);
  cudaStream_t st = as_stream(stream);
  int rc = tr_reserve_batch(&t, B);
  if (rc) return rc;
  if ((rc = tr_forward(&t, which ? t.swa : t.p, which ? t.swa_bn_mean : t.bn_mean, which ? t.swa_bn_var : t.bn_var, X, B, 0, 0.f, nullptr,
                       nullptr, 1.f, st)))
    return rc;
  VQA_CUDA(cudaMemcpyAsync(pred, t.pred, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return B200VQA_OK;
}

// torch.optim.swa_utils.update_bn, one batch: batch_index 0 resets the running statistics (mean 0, var 1), then each batch
// enters the cumulative average (momentum = 1 / (batches seen)).
extern "C" int b200vqa_trainer_update_bn(b200vqa_trainer_t* tr, int which, const float* X, int B, int batch_index, void* stream) {
  if (!tr || !X || B < 2 || which < 0 || which > 1 || batch_index < 0) return B200VQA_EINVAL;
  Trainer& t = tr->t;
  CtxScope scope(tr->ctx);
  cudaStream_t st = as_stream(stream);
  int rc = tr_reserve_batch(&t, B);
  if (rc) return rc;
  const float* params = which ? t.swa : t.p;
  float* rm = which ? t.swa_bn_mean : t.bn_mean;
  float* rv = which ? t.swa_bn_var : t.bn_var;
  if ((rc = tr_contract(&t, X, t.in, 1, params + t.o_fc1w, 1, t.in, t.z1, B, t.h1, t.in, params + t.o_fc1b, st))) return rc;
  tr_bn_gelu_drop_fwd<<<t.h1, 256, 0, st>>>(t.z1, B, t.h1, params + t.o_bnw, params + t.o_bnb, rm, rv, 1.f / (float)(batch_index + 1), 1, 0,
                                             t.xhat, t.bmean, t.brstd, nullptr, 1.f, t.a1);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}
