// SIMT check kernels: the same arithmetic as the tcgen05 epilogues, one thread per output.
// Selected with b200vqa_set_gemm_impl(h, 1); used only to validate the tensor-core path on
// the device (tests) - never the default.
#pragma once
#include "gemm_tcgen05.cuh"

namespace b200vqa {

__global__ void ref_gemm_rowmajor(const __half* __restrict__ A, const __half* __restrict__ B, const float* __restrict__ bias,
                                  const float* residual, void* out, int M, int N, int K, int ldo, int act, int out_is_f32) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(__half2float(A[(size_t)m * K + k]), __half2float(B[(size_t)n * K + k]), acc);
  if (bias) acc += bias[n];
  if (act == ACT_GELU) acc = gelu_erf(acc); else if (act == ACT_RELU) acc = fmaxf(acc, 0.f);
  const size_t o = (size_t)m * ldo + n;
  if (residual) acc += residual[o];
  if (out_is_f32) static_cast<float*>(out)[o] = acc; else static_cast<__half*>(out)[o] = __float2half_rn(acc);
}

// Direct convolution, NHWC fp16 in/out, weights [Cout][R][S][Cin]; one thread per (pixel, channel).
// gap_sum (optional, [Nimg][Cout], pre-zeroed) is accumulated with atomics (check path only).
__global__ void ref_conv_nhwc(const __half* __restrict__ in, const __half* __restrict__ w, const float* __restrict__ scale,
                              const float* __restrict__ shift, const __half* __restrict__ identity, __half* __restrict__ out,
                              float* gap_sum, int gap_raw, int Nimg, int Hin, int Win, int Cin, int Hout, int Wout, int Cout,
                              int R, int S, int stride, int pad, int act) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)Nimg * Hout * Wout * Cout;
  if (idx >= total) return;
  const int c = idx % Cout;
  size_t pix = idx / Cout;
  const int x = pix % Wout; pix /= Wout;
  const int y = pix % Hout;
  const int n = pix / Hout;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) {
    const int iy = y * stride + r - pad;
    if (iy < 0 || iy >= Hin) continue;
    for (int s = 0; s < S; ++s) {
      const int ix = x * stride + s - pad;
      if (ix < 0 || ix >= Win) continue;
      const __half* ip = in + (((size_t)n * Hin + iy) * Win + ix) * Cin;
      const __half* wp = w + (((size_t)c * R + r) * S + s) * Cin;
      for (int k = 0; k < Cin; ++k) acc = fmaf(__half2float(ip[k]), __half2float(wp[k]), acc);
    }
  }
  float val = fmaf(acc, scale[c], shift[c]);
  if (identity) val += __half2float(identity[idx]);
  if (act == ACT_RELU) val = fmaxf(val, 0.f);
  out[idx] = __float2half_rn(val);
  if (gap_sum) atomicAdd(&gap_sum[(size_t)n * Cout + c], gap_raw ? acc : val);
}

}  // namespace b200vqa
