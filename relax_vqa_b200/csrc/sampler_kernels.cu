// Frame sampling from raw yuv420p video: the pixel half of the reference's
//   ffmpeg -s WxH -pix_fmt yuv420p -i x.yuv -vf select='not(mod(n,k))' ... x_%d.png        (src/video_frames_extract.py:29-49)
//   ffmpeg ... -vf select='not(mod(n-1,k))' ... x_%d_next.png                             (:76-100)
// i.e. libswscale's unscaled yuv420p -> rgb24 / bgr24 conversion of the selected frames (the frame selection itself is
// index arithmetic on the host, relax_vqa_b200/video_frames_extract.py).  Third-party algorithm (FFmpeg libswscale, not in
// /root/reference): x86 SIMD path of yuv2rgb (libswscale/x86/yuv_2_rgb.asm with the coefficients of
// ff_yuv2rgb_c_init_tables, ITU-R BT.601, limited range): chroma is NOT interpolated (each U, V sample serves its 2 x 2
// luma block), all arithmetic is signed 16-bit with pmulhw (high half of the 32-bit product):
//     y = (((Y << 3) - 128) * 9539) >> 16          u = (U << 3) - 1024      v = (V << 3) - 1024
//     R = sat8(y + ((v * 13075) >> 16))   G = sat8(y + ((u * -3209) >> 16) + ((v * -6660) >> 16))   B = sat8(y + ((u * 16525) >> 16))
// Bit-identical to swscale 9.1 (the FFmpeg inside this image's OpenCV) on random planes: oracle/sampler.py,
// tests/test_oracle_sampler.py.
#include "common.cuh"

namespace b200vqa {

__device__ __forceinline__ int mulhi16(int a, int b) { return (a * b) >> 16; }      // pmulhw on sign-extended 16-bit lanes
__device__ __forceinline__ uint32_t sat8(int v) { return (uint32_t)min(max(v, 0), 255); }

// one thread = 4 horizontally adjacent pixels of one row (one 32-bit word of Y, two U and two V samples, 12 output bytes)
__global__ void __launch_bounds__(256)
k0_yuv420p_to_bgr(const uint8_t* __restrict__ yuv, int H, int W, uint8_t* __restrict__ bgr) {
  const int wq = W >> 2;
  const int xq = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (xq >= wq) return;
  const size_t frame_bytes = (size_t)H * W * 3 / 2;
  const uint8_t* base = yuv + (size_t)blockIdx.z * frame_bytes;
  const uint8_t* up = base + (size_t)H * W + (size_t)(y >> 1) * (W >> 1) + 2 * xq;
  const uint8_t* vp = up + (size_t)(H >> 1) * (W >> 1);
  const uint32_t yw = *reinterpret_cast<const uint32_t*>(base + (size_t)y * W + 4 * xq);
  const uint32_t uw = *reinterpret_cast<const uint16_t*>(up), vw = *reinterpret_cast<const uint16_t*>(vp);
  uint8_t px[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int Y = (yw >> (8 * i)) & 0xff, U = (uw >> (8 * (i >> 1))) & 0xff, V = (vw >> (8 * (i >> 1))) & 0xff;
    const int yy = mulhi16((Y << 3) - 128, 9539), u = (U << 3) - 1024, v = (V << 3) - 1024;
    px[3 * i] = (uint8_t)sat8(yy + mulhi16(u, 16525));
    px[3 * i + 1] = (uint8_t)sat8(yy + mulhi16(u, -3209) + mulhi16(v, -6660));
    px[3 * i + 2] = (uint8_t)sat8(yy + mulhi16(v, 13075));
  }
  uint32_t* o = reinterpret_cast<uint32_t*>(bgr + ((size_t)blockIdx.z * H * W + (size_t)y * W + 4 * xq) * 3);
  o[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
  o[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
  o[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
}

}  // namespace b200vqa

using namespace b200vqa;

extern "C" int b200vqa_yuv420p_to_bgr(const uint8_t* yuv, int B, int H, int W, uint8_t* bgr, void* stream) {
  if (!yuv || !bgr || B <= 0 || H <= 0 || W <= 0 || (W & 3) || (H & 1)) return B200VQA_EINVAL;     // swscale's unscaled SIMD path: even height
  k0_yuv420p_to_bgr<<<dim3(cdiv(W >> 2, 256), H, B), 256, 0, as_stream(stream)>>>(yuv, H, W, bgr);
  VQA_LAUNCH_CHECK();
  return B200VQA_OK;
}
