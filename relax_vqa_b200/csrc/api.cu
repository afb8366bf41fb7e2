// Context management and error reporting for libb200vqa.
#include <string.h>
#include <string>
#include <stdlib.h>
#include "context.h"

namespace b200vqa {

thread_local int64_t* g_launch_counter = nullptr;
thread_local b200vqa_ctx* g_ctx = nullptr;
static thread_local char g_last_error[512] = "";

void set_last_error(const char* what, cudaError_t e) {
  snprintf(g_last_error, sizeof g_last_error, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

int DeviceBuffer::reserve(size_t n) {
  if (n <= bytes) return B200VQA_OK;
  if (ptr) { cudaFree(ptr); ptr = nullptr; bytes = 0; }
  size_t want = n + n / 8 + 256;
  cudaError_t e = cudaMalloc(&ptr, want);
  if (e != cudaSuccess) { set_last_error("workspace cudaMalloc", e); ptr = nullptr; return B200VQA_ENOMEM; }
  bytes = want;
  return B200VQA_OK;
}
void DeviceBuffer::release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }

}  // namespace b200vqa

using namespace b200vqa;

extern "C" int b200vqa_version(void) { return 100; }

extern "C" const char* b200vqa_error_string(int code) {
  switch (code) {
    case B200VQA_OK: return "ok";
    case B200VQA_EINVAL: return "invalid argument";
    case B200VQA_ECUDA: return "CUDA error";
    case B200VQA_ENOTLOADED: return "weights not loaded";
    case B200VQA_ENOMEM: return "out of device memory";
    default: return "unknown error";
  }
}

extern "C" const char* b200vqa_last_error(void) { return g_last_error; }

extern "C" int b200vqa_create(int device, b200vqa_t** out) {
  if (!out) return B200VQA_EINVAL;
  int count = 0;
  VQA_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return B200VQA_EINVAL;
  VQA_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VQA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    snprintf(g_last_error, sizeof g_last_error, "device %d is sm_%d%d; libb200vqa is built for sm_100a only", device, prop.major, prop.minor);
    return B200VQA_ECUDA;
  }
  int rc;
  if ((rc = flow_init_device_attrs()) || (rc = gemm_init_device_attrs()) || (rc = vit_init_device_attrs())) return rc;
  b200vqa_ctx* h = new b200vqa_ctx();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("B200VQA_FLOW_IMPL")) h->flow_impl = atoi(e);      // A/B switch for tools / bench
  *out = h;
  return B200VQA_OK;
}

extern "C" int b200vqa_destroy(b200vqa_t* h) {
  if (!h) return B200VQA_EINVAL;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->resize_tables) { cudaFree(kv.second.d_bounds); cudaFree(kv.second.d_kk); cudaFree(kv.second.d_kkT); }
  h->ws_resize.release(); h->ws_flow.release(); h->ws_resnet.release(); h->ws_vit.release(); h->ws_head.release(); h->ws_misc.release();
  for (auto& ev : h->prof_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto& ev : h->prof_events_flow) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto& ev : h->prof_pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  free_resnet(h->resnet); free_vit(h->vit); free_head(h->head);
  delete h;
  return B200VQA_OK;
}

extern "C" int64_t b200vqa_launch_count(b200vqa_t* h) { return h ? h->launches : -1; }

extern "C" int b200vqa_set_gemm_impl(b200vqa_t* h, int impl) {
  if (!h || impl < 0 || impl > 2) return B200VQA_EINVAL;
  h->gemm_impl = impl;
  return B200VQA_OK;
}

extern "C" int b200vqa_set_attn_impl(b200vqa_t* h, int impl) {
  if (!h || impl < 0 || impl > 2) return B200VQA_EINVAL;
  h->attn_impl = impl;
  return B200VQA_OK;
}

extern "C" int b200vqa_set_gemm_sms(b200vqa_t* h, int sms) {
  if (!h || sms < 0 || (sms & 1)) return B200VQA_EINVAL;
  h->gemm_sms = sms;
  return B200VQA_OK;
}

extern "C" int b200vqa_set_flow_impl(b200vqa_t* h, int impl) {
  if (!h || impl < 0 || impl > 5) return B200VQA_EINVAL;
  h->flow_impl = impl;
  return B200VQA_OK;
}

extern "C" int b200vqa_set_profiling(b200vqa_t* h, int on) {
  if (!h) return B200VQA_EINVAL;
  h->profiling = on ? 1 : 0;
  return B200VQA_OK;
}

extern "C" int b200vqa_profile_read_flow(b200vqa_t* h, double* ms_out, int64_t* launches, double* bytes) {
  if (!h) return B200VQA_EINVAL;
  VQA_CUDA(cudaSetDevice(h->device));
  VQA_CUDA(cudaDeviceSynchronize());
  double ms = 0.0;
  for (auto& ev : h->prof_events_flow) {
    float t = 0.f;
    VQA_CUDA(cudaEventElapsedTime(&t, ev.first, ev.second));
    ms += t;
    h->prof_pool.push_back(ev);
  }
  if (ms_out) *ms_out = ms;
  if (launches) *launches = (int64_t)h->prof_events_flow.size();
  if (bytes) *bytes = h->prof_bytes_flow;
  h->prof_events_flow.clear();
  h->prof_bytes_flow = 0.0;
  return B200VQA_OK;
}

extern "C" int b200vqa_profile_read(b200vqa_t* h, double* gemm_ms, int64_t* gemm_launches, double* gemm_flops) {
  if (!h) return B200VQA_EINVAL;
  VQA_CUDA(cudaSetDevice(h->device));
  VQA_CUDA(cudaDeviceSynchronize());
  double ms = 0.0;
  for (auto& ev : h->prof_events) {
    float t = 0.f;
    VQA_CUDA(cudaEventElapsedTime(&t, ev.first, ev.second));
    ms += t;
    h->prof_pool.push_back(ev);
  }
  if (gemm_ms) *gemm_ms = ms;
  if (gemm_launches) *gemm_launches = (int64_t)h->prof_events.size();
  if (gemm_flops) *gemm_flops = h->prof_flops;
  h->prof_events.clear();
  h->prof_flops = 0.0;
  return B200VQA_OK;
}
