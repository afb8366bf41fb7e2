// The opaque b200vqa_ctx: device, weights, resize tables, grow-only workspaces.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "common.cuh"

namespace b200vqa {

struct ResizeTable {          // Pillow coefficients for one (in_size -> 224, filter)
  int ksize = 0;
  int* d_bounds = nullptr;    // [224][2] (xmin, count)
  int* d_kk = nullptr;        // [224][ksize] 22-bit fixed point (row-major: vertical pass, broadcast reads)
  int* d_kkT = nullptr;       // [ksize][224] tap-major copy (horizontal pass: coalesced across output columns)
};

struct DeviceBuffer {         // grow-only scratch
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t n);
  void release();
};

struct ResNetWeights;         // nn_resnet.cu
struct ViTWeights;            // nn_vit.cu
struct HeadWeights;           // nn_head.cu

}  // namespace b200vqa

struct b200vqa_ctx {
  int device = 0;
  int sm_count = 148;
  int gemm_sms = 0;                        // > 0: persistent tcgen05 grids use this many SMs (b200vqa_set_gemm_sms)
  int conv_no_halo = 0;                    // set while the raw-map boundary path runs (generic per-tap convolutions there)
  int gemm_impl = 0;                       // 0 tcgen05, 1 SIMT check kernels
  int attn_impl = 0;                       // 0 tcgen05 / TMEM attention, 1 warp-level mma.sync attention (A/B)
  int flow_impl = 0;                       // 0 streaming strip kernels (f64 running sums), 1 same with Kahan fp32 sums, 2 tile kernels,
                                           // 3 192-thread strips, 4 software-pipelined variant (march3), 5 same with the first version's solve
  int64_t launches = 0;
  int profiling = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;          // class 0: tcgen05 GEMM / conv launches
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events_flow;     // class 1: k4_flow_iter launches
  double prof_bytes_flow = 0.0;                                          // algorithmic bytes of those launches
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;
  double prof_flops = 0.0;
  std::map<std::pair<int, int>, b200vqa::ResizeTable> resize_tables;   // (in_size, filter)
  b200vqa::DeviceBuffer ws_resize, ws_flow, ws_resnet, ws_vit, ws_head, ws_misc;
  b200vqa::ResNetWeights* resnet = nullptr;
  b200vqa::ViTWeights* vit = nullptr;
  b200vqa::HeadWeights* head = nullptr;
};

namespace b200vqa {
extern thread_local b200vqa_ctx* g_ctx;
struct CtxScope {             // routes launch counting to the context for the current call
  explicit CtxScope(b200vqa_ctx* h) {
    g_launch_counter = h ? &h->launches : nullptr; g_ctx = h;
    if (h) cudaSetDevice(h->device);     // every handle-taking entry point runs on the context's device, whatever the caller's current one
  }
  ~CtxScope() { g_launch_counter = nullptr; g_ctx = nullptr; }
};
// SMs a persistent tcgen05 grid of this context may occupy (even, >= 2)
inline int gemm_grid_sms(const b200vqa_ctx* h) { return (h->gemm_sms > 0 && h->gemm_sms < h->sm_count) ? h->gemm_sms : h->sm_count; }
// cudaFuncSetAttribute is per device: each translation unit sets its kernels' attributes for the CURRENT device
int flow_init_device_attrs();
int gemm_init_device_attrs();
int vit_init_device_attrs();
void free_resnet(ResNetWeights*);
void free_vit(ViTWeights*);
void free_head(HeadWeights*);
}  // namespace b200vqa
