// Shared helpers for libb200vqa (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b200vqa.h"

namespace b200vqa {

void set_last_error(const char* what, cudaError_t e);
extern thread_local int64_t* g_launch_counter;   // points into the active context (or a dummy)

inline void count_launch(int n = 1) { if (g_launch_counter) *g_launch_counter += n; }

#define VQA_CUDA(expr)                                              \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) {                                        \
      ::b200vqa::set_last_error(#expr, _e);                         \
      return B200VQA_ECUDA;                                         \
    }                                                               \
  } while (0)

#define VQA_LAUNCH_CHECK()                                          \
  do {                                                              \
    ::b200vqa::count_launch();                                      \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) {                                        \
      ::b200vqa::set_last_error("kernel launch", _e);               \
      return B200VQA_ECUDA;                                         \
    }                                                               \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace b200vqa
