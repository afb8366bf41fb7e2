// Minimal repro for the one hazard compute-sanitizer racecheck reports in libb200vqa (gemm2cta_tcgen05_kernel):
// a cluster of two CTAs, one warp per CTA executes tcgen05.alloc.cta_group::2 (the CUTLASS Allocator2Sm protocol), every
// thread reads the slot only after a CTA barrier AND a cluster barrier, then the pair frees the columns.  Nothing else touches
// shared memory.  racecheck flags the alloc instruction's own shared-memory operand against the hardware's write of the
// allocated address (performed for both CTAs of the pair, hence "Write Thread (block rank 1)"): the report is about the
// instruction itself, not about the kernel around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o racecheck_repro tools/racecheck_repro.cu
//   compute-sanitizer --tool racecheck ./racecheck_repro
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) alloc_pair(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  if (threadIdx.x == 0) out[blockIdx.x] = base;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
  }
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 8 * sizeof(uint32_t));
  alloc_pair<<<8, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  uint32_t h[8];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("status %s; tmem base per CTA:", cudaGetErrorString(e));
  for (int i = 0; i < 8; ++i) printf(" %u", h[i]);
  printf("\n");
  return e != cudaSuccess;
}
