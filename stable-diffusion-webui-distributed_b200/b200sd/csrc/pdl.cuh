// pdl.cuh — programmatic dependent launch (PDL) for the kernel chains of one UNet evaluation / VAE pass.
//
// A UNet evaluation is ~350 kernels on one stream (replayed from a CUDA graph).  Back to back, every boundary costs the
// launch latency of the next grid plus its prologue (mbarrier init, TMEM allocation, tensor-map prefetch) — a few
// microseconds each, i.e. 1-2 ms per evaluation whatever the batch, which is 2 % of an evaluation at the benchmark batch
// and a quarter of it at one image per GPU.  With PDL the next grid's CTAs become resident as soon as every CTA of the
// running grid has executed `griddepcontrol.launch_dependents` (placed at kernel entry) and SM resources free up; they run
// their prologue and then block in `griddepcontrol.wait` until the previous grid has COMPLETED and its memory is visible.
//
// Rules every kernel launched through launch_pdl() follows (they make the chain transitively safe):
//   * pdl_wait() is executed by every thread before the first global-memory access (read OR write) of the kernel;
//   * nothing before pdl_wait() touches global memory (shared memory, TMEM, barriers, descriptor prefetch only).
// Kernels launched the ordinary way are unaffected (griddepcontrol.* are no-ops for them), and B200SD_PDL=0 turns the
// launch attribute off at run time.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace b200sd {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200SD_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace b200sd
