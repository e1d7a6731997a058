// Kernel launch helper: every kernel of the one-token decode chain is launched with the programmatic-dependent-launch
// attribute, so kernel N+1's prologue (barrier init, TMEM allocation, descriptor prefetch, WEIGHT tile loads) overlaps
// kernel N's tail; each kernel calls pdl_launch_dependents() first thing and pdl_wait() before it touches any
// activation / state buffer.  pdl_wait() returns once the predecessor grid has completed and flushed, and since every
// predecessor waited the same way before its own side effects, everything older is complete too.
#pragma once
#include <cstdlib>
#include <utility>
#include "common.cuh"

namespace dmg {

extern long long g_launch_count;

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("DMG_NO_PDL") ? 0 : 1;
  return on == 1;
}

template <class... KArgs, class... Args>
int launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    na++;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DMG_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
  g_launch_count++;
  return 0;
}

// Plain stream-ordered launch (training kernels: no programmatic dependent launch, so no pdl_wait() obligations).
template <class... KArgs, class... Args>
int launch_np(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  DMG_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
  g_launch_count++;
  return 0;
}

}  // namespace dmg
