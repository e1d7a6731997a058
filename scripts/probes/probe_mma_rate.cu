// Probe: how long does a back-to-back stream of small tcgen05.mma instructions take?  One CTA, operands resident in shared memory
// (K-major, 128B swizzle), `count` MMAs of M x N x 16 issued by one thread, one tcgen05.commit at the end or one per `commit_every`
// MMAs, `chains` independent accumulators used round-robin.  Prints cycles per MMA (clock64 around issue -> final mbarrier wait).
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/probes/probe_mma_rate.bin scripts/probes/probe_mma_rate.cu -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepmusicgeneration_b200/csrc/common.cuh"

using namespace dmg;

__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) probe(int M, int N, int count, int chains, int commit_every, int a_tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                       // a_tiles x 16 KB of "weights"
  uint8_t* sB = smem + a_tiles * 16384;     // 32 KB activation tile (N <= 256 rows x 64 k)
  uint64_t* bar = (uint64_t*)(sB + 32768);
  uint64_t* bar2 = bar + 1;
  uint32_t* holder = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (a_tiles * 16384 + 32768) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + i;   // arbitrary finite bf16 pairs
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    int commits = 0;
    // k-blocks of 4 MMAs, fully unrolled inner loop; a_tiles and chains are powers of two (masks, no division)
    const uint64_t dhi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
    const int nkb = count >> 2, tmask = a_tiles - 1, cmask = chains - 1, cemask = (commit_every >> 2) - 1;   // commit_every: a power of two >= 4
    for (int kb = 0; kb < nkb; kb++) {
      const uint32_t a = a0 + (uint32_t)((kb & tmask) * 16384), acc = kb >= chains ? 1u : 0u;
      const uint32_t d = tmem_base + (uint32_t)((kb & cmask) * N);
#pragma unroll
      for (int k = 0; k < 4; k++)
        umma_bf16(d, dhi | (uint64_t)(((a + k * 32) & 0x3FFFFu) >> 4), dhi | (uint64_t)(((b0 + k * 32) & 0x3FFFFu) >> 4), idesc, k ? 1u : acc);
      if (commit_every > 0 && ((kb + 1) & cemask) == 0 && kb + 1 < nkb) { umma_commit(bar2); commits++; }
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0; out[2] = commits;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int smem_bytes = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  struct Cfg { int M, N, count, chains, commit_every, a_tiles; };
  std::vector<Cfg> cfgs = {
      {128, 32, 256, 1, 0, 8},  {128, 32, 256, 4, 0, 8},  {128, 32, 256, 8, 0, 8},  {128, 32, 256, 4, 4, 8},  {128, 32, 256, 4, 8, 8},
      {128, 32, 256, 4, 16, 8}, {128, 32, 256, 4, 32, 8}, {64, 32, 256, 4, 0, 8},   {64, 32, 256, 8, 4, 8},   {128, 16, 256, 4, 0, 8},
      {128, 64, 256, 4, 0, 8},  {128, 128, 256, 2, 0, 8}, {128, 256, 256, 1, 0, 8}, {128, 256, 256, 2, 0, 8}, {128, 32, 1024, 8, 0, 8},
      {128, 32, 1024, 8, 8, 8}, {64, 16, 1024, 8, 0, 8},
  };
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; rep++) {
      probe<<<1, 128, smem_bytes>>>(c.M, c.N, c.count, c.chains, c.commit_every, c.a_tiles, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("M %3d N %3d count %4d chains %d commit_every %d a_tiles %d: issue %6lld cyc (%.1f / MMA), complete %6lld cyc (%.1f / MMA)\n", c.M, c.N,
           c.count, c.chains, c.commit_every, c.a_tiles, h[0], (double)h[0] / c.count, h[1], (double)h[1] / c.count);
  }
  return 0;
}
