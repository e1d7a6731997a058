// Probe for the round-2 plan: tcgen05.mma with the A operand in TENSOR MEMORY (the softmax warps would
// write P with tcgen05.st instead of through shared memory).  D[128 x 128] = A[128 x 64] (bf16, TMEM) x B[128 x 64]^T (bf16, K-major
// SW128 in shared memory).  Assumption under test: A lies lane = row, two bf16 per 32-bit column (low half = even k), 8 columns per
// K = 16 step.  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/probes/probe_a_tmem.bin scripts/probes/probe_a_tmem.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../deepmusicgeneration_b200/csrc/common.cuh"

using namespace dmg;

__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {   // K-major, 128B swizzle
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_a_tmem(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const uint16_t* A, const uint16_t* B, float* out, int col_step) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;
  uint64_t* bar = (uint64_t*)(smem + 16384);
  uint32_t* holder = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < 128 * 64; idx += 128) {
    const int row = idx >> 6, col = idx & 63;
    const uint32_t off = (row >> 3) * 1024 + (row & 7) * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
    *(uint16_t*)(sB + off) = B[idx];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<256>(holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  const uint32_t t_lane = tmem_base + ((uint32_t)(32 * warp) << 16);
  {   // this thread's row of A: 64 bf16 -> 32 packed registers -> TMEM columns 128 .. 159 of its lane
    uint32_t r[32];
    const int row = 32 * warp + lane;
    for (int i = 0; i < 32; i++) r[i] = (uint32_t)A[row * 64 + 2 * i] | ((uint32_t)A[row * 64 + 2 * i + 1] << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(t_lane + 128),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // f32 D, bf16 A / B, N 128, M 128
    for (int k = 0; k < 4; k++) umma_a_tmem(tmem_base, tmem_base + 128 + col_step * k, desc_k(smem_u32(sB) + k * 32), idesc, (uint32_t)(k > 0));
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int ch = 0; ch < 4; ch++) {
    uint32_t x[32];
    tmem_ld_32x32(t_lane + 32 * ch, x);
    tmem_ld_wait();
    for (int i = 0; i < 32; i++) out[(32 * warp + lane) * 128 + 32 * ch + i] = __uint_as_float(x[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

static float b2f(uint16_t h) { uint32_t w = (uint32_t)h << 16; float f; memcpy(&f, &w, 4); return f; }
static uint16_t f2b(float f) { uint32_t w; memcpy(&w, &f, 4); return (uint16_t)((w + 0x7fff + ((w >> 16) & 1)) >> 16); }

int main() {
  std::vector<uint16_t> A(128 * 64), B(128 * 64);
  srand(11);
  for (auto& v : A) v = f2b((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = f2b((rand() % 2001 - 1000) / 1000.f);
  std::vector<float> ref(128 * 128), ref_swapped(128 * 128);
  for (int r = 0; r < 128; r++)
    for (int c = 0; c < 128; c++) {
      float s = 0.f, t = 0.f;
      for (int k = 0; k < 64; k++) { s += b2f(A[r * 64 + k]) * b2f(B[c * 64 + k]); t += b2f(A[r * 64 + (k ^ 1)]) * b2f(B[c * 64 + k]); }
      ref[r * 128 + c] = s; ref_swapped[r * 128 + c] = t;
    }
  uint16_t *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 128 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024);
  for (int col_step : {8, 16}) {   // TMEM columns per K = 16 step of A: 8 if two bf16 share a column
    cudaMemset(dO, 0xff, 128 * 128 * 4);
    probe<<<1, 128, 24 * 1024>>>(dA, dB, dO, col_step);
    cudaError_t e = cudaDeviceSynchronize();
    printf("=== A in TMEM, %d columns per K=16 step: %s\n", col_step, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> O(128 * 128);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double e0 = 0, e1 = 0;
    for (int i = 0; i < 128 * 128; i++) { e0 = fmax(e0, fabs(O[i] - ref[i])); e1 = fmax(e1, fabs(O[i] - ref_swapped[i])); }
    printf("max |err| vs reference %.4g | vs reference with the bf16 pairs of A swapped %.4g\n", e0, e1);
    for (int r : {0, 37, 127})
      printf("row %3d ref %8.4f %8.4f %8.4f | got %8.4f %8.4f %8.4f | col 127: ref %8.4f got %8.4f\n", r, ref[r * 128], ref[r * 128 + 1],
             ref[r * 128 + 2], O[r * 128], O[r * 128 + 1], O[r * 128 + 2], ref[r * 128 + 127], O[r * 128 + 127]);
  }
  return 0;
}
