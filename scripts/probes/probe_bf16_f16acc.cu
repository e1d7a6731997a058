// Probe for the round-2 plan: tcgen05.mma kind::f16 with BF16 operands and FP16 accumulators - is the
// combination accepted, and are the results right?  (If so the Rd cache and the q + v tile can stay bf16.)
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/probes/probe_bf16_f16acc.bin scripts/probes/probe_bf16_f16acc.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
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

__global__ void __launch_bounds__(128, 1) probe(const uint16_t* A, const uint16_t* B, uint32_t* out_plain, uint32_t* out_pack, int d_f32) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem; uint8_t* sB = smem + 16384;
  uint64_t* bar = (uint64_t*)(smem + 32768);
  uint32_t* holder = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < 128 * 64; idx += 128) {
    const int row = idx >> 6, col = idx & 63;
    const uint32_t off = (row >> 3) * 1024 + (row & 7) * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
    *(uint16_t*)(sA + off) = A[idx];
    *(uint16_t*)(sB + off) = B[idx];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<128>(holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  if (tid == 0) {
    // D format bits [4,6): 0 = f16, 1 = f32; A / B format bits [7,10) / [10,13): 0 = f16; N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = ((uint32_t)(d_f32 ? 1 : 0) << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // bf16 A / B
    for (int k = 0; k < 4; k++) umma_bf16(tmem_base, desc_k(smem_u32(sA) + k * 32), desc_k(smem_u32(sB) + k * 32), idesc, (uint32_t)(k > 0));
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t t_lane = tmem_base + ((uint32_t)(32 * warp) << 16);
  for (int ch = 0; ch < 4; ch++) {
    uint32_t x[32];
    tmem_ld_32x32(t_lane + 32 * ch, x);
    tmem_ld_wait();
    for (int i = 0; i < 32; i++) out_plain[(32 * warp + lane) * 128 + 32 * ch + i] = x[i];
  }
  for (int ch = 0; ch < 2; ch++) {   // .pack::16b: 32 registers <- 64 columns (two adjacent columns' low halves per register)?
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(t_lane + 64 * ch)
        : "memory");
    tmem_ld_wait();
    for (int i = 0; i < 32; i++) out_pack[(32 * warp + lane) * 64 + 32 * ch + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

static float h2f(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }
static float b2f(uint16_t h) { uint32_t w = (uint32_t)h << 16; float f; memcpy(&f, &w, 4); return f; }
static uint16_t f2b(float f) { uint32_t w; memcpy(&w, &f, 4); return (uint16_t)((w + 0x7fff + ((w >> 16) & 1)) >> 16); }

int main() {
  std::vector<uint16_t> A(128 * 64), B(128 * 64);
  std::vector<float> ref(128 * 128);
  srand(7);
  for (auto& v : A) v = f2b((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = f2b((rand() % 2001 - 1000) / 1000.f);
  for (int r = 0; r < 128; r++)
    for (int c = 0; c < 128; c++) {
      float s = 0.f;
      for (int k = 0; k < 64; k++) s += b2f(A[r * 64 + k]) * b2f(B[c * 64 + k]);
      ref[r * 128 + c] = s;
    }
  uint16_t *dA, *dB; uint32_t *dP, *dQ;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dP, 128 * 128 * 4); cudaMalloc(&dQ, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  for (int d_f32 = 1; d_f32 >= 0; d_f32--) {
    cudaMemset(dP, 0xff, 128 * 128 * 4); cudaMemset(dQ, 0xff, 128 * 64 * 4);
    probe<<<1, 128, 40 * 1024>>>(dA, dB, dP, dQ, d_f32);
    cudaError_t e = cudaDeviceSynchronize();
    printf("=== D %s: %s\n", d_f32 ? "f32" : "f16", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint32_t> P(128 * 128), Q(128 * 64);
    cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(Q.data(), dQ, Q.size() * 4, cudaMemcpyDeviceToHost);
    double e_f32 = 0, e_lo = 0, e_packed_cells = 0, e_pack_ld = 0;
    for (int r = 0; r < 128; r++)
      for (int c = 0; c < 128; c++) {
        const float rf = ref[r * 128 + c];
        float f; uint32_t w = P[r * 128 + c]; memcpy(&f, &w, 4);
        e_f32 = fmax(e_f32, fabs(f - rf));                                              // one fp32 per cell
        e_lo = fmax(e_lo, fabs(h2f((uint16_t)(w & 0xffff)) - rf));                      // one fp16 per cell, low half
        const uint32_t wp = P[r * 128 + c / 2];
        e_packed_cells = fmax(e_packed_cells, fabs(h2f((uint16_t)((c & 1) ? wp >> 16 : wp & 0xffff)) - rf));   // two fp16 per cell: N/2 columns
        const uint32_t wq = Q[r * 64 + c / 2];
        e_pack_ld = fmax(e_pack_ld, fabs(h2f((uint16_t)((c & 1) ? wq >> 16 : wq & 0xffff)) - rf));            // .pack::16b: reg i = columns 2i, 2i+1
      }
    printf("max |err| if cells hold: fp32 %.4g | fp16 in low half %.4g | two fp16 per cell (N/2 columns) %.4g | pack::16b regs (cols 2i,2i+1) %.4g\n",
           e_f32, e_lo, e_packed_cells, e_pack_ld);
    for (int r : {0, 37, 127}) {
      printf("row %3d ref %8.4f %8.4f %8.4f %8.4f | plain words %08x %08x %08x %08x .. col64 %08x col127 %08x | pack words %08x %08x\n", r,
             ref[r * 128], ref[r * 128 + 1], ref[r * 128 + 2], ref[r * 128 + 3], P[r * 128], P[r * 128 + 1], P[r * 128 + 2], P[r * 128 + 3],
             P[r * 128 + 64], P[r * 128 + 127], Q[r * 64], Q[r * 64 + 1]);
    }
  }
  return 0;
}
