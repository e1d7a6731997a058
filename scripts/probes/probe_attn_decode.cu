// Stand-alone timing harness of the decode attention body (attention_decode2.cuh) at the C2 geometry: B streams x 8 heads over a
// 512-slot bf16 K/V ring.  Sweeps the grid size and the ring depth and prints, per configuration, the launch time, the DRAM rate
// and where consumer warp 0 spends its cycles (waiting for the query / K tiles / V tiles / named barriers vs. everything else).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DD2_PROFILE -I deepmusicgeneration_b200/csrc \
//        -o scripts/probes/probe_attn_decode.bin scripts/probes/probe_attn_decode.cu -L deepmusicgeneration_b200 -ldmg_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../../deepmusicgeneration_b200' -lcuda
#include <cstdio>
#include <vector>
#include <algorithm>
#include "attention_decode3.cuh"
using namespace dmg;

template <int G>
__global__ void __launch_bounds__((4 * G + 1) * 32, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmR,
             const AttnDecodeArgs a, int n_stages) {
  extern __shared__ __align__(1024) uint8_t d2_smem[];
  attn_decode2_body<G>(tmK, tmV, tmR, a, n_stages, 0, (int)blockIdx.x, (int)gridDim.x, d2_smem);
}

template <int T>
__global__ void __launch_bounds__(T * 5 * 32, 1)
probe3_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmR,
              const AttnDecodeArgs a, int n_stages) {
  extern __shared__ __align__(1024) uint8_t d3_smem[];
  attn_decode3_body<T>(tmK, tmV, tmR, a, n_stages, 0, (int)blockIdx.x, (int)gridDim.x, d3_smem);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 256, H = 8, M = 512, Dcap = 576, NSETS = 3;
  const size_t ring = (size_t)B * H * M * 64;
  bf16 *k[NSETS], *v[NSETS], *rd, *out;
  float *qkv, *u, *vv;
  int* st;
  for (int s = 0; s < NSETS; s++) { CK(cudaMalloc(&k[s], ring * 2)); CK(cudaMalloc(&v[s], ring * 2)); CK(cudaMemset(k[s], 0, ring * 2)); CK(cudaMemset(v[s], 0, ring * 2)); }
  CK(cudaMalloc(&rd, (size_t)H * Dcap * 64 * 2)); CK(cudaMemset(rd, 0, (size_t)H * Dcap * 64 * 2));
  CK(cudaMalloc(&out, (size_t)B * H * 64 * 2));
  CK(cudaMalloc(&qkv, (size_t)B * 3 * H * 64 * 4)); CK(cudaMemset(qkv, 0, (size_t)B * 3 * H * 64 * 4));
  CK(cudaMalloc(&u, H * 64 * 4)); CK(cudaMalloc(&vv, H * 64 * 4)); CK(cudaMemset(u, 0, H * 64 * 4)); CK(cudaMemset(vv, 0, H * 64 * 4));
  int hst[2] = {1000, 512};
  CK(cudaMalloc(&st, 8)); CK(cudaMemcpy(st, hst, 8, cudaMemcpyHostToDevice));
  unsigned long long* prof;
  CK(cudaMalloc(&prof, 256 * 14 * 8));
#ifdef D2_PROFILE
  CK(cudaMemcpyToSymbol(d2_prof_ptr, &prof, sizeof(prof)));
#endif
  CK(cudaMemset(prof, 0, 256 * 14 * 8));
  TensorMap2D tmK[NSETS], tmV[NSETS], tmR;
  for (int s = 0; s < NSETS; s++) {
    if (make_tmap_bf16(&tmK[s], k[s], 64, (long long)B * H * M, 64, 64)) return 1;
    if (make_tmap_bf16(&tmV[s], v[s], 64, (long long)B * H * M, 64, 64)) return 1;
  }
  if (make_tmap_bf16(&tmR, rd, 64, (long long)H * Dcap, 64, 64)) return 1;
  cudaStream_t stream; CK(cudaStreamCreate(&stream));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grids[] = {148, 88, 37};
  const int stages[] = {8, 4, 2};
  printf("B=%d H=%d M=%d: %.1f MB of K/V per launch\n", B, H, M, 2.0 * ring * 2 / 1e6);
  for (int G : {2}) for (int ns : {8}) for (int grid : grids) {
    const D2Layout L = d2_layout(M, G, ns);
    if (L.total > 227 * 1024) continue;
    if (G == 4) CK(cudaFuncSetAttribute(probe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    else if (G == 2) CK(cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    else CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int reps = 12;
    float best = 1e9f, total = 0;
    for (int r = 0; r < reps + 3; r++) {
      const int s = r % NSETS;
      AttnDecodeArgs a{qkv, k[s], v[s], rd, u, vv, out, st, B, H, M, Dcap, 0.125f, 0, 1};
      CK(cudaEventRecord(e0, stream));
      if (G == 4) launch_k(probe_kernel<4>, dim3(grid), dim3(17 * 32), L.total, stream, 1, *(const CUtensorMap*)tmK[s].bytes, *(const CUtensorMap*)tmV[s].bytes, *(const CUtensorMap*)tmR.bytes, a, ns);
      else if (G == 2) launch_k(probe_kernel<2>, dim3(grid), dim3(9 * 32), L.total, stream, 1, *(const CUtensorMap*)tmK[s].bytes, *(const CUtensorMap*)tmV[s].bytes, *(const CUtensorMap*)tmR.bytes, a, ns);
      else launch_k(probe_kernel<1>, dim3(grid), dim3(5 * 32), L.total, stream, 1, *(const CUtensorMap*)tmK[s].bytes, *(const CUtensorMap*)tmV[s].bytes, *(const CUtensorMap*)tmR.bytes, a, ns);
      CK(cudaEventRecord(e1, stream));
      CK(cudaStreamSynchronize(stream));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 3) { best = std::min(best, ms); total += ms; }
    }
    std::vector<unsigned long long> hp(grid * 14);
    CK(cudaMemcpy(hp.data(), prof, grid * 14 * 8, cudaMemcpyDeviceToHost));
    double acc[14] = {0};
    for (int c = 0; c < grid; c++) for (int q = 0; q < 14; q++) acc[q] += (double)hp[c * 14 + q] / grid;
    const double us = total / reps * 1e3, items = acc[5];
    printf("G=%d stages=%d (%3d KB ring) grid=%3d: %7.2f us (best %7.2f)  %5.2f TB/s  %5.1f GB/s per SM | per item: %6.0f cycles = wait q %5.0f + K %5.0f + V %5.0f + bars %5.0f + other %5.0f\n",
           G, ns, ns * L.tile_bytes / 1024, grid, us, best * 1e3, 2.0 * ring * 2 / us / 1e6, 2.0 * ring * 2 / us / 1e3 / grid, acc[0] / items, acc[1] / items,
           acc[2] / items, acc[3] / items, acc[4] / items, (acc[0] - acc[1] - acc[2] - acc[3] - acc[4]) / items);
    printf("      sections (incl. their waits): q wait %5.0f | fragments %5.0f | K phase %5.0f | own key + bar1 %5.0f | softmax + bar2 %5.0f | V phase %5.0f | epilogue %5.0f | loop %5.0f\n",
           acc[6] / items, acc[7] / items, acc[8] / items, acc[9] / items, acc[10] / items, acc[11] / items, acc[12] / items, acc[13] / items);
  }
  // third generation with the per-section cycle marks of team 0 / warp 0, two and three teams
  for (int T : {2, 3}) for (int grid : {148, 88, 37}) {
    const int ns = d3_pick_stages(M, T);
    const D3Layout L = d3_layout(M, ns, T);
    if (T == 2) CK(cudaFuncSetAttribute(probe3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    else CK(cudaFuncSetAttribute(probe3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int per = D3_CAP * grid / H;
    float total = 0;
    const int reps = 8;
    for (int r = 0; r < reps + 2; r++) {
      const int s = r % NSETS;
      AttnDecodeArgs a{qkv, k[s], v[s], rd, u, vv, out, st, per < B ? per : B, H, M, Dcap, 0.125f, 0, 1};
      CK(cudaEventRecord(e0, stream));
      if (T == 2) launch_k(probe3_kernel<2>, dim3(grid), dim3(2 * 160), L.total, stream, 1, *(const CUtensorMap*)tmK[s].bytes, *(const CUtensorMap*)tmV[s].bytes, *(const CUtensorMap*)tmR.bytes, a, ns);
      else launch_k(probe3_kernel<3>, dim3(grid), dim3(3 * 160), L.total, stream, 1, *(const CUtensorMap*)tmK[s].bytes, *(const CUtensorMap*)tmV[s].bytes, *(const CUtensorMap*)tmR.bytes, a, ns);
      CK(cudaEventRecord(e1, stream));
      CK(cudaStreamSynchronize(stream));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 2) total += ms;
    }
    std::vector<unsigned long long> hp(grid * 14);
    CK(cudaMemcpy(hp.data(), prof, grid * 14 * 8, cudaMemcpyDeviceToHost));
    double acc[14] = {0};
    for (int c = 0; c < grid; c++) for (int q = 0; q < 14; q++) acc[q] += (double)hp[c * 14 + q] / grid;
    const double items = acc[5], nstreams = per < B ? per : B, us = total / reps * 1e3;
    printf("v3 T=%d grid=%3d streams=%3.0f stages=%d: %7.2f us  %5.2f TB/s %5.1f GB/s per SM | team 0: %4.1f items, total %6.0f cycles; prologue %5.0f; per item: q wait %5.0f | fragments+own %5.0f | K phase %5.0f (waits %5.0f) | softmax %5.0f | V phase %5.0f (waits %5.0f) | epilogue %5.0f\n",
           T, grid, nstreams, ns, us, nstreams * H * M * 256.0 / us / 1e6, nstreams * H * M * 256.0 / us / 1e3 / grid, items, acc[0], acc[13], acc[6] / items, acc[7] / items, acc[8] / items, acc[2] / items, acc[10] / items, acc[11] / items, acc[3] / items, acc[12] / items);
  }
  // the library's launchers: second generation (force_v2) against the third (teams + rel-pos table), same buffers
  for (int v2 : {1, 0}) for (int sms : {148, 88, 37}) {
    const int reps = 12;
    float best = 1e9f, total = 0;
    for (int r = 0; r < reps + 3; r++) {
      const int s = r % NSETS;
      AttnDecodeArgs a{qkv, k[s], v[s], rd, u, vv, out, st, B, H, M, Dcap, 0.125f, v2, 1};
      CK(cudaEventRecord(e0, stream));
      if (attn_decode2(&tmK[s], &tmV[s], &tmR, a, 0, sms, stream)) return 1;
      CK(cudaEventRecord(e1, stream));
      CK(cudaStreamSynchronize(stream));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 3) { best = std::min(best, ms); total += ms; }
    }
    const double us = total / reps * 1e3;
    printf("library %s, %3d CTAs: %7.2f us (best %7.2f)  %5.2f TB/s  %5.1f GB/s per SM\n", v2 ? "v2" : "v3", sms, us, best * 1e3,
           2.0 * ring * 2 / us / 1e6, 2.0 * ring * 2 / us / 1e3 / sms);
  }
  return 0;
}
