// The one-token step between two attention launches as ONE kernel (SURVEY.md 2.2 K10-K12 + K3 of the next layer):
//
//     x   = LayerNorm(x + out(attn))                     fastai MultiHeadRelativeAttention.forward   (after _apply_attention)
//     x   = LayerNorm(x + W2 gelu(W1 x + b1) + b2)       fastai feed_forward (SequentialEx + MergeLayer + LayerNorm)
//     qkv = Wqkv' x                                      the NEXT layer's fused q|k|v projection (deep_music_genre.py:1641-1644 loop body)
//
// The unfused path runs these as 6 dependent launches per layer (4 split-K GEMMs + 2 LayerNorms, ~6 us each: every one of them
// is bound by its launch-to-launch latency chain, not by bytes or flops - profiles/README.md).  Here a CLUSTER of 8 CTAs owns 16
// generation streams ("rows") for the whole chain, so nothing but those 16 rows ever has to be exchanged:
//
//  * operands are swapped: D[feature, row] = W[feature, k] * X[row, k]^T - the weights are the M = 64 / 128 operand of
//    tcgen05.mma (K-major as nn.Linear stores them), the 16 rows are the N = 16 operand; accumulators live in TMEM, lane = feature;
//  * every GEMM's output features are split over the 8 CTAs of the cluster, so each CTA streams 1/8 of the layer's weights
//    (768 KB at C2) through an 8-stage TMA ring - the weight stream never waits for activations and runs ahead across phases;
//  * the three exchanges per layer (out-projection -> LayerNorm, FFN-up -> FFN-down, FFN-down -> LayerNorm) go through small
//    L2-resident scratch rows and a hardware cluster barrier (release / acquire); LayerNorm is computed redundantly by all 8 CTAs
//    (16 x 512 elements), its fp32 result stays in registers as the residual of the next LayerNorm and is written as the bf16
//    B operand straight into shared memory in the canonical 128B-swizzled K-major layout;
//  * launched with programmatic dependent launch: barrier init, TMEM allocation and the first weight tiles overlap the tail of
//    the attention kernel.
//
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = tcgen05.mma issuer (one lane), warps 2-5 = TMEM epilogues + LayerNorm.
#include <cuda.h>

#include "kernels.cuh"
#include "launch.cuh"

namespace dmg {
namespace {

constexpr int DL_STAGES = 8;
constexpr int DL_A_BYTES = 128 * 128;               // weight tile: up to 128 features x 64 bf16
constexpr int DL_B_BYTES = DL_ROWS * 128;           // activation tile: 16 rows x 64 bf16
constexpr int DL_STAGE = DL_A_BYTES + DL_B_BYTES;   // 18 KB (a multiple of 1024: every tile base keeps the 128B-swizzle phase)
constexpr int DL_D = 512;                           // d_model this kernel is specialised for (LayerNorm thread mapping)
constexpr int DL_XA_BYTES = DL_ROWS * DL_D * 2;     // resident B operand: the 16 rows after a LayerNorm, 8 k-blocks of 2 KB
constexpr int DL_THREADS = 192;
constexpr int DL_SMEM = DL_STAGES * DL_STAGE + DL_XA_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
constexpr int DL_TMEM_COLS = 256;
constexpr int COL_A = 0, COL_B = 16, COL_C = 144, COL_D = 160;   // accumulator columns per phase (16 per tile)

__device__ __forceinline__ uint64_t dl_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t dl_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t dl_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// split-phase cluster barrier, every thread of the cluster takes part (non-aligned forms: the one-lane roles diverge)
__device__ __forceinline__ void dl_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void dl_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void dl_fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void dl_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg((const float4*)p); }

// LayerNorm over the 512 features of one row held by 8 consecutive lanes (64 values each); eps 1e-5, two-pass
__device__ __forceinline__ void dl_layernorm(float (&z)[64], const float* __restrict__ w, const float* __restrict__ b, int t) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; i++) s += z[i];
  s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float mean = s * (1.f / DL_D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 64; i++) { const float c = z[i] - mean; q += c * c; }
  q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
  const float rstd = rsqrtf(q * (1.f / DL_D) + 1e-5f);
#pragma unroll
  for (int kb = 0; kb < 8; kb++) {
    const int c = 64 * kb + 8 * t;
    const float4 w0 = __ldg((const float4*)(w + c)), w1 = __ldg((const float4*)(w + c + 4));
    const float4 b0 = __ldg((const float4*)(b + c)), b1 = __ldg((const float4*)(b + c + 4));
    float* v = z + 8 * kb;
    v[0] = (v[0] - mean) * rstd * w0.x + b0.x; v[1] = (v[1] - mean) * rstd * w0.y + b0.y;
    v[2] = (v[2] - mean) * rstd * w0.z + b0.z; v[3] = (v[3] - mean) * rstd * w0.w + b0.w;
    v[4] = (v[4] - mean) * rstd * w1.x + b1.x; v[5] = (v[5] - mean) * rstd * w1.y + b1.y;
    v[6] = (v[6] - mean) * rstd * w1.z + b1.z; v[7] = (v[7] - mean) * rstd * w1.w + b1.w;
  }
}
// the row as the bf16 B operand: k-block kb = columns [64 kb, +64), 16 rows x 128 B, 16-byte chunk index XOR (row & 7)
__device__ __forceinline__ void dl_store_xa(uint8_t* xa_s, const float (&z)[64], int r, int t) {
#pragma unroll
  for (int kb = 0; kb < 8; kb++) {
    const float* v = z + 8 * kb;
    *(uint4*)(xa_s + kb * 2048 + r * 128 + ((t ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

__global__ void __launch_bounds__(DL_THREADS, 1)
decode_layer_kernel(const __grid_constant__ CUtensorMap tmAttn, const __grid_constant__ CUtensorMap tmH,
                    const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmWq, const DecodeLayerArgs a) {
  extern __shared__ __align__(1024) uint8_t dl_smem[];
  uint8_t* tiles = dl_smem + ((1024u - (smem_u32(dl_smem) & 1023u)) & 1023u);
  uint8_t* xa_s = tiles + DL_STAGES * DL_STAGE;
  uint64_t* full = (uint64_t*)(xa_s + DL_XA_BYTES);
  uint64_t* empty = full + DL_STAGES;
  uint64_t* tmem_full = empty + DL_STAGES;   // [4]: one per phase, single use
  uint64_t* xa_ready = tmem_full + 4;        // [2]: B operand written (after LayerNorm 1 / LayerNorm 2)
  uint32_t* tmem_holder = (uint32_t*)(xa_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)dl_cluster_ctarank();
  const int row0 = (blockIdx.x / DL_CLUSTER) * DL_ROWS;
  const bool body = (a.mode & 1) != 0, next = (a.mode & 2) != 0;
  const int nkb_A = a.HD >> 6, nkb_d = DL_D >> 6, nkb_C = a.di >> 6;
  const int fB = a.di / DL_CLUSTER, nB = fB >> 7;                         // FFN-up features of this CTA, in 128-feature tiles
  const int fD = a.n3 / DL_CLUSTER, nD128 = fD >> 7, nD64 = (fD & 127) >> 6;   // next-layer q|k|v features of this CTA
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmAttn); tma_prefetch_desc(&tmH);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < DL_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int p = 0; p < 4; p++) mbar_init(&tmem_full[p], 1);
    mbar_init(&xa_ready[0], 128); mbar_init(&xa_ready[1], 128);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<DL_TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================================================== TMA producer
    int it = 0;
    auto stage = [&](int i) { return tiles + (i % DL_STAGES) * DL_STAGE; };
    auto wait_empty = [&](int i) { if (i >= DL_STAGES) mbar_wait(&empty[i % DL_STAGES], (uint32_t)((i / DL_STAGES) - 1) & 1u); };
    auto w64 = [&](int i, const CUtensorMap* tm, int kb, int frow, uint32_t extra) {   // one 64-feature weight box
      wait_empty(i);
      mbar_expect_tx(&full[i % DL_STAGES], 64 * 128 + extra);
      tma_load_2d(stage(i), tm, kb * 64, frow, &full[i % DL_STAGES]);
    };
    auto w128 = [&](int i, const CUtensorMap* tm, int kb, int frow) {                  // two boxes = one 128-feature tile
      wait_empty(i);
      mbar_expect_tx(&full[i % DL_STAGES], 128 * 128);
      tma_load_2d(stage(i), tm, kb * 64, frow, &full[i % DL_STAGES]);
      tma_load_2d(stage(i) + 64 * 128, tm, kb * 64, frow + 64, &full[i % DL_STAGES]);
    };
    auto act = [&](int i, const CUtensorMap* tm, int kb) { tma_load_2d(stage(i) + DL_A_BYTES, tm, kb * 64, row0, &full[i % DL_STAGES]); };
    if (body) {
      // ---- phase A: out-projection rows [64 crank, +64); the weight boxes go out before the attention kernel has finished
      if (lane == 0) {
        const int pre = nkb_A < DL_STAGES ? nkb_A : DL_STAGES;
        for (int kb = 0; kb < pre; kb++) w64(it + kb, &tmWo, kb, 64 * crank, DL_B_BYTES);
        pdl_wait();
        for (int kb = 0; kb < pre; kb++) act(it + kb, &tmAttn, kb);
        for (int kb = pre; kb < nkb_A; kb++) { w64(it + kb, &tmWo, kb, 64 * crank, DL_B_BYTES); act(it + kb, &tmAttn, kb); }
      }
      it += nkb_A;
      __syncwarp();
      dl_arrive();                                                     // #1 (nothing of ours to publish)
      // ---- phase B: FFN-up features [fB crank, +fB); B operand = the resident LayerNorm rows
      if (lane == 0)
        for (int t = 0; t < nB; t++)
          for (int kb = 0; kb < nkb_d; kb++) w128(it + t * nkb_d + kb, &tmW1, kb, fB * crank + 128 * t);
      it += nB * nkb_d;
      // ---- phase C: FFN-down rows [64 crank, +64) over the whole K = d_inner; H rows come from all CTAs of the cluster
      const int pre = nkb_C < DL_STAGES ? nkb_C : DL_STAGES;
      if (lane == 0)
        for (int kb = 0; kb < pre; kb++) w64(it + kb, &tmW2, kb, 64 * crank, DL_B_BYTES);
      __syncwarp();
      dl_wait();                                                       // #1
      dl_arrive();                                                     // #2
      dl_wait();                                                       // #2: every CTA's slice of H is in global memory
      if (lane == 0) {
        for (int kb = 0; kb < pre; kb++) act(it + kb, &tmH, kb);
        for (int kb = pre; kb < nkb_C; kb++) { w64(it + kb, &tmW2, kb, 64 * crank, DL_B_BYTES); act(it + kb, &tmH, kb); }
      }
      it += nkb_C;
      __syncwarp();
      dl_arrive();                                                     // #3
    }
    if (next && lane == 0) {
      // ---- phase D: the next layer's q|k|v features [fD crank, +fD)
      for (int t = 0; t < nD128; t++)
        for (int kb = 0; kb < nkb_d; kb++) w128(it++, &tmWq, kb, fD * crank + 128 * t);
      for (int t = 0; t < nD64; t++)
        for (int kb = 0; kb < nkb_d; kb++) w64(it++, &tmWq, kb, fD * crank + 128 * nD128 + 64 * t, 0);
    }
    __syncwarp();
    if (body) dl_wait();                                               // #3
  } else if (warp == 1) {
    // ================================================================== tcgen05.mma issuer
    int it = 0;
    const uint32_t xa_addr = smem_u32(xa_s);
    auto kblock = [&](int i, uint32_t col, int M, uint32_t b_addr, bool first) {   // b_addr == 0: the activation tile of the stage
      const int s = i % DL_STAGES;
      mbar_wait(&full[s], (uint32_t)(i / DL_STAGES) & 1u);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(tiles + s * DL_STAGE);
        const uint32_t b = b_addr ? b_addr : a_addr + DL_A_BYTES;
        const uint32_t idesc = M == 64 ? dl_idesc(64, DL_ROWS) : dl_idesc(128, DL_ROWS);
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_bf16(tmem_base + col, dl_desc_sw128(a_addr + k * 32), dl_desc_sw128(b + k * 32), idesc, (uint32_t)(!first || k != 0));
        umma_commit(&empty[s]);
      }
      __syncwarp();
    };
    if (body) {
      for (int kb = 0; kb < nkb_A; kb++) kblock(it++, COL_A, 64, 0, kb == 0);
      if (lane == 0) umma_commit(&tmem_full[0]);
      __syncwarp();
      dl_arrive();                                                     // #1
      mbar_wait(&xa_ready[0], 0);                                      // LayerNorm 1 rows are in shared memory
      tc_fence_after();
      for (int t = 0; t < nB; t++)
        for (int kb = 0; kb < nkb_d; kb++) kblock(it++, COL_B + 16 * t, 128, xa_addr + kb * 2048, kb == 0);
      if (lane == 0) umma_commit(&tmem_full[1]);
      __syncwarp();
      dl_wait();                                                       // #1
      dl_arrive();                                                     // #2
      for (int kb = 0; kb < nkb_C; kb++) kblock(it++, COL_C, 64, 0, kb == 0);
      if (lane == 0) umma_commit(&tmem_full[2]);
      __syncwarp();
      dl_wait();                                                       // #2
      dl_arrive();                                                     // #3
    }
    if (next) {
      mbar_wait(&xa_ready[1], 0);
      tc_fence_after();
      for (int t = 0; t < nD128; t++)
        for (int kb = 0; kb < nkb_d; kb++) kblock(it++, COL_D + 16 * t, 128, xa_addr + kb * 2048, kb == 0);
      for (int t = 0; t < nD64; t++)
        for (int kb = 0; kb < nkb_d; kb++) kblock(it++, COL_D + 16 * (nD128 + t), 64, xa_addr + kb * 2048, kb == 0);
      if (lane == 0) umma_commit(&tmem_full[3]);
      __syncwarp();
    }
    if (body) dl_wait();                                               // #3
  } else {
    // ================================================================== epilogues + LayerNorm (128 threads)
    const int q = warp & 3;                               // TMEM lane quarter this warp may read
    const int te = threadIdx.x - 64, r = te >> 3, t = te & 7;   // LayerNorm role: row r, 16-byte chunk t of every k-block
    const int row = row0 + r;
    const bool valid = row < a.B;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    float z[64];
    pdl_wait();                                           // the residual stream and the attention output come from the predecessors
#pragma unroll
    for (int kb = 0; kb < 8; kb++) {
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (valid) { x0 = ldcg4(a.x32 + (size_t)row * DL_D + 64 * kb + 8 * t); x1 = ldcg4(a.x32 + (size_t)row * DL_D + 64 * kb + 8 * t + 4); }
      float* v = z + 8 * kb;
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    }
    uint32_t acc[16];
    if (body) {
      // ---- out-projection slice -> scratch P[row][feature]
      mbar_wait(&tmem_full[0], 0);
      tc_fence_after();
      tmem_ld_32x16(tq + COL_A, acc);
      tmem_ld_wait();
      if (lane < 16) {                                    // M = 64: feature 16 q + lane sits in TMEM lane 32 q + lane
        const int f = 64 * crank + 16 * q + lane;
        const float bias = a.bo ? __ldg(a.bo + f) : 0.f;
#pragma unroll
        for (int rr = 0; rr < DL_ROWS; rr++) a.P[(size_t)(row0 + rr) * DL_D + f] = __uint_as_float(acc[rr]) + bias;
      }
      tc_fence_before();
      __syncwarp();
      dl_arrive();                                                     // #1: P slices published
      dl_wait();
      // ---- LayerNorm 1 (every CTA, all 512 features of its cluster's 16 rows)
#pragma unroll
      for (int kb = 0; kb < 8; kb++) {
        const float4 p0 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t), p1 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t + 4);
        float* v = z + 8 * kb;
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
      }
      dl_layernorm(z, a.ln1w, a.ln1b, t);
      dl_store_xa(xa_s, z, r, t);
      dl_fence_async_smem();
      mbar_arrive(&xa_ready[0]);
      // ---- FFN-up slice: + b1, tanh-GeLU, bf16 -> scratch H[row][feature]
      mbar_wait(&tmem_full[1], 0);
      tc_fence_after();
      for (int tt = 0; tt < nB; tt++) {
        tmem_ld_32x16(tq + COL_B + 16 * tt, acc);
        tmem_ld_wait();
        const int f = fB * crank + 128 * tt + 32 * q + lane;
        const float bias = a.b1 ? __ldg(a.b1 + f) : 0.f;
#pragma unroll
        for (int rr = 0; rr < DL_ROWS; rr++)
          a.H[(size_t)(row0 + rr) * a.di + f] = __float2bfloat16_rn(gelu_tanh(__uint_as_float(acc[rr]) + bias));
      }
      tc_fence_before();
      dl_fence_async_all();                               // H is read by the TMA engine of the peer CTAs
      __syncwarp();
      dl_arrive();                                                     // #2
      dl_wait();
      // ---- FFN-down slice -> scratch P[row][feature] (every LayerNorm-1 read of P happened before barrier #2)
      mbar_wait(&tmem_full[2], 0);
      tc_fence_after();
      tmem_ld_32x16(tq + COL_C, acc);
      tmem_ld_wait();
      if (lane < 16) {
        const int f = 64 * crank + 16 * q + lane;
        const float bias = a.b2 ? __ldg(a.b2 + f) : 0.f;
#pragma unroll
        for (int rr = 0; rr < DL_ROWS; rr++) a.P[(size_t)(row0 + rr) * DL_D + f] = __uint_as_float(acc[rr]) + bias;
      }
      tc_fence_before();
      __syncwarp();
      dl_arrive();                                                     // #3
      dl_wait();
      // ---- LayerNorm 2
#pragma unroll
      for (int kb = 0; kb < 8; kb++) {
        const float4 p0 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t), p1 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t + 4);
        float* v = z + 8 * kb;
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
      }
      dl_layernorm(z, a.ln2w, a.ln2b, t);
      if (crank == 0 && valid) {                          // the residual stream of the next layer (one writer per row)
#pragma unroll
        for (int kb = 0; kb < 8; kb++) {
          const float* v = z + 8 * kb;
          float* dst = a.x32 + (size_t)row * DL_D + 64 * kb + 8 * t;
          *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
          *(float4*)(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
          if (a.xa_out)
            *(uint4*)(a.xa_out + (size_t)row * DL_D + 64 * kb + 8 * t) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
      }
    }
    if (next) {
      dl_store_xa(xa_s, z, r, t);                         // body: LayerNorm 2 rows; first call of a step: the embedded rows
      dl_fence_async_smem();
      mbar_arrive(&xa_ready[1]);
      // ---- next layer's q|k|v slice -> fp32 [row][3 HD] (the input of the decode-attention kernel)
      mbar_wait(&tmem_full[3], 0);
      tc_fence_after();
      for (int tt = 0; tt < nD128 + nD64; tt++) {
        tmem_ld_32x16(tq + COL_D + 16 * tt, acc);
        tmem_ld_wait();
        const bool half = tt >= nD128;
        const int f = fD * crank + (half ? 128 * nD128 + 64 * (tt - nD128) + 16 * q + lane : 128 * tt + 32 * q + lane);
        if (!half || lane < 16) {
          const float bias = a.bq ? __ldg(a.bq + f) : 0.f;
#pragma unroll
          for (int rr = 0; rr < DL_ROWS; rr++)
            if (row0 + rr < a.B) a.qkv[(size_t)(row0 + rr) * a.n3 + f] = __uint_as_float(acc[rr]) + bias;
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc<DL_TMEM_COLS>(tmem_base);
}

}  // namespace

bool decode_layer_supported(int d, int HD, int di, int n3) {
  if (d != DL_D || HD % 64 || HD <= 0) return false;
  if (di % (DL_CLUSTER * 128) || di / DL_CLUSTER / 128 > 8) return false;
  if (n3 % (DL_CLUSTER * 64) || (n3 / DL_CLUSTER + 127) / 128 > 6) return false;
  return true;
}

int decode_layer(const TensorMap2D* tmAttn, const TensorMap2D* tmH, const TensorMap2D* tmWo, const TensorMap2D* tmW1,
                 const TensorMap2D* tmW2, const TensorMap2D* tmWq, const DecodeLayerArgs& a, cudaStream_t st) {
  DMG_CHECK(decode_layer_supported(a.d, a.HD, a.di, a.n3), "decode_layer: geometry d=%d HD=%d di=%d n3=%d not supported", a.d, a.HD, a.di, a.n3);
  static bool configured = false;
  if (!configured) {
    DMG_CUDA_OK(cudaFuncSetAttribute(decode_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
    configured = true;
  }
  const int clusters = (a.B + DL_ROWS - 1) / DL_ROWS;
  return launch_k(decode_layer_kernel, dim3(clusters * DL_CLUSTER), dim3(DL_THREADS), (size_t)DL_SMEM, st, DL_CLUSTER,
                  *(const CUtensorMap*)tmAttn->bytes, *(const CUtensorMap*)tmH->bytes, *(const CUtensorMap*)tmWo->bytes,
                  *(const CUtensorMap*)tmW1->bytes, *(const CUtensorMap*)tmW2->bytes, *(const CUtensorMap*)tmWq->bytes, a);
}

}  // namespace dmg
