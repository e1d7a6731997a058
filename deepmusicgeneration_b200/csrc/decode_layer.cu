// The one-token step between two attention launches as ONE kernel (SURVEY.md 2.2 K10-K12 + K3 of the next layer):
//
//     x   = LayerNorm(x + out(attn))                     fastai MultiHeadRelativeAttention.forward   (after _apply_attention)
//     x   = LayerNorm(x + W2 gelu(W1 x + b1) + b2)       fastai feed_forward (SequentialEx + MergeLayer + LayerNorm)
//     qkv = Wqkv' x                                      the NEXT layer's fused q|k|v projection (deep_music_genre.py:1641-1644 loop body)
//
// The unfused path runs these as 6 dependent launches per layer (4 split-K GEMMs + 2 LayerNorms, ~6 us each: every one of them
// is bound by its launch-to-launch latency chain, not by bytes or flops - profiles/README.md).  Here a CLUSTER of 8 CTAs owns 32
// generation streams ("rows") for the whole chain, so nothing but those 16 rows ever has to be exchanged:
//
//  * operands are swapped: D[feature, row] = W[feature, k] * X[row, k]^T - the weights are the M = 64 / 128 operand of
//    tcgen05.mma (K-major as nn.Linear stores them), the 32 rows are the N = 32 operand; accumulators live in TMEM, lane = feature.
//    An MMA this small is bound by the latency of its accumulate chain (measured: ~180 cycles per dependent MMA with one
//    accumulator per tile, profiles/README.md r2d), so every tile keeps 4-8 INDEPENDENT partial accumulators (one per 16-wide K
//    step of a k-block, times two alternating k-blocks for single-tile phases; tiles of a phase are interleaved) that the
//    epilogue adds up;
//  * every GEMM's output features are split over the 8 CTAs of the cluster, so each CTA streams 1/8 of the layer's weights
//    (768 KB at C2) through an 8-stage TMA ring - the weight stream never waits for activations and runs ahead across phases;
//  * the three exchanges per layer (out-projection -> LayerNorm, FFN-up -> FFN-down, FFN-down -> LayerNorm) go through small
//    L2-resident scratch rows and a hardware cluster barrier (release / acquire); LayerNorm is computed redundantly by all 8 CTAs
//    (32 x 512 elements), its fp32 result stays in registers as the residual of the next LayerNorm and is written as the bf16
//    B operand straight into shared memory in the canonical 128B-swizzled K-major layout;
//  * launched with programmatic dependent launch: barrier init, TMEM allocation and the first weight tiles overlap the tail of
//    the attention kernel.
//
// Warp roles: warp 0 = TMA producer (one lane), warps 1-4 = tcgen05.mma issuers (one lane each, round-robin over stage uses: at this tile size
// the scalar issue loop of ONE thread - barrier wait, descriptors, commit - costs more than the tensor core needs per MMA),
// the last 8 warps = TMEM epilogues + LayerNorm.
// Grid: 8 CTAs per 32 streams - 64 CTAs at the benchmark's 256 streams (at most 15 clusters of 8 are co-resident on a B200, so
// 16-row clusters would run in two waves).
#include <cuda.h>

#include "kernels.cuh"
#include "launch.cuh"

namespace dmg {
namespace {

constexpr int DL_STAGES = 8;
constexpr int DL_A_BYTES = 128 * 128;               // weight tile: up to 128 features x 64 bf16
constexpr int DL_B_BYTES = DL_ROWS * 128;           // activation tile: 32 rows x 64 bf16
constexpr int DL_STAGE = DL_A_BYTES + DL_B_BYTES;   // 20 KB (a multiple of 1024: every tile base keeps the 128B-swizzle phase)
constexpr int DL_D = 512;                           // d_model this kernel is specialised for (LayerNorm thread mapping)
constexpr int DL_XA_BYTES = DL_ROWS * DL_D * 2;     // resident B operand: the 32 rows after a LayerNorm, 8 k-blocks of 4 KB
constexpr int DL_MMAW = 2;                           // MMA-issuing warps (round-robin over stage uses; measured: 1 -> 2 warps -30 % MMA time, 4 no better)
constexpr int DL_THREADS = 32 * (1 + DL_MMAW + 8);  // producer warp + MMA-issuing warps + 8 epilogue warps
constexpr int DL_EPI0 = 32 * (1 + DL_MMAW);         // first epilogue thread
constexpr int DL_NCHAIN = 4 * DL_MMAW;                       // partial accumulators of a phase (all tiles together): 16 x 32 columns = TMEM
constexpr int DL_EPI = 256;                         // epilogue / LayerNorm threads (8 per row)
constexpr int DL_LN_BYTES = 4 * DL_D * 4;           // ln1 w, b and ln2 w, b staged in shared memory
constexpr int DL_SMEM = DL_STAGES * DL_STAGE + DL_XA_BYTES + DL_LN_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
constexpr int DL_TMEM_COLS = 256;                   // phases reuse the columns: each phase's MMAs start after the previous epilogue
constexpr int DL_CHAIN = DL_ROWS;                   // TMEM columns of one partial accumulator (N = 32 rows)

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
// bounded spin without the printf of mbar_wait (a protocol bug must trap, not hang the GPU box)
__device__ __forceinline__ void dl_spin(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity))
    if (++n > (1u << 27)) __trap();
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg((const float4*)p); }
// optional per-role timeline (DecodeLayerArgs::dbg, scripts/probe_decode_layer.py): slot = role * 16 + mark, value = %globaltimer ns
__device__ __forceinline__ void dl_mark(unsigned long long* dbg, int role, int mark) {
  if (dbg == nullptr || blockIdx.x != 0) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  dbg[role * 16 + mark] = t;
}

// tanh-GeLU with the hardware tanh (error ~5e-4, below the bf16 rounding of the value that is stored)
__device__ __forceinline__ float dl_gelu(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.7978845608028654f * (x + 0.044715f * x * x * x)));
  return 0.5f * x * (1.f + t);
}

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
    const float4 w0 = *(const float4*)(w + c), w1 = *(const float4*)(w + c + 4);      // shared memory
    const float4 b0 = *(const float4*)(b + c), b1 = *(const float4*)(b + c + 4);
    float* v = z + 8 * kb;
    v[0] = (v[0] - mean) * rstd * w0.x + b0.x; v[1] = (v[1] - mean) * rstd * w0.y + b0.y;
    v[2] = (v[2] - mean) * rstd * w0.z + b0.z; v[3] = (v[3] - mean) * rstd * w0.w + b0.w;
    v[4] = (v[4] - mean) * rstd * w1.x + b1.x; v[5] = (v[5] - mean) * rstd * w1.y + b1.y;
    v[6] = (v[6] - mean) * rstd * w1.z + b1.z; v[7] = (v[7] - mean) * rstd * w1.w + b1.w;
  }
}
// the row as the bf16 B operand: k-block kb = columns [64 kb, +64), 32 rows x 128 B, 16-byte chunk index XOR (row & 7)
__device__ __forceinline__ void dl_store_xa(uint8_t* xa_s, const float (&z)[64], int r, int t) {
#pragma unroll
  for (int kb = 0; kb < 8; kb++) {
    const float* v = z + 8 * kb;
    *(uint4*)(xa_s + kb * DL_B_BYTES + r * 128 + ((t ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// Sum of the `nacc` partial accumulators of one tile for this thread's TMEM lane and its 16 rows (columns [16 half, +16) of
// every 32-column chain)
__device__ __forceinline__ void dl_gather(uint32_t taddr, int nacc, float (&out)[16]) {
  uint32_t r[16];
  tmem_ld_32x16(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = __uint_as_float(r[i]);
  for (int c = 1; c < nacc; c++) {
    tmem_ld_32x16(taddr + c * DL_CHAIN, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] += __uint_as_float(r[i]);
  }
}

__global__ void __launch_bounds__(DL_THREADS, 1)
decode_layer_kernel(const __grid_constant__ CUtensorMap tmAttn, const __grid_constant__ CUtensorMap tmH,
                    const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmWq, const DecodeLayerArgs a) {
  extern __shared__ __align__(1024) uint8_t dl_smem[];
  uint8_t* tiles = dl_smem + ((1024u - (smem_u32(dl_smem) & 1023u)) & 1023u);
  uint8_t* xa_s = tiles + DL_STAGES * DL_STAGE;
  float* ln_s = (float*)(xa_s + DL_XA_BYTES);                // [4][512]: ln1 w, ln1 b, ln2 w, ln2 b
  uint64_t* full = (uint64_t*)(xa_s + DL_XA_BYTES + DL_LN_BYTES);
  uint64_t* empty = full + DL_STAGES;
  uint64_t* tmem_full = empty + DL_STAGES;   // [4]: one per phase, single use
  uint64_t* xa_ready = tmem_full + 4;        // [2]: B operand written (after LayerNorm 1 / LayerNorm 2)
  uint32_t* tmem_holder = (uint32_t*)(xa_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)dl_cluster_ctarank();
  const int row0 = a.row_base + (blockIdx.x / DL_CLUSTER) * DL_ROWS;
  const bool body = (a.mode & 1) != 0, next = (a.mode & 2) != 0;
  const int nkb_A = a.HD >> 6, nkb_d = DL_D >> 6, nkb_C = a.di >> 6;
  const int fB = a.di / DL_CLUSTER, nB = fB >> 7;                         // FFN-up features of this CTA, in 128-feature tiles
  const int fD = a.n3 / DL_CLUSTER, nD128 = fD >> 7, nD64 = (fD & 127) >> 6;   // next-layer q|k|v features of this CTA
  const int nD = nD128 + nD64;
  // partial accumulators per tile: one per 16-wide K step of a k-block; single-tile phases also alternate two k-blocks
  const int accB = DL_NCHAIN / nB, accD = DL_NCHAIN / (nD ? nD : 1);   // chains per tile (4); single-tile phases use all 8
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmAttn); tma_prefetch_desc(&tmH);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < DL_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }   // 4 issuing lanes commit per stage
    for (int p = 0; p < 4; p++) mbar_init(&tmem_full[p], 4 * DL_MMAW);   // every MMA-issuing lane commits
    mbar_init(&xa_ready[0], DL_EPI); mbar_init(&xa_ready[1], DL_EPI);
    mbar_fence_init();
  }
  if (warp == 1 + DL_MMAW) tmem_alloc<DL_TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================================================== TMA producer
    // Eight lanes issue in lock-step, lane j serving the stage uses i with i % 8 == j of a round of eight consecutive uses: one
    // thread's scalar issue loop (barrier wait, expect_tx, two TMA instructions) was measured to be the bottleneck of phase C.
    if (lane == 0) dl_mark(a.dbg, 0, 0);
    const bool act_lane = lane < DL_STAGES;
    auto stage = [&](int i) { return tiles + (i & (DL_STAGES - 1)) * DL_STAGE; };
    auto wait_empty = [&](int i) { if (i >= DL_STAGES) dl_spin(&empty[i & (DL_STAGES - 1)], (uint32_t)((i / DL_STAGES) - 1) & 1u); };
    // weights of stage use i: `boxes` 64-feature boxes starting at feature row `frow`, k-block kb; `extra` = bytes of an activation box
    auto wload = [&](int i, const CUtensorMap* tm, int kb, int frow, int boxes, uint32_t extra) {
      wait_empty(i);
      uint64_t* fb = &full[i & (DL_STAGES - 1)];
      mbar_expect_tx(fb, (uint32_t)(boxes * 64 * 128) + extra);
      tma_load_2d(stage(i), tm, kb * 64, frow, fb);
      if (boxes == 2) tma_load_2d(stage(i) + 64 * 128, tm, kb * 64, frow + 64, fb);
    };
    auto aload = [&](int i, const CUtensorMap* tm, int kb) { tma_load_2d(stage(i) + DL_A_BYTES, tm, kb * 64, row0, &full[i & (DL_STAGES - 1)]); };
    int it = 0;
    if (body) {
      // ---- phase A: out-projection rows [64 crank, +64); the first round's weight boxes go out before the attention kernel has finished
      for (int r = 0; r < nkb_A; r += DL_STAGES) {
        const int kb = r + lane;
        const bool on = act_lane && kb < nkb_A;
        if (on) wload(it + kb, &tmWo, kb, 64 * crank, 1, DL_B_BYTES);
        if (r == 0) pdl_wait();
        if (on) aload(it + kb, &tmAttn, kb);
        __syncwarp();
      }
      it += nkb_A;
      if (lane == 0) dl_mark(a.dbg, 0, 1);
      dl_arrive();                                                     // #1 (nothing of ours to publish)
      // ---- phase B: FFN-up features [fB crank, +fB), tiles interleaved; B operand = the resident LayerNorm rows
      for (int r = 0; r < nB * nkb_d; r += DL_STAGES) {
        const int j = r + lane;
        if (act_lane && j < nB * nkb_d) wload(it + j, &tmW1, j / nB, fB * crank + 128 * (j % nB), 2, 0);
        __syncwarp();
      }
      it += nB * nkb_d;
      if (lane == 0) dl_mark(a.dbg, 0, 2);
      // ---- phase C: FFN-down rows [64 crank, +64) over the whole K = d_inner; H rows come from all CTAs of the cluster
      for (int r = 0; r < nkb_C; r += DL_STAGES) {
        const int kb = r + lane;
        const bool on = act_lane && kb < nkb_C;
        if (on) wload(it + kb, &tmW2, kb, 64 * crank, 1, DL_B_BYTES);
        if (r == 0) {
          if (lane == 0) dl_mark(a.dbg, 0, 3);
          __syncwarp();
          dl_wait();                                                   // #1
          dl_arrive();                                                 // #2
          dl_wait();                                                   // #2: every CTA's slice of H is in global memory
          if (lane == 0) dl_mark(a.dbg, 0, 4);
        }
        if (on) aload(it + kb, &tmH, kb);
        __syncwarp();
      }
      it += nkb_C;
      if (lane == 0) dl_mark(a.dbg, 0, 5);
      dl_arrive();                                                     // #3
    }
    if (next) {
      // ---- phase D: the next layer's q|k|v features [fD crank, +fD), tiles interleaved (128-feature tiles first, then 64)
      for (int r = 0; r < nD * nkb_d; r += DL_STAGES) {
        const int j = r + lane;
        if (act_lane && j < nD * nkb_d) {
          const int kb = j / nD, t = j % nD;
          if (t < nD128) wload(it + j, &tmWq, kb, fD * crank + 128 * t, 2, 0);
          else wload(it + j, &tmWq, kb, fD * crank + 128 * nD128 + 64 * (t - nD128), 1, 0);
        }
        __syncwarp();
      }
      it += nD * nkb_d;
    }
    if (lane == 0) dl_mark(a.dbg, 0, 6);
    __syncwarp();
    if (body) dl_wait();                                               // #3
  } else if (warp <= DL_MMAW) {
    // ================================================================== tcgen05.mma issuers (warp w: stage uses i with i % DL_MMAW == w - 1)
    // ONE lane runs the whole issue loop with as few scalar instructions per MMA as possible: at N = 32 the tensor core retires an
    // MMA every 64 cycles (scripts/probes/probe_mma_rate.cu: 63 cycles for any N <= 128), so the issue thread is the bottleneck.
    int it = 0;
    const int mine = warp - 1;
    const uint32_t xa_addr = smem_u32(xa_s), tiles_addr = smem_u32(tiles);
    constexpr uint64_t DHI = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
    // k-block `kb` of a tile whose partial accumulators start at column `col`: MMA k goes to chain (kb & kbx) * 4 + k, where
    // kbx + 1 = k-blocks that rotate through the tile's chains (DL_NCHAIN / 4 / tiles of the phase): stage uses that different
    // warps issue concurrently never share an accumulator
    auto kblock = [&](int i, uint32_t col, uint32_t idesc, uint32_t b_addr, int kb, int kbx) {   // b_addr == 0: the stage's activation tile
      if ((i & (DL_MMAW - 1)) != mine) return;
      const int s = i & (DL_STAGES - 1);
      dl_spin(&full[s], (uint32_t)(i / DL_STAGES) & 1u);
      tc_fence_after();
      // lanes 0-3 issue the four 16-wide K steps of the k-block at once, each into its own accumulator chain, and each commits
      // its own MMA to the stage's `empty` barrier (tcgen05.commit tracks the issuing THREAD; the barrier counts 4 arrivals)
      const uint32_t a_addr = tiles_addr + s * DL_STAGE + lane * 32;
      const uint32_t b = (b_addr ? b_addr : tiles_addr + s * DL_STAGE + DL_A_BYTES) + lane * 32;
      const uint32_t d = tmem_base + col + (uint32_t)((((kb & kbx) << 2) + lane) * DL_CHAIN);
      umma_bf16(d, DHI | (uint64_t)((a_addr & 0x3FFFFu) >> 4), DHI | (uint64_t)((b & 0x3FFFFu) >> 4), idesc, kb > kbx ? 1u : 0u);
      umma_commit(&empty[s]);
    };
    constexpr uint32_t ID64 = dl_idesc(64, DL_ROWS), ID128 = dl_idesc(128, DL_ROWS);
    static_assert((DL_STAGES & (DL_STAGES - 1)) == 0, "stage ring must be a power of two");
    if (lane == 0 && mine == 0) dl_mark(a.dbg, 1, 0);
    if (body) {
      if (lane < 4) {
        for (int kb = 0; kb < nkb_A; kb++) kblock(it++, 0, ID64, 0, kb, DL_NCHAIN / 4 - 1);
        dl_mark(a.dbg, 1, 1);
        umma_commit(&tmem_full[0]);
      }
      __syncwarp();
      dl_arrive();                                                     // #1
      if (lane < 4) {
        dl_spin(&xa_ready[0], 0);                                      // LayerNorm 1 rows are in shared memory
        tc_fence_after();
        dl_mark(a.dbg, 1, 2);
        for (int kb = 0; kb < nkb_d; kb++)
          for (int t = 0; t < nB; t++) kblock(it++, (uint32_t)(t * accB * DL_CHAIN), ID128, xa_addr + kb * DL_B_BYTES, kb, accB / 4 - 1);
        dl_mark(a.dbg, 1, 3);
        umma_commit(&tmem_full[1]);
      }
      __syncwarp();
      dl_wait();                                                       // #1
      dl_arrive();                                                     // #2
      if (lane < 4) {
        for (int kb = 0; kb < nkb_C; kb++) kblock(it++, 0, ID64, 0, kb, DL_NCHAIN / 4 - 1);
        dl_mark(a.dbg, 1, 5);
        umma_commit(&tmem_full[2]);
      }
      __syncwarp();
      dl_wait();                                                       // #2
      dl_arrive();                                                     // #3
    }
    if (next && lane < 4) {
      dl_spin(&xa_ready[1], 0);
      tc_fence_after();
      dl_mark(a.dbg, 1, 6);
      for (int kb = 0; kb < nkb_d; kb++) {
        for (int t = 0; t < nD128; t++) kblock(it++, (uint32_t)(t * accD * DL_CHAIN), ID128, xa_addr + kb * DL_B_BYTES, kb, accD / 4 - 1);
        for (int t = 0; t < nD64; t++) kblock(it++, (uint32_t)((nD128 + t) * accD * DL_CHAIN), ID64, xa_addr + kb * DL_B_BYTES, kb, accD / 4 - 1);
      }
      dl_mark(a.dbg, 1, 7);
      umma_commit(&tmem_full[3]);
    }
    __syncwarp();
    if (body) dl_wait();                                               // #3
  } else {
    // ================================================================== epilogues + LayerNorm (256 threads)
    const int q = warp & 3;                               // TMEM lane quarter this warp may read (warp % 4)
    const int half = (warp - 1 - DL_MMAW) >> 2;                     // which 16 of the 32 rows (accumulator columns) this warp unloads
    const int te = threadIdx.x - DL_EPI0, r = te >> 3, t = te & 7;   // LayerNorm role: row r, 16-byte chunk t of every k-block
    const int row = row0 + r;
    const bool valid = row < a.B;
    const int erow0 = row0 + 16 * half;                   // first row of this warp's epilogue columns
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * half);
    float z[64];
    const bool mk = te == 0;
    if (body) {                                           // LayerNorm parameters never change: stage them before the dependency wait
      const float* src[4] = {a.ln1w, a.ln1b, a.ln2w, a.ln2b};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        *(float2*)(ln_s + i * DL_D + 2 * te) = __ldg((const float2*)(src[i] + 2 * te));
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // the 256 epilogue threads only
    }
    if (mk) dl_mark(a.dbg, 2, 0);
    pdl_wait();                                           // the residual stream and the attention output come from the predecessors
    if (mk) dl_mark(a.dbg, 2, 1);
#pragma unroll
    for (int kb = 0; kb < 8; kb++) {
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (valid) { x0 = ldcg4(a.x32 + (size_t)row * DL_D + 64 * kb + 8 * t); x1 = ldcg4(a.x32 + (size_t)row * DL_D + 64 * kb + 8 * t + 4); }
      float* v = z + 8 * kb;
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    }
    float acc[16];
    if (body) {
      // ---- out-projection slice -> scratch P[row][feature]
      mbar_wait(&tmem_full[0], 0);
      tc_fence_after();
      if (mk) dl_mark(a.dbg, 2, 2);
      dl_gather(tq, 4 * nkb_A < DL_NCHAIN ? 4 * nkb_A : DL_NCHAIN, acc);
      if (lane < 16) {                                    // M = 64: feature 16 q + lane sits in TMEM lane 32 q + lane
        const int f = 64 * crank + 16 * q + lane;
        const float bias = a.bo ? __ldg(a.bo + f) : 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; rr++) a.P[(size_t)(erow0 + rr) * DL_D + f] = acc[rr] + bias;
      }
      tc_fence_before();
      __syncwarp();
      if (mk) dl_mark(a.dbg, 2, 3);
      dl_arrive();                                                     // #1: P slices published
      dl_wait();
      if (mk) dl_mark(a.dbg, 2, 4);
      // ---- LayerNorm 1 (every CTA, all 512 features of its cluster's 32 rows)
#pragma unroll
      for (int kb = 0; kb < 8; kb++) {
        const float4 p0 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t), p1 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t + 4);
        float* v = z + 8 * kb;
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
      }
      dl_layernorm(z, ln_s, ln_s + DL_D, t);
      dl_store_xa(xa_s, z, r, t);
      dl_fence_async_smem();
      mbar_arrive(&xa_ready[0]);
      if (mk) dl_mark(a.dbg, 2, 5);
      // ---- FFN-up slice: + b1, tanh-GeLU, bf16 -> scratch H[row][feature]
      mbar_wait(&tmem_full[1], 0);
      tc_fence_after();
      if (mk) dl_mark(a.dbg, 2, 6);
      for (int tt = 0; tt < nB; tt++) {
        dl_gather(tq + tt * accB * DL_CHAIN, nkb_d * 4 < accB ? nkb_d * 4 : accB, acc);
        const int f = fB * crank + 128 * tt + 32 * q + lane;
        const float bias = a.b1 ? __ldg(a.b1 + f) : 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; rr++) a.H[(size_t)(erow0 + rr) * a.di + f] = __float2bfloat16_rn(dl_gelu(acc[rr] + bias));
      }
      tc_fence_before();
      dl_fence_async_all();                               // H is read by the TMA engine of the peer CTAs
      __syncwarp();
      if (mk) dl_mark(a.dbg, 2, 7);
      dl_arrive();                                                     // #2
      dl_wait();
      if (mk) dl_mark(a.dbg, 2, 8);
      // ---- FFN-down slice -> scratch P[row][feature] (every LayerNorm-1 read of P happened before barrier #2)
      mbar_wait(&tmem_full[2], 0);
      tc_fence_after();
      if (mk) dl_mark(a.dbg, 2, 9);
      dl_gather(tq, 4 * nkb_C < DL_NCHAIN ? 4 * nkb_C : DL_NCHAIN, acc);
      if (lane < 16) {
        const int f = 64 * crank + 16 * q + lane;
        const float bias = a.b2 ? __ldg(a.b2 + f) : 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; rr++) a.P[(size_t)(erow0 + rr) * DL_D + f] = acc[rr] + bias;
      }
      tc_fence_before();
      __syncwarp();
      if (mk) dl_mark(a.dbg, 2, 10);
      dl_arrive();                                                     // #3
      dl_wait();
      if (mk) dl_mark(a.dbg, 2, 11);
      // ---- LayerNorm 2
#pragma unroll
      for (int kb = 0; kb < 8; kb++) {
        const float4 p0 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t), p1 = ldcg4(a.P + (size_t)row * DL_D + 64 * kb + 8 * t + 4);
        float* v = z + 8 * kb;
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
      }
      dl_layernorm(z, ln_s + 2 * DL_D, ln_s + 3 * DL_D, t);
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
      if (mk) dl_mark(a.dbg, 2, 12);
      // ---- next layer's q|k|v slice -> fp32 [row][3 HD] (the input of the decode-attention kernel)
      mbar_wait(&tmem_full[3], 0);
      tc_fence_after();
      if (mk) dl_mark(a.dbg, 2, 13);
      for (int tt = 0; tt < nD; tt++) {
        dl_gather(tq + tt * accD * DL_CHAIN, accD, acc);
        const bool h64 = tt >= nD128;
        const int f = fD * crank + (h64 ? 128 * nD128 + 64 * (tt - nD128) + 16 * q + lane : 128 * tt + 32 * q + lane);
        if (!h64 || lane < 16) {
          const float bias = a.bq ? __ldg(a.bq + f) : 0.f;
#pragma unroll
          for (int rr = 0; rr < 16; rr++)
            if (erow0 + rr < a.B) a.qkv[(size_t)(erow0 + rr) * a.n3 + f] = acc[rr] + bias;
        }
      }
      tc_fence_before();
      if (mk) dl_mark(a.dbg, 2, 14);
    }
  }
  __syncthreads();
  if (warp == 1 + DL_MMAW) tmem_dealloc<DL_TMEM_COLS>(tmem_base);
}

}  // namespace

bool decode_layer_supported(int d, int HD, int di, int n3) {
  if (d != DL_D || HD % 64 || HD <= 0) return false;
  if (di % (DL_CLUSTER * 128)) return false;
  const int nB = di / DL_CLUSTER / 128, nD = (n3 % (DL_CLUSTER * 64)) ? 0 : (n3 / DL_CLUSTER + 127) / 128;
  // two tiles per multi-tile phase (d_inner 2048, 3 * HD = 1536): each tile owns 4 accumulator chains and one of the two
  // MMA-issuing warps; other geometries take the unfused launches
  if (nB != 2 || nD != 2) return false;
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
  const int clusters = (a.B - a.row_base + DL_ROWS - 1) / DL_ROWS;
  if (clusters <= 0) return 0;
  return launch_k(decode_layer_kernel, dim3(clusters * DL_CLUSTER), dim3(DL_THREADS), (size_t)DL_SMEM, st, DL_CLUSTER,
                  *(const CUtensorMap*)tmAttn->bytes, *(const CUtensorMap*)tmH->bytes, *(const CUtensorMap*)tmWo->bytes,
                  *(const CUtensorMap*)tmW1->bytes, *(const CUtensorMap*)tmW2->bytes, *(const CUtensorMap*)tmWq->bytes, a);
}

}  // namespace dmg
