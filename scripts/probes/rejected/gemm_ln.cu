// Decode-step fusion: x32 = LayerNorm(x32 + A W^T + bias) ; xa = bf16(x32)   in ONE kernel (tcgen05 + cluster).
//
// Replaces, for the one-token step (rows = streams <= 512), the pair {out-projection or FFN-down GEMM, residual_layernorm}
// of fastai's MultiHeadRelativeAttention.forward / feed_forward (ln(x + drop(out(attn))),
// LN(x + W2 gelu(W1 x + b1) + b2)).  profiles/r1b_launches_decode_step.csv: those two pairs cost 9 + 12.5 us per layer as
// four launches whose data (256 x 512 activations) never leaves L2.
//
// One 8-CTA cluster owns 128 rows; CTA r computes columns [r*d/8, (r+1)*d/8) over the full K.  The A tile (activations,
// shared by all eight) is fetched once per cluster: CTA r loads rows [16r, 16r+16) of every k-block and MULTICASTS them to
// all eight CTAs; a stage is recycled when the MMAs of all eight CTAs have retired (tcgen05.commit multicast onto every
// CTA's "empty" barrier).  LayerNorm needs whole rows: each CTA reduces its slice, the eight partial sums are exchanged
// through distributed shared memory (exact two-pass mean / variance, two cluster barriers).
#include <cuda.h>
#include "kernels.cuh"
#include "launch.cuh"

namespace dmg {

namespace {

constexpr int GL_CL = 8;   // CTAs per cluster = column slices

__device__ __forceinline__ uint32_t gl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void gl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float gl_ld_remote(const float* local, uint32_t rank) {
  uint32_t a;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void gl_tma_multicast(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void gl_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint64_t gl_desc(uint32_t smem_addr) {   // K-major, 128B swizzle (as gemm.cu)
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN>
struct GlSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int STAGE = A_BYTES + BN * 64 * 2;
  static constexpr int STAGES = (192 * 1024) / STAGE > 8 ? 8 : (192 * 1024) / STAGE;
  static constexpr int TOTAL = STAGES * STAGE + 1024 /*align*/ + 2 * 128 * 4 /*partials*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
               float* __restrict__ x32, const float* __restrict__ ln_w, const float* __restrict__ ln_b, bf16* __restrict__ xa, int M,
               int d, int K) {
  using L = GlSmem<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t gl_smem_raw[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)gl_smem_raw + 1023) & ~(uintptr_t)1023);
  float* part1 = (float*)(tiles + STAGES * L::STAGE);
  float* part2 = part1 + 128;
  uint64_t* full = (uint64_t*)(part2 + 128);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_holder = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = gl_rank();
  const int m0 = (blockIdx.x / GL_CL) * 128, n0 = (int)rank * BN;
  const int num_kb = K / 64;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GL_CL);        // one arrival per CTA of the cluster: its MMAs on this stage have retired
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  gl_cluster_sync();                      // barriers of every CTA are initialised before any remote arrival / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], L::STAGE);
        uint8_t* a_dst = tiles + s * L::STAGE;
        tma_load_2d(a_dst + L::A_BYTES, &tmW, kb * 64, n0, &full[s]);           // weights: independent of the predecessor
        if (kb == 0) pdl_wait();                                                 // activations belong to the predecessor until here
        gl_tma_multicast(a_dst + rank * 2048, &tmA, kb * 64, m0 + (int)rank * 16, &full[s], (uint16_t)0xFF);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
    for (int kb = 0; kb < num_kb; kb++) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(tiles + s * L::STAGE);
        const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_bf16(tmem_base, gl_desc(a_addr + k * 32), gl_desc(b_addr + k * 32), idesc, (uint32_t)((kb | k) != 0));
        gl_commit_multicast(&empty[s], (uint16_t)0xFF);      // every CTA's producer learns that this CTA is done with stage s
        if (kb == num_kb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  }

  __syncwarp();
  // ---- epilogue (warps 2..5): bias + residual, LayerNorm over the cluster, stores.  Warps 0/1 only join the barriers.
  float v[BN];
  float mean = 0.f, rstd = 0.f;
  const int q = warp & 3;
  const int row = m0 + q * 32 + lane;
  const bool epi = warp >= 2;
  const bool valid = epi && row < M;
  if (epi) {
    pdl_wait();                            // x32 is read and rewritten here
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 res = make_float4(0.f, 0.f, 0.f, 0.f), bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) res = *(const float4*)(x32 + (size_t)row * d + n0 + c + j);
        if (bias) bq = __ldg((const float4*)(bias + n0 + c + j));
        v[c + j] = __uint_as_float(r[j]) + res.x + bq.x;
        v[c + j + 1] = __uint_as_float(r[j + 1]) + res.y + bq.y;
        v[c + j + 2] = __uint_as_float(r[j + 2]) + res.z + bq.z;
        v[c + j + 3] = __uint_as_float(r[j + 3]) + res.w + bq.w;
        s += v[c + j] + v[c + j + 1] + v[c + j + 2] + v[c + j + 3];
      }
    }
    part1[q * 32 + lane] = s;
  }
  tc_fence_before();
  gl_cluster_sync();
  if (epi) {
    float tot = 0.f;
#pragma unroll
    for (int r2 = 0; r2 < GL_CL; r2++) tot += gl_ld_remote(part1 + q * 32 + lane, (uint32_t)r2);
    mean = tot / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < BN; j++) { const float t = v[j] - mean; sq += t * t; }
    part2[q * 32 + lane] = sq;
  }
  gl_cluster_sync();
  if (epi) {
    float tot = 0.f;
#pragma unroll
    for (int r2 = 0; r2 < GL_CL; r2++) tot += gl_ld_remote(part2 + q * 32 + lane, (uint32_t)r2);
    rstd = rsqrtf(tot / (float)d + 1e-5f);
    if (valid) {
#pragma unroll
      for (int j = 0; j < BN; j += 8) {
        const float4 w0 = __ldg((const float4*)(ln_w + n0 + j)), w1 = __ldg((const float4*)(ln_w + n0 + j + 4));
        const float4 b0 = __ldg((const float4*)(ln_b + n0 + j)), b1 = __ldg((const float4*)(ln_b + n0 + j + 4));
        float o[8];
        o[0] = (v[j] - mean) * rstd * w0.x + b0.x; o[1] = (v[j + 1] - mean) * rstd * w0.y + b0.y;
        o[2] = (v[j + 2] - mean) * rstd * w0.z + b0.z; o[3] = (v[j + 3] - mean) * rstd * w0.w + b0.w;
        o[4] = (v[j + 4] - mean) * rstd * w1.x + b1.x; o[5] = (v[j + 5] - mean) * rstd * w1.y + b1.y;
        o[6] = (v[j + 6] - mean) * rstd * w1.z + b1.z; o[7] = (v[j + 7] - mean) * rstd * w1.w + b1.w;
        float* xo = x32 + (size_t)row * d + n0 + j;
        *(float4*)xo = make_float4(o[0], o[1], o[2], o[3]);
        *(float4*)(xo + 4) = make_float4(o[4], o[5], o[6], o[7]);
        *(uint4*)(xa + (size_t)row * d + n0 + j) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
  }
  gl_cluster_sync();                      // nobody retires while a peer may still read its partials / multicast into it
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BN>
int launch_gl(const TensorMap2D* tmA16, const TensorMap2D* tmW, const float* bias, float* x32, const float* ln_w, const float* ln_b,
              bf16* xa, int M, int d, int K, cudaStream_t st) {
  using L = GlSmem<BN>;
  static bool configured = false;
  if (!configured) {
    DMG_CUDA_OK(cudaFuncSetAttribute(gemm_ln_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  return launch_k(gemm_ln_kernel<BN>, dim3(GL_CL * ((M + 127) / 128)), dim3(192), (size_t)L::TOTAL, st, GL_CL,
                  *(const CUtensorMap*)tmA16->bytes, *(const CUtensorMap*)tmW->bytes, bias, x32, ln_w, ln_b, xa, M, d, K);
}

}  // namespace

bool gemm_ln_supported(int d, int K) { return (d == 512 || d == 1024) && K % 64 == 0 && K >= 64; }

// tmA16: activation map with 16-row boxes; tmW: weight map with (d/8)-row boxes
int gemm_ln(const TensorMap2D* tmA16, const TensorMap2D* tmW, const float* bias, float* x32, const float* ln_w, const float* ln_b,
            bf16* xa, int M, int d, int K, cudaStream_t st) {
  DMG_CHECK(gemm_ln_supported(d, K), "gemm_ln: d=%d K=%d unsupported", d, K);
  if (M <= 0) return 0;
  if (d == 512) return launch_gl<64>(tmA16, tmW, bias, x32, ln_w, ln_b, xa, M, d, K, st);
  return launch_gl<128>(tmA16, tmW, bias, x32, ln_w, ln_b, xa, M, d, K, st);
}

}  // namespace dmg
