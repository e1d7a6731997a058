// GEMM kernels: C[M,N] = A[M,K] * W[N,K]^T (+bias)(tanh-GeLU).
//
//  * gemm_tc_kernel  - the bf16 product path: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ->
//    tcgen05.mma (cta_group::1, M=128) with fp32 accumulators in TMEM -> tcgen05.ld epilogue.
//    Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator + epilogue, warps 3-5 epilogue.
//  * gemm_simt_kernel - fp32 parity mode (and a bf16 debugging reference): plain FFMA tiles.
//
// Replaces the eager nn.Linear calls of the reference hot path: fastai MultiHeadRelativeAttention.attention /
// .out / .r_attn, feed_forward's two Linears, LinearDecoder.decoder and
// q_wgt/k_wgt/v_wgt/r_attn of deep_music_remix.py:2034-2040.
#include <cuda.h>
#include <cstdarg>
#include "kernels.cuh"
#include "launch.cuh"

namespace dmg {

long long g_launch_count = 0;

// ---------------------------------------------------------------------------------------------
// SIMT tile GEMM (64x64x16, 256 threads, 4x4 per thread)
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W,
                                                        int ldw, const float* __restrict__ bias, void* __restrict__ C,
                                                        int ldc, int M, int N, int K, int gelu, int out_bf16) {
  __shared__ float As[16][68];
  __shared__ float Ws[16][68];
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  const int lr = t >> 2, lk = (t & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int k = k0 + lk + i;
      int ra = m0 + lr, rw = n0 + lr;
      As[lk + i][lr] = (ra < M && k < K) ? to_f32(A[(size_t)ra * lda + k]) : 0.f;
      Ws[lk + i][lr] = (rw < N && k < K) ? to_f32(W[(size_t)rw * ldw + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int r = m0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int c = n0 + tx * 4 + j;
      if (c >= N) continue;
      float v = acc[i][j] + (bias ? bias[c] : 0.f);
      if (gelu) v = gelu_tanh(v);
      if (out_bf16) ((bf16*)C)[(size_t)r * ldc + c] = __float2bfloat16_rn(v);
      else ((float*)C)[(size_t)r * ldc + c] = v;
    }
  }
}

template <class T>
int gemm_simt(const T* A, int lda, const T* W, int ldw, const float* bias, void* C, int ldc, int M, int N, int K,
              int gelu, int out_bf16, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  return launch_k(gemm_simt_kernel<T>, grid, dim3(256), 0, st, 1, A, lda, W, ldw, bias, C, ldc, M, N, K, gelu, out_bf16);
}
template int gemm_simt<float>(const float*, int, const float*, int, const float*, void*, int, int, int, int, int, int,
                              cudaStream_t);
template int gemm_simt<bf16>(const bf16*, int, const bf16*, int, const float*, void*, int, int, int, int, int, int,
                             cudaStream_t);

// ---------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int make_tmap_bf16(TensorMap2D* out, const void* base, long long cols, long long rows, long long ld, int box_rows) {
  static_assert(sizeof(CUtensorMap) <= sizeof(TensorMap2D), "CUtensorMap size");
  PFN_encodeTiled enc = get_encode();
  DMG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  DMG_CHECK(cols % 64 == 0, "tensor map: inner dimension %lld is not a multiple of 64", cols);
  DMG_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "tensor map: base/stride not 16-byte aligned");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc((CUtensorMap*)out->bytes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// General form (training GEMMs): any inner length (out-of-bounds box elements read as zero), 64-element inner box.
int make_tmap_bf16_ex(TensorMap2D* out, const void* base, long long inner, long long rows, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  DMG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  DMG_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "tensor map: base/stride not 16-byte aligned (ld=%lld)", ld);
  DMG_CHECK(box_rows >= 1 && box_rows <= 256, "tensor map: box_rows %d outside [1, 256]", box_rows);
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc((CUtensorMap*)out->bytes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner %lld rows %lld ld %lld box %d)", (int)r,
            inner, rows, ld, box_rows);
  return 0;
}

// Output-side map (TMA stores of the training GEMM epilogue): 32-column x 32-row boxes (64-byte rows), 64B swizzle so the
// row-per-lane staging writes are bank-conflict free.
int make_tmap_bf16_store(TensorMap2D* out, const void* base, long long inner, long long rows, long long ld) {
  PFN_encodeTiled enc = get_encode();
  DMG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  DMG_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "store tensor map: base/stride not 16-byte aligned (ld=%lld)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc((CUtensorMap*)out->bytes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (store map) failed with CUresult %d (inner %lld rows %lld ld %lld)", (int)r, inner,
            rows, ld);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 GEMM
// ---------------------------------------------------------------------------------------------
// K-major operand tile in shared memory written by TMA with 128B swizzle: rows of 128 bytes (64 bf16), 8-row
// groups 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN, int STAGES>
struct GemmTcSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const float* __restrict__ bias, void* __restrict__ C, int ldc, int M, int N, int K, int gelu,
               int out_bf16) {
  using L = GemmTcSmem<BN, STAGES>;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(tiles + L::TILE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_holder = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * 128;
  const int num_kb = K / 64;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();   // activations (A) and the output buffer belong to the predecessor chain until here

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], L::STAGE_BYTES);
        uint8_t* a_dst = tiles + s * L::STAGE_BYTES;
        tma_load_2d(a_dst, &tmA, kb * 64, m0, &full[s]);
        tma_load_2d(a_dst + L::A_BYTES, &tmW, kb * 64, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    // instruction descriptor: D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), K-major both, N>>3 @17, M>>4 @24
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
    for (int kb = 0; kb < num_kb; kb++) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      // descriptors in uniform control flow, issue under elect_one (gemm_train.cu: behind `if (lane == 0)` every MMA is wrapped in a
      // vector-to-uniform move loop)
      const uint32_t a_addr = smem_u32(tiles + s * L::STAGE_BYTES);
      const uint64_t da0 = umma_desc_sw128(a_addr), db0 = umma_desc_sw128(a_addr + L::A_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; k++) {   // UMMA_K = 16 bf16 = 32 bytes inside the 128B swizzle atom
          umma_bf16(tmem_base, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        }
        umma_commit(&empty[s]);                       // frees the smem stage when these MMAs retire
        if (kb == num_kb - 1) umma_commit(tmem_full);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;                 // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
      const int col0 = n0 + c;
      if (row < M && col0 < N) {
        float vals[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float x = __uint_as_float(r[j]);
          if (bias) x += (col0 + j < N) ? bias[col0 + j] : 0.f;
          if (gelu) x = gelu_tanh(x);
          vals[j] = x;
        }
        if (col0 + 32 <= N) {
          if (out_bf16) {
            uint4* dst = (uint4*)((bf16*)C + (size_t)row * ldc + col0);
#pragma unroll
            for (int j = 0; j < 4; j++)
              dst[j] = make_uint4(pack_bf16x2(vals[8 * j], vals[8 * j + 1]), pack_bf16x2(vals[8 * j + 2], vals[8 * j + 3]),
                                  pack_bf16x2(vals[8 * j + 4], vals[8 * j + 5]), pack_bf16x2(vals[8 * j + 6], vals[8 * j + 7]));
          } else {
            float4* dst = (float4*)((float*)C + (size_t)row * ldc + col0);
#pragma unroll
            for (int j = 0; j < 8; j++) dst[j] = make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (col0 + j < N) {
              if (out_bf16) ((bf16*)C)[(size_t)row * ldc + col0 + j] = __float2bfloat16_rn(vals[j]);
              else ((float*)C)[(size_t)row * ldc + col0 + j] = vals[j];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BN, int STAGES>
static int launch_tc(const TensorMap2D* tmA, const TensorMap2D* tmW, const float* bias, void* C, int ldc, int M, int N,
                     int K, int gelu, int out_bf16, cudaStream_t st) {
  using L = GemmTcSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    DMG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + 127) / 128);
  return launch_k(gemm_tc_kernel<BN, STAGES>, grid, dim3(192), L::TOTAL, st, 1, *(const CUtensorMap*)tmA->bytes,
                  *(const CUtensorMap*)tmW->bytes, bias, C, ldc, M, N, K, gelu, out_bf16);
}

int gemm_tc(const TensorMap2D* tmA, const TensorMap2D* tmW, int BN, const float* bias, void* C, int ldc, int M, int N,
            int K, int gelu, int out_bf16, cudaStream_t st) {
  DMG_CHECK(K % 64 == 0 && K >= 64, "gemm_tc: K=%d must be a positive multiple of 64", K);
  if (M <= 0 || N <= 0) return 0;
  if (out_bf16) DMG_CHECK(ldc % 8 == 0, "gemm_tc: bf16 output needs ldc %% 8 == 0 (ldc=%d)", ldc);
  else DMG_CHECK(ldc % 4 == 0, "gemm_tc: fp32 output needs ldc %% 4 == 0 (ldc=%d)", ldc);
  if (BN == 32) return launch_tc<32, 8>(tmA, tmW, bias, C, ldc, M, N, K, gelu, out_bf16, st);
  if (BN == 128) return launch_tc<128, 6>(tmA, tmW, bias, C, ldc, M, N, K, gelu, out_bf16, st);
  DMG_CHECK(false, "gemm_tc: unsupported BN=%d", BN);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Skinny-M GEMM (decode: M = number of streams <= 512): split-K over a thread-block cluster.
//
// profiles/r1a_launches_decode_step.csv: with one CTA per 128x32 output tile the four per-layer GEMMs took
// 9-18 us each - 16..64 CTAs each pulling 160-640 KB through one SM's TMA path, the A tile re-read by every
// N-tile.  Here KSPLIT CTAs of a cluster each own K/KSPLIT of the reduction for the same 128xBN tile (40-80 KB per
// CTA, 128-512 CTAs per GEMM), accumulate in TMEM, park the partial tile in their shared memory and reduce it
// through distributed shared memory: CTA r finalises rows [r*128/KSPLIT, (r+1)*128/KSPLIT) by summing the KSPLIT
// partials (ld.shared::cluster), then applies bias / GeLU and stores.  Deterministic (fixed summation order).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split-phase exit barrier without memory ordering: it only has to keep a CTA's shared memory alive until every peer has
// finished READING it (the release form costs an ERRBAR that waits for this CTA's global stores: 29 % of the stall samples
// of a decode GEMM, profiles/README.md round 1d)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_only() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

constexpr int SK_MAXKB = 8;                       // k-blocks (of 64) per CTA
// BN = tile width (32, 64 or 128 output columns).  Every CTA of a wide GEMM re-reads its 128 x K/KSPLIT slice of the
// ACTIVATIONS for each column tile (16 KB per k-block against 4 KB of weights at BN = 32): measured at C2 the four
// per-layer GEMMs cost ~5 us + 2.5 us per 512 output columns, the slope being that redundant L2 -> SM traffic.  Wider
// tiles for the wide GEMMs cut it 1.7x (BN 64) / 2.5x (BN 128) and still leave >= 96 CTAs.
__host__ __device__ constexpr int sk_stage_bytes(int BN) { return 128 * 64 * 2 + BN * 64 * 2; }
__host__ __device__ constexpr int sk_red_ld(int BN) { return BN + 4; }   // padded row stride of the partial tile (floats)
// The fp32 partial tile reuses the operand stages when it fits (it is written only after the last MMA has retired);
// 2 k-blocks at BN = 32 -> 42 KB per CTA -> five CTAs per SM.
__host__ __device__ constexpr int sk_smem_bytes(int BN, int num_kb) {
  return (num_kb * sk_stage_bytes(BN) > 128 * sk_red_ld(BN) * 4 ? num_kb * sk_stage_bytes(BN) : 128 * sk_red_ld(BN) * 4) + 1024 + 256;
}

template <int KSPLIT, int BN>
__global__ void __launch_bounds__(192)
gemm_tc_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                      const float* __restrict__ bias, void* __restrict__ C, int ldc, int M, int N, int K, int gelu,
                      int out_bf16) {
  constexpr int STAGE = sk_stage_bytes(BN), RED_LD = sk_red_ld(BN);
  extern __shared__ __align__(1024) uint8_t sk_smem[];
  uint8_t* tiles = sk_smem + ((1024u - (smem_u32(sk_smem) & 1023u)) & 1023u);
  const int num_kb = K / 64 / KSPLIT;            // k-blocks of this CTA (<= SK_MAXKB)
  float* red = (float*)tiles;                     // aliases the operand stages (free once tmem_full has fired)
  uint64_t* full = (uint64_t*)(tiles + sk_smem_bytes(BN, num_kb) - 1024 - 256);
  uint64_t* tmem_full = full + SK_MAXKB;
  uint32_t* tmem_holder = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t krank = cluster_ctarank();
  const int n0 = (blockIdx.x / KSPLIT) * BN, m0 = blockIdx.y * 128;
  const int kb0 = krank * num_kb;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < num_kb; s++) mbar_init(&full[s], 1);
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<BN>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {   // every stage is used once: issue all loads up front - the WEIGHT tiles even before the
                       // predecessor kernel has finished (they never change), the activation tiles after pdl_wait
      for (int kb = 0; kb < num_kb; kb++) {
        mbar_expect_tx(&full[kb], STAGE);
#pragma unroll
        for (int j = 0; j < BN / 32; j++)          // the weight map has 32-row boxes; consecutive boxes form one K-major tile
          tma_load_2d(tiles + kb * STAGE + 128 * 64 * 2 + j * 32 * 128, &tmW, (kb0 + kb) * 64, n0 + 32 * j, &full[kb]);
      }
      pdl_wait();
      for (int kb = 0; kb < num_kb; kb++) tma_load_2d(tiles + kb * STAGE, &tmA, (kb0 + kb) * 64, m0, &full[kb]);
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
    for (int kb = 0; kb < num_kb; kb++) {
      mbar_wait(&full[kb], 0);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(tiles + kb * STAGE);
      const uint64_t da0 = umma_desc_sw128(a_addr), db0 = umma_desc_sw128(a_addr + 128 * 64 * 2);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_bf16(tmem_base, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        if (kb == num_kb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // partial tile: TMEM -> registers -> this CTA's shared memory (padded rows: conflict-free 16-byte stores)
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
      float4* dst = (float4*)(red + (q * 32 + lane) * RED_LD + c);
#pragma unroll
      for (int j = 0; j < 8; j++)
        dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
    }
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();          // all partial tiles of the cluster are in shared memory
  pdl_wait();                  // the output buffer may still be read by the predecessor chain until here

  if (warp >= 2) {
    constexpr int ROWS = 128 / KSPLIT;          // rows finalised by this CTA
    constexpr int TPR = 128 / ROWS;             // threads per row (4 or 8)
    constexpr int CPT = BN / TPR;               // columns per thread (a multiple of 4)
    const int te = threadIdx.x - 64;            // 0..127
    const int rr = te / TPR, cg = (te % TPR) * CPT;
    const int row_l = krank * ROWS + rr;
    const int row = m0 + row_l;
    float v[CPT];
#pragma unroll
    for (int c4 = 0; c4 < CPT; c4 += 4) {            // all remote reads first ...
      const uint32_t laddr = smem_u32(red + row_l * RED_LD + cg + c4);
      v[c4] = 0.f; v[c4 + 1] = 0.f; v[c4 + 2] = 0.f; v[c4 + 3] = 0.f;
#pragma unroll
      for (int r2 = 0; r2 < KSPLIT; r2++) {
        const float4 x = ld_dsmem_f4(mapa_shared(laddr, (uint32_t)r2));
        v[c4] += x.x; v[c4 + 1] += x.y; v[c4 + 2] += x.z; v[c4 + 3] += x.w;
      }
    }
    // the sums must be complete (= every remote load has returned) before the arrival is issued: pin them in front of it
#pragma unroll
    for (int c = 0; c < CPT; c++) asm volatile("" ::"f"(v[c]) : "memory");
    __syncwarp();
    cluster_arrive_relaxed();                        // ... then this warp is done with the peers' shared memory
#pragma unroll
    for (int c4 = 0; c4 < CPT; c4 += 4) {
      const int col0 = n0 + cg + c4;
      if (row < M && col0 < N) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (bias && col0 + j < N) v[c4 + j] += bias[col0 + j];
          if (gelu) v[c4 + j] = gelu_tanh(v[c4 + j]);
        }
        if (col0 + 4 <= N) {
          if (out_bf16) *(uint2*)((bf16*)C + (size_t)row * ldc + col0) = make_uint2(pack_bf16x2(v[c4], v[c4 + 1]), pack_bf16x2(v[c4 + 2], v[c4 + 3]));
          else *(float4*)((float*)C + (size_t)row * ldc + col0) = make_float4(v[c4], v[c4 + 1], v[c4 + 2], v[c4 + 3]);
        } else {
          for (int j = 0; j < 4; j++)
            if (col0 + j < N) {
              if (out_bf16) ((bf16*)C)[(size_t)row * ldc + col0 + j] = __float2bfloat16_rn(v[c4 + j]);
              else ((float*)C)[(size_t)row * ldc + col0 + j] = v[c4 + j];
            }
        }
      }
    }
  } else {
    __syncwarp();
    cluster_arrive_relaxed();
  }
  cluster_wait_only();         // nobody may exit while a peer still reads its shared memory
  if (warp == 2) tmem_dealloc<BN>(tmem_base);
}

template <int KSPLIT, int BN>
static int launch_splitk(const TensorMap2D* tmA, const TensorMap2D* tmW, const float* bias, void* C, int ldc, int M, int N,
                         int K, int gelu, int out_bf16, cudaStream_t st) {
  const int num_kb = K / 64 / KSPLIT;
  const int smem = sk_smem_bytes(BN, num_kb);
  static int configured = 0;
  if (configured < smem) {
    DMG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_splitk_kernel<KSPLIT, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  return launch_k(gemm_tc_splitk_kernel<KSPLIT, BN>, dim3(((N + BN - 1) / BN) * KSPLIT, (M + 127) / 128, 1), dim3(192), smem, st,
                  KSPLIT, *(const CUtensorMap*)tmA->bytes, *(const CUtensorMap*)tmW->bytes, bias, C, ldc, M, N, K, gelu, out_bf16);
}

// Picks the split: K/64 k-blocks spread over 4 or 8 CTAs (at most SK_MAXKB blocks each).  Returns 0 when the shape
// does not qualify (caller falls back to gemm_tc).
int gemm_tc_splitk_ways(int K) {
  if (K % 64) return 0;
  const int kb = K / 64;
  static const int min8 = getenv("DMG_SPLITK_8") ? 1 : 2;   // DMG_SPLITK_8: 8-way split already at K = 512 (timing experiments)
  if (kb % 8 == 0 && kb / 8 >= min8 && kb / 8 <= SK_MAXKB) return 8;
  if (kb % 4 == 0 && kb / 4 >= 1 && kb / 4 <= SK_MAXKB) return 4;
  return 0;
}

int gemm_tc_splitk(const TensorMap2D* tmA, const TensorMap2D* tmW32, const float* bias, void* C, int ldc, int M, int N,
                   int K, int gelu, int out_bf16, cudaStream_t st) {
  const int ways = gemm_tc_splitk_ways(K);
  DMG_CHECK(ways != 0, "gemm_tc_splitk: K=%d does not split", K);
  if (M <= 0 || N <= 0) return 0;
  if (out_bf16) DMG_CHECK(ldc % 8 == 0, "gemm_tc_splitk: bf16 output needs ldc %% 8 == 0 (ldc=%d)", ldc);
  else DMG_CHECK(ldc % 4 == 0, "gemm_tc_splitk: fp32 output needs ldc %% 4 == 0 (ldc=%d)", ldc);
  // tile width: measured at C2 (256 rows; ms per decode step): BN 32 -> 1.401, BN 64 -> 1.474, BN 128 -> 1.790 - the
  // GEMMs of the one-token step are bound by the per-CTA latency chain (TMA -> MMA -> TMEM -> DSMEM reduction), not by L2
  // traffic, so many small CTAs (five resident per SM) win.  DMG_SPLITK_BN overrides (timing experiments).
  static const int force_bn = getenv("DMG_SPLITK_BN") ? atoi(getenv("DMG_SPLITK_BN")) : 0;
  int BN = 32;
  if (force_bn == 64 || force_bn == 128) BN = (N % force_bn == 0) ? force_bn : 32;
#define DMG_SK(W_, B_) launch_splitk<W_, B_>(tmA, tmW32, bias, C, ldc, M, N, K, gelu, out_bf16, st)
  if (ways == 8) return BN == 128 ? DMG_SK(8, 128) : BN == 64 ? DMG_SK(8, 64) : DMG_SK(8, 32);
  return BN == 128 ? DMG_SK(4, 128) : BN == 64 ? DMG_SK(4, 64) : DMG_SK(4, 32);
#undef DMG_SK
}

}  // namespace dmg
