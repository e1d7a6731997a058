// Persistent tcgen05 GEMM for the training path (forward, input-gradient and weight-gradient contractions).
//
//   C[M,N] = A * B, reduction over K, bf16 operands, fp32 accumulation in TMEM.
//   One CTA per SM loops over 128 x BN output tiles (x split-K slices).  Warp 0 = TMA producer, warp 1 = single-thread
//   tcgen05.mma issuer, warps 2..9 = epilogue (TMEM -> registers -> fused epilogue -> global).  Two accumulator stages
//   in TMEM (2 x BN columns) let the epilogue of tile i overlap the main loop of tile i+1.
//
//   Both operands can be "K-major" (reduction index contiguous in memory: activations [rows, features] as A, nn.Linear
//   weights [out, in] as B) or "MN-major" (reduction index is the slow one: dY^T X weight gradients take both operands
//   that way, dX = dY W takes W that way), so no transposed copy of an activation or a weight is ever made.  MN-major
//   tiles are fetched as 64(mn) x 64(k) TMA boxes with the 128B swizzle; the UMMA descriptor walks them with
//   LBO = 8192 B (next 64-wide mn block) and SBO = 1024 B (next group of 8 k rows).
//
// Replaces, for training, the nn.Linear forward/backward calls of fastai's MultiHeadRelativeAttention / feed_forward /
// LinearDecoder under autograd.
#include <cuda.h>
#include <map>
#include <tuple>
#include "kernels.cuh"
#include "launch.cuh"
#include "train_kernels.cuh"

namespace dmg {

namespace {

constexpr int GT_BM = 128;
constexpr int GT_EPI_WARPS = 8;
constexpr int GT_THREADS = 64 + GT_EPI_WARPS * 32;

__host__ __device__ constexpr int gt_stage_bytes(int BN) { return GT_BM * 64 * 2 + BN * 64 * 2; }
__host__ __device__ constexpr int gt_stages(int BN) { return BN == 256 ? 4 : (BN == 128 ? 6 : 8); }
__host__ __device__ constexpr int gt_smem_bytes(int BN) { return gt_stages(BN) * gt_stage_bytes(BN) + 1024 + 256; }

// shared-memory matrix descriptor, 128B swizzle, sm_100 version bit; lbo/sbo in bytes
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

struct GemmTParams {
  int M, N, K;
  int a_mn, b_mn;
  int splitk, kb_per_split, num_kb;
  int num_m, num_n;
  const float* bias;
  int act;
  const void* aux; long long ld_aux; int aux_mode;
  void* out; long long ldc; int out_mode;
  bf16* out2; long long ld2;
  uint32_t drop_thresh, drop_seed; float drop_scale;
  int groups; int a_gs, b_gs; long long c_gs;
};

template <int BN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_train_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTParams p) {
  constexpr int STAGES = gt_stages(BN);
  constexpr int STAGE_BYTES = gt_stage_bytes(BN);
  constexpr int A_BYTES = GT_BM * 64 * 2;
  constexpr int TMEM_COLS = 2 * BN;   // 128 / 256 / 512: powers of two
  extern __shared__ uint8_t gt_smem_raw[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)gt_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_holder = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mn = p.num_m * p.num_n;
  const int tiles_g = tiles_mn * p.splitk;
  const int total = tiles_g * p.groups;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], GT_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int grp = t / tiles_g, tg = t - grp * tiles_g;
        const int ks = tg / tiles_mn, r = tg - ks * tiles_mn;
        const int m0 = (r % p.num_m) * GT_BM + grp * p.a_gs, n0 = (r / p.num_m) * BN + grp * p.b_gs;
        const int kb_lo = ks * p.kb_per_split;
        const int kb_hi = min(p.num_kb, kb_lo + p.kb_per_split);
        for (int kb = kb_lo; kb < kb_hi; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* a_dst = tiles + s * STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_BYTES;
          if (!p.a_mn) {
            tma_load_2d(a_dst, &tmA, kb * 64, m0, &full[s]);
          } else {
#pragma unroll
            for (int j = 0; j < GT_BM / 64; j++) tma_load_2d(a_dst + j * 8192, &tmA, m0 + j * 64, kb * 64, &full[s]);
          }
          if (!p.b_mn) {
            tma_load_2d(b_dst, &tmB, kb * 64, n0, &full[s]);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; j++) tma_load_2d(b_dst + j * 8192, &tmB, n0 + j * 64, kb * 64, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
    const uint32_t a_lbo = p.a_mn ? 8192u : 16u, a_kstep = p.a_mn ? 2048u : 32u;
    const uint32_t b_lbo = p.b_mn ? 8192u : 16u, b_kstep = p.b_mn ? 2048u : 32u;
    uint32_t it = 0, tile_i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, tile_i++) {
      const int ks = (t % tiles_g) / tiles_mn;
      const int kb_lo = ks * p.kb_per_split;
      const int kb_hi = min(p.num_kb, kb_lo + p.kb_per_split);
      const uint32_t acc = tile_i & 1, acc_ph = (tile_i >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_ph ^ 1);     // the epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb_lo; kb < kb_hi; kb++, it++) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(tiles + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_bf16(d_tmem, umma_desc(a_addr + k * a_kstep, a_lbo, 1024u), umma_desc(b_addr + k * b_kstep, b_lbo, 1024u),
                      idesc, (uint32_t)((kb > kb_lo) || k > 0));
          umma_commit(&empty[s]);
          if (kb == kb_hi - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
      if (kb_hi <= kb_lo && lane == 0) umma_commit(&tmem_full[acc]);   // empty K slice (cannot happen with our splits)
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int half = ew >> 2;                // column half of the tile
    constexpr int HALF_COLS = BN / 2;
    uint32_t tile_i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, tile_i++) {
      const int grp = t / tiles_g, tg = t - grp * tiles_g;
      const int ks = tg / tiles_mn, r = tg - ks * tiles_mn;
      const int m0 = (r % p.num_m) * GT_BM, n0 = (r / p.num_m) * BN;
      const long long c_off = (long long)grp * p.c_gs;
      const uint32_t acc = tile_i & 1, acc_ph = (tile_i >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool first_split = ks == 0;
#pragma unroll 1
      for (int c = 0; c < HALF_COLS; c += 32) {
        uint32_t rr[32];
        const int ccol = half * HALF_COLS + c;
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)ccol, rr);
        tmem_ld_wait();
        if (c + 32 >= HALF_COLS) {            // last read of this accumulator stage by this warp: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int col0 = n0 + ccol;
        if (row >= p.M || col0 >= p.N) continue;
        const int nval = min(32, p.N - col0);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(rr[j]);
        if (p.bias && first_split) {
          if (nval == 32) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float4 bq = __ldg((const float4*)(p.bias + col0) + j);
              v[4 * j] += bq.x; v[4 * j + 1] += bq.y; v[4 * j + 2] += bq.z; v[4 * j + 3] += bq.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] += (j < nval) ? __ldg(p.bias + col0 + j) : 0.f;
          }
        }
        if (p.out2) {
          bf16* o2 = p.out2 + (size_t)row * p.ld2 + col0;
          if (nval == 32) {
#pragma unroll
            for (int j = 0; j < 4; j++)
              ((uint4*)o2)[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (j < nval) o2[j] = __float2bfloat16_rn(v[j]);
          }
        }
        if (p.act) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = gelu_tanh_fast(v[j]);
        }
        if (p.aux_mode == GEMM_AUX_GELU_GRAD || p.aux_mode == GEMM_AUX_ADD_BF16) {
          const bf16* ax = (const bf16*)p.aux + (size_t)row * p.ld_aux + col0;
          float a[32];
          if (nval == 32) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const uint4 w = __ldg((const uint4*)ax + j);
              a[8 * j] = bf16lo(w.x); a[8 * j + 1] = bf16hi(w.x); a[8 * j + 2] = bf16lo(w.y); a[8 * j + 3] = bf16hi(w.y);
              a[8 * j + 4] = bf16lo(w.z); a[8 * j + 5] = bf16hi(w.z); a[8 * j + 6] = bf16lo(w.w); a[8 * j + 7] = bf16hi(w.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) a[j] = (j < nval) ? __bfloat162float(ax[j]) : 0.f;
          }
          if (p.aux_mode == GEMM_AUX_GELU_GRAD) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] *= gelu_tanh_grad_fast(a[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] += a[j];
          }
        } else if (p.aux_mode == GEMM_AUX_ADD_F32) {
          const float* ax = (const float*)p.aux + (size_t)row * p.ld_aux + col0;
          if (nval == 32) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float4 w = __ldg((const float4*)ax + j);
              v[4 * j] += w.x; v[4 * j + 1] += w.y; v[4 * j + 2] += w.z; v[4 * j + 3] += w.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] += (j < nval) ? ax[j] : 0.f;
          }
        }
        if (p.drop_thresh) {
          // element index = row * N + col: pairs (2p, 2p+1) share one hash (N is even whenever dropout is used)
          const uint32_t e0 = (uint32_t)row * (uint32_t)p.N + (uint32_t)col0;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint32_t h = drop_pair_bits(p.drop_seed, (e0 + j) >> 1);
            v[j] = ((h & 0xFFFFu) >= p.drop_thresh) ? v[j] * p.drop_scale : 0.f;
            v[j + 1] = ((h >> 16) >= p.drop_thresh) ? v[j + 1] * p.drop_scale : 0.f;
          }
        }
        if (p.out_mode == GEMM_OUT_BF16) {
          bf16* o = (bf16*)p.out + (size_t)row * p.ldc + col0 + c_off;
          if (nval == 32) {
#pragma unroll
            for (int j = 0; j < 4; j++)
              ((uint4*)o)[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                          pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (j < nval) o[j] = __float2bfloat16_rn(v[j]);
          }
        } else if (p.out_mode == GEMM_OUT_F32) {
          float* o = (float*)p.out + (size_t)row * p.ldc + col0 + c_off;
          if (nval == 32) {
#pragma unroll
            for (int j = 0; j < 8; j++) ((float4*)o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (j < nval) o[j] = v[j];
          }
        } else {
          float* o = (float*)p.out + (size_t)row * p.ldc + col0 + c_off;
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (j < nval) atomicAdd(o + j, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- tensor-map cache: the training step re-issues the same few hundred (buffer, shape) pairs every iteration
typedef std::tuple<const void*, long long, long long, long long, int> TmKey;
std::map<TmKey, TensorMap2D>& tm_cache() {
  static std::map<TmKey, TensorMap2D> c;
  return c;
}
int get_tmap(const void* base, long long inner, long long rows, long long ld, int box_rows, const TensorMap2D** out) {
  TmKey k(base, inner, rows, ld, box_rows);
  auto& c = tm_cache();
  auto it = c.find(k);
  if (it == c.end()) {
    TensorMap2D tm;
    if (make_tmap_bf16_ex(&tm, base, inner, rows, ld, box_rows)) return -1;
    if (c.size() > 8192) c.clear();
    it = c.emplace(k, tm).first;
  }
  *out = &it->second;
  return 0;
}

template <int BN>
int launch_gt(const TensorMap2D* ta, const TensorMap2D* tb, const GemmTParams& p, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    DMG_CUDA_OK(cudaFuncSetAttribute(gemm_train_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes(BN)));
    configured = true;
  }
  return launch_np(gemm_train_kernel<BN>, dim3(grid), dim3(GT_THREADS), (size_t)gt_smem_bytes(BN), st,
                   *(const CUtensorMap*)ta->bytes, *(const CUtensorMap*)tb->bytes, p);
}

}  // namespace

int gemm_bf16_tc(const bf16* A, int a_mn, long long lda, const bf16* B, int b_mn, long long ldb, int M, int N, int K,
                 int splitk, const GemmEpi& e, int num_sms, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  DMG_CHECK(K > 0, "gemm_bf16_tc: K=%d", K);
  DMG_CHECK(A && B && e.out, "gemm_bf16_tc: null operand");
  DMG_CHECK(lda % 8 == 0 && ldb % 8 == 0, "gemm_bf16_tc: operand row strides must be multiples of 8 (lda=%lld ldb=%lld)", lda, ldb);
  DMG_CHECK(splitk >= 1 && (splitk == 1 || e.out_mode == GEMM_OUT_ATOMIC), "gemm_bf16_tc: split-K needs the atomic output mode");
  DMG_CHECK(e.out_mode == GEMM_OUT_BF16 ? e.ldc % 8 == 0 : e.ldc % 4 == 0, "gemm_bf16_tc: output row stride %lld misaligned", e.ldc);
  DMG_CHECK(!e.out2 || e.ld2 % 8 == 0, "gemm_bf16_tc: second output row stride misaligned");
  DMG_CHECK(e.aux_mode == GEMM_AUX_NONE || (e.aux && (e.aux_mode == GEMM_AUX_ADD_F32 ? e.ld_aux % 4 == 0 : e.ld_aux % 8 == 0)),
            "gemm_bf16_tc: aux operand missing or misaligned");
  DMG_CHECK(!e.drop_thresh || N % 2 == 0, "gemm_bf16_tc: fused dropout needs an even N");
  // tile width: widest tile that still yields at least ~one wave of tiles
  const int num_m = (M + GT_BM - 1) / GT_BM;
  int BN = 256;
  const int groups = e.groups < 1 ? 1 : e.groups;
  while (BN > 64 && ((long long)num_m * ((N + BN - 1) / BN) * splitk * groups < num_sms || N <= BN / 2)) BN >>= 1;
  const int num_n = (N + BN - 1) / BN;
  GemmTParams p;
  p.M = M; p.N = N; p.K = K; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.num_kb = (K + 63) / 64;
  if (splitk > p.num_kb) splitk = p.num_kb;
  p.kb_per_split = (p.num_kb + splitk - 1) / splitk;
  p.splitk = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.num_m = num_m; p.num_n = num_n;
  p.bias = e.bias; p.act = e.act; p.aux = e.aux; p.ld_aux = e.ld_aux; p.aux_mode = e.aux_mode;
  p.out = e.out; p.ldc = e.ldc; p.out_mode = e.out_mode; p.out2 = e.out2; p.ld2 = e.ld2;
  p.drop_thresh = e.drop_thresh; p.drop_seed = e.drop_seed; p.drop_scale = e.drop_scale;
  p.groups = e.groups < 1 ? 1 : e.groups; p.a_gs = (int)e.a_gs; p.b_gs = (int)e.b_gs; p.c_gs = e.c_gs;
  DMG_CHECK(p.groups == 1 || (!e.out2 && e.aux_mode == GEMM_AUX_NONE && !e.bias && !e.drop_thresh), "gemm_bf16_tc: grouped launches take the plain epilogue only");
  const TensorMap2D *ta = nullptr, *tb = nullptr;
  const long long Mext = M + (long long)(groups - 1) * e.a_gs, Next = N + (long long)(groups - 1) * e.b_gs;   // tensor extents
  if (!a_mn) { if (get_tmap(A, K, Mext, lda, GT_BM, &ta)) return -1; }
  else       { if (get_tmap(A, Mext, K, lda, 64, &ta)) return -1; }
  if (!b_mn) { if (get_tmap(B, K, Next, ldb, BN, &tb)) return -1; }
  else       { if (get_tmap(B, Next, K, ldb, 64, &tb)) return -1; }
  const long long total = (long long)num_m * num_n * p.splitk * groups;
  const int grid = (int)(total < num_sms ? total : num_sms);
  if (BN == 256) return launch_gt<256>(ta, tb, p, grid, st);
  if (BN == 128) return launch_gt<128>(ta, tb, p, grid, st);
  return launch_gt<64>(ta, tb, p, grid, st);
}

}  // namespace dmg
