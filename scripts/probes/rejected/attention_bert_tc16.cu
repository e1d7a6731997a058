// attn_bert_tc_kernel (attention_bert_tc.cu) with SIXTEEN softmax warps - four per scheduler instead of two - for latency hiding.
//
// Same tiles, TMEM layout (AC 128 | strip 256 | O 2 x 64 columns), barriers, TMA producer and MMA issuer as the eight-warp kernel;
// replaces MemMultiHeadRelativeAttentionKV._apply_attention (deep_music_remix.py:2078-2104) for bf16 sequences of >= 128 tokens.
// The softmax work of a (query row, key half) is shared by a PAIR of threads (same lane, warps w and w + 4): each takes 32 of the
// half's 64 keys.  Their two 64-column strip windows overlap in 32 columns, so the pair reads 96 distinct strip columns ONCE (each
// thread 48: one tcgen05.ld.x32 + one .x16 instead of two .x32), converts them to fp16 and stores them into one shared 96-half line
// (the skew is then a 2-byte read at line[32 (1 - sub) + 32 + lane - jj]).  The pair agrees on the running maximum through shared
// memory (64-thread named barrier), writes its 2 x 32 probabilities into the same swizzled P row, and splits the fold of the
// P V partial by d_head halves (32 accumulator registers per thread instead of 64).  Registers: 104 per softmax thread (640 threads x 96 at launch; setmaxnreg only moves registers inside the CTA).
// fp16 strip: |BD| < 65504 saturates, 11 bits of mantissa against the 8 of the bf16 operands.
#include "attention_bert_tc_common.cuh"

namespace dmg {

namespace {

using namespace bert_tc;

constexpr int B6_SOFT_WARPS = 16;
constexpr int B6_THREADS = (B6_SOFT_WARPS + 4) * 32;   // + one auxiliary warpgroup: TMA warp, MMA warp, two idle warps
constexpr int B6_MROW = 76;                            // floats per row of the final merge buffer: 2 x (32 + m + l + pad) + 4
constexpr int C_XCH = BO_STRIP + 256 * BT_LINE16;      // after the 256 pair lines: 256 pairs x 2 floats (local maxima)
constexpr int C_BAR = C_XCH + 256 * 2 * 4;
constexpr int B6_SMEM = C_BAR + 256 + 1024;
static_assert(B6_SMEM <= 227 * 1024, "shared memory budget");
static_assert(128 * B6_MROW * 4 <= 3 * BT16K, "merge buffer aliases the three q tiles");

__device__ __forceinline__ void b6_pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(B6_THREADS, 1)
attn_bert_tc16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR, const BertTcArgs a) {
  extern __shared__ __align__(1024) uint8_t b6_smem_raw[];
  uint8_t* smem = b6_smem_raw + ((1024u - (smem_u32(b6_smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = (uint64_t*)(smem + C_BAR);
  uint32_t* tmem_holder = (uint32_t*)(bar + Q_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nT = (a.T + 127) / 128;                  // a ragged last tile: keys >= T are masked, rows >= T are not stored
  const int it = blockIdx.x % nT;
  const int bh = blockIdx.x / nT, b = bh / a.H, h = bh % a.H;
  const int i0 = it * 128, HD = a.H * 64;
  const int NT = nT;

  if (warp == B6_SOFT_WARPS && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmR);
    bt_init_barriers(bar, B6_SOFT_WARPS);
  }
  if (warp == B6_SOFT_WARPS + 1) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_launch_dependents();

  if (warp >= B6_SOFT_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == B6_SOFT_WARPS) {
      // =========================================== TMA producer ===========================================
      if (lane == 0) bt_producer(smem, bar, tmX, tmR, a, b, h, it, NT);
    } else if (warp == B6_SOFT_WARPS + 1) {
      // =========================================== MMA issuer ===========================================
      if (lane == 0) bt_mma_issuer(smem, bar, tmem_base, it, NT);
    }
  } else {
    // =========================================== softmax warps ===========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");   // 640 x 96 at launch = 512 x 104 + 128 x 56 + 1024 (the pool is what the CTA holds)
    const int q4 = warp & 3, kq = warp >> 2, hf = kq >> 1, sub = kq & 1;   // TMEM lane quarter; key quarter = (half, 32-key part)
    const int r = q4 * 32 + lane;                   // query row of this thread inside the tile
    const int row = i0 + r;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int pair = (hf * 4 + q4) * 32 + lane;
    const int pair_bar = 1 + hf * 4 + q4;           // named barriers 1..8: the two warps of a pair
    uint8_t* line = smem + BO_STRIP + (size_t)pair * BT_LINE16;
    float* xch = (float*)(smem + C_XCH) + pair * 2;
    const float c = a.scale * BT_LOG2E;

    {   // q + u, q + v and q_next + v in the canonical swizzled layout: eight consecutive threads take the eight 16-byte chunks of a row
      const int tid = threadIdx.x, pc = tid & 7, rb = tid >> 3;       // rb 0..63
      const int col = 8 * (pc ^ (rb & 7));
      const float4 ua = __ldg((const float4*)(a.u + h * 64 + col)), ub = __ldg((const float4*)(a.u + h * 64 + col + 4));
      const float4 va = __ldg((const float4*)(a.v + h * 64 + col)), vb = __ldg((const float4*)(a.v + h * 64 + col + 4));
      const float uu[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w}, vv8[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
      mbar_wait(&bar[Q_QFULL], 0);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int qr = rb + 64 * k;
        const uint32_t off = (uint32_t)((qr >> 3) * 1024 + (qr & 7) * 128 + pc * 16);
        const uint4 raw = *(const uint4*)(smem + BO_P + off);
        const uint4 rawn = *(const uint4*)(smem + BO_P + BT16K + off);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w}, wn[4] = {rawn.x, rawn.y, rawn.z, rawn.w};
        uint32_t ou[4], ov[4], on[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float u0 = uu[2 * e], u1 = uu[2 * e + 1], v0 = vv8[2 * e], v1 = vv8[2 * e + 1];
          ou[e] = pack_bf16x2(bf16lo(w[e]) + u0, bf16hi(w[e]) + u1);
          ov[e] = pack_bf16x2(bf16lo(w[e]) + v0, bf16hi(w[e]) + v1);
          on[e] = pack_bf16x2(bf16lo(wn[e]) + v0, bf16hi(wn[e]) + v1);
        }
        *(uint4*)(smem + BO_QU + off) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        *(uint4*)(smem + BO_QV + off) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
        *(uint4*)(smem + BO_QVN + off) = make_uint4(on[0], on[1], on[2], on[3]);
      }
      bt_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[Q_QREADY]);
    }

    float o[32];                                     // running output, d_head columns [32 sub, 32 sub + 32) of key half hf
#pragma unroll
    for (int i = 0; i < 32; i++) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    uint8_t* prow = smem + BO_P + hf * BT16K + (r >> 3) * 1024 + (r & 7) * 128;
    // this thread's keys inside a tile: kbase + jj, jj = 0..31
    const int kbase = 64 * hf + 32 * sub;
    // the zero pad of _line_shift: key j = i + 1 (tile (row+1)/128, local key (row+1)%128)
    const int zero_tile = (row + 1) >> 7, zero_jj = ((((row + 1) & 127) >> 5) == 2 * hf + sub) ? ((row + 1) & 31) : -1000;
    // strip window of the pair: TMEM strip columns [wbase, wbase + 96); this thread loads [wbase + 48 sub', +48) with sub' = 1 - sub
    // (sub 0 = the first 32 keys of the half = the HIGHER columns) and reads halves line[32 (1 - sub) + 32 + lane - jj]
    const uint32_t wbase = (uint32_t)(64 - 64 * hf + 32 * q4);
    const uint32_t ld_off = 48u * (uint32_t)(1 - sub);
    const int S_h = 32 * (1 - sub) + 32 + lane;                 // half index of this thread's first key (jj = 0) in the line
    const __half* rd = (const __half*)line + S_h;
    const uint32_t* rdw = (const uint32_t*)line + (S_h >> 1);   // word of halves S' - 1, S' with S' = S_h | 1
    const uint32_t rd_shift = (S_h & 1) ? 0u : 16u;

    for (int n = 0; n < NT; n++) {
      float s[32];
      mbar_wait(&bar[Q_SFULL], n & 1);
      tc_fence_after();
      {   // 48 strip columns -> fp16 -> the pair's line
        uint32_t x0[32], x1[16];
        tmem_ld_32x32(t_lane + BTM_STRIP + wbase + ld_off, x0);
        tmem_ld_32x16(t_lane + BTM_STRIP + wbase + ld_off + 32, x1);
        tmem_ld_wait();
        uint8_t* dst = line + 2 * ld_off;
#pragma unroll
        for (int k = 0; k < 4; k++)
          *(uint4*)(dst + 16 * k) = make_uint4(bt_f16x2_sat(__uint_as_float(x0[8 * k]), __uint_as_float(x0[8 * k + 1])),
                                               bt_f16x2_sat(__uint_as_float(x0[8 * k + 2]), __uint_as_float(x0[8 * k + 3])),
                                               bt_f16x2_sat(__uint_as_float(x0[8 * k + 4]), __uint_as_float(x0[8 * k + 5])),
                                               bt_f16x2_sat(__uint_as_float(x0[8 * k + 6]), __uint_as_float(x0[8 * k + 7])));
#pragma unroll
        for (int k = 0; k < 2; k++)
          *(uint4*)(dst + 64 + 16 * k) = make_uint4(bt_f16x2_sat(__uint_as_float(x1[8 * k]), __uint_as_float(x1[8 * k + 1])),
                                                    bt_f16x2_sat(__uint_as_float(x1[8 * k + 2]), __uint_as_float(x1[8 * k + 3])),
                                                    bt_f16x2_sat(__uint_as_float(x1[8 * k + 4]), __uint_as_float(x1[8 * k + 5])),
                                                    bt_f16x2_sat(__uint_as_float(x1[8 * k + 6]), __uint_as_float(x1[8 * k + 7])));
      }
      {   // content term of this thread's 32 keys
        uint32_t x[32];
        tmem_ld_32x32(t_lane + BTM_AC + kbase, x);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[Q_SFREE]);
#pragma unroll
        for (int i = 0; i < 32; i++) s[i] = __uint_as_float(x[i]);
      }
      b6_pair_sync(pair_bar);                        // both halves of the line are written
      const int zj = n == zero_tile ? zero_jj : -1000;
      if ((unsigned)zj < 32u) const_cast<__half*>(rd)[-zj] = __float2half(0.f);   // the zero pad: this thread's own slot of the line
      asm volatile("" ::: "memory");            // the word reads below alias the half store above
      {   // the 32 halves line[S - jj] as 17 aligned words (a quarter of the shared-memory wavefronts of 2-byte reads): word k holds
          // halves S' - 2k - 1 (low) and S' - 2k (high), S' = S | 1; lanes with an even S shift the word chain by one half
        uint32_t w[17];
#pragma unroll
        for (int k = 0; k < 17; k++) w[k] = rdw[-k];
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint32_t x = __funnelshift_l(w[k + 1], w[k], rd_shift);      // (w[k] << shift) | (w[k+1] >> (32 - shift)), shift 0 or 16
          const __half2 h2 = *reinterpret_cast<const __half2*>(&x);
          s[2 * k] += __high2float(h2);
          s[2 * k + 1] += __low2float(h2);
        }
      }
      if (n == NT - 1 && (a.T & 127)) {              // ragged last tile: keys past the sequence end
        const int jlim = a.T - n * 128 - kbase;
#pragma unroll
        for (int jj = 0; jj < 32; jj++) s[jj] = jj < jlim ? s[jj] : -INFINITY;
      }

      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 32; jj++) mx = fmaxf(mx, s[jj]);
      xch[sub] = mx;

      if (n > 0) {                                   // fold the previous tile's P V into the running output (this thread's d_head half)
        mbar_wait(&bar[Q_OFULL0 + hf], (n - 1) & 1);
        tc_fence_after();
        uint32_t x[32];
        tmem_ld_32x32(t_lane + BTM_O + 64 * hf + 32 * sub, x);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(x[i]));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[Q_OFREE0 + hf]);
      }

      b6_pair_sync(pair_bar);                        // local maxima exchanged; the partner is done with the strip line
      const float m_new = fmaxf(m_run, fmaxf(mx, xch[sub ^ 1]));
      const float alpha = ex2_fast((m_run - m_new) * c);        // exp2(-inf) = 0 on the first tile
      const float neg_mc = -m_new * c;
      m_run = m_new;
      float rs = 0.f;
#pragma unroll
      for (int ck = 0; ck < 4; ck++) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int pp = 4 * ck + e;
          const float p0 = ex2_fast(fmaf(s[2 * pp], c, neg_mc)), p1 = ex2_fast(fmaf(s[2 * pp + 1], c, neg_mc));
          rs += p0 + p1;
          pk[e] = pack_bf16x2(p0, p1);
        }
        *(uint4*)(prow + (((4 * sub + ck) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_run = l_run * alpha + rs;
      alpha_prev = alpha;
      bt_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[Q_PFULL0 + hf]);
    }
    mbar_wait(&bar[Q_OFULL0 + hf], (NT - 1) & 1);
    tc_fence_after();
    {
      uint32_t x[32];
      tmem_ld_32x32(t_lane + BTM_O + 64 * hf + 32 * sub, x);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(x[i]));
    }
    tc_fence_before();

    // merge the two key halves of every row: half 1 leaves its output slice, maximum and row sums in the (dead) q tiles
    float* mrow = (float*)(smem + BO_QU) + (size_t)r * B6_MROW;
    if (hf == 1) {
#pragma unroll
      for (int k = 0; k < 8; k++) *(float4*)(mrow + 36 * sub + 4 * k) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
      mrow[36 * sub + 32] = m_run;
      mrow[36 * sub + 33] = l_run;
    } else {
      mrow[72 + sub] = l_run;                        // (not xch: the partner may still be reading this tile's maximum from it)
    }
    asm volatile("bar.sync 9, 512;" ::: "memory");
    if (hf == 0 && row < a.T) {
      const float l0 = l_run + mrow[72 + (sub ^ 1)];
      const float m1 = mrow[36 * sub + 32], l1 = mrow[33] + mrow[36 + 33];
      const float m = fmaxf(m_run, m1);
      const float w0 = ex2_fast((m_run - m) * c), w1 = ex2_fast((m1 - m) * c);
      const float inv = 1.f / (l0 * w0 + l1 * w1);
      const float* other = mrow + 36 * sub;
      bf16* orow = a.out + ((long long)b * a.T + row) * HD + h * 64 + 32 * sub;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int d0 = 8 * k + 2 * e;
          w[e] = pack_bf16x2((o[d0] * w0 + other[d0] * w1) * inv, (o[d0 + 1] * w0 + other[d0 + 1] * w1) * inv);
        }
        *(uint4*)(orow + 8 * k) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == B6_SOFT_WARPS + 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace

// same contract as attn_bert_tc (attention_bert_tc.cu)
int attn_bert_tc16(const bf16* qkv, const bf16* rd, int Dcap, const float* u, const float* v, bf16* out, int B, int T, int H, float scale,
                   cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    DMG_CUDA_OK(cudaFuncSetAttribute(attn_bert_tc16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B6_SMEM));
    configured = true;
  }
  const int HD = H * 64;
  const TensorMap2D *tx = nullptr, *tr = nullptr;
  if (train_get_tmap(qkv, 3 * HD, (long long)B * T, 3 * HD, 128, &tx)) return -1;
  if (train_get_tmap(rd, 64, (long long)H * Dcap, 64, 128, &tr)) return -1;
  BertTcArgs a;
  a.u = u; a.v = v; a.out = out; a.B = B; a.T = T; a.H = H; a.Dcap = Dcap; a.scale = scale;
  return launch_k(attn_bert_tc16_kernel, dim3(B * H * ((T + 127) / 128)), dim3(B6_THREADS), (size_t)B6_SMEM, st, 1,
                  *(const CUtensorMap*)tx->bytes, *(const CUtensorMap*)tr->bytes, a);
}

}  // namespace dmg
