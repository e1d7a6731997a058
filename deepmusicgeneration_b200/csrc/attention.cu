// Relative-positional attention with segment-recurrence memory (Transformer-XL) - fused kernels.
//
// Reference being replaced: fastai MultiHeadRelativeAttention._apply_attention and its
// in-repo twin MemMultiHeadRelativeAttentionKV._apply_attention (deep_music_remix.py:2078-2104):
//     AC = (q+u) K^T ; BD = _line_shift((q+v) R^T) ; P = softmax((AC+BD)/sqrt(Dh) + mask) ; out = P V
// _line_shift is never materialised: with Rd[dist] = r_attn(PositionalEncoding(dist)) cached per layer,
//     j <= M+i   : BD[i,j] = (q_i+v)     . Rd[M+i-j]
//     j == M+i+1 : BD[i,j] = 0                                   (BERT encoder only; masked in the causal TXL)
//     j >  M+i+1 : BD[i,j] = (q_{i+1}+v) . Rd[S+1+i-j]           (the live wrap-around of the unmasked encoder)
// (SURVEY.md App. A.4).  K/V of past tokens live in a per-(stream, head) ring in HBM, slot = token index mod M.
//
//  * attn_decode_kernel  - x_len == 1, bf16 ring: the HBM-bound kernel of batched generation.  One CTA per
//    (stream, head); K, R and V tiles stream through a 4-stage shared-memory ring filled by the TMA engine
//    (cp.async.bulk + mbarrier complete_tx); exact two-pass softmax over the M+1 scores held in shared memory;
//    the new token's K/V are appended to the ring by the same CTA after its last read of the oldest slot.
//  * attn_general_kernel - any x_len (prefill, training-shape forward, BERT encoder), fp32 or bf16 ring,
//    flash-style online softmax over 32x32 tiles on the FFMA pipe.
#include "kernels.cuh"
#include "launch.cuh"

namespace dmg {

// =============================================================================================
// decode kernel
// =============================================================================================
constexpr int DEC_STAGES = 4;
constexpr int DEC_STAGE_BYTES = 16384;   // K item: 64 keys K + 64 rows R;  V item: 128 keys
constexpr float LOG2E = 1.4426950408889634f;

__host__ __device__ inline int dec_smem_bytes(int M) {
  return DEC_STAGES * DEC_STAGE_BYTES + (M + 8) * 4 /*scores*/ + 16 * 64 * 4 /*group partials*/ + 128 * 4 /*qu,qv*/ +
         64 /*reduction scratch*/ + DEC_STAGES * 8 /*mbarriers*/ + 128 /*alignment*/;
}

__global__ void __launch_bounds__(128) attn_decode_kernel(AttnDecodeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint8_t* stages = base;
  float* sc = (float*)(stages + DEC_STAGES * DEC_STAGE_BYTES);
  const int M = a.M;
  float* red = sc + (M + 8);
  float* qu = red + 16 * 64;
  float* qv = qu + 64;
  float* scratch = qv + 64;                    // 16 floats
  uint64_t* full = (uint64_t*)(scratch + 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int H = a.H, HD = H * 64;
  pdl_launch_dependents();
  pdl_wait();
  const int pos_total = a.dev_state[0], mc = a.dev_state[1];
  const int head = pos_total % M;
  const int nK = M / 64, nV = M / 128, nItems = nK + nV;

  const bf16* kbase = a.kring + ((size_t)b * H + h) * M * 64;
  const bf16* vbase = a.vring + ((size_t)b * H + h) * M * 64;
  const bf16* rbase = a.rd + (size_t)h * a.Dcap * 64;
  const float* qrow = a.qkv + (size_t)b * 3 * HD + h * 64;

  if (tid == 0) {
    for (int s = 0; s < DEC_STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  if (tid < 64) {
    const float q = qrow[tid];
    qu[tid] = q + a.u[h * 64 + tid];
    qv[tid] = q + a.v[h * 64 + tid];
  }
  __syncthreads();

  auto issue = [&](int item) {
    const int s = item % DEC_STAGES;
    uint8_t* dst = stages + s * DEC_STAGE_BYTES;
    mbar_expect_tx(&full[s], DEC_STAGE_BYTES);
    if (item < nK) {
      const int p0 = item * 64;
      bulk_g2s(dst, kbase + (size_t)p0 * 64, 8192, &full[s]);
      uint8_t* rdst = dst + 8192;
      // slot p0+r holds the key at distance (p<head ? head-p : M+head-p); shared row 63-r <-> slot p0+r
      if (p0 + 64 <= head) {
        bulk_g2s(rdst, rbase + (size_t)(head - p0 - 63) * 64, 8192, &full[s]);
      } else if (p0 >= head) {
        bulk_g2s(rdst, rbase + (size_t)(M + head - p0 - 63) * 64, 8192, &full[s]);
      } else {
        const int n1 = head - p0;   // slots below head: distances n1..1 -> rows 64-n1..63
        bulk_g2s(rdst + (64 - n1) * 128, rbase + (size_t)1 * 64, n1 * 128, &full[s]);
        bulk_g2s(rdst, rbase + (size_t)(M + n1 - 63) * 64, (64 - n1) * 128, &full[s]);
      }
    } else {
      const int p0 = (item - nK) * 128;
      bulk_g2s(dst, vbase + (size_t)p0 * 64, 16384, &full[s]);
    }
  };
  if (tid == 0) {
    for (int it = 0; it < DEC_STAGES && it < nItems; it++) issue(it);
  }

  const int g = tid >> 3, c = tid & 7;
  float quf[8], qvf[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    quf[j] = qu[c * 8 + j];
    qvf[j] = qv[c * 8 + j];
  }
  const float sscale = a.scale * LOG2E;

  // ---------------- phase 1: scores over the ring ----------------
  for (int it = 0; it < nK; it++) {
    const int s = it % DEC_STAGES;
    mbar_wait(&full[s], (it / DEC_STAGES) & 1);
    const uint8_t* st = stages + s * DEC_STAGE_BYTES;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int r = g + 16 * t;
      const uint4 kw = *(const uint4*)(st + r * 128 + c * 16);
      const uint4 rw = *(const uint4*)(st + 8192 + (63 - r) * 128 + c * 16);
      float acc = quf[0] * bf16lo(kw.x);
      acc = fmaf(quf[1], bf16hi(kw.x), acc);
      acc = fmaf(quf[2], bf16lo(kw.y), acc);
      acc = fmaf(quf[3], bf16hi(kw.y), acc);
      acc = fmaf(quf[4], bf16lo(kw.z), acc);
      acc = fmaf(quf[5], bf16hi(kw.z), acc);
      acc = fmaf(quf[6], bf16lo(kw.w), acc);
      acc = fmaf(quf[7], bf16hi(kw.w), acc);
      float acc2 = qvf[0] * bf16lo(rw.x);
      acc2 = fmaf(qvf[1], bf16hi(rw.x), acc2);
      acc2 = fmaf(qvf[2], bf16lo(rw.y), acc2);
      acc2 = fmaf(qvf[3], bf16hi(rw.y), acc2);
      acc2 = fmaf(qvf[4], bf16lo(rw.z), acc2);
      acc2 = fmaf(qvf[5], bf16hi(rw.z), acc2);
      acc2 = fmaf(qvf[6], bf16lo(rw.w), acc2);
      acc2 = fmaf(qvf[7], bf16hi(rw.w), acc2);
      acc += acc2;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (c == 0) {
        const int p = it * 64 + r;
        const int dist = p < head ? head - p : M + head - p;
        sc[p] = dist <= mc ? acc * sscale : -INFINITY;
      }
    }
    __syncthreads();
    if (tid == 0 && it + DEC_STAGES < nItems) issue(it + DEC_STAGES);
  }

  // the new token itself (distance 0, always visible); its K/V are still fp32 in the qkv buffer
  if (warp == 0) {
    const float k0 = qrow[HD + lane], k1 = qrow[HD + lane + 32];
    const float r0 = __bfloat162float(rbase[lane]), r1 = __bfloat162float(rbase[lane + 32]);
    float acc = qu[lane] * k0 + qu[lane + 32] * k1 + qv[lane] * r0 + qv[lane + 32] * r1;
    acc = warp_sum(acc);
    if (lane == 0) sc[M] = acc * sscale;
  }
  __syncthreads();

  // ---------------- exact softmax over M+1 scores (kept un-normalised; 1/sum applied at the end) ----------------
  float mx = -INFINITY;
  for (int j = tid; j <= M; j += 128) mx = fmaxf(mx, sc[j]);
  mx = warp_max(mx);
  if (lane == 0) scratch[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(scratch[0], scratch[1]), fmaxf(scratch[2], scratch[3]));
  float sum = 0.f;
  for (int j = tid; j <= M; j += 128) {
    const float p = exp2f(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) scratch[4 + warp] = sum;
  __syncthreads();
  sum = scratch[4] + scratch[5] + scratch[6] + scratch[7];

  // ---------------- phase 2: P.V ----------------
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; j++) o[j] = 0.f;
  for (int iv = 0; iv < nV; iv++) {
    const int it = nK + iv;
    const int s = it % DEC_STAGES;
    mbar_wait(&full[s], (it / DEC_STAGES) & 1);
    const uint8_t* st = stages + s * DEC_STAGE_BYTES;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int r = g + 16 * t;
      const float p = sc[iv * 128 + r];
      const uint4 vw = *(const uint4*)(st + r * 128 + c * 16);
      o[0] = fmaf(p, bf16lo(vw.x), o[0]);
      o[1] = fmaf(p, bf16hi(vw.x), o[1]);
      o[2] = fmaf(p, bf16lo(vw.y), o[2]);
      o[3] = fmaf(p, bf16hi(vw.y), o[3]);
      o[4] = fmaf(p, bf16lo(vw.z), o[4]);
      o[5] = fmaf(p, bf16hi(vw.z), o[5]);
      o[6] = fmaf(p, bf16lo(vw.w), o[6]);
      o[7] = fmaf(p, bf16hi(vw.w), o[7]);
    }
    __syncthreads();
    if (tid == 0 && it + DEC_STAGES < nItems) issue(it + DEC_STAGES);
  }
  if (g == 0) {
    const float p = sc[M];
#pragma unroll
    for (int j = 0; j < 8; j++) o[j] = fmaf(p, qrow[2 * HD + c * 8 + j], o[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) red[g * 64 + c * 8 + j] = o[j];
  __syncthreads();
  if (tid < 64) {
    float tot = 0.f;
#pragma unroll
    for (int gg = 0; gg < 16; gg++) tot += red[gg * 64 + tid];
    a.out[(size_t)b * HD + h * 64 + tid] = __float2bfloat16_rn(tot / sum);
  } else {
    // ring append (K13): this CTA is the only reader of the (stream, head) ring and has finished with slot `head`
    const int e = tid - 64;
    const size_t o_ = ((size_t)b * H + h) * M * 64 + (size_t)head * 64 + e;
    a.kring[o_] = __float2bfloat16_rn(qrow[HD + e]);
    a.vring[o_] = __float2bfloat16_rn(qrow[2 * HD + e]);
  }
}

bool attn_decode_supported(int Dh, int M) { return Dh == 64 && M >= 128 && M % 128 == 0 && dec_smem_bytes(M) <= 227 * 1024; }

int attn_decode(const AttnDecodeArgs& a, cudaStream_t st) {
  DMG_CHECK(a.Dcap >= a.M + 1, "attn_decode: rel-pos cache too small (%d < %d)", a.Dcap, a.M + 1);
  const int smem = dec_smem_bytes(a.M);
  static int configured = 0;
  if (configured < smem) {
    DMG_CUDA_OK(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  return launch_k(attn_decode_kernel, dim3(a.H, a.B), dim3(128), smem, st, 1, a);
}

// =============================================================================================
// general kernel
// =============================================================================================
constexpr int GQ = 32, GK = 32, GLD = 68;   // query tile, key tile, padded row stride (floats)

template <bool BERT>
__host__ __device__ constexpr int gen_smem_floats() {
  return GQ * GLD /*Qu*/ + (GQ + 1) * GLD /*Qv*/ + GK * GLD /*K*/ + GK * GLD /*V*/ + 63 * GLD /*R lower*/ +
         (BERT ? 63 * GLD : 0) /*R upper*/ + GQ * (GK + 1) /*P*/;
}

template <class T, bool BERT>
__global__ void __launch_bounds__(128) attn_general_kernel(AttnGeneralArgs a) {
  extern __shared__ float sm[];
  float* Qu = sm;
  float* Qv = Qu + GQ * GLD;
  float* Ks = Qv + (GQ + 1) * GLD;
  float* Vs = Ks + GK * GLD;
  float* RL = Vs + GK * GLD;
  float* RU = RL + 63 * GLD;
  float* P = RU + (BERT ? 63 * GLD : 0);

  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const int i0 = blockIdx.x * GQ, h = blockIdx.y, b = blockIdx.z;
  const int T_len = a.T, H = a.H, HD = H * 64, M = a.M;
  const int m = a.mem_count, S = m + T_len;
  const T* kring = (const T*)a.kring;
  const T* vring = (const T*)a.vring;
  const T* rd = (const T*)a.rd + (size_t)h * a.Dcap * 64;

  // q (+u / +v) rows; Qv holds one extra row (i0+32) for the wrap-around term
  for (int idx = tid; idx < (GQ + 1) * 16; idx += 128) {
    const int r = idx >> 4, c4 = (idx & 15) * 4;
    const int i = i0 + r;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < T_len) q = *(const float4*)(a.qkv + ((size_t)b * T_len + i) * 3 * HD + h * 64 + c4);
    const float4 uu = *(const float4*)(a.u + h * 64 + c4), vv = *(const float4*)(a.v + h * 64 + c4);
    if (r < GQ) *(float4*)(Qu + r * GLD + c4) = make_float4(q.x + uu.x, q.y + uu.y, q.z + uu.z, q.w + uu.w);
    *(float4*)(Qv + r * GLD + c4) = make_float4(q.x + vv.x, q.y + vv.y, q.z + vv.z, q.w + vv.w);
  }

  const int qi = tid >> 2, kg = tid & 3;
  const int i = i0 + qi;
  float m_run = -INFINITY, l_run = 0.f;
  float o[16];
#pragma unroll
  for (int e = 0; e < 16; e++) o[e] = 0.f;

  int jend = S;
  if (!BERT) jend = min(S, m + i0 + GQ);   // every mask the reference builds hides keys with jx > i

  for (int j0 = 0; j0 < jend; j0 += GK) {
    __syncthreads();
    // ---- K / V tile: ring slots for memory, the qkv buffer for the segment itself
    {
      const int r = tid >> 2, c16 = (tid & 3) * 16;
      const int j = j0 + r;
      float kv[16], vv[16];
      if (j < S) {
        if (j < m) {
          const int slot = (int)((a.pos_total - m + j) % M);
          const size_t off = (((size_t)(a.b0 + b) * H + h) * M + slot) * 64 + c16;
#pragma unroll
          for (int e = 0; e < 16; e++) {
            kv[e] = to_f32(kring[off + e]);
            vv[e] = to_f32(vring[off + e]);
          }
        } else {
          const float* row = a.qkv + ((size_t)b * T_len + (j - m)) * 3 * HD + h * 64 + c16;
#pragma unroll
          for (int e = 0; e < 16; e++) {
            kv[e] = row[HD + e];
            vv[e] = row[2 * HD + e];
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; e++) kv[e] = vv[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 16; e++) {
        Ks[r * GLD + c16 + e] = kv[e];
        Vs[r * GLD + c16 + e] = vv[e];
      }
    }
    // ---- relative-position key windows: row w <-> (qi - jl + 31)
    {
      const int dLo = m + i0 - j0 - 31;
      const int dUp = S + 1 + i0 - j0 - 31;
      for (int idx = tid; idx < 63 * 16; idx += 128) {
        const int w = idx >> 4, c4 = (idx & 15) * 4;
        int dist = dLo + w;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dist >= 0 && dist < a.Dcap) {
          const T* p = rd + (size_t)dist * 64 + c4;
          val = make_float4(to_f32(p[0]), to_f32(p[1]), to_f32(p[2]), to_f32(p[3]));
        }
        *(float4*)(RL + w * GLD + c4) = val;
        if (BERT) {
          dist = dUp + w;
          val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (dist >= 0 && dist < a.Dcap) {
            const T* p = rd + (size_t)dist * 64 + c4;
            val = make_float4(to_f32(p[0]), to_f32(p[1]), to_f32(p[2]), to_f32(p[3]));
          }
          *(float4*)(RU + w * GLD + c4) = val;
        }
      }
    }
    __syncthreads();

    // ---- scores
    float sv[8];
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const int jl = kg + 4 * jj;
      const int j = j0 + jl;
      bool valid = (i < T_len) && (j < S);
      const float* rrow = RL + (qi - jl + 31) * GLD;
      const float* qvrow = Qv + qi * GLD;
      float bdw = 1.f;
      if (BERT) {
        if (j == m + i + 1) bdw = 0.f;
        else if (j > m + i + 1) {
          rrow = RU + (qi - jl + 31) * GLD;
          qvrow = Qv + (qi + 1) * GLD;
        }
      } else {
        const int jx = j - m;
        if (jx > 0) valid = valid && ((jx / a.win - i / a.win) < a.k);
      }
      const float* qurow = Qu + qi * GLD;
      const float* krow = Ks + jl * GLD;
      float ac = 0.f, bd = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < 64; d4 += 4) {
        const float4 qa = *(const float4*)(qurow + d4), kk = *(const float4*)(krow + d4);
        const float4 qb = *(const float4*)(qvrow + d4), rr = *(const float4*)(rrow + d4);
        ac = fmaf(qa.x, kk.x, ac); ac = fmaf(qa.y, kk.y, ac); ac = fmaf(qa.z, kk.z, ac); ac = fmaf(qa.w, kk.w, ac);
        bd = fmaf(qb.x, rr.x, bd); bd = fmaf(qb.y, rr.y, bd); bd = fmaf(qb.z, rr.z, bd); bd = fmaf(qb.w, rr.w, bd);
      }
      const float s = valid ? (ac + bdw * bd) * a.scale : -INFINITY;
      sv[jj] = s;
      tmax = fmaxf(tmax, s);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float m_new = fmaxf(m_run, tmax);
    const float corr = (m_new == -INFINITY) ? 1.f : expf(m_run - m_new);
    float psum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const float p = (m_new == -INFINITY) ? 0.f : expf(sv[jj] - m_new);
      P[qi * (GK + 1) + kg + 4 * jj] = p;
      psum += p;
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l_run = l_run * corr + psum;
    m_run = m_new;
#pragma unroll
    for (int e = 0; e < 16; e++) o[e] *= corr;
    __syncwarp();

    // ---- P.V : this thread owns output dims [kg*16, kg*16+16) of query qi
#pragma unroll 4
    for (int jl = 0; jl < GK; jl++) {
      const float p = P[qi * (GK + 1) + jl];
      const float* vrow = Vs + jl * GLD + kg * 16;
#pragma unroll
      for (int e4 = 0; e4 < 16; e4 += 4) {
        const float4 vv = *(const float4*)(vrow + e4);
        o[e4] = fmaf(p, vv.x, o[e4]); o[e4 + 1] = fmaf(p, vv.y, o[e4 + 1]);
        o[e4 + 2] = fmaf(p, vv.z, o[e4 + 2]); o[e4 + 3] = fmaf(p, vv.w, o[e4 + 3]);
      }
    }
  }
  if (i < T_len) {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    T* out = (T*)a.out + ((size_t)b * T_len + i) * HD + h * 64 + kg * 16;
#pragma unroll
    for (int e = 0; e < 16; e++) out[e] = from_f32<T>(o[e] * inv);
  }
}

template <class T, bool BERT>
static int launch_general(const AttnGeneralArgs& a, cudaStream_t st) {
  const int smem = gen_smem_floats<BERT>() * 4;
  static bool configured = false;
  if (!configured) {
    DMG_CUDA_OK(cudaFuncSetAttribute(attn_general_kernel<T, BERT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((a.T + GQ - 1) / GQ, a.H, a.B);
  return launch_k(attn_general_kernel<T, BERT>, grid, dim3(128), smem, st, 1, a);
}

template <class T>
int attn_general(const AttnGeneralArgs& a, cudaStream_t st) {
  DMG_CHECK(a.win >= 1, "attn_general: window size must be >= 1");
  if (a.B <= 0 || a.T <= 0) return 0;
  return a.bert ? launch_general<T, true>(a, st) : launch_general<T, false>(a, st);
}
template int attn_general<float>(const AttnGeneralArgs&, cudaStream_t);
template int attn_general<bf16>(const AttnGeneralArgs&, cudaStream_t);

}  // namespace dmg
