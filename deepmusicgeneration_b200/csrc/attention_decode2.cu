// Stand-alone launch of the decode attention (x_len == 1 over the bf16 K/V ring); the kernel body lives in attention_decode2.cuh.
#include "attention_decode3.cuh"

namespace dmg {

bool attn_decode2_supported(int Dh, int M) { return Dh == 64 && M >= 64 && M % 64 == 0 && d2_pick_groups(M) > 0; }

template <int G>
__global__ void __launch_bounds__((4 * G + 1) * 32, 1)
attn_decode2_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmR, const AttnDecodeArgs a, int n_stages, int b0) {
  extern __shared__ __align__(1024) uint8_t d2_smem[];
  attn_decode2_body<G>(tmK, tmV, tmR, a, n_stages, b0, (int)blockIdx.x, (int)gridDim.x, d2_smem);
}

template <int G>
static int launch_d2(const TensorMap2D* tmK, const TensorMap2D* tmV, const TensorMap2D* tmR, const AttnDecodeArgs& a, int b0,
                     int num_sms, cudaStream_t st) {
  const int ns = d2_pick_stages(a.M, G);
  const D2Layout L = d2_layout(a.M, G, ns);
  static int configured = 0;
  if (configured < L.total) {
    DMG_CUDA_OK(cudaFuncSetAttribute(attn_decode2_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    configured = L.total;
  }
  DMG_CHECK(ns >= 2, "attn_decode2: no stage count fits (M=%d)", a.M);
  const long long NI = (long long)a.B * a.H;
  const int grid = (int)(NI < num_sms ? NI : num_sms);
  // a.kring / a.vring / a.qkv / a.out are chunk-local (offset by b0 streams); the TMA row coordinate adds b0
  return launch_k(attn_decode2_kernel<G>, dim3(grid), dim3((4 * G + 1) * 32), L.total, st, 1, *(const CUtensorMap*)tmK->bytes,
                  *(const CUtensorMap*)tmV->bytes, *(const CUtensorMap*)tmR->bytes, a, ns, b0);
}

// ---- third generation (attention_decode3.cuh): two consumer teams, rel-pos scores from a per-CTA table ----
bool attn_decode3_supported(int Dh, int M) { return Dh == 64 && M >= D3_KEYS && M <= 512 && M % D3_KEYS == 0 && d3_pick_stages(M) >= 2; }

__global__ void __launch_bounds__(D3_THREADS, 1)
attn_decode3_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmR, const AttnDecodeArgs a, int n_stages, int b0) {
  extern __shared__ __align__(1024) uint8_t d3_smem[];
  attn_decode3_body<D3_TEAMS>(tmK, tmV, tmR, a, n_stages, b0, (int)blockIdx.x, (int)gridDim.x, d3_smem);
}

// A CTA holds the rel-pos score table of at most D3_CAP items: larger batches go out as several launches over chunks of streams.
static int attn_decode3(const TensorMap2D* tmK, const TensorMap2D* tmV, const TensorMap2D* tmR, const AttnDecodeArgs& a_in, int b0,
                        int num_sms, cudaStream_t st) {
  const int ns = d3_pick_stages(a_in.M);
  const D3Layout L = d3_layout(a_in.M, ns);
  static int configured = 0;
  if (configured < L.total) {
    DMG_CUDA_OK(cudaFuncSetAttribute(attn_decode3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    configured = L.total;
  }
  const int HD = a_in.H * 64;
  const int per = (int)((long long)D3_CAP * num_sms / a_in.H);          // streams per launch
  DMG_CHECK(per >= 1, "attn_decode3: %d heads do not fit %d CTAs of %d items", a_in.H, num_sms, D3_CAP);
  for (int c0 = 0; c0 < a_in.B; c0 += per) {
    AttnDecodeArgs a = a_in;
    a.B = a_in.B - c0 < per ? a_in.B - c0 : per;
    a.qkv = a_in.qkv + (size_t)c0 * 3 * HD;
    a.out = a_in.out + (size_t)c0 * HD;
    a.kring = a_in.kring + (size_t)c0 * a_in.H * a_in.M * 64;
    a.vring = a_in.vring + (size_t)c0 * a_in.H * a_in.M * 64;
    const long long NI = (long long)a.B * a.H;
    const int grid = (int)(NI < num_sms ? NI : num_sms);
    if (launch_k(attn_decode3_kernel, dim3(grid), dim3(D3_THREADS), L.total, st, 1, *(const CUtensorMap*)tmK->bytes,
                 *(const CUtensorMap*)tmV->bytes, *(const CUtensorMap*)tmR->bytes, a, ns, b0 + c0)) return -1;
  }
  return 0;
}

int attn_decode2(const TensorMap2D* tmK, const TensorMap2D* tmV, const TensorMap2D* tmR, const AttnDecodeArgs& a_in, int b0,
                 int num_sms, cudaStream_t st) {
  static const int no_early = getenv("DMG_NO_EARLY_KV") ? 1 : 0;
  static const int env_v2 = getenv("DMG_ATTN_DECODE_V2") ? 1 : 0;
  AttnDecodeArgs a = a_in;
  a.no_early_kv = no_early;
  DMG_CHECK(a.Dcap >= a.M + 1, "attn_decode2: rel-pos cache too small (%d < %d)", a.Dcap, a.M + 1);
  if (!a.force_v2 && !env_v2 && attn_decode3_supported(64, a.M))
    return attn_decode3(tmK, tmV, tmR, a, b0, num_sms, st);
  const int G = d2_pick_groups(a.M);
  DMG_CHECK(G > 0, "attn_decode2: mem_len %d not supported", a.M);
  if (G == 4) return launch_d2<4>(tmK, tmV, tmR, a, b0, num_sms, st);
  if (G == 2) return launch_d2<2>(tmK, tmV, tmR, a, b0, num_sms, st);
  return launch_d2<1>(tmK, tmV, tmR, a, b0, num_sms, st);
}

}  // namespace dmg
