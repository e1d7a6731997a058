// Training data feed on the device.
//
// preload_fill_kernel: MusicPreloader.__getitem__ / fill_row (deep_music_genre.py:1088-1125) for ALL `bs` rows of one batch in
// one launch.  The tokenised corpus stays in HBM as a flat ragged array (tokens, positions, offsets); a batch row is a walk
// over consecutive items of the (shuffled) CircularIndex (:1005-1014) starting at the row's (ro, ri) cursor, `bptt + y_offset`
// tokens long, with the per-item random transpose (MusicItem.transpose :1247 -> tfm_transpose :1541-1544) applied on the fly
// and x / y = the row without its last / first y_offset tokens (batch_position_tfm :1129-1136 splits index and position).
// One CTA per row: thread 0 walks the items and lists the copy segments, all threads copy.  HBM-bound by construction
// (4 bytes read + 16 bytes written per token), latency-bound in practice (a batch is bs * bptt tokens).
//
// mask_tfm_kernel: the BERT-style batch masking of the remix encoder (deep_music_remix.py:1208-1223), element-wise with a
// counter-based generator (the reference draws torch.rand / torch.randint; the draws can be exported for the oracle).
#include "kernels.cuh"
#include "launch.cuh"
#include "train_kernels.cuh"

namespace dmg {

namespace {

constexpr int PL_MAXSEG = 64;

struct PreloadArgs {
  const int32_t* tokens; const int32_t* positions; const long long* offsets; const long long* perm; const int32_t* transpose;
  long long* ro; long long* ri;
  long long *x, *y, *xpos;
  int n_items, forward, note_lo, note_hi, bs, bptt, y_offset;
};

__global__ void __launch_bounds__(128) preload_fill_kernel(const PreloadArgs a) {
  __shared__ long long seg_src[PL_MAXSEG];
  __shared__ int seg_dst[PL_MAXSEG], seg_n[PL_MAXSEG], seg_tv[PL_MAXSEG];
  __shared__ int nseg, done;
  __shared__ long long s_ro, s_ri;
  __shared__ int s_ibuf, s_n;
  const int j = blockIdx.x, L = a.bptt + a.y_offset;
  if (threadIdx.x == 0) { s_ro = a.ro[j] - 1; s_ri = a.ri[j]; s_ibuf = 0; s_n = 0; done = 0; }
  __syncthreads();
  while (true) {
    if (threadIdx.x == 0) {                         // fill_row's while loop, up to PL_MAXSEG items at a time
      long long ro = s_ro, ri = s_ri;
      int ibuf = s_ibuf, n = s_n, ns = 0;
      while (ibuf < L && ns < PL_MAXSEG) {
        ro += 1;
        const long long m = ro % a.n_items;
        const long long ix = a.perm[a.forward ? m : a.n_items - 1 - m];
        const long long base = a.offsets[ix];
        const int len = (int)(a.offsets[ix + 1] - base);
        if (a.forward) {
          ri = ibuf ? 0 : ri;
          n = min((int)(len - ri), L - ibuf);
          seg_src[ns] = base + ri;                  // row[ibuf + t] = rag[ri + t]
        } else {
          ri = ibuf ? len : ri;
          n = min((int)ri, L - ibuf);
          seg_src[ns] = base + ri - 1;              // row[ibuf + t] = rag[ri - 1 - t]
        }
        seg_dst[ns] = ibuf; seg_n[ns] = n; seg_tv[ns] = a.transpose ? a.transpose[ix] : 0;
        ns++;
        ibuf += n;
      }
      s_ro = ro; s_ri = ri; s_ibuf = ibuf; s_n = n; nseg = ns;
      if (ibuf >= L) {
        done = 1;
        a.ro[j] = ro;
        a.ri[j] = ri + (a.forward ? (n - 1) : -(n - 1));   // overlap = 1: the next batch re-reads this row's last token
      }
    }
    __syncthreads();
    for (int sgi = 0; sgi < nseg; sgi++) {
      const long long src = seg_src[sgi];
      const int dst = seg_dst[sgi], n = seg_n[sgi], tv = seg_tv[sgi];
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const long long s = a.forward ? src + t : src - t;
        int tok = a.tokens[s];
        if (tok >= a.note_lo && tok < a.note_hi) tok += tv;
        const int p = dst + t;
        if (p < a.bptt) {
          a.x[(long long)j * a.bptt + p] = tok;
          if (a.xpos) a.xpos[(long long)j * a.bptt + p] = a.positions[s];
        }
        if (p >= a.y_offset) a.y[(long long)j * a.bptt + (p - a.y_offset)] = tok;
      }
    }
    __syncthreads();
    if (done) break;
  }
}

struct MaskTfmArgs {
  long long *x, *y;
  float* rand_out; long long* wrong_out;
  long long n;
  int lo, hi, mask_idx, pad_idx;
  float p, p8, p9;
  uint32_t seed;
};

__global__ void __launch_bounds__(256) mask_tfm_kernel(const MaskTfmArgs a) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  const uint32_t h0 = mix32((uint32_t)i ^ a.seed), h1 = mix32(h0 ^ 0x9E3779B9u);
  float r = (float)(h0 >> 8) * (1.0f / 16777216.0f);                      // uniform [0, 1), 24 bits like torch.rand
  const long long wrong = a.lo + (long long)(h1 % (uint32_t)(a.hi - a.lo));
  if (a.rand_out) a.rand_out[i] = r;
  if (a.wrong_out) a.wrong_out[i] = wrong;
  const long long xv = a.x[i];
  if (xv < a.lo || xv >= a.hi) r = 1.0f;
  if (r > a.p) a.y[i] = a.pad_idx;                                         // unchanged tokens leave the loss
  if (r <= a.p8) a.x[i] = a.mask_idx;                                      // 80 %: mask token
  else if (r <= a.p9) a.x[i] = wrong;                                      // 10 %: wrong token (the last 10 % stay as they are)
}

}  // namespace

}  // namespace dmg

extern "C" {

int dmg_preload_fill(const int32_t* tokens, const int32_t* positions, const int64_t* offsets, const int64_t* perm, int n_items,
                     int forward, const int32_t* transpose, int note_lo, int note_hi, int64_t* ro, int64_t* ri, int bs, int bptt,
                     int y_offset, int64_t* x, int64_t* y, int64_t* xpos, void* stream) {
  using namespace dmg;
  DMG_CHECK(tokens && offsets && perm && ro && ri && x && y, "dmg_preload_fill: null argument");
  DMG_CHECK(n_items > 0 && bs > 0 && bptt > 0 && y_offset >= 0 && y_offset <= bptt, "dmg_preload_fill: bad sizes (n_items %d, bs %d, bptt %d, y_offset %d)",
            n_items, bs, bptt, y_offset);
  DMG_CHECK(!xpos || positions, "dmg_preload_fill: xpos requested without a position array");
  DMG_CHECK(forward || !xpos, "dmg_preload_fill: a backwards epoch with encode_position fails in the reference too (row.size, deep_music_genre.py:1120)");
  PreloadArgs a;
  a.tokens = tokens; a.positions = positions; a.offsets = (const long long*)offsets; a.perm = (const long long*)perm; a.transpose = transpose;
  a.ro = (long long*)ro; a.ri = (long long*)ri; a.x = (long long*)x; a.y = (long long*)y; a.xpos = (long long*)xpos;
  a.n_items = n_items; a.forward = forward ? 1 : 0; a.note_lo = note_lo; a.note_hi = note_hi; a.bs = bs; a.bptt = bptt; a.y_offset = y_offset;
  return launch_np(preload_fill_kernel, dim3(bs), dim3(128), 0, (cudaStream_t)stream, a);
}

int dmg_mask_tfm(int64_t* x, int64_t* y, int64_t n, int mask_lo, int mask_hi, int mask_idx, int pad_idx, double p, uint32_t seed,
                 float* rand_out, int64_t* wrong_out, void* stream) {
  using namespace dmg;
  DMG_CHECK(x && y && n >= 0 && mask_hi > mask_lo, "dmg_mask_tfm: bad argument");
  if (n == 0) return 0;
  MaskTfmArgs a;
  a.x = (long long*)x; a.y = (long long*)y; a.rand_out = rand_out; a.wrong_out = (long long*)wrong_out; a.n = n;
  a.lo = mask_lo; a.hi = mask_hi; a.mask_idx = mask_idx; a.pad_idx = pad_idx;
  a.p = (float)p; a.p8 = (float)(p * .8); a.p9 = (float)(p * .9);   // the reference compares a float32 tensor with python doubles
  a.seed = seed;
  return launch_np(mask_tfm_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, a);
}

}  // extern "C"
