// Warp-level tensor-core helpers (mma.sync m16n8k16 bf16, ldmatrix, cp.async) and the 64x64 bf16 shared-memory tile
// with a 16-byte-chunk XOR swizzle used by the training attention kernels.
#pragma once
#include "common.cuh"

namespace dmg {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D (fp32, 16x8) += A (bf16 16x16, row) * B (bf16 16x8, col)
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- 64 x 64 bf16 tile: row r = 128 bytes, 16-byte chunk c stored at chunk position (c ^ (r & 7))
constexpr int TILE_BYTES = 64 * 64 * 2;
__device__ __forceinline__ uint32_t tile_off(int r, int chunk) { return (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4)); }

// 128 threads copy a 64-row x 64-col bf16 block (row stride ld elements) into a tile: thread t moves the 16-byte chunk
// (t & 7) of rows (t >> 3) + 16 k, k = 0..3 - the swizzled chunk position is the same for all four (row & 7 is unchanged)
__device__ __forceinline__ void tile_load_async(uint8_t* tile, const bf16* g, long long ld, int tid, int nthreads) {
  (void)nthreads;
  const int r = tid >> 3, c = tid & 7;
  uint8_t* dst = tile + tile_off(r, c);
  const bf16* src = g + (long long)r * ld + c * 8;
#pragma unroll
  for (int k = 0; k < 4; k++) cp_async16(dst + k * 2048, src + (long long)k * 16 * ld);
}

// Per-lane byte offsets of the ldmatrix row addresses inside a tile, computed once per kernel: the XOR swizzle only mixes
// lane-constant bits (row & 7 == lane & 7 for every fragment shape below) with the 16-byte chunk index.
struct LaneOff {
  uint32_t b[4];    // [ks] B fragment from a [n][k] tile:  tile + n0*128 + b[ks]     (k0 = 16*ks)
  uint32_t bt[4];   // [np] B fragment from a [k][n] tile:  tile + k0*128 + bt[np]    (n0 = 16*np)
  uint32_t a[4];    // [ks] A fragment from a [m][k] tile:  tile + m0*128 + a[ks]
  uint32_t at;      //      A fragment from a [k][m] tile:  tile + k0*128 + at        (m0 = 16*warp, folded in)
};
__device__ __forceinline__ void lane_off_init(LaneOff& L, int lane, int warp) {
  const uint32_t sw = lane & 7;
  const uint32_t rB = (lane & 7) + ((lane >> 4) << 3), kb = (lane >> 3) & 1;      // frag_b / frag_a_t row, chunk bit
  const uint32_t rT = (lane & 7) + (((lane >> 3) & 1) << 3), hb = lane >> 4;      // frag_b_t row, chunk bit
#pragma unroll
  for (int i = 0; i < 4; i++) {
    L.b[i] = rB * 128 + (((2 * i + kb) ^ sw) << 4);
    L.bt[i] = rT * 128 + (((2 * i + hb) ^ sw) << 4);
    L.a[i] = (lane & 15) * 128 + (((2 * i + hb) ^ sw) << 4);
  }
  L.at = rB * 128 + (((2 * warp + kb) ^ sw) << 4);
}
// n0, k0, m0: multiples of 16 (compile-time constants at the call sites -> folded into the ldmatrix immediate)
__device__ __forceinline__ void frag_b(uint32_t tile_addr, int n0, int ks, const LaneOff& L, uint32_t (&r)[4]) {
  ldsm_x4(tile_addr + n0 * 128 + L.b[ks], r);
}
__device__ __forceinline__ void frag_b_t(uint32_t tile_addr, int np, int k0, const LaneOff& L, uint32_t (&r)[4]) {
  ldsm_x4_t(tile_addr + k0 * 128 + L.bt[np], r);
}
__device__ __forceinline__ void frag_a(uint32_t tile_addr, int m0, int ks, const LaneOff& L, uint32_t (&a)[4]) {
  ldsm_x4(tile_addr + m0 * 128 + L.a[ks], a);
}
__device__ __forceinline__ void frag_a_t(uint32_t tile_addr, int k0, const LaneOff& L, uint32_t (&a)[4]) {
  ldsm_x4_t(tile_addr + k0 * 128 + L.at, a);
}

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace dmg
