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

// all `nthreads` threads of the CTA copy a 64-row x 64-col bf16 block (row stride ld elements) into a tile
__device__ __forceinline__ void tile_load_async(uint8_t* tile, const bf16* g, long long ld, int tid, int nthreads) {
  for (int i = tid; i < 512; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    cp_async16(tile + tile_off(r, c), g + (long long)r * ld + c * 8);
  }
}

// A fragment (16 rows m0.., 16 k k0..) from a tile stored [m][k]
__device__ __forceinline__ void frag_a(uint32_t tile_addr, int m0, int k0, int lane, uint32_t (&a)[4]) {
  ldsm_x4(tile_addr + tile_off(m0 + (lane & 15), (k0 >> 3) + (lane >> 4)), a);
}
// A fragment from a tile stored [k][m] (transposed read)
__device__ __forceinline__ void frag_a_t(uint32_t tile_addr, int m0, int k0, int lane, uint32_t (&a)[4]) {
  ldsm_x4_t(tile_addr + tile_off(k0 + (lane & 7) + ((lane >> 4) << 3), (m0 >> 3) + ((lane >> 3) & 1)), a);
}
// B fragments of two neighbouring n-tiles (n0..n0+15) x 16 k from a tile stored [n][k]: (r0,r1) -> n-tile 0, (r2,r3) -> n-tile 1
__device__ __forceinline__ void frag_b(uint32_t tile_addr, int n0, int k0, int lane, uint32_t (&r)[4]) {
  ldsm_x4(tile_addr + tile_off(n0 + (lane & 7) + ((lane >> 4) << 3), (k0 >> 3) + ((lane >> 3) & 1)), r);
}
// same from a tile stored [k][n] (transposed read)
__device__ __forceinline__ void frag_b_t(uint32_t tile_addr, int n0, int k0, int lane, uint32_t (&r)[4]) {
  ldsm_x4_t(tile_addr + tile_off(k0 + (lane & 7) + (((lane >> 3) & 1) << 3), (n0 >> 3) + (lane >> 4)), r);
}

}  // namespace dmg
