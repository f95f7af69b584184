// Warp-level tensor-core building blocks for the small-head-dim attention kernels (window attention d = 32,
// cross-modal attention d = r <= 96): mma.sync m16n8k16 bf16 -> fp32, ldmatrix, cp.async.
//
// Why not tcgen05 here: with d = 12..48 these kernels execute 64..160 MMA flops per softmax exponential, so
// they are bound by the MUFU / FMA pipes (16 ex2 per clock per SM), not by the tensor pipe -- the register
// resident flash formulation (scores never leave registers between the two matmuls) is what minimises the
// non-MMA work per score.  The GEMMs (gemm.cu) are the tcgen05 / TMEM / TMA kernels.
#pragma once
#include "common.cuh"

namespace stgcma {
namespace mma {

// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
//   lane = 4*g + t.  A: a0 (row g, k 2t..2t+1) a1 (row g+8, same k) a2 (row g, k 2t+8..) a3 (row g+8, k 2t+8..)
//   B: b0 (k 2t..2t+1, n g) b1 (k 2t+8.., n g).   C/D: c0 c1 (row g, n 2t, 2t+1)  c2 c3 (row g+8, same n)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// four 8x8 b16 matrices; lane i supplies the address of row (i % 8) of matrix (i / 8); r[j] <- matrix j,
// distributed so that lane holds (row lane/4, cols 2*(lane%4), +1)   [.trans: (row 2*(lane%4), +1 ; col lane/4)]
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// A fragment (16 rows x 16 k) from a row-major [row][k] bf16 tile with row stride `ld` BYTES
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const uint8_t* tile, int ld, int row0, int k0, int lane) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int k = k0 + (lane >> 4) * 8;
  ldsm_x4(a, smem_addr(tile + r * ld + k * 2));
}
// B fragments of TWO n-tiles (n0..n0+15) for one k-step from a [n][k] row-major tile (keys as rows, features
// as columns: the "K" operand of Q.K^T):  b[0],b[1] -> n-tile n0 ; b[2],b[3] -> n-tile n0+8
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], const uint8_t* tile, int ld, int n0, int k0, int lane) {
  const int n = n0 + (lane & 7) + (lane >> 4) * 8;
  const int k = k0 + ((lane >> 3) & 1) * 8;
  ldsm_x4(b, smem_addr(tile + n * ld + k * 2));
}
// B fragments of TWO n-tiles (n0..n0+15) for one k-step (k0..k0+15) from a [k][n] row-major tile (keys as rows
// = the contraction index, features as columns: the "V" operand of P.V):  b[0],b[1] -> n0 ; b[2],b[3] -> n0+8
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], const uint8_t* tile, int ld, int k0, int n0, int lane) {
  const int k = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int n = n0 + (lane >> 4) * 8;
  ldsm_x4_t(b, smem_addr(tile + k * ld + n * 2));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// max / sum over the 4 lanes that share an accumulator row
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

constexpr float kLog2e = 1.4426950408889634f;

}  // namespace mma
}  // namespace stgcma
