// Building blocks of the attention-type kernels on tcgen05 / TMEM (mha_tc.cu, xmodal_tc5.cu): canonical NO-SWIZZLE
// core-matrix shared-memory images and their descriptors (one image serves as K-major and as MN-major operand), MMA with
// the A operand in TMEM, TMEM load / store helpers, cp.async tile loaders.
#pragma once
#include "tc_common.cuh"

namespace stgcma {
namespace tc5 {

using namespace tc;

constexpr float kLog2e = 1.4426950408889634f;

// canonical no-swizzle descriptor: lbo / sbo in bytes
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}

// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_bf16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// rows [r0, r0 + nrows) of a sequence (row stride `ld` elements, HD columns from `src`) -> canonical core-matrix image:
// element (r, c) at (r / 8) * RG + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2 bytes.  Rows >= r_valid are zero-filled.
// A warp instruction covers 8 rows x 64 contiguous bytes (4 lanes per row): 8 global lines, all sectors fully used, and
// a 4-way shared-memory store conflict -- against 32 half-used lines when 8 lanes walk 8 rows of one 16-byte column
// (the first version: the L1 wavefronts of the copy were a third of the kernel, ncu profiles/r02b).
template <int HD>
__device__ __forceinline__ void load_rows(uint32_t dst, const bf16* __restrict__ src, int64_t ld, int r0, int nrows, int r_valid,
                                          int tid, int nthr) {
  constexpr int CH = HD / 8, RG = CH * 128, Q4 = CH / 4;
  static_assert(CH % 4 == 0, "head dim must be a multiple of 32");
  const int ld32 = (int)ld;
  for (int e = tid; e < nrows * CH; e += nthr) {
    const int kc4 = e & 3, r8 = (e >> 2) & 7, blk = e >> 5;
    const int kq = blk % Q4, rg = blk / Q4;
    const int kc = kq * 4 + kc4, r = rg * 8 + r8;
    const uint32_t d = dst + rg * RG + kc * 128 + r8 * 16;
    if (r0 + r < r_valid) cp16(d, src + (r0 + r) * ld32 + kc * 8);
    else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0u) : "memory");
  }
}

__device__ __forceinline__ void tmem_ld16u(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// two 16-column loads in flight before the wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float (&a)[16], float (&b)[16]) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(tb)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(q[i]);
  }
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// 16 fp32 accumulator columns of this thread's row -> bf16, 32 bytes to global memory
__device__ __forceinline__ void store16(bf16* dst, const float (&v)[16], float sc) {
  uint4 a, b;
  a.x = pack2(v[0] * sc, v[1] * sc); a.y = pack2(v[2] * sc, v[3] * sc); a.z = pack2(v[4] * sc, v[5] * sc); a.w = pack2(v[6] * sc, v[7] * sc);
  b.x = pack2(v[8] * sc, v[9] * sc); b.y = pack2(v[10] * sc, v[11] * sc); b.z = pack2(v[12] * sc, v[13] * sc); b.w = pack2(v[14] * sc, v[15] * sc);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}


// issue-only TMEM loads (32 / 16 columns of this thread's lane): several in flight, then ONE tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc5
}  // namespace stgcma
