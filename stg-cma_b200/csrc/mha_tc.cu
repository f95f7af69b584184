// Multi-head attention core on the 5th-generation tensor cores (tcgen05 / TMEM): the spatial attention of the CLIP-ViT
// blocks (nn.MultiheadAttention over the 197 tokens of a video frame, CLIP_AVE.py:385 / :106-108), forward.
//
//   one CTA = one (sequence, head); sequences of 65 .. 208 tokens, head_dim in {32, 64, 96, 128}
//   shared memory : Q tile (128 rows), K, V (all keys) in the canonical NO-SWIZZLE core-matrix layout -- 8 rows x 16 bytes
//                   per core matrix, the core matrices of an 8-row group contiguous along the head dim.  Filled with
//                   cp.async straight from the strided qkv rows (no TMA: the rows of a sequence are a gather).  The same
//                   K/V image serves as a K-major operand (Q K^T: contraction over the head dim) and as an MN-major operand
//                   (P V: contraction over the keys) -- only the descriptor's LBO / SBO and the major bit differ.
//   TMEM (256 col): S = Q K^T  fp32 [128 lanes x KP columns]  (tcgen05.mma, A and B from shared memory)
//                   P = exp2(S - max) bf16, written back over S by the softmax threads (tcgen05.st, one row per thread:
//                       no shuffles, no ldmatrix) and consumed as the A operand of the second MMA straight from TMEM
//                   O = P V     fp32 [128 x HD] in the dead upper half of S
//   warps 0-3     : softmax + epilogue (thread = query row = TMEM lane); warp 4: cp.async of the next Q tile + MMA issue
//   Two CTAs per SM (104 KB shared memory, 256 TMEM columns each): while one CTA's softmax occupies the MUFU / FMA pipes
//   the other CTA's MMAs occupy the tensor pipe.
// Shorter / longer sequences and the other head dims run on the mma.sync kernels of mha.cu.
#include "tc_common.cuh"

namespace stgcma {
namespace mhatc {

using namespace tc;

constexpr int kThreads = 160;
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
// Eight consecutive threads fill one 128-byte core matrix (conflict-free), consecutive groups of 8 walk the head dim.
template <int HD>
__device__ __forceinline__ void load_rows(uint32_t dst, const bf16* __restrict__ src, int64_t ld, int r0, int nrows, int r_valid,
                                          int tid, int nthr) {
  constexpr int CH = HD / 8, RG = CH * 128;
  for (int e = tid; e < nrows * CH; e += nthr) {
    const int r8 = e & 7, kc = (e >> 3) % CH, rg = (e >> 3) / CH;
    const int r = rg * 8 + r8;
    const uint32_t d = dst + rg * RG + kc * 128 + r8 * 16;
    if (r0 + r < r_valid) cp16(d, src + (int64_t)(r0 + r) * ld + kc * 8);
    else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0u) : "memory");
  }
}

template <int HD>
__global__ void __launch_bounds__(kThreads, 2) mha_tc_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                float* __restrict__ lse, int seq_len, int heads, int KP,
                                                                float scale_log2) {
  constexpr int CH = HD / 8, RG = CH * 128;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);
  const uint32_t Qs = smem_u32(sm);                       // 128 rows
  const uint32_t Ks = Qs + 16 * RG;                       // KP rows
  const uint32_t Vs = Ks + (KP / 8) * RG;                 // KP rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 16 * RG + 2 * (KP / 8) * RG);
  const uint32_t s_full = smem_u32(bars), p_full = smem_u32(bars + 1), o_full = smem_u32(bars + 2), o_free = smem_u32(bars + 3);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t s = blockIdx.x;
  const int h = blockIdx.y;
  const int D = heads * HD;
  const int64_t ldq = 3 * (int64_t)D;
  const bf16* qbase = qkv + s * seq_len * ldq + h * HD;
  const int ntiles = (seq_len + 127) >> 7;

  if (tid == 0) {
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  load_rows<HD>(Ks, qbase + D, ldq, 0, KP, seq_len, tid, kThreads);
  load_rows<HD>(Vs, qbase + 2 * D, ldq, 0, KP, seq_len, tid, kThreads);
  load_rows<HD>(Qs, qbase, ldq, 0, 128, seq_len, tid, kThreads);
  cp_commit_wait();
  fence_async_smem();          // generic-proxy writes (cp.async, st.shared) -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t o_col = (uint32_t)(KP >> 1);             // O lives right after P (bf16: KP / 2 columns) inside the S region

  if (warp == 4) {
    // ===================== cp.async of the next Q tile + MMA issue (warp-uniform, one elected lane issues) =====================
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int t = 0; t < ntiles; ++t) {
      if (t > 0) {
        mbar_wait(o_free, (uint32_t)((t - 1) & 1));       // the epilogue has drained O of the previous tile: TMEM is free
        tc_fence_after();
      }
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)                // S = Q K^T: both operands K-major (LBO 128 B, SBO one row group)
        mma_bf16_elect(tmem, desc_nosw(Qs + ks * 256, 128, RG), desc_nosw(Ks + ks * 256, 128, RG), idesc_s, (uint32_t)(ks != 0));
      mma_commit_elect(s_full);
      mbar_wait(s_full, (uint32_t)(t & 1));               // (own commit) the MMAs have read Qs: it can be refilled
      if (t + 1 < ntiles) {
        load_rows<HD>(Qs, qbase, ldq, (t + 1) * 128, 128, seq_len, lane, 32);
        cp_commit_wait();
        fence_async_smem();
        __syncwarp();
      }
      mbar_wait(p_full, (uint32_t)(t & 1));               // every softmax thread has stored its P row
      tc_fence_after();
      for (int ks = 0; ks < KP / 16; ++ks)                // O = P V: A from TMEM (8 columns per k-step), B = V MN-major
        mma_bf16_ts_elect(tmem + o_col, tmem + ks * 8, desc_nosw(Vs + ks * 2 * RG, RG, 128), idesc_o, (uint32_t)(ks != 0));
      mma_commit_elect(o_full);
    }
  } else {
    // ===================== softmax + epilogue: thread = query row = TMEM lane =====================
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int t = 0; t < ntiles; ++t) {
      const int i = t * 128 + warp * 32 + lane;           // query row of this thread
      const bool warp_live = t * 128 + warp * 32 < seq_len;
      mbar_wait_relaxed(s_full, (uint32_t)(t & 1));
      tc_fence_after();
      float m = -INFINITY, l = 0.f;
      if (warp_live) {
        // pass 1: row maximum over the valid keys
        for (int c0 = 0; c0 < KP; c0 += 32) {
          float v[32];
          if (c0 + 32 <= KP) tmem_ld32(trow + c0, v);
          else {
            tmem_ld16(trow + c0, v);
#pragma unroll
            for (int k = 16; k < 32; ++k) v[k] = -INFINITY;
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) m = fmaxf(m, (c0 + k < seq_len) ? v[k] : -INFINITY);
        }
        const float mb = m * scale_log2;
        // pass 2: P = exp2(S c - m c) as bf16 pairs, written over the low half of S (chunk c overwrites columns that
        // chunk <= c has already consumed)
        for (int c0 = 0; c0 < KP; c0 += 32) {
          float v[32];
          const bool full = c0 + 32 <= KP;
          if (full) tmem_ld32(trow + c0, v);
          else tmem_ld16(trow + c0, v);
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float p0 = 0.f, p1 = 0.f;
            if (full || k < 8) {
              p0 = (c0 + 2 * k < seq_len) ? ex2f(fmaf(v[2 * k], scale_log2, -mb)) : 0.f;
              p1 = (c0 + 2 * k + 1 < seq_len) ? ex2f(fmaf(v[2 * k + 1], scale_log2, -mb)) : 0.f;
            }
            l += p0 + p1;
            pk[k] = pack2(p0, p1);
          }
          if (full) tmem_st16(trow + (c0 >> 1), pk);
          else tmem_st8(trow + (c0 >> 1), pk);
        }
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait_relaxed(o_full, (uint32_t)(t & 1));
      tc_fence_after();
      if (warp_live) {       // warp-uniform: every lane takes part in the .sync.aligned TMEM loads, valid rows store
        const bool ok = i < seq_len;
        const float inv = 1.f / l;
        bf16* orow = out + (s * seq_len + (ok ? i : 0)) * D + h * HD;
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
          float v[32];
          tmem_ld32(trow + o_col + c0, v);
          if (ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = pack2(v[8 * q] * inv, v[8 * q + 1] * inv);
              u.y = pack2(v[8 * q + 2] * inv, v[8 * q + 3] * inv);
              u.z = pack2(v[8 * q + 4] * inv, v[8 * q + 5] * inv);
              u.w = pack2(v[8 * q + 6] * inv, v[8 * q + 7] * inv);
              *reinterpret_cast<uint4*>(orow + c0 + 8 * q) = u;
            }
          }
        }
        if (ok) lse[(s * seq_len + i) * heads + h] = (m * scale_log2 + log2f(l)) * 0.6931471805599453f;
      }
      tc_fence_before();
      mbar_arrive(o_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <int HD>
static int fwd_launch(const void* qkv, void* out, float* lse, int64_t n_seq, int seq_len, int heads, cudaStream_t st) {
  constexpr int RG = HD / 8 * 128;
  const int KP = (seq_len + 15) / 16 * 16;
  const size_t smem = 128 + (size_t)(16 + 2 * (KP / 8)) * RG + 64;
  if (smem > 113 * 1024) return -1;
  auto kern = mha_tc_fwd_kernel<HD>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024); });
  if (attr_err != cudaSuccess) {
    set_error("mha_fwd(tcgen05): smem attr: %s", cudaGetErrorString(attr_err));
    return STGCMA_ECUDA;
  }
  dim3 grid((unsigned)n_seq, (unsigned)heads);
  kern<<<grid, kThreads, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, seq_len, heads, KP, kLog2e / sqrtf((float)HD));
  return check_launch("mha_fwd(tcgen05)");
}

}  // namespace mhatc

// -1: shape not covered by the tcgen05 kernel (the caller uses the mma.sync kernels of mha.cu)
int mha_tc_fwd_bf16(const void* qkv, void* out, float* lse, int64_t n_seq, int seq_len, int heads, int hd, cudaStream_t st) {
  if (seq_len <= 64 || seq_len > 208 || n_seq >= (1ll << 31)) return -1;
  switch (hd) {
    case 32: return mhatc::fwd_launch<32>(qkv, out, lse, n_seq, seq_len, heads, st);
    case 64: return mhatc::fwd_launch<64>(qkv, out, lse, n_seq, seq_len, heads, st);
    case 96: return mhatc::fwd_launch<96>(qkv, out, lse, n_seq, seq_len, heads, st);
    case 128: return mhatc::fwd_launch<128>(qkv, out, lse, n_seq, seq_len, heads, st);
    default: return -1;
  }
}

}  // namespace stgcma
