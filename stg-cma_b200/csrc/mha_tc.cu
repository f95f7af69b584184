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
#include <stdlib.h>
#include "tc5_attn.cuh"

namespace stgcma {
namespace mhatc {

using namespace tc5;

constexpr int kThreads = 160;

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
  // heads fastest: the CTAs of one sequence run at the same time on neighbouring SMs and together read whole qkv rows
  const int64_t s = blockIdx.y;
  const int h = blockIdx.x;
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
          if (c0 + 32 <= seq_len) {       // only the last chunk(s) hold masked columns
#pragma unroll
            for (int k = 0; k < 32; k += 2) m = fmaxf(m, fmaxf(v[k], v[k + 1]));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) m = fmaxf(m, (c0 + k < seq_len) ? v[k] : -INFINITY);
          }
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
          if (c0 + 32 <= seq_len) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float p0 = ex2f(fmaf(v[2 * k], scale_log2, -mb)), p1 = ex2f(fmaf(v[2 * k + 1], scale_log2, -mb));
              l += p0 + p1;
              pk[k] = pack2(p0, p1);
            }
          } else {
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
  dim3 grid((unsigned)heads, (unsigned)n_seq);
  kern<<<grid, kThreads, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, seq_len, heads, KP, kLog2e / sqrtf((float)HD));
  return check_launch("mha_fwd(tcgen05)");
}


// ---------------------------------------------------------------------------------------------
// backward: one CTA = one (sequence, head), Q / K / V / dO resident in shared memory (one CTA per SM), 512 TMEM columns.
//   phase A, per 128-key tile   : S^T = K_t Q^T, dP^T = V_t dO^T                      (lanes = keys, columns = queries)
//                                 P^T = exp2(S^T c - lse_i), dS^T = P^T (dP^T - D_i)  (bf16, back into TMEM)
//                                 dV_t = P^T dO, dK_t = dS^T Q                        (A from TMEM, B = dO / Q MN-major)
//   phase B, per 128-query tile : S = Q_t K^T, dP = dO_t V^T -> dS -> dQ_t = dS K     (lanes = queries)
//   The probabilities are recomputed in both orientations (like the mma.sync kernels): the transposed products need
//   the keys on the TMEM lanes.  8 element-wise warps: warps w and w + 4 share the TMEM lanes 32 (w % 4) .. + 31 and
//   split the columns; warp 8 issues the MMAs.  D_i = <dO_i, O_i> and lse_i sit in shared memory (column broadcasts
//   in phase A, per-thread scalars in phase B).  The 1 / sqrt(hd) factor is applied to dQ / dK in the epilogue.
// TMEM map (KP <= 208 columns per score tile): S @ 0, dP @ KP; bf16 results of the first column half over the
// consumed prefix of their own fp32 tile, of the second half in the spare columns [2 KP, 2 KP + KP - h0); outputs in
// dead fp32 columns: dV @ KP/2, dK @ KP + KP/2 (phase A), dQ @ 0 (phase B).
constexpr int kBwdThreads = 288;

template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1) mha_tc_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                                   const float* __restrict__ lse, const bf16* __restrict__ dout,
                                                                   bf16* __restrict__ dqkv, int seq_len, int heads, int KP,
                                                                   float scale, float scale_log2, long long* dbg) {
  constexpr int CH = HD / 8, RG = CH * 128;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);
  // phase trace of one CTA (STGCMA_MHA_TRACE=1): clock64 stamps of element-wise warp 0
  const bool trace = dbg != nullptr && blockIdx.y == 300 && blockIdx.x == 3 && threadIdx.x == 0;
  int nstamp = 0;
#define STAMP() do { if (trace) dbg[nstamp++] = clock64(); } while (0)
  STAMP();
  const uint32_t img = (uint32_t)(KP / 8) * RG;             // bytes of one [KP x HD] operand image
  const uint32_t Ks = smem_u32(sm), Vs = Ks + img, Qs = Vs + img, Os = Qs + img;   // Os = dO; 48 slack rows follow it
  uint8_t* tail = sm + 4 * (size_t)img + 6 * RG;
  float* lse2 = reinterpret_cast<float*>(tail);             // [KP]  lse * log2(e); +inf beyond the sequence
  float* Dsm = lse2 + KP;                                   // [KP]  <dO_i, O_i>
  uint64_t* bars = reinterpret_cast<uint64_t*>(Dsm + KP);
  const uint32_t sc_full = smem_u32(bars), e_full = smem_u32(bars + 1), o_full = smem_u32(bars + 2), t_free = smem_u32(bars + 3);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  uint8_t* stage = reinterpret_cast<uint8_t*>(bars + 8);    // [2][128][kPitch]: output rows on their way to coalesced stores
  constexpr int kPitch = HD * 2 + 16;                       // == 16 (mod 32) bytes: conflict-free 16-byte row-strided stores

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // heads fastest: the CTAs of one sequence run at the same time on neighbouring SMs and together read whole qkv rows
  const int64_t s = blockIdx.y;
  const int h = blockIdx.x;
  const int D = heads * HD;
  const int64_t ldq = 3 * (int64_t)D;
  const int64_t row0 = s * seq_len;
  const bf16* qbase = qkv + row0 * ldq + h * HD;
  const bf16* dobase = dout + row0 * D + h * HD;
  const int ntiles = (seq_len + 127) >> 7;

  if (tid == 0) {
    mbar_init(sc_full, 1);
    mbar_init(e_full, 256);
    mbar_init(o_full, 1);
    mbar_init(t_free, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // D_i = <dO_i, O_i> and lse_i: 4 lanes per row, a contiguous quarter of the head dim each.  The global loads are
  // ISSUED before the bulk copies and consumed after them: issued behind the copies they queue behind 170 KB of cp.async in
  // the SM's load pipe (12k cycles in the phase trace, profiles/r02b); consumed right away they would delay the copies.
  constexpr int kDI = 3, kDC = CH / 4;                      // KP <= 216 rows / 72 row slots
  uint4 ob[kDI][kDC], db[kDI][kDC];
  float lrow[kDI];
#pragma unroll
  for (int k = 0; k < kDI; ++k) {
    const int i = (tid >> 2) + k * (kBwdThreads / 4), part = tid & 3;
    lrow[k] = INFINITY;
#pragma unroll
    for (int c = 0; c < kDC; ++c) ob[k][c] = db[k][c] = make_uint4(0, 0, 0, 0);
    if (i < seq_len) {
      const uint4* orow = reinterpret_cast<const uint4*>(o + (row0 + i) * D + h * HD) + part * kDC;
      const uint4* drow = reinterpret_cast<const uint4*>(dobase + (int64_t)i * D) + part * kDC;
#pragma unroll
      for (int c = 0; c < kDC; ++c) {
        ob[k][c] = __ldg(orow + c);
        db[k][c] = __ldg(drow + c);
      }
      if (part == 0) lrow[k] = __ldg(lse + (row0 + i) * heads + h) * kLog2e;
    }
  }
  load_rows<HD>(Ks, qbase + D, ldq, 0, KP, seq_len, tid, kBwdThreads);
  load_rows<HD>(Vs, qbase + 2 * D, ldq, 0, KP, seq_len, tid, kBwdThreads);
  load_rows<HD>(Qs, qbase, ldq, 0, KP, seq_len, tid, kBwdThreads);
  load_rows<HD>(Os, dobase, D, 0, KP + 48, seq_len, tid, kBwdThreads);     // + the slack rows (zero)
  asm volatile("cp.async.commit_group;" ::: "memory");
  STAMP();
#pragma unroll
  for (int k = 0; k < kDI; ++k) {
    const int i = (tid >> 2) + k * (kBwdThreads / 4);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < kDC; ++c) {
      float ov[8], dv[8];
      unpack8(ob[k][c], ov);
      unpack8(db[k][c], dv);
#pragma unroll
      for (int q = 0; q < 8; ++q) a = fmaf(ov[q], dv[q], a);
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if ((tid & 3) == 0 && i < KP) {
      Dsm[i] = a;
      lse2[i] = lrow[k];
    }
  }
  STAMP();
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  STAMP();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  STAMP();
  const uint32_t tmem = *tmem_slot;
  const int n16 = KP >> 4;
  const int h0 = ((n16 + 1) >> 1) << 4;                     // columns [0, h0): warps 0-3, [h0, KP): warps 4-7
  const uint32_t cS = 0, cP = (uint32_t)KP, cX = (uint32_t)(2 * KP);      // S, dP, spare columns
  const uint32_t cV = (uint32_t)(KP >> 1), cK = (uint32_t)(KP + (KP >> 1));   // dV / dK accumulators (phase A); dQ @ 0 (phase B)
  const int xw = (KP - h0) >> 1;                            // bf16 columns of the second half

  if (warp == 8) {
    // ===================== MMA issue =====================
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < 2 * ntiles; ++it) {
      const bool phaseA = it < ntiles;
      const int t = phaseA ? it : it - ntiles;
      if (it > 0) {
        mbar_wait(t_free, (uint32_t)((it - 1) & 1));
        tc_fence_after();
      }
      const uint32_t a0 = (phaseA ? Ks : Qs) + (uint32_t)t * 16 * RG, a1 = (phaseA ? Vs : Os) + (uint32_t)t * 16 * RG;
      const uint32_t b0 = phaseA ? Qs : Ks, b1 = phaseA ? Os : Vs;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        mma_bf16_elect(tmem + cS, desc_nosw(a0 + ks * 256, 128, RG), desc_nosw(b0 + ks * 256, 128, RG), idesc_s, (uint32_t)(ks != 0));
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        mma_bf16_elect(tmem + cP, desc_nosw(a1 + ks * 256, 128, RG), desc_nosw(b1 + ks * 256, 128, RG), idesc_s, (uint32_t)(ks != 0));
      mma_commit_elect(sc_full);
      mbar_wait(e_full, (uint32_t)(it & 1));
      tc_fence_after();
      for (int ks = 0; ks < n16; ++ks) {
        // A k-step = 16 contraction indices = 8 bf16 columns; second column half lives in the spare region
        const bool lo = ks * 16 < h0;
        const uint32_t off = lo ? (uint32_t)(ks * 8) : (uint32_t)((ks * 16 - h0) >> 1);
        if (phaseA) {
          mma_bf16_ts_elect(tmem + cV, tmem + (lo ? cS : cX) + off, desc_nosw(Os + ks * 2 * RG, RG, 128), idesc_o, (uint32_t)(ks != 0));
          mma_bf16_ts_elect(tmem + cK, tmem + (lo ? cP : cX + xw) + off, desc_nosw(Qs + ks * 2 * RG, RG, 128), idesc_o, (uint32_t)(ks != 0));
        } else {
          mma_bf16_ts_elect(tmem + cS, tmem + (lo ? cP : cX) + off, desc_nosw(Ks + ks * 2 * RG, RG, 128), idesc_o, (uint32_t)(ks != 0));
        }
      }
      mma_commit_elect(o_full);
    }
  } else {
    // ===================== element-wise + epilogue warps =====================
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    const int c_lo = half ? h0 : 0, c_hi = half ? KP : h0;
    for (int it = 0; it < 2 * ntiles; ++it) {
      const bool phaseA = it < ntiles;
      const int t = phaseA ? it : it - ntiles;
      const int r = t * 128 + quad * 32 + lane;               // key (phase A) / query (phase B) row of this thread
      const bool warp_live = t * 128 + quad * 32 < seq_len;
      mbar_wait_relaxed(sc_full, (uint32_t)(it & 1));
      tc_fence_after();
      STAMP();
      if (warp_live) {
        const float my_l = lse2[r < KP ? r : 0], my_D = Dsm[r < KP ? r : 0];
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
          float sv[16], dp[16];
          tmem_ld16x2(trow + cS + c0, trow + cP + c0, sv, dp);
          uint32_t pk[8], dk[8];
          if (phaseA) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 l4 = *reinterpret_cast<const float4*>(lse2 + c0 + 4 * q4);
              const float4 d4 = *reinterpret_cast<const float4*>(Dsm + c0 + 4 * q4);
              const float p0 = ex2f(fmaf(sv[4 * q4], scale_log2, -l4.x)), p1 = ex2f(fmaf(sv[4 * q4 + 1], scale_log2, -l4.y));
              const float p2 = ex2f(fmaf(sv[4 * q4 + 2], scale_log2, -l4.z)), p3 = ex2f(fmaf(sv[4 * q4 + 3], scale_log2, -l4.w));
              pk[2 * q4] = pack2(p0, p1);
              pk[2 * q4 + 1] = pack2(p2, p3);
              dk[2 * q4] = pack2(p0 * (dp[4 * q4] - d4.x), p1 * (dp[4 * q4 + 1] - d4.y));
              dk[2 * q4 + 1] = pack2(p2 * (dp[4 * q4 + 2] - d4.z), p3 * (dp[4 * q4 + 3] - d4.w));
            }
            const uint32_t off = half ? (uint32_t)((c0 - h0) >> 1) : (uint32_t)(c0 >> 1);
            tmem_st8u(trow + (half ? cX : cS) + off, pk);
            tmem_st8u(trow + (half ? cX + xw : cP) + off, dk);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float p0 = ex2f(fmaf(sv[2 * k], scale_log2, -my_l)), p1 = ex2f(fmaf(sv[2 * k + 1], scale_log2, -my_l));
              dk[k] = pack2(p0 * (dp[2 * k] - my_D), p1 * (dp[2 * k + 1] - my_D));
            }
            const uint32_t off = half ? (uint32_t)((c0 - h0) >> 1) : (uint32_t)(c0 >> 1);
            tmem_st8u(trow + (half ? cX : cP) + off, dk);
          }
        }
        tmem_wait_st();
      }
      tc_fence_before();
      STAMP();
      mbar_arrive(e_full);
      mbar_wait_relaxed(o_full, (uint32_t)(it & 1));
      tc_fence_after();
      STAMP();
      // TMEM -> registers -> bf16 rows in the staging tile; the accumulators are free as soon as they are read
      {
        const int rl = quad * 32 + lane;
        if (warp_live) {
          if (phaseA) {         // warps 0-3: dV rows, warps 4-7: dK rows (scaled)
            uint8_t* dstrow = stage + ((size_t)half * 128 + rl) * kPitch;
            const uint32_t src = trow + (half ? cK : cV);
            const float sc = half ? scale : 1.0f;
#pragma unroll
            for (int c0 = 0; c0 < HD; c0 += 16) {
              float v[16];
              tmem_ld16u(src + c0, v);
              store16(reinterpret_cast<bf16*>(dstrow) + c0, v, sc);
            }
          } else {              // dQ rows: warps 0-3 the first half of the head dim, warps 4-7 the second
            uint8_t* dstrow = stage + (size_t)rl * kPitch;
#pragma unroll
            for (int c0 = 0; c0 < HD / 2; c0 += 16) {
              float v[16];
              const int c = half * (HD / 2) + c0;
              tmem_ld16u(trow + cS + c, v);
              store16(reinterpret_cast<bf16*>(dstrow) + c, v, scale);
            }
          }
        }
        tc_fence_before();
        STAMP();
        mbar_arrive(t_free);
        asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 element-wise warps: staging tile complete
        // coalesced copy-out: 16 bytes per lane, HD / 8 lanes per row
        const int nmat = phaseA ? 2 : 1;
        const int et = tid;                                     // 0 .. 255
        for (int e = et; e < nmat * 128 * CH; e += 256) {
          const int m = e / (128 * CH), rem = e - m * 128 * CH;
          const int rr = rem / CH, c = rem - rr * CH;
          const int gr = t * 128 + rr;
          if (gr < seq_len) {
            const uint4 u = *reinterpret_cast<const uint4*>(stage + ((size_t)m * 128 + rr) * kPitch + c * 16);
            const int col = phaseA ? (m ? D : 2 * D) : 0;
            *reinterpret_cast<uint4*>(dqkv + (row0 + gr) * ldq + col + h * HD + c * 8) = u;
          }
        }
      }
    }
  }
#undef STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int HD>
static int bwd_launch(const void* qkv, const void* o, const float* lse, const void* dout, void* dqkv, int64_t n_seq, int seq_len,
                      int heads, cudaStream_t st) {
  constexpr int RG = HD / 8 * 128;
  const int KP = (seq_len + 15) / 16 * 16;
  const size_t smem = 128 + (size_t)(4 * (KP / 8) + 6) * RG + 2 * KP * sizeof(float) + 64 + 2 * 128 * (HD * 2 + 16);
  // the dV / dK accumulators live in the dead upper halves of the two fp32 score tiles: needs HD <= KP / 2
  if (smem > 227 * 1024 || HD % 32 || HD > KP / 2) return -1;
  auto kern = mha_tc_bwd_kernel<HD>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  if (attr_err != cudaSuccess) {
    set_error("mha_bwd(tcgen05): smem attr: %s", cudaGetErrorString(attr_err));
    return STGCMA_ECUDA;
  }
  dim3 grid((unsigned)heads, (unsigned)n_seq);
  const float scale = 1.0f / sqrtf((float)HD);
  long long* dbg = nullptr;
  const bool trace = getenv("STGCMA_MHA_TRACE") != nullptr;
  if (trace) cudaMalloc(&dbg, 64 * sizeof(long long));
  kern<<<grid, kBwdThreads, smem, st>>>((const bf16*)qkv, (const bf16*)o, lse, (const bf16*)dout, (bf16*)dqkv, seq_len, heads, KP,
                                        scale, scale * kLog2e, dbg);
  if (trace) {   // debugging aid: phase stamps (cycles since the CTA started) of element-wise warp 0 of CTA (300, 3)
    long long h_[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(h_, dbg, sizeof(h_), cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    fprintf(stderr, "mha_tc_bwd trace:");
    for (int i = 1; i < 5 + 4 * 2 * ((seq_len + 127) / 128); ++i) fprintf(stderr, " %lld", h_[i] - h_[0]);
    fprintf(stderr, "\n");
  }
  return check_launch("mha_bwd(tcgen05)");
}

}  // namespace mhatc

// -1: shape not covered by the tcgen05 kernel (the caller uses the mma.sync kernels of mha.cu)
int mha_tc_fwd_bf16(const void* qkv, void* out, float* lse, int64_t n_seq, int seq_len, int heads, int hd, cudaStream_t st) {
  if (seq_len <= 64 || seq_len > 208 || n_seq >= 65536) return -1;
  switch (hd) {
    case 32: return mhatc::fwd_launch<32>(qkv, out, lse, n_seq, seq_len, heads, st);
    case 64: return mhatc::fwd_launch<64>(qkv, out, lse, n_seq, seq_len, heads, st);
    case 96: return mhatc::fwd_launch<96>(qkv, out, lse, n_seq, seq_len, heads, st);
    case 128: return mhatc::fwd_launch<128>(qkv, out, lse, n_seq, seq_len, heads, st);
    default: return -1;
  }
}

int mha_tc_bwd_bf16(const void* qkv, const void* o, const float* lse, const void* dout, void* dqkv, int64_t n_seq, int seq_len,
                    int heads, int hd, cudaStream_t st) {
  if (seq_len <= 64 || seq_len > 208 || n_seq >= 65536) return -1;
  switch (hd) {
    case 32: return mhatc::bwd_launch<32>(qkv, o, lse, dout, dqkv, n_seq, seq_len, heads, st);
    case 64: return mhatc::bwd_launch<64>(qkv, o, lse, dout, dqkv, n_seq, seq_len, heads, st);
    case 96: return mhatc::bwd_launch<96>(qkv, o, lse, dout, dqkv, n_seq, seq_len, heads, st);
    case 128: return mhatc::bwd_launch<128>(qkv, o, lse, dout, dqkv, n_seq, seq_len, heads, st);
    default: return -1;
  }
}

}  // namespace stgcma
