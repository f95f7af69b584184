// Swin window attention (WindowAttention.forward, temporal=False; Swin_AVE.py:238-276), bf16 production
// kernels on the warp-level tensor cores (mma.cuh).  roll / window_partition / window_reverse are addressing:
// a CTA gathers the ws*ws (<= 49) token rows of one (frame, window) for a group of HG heads straight from the
// token-major qkv tensor with cp.async, keeps q / k / v (and dO in the backward) in shared memory, and each of
// its 4 warps owns one 16-row block of the (padded to 64) window for every head of the group.  The relative
// position bias is a shared-memory table lookup row_offset + column_offset (pad columns point at a -inf run, so
// there is no validity test per score); the 0 / -100 shift mask is a per-lane bit mask computed once per window and
// applied only in the windows that touch the shifted border.
//   forward : S = scale q k^T + bias + mask -> softmax -> P v, all in registers (14 + 16 MMAs per 16 x 49 block)
//   backward: phase A (warp = query rows)  P, dP = dO v^T, delta = rowsum(P dP), dS, dq = scale dS k; P and dS -> smem
//             phase B (warp = key rows)    dk = scale dS^T q, dv = P^T dO from the ldmatrix.trans'ed P / dS blocks
//   (round 1 recomputed the scores transposed in phase B: 2050 instructions per warp and head, issue bound at 0.23 of
//    the HBM roof -- ncu profiles/r01f_ncu_win_bwd0.txt; this version issues about a quarter of that)
#include <stdlib.h>
#include "mma.cuh"

namespace stgcma {
namespace wtc {

using namespace mma;

constexpr int HD = 32;
constexpr int NP = 64;       // padded window length
constexpr int NTH = 128;     // 4 warps
constexpr float kScale = 0.17677669529663689f;  // 32^-0.5

// Tiles hold N + 1 rows (N = ws * ws <= 49 tokens and one ZERO row): every fragment load clamps its row index to the zero
// row instead of reading 64 padded rows -- 4 x 50 instead of 4 x 64 tile rows are what lets a fourth backward CTA (a sixth
// forward CTA) fit an SM's shared memory.
__device__ __forceinline__ void load_a_c(uint32_t (&a)[4], const uint8_t* tile, int ld, int row0, int k0, int lane, int rmax) {
  int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  r = r < rmax ? r : rmax;
  const int k = k0 + (lane >> 4) * 8;
  ldsm_x4(a, smem_addr(tile + r * ld + k * 2));
}
__device__ __forceinline__ void load_b_nk_c(uint32_t (&b)[4], const uint8_t* tile, int ld, int n0, int k0, int lane, int rmax) {
  int n = n0 + (lane & 7) + (lane >> 4) * 8;
  n = n < rmax ? n : rmax;
  const int k = k0 + ((lane >> 3) & 1) * 8;
  ldsm_x4(b, smem_addr(tile + n * ld + k * 2));
}
__device__ __forceinline__ void load_b_kn_c(uint32_t (&b)[4], const uint8_t* tile, int ld, int k0, int n0, int lane, int rmax) {
  int k = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  k = k < rmax ? k : rmax;
  const int n = n0 + (lane >> 4) * 8;
  ldsm_x4_t(b, smem_addr(tile + k * ld + n * 2));
}

struct Win {
  int16_t tok[NP];           // token row (within the frame) of window slot i
  int16_t rowoff[NP];        // yi * (2ws-1) + xi                         (0 for pad slots)
  int16_t reg[NP];           // shift-mask region id
};

// Shared-memory bias table of one head: [0, T2v) = rpb * log2(e), [T2v, T2v + K0] = -inf, so that the lookup
// row_base + column_offset of a PAD column (offset T2v) lands on -inf for every row (row offsets are 0..K0):
// no per-score validity test.  TS = padded stride.
__host__ __device__ __forceinline__ int tbl_stride(int ws) {
  return ((2 * ws - 1) * (2 * ws - 1) + (ws - 1) * (2 * ws - 1) + ws + 3) & ~3;
}

__device__ __forceinline__ void win_setup(Win& w, int wy, int wx, int ws, int shift, int R, int N, int tid) {
  if (tid < NP) {
    const int i = tid;
    if (i < N) {
      const int iy = i / ws, ix = i - iy * ws;
      w.tok[i] = (int16_t)(win_src(wy, iy, ws, shift, R) * R + win_src(wx, ix, ws, shift, R));
      w.rowoff[i] = (int16_t)(iy * (2 * ws - 1) + ix);
      w.reg[i] = (int16_t)(shift > 0 ? 3 * win_region1(wy * ws + iy, ws, shift, R) + win_region1(wx * ws + ix, ws, shift, R) : 0);
    } else {
      w.tok[i] = 0; w.rowoff[i] = 0; w.reg[i] = 0;
    }
  }
}

template <int HG>
__device__ __forceinline__ void load_tbl(float* tbl, const float* __restrict__ rpb, int TS, int T2v, int heads, int h0, int tid) {
  // TS <= 2 * NTH for ws <= 7: every thread issues its 2 * HG independent loads before the first store
  float v[HG][2];
#pragma unroll
  for (int hh = 0; hh < HG; ++hh)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int i = tid + p * NTH;
      v[hh][p] = i < T2v ? __ldg(rpb + i * heads + h0 + hh) : -INFINITY;
    }
#pragma unroll
  for (int hh = 0; hh < HG; ++hh)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int i = tid + p * NTH;
      if (i < TS) tbl[hh * TS + i] = v[hh][p] * kLog2e;
    }
}

// Row bookkeeping of the gather / scatter loops.  Thread layout: CH = HG*4 consecutive threads move the CH 16-byte chunks
// of one row, RPP = NTH / CH rows per pass; the token index of each of the thread's rows is looked up ONCE per window
// (it is the same for q, k, v, dO and the three outputs) -- these loops were 23 % of the kernel's instructions.
template <int HG>
struct RowMap {
  static constexpr int CH = HG * 4, LD = HG * 64 + 16, RPP = NTH / CH, PASSES = (NP + RPP - 1) / RPP;
  int tok[PASSES];   // token row of pass p, -1: the zero row N (zero fill), -2: beyond the tile / idle thread
  int r0, c;
  __device__ __forceinline__ void init(const Win& w, int N, int tid) {
    r0 = tid / CH;
    c = tid - r0 * CH;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int r = r0 + p * RPP;
      tok[p] = (r0 >= RPP || r > N) ? -2 : (r < N ? (int)w.tok[r] : -1);   // row N: the zero row
    }
  }
  // [N rows][HG*32] slice (column offset col0 of rows of `ld_g` elements) -> [N + 1][LD] tile; row N zero
  __device__ __forceinline__ void gather(uint8_t* tile, const bf16* __restrict__ src, int64_t frame_row0, int ld_g, int col0) const {
    const bf16* base = src + frame_row0 * ld_g + col0 + c * 8;
    uint8_t* dst = tile + r0 * LD + c * 16;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      if (tok[p] >= 0) cp_async16(dst + p * RPP * LD, base + tok[p] * ld_g);
      else if (tok[p] == -1) *reinterpret_cast<uint4*>(dst + p * RPP * LD) = make_uint4(0, 0, 0, 0);
    }
  }
  // coalesced scatter of a staged tile to the token-major destination (a row's HG*64 B segment is contiguous)
  __device__ __forceinline__ void scatter(const uint8_t* tile, bf16* __restrict__ dst, int64_t frame_row0, int ld_g, int col0) const {
    bf16* base = dst + frame_row0 * ld_g + col0 + c * 8;
    const uint8_t* sp = tile + r0 * LD + c * 16;
#pragma unroll
    for (int p = 0; p < PASSES; ++p)
      if (tok[p] >= 0) *reinterpret_cast<uint4*>(base + tok[p] * ld_g) = *reinterpret_cast<const uint4*>(sp + p * RPP * LD);
  }
};

// Per-lane metadata of the 14 score columns a lane owns (c = i*8 + 2t + e), head independent, hoisted out of the head
// loop: ocb = BYTE offset of the column's bias-table term (coloff(c) = K0 - rowoff(c); T2v for pad columns), and one mask
// bit per (row, column): set where the shift-mask regions differ (the reference's -100).
__device__ __forceinline__ void col_meta(int (&ocb)[14], uint32_t& ma, uint32_t& mb, const Win& w, int N, int t, int K0, int T2v,
                                         int ra, int rb, bool masked) {
  ma = 0u; mb = 0u;
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = i * 8 + 2 * t + e, k = 2 * i + e;
      ocb[k] = c < N ? (K0 - (int)w.rowoff[c]) * 4 : T2v * 4;
    }
  if (masked) {
    const int ga = w.reg[ra < N ? ra : 0], gb = w.reg[rb < N ? rb : 0];
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = i * 8 + 2 * t + e, k = 2 * i + e;
        if (c < N) {
          const int gc = w.reg[c];
          ma |= (uint32_t)(ga != gc) << k;
          mb |= (uint32_t)(gb != gc) << k;
        }
      }
  }
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

constexpr float kScaleLog2e = kScale * kLog2e;
constexpr float kMaskLog2e = 100.0f * kLog2e;

// s (7 n-tiles of a 16 x 56 block, raw q.k accumulators) -> log2-domain logits with bias (and mask) applied:
// 2 address adds + 2 loads + 2 FMAs per column pair.  pa / pb: shared-memory byte address of tbl[head] + rowoff(row).
template <bool MASKED>
__device__ __forceinline__ void bias_mask(float (&s)[7][4], uint32_t pa, uint32_t pb, const int (&ocb)[14], uint32_t ma,
                                          uint32_t mb) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 2 * i + e;
      float va = fmaf(s[i][e], kScaleLog2e, lds_f32(pa + ocb[k]));
      float vb = fmaf(s[i][2 + e], kScaleLog2e, lds_f32(pb + ocb[k]));
      if (MASKED) {
        if (ma & (1u << k)) va -= kMaskLog2e;
        if (mb & (1u << k)) vb -= kMaskLog2e;
      }
      s[i][e] = va;
      s[i][2 + e] = vb;
    }
  }
}

// write this lane's part of a 16 x 32 fp32 accumulator block (rows ra / rb of the tile, head h) as bf16 into a
// [NP][LD] shared tile (the warp's own rows: conflict-free, see the LD padding)
__device__ __forceinline__ void stage_rows(uint8_t* tile, int LD, const float (&o)[4][4], int ra, int rb, int h, int t, int N) {
  uint8_t* pa = tile + ra * LD + (h * HD + 2 * t) * 2;
  uint8_t* pb = tile + rb * LD + (h * HD + 2 * t) * 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (ra < N) *reinterpret_cast<uint32_t*>(pa + i * 16) = pack_bf16(o[i][0], o[i][1]);
    if (rb < N) *reinterpret_cast<uint32_t*>(pb + i * 16) = pack_bf16(o[i][2], o[i][3]);
  }
}
__device__ __forceinline__ void stage_rows(uint8_t* tile, int LD, const float (&o)[4][4], float sa, float sb, int ra, int rb,
                                           int h, int t, int N) {
  uint8_t* pa = tile + ra * LD + (h * HD + 2 * t) * 2;
  uint8_t* pb = tile + rb * LD + (h * HD + 2 * t) * 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (ra < N) *reinterpret_cast<uint32_t*>(pa + i * 16) = pack_bf16(o[i][0] * sa, o[i][1] * sa);
    if (rb < N) *reinterpret_cast<uint32_t*>(pb + i * 16) = pack_bf16(o[i][2] * sb, o[i][3] * sb);
  }
}

// 16 x 56 raw scores: A = 16 rows of `ta` (row block rb), B = rows of `tb` as [n][k]; features h*32 .. +31
__device__ __forceinline__ void scores(float (&s)[7][4], const uint8_t* ta, const uint8_t* tb, int LD, int rb, int h, int lane,
                                       int N) {
#pragma unroll
  for (int i = 0; i < 7; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
  uint32_t a[2][4];
  load_a_c(a[0], ta, LD, rb * 16, h * HD, lane, N);
  load_a_c(a[1], ta, LD, rb * 16, h * HD + 16, lane, N);
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t b[4];
      load_b_nk_c(b, tb, LD, n2 * 16, h * HD + ks * 16, lane, N);
      mma16816(s[2 * n2], a[ks], b[0], b[1]);
      if (n2 < 3) mma16816(s[2 * n2 + 1], a[ks], b[2], b[3]);
    }
  }
}

// bf16 pairs of a 16 x 56 fp32 block in the accumulator layout: pk[2i] = (row g: cols 8i+2t, +1), pk[2i+1] = row g+8.
// pk[4kk .. 4kk+3] is the A fragment of k-step kk of the block used as a left operand (pk[14], pk[15] = 0).
__device__ __forceinline__ void pack_block(uint32_t (&pk)[16], const float (&p)[7][4]) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    pk[2 * i] = pack_bf16(p[i][0], p[i][1]);
    pk[2 * i + 1] = pack_bf16(p[i][2], p[i][3]);
  }
  pk[14] = 0u;
  pk[15] = 0u;
}

// o[4 n-tiles of 32 features] += P (16 x 64 as packed A fragments) . tile[k][n]
__device__ __forceinline__ void pv(float (&o)[4][4], const uint32_t (&pk)[16], const uint8_t* tv, int LD, int h, int lane, int N) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t pa[4] = {pk[4 * kk], pk[4 * kk + 1], pk[4 * kk + 2], pk[4 * kk + 3]};
#pragma unroll
    for (int n2 = 0; n2 < 2; ++n2) {
      uint32_t b[4];
      load_b_kn_c(b, tv, LD, kk * 16, h * HD + n2 * 16, lane, N);
      mma16816(o[2 * n2], pa, b[0], b[1]);
      mma16816(o[2 * n2 + 1], pa, b[2], b[3]);
    }
  }
}

// the packed 16 x 56 block of row block `rb` -> a [N + 2][LDP] bf16 shared tile (stmatrix: the accumulator layout IS the
// 8 x 8 matrix fragment layout).  Rows <= N are kept (row N: a pad query, finite, contributes nothing because q and dO are
// zero there); the other pad rows of the last row block all land on the dump row N + 1.  The tile holds 56 key columns; phase
// B's key block 48..63 reads columns 56..63 from the next row (finite values of the same tile): they only feed discarded rows.
constexpr int LDP = 112;   // 56 keys, 7 x 16 B: an odd chunk count keeps ldmatrix / stmatrix conflict-free
__device__ __forceinline__ void store_block(uint8_t* tile, const uint32_t (&pk)[16], int rb, int lane, int N) {
  const int j = lane >> 3;
  int r = rb * 16 + (j & 1) * 8 + (lane & 7);
  r = r <= N ? r : N + 1;
  const uint32_t base = smem_addr(tile + r * LDP + (j >> 1) * 16);
#pragma unroll
  for (int p = 0; p < 3; ++p)
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(base + p * 32), "r"(pk[4 * p]),
                 "r"(pk[4 * p + 1]), "r"(pk[4 * p + 2]), "r"(pk[4 * p + 3]) : "memory");
  asm volatile("stmatrix.sync.aligned.m8n8.x2.shared.b16 [%0], {%1,%2};" ::"r"(base + 96), "r"(pk[12]), "r"(pk[13]) : "memory");
}

// A fragment (16 m x 16 k) of X^T where X is stored [k][m] row-major (row stride ld bytes); k rows clamped to rmax
__device__ __forceinline__ void load_a_t(uint32_t (&a)[4], const uint8_t* tile, int ld, int m0, int k0, int lane, int rmax) {
  int k = k0 + (lane & 7) + (lane >> 4) * 8;
  k = k < rmax ? k : rmax;
  const int m = m0 + ((lane >> 3) & 1) * 8;
  ldsm_x4_t(a, smem_addr(tile + k * ld + m * 2));
}

// o[4 n-tiles of 32 features] += X^T (rows = columns rb*16.. of the [N + 2][LDP] tile sx, k = its query rows) . tile[k][n]
__device__ __forceinline__ void xtv(float (&o)[4][4], const uint8_t* sx, const uint8_t* tile, int LD, int rb, int h, int lane,
                                    int N) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    load_a_t(a, sx, LDP, rb * 16, kk * 16, lane, N);
#pragma unroll
    for (int n2 = 0; n2 < 2; ++n2) {
      uint32_t b[4];
      load_b_kn_c(b, tile, LD, kk * 16, h * HD + n2 * 16, lane, N);
      mma16816(o[2 * n2], a, b[0], b[1]);
      mma16816(o[2 * n2 + 1], a, b[2], b[3]);
    }
  }
}

// log2-domain logits -> unnormalised probabilities in place; returns the row sums (rows ra / rb of the lane)
__device__ __forceinline__ void softmax_rows(float (&s)[7][4], float& l0, float& l1) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    m0 = fmaxf(m0, fmaxf(s[i][0], s[i][1]));
    m1 = fmaxf(m1, fmaxf(s[i][2], s[i][3]));
  }
  m0 = quad_max(m0);
  m1 = quad_max(m1);
  l0 = 0.f; l1 = 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    s[i][0] = ex2(s[i][0] - m0);
    s[i][1] = ex2(s[i][1] - m0);
    s[i][2] = ex2(s[i][2] - m1);
    s[i][3] = ex2(s[i][3] - m1);
    l0 += s[i][0] + s[i][1];
    l1 += s[i][2] + s[i][3];
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
}

template <int HG>
__global__ void __launch_bounds__(NTH) window_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ rpb,
                                                         bf16* __restrict__ out, int R, int ws, int shift, int heads) {
  constexpr int LD = HG * 64 + 16;
  extern __shared__ __align__(16) uint8_t wsm[];
  const int N = ws * ws, NT = N + 1;                     // tile rows: N tokens + the zero row
  uint8_t* tq = wsm;
  uint8_t* tk = tq + NT * LD;
  uint8_t* tv = tk + NT * LD;
  float* tbl = reinterpret_cast<float*>(tv + NT * LD);   // [HG][TS]
  const int TS = tbl_stride(ws), T2v = (2 * ws - 1) * (2 * ws - 1);
  Win& w = *reinterpret_cast<Win*>(tbl + HG * TS);
  const int nwr = R / ws, nW = nwr * nwr, C = heads * HD, L = R * R;
  const uint32_t ngrp = (uint32_t)(heads / HG);
  uint32_t b = blockIdx.x;                      // host checks frames * nW * ngrp < 2^31: 32-bit index math
  const int hg = (int)(b % ngrp); b /= ngrp;
  const int wi = (int)(b % (uint32_t)nW);
  const int64_t frame = b / (uint32_t)nW;
  const int wy = wi / nwr, wx = wi - wy * nwr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  win_setup(w, wy, wx, ws, shift, R, N, tid);
  __syncthreads();
  const int64_t row0 = frame * L;
  RowMap<HG> rm;
  rm.init(w, N, tid);
  rm.gather(tq, qkv, row0, 3 * C, hg * HG * HD);
  rm.gather(tk, qkv, row0, 3 * C, C + hg * HG * HD);
  rm.gather(tv, qkv, row0, 3 * C, 2 * C + hg * HG * HD);
  cp_async_commit();
  load_tbl<HG>(tbl, rpb, TS, T2v, heads, hg * HG, tid);   // overlaps the gather
  const int ra = warp * 16 + g, rb = ra + 8;
  const int K0 = (ws - 1) * (2 * ws - 1) + (ws - 1);
  const bool masked = shift > 0 && (wy == nwr - 1 || wx == nwr - 1);   // interior windows of a shifted block: one region
  int ocb[14];
  uint32_t ma, mb;
  col_meta(ocb, ma, mb, w, N, t, K0, T2v, ra, rb, masked);
  const uint32_t pa0 = smem_addr(tbl) + 4 * (int)w.rowoff[ra < N ? ra : 0], pb0 = smem_addr(tbl) + 4 * (int)w.rowoff[rb < N ? rb : 0];
  cp_async_wait<0>();
  __syncthreads();

#pragma unroll 1
  for (int h = 0; h < HG; ++h) {
    float s[7][4];
    scores(s, tq, tk, LD, warp, h, lane, N);
    if (masked) bias_mask<true>(s, pa0 + h * TS * 4, pb0 + h * TS * 4, ocb, ma, mb);
    else bias_mask<false>(s, pa0 + h * TS * 4, pb0 + h * TS * 4, ocb, ma, mb);
    float l0, l1;
    softmax_rows(s, l0, l1);
    uint32_t pk[16];
    pack_block(pk, s);
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    pv(o, pk, tv, LD, h, lane, N);
    // q rows of this warp / head are dead (only this warp read them): they stage the output
    __syncwarp();
    stage_rows(tq, LD, o, 1.f / l0, 1.f / l1, ra, rb, h, t, N);
  }
  __syncthreads();
  rm.scatter(tq, out, row0, C, hg * HG * HD);
}

// Backward.  Per head: (A) the warp's 16 query rows: P, dP = dO v^T, delta, dS, dq = dS k in registers; P and dS also go to
// shared memory (bf16, stmatrix).  (B) the warp's 16 key rows: dk = dS^T q and dv = P^T dO with the transposed blocks read
// back by ldmatrix.trans -- nothing is recomputed.  Outputs are staged in the input tiles as their head slices die:
// dq -> k, dk -> v (both dead after phase A of the head), dv -> dO (dead after phase B).
template <int HG>
__global__ void __launch_bounds__(NTH, 4) window_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ rpb,
                                                         const bf16* __restrict__ dout, bf16* __restrict__ dqkv, int R, int ws,
                                                         int shift, int heads) {
  constexpr int LD = HG * 64 + 16;
  extern __shared__ __align__(16) uint8_t wsm[];
  const int N = ws * ws, NT = N + 1;                     // tile rows: N tokens + the zero row
  uint8_t* tq = wsm;
  uint8_t* tk = tq + NT * LD;
  uint8_t* tv = tk + NT * LD;
  uint8_t* td = tv + NT * LD;
  uint8_t* sp = td + NT * LD;                            // P  [N + 2 queries][LDP]  (row N: pad query, row N + 1: dump)
  uint8_t* sd = sp + (N + 2) * LDP;                      // dS [N + 2 queries][LDP]
  float* tbl = reinterpret_cast<float*>(sd + (N + 2) * LDP);   // [HG][TS]
  const int TS = tbl_stride(ws), T2v = (2 * ws - 1) * (2 * ws - 1);
  Win& w = *reinterpret_cast<Win*>(tbl + HG * TS);
  const int nwr = R / ws, nW = nwr * nwr, C = heads * HD, L = R * R;
  const uint32_t ngrp = (uint32_t)(heads / HG);
  uint32_t b = blockIdx.x;                      // host checks frames * nW * ngrp < 2^31: 32-bit index math
  const int hg = (int)(b % ngrp); b /= ngrp;
  const int wi = (int)(b % (uint32_t)nW);
  const int64_t frame = b / (uint32_t)nW;
  const int wy = wi / nwr, wx = wi - wy * nwr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  win_setup(w, wy, wx, ws, shift, R, N, tid);
  __syncthreads();
  const int64_t row0 = frame * L;
  RowMap<HG> rm;
  rm.init(w, N, tid);
  rm.gather(tq, qkv, row0, 3 * C, hg * HG * HD);
  rm.gather(tk, qkv, row0, 3 * C, C + hg * HG * HD);
  rm.gather(tv, qkv, row0, 3 * C, 2 * C + hg * HG * HD);
  rm.gather(td, dout, row0, C, hg * HG * HD);
  cp_async_commit();
  load_tbl<HG>(tbl, rpb, TS, T2v, heads, hg * HG, tid);   // overlaps the gather
  const int ra = warp * 16 + g, rb = ra + 8;
  const int K0 = (ws - 1) * (2 * ws - 1) + (ws - 1);
  const bool masked = shift > 0 && (wy == nwr - 1 || wx == nwr - 1);
  int ocb[14];
  uint32_t ma, mb;
  col_meta(ocb, ma, mb, w, N, t, K0, T2v, ra, rb, masked);
  const uint32_t pa0 = smem_addr(tbl) + 4 * (int)w.rowoff[ra < N ? ra : 0], pb0 = smem_addr(tbl) + 4 * (int)w.rowoff[rb < N ? rb : 0];
  cp_async_wait<0>();
  __syncthreads();

#pragma unroll 1
  for (int h = 0; h < HG; ++h) {
    float oq[4][4];
    {
      // ---------------- phase A: this warp's rows are queries ----------------
      float s[7][4], dp[7][4];
      scores(s, tq, tk, LD, warp, h, lane, N);
      if (masked) bias_mask<true>(s, pa0 + h * TS * 4, pb0 + h * TS * 4, ocb, ma, mb);
      else bias_mask<false>(s, pa0 + h * TS * 4, pb0 + h * TS * 4, ocb, ma, mb);
      float l0, l1;
      softmax_rows(s, l0, l1);
      const float inv0 = 1.f / l0, inv1 = 1.f / l1;
      scores(dp, td, tv, LD, warp, h, lane, N);   // dP = dO . v^T
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        d0 += s[i][0] * dp[i][0] + s[i][1] * dp[i][1];
        d1 += s[i][2] * dp[i][2] + s[i][3] * dp[i][3];
      }
      d0 = -quad_sum(d0) * inv0 * kScale;      // dS carries the softmax scale: dq and dk need none at staging
      d1 = -quad_sum(d1) * inv1 * kScale;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        s[i][0] *= inv0; s[i][1] *= inv0; s[i][2] *= inv1; s[i][3] *= inv1;
        dp[i][0] = s[i][0] * fmaf(dp[i][0], kScale, d0); dp[i][1] = s[i][1] * fmaf(dp[i][1], kScale, d0);
        dp[i][2] = s[i][2] * fmaf(dp[i][2], kScale, d1); dp[i][3] = s[i][3] * fmaf(dp[i][3], kScale, d1);
      }
      uint32_t pk[16];
      pack_block(pk, s);
      store_block(sp, pk, warp, lane, N);
      pack_block(pk, dp);
      store_block(sd, pk, warp, lane, N);
#pragma unroll
      for (int i = 0; i < 4; ++i) oq[i][0] = oq[i][1] = oq[i][2] = oq[i][3] = 0.f;
      pv(oq, pk, tk, LD, h, lane, N);             // dq = dS . k
    }
    __syncthreads();                           // P / dS complete; k and v slices of head h are dead
    stage_rows(tk, LD, oq, ra, rb, h, t, N);
    // ---------------- phase B: this warp's rows are keys ----------------
    float ok[4][4], ov[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok[i][0] = ok[i][1] = ok[i][2] = ok[i][3] = 0.f;
      ov[i][0] = ov[i][1] = ov[i][2] = ov[i][3] = 0.f;
    }
    xtv(ok, sd, tq, LD, warp, h, lane, N);        // dk = dS^T . q
    xtv(ov, sp, td, LD, warp, h, lane, N);        // dv = P^T . dO
    stage_rows(tv, LD, ok, ra, rb, h, t, N);
    __syncthreads();                           // q and dO slices of head h are dead; P / dS may be overwritten
    stage_rows(td, LD, ov, ra, rb, h, t, N);
  }
  __syncthreads();
  rm.scatter(tk, dqkv, row0, 3 * C, hg * HG * HD);
  rm.scatter(tv, dqkv, row0, 3 * C, C + hg * HG * HD);
  rm.scatter(td, dqkv, row0, 3 * C, 2 * C + hg * HG * HD);
}

static int pick_hg(int heads) {
  static const int force = getenv("STGCMA_WIN_HG") ? atoi(getenv("STGCMA_WIN_HG")) : 0;   // tuning knob
  if (force > 0 && heads % force == 0) return force;
  if (heads % 3 == 0) return 3;   // backward: 4 tiles of 50 rows + P / dS = 55 KB -> 4 CTAs per SM
  if (heads % 2 == 0) return 2;
  return 1;
}

template <int HG>
static int fwd_launch(const void* qkv, const float* rpb, void* out, int64_t frames, int R, int ws, int shift, int heads,
                      cudaStream_t st) {
  constexpr int LD = HG * 64 + 16;
  const size_t smem = (size_t)3 * (ws * ws + 1) * LD + sizeof(float) * HG * tbl_stride(ws) + sizeof(Win);
  auto kern = window_fwd_kernel<HG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("window_attn_fwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws) * (heads / HG));
  kern<<<grid, NTH, smem, st>>>((const bf16*)qkv, rpb, (bf16*)out, R, ws, shift, heads);
  return check_launch("window_attn_fwd(tc)");
}

template <int HG>
static int bwd_launch(const void* qkv, const float* rpb, const void* dout, void* dqkv, int64_t frames, int R, int ws, int shift,
                      int heads, cudaStream_t st) {
  constexpr int LD = HG * 64 + 16;
  const size_t smem = (size_t)4 * (ws * ws + 1) * LD + (size_t)2 * (ws * ws + 2) * LDP + sizeof(float) * HG * tbl_stride(ws) + sizeof(Win);
  auto kern = window_bwd_kernel<HG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("window_attn_bwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws) * (heads / HG));
  kern<<<grid, NTH, smem, st>>>((const bf16*)qkv, rpb, (const bf16*)dout, (bf16*)dqkv, R, ws, shift, heads);
  return check_launch("window_attn_bwd(tc)");
}

}  // namespace wtc

int window_attn_fwd_bf16(const void* qkv, const float* rpb, void* out, int64_t frames, int R, int ws, int shift, int heads,
                         cudaStream_t st) {
  switch (wtc::pick_hg(heads)) {
    case 3: return wtc::fwd_launch<3>(qkv, rpb, out, frames, R, ws, shift, heads, st);
    case 2: return wtc::fwd_launch<2>(qkv, rpb, out, frames, R, ws, shift, heads, st);
    default: return wtc::fwd_launch<1>(qkv, rpb, out, frames, R, ws, shift, heads, st);
  }
}

int window_attn_bwd_bf16(const void* qkv, const float* rpb, const void* dout, void* dqkv, int64_t frames, int R, int ws,
                         int shift, int heads, cudaStream_t st) {
  switch (wtc::pick_hg(heads)) {
    case 3: return wtc::bwd_launch<3>(qkv, rpb, dout, dqkv, frames, R, ws, shift, heads, st);
    case 2: return wtc::bwd_launch<2>(qkv, rpb, dout, dqkv, frames, R, ws, shift, heads, st);
    default: return wtc::bwd_launch<1>(qkv, rpb, dout, dqkv, frames, R, ws, shift, heads, st);
  }
}

}  // namespace stgcma
