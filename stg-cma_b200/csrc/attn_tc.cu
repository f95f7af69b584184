// Swin window attention (WindowAttention.forward, temporal=False; Swin_AVE.py:238-276), bf16 production
// kernels on the warp-level tensor cores (mma.cuh).  roll / window_partition / window_reverse are addressing:
// a CTA gathers the ws*ws (<= 49) token rows of one (frame, window) for a group of HG heads straight from the
// token-major qkv tensor with cp.async, keeps q / k / v (and dO in the backward) in shared memory, and each of
// its 4 warps owns one 16-row block of the (padded to 64) window for every head of the group.  The relative
// position bias is looked up from a per-head table in shared memory; the 0 / -100 shift mask is a per-thread
// bit mask computed once from the window coordinates.
//   forward : S = scale q k^T + bias + mask -> softmax -> P v, all in registers (14 + 16 MMAs per 16 x 49 block)
//   backward: path A (warp = query rows)  recompute P, dP = dO v^T, delta = rowsum(P dP), dq = scale dS k
//             path B (warp = key rows)    recompute P^T with the saved row statistics, dk = scale dS^T q, dv = P^T dO
//             (recomputing the 49 x 49 scores transposed is cheaper than transposing register fragments)
#include <stdlib.h>
#include "mma.cuh"

namespace stgcma {
namespace wtc {

using namespace mma;

constexpr int HD = 32;
constexpr int NP = 64;       // padded window length
constexpr int NTH = 128;     // 4 warps
constexpr float kScale = 0.17677669529663689f;  // 32^-0.5

struct Win {
  int16_t tok[NP];           // token row (within the frame) of window slot i
  int16_t rowoff[NP];        // yi * (2ws-1) + xi                         (0 for pad slots)
  int16_t coloff[NP];        // (ws-1-yj) * (2ws-1) + (ws-1-xj)           (0 for pad slots)
  int16_t reg[NP];           // shift-mask region id
};

__device__ __forceinline__ void win_setup(Win& w, int wy, int wx, int ws, int shift, int R, int N, int tid) {
  if (tid < NP) {
    const int i = tid;
    if (i < N) {
      const int iy = i / ws, ix = i - iy * ws;
      w.tok[i] = (int16_t)(win_src(wy, iy, ws, shift, R) * R + win_src(wx, ix, ws, shift, R));
      w.rowoff[i] = (int16_t)(iy * (2 * ws - 1) + ix);
      w.coloff[i] = (int16_t)((ws - 1 - iy) * (2 * ws - 1) + (ws - 1 - ix));
      w.reg[i] = (int16_t)(shift > 0 ? 3 * win_region1(wy * ws + iy, ws, shift, R) + win_region1(wx * ws + ix, ws, shift, R) : 0);
    } else {
      w.tok[i] = 0; w.rowoff[i] = 0; w.coloff[i] = 0; w.reg[i] = 0;
    }
  }
}

// gather [N rows][HG*32] slices (column offset col0 of a row of `ld_g` elements) into a [NP][LD] tile; pad rows zero.
// Thread layout: CH = HG*4 consecutive threads copy the CH 16-byte chunks of one row, NTH / CH rows per pass -- the
// (row, chunk) split is computed once per thread instead of a division per chunk (this loop was ~13 % of the kernel's
// instructions; ncu source page, profiles/r01f).
template <int HG>
__device__ __forceinline__ void gather_async(uint8_t* tile, const bf16* __restrict__ src, int64_t frame_row0, int64_t ld_g,
                                             int col0, const Win& w, int N, int tid) {
  constexpr int CH = HG * 4, LD = HG * 64 + 16, RPP = NTH / CH;   // rows per pass
  const int r0 = tid / CH, c = tid - r0 * CH;
  if (r0 >= RPP) return;                                            // NTH % CH leftover threads
  const bf16* base = src + frame_row0 * ld_g + col0 + c * 8;
  uint8_t* dst = tile + r0 * LD + c * 16;
#pragma unroll 1
  for (int r = r0; r < NP; r += RPP, dst += RPP * LD) {
    if (r < N) cp_async16(dst, base + (int64_t)w.tok[r] * ld_g);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}

// Per-lane column metadata of the 14 score columns a lane owns (c = i*8 + 2t + e), head independent and hoisted out of
// the head loops: rowoff(c) | region(c) << 8, or -1 for pad columns.  coloff(c) = K0 - rowoff(c), K0 = (ws-1)*(2ws-1)+ws-1.
__device__ __forceinline__ void col_meta(int (&cm)[14], const Win& w, int N, int t) {
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = i * 8 + 2 * t + e;
      cm[2 * i + e] = c < N ? ((int)w.rowoff[c] | ((int)w.reg[c] << 8)) : -1;
    }
}

// s (7 n-tiles of a 16 x 56 block, raw q.k accumulators) -> log2-domain logits with bias and mask applied.
// ROWS_ARE_QUERIES: idx = rowoff(row) + coloff(col); else idx = rowoff(col) + coloff(row) (transposed scores).
// oa / ob: bias offsets of this thread's two rows (rowoff resp. coloff), ga / gb their mask regions.
template <bool ROWS_ARE_QUERIES>
__device__ __forceinline__ void apply_bias_mask(float (&s)[7][4], const float* __restrict__ tbl, const int (&cm)[14], int oa,
                                                int ob, int ga, int gb, int K0, bool shifted) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int m = cm[2 * i + e];
      if (m >= 0) {
        const int oc = ROWS_ARE_QUERIES ? K0 - (m & 255) : (m & 255);
        const int gc = m >> 8;
        float va = s[i][e] * kScale + tbl[oa + oc];
        float vb = s[i][2 + e] * kScale + tbl[ob + oc];
        if (shifted) {
          if (ga != gc) va -= 100.0f;
          if (gb != gc) vb -= 100.0f;
        }
        s[i][e] = va * kLog2e;
        s[i][2 + e] = vb * kLog2e;
      } else {
        s[i][e] = -INFINITY;
        s[i][2 + e] = -INFINITY;
      }
    }
  }
}

// write this lane's part of a 16 x 32 fp32 accumulator block (rows ra / rb of the tile, head h) as bf16 into a
// [NP][LD] shared tile (the warp's own rows: conflict-free, see the LD padding)
__device__ __forceinline__ void stage_rows(uint8_t* tile, int LD, const float (&o)[4][4], float sa, float sb, int ra, int rb,
                                           int h, int t) {
  uint8_t* pa = tile + ra * LD + (h * HD + 2 * t) * 2;
  uint8_t* pb = tile + rb * LD + (h * HD + 2 * t) * 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    *reinterpret_cast<uint32_t*>(pa + i * 16) = pack_bf16(o[i][0] * sa, o[i][1] * sa);
    *reinterpret_cast<uint32_t*>(pb + i * 16) = pack_bf16(o[i][2] * sb, o[i][3] * sb);
  }
}

// coalesced scatter of a staged [N rows][HG*32] tile to the token-major destination (16 B per lane, a row's HG*64 B
// segment is contiguous in global memory); same thread layout as gather_async
template <int HG>
__device__ __forceinline__ void scatter_rows(const uint8_t* tile, bf16* __restrict__ dst, int64_t frame_row0, int64_t ld_g,
                                             int col0, const Win& w, int N, int tid) {
  constexpr int CH = HG * 4, LD = HG * 64 + 16, RPP = NTH / CH;
  const int r0 = tid / CH, c = tid - r0 * CH;
  if (r0 >= RPP) return;
  bf16* base = dst + frame_row0 * ld_g + col0 + c * 8;
  const uint8_t* sp = tile + r0 * LD + c * 16;
#pragma unroll 1
  for (int r = r0; r < N; r += RPP, sp += RPP * LD)
    *reinterpret_cast<uint4*>(base + (int64_t)w.tok[r] * ld_g) = *reinterpret_cast<const uint4*>(sp);
}

// 16 x 56 raw scores: A = 16 rows of `ta` (row block rb), B = rows of `tb` as [n][k]; features h*32 .. +31
__device__ __forceinline__ void scores(float (&s)[7][4], const uint8_t* ta, const uint8_t* tb, int LD, int rb, int h, int lane) {
#pragma unroll
  for (int i = 0; i < 7; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
  uint32_t a[2][4];
  load_a(a[0], ta, LD, rb * 16, h * HD, lane);
  load_a(a[1], ta, LD, rb * 16, h * HD + 16, lane);
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t b[4];
      load_b_nk(b, tb, LD, n2 * 16, h * HD + ks * 16, lane);
      mma16816(s[2 * n2], a[ks], b[0], b[1]);
      if (n2 < 3) mma16816(s[2 * n2 + 1], a[ks], b[2], b[3]);
    }
  }
}

// o[4 n-tiles of 32 features] += P (16 x 64, from 7 n-tiles of probabilities; cols >= 56 are zero) . tile[k][n]
__device__ __forceinline__ void pv(float (&o)[4][4], const float (&p)[7][4], const uint8_t* tv, int LD, int h, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    pa[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    if (kk < 3) {
      pa[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
      pa[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    } else {
      pa[2] = 0u;
      pa[3] = 0u;
    }
#pragma unroll
    for (int n2 = 0; n2 < 2; ++n2) {
      uint32_t b[4];
      load_b_kn(b, tv, LD, kk * 16, h * HD + n2 * 16, lane);
      mma16816(o[2 * n2], pa, b[0], b[1]);
      mma16816(o[2 * n2 + 1], pa, b[2], b[3]);
    }
  }
}

template <int HG>
__global__ void __launch_bounds__(NTH) window_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ rpb,
                                                         bf16* __restrict__ out, int R, int ws, int shift, int heads) {
  constexpr int LD = HG * 64 + 16;
  extern __shared__ __align__(16) uint8_t wsm[];
  uint8_t* tq = wsm;
  uint8_t* tk = tq + NP * LD;
  uint8_t* tv = tk + NP * LD;
  float* tbl = reinterpret_cast<float*>(tv + NP * LD);   // [HG][T2]
  const int T2 = ((2 * ws - 1) * (2 * ws - 1) + 3) & ~3;   // table stride, padded so what follows stays 16 B aligned
  const int T2v = (2 * ws - 1) * (2 * ws - 1);
  Win& w = *reinterpret_cast<Win*>(tbl + HG * T2);
  const int nwr = R / ws, nW = nwr * nwr, N = ws * ws, C = heads * HD, L = R * R, ngrp = heads / HG;
  int64_t b = blockIdx.x;
  const int hg = (int)(b % ngrp); b /= ngrp;
  const int wi = (int)(b % nW);
  const int64_t frame = b / nW;
  const int wy = wi / nwr, wx = wi - wy * nwr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  win_setup(w, wy, wx, ws, shift, R, N, tid);
  __syncthreads();
  const int64_t row0 = frame * L;
  gather_async<HG>(tq, qkv, row0, 3 * C, hg * HG * HD, w, N, tid);
  gather_async<HG>(tk, qkv, row0, 3 * C, C + hg * HG * HD, w, N, tid);
  gather_async<HG>(tv, qkv, row0, 3 * C, 2 * C + hg * HG * HD, w, N, tid);
  cp_async_commit();
  for (int e = tid; e < HG * T2; e += NTH) {   // overlaps the gather
    const int hh = e / T2, i = e - hh * T2;
    tbl[e] = i < T2v ? __ldg(rpb + (int64_t)i * heads + hg * HG + hh) : 0.f;
  }
  const int ra = warp * 16 + g, rb = ra + 8;
  const int K0 = (ws - 1) * (2 * ws - 1) + (ws - 1);
  int cm[14];
  col_meta(cm, w, N, t);
  const int rav = ra < N ? ra : 0, rbv = rb < N ? rb : 0;
  const int oa = w.rowoff[rav], ob = w.rowoff[rbv], ga = w.reg[rav], gb = w.reg[rbv];
  cp_async_wait<0>();
  __syncthreads();

#pragma unroll 1
  for (int h = 0; h < HG; ++h) {
    float s[7][4];
    scores(s, tq, tk, LD, warp, h, lane);
    apply_bias_mask<true>(s, tbl + h * T2, cm, oa, ob, ga, gb, K0, shift > 0);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      m0 = fmaxf(m0, fmaxf(s[i][0], s[i][1]));
      m1 = fmaxf(m1, fmaxf(s[i][2], s[i][3]));
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      s[i][0] = ex2(s[i][0] - m0);
      s[i][1] = ex2(s[i][1] - m0);
      s[i][2] = ex2(s[i][2] - m1);
      s[i][3] = ex2(s[i][3] - m1);
      l0 += s[i][0] + s[i][1];
      l1 += s[i][2] + s[i][3];
    }
    const float inv0 = 1.f / quad_sum(l0), inv1 = 1.f / quad_sum(l1);
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    pv(o, s, tv, LD, h, lane);
    // q rows of this warp / head are dead (only this warp read them): they stage the output
    __syncwarp();
    stage_rows(tq, LD, o, inv0, inv1, ra, rb, h, t);
  }
  __syncthreads();
  scatter_rows<HG>(tq, out, row0, C, hg * HG * HD, w, N, tid);
}

template <int HG>
__global__ void __launch_bounds__(NTH) window_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ rpb,
                                                         const bf16* __restrict__ dout, bf16* __restrict__ dqkv, int R, int ws,
                                                         int shift, int heads) {
  constexpr int LD = HG * 64 + 16;
  extern __shared__ __align__(16) uint8_t wsm[];
  uint8_t* tq = wsm;
  uint8_t* tk = tq + NP * LD;
  uint8_t* tv = tk + NP * LD;
  uint8_t* td = tv + NP * LD;
  uint8_t* tdq = td + NP * LD;                           // dq staging (path A output; q itself is still needed by path B)
  float* tbl = reinterpret_cast<float*>(tdq + NP * LD);  // [HG][T2]
  const int T2 = ((2 * ws - 1) * (2 * ws - 1) + 3) & ~3;
  const int T2v = (2 * ws - 1) * (2 * ws - 1);
  float* s_lse = tbl + HG * T2;                          // [HG][NP]  log2-domain logsumexp per query row (+inf on pad rows)
  float* s_delta = s_lse + HG * NP;                      // [HG][NP]
  Win& w = *reinterpret_cast<Win*>(s_delta + HG * NP);
  const int nwr = R / ws, nW = nwr * nwr, N = ws * ws, C = heads * HD, L = R * R, ngrp = heads / HG;
  int64_t b = blockIdx.x;
  const int hg = (int)(b % ngrp); b /= ngrp;
  const int wi = (int)(b % nW);
  const int64_t frame = b / nW;
  const int wy = wi / nwr, wx = wi - wy * nwr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  win_setup(w, wy, wx, ws, shift, R, N, tid);
  __syncthreads();
  const int64_t row0 = frame * L;
  gather_async<HG>(tq, qkv, row0, 3 * C, hg * HG * HD, w, N, tid);
  gather_async<HG>(tk, qkv, row0, 3 * C, C + hg * HG * HD, w, N, tid);
  gather_async<HG>(tv, qkv, row0, 3 * C, 2 * C + hg * HG * HD, w, N, tid);
  gather_async<HG>(td, dout, row0, C, hg * HG * HD, w, N, tid);
  cp_async_commit();
  for (int e = tid; e < HG * T2; e += NTH) {   // overlaps the gather
    const int hh = e / T2, i = e - hh * T2;
    tbl[e] = i < T2v ? __ldg(rpb + (int64_t)i * heads + hg * HG + hh) : 0.f;
  }
  const int ra = warp * 16 + g, rb = ra + 8;
  const int K0 = (ws - 1) * (2 * ws - 1) + (ws - 1);
  int cm[14];
  col_meta(cm, w, N, t);
  const int rav = ra < N ? ra : 0, rbv = rb < N ? rb : 0;
  const int roa = w.rowoff[rav], rob = w.rowoff[rbv], ga = w.reg[rav], gb = w.reg[rbv];
  cp_async_wait<0>();
  __syncthreads();

  // ---------------- path A: this warp's rows are queries ----------------
#pragma unroll 1
  for (int h = 0; h < HG; ++h) {
    float s[7][4], dp[7][4];
    scores(s, tq, tk, LD, warp, h, lane);
    apply_bias_mask<true>(s, tbl + h * T2, cm, roa, rob, ga, gb, K0, shift > 0);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      m0 = fmaxf(m0, fmaxf(s[i][0], s[i][1]));
      m1 = fmaxf(m1, fmaxf(s[i][2], s[i][3]));
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float l0 = 0.f, l1 = 0.f;
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
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    scores(dp, td, tv, LD, warp, h, lane);   // dP = dO . v^T
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      s[i][0] *= inv0; s[i][1] *= inv0; s[i][2] *= inv1; s[i][3] *= inv1;
      d0 += s[i][0] * dp[i][0] + s[i][1] * dp[i][1];
      d1 += s[i][2] * dp[i][2] + s[i][3] * dp[i][3];
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      s[i][0] *= (dp[i][0] - d0); s[i][1] *= (dp[i][1] - d0);
      s[i][2] *= (dp[i][2] - d1); s[i][3] *= (dp[i][3] - d1);
    }
    if (t == 0) {
      s_lse[h * NP + ra] = ra < N ? m0 + log2f(l0) : INFINITY;
      s_lse[h * NP + rb] = rb < N ? m1 + log2f(l1) : INFINITY;
      s_delta[h * NP + ra] = d0;
      s_delta[h * NP + rb] = d1;
    }
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    pv(o, s, tk, LD, h, lane);               // dq = dS . k
    stage_rows(tdq, LD, o, kScale, kScale, ra, rb, h, t);
  }
  __syncthreads();
  // ---------------- path B: this warp's rows are keys; columns are queries ----------------
  const int coa = K0 - roa, cob = K0 - rob;
#pragma unroll 1
  for (int h = 0; h < HG; ++h) {
    float s[7][4], dp[7][4];
    scores(s, tk, tq, LD, warp, h, lane);    // S^T = k . q^T
    apply_bias_mask<false>(s, tbl + h * T2, cm, coa, cob, ga, gb, K0, shift > 0);
    scores(dp, tv, td, LD, warp, h, lane);   // dP^T = v . dO^T
    const float* lse_h = s_lse + h * NP;
    const float* del_h = s_delta + h * NP;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int c = i * 8 + 2 * t;
      const float2 ls = *reinterpret_cast<const float2*>(lse_h + c);
      const float2 dl = *reinterpret_cast<const float2*>(del_h + c);
      // P^T[j, i] = exp2(s - lse_i)  (0 for pad queries: lse = +inf, and for pad keys the row is discarded)
      const float p0 = ex2(s[i][0] - ls.x), p1 = ex2(s[i][1] - ls.y), p2 = ex2(s[i][2] - ls.x), p3 = ex2(s[i][3] - ls.y);
      s[i][0] = p0; s[i][1] = p1; s[i][2] = p2; s[i][3] = p3;
      dp[i][0] = p0 * (dp[i][0] - dl.x); dp[i][1] = p1 * (dp[i][1] - dl.y);
      dp[i][2] = p2 * (dp[i][2] - dl.x); dp[i][3] = p3 * (dp[i][3] - dl.y);
    }
    float ok[4][4], ov[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok[i][0] = ok[i][1] = ok[i][2] = ok[i][3] = 0.f;
      ov[i][0] = ov[i][1] = ov[i][2] = ov[i][3] = 0.f;
    }
    pv(ok, dp, tq, LD, h, lane);             // dk = dS^T . q
    pv(ov, s, td, LD, h, lane);              // dv = P^T . dO
    // k / v rows of this warp / head are dead now (path B reads them only as this warp's A operand; path A, which
    // read every row, is complete for all warps): they stage dk / dv
    __syncwarp();
    stage_rows(tk, LD, ok, kScale, kScale, ra, rb, h, t);
    stage_rows(tv, LD, ov, 1.f, 1.f, ra, rb, h, t);
  }
  __syncthreads();
  scatter_rows<HG>(tdq, dqkv, row0, 3 * C, hg * HG * HD, w, N, tid);
  scatter_rows<HG>(tk, dqkv, row0, 3 * C, C + hg * HG * HD, w, N, tid);
  scatter_rows<HG>(tv, dqkv, row0, 3 * C, 2 * C + hg * HG * HD, w, N, tid);
}

static int pick_hg(int heads) {
  static const int force = getenv("STGCMA_WIN_HG") ? atoi(getenv("STGCMA_WIN_HG")) : 0;   // tuning knob
  if (force > 0 && heads % force == 0) return force;
  if (heads % 3 == 0) return 3;   // HG = 4 would need 88 KB in the backward (5 tiles): 2 CTAs per SM instead of 3
  if (heads % 2 == 0) return 2;
  return 1;
}

template <int HG>
static int fwd_launch(const void* qkv, const float* rpb, void* out, int64_t frames, int R, int ws, int shift, int heads,
                      cudaStream_t st) {
  constexpr int LD = HG * 64 + 16;
  const int T2 = ((2 * ws - 1) * (2 * ws - 1) + 3) & ~3;
  const size_t smem = (size_t)3 * NP * LD + sizeof(float) * HG * T2 + sizeof(Win);
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
  const int T2 = ((2 * ws - 1) * (2 * ws - 1) + 3) & ~3;
  const size_t smem = (size_t)5 * NP * LD + sizeof(float) * (HG * T2 + 2 * HG * NP) + sizeof(Win);
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
