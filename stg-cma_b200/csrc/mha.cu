// Multi-head attention core of the CLIP-ViT blocks: nn.MultiheadAttention(d_model, n_head)(x, x, x) without mask
// (CLIP_AVE.py:54, :106-108) between the packed in-projection and the out-projection (both are stgcma_gemm calls):
//
//     o[:, h] = softmax(q_h k_h^T / sqrt(hd)) v_h          per sequence and head
//
// The same kernels serve the spatial attention (sequence = the 197 video / 49 audio tokens of a frame, rows contiguous)
// and the temporal attention (sequence = the T frames of one (clip, token): rows `inner` apart), CLIP_AVE.py:369-385.
//
// bf16 production kernels: warp-level tensor cores (mma.sync m16n8k16), flash formulation.  Sequences are short
// (<= 257 tokens), so the K / V rows of a (sequence, head) are staged in shared memory ONCE per CTA and every warp owns
// 16-row query tiles: scores, probabilities and the output accumulator never leave registers.
//   forward : out, lse (natural log) per (row, head)
//   backward: two kernels, both atomics-free
//     dq  : query-tile owner.   D_i = <dO_i, O_i>; P = exp(s - lse_i); dS = P (dO_i.v_j - D_i); dQ_i = scale * dS K
//     dkv : key-tile owner.     P^T and dS^T recomputed from (k_i, q_j), (v_i, dO_j): dV_i = P^T dO, dK_i = scale * dS^T Q
// fp32 check mode: one thread per row, everything from global memory (numerics reference, not a fast path).
#include <stdlib.h>
#include <mutex>
#include "mma.cuh"

namespace stgcma {
namespace mha {

using namespace mma;

template <int HD>
struct Cfg {
  static constexpr int LD = HD * 2 + 16;  // smem row stride in bytes (odd multiple of 16: conflict-free ldmatrix)
  static constexpr int KS = HD / 16;
  static constexpr int CH = HD / 8;       // 16-byte chunks per row
};

struct Seq {
  int64_t n_seq;
  int seq_len, inner;
  // row of token i of sequence s:  inner = 1 -> contiguous sequences; inner = N -> the T frames of one (clip, token)
  __device__ __forceinline__ int64_t row(int64_t s, int i) const {
    const int64_t o = s / inner, r = s - o * inner;
    return o * (int64_t)seq_len * inner + r + (int64_t)i * inner;
  }
};

// rows [i0, i0 + nrows) of sequence s, columns [0, HD) of `base` (row stride ld elements) -> smem tile; rows beyond the
// sequence (or sequences beyond n_seq) are zero-filled
template <int HD>
__device__ __forceinline__ void load_tile(uint8_t* tile, const bf16* __restrict__ base, int64_t ld, const Seq& sq, int64_t s,
                                          int i0, int nrows, int tid, int nthr) {
  constexpr int CH = Cfg<HD>::CH, LD = Cfg<HD>::LD;
  const bool live = s < sq.n_seq;
  const bf16* row0 = base + (live ? sq.row(s, 0) : 0) * ld;      // one 64-bit division per call, not per 16-byte chunk
  const int64_t step = (int64_t)sq.inner * ld;
  for (int e = tid; e < nrows * CH; e += nthr) {
    const int r = e / CH, c = e - r * CH;
    uint8_t* dst = tile + r * LD + c * 16;
    const int i = i0 + r;
    if (live && i < sq.seq_len) cp_async16(dst, row0 + i * step + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}

// 16 x HD bf16 tile in smem -> rows [i0, i0 + 16) of sequence s (16-byte coalesced stores), one warp
template <int HD>
__device__ __forceinline__ void store_tile(const uint8_t* tile, bf16* __restrict__ base, int64_t ld, const Seq& sq, int64_t s,
                                           int i0, int lane) {
  constexpr int CH = Cfg<HD>::CH, LD = Cfg<HD>::LD;
  bf16* row0 = base + sq.row(s, 0) * ld;
  const int64_t step = (int64_t)sq.inner * ld;
  for (int e = lane; e < 16 * CH; e += 32) {
    const int r = e / CH, c = e - r * CH;
    const int i = i0 + r;
    if (i < sq.seq_len) *reinterpret_cast<uint4*>(row0 + i * step + c * 8) = *reinterpret_cast<const uint4*>(tile + r * LD + c * 16);
  }
}

// accumulator fragments (rows g / g + 8 of the warp's 16-row tile) -> bf16 smem tile
template <int HD>
__device__ __forceinline__ void acc_to_tile(uint8_t* tile, const float (&acc)[2 * Cfg<HD>::KS][4], float m0, float m1, int lane) {
  constexpr int LD = Cfg<HD>::LD;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2 * Cfg<HD>::KS; ++i) {
    const int col = i * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(tile + g * LD + col * 2) = pack_bf16(acc[i][0] * m0, acc[i][1] * m0);
    *reinterpret_cast<uint32_t*>(tile + (g + 8) * LD + col * 2) = pack_bf16(acc[i][2] * m1, acc[i][3] * m1);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) mha_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                      float* __restrict__ lse, Seq sq, int heads, int SPC, int KP,
                                                      float scale_log2) {
  constexpr int LD = Cfg<HD>::LD, KS = Cfg<HD>::KS;
  extern __shared__ __align__(16) uint8_t msm[];
  const int NW = blockDim.x >> 5;
  uint8_t* Ks = msm;                       // [SPC][KP][LD]
  uint8_t* Vs = Ks + SPC * KP * LD;        // [SPC][KP][LD]
  uint8_t* Qs = Vs + SPC * KP * LD;        // [NW][16][LD]
  const int h = blockIdx.y;
  const int64_t s0 = (int64_t)blockIdx.x * SPC;
  const int D = heads * HD;
  const int64_t ldq = 3 * (int64_t)D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int q = 0; q < SPC; ++q) {
    load_tile<HD>(Ks + q * KP * LD, qkv + D + h * HD, ldq, sq, s0 + q, 0, KP, tid, blockDim.x);
    load_tile<HD>(Vs + q * KP * LD, qkv + 2 * D + h * HD, ldq, sq, s0 + q, 0, KP, tid, blockDim.x);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int QT = KP >> 4;
  uint8_t* qs = Qs + warp * 16 * LD;
  for (int item = warp; item < SPC * QT; item += NW) {
    const int q = item / QT, qt = item - q * QT;
    const int64_t s = s0 + q;
    if (s >= sq.n_seq || qt * 16 >= sq.seq_len) continue;
    const uint8_t* ks_ = Ks + q * KP * LD;
    const uint8_t* vs_ = Vs + q * KP * LD;
    load_tile<HD>(qs, qkv + h * HD, ldq, sq, s, qt * 16, 16, lane, 32);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) load_a(qa[ks], qs, LD, 0, ks * 16, lane);
    float acc[2 * KS][4];
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    for (int c0 = 0; c0 < KP; c0 += 64) {
      const int nb = min(4, (KP - c0) >> 4);
      float sc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        if (n2 < nb) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t b[4];
            load_b_nk(b, ks_, LD, c0 + n2 * 16, ks * 16, lane);
            mma16816(sc[2 * n2], qa[ks], b[0], b[1]);
            mma16816(sc[2 * n2 + 1], qa[ks], b[2], b[3]);
          }
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = sc[i][e] * scale_log2;
          if (c0 + i * 8 + 2 * t + (e & 1) >= sq.seq_len) v = -INFINITY;
          sc[i][e] = v;
        }
        mx0 = fmaxf(mx0, fmaxf(sc[i][0], sc[i][1]));
        mx1 = fmaxf(mx1, fmaxf(sc[i][2], sc[i][3]));
      }
      mx0 = quad_max(mx0);
      mx1 = quad_max(mx1);
      const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
      const float c0s = ex2(m_run[0] - mn0), c1s = ex2(m_run[1] - mn1);
      m_run[0] = mn0;
      m_run[1] = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sc[i][0] = ex2(sc[i][0] - mn0);
        sc[i][1] = ex2(sc[i][1] - mn0);
        sc[i][2] = ex2(sc[i][2] - mn1);
        sc[i][3] = ex2(sc[i][3] - mn1);
        ps0 += sc[i][0] + sc[i][1];
        ps1 += sc[i][2] + sc[i][3];
      }
      l_run[0] = l_run[0] * c0s + ps0;
      l_run[1] = l_run[1] * c1s + ps1;
#pragma unroll
      for (int i = 0; i < 2 * KS; ++i) {
        acc[i][0] *= c0s;
        acc[i][1] *= c0s;
        acc[i][2] *= c1s;
        acc[i][3] *= c1s;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk < nb) {
          uint32_t pa[4];
          pa[0] = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
          pa[1] = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
          pa[2] = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
          pa[3] = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
          for (int n2 = 0; n2 < KS; ++n2) {
            uint32_t b[4];
            load_b_kn(b, vs_, LD, c0 + kk * 16, n2 * 16, lane);
            mma16816(acc[2 * n2], pa, b[0], b[1]);
            mma16816(acc[2 * n2 + 1], pa, b[2], b[3]);
          }
        }
      }
    }
    const float l0 = quad_sum(l_run[0]), l1 = quad_sum(l_run[1]);
    __syncwarp();   // every lane has its q fragments: the staging tile can be overwritten with the output
    acc_to_tile<HD>(qs, acc, 1.f / l0, 1.f / l1, lane);
    __syncwarp();
    store_tile<HD>(qs, out + h * HD, D, sq, s, qt * 16, lane);
    if (t == 0) {
      const int ia = qt * 16 + g, ib = ia + 8;
      if (ia < sq.seq_len) lse[sq.row(s, ia) * heads + h] = (m_run[0] + log2f(l0)) * 0.6931471805599453f;
      if (ib < sq.seq_len) lse[sq.row(s, ib) * heads + h] = (m_run[1] + log2f(l1)) * 0.6931471805599453f;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward, dQ (query-tile owner); also writes D = <dO, O> per (row, head) for the dK / dV kernel
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) mha_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                         const float* __restrict__ lse, const bf16* __restrict__ dout,
                                                         bf16* __restrict__ dqkv, float* __restrict__ Dbuf, Seq sq, int heads,
                                                         int SPC, int KP, float scale, float scale_log2) {
  constexpr int LD = Cfg<HD>::LD, KS = Cfg<HD>::KS;
  extern __shared__ __align__(16) uint8_t msm[];
  const int NW = blockDim.x >> 5;
  uint8_t* Ks = msm;
  uint8_t* Vs = Ks + SPC * KP * LD;
  uint8_t* Ws = Vs + SPC * KP * LD;        // [NW][2][16][LD]: q tile, dO tile
  const int h = blockIdx.y;
  const int64_t s0 = (int64_t)blockIdx.x * SPC;
  const int D = heads * HD;
  const int64_t ldq = 3 * (int64_t)D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int q = 0; q < SPC; ++q) {
    load_tile<HD>(Ks + q * KP * LD, qkv + D + h * HD, ldq, sq, s0 + q, 0, KP, tid, blockDim.x);
    load_tile<HD>(Vs + q * KP * LD, qkv + 2 * D + h * HD, ldq, sq, s0 + q, 0, KP, tid, blockDim.x);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int QT = KP >> 4;
  uint8_t* qs = Ws + warp * 32 * LD;
  uint8_t* ds = qs + 16 * LD;
  for (int item = warp; item < SPC * QT; item += NW) {
    const int q = item / QT, qt = item - q * QT;
    const int64_t s = s0 + q;
    if (s >= sq.n_seq || qt * 16 >= sq.seq_len) continue;
    const uint8_t* ks_ = Ks + q * KP * LD;
    const uint8_t* vs_ = Vs + q * KP * LD;
    load_tile<HD>(qs, qkv + h * HD, ldq, sq, s, qt * 16, 16, lane, 32);
    load_tile<HD>(ds, dout + h * HD, D, sq, s, qt * 16, 16, lane, 32);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    // D_i = <dO_i, O_i>: lane (2r + half) sums half of row r; O straight from global memory
    float Dv;
    {
      const int r = lane >> 1, half = lane & 1, i = qt * 16 + r;
      float a = 0.f;
      if (i < sq.seq_len) {
        const bf16* orow = o + sq.row(s, i) * D + h * HD + half * (HD / 2);
        const uint8_t* drow = ds + r * LD + half * HD;
#pragma unroll
        for (int c = 0; c < HD / 2; c += 8) {
          float ov[8], dv[8];
          load8(orow + c, ov);
          unpack8(*reinterpret_cast<const uint4*>(drow + c * 2), dv);
#pragma unroll
          for (int k = 0; k < 8; ++k) a = fmaf(ov[k], dv[k], a);
        }
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      Dv = a;
      if (half == 0 && i < sq.seq_len) Dbuf[sq.row(s, i) * heads + h] = a;
    }
    const float D0 = __shfl_sync(0xffffffffu, Dv, 2 * g), D1 = __shfl_sync(0xffffffffu, Dv, 2 * (g + 8));
    const int ia = qt * 16 + g, ib = ia + 8;
    const float lse0 = ia < sq.seq_len ? lse[sq.row(s, ia) * heads + h] * kLog2e : INFINITY;
    const float lse1 = ib < sq.seq_len ? lse[sq.row(s, ib) * heads + h] * kLog2e : INFINITY;
    uint32_t qa[KS][4], da[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      load_a(qa[ks], qs, LD, 0, ks * 16, lane);
      load_a(da[ks], ds, LD, 0, ks * 16, lane);
    }
    float acc[2 * KS][4];
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    for (int c0 = 0; c0 < KP; c0 += 32) {
      const int nb = min(2, (KP - c0) >> 4);
      float sc[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[i][e] = dp[i][e] = 0.f;
#pragma unroll
      for (int n2 = 0; n2 < 2; ++n2) {
        if (n2 < nb) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t bk[4], bv[4];
            load_b_nk(bk, ks_, LD, c0 + n2 * 16, ks * 16, lane);
            load_b_nk(bv, vs_, LD, c0 + n2 * 16, ks * 16, lane);
            mma16816(sc[2 * n2], qa[ks], bk[0], bk[1]);
            mma16816(sc[2 * n2 + 1], qa[ks], bk[2], bk[3]);
            mma16816(dp[2 * n2], da[ks], bv[0], bv[1]);
            mma16816(dp[2 * n2 + 1], da[ks], bv[2], bv[3]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = c0 + i * 8 + 2 * t + (e & 1) < sq.seq_len;
          const float P = ok ? ex2(sc[i][e] * scale_log2 - ((e & 2) ? lse1 : lse0)) : 0.f;
          sc[i][e] = P * (dp[i][e] - ((e & 2) ? D1 : D0)) * scale;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (kk < nb) {
          uint32_t wa[4];
          wa[0] = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
          wa[1] = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
          wa[2] = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
          wa[3] = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
          for (int n2 = 0; n2 < KS; ++n2) {
            uint32_t b[4];
            load_b_kn(b, ks_, LD, c0 + kk * 16, n2 * 16, lane);
            mma16816(acc[2 * n2], wa, b[0], b[1]);
            mma16816(acc[2 * n2 + 1], wa, b[2], b[3]);
          }
        }
      }
    }
    __syncwarp();
    acc_to_tile<HD>(qs, acc, 1.f, 1.f, lane);
    __syncwarp();
    store_tile<HD>(qs, dqkv + h * HD, ldq, sq, s, qt * 16, lane);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward, dK / dV (key-tile owner)
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) mha_bwd_dkv_kernel(const bf16* __restrict__ qkv, const float* __restrict__ lse,
                                                          const bf16* __restrict__ dout, const float* __restrict__ Dbuf,
                                                          bf16* __restrict__ dqkv, Seq sq, int heads, int SPC, int KP,
                                                          float scale, float scale_log2) {
  constexpr int LD = Cfg<HD>::LD, KS = Cfg<HD>::KS;
  extern __shared__ __align__(16) uint8_t msm[];
  const int NW = blockDim.x >> 5;
  uint8_t* Qs = msm;                                   // [SPC][KP][LD]
  uint8_t* Os = Qs + SPC * KP * LD;                    // [SPC][KP][LD]  dO
  uint8_t* Ws = Os + SPC * KP * LD;                    // [NW][2][16][LD]: k tile, v tile
  float* lse2 = reinterpret_cast<float*>(Ws + NW * 32 * LD);   // [SPC][KP]  lse * log2e (+inf for padding rows)
  float* Dsm = lse2 + SPC * KP;                        // [SPC][KP]
  const int h = blockIdx.y;
  const int64_t s0 = (int64_t)blockIdx.x * SPC;
  const int D = heads * HD;
  const int64_t ldq = 3 * (int64_t)D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int q = 0; q < SPC; ++q) {
    load_tile<HD>(Qs + q * KP * LD, qkv + h * HD, ldq, sq, s0 + q, 0, KP, tid, blockDim.x);
    load_tile<HD>(Os + q * KP * LD, dout + h * HD, D, sq, s0 + q, 0, KP, tid, blockDim.x);
  }
  cp_async_commit();
  for (int e = tid; e < SPC * KP; e += blockDim.x) {
    const int q = e / KP, i = e - q * KP;
    const bool ok = i < sq.seq_len && s0 + q < sq.n_seq;
    const int64_t r = ok ? sq.row(s0 + q, i) * heads + h : 0;
    lse2[e] = ok ? lse[r] * kLog2e : INFINITY;
    Dsm[e] = ok ? Dbuf[r] : 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();
  const int QT = KP >> 4;
  uint8_t* kt_ = Ws + warp * 32 * LD;
  uint8_t* vt_ = kt_ + 16 * LD;
  for (int item = warp; item < SPC * QT; item += NW) {
    const int q = item / QT, kt = item - q * QT;
    const int64_t s = s0 + q;
    if (s >= sq.n_seq || kt * 16 >= sq.seq_len) continue;
    const uint8_t* qs_ = Qs + q * KP * LD;
    const uint8_t* os_ = Os + q * KP * LD;
    const float* cl = lse2 + q * KP;
    const float* cD = Dsm + q * KP;
    load_tile<HD>(kt_, qkv + D + h * HD, ldq, sq, s, kt * 16, 16, lane, 32);
    load_tile<HD>(vt_, qkv + 2 * D + h * HD, ldq, sq, s, kt * 16, 16, lane, 32);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    uint32_t ka[KS][4], va[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      load_a(ka[ks], kt_, LD, 0, ks * 16, lane);
      load_a(va[ks], vt_, LD, 0, ks * 16, lane);
    }
    float accK[2 * KS][4], accV[2 * KS][4];
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) {
      accK[i][0] = accK[i][1] = accK[i][2] = accK[i][3] = 0.f;
      accV[i][0] = accV[i][1] = accV[i][2] = accV[i][3] = 0.f;
    }
    for (int c0 = 0; c0 < KP; c0 += 32) {
      const int nb = min(2, (KP - c0) >> 4);
      float sc[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[i][e] = dp[i][e] = 0.f;
#pragma unroll
      for (int n2 = 0; n2 < 2; ++n2) {
        if (n2 < nb) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t bq[4], bo[4];
            load_b_nk(bq, qs_, LD, c0 + n2 * 16, ks * 16, lane);
            load_b_nk(bo, os_, LD, c0 + n2 * 16, ks * 16, lane);
            mma16816(sc[2 * n2], ka[ks], bq[0], bq[1]);       // s^T = <k_i, q_j>
            mma16816(sc[2 * n2 + 1], ka[ks], bq[2], bq[3]);
            mma16816(dp[2 * n2], va[ks], bo[0], bo[1]);       // dP^T = <v_i, dO_j>
            mma16816(dp[2 * n2 + 1], va[ks], bo[2], bo[3]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= 2 * nb) continue;                  // columns beyond KP: not computed, not consumed (and beyond lse2 / Dsm)
        const int jc = c0 + i * 8 + 2 * t;          // < KP: padded entries carry lse = +inf -> P = 0
        const float2 l2 = *reinterpret_cast<const float2*>(cl + jc);
        const float2 d2 = *reinterpret_cast<const float2*>(cD + jc);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float P = ex2(sc[i][e] * scale_log2 - ((e & 1) ? l2.y : l2.x));
          dp[i][e] = P * (dp[i][e] - ((e & 1) ? d2.y : d2.x)) * scale;   // dS^T
          sc[i][e] = P;                                                  // P^T
        }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (kk < nb) {
          uint32_t pa[4], wa[4];
          pa[0] = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
          pa[1] = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
          pa[2] = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
          pa[3] = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
          wa[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
          wa[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
          wa[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
          wa[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
          for (int n2 = 0; n2 < KS; ++n2) {
            uint32_t bo[4], bq[4];
            load_b_kn(bo, os_, LD, c0 + kk * 16, n2 * 16, lane);
            load_b_kn(bq, qs_, LD, c0 + kk * 16, n2 * 16, lane);
            mma16816(accV[2 * n2], pa, bo[0], bo[1]);         // dV += P^T dO
            mma16816(accV[2 * n2 + 1], pa, bo[2], bo[3]);
            mma16816(accK[2 * n2], wa, bq[0], bq[1]);         // dK += dS^T Q
            mma16816(accK[2 * n2 + 1], wa, bq[2], bq[3]);
          }
        }
      }
    }
    __syncwarp();
    acc_to_tile<HD>(kt_, accK, 1.f, 1.f, lane);
    acc_to_tile<HD>(vt_, accV, 1.f, 1.f, lane);
    __syncwarp();
    store_tile<HD>(kt_, dqkv + D + h * HD, ldq, sq, s, kt * 16, lane);
    store_tile<HD>(vt_, dqkv + 2 * D + h * HD, ldq, sq, s, kt * 16, lane);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 check mode (and the numerics reference of the kernels above): one thread per row
// ---------------------------------------------------------------------------------------------
constexpr int kMaxHd = 128;

template <typename T>
__global__ void __launch_bounds__(128) mha_fwd_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse,
                                                           Seq sq, int heads, int hd, float scale) {
  const int64_t s = blockIdx.x;
  const int h = blockIdx.y;
  const int D = heads * hd;
  const int64_t ldq = 3 * (int64_t)D;
  for (int i = threadIdx.x; i < sq.seq_len; i += blockDim.x) {
    const int64_t ri = sq.row(s, i);
    const T* qi = qkv + ri * ldq + h * hd;
    float q[kMaxHd], acc[kMaxHd];
    for (int c = 0; c < hd; ++c) { q[c] = ld_f(qi + c) * scale; acc[c] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < sq.seq_len; ++j) {
      const T* kj = qkv + sq.row(s, j) * ldq + D + h * hd;
      float d = 0.f;
      for (int c = 0; c < hd; ++c) d = fmaf(q[c], ld_f(kj + c), d);
      const float mn = fmaxf(m, d);
      const float corr = __expf(m - mn), p = __expf(d - mn);
      l = l * corr + p;
      const T* vj = kj + D;
      for (int c = 0; c < hd; ++c) acc[c] = fmaf(p, ld_f(vj + c), acc[c] * corr);
      m = mn;
    }
    const float inv = 1.f / l;
    T* oi = out + ri * D + h * hd;
    for (int c = 0; c < hd; ++c) st_f(oi + c, acc[c] * inv);
    lse[ri * heads + h] = m + logf(l);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) mha_bwd_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ o,
                                                           const float* __restrict__ lse, const T* __restrict__ dout,
                                                           T* __restrict__ dqkv, Seq sq, int heads, int hd, float scale) {
  const int64_t s = blockIdx.x;
  const int h = blockIdx.y;
  const int D = heads * hd;
  const int64_t ldq = 3 * (int64_t)D;
  for (int i = threadIdx.x; i < sq.seq_len; i += blockDim.x) {
    const int64_t ri = sq.row(s, i);
    const T* qi = qkv + ri * ldq + h * hd;
    const T* ki = qi + D;
    const T* vi = ki + D;
    const T* doi = dout + ri * D + h * hd;
    const T* oi = o + ri * D + h * hd;
    float a[kMaxHd], b[kMaxHd];
    // ---- query role: dq_i ----
    float Di = 0.f;
    for (int c = 0; c < hd; ++c) { Di = fmaf(ld_f(doi + c), ld_f(oi + c), Di); a[c] = 0.f; }
    const float li = lse[ri * heads + h];
    for (int j = 0; j < sq.seq_len; ++j) {
      const T* kj = qkv + sq.row(s, j) * ldq + D + h * hd;
      const T* vj = kj + D;
      float sc = 0.f, dp = 0.f;
      for (int c = 0; c < hd; ++c) { sc = fmaf(ld_f(qi + c), ld_f(kj + c), sc); dp = fmaf(ld_f(doi + c), ld_f(vj + c), dp); }
      const float p = __expf(sc * scale - li);
      const float dsv = p * (dp - Di) * scale;
      for (int c = 0; c < hd; ++c) a[c] = fmaf(dsv, ld_f(kj + c), a[c]);
    }
    T* dq = dqkv + ri * ldq + h * hd;
    for (int c = 0; c < hd; ++c) st_f(dq + c, a[c]);
    // ---- key role: dk_i, dv_i ----
    for (int c = 0; c < hd; ++c) a[c] = b[c] = 0.f;
    for (int j = 0; j < sq.seq_len; ++j) {
      const int64_t rj = sq.row(s, j);
      const T* qj = qkv + rj * ldq + h * hd;
      const T* doj = dout + rj * D + h * hd;
      const T* oj = o + rj * D + h * hd;
      float sc = 0.f, dp = 0.f, Dj = 0.f;
      for (int c = 0; c < hd; ++c) {
        sc = fmaf(ld_f(qj + c), ld_f(ki + c), sc);
        dp = fmaf(ld_f(doj + c), ld_f(vi + c), dp);
        Dj = fmaf(ld_f(doj + c), ld_f(oj + c), Dj);
      }
      const float p = __expf(sc * scale - lse[rj * heads + h]);
      const float dsv = p * (dp - Dj) * scale;
      for (int c = 0; c < hd; ++c) { a[c] = fmaf(dsv, ld_f(qj + c), a[c]); b[c] = fmaf(p, ld_f(doj + c), b[c]); }
    }
    T* dk = dq + D;
    T* dv = dk + D;
    for (int c = 0; c < hd; ++c) { st_f(dk + c, a[c]); st_f(dv + c, b[c]); }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Plan {
  int SPC, NW, KP;
  size_t smem_fwd, smem_dq, smem_dkv;
};

template <int HD>
static Plan make_plan(int seq_len) {
  constexpr int LD = Cfg<HD>::LD;
  Plan p;
  p.KP = (seq_len + 15) / 16 * 16;
  const int QT = p.KP / 16;
  if (QT >= 8) {
    p.SPC = 1;
    const int rounds = (QT + 7) / 8;
    p.NW = (QT + rounds - 1) / rounds;
  } else {
    p.SPC = 8 / QT;
    p.NW = p.SPC * QT;
  }
  p.smem_fwd = (size_t)(2 * p.SPC * p.KP + p.NW * 16) * LD;
  p.smem_dq = (size_t)(2 * p.SPC * p.KP + p.NW * 32) * LD;
  p.smem_dkv = p.smem_dq + (size_t)2 * p.SPC * p.KP * sizeof(float);
  return p;
}

template <typename K>
static int raise_smem(K kern, size_t smem, const char* what) {
  if (smem > 232448) {
    set_error("%s: sequence too long for the shared-memory resident formulation (%zu B)", what, smem);
    return STGCMA_EINVAL;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("%s: smem attr: %s", what, cudaGetErrorString(e));
    return STGCMA_ECUDA;
  }
  return STGCMA_OK;
}

template <int HD>
static int fwd_bf16(const void* qkv, void* out, float* lse, Seq sq, int heads, cudaStream_t st) {
  const Plan p = make_plan<HD>(sq.seq_len);
  auto kern = mha_fwd_kernel<HD>;
  int rc = raise_smem(kern, p.smem_fwd, "mha_fwd");
  if (rc) return rc;
  dim3 grid((unsigned)ceil_div(sq.n_seq, p.SPC), (unsigned)heads);
  const float scale = 1.0f / sqrtf((float)HD);
  kern<<<grid, p.NW * 32, p.smem_fwd, st>>>((const bf16*)qkv, (bf16*)out, lse, sq, heads, p.SPC, p.KP, scale * kLog2e);
  return check_launch("mha_fwd(bf16)");
}

template <int HD>
static int bwd_bf16(const void* qkv, const void* o, const float* lse, const void* dout, void* dqkv, float* work, Seq sq,
                    int heads, cudaStream_t st) {
  const Plan p = make_plan<HD>(sq.seq_len);
  auto k1 = mha_bwd_dq_kernel<HD>;
  auto k2 = mha_bwd_dkv_kernel<HD>;
  int rc = raise_smem(k1, p.smem_dq, "mha_bwd(dq)");
  if (rc) return rc;
  rc = raise_smem(k2, p.smem_dkv, "mha_bwd(dkv)");
  if (rc) return rc;
  dim3 grid((unsigned)ceil_div(sq.n_seq, p.SPC), (unsigned)heads);
  const float scale = 1.0f / sqrtf((float)HD);
  k1<<<grid, p.NW * 32, p.smem_dq, st>>>((const bf16*)qkv, (const bf16*)o, lse, (const bf16*)dout, (bf16*)dqkv, work, sq, heads,
                                        p.SPC, p.KP, scale, scale * kLog2e);
  rc = check_launch("mha_bwd(dq, bf16)");
  if (rc) return rc;
  k2<<<grid, p.NW * 32, p.smem_dkv, st>>>((const bf16*)qkv, lse, (const bf16*)dout, work, (bf16*)dqkv, sq, heads, p.SPC, p.KP,
                                         scale, scale * kLog2e);
  return check_launch("mha_bwd(dkv, bf16)");
}

static int check_args(const char* what, int64_t n_seq, int seq_len, int inner, int heads, int hd) {
  STG_REQUIRE(n_seq > 0 && seq_len > 0 && inner > 0 && heads > 0, "%s: empty problem", what);
  STG_REQUIRE(n_seq % inner == 0, "%s: n_seq=%lld must be a multiple of inner=%d", what, (long long)n_seq, inner);
  STG_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= kMaxHd, "%s: head_dim %d not in {16, 32, .., 128}", what, hd);
  STG_REQUIRE(n_seq < (1ll << 31) && heads < 65536, "%s: grid too large", what);
  return STGCMA_OK;
}

}  // namespace mha

// temporal_tc.cu: whole-sequence-in-one-tile kernels for T <= 16 frames of a [clips, T, tokens] layout (-1: not covered)
int temporal_mha_fwd_bf16(const void* qkv, void* out, float* lse, int B, int T, int L, int heads, int hd, cudaStream_t st);
int temporal_mha_bwd_bf16(const void* qkv, const void* dout, void* dqkv, int B, int T, int L, int heads, int hd, cudaStream_t st);
// mha_tc.cu: tcgen05 / TMEM kernels for contiguous sequences of 65 .. 208 tokens (-1: not covered)
int mha_tc_fwd_bf16(const void* qkv, void* out, float* lse, int64_t n_seq, int seq_len, int heads, int hd, cudaStream_t st);
int mha_tc_bwd_bf16(const void* qkv, const void* o, const float* lse, const void* dout, void* dqkv, int64_t n_seq, int seq_len,
                    int heads, int hd, cudaStream_t st);
}  // namespace stgcma

using namespace stgcma;

// head dims with a bf16 instantiation: 16 (test models), 32, 64 (ViT-L/14: 1024 / 16; ViT-B/16 with its native 12 heads),
// 96 (ViT-B/16 as the reference instantiates it: 8 heads, run_adapt_ave29.py:137), 128
#define MHA_DISPATCH(HD_, CALL, ...)                                        \
  switch (HD_) {                                                            \
    case 16: return mha::CALL<16>(__VA_ARGS__);                             \
    case 32: return mha::CALL<32>(__VA_ARGS__);                             \
    case 64: return mha::CALL<64>(__VA_ARGS__);                             \
    case 96: return mha::CALL<96>(__VA_ARGS__);                             \
    case 128: return mha::CALL<128>(__VA_ARGS__);                           \
    default: STG_REQUIRE(false, "mha(bf16): head_dim %d not in {16, 32, 64, 96, 128}", HD_); \
  }

extern "C" int stgcma_mha_fwd(const void* qkv, void* out, float* lse, int64_t n_seq, int seq_len, int inner, int heads, int hd,
                              int dtype, void* stream) {
  int rc = mha::check_args("mha_fwd", n_seq, seq_len, inner, heads, hd);
  if (rc) return rc;
  STG_REQUIRE(qkv && out && lse, "mha_fwd: null operand");
  cudaStream_t st = (cudaStream_t)stream;
  const mha::Seq sq{n_seq, seq_len, inner};
  if (dtype == STGCMA_F32) {
    dim3 grid((unsigned)n_seq, (unsigned)heads);
    mha::mha_fwd_simt_kernel<float><<<grid, 128, 0, st>>>((const float*)qkv, (float*)out, lse, sq, heads, hd, 1.0f / sqrtf((float)hd));
    return check_launch("mha_fwd(fp32)");
  }
  STG_REQUIRE(dtype == STGCMA_BF16, "mha_fwd: bad dtype %d", dtype);
  STG_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "mha_fwd: operands must be 16-byte aligned");
  if (inner > 1 && seq_len <= 16 && n_seq / inner < (1 << 30)) {
    // temporal attention: an HBM stream with a 16 x 16 score tile per (token, head) -- the tile kernels of temporal_tc.cu
    rc = temporal_mha_fwd_bf16(qkv, out, lse, (int)(n_seq / inner), seq_len, inner, heads, hd, st);
    if (rc != -1) return rc;
  }
  if (inner == 1 && !getenv("STGCMA_MHA_NO_TC")) {
    rc = mha_tc_fwd_bf16(qkv, out, lse, n_seq, seq_len, heads, hd, st);
    if (rc != -1) return rc;
  }
  MHA_DISPATCH(hd, fwd_bf16, qkv, out, lse, sq, heads, st);
}

extern "C" int stgcma_mha_bwd(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, float* work,
                              int64_t n_seq, int seq_len, int inner, int heads, int hd, int dtype, void* stream) {
  int rc = mha::check_args("mha_bwd", n_seq, seq_len, inner, heads, hd);
  if (rc) return rc;
  STG_REQUIRE(qkv && out && lse && dout && dqkv, "mha_bwd: null operand");
  cudaStream_t st = (cudaStream_t)stream;
  const mha::Seq sq{n_seq, seq_len, inner};
  if (dtype == STGCMA_F32) {
    dim3 grid((unsigned)n_seq, (unsigned)heads);
    mha::mha_bwd_simt_kernel<float><<<grid, 128, 0, st>>>((const float*)qkv, (const float*)out, lse, (const float*)dout,
                                                          (float*)dqkv, sq, heads, hd, 1.0f / sqrtf((float)hd));
    return check_launch("mha_bwd(fp32)");
  }
  STG_REQUIRE(dtype == STGCMA_BF16, "mha_bwd: bad dtype %d", dtype);
  STG_REQUIRE(work, "mha_bwd(bf16): needs the [rows, heads] fp32 work buffer");
  STG_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)dout & 15) == 0 && ((uintptr_t)dqkv & 15) == 0,
              "mha_bwd: operands must be 16-byte aligned");
  if (inner > 1 && seq_len <= 16 && n_seq / inner < (1 << 30)) {
    rc = temporal_mha_bwd_bf16(qkv, dout, dqkv, (int)(n_seq / inner), seq_len, inner, heads, hd, st);
    if (rc != -1) return rc;
  }
  if (inner == 1 && !getenv("STGCMA_MHA_NO_TC")) {
    rc = mha_tc_bwd_bf16(qkv, out, lse, dout, dqkv, n_seq, seq_len, heads, hd, st);
    if (rc != -1) return rc;
  }
  MHA_DISPATCH(hd, bwd_bf16, qkv, out, lse, dout, dqkv, work, sq, heads, st);
}
