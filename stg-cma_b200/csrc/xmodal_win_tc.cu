// Windowed cross-modal token attention of the S-adapters (Swin_AVE.py:753-760), bf16 production kernels.
//
//   per (frame, window w):   hv' = hv + gate_v * softmax(hv ha^T) ha        ha' = ha + gate_a * softmax(ha hv^T) hv
//
// One CTA (8 warps) owns one window pair: the <= 49 video and <= 49 audio token rows (gathered through the shifted
// window map, padded to 64 zero rows) sit in shared memory; warps 0-3 are the video-query direction, warps 4-7 the
// audio-query direction, 16 query rows per warp.  Scores, probabilities and accumulators live in registers
// (mma.sync m16n8k16, mma.cuh); S^T is simply recomputed by the other direction's warps (the flops are negligible:
// 0.2 % of the step), so there is no column softmax and no exchange between warps except the row statistics of the
// backward.  This replaces a CUDA-core kernel that kept S, P and dS in fp32 shared memory (r01b: 11 % of the step).
//   backward ("row owner", same algebra as xmodal_tc.cu):
//       dh_i = dout_i + sum_j [ (dS1_ij + dS2_ij) h_j + g_o P2_ij dout_j ]
//       dS1_ij = P1_ij (g_z <dout_i, h_j> - D_i)     P1_ij = exp(s_ij - lse_z[i])
//       dS2_ij = P2_ij (g_o <dout_j, h_i> - D_j)     P2_ij = exp(s_ij - lse_o[j])      D = gate * <dout, attn_out>
//   lse and D of all 2 x 64 rows are recomputed inside the CTA (phase 1) and exchanged through shared memory.
#include "mma.cuh"
#include "gelu_poly.cuh"

namespace stgcma {
namespace xwtc {

using namespace mma;

constexpr int NR = 64;   // padded rows per modality

template <int RP>
struct Cfg {
  static constexpr int LD = RP * 2 + 16;  // smem row stride in bytes (odd multiple of 16: conflict-free ldmatrix)
  static constexpr int KS = RP / 16;
  static constexpr int CH = RP / 8;       // 16-byte chunks per row
};

__device__ __forceinline__ int win_token(int wy, int wx, int i, int ws, int shift, int R) {
  const int iy = i / ws, ix = i - iy * ws;
  return win_src(wy, iy, ws, shift, R) * R + win_src(wx, ix, ws, shift, R);
}

// scores of this warp's 16 own rows against the 64 rows of `oth`, in the log2 domain, invalid columns = -inf
template <int RP>
__device__ __forceinline__ void scores16(float (&s)[8][4], const uint32_t (&qa)[Cfg<RP>::KS][4], const uint8_t* oth, int N,
                                         int lane) {
  constexpr int LD = Cfg<RP>::LD, KS = Cfg<RP>::KS;
  const int t = lane & 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t b[4];
      load_b_nk(b, oth, LD, n2 * 16, ks * 16, lane);
      mma16816(s[2 * n2], qa[ks], b[0], b[1]);
      mma16816(s[2 * n2 + 1], qa[ks], b[2], b[3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = i * 8 + 2 * t + (e & 1);
      s[i][e] = col < N ? s[i][e] * kLog2e : -INFINITY;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int RP>
__global__ void __launch_bounds__(256) xw_fwd_kernel(const bf16* __restrict__ h, const float* __restrict__ gate,
                                                     bf16* __restrict__ out, bf16* __restrict__ attn_out, int64_t frames, int R,
                                                     int ws, int shift) {
  constexpr int LD = Cfg<RP>::LD, KS = Cfg<RP>::KS, CH = Cfg<RP>::CH;
  extern __shared__ __align__(16) uint8_t xsm[];
  uint8_t* Hs = xsm;                      // [2][64][LD]
  uint8_t* Os = Hs + 2 * NR * LD;         // [2][64][LD]  out staging
  uint8_t* As = Os + 2 * NR * LD;         // [2][64][LD]  attn_out staging
  int* tok = reinterpret_cast<int*>(As + 2 * NR * LD);
  const int N = ws * ws, L = R * R;
  const int nwr = R / ws, nW = nwr * nwr;
  const int w = (int)(blockIdx.x % nW);
  const int64_t frame = blockIdx.x / nW;
  const int wy = w / nwr, wx = w - wy * nwr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  if (tid < NR) tok[tid] = tid < N ? win_token(wy, wx, tid, ws, shift, R) : 0;
  __syncthreads();
  const int64_t mod_stride = frames * (int64_t)L * RP;
  for (int e = tid; e < 2 * NR * CH; e += 256) {
    const int mod = e / (NR * CH), rem = e - mod * NR * CH, r = rem / CH, c = rem - r * CH;
    uint8_t* dst = Hs + (mod * NR + r) * LD + c * 16;
    if (r < N) cp_async16(dst, h + mod * mod_stride + (frame * L + tok[r]) * (int64_t)RP + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int z = warp >> 2, row0 = (warp & 3) * 16;
  const uint8_t* own = Hs + z * NR * LD;
  const uint8_t* oth = Hs + (1 - z) * NR * LD;
  if (row0 < N) {
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) load_a(qa[ks], own, LD, row0, ks * 16, lane);
    float s[8][4];
    scores16<RP>(s, qa, oth, N, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = ex2(s[i][0] - mx0);
      s[i][1] = ex2(s[i][1] - mx0);
      s[i][2] = ex2(s[i][2] - mx1);
      s[i][3] = ex2(s[i][3] - mx1);
      l0 += s[i][0] + s[i][1];
      l1 += s[i][2] + s[i][3];
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    float acc[2 * KS][4];
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < KS; ++n2) {
        uint32_t b[4];
        load_b_kn(b, oth, LD, kk * 16, n2 * 16, lane);
        mma16816(acc[2 * n2], pa, b[0], b[1]);
        mma16816(acc[2 * n2 + 1], pa, b[2], b[3]);
      }
    }
    const float gz = gate[z], inv0 = 1.f / l0, inv1 = 1.f / l1;
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) {
      const int col = i * 8 + 2 * t;
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int lr = row0 + g + hr * 8;
        const float inv = hr ? inv1 : inv0;
        const float a0 = acc[i][2 * hr] * inv, a1 = acc[i][2 * hr + 1] * inv;
        const __nv_bfloat162 o = *reinterpret_cast<const __nv_bfloat162*>(own + lr * LD + col * 2);
        *reinterpret_cast<uint32_t*>(As + (z * NR + lr) * LD + col * 2) = pack_bf16(a0, a1);
        *reinterpret_cast<uint32_t*>(Os + (z * NR + lr) * LD + col * 2) =
            pack_bf16(__bfloat162float(o.x) + gz * a0, __bfloat162float(o.y) + gz * a1);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < 2 * NR * CH; e += 256) {
    const int mod = e / (NR * CH), rem = e - mod * NR * CH, r = rem / CH, c = rem - r * CH;
    if (r >= N) continue;
    const int64_t off = mod * mod_stride + (frame * L + tok[r]) * (int64_t)RP + c * 8;
    *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(Os + (mod * NR + r) * LD + c * 16);
    *reinterpret_cast<uint4*>(attn_out + off) = *reinterpret_cast<const uint4*>(As + (mod * NR + r) * LD + c * 16);
  }
}

// ---------------------------------------------------------------------------------------------
template <int RP>
__global__ void __launch_bounds__(256, (RP <= 16 ? 4 : RP <= 32 ? 3 : 0)) xw_bwd_kernel(const bf16* __restrict__ h, const bf16* __restrict__ hpre,
                                                     const float* __restrict__ gate, const bf16* __restrict__ attn_out,
                                                     const bf16* __restrict__ dout, bf16* __restrict__ dh,
                                                     float* __restrict__ dgate, int64_t frames, int R, int ws, int shift) {
  constexpr int LD = Cfg<RP>::LD, KS = Cfg<RP>::KS, CH = Cfg<RP>::CH;
  extern __shared__ __align__(16) uint8_t xsm[];
  uint8_t* Hs = xsm;                      // [2][64][LD]
  uint8_t* Ds = Hs + 2 * NR * LD;         // [2][64][LD]  dout
  uint8_t* Os = Ds + 2 * NR * LD;         // [2][64][LD]  dh staging
  uint8_t* Ps = Os + 2 * NR * LD;         // [2][64][LD]  hpre (pre-GELU features) of the own rows
  float* lse_s = reinterpret_cast<float*>(Ps + 2 * NR * LD);   // [2][64] log2 domain, +inf on padding rows
  float* D_s = lse_s + 2 * NR;                                   // [2][64] gate * <dout, attn_out>
  float* dot_s = D_s + 2 * NR;                                   // [2][64] <dout, attn_out>
  // per-chunk partial dot products, summed in a FIXED order below: fp32 atomics here made D (and through the bf16 rounding
  // of dS every gradient upstream of this sub-block) differ from run to run in the last bit -- harmless for the adapters,
  // but the temporal bias tables amplify it to 0.5 - 0.9 % (tests/test_step_gpu.py::test_full_width_properties)
  float* part_s = dot_s + 2 * NR;                                // [2][64][CH]
  int* tok = reinterpret_cast<int*>(part_s + 2 * NR * CH);
  const int N = ws * ws, L = R * R;
  const int nwr = R / ws, nW = nwr * nwr;
  const int w = (int)(blockIdx.x % nW);
  const int64_t frame = blockIdx.x / nW;
  const int wy = w / nwr, wx = w - wy * nwr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  if (tid < NR) tok[tid] = tid < N ? win_token(wy, wx, tid, ws, shift, R) : 0;
  __syncthreads();
  const int64_t mod_stride = frames * (int64_t)L * RP;
  for (int e = tid; e < 2 * NR * CH; e += 256) {
    const int mod = e / (NR * CH), rem = e - mod * NR * CH, r = rem / CH, c = rem - r * CH;
    uint8_t* hd = Hs + (mod * NR + r) * LD + c * 16;
    uint8_t* dd = Ds + (mod * NR + r) * LD + c * 16;
    if (r < N) {
      const int64_t off = mod * mod_stride + (frame * L + tok[r]) * (int64_t)RP + c * 8;
      cp_async16(hd, h + off);
      if (hpre) cp_async16(Ps + (mod * NR + r) * LD + c * 16, hpre + off);
      const uint4 dv = *reinterpret_cast<const uint4*>(dout + off);
      const uint4 av = *reinterpret_cast<const uint4*>(attn_out + off);
      *reinterpret_cast<uint4*>(dd) = dv;
      float d8[8], a8[8];
      unpack8(dv, d8);
      unpack8(av, a8);
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(d8[i], a8[i], dot);
      part_s[(mod * NR + r) * CH + c] = dot;
    } else {
      *reinterpret_cast<uint4*>(hd) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dd) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int z = warp >> 2, row0 = (warp & 3) * 16;
  const uint8_t* own_h = Hs + z * NR * LD;
  const uint8_t* own_d = Ds + z * NR * LD;
  const uint8_t* ch = Hs + (1 - z) * NR * LD;
  const uint8_t* cd = Ds + (1 - z) * NR * LD;
  const float g_own = gate[z], g_oth = gate[1 - z];
  if (tid < 2 * NR) {
    float d = 0.f;
    if ((tid & (NR - 1)) < N)
#pragma unroll
      for (int c = 0; c < CH; ++c) d += part_s[tid * CH + c];
    dot_s[tid] = d;
    D_s[tid] = gate[tid / NR] * d;
  }

  uint32_t ha[KS][4], da[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    load_a(ha[ks], own_h, LD, row0, ks * 16, lane);
    load_a(da[ks], own_d, LD, row0, ks * 16, lane);
  }
  // ---- phase 1: row statistics of the own rows ----
  float lse0, lse1;
  {
    float s[8][4];
    scores16<RP>(s, ha, ch, N, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      l0 += ex2(s[i][0] - mx0) + ex2(s[i][1] - mx0);
      l1 += ex2(s[i][2] - mx1) + ex2(s[i][3] - mx1);
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const int ra = row0 + g, rb = ra + 8;
    lse0 = ra < N ? mx0 + log2f(l0) : INFINITY;
    lse1 = rb < N ? mx1 + log2f(l1) : INFINITY;
    if (t == 0) {
      lse_s[z * NR + ra] = lse0;
      lse_s[z * NR + rb] = lse1;
    }
  }
  __syncthreads();
  if (warp < 2) {   // dgate[z] += sum_rows <dout, attn_out>
    float v = dot_s[warp * NR + lane] + dot_s[warp * NR + lane + 32];
    v = warp_sum(v);
    if (lane == 0) atomicAdd(dgate + warp, v);
  }
  const float* col_lse = lse_s + (1 - z) * NR;
  const float* col_D = D_s + (1 - z) * NR;
  const float D0 = D_s[z * NR + row0 + g], D1 = D_s[z * NR + row0 + g + 8];

  // ---- phase 2 ----
  float acc[2 * KS][4];
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  if (row0 < N) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float s[4][4], p1[4][4], p2[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[i][e] = p1[i][e] = p2[i][e] = 0.f;
#pragma unroll
      for (int n2 = 0; n2 < 2; ++n2) {
        const int n0 = half * 32 + n2 * 16;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t bh[4], bd[4];
          load_b_nk(bh, ch, LD, n0, ks * 16, lane);
          load_b_nk(bd, cd, LD, n0, ks * 16, lane);
          mma16816(s[2 * n2], ha[ks], bh[0], bh[1]);        // s   = <h_i, h_j>
          mma16816(s[2 * n2 + 1], ha[ks], bh[2], bh[3]);
          mma16816(p1[2 * n2], da[ks], bh[0], bh[1]);       // dp1 = <dout_i, h_j>
          mma16816(p1[2 * n2 + 1], da[ks], bh[2], bh[3]);
          mma16816(p2[2 * n2], ha[ks], bd[0], bd[1]);       // dp2 = <h_i, dout_j>
          mma16816(p2[2 * n2 + 1], ha[ks], bd[2], bd[3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int jc = half * 32 + i * 8 + 2 * t;
        const float2 cl = *reinterpret_cast<const float2*>(col_lse + jc);
        const float2 cD = *reinterpret_cast<const float2*>(col_D + jc);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sl = s[i][e] * kLog2e;
          const float P1 = ex2(sl - ((e & 2) ? lse1 : lse0));
          const float P2 = ex2(sl - ((e & 1) ? cl.y : cl.x));
          const float wv = P1 * (g_own * p1[i][e] - ((e & 2) ? D1 : D0)) + P2 * (g_oth * p2[i][e] - ((e & 1) ? cD.y : cD.x));
          s[i][e] = wv;
          p2[i][e] = g_oth * P2;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t wa[4], qa[4];
        wa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        wa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        wa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        wa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        qa[0] = pack_bf16(p2[2 * kk][0], p2[2 * kk][1]);
        qa[1] = pack_bf16(p2[2 * kk][2], p2[2 * kk][3]);
        qa[2] = pack_bf16(p2[2 * kk + 1][0], p2[2 * kk + 1][1]);
        qa[3] = pack_bf16(p2[2 * kk + 1][2], p2[2 * kk + 1][3]);
        const int k0 = half * 32 + kk * 16;
#pragma unroll
        for (int n2 = 0; n2 < KS; ++n2) {
          uint32_t bh[4], bd[4];
          load_b_kn(bh, ch, LD, k0, n2 * 16, lane);
          load_b_kn(bd, cd, LD, k0, n2 * 16, lane);
          mma16816(acc[2 * n2], wa, bh[0], bh[1]);          // += (dS1 + dS2) h_j
          mma16816(acc[2 * n2 + 1], wa, bh[2], bh[3]);
          mma16816(acc[2 * n2], qa, bd[0], bd[1]);          // += g_oth P2 dout_j
          mma16816(acc[2 * n2 + 1], qa, bd[2], bd[3]);
        }
      }
    }
  }
  // ---- dh = (dout_own + acc) * gelu'(hpre) -> staging -> coalesced 16-byte stores ----
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) {
    const int col = i * 8 + 2 * t;
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int lr = row0 + g + hr * 8;
      const __nv_bfloat162 dd = *reinterpret_cast<const __nv_bfloat162*>(own_d + lr * LD + col * 2);
      float v0 = __bfloat162float(dd.x) + acc[i][2 * hr], v1 = __bfloat162float(dd.y) + acc[i][2 * hr + 1];
      if (hpre) {
        const __nv_bfloat162 hp = *reinterpret_cast<const __nv_bfloat162*>(Ps + (z * NR + lr) * LD + col * 2);
        v0 *= gelu_grad_poly(__bfloat162float(hp.x));
        v1 *= gelu_grad_poly(__bfloat162float(hp.y));
      }
      *reinterpret_cast<uint32_t*>(Os + (z * NR + lr) * LD + col * 2) = pack_bf16(v0, v1);
    }
  }
  __syncthreads();
  for (int e = tid; e < 2 * NR * CH; e += 256) {
    const int mod = e / (NR * CH), rem = e - mod * NR * CH, r = rem / CH, c = rem - r * CH;
    if (r >= N) continue;
    const int64_t off = mod * mod_stride + (frame * L + tok[r]) * (int64_t)RP + c * 8;
    *reinterpret_cast<uint4*>(dh + off) = *reinterpret_cast<const uint4*>(Os + (mod * NR + r) * LD + c * 16);
  }
}

template <int RP>
static int fwd_launch(const void* h, const float* gate, void* out, void* attn_out, int64_t frames, int R, int ws, int shift,
                      cudaStream_t st) {
  constexpr int LD = Cfg<RP>::LD;
  const size_t smem = (size_t)6 * NR * LD + NR * sizeof(int);
  auto kern = xw_fwd_kernel<RP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("xmodal_window_fwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws));
  kern<<<grid, 256, smem, st>>>((const bf16*)h, gate, (bf16*)out, (bf16*)attn_out, frames, R, ws, shift);
  return check_launch("xmodal_window_fwd(tc)");
}

template <int RP>
static int bwd_launch(const void* h, const void* hpre, const float* gate, const void* attn_out, const void* dout, void* dh,
                      float* dgate, int64_t frames, int R, int ws, int shift, cudaStream_t st) {
  constexpr int LD = Cfg<RP>::LD;
  const size_t smem = (size_t)8 * NR * LD + (6 + 2 * Cfg<RP>::CH) * NR * sizeof(float) + NR * sizeof(int);
  auto kern = xw_bwd_kernel<RP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("xmodal_window_bwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws));
  kern<<<grid, 256, smem, st>>>((const bf16*)h, (const bf16*)hpre, gate, (const bf16*)attn_out, (const bf16*)dout, (bf16*)dh,
                                dgate, frames, R, ws, shift);
  return check_launch("xmodal_window_bwd(tc)");
}

}  // namespace xwtc

// entry points used by xmodal.cu's bf16 dispatch; return -1 for an unsupported rp
int xmodal_window_fwd_bf16(const void* h, const float* gate, void* out, void* attn_out, int64_t frames, int R, int ws,
                           int shift, int rp, cudaStream_t st) {
  switch (rp) {
    case 16: return xwtc::fwd_launch<16>(h, gate, out, attn_out, frames, R, ws, shift, st);
    case 32: return xwtc::fwd_launch<32>(h, gate, out, attn_out, frames, R, ws, shift, st);
    case 48: return xwtc::fwd_launch<48>(h, gate, out, attn_out, frames, R, ws, shift, st);
    case 64: return xwtc::fwd_launch<64>(h, gate, out, attn_out, frames, R, ws, shift, st);
    case 96: return xwtc::fwd_launch<96>(h, gate, out, attn_out, frames, R, ws, shift, st);
    default: return -1;
  }
}

int xmodal_window_bwd_bf16(const void* h, const void* hpre, const float* gate, const void* attn_out, const void* dout,
                           void* dh, float* dgate, int64_t frames, int R, int ws, int shift, int rp, cudaStream_t st) {
  switch (rp) {
    case 16: return xwtc::bwd_launch<16>(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, st);
    case 32: return xwtc::bwd_launch<32>(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, st);
    case 48: return xwtc::bwd_launch<48>(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, st);
    case 64: return xwtc::bwd_launch<64>(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, st);
    case 96: return xwtc::bwd_launch<96>(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, st);
    default: return -1;
  }
}

}  // namespace stgcma
