// Global cross-modal token attention of the FFN adapters (Swin_AVE.py:801-808), bf16 production kernels.
//
//   hv' = hv + gate_v * softmax(hv ha^T) ha         ha' = ha + gate_a * softmax(ha hv^T) hv
//
// Flash formulation on the warp-level tensor cores (mma.cuh explains why mma.sync and not tcgen05 for d = r):
// a CTA owns NW*16 rows of one modality of one frame (one warp = 16 rows), streams the other modality's rows
// through a double-buffered cp.async pipeline in tiles of 64, and keeps scores, probabilities and the output
// accumulator in registers.  The L x L score matrix (3136^2 per frame at stage 0; the reference materialises
// it, 787 MB per block per clip) never exists.
//   forward : out, attn_out, lse (natural log, one per row and direction)
//   backward: "row owner" scheme -- the CTA that owns rows i of modality z accumulates BOTH roles of those rows
//             (queries of direction z, keys/values of direction 1-z) while streaming modality 1-z, so dh needs
//             no atomics:   dh_i = dout_i + sum_j [ (dS1_ij + dS2_ij) h_j + g_o P2_ij dout_j ]
//             dS1_ij = P1_ij (g_z <dout_i, h_j> - D_i)   P1_ij = exp(s_ij - lse_z[i])      (i queries j)
//             dS2_ij = P2_ij (g_o <dout_j, h_i> - D_j)   P2_ij = exp(s_ij - lse_o[j])      (j queries i)
//             with D = gate * <dout, attn_out> per row (xmodal_prep_kernel in xmodal.cu), s_ij = <h_i, h_j>.
#include "mma.cuh"

namespace stgcma {
namespace xtc {

using namespace mma;

constexpr int BN = 64;  // streamed rows per tile

template <int RP>
struct Cfg {
  static constexpr int LD = RP * 2 + 16;  // smem row stride in bytes: an odd multiple of 16 -> conflict-free ldmatrix
  static constexpr int KS = RP / 16;      // k-steps over the feature dim
  static constexpr int CH = RP / 8;       // 16-byte chunks per row
};

// rows [row0, row0 + nrows) of a [L][RP] bf16 matrix -> smem tile (rows >= L are zero-filled)
template <int RP>
__device__ __forceinline__ void load_rows_async(uint8_t* tile, const bf16* __restrict__ src, int row0, int nrows, int L,
                                                int tid, int nthr) {
  constexpr int CH = Cfg<RP>::CH, LD = Cfg<RP>::LD;
  for (int e = tid; e < nrows * CH; e += nthr) {
    const int r = e / CH, c = e - r * CH;
    uint8_t* dst = tile + r * LD + c * 16;
    if (row0 + r < L) cp_async16(dst, src + (int64_t)(row0 + r) * RP + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int RP, int NW>
__global__ void __launch_bounds__(NW * 32) xg_fwd_kernel(const bf16* __restrict__ h, const float* __restrict__ gate,
                                                         bf16* __restrict__ out, bf16* __restrict__ attn_out,
                                                         float* __restrict__ lse, int64_t frames, int L, int La) {
  constexpr int LD = Cfg<RP>::LD, KS = Cfg<RP>::KS, BM = NW * 16;
  extern __shared__ __align__(16) uint8_t xsm[];
  uint8_t* own = xsm;                 // [BM][LD]
  uint8_t* oth = own + BM * LD;       // 3 x [BN][LD]  (ring: one block-wide barrier per streamed tile instead of two)
  const int z = blockIdx.z;
  const int64_t frame = blockIdx.y;
  const int r0 = blockIdx.x * BM;
  // rows per frame: L for the first modality (video), La for the second (audio; = L in the Swin model, 197 / 49 in CLIP)
  const int Lown = z ? La : L, Loth = z ? L : La;
  if (r0 >= Lown) return;
  const int64_t row_own0 = (z ? frames * (int64_t)L : 0) + frame * (int64_t)Lown;
  const int64_t row_oth0 = (z ? 0 : frames * (int64_t)L) + frame * (int64_t)Loth;
  const bf16* hown = h + row_own0 * RP;
  const bf16* hoth = h + row_oth0 * RP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nkt = (Loth + BN - 1) / BN;

  load_rows_async<RP>(own, hown, r0, BM, Lown, tid, NW * 32);
  load_rows_async<RP>(oth, hoth, 0, BN, Loth, tid, NW * 32);
  cp_async_commit();
  if (nkt > 1) load_rows_async<RP>(oth + BN * LD, hoth, BN, BN, Loth, tid, NW * 32);
  cp_async_commit();

  uint32_t qa[KS][4];
  float acc[2 * KS][4];
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};  // log2 domain; l = per-thread partial sums

  for (int kt = 0; kt < nkt; ++kt) {
    uint8_t* cur = oth + (kt % 3) * BN * LD;
    cp_async_wait<1>();   // tile kt has landed (tile kt + 1 may still be in flight)
    __syncthreads();      // ... for every thread, and every warp is done with tile kt - 1, whose buffer is refilled now
    if (kt + 2 < nkt) load_rows_async<RP>(oth + ((kt + 2) % 3) * BN * LD, hoth, (kt + 2) * BN, BN, Loth, tid, NW * 32);
    cp_async_commit();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) load_a(qa[ks], own, LD, warp * 16, ks * 16, lane);
    }
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t b[4];
        load_b_nk(b, cur, LD, n2 * 16, ks * 16, lane);
        mma16816(s[2 * n2], qa[ks], b[0], b[1]);
        mma16816(s[2 * n2 + 1], qa[ks], b[2], b[3]);
      }
    }
    const int jbase = kt * BN;
    const bool tail = jbase + BN > Loth;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = s[i][e] * kLog2e;
        if (tail && jbase + i * 8 + 2 * t + (e & 1) >= Loth) v = -INFINITY;
        s[i][e] = v;
      }
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
    const float c0 = ex2(m_run[0] - mn0), c1 = ex2(m_run[1] - mn1);  // ex2(-inf) = 0 on the first tile
    m_run[0] = mn0;
    m_run[1] = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = ex2(s[i][0] - mn0);
      s[i][1] = ex2(s[i][1] - mn0);
      s[i][2] = ex2(s[i][2] - mn1);
      s[i][3] = ex2(s[i][3] - mn1);
      ps0 += s[i][0] + s[i][1];
      ps1 += s[i][2] + s[i][3];
    }
    l_run[0] = l_run[0] * c0 + ps0;
    l_run[1] = l_run[1] * c1 + ps1;
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) {
      acc[i][0] *= c0;
      acc[i][1] *= c0;
      acc[i][2] *= c1;
      acc[i][3] *= c1;
    }
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
        load_b_kn(b, cur, LD, kk * 16, n2 * 16, lane);
        mma16816(acc[2 * n2], pa, b[0], b[1]);
        mma16816(acc[2 * n2 + 1], pa, b[2], b[3]);
      }
    }
  }

  const float gz = gate[z];
  const float l0 = quad_sum(l_run[0]), l1 = quad_sum(l_run[1]);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int ra = r0 + warp * 16 + g, rb = ra + 8;
  const int64_t base = row_own0 * RP;
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) {
    const int col = i * 8 + 2 * t;
    if (ra < Lown) {
      const float a0 = acc[i][0] * inv0, a1 = acc[i][1] * inv0;
      const __nv_bfloat162 o = *reinterpret_cast<const __nv_bfloat162*>(own + (warp * 16 + g) * LD + col * 2);
      *reinterpret_cast<uint32_t*>(attn_out + base + (int64_t)ra * RP + col) = pack_bf16(a0, a1);
      *reinterpret_cast<uint32_t*>(out + base + (int64_t)ra * RP + col) =
          pack_bf16(__bfloat162float(o.x) + gz * a0, __bfloat162float(o.y) + gz * a1);
    }
    if (rb < Lown) {
      const float a0 = acc[i][2] * inv1, a1 = acc[i][3] * inv1;
      const __nv_bfloat162 o = *reinterpret_cast<const __nv_bfloat162*>(own + (warp * 16 + g + 8) * LD + col * 2);
      *reinterpret_cast<uint32_t*>(attn_out + base + (int64_t)rb * RP + col) = pack_bf16(a0, a1);
      *reinterpret_cast<uint32_t*>(out + base + (int64_t)rb * RP + col) =
          pack_bf16(__bfloat162float(o.x) + gz * a0, __bfloat162float(o.y) + gz * a1);
    }
  }
  if (t == 0) {
    float* lp = lse + row_own0;
    if (ra < Lown) lp[ra] = (m_run[0] + log2f(l0)) * 0.6931471805599453f;
    if (rb < Lown) lp[rb] = (m_run[1] + log2f(l1)) * 0.6931471805599453f;
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int RP, int NW>
__global__ void __launch_bounds__(NW * 32, (RP <= 16 ? 8 : 0)) xg_bwd_kernel(const bf16* __restrict__ h, const bf16* __restrict__ hpre,
                                                         const float* __restrict__ gate, const float* __restrict__ lse,
                                                         const float* __restrict__ D, const bf16* __restrict__ dout,
                                                         bf16* __restrict__ dh, int64_t frames, int L, int La) {
  constexpr int LD = Cfg<RP>::LD, KS = Cfg<RP>::KS, BM = NW * 16;
  extern __shared__ __align__(16) uint8_t xsm[];
  uint8_t* own_h = xsm;                       // [BM][LD]
  uint8_t* own_d = own_h + BM * LD;           // [BM][LD]  dout of the own rows
  uint8_t* oth_h = own_d + BM * LD;           // 3 x [BN][LD]  (ring of 3: one block-wide barrier per streamed tile)
  uint8_t* oth_d = oth_h + 3 * BN * LD;       // 3 x [BN][LD]
  float* col_lse = reinterpret_cast<float*>(oth_d + 3 * BN * LD);  // 3 x [BN]  lse (log2 domain) of the streamed rows
  float* col_D = col_lse + 3 * BN;                                  // 3 x [BN]
  const int z = blockIdx.z;
  const int64_t frame = blockIdx.y;
  const int r0 = blockIdx.x * BM;
  const int Lown = z ? La : L, Loth = z ? L : La;
  if (r0 >= Lown) return;
  const int64_t row_own0 = (z ? frames * (int64_t)L : 0) + frame * (int64_t)Lown;
  const int64_t row_oth0 = (z ? 0 : frames * (int64_t)L) + frame * (int64_t)Loth;
  const int64_t off_own = row_own0 * RP, off_oth = row_oth0 * RP;
  const float g_own = gate[z], g_oth = gate[1 - z];
  const float* lse_own_p = lse + row_own0;
  const float* lse_oth_p = lse + row_oth0;
  const float* D_own_p = D + row_own0;
  const float* D_oth_p = D + row_oth0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nkt = (Loth + BN - 1) / BN;

  auto load_cols = [&](int kt) {
    const int buf = kt % 3;
    load_rows_async<RP>(oth_h + buf * BN * LD, h + off_oth, kt * BN, BN, Loth, tid, NW * 32);
    load_rows_async<RP>(oth_d + buf * BN * LD, dout + off_oth, kt * BN, BN, Loth, tid, NW * 32);
    if (tid < BN) {
      const int j = kt * BN + tid;
      const bool ok = j < Loth;
      // invalid columns get lse = +inf -> P2 = exp2(s - inf) = 0
      col_lse[buf * BN + tid] = ok ? lse_oth_p[j] * kLog2e : INFINITY;
      col_D[buf * BN + tid] = ok ? D_oth_p[j] : 0.f;
    }
  };

  load_rows_async<RP>(own_h, h + off_own, r0, BM, Lown, tid, NW * 32);
  load_rows_async<RP>(own_d, dout + off_own, r0, BM, Lown, tid, NW * 32);
  load_cols(0);
  cp_async_commit();
  if (nkt > 1) load_cols(1);
  cp_async_commit();

  const int ra = r0 + warp * 16 + g, rb = ra + 8;
  // invalid own rows: lse = +inf -> P1 = 0 ; their P2 contributions are discarded with the row
  const float lse0 = ra < Lown ? lse_own_p[ra] * kLog2e : INFINITY, lse1 = rb < Lown ? lse_own_p[rb] * kLog2e : INFINITY;
  const float D0 = ra < Lown ? D_own_p[ra] : 0.f, D1 = rb < Lown ? D_own_p[rb] : 0.f;

  uint32_t ha[KS][4], da[KS][4];
  float acc[2 * KS][4];
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt % 3;
    const uint8_t* ch = oth_h + buf * BN * LD;
    const uint8_t* cd = oth_d + buf * BN * LD;
    cp_async_wait<1>();   // tile kt has landed
    __syncthreads();      // ... for everyone; tile kt - 1 is no longer read: its buffers are refilled with tile kt + 2
    if (kt + 2 < nkt) load_cols(kt + 2);
    cp_async_commit();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        load_a(ha[ks], own_h, LD, warp * 16, ks * 16, lane);
        load_a(da[ks], own_d, LD, warp * 16, ks * 16, lane);
      }
    }
    // the 64 streamed rows are processed in two halves of 32 to bound the live score registers
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
      // elementwise: w = dS1 + dS2 (-> s), q = g_oth * P2 (-> p2)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int jc = half * 32 + i * 8 + 2 * t;
        const float2 cl = *reinterpret_cast<const float2*>(col_lse + buf * BN + jc);
        const float2 cD = *reinterpret_cast<const float2*>(col_D + buf * BN + jc);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sl = s[i][e] * kLog2e;
          const float P1 = ex2(sl - ((e & 2) ? lse1 : lse0));
          const float P2 = ex2(sl - ((e & 1) ? cl.y : cl.x));
          const float w = P1 * (g_own * p1[i][e] - ((e & 2) ? D1 : D0)) + P2 * (g_oth * p2[i][e] - ((e & 1) ? cD.y : cD.x));
          s[i][e] = w;
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

#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) {
    const int col = i * 8 + 2 * t;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int r = hrow ? rb : ra;
      if (r >= Lown) continue;
      const int lr = warp * 16 + g + hrow * 8;
      const __nv_bfloat162 dd = *reinterpret_cast<const __nv_bfloat162*>(own_d + lr * LD + col * 2);
      float v0 = __bfloat162float(dd.x) + acc[i][2 * hrow], v1 = __bfloat162float(dd.y) + acc[i][2 * hrow + 1];
      const int64_t o = off_own + (int64_t)r * RP + col;
      if (hpre) {
        const __nv_bfloat162 hp = *reinterpret_cast<const __nv_bfloat162*>(hpre + o);
        v0 *= gelu_grad_fast(__bfloat162float(hp.x));
        v1 *= gelu_grad_fast(__bfloat162float(hp.y));
      }
      *reinterpret_cast<uint32_t*>(dh + o) = pack_bf16(v0, v1);
    }
  }
}

template <int RP, int NW>
static int fwd_launch(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames, int L,
                      int La, cudaStream_t st) {
  constexpr int LD = Cfg<RP>::LD, BM = NW * 16;
  const size_t smem = (size_t)(BM + 3 * BN) * LD;
  auto kern = xg_fwd_kernel<RP, NW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("xmodal_global_fwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  dim3 grid((unsigned)ceil_div(L > La ? L : La, BM), (unsigned)frames, 2);
  kern<<<grid, NW * 32, smem, st>>>((const bf16*)h, gate, (bf16*)out, (bf16*)attn_out, lse, frames, L, La);
  return check_launch("xmodal_global_fwd(tc)");
}

template <int RP, int NW>
static int bwd_launch(const void* h, const void* hpre, const float* gate, const float* lse, const float* D,
                      const void* dout, void* dh, int64_t frames, int L, int La, cudaStream_t st) {
  constexpr int LD = Cfg<RP>::LD, BM = NW * 16;
  const size_t smem = (size_t)(2 * BM + 6 * BN) * LD + 6 * BN * sizeof(float);
  auto kern = xg_bwd_kernel<RP, NW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("xmodal_global_bwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  dim3 grid((unsigned)ceil_div(L > La ? L : La, BM), (unsigned)frames, 2);
  kern<<<grid, NW * 32, smem, st>>>((const bf16*)h, (const bf16*)hpre, gate, lse, D, (const bf16*)dout, (bf16*)dh, frames, L, La);
  return check_launch("xmodal_global_bwd(tc)");
}

}  // namespace xtc

// xmodal_tc5.cu: tcgen05 / TMEM kernel for long rows (-1: shape not covered)
int xmodal_global_fwd_tc5(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames, int L, int La,
                          int rp, cudaStream_t st);

// entry points used by xmodal.cu's bf16 dispatch; return -1 for an unsupported rp
int xmodal_global_fwd_bf16(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames, int L,
                           int La, int rp, cudaStream_t st) {
  {
    const int rc = xmodal_global_fwd_tc5(h, gate, out, attn_out, lse, frames, L, La, rp, st);
    if (rc != -1) return rc;
  }
  // 4 warps (64 own rows) per CTA: measured against 8 and 2 warps at L = 3136 (32 clips): 5.1 / 5.7 / 5.5 ms forward,
  // 10.4 / 11.1 / 11.2 ms backward per step -- two block-wide barriers per 64-row tile are cheaper with fewer warps, and
  // the streamed tiles are L2 hits whichever CTA re-reads them
#define XTC_FWD(RP_)                                                                               \
  case RP_:                                                                                        \
    return xtc::fwd_launch<RP_, 4>(h, gate, out, attn_out, lse, frames, L, La, st);
  switch (rp) {
    XTC_FWD(16) XTC_FWD(32) XTC_FWD(48) XTC_FWD(64) XTC_FWD(96)
    default: return -1;
  }
#undef XTC_FWD
}

int xmodal_global_bwd_bf16(const void* h, const void* hpre, const float* gate, const float* lse, const float* D,
                           const void* dout, void* dh, int64_t frames, int L, int La, int rp, cudaStream_t st) {
#define XTC_BWD(RP_)                                                                                   \
  case RP_:                                                                                            \
    return xtc::bwd_launch<RP_, 4>(h, hpre, gate, lse, D, dout, dh, frames, L, La, st);
  switch (rp) {
    XTC_BWD(16) XTC_BWD(32) XTC_BWD(48) XTC_BWD(64) XTC_BWD(96)
    default: return -1;
  }
#undef XTC_BWD
}

}  // namespace stgcma
