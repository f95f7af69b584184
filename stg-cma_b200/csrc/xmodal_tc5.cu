// Global cross-modal token attention of the FFN adapters (Swin_AVE.py:801-808) on tcgen05 / TMEM, forward.
//
//   hv' = hv + gate_v * softmax(hv ha^T) ha         ha' = ha + gate_a * softmax(ha hv^T) hv          (logits unscaled)
//
// The contraction is tiny (d = r padded to 16 .. 64), so the kernel is bound by the SOFTMAX, not by the tensor pipe:
// 3136^2 exponentials per frame and direction at Swin stage 0.  What the TMEM formulation buys is the instruction count
// per score: the score tile comes out of the tensor core with ONE ROW PER THREAD (tcgen05.ld), so the row maximum / sum
// need no quad shuffles and the probabilities go back as the A operand of P.V through tcgen05.st -- no ldmatrix, no
// fragment packing (8.4 warp instructions per score in the mma.sync kernel of xmodal_tc.cu, ~4 here), and a third of
// the exponentials is evaluated on the FMA pipe (Cody-Waite + degree-3 polynomial, 1.1e-4 relative: far below the bf16
// rounding of P) because MUFU.EX2 at 16 / clk / SM is the roof of the kernel.
//
//   CTA = (frame, direction, 256 query rows): two 128-row query tiles, 8 softmax warps (warps w and w + 4 share the TMEM
//         lanes of quadrant w % 4 and own one query tile each), warp 8 issues the MMAs, warp 9 streams the other modality's
//         rows (keys = values) through a 4-stage cp.async ring of 96-row tiles in the canonical no-swizzle layout.
//   two passes over the key tiles: pass 1 row maxima (S recomputed by the tensor core: 1 MMA of K = 16 per tile at
//         stage 0), pass 2 P = exp2(S - max), row sums, O += P.H  -- no online rescaling of O.
//   TMEM: S[tile q][buffer b] 96 fp32 columns each (double buffered: the MMA of key tile g + 1 runs under the softmax of
//         tile g), P (bf16) over the first 48 columns of its S buffer, O[q] = 128 x rp fp32.
#include <stdlib.h>
#include "tc5_attn.cuh"

namespace stgcma {
namespace xg5 {

using namespace tc5;

constexpr int KN = 96;          // keys per tile
constexpr int NST = 4;          // ring stages
constexpr int kThreads = 320;   // 8 softmax warps + MMA warp + producer warp

// 2^x for x <= 0 on the FMA / ALU pipes: round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax of 2^f (1.1e-4
// relative), exponent add through the integer view.  x is clamped at -126 (the result is then ~1e-38: zero for us).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05459282f, 0.24221784f);
  p = fmaf(p, f, 0.6933686f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// rows [r0, r0 + nrows) of a contiguous [L][RP] bf16 matrix -> canonical core-matrix image (rows >= L zero-filled)
template <int RP>
__device__ __forceinline__ void load_rows_c(uint32_t dst, const bf16* __restrict__ src, int r0, int nrows, int L, int tid, int nthr) {
  constexpr int CH = RP / 8, RG = CH * 128;
  for (int e = tid; e < nrows * CH; e += nthr) {
    const int r = e / CH, kc = e - r * CH;
    const uint32_t d = dst + (r >> 3) * RG + kc * 128 + (r & 7) * 16;
    if (r0 + r < L) cp16(d, src + (int64_t)(r0 + r) * RP + kc * 8);
    else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0u) : "memory");
  }
}

template <int RP, int POLY>
__global__ void __launch_bounds__(kThreads, 1) xg5_fwd_kernel(const bf16* __restrict__ h, const float* __restrict__ gate,
                                                              bf16* __restrict__ out, bf16* __restrict__ attn_out,
                                                              float* __restrict__ lse, int64_t frames, int L, int La,
                                                              long long* dbg) {
  constexpr int CH = RP / 8, RG = CH * 128;
  const bool trace = dbg != nullptr && blockIdx.x == 3 && blockIdx.y == 100 && blockIdx.z == 0 && threadIdx.x == 256;
  int nstamp = 0;
#define STAMP() do { if (trace && nstamp < 60) dbg[nstamp++] = clock64(); } while (0)
  STAMP();
  constexpr uint32_t kQTile = 16 * RG, kKTile = (KN / 8) * RG;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);
  const uint32_t Qs = smem_u32(sm);                       // 2 query tiles
  const uint32_t Ring = Qs + 2 * kQTile;                  // NST key tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * kQTile + NST * kKTile);
  // kfull[NST] kfree[NST] sfull[2] pfull[2][2] ofull
  const uint32_t kfull = smem_u32(bars), kfree = smem_u32(bars + NST), sfull = smem_u32(bars + 2 * NST),
                 pfull = smem_u32(bars + 2 * NST + 2), ofull = smem_u32(bars + 2 * NST + 6);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 7);

  const int z = blockIdx.z;
  const int64_t frame = blockIdx.y;
  const int Lown = z ? La : L, Loth = z ? L : La;
  const int q0 = blockIdx.x * 256;
  if (q0 >= Lown) return;
  const int64_t row_own0 = (z ? frames * (int64_t)L : 0) + frame * (int64_t)Lown;
  const int64_t row_oth0 = (z ? 0 : frames * (int64_t)L) + frame * (int64_t)Loth;
  const bf16* hown = h + row_own0 * RP;
  const bf16* hoth = h + row_oth0 * RP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkt = (Loth + KN - 1) / KN;
  const int ntot = 2 * nkt;                               // pass 1 tiles, then pass 2 tiles

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(kfull + 8 * i, 32);
      mbar_init(kfree + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(sfull + 8 * i, 1);
    for (int i = 0; i < 4; ++i) mbar_init(pfull + 8 * i, 128);
    mbar_init(ofull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  load_rows_c<RP>(Qs, hown, q0, 256, Lown, tid, kThreads);
  cp_commit_wait();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t cO = 4 * KN;                              // O[q] at cO + q * RP

  if (warp == 9) {
    // ===================== producer: key tiles through the ring, two cp.async groups in flight =====================
    for (int g = 0; g < ntot; ++g) {
      const int st = g % NST;
      if (g >= NST) mbar_wait(kfree + 8 * st, (uint32_t)(((g / NST) - 1) & 1));
      const int kt = g < nkt ? g : g - nkt;
      load_rows_c<RP>(Ring + st * kKTile, hoth, kt * KN, KN, Loth, lane, 32);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (g >= 1) {   // (signalling each tile only after its own copies, one group in flight, measured slower: 7.1 vs 5.7 ms)
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        fence_async_smem();
        mbar_arrive(kfull + 8 * ((g - 1) % NST));
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_async_smem();
    mbar_arrive(kfull + 8 * ((ntot - 1) % NST));
  } else if (warp == 8) {
    // ===================== MMA issue =====================
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(RP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // descriptors: bases built once, the k-step / stage / tile offsets are added to the 16-byte-unit address field
    const uint64_t dQ = desc_nosw(Qs, 128, RG), dKk = desc_nosw(Ring, 128, RG), dKmn = desc_nosw(Ring, RG, 128);
    // retire tile g: (pass 2) O[q] += P[q][b] . H_tile for both query tiles; then the ring stage is free
    auto retire = [&](int g) {
      const int b = g & 1, st = g % NST;
      mbar_wait(pfull + 8 * (0 * 2 + b), (uint32_t)((g >> 1) & 1));
      mbar_wait(pfull + 8 * (1 * 2 + b), (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      if (g >= nkt + 2 && g < nkt + 8) STAMP();
      if (g >= nkt) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int ks = 0; ks < KN / 16; ++ks)
            mma_bf16_ts_elect(tmem + cO + q * RP, tmem + (uint32_t)((q * 2 + b) * KN + ks * 8),
                              dKmn + (uint64_t)((st * kKTile + ks * 2 * RG) >> 4), idesc_o, (uint32_t)((g > nkt) || ks != 0));
      }
      mma_commit_elect(kfree + 8 * st);
    };
    for (int g = 0; g < ntot; ++g) {
      const int b = g & 1, st = g % NST;
      const bool tr = (g >= nkt + 4 && g < nkt + 10);
      if (tr) STAMP();
      if (g >= 2) retire(g - 2);                           // frees S buffer b (in issue order) and ring stage (g - 2) % NST
      if (tr) STAMP();
      mbar_wait(kfull + 8 * st, (uint32_t)((g / NST) & 1));
      tc_fence_after();
      if (tr) STAMP();
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int ks = 0; ks < RP / 16; ++ks)
          mma_bf16_elect(tmem + (uint32_t)((q * 2 + b) * KN), dQ + (uint64_t)((q * kQTile + ks * 256) >> 4),
                         dKk + (uint64_t)((st * kKTile + ks * 256) >> 4), idesc_s, (uint32_t)(ks != 0));
      mma_commit_elect(sfull + 8 * b);
      if (tr) STAMP();
    }
    retire(ntot - 2);
    retire(ntot - 1);
    mma_commit_elect(ofull);
  } else {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int quad = warp & 3, q = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    const int i = q0 + q * 128 + quad * 32 + lane;        // own row within the frame
    const bool warp_live = q0 + q * 128 + quad * 32 < Lown;
    float m = -INFINITY, l = 0.f, mb = 0.f;
    for (int g = 0; g < ntot; ++g) {
      const int b = g & 1;
      const int kt = g < nkt ? g : g - nkt;
      const uint32_t sS = trow + (uint32_t)((q * 2 + b) * KN);
      if (g == nkt) mb = m * kLog2e;
      if (g < 6 || g == nkt - 1 || (g >= nkt && g < nkt + 6) || g == ntot - 1) STAMP();
      mbar_wait_relaxed(sfull + 8 * b, (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      if (g < 6 || g == nkt - 1 || (g >= nkt && g < nkt + 6) || g == ntot - 1) STAMP();
      if (warp_live) {
        const int jrem = Loth - kt * KN;                   // valid keys in this tile (>= KN: all)
        if (g < nkt) {
          // ---- pass 1: running row maximum ----
#pragma unroll
          for (int c0 = 0; c0 < KN; c0 += 32) {
            float v[32];
            tmem_ld32(sS + c0, v);
            if (c0 + 32 <= jrem) {
#pragma unroll
              for (int k = 0; k < 32; k += 2) m = fmaxf(m, fmaxf(v[k], v[k + 1]));
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k) m = fmaxf(m, (c0 + k < jrem) ? v[k] : -INFINITY);
            }
          }
        } else {
          // ---- pass 2: P = exp2(S log2e - max log2e) -> bf16 pairs over the first half of the S buffer ----
          // (one 32-column chunk at a time: all three TMEM loads in flight before one wait measured SLOWER, 6.3 vs 5.7 ms)
#pragma unroll
          for (int c0 = 0; c0 < KN; c0 += 32) {
            float v[32];
            tmem_ld32(sS + c0, v);
            uint32_t pk[16];
            if (c0 + 32 <= jrem) {
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float x0 = fmaf(v[2 * k], kLog2e, -mb), x1 = fmaf(v[2 * k + 1], kLog2e, -mb);
                const float p0 = ex2f(x0);
                const float p1 = (POLY && (k % 3) != 1) ? ex2_poly(x1) : ex2f(x1);   // 2/3 of the odd columns = 1/3 of all
                l += p0 + p1;
                pk[k] = pack2(p0, p1);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float p0 = (c0 + 2 * k < jrem) ? ex2f(fmaf(v[2 * k], kLog2e, -mb)) : 0.f;
                const float p1 = (c0 + 2 * k + 1 < jrem) ? ex2f(fmaf(v[2 * k + 1], kLog2e, -mb)) : 0.f;
                l += p0 + p1;
                pk[k] = pack2(p0, p1);
              }
            }
            tmem_st16(sS + (c0 >> 1), pk);
          }
          tmem_wait_st();
        }
      }
      tc_fence_before();
      mbar_arrive(pfull + 8 * (q * 2 + b));
    }
    STAMP();
    mbar_wait_relaxed(ofull, 0u);
    tc_fence_after();
    STAMP();
    if (warp_live) {
      const bool ok = i < Lown;
      const float inv = 1.f / l, gz = gate[z];
      const int64_t off = (row_own0 + (ok ? i : 0)) * RP;
#pragma unroll
      for (int c0 = 0; c0 < RP; c0 += 16) {
        float v[16];
        tmem_ld16u(trow + cO + q * RP + c0, v);
        if (ok) {
          float hv[16], a[16];
          load8(h + off + c0, *reinterpret_cast<float(*)[8]>(&hv[0]));
          load8(h + off + c0 + 8, *reinterpret_cast<float(*)[8]>(&hv[8]));
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            a[k] = v[k] * inv;
            hv[k] = fmaf(gz, a[k], hv[k]);
          }
          store16(attn_out + off + c0, a, 1.0f);
          store16(out + off + c0, hv, 1.0f);
        }
      }
      if (ok) lse[row_own0 + i] = (mb + log2f(l)) * 0.6931471805599453f;
    }
  }
  STAMP();
#undef STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int RP>
static int fwd_launch(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames, int L, int La,
                      cudaStream_t st) {
  constexpr int RG = RP / 8 * 128;
  const size_t smem = 128 + (size_t)(2 * 16 + NST * (KN / 8)) * RG + 128;
  static const bool poly = getenv("STGCMA_XG_NO_POLY") == nullptr;
  auto kern = poly ? xg5_fwd_kernel<RP, 1> : xg5_fwd_kernel<RP, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("xmodal_global_fwd(tcgen05): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  dim3 grid((unsigned)ceil_div(L > La ? L : La, 256), (unsigned)frames, 2);
  long long* dbg = nullptr;
  const bool trace = getenv("STGCMA_XG_TRACE") != nullptr;
  if (trace) { cudaMalloc(&dbg, 64 * sizeof(long long)); cudaMemset(dbg, 0, 64 * sizeof(long long)); }
  kern<<<grid, kThreads, smem, st>>>((const bf16*)h, gate, (bf16*)out, (bf16*)attn_out, lse, frames, L, La, dbg);
  if (trace) {     // debugging aid: cycle stamps of softmax warp 0 of one CTA (see STAMP() in the kernel)
    long long h_[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(h_, dbg, sizeof(h_), cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    fprintf(stderr, "xg5_fwd trace (L %d rp %d):", L, RP);
    for (int i = 1; i < 64 && h_[i]; ++i) fprintf(stderr, " %lld", h_[i] - h_[0]);
    fprintf(stderr, "\n");
  }
  return check_launch("xmodal_global_fwd(tcgen05)");
}

}  // namespace xg5

// -1: shape not covered (the caller uses the mma.sync kernels of xmodal_tc.cu)
int xmodal_global_fwd_tc5(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames, int L, int La,
                          int rp, cudaStream_t st) {
  // OPT-IN (STGCMA_XG_TC=1).  Measured on B200 at Swin-T stage 0 (640 frame pairs, L = 3136, rp = 16; scripts/xg_bench.py,
  // phase traces STGCMA_XG_TRACE=1): 5.7 - 6.5 ms against 4.7 ms of the mma.sync kernel in xmodal_tc.cu.  The exponentials of
  // pass 2 do run at the MUFU roof (1570 of 1536 cycles per 2 x 128 x 96 tile), but the two-pass scheme reads every score
  // tile twice from TMEM (~160 B / clk / SM measured: 600 cycles per tile and pass) and the per-tile MMA <-> softmax
  // hand-shakes (12 P.V MMAs of N = 16 cost ~40 issue cycles each) are not hidden by two S buffers.  Long rows only.
  if ((L < 512 && La < 512) || !getenv("STGCMA_XG_TC")) return -1;
  switch (rp) {
    case 16: return xg5::fwd_launch<16>(h, gate, out, attn_out, lse, frames, L, La, st);
    case 32: return xg5::fwd_launch<32>(h, gate, out, attn_out, lse, frames, L, La, st);
    case 48: return xg5::fwd_launch<48>(h, gate, out, attn_out, lse, frames, L, La, st);
    case 64: return xg5::fwd_launch<64>(h, gate, out, attn_out, lse, frames, L, La, st);
    default: return -1;
  }
}

}  // namespace stgcma
