// nn.Linear family for the STG-CMA adapter train step.
//
//   stgcma_gemm      out = epilogue(A . W^T + bias)
//       bf16: persistent warp-specialised tcgen05 kernel -- TMA (SWIZZLE_128B, 3-D maps so the
//             per-modality "groups" never bleed into each other) -> 128 x BN x 64 smem stages ->
//             tcgen05.mma (cta_group::1, M=128, fp32 accumulators double-buffered in TMEM) ->
//             8 epilogue warps: operand loads issued first, tcgen05.ld -> smem transpose -> bias ->
//             residual / GELU / dGELU / row-scale -> coalesced 128-bit stores.
//       fp32: plain CUDA-core tiled GEMM (the check mode; also the numerics reference).
//   stgcma_gemm_tn   dW += A^T . B   (trainable adapter / head weights), split over rows,
//                    fp32 atomics into the gradient bucket.
#include <stdlib.h>
#include "tc_common.cuh"
#include "gelu_poly.cuh"

namespace stgcma {

// =========================================================================================
// Epilogue (shared)
// =========================================================================================
template <typename T>
__device__ __forceinline__ void epi_elem(const EpiParams& p, int64_t row, int col, float v) {
  const int64_t off = row * p.ldo + col;
  T* out = reinterpret_cast<T*>(p.out);
  if (p.rowscale) v *= p.rowscale[row];
  switch (p.mode) {
    case STGCMA_EPI_LINEAR: {
      if (p.res1) v += ld_f(reinterpret_cast<const T*>(p.res1) + off);
      if (p.res2) v += ld_f(reinterpret_cast<const T*>(p.res2) + off);
      st_f(out + off, v);
    } break;
    case STGCMA_EPI_GELU: {
      if (p.out_pre) st_f(reinterpret_cast<T*>(p.out_pre) + off, v);
      st_f(out + off, gelu_erf(v));
    } break;
    case STGCMA_EPI_DGELU: {
      const float a = ld_f(reinterpret_cast<const T*>(p.aux) + off);
      st_f(out + off, v * gelu_erf_grad(a));
    } break;
    case STGCMA_EPI_QGELU: {
      if (p.out_pre) st_f(reinterpret_cast<T*>(p.out_pre) + off, v);
      st_f(out + off, qgelu(v));
    } break;
    case STGCMA_EPI_DQGELU: {
      const float a = ld_f(reinterpret_cast<const T*>(p.aux) + off);
      st_f(out + off, v * qgelu_grad(a));
    } break;
    case STGCMA_EPI_GELU_G: {
      if (p.out_pre) st_f(reinterpret_cast<T*>(p.out_pre) + off, gelu_erf_grad(v));
      st_f(out + off, gelu_erf(v));
    } break;
    case STGCMA_EPI_QGELU_G: {
      if (p.out_pre) st_f(reinterpret_cast<T*>(p.out_pre) + off, qgelu_grad(v));
      st_f(out + off, qgelu(v));
    } break;
    default: {
      const float a = ld_f(reinterpret_cast<const T*>(p.aux) + off);
      st_f(out + off, v * a);
    } break;
  }
}

// GELU + GELU' of the derivative-saving epilogue: the tanh form by default (STGCMA_GELU_ERF=1 at compile time keeps the erf form)
__device__ __forceinline__ void gelu_val_grad_act(float x, float& y, float& dy) {
#ifdef STGCMA_GELU_ERF
  gelu_val_grad_fast(x, y, dy);
#else
  gelu_val_grad_tanh(x, y, dy);
#endif
}

// Epilogue of the tcgen05 kernel.  An epilogue warp owns 32 rows x 32 columns per step.  The 32x32b TMEM load
// hands every lane 32 columns of ITS OWN row; accessing global memory that way would touch 32 cache lines per
// warp instruction (measured: the LSU wavefront rate then caps the kernel at ~2 TB/s), so the chunk is
// transposed through a padded shared-memory staging tile into a coalesced layout -- 4 lanes x 8 columns per
// row, 8 rows per pass, 4 passes -- in which every residual / aux load and every store is a 128-bit access
// and a warp instruction covers 64..128 contiguous bytes per row.  All operand loads of a chunk are issued
// first (`epi_load`: before the TMEM read, and for the first chunk of a tile before the accumulator is even
// ready) and consumed afterwards (`epi_apply`); the epilogue is latency-bound otherwise.
// Activation tensors (res2 / aux / out_pre) are bf16; `out` and `res1` are bf16, or fp32 when p.out_f32.
// The epilogue variant is a compile-time property of the kernel (MODE, OUT_F32, NRES = number of residual
// operands): with run-time flags the chunk loop executed ~400 instructions per 32 x 32 chunk, 80 % of them
// uniform branches and constant-bank loads (ncu source page, profiles/r01b), and was issue-bound.
struct EpiRegs {
  uint4 a[4][2];  // per pass: res1 (fp32: 2 x 16 B, bf16: 1 x 16 B)  or aux (bf16)
  uint4 b[4];     // per pass: res2 (bf16)
  float rs[4];    // per pass: rowscale
};

// off0: element offset (row * ldo + col) of this lane in pass 0; pass ps is pass_stride elements further.
template <int MODE, bool OUT_F32, int NRES>
__device__ __forceinline__ void epi_load(const EpiParams& p, int64_t off0, int64_t pass_stride, int64_t row0, int rpp,
                                         int nvalid, EpiRegs& r) {
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    if (ps >= nvalid) continue;
    const int64_t off = off0 + ps * pass_stride;
    r.rs[ps] = p.rowscale ? __ldg(p.rowscale + row0 + ps * rpp) : 1.f;
    if (MODE == STGCMA_EPI_LINEAR) {
      if (NRES >= 1) {
        if (OUT_F32) {
          const uint4* s4 = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.res1) + off);
          r.a[ps][0] = s4[0];
          r.a[ps][1] = s4[1];
        } else {
          r.a[ps][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.res1) + off);
        }
      }
      if (NRES >= 2) r.b[ps] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.res2) + off);
    } else if (MODE != STGCMA_EPI_GELU && MODE != STGCMA_EPI_QGELU && MODE != STGCMA_EPI_GELU_G && MODE != STGCMA_EPI_QGELU_G) {
      r.a[ps][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.aux) + off);
    }
  }
}

// v: acc + bias of 8 consecutive columns of one row (pass ps of the chunk)
template <int MODE, bool OUT_F32, int NRES>
__device__ __forceinline__ void epi_apply(const EpiParams& p, int64_t off, int ps, float (&v)[8], const EpiRegs& r) {
  if (p.rowscale) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= r.rs[ps];
  }
  if (MODE == STGCMA_EPI_LINEAR) {
    if (NRES >= 1) {
      if (OUT_F32) {
        const float4 x = *reinterpret_cast<const float4*>(&r.a[ps][0]), y = *reinterpret_cast<const float4*>(&r.a[ps][1]);
        v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
      } else {
        float t[8];
        unpack8(r.a[ps][0], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
      }
    }
    if (NRES >= 2) {
      float t[8];
      unpack8(r.b[ps], t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
  } else if (MODE == STGCMA_EPI_GELU || MODE == STGCMA_EPI_QGELU) {
    if (p.out_pre) store8(reinterpret_cast<bf16*>(p.out_pre) + off, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (MODE == STGCMA_EPI_GELU) ? gelu_fast(v[i]) : qgelu(v[i]);
  } else if (MODE == STGCMA_EPI_GELU_G || MODE == STGCMA_EPI_QGELU_G) {
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == STGCMA_EPI_GELU_G) gelu_val_grad_act(v[i], v[i], d[i]);
      else qgelu_val_grad(v[i], v[i], d[i]);
    }
    if (p.out_pre) store8(reinterpret_cast<bf16*>(p.out_pre) + off, d);
  } else {
    float t[8];
    unpack8(r.a[ps][0], t);
#pragma unroll
    // (round 1 kept the erf forms here because the polynomial forms "moved one temporal-bias-table gradient from 1.9e-2 to
    //  2.4e-2"; the round-2 noise analysis -- DESIGN.md section 4 -- shows that was a re-draw of the bf16 storage noise, not an
    //  accuracy loss.  These dX epilogues have N = rp <= 96 columns and are not the bottleneck of their GEMMs, so they stay.)
    for (int i = 0; i < 8; ++i)
      v[i] *= (MODE == STGCMA_EPI_DGELU) ? gelu_grad_fast(t[i]) : (MODE == STGCMA_EPI_DQGELU) ? qgelu_grad(t[i]) : t[i];
  }
  if (OUT_F32) store8(reinterpret_cast<float*>(p.out) + off, v);
  else store8(reinterpret_cast<bf16*>(p.out) + off, v);
}

// =========================================================================================
// CUDA-core GEMM (fp32 check mode; generic T)
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ W, int64_t group_rows,
                                                        int N, int K, int64_t lda, int64_t ldw, EpiParams ep) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int g = blockIdx.z;
  const T* Ag = A + (int64_t)g * group_rows * lda;
  const T* Wg = W + (int64_t)g * N * ldw;
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      const int64_t m = m0 + lr;
      As[lk + i][lr] = (m < group_rows && k < K) ? ld_f(Ag + m * lda + k) : 0.f;
      const int n = n0 + lr;
      Ws[lk + i][lr] = (n < N && k < K) ? ld_f(Wg + (int64_t)n * ldw + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= group_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = acc[i][j] + (ep.bias ? ep.bias[(int64_t)g * N + n] : 0.f);
      epi_elem<T>(ep, (int64_t)g * group_rows + m, n, v);
    }
  }
}

// =========================================================================================
// tcgen05 / TMEM / TMA GEMM (bf16 in, fp32 accumulate)
// =========================================================================================
namespace tc {

// Epilogue warps: two per TMEM lane quarter, alternating 32-column chunks -- or four (OCC = 3: one CTA per SM with 16
// epilogue warps) for the wide-N shapes with K < 1024 whose epilogue, not the MMA, sets the tile time.
constexpr int epi_warps(int occ) { return occ == 3 ? 16 : 8; }
constexpr int kStageRowF = 36;           // staging row stride in floats (32 + 4 pad: conflict-free 128-bit rows)
constexpr int staging_bytes(int occ) { return epi_warps(occ) * 32 * kStageRowF * 4; }
constexpr int kABytes = BM * BK * 2;     // 16 KiB

struct Params {
  int64_t group_rows;
  int N, K, groups;
  int BN, n_tiles, m_tiles_g, num_tiles, k_blocks, stages;
  uint32_t tmem_cols;
  EpiParams epi;
};

// OCC: CTAs per SM the variant is compiled for.  OCC = 2 (BN <= 128, half the smem / TMEM each) is used for the
// small-K, epilogue-bound shapes: two tiles in flight per SM hide the epilogue latency; OCC = 1 (BN <= 256) for the
// large-K, tensor-bound ones (a 128 x 256 tile reads 25 % less shared memory per flop than two 128 x 128 tiles).
// TMAOUT: 0 = transposing epilogue; 1 = plain bf16 output through TMA stores; 2 = the derivative-saving activation epilogues
// (GELU_G / QGELU_G) through TMA stores: activation -> `out` (tmC), derivative -> `out_pre` (tmC2), two staging tiles.
template <int MODE, bool OUT_F32, int NRES, int OCC, int TMAOUT>
__global__ void __launch_bounds__(64 + epi_warps(OCC) * 32, OCC == 3 ? 1 : OCC)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for SWIZZLE_128B, by pointer arithmetic on the shared pointer (keeps ld/st.shared addressing)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kEpiWarps = epi_warps(OCC), kStagingBytes = staging_bytes(OCC);
  const uint32_t stage_bytes = kABytes + (uint32_t)p.BN * 128u;
  float* staging = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(staging) + kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* tfull = bars + 2 * p.stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(tfull + s), 1);
      mbar_init(smem_u32(tempty + s), kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int g = mt / p.m_tiles_g;
        const int m_blk = mt - g * p.m_tiles_g;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(smem_u32(empty + stage), phase ^ 1u);
          const uint32_t fb = smem_u32(full + stage);
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          tma_load_3d(sa, &tmA, fb, kb * BK, m_blk * BM, g);
          tma_load_3d(sa + kABytes, &tmW, fb, kb * BK, n_blk * p.BN, g);
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform loop, one elected lane per instruction (mma_bf16_elect) =====================
    {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(tempty + acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(smem_u32(full + stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_sdesc(sa);
          const uint64_t bdesc = make_sdesc(sa + kABytes);
          int ksteps = (p.K - kb * BK + 15) >> 4;
          if (ksteps > BK / 16) ksteps = BK / 16;
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle atom: +2 in 16 B units
            mma_bf16_elect(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          }
          mma_commit_elect(smem_u32(empty + stage));
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
        }
        mma_commit_elect(smem_u32(tfull + acc));
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (TMEM -> regs -> smem transpose -> global) =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int par = (warp - 2) >> 2;    // chunks par, par + NPAR, ... of the tile
    constexpr int NPAR = kEpiWarps / 4;
    const int nchunks = (p.BN + 31) >> 5;
    float* stg = staging + (size_t)(warp - 2) * 32 * kStageRowF;
    const EpiParams ep = p.epi;
    uint32_t acc = 0, acc_phase = 0;
    if (TMAOUT) {
      // Plain bf16 output (out = acc + bias): the tile is written lane-per-row, straight from the TMEM load, into a
      // K-major SWIZZLE_128B staging tile and leaves with one TMA store per 64 columns -- ~2 instructions per element
      // instead of ~17 through the transposing epilogue below, which was issue-bound at K = 96 (ncu: 51 % issue
      // slots busy, 41 % DRAM throughput; profiles/r01e).  Row / column tails are clipped by the tensor map.
      uint8_t* stile = reinterpret_cast<uint8_t*>(staging);
      uint8_t* stile2 = stile + ((p.BN + 63) >> 6) * kABytes;   // TMAOUT == 2: the derivative tile
      const int trow = q * 32 + lane;
      bool first = true;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int g = mt / p.m_tiles_g;
        const int m_blk = mt - g * p.m_tiles_g;
        const uint32_t t_row = tmem_base + acc * (uint32_t)p.BN + ((uint32_t)(q * 32) << 16);
        mbar_wait(smem_u32(tfull + acc), acc_phase);
        tc_fence_after();
        // the previous tile's TMA store must have finished READING the staging tile
        if (!first) {
          if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        }
        first = false;
        for (int ch = par; ch < nchunks; ch += NPAR) {
          const int c0 = ch << 5;
          const bool wide = p.BN - c0 >= 32;
          float v[32];
          if (wide) tmem_ld32(t_row + (uint32_t)c0, v);
          else tmem_ld16(t_row + (uint32_t)c0, v);
          const int col = n_blk * p.BN + c0;
          const float* bp = ep.bias ? ep.bias + (int64_t)g * p.N + col : nullptr;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j >= 2 && !wide) break;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = v[8 * j + i];
            if (bp && col + 8 * j < p.N) {   // N % 8 == 0: an 8-column group is inside or outside as a whole
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp + 8 * j)), b1 = __ldg(reinterpret_cast<const float4*>(bp + 8 * j + 4));
              o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
            }
            const int cc = c0 + 8 * j;
            const uint32_t soff = (uint32_t)((cc >> 6) * kABytes + trow * 128 + ((((cc & 63) >> 3) ^ (trow & 7)) << 4));
            Vec<bf16> pk;
            if (TMAOUT == 2) {
              Vec<bf16> dk;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float y, d;
                if (MODE == STGCMA_EPI_GELU_G) gelu_val_grad_act(o[i], y, d);
                else qgelu_val_grad(o[i], y, d);
                pk.v[i] = y;
                dk.v[i] = d;
              }
              dk.store(reinterpret_cast<bf16*>(stile2 + soff));
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) pk.v[i] = o[i];
            }
            pk.store(reinterpret_cast<bf16*>(stile + soff));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tempty + acc));   // accumulator drained: the MMA warp may reuse it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (warp == 2 && lane == 0) {
          for (int cb = 0; cb * 64 < p.BN; ++cb) {
            if (n_blk * p.BN + cb * 64 >= p.N) break;
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(smem_u32(stile + cb * kABytes)), "r"(n_blk * p.BN + cb * 64),
                           "r"(m_blk * BM), "r"(g)
                         : "memory");
            if (TMAOUT == 2)
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmC2)), "r"(smem_u32(stile2 + cb * kABytes)), "r"(n_blk * p.BN + cb * 64),
                             "r"(m_blk * BM), "r"(g)
                           : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int g = mt / p.m_tiles_g;
      const int m_blk = mt - g * p.m_tiles_g;
      const uint32_t t_row = tmem_base + acc * (uint32_t)p.BN + ((uint32_t)(q * 32) << 16);
      const int64_t wrow = (int64_t)m_blk * BM + q * 32;   // first row (within the group) of this warp
      bool waited = false;
      for (int ch = par; ch < nchunks; ch += NPAR) {
        const int c0 = ch << 5;
        const bool wide = p.BN - c0 >= 32;
        const int lpr = wide ? 4 : 2;       // lanes per row (8 columns each)
        const int rpp = wide ? 8 : 16;      // rows per pass
        const int r_in = wide ? (lane >> 2) : (lane >> 1), seg = wide ? (lane & 3) : (lane & 1);
        const int col = n_blk * p.BN + c0 + seg * 8;
        const bool col_ok = col < p.N;
        // passes whose row is inside the group: rows wrow + r_in + ps * rpp, ps < (wide ? 4 : 2)
        const int64_t rows_left = p.group_rows - (wrow + r_in);
        int nvalid = rows_left <= 0 ? 0 : (int)((rows_left + rpp - 1) / rpp);
        nvalid = nvalid > 32 / rpp ? 32 / rpp : nvalid;
        if (!col_ok) nvalid = 0;
        const int64_t grow0 = (int64_t)g * p.group_rows + wrow + r_in;
        const int64_t off0 = grow0 * ep.ldo + col;
        const int64_t pstride = (int64_t)rpp * ep.ldo;
        EpiRegs er;
        epi_load<MODE, OUT_F32, NRES>(ep, off0, pstride, grow0, rpp, nvalid, er);
        float bias8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (ep.bias && nvalid > 0) load8(ep.bias + (int64_t)g * p.N + col, bias8);
        if (!waited) {
          mbar_wait(smem_u32(tfull + acc), acc_phase);
          tc_fence_after();
          waited = true;
        }
        float v[32];
        if (wide) tmem_ld32(t_row + (uint32_t)c0, v);
        else tmem_ld16(t_row + (uint32_t)c0, v);
        if (ch + NPAR >= nchunks) {
          // all TMEM reads of this warp for this accumulator are done: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(tempty + acc));
        }
        float4* srow = reinterpret_cast<float4*>(stg + lane * kStageRowF);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < 4 || wide) srow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (nvalid > 0) {
          const float* sbase = stg + r_in * kStageRowF + seg * 8;
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            if (ps >= nvalid) continue;
            const float4* s4 = reinterpret_cast<const float4*>(sbase + ps * rpp * kStageRowF);
            const float4 x = s4[0], y = s4[1];
            float o[8] = {x.x + bias8[0], x.y + bias8[1], x.z + bias8[2], x.w + bias8[3],
                          y.x + bias8[4], y.y + bias8[5], y.z + bias8[6], y.w + bias8[7]};
            epi_apply<MODE, OUT_F32, NRES>(ep, off0 + ps * pstride, ps, o, er);
          }
        }
        __syncwarp();
      }
      if (!waited) {
        // this warp owns no chunk of the tile (BN <= 32 and par == 1): it still takes part in the handshake
        mbar_wait(smem_u32(tfull + acc), acc_phase);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tempty + acc));
      }
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

static int launch(const void* A, const void* W, int64_t M, int N, int K, int64_t lda, int64_t ldw, int groups,
                  const EpiParams& ep, cudaStream_t st) {
  Params p;
  p.group_rows = M / groups;
  p.N = N;
  p.K = K;
  p.groups = groups;
  // tile width: the multiple of 16 (<= bn_max) that wastes the fewest padded columns, larger on ties
  // One CTA per SM with tiles up to 128 x 256 and a 4-stage ring when the tensor pipe is the limit: K >= 1024, or
  // K >= 256 with a wide N unless the epilogue evaluates GELU / GELU' at K < 768 (then two CTAs' epilogue warps per SM
  // matter more than the deeper ring).  Measured per shape with scripts/gemm_shapes.py (profiles/r01f): qkv stage 2
  // 775 -> 859 TFLOP/s, qkv stage 3 908 -> 1176, proj stage 3 780 -> 961; mlp.fc1 + GELU stage 2 532 -> 509 (kept at 2).
  const bool act_epi = ep.mode == STGCMA_EPI_GELU || ep.mode == STGCMA_EPI_DGELU || ep.mode == STGCMA_EPI_QGELU ||
                       ep.mode == STGCMA_EPI_DQGELU || ep.mode == STGCMA_EPI_GELU_G || ep.mode == STGCMA_EPI_QGELU_G;
  int occ = (K >= 1024 || (K >= 256 && N >= 384 && (!act_epi || K >= 768))) ? 1 : 2;
  // wide-N outputs behind a short K: a 128 x 256 tile's MMAs take ~3 k cycles at K = 384 while 8 epilogue warps need 2 - 4x
  // that to drain it (TMEM -> transpose -> GELU' / aux multiply -> two [M, N] tensors): 16 epilogue warps per CTA
  static const int epw16 = getenv("STGCMA_GEMM_EPW16") ? atoi(getenv("STGCMA_GEMM_EPW16")) : 1;
  // measured (scripts/gemm_shapes.py, 64 clips): fc1 + GELU' stage 2 3.50 -> 3.16 ms, its MULAUX backward 3.61 -> 3.13 ms, the
  // stage-3 pair 0.70 / 0.68 -> 0.66 / 0.61 ms; the plain wide-N GEMMs (qkv) lose 1 - 4 % to the shallower ring: not routed here
  if (epw16 && ep.mode != STGCMA_EPI_LINEAR && K >= 256 && K < 1024 && N >= 1024) occ = 3;
  // The derivative-saving activation epilogues of those shapes are ISSUE bound: 40 instructions per output element at 78 % of the
  // issue slots (ncu, profiles/r02d_ncu_gemm_fc1g_s2.txt), ~30 of them the erf / expf / rcp evaluation -- hence the tanh form
  // (gelu_val_grad_act: 3.15 -> 2.66 ms per step at stage 2).  Opt-in experiment (STGCMA_GEMM_ACT_TMA=1): lane-per-row into two
  // swizzled staging tiles + two TMA stores with 128-wide tiles; fewer instructions (35 -> 19 per element) but measured SLOWER
  // (3.20 ms: two block barriers per tile around the single-buffered staging tiles, twice the tiles), so it is off.
  static const int act_tma = getenv("STGCMA_GEMM_ACT_TMA") ? atoi(getenv("STGCMA_GEMM_ACT_TMA")) : 0;
  const bool tma_act = act_tma && occ == 3 && (ep.mode == STGCMA_EPI_GELU_G || ep.mode == STGCMA_EPI_QGELU_G) && ep.out_pre &&
                       !ep.out_f32 && !ep.rowscale && N % 128 == 0;
  const int bn_max = (occ != 2 && !tma_act) ? 256 : 128;
  int best_bn = 16;
  int64_t best_cost = -1;
  for (int bn = bn_max; bn >= 16; bn -= 16) {
    const int64_t cost = ceil_div(N, bn) * (int64_t)bn + (bn < 64 && N > bn ? N : 0);   // avoid slivers when N is large
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_bn = bn; }
  }
  // Plain bf16 output -> TMA-store epilogue.  Its 64-column store boxes are clipped only at the tensor's edge, so
  // either the whole N fits one tile or the tile width is a multiple of 64 (and <= 128: two staging blocks).
  // (Forcing BN = 128 for N = 288 to qualify was slower than BN = 96 with the transposing epilogue: measured.)
  const bool tma_cand = ep.mode == STGCMA_EPI_LINEAR && !ep.out_f32 && !ep.res1 && !ep.res2 && !ep.rowscale;
  // two CTAs per SM, plain output, several n-tiles: a tile width of 64 / 128 keeps the TMA-store epilogue (~2 instructions per
  // element) where the padding-optimal width (e.g. 96 for N = 192 / 576) would fall back to the transposing one (~17)
  static const int tma64 = getenv("STGCMA_GEMM_TMA64") ? atoi(getenv("STGCMA_GEMM_TMA64")) : 1;
  if (tma64 && occ == 2 && tma_cand && N > 128 && N % 64 == 0 && best_bn % 64 != 0) best_bn = (N % 128 == 0) ? 128 : 64;
  p.BN = best_bn;
  p.n_tiles = (int)ceil_div(N, p.BN);
  p.m_tiles_g = (int)ceil_div(p.group_rows, BM);
  const int64_t tiles = (int64_t)groups * p.m_tiles_g * p.n_tiles;
  STG_REQUIRE(tiles < (1ll << 31), "gemm: too many tiles");
  p.num_tiles = (int)tiles;
  p.k_blocks = (int)ceil_div(K, BK);
  const int stage_bytes = kABytes + p.BN * 128;
  const int smem_cap = occ != 2 ? 232448 : (233472 / 2 - 1024);   // per CTA; 1 KB per CTA is reserved by the driver
  const int kStagingBytes = staging_bytes(occ), kThreads = 64 + epi_warps(occ) * 32;
  const int budget = smem_cap - 1024 - kStagingBytes - 256;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.BN)) cols <<= 1;
  p.tmem_cols = cols;
  p.epi = ep;

  CUtensorMap tmA, tmW, tmC, tmC2;
  int rc = make_map(&tmA, A, p.group_rows, K, lda, groups, BM);
  if (rc) return rc;
  rc = make_map(&tmW, W, N, K, ldw, groups, p.BN);
  if (rc) return rc;
  // plain bf16 output -> TMA-store epilogue (needs the two-block staging tile: BN <= 128)
  const bool tma_out = tma_cand && p.BN <= 128 && (p.n_tiles == 1 || p.BN % 64 == 0);
  tmC = tmA;
  tmC2 = tmA;
  if (tma_out || tma_act) {
    rc = make_map(&tmC, ep.out, p.group_rows, N, ep.ldo, groups, BM);
    if (rc) return rc;
  }
  if (tma_act) {
    rc = make_map(&tmC2, ep.out_pre, p.group_rows, N, ep.ldo, groups, BM);
    if (rc) return rc;
  }

  const size_t smem = 1024 + (size_t)stages * stage_bytes + kStagingBytes + 256;
  const int ctas_per_sm = occ == 2 ? 2 : 1;
  const int grid = p.num_tiles < ctas_per_sm * sm_count() ? p.num_tiles : ctas_per_sm * sm_count();
  const int nres = (ep.mode == STGCMA_EPI_LINEAR) ? (ep.res1 ? (ep.res2 ? 2 : 1) : 0) : 0;
  if (ep.mode == STGCMA_EPI_LINEAR && ep.res2 && !ep.res1) {
    set_error("gemm(bf16): res2 without res1 is not supported");
    return STGCMA_EINVAL;
  }
#define STG_TC_LAUNCH(MODE_, F32_, NRES_)                                                                      \
  do {                                                                                                         \
    if (occ == 3) STG_TC_LAUNCH2(MODE_, F32_, NRES_, 3, 0);                                                    \
    if (occ == 1) STG_TC_LAUNCH2(MODE_, F32_, NRES_, 1, 0);                                                    \
    STG_TC_LAUNCH2(MODE_, F32_, NRES_, 2, 0);                                                                  \
  } while (0)
#define STG_TC_LAUNCH2(MODE_, F32_, NRES_, OCC_, TMA_)                                                         \
  do {                                                                                                         \
    auto kern = gemm_tc_kernel<MODE_, F32_, NRES_, OCC_, TMA_>;                                                    \
    static std::once_flag once;                                                                                \
    static cudaError_t attr_err = cudaSuccess;                                                                 \
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, OCC_ != 2 ? 232448 : 115712); }); \
    if (attr_err != cudaSuccess) {                                                                             \
      set_error("gemm: cannot raise dynamic smem limit: %s", cudaGetErrorString(attr_err));                    \
      return STGCMA_ECUDA;                                                                                     \
    }                                                                                                          \
    kern<<<grid, kThreads, smem, st>>>(tmA, tmW, tmC, tmC2, p);                                                   \
    return check_launch("gemm(tcgen05)");                                                                      \
  } while (0)
  const bool f32 = ep.out_f32 != 0;
  if (tma_out) {
    if (occ == 1) STG_TC_LAUNCH2(STGCMA_EPI_LINEAR, false, 0, 1, 1);
    if (occ == 2) STG_TC_LAUNCH2(STGCMA_EPI_LINEAR, false, 0, 2, 1);
  }
  if (tma_act) {
    if (ep.mode == STGCMA_EPI_GELU_G) STG_TC_LAUNCH2(STGCMA_EPI_GELU_G, false, 0, 3, 2);
    STG_TC_LAUNCH2(STGCMA_EPI_QGELU_G, false, 0, 3, 2);
  }
  switch (ep.mode) {
    case STGCMA_EPI_LINEAR:
      if (f32) {
        if (nres == 0) STG_TC_LAUNCH(STGCMA_EPI_LINEAR, true, 0);
        if (nres == 1) STG_TC_LAUNCH(STGCMA_EPI_LINEAR, true, 1);
        STG_TC_LAUNCH(STGCMA_EPI_LINEAR, true, 2);
      } else {
        if (nres == 0) STG_TC_LAUNCH(STGCMA_EPI_LINEAR, false, 0);
        if (nres == 1) STG_TC_LAUNCH(STGCMA_EPI_LINEAR, false, 1);
        STG_TC_LAUNCH(STGCMA_EPI_LINEAR, false, 2);
      }
    case STGCMA_EPI_GELU:
      if (f32) STG_TC_LAUNCH(STGCMA_EPI_GELU, true, 0);
      STG_TC_LAUNCH(STGCMA_EPI_GELU, false, 0);
    case STGCMA_EPI_DGELU:
      if (f32) STG_TC_LAUNCH(STGCMA_EPI_DGELU, true, 0);
      STG_TC_LAUNCH(STGCMA_EPI_DGELU, false, 0);
    case STGCMA_EPI_QGELU:
      STG_REQUIRE(!f32, "gemm(bf16): QuickGELU epilogues write the activation dtype");
      STG_TC_LAUNCH(STGCMA_EPI_QGELU, false, 0);
    case STGCMA_EPI_DQGELU:
      STG_REQUIRE(!f32, "gemm(bf16): QuickGELU epilogues write the activation dtype");
      STG_TC_LAUNCH(STGCMA_EPI_DQGELU, false, 0);
    case STGCMA_EPI_GELU_G:
      STG_REQUIRE(!f32, "gemm(bf16): derivative-saving epilogues write the activation dtype");
      STG_TC_LAUNCH(STGCMA_EPI_GELU_G, false, 0);
    case STGCMA_EPI_QGELU_G:
      STG_REQUIRE(!f32, "gemm(bf16): derivative-saving epilogues write the activation dtype");
      STG_TC_LAUNCH(STGCMA_EPI_QGELU_G, false, 0);
    default:
      if (f32) STG_TC_LAUNCH(STGCMA_EPI_MULAUX, true, 0);
      STG_TC_LAUNCH(STGCMA_EPI_MULAUX, false, 0);
  }
#undef STG_TC_LAUNCH
#undef STG_TC_LAUNCH2
}

}  // namespace tc

// =========================================================================================
// dW += A^T B  (CUDA-core version; generic T)
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(256) gemm_tn_simt_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                           const float* __restrict__ rowscale, float* __restrict__ dW,
                                                           float* __restrict__ csumA, float* __restrict__ csumB, int64_t M,
                                                           int Ka, int Nb, int64_t lda, int64_t ldb, int64_t ldo,
                                                           int64_t rows_per_split) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  const int64_t mbeg = (int64_t)blockIdx.z * rows_per_split;
  int64_t mend = mbeg + rows_per_split;
  if (mend > M) mend = M;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 4, lc = (tid & 15) * 4;  // load: 16 rows x 64 cols, 4 consecutive cols per thread
  float acc[4][4] = {};
  float sa[4] = {}, sb[4] = {};
  for (int64_t m0 = mbeg; m0 < mend; m0 += 16) {
    const int64_t m = m0 + lr;
    const float rs = (m < mend && rowscale) ? rowscale[m] : 1.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + lc + c, j = j0 + lc + c;
      As[lr][lc + c] = (m < mend && i < Ka) ? ld_f(A + m * lda + i) * rs : 0.f;
      Bs[lr][lc + c] = (m < mend && j < Nb) ? ld_f(B + m * ldb + j) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sa[i] += a[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (csumB) {
        const float rk = (rowscale && m0 + k < mend) ? rowscale[m0 + k] : 1.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) sb[j] += b[j] * rk;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = i0 + ty * 4 + i;
    if (ii >= Ka) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj < Nb) atomicAdd(dW + (int64_t)ii * ldo + jj, acc[i][j]);
    }
    if (csumA && blockIdx.y == 0 && tx == 0) atomicAdd(csumA + ii, sa[i]);
  }
  if (csumB && blockIdx.x == 0 && ty == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj < Nb) atomicAdd(csumB + jj, sb[j]);
    }
  }
}

template <typename T>
static int gemm_tn_launch(const void* A, const void* B, const float* rowscale, float* dW, float* csumA, float* csumB,
                          int64_t M, int Ka, int Nb, int64_t lda, int64_t ldb, int64_t ldo, cudaStream_t st) {
  const int ti = (int)ceil_div(Ka, 64), tj = (int)ceil_div(Nb, 64);
  int64_t split = ceil_div((int64_t)4 * sm_count(), (int64_t)ti * tj);
  const int64_t max_split = ceil_div(M, 256);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  int64_t rps = ceil_div(M, split);
  rps = ceil_div(rps, 16) * 16;
  split = ceil_div(M, rps);
  dim3 grid(ti, tj, (unsigned)split);
  gemm_tn_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)A, (const T*)B, rowscale, dW, csumA, csumB, M, Ka, Nb, lda, ldb,
                                               ldo, rps);
  return check_launch("gemm_tn");
}

// skinny-K residual GEMM (gemm_skinny.cu); -1 = not such a call
int gemm_skinny_bf16(const void* A, const void* W, int64_t M, int N, int K, int64_t lda, int64_t ldw, int groups,
                     const EpiParams& ep, cudaStream_t st);
// bf16 tensor-core version (gemm_tn_tc.cu); -1 = arguments do not fit it
int gemm_tn_bf16(const void* A, const void* B, const float* rowscale, float* dW, float* csumA, int64_t M, int Ka, int Nb,
                 int64_t lda, int64_t ldb, int64_t ldo, cudaStream_t st);

}  // namespace stgcma

using namespace stgcma;

extern "C" int stgcma_gemm(const void* A, const void* W, const float* bias, const void* res1, const void* res2,
                           const float* rowscale, const void* aux, void* out, void* out_pre, int64_t M, int N, int K,
                           int64_t lda, int64_t ldw, int64_t ldo, int groups, int epilogue, int dtype, int out_dtype,
                           void* stream) {
  STG_REQUIRE(out_dtype == dtype || out_dtype == STGCMA_F32, "gemm: out_dtype must be the activation dtype or fp32");
  STG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%lld N=%d K=%d", (long long)M, N, K);
  STG_REQUIRE(groups >= 1 && M % groups == 0, "gemm: M=%lld not divisible by groups=%d", (long long)M, groups);
  STG_REQUIRE(epilogue >= STGCMA_EPI_LINEAR && epilogue <= STGCMA_EPI_QGELU_G, "gemm: bad epilogue %d", epilogue);
  STG_REQUIRE((epilogue != STGCMA_EPI_DGELU && epilogue != STGCMA_EPI_MULAUX && epilogue != STGCMA_EPI_DQGELU) || aux,
              "gemm: epilogue %d needs aux", epilogue);
  STG_REQUIRE(A && W && out, "gemm: null operand");
  cudaStream_t st = (cudaStream_t)stream;
  EpiParams ep;
  ep.bias = bias; ep.res1 = res1; ep.res2 = res2; ep.rowscale = rowscale; ep.aux = aux;
  ep.out = out; ep.out_pre = out_pre; ep.ldo = ldo; ep.mode = epilogue;
  ep.out_f32 = (out_dtype == STGCMA_F32);
  if (dtype == STGCMA_F32) {
    const int64_t gr = M / groups;
    dim3 grid((unsigned)ceil_div(gr, 64), (unsigned)ceil_div(N, 64), groups);
    gemm_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)A, (const float*)W, gr, N, K, lda, ldw, ep);
    return check_launch("gemm(fp32)");
  }
  STG_REQUIRE(dtype == STGCMA_BF16, "gemm: bad dtype %d", dtype);
  STG_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm(bf16): K=%d lda=%lld ldw=%lld must be multiples of 8", K,
              (long long)lda, (long long)ldw);
  STG_REQUIRE(N % 8 == 0 && ldo % 8 == 0, "gemm(bf16): N=%d ldo=%lld must be multiples of 8", N, (long long)ldo);
  STG_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                  ((uintptr_t)res1 & 15) == 0 && ((uintptr_t)res2 & 15) == 0 && ((uintptr_t)aux & 15) == 0 &&
                  ((uintptr_t)out_pre & 15) == 0 && ((uintptr_t)bias & 15) == 0,
              "gemm(bf16): operands must be 16-byte aligned");
  if (ep.res1 && K <= 128) {   // adapter up-projection + residual adds: an HBM stream, not a tensor-core problem
    const int rc = gemm_skinny_bf16(A, W, M, N, K, lda, ldw, groups, ep, st);
    if (rc != -1) return rc;
  }
  return tc::launch(A, W, M, N, K, lda, ldw, groups, ep, st);
}

extern "C" int stgcma_gemm_tn(const void* A, const void* B, const float* rowscale, float* dW, float* csumA, float* csumB,
                              int64_t M, int Ka, int Nb, int64_t lda, int64_t ldb, int64_t ldo, int dtype, void* stream) {
  STG_REQUIRE(M > 0 && Ka > 0 && Nb > 0, "gemm_tn: empty problem");
  STG_REQUIRE(A && B && dW, "gemm_tn: null operand");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == STGCMA_F32) return gemm_tn_launch<float>(A, B, rowscale, dW, csumA, csumB, M, Ka, Nb, lda, ldb, ldo, st);
  if (dtype == STGCMA_BF16) {
    if (!csumB) {
      const int rc = gemm_tn_bf16(A, B, rowscale, dW, csumA, M, Ka, Nb, lda, ldb, ldo, st);
      if (rc != -1) return rc;
    }
    return gemm_tn_launch<bf16>(A, B, rowscale, dW, csumA, csumB, M, Ka, Nb, lda, ldb, ldo, st);
  }
  STG_REQUIRE(false, "gemm_tn: bad dtype %d", dtype);
}
