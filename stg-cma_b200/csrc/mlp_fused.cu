// Fused frozen-MLP kernels of the (G) sub-block (Swin_AVE.py:790-797), bf16 production mode.
//
//   forward :  xn = LN2(x)  ->  hid = GELU(xn W1^T + b1)  ->  y = hid W2^T + b2  ->  hpre = y Wa^T + ba, h = GELU(hpre)
//              (norm2 :353/:790, mlp.fc1 / act / fc2 :121-127, S_Adapter[_Audio].D_fc1 + act :796-797)
//   backward:  dy = dhpre Wa + dxb ;  dfpre = (dy W2) * gelu'(xn W1^T + b1) ;  dxn = dfpre W1       (recompute, no saved hidden)
//
// The 4C-wide hidden activation is the largest tensor of a Swin block; unfused it is written and re-read three
// times per step (pre-activation, activation, its gradient).  Here it never leaves the SM: a CTA owns a 128-row
// tile, keeps LN2(x) (bf16, K-major SWIZZLE_128B) resident in shared memory as the A operand, and walks the hidden
// dimension in chunks of HB columns:
//       G1(j): acc1[j&1] (TMEM) = xn . W1[chunk j]^T              tcgen05.mma, K = C
//       epi(j): TMEM -> registers -> +b1 -> GELU -> bf16 -> shared memory (K-major SWIZZLE_128B A operand)
//       G2(j): acc2 (TMEM, 128 x C) += hid_j . W2[:, chunk j]^T   tcgen05.mma, K = HB
// with G1(j+2) issued right behind G2(j), so the tensor pipe computes the next chunk while the 8 epilogue warps
// run the activation of the current one.  Weights stream through a TMA ring (they are L2 resident: every tile
// reads the same ones).  After the last chunk y goes out by TMA store from the (re-used) A tile, which is also
// the A operand of the small adapter GEMM G3 (N = rp).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 = workers
// (LayerNorm prologue, activation epilogues, output epilogue); TMEM lane quarter of a worker = warp % 4.
#include <stdlib.h>
#include "tc_common.cuh"
#include "gelu_poly.cuh"

namespace stgcma {
namespace mlp {

using namespace tc;

// worker warps = 4 * NWQ (NWQ per TMEM lane quarter); block = 64 + 128 * NWQ threads.  NWQ = 4 (16 workers) when the
// CTA is alone on its SM: with 8 workers the prologue / epilogues ran at ~0.25 IPC per scheduler (2 warps of mostly
// serial code per scheduler; phase trace in profiles/r01c) -- latency-bound, not issue-bound.
constexpr int kBlkBytes = BM * 128;   // one [128 rows x 64 bf16] swizzled k-block

// Debug tracing (stgcma_debug_trace): CTA 0 records clock64() at phase boundaries, [role][event] with 512 slots per
// role (0 = worker warp 0, 1 = MMA issuer, 2 = TMA producer).  Null in production: one predicated branch per phase.
static unsigned long long* g_trace = nullptr;
struct Tracer {
  unsigned long long* p;
  int n;
  __device__ __forceinline__ Tracer(unsigned long long* base, int role) : p(base && blockIdx.x == 0 ? base + role * 512 : nullptr), n(0) {}
  __device__ __forceinline__ void mark(int id) {
    if (p && n < 511) { p[n++] = ((unsigned long long)id << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); }
  }
};

struct FwdParams {
  const float* x;        // [M, C] fp32 residual stream
  const float* gamma;    // [C]
  const float* beta;     // [C]
  const float* b1;       // [H]
  const float* b2;       // [C]
  const float* ba;       // [groups, rp]
  bf16* hpre;            // [M, rp]
  bf16* h;               // [M, rp]
  float* mean;           // [M]
  float* rstd;           // [M]
  int64_t group_rows;
  int groups, C, H, rp;
  int HB, NC, KB, HKB;   // hidden chunk width, number of chunks, k-blocks of C, k-blocks of HB
  int nsplit, CN;        // G2 is issued as nsplit MMAs of N = CN = C / nsplit (UMMA N <= 256)
  int m_tiles_g, num_tiles;
  int ring;
  uint32_t slot_bytes, tmem_cols;
  unsigned long long* trace;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NW>
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); }

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// byte offset of element (row, col) of a K-major SWIZZLE_128B operand made of [128 x 64] bf16 k-blocks; col % 8 == 0
__device__ __forceinline__ uint32_t sw_off(int row, int col) {
  return (uint32_t)((col >> 6) * kBlkBytes + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint32_t idesc_f16(int N) {   // D = f32, A = B = bf16, K-major both, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm prologue: rows [row0, row0 + 128) of the group -> bf16 K-major swizzled A tile.
// Worker warp w owns tile rows 16 w .. 16 w + 15; a row is read with coalesced 128-bit loads (lane = 4 columns,
// NV = ceil(C / 128) vectors per lane), 4 rows in flight per lane.  `affine` = apply gamma / beta.
// ---------------------------------------------------------------------------------------------
template <int NV, int NW>
__device__ __forceinline__ void ln_rows_to_smem(const float* __restrict__ xg, int64_t grow0, int64_t rows_valid, int C,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                uint8_t* atile, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                int wid, int lane, Tracer* tr = nullptr) {
  // rows in flight per lane: every load of a batch is issued before the first use (the prologue is a pure
  // latency chain otherwise: measured 8.4 us per tile with 4 rows in flight, profiles/r01c)
  constexpr int RPW = BM / NW;                                   // rows per worker warp
  constexpr int RB = (RPW == 8 && NV > 1) ? 4 : 8;               // register budget: 112 / thread with 16 workers
  float4 gm[NV], bt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 4 * lane + 128 * i;
    gm[i] = c < C ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = c < C ? __ldg(reinterpret_cast<const float4*>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invC = 1.0f / (float)C;
  float my_mean = 0.f, my_rstd = 0.f;
#pragma unroll 1
  for (int r0 = 0; r0 < RPW; r0 += RB) {
    float4 v[RB][NV];
    float mu[RB], rs[RB];
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      const int r = wid * RPW + r0 + rr;
      const bool ok = r < rows_valid;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = 4 * lane + 128 * i;
        v[rr][i] = (ok && c < C) ? __ldg(reinterpret_cast<const float4*>(xg + (grow0 + r) * (int64_t)C + c))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (mean_in) {
        mu[rr] = ok ? __ldg(mean_in + grow0 + r) : 0.f;
        rs[rr] = ok ? __ldg(rstd_in + grow0 + r) : 0.f;
      }
    }
    if (tr) tr->mark(20);
    if (!mean_in) {
      // the RB reductions are independent: their shuffles interleave
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[rr][i].x + v[rr][i].y) + (v[rr][i].z + v[rr][i].w);
        mu[rr] = s;
      }
      if (tr) tr->mark(21);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) mu[rr] += __shfl_xor_sync(0xffffffffu, mu[rr], o);
      }
      if (tr) tr->mark(22);
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        mu[rr] *= invC;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (4 * lane + 128 * i < C) {
            const float a = v[rr][i].x - mu[rr], b = v[rr][i].y - mu[rr], c2 = v[rr][i].z - mu[rr], d = v[rr][i].w - mu[rr];
            q += (a * a + b * b) + (c2 * c2 + d * d);
          }
        }
        rs[rr] = q;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) rs[rr] += __shfl_xor_sync(0xffffffffu, rs[rr], o);
      }
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        rs[rr] = rsqrtf(rs[rr] * invC + 1e-5f);
        if (lane == r0 + rr) { my_mean = mu[rr]; my_rstd = rs[rr]; }
      }
    }
    if (tr) tr->mark(23);
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      const int r = wid * RPW + r0 + rr;
      const bool ok = r < rows_valid;
      const float mean = mu[rr], rstd = rs[rr];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = 4 * lane + 128 * i;
        if (c < C) {
          uint2 o;
          if (ok) {
            const float a = (v[rr][i].x - mean) * rstd * gm[i].x + bt[i].x, b = (v[rr][i].y - mean) * rstd * gm[i].y + bt[i].y;
            const float c2 = (v[rr][i].z - mean) * rstd * gm[i].z + bt[i].z, d = (v[rr][i].w - mean) * rstd * gm[i].w + bt[i].w;
            o = make_uint2(pack2(a, b), pack2(c2, d));
          } else {
            o = make_uint2(0u, 0u);
          }
          *reinterpret_cast<uint2*>(atile + sw_off(r, c & ~7) + (c & 7) * 2) = o;
        }
      }
    }
  }
  if (tr) tr->mark(24);
  if (mean_out && lane < RPW) {
    const int r = wid * RPW + lane;
    if (r < rows_valid) {
      mean_out[grow0 + r] = my_mean;
      rstd_out[grow0 + r] = my_rstd;
    }
  }
}

// L2 prefetch of the fp32 rows of the tile this CTA will process next (hides the HBM latency of its prologue)
__device__ __forceinline__ void prefetch_rows_l2(const float* xg, int64_t grow0, int64_t rows_valid, int C, int tid, int nthr) {
  const int64_t nrows = rows_valid < BM ? rows_valid : BM;
  if (nrows <= 0) return;
  const int64_t lines = (nrows * C * 4 + 127) >> 7;
  const char* base = reinterpret_cast<const char*>(xg + grow0 * (int64_t)C);
  for (int64_t i = tid; i < lines; i += nthr) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (i << 7)));
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
template <int OCC, int NV, int NWQ>
__global__ void __launch_bounds__(64 + 128 * NWQ, OCC)
mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
               const __grid_constant__ CUtensorMap tmWa, const __grid_constant__ CUtensorMap tmY, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_tile = smem;                                        // KB k-blocks: LN2(x), later y
  uint8_t* h_buf = a_tile + (size_t)p.KB * kBlkBytes;            // 2 x HKB k-blocks: GELU(hidden chunk)
  uint8_t* ring = h_buf + (size_t)2 * p.HKB * kBlkBytes;         // weight k-blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.ring * p.slot_bytes);
  uint64_t* r_full = bars;                 // [ring]  TMA -> MMA
  uint64_t* r_empty = bars + p.ring;       // [ring]  MMA -> TMA
  uint64_t* fixed = bars + 2 * p.ring;
  uint64_t* a_full = fixed + 0;            // workers -> MMA : A tile written
  uint64_t* acc1_full = fixed + 1;         // [2] MMA -> workers
  uint64_t* h_full = fixed + 3;            // [2] workers -> MMA : hidden chunk in smem, acc1[b] drained
  uint64_t* acc2_full = fixed + 5;         // MMA -> workers
  uint64_t* y_full = fixed + 6;            // workers -> MMA : y tile in smem
  uint64_t* acc3_full = fixed + 7;         // MMA -> workers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fixed + 8);
  constexpr int NW = 4 * NWQ;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW2)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWa)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    for (int s = 0; s < p.ring; ++s) {
      mbar_init(smem_u32(r_full + s), 1);
      mbar_init(smem_u32(r_empty + s), 1);
    }
    mbar_init(smem_u32(a_full), NW);
    mbar_init(smem_u32(acc1_full + 0), 1);
    mbar_init(smem_u32(acc1_full + 1), 1);
    mbar_init(smem_u32(h_full + 0), NW);
    mbar_init(smem_u32(h_full + 1), NW);
    mbar_init(smem_u32(acc2_full), 1);
    mbar_init(smem_u32(y_full), NW);
    mbar_init(smem_u32(acc3_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
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
  const uint32_t t_acc1 = tmem_base;                          // 2 x HB columns
  const uint32_t t_acc2 = tmem_base + 2u * (uint32_t)p.HB;    // C columns
  const uint32_t t_acc3 = tmem_base;                          // rp columns (acc1[0] is free by then)

  if (warp == 0) {
    // ===================== TMA producer: weight k-blocks in the MMA issue order =====================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      Tracer tr(p.trace, 2);
      auto put = [&](const CUtensorMap* map, uint32_t bytes, int c0, int c1, int c2) {
        mbar_wait_relaxed(smem_u32(r_empty + slot), phase ^ 1u);
        const uint32_t fb = smem_u32(r_full + slot);
        mbar_expect_tx(fb, bytes);
        tma_load_3d(smem_u32(ring + (size_t)slot * p.slot_bytes), map, fb, c0, c1, c2);
        if (++slot == (uint32_t)p.ring) { slot = 0; phase ^= 1u; }
        tr.mark(1);
      };
      auto put_w1 = [&](int j) {
        for (int kb = 0; kb < p.KB; ++kb) put(&tmW1, (uint32_t)p.HB * 128u, kb * BK, j * p.HB, 0);
      };
      auto put_w2 = [&](int j) {
        for (int kb = 0; kb < p.HKB; ++kb)
          for (int ns = 0; ns < p.nsplit; ++ns) put(&tmW2, (uint32_t)p.CN * 128u, j * p.HB + kb * BK, ns * p.CN, 0);
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int g = tile / p.m_tiles_g;
        put_w1(0);
        if (p.NC > 1) put_w1(1);
        for (int j = 0; j < p.NC; ++j) {
          put_w2(j);
          if (j + 2 < p.NC) put_w1(j + 2);
        }
        for (int kb = 0; kb < p.KB; ++kb) put(&tmWa, (uint32_t)p.rp * 128u, kb * BK, 0, g);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {   // the whole warp runs the (uniform) issue loop; one elected lane executes each tcgen05 instruction
      uint32_t slot = 0, phase = 0;
      uint32_t ph_a = 0, ph_h[2] = {0, 0}, ph_y = 0;
      Tracer tr(lane == 0 ? p.trace : nullptr, 1);
      const uint32_t id1 = idesc_f16(p.HB), id2 = idesc_f16(p.CN), id3 = idesc_f16(p.rp);
      const uint32_t a_base = smem_u32(a_tile), h_base = smem_u32(h_buf);
      auto take = [&]() -> uint32_t {
        tr.mark(10);
        mbar_wait(smem_u32(r_full + slot), phase);
        tc_fence_after();
        tr.mark(11);
        return smem_u32(ring + (size_t)slot * p.slot_bytes);
      };
      auto release = [&]() {
        mma_commit_elect(smem_u32(r_empty + slot));
        if (++slot == (uint32_t)p.ring) { slot = 0; phase ^= 1u; }
        tr.mark(12);
      };
      // D[128 x N] (+)= A[kb-th k-block of the resident tile] . B[slot]^T over the valid k-steps of the block
      auto gemm_kblock = [&](uint32_t d, uint32_t a_blk, uint32_t b_addr, uint32_t idesc, int ksteps, bool first) {
        const uint64_t ad = make_sdesc(a_blk), bd = make_sdesc(b_addr);
        for (int k = 0; k < ksteps; ++k)
          mma_bf16_elect(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)(!first || k != 0));
      };
      auto issue_g1 = [&](int j) {   // acc1[j & 1] = xn . W1[chunk j]^T
        const uint32_t d = t_acc1 + (uint32_t)((j & 1) * p.HB);
        for (int kb = 0; kb < p.KB; ++kb) {
          const uint32_t b = take();
          int ks = (p.C - kb * BK) >> 4;
          if (ks > 4) ks = 4;
          gemm_kblock(d, a_base + (uint32_t)kb * kBlkBytes, b, id1, ks, kb == 0);
          release();
        }
        mma_commit_elect(smem_u32(acc1_full + (j & 1)));
      };
      auto issue_g2 = [&](int j) {   // acc2 += hid_j . W2[:, chunk j]^T
        for (int kb = 0; kb < p.HKB; ++kb)
          for (int ns = 0; ns < p.nsplit; ++ns) {
            const uint32_t b = take();
            gemm_kblock(t_acc2 + (uint32_t)(ns * p.CN), h_base + (uint32_t)(((j & 1) * p.HKB + kb)) * kBlkBytes, b, id2, 4,
                        j == 0 && kb == 0);
            release();
          }
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        tr.mark(1);
        mbar_wait(smem_u32(a_full), ph_a);
        ph_a ^= 1u;
        tc_fence_after();
        tr.mark(2);
        issue_g1(0);
        if (p.NC > 1) issue_g1(1);
        tr.mark(3);
        for (int j = 0; j < p.NC; ++j) {
          const int b = j & 1;
          mbar_wait(smem_u32(h_full + b), ph_h[b]);
          ph_h[b] ^= 1u;
          tc_fence_after();
          tr.mark(4);
          issue_g2(j);
          if (j + 2 < p.NC) issue_g1(j + 2);
          tr.mark(5);
        }
        mma_commit_elect(smem_u32(acc2_full));
        mbar_wait(smem_u32(y_full), ph_y);
        ph_y ^= 1u;
        tc_fence_after();
        tr.mark(6);
        for (int kb = 0; kb < p.KB; ++kb) {   // acc3 = y . Wa[g]^T
          const uint32_t b = take();
          int ks = (p.C - kb * BK) >> 4;
          if (ks > 4) ks = 4;
          gemm_kblock(t_acc3, a_base + (uint32_t)kb * kBlkBytes, b, id3, ks, kb == 0);
          release();
        }
        mma_commit_elect(smem_u32(acc3_full));
      }
    }
  } else {
    // ===================== workers =====================
    const int wid = warp - 2;
    const int q = warp & 3;            // TMEM lane quarter
    const int par = wid >> 2;          // which of the NWQ warps of the quarter
    const int trow = q * 32 + lane;    // tile row of this thread in TMEM-lane order
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t ph1[2] = {0, 0}, ph2 = 0, ph3 = 0;
    Tracer tr(wid == 0 && lane == 0 ? p.trace : nullptr, 0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int g = tile / p.m_tiles_g;
      const int m_blk = tile - g * p.m_tiles_g;
      const int64_t grow0 = (int64_t)g * p.group_rows + (int64_t)m_blk * BM;
      const int64_t rows_valid = p.group_rows - (int64_t)m_blk * BM;   // may exceed 128
      // ---- LayerNorm prologue ----
      tr.mark(1);
      {
        const int nt = tile + gridDim.x;
        if (nt < p.num_tiles) {
          const int ng = nt / p.m_tiles_g, nm = nt - ng * p.m_tiles_g;
          prefetch_rows_l2(p.x, (int64_t)ng * p.group_rows + (int64_t)nm * BM, p.group_rows - (int64_t)nm * BM, p.C,
                           wid * 32 + lane, NW * 32);
        }
      }
      ln_rows_to_smem<NV, NW>(p.x, grow0, rows_valid, p.C, p.gamma, p.beta, a_tile, p.mean, p.rstd, nullptr, nullptr, wid, lane, &tr);
      tr.mark(25);
      fence_async_smem();
      tr.mark(26);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(a_full));
      tr.mark(2);
      // ---- activation epilogues ----
      for (int j = 0; j < p.NC; ++j) {
        const int b = j & 1;
        mbar_wait_relaxed(smem_u32(acc1_full + b), ph1[b]);
        ph1[b] ^= 1u;
        tc_fence_after();
        tr.mark(3);
        uint8_t* hb = h_buf + (size_t)b * p.HKB * kBlkBytes;
        for (int c0 = par * 16; c0 < p.HB; c0 += 16 * NWQ) {
          float v[32];
          tmem_ld16(t_acc1 + (uint32_t)(b * p.HB + c0) + lane_off, v);
          const float* bp = p.b1 + j * p.HB + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bp + i));
            v[i] = gelu_tanh(v[i] + bb.x);
            v[i + 1] = gelu_tanh(v[i + 1] + bb.y);
            v[i + 2] = gelu_tanh(v[i + 2] + bb.z);
            v[i + 3] = gelu_tanh(v[i + 3] + bb.w);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            const uint4 o = make_uint4(pack2(v[i], v[i + 1]), pack2(v[i + 2], v[i + 3]), pack2(v[i + 4], v[i + 5]),
                                       pack2(v[i + 6], v[i + 7]));
            *reinterpret_cast<uint4*>(hb + sw_off(trow, c0 + i)) = o;
          }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(h_full + b));
        tr.mark(4);
      }
      // ---- y = acc2 + b2 -> A tile (bf16, swizzled) -> TMA store; also the A operand of the adapter GEMM ----
      mbar_wait_relaxed(smem_u32(acc2_full), ph2);
      ph2 ^= 1u;
      tc_fence_after();
      tr.mark(5);
      for (int c0 = par * 16; c0 < p.C; c0 += 16 * NWQ) {
        float v[32];
        tmem_ld16(t_acc2 + (uint32_t)c0 + lane_off, v);
        const float* bp = p.b2 + c0;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp + i)), b1 = __ldg(reinterpret_cast<const float4*>(bp + i + 4));
          const uint4 o = make_uint4(pack2(v[i] + b0.x, v[i + 1] + b0.y), pack2(v[i + 2] + b0.z, v[i + 3] + b0.w),
                                     pack2(v[i + 4] + b1.x, v[i + 5] + b1.y), pack2(v[i + 6] + b1.z, v[i + 7] + b1.w));
          *reinterpret_cast<uint4*>(a_tile + sw_off(trow, c0 + i)) = o;
        }
      }
      tc_fence_before();
      fence_async_smem();
      worker_bar<NW>();
      if (wid == 0 && lane == 0) {
        for (int kb = 0; kb < p.KB; ++kb) tma_store_3d(&tmY, smem_u32(a_tile + (size_t)kb * kBlkBytes), kb * BK, m_blk * BM, g);
        tma_store_commit();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(y_full));
      tr.mark(6);
      // ---- adapter epilogue: hpre = acc3 + ba, h = GELU(hpre) ----
      mbar_wait_relaxed(smem_u32(acc3_full), ph3);
      ph3 ^= 1u;
      tc_fence_after();
      tr.mark(7);
      const bool row_ok = trow < rows_valid;
      for (int c0 = par * 16; c0 < p.rp; c0 += 16 * NWQ) {
        float v[32];
        tmem_ld16(t_acc3 + (uint32_t)c0 + lane_off, v);
        const float* bp = p.ba + g * p.rp + c0;
        float pre[16], act[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pre[i] = v[i] + __ldg(bp + i);
          act[i] = gelu_tanh(pre[i]);
        }
        if (row_ok) {
          const int64_t o = (grow0 + trow) * (int64_t)p.rp + c0;
          float t8[8];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = pre[hh * 8 + i];
            store8(p.hpre + o + hh * 8, t8);
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = act[hh * 8 + i];
            store8(p.h + o + hh * 8, t8);
          }
        }
      }
      tc_fence_before();
      tr.mark(8);
      if (wid == 0 && lane == 0) tma_store_wait_read();   // the y tile has left shared memory
      worker_bar<NW>();                                       // A tile and acc1 / acc3 are free for the next tile
      tr.mark(9);
    }
    if (wid == 0 && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// backward kernel:  dxn = [ ((dhpre Wa + dxb) W2) * gelu'(LN2(x) W1^T + b1) ] W1
//   resident A operands: XN = LN2(x) (recomputed from x and the saved statistics) and DY = dhpre Wa + dxb;
//   per hidden chunk j:  G1a: accF[j&1] = XN . W1[chunk]^T      G1b: accH[j&1] = DY . W2t[chunk]^T      (K = C)
//                        epi : dF = accH * gelu'(accF + b1) -> bf16 -> shared memory
//                        G2  : accX (128 x C) += dF . W1t[:, chunk]^T                                    (K = HB)
//   dhpre's tile (the A operand of G0: accX = dhpre . Wat^T, K = rp) borrows the dF[0] buffer.
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  const float* x;        // [M, C] fp32: input of LN2
  const float* gamma;
  const float* beta;
  const float* mean;     // [M] saved by the forward
  const float* rstd;
  const float* b1;       // [H]
  int64_t group_rows;
  int groups, C, H, rp;
  int HB, NC, KB, HKB, RKB;   // RKB: k-blocks of rp
  int nsplit, CN;
  int m_tiles_g, num_tiles;
  int ring;
  uint32_t slot_bytes, tmem_cols;
};

template <int NV, int NWQ>
__global__ void __launch_bounds__(64 + 128 * NWQ, 1)
mlp_bwd_kernel(const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDxb,
               const __grid_constant__ CUtensorMap tmWat, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2t, const __grid_constant__ CUtensorMap tmW1t,
               const __grid_constant__ CUtensorMap tmDxn, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xn_tile = smem;                                       // KB k-blocks: LN2(x); at the end dxn
  uint8_t* dy_tile = xn_tile + (size_t)p.KB * kBlkBytes;         // KB k-blocks: dxb, then dy
  uint8_t* df_buf = dy_tile + (size_t)p.KB * kBlkBytes;          // 2 x HKB k-blocks: dF chunk; [0] first holds dhpre
  const int df_blocks = 2 * p.HKB > p.RKB ? 2 * p.HKB : p.RKB;
  uint8_t* ring = df_buf + (size_t)df_blocks * kBlkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.ring * p.slot_bytes);
  uint64_t* r_full = bars;
  uint64_t* r_empty = bars + p.ring;
  uint64_t* fixed = bars + 2 * p.ring;
  uint64_t* in_full = fixed + 0;           // TMA -> MMA, workers : dhpre and dxb tiles landed
  uint64_t* in_empty = fixed + 1;          // MMA -> TMA : every MMA of the tile is complete
  uint64_t* dyacc_full = fixed + 2;        // MMA -> workers : accX = dhpre . Wat^T
  uint64_t* ady_full = fixed + 3;          // workers -> MMA : XN and DY written
  uint64_t* acc_full = fixed + 4;          // [2] MMA -> workers : accF[b], accH[b]
  uint64_t* df_full = fixed + 6;           // [2] workers -> MMA
  uint64_t* dxn_full = fixed + 8;          // MMA -> workers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fixed + 9);
  constexpr int NW = 4 * NWQ;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    const CUtensorMap* maps[7] = {&tmDh, &tmDxb, &tmWat, &tmW1, &tmW2t, &tmW1t, &tmDxn};
    for (int i = 0; i < 7; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps[i])) : "memory");
    for (int s = 0; s < p.ring; ++s) {
      mbar_init(smem_u32(r_full + s), 1);
      mbar_init(smem_u32(r_empty + s), 1);
    }
    mbar_init(smem_u32(in_full), 1);
    mbar_init(smem_u32(in_empty), 1);
    mbar_init(smem_u32(dyacc_full), 1);
    mbar_init(smem_u32(ady_full), NW);
    mbar_init(smem_u32(acc_full + 0), 1);
    mbar_init(smem_u32(acc_full + 1), 1);
    mbar_init(smem_u32(df_full + 0), NW);
    mbar_init(smem_u32(df_full + 1), NW);
    mbar_init(smem_u32(dxn_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
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
  const uint32_t t_accF = tmem_base;                          // 2 x HB
  const uint32_t t_accH = tmem_base + 2u * (uint32_t)p.HB;    // 2 x HB
  const uint32_t t_accX = tmem_base + 4u * (uint32_t)p.HB;    // C

  if (warp == 0) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0, ph_in = 0;
      auto put = [&](const CUtensorMap* map, uint32_t bytes, int c0, int c1, int c2) {
        mbar_wait_relaxed(smem_u32(r_empty + slot), phase ^ 1u);
        const uint32_t fb = smem_u32(r_full + slot);
        mbar_expect_tx(fb, bytes);
        tma_load_3d(smem_u32(ring + (size_t)slot * p.slot_bytes), map, fb, c0, c1, c2);
        if (++slot == (uint32_t)p.ring) { slot = 0; phase ^= 1u; }
      };
      auto put_g1 = [&](int j) {
        for (int kb = 0; kb < p.KB; ++kb) put(&tmW1, (uint32_t)p.HB * 128u, kb * BK, j * p.HB, 0);
        for (int kb = 0; kb < p.KB; ++kb) put(&tmW2t, (uint32_t)p.HB * 128u, kb * BK, j * p.HB, 0);
      };
      auto put_g2 = [&](int j) {
        for (int kb = 0; kb < p.HKB; ++kb)
          for (int ns = 0; ns < p.nsplit; ++ns) put(&tmW1t, (uint32_t)p.CN * 128u, j * p.HB + kb * BK, ns * p.CN, 0);
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int g = tile / p.m_tiles_g;
        const int m_blk = tile - g * p.m_tiles_g;
        // activation tiles of this CTA: free once every MMA of the previous tile has completed
        mbar_wait(smem_u32(in_empty), ph_in ^ 1u);
        ph_in ^= 1u;
        const uint32_t fb = smem_u32(in_full);
        mbar_expect_tx(fb, (uint32_t)(p.RKB + p.KB) * kBlkBytes);
        for (int kb = 0; kb < p.RKB; ++kb) tma_load_3d(smem_u32(df_buf + (size_t)kb * kBlkBytes), &tmDh, fb, kb * BK, m_blk * BM, g);
        for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(smem_u32(dy_tile + (size_t)kb * kBlkBytes), &tmDxb, fb, kb * BK, m_blk * BM, g);
        // weights, in the MMA issue order
        for (int kb = 0; kb < p.RKB; ++kb)
          for (int ns = 0; ns < p.nsplit; ++ns) put(&tmWat, (uint32_t)p.CN * 128u, kb * BK, ns * p.CN, g);
        put_g1(0);
        if (p.NC > 1) put_g1(1);
        for (int j = 0; j < p.NC; ++j) {
          put_g2(j);
          if (j + 2 < p.NC) put_g1(j + 2);
        }
      }
    }
  } else if (warp == 1) {
    {   // whole-warp uniform issue (see mma_bf16_elect)
      uint32_t slot = 0, phase = 0;
      uint32_t ph_in = 0, ph_ady = 0, ph_df[2] = {0, 0};
      const uint32_t id1 = idesc_f16(p.HB), id2 = idesc_f16(p.CN);
      const uint32_t xn_base = smem_u32(xn_tile), dy_base = smem_u32(dy_tile), df_base = smem_u32(df_buf);
      auto take = [&]() -> uint32_t {
        mbar_wait(smem_u32(r_full + slot), phase);
        tc_fence_after();
        return smem_u32(ring + (size_t)slot * p.slot_bytes);
      };
      auto release = [&]() {
        mma_commit_elect(smem_u32(r_empty + slot));
        if (++slot == (uint32_t)p.ring) { slot = 0; phase ^= 1u; }
      };
      auto gemm_kblock = [&](uint32_t d, uint32_t a_blk, uint32_t b_addr, uint32_t idesc, int ksteps, bool first) {
        const uint64_t ad = make_sdesc(a_blk), bd = make_sdesc(b_addr);
        for (int k = 0; k < ksteps; ++k)
          mma_bf16_elect(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)(!first || k != 0));
      };
      auto issue_g1 = [&](int j) {
        const int b = j & 1;
        for (int which = 0; which < 2; ++which) {
          const uint32_t d = (which ? t_accH : t_accF) + (uint32_t)(b * p.HB);
          const uint32_t a = which ? dy_base : xn_base;
          for (int kb = 0; kb < p.KB; ++kb) {
            const uint32_t bs = take();
            int ks = (p.C - kb * BK) >> 4;
            if (ks > 4) ks = 4;
            gemm_kblock(d, a + (uint32_t)kb * kBlkBytes, bs, id1, ks, kb == 0);
            release();
          }
        }
        mma_commit_elect(smem_u32(acc_full + b));
      };
      auto issue_g2 = [&](int j) {
        for (int kb = 0; kb < p.HKB; ++kb)
          for (int ns = 0; ns < p.nsplit; ++ns) {
            const uint32_t bs = take();
            gemm_kblock(t_accX + (uint32_t)(ns * p.CN), df_base + (uint32_t)((j & 1) * p.HKB + kb) * kBlkBytes, bs, id2, 4,
                        j == 0 && kb == 0);
            release();
          }
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        // G0: accX = dhpre . Wat^T  (K = rp)
        mbar_wait(smem_u32(in_full), ph_in);
        ph_in ^= 1u;
        tc_fence_after();
        for (int kb = 0; kb < p.RKB; ++kb)
          for (int ns = 0; ns < p.nsplit; ++ns) {
            const uint32_t bs = take();
            int ks = (p.rp - kb * BK) >> 4;
            if (ks > 4) ks = 4;
            gemm_kblock(t_accX + (uint32_t)(ns * p.CN), df_base + (uint32_t)kb * kBlkBytes, bs, id2, ks, kb == 0);
            release();
          }
        mma_commit_elect(smem_u32(dyacc_full));
        mbar_wait(smem_u32(ady_full), ph_ady);
        ph_ady ^= 1u;
        tc_fence_after();
        issue_g1(0);
        if (p.NC > 1) issue_g1(1);
        for (int j = 0; j < p.NC; ++j) {
          const int b = j & 1;
          mbar_wait(smem_u32(df_full + b), ph_df[b]);
          ph_df[b] ^= 1u;
          tc_fence_after();
          issue_g2(j);
          if (j + 2 < p.NC) issue_g1(j + 2);
        }
        mma_commit_elect(smem_u32(dxn_full));
        mma_commit_elect(smem_u32(in_empty));
      }
    }
  } else {
    const int wid = warp - 2;
    const int q = warp & 3;
    const int par = wid >> 2;
    const int trow = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t ph_in = 0, ph_dyacc = 0, ph_acc[2] = {0, 0}, ph_dxn = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int g = tile / p.m_tiles_g;
      const int m_blk = tile - g * p.m_tiles_g;
      const int64_t grow0 = (int64_t)g * p.group_rows + (int64_t)m_blk * BM;
      const int64_t rows_valid = p.group_rows - (int64_t)m_blk * BM;
      {
        const int nt = tile + gridDim.x;
        if (nt < p.num_tiles) {
          const int ng = nt / p.m_tiles_g, nm = nt - ng * p.m_tiles_g;
          prefetch_rows_l2(p.x, (int64_t)ng * p.group_rows + (int64_t)nm * BM, p.group_rows - (int64_t)nm * BM, p.C,
                           wid * 32 + lane, NW * 32);
        }
      }
      // ---- XN = LN2(x) from the saved statistics ----
      ln_rows_to_smem<NV, NW>(p.x, grow0, rows_valid, p.C, p.gamma, p.beta, xn_tile, nullptr, nullptr, p.mean, p.rstd, wid, lane);
      // ---- DY = accX + dxb (in place over the TMA-loaded dxb tile) ----
      mbar_wait_relaxed(smem_u32(in_full), ph_in);
      ph_in ^= 1u;
      mbar_wait_relaxed(smem_u32(dyacc_full), ph_dyacc);
      ph_dyacc ^= 1u;
      tc_fence_after();
      for (int c0 = par * 16; c0 < p.C; c0 += 16 * NWQ) {
        float v[32];
        tmem_ld16(t_accX + (uint32_t)c0 + lane_off, v);
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4* sp = reinterpret_cast<uint4*>(dy_tile + sw_off(trow, c0 + i));
          float d8[8];
          unpack8(*sp, d8);
          *sp = make_uint4(pack2(v[i] + d8[0], v[i + 1] + d8[1]), pack2(v[i + 2] + d8[2], v[i + 3] + d8[3]),
                           pack2(v[i + 4] + d8[4], v[i + 5] + d8[5]), pack2(v[i + 6] + d8[6], v[i + 7] + d8[7]));
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(ady_full));
      // ---- dF chunks ----
      for (int j = 0; j < p.NC; ++j) {
        const int b = j & 1;
        mbar_wait_relaxed(smem_u32(acc_full + b), ph_acc[b]);
        ph_acc[b] ^= 1u;
        tc_fence_after();
        uint8_t* db = df_buf + (size_t)b * p.HKB * kBlkBytes;
        for (int c0 = par * 16; c0 < p.HB; c0 += 16 * NWQ) {
          float f[32], hgrad[32];
          tmem_ld16(t_accF + (uint32_t)(b * p.HB + c0) + lane_off, f);
          tmem_ld16(t_accH + (uint32_t)(b * p.HB + c0) + lane_off, hgrad);
          const float* bp = p.b1 + j * p.HB + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bp + i));
            f[i] = hgrad[i] * gelu_grad_poly(f[i] + bb.x);
            f[i + 1] = hgrad[i + 1] * gelu_grad_poly(f[i + 1] + bb.y);
            f[i + 2] = hgrad[i + 2] * gelu_grad_poly(f[i + 2] + bb.z);
            f[i + 3] = hgrad[i + 3] * gelu_grad_poly(f[i + 3] + bb.w);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            const uint4 o = make_uint4(pack2(f[i], f[i + 1]), pack2(f[i + 2], f[i + 3]), pack2(f[i + 4], f[i + 5]),
                                       pack2(f[i + 6], f[i + 7]));
            *reinterpret_cast<uint4*>(db + sw_off(trow, c0 + i)) = o;
          }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(df_full + b));
      }
      // ---- dxn = accX -> bf16 -> XN tile region -> TMA store ----
      mbar_wait_relaxed(smem_u32(dxn_full), ph_dxn);
      ph_dxn ^= 1u;
      tc_fence_after();
      for (int c0 = par * 16; c0 < p.C; c0 += 16 * NWQ) {
        float v[32];
        tmem_ld16(t_accX + (uint32_t)c0 + lane_off, v);
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          const uint4 o = make_uint4(pack2(v[i], v[i + 1]), pack2(v[i + 2], v[i + 3]), pack2(v[i + 4], v[i + 5]),
                                     pack2(v[i + 6], v[i + 7]));
          *reinterpret_cast<uint4*>(xn_tile + sw_off(trow, c0 + i)) = o;
        }
      }
      tc_fence_before();
      fence_async_smem();
      worker_bar<NW>();
      if (wid == 0 && lane == 0) {
        for (int kb = 0; kb < p.KB; ++kb) tma_store_3d(&tmDxn, smem_u32(xn_tile + (size_t)kb * kBlkBytes), kb * BK, m_blk * BM, g);
        tma_store_commit();
        tma_store_wait_read();
      }
      worker_bar<NW>();   // the XN region may be rewritten
    }
    if (wid == 0 && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

static uint32_t pow2_cols(int need) {
  uint32_t c = 32;
  while (c < (uint32_t)need) c <<= 1;
  return c;
}

}  // namespace mlp
}  // namespace stgcma

using namespace stgcma;

extern "C" int stgcma_mlp_fwd(const float* x, const float* gamma, const float* beta, const void* W1, const float* b1,
                              const void* W2, const float* b2, const void* Wa, const float* ba, void* y, void* hpre,
                              void* h, float* mean, float* rstd, int64_t M, int C, int H, int rp, int groups, int dtype,
                              void* stream) {
  using namespace stgcma::mlp;
  STG_REQUIRE(dtype == STGCMA_BF16, "mlp_fwd: only the bf16 production mode is fused (the fp32 check mode composes "
                                    "layernorm_fwd + gemm)");
  STG_REQUIRE(x && gamma && beta && W1 && b1 && W2 && b2 && Wa && ba && y && hpre && h && mean && rstd, "mlp_fwd: null operand");
  STG_REQUIRE(M > 0 && groups >= 1 && M % groups == 0, "mlp_fwd: M=%lld not divisible by groups=%d", (long long)M, groups);
  STG_REQUIRE(C % 32 == 0 && C >= 32 && C <= 384, "mlp_fwd: C=%d must be a multiple of 32 in [32, 384]", C);
  STG_REQUIRE(H % 64 == 0 && H >= 64, "mlp_fwd: H=%d must be a multiple of 64", H);
  STG_REQUIRE(rp % 16 == 0 && rp >= 16 && rp <= 128, "mlp_fwd: rp=%d must be a multiple of 16 in [16, 128]", rp);
  cudaStream_t st = (cudaStream_t)stream;
  FwdParams p;
  p.x = x; p.gamma = gamma; p.beta = beta; p.b1 = b1; p.b2 = b2; p.ba = ba;
  p.hpre = (bf16*)hpre; p.h = (bf16*)h; p.mean = mean; p.rstd = rstd; p.trace = g_trace;
  p.group_rows = M / groups; p.groups = groups; p.C = C; p.H = H; p.rp = rp;
  p.KB = (int)ceil_div(C, BK);
  // hidden chunk: 128 columns when TMEM (2 HB + C <= 512) and H allow it and the CTA is alone on its SM, else 64
  // Two CTAs per SM when they fit (C <= 128): a tile is a chain of dependent phases (prologue -> per-chunk MMA /
  // activation hand-offs -> output) paced by the single MMA-issuing thread, so a second resident CTA fills the
  // bubbles: measured 592 us vs 935 us (one CTA, 16 workers, 20-slot ring) at Swin-T stage 0, 16 clips
  // (profiles/r01c).  STGCMA_MLP_OCC=1 forces one CTA per SM (debug).
  static const int occ_env = [] { const char* e = getenv("STGCMA_MLP_OCC"); return e ? atoi(e) : 2; }();
  const int occ = (occ_env == 2 && C <= 128) ? 2 : 1;
  // G2's B operand (N = C rows of W2) is issued in nsplit pieces: UMMA N <= 256; for small C two pieces keep the
  // ring slots at 8 KB (more slots in flight for the same shared memory)
  p.nsplit = (C > 256 || (C <= 128 && C % 32 == 0 && C > 64)) ? 2 : 1;
  p.CN = C / p.nsplit;
  STG_REQUIRE(p.CN % 16 == 0, "mlp_fwd: C / nsplit = %d must be a multiple of 16", p.CN);
  p.HB = (C > 128 && H % 128 == 0 && 2 * 128 + C <= 512) ? 128 : 64;
  p.NC = H / p.HB;
  p.HKB = p.HB / BK;
  p.m_tiles_g = (int)ceil_div(p.group_rows, BM);
  const int64_t tiles = (int64_t)groups * p.m_tiles_g;
  STG_REQUIRE(tiles < (1ll << 31), "mlp_fwd: too many tiles");
  p.num_tiles = (int)tiles;
  int slot_rows = p.HB > p.CN ? p.HB : p.CN;
  if (rp > slot_rows) slot_rows = rp;
  p.slot_bytes = (uint32_t)slot_rows * 128u;
  p.tmem_cols = pow2_cols(2 * p.HB + C);
  STG_REQUIRE(p.tmem_cols <= (occ == 2 ? 256u : 512u), "mlp_fwd: TMEM budget exceeded (C=%d HB=%d)", C, p.HB);
  const int smem_cap = occ == 1 ? 232448 : 115712;
  const int fixed_bytes = 1024 + (p.KB + 2 * p.HKB) * kBlkBytes + 512;
  int ring = (smem_cap - fixed_bytes) / (int)p.slot_bytes;
  if (ring > 24) ring = 24;
  STG_REQUIRE(ring >= 2, "mlp_fwd: shared memory budget exceeded (C=%d HB=%d)", C, p.HB);
  p.ring = ring;
  const size_t smem = (size_t)fixed_bytes + (size_t)ring * p.slot_bytes;

  CUtensorMap tmW1, tmW2, tmWa, tmY;
  int rc = make_map(&tmW1, W1, H, C, C, 1, p.HB);
  if (rc) return rc;
  rc = make_map(&tmW2, W2, C, H, H, 1, p.CN);
  if (rc) return rc;
  rc = make_map(&tmWa, Wa, rp, C, C, groups, rp);
  if (rc) return rc;
  rc = make_map(&tmY, y, p.group_rows, C, C, groups, BM);
  if (rc) return rc;
  const int grid = p.num_tiles < occ * sm_count() ? p.num_tiles : occ * sm_count();
#define STG_MLP_FWD(OCC_, NV_, NWQ_)                                                                                        \
  do {                                                                                                             \
    auto kern = mlp_fwd_kernel<OCC_, NV_, NWQ_>;                                                                           \
    static std::once_flag once;                                                                                    \
    static cudaError_t attr_err = cudaSuccess;                                                                     \
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, OCC_ == 1 ? 232448 : 115712); }); \
    if (attr_err != cudaSuccess) {                                                                                 \
      set_error("mlp_fwd: cannot raise dynamic smem limit: %s", cudaGetErrorString(attr_err));                     \
      return STGCMA_ECUDA;                                                                                         \
    }                                                                                                              \
    kern<<<grid, 64 + 128 * NWQ_, smem, st>>>(tmW1, tmW2, tmWa, tmY, p);                                                  \
    return check_launch("mlp_fwd");                                                                                \
  } while (0)
  if (occ == 2) STG_MLP_FWD(2, 1, 2);
  if (C <= 128) STG_MLP_FWD(1, 1, 4);
  if (C <= 256) STG_MLP_FWD(1, 2, 4);
  STG_MLP_FWD(1, 3, 4);
#undef STG_MLP_FWD
}

extern "C" int stgcma_mlp_bwd(const void* dhpre, const void* dxb, const float* x, const float* gamma, const float* beta,
                              const float* mean, const float* rstd, const void* Wat, const void* W1, const float* b1,
                              const void* W2t, const void* W1t, void* dxn, int64_t M, int C, int H, int rp, int groups,
                              int dtype, void* stream) {
  using namespace stgcma::mlp;
  STG_REQUIRE(dtype == STGCMA_BF16, "mlp_bwd: only the bf16 production mode is fused");
  STG_REQUIRE(dhpre && dxb && x && gamma && beta && mean && rstd && Wat && W1 && b1 && W2t && W1t && dxn, "mlp_bwd: null operand");
  STG_REQUIRE(M > 0 && groups >= 1 && M % groups == 0, "mlp_bwd: M=%lld not divisible by groups=%d", (long long)M, groups);
  STG_REQUIRE(C % 32 == 0 && C >= 32 && C <= 256, "mlp_bwd: C=%d must be a multiple of 32 in [32, 256]", C);
  STG_REQUIRE(H % 64 == 0 && H >= 64, "mlp_bwd: H=%d must be a multiple of 64", H);
  STG_REQUIRE(rp % 16 == 0 && rp >= 16 && rp <= 128, "mlp_bwd: rp=%d must be a multiple of 16 in [16, 128]", rp);
  cudaStream_t st = (cudaStream_t)stream;
  BwdParams p;
  p.x = x; p.gamma = gamma; p.beta = beta; p.mean = mean; p.rstd = rstd; p.b1 = b1;
  p.group_rows = M / groups; p.groups = groups; p.C = C; p.H = H; p.rp = rp;
  p.KB = (int)ceil_div(C, BK);
  p.RKB = (int)ceil_div(rp, BK);
  p.nsplit = (C <= 128 && C % 32 == 0 && C > 64) ? 2 : 1;
  p.CN = C / p.nsplit;
  p.HB = 64;
  p.NC = H / p.HB;
  p.HKB = 1;
  p.m_tiles_g = (int)ceil_div(p.group_rows, BM);
  const int64_t tiles = (int64_t)groups * p.m_tiles_g;
  STG_REQUIRE(tiles < (1ll << 31), "mlp_bwd: too many tiles");
  p.num_tiles = (int)tiles;
  const int slot_rows = p.HB > p.CN ? p.HB : p.CN;
  p.slot_bytes = (uint32_t)slot_rows * 128u;
  p.tmem_cols = pow2_cols(4 * p.HB + C);
  STG_REQUIRE(p.tmem_cols <= 512u, "mlp_bwd: TMEM budget exceeded (C=%d)", C);
  const int df_blocks = 2 * p.HKB > p.RKB ? 2 * p.HKB : p.RKB;
  const int fixed_bytes = 1024 + (2 * p.KB + df_blocks) * kBlkBytes + 512;
  int ring = (232448 - fixed_bytes) / (int)p.slot_bytes;
  if (ring > 24) ring = 24;
  STG_REQUIRE(ring >= 2, "mlp_bwd: shared memory budget exceeded (C=%d)", C);
  p.ring = ring;
  const size_t smem = (size_t)fixed_bytes + (size_t)ring * p.slot_bytes;

  CUtensorMap tmDh, tmDxb, tmWat, tmW1, tmW2t, tmW1t, tmDxn;
  int rc = make_map(&tmDh, dhpre, p.group_rows, rp, rp, groups, BM);
  if (rc) return rc;
  rc = make_map(&tmDxb, dxb, p.group_rows, C, C, groups, BM);
  if (rc) return rc;
  rc = make_map(&tmWat, Wat, C, rp, rp, groups, p.CN);     // [groups, C, rp]: dy = dhpre . Wat^T
  if (rc) return rc;
  rc = make_map(&tmW1, W1, H, C, C, 1, p.HB);              // [H, C]
  if (rc) return rc;
  rc = make_map(&tmW2t, W2t, H, C, C, 1, p.HB);            // [H, C] = fc2.weight^T
  if (rc) return rc;
  rc = make_map(&tmW1t, W1t, C, H, H, 1, p.CN);            // [C, H] = fc1.weight^T
  if (rc) return rc;
  rc = make_map(&tmDxn, dxn, p.group_rows, C, C, groups, BM);
  if (rc) return rc;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
#define STG_MLP_BWD(NV_, NWQ_)                                                                                           \
  do {                                                                                                             \
    auto kern = mlp_bwd_kernel<NV_, NWQ_>;                                                                               \
    static std::once_flag once;                                                                                    \
    static cudaError_t attr_err = cudaSuccess;                                                                     \
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448); }); \
    if (attr_err != cudaSuccess) {                                                                                 \
      set_error("mlp_bwd: cannot raise dynamic smem limit: %s", cudaGetErrorString(attr_err));                     \
      return STGCMA_ECUDA;                                                                                         \
    }                                                                                                              \
    kern<<<grid, 64 + 128 * NWQ_, smem, st>>>(tmDh, tmDxb, tmWat, tmW1, tmW2t, tmW1t, tmDxn, p);                          \
    return check_launch("mlp_bwd");                                                                                \
  } while (0)
  if (C <= 128) STG_MLP_BWD(1, 4);
  STG_MLP_BWD(2, 4);
#undef STG_MLP_BWD
}

/* debug only: CTA 0 of the next stgcma_mlp_fwd launches records clock64() phase marks into buf (3 x 512 u64, device) */
extern "C" int stgcma_debug_trace(void* buf) {
  stgcma::mlp::g_trace = (unsigned long long*)buf;
  return STGCMA_OK;
}
