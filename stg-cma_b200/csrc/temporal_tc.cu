// Temporal attention (WindowAttention.forward, temporal=True; Swin_AVE.py:244-255), bf16 production kernels.
//
//   sequence = the T frames of one (modality g, clip b, token n); per head (dim 32):
//       S[i][j] = (q_i . k_j) / sqrt(32) + tbl_g[i - j + T - 1][head]      P = softmax_j S      o_i = sum_j P[i][j] v_j
//
// The '(b t) n c -> (b n) t c' rearranges (:706-716) are folded into the addressing: a CTA owns NT consecutive
// tokens of one (g, b) and loads, for each frame t, the NT * 3C contiguous channels of those tokens with 16-byte
// cp.async -- every global access of the kernel is a full-line burst (the previous kernel read 64-byte pieces).
// One warp processes one (token, head) pair at a time on the warp-level tensor cores (mma.sync m16n8k16; T <= 16 is
// one 16 x 16 score tile): Q K^T, softmax in registers, P V.  Results overwrite the task's own q (fwd) / q, k, v
// (bwd) slots of the shared tile, which is then written back with coalesced 16-byte stores.
//   backward: dS = P o (dP - rowsum(P o dP)),  dQ = dS K / sqrt(32),  dK = dS^T Q / sqrt(32),  dV = P^T dO, and the
//   gradient of the TRAINABLE bias table, dtbl[i - j + T - 1][head] += dS[i][j].  The transposed products use a
//   second, transposed score tile (K Q^T) instead of transposing fragments.  dtbl is accumulated in shared memory
//   over the CTA's whole (persistent) tile loop and flushed once: a few thousand global atomics per launch.
#include "mma.cuh"

namespace stgcma {
namespace ttc {

using namespace mma;

// head dim: 32 for every Swin configuration; the CLIP-ViT temporal attention (nn.MultiheadAttention over T frames,
// CLIP_AVE.py:369-377: no bias) runs the same kernels with HD = 96 (ViT-B/16 with the reference's 8 heads), 64, 16, 128
// and BIAS = false through stgcma_mha_fwd / _bwd (mha.cu dispatches here when seq_len <= 16 and inner > 1).
template <int HD> struct Scale;
template <> struct Scale<16> { static constexpr float v = 0.25f; };
template <> struct Scale<32> { static constexpr float v = 0.17677669529663689f; };
template <> struct Scale<64> { static constexpr float v = 0.125f; };
template <> struct Scale<96> { static constexpr float v = 0.10206207261596575f; };
template <> struct Scale<128> { static constexpr float v = 0.08838834764831845f; };

struct Geo {
  int T, L, heads, C, NT;
  int row3;      // bytes of one token row of qkv   (3C * 2)
  int row1;      // bytes of one token row of out   (C * 2)
  int ts3, ts1;  // byte stride between frames in the shared tiles (NT * row + pad, == 16 mod 128)
};

__device__ __forceinline__ int clampi(int r, int T) { return r < T ? r : T - 1; }

// fragments from a [t][.] tile whose frame stride is `ts` bytes; rows (frames) >= T are clamped to T - 1
__device__ __forceinline__ void lda(uint32_t (&a)[4], const uint8_t* base, int ts, int k0, int T, int lane) {
  const int r = clampi((lane & 7) + ((lane >> 3) & 1) * 8, T);
  ldsm_x4(a, smem_addr(base + r * ts + (k0 + (lane >> 4) * 8) * 2));
}
__device__ __forceinline__ void ldb_nk(uint32_t (&b)[4], const uint8_t* base, int ts, int k0, int T, int lane) {
  const int n = clampi((lane & 7) + (lane >> 4) * 8, T);
  ldsm_x4(b, smem_addr(base + n * ts + (k0 + ((lane >> 3) & 1) * 8) * 2));
}
__device__ __forceinline__ void ldb_kn(uint32_t (&b)[4], const uint8_t* base, int ts, int n0, int T, int lane) {
  const int k = clampi((lane & 7) + ((lane >> 3) & 1) * 8, T);
  ldsm_x4_t(b, smem_addr(base + k * ts + (n0 + (lane >> 4) * 8) * 2));
}

// D[16 x 16] = A[16 x HD] . B[16 x HD]^T   (A rows from `a`, B rows from `b`, both [t][HD] slices)
template <int HD>
__device__ __forceinline__ void mm_nt(float (&d)[2][4], const uint8_t* a, int tsa, const uint8_t* b, int tsb, int T, int lane) {
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    uint32_t fa[4], fb[4];
    lda(fa, a, tsa, ks * 16, T, lane);
    ldb_nk(fb, b, tsb, ks * 16, T, lane);
    mma16816(d[0], fa, fb[0], fb[1]);
    mma16816(d[1], fa, fb[2], fb[3]);
  }
}
// packs a 16 x 16 tile in accumulator layout into the A fragment of the next product
__device__ __forceinline__ void pack_a(uint32_t (&pa)[4], const float (&p)[2][4]) {
  pa[0] = pack_bf16(p[0][0], p[0][1]);
  pa[1] = pack_bf16(p[0][2], p[0][3]);
  pa[2] = pack_bf16(p[1][0], p[1][1]);
  pa[3] = pack_bf16(p[1][2], p[1][3]);
}
// D[16 x 32] = P[16 x 16] . B[16 x (n0 .. n0 + 32)]   (B rows = contraction index): one 32-column block of the output
__device__ __forceinline__ void mm_pv32(float (&d)[4][4], const uint32_t (&pa)[4], const uint8_t* b, int tsb, int n0, int T, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
#pragma unroll
  for (int n2 = 0; n2 < 2; ++n2) {
    uint32_t fb[4];
    ldb_kn(fb, b, tsb, n0 + n2 * 16, T, lane);
    mma16816(d[2 * n2], pa, fb[0], fb[1]);
    mma16816(d[2 * n2 + 1], pa, fb[2], fb[3]);
  }
}
// 16-column block (HD = 16, or the tail of HD = 16 mod 32)
__device__ __forceinline__ void mm_pv16(float (&d)[4][4], const uint32_t (&pa)[4], const uint8_t* b, int tsb, int n0, int T, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  uint32_t fb[4];
  ldb_kn(fb, b, tsb, n0, T, lane);
  mma16816(d[0], pa, fb[0], fb[1]);
  mma16816(d[1], pa, fb[2], fb[3]);
}

// writes NB 8-column accumulator blocks (rows = frames) into columns [n0, n0 + 8 NB) of a [t][HD] slice, rows < T only
template <int NB>
__device__ __forceinline__ void store_cols(uint8_t* dst, int ts, const float (&d)[4][4], float scale, int n0, int T, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int col = n0 + i * 8 + 2 * t;
    if (g < T) *reinterpret_cast<uint32_t*>(dst + g * ts + col * 2) = pack_bf16(d[i][0] * scale, d[i][1] * scale);
    if (g + 8 < T) *reinterpret_cast<uint32_t*>(dst + (g + 8) * ts + col * 2) = pack_bf16(d[i][2] * scale, d[i][3] * scale);
  }
}

// copies NT token rows x T frames between global ([T frames L rows apart][row bytes]) and the shared tile
template <bool TO_SMEM>
__device__ __forceinline__ void copy_tile(uint8_t* tile, int ts, uint8_t* gbase, int64_t frame_stride_bytes, int rowbytes,
                                          int nt, int T, int tid, int nthr) {
  const int cpr = nt * rowbytes / 16;   // 16-byte chunks per frame (the nt rows are contiguous)
  // (frame, chunk) of element e = tid + k * nthr advanced incrementally: one division per call instead of one per chunk
  // (the per-chunk division was ~30 of the ~45 instructions of this loop's body)
  const int dt = nthr / cpr, dc = nthr - dt * cpr;
  int t = tid / cpr, c = tid - t * cpr;
  while (t < T) {
    uint8_t* s = tile + t * ts + c * 16;
    uint8_t* gp = gbase + t * frame_stride_bytes + (int64_t)c * 16;
    if (TO_SMEM) cp_async16(s, gp);
    else *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(s);
    t += dt;
    c += dc;
    if (c >= cpr) { c -= cpr; ++t; }
  }
}

// softmax of a 16 x 16 score tile held in accumulator layout; returns P in place and the row lse (log2 domain)
template <int HD, bool BIAS>
__device__ __forceinline__ void bias_softmax(float (&s)[2][4], const float* tbs, int heads, int T, int lane, float& lse0,
                                             float& lse1) {
  constexpr float kScale = Scale<HD>::v;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = g + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
      s[nt][e] = (j < T && i < T) ? (s[nt][e] * kScale + (BIAS ? tbs[(i - j + T - 1) * heads] : 0.f)) * kLog2e : -INFINITY;
    }
  float m0 = quad_max(fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1])));
  float m1 = quad_max(fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3])));
  if (m0 == -INFINITY) m0 = 0.f;   // padding rows (i >= T): all -inf
  if (m1 == -INFINITY) m1 = 0.f;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    s[nt][0] = ex2(s[nt][0] - m0);
    s[nt][1] = ex2(s[nt][1] - m0);
    s[nt][2] = ex2(s[nt][2] - m1);
    s[nt][3] = ex2(s[nt][3] - m1);
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    s[nt][0] *= i0;
    s[nt][1] *= i0;
    s[nt][2] *= i1;
    s[nt][3] *= i1;
  }
  lse0 = l0 > 0.f ? m0 + log2f(l0) : INFINITY;
  lse1 = l1 > 0.f ? m1 + log2f(l1) : INFINITY;
}

__device__ __forceinline__ void tile_coords(int64_t tile, int tiles_per_seq, int B, int NT, int& g, int& b, int& n0) {
  const int tl = (int)(tile % tiles_per_seq);
  const int64_t gb = tile / tiles_per_seq;
  b = (int)(gb % B);
  g = (int)(gb / B);
  n0 = tl * NT;
}

// ---------------------------------------------------------------------------------------------
// o = P V written straight into this task's q slot, one 32-column block at a time (q is dead once S exists)
template <int HD>
__device__ __forceinline__ void pv_into(uint8_t* dst, int ts_dst, const uint32_t (&pa)[4], const uint8_t* b, int tsb, float scale,
                                        int T, int lane) {
#pragma unroll
  for (int n0 = 0; n0 + 32 <= HD; n0 += 32) {
    float d[4][4];
    mm_pv32(d, pa, b, tsb, n0, T, lane);
    store_cols<4>(dst, ts_dst, d, scale, n0, T, lane);
  }
  if (HD % 32) {
    float d[4][4];
    mm_pv16(d, pa, b, tsb, HD - 16, T, lane);
    store_cols<2>(dst, ts_dst, d, scale, HD - 16, T, lane);
  }
}

template <int HD, bool BIAS>
__global__ void __launch_bounds__(256) tt_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ tbl,
                                                     bf16* __restrict__ out, float* __restrict__ lse_out, int G, int B, Geo geo) {
  extern __shared__ __align__(16) uint8_t tsm[];
  uint8_t* tile = tsm;                                               // [T][NT][3C] (+pad per frame)
  float* tb_s = reinterpret_cast<float*>(tile + geo.T * geo.ts3);    // [G][2T-1][heads]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = geo.T, L = geo.L, NT = geo.NT, heads = geo.heads;
  const int nb = 2 * T - 1;
  if (BIAS)
    for (int i = tid; i < G * nb * heads; i += 256) tb_s[i] = tbl[i];
  const int tiles_per_seq = (L + NT - 1) / NT;
  const int64_t ntiles = (int64_t)G * B * tiles_per_seq;
  for (int64_t tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    int g, b, n0;
    tile_coords(tile_i, tiles_per_seq, B, NT, g, b, n0);
    const int nt = (L - n0) < NT ? (L - n0) : NT;
    const int64_t row0 = ((int64_t)g * B + b) * T * L + n0;   // token row of frame 0
    __syncthreads();   // the previous tile has been written back; tb_s is loaded
    copy_tile<true>(tile, geo.ts3, (uint8_t*)(qkv + row0 * 3 * geo.C), (int64_t)L * geo.row3, geo.row3, nt, T, tid, 256);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int task = warp; task < nt * heads; task += 8) {
      const int n = task / heads, h = task - n * heads;
      uint8_t* qb = tile + n * geo.row3 + h * (HD * 2);
      const uint8_t* kb = qb + geo.C * 2;
      const uint8_t* vb = qb + geo.C * 4;
      const float* tbs = tb_s + (g * nb) * heads + h;
      float s[2][4], lse0, lse1;
      mm_nt<HD>(s, qb, geo.ts3, kb, geo.ts3, T, lane);
      bias_softmax<HD, BIAS>(s, tbs, heads, T, lane, lse0, lse1);
      if (lse_out && (lane & 3) == 0) {      // natural-log lse per (row, head): the stgcma_mha_fwd contract
        const int gq = lane >> 2;
        if (gq < T) lse_out[(row0 + (int64_t)gq * L + n) * heads + h] = lse0 * 0.6931471805599453f;
        if (gq + 8 < T) lse_out[(row0 + (int64_t)(gq + 8) * L + n) * heads + h] = lse1 * 0.6931471805599453f;
      }
      uint32_t pa[4];
      pack_a(pa, s);
      __syncwarp();      // every lane has read q (mm_nt): o may overwrite this task's q slot
      pv_into<HD>(qb, geo.ts3, pa, vb, geo.ts3, 1.0f, T, lane);
    }
    __syncthreads();
    // write back the q slots ([t][n][0..C)) as out[t][n][C]
    {
      const int cpr = geo.row1 / 16;
      for (int e = tid; e < T * nt * cpr; e += 256) {
        const int t = e / (nt * cpr), rem = e - t * nt * cpr, n = rem / cpr, c = rem - n * cpr;
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out + (row0 + (int64_t)t * L + n) * geo.C) + c * 16) =
            *reinterpret_cast<const uint4*>(tile + t * geo.ts3 + n * geo.row3 + c * 16);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FIXED_HEAD (BIAS only): the block has a multiple of `heads` warps and warp w always works on head w % heads, so the
// bias-table gradient dS is summed element-wise in REGISTERS over the CTA's whole life (per modality) and folded into
// the diagonals once at the end.  The generic variant adds every dS element to its diagonal in shared memory as it goes:
// fp32 shared atomics are CAS loops and the ~10 elements of a diagonal collide inside the warp (ncu, profiles/r01f:
// 7 retries per atomic, ~25 % of the kernel's instructions).
// The three output products are evaluated one 32-column block at a time and written IN PLACE (dQ over q, dK over k,
// dV over v): column block c of dQ = dS K needs only column block c of K, and of dK = dS^T Q only that of Q, so blocks
// already overwritten are never read again -- no output tile is ever held whole in registers (144 at HD = 96).
template <bool FIXED_HEAD, int NTHR, int MINB, int HD, bool BIAS>
__global__ void __launch_bounds__(NTHR, MINB) tt_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ tbl,
                                                     const bf16* __restrict__ dout, bf16* __restrict__ dqkv,
                                                     float* __restrict__ dtbl, int G, int B, Geo geo) {
  constexpr float kScale = Scale<HD>::v;
  extern __shared__ __align__(16) uint8_t tsm[];
  uint8_t* tile = tsm;                                               // [T][NT][3C]
  uint8_t* dtile = tile + geo.T * geo.ts3;                           // [T][NT][C]  dout
  float* tb_s = reinterpret_cast<float*>(dtile + geo.T * geo.ts1);   // [G][2T-1][heads]
  float* db_s = tb_s + G * (2 * geo.T - 1) * geo.heads;              // [G][2T-1][heads] gradient accumulator
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int T = geo.T, L = geo.L, NT = geo.NT, heads = geo.heads;
  const int nb = 2 * T - 1;
  const int nthr = blockDim.x, nwarp = nthr >> 5;
  if (BIAS)
    for (int i = tid; i < G * nb * heads; i += nthr) {
      tb_s[i] = tbl[i];
      db_s[i] = 0.f;
    }
  float dsacc[2][2][4];   // [modality][n-tile][e]   (FIXED_HEAD)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 2; ++i) dsacc[a][i][0] = dsacc[a][i][1] = dsacc[a][i][2] = dsacc[a][i][3] = 0.f;
  const int task0 = warp;
  const int task_stride = nwarp;   // FIXED_HEAD: a multiple of heads, so task % heads never changes
  const int tiles_per_seq = (L + NT - 1) / NT;
  const int64_t ntiles = (int64_t)G * B * tiles_per_seq;
  for (int64_t tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    int g, b, n0;
    tile_coords(tile_i, tiles_per_seq, B, NT, g, b, n0);
    const int nt = (L - n0) < NT ? (L - n0) : NT;
    const int64_t row0 = ((int64_t)g * B + b) * T * L + n0;
    __syncthreads();
    copy_tile<true>(tile, geo.ts3, (uint8_t*)(qkv + row0 * 3 * geo.C), (int64_t)L * geo.row3, geo.row3, nt, T, tid, nthr);
    copy_tile<true>(dtile, geo.ts1, (uint8_t*)(dout + row0 * geo.C), (int64_t)L * geo.row1, geo.row1, nt, T, tid, nthr);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int task = task0; task < nt * heads; task += task_stride) {
      const int n = task / heads, h = task - n * heads;
      uint8_t* qb = tile + n * geo.row3 + h * (HD * 2);
      uint8_t* kb = qb + geo.C * 2;
      uint8_t* vb = qb + geo.C * 4;
      const uint8_t* ob = dtile + n * geo.row1 + h * (HD * 2);
      const float* tbs = tb_s + (g * nb) * heads + h;
      float* dbs = db_s + (g * nb) * heads + h;
      // ---- direct tile: rows i (queries), columns j (keys) ----
      float p[2][4], dp[2][4], lse0, lse1;
      mm_nt<HD>(p, qb, geo.ts3, kb, geo.ts3, T, lane);
      bias_softmax<HD, BIAS>(p, tbs, heads, T, lane, lse0, lse1);
      mm_nt<HD>(dp, ob, geo.ts1, vb, geo.ts3, T, lane);          // dP = dO V^T
      float dl0 = quad_sum(p[0][0] * dp[0][0] + p[0][1] * dp[0][1] + p[1][0] * dp[1][0] + p[1][1] * dp[1][1]);
      float dl1 = quad_sum(p[0][2] * dp[0][2] + p[0][3] * dp[0][3] + p[1][2] * dp[1][2] + p[1][3] * dp[1][3]);
      float ds[2][4];
#pragma unroll
      for (int nt2 = 0; nt2 < 2; ++nt2)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = gq + (e >> 1) * 8, j = nt2 * 8 + 2 * tq + (e & 1);
          const float v = p[nt2][e] * (dp[nt2][e] - ((e & 2) ? dl1 : dl0));   // p == 0 on padding
          ds[nt2][e] = v;
          if (BIAS) {
            if (FIXED_HEAD) {
              if (g) dsacc[1][nt2][e] += v; else dsacc[0][nt2][e] += v;       // p == 0 outside the T x T block
            } else if (i < T && j < T) {
              atomicAdd(dbs + (i - j + T - 1) * heads, v);   // the bias enters the logits un-scaled
            }
          }
        }
      uint32_t ds_a[4];
      pack_a(ds_a, ds);
      // ---- transposed tile: rows j (keys), columns i (queries) ----
      float pt[2][4], dpt[2][4];
      mm_nt<HD>(pt, kb, geo.ts3, qb, geo.ts3, T, lane);          // S^T = K Q^T
      mm_nt<HD>(dpt, vb, geo.ts3, ob, geo.ts1, T, lane);         // dP^T = V dO^T
      // lse / dl of query i live in the quad of lanes 4 (i % 8) .. + 3 (lse0 / dl0: i < 8, lse1 / dl1: i >= 8)
#pragma unroll
      for (int nt2 = 0; nt2 < 2; ++nt2)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = gq + (e >> 1) * 8, i = nt2 * 8 + 2 * tq + (e & 1);
          const int src = 4 * (2 * tq + (e & 1));
          const float lse_i = __shfl_sync(0xffffffffu, nt2 ? lse1 : lse0, src);
          const float dl_i = __shfl_sync(0xffffffffu, nt2 ? dl1 : dl0, src);
          float pv = 0.f;
          if (i < T && j < T) pv = ex2((pt[nt2][e] * kScale + (BIAS ? tbs[(i - j + T - 1) * heads] : 0.f)) * kLog2e - lse_i);
          pt[nt2][e] = pv;
          dpt[nt2][e] = pv * (dpt[nt2][e] - dl_i);               // dS^T
        }
      uint32_t pt_a[4], dst_a[4];
      pack_a(pt_a, pt);
      pack_a(dst_a, dpt);
      __syncwarp();      // every score-level product has read q, k, v, dO: the slots may now be overwritten block by block
      pv_into<HD>(vb, geo.ts3, pt_a, ob, geo.ts1, 1.0f, T, lane);      // dV = P^T dO        (reads dO only)
#pragma unroll
      for (int n0 = 0; n0 + 32 <= HD; n0 += 32) {
        float dq[4][4], dk[4][4];
        mm_pv32(dq, ds_a, kb, geo.ts3, n0, T, lane);                   // dQ[:, blk] = dS K[:, blk]
        mm_pv32(dk, dst_a, qb, geo.ts3, n0, T, lane);                  // dK[:, blk] = dS^T Q[:, blk]
        __syncwarp();
        store_cols<4>(qb, geo.ts3, dq, kScale, n0, T, lane);
        store_cols<4>(kb, geo.ts3, dk, kScale, n0, T, lane);
      }
      if (HD % 32) {
        float dq[4][4], dk[4][4];
        mm_pv16(dq, ds_a, kb, geo.ts3, HD - 16, T, lane);
        mm_pv16(dk, dst_a, qb, geo.ts3, HD - 16, T, lane);
        __syncwarp();
        store_cols<2>(qb, geo.ts3, dq, kScale, HD - 16, T, lane);
        store_cols<2>(kb, geo.ts3, dk, kScale, HD - 16, T, lane);
      }
    }
    __syncthreads();
    copy_tile<false>(tile, geo.ts3, (uint8_t*)(dqkv + row0 * 3 * geo.C), (int64_t)L * geo.row3, geo.row3, nt, T, tid, nthr);
  }
  if (BIAS) {
    if (FIXED_HEAD) {
      const int h = warp % heads;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (a >= G) break;
#pragma unroll
        for (int nt2 = 0; nt2 < 2; ++nt2)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = gq + (e >> 1) * 8, j = nt2 * 8 + 2 * tq + (e & 1);
            if (i < T && j < T) atomicAdd(db_s + (a * nb + (i - j + T - 1)) * heads + h, dsacc[a][nt2][e]);
          }
      }
    }
    __syncthreads();
    for (int i = tid; i < G * nb * heads; i += nthr) {
      const float v = db_s[i];
      if (v != 0.f) atomicAdd(dtbl + i, v);
    }
  }
}

static bool make_geo(Geo& geo, int T, int L, int heads, int HD) {
  geo.T = T;
  geo.L = L;
  geo.heads = heads;
  geo.C = heads * HD;
  geo.row3 = 3 * geo.C * 2;
  geo.row1 = geo.C * 2;
  // tokens per tile: keep the qkv tile around 48 KB so that several CTAs share an SM
  int nt = (48 * 1024) / (T * geo.row3);
  if (nt < 1) nt = 1;
  if (nt > 8) nt = 8;
  if (nt > L) nt = L;
  geo.NT = nt;
  auto pad16 = [](int bytes) { return bytes + ((16 - (bytes % 128)) + 128) % 128; };   // stride == 16 (mod 128)
  geo.ts3 = pad16(nt * geo.row3);
  geo.ts1 = pad16(nt * geo.row1);
  return true;
}

template <int HD, bool BIAS>
static int fwd_launch(const void* qkv, const float* tbl, void* out, float* lse, int G, int B, int T, int L, int heads, cudaStream_t st) {
  Geo geo;
  make_geo(geo, T, L, heads, HD);
  const size_t smem = (size_t)T * geo.ts3 + (BIAS ? (size_t)G * (2 * T - 1) * heads * sizeof(float) : 0);
  if (smem > 200 * 1024) return -1;
  auto kern = tt_fwd_kernel<HD, BIAS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("temporal_attn_fwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  const int64_t ntiles = (int64_t)G * B * ceil_div(L, geo.NT);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  kern<<<grid, 256, smem, st>>>((const bf16*)qkv, tbl, (bf16*)out, lse, G, B, geo);
  return check_launch("temporal_attn_fwd(tc)");
}

template <int HD>
static int bwd_launch_nobias(const void* qkv, const void* dout, void* dqkv, int B, int T, int L, int heads, cudaStream_t st) {
  Geo geo;
  make_geo(geo, T, L, heads, HD);
  const size_t smem = (size_t)T * (geo.ts3 + geo.ts1);
  if (smem > 200 * 1024) return -1;
  auto kern = tt_bwd_kernel<false, 256, 2, HD, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("mha_bwd(temporal tile): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  const int64_t ntiles = (int64_t)B * ceil_div(L, geo.NT);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  kern<<<grid, 256, smem, st>>>((const bf16*)qkv, nullptr, (const bf16*)dout, (bf16*)dqkv, nullptr, 1, B, geo);
  return check_launch("mha_bwd(temporal tile)");
}

}  // namespace ttc

int temporal_attn_fwd_bf16(const void* qkv, const float* tbl, void* out, int G, int B, int T, int L, int heads,
                           cudaStream_t st) {
  if (T > 16 || T < 1) return -1;
  return ttc::fwd_launch<32, true>(qkv, tbl, out, nullptr, G, B, T, L, heads, st);
}

int temporal_attn_bwd_bf16(const void* qkv, const float* tbl, const void* dout, void* dqkv, float* dtbl, int G, int B, int T,
                           int L, int heads, cudaStream_t st) {
  using namespace ttc;
  constexpr int HD = 32;
  if (T > 16 || T < 1) return -1;
  Geo geo;
  make_geo(geo, T, L, heads, HD);
  const size_t smem = (size_t)T * (geo.ts3 + geo.ts1) + (size_t)2 * G * (2 * T - 1) * heads * sizeof(float);
  if (smem > 200 * 1024) return -1;
  // warps per block: a multiple of `heads` near 8 when there is one (-> tt_bwd_kernel<true>), else 8
  int nw = 8;
  auto kern = tt_bwd_kernel<false, 256, 2, HD, true>;
  if (G <= 2 && heads <= 16) {
    nw = heads <= 8 ? (8 / heads) * heads : heads;
    if (nw == 6) kern = tt_bwd_kernel<true, 192, 3, HD, true>;
    else if (nw == 8) kern = tt_bwd_kernel<true, 256, 2, HD, true>;
    else if (nw <= 12) kern = tt_bwd_kernel<true, 384, 1, HD, true>;
    else kern = tt_bwd_kernel<true, 512, 1, HD, true>;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("temporal_attn_bwd(tc): smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  const int64_t ntiles = (int64_t)G * B * ceil_div(L, geo.NT);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  kern<<<grid, nw * 32, smem, st>>>((const bf16*)qkv, tbl, (const bf16*)dout, (bf16*)dqkv, dtbl, G, B, geo);
  return check_launch("temporal_attn_bwd(tc)");
}

// CLIP-ViT temporal attention (no bias; mha.cu routes here when the sequences are the T <= 16 frames of a
// [clips, T, tokens] row layout): -1 = shape not covered, the caller falls back to the generic flash kernels
int temporal_mha_fwd_bf16(const void* qkv, void* out, float* lse, int B, int T, int L, int heads, int hd, cudaStream_t st) {
  if (T > 16 || T < 1) return -1;
  switch (hd) {
    case 16: return ttc::fwd_launch<16, false>(qkv, nullptr, out, lse, 1, B, T, L, heads, st);
    case 32: return ttc::fwd_launch<32, false>(qkv, nullptr, out, lse, 1, B, T, L, heads, st);
    case 64: return ttc::fwd_launch<64, false>(qkv, nullptr, out, lse, 1, B, T, L, heads, st);
    case 96: return ttc::fwd_launch<96, false>(qkv, nullptr, out, lse, 1, B, T, L, heads, st);
    case 128: return ttc::fwd_launch<128, false>(qkv, nullptr, out, lse, 1, B, T, L, heads, st);
    default: return -1;
  }
}

int temporal_mha_bwd_bf16(const void* qkv, const void* dout, void* dqkv, int B, int T, int L, int heads, int hd, cudaStream_t st) {
  if (T > 16 || T < 1) return -1;
  switch (hd) {
    case 16: return ttc::bwd_launch_nobias<16>(qkv, dout, dqkv, B, T, L, heads, st);
    case 32: return ttc::bwd_launch_nobias<32>(qkv, dout, dqkv, B, T, L, heads, st);
    case 64: return ttc::bwd_launch_nobias<64>(qkv, dout, dqkv, B, T, L, heads, st);
    case 96: return ttc::bwd_launch_nobias<96>(qkv, dout, dqkv, B, T, L, heads, st);
    case 128: return ttc::bwd_launch_nobias<128>(qkv, dout, dqkv, B, T, L, heads, st);
    default: return -1;
  }
}

}  // namespace stgcma
