// dW += A^T . B for the trainable adapter weights (bf16 production path, warp-level tensor cores).
//   A [M, Ka] (row stride lda), B [M, Nb] (row stride ldb), contraction over the M token rows; dW fp32 [Ka, Nb].
// The shapes are extreme: M = 10^5..10^6 rows against 12..768 output rows / columns, i.e. a streaming
// reduction that must run at HBM speed (bytes = 2 M (Ka + Nb)); the CUDA-core version reached 350 GB/s.
// A CTA owns a 128 x 64 or 64 x 128 block of dW (whichever operand is the wide one) and a contiguous slice of M; the
// grid is one wave of 2 CTAs per SM; 64-row chunks of A and B stream through a 4-stage cp.async pipeline; A^T fragments come from ldmatrix.trans on the [m][i] tile, B fragments from
// ldmatrix.trans on the [m][j] tile (the contraction index is the row index of both).  Partial sums are added
// to dW with fp32 atomics (the destination is a slice of the pre-zeroed gradient bucket).  The bias gradient
// (column sums of A) is one extra MMA per step against an all-ones B fragment held in registers.
#include "mma.cuh"

namespace stgcma {
namespace tntc {

using namespace mma;

constexpr int MC = 64;    // token rows per chunk
constexpr int NST = 4;    // pipeline stages (3 chunks = ~40 KB per CTA in flight, 2 CTAs per SM)
constexpr int NTH = 256;  // 8 warps, each owns a 16 x 64 block of the CTA's IB x JB tile of dW

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// IB x JB = 128 x 64 (A wide, B narrow: dW2 = dOut^T . h) or 64 x 128 (A narrow, B wide: dW1 = dHpre^T . x).
// RS: rows of A are scaled by rowscale[m] (DropPath keep mask of the T adapter) before the product and the column sum.
template <int IB, int JB, bool RS>
__global__ void __launch_bounds__(NTH, 2)
gemm_tn_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, const float* __restrict__ rowscale,
               float* __restrict__ dW, float* __restrict__ csumA, int64_t M, int Ka, int Nb, int64_t lda, int64_t ldb,
               int64_t ldo, int64_t rows_per_cta) {
  constexpr int LDA_S = IB * 2 + 16, LDB_S = JB * 2 + 16;   // +16 B: ldmatrix rows land in distinct banks
  constexpr int STAGE = MC * (LDA_S + LDB_S) + MC * 4;
  constexpr int WI = IB / 16;
  static_assert(WI * (JB / 64) == NTH / 32, "8 warps of 16 x 64");
  extern __shared__ __align__(16) uint8_t tsm[];
  const int i_blk = blockIdx.x * IB, j_blk = blockIdx.y * JB;
  const int64_t m_beg = (int64_t)blockIdx.z * rows_per_cta;
  int64_t m_end = m_beg + rows_per_cta;
  if (m_end > M) m_end = M;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wi = warp % WI, wj = warp / WI;
  const int ka8 = (Ka + 7) & ~7, nb8 = (Nb + 7) & ~7;
  const int a_chunks = (min(IB, ka8 - i_blk) + 7) >> 3;    // valid 16-byte chunks per A row
  const int b_chunks = (min(JB, nb8 - j_blk) + 7) >> 3;
  const int nchunk = (int)((m_end - m_beg + MC - 1) / MC);

  // columns past Ka / Nb are never loaded: they stay zero from here on
  for (int e = tid; e < NST * STAGE / 16; e += NTH) reinterpret_cast<uint4*>(tsm)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  auto load_chunk = [&](int c) {
    uint8_t* da = tsm + (c % NST) * STAGE;
    uint8_t* db = da + MC * LDA_S;
    const int64_t m0 = m_beg + (int64_t)c * MC;
    const bool full = m0 + MC <= m_end;
    for (int e = tid; e < MC * (IB / 8); e += NTH) {
      const int r = e / (IB / 8), ch = e % (IB / 8);
      if (ch >= a_chunks) continue;
      uint8_t* dst = da + r * LDA_S + ch * 16;
      if (full || m0 + r < m_end) cp_async16(dst, A + (m0 + r) * lda + i_blk + ch * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    for (int e = tid; e < MC * (JB / 8); e += NTH) {
      const int r = e / (JB / 8), ch = e % (JB / 8);
      if (ch >= b_chunks) continue;
      uint8_t* dst = db + r * LDB_S + ch * 16;
      if (full || m0 + r < m_end) cp_async16(dst, B + (m0 + r) * ldb + j_blk + ch * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    if (RS && tid < MC && m0 + tid < m_end) cp_async4(db + MC * LDB_S + tid * 4, rowscale + m0 + tid);
  };

  float acc[8][4], accs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const bool want_sum = csumA != nullptr && blockIdx.y == 0 && wj == 0;
  const uint32_t ones = 0x3f803f80u;   // bf16 (1.0, 1.0)
  const bool i_active = i_blk + wi * 16 < Ka;          // this warp's 16 rows hold real output
  const int n_act = min(8, (Nb - j_blk - wj * 64 + 7) >> 3);   // n8 tiles of this warp that hold real output

#pragma unroll
  for (int c = 0; c < NST - 1; ++c) {
    if (c < nchunk) load_chunk(c);
    cp_async_commit();
  }
  for (int c = 0; c < nchunk; ++c) {
    cp_async_wait<NST - 2>();
    __syncthreads();   // chunk c has landed for every thread; every warp is done with chunk c - 1
    if (c + NST - 1 < nchunk) load_chunk(c + NST - 1);   // refills the stage chunk c - 1 used
    cp_async_commit();
    const uint8_t* ca = tsm + (c % NST) * STAGE;
    const uint8_t* cb = ca + MC * LDA_S;
    const float* rs = reinterpret_cast<const float*>(cb + MC * LDB_S);
    if (i_active && n_act > 0) {
#pragma unroll
      for (int ks = 0; ks < MC / 16; ++ks) {
        // A^T fragment (rows i, k = token rows) by ldmatrix.trans on the [m][i] tile
        uint32_t a[4];
        {
          const int id = lane >> 3;
          const int m = ks * 16 + (lane & 7) + ((id >> 1) & 1) * 8;
          const int i = wi * 16 + (id & 1) * 8;
          ldsm_x4_t(a, smem_addr(ca + m * LDA_S + i * 2));
        }
        if (RS) {
          const float2 s01 = *reinterpret_cast<const float2*>(rs + ks * 16 + 2 * t);
          const float2 s23 = *reinterpret_cast<const float2*>(rs + ks * 16 + 8 + 2 * t);
          const uint32_t p01 = pack_bf16(s01.x, s01.y), p23 = pack_bf16(s23.x, s23.y);
          a[0] = mul_bf16x2(a[0], p01); a[1] = mul_bf16x2(a[1], p01);
          a[2] = mul_bf16x2(a[2], p23); a[3] = mul_bf16x2(a[3], p23);
        }
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
          if (2 * n2 < n_act) {
            uint32_t b[4];
            load_b_kn(b, cb, LDB_S, ks * 16, wj * 64 + n2 * 16, lane);
            mma16816(acc[2 * n2], a, b[0], b[1]);
            mma16816(acc[2 * n2 + 1], a, b[2], b[3]);
          }
        }
        if (want_sum) mma16816(accs, a, ones, ones);
      }
    }
  }

  if (i_active) {
    const int ia = i_blk + wi * 16 + g, ib = ia + 8;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int j = j_blk + wj * 64 + n * 8 + 2 * t;
      if (ia < Ka) {
        if (j < Nb) atomicAdd(dW + (int64_t)ia * ldo + j, acc[n][0]);
        if (j + 1 < Nb) atomicAdd(dW + (int64_t)ia * ldo + j + 1, acc[n][1]);
      }
      if (ib < Ka) {
        if (j < Nb) atomicAdd(dW + (int64_t)ib * ldo + j, acc[n][2]);
        if (j + 1 < Nb) atomicAdd(dW + (int64_t)ib * ldo + j + 1, acc[n][3]);
      }
    }
    if (want_sum && t == 0) {
      if (ia < Ka) atomicAdd(csumA + ia, accs[0]);
      if (ib < Ka) atomicAdd(csumA + ib, accs[2]);
    }
  }
}

template <int IB, int JB, bool RS>
static int launch_tn(const void* A, const void* B, const float* rowscale, float* dW, float* csumA, int64_t M, int Ka, int Nb,
                     int64_t lda, int64_t ldb, int64_t ldo, cudaStream_t st) {
  const int gi = (int)ceil_div(Ka, IB), gj = (int)ceil_div(Nb, JB);
  // one wave of 2 CTAs per SM, each streaming a contiguous slice of the token rows
  int64_t split = (int64_t)2 * sm_count() / ((int64_t)gi * gj);
  const int64_t max_split = ceil_div(M, 4 * MC);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  int64_t rpc = ceil_div(M, split);
  rpc = ceil_div(rpc, MC) * MC;
  split = ceil_div(M, rpc);
  constexpr size_t smem = (size_t)NST * (MC * (IB * 2 + 16 + JB * 2 + 16) + MC * 4);
  auto kern = gemm_tn_kernel<IB, JB, RS>;
  static cudaError_t attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr_err != cudaSuccess) { set_error("gemm_tn(tc): smem attr: %s", cudaGetErrorString(attr_err)); return STGCMA_ECUDA; }
  dim3 grid(gi, gj, (unsigned)split);
  kern<<<grid, NTH, smem, st>>>((const bf16*)A, (const bf16*)B, rowscale, dW, csumA, M, Ka, Nb, lda, ldb, ldo, rpc);
  return check_launch("gemm_tn(tc)");
}

}  // namespace tntc

// returns -1 when the arguments do not fit the tensor-core kernel (caller falls back to the CUDA-core one)
int gemm_tn_bf16(const void* A, const void* B, const float* rowscale, float* dW, float* csumA, int64_t M, int Ka, int Nb,
                 int64_t lda, int64_t ldb, int64_t ldo, cudaStream_t st) {
  using namespace tntc;
  if (lda % 8 || ldb % 8 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return -1;
  if (((Ka + 7) & ~7) > lda || ((Nb + 7) & ~7) > ldb) return -1;   // the 8-column padding must be readable
  if (Nb <= 64) {
    if (rowscale) return launch_tn<128, 64, true>(A, B, rowscale, dW, csumA, M, Ka, Nb, lda, ldb, ldo, st);
    return launch_tn<128, 64, false>(A, B, rowscale, dW, csumA, M, Ka, Nb, lda, ldb, ldo, st);
  }
  if (rowscale) return launch_tn<64, 128, true>(A, B, rowscale, dW, csumA, M, Ka, Nb, lda, ldb, ldo, st);
  return launch_tn<64, 128, false>(A, B, rowscale, dW, csumA, M, Ka, Nb, lda, ldb, ldo, st);
}

}  // namespace stgcma
