// dW += A^T . B for the trainable adapter weights (bf16 production path, warp-level tensor cores).
//   A [M, Ka] (row stride lda), B [M, Nb] (row stride ldb), contraction over the M token rows; dW fp32 [Ka, Nb].
// The shapes are extreme: M = 10^5..10^6 rows against 12..768 output rows / columns, i.e. a streaming
// reduction that must run at HBM speed (bytes = 2 M (Ka + Nb)); the CUDA-core version reached 350 GB/s.
// A CTA owns a 64 x 128 block of dW and a contiguous slice of M; 64-row chunks of A and B stream through a
// 3-stage cp.async pipeline; A^T fragments come from ldmatrix.trans on the [m][i] tile, B fragments from
// ldmatrix.trans on the [m][j] tile (the contraction index is the row index of both).  Partial sums are added
// to dW with fp32 atomics (the destination is a slice of the pre-zeroed gradient bucket).  The bias gradient
// (column sums of A) is one extra MMA per step against an all-ones B fragment held in registers.
#include "mma.cuh"

namespace stgcma {
namespace tntc {

using namespace mma;

constexpr int IB = 64;    // dW rows per CTA  (4 m16 tiles)
constexpr int JB = 128;   // dW cols per CTA  (16 n8 tiles)
constexpr int MC = 64;    // token rows per chunk
constexpr int NST = 3;    // pipeline stages
constexpr int LDA_S = IB * 2 + 16;   // 144 B
constexpr int LDB_S = JB * 2 + 16;   // 272 B
constexpr int NTH = 256;

// A fragment of A^T (rows i0..i0+15, k = m0..m0+15) from the [m][i] tile
__device__ __forceinline__ void load_at(uint32_t (&a)[4], const uint8_t* tile, int m0, int i0, int lane) {
  const int id = lane >> 3;
  const int m = m0 + (lane & 7) + ((id >> 1) & 1) * 8;
  const int i = i0 + (id & 1) * 8;
  ldsm_x4_t(a, smem_addr(tile + m * LDA_S + i * 2));
}

__global__ void __launch_bounds__(NTH) gemm_tn_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B,
                                                      float* __restrict__ dW, float* __restrict__ csumA, int64_t M, int Ka,
                                                      int Nb, int64_t lda, int64_t ldb, int64_t ldo, int64_t rows_per_cta) {
  extern __shared__ __align__(16) uint8_t tsm[];
  uint8_t* sa = tsm;                          // NST x [MC][LDA_S]
  uint8_t* sb = sa + NST * MC * LDA_S;        // NST x [MC][LDB_S]
  const int i_blk = blockIdx.x * IB, j_blk = blockIdx.y * JB;
  const int64_t m_beg = (int64_t)blockIdx.z * rows_per_cta;
  int64_t m_end = m_beg + rows_per_cta;
  if (m_end > M) m_end = M;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wi = warp & 3, wj = warp >> 2;    // m16 tile of the block, 64-column half of the block
  const int ka8 = (Ka + 7) & ~7, nb8 = (Nb + 7) & ~7;
  const int a_chunks = (min(IB, ka8 - i_blk) + 7) >> 3;    // valid 16-byte chunks per A row (rest zero-filled)
  const int b_chunks = (min(JB, nb8 - j_blk) + 7) >> 3;
  const int nchunk = (int)((m_end - m_beg + MC - 1) / MC);

  auto load_chunk = [&](int c) {
    const int st = c % NST;
    uint8_t* da = sa + st * MC * LDA_S;
    uint8_t* db = sb + st * MC * LDB_S;
    const int64_t m0 = m_beg + (int64_t)c * MC;
    for (int e = tid; e < MC * (IB / 8); e += NTH) {
      const int r = e >> 3, ch = e & 7;
      uint8_t* dst = da + r * LDA_S + ch * 16;
      if (m0 + r < m_end && ch < a_chunks) cp_async16(dst, A + (m0 + r) * lda + i_blk + ch * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    for (int e = tid; e < MC * (JB / 8); e += NTH) {
      const int r = e >> 4, ch = e & 15;
      uint8_t* dst = db + r * LDB_S + ch * 16;
      if (m0 + r < m_end && ch < b_chunks) cp_async16(dst, B + (m0 + r) * ldb + j_blk + ch * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
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
    if (c + NST - 1 < nchunk) load_chunk(c + NST - 1);
    cp_async_commit();
    cp_async_wait<NST - 1>();
    __syncthreads();
    const uint8_t* ca = sa + (c % NST) * MC * LDA_S;
    const uint8_t* cb = sb + (c % NST) * MC * LDB_S;
    if (i_active && n_act > 0) {
#pragma unroll
      for (int ks = 0; ks < MC / 16; ++ks) {
        uint32_t a[4];
        load_at(a, ca, ks * 16, wi * 16, lane);
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
    __syncthreads();
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

}  // namespace tntc

// returns -1 when the arguments do not fit the tensor-core kernel (caller falls back to the CUDA-core one)
int gemm_tn_bf16(const void* A, const void* B, float* dW, float* csumA, int64_t M, int Ka, int Nb, int64_t lda, int64_t ldb,
                 int64_t ldo, cudaStream_t st) {
  using namespace tntc;
  if (lda % 8 || ldb % 8 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return -1;
  if (((Ka + 7) & ~7) > lda || ((Nb + 7) & ~7) > ldb) return -1;   // the 8-column padding must be readable
  const int gi = (int)ceil_div(Ka, IB), gj = (int)ceil_div(Nb, JB);
  int64_t split = ceil_div((int64_t)4 * sm_count(), (int64_t)gi * gj);
  const int64_t max_split = ceil_div(M, 4 * MC);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  int64_t rpc = ceil_div(M, split);
  rpc = ceil_div(rpc, MC) * MC;
  split = ceil_div(M, rpc);
  const size_t smem = (size_t)NST * MC * (LDA_S + LDB_S);
  static cudaError_t attr_err = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr_err != cudaSuccess) { set_error("gemm_tn(tc): smem attr: %s", cudaGetErrorString(attr_err)); return STGCMA_ECUDA; }
  dim3 grid(gi, gj, (unsigned)split);
  gemm_tn_kernel<<<grid, NTH, smem, st>>>((const bf16*)A, (const bf16*)B, dW, csumA, M, Ka, Nb, lda, ldb, ldo, rpc);
  return check_launch("gemm_tn(tc)");
}

}  // namespace stgcma
