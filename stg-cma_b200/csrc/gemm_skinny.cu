// out = res1 [+ res2] + rowscale * (A . W^T + bias) for SKINNY contractions (K = adapter width r, 16..128) --
// the adapters' up-projection D_fc2 with the block's residual adds (Swin_AVE.py:57, :762-763, :783, :810-811) and
// the matching dX of the down-projection (dAt = dHpre . W1 + dX).
//
// With K = 16..48 against N = C = 96..768 this is not a GEMM problem but an HBM stream: per output row it reads
// 2 K bytes of A and 4..6 C bytes of residuals and writes 2..4 C bytes; the tensor work is ~1 % of the time.  The
// tcgen05 kernel (gemm.cu) spends its epilogue transposing TMEM rows into a coalesced layout, one 32 x 32 chunk per
// warp at a time, and reached 3.3-3.5 TB/s on these shapes (profiles/r01f).  Here every warp owns a 16-row x 48..128
// column block: the residual operands are loaded straight into the (column-permuted) mma.sync accumulator layout,
// ALL of a block's loads are issued before the first is consumed, and 16 resident warps per SM keep ~75 KB in flight.  W (one modality group per CTA) and the bias live in shared memory.
#include "mma.cuh"

namespace stgcma {
namespace skinny {

using namespace mma;

constexpr int NTH = 256;

__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

// NT: n8 tiles per block (fp32 stream: 6 -> 48 columns when N % 48 == 0, else 8; bf16: 12 or 16), handled as NT/2 PAIRS: the n index
// of the B operand is permuted when W is copied to shared memory so that the two tiles of a pair give lane t the four
// CONSECUTIVE output columns 16p + 4t .. 4t+3 (tile A: MMA column 2t+e -> 4t+e, tile B: -> 4t+2+e).  Residual loads
// and stores are then 16 bytes (fp32) / 8 bytes (bf16) per lane and a quad covers 64 / 32 contiguous bytes of a row.
// F32: out / res1 are fp32 (the residual stream), else bf16.  NRES: 1 (res1) or 2 (res1 + res2, bf16).
template <int NT, bool F32, int NRES>
__global__ void __launch_bounds__(NTH, 2)
skinny_kernel(const bf16* __restrict__ A, const bf16* __restrict__ W, const float* __restrict__ bias,
              const void* __restrict__ res1, const bf16* __restrict__ res2, const float* __restrict__ rowscale,
              void* __restrict__ out, int64_t group_rows, int N, int K, int64_t lda, int64_t ldw, int64_t ldo) {
  constexpr int NPR = NT / 2;
  extern __shared__ __align__(16) uint32_t ssm[];
  const int grp = blockIdx.y;
  const int wst = K / 2 + 4;                 // words per W row: +4 keeps the 8 rows of a B fragment in distinct banks
  uint32_t* sW = ssm;                        // [N][wst], rows in MMA order (see above)
  float* sB = reinterpret_cast<float*>(sW + (size_t)N * wst);   // [N], natural order
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  {
    const int cpr = K / 8;
    const bf16* Wg = W + (int64_t)grp * N * ldw;
    for (int e = tid; e < N * cpr; e += NTH) {
      const int srow = e / cpr, c = e - srow * cpr;
      const int q = srow & 15, j = q & 7;
      const int n = (srow & ~15) + 4 * (j >> 1) + (j & 1) + 2 * (q >> 3);   // output column held by smem row `srow`
      *reinterpret_cast<uint4*>(sW + (size_t)srow * wst + c * 4) = __ldg(reinterpret_cast<const uint4*>(Wg + (int64_t)n * ldw + c * 8));
    }
    for (int n = tid; n < N; n += NTH) sB[n] = bias ? __ldg(bias + (int64_t)grp * N + n) : 0.f;
  }
  __syncthreads();

  const int nchunks = (N + NT * 8 - 1) / (NT * 8);
  const int64_t slabs = (group_rows + 15) >> 4;
  const int64_t items = slabs * nchunks;
  const int ksteps = K >> 4;
  for (int64_t item = (int64_t)blockIdx.x * (NTH / 32) + warp; item < items; item += (int64_t)gridDim.x * (NTH / 32)) {
    const int64_t slab = item / nchunks;
    const int ch = (int)(item - slab * nchunks);
    const int64_t r0 = slab * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < group_rows, v1 = r1 < group_rows;
    const int64_t gr0 = (int64_t)grp * group_rows + r0, gr1 = gr0 + 8;
    const int col0 = ch * NT * 8 + 4 * t;    // first of this lane's 4 columns in pair 0
    // ---- issue every global load of the block ----
    float4 x0[NPR], x1[NPR];   // res1: fp32 x4, or 4 bf16 packed in .x/.y
    uint2 y0[NPR], y1[NPR];    // res2: 4 bf16
#pragma unroll
    for (int p = 0; p < NPR; ++p) {
      const int col = col0 + p * 16;
      const bool ok = col < N;   // N % 16 == 0: a pair is inside or outside as a whole
      if (F32) {
        const float* rp = reinterpret_cast<const float*>(res1);
        x0[p] = (ok && v0) ? *reinterpret_cast<const float4*>(rp + gr0 * ldo + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        x1[p] = (ok && v1) ? *reinterpret_cast<const float4*>(rp + gr1 * ldo + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const bf16* rp = reinterpret_cast<const bf16*>(res1);
        const uint2 u0 = (ok && v0) ? *reinterpret_cast<const uint2*>(rp + gr0 * ldo + col) : make_uint2(0u, 0u);
        const uint2 u1 = (ok && v1) ? *reinterpret_cast<const uint2*>(rp + gr1 * ldo + col) : make_uint2(0u, 0u);
        x0[p].x = __uint_as_float(u0.x); x0[p].y = __uint_as_float(u0.y);
        x1[p].x = __uint_as_float(u1.x); x1[p].y = __uint_as_float(u1.y);
      }
      if (NRES >= 2) {
        y0[p] = (ok && v0) ? *reinterpret_cast<const uint2*>(res2 + gr0 * ldo + col) : make_uint2(0u, 0u);
        y1[p] = (ok && v1) ? *reinterpret_cast<const uint2*>(res2 + gr1 * ldo + col) : make_uint2(0u, 0u);
      }
    }
    float rs0 = 1.f, rs1 = 1.f;
    if (rowscale) {
      rs0 = v0 ? __ldg(rowscale + gr0) : 0.f;
      rs1 = v1 ? __ldg(rowscale + gr1) : 0.f;
    }
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    const bf16* a0p = A + gr0 * lda + 2 * t;
    const bf16* a1p = A + gr1 * lda + 2 * t;
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
      uint32_t a[4];
      a[0] = v0 ? *reinterpret_cast<const uint32_t*>(a0p + ks * 16) : 0u;
      a[1] = v1 ? *reinterpret_cast<const uint32_t*>(a1p + ks * 16) : 0u;
      a[2] = v0 ? *reinterpret_cast<const uint32_t*>(a0p + ks * 16 + 8) : 0u;
      a[3] = v1 ? *reinterpret_cast<const uint32_t*>(a1p + ks * 16 + 8) : 0u;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int srow = ch * NT * 8 + n * 8 + g;
        if (srow - g < N) {
          const uint32_t* wp = sW + (size_t)srow * wst + ks * 8 + t;
          mma16816(acc[n], a, wp[0], wp[4]);
        }
      }
    }
    // ---- epilogue: (acc + bias) * rowscale + res1 + res2 ----
#pragma unroll
    for (int p = 0; p < NPR; ++p) {
      const int col = col0 + p * 16;
      if (col >= N) continue;
      const float4 b = *reinterpret_cast<const float4*>(sB + col);
      float o0[4] = {(acc[2 * p][0] + b.x) * rs0, (acc[2 * p][1] + b.y) * rs0, (acc[2 * p + 1][0] + b.z) * rs0, (acc[2 * p + 1][1] + b.w) * rs0};
      float o1[4] = {(acc[2 * p][2] + b.x) * rs1, (acc[2 * p][3] + b.y) * rs1, (acc[2 * p + 1][2] + b.z) * rs1, (acc[2 * p + 1][3] + b.w) * rs1};
      if (F32) {
        o0[0] += x0[p].x; o0[1] += x0[p].y; o0[2] += x0[p].z; o0[3] += x0[p].w;
        o1[0] += x1[p].x; o1[1] += x1[p].y; o1[2] += x1[p].z; o1[3] += x1[p].w;
      } else {
        const float2 p00 = bf2_to_f2(__float_as_uint(x0[p].x)), p01 = bf2_to_f2(__float_as_uint(x0[p].y));
        const float2 p10 = bf2_to_f2(__float_as_uint(x1[p].x)), p11 = bf2_to_f2(__float_as_uint(x1[p].y));
        o0[0] += p00.x; o0[1] += p00.y; o0[2] += p01.x; o0[3] += p01.y;
        o1[0] += p10.x; o1[1] += p10.y; o1[2] += p11.x; o1[3] += p11.y;
      }
      if (NRES >= 2) {
        const float2 q00 = bf2_to_f2(y0[p].x), q01 = bf2_to_f2(y0[p].y), q10 = bf2_to_f2(y1[p].x), q11 = bf2_to_f2(y1[p].y);
        o0[0] += q00.x; o0[1] += q00.y; o0[2] += q01.x; o0[3] += q01.y;
        o1[0] += q10.x; o1[1] += q10.y; o1[2] += q11.x; o1[3] += q11.y;
      }
      if (F32) {
        float* op = reinterpret_cast<float*>(out);
        if (v0) *reinterpret_cast<float4*>(op + gr0 * ldo + col) = make_float4(o0[0], o0[1], o0[2], o0[3]);
        if (v1) *reinterpret_cast<float4*>(op + gr1 * ldo + col) = make_float4(o1[0], o1[1], o1[2], o1[3]);
      } else {
        bf16* op = reinterpret_cast<bf16*>(out);
        if (v0) *reinterpret_cast<uint2*>(op + gr0 * ldo + col) = make_uint2(pack_bf16(o0[0], o0[1]), pack_bf16(o0[2], o0[3]));
        if (v1) *reinterpret_cast<uint2*>(op + gr1 * ldo + col) = make_uint2(pack_bf16(o1[0], o1[1]), pack_bf16(o1[2], o1[3]));
      }
    }
  }
}

template <int NT, bool F32, int NRES>
static int launch(const void* A, const void* W, int64_t M, int N, int K, int64_t lda, int64_t ldw, int groups,
                  const EpiParams& ep, size_t smem, cudaStream_t st) {
  auto kern = skinny_kernel<NT, F32, NRES>;
  static cudaError_t attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (attr_err != cudaSuccess) { set_error("gemm(skinny): smem attr: %s", cudaGetErrorString(attr_err)); return STGCMA_ECUDA; }
  const int64_t group_rows = M / groups;
  const int64_t items = ((group_rows + 15) / 16) * ((N + NT * 8 - 1) / (NT * 8));
  int64_t ctas = (items + NTH / 32 - 1) / (NTH / 32);
  const int64_t cap = ((int64_t)2 * sm_count() + groups - 1) / groups;
  if (ctas > cap) ctas = cap;
  dim3 grid((unsigned)ctas, (unsigned)groups);
  kern<<<grid, NTH, smem, st>>>((const bf16*)A, (const bf16*)W, ep.bias, ep.res1, (const bf16*)ep.res2, ep.rowscale, ep.out,
                                group_rows, N, K, lda, ldw, ep.ldo);
  return check_launch("gemm(skinny)");
}

}  // namespace skinny

// returns -1 when the call is not a skinny residual GEMM (the caller then uses the tcgen05 kernel)
int gemm_skinny_bf16(const void* A, const void* W, int64_t M, int N, int K, int64_t lda, int64_t ldw, int groups,
                     const EpiParams& ep, cudaStream_t st) {
  using namespace skinny;
  if (ep.mode != STGCMA_EPI_LINEAR || !ep.res1 || ep.out_pre || K % 16 || K > 128 || N % 16 || (ep.ldo & 3)) return -1;
  const size_t smem = (size_t)N * (K / 2 + 4) * 4 + (size_t)N * 4;
  if (smem > 100 * 1024) return -1;
  // bf16 out / res1 (dX of D_fc1): a third of the bytes per column, so the register-staged loads of 16 warps per SM no
  // longer cover the memory latency at large N (measured: 2.9 TB/s at N = 384 against 3.4 TB/s for the tcgen05 kernel,
  // but 4.7 against 2.9 TB/s at N = 96)
  if (!ep.out_f32 && N > 128) return -1;
  const bool two = ep.res2 != nullptr;
  // bf16 out / res1 (the dX variant) moves a third of the bytes per column: twice as many columns per warp block keep
  // the same number of bytes in flight
  if (ep.out_f32) {
    if (N % 48 == 0) {
      if (two) return launch<6, true, 2>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
      return launch<6, true, 1>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
    }
    if (two) return launch<8, true, 2>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
    return launch<8, true, 1>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
  }
  if (N % 96 == 0) {
    if (two) return launch<12, false, 2>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
    return launch<12, false, 1>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
  }
  if (two) return launch<16, false, 2>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
  return launch<16, false, 1>(A, W, M, N, K, lda, ldw, groups, ep, smem, st);
}

}  // namespace stgcma
