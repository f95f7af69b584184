// Memory-bound kernels of the Swin AVE adapter train step: patch embed (+LayerNorm),
// LayerNorm forward / frozen-affine backward (optionally fused with the PatchMerging 2x2
// gather / scatter), head mean-pool + cat.  All 128-bit vectorised, fp32 math.
#include "common.cuh"

namespace stgcma {

// =========================================================================================
// LayerNorm: one warp per row.
// =========================================================================================
template <bool MERGE>
struct RowAddr {
  // element offset of channel c of logical row `row`
  int64_t base[4];
  int C4;
  __device__ __forceinline__ RowAddr(int64_t row, int C, int R) {
    if (MERGE) {
      C4 = C >> 2;
      const int Rh = R >> 1;
      const int64_t f = row / (Rh * Rh);
      const int rem = (int)(row - f * Rh * Rh);
      const int yy = rem / Rh, xx = rem - yy * Rh;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int dy = q & 1, dx = q >> 1;  // cat order x0,x1,x2,x3 = (0,0),(1,0),(0,1),(1,1)  Swin_AVE.py:971-975
        base[q] = ((f * R + (2 * yy + dy)) * R + (2 * xx + dx)) * (int64_t)C4;
      }
    } else {
      C4 = C;
      base[0] = row * (int64_t)C;
    }
  }
  __device__ __forceinline__ int64_t at(int c) const {
    if (MERGE) {
      const int q = c / C4;
      return base[q] + (c - q * C4);
    }
    return base[0] + c;
  }
};

// TX: type of x (the residual stream), TY: type of y.  8 channels per lane per step.
template <typename TX, typename TY, bool MERGE>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, TY* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int64_t rows, int C, int R) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  RowAddr<MERGE> addr(row, C, R);
  float s = 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    float t[8];
    load8(x + addr.at(c), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += t[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    float t[8];
    load8(x + addr.at(c), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = t[i] - mean;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
  TY* yrow = y + row * (int64_t)C;
  for (int c = lane * 8; c < C; c += 256) {
    float t[8];
    load8(x + addr.at(c), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = (t[i] - mean) * rstd * __ldg(gamma + c + i) + __ldg(beta + c + i);
    store8(yrow + c, t);
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// TD: type of dy, TX: type of x / dres / dx (stream), TC: type of the optional operand copy of dx.
template <typename TD, typename TX, typename TC, bool MERGE>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const TD* __restrict__ dy, const TX* __restrict__ x,
                                                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                     const float* __restrict__ gamma, const TX* dres, TX* dx,
                                                     TC* __restrict__ dx_copy, int64_t rows, int C, int R) {  // dres may alias dx
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  RowAddr<MERGE> addr(row, C, R);
  const float mean = mean_in[row], rstd = rstd_in[row];
  const TD* dyrow = dy + row * (int64_t)C;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    float g[8], t[8];
    load8(dyrow + c, g);
    load8(x + addr.at(c), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gg = g[i] * __ldg(gamma + c + i);
      s1 += gg;
      s2 += gg * (t[i] - mean) * rstd;
    }
  }
  s1 = warp_sum(s1) / (float)C;
  s2 = warp_sum(s2) / (float)C;
  for (int c = lane * 8; c < C; c += 256) {
    float g[8], t[8], o[8];
    load8(dyrow + c, g);
    const int64_t off = addr.at(c);
    load8(x + off, t);
    if (dres) load8(dres + off, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gg = g[i] * __ldg(gamma + c + i);
      const float xh = (t[i] - mean) * rstd;
      const float d = rstd * (gg - s1 - xh * s2);
      o[i] = dres ? o[i] + d : d;
    }
    store8(dx + off, o);
    if (dx_copy) store8(dx_copy + off, o);
  }
}

// -----------------------------------------------------------------------------------------
// Row-in-registers variants (the ones the engine hits): LPR lanes own a row (32 / LPR rows per warp), each lane
// holds VPL <= VMAX 4-element vectors, so x (and dy) are read from global memory exactly once and every lane of
// the warp is busy for any C.  The warp-per-row kernels above kept 12 of 32 lanes busy at C = 96 and walked the
// row three times: 26 % (fwd) / 52 % (bwd) of HBM peak (profiles/r01d).
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TX, typename TY, bool MERGE, int LPR, int VMAX>
__global__ void __launch_bounds__(256) ln_fwd_rr_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, TY* __restrict__ y,
                                                        float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                        int64_t rows, int C, int R, int vpl) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool ok = row < rows;
  RowAddr<MERGE> addr(ok ? row : 0, C, R);
  float v[VMAX][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    if (i < vpl && ok) {
      load4(x + addr.at(4 * (sub + LPR * i)), v[i]);
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    } else {
      v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
    }
  }
  const float invC = 1.f / (float)C;
  const float mean = group_sum<LPR>(s) * invC;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i)
    if (i < vpl) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = v[i][e] - mean;
        q += d * d;
      }
    }
  const float rstd = rsqrtf(group_sum<LPR>(q) * invC + 1e-5f);
  if (!ok) return;
  TY* yrow = y + row * (int64_t)C;
#pragma unroll
  for (int i = 0; i < VMAX; ++i)
    if (i < vpl) {
      const int c = 4 * (sub + LPR * i);
      float g[4], b[4], o[4];
      load4(gamma + c, g);
      load4(beta + c, b);
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (v[i][e] - mean) * rstd * g[e] + b[e];
      store4(yrow + c, o);
    }
  if (sub == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

template <typename TD, typename TX, typename TC, bool MERGE, int LPR, int VMAX>
__global__ void __launch_bounds__(256) ln_bwd_rr_kernel(const TD* __restrict__ dy, const TX* __restrict__ x,
                                                        const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                        const float* __restrict__ gamma, const TX* dres, TX* dx,
                                                        TC* __restrict__ dx_copy, int64_t rows, int C, int R, int vpl) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool ok = row < rows;
  RowAddr<MERGE> addr(ok ? row : 0, C, R);
  const float mean = ok ? mean_in[row] : 0.f, rstd = ok ? rstd_in[row] : 0.f;
  const TD* dyrow = dy + (ok ? row : 0) * (int64_t)C;
  float gg[VMAX][4], xh[VMAX][4];
  // Rows up to 1024 channels (VMAX <= 8): the residual gradient is loaded together with dy / x, before the row reduction, so that a
  // row has ONE exposed memory round trip instead of two (every load of the row is in flight at once).  dres may alias
  // dx: this lane still reads its own elements before it writes them.
  constexpr bool EARLY = VMAX <= 8;
  float rr[EARLY ? VMAX : 1][4];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    if (i < vpl && ok) {
      const int c = 4 * (sub + LPR * i);
      float d4[4], x4[4], g4[4];
      load4(dyrow + c, d4);
      load4(x + addr.at(c), x4);
      if (EARLY && dres) load4(dres + addr.at(c), rr[EARLY ? i : 0]);
      load4(gamma + c, g4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        gg[i][e] = d4[e] * g4[e];
        xh[i][e] = (x4[e] - mean) * rstd;
        s1 += gg[i][e];
        s2 += gg[i][e] * xh[i][e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) gg[i][e] = xh[i][e] = 0.f;
    }
  }
  const float invC = 1.f / (float)C;
  s1 = group_sum<LPR>(s1) * invC;
  s2 = group_sum<LPR>(s2) * invC;
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < VMAX; ++i)
    if (i < vpl) {
      const int64_t off = addr.at(4 * (sub + LPR * i));
      float o[4];
      if (dres) {
        float r4[4];
        if (EARLY) {
#pragma unroll
          for (int e = 0; e < 4; ++e) r4[e] = rr[EARLY ? i : 0][e];
        } else {
          load4(dres + off, r4);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = r4[e] + rstd * (gg[i][e] - s1 - xh[i][e] * s2);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = rstd * (gg[i][e] - s1 - xh[i][e] * s2);
      }
      store4(dx + off, o);
      if (dx_copy) store4(dx_copy + off, o);
    }
}

// (LPR, VMAX) for a row of C channels; false when C does not fit the row-in-registers kernels
static bool rr_config(int C, int& lpr, int& vpl, int& vmax) {
  if (C % 4) return false;
  const int per4 = C / 4;
  lpr = per4 <= 32 ? 8 : (per4 <= 64 ? 16 : 32);
  if (per4 % lpr) return false;
  vpl = per4 / lpr;
  if (vpl > 12) return false;
  vmax = (vpl <= 4) ? 4 : (vpl <= 8 ? 8 : 12);    // 8: C = 544 .. 1024 (CLIP ViT-B/16 D = 768, Swin-T stage 3)
  if (vmax > 4 && lpr != 32) return false;
  return true;
}

template <typename TX, typename TY, bool MERGE, int LPR, int VMAX>
static void ln_fwd_rr_launch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                             int64_t rows, int C, int R, int vpl, cudaStream_t st) {
  const int rpb = 8 * (32 / LPR);
  ln_fwd_rr_kernel<TX, TY, MERGE, LPR, VMAX><<<(unsigned)ceil_div(rows, rpb), 256, 0, st>>>((const TX*)x, gamma, beta, (TY*)y, mean,
                                                                                         rstd, rows, C, R, vpl);
}

template <typename TX, typename TY, bool MERGE>
static bool ln_fwd_rr_dispatch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                               int64_t rows, int C, int R, cudaStream_t st) {
  int lpr, vpl, vmax;
  if (!rr_config(C, lpr, vpl, vmax)) return false;
  if (vmax == 12) ln_fwd_rr_launch<TX, TY, MERGE, 32, 12>(x, gamma, beta, y, mean, rstd, rows, C, R, vpl, st);
  else if (vmax == 8) ln_fwd_rr_launch<TX, TY, MERGE, 32, 8>(x, gamma, beta, y, mean, rstd, rows, C, R, vpl, st);
  else if (lpr == 8) ln_fwd_rr_launch<TX, TY, MERGE, 8, 4>(x, gamma, beta, y, mean, rstd, rows, C, R, vpl, st);
  else if (lpr == 16) ln_fwd_rr_launch<TX, TY, MERGE, 16, 4>(x, gamma, beta, y, mean, rstd, rows, C, R, vpl, st);
  else ln_fwd_rr_launch<TX, TY, MERGE, 32, 4>(x, gamma, beta, y, mean, rstd, rows, C, R, vpl, st);
  return true;
}

template <typename TX, typename TY>
static int ln_fwd_launch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                         int64_t rows, int C, int merge, int R, cudaStream_t st) {
  if (merge ? ln_fwd_rr_dispatch<TX, TY, true>(x, gamma, beta, y, mean, rstd, rows, C, R, st)
            : ln_fwd_rr_dispatch<TX, TY, false>(x, gamma, beta, y, mean, rstd, rows, C, R, st))
    return check_launch("layernorm_fwd");
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div(rows, wpb)), block(wpb * 32);
  if (merge)
    ln_fwd_kernel<TX, TY, true><<<grid, block, 0, st>>>((const TX*)x, gamma, beta, (TY*)y, mean, rstd, rows, C, R);
  else
    ln_fwd_kernel<TX, TY, false><<<grid, block, 0, st>>>((const TX*)x, gamma, beta, (TY*)y, mean, rstd, rows, C, R);
  return check_launch("layernorm_fwd");
}

template <typename TD, typename TX, typename TC, bool MERGE, int LPR, int VMAX>
static void ln_bwd_rr_launch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                             const void* dres, void* dx, void* dx_copy, int64_t rows, int C, int R, int vpl, cudaStream_t st) {
  const int rpb = 8 * (32 / LPR);
  ln_bwd_rr_kernel<TD, TX, TC, MERGE, LPR, VMAX><<<(unsigned)ceil_div(rows, rpb), 256, 0, st>>>(
      (const TD*)dy, (const TX*)x, mean, rstd, gamma, (const TX*)dres, (TX*)dx, (TC*)dx_copy, rows, C, R, vpl);
}

template <typename TD, typename TX, typename TC, bool MERGE>
static bool ln_bwd_rr_dispatch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                               const void* dres, void* dx, void* dx_copy, int64_t rows, int C, int R, cudaStream_t st) {
  int lpr, vpl, vmax;
  if (!rr_config(C, lpr, vpl, vmax)) return false;
  if (vmax == 12) ln_bwd_rr_launch<TD, TX, TC, MERGE, 32, 12>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, R, vpl, st);
  else if (vmax == 8) ln_bwd_rr_launch<TD, TX, TC, MERGE, 32, 8>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, R, vpl, st);
  else if (lpr == 8) ln_bwd_rr_launch<TD, TX, TC, MERGE, 8, 4>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, R, vpl, st);
  else if (lpr == 16) ln_bwd_rr_launch<TD, TX, TC, MERGE, 16, 4>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, R, vpl, st);
  else ln_bwd_rr_launch<TD, TX, TC, MERGE, 32, 4>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, R, vpl, st);
  return true;
}

template <typename TD, typename TX, typename TC>
static int ln_bwd_launch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                         const void* dres, void* dx, void* dx_copy, int64_t rows, int C, int merge, int R, cudaStream_t st) {
  if (merge ? ln_bwd_rr_dispatch<TD, TX, TC, true>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, R, st)
            : ln_bwd_rr_dispatch<TD, TX, TC, false>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, R, st))
    return check_launch("layernorm_bwd");
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div(rows, wpb)), block(wpb * 32);
  if (merge)
    ln_bwd_kernel<TD, TX, TC, true><<<grid, block, 0, st>>>((const TD*)dy, (const TX*)x, mean, rstd, gamma, (const TX*)dres,
                                                            (TX*)dx, (TC*)dx_copy, rows, C, R);
  else
    ln_bwd_kernel<TD, TX, TC, false><<<grid, block, 0, st>>>((const TD*)dy, (const TX*)x, mean, rstd, gamma, (const TX*)dres,
                                                             (TX*)dx, (TC*)dx_copy, rows, C, R);
  return check_launch("layernorm_bwd");
}

// =========================================================================================
// Patch embed: Conv3d k=s=(1,4,4) == per-token Linear(Cin*16 -> C), + bias + LayerNorm(C).
// One block streams (frame, patch-row) strips: the 4 x W x Cin input strip is staged in smem
// with coalesced loads, the [Cin*16, C] transposed weight stays resident in smem, each warp
// produces whole tokens (lane = channel mod 32) so the LayerNorm is a warp reduction.
// =========================================================================================
template <typename T, int PER>
__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, T* __restrict__ out, int B,
                                                          int Cin, int Tn, int H, int W, int C) {
  extern __shared__ float smem[];
  const int K = Cin * 16;
  float* wt = smem;                 // [K][C]
  float* sb = wt + K * C;           // bias, gamma, beta: [3][C]
  float* sin0 = sb + 3 * C;         // 2 x [Cin*4][W]: the next strip is copied in (cp.async) while this one is computed
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  for (int i = tid; i < K * C; i += blockDim.x) {
    const int c = i / K, k = i - c * K;  // w is [C][K]
    wt[k * C + c] = w[i];
  }
  for (int i = tid; i < C; i += blockDim.x) {
    sb[i] = bias[i];
    sb[C + i] = gamma[i];
    sb[2 * C + i] = beta[i];
  }
  const int Hp = H >> 2, Wp = W >> 2;
  const int64_t nstrips = (int64_t)B * Tn * Hp;
  const int W4 = W >> 2, nvec = Cin * 4 * W4;      // 16-byte chunks of a strip (W % 4 == 0, img 16-byte aligned: host checks)
  // The strip copy used to be a scalar global -> shared loop with a division per element and its load latency exposed
  // once per strip (10 dependent round trips: 3/4 of the kernel's time, 0.19 of the HBM roof); now 16-byte cp.async into the
  // other half of a double buffer, issued one strip ahead.
  auto copy_strip = [&](int64_t strip, float* dst) {
    const int y = (int)(strip % Hp);
    const int64_t f = strip / Hp;
    const int t = (int)(f % Tn);
    const int b = (int)(f / Tn);
    for (int i = tid; i < nvec; i += blockDim.x) {
      const int rowi = i / W4, x4 = i - rowi * W4;
      const int ci = rowi >> 2, ky = rowi & 3;
      const float* src = img + ((((int64_t)b * Cin + ci) * Tn + t) * H + (4 * y + ky)) * W + 4 * x4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + rowi * W + 4 * x4)),
                   "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if ((int64_t)blockIdx.x < nstrips) copy_strip(blockIdx.x, sin0);
  for (int64_t strip = blockIdx.x; strip < nstrips; strip += gridDim.x, buf ^= 1) {
    const int y = (int)(strip % Hp);
    const int64_t f = strip / Hp;  // frame index b*T + t
    const float* sin = sin0 + buf * (Cin * 4 * W);
    __syncthreads();               // every warp is done with the buffer the next copy overwrites (and wt / sb are loaded)
    if (strip + gridDim.x < nstrips) {
      copy_strip(strip + gridDim.x, sin0 + (buf ^ 1) * (Cin * 4 * W));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // Register tile: a warp owns TT tokens (x = warp + nwarp * i) x `per` channels per lane, so that every staged
    // input value and every weight is read from shared memory once per TT (resp. once per token tile) instead of
    // once per FMA: the first version issued 4 LDS per 3 FMA and ran 20x off the HBM roofline (profiles/r01d).
    constexpr int TT = 7;
    for (int x0 = 0; x0 < Wp; x0 += TT * nwarp) {
      float acc[TT][PER];
#pragma unroll
      for (int i = 0; i < TT; ++i)
#pragma unroll
        for (int j = 0; j < PER; ++j) acc[i][j] = sb[lane + 32 * j];
      for (int k4 = 0; k4 < K; k4 += 4) {       // k4 = (ci, ky): the 4 kx taps are contiguous in the strip
        const int rowi = k4 >> 2;
        float wv[4][PER];
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
#pragma unroll
          for (int j = 0; j < PER; ++j) wv[kx][j] = wt[(k4 + kx) * C + lane + 32 * j];
#pragma unroll
        for (int i = 0; i < TT; ++i) {
          const int x = x0 + warp + nwarp * i;
          const float4 xv = x < Wp ? *reinterpret_cast<const float4*>(sin + rowi * W + 4 * x) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < PER; ++j) {
            acc[i][j] = fmaf(xv.x, wv[0][j], acc[i][j]);
            acc[i][j] = fmaf(xv.y, wv[1][j], acc[i][j]);
            acc[i][j] = fmaf(xv.z, wv[2][j], acc[i][j]);
            acc[i][j] = fmaf(xv.w, wv[3][j], acc[i][j]);
          }
        }
      }
      float mu[TT], rs[TT];
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) sum += acc[i][j];
        mu[i] = sum;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < TT; ++i) mu[i] += __shfl_xor_sync(0xffffffffu, mu[i], o);
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        mu[i] /= (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const float d = acc[i][j] - mu[i];
          q += d * d;
        }
        rs[i] = q;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < TT; ++i) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        const int x = x0 + warp + nwarp * i;
        if (x >= Wp) continue;
        const float rstd = rsqrtf(rs[i] / (float)C + 1e-5f);
        T* o = out + ((f * Hp + y) * (int64_t)Wp + x) * C;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int c = lane + 32 * j;
          st_f(o + c, (acc[i][j] - mu[i]) * rstd * sb[C + c] + sb[2 * C + c]);
        }
      }
    }
  }
}

// =========================================================================================
// Head: mean over the N tokens of each frame, feat = [audio | video].
// =========================================================================================
template <typename T>
__global__ void head_pool_fwd_kernel(const T* __restrict__ xn, T* __restrict__ feat, int64_t frames, int N, int C) {
  const int64_t f = blockIdx.x;
  const int m = blockIdx.y;  // 0 video, 1 audio
  const T* src = xn + ((int64_t)m * frames + f) * N * C;
  T* dst = feat + f * 2 * (int64_t)C + (m == 1 ? 0 : C);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += ld_f(src + (int64_t)n * C + c);
    st_f(dst + c, s / (float)N);
  }
}

template <typename T>
__global__ void head_pool_bwd_kernel(const T* __restrict__ dfeat, T* __restrict__ dxn, int64_t frames, int N, int C) {
  const int64_t f = blockIdx.x;
  const int m = blockIdx.y;
  T* dst = dxn + ((int64_t)m * frames + f) * N * C;
  const T* src = dfeat + f * 2 * (int64_t)C + (m == 1 ? 0 : C);
  const float inv = 1.f / (float)N;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = ld_f(src + c) * inv;
    for (int n = 0; n < N; ++n) st_f(dst + (int64_t)n * C + c, g);
  }
}

// =========================================================================================
// Trainable-parameter refresh: a table of small 2-D copies (cast / pad / transpose) in one launch.
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(256) pack_params_kernel(const StgcmaPackEntry* __restrict__ table) {
  const StgcmaPackEntry e = table[blockIdx.x];
  const int64_t n = (int64_t)e.rows * e.cols;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.y * blockDim.x) {
    const int r = (int)(i / e.cols), c = (int)(i - (int64_t)r * e.cols);
    const float v = e.src[(int64_t)r * e.src_ld + c];
    const int64_t o = e.transpose ? (int64_t)c * e.dst_ld + r : (int64_t)r * e.dst_ld + c;
    if (e.dst_f32) reinterpret_cast<float*>(e.dst)[o] = v;
    else st_f(reinterpret_cast<T*>(e.dst) + o, v);
  }
}

// =========================================================================================
// Adam over a table of tensors (one launch for the whole optimizer step)
// =========================================================================================
__global__ void __launch_bounds__(256) adam_kernel(const StgcmaAdamEntry* __restrict__ table, const StgcmaAdamGroups hp,
                                                   float grad_scale) {
  const StgcmaAdamEntry e = table[blockIdx.x];
  const int gi = e.group;
  const float lr = hp.lr[gi], wd = hp.weight_decay[gi], b1 = hp.beta1[gi], b2 = hp.beta2[gi], eps = hp.eps[gi];
  const float step_size = lr / hp.bias_corr1[gi], bc2s = hp.bias_corr2_sqrt[gi];
  const float omb1 = hp.one_minus_beta1[gi], omb2 = hp.one_minus_beta2[gi];
  (void)b1;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < e.numel; i += (int64_t)gridDim.y * blockDim.x) {
    const float p = e.param[i];
    float g = e.grad[i] * grad_scale;
    g = fmaf(wd, p, g);                                   // L2 weight decay on the gradient (torch.optim.Adam)
    const float m = fmaf(omb1, g - e.exp_avg[i], e.exp_avg[i]);              // lerp, as torch does
    const float v = fmaf(omb2 * g, g, b2 * e.exp_avg_sq[i]);
    e.exp_avg[i] = m;
    e.exp_avg_sq[i] = v;
    const float denom = sqrtf(v) / bc2s + eps;
    e.param[i] = p - step_size * (m / denom);
  }
}

// =========================================================================================
// One-shot all-reduce of the gradient bucket over peer-mapped (symmetric) memory
// =========================================================================================
constexpr int kMaxPeers = 16;
struct PeerPtrs { const float* p[kMaxPeers]; };

__global__ void __launch_bounds__(256) bucket_allreduce_kernel(const float* const* __restrict__ peers_dev, int world,
                                                               float* __restrict__ out, int64_t n4, int64_t n, float scale) {
  __shared__ const float* sp[kMaxPeers];
  if (threadIdx.x < world) sp[threadIdx.x] = peers_dev[threadIdx.x];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int r = 0; r < world; ++r) {
      // peer memory is written by other GPUs: volatile-strength loads that bypass L1 (the barrier before the launch orders them)
      float4 v;
      asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(reinterpret_cast<const float4*>(sp[r]) + i));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
  }
  // tail (n % 4 floats)
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) {
      float v;
      asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(v) : "l"(sp[r] + i));
      a += v;
    }
    out[i] = a * scale;
  }
}

// NVLS: the switch adds the W copies (multimem.ld_reduce on the multicast mapping of the symmetric allocation)
__global__ void __launch_bounds__(256) bucket_allreduce_mc_kernel(const float* __restrict__ mc, float* __restrict__ out,
                                                                  int64_t n4, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(reinterpret_cast<const float4*>(mc) + i) : "memory");
    reinterpret_cast<float4*>(out)[i] = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}

// two-shot NVLS: reduce the rank's slice in the switch, broadcast it to every rank's output copy
__global__ void __launch_bounds__(256) bucket_allreduce_nvls_kernel(const float* __restrict__ mc_in, float* __restrict__ mc_out,
                                                                    int64_t lo4, int64_t hi4, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(reinterpret_cast<const float4*>(mc_in) + i) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(reinterpret_cast<float4*>(mc_out) + i),
                 "f"(v.x * scale), "f"(v.y * scale), "f"(v.z * scale), "f"(v.w * scale) : "memory");
  }
}

// GradScaler variant: scale / found_inf are device scalars, bias corrections from the device-side count of applied steps
struct AdamSteps { int32_t t[STGCMA_ADAM_MAX_GROUPS]; };
__global__ void __launch_bounds__(256) adam_amp_kernel(const StgcmaAdamEntry* __restrict__ table, const StgcmaAdamGroups hp,
                                                       const AdamSteps steps, const float* __restrict__ scale,
                                                       const float* __restrict__ found_inf, int32_t* __restrict__ skipped) {
  // every block reads *skipped BEFORE the barrier-free early exit below can change it: only the found_inf path writes it,
  // and on that path nobody uses the value
  const int skipped_now = *skipped;
  if (*found_inf != 0.f) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *skipped = skipped_now + 1;
    return;
  }
  const StgcmaAdamEntry e = table[blockIdx.x];
  const int gi = e.group;
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double t = (double)(steps.t[gi] - skipped_now);
    s_bc[0] = (float)(1.0 - pow((double)hp.beta1[gi], t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)hp.beta2[gi], t));
  }
  __syncthreads();
  const float lr = hp.lr[gi], wd = hp.weight_decay[gi], b2 = hp.beta2[gi], eps = hp.eps[gi];
  const float step_size = lr / s_bc[0], bc2s = s_bc[1];
  const float omb1 = hp.one_minus_beta1[gi], omb2 = hp.one_minus_beta2[gi];
  const float inv = 1.0f / *scale;
  float* grad_rw = const_cast<float*>(e.grad);
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < e.numel; i += (int64_t)gridDim.y * blockDim.x) {
    const float p = e.param[i];
    float g = e.grad[i] * inv;
    grad_rw[i] = g;                                       // torch's fused Adam leaves the un-scaled gradient behind
    g = fmaf(wd, p, g);
    const float m = fmaf(omb1, g - e.exp_avg[i], e.exp_avg[i]);
    const float v = fmaf(omb2 * g, g, b2 * e.exp_avg_sq[i]);
    e.exp_avg[i] = m;
    e.exp_avg_sq[i] = v;
    const float denom = sqrtf(v) / bc2s + eps;
    e.param[i] = p - step_size * (m / denom);
  }
}

}  // namespace stgcma

using namespace stgcma;

extern "C" int stgcma_bucket_allreduce(const float* const* peers_dev, int world, const float* mc, float* out, int64_t n,
                                       float scale, void* stream) {
  STG_REQUIRE(out && n > 0 && world >= 1 && world <= kMaxPeers, "bucket_allreduce: bad arguments (world=%d)", world);
  STG_REQUIRE(mc || peers_dev, "bucket_allreduce: no source");
  STG_REQUIRE(((uintptr_t)out & 15) == 0, "bucket_allreduce: out must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (mc) {
    STG_REQUIRE(n % 4 == 0 && ((uintptr_t)mc & 15) == 0, "bucket_allreduce: the multicast path needs n %% 4 == 0");
    bucket_allreduce_mc_kernel<<<(unsigned)blocks, 256, 0, st>>>(mc, out, n4, scale);
  } else {
    bucket_allreduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(peers_dev, world, out, n4, n, scale);
  }
  return check_launch("bucket_allreduce");
}

extern "C" int stgcma_bucket_allreduce_nvls(const float* mc_in, float* mc_out, int64_t n, int rank, int world, float scale,
                                            void* stream) {
  STG_REQUIRE(mc_in && mc_out && n > 0 && world >= 1 && rank >= 0 && rank < world, "bucket_allreduce_nvls: bad arguments");
  STG_REQUIRE(n % (4 * (int64_t)world) == 0, "bucket_allreduce_nvls: n=%lld must be a multiple of 4 * world", (long long)n);
  STG_REQUIRE((((uintptr_t)mc_in | (uintptr_t)mc_out) & 15) == 0, "bucket_allreduce_nvls: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per4 = n / 4 / world, lo4 = per4 * rank, hi4 = lo4 + per4;
  int64_t blocks = (per4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  bucket_allreduce_nvls_kernel<<<(unsigned)blocks, 256, 0, st>>>(mc_in, mc_out, lo4, hi4, scale);
  return check_launch("bucket_allreduce_nvls");
}

extern "C" int stgcma_adam_step_amp(const StgcmaAdamEntry* table_dev, int n, const StgcmaAdamGroups* groups_host,
                                    const int32_t* steps_host, const float* scale_dev, const float* found_inf_dev,
                                    int32_t* skipped_dev, void* stream) {
  STG_REQUIRE(n > 0 && table_dev && groups_host && steps_host, "adam_step_amp: empty table");
  STG_REQUIRE(scale_dev && found_inf_dev && skipped_dev, "adam_step_amp: scale / found_inf / skipped must be device pointers");
  cudaStream_t st = (cudaStream_t)stream;
  AdamSteps steps;
  for (int i = 0; i < STGCMA_ADAM_MAX_GROUPS; ++i) steps.t[i] = steps_host[i];
  dim3 grid(n, 16);
  adam_amp_kernel<<<grid, 256, 0, st>>>(table_dev, *groups_host, steps, scale_dev, found_inf_dev, skipped_dev);
  return check_launch("adam_step_amp");
}

extern "C" int stgcma_adam_step(const StgcmaAdamEntry* table_dev, int n, const StgcmaAdamGroups* groups_host, float grad_scale,
                                void* stream) {
  STG_REQUIRE(n > 0 && table_dev && groups_host, "adam_step: empty table");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n, 16);
  adam_kernel<<<grid, 256, 0, st>>>(table_dev, *groups_host, grad_scale);
  return check_launch("adam_step");
}

extern "C" int stgcma_pack_params(const StgcmaPackEntry* table_dev, int n, int dtype, void* stream) {
  STG_REQUIRE(n > 0 && table_dev, "pack_params: empty table");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n, 8);
  if (dtype == STGCMA_F32) pack_params_kernel<float><<<grid, 256, 0, st>>>(table_dev);
  else if (dtype == STGCMA_BF16) pack_params_kernel<bf16><<<grid, 256, 0, st>>>(table_dev);
  else STG_REQUIRE(false, "pack_params: bad dtype %d", dtype);
  return check_launch("pack_params");
}

extern "C" int stgcma_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                    float* rstd, int64_t rows, int C, int merge, int R, int xdtype, int dtype, void* stream) {
  STG_REQUIRE(rows > 0 && C > 0, "layernorm_fwd: empty input");
  STG_REQUIRE(C % 8 == 0 && (!merge || (C % 32 == 0 && R % 2 == 0)), "layernorm_fwd: C=%d must be a multiple of 8 (32 when merging)", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (xdtype == STGCMA_F32 && dtype == STGCMA_F32) return ln_fwd_launch<float, float>(x, gamma, beta, y, mean, rstd, rows, C, merge, R, st);
  if (xdtype == STGCMA_F32 && dtype == STGCMA_BF16) return ln_fwd_launch<float, bf16>(x, gamma, beta, y, mean, rstd, rows, C, merge, R, st);
  if (xdtype == STGCMA_BF16 && dtype == STGCMA_BF16) return ln_fwd_launch<bf16, bf16>(x, gamma, beta, y, mean, rstd, rows, C, merge, R, st);
  STG_REQUIRE(false, "layernorm_fwd: unsupported dtypes x=%d y=%d", xdtype, dtype);
}

extern "C" int stgcma_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                                    const float* gamma, const void* dres, void* dx, void* dx_copy, int64_t rows, int C,
                                    int merge, int R, int dydtype, int xdtype, int dtype, void* stream) {
  STG_REQUIRE(rows > 0 && C > 0, "layernorm_bwd: empty input");
  STG_REQUIRE(C % 8 == 0 && (!merge || (C % 32 == 0 && R % 2 == 0)), "layernorm_bwd: C=%d must be a multiple of 8 (32 when merging)", C);
  STG_REQUIRE(!(merge && dres), "layernorm_bwd: dres is not supported together with merge");
  STG_REQUIRE(mean && rstd, "layernorm_bwd: saved statistics required");
  cudaStream_t st = (cudaStream_t)stream;
#define LN_BWD_CASE(DY, X, CP, TD, TX, TC)                              \
  if (dydtype == DY && xdtype == X && dtype == CP)                      \
    return ln_bwd_launch<TD, TX, TC>(dy, x, mean, rstd, gamma, dres, dx, dx_copy, rows, C, merge, R, st);
  LN_BWD_CASE(STGCMA_F32, STGCMA_F32, STGCMA_F32, float, float, float)
  LN_BWD_CASE(STGCMA_F32, STGCMA_F32, STGCMA_BF16, float, float, bf16)
  LN_BWD_CASE(STGCMA_BF16, STGCMA_F32, STGCMA_BF16, bf16, float, bf16)
  LN_BWD_CASE(STGCMA_BF16, STGCMA_BF16, STGCMA_BF16, bf16, bf16, bf16)
#undef LN_BWD_CASE
  STG_REQUIRE(false, "layernorm_bwd: unsupported dtypes dy=%d x=%d copy=%d", dydtype, xdtype, dtype);
}

extern "C" int stgcma_patch_embed_fwd(const float* img, const float* w, const float* bias, const float* gamma,
                                      const float* beta, void* out, int B, int Cin, int T, int H, int W, int C,
                                      int dtype, void* stream) {
  STG_REQUIRE(B > 0 && T > 0, "patch_embed_fwd: empty batch");
  STG_REQUIRE(H % 4 == 0 && W % 4 == 0, "patch_embed_fwd: H, W must be multiples of the 4x4 patch");
  STG_REQUIRE(C % 32 == 0 && C <= 256, "patch_embed_fwd: C=%d must be a multiple of 32 and <= 256", C);
  STG_REQUIRE(Cin >= 1 && Cin <= 4, "patch_embed_fwd: Cin=%d unsupported", Cin);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(float) * ((size_t)Cin * 16 * C + 3 * C + (size_t)2 * Cin * 4 * W);   // weights + 2 strip buffers
  const int64_t strips = (int64_t)B * T * (H / 4);
  const int grid = (int)(strips < (int64_t)sm_count() * 4 ? strips : (int64_t)sm_count() * 4);
  STG_REQUIRE(W % 4 == 0 && ((uintptr_t)img & 15) == 0, "patch_embed_fwd: img must be 16-byte aligned");
#define STG_PE_LAUNCH(T_, PER_)                                                                                          \
  do {                                                                                                                   \
    auto kern = patch_embed_kernel<T_, PER_>;                                                                            \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
    if (e != cudaSuccess) { set_error("patch_embed_fwd: smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }   \
    kern<<<grid, 256, smem, st>>>(img, w, bias, gamma, beta, (T_*)out, B, Cin, T, H, W, C);                               \
  } while (0)
#define STG_PE_DISPATCH(T_)                                      \
  switch (C >> 5) {                                              \
    case 1: STG_PE_LAUNCH(T_, 1); break;                         \
    case 2: STG_PE_LAUNCH(T_, 2); break;                         \
    case 3: STG_PE_LAUNCH(T_, 3); break;                         \
    case 4: STG_PE_LAUNCH(T_, 4); break;                         \
    case 5: STG_PE_LAUNCH(T_, 5); break;                         \
    case 6: STG_PE_LAUNCH(T_, 6); break;                         \
    case 7: STG_PE_LAUNCH(T_, 7); break;                         \
    default: STG_PE_LAUNCH(T_, 8); break;                        \
  }
  if (dtype == STGCMA_F32) {
    STG_PE_DISPATCH(float);
  } else if (dtype == STGCMA_BF16) {
    STG_PE_DISPATCH(bf16);
  } else {
    STG_REQUIRE(false, "patch_embed_fwd: bad dtype %d", dtype);
  }
#undef STG_PE_DISPATCH
#undef STG_PE_LAUNCH
  return check_launch("patch_embed_fwd");
}

extern "C" int stgcma_head_pool_fwd(const void* xn, void* feat, int64_t frames, int N, int C, int dtype, void* stream) {
  STG_REQUIRE(frames > 0 && N > 0 && C > 0, "head_pool_fwd: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)frames, 2);
  if (dtype == STGCMA_F32) head_pool_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)xn, (float*)feat, frames, N, C);
  else if (dtype == STGCMA_BF16) head_pool_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)xn, (bf16*)feat, frames, N, C);
  else STG_REQUIRE(false, "head_pool_fwd: bad dtype %d", dtype);
  return check_launch("head_pool_fwd");
}

extern "C" int stgcma_head_pool_bwd(const void* dfeat, void* dxn, int64_t frames, int N, int C, int dtype, void* stream) {
  STG_REQUIRE(frames > 0 && N > 0 && C > 0, "head_pool_bwd: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)frames, 2);
  if (dtype == STGCMA_F32) head_pool_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dfeat, (float*)dxn, frames, N, C);
  else if (dtype == STGCMA_BF16) head_pool_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dfeat, (bf16*)dxn, frames, N, C);
  else STG_REQUIRE(false, "head_pool_bwd: bad dtype %d", dtype);
  return check_launch("head_pool_bwd");
}
