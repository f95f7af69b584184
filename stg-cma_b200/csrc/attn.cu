// Swin window attention and temporal attention of the STG-CMA block, forward and backward
// (frozen qkv/proj weights: the backward only produces d(qkv) and, for the temporal variant,
// the gradient of the trainable temporal_position_bias_table).
//
// Window partition / cyclic shift / window reverse are pure addressing here: a block gathers
// the 49 tokens of its (shifted) window straight from the token-major qkv tensor and scatters
// its result back to the tokens' own rows.  The 0/-100 shift mask is computed from coordinates.
// head_dim is 32 for every Swin configuration, so one lane <-> one head channel.
#include "common.cuh"

namespace stgcma {

constexpr int HD = 32;
constexpr int WN_MAX = 49;   // 7 x 7 window
constexpr int T_MAX = 16;

// token row (within the frame) of window slot i
__device__ __forceinline__ int win_token(int wy, int wx, int i, int ws, int shift, int R) {
  const int iy = i / ws, ix = i - iy * ws;
  return win_src(wy, iy, ws, shift, R) * R + win_src(wx, ix, ws, shift, R);
}

// S = scale*q.k^T + rpb + mask, softmax over j.  q,k in smem [N][33]; result P in smem [N][N+1].
__device__ __forceinline__ void window_scores_softmax(const float (*q)[HD + 1], const float (*k)[HD + 1], float* P, int N, int ws,
                                                      int shift, int R, int wy, int wx, int heads, int h,
                                                      const float* __restrict__ rpb) {
  const float scale = 0.17677669529663689f;  // 32^-0.5
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int e = tid; e < N * N; e += nthr) {
    const int i = e / N, j = e - i * N;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(q[i][d], k[j][d], s);
    s *= scale;
    const int yi = i / ws, xi = i - yi * ws, yj = j / ws, xj = j - yj * ws;
    s += __ldg(rpb + ((yi - yj + ws - 1) * (2 * ws - 1) + (xi - xj + ws - 1)) * heads + h);
    if (shift > 0) {
      const int ri = 3 * win_region1(wy * ws + yi, ws, shift, R) + win_region1(wx * ws + xi, ws, shift, R);
      const int rj = 3 * win_region1(wy * ws + yj, ws, shift, R) + win_region1(wx * ws + xj, ws, shift, R);
      if (ri != rj) s += -100.0f;
    }
    P[i * (N + 1) + j] = s;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  for (int i = warp; i < N; i += nwarp) {
    float* row = P + i * (N + 1);
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = __expf(row[j] - m);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < N; j += 32) row[j] *= inv;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(128) window_attn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ rpb,
                                                              T* __restrict__ out, int R, int ws, int shift, int heads) {
  __shared__ float q[WN_MAX][HD + 1], k[WN_MAX][HD + 1], v[WN_MAX][HD + 1];
  __shared__ float P[WN_MAX * (WN_MAX + 1)];
  __shared__ int tok[WN_MAX];
  const int nwr = R / ws, nW = nwr * nwr, N = ws * ws, C = heads * HD, L = R * R;
  int64_t b = blockIdx.x;
  const int h = (int)(b % heads); b /= heads;
  const int w = (int)(b % nW);
  const int64_t frame = b / nW;
  const int wy = w / nwr, wx = w - wy * nwr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < N) tok[tid] = win_token(wy, wx, tid, ws, shift, R);
  __syncthreads();
  for (int i = warp; i < N; i += 4) {
    const T* src = qkv + (frame * L + tok[i]) * (int64_t)(3 * C) + h * HD + lane;
    q[i][lane] = ld_f(src);
    k[i][lane] = ld_f(src + C);
    v[i][lane] = ld_f(src + 2 * C);
  }
  __syncthreads();
  window_scores_softmax(q, k, P, N, ws, shift, R, wy, wx, heads, h, rpb);
  for (int i = warp; i < N; i += 4) {
    const float* prow = P + i * (N + 1);
    float o = 0.f;
    for (int j = 0; j < N; ++j) o = fmaf(prow[j], v[j][lane], o);
    st_f(out + (frame * L + tok[i]) * (int64_t)C + h * HD + lane, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) window_attn_bwd_kernel(const T* __restrict__ qkv, const float* __restrict__ rpb,
                                                              const T* __restrict__ dout, T* __restrict__ dqkv, int R, int ws,
                                                              int shift, int heads) {
  __shared__ float q[WN_MAX][HD + 1], k[WN_MAX][HD + 1], v[WN_MAX][HD + 1], dO[WN_MAX][HD + 1];
  __shared__ float P[WN_MAX * (WN_MAX + 1)], dS[WN_MAX * (WN_MAX + 1)];
  __shared__ int tok[WN_MAX];
  const float scale = 0.17677669529663689f;
  const int nwr = R / ws, nW = nwr * nwr, N = ws * ws, C = heads * HD, L = R * R;
  int64_t b = blockIdx.x;
  const int h = (int)(b % heads); b /= heads;
  const int w = (int)(b % nW);
  const int64_t frame = b / nW;
  const int wy = w / nwr, wx = w - wy * nwr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < N) tok[tid] = win_token(wy, wx, tid, ws, shift, R);
  __syncthreads();
  for (int i = warp; i < N; i += 4) {
    const int64_t row = frame * L + tok[i];
    const T* src = qkv + row * (int64_t)(3 * C) + h * HD + lane;
    q[i][lane] = ld_f(src);
    k[i][lane] = ld_f(src + C);
    v[i][lane] = ld_f(src + 2 * C);
    dO[i][lane] = ld_f(dout + row * (int64_t)C + h * HD + lane);
  }
  __syncthreads();
  window_scores_softmax(q, k, P, N, ws, shift, R, wy, wx, heads, h, rpb);
  // dP = dO . v^T
  for (int e = tid; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(dO[i][d], v[j][d], s);
    dS[i * (N + 1) + j] = s;
  }
  __syncthreads();
  // dS = P * (dP - rowsum(P * dP))
  for (int i = warp; i < N; i += 4) {
    float* drow = dS + i * (N + 1);
    const float* prow = P + i * (N + 1);
    float dl = 0.f;
    for (int j = lane; j < N; j += 32) dl += prow[j] * drow[j];
    dl = warp_sum(dl);
    for (int j = lane; j < N; j += 32) drow[j] = prow[j] * (drow[j] - dl);
  }
  __syncthreads();
  for (int i = warp; i < N; i += 4) {
    // row i plays "query i" for dQ and "key i" for dK / dV
    float dq = 0.f, dk = 0.f, dv = 0.f;
    const float* dsrow = dS + i * (N + 1);
    for (int j = 0; j < N; ++j) {
      dq = fmaf(dsrow[j], k[j][lane], dq);
      dk = fmaf(dS[j * (N + 1) + i], q[j][lane], dk);
      dv = fmaf(P[j * (N + 1) + i], dO[j][lane], dv);
    }
    T* dst = dqkv + (frame * L + tok[i]) * (int64_t)(3 * C) + h * HD + lane;
    st_f(dst, dq * scale);
    st_f(dst + C, dk * scale);
    st_f(dst + 2 * C, dv);
  }
}

// =========================================================================================
// Temporal attention: one warp per (sequence, head); a block = 4 sequences of one (modality, head)
// so that the dBias reduction is block-local.
// =========================================================================================
struct TemporalSmem {
  float q[T_MAX][HD + 1], k[T_MAX][HD + 1], v[T_MAX][HD + 1];
  float P[T_MAX][T_MAX + 1];
};
struct TemporalSmemBwd {
  float dO[T_MAX][HD + 1];
  float dS[T_MAX][T_MAX + 1];
};

__device__ __forceinline__ void temporal_scores_softmax(TemporalSmem& s, int T, const float* __restrict__ tbl_gh, int heads,
                                                        int lane) {
  const float scale = 0.17677669529663689f;
  for (int e = lane; e < T * T; e += 32) {
    const int i = e / T, j = e - i * T;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) a = fmaf(s.q[i][d], s.k[j][d], a);
    s.P[i][j] = a * scale + __ldg(tbl_gh + (i - j + T - 1) * heads);
  }
  __syncwarp();
  if (lane < T) {
    float m = -INFINITY;
    for (int j = 0; j < T; ++j) m = fmaxf(m, s.P[lane][j]);
    float sum = 0.f;
    for (int j = 0; j < T; ++j) {
      const float e = __expf(s.P[lane][j] - m);
      s.P[lane][j] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int j = 0; j < T; ++j) s.P[lane][j] *= inv;
  }
  __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(128) temporal_attn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ tbl,
                                                                T* __restrict__ out, int B, int Tn, int L, int heads) {
  __shared__ TemporalSmem sm[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y / heads, h = blockIdx.y - g * heads;
  const int64_t seq = (int64_t)blockIdx.x * 4 + warp;  // (b, n)
  if (seq >= (int64_t)B * L) return;
  const int64_t b = seq / L;
  const int n = (int)(seq - b * L);
  const int C = heads * HD;
  TemporalSmem& s = sm[warp];
  const int64_t row0 = ((int64_t)g * B + b) * Tn * L + n;  // row of t = 0; rows of t are L apart
  for (int t = 0; t < Tn; ++t) {
    const T* src = qkv + (row0 + (int64_t)t * L) * (3 * C) + h * HD + lane;
    s.q[t][lane] = ld_f(src);
    s.k[t][lane] = ld_f(src + C);
    s.v[t][lane] = ld_f(src + 2 * C);
  }
  __syncwarp();
  temporal_scores_softmax(s, Tn, tbl + ((int64_t)g * (2 * Tn - 1)) * heads + h, heads, lane);
  for (int i = 0; i < Tn; ++i) {
    float o = 0.f;
    for (int j = 0; j < Tn; ++j) o = fmaf(s.P[i][j], s.v[j][lane], o);
    st_f(out + (row0 + (int64_t)i * L) * C + h * HD + lane, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) temporal_attn_bwd_kernel(const T* __restrict__ qkv, const float* __restrict__ tbl,
                                                                const T* __restrict__ dout, T* __restrict__ dqkv,
                                                                float* __restrict__ dtbl, int B, int Tn, int L, int heads) {
  __shared__ TemporalSmem sm[4];
  __shared__ TemporalSmemBwd sb[4];
  __shared__ float dbias[2 * T_MAX];
  const float scale = 0.17677669529663689f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y / heads, h = blockIdx.y - g * heads;
  if (threadIdx.x < 2 * T_MAX) dbias[threadIdx.x] = 0.f;
  __syncthreads();
  const int64_t seq = (int64_t)blockIdx.x * 4 + warp;
  const bool active = seq < (int64_t)B * L;
  if (active) {
    const int64_t b = seq / L;
    const int n = (int)(seq - b * L);
    const int C = heads * HD;
    TemporalSmem& s = sm[warp];
    TemporalSmemBwd& d = sb[warp];
    const int64_t row0 = ((int64_t)g * B + b) * Tn * L + n;
    for (int t = 0; t < Tn; ++t) {
      const int64_t row = row0 + (int64_t)t * L;
      const T* src = qkv + row * (3 * C) + h * HD + lane;
      s.q[t][lane] = ld_f(src);
      s.k[t][lane] = ld_f(src + C);
      s.v[t][lane] = ld_f(src + 2 * C);
      d.dO[t][lane] = ld_f(dout + row * C + h * HD + lane);
    }
    __syncwarp();
    temporal_scores_softmax(s, Tn, tbl + ((int64_t)g * (2 * Tn - 1)) * heads + h, heads, lane);
    for (int e = lane; e < Tn * Tn; e += 32) {
      const int i = e / Tn, j = e - i * Tn;
      float a = 0.f;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) a = fmaf(d.dO[i][dd], s.v[j][dd], a);
      d.dS[i][j] = a;
    }
    __syncwarp();
    if (lane < Tn) {
      float dl = 0.f;
      for (int j = 0; j < Tn; ++j) dl += s.P[lane][j] * d.dS[lane][j];
      for (int j = 0; j < Tn; ++j) {
        const float ds = s.P[lane][j] * (d.dS[lane][j] - dl);
        d.dS[lane][j] = ds;
        atomicAdd(&dbias[lane - j + Tn - 1], ds);  // bias enters the logits un-scaled
      }
    }
    __syncwarp();
    for (int i = 0; i < Tn; ++i) {
      float dq = 0.f, dk = 0.f, dv = 0.f;
      for (int j = 0; j < Tn; ++j) {
        dq = fmaf(d.dS[i][j], s.k[j][lane], dq);
        dk = fmaf(d.dS[j][i], s.q[j][lane], dk);
        dv = fmaf(s.P[j][i], d.dO[j][lane], dv);
      }
      T* dst = dqkv + (row0 + (int64_t)i * L) * (3 * C) + h * HD + lane;
      st_f(dst, dq * scale);
      st_f(dst + C, dk * scale);
      st_f(dst + 2 * C, dv);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * Tn - 1) atomicAdd(dtbl + ((int64_t)g * (2 * Tn - 1) + threadIdx.x) * heads + h, dbias[threadIdx.x]);
}

// bf16 production kernels (tensor cores), temporal_tc.cu; -1 = shape not covered
int temporal_attn_fwd_bf16(const void* qkv, const float* tbl, void* out, int G, int B, int T, int L, int heads,
                           cudaStream_t st);
int temporal_attn_bwd_bf16(const void* qkv, const float* tbl, const void* dout, void* dqkv, float* dtbl, int G, int B, int T,
                           int L, int heads, cudaStream_t st);
// bf16 production kernels (tensor cores), attn_tc.cu
int window_attn_fwd_bf16(const void* qkv, const float* rpb, void* out, int64_t frames, int R, int ws, int shift, int heads,
                         cudaStream_t st);
int window_attn_bwd_bf16(const void* qkv, const float* rpb, const void* dout, void* dqkv, int64_t frames, int R, int ws,
                         int shift, int heads, cudaStream_t st);

}  // namespace stgcma

using namespace stgcma;

static int check_window_args(const char* what, int64_t frames, int R, int ws, int shift, int heads) {
  STG_REQUIRE(frames > 0, "%s: empty input", what);
  STG_REQUIRE(ws >= 1 && ws * ws <= WN_MAX && R % ws == 0, "%s: window %d on a %d x %d map unsupported (max 7x7)", what, ws, R, R);
  STG_REQUIRE(shift >= 0 && shift < ws, "%s: shift %d must be in [0, ws)", what, shift);
  STG_REQUIRE(heads >= 1, "%s: heads", what);
  const int64_t blocks = frames * (R / ws) * (R / ws) * heads;
  STG_REQUIRE(blocks < (1ll << 31), "%s: grid too large", what);
  return STGCMA_OK;
}

extern "C" int stgcma_window_attn_fwd(const void* qkv, const float* rpb, void* out, int64_t frames, int R, int ws, int shift,
                                      int heads, int dtype, void* stream) {
  int rc = check_window_args("window_attn_fwd", frames, R, ws, shift, heads);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws) * heads);
  if (dtype == STGCMA_F32)
    window_attn_fwd_kernel<float><<<grid, 128, 0, st>>>((const float*)qkv, rpb, (float*)out, R, ws, shift, heads);
  else if (dtype == STGCMA_BF16)
    return window_attn_fwd_bf16(qkv, rpb, out, frames, R, ws, shift, heads, st);
  else
    STG_REQUIRE(false, "window_attn_fwd: bad dtype %d", dtype);
  return check_launch("window_attn_fwd");
}

extern "C" int stgcma_window_attn_bwd(const void* qkv, const float* rpb, const void* dout, void* dqkv, int64_t frames, int R,
                                      int ws, int shift, int heads, int dtype, void* stream) {
  int rc = check_window_args("window_attn_bwd", frames, R, ws, shift, heads);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws) * heads);
  if (dtype == STGCMA_F32)
    window_attn_bwd_kernel<float><<<grid, 128, 0, st>>>((const float*)qkv, rpb, (const float*)dout, (float*)dqkv, R, ws, shift, heads);
  else if (dtype == STGCMA_BF16)
    return window_attn_bwd_bf16(qkv, rpb, dout, dqkv, frames, R, ws, shift, heads, st);
  else
    STG_REQUIRE(false, "window_attn_bwd: bad dtype %d", dtype);
  return check_launch("window_attn_bwd");
}

extern "C" int stgcma_temporal_attn_fwd(const void* qkv, const float* tbl, void* out, int G, int B, int T, int L, int heads,
                                        int dtype, void* stream) {
  STG_REQUIRE(G >= 1 && B > 0 && L > 0 && heads >= 1, "temporal_attn_fwd: empty input");
  STG_REQUIRE(T >= 1 && T <= T_MAX, "temporal_attn_fwd: T=%d unsupported (max %d)", T, T_MAX);
  STG_REQUIRE(G * heads <= 65535, "temporal_attn_fwd: G*heads too large");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div((int64_t)B * L, 4), G * heads);
  if (dtype == STGCMA_F32)
    temporal_attn_fwd_kernel<float><<<grid, 128, 0, st>>>((const float*)qkv, tbl, (float*)out, B, T, L, heads);
  else if (dtype == STGCMA_BF16) {
    const int rc = temporal_attn_fwd_bf16(qkv, tbl, out, G, B, T, L, heads, st);   // tensor-core kernel
    if (rc != -1) return rc;
    temporal_attn_fwd_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)qkv, tbl, (bf16*)out, B, T, L, heads);
  }
  else
    STG_REQUIRE(false, "temporal_attn_fwd: bad dtype %d", dtype);
  return check_launch("temporal_attn_fwd");
}

extern "C" int stgcma_temporal_attn_bwd(const void* qkv, const float* tbl, const void* dout, void* dqkv, float* dtbl, int G,
                                        int B, int T, int L, int heads, int dtype, void* stream) {
  STG_REQUIRE(G >= 1 && B > 0 && L > 0 && heads >= 1, "temporal_attn_bwd: empty input");
  STG_REQUIRE(T >= 1 && T <= T_MAX, "temporal_attn_bwd: T=%d unsupported (max %d)", T, T_MAX);
  STG_REQUIRE(G * heads <= 65535, "temporal_attn_bwd: G*heads too large");
  STG_REQUIRE(dtbl, "temporal_attn_bwd: dtbl required (the bias tables are trainable)");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div((int64_t)B * L, 4), G * heads);
  if (dtype == STGCMA_F32)
    temporal_attn_bwd_kernel<float><<<grid, 128, 0, st>>>((const float*)qkv, tbl, (const float*)dout, (float*)dqkv, dtbl, B, T, L, heads);
  else if (dtype == STGCMA_BF16) {
    const int rc = temporal_attn_bwd_bf16(qkv, tbl, dout, dqkv, dtbl, G, B, T, L, heads, st);
    if (rc != -1) return rc;
    temporal_attn_bwd_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)qkv, tbl, (const bf16*)dout, (bf16*)dqkv, dtbl, B, T, L, heads);
  }
  else
    STG_REQUIRE(false, "temporal_attn_bwd: bad dtype %d", dtype);
  return check_launch("temporal_attn_bwd");
}
