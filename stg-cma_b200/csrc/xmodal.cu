// Cross-modal audio<->video token attention of the STG-CMA adapters (Swin_AVE.py:753-760,
// :801-808), forward and backward, CUDA-core implementation (fp32 math, any storage type).
//
//   hv' = hv + gate_v * softmax(hv ha^T) ha         ha' = ha + gate_a * softmax(ha hv^T) hv
//
// No projections, no scale; both directions read the un-updated features, so ONE score matrix
// S = hv ha^T serves both (row softmax for video queries, column softmax for audio queries).
//   window variant: token groups = the (shifted) ws x ws windows; a block owns one (frame, window)
//                   and keeps S, both probability matrices and all operands in shared memory.
//   global variant: token group = all L tokens of a frame (L = 3136 at stage 0).  Flash-style:
//                   a block owns 64 rows of one modality and streams the other modality's rows in
//                   tiles of 64 with an online softmax; the L x L scores never exist in memory.
//                   The backward gives each block 64 rows of one modality and makes it compute
//                   both roles of those rows (queries of their own direction, keys/values of the
//                   other direction) from the same S tile, so no atomics are needed on dh.
#include "common.cuh"

namespace stgcma {

constexpr int XW_MAX = 49;

__device__ __forceinline__ int xwin_token(int wy, int wx, int i, int ws, int shift, int R) {
  const int iy = i / ws, ix = i - iy * ws;
  return win_src(wy, iy, ws, shift, R) * R + win_src(wx, ix, ws, shift, R);
}

// ---------------------------------------------------------------------------------------------
// window variant
// ---------------------------------------------------------------------------------------------
// dynamic smem layout (floats): hv[N][rp+1] ha[N][rp+1] Pv[N][N+1] Pa[N][N+1] (+bwd: gv[N][rp+1]
// ga[N][rp+1] dS[N][N+1]) tok[N]
__device__ __forceinline__ void xwin_scores(const float* hv, const float* ha, float* Pv, float* Pa, float* red, int N, int rp) {
  const int tid = threadIdx.x, nthr = blockDim.x, ldh = rp + 1, ldp = N + 1;
  for (int e = tid; e < N * N; e += nthr) {
    const int i = e / N, j = e - i * N;
    float s = 0.f;
    for (int d = 0; d < rp; ++d) s = fmaf(hv[i * ldh + d], ha[j * ldh + d], s);
    Pv[i * ldp + j] = s;
    Pa[i * ldp + j] = s;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  // rows of Pv: softmax over j;  columns of Pa: softmax over i
  for (int r = warp; r < 2 * N; r += nwarp) {
    const bool col = r >= N;
    const int idx = col ? r - N : r;
    float* base = col ? Pa + idx : Pv + idx * ldp;
    const int stride = col ? ldp : 1;
    float m = -INFINITY;
    for (int t = lane; t < N; t += 32) m = fmaxf(m, base[t * stride]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < N; t += 32) {
      const float e = __expf(base[t * stride] - m);
      base[t * stride] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int t = lane; t < N; t += 32) base[t * stride] *= inv;
  }
  __syncthreads();
  (void)red;
}

template <typename T>
__global__ void __launch_bounds__(128) xmodal_window_fwd_kernel(const T* __restrict__ h, const float* __restrict__ gate,
                                                                T* __restrict__ out, T* __restrict__ attn_out, int64_t frames,
                                                                int R, int ws, int shift, int rp) {
  extern __shared__ float xs[];
  const int N = ws * ws, ldh = rp + 1, ldp = N + 1, L = R * R;
  float* hv = xs;
  float* ha = hv + N * ldh;
  float* Pv = ha + N * ldh;
  float* Pa = Pv + N * ldp;
  int* tok = reinterpret_cast<int*>(Pa + N * ldp);
  const int nwr = R / ws, nW = nwr * nwr;
  const int w = (int)(blockIdx.x % nW);
  const int64_t frame = blockIdx.x / nW;
  const int wy = w / nwr, wx = w - wy * nwr;
  const int tid = threadIdx.x;
  if (tid < N) tok[tid] = xwin_token(wy, wx, tid, ws, shift, R);
  __syncthreads();
  const int64_t mod_stride = frames * L * (int64_t)rp;
  for (int e = tid; e < N * rp; e += blockDim.x) {
    const int i = e / rp, d = e - i * rp;
    const int64_t off = (frame * L + tok[i]) * (int64_t)rp + d;
    hv[i * ldh + d] = ld_f(h + off);
    ha[i * ldh + d] = ld_f(h + mod_stride + off);
  }
  __syncthreads();
  xwin_scores(hv, ha, Pv, Pa, nullptr, N, rp);
  const float gv = gate[0], ga = gate[1];
  for (int e = tid; e < N * rp; e += blockDim.x) {
    const int i = e / rp, d = e - i * rp;
    float a2v = 0.f, v2a = 0.f;
    for (int j = 0; j < N; ++j) {
      a2v = fmaf(Pv[i * ldp + j], ha[j * ldh + d], a2v);   // video query i over audio keys j
      v2a = fmaf(Pa[j * ldp + i], hv[j * ldh + d], v2a);   // audio query i over video keys j
    }
    const int64_t off = (frame * L + tok[i]) * (int64_t)rp + d;
    st_f(out + off, hv[i * ldh + d] + gv * a2v);
    st_f(out + mod_stride + off, ha[i * ldh + d] + ga * v2a);
    st_f(attn_out + off, a2v);
    st_f(attn_out + mod_stride + off, v2a);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) xmodal_window_bwd_kernel(const T* __restrict__ h, const T* __restrict__ hpre,
                                                                const float* __restrict__ gate, const T* __restrict__ attn_out,
                                                                const T* __restrict__ dout, T* __restrict__ dh,
                                                                float* __restrict__ dgate, int64_t frames, int R, int ws,
                                                                int shift, int rp) {
  extern __shared__ float xs[];
  const int N = ws * ws, ldh = rp + 1, ldp = N + 1, L = R * R;
  float* hv = xs;
  float* ha = hv + N * ldh;
  float* gov = ha + N * ldh;   // gate_v * dout_v
  float* goa = gov + N * ldh;  // gate_a * dout_a
  float* Pv = goa + N * ldh;
  float* Pa = Pv + N * ldp;
  float* dS = Pa + N * ldp;
  float* red = dS + N * ldp;   // [8]
  int* tok = reinterpret_cast<int*>(red + 8);
  const int nwr = R / ws, nW = nwr * nwr;
  const int w = (int)(blockIdx.x % nW);
  const int64_t frame = blockIdx.x / nW;
  const int wy = w / nwr, wx = w - wy * nwr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < N) tok[tid] = xwin_token(wy, wx, tid, ws, shift, R);
  if (tid < 8) red[tid] = 0.f;
  __syncthreads();
  const int64_t mod_stride = frames * L * (int64_t)rp;
  const float gv = gate[0], ga = gate[1];
  float dgv = 0.f, dga = 0.f;
  for (int e = tid; e < N * rp; e += blockDim.x) {
    const int i = e / rp, d = e - i * rp;
    const int64_t off = (frame * L + tok[i]) * (int64_t)rp + d;
    hv[i * ldh + d] = ld_f(h + off);
    ha[i * ldh + d] = ld_f(h + mod_stride + off);
    const float dv = ld_f(dout + off), da = ld_f(dout + mod_stride + off);
    gov[i * ldh + d] = gv * dv;
    goa[i * ldh + d] = ga * da;
    dgv = fmaf(dv, ld_f(attn_out + off), dgv);
    dga = fmaf(da, ld_f(attn_out + mod_stride + off), dga);
  }
  dgv = warp_sum(dgv);
  dga = warp_sum(dga);
  if (lane == 0) {
    atomicAdd(&red[0], dgv);
    atomicAdd(&red[1], dga);
  }
  __syncthreads();
  xwin_scores(hv, ha, Pv, Pa, nullptr, N, rp);
  // dPv[i][j] = gov_i . ha_j ; dPa[i][j] = goa_j . hv_i  (stored in dS / reuse below)
  // first dSv into dS
  for (int e = tid; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    float a = 0.f;
    for (int d = 0; d < rp; ++d) a = fmaf(gov[i * ldh + d], ha[j * ldh + d], a);
    dS[i * ldp + j] = a;
  }
  __syncthreads();
  for (int i = warp; i < N; i += 4) {
    float dl = 0.f;
    for (int j = lane; j < N; j += 32) dl += Pv[i * ldp + j] * dS[i * ldp + j];
    dl = warp_sum(dl);
    for (int j = lane; j < N; j += 32) dS[i * ldp + j] = Pv[i * ldp + j] * (dS[i * ldp + j] - dl);
  }
  __syncthreads();
  // audio direction: column j of Pa is the softmax of audio query j over video keys i
  for (int j = warp; j < N; j += 4) {
    float dpa[2];  // N <= 49 -> at most 2 rows per lane
    float dl = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = lane + 32 * t;
      dpa[t] = 0.f;
      if (i < N) {
        float a = 0.f;
        for (int d = 0; d < rp; ++d) a = fmaf(goa[j * ldh + d], hv[i * ldh + d], a);
        dpa[t] = a;
        dl += Pa[i * ldp + j] * a;
      }
    }
    dl = warp_sum(dl);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = lane + 32 * t;
      if (i < N) dS[i * ldp + j] += Pa[i * ldp + j] * (dpa[t] - dl);
    }
  }
  __syncthreads();
  for (int e = tid; e < N * rp; e += blockDim.x) {
    const int i = e / rp, d = e - i * rp;
    float gh_v = 0.f, gh_a = 0.f;
    for (int j = 0; j < N; ++j) {
      // video row i: dS[i][j] ha_j  +  Pa[i][j] goa_j   (value role in the audio direction)
      gh_v = fmaf(dS[i * ldp + j], ha[j * ldh + d], gh_v);
      gh_v = fmaf(Pa[i * ldp + j], goa[j * ldh + d], gh_v);
      // audio row i: dS[j][i] hv_j  +  Pv[j][i] gov_j
      gh_a = fmaf(dS[j * ldp + i], hv[j * ldh + d], gh_a);
      gh_a = fmaf(Pv[j * ldp + i], gov[j * ldh + d], gh_a);
    }
    const int64_t off = (frame * L + tok[i]) * (int64_t)rp + d;
    float rv = ld_f(dout + off) + gh_v;
    float ra = ld_f(dout + mod_stride + off) + gh_a;
    if (hpre) {
      rv *= gelu_erf_grad(ld_f(hpre + off));
      ra *= gelu_erf_grad(ld_f(hpre + mod_stride + off));
    }
    st_f(dh + off, rv);
    st_f(dh + mod_stride + off, ra);
  }
  if (tid == 0) {
    atomicAdd(dgate + 0, red[0]);
    atomicAdd(dgate + 1, red[1]);
  }
}

// ---------------------------------------------------------------------------------------------
// global variant (flash-style), CUDA cores
// ---------------------------------------------------------------------------------------------
constexpr int XG_TILE = 64;

// z = 0: own = video rows (queries), other = audio;  z = 1: own = audio, other = video.
template <typename T, int RP>
__global__ void __launch_bounds__(128) xmodal_global_fwd_kernel(const T* __restrict__ h, const float* __restrict__ gate,
                                                                T* __restrict__ out, T* __restrict__ attn_out,
                                                                float* __restrict__ lse, int64_t frames, int L, int La) {
  extern __shared__ float xs[];
  constexpr int LDH = RP + 1, LDP = XG_TILE + 1;
  float* own = xs;
  float* oth = own + XG_TILE * LDH;
  float* P = oth + XG_TILE * LDH;
  const int z = blockIdx.z;
  const int64_t frame = blockIdx.y;
  const int r0 = blockIdx.x * XG_TILE;
  // rows per frame: L (first modality, video) and La (second, audio); the rows of the two modalities are stacked
  const int Lown = z ? La : L, Loth = z ? L : La;
  if (r0 >= Lown) return;
  const int64_t row_own0 = (z ? frames * (int64_t)L : 0) + frame * (int64_t)Lown;
  const int64_t row_oth0 = (z ? 0 : frames * (int64_t)L) + frame * (int64_t)Loth;
  const T* hown = h + row_own0 * RP;
  const T* hoth = h + row_oth0 * RP;
  const int tid = threadIdx.x;
  for (int e = tid; e < XG_TILE * RP; e += 128) {
    const int i = e / RP, d = e - i * RP;
    own[i * LDH + d] = (r0 + i < Lown) ? ld_f(hown + (int64_t)(r0 + i) * RP + d) : 0.f;
  }
  const int row = tid >> 1, half = tid & 1;
  float acc[RP / 2];
#pragma unroll
  for (int d = 0; d < RP / 2; ++d) acc[d] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  for (int c0 = 0; c0 < Loth; c0 += XG_TILE) {
    __syncthreads();
    for (int e = tid; e < XG_TILE * RP; e += 128) {
      const int j = e / RP, d = e - j * RP;
      oth[j * LDH + d] = (c0 + j < Loth) ? ld_f(hoth + (int64_t)(c0 + j) * RP + d) : 0.f;
    }
    __syncthreads();
    float s[XG_TILE / 2];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < XG_TILE / 2; ++jj) {
      const int j = half + 2 * jj;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < RP; ++d) a = fmaf(own[row * LDH + d], oth[j * LDH + d], a);
      a = (c0 + j < Loth) ? a : -INFINITY;
      s[jj] = a;
      mx = fmaxf(mx, a);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    const float m_new = fmaxf(m_run, mx);
    const float corr = __expf(m_run - m_new);  // exp(-inf) = 0 on the first tile
    float psum = 0.f;
#pragma unroll
    for (int jj = 0; jj < XG_TILE / 2; ++jj) {
      const float p = __expf(s[jj] - m_new);
      P[row * LDP + half + 2 * jj] = p;
      psum += p;
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    l_run = l_run * corr + psum;
    m_run = m_new;
    __syncwarp();
#pragma unroll
    for (int d = 0; d < RP / 2; ++d) acc[d] *= corr;
    for (int j = 0; j < XG_TILE; ++j) {
      const float p = P[row * LDP + j];
#pragma unroll
      for (int d = 0; d < RP / 2; ++d) acc[d] = fmaf(p, oth[j * LDH + half + 2 * d], acc[d]);
    }
  }
  if (r0 + row < Lown) {
    const float inv = 1.f / l_run;
    const float g = gate[z];
    const int64_t base = (row_own0 + r0 + row) * RP;
#pragma unroll
    for (int d = 0; d < RP / 2; ++d) {
      const int dd = half + 2 * d;
      const float a = acc[d] * inv;
      st_f(attn_out + base + dd, a);
      st_f(out + base + dd, own[row * LDH + dd] + g * a);
    }
    if (half == 0) lse[row_own0 + r0 + row] = m_run + __logf(l_run);
  }
}

// D[m][row] = gate[m] * <dout_row, attn_row>;  dgate[m] += sum_rows <dout_row, attn_row>
template <typename T>
__global__ void __launch_bounds__(256) xmodal_prep_kernel(const T* __restrict__ dout, const T* __restrict__ attn_out,
                                                          const float* __restrict__ gate, float* __restrict__ D,
                                                          float* __restrict__ dgate, int64_t rows0, int64_t rows1, int rp) {
  __shared__ float red[8];
  const int m = blockIdx.y;
  const int64_t rows_per_mod = m ? rows1 : rows0, row_base = m ? rows0 : 0;   // modality 1 follows the rows0 rows of modality 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float g = gate[m];
  float tot = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows_per_mod; row += (int64_t)gridDim.x * 8) {
    const int64_t base = (row_base + row) * rp;
    float a = 0.f;
    for (int d = lane; d < rp; d += 32) a = fmaf(ld_f(dout + base + d), ld_f(attn_out + base + d), a);
    a = warp_sum(a);
    if (lane == 0) D[row_base + row] = g * a;
    tot += a;
  }
  if (lane == 0) red[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(dgate + m, s);
  }
}

// z = 0: produce dh for video rows; z = 1: for audio rows.
template <typename T, int RP>
__global__ void __launch_bounds__(128) xmodal_global_bwd_kernel(const T* __restrict__ h, const T* __restrict__ hpre,
                                                                const float* __restrict__ gate, const float* __restrict__ lse,
                                                                const float* __restrict__ D, const T* __restrict__ dout,
                                                                T* __restrict__ dh, int64_t frames, int L, int La) {
  extern __shared__ float xs[];
  constexpr int LDH = RP + 1, LDP = XG_TILE + 1;
  float* own = xs;                        // h of the own rows
  float* gown = own + XG_TILE * LDH;      // gate_own * dout of the own rows
  float* oth = gown + XG_TILE * LDH;
  float* goth = oth + XG_TILE * LDH;
  float* Wm = goth + XG_TILE * LDH;       // dS1 + dS2
  float* P2 = Wm + XG_TILE * LDP;
  float* lse_o = P2 + XG_TILE * LDP;      // [64] lse of other rows (as queries of their direction)
  float* D_o = lse_o + XG_TILE;
  const int z = blockIdx.z;
  const int64_t frame = blockIdx.y;
  const int r0 = blockIdx.x * XG_TILE;
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
  const int tid = threadIdx.x;
  for (int e = tid; e < XG_TILE * RP; e += 128) {
    const int i = e / RP, d = e - i * RP;
    const bool ok = r0 + i < Lown;
    own[i * LDH + d] = ok ? ld_f(h + off_own + (int64_t)(r0 + i) * RP + d) : 0.f;
    gown[i * LDH + d] = ok ? g_own * ld_f(dout + off_own + (int64_t)(r0 + i) * RP + d) : 0.f;
  }
  const int row = tid >> 1, half = tid & 1;
  const bool row_ok = r0 + row < Lown;
  const float my_lse = row_ok ? lse_own_p[r0 + row] : 0.f;
  const float my_D = row_ok ? D_own_p[r0 + row] : 0.f;
  float acc[RP / 2];
#pragma unroll
  for (int d = 0; d < RP / 2; ++d) acc[d] = 0.f;
  for (int c0 = 0; c0 < Loth; c0 += XG_TILE) {
    __syncthreads();
    for (int e = tid; e < XG_TILE * RP; e += 128) {
      const int j = e / RP, d = e - j * RP;
      const bool ok = c0 + j < Loth;
      oth[j * LDH + d] = ok ? ld_f(h + off_oth + (int64_t)(c0 + j) * RP + d) : 0.f;
      goth[j * LDH + d] = ok ? g_oth * ld_f(dout + off_oth + (int64_t)(c0 + j) * RP + d) : 0.f;
    }
    if (tid < XG_TILE) {
      const bool ok = c0 + tid < Loth;
      lse_o[tid] = ok ? lse_oth_p[c0 + tid] : 0.f;
      D_o[tid] = ok ? D_oth_p[c0 + tid] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < XG_TILE / 2; ++jj) {
      const int j = half + 2 * jj;
      float s = 0.f, dp1 = 0.f, dp2 = 0.f;
#pragma unroll
      for (int d = 0; d < RP; ++d) {
        const float ho = own[row * LDH + d], hj = oth[j * LDH + d];
        s = fmaf(ho, hj, s);
        dp1 = fmaf(gown[row * LDH + d], hj, dp1);
        dp2 = fmaf(goth[j * LDH + d], ho, dp2);
      }
      float w = 0.f, p2 = 0.f;
      if (row_ok && c0 + j < Loth) {
        const float p1 = __expf(s - my_lse);       // own row as query
        p2 = __expf(s - lse_o[j]);                 // other row j as query, own row as key/value
        w = p1 * (dp1 - my_D) + p2 * (dp2 - D_o[j]);
      }
      Wm[row * LDP + j] = w;
      P2[row * LDP + j] = p2;
    }
    __syncwarp();
    for (int j = 0; j < XG_TILE; ++j) {
      const float w = Wm[row * LDP + j], p2 = P2[row * LDP + j];
#pragma unroll
      for (int d = 0; d < RP / 2; ++d) {
        const int dd = half + 2 * d;
        acc[d] = fmaf(w, oth[j * LDH + dd], acc[d]);
        acc[d] = fmaf(p2, goth[j * LDH + dd], acc[d]);
      }
    }
  }
  if (row_ok) {
    const int64_t base = off_own + (int64_t)(r0 + row) * RP;
#pragma unroll
    for (int d = 0; d < RP / 2; ++d) {
      const int dd = half + 2 * d;
      float r = ld_f(dout + base + dd) + acc[d];
      if (hpre) r *= gelu_erf_grad(ld_f(hpre + base + dd));
      st_f(dh + base + dd, r);
    }
  }
}

template <typename T, int RP>
static int xg_fwd_launch(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames, int L,
                         int La, cudaStream_t st) {
  const size_t smem = sizeof(float) * (2 * XG_TILE * (RP + 1) + XG_TILE * (XG_TILE + 1));
  auto kern = xmodal_global_fwd_kernel<T, RP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("xmodal_global_fwd: smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  dim3 grid((unsigned)ceil_div(L > La ? L : La, XG_TILE), (unsigned)frames, 2);
  kern<<<grid, 128, smem, st>>>((const T*)h, gate, (T*)out, (T*)attn_out, lse, frames, L, La);
  return check_launch("xmodal_global_fwd");
}

template <typename T, int RP>
static int xg_bwd_launch(const void* h, const void* hpre, const float* gate, const float* lse, const float* D, const void* dout,
                         void* dh, int64_t frames, int L, int La, cudaStream_t st) {
  const size_t smem = sizeof(float) * (4 * XG_TILE * (RP + 1) + 2 * XG_TILE * (XG_TILE + 1) + 2 * XG_TILE);
  auto kern = xmodal_global_bwd_kernel<T, RP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("xmodal_global_bwd: smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
  dim3 grid((unsigned)ceil_div(L > La ? L : La, XG_TILE), (unsigned)frames, 2);
  kern<<<grid, 128, smem, st>>>((const T*)h, (const T*)hpre, gate, lse, D, (const T*)dout, (T*)dh, frames, L, La);
  return check_launch("xmodal_global_bwd");
}

#define XG_DISPATCH_RP(rp, CALL)                    \
  switch (rp) {                                     \
    case 16: { constexpr int RP_ = 16; CALL; } break; \
    case 32: { constexpr int RP_ = 32; CALL; } break; \
    case 48: { constexpr int RP_ = 48; CALL; } break; \
    case 64: { constexpr int RP_ = 64; CALL; } break; \
    case 96: { constexpr int RP_ = 96; CALL; } break; \
    default: rc = -1;                               \
  }

// bf16 production kernels (tensor cores), xmodal_tc.cu
int xmodal_global_fwd_bf16(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames, int L, int La,
                           int rp, cudaStream_t st);
int xmodal_global_bwd_bf16(const void* h, const void* hpre, const float* gate, const float* lse, const float* D,
                           const void* dout, void* dh, int64_t frames, int L, int La, int rp, cudaStream_t st);
// xmodal_win_tc.cu
int xmodal_window_fwd_bf16(const void* h, const float* gate, void* out, void* attn_out, int64_t frames, int R, int ws,
                           int shift, int rp, cudaStream_t st);
int xmodal_window_bwd_bf16(const void* h, const void* hpre, const float* gate, const void* attn_out, const void* dout,
                           void* dh, float* dgate, int64_t frames, int R, int ws, int shift, int rp, cudaStream_t st);

}  // namespace stgcma

using namespace stgcma;

static int check_xw(const char* what, int64_t frames, int R, int ws, int shift, int rp) {
  STG_REQUIRE(frames > 0, "%s: empty input", what);
  STG_REQUIRE(ws >= 1 && ws * ws <= XW_MAX && R % ws == 0, "%s: window %d on %d x %d unsupported", what, ws, R, R);
  STG_REQUIRE(shift >= 0 && shift < ws, "%s: bad shift %d", what, shift);
  STG_REQUIRE(rp >= 8 && rp <= 128 && rp % 8 == 0, "%s: rp=%d unsupported", what, rp);
  STG_REQUIRE(frames * (R / ws) * (R / ws) < (1ll << 31), "%s: grid too large", what);
  return STGCMA_OK;
}

extern "C" int stgcma_xmodal_window_fwd(const void* h, const float* gate, void* out, void* attn_out, int64_t frames, int R,
                                        int ws, int shift, int rp, int dtype, void* stream) {
  int rc = check_xw("xmodal_window_fwd", frames, R, ws, shift, rp);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = ws * ws;
  const size_t smem = sizeof(float) * (2 * N * (rp + 1) + 2 * N * (N + 1)) + sizeof(int) * N;
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws));
  cudaError_t e;
  if (dtype == STGCMA_F32) {
    e = cudaFuncSetAttribute(xmodal_window_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("xmodal_window_fwd: smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
    xmodal_window_fwd_kernel<float><<<grid, 128, smem, st>>>((const float*)h, gate, (float*)out, (float*)attn_out, frames, R, ws, shift, rp);
  } else if (dtype == STGCMA_BF16) {
    rc = xmodal_window_fwd_bf16(h, gate, out, attn_out, frames, R, ws, shift, rp, st);   // tensor-core kernel
    if (rc != -1) return rc;
    e = cudaFuncSetAttribute(xmodal_window_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("xmodal_window_fwd: smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
    xmodal_window_fwd_kernel<bf16><<<grid, 128, smem, st>>>((const bf16*)h, gate, (bf16*)out, (bf16*)attn_out, frames, R, ws, shift, rp);
  } else {
    STG_REQUIRE(false, "xmodal_window_fwd: bad dtype %d", dtype);
  }
  return check_launch("xmodal_window_fwd");
}

extern "C" int stgcma_xmodal_window_bwd(const void* h, const void* hpre, const float* gate, const void* attn_out,
                                        const void* dout, void* dh, float* dgate, int64_t frames, int R, int ws, int shift,
                                        int rp, int dtype, void* stream) {
  int rc = check_xw("xmodal_window_bwd", frames, R, ws, shift, rp);
  if (rc) return rc;
  STG_REQUIRE(dgate, "xmodal_window_bwd: dgate required");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = ws * ws;
  const size_t smem = sizeof(float) * (4 * N * (rp + 1) + 3 * N * (N + 1) + 8) + sizeof(int) * N;
  const unsigned grid = (unsigned)(frames * (R / ws) * (R / ws));
  cudaError_t e;
  if (dtype == STGCMA_F32) {
    e = cudaFuncSetAttribute(xmodal_window_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("xmodal_window_bwd: smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
    xmodal_window_bwd_kernel<float><<<grid, 128, smem, st>>>((const float*)h, (const float*)hpre, gate, (const float*)attn_out,
                                                             (const float*)dout, (float*)dh, dgate, frames, R, ws, shift, rp);
  } else if (dtype == STGCMA_BF16) {
    rc = xmodal_window_bwd_bf16(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, rp, st);
    if (rc != -1) return rc;
    e = cudaFuncSetAttribute(xmodal_window_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("xmodal_window_bwd: smem attr: %s", cudaGetErrorString(e)); return STGCMA_ECUDA; }
    xmodal_window_bwd_kernel<bf16><<<grid, 128, smem, st>>>((const bf16*)h, (const bf16*)hpre, gate, (const bf16*)attn_out,
                                                            (const bf16*)dout, (bf16*)dh, dgate, frames, R, ws, shift, rp);
  } else {
    STG_REQUIRE(false, "xmodal_window_bwd: bad dtype %d", dtype);
  }
  return check_launch("xmodal_window_bwd");
}

extern "C" int stgcma_xmodal_global_fwd(const void* h, const float* gate, void* out, void* attn_out, float* lse, int64_t frames,
                                        int L, int La, int rp, int dtype, void* stream) {
  if (La <= 0) La = L;
  STG_REQUIRE(frames > 0 && frames <= 65535 && L > 0, "xmodal_global_fwd: bad sizes frames=%lld L=%d", (long long)frames, L);
  STG_REQUIRE(lse, "xmodal_global_fwd: lse required");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = STGCMA_OK;
  if (dtype == STGCMA_F32) {
    XG_DISPATCH_RP(rp, rc = (xg_fwd_launch<float, RP_>(h, gate, out, attn_out, lse, frames, L, La, st)));
  } else if (dtype == STGCMA_BF16) {
    rc = xmodal_global_fwd_bf16(h, gate, out, attn_out, lse, frames, L, La, rp, st);
  } else {
    STG_REQUIRE(false, "xmodal_global_fwd: bad dtype %d", dtype);
  }
  STG_REQUIRE(rc != -1, "xmodal_global_fwd: rp=%d unsupported (16, 32, 48, 64, 96)", rp);
  return rc;
}

extern "C" int stgcma_xmodal_global_bwd(const void* h, const void* hpre, const float* gate, const void* attn_out,
                                        const float* lse, const void* dout, void* dh, float* dgate, float* work, int64_t frames,
                                        int L, int La, int rp, int dtype, void* stream) {
  if (La <= 0) La = L;
  STG_REQUIRE(frames > 0 && frames <= 65535 && L > 0, "xmodal_global_bwd: bad sizes frames=%lld L=%d", (long long)frames, L);
  STG_REQUIRE(lse && dgate && work, "xmodal_global_bwd: lse, dgate and work ([2, frames, L] fp32) required");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows0 = frames * L, rows1 = frames * (int64_t)La;
  int pg = (int)ceil_div(rows0 > rows1 ? rows0 : rows1, 8);
  if (pg > sm_count() * 8) pg = sm_count() * 8;
  dim3 pgrid(pg, 2);
  int rc = STGCMA_OK;
  if (dtype == STGCMA_F32) {
    xmodal_prep_kernel<float><<<pgrid, 256, 0, st>>>((const float*)dout, (const float*)attn_out, gate, work, dgate, rows0, rows1, rp);
    rc = check_launch("xmodal_prep");
    if (rc) return rc;
    XG_DISPATCH_RP(rp, rc = (xg_bwd_launch<float, RP_>(h, hpre, gate, lse, work, dout, dh, frames, L, La, st)));
  } else if (dtype == STGCMA_BF16) {
    xmodal_prep_kernel<bf16><<<pgrid, 256, 0, st>>>((const bf16*)dout, (const bf16*)attn_out, gate, work, dgate, rows0, rows1, rp);
    rc = check_launch("xmodal_prep");
    if (rc) return rc;
    rc = xmodal_global_bwd_bf16(h, hpre, gate, lse, work, dout, dh, frames, L, La, rp, st);
  } else {
    STG_REQUIRE(false, "xmodal_global_bwd: bad dtype %d", dtype);
  }
  STG_REQUIRE(rc != -1, "xmodal_global_bwd: rp=%d unsupported (16, 32, 48, 64, 96)", rp);
  return rc;
}
