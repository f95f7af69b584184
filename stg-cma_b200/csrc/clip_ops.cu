// Memory-bound kernels of the CLIP-ViT variant of the adapter train step (MM_CLIP_AVE.forward, CLIP_AVE.py:1089-1140):
//   im2col            conv1 / conv1_audio (k = s = patch, no bias) as a gather in front of stgcma_gemm       :1091, :1107
//   clip_embed_fwd    [class token | patches] + positional embedding + temporal embedding -> fp32 stream      :1094-1103
//   clip_temb_bwd     gradient of the (trainable) temporal embeddings: a [B, N] reduction of d(stream)        :749-752
//   cls_ln_fwd / bwd  ln_post (trainable affine) on the class-token rows only + cat(audio, video)            :1124-1137
#include "common.cuh"

namespace stgcma {

// out[((b*T + t)*Hp + py)*Wp + px][(c*P + ky)*P + kx] = img[b][c][t][py*P + ky][px*P + kx]
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int Cin, int Tn,
                                                     int H, int W, int P, int Hp, int Wp, int64_t ldo, int64_t total2) {
  const int K = Cin * P * P, K2 = K >> 1;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / K2;
    const int k = (int)(idx - row * K2) * 2;
    const int kx = k % P, ky = (k / P) % P, c = k / (P * P);
    int64_t r = row;
    const int px = (int)(r % Wp); r /= Wp;
    const int py = (int)(r % Hp); r /= Hp;
    const int t = (int)(r % Tn);
    const int b = (int)(r / Tn);
    const float* src = img + ((((int64_t)b * Cin + c) * Tn + t) * H + (py * P + ky)) * (int64_t)W + px * P + kx;
    const float2 v = *reinterpret_cast<const float2*>(src);
    T* dst = out + row * ldo + k;
    st_f(dst, v.x);
    st_f(dst + 1, v.y);
  }
}

// x0[(bt*N + i)][:] = (i == 0 ? cls : patches[bt*(N-1) + i - 1]) + pos[i] + temb[bt % T]
template <typename T>
__global__ void __launch_bounds__(256) clip_embed_kernel(const T* __restrict__ patches, const float* __restrict__ cls,
                                                         const float* __restrict__ pos, const float* __restrict__ temb,
                                                         float* __restrict__ x0, int64_t BT, int Tn, int N, int D) {
  const int D4 = D >> 2;
  const int64_t total = BT * N * D4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / D4;
    const int c = (int)(idx - row * D4) * 4;
    const int64_t bt = row / N;
    const int i = (int)(row - bt * N);
    const int t = (int)(bt % Tn);
    float v[4];
    if (i == 0) {
      const float4 a = *reinterpret_cast<const float4*>(cls + c);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
      const T* p = patches + (bt * (N - 1) + (i - 1)) * (int64_t)D + c;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ld_f(p + k);
    }
    const float4 pp = *reinterpret_cast<const float4*>(pos + (int64_t)i * D + c);
    const float4 tt = *reinterpret_cast<const float4*>(temb + (int64_t)t * D + c);
    *reinterpret_cast<float4*>(x0 + row * D + c) = make_float4(v[0] + pp.x + tt.x, v[1] + pp.y + tt.y, v[2] + pp.z + tt.z, v[3] + pp.w + tt.w);
  }
}

// dtemb[t][:] += sum_{b, i} dx0[((b*T + t)*N + i)][:]        grid (T, column blocks, splits)
__global__ void __launch_bounds__(256) clip_temb_bwd_kernel(const float* __restrict__ dx0, float* __restrict__ dtemb, int64_t B,
                                                            int Tn, int N, int D) {
  const int t = blockIdx.x;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const int64_t rows = B * N;   // (b, i) pairs of this t
  const int64_t per = (rows + gridDim.z - 1) / gridDim.z;
  const int64_t r0 = blockIdx.z * per, r1 = r0 + per < rows ? r0 + per : rows;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t b = r / N;
    const int i = (int)(r - b * N);
    const float4 v = __ldg(reinterpret_cast<const float4*>(dx0 + ((b * Tn + t) * N + i) * (int64_t)D + c));
    a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
  }
  float* d = dtemb + (int64_t)t * D + c;
  atomicAdd(d, a0); atomicAdd(d + 1, a1); atomicAdd(d + 2, a2); atomicAdd(d + 3, a3);
}

// one warp per (modality z, frame f): LayerNorm of the class-token row -> feat[f][(1 - z) * D ..]  (audio first)
__global__ void __launch_bounds__(256) cls_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ feat,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                         int64_t frames, int Nv, int Na, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= 2 * frames) return;
  const int z = w >= frames;
  const int64_t f = z ? w - frames : w;
  const float* xr = x + (z ? frames * (int64_t)Nv + f * Na : f * Nv) * (int64_t)D;
  float s = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    q += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean) + (v.z - mean) * (v.z - mean) + (v.w - mean) * (v.w - mean);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
  float* o = feat + f * 2 * (int64_t)D + (1 - z) * D;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    *reinterpret_cast<float4*>(o + c) = make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y,
                                                    (v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w);
  }
  if (lane == 0) {
    mean_out[w] = mean;
    rstd_out[w] = rstd;
  }
}

// dx (class-token rows only; the caller zero-fills the rest), dgamma += dy * xhat, dbeta += dy
template <typename TC>
__global__ void __launch_bounds__(256) cls_ln_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ x,
                                                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                         const float* __restrict__ gamma, float* __restrict__ dx,
                                                         TC* __restrict__ dx_copy, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, int64_t frames, int Nv, int Na, int D) {
  extern __shared__ float cls_sm[];   // [2][D]: per-block partial dgamma, dbeta
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) cls_sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w < 2 * frames) {
    const int z = w >= frames;
    const int64_t f = z ? w - frames : w;
    const int64_t off = (z ? frames * (int64_t)Nv + f * Na : f * Nv) * (int64_t)D;
    const float* xr = x + off;
    const float* dy = dfeat + f * 2 * (int64_t)D + (1 - z) * D;
    const float mean = mean_in[w], rstd = rstd_in[w];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float gg = dy[c] * gamma[c];
      s1 += gg;
      s2 += gg * (xr[c] - mean) * rstd;
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    for (int c = lane; c < D; c += 32) {
      const float xh = (xr[c] - mean) * rstd;
      const float d = dy[c];
      const float dv = rstd * (d * gamma[c] - s1 - xh * s2);
      dx[off + c] = dv;
      if (dx_copy) st_f(dx_copy + off + c, dv);
      atomicAdd(cls_sm + c, d * xh);
      atomicAdd(cls_sm + D + c, d);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(dgamma + c, cls_sm[c]);
    atomicAdd(dbeta + c, cls_sm[D + c]);
  }
}

}  // namespace stgcma

using namespace stgcma;

extern "C" int stgcma_im2col(const float* img, void* out, int B, int Cin, int T, int H, int W, int P, int64_t ldo, int dtype,
                             void* stream) {
  STG_REQUIRE(img && out, "im2col: null operand");
  STG_REQUIRE(B > 0 && Cin > 0 && T > 0 && P > 0 && H >= P && W >= P, "im2col: bad shape");
  STG_REQUIRE(P % 2 == 0 && W % 2 == 0 && ((uintptr_t)img & 7) == 0, "im2col: patch size and image width must be even");
  const int Hp = (H - P) / P + 1, Wp = (W - P) / P + 1;
  const int K = Cin * P * P;
  STG_REQUIRE(ldo >= K, "im2col: ldo=%lld < K=%d", (long long)ldo, K);
  const int64_t rows = (int64_t)B * T * Hp * Wp;
  const int64_t total2 = rows * (K / 2);
  const int grid = (int)(ceil_div(total2, 256) < (int64_t)sm_count() * 16 ? ceil_div(total2, 256) : (int64_t)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == STGCMA_F32) im2col_kernel<float><<<grid, 256, 0, st>>>(img, (float*)out, B, Cin, T, H, W, P, Hp, Wp, ldo, total2);
  else if (dtype == STGCMA_BF16) im2col_kernel<bf16><<<grid, 256, 0, st>>>(img, (bf16*)out, B, Cin, T, H, W, P, Hp, Wp, ldo, total2);
  else STG_REQUIRE(false, "im2col: bad dtype %d", dtype);
  return check_launch("im2col");
}

extern "C" int stgcma_clip_embed_fwd(const void* patches, const float* cls, const float* pos, const float* temb, float* x0,
                                     int64_t BT, int T, int N, int D, int dtype, void* stream) {
  STG_REQUIRE(patches && cls && pos && temb && x0, "clip_embed_fwd: null operand");
  STG_REQUIRE(BT > 0 && T > 0 && BT % T == 0 && N > 1 && D % 4 == 0, "clip_embed_fwd: bad shape");
  const int64_t total = BT * N * (D / 4);
  const int grid = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 16 ? ceil_div(total, 256) : (int64_t)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == STGCMA_F32) clip_embed_kernel<float><<<grid, 256, 0, st>>>((const float*)patches, cls, pos, temb, x0, BT, T, N, D);
  else if (dtype == STGCMA_BF16) clip_embed_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)patches, cls, pos, temb, x0, BT, T, N, D);
  else STG_REQUIRE(false, "clip_embed_fwd: bad dtype %d", dtype);
  return check_launch("clip_embed_fwd");
}

extern "C" int stgcma_clip_temb_bwd(const float* dx0, float* dtemb, int64_t B, int T, int N, int D, void* stream) {
  STG_REQUIRE(dx0 && dtemb, "clip_temb_bwd: null operand");
  STG_REQUIRE(B > 0 && T > 0 && N > 0 && D % 4 == 0, "clip_temb_bwd: bad shape");
  const int colblocks = (int)ceil_div(D / 4, 256);
  int splits = (int)ceil_div(4 * sm_count(), (int64_t)T * colblocks);
  if (splits > B * N) splits = (int)(B * N);
  if (splits < 1) splits = 1;
  dim3 grid(T, colblocks, splits);
  clip_temb_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dx0, dtemb, B, T, N, D);
  return check_launch("clip_temb_bwd");
}

extern "C" int stgcma_cls_ln_fwd(const float* x, const float* gamma, const float* beta, float* feat, float* mean, float* rstd,
                                 int64_t frames, int Nv, int Na, int D, void* stream) {
  STG_REQUIRE(x && gamma && beta && feat && mean && rstd, "cls_ln_fwd: null operand");
  STG_REQUIRE(frames > 0 && Nv > 0 && Na > 0 && D % 4 == 0, "cls_ln_fwd: bad shape");
  cls_ln_fwd_kernel<<<(unsigned)ceil_div(2 * frames, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, feat, mean, rstd, frames,
                                                                                        Nv, Na, D);
  return check_launch("cls_ln_fwd");
}

extern "C" int stgcma_cls_ln_bwd(const float* dfeat, const float* x, const float* mean, const float* rstd, const float* gamma,
                                 float* dx, void* dx_copy, float* dgamma, float* dbeta, int64_t frames, int Nv, int Na, int D,
                                 int dtype, void* stream) {
  STG_REQUIRE(dfeat && x && mean && rstd && gamma && dx && dgamma && dbeta, "cls_ln_bwd: null operand");
  STG_REQUIRE(frames > 0 && Nv > 0 && Na > 0 && D > 0 && D <= 4096, "cls_ln_bwd: bad shape");
  const unsigned grid = (unsigned)ceil_div(2 * frames, 8);
  const size_t smem = 2 * D * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == STGCMA_BF16)
    cls_ln_bwd_kernel<bf16><<<grid, 256, smem, st>>>(dfeat, x, mean, rstd, gamma, dx, (bf16*)dx_copy, dgamma, dbeta, frames, Nv, Na, D);
  else if (dtype == STGCMA_F32)
    cls_ln_bwd_kernel<float><<<grid, 256, smem, st>>>(dfeat, x, mean, rstd, gamma, dx, (float*)dx_copy, dgamma, dbeta, frames, Nv, Na, D);
  else STG_REQUIRE(false, "cls_ln_bwd: bad dtype %d", dtype);
  return check_launch("cls_ln_bwd");
}
