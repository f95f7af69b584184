// Shared device / host helpers for libstgcma (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/stgcma.h"

namespace stgcma {

// ---- error plumbing (thread local; no exceptions across the C ABI) --------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> status

#define STG_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      stgcma::set_error(__VA_ARGS__);       \
      return STGCMA_EINVAL;                 \
    }                                       \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
int sm_count();

// ---- storage <-> fp32 -------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float ld_f(const float* p) { return *p; }
__device__ __forceinline__ float ld_f(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 128-bit vector of storage elements, unpacked to fp32.  Vec<T>::N elements per 16 bytes.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// 8 consecutive elements <-> fp32 registers (16 B for bf16, 2 x 16 B for fp32); p must be 16 B aligned.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  Vec<bf16> t;
  t.load(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = t.v[i];
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  Vec<bf16> t;
#pragma unroll
  for (int i = 0; i < 8; ++i) t.v[i] = v[i];
  t.store(p);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// ---- math ------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// bf16 production path: erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32 round-off level; GELU and
// GELU' below are within 7e-7 absolute of the erff versions over [-8, 8]) at about half the instructions of erff:
// one MUFU.RCP, one MUFU.EX2 and 8 FMAs, and the exponential is shared with the Gaussian term of the gradient.
// The fp32 check mode keeps erff.
__device__ __forceinline__ void erf_exp_fast(float x, float& erf_z, float& exp_mz2) {   // z = x / sqrt(2)
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  exp_mz2 = __expf(-z * z);
  const float e = fmaf(-p * t, exp_mz2, 1.0f);
  erf_z = copysignf(e, x);
}
__device__ __forceinline__ float gelu_fast(float x) {
  float e, g;
  erf_exp_fast(x, e, g);
  return 0.5f * x * (1.0f + e);
}
// QuickGELU of the CLIP MLP (CLIP_AVE.py:41-43): x * sigmoid(1.702 x), and its derivative s (1 + 1.702 x (1 - s))
__device__ __forceinline__ float qgelu(float x) { return x * __fdividef(1.0f, 1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float qgelu_grad(float x) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-1.702f * x));
  return s * fmaf(1.702f * x, 1.0f - s, 1.0f);
}
// activation and derivative from ONE evaluation of the shared transcendental terms
__device__ __forceinline__ void gelu_val_grad_fast(float x, float& y, float& dy) {
  float e, g;
  erf_exp_fast(x, e, g);
  const float cdf = 0.5f * (1.0f + e);
  y = x * cdf;
  dy = fmaf(x * 0.39894228040143268f, g, cdf);
}
__device__ __forceinline__ void qgelu_val_grad(float x, float& y, float& dy) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-1.702f * x));
  y = x * s;
  dy = s * fmaf(1.702f * x, 1.0f - s, 1.0f);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float e, g;
  erf_exp_fast(x, e, g);
  return fmaf(x * 0.39894228040143268f, g, 0.5f * (1.0f + e));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Swin shifted-window helpers (Swin_AVE.py:368-387, :728-740): token (y, x) of the un-shifted
// R x R map that sits at slot (iy, ix) of window (wy, wx) after roll(-shift), and the mask region
// id of a SHIFTED coordinate.
__device__ __forceinline__ int win_src(int w, int i, int ws, int shift, int R) {
  int u = w * ws + i + shift;
  return u >= R ? u - R : u;
}
__device__ __forceinline__ int win_region1(int u, int ws, int shift, int R) {
  return (u >= R - ws ? 1 : 0) + (u >= R - shift ? 1 : 0);
}

// GEMM epilogue, shared by the SIMT (fp32 check) and tcgen05 (bf16) kernels.
struct EpiParams {
  const float* bias;      // [groups, N] or null
  const void* res1;       // [M, ldo] or null
  const void* res2;
  const float* rowscale;  // [M] or null
  const void* aux;        // [M, ldo] (DGELU / MULAUX)
  void* out;
  void* out_pre;          // GELU: optional pre-activation output
  int64_t ldo;
  int mode;
  int out_f32;            // 1: `out` and `res1` are fp32 (the residual stream) whatever the activation dtype
};

}  // namespace stgcma
