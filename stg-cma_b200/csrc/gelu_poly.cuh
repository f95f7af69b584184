// MUFU-free GELU / GELU' for the bf16 epilogues of the fused MLP kernels (mlp_fused.cu).
//
// nn.GELU() (exact erf form, Swin_AVE.py:116-127) costs two MUFU ops per element when evaluated through erf
// (rcp + ex2); the MUFU pipe issues 16 per clock per SM, so at Swin stage 0 (K = 96) the activation, not the
// tensor pipe, bounds the fused kernel.  Here Phi(x) - 1/2 and gelu'(x) - 1/2 (both odd) are near-minimax odd
// polynomials on a clamped range, 11 / 13 FMA-pipe instructions, no MUFU:
//     gelu(x)  = x * sat(1/2 + xc P(xc^2)),  xc = clamp(x, -4, 4)        max abs error 2.0e-4 (|Phi| error 4.9e-5)
//     gelu'(x) =         1/2 + xc Q(xc^2),   xc = clamp(x, -4.5, 4.5)    max abs error 1.5e-4
// i.e. 10x below the bf16 rounding (2^-9 relative) applied to the result.  Coefficients: scripts/fit_gelu_poly.py.
// The fp32 check mode keeps erff (common.cuh).
#pragma once

namespace stgcma {

__device__ __forceinline__ float gelu_poly(float x) {
  const float xc = fminf(fmaxf(x, -4.0f), 4.0f);
  const float t = xc * xc;
  float p = -1.581098869e-09f;
  p = fmaf(p, t, 1.217303207e-07f);
  p = fmaf(p, t, -4.101366426e-06f);
  p = fmaf(p, t, 8.067483337e-05f);
  p = fmaf(p, t, -1.048278778e-03f);
  p = fmaf(p, t, 9.665444308e-03f);
  p = fmaf(p, t, -6.617897502e-02f);
  p = fmaf(p, t, 3.988687775e-01f);
  float phi;
  asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(phi) : "f"(xc), "f"(p));   // sat(xc * p + 0.5)
  return x * phi;
}

// Forward GELU with one MUFU op: Phi(x) = 1/2 + 1/2 tanh(x (c0 + c1 x^2 + c2 x^4)) -- the odd polynomial is a fit of
// atanh(erf(x / sqrt 2)) (|Phi| error 5.0e-5, scripts/fit_gelu_poly.py), tanh is MUFU.TANH (relative error 2^-11,
// i.e. |Phi| error <= 2.4e-4): 8 issue slots per element instead of 12, and the MUFU pipe (16 / clk / SM) carries
// exactly one of them.  x^2 is clamped at 80 so the polynomial stays monotone for |x| > 9.
__device__ __forceinline__ float gelu_tanh(float x) {
  const float t = fminf(x * x, 80.0f);
  const float g = x * fmaf(t, fmaf(t, -3.51519793e-04f, 3.70056659e-02f), 7.97507861e-01f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(g));
  return x * fmaf(0.5f, th, 0.5f);
}

// Activation AND derivative from the tanh form: Phi as in gelu_tanh (|error| <= 2.4e-4), the Gaussian term of the derivative
// from one ex2 -- 12 issue slots (2 MUFU) instead of the ~30 of the erf / expf / rcp form (ncu, profiles/r02d_ncu_gemm_fc1g_s2.txt:
// the derivative-saving fc1 GEMMs of stages 2-3 spend 3/4 of their epilogue instructions there).  Both results are rounded to
// bf16 (2^-9 relative) right after.
__device__ __forceinline__ void gelu_val_grad_tanh(float x, float& y, float& dy) {
  const float t = fminf(x * x, 80.0f);
  const float g = x * fmaf(t, fmaf(t, -3.51519793e-04f, 3.70056659e-02f), 7.97507861e-01f);
  float th, e;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(g));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * -0.72134752044448170f));   // exp(-x^2 / 2)
  const float cdf = fmaf(0.5f, th, 0.5f);
  y = x * cdf;
  dy = fmaf(x * 0.39894228040143268f, e, cdf);
}

__device__ __forceinline__ float gelu_grad_poly(float x) {
  const float xc = fminf(fmaxf(x, -4.5f), 4.5f);
  const float t = xc * xc;
  float q = -2.211075221e-11f;
  q = fmaf(q, t, 2.521566252e-09f);
  q = fmaf(q, t, -1.268147612e-07f);
  q = fmaf(q, t, 3.722128647e-06f);
  q = fmaf(q, t, -7.122430330e-05f);
  q = fmaf(q, t, 9.405587952e-04f);
  q = fmaf(q, t, -8.815928839e-03f);
  q = fmaf(q, t, 5.860945634e-02f);
  q = fmaf(q, t, -2.649258273e-01f);
  q = fmaf(q, t, 7.976261832e-01f);
  return fmaf(xc, q, 0.5f);
}

}  // namespace stgcma
