// Row-wise helpers shared by the head prologue (head_prologue.cu) and the trunk's backward kernels (trunk_bwd.cu):
// one warp per row of H = 256 * NCH features, lane l owning the 8 consecutive columns [256 c + 8 l, +8) of chunk c.
#pragma once

#include "common.cuh"

namespace splade {
namespace rowwise {

// Exact-erf GELU and its derivative from ONE exponential: Abramowitz-Stegun 7.1.26, erfc(|x|) = poly(t) e^{-x^2} with
// t = 1 / (1 + p |x|) (absolute error <= 3e-7 in fp32 - four decimal orders below the bf16 rounding that follows), and
// e^{-x^2} with x = y / sqrt 2 IS the Gaussian the derivative needs: Phi'(y) = e^{-y^2/2} / sqrt(2 pi).  It costs two
// MUFU operations per element (rcp, ex2): right for these memory-bound row-wise kernels, where the backward wants cdf
// and pdf together; inside the GEMM epilogues (MUFU-starved, trunk_gemm.cu) erff's pure FMA polynomial is faster
// (measured: GeGLU GEMM 51.8 -> 65.7 us with this form), so they keep erff.
__device__ __forceinline__ void gelu_parts(float y, float* cdf, float* pdf) {
  const float ax = fabsf(y) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));  // rcp.approx (1 ulp): the IEEE reciprocal is 8 instructions
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float e = __expf(-ax * ax);
  const float half_erfc = 0.5f * t * poly * e;  // 0.5 erfc(|x|)
  *cdf = y >= 0.f ? 1.0f - half_erfc : half_erfc;
  *pdf = 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_erf(float y) {
  float cdf, pdf;
  gelu_parts(y, &cdf, &pdf);
  return y * cdf;
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(cvt_bf16x2(f[1], f[0]), cvt_bf16x2(f[3], f[2]), cvt_bf16x2(f[5], f[4]), cvt_bf16x2(f[7], f[6]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

}  // namespace rowwise
}  // namespace splade
