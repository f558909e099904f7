// Head prologue on the training / default path: ModernBertPredictionHead
//   hidden = LayerNorm(GELU(dense(x)))        transformers modeling_modernbert.py:493-502,
// called from the reference at src/model/splade_modern.py:69-73, forward AND backward.
//
// Arithmetic = the reference under torch.autocast(bf16):
//   y = bf16(x16 @ Wd16^T)                    tcgen05 GEMM (trunk_gemm.cu, plain bf16 epilogue)
//   g = bf16(gelu_erf(y))                     torch's GELU on a bf16 tensor: fp32 math, one rounding
//   hidden = bf16(LayerNorm_fp32(g) * gamma)  autocast runs layer_norm in fp32; the decoder Linear casts it to bf16
// Only y and the per-row (mean, rstd) are kept for the backward; g is recomputed from y (GELU is not invertible, the
// dense output is what torch's autograd keeps too).
//
// Backward, one fused pass over the rows (HBM bound: reads dHidden fp32 + y bf16, writes dy bf16):
//   a = dHidden * gamma,  xhat = (g - mean) * rstd
//   dg = rstd * (a - mean_H(a) - xhat * mean_H(a * xhat))      LayerNorm backward (fp32)
//   dy = bf16( bf16(dg) * (Phi(y) + y * phi(y)) )              GELU backward on bf16 tensors, as autocast does it
//   dgamma = sum_rows dHidden * xhat                            per-warp registers -> per-CTA partial -> fixed-order sum
// then dx = bf16(dy @ Wd16) on the same tcgen05 GEMM (the operand is Wd16^T, which the caller keeps next to Wd16).
// The dense weight gradient dWd = dy^T x16 is a plain [H,T]x[T,H] GEMM left to the caller's BLAS.
#include "../../include/splade_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "rowwise.cuh"

#include <math.h>

namespace splade {
namespace {

constexpr int PRO_THREADS = 256;  // 8 warps, one row per warp at a time

using namespace rowwise;

// H = 256 * NCH; lane l owns columns [256 c + 8 l, 256 c + 8 l + 8) of every chunk c: every warp access is a
// contiguous 512 B (bf16) or 1 KB (fp32).
template <int NCH>
__global__ void __launch_bounds__(PRO_THREADS)
prologue_act_ln_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma, float eps, int T,
                       __nv_bfloat16* __restrict__ hidden, float2* __restrict__ stats) {
  constexpr int H = 256 * NCH;
  const int lane = static_cast<int>(threadIdx.x & 31u);
  float gam[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gamma + 256 * c + 8 * lane));
    const float4 b = __ldg(reinterpret_cast<const float4*>(gamma + 256 * c + 8 * lane) + 1);
    gam[c][0] = a.x, gam[c][1] = a.y, gam[c][2] = a.z, gam[c][3] = a.w;
    gam[c][4] = b.x, gam[c][5] = b.y, gam[c][6] = b.z, gam[c][7] = b.w;
  }
  for (int row = blockIdx.x * (PRO_THREADS / 32) + static_cast<int>(threadIdx.x >> 5); row < T;
       row += gridDim.x * (PRO_THREADS / 32)) {
    const __nv_bfloat16* src = y + static_cast<size_t>(row) * H;
    float g[NCH][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + 256 * c + 8 * lane);
      unpack8(v, g[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[c][j] = round_bf16(gelu_erf(g[c][j]));
        sum += g[c][j];
      }
    }
    const float mean = warp_sum(sum) * (1.0f / H);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = g[c][j] - mean;
        var += d * d;
      }
    const float rstd = rsqrtf(warp_sum(var) * (1.0f / H) + eps);
    __nv_bfloat16* dst = hidden + static_cast<size_t>(row) * H;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (g[c][j] - mean) * rstd * gam[c][j];
      *reinterpret_cast<uint4*>(dst + 256 * c + 8 * lane) = pack8(o);
    }
    if (lane == 0) stats[row] = make_float2(mean, rstd);
  }
}

template <int NCH>
__global__ void __launch_bounds__(PRO_THREADS, 2)  // two CTAs per SM also at H = 768 (128 registers)
prologue_bwd_kernel(const float* __restrict__ d_hidden, const __nv_bfloat16* __restrict__ y,
                    const float2* __restrict__ stats, const float* __restrict__ gamma, int T,
                    __nv_bfloat16* __restrict__ dy, float* __restrict__ dgamma_partial) {
  constexpr int H = 256 * NCH;
  __shared__ float s_part[PRO_THREADS / 32][H];
  const int lane = static_cast<int>(threadIdx.x & 31u), warp = static_cast<int>(threadIdx.x >> 5);
  float gam[NCH][8], acc[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gamma + 256 * c + 8 * lane));
    const float4 b = __ldg(reinterpret_cast<const float4*>(gamma + 256 * c + 8 * lane) + 1);
    gam[c][0] = a.x, gam[c][1] = a.y, gam[c][2] = a.z, gam[c][3] = a.w;
    gam[c][4] = b.x, gam[c][5] = b.y, gam[c][6] = b.z, gam[c][7] = b.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  }
  for (int row = blockIdx.x * (PRO_THREADS / 32) + warp; row < T; row += gridDim.x * (PRO_THREADS / 32)) {
    const float* gsrc = d_hidden + static_cast<size_t>(row) * H;
    const __nv_bfloat16* ysrc = y + static_cast<size_t>(row) * H;
    const float2 st = __ldg(&stats[row]);
    float yv[NCH][8], a[NCH][8], xh[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float4 d0 = __ldcs(reinterpret_cast<const float4*>(gsrc + 256 * c + 8 * lane));
      const float4 d1 = __ldcs(reinterpret_cast<const float4*>(gsrc + 256 * c + 8 * lane) + 1);
      const float dh[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      const uint4 v = *reinterpret_cast<const uint4*>(ysrc + 256 * c + 8 * lane);
      unpack8(v, yv[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float cdf, pdf;
        gelu_parts(yv[c][j], &cdf, &pdf);
        const float g = round_bf16(yv[c][j] * cdf);
        yv[c][j] = fmaf(yv[c][j], pdf, cdf);  // from here on: gelu'(y)
        xh[c][j] = (g - st.x) * st.y;
        a[c][j] = dh[j] * gam[c][j];
        s1 += a[c][j];
        s2 += a[c][j] * xh[c][j];
        acc[c][j] += dh[j] * xh[c][j];
      }
    }
    s1 = warp_sum(s1) * (1.0f / H);
    s2 = warp_sum(s2) * (1.0f / H);
    __nv_bfloat16* dst = dy + static_cast<size_t>(row) * H;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dg = round_bf16(st.y * (a[c][j] - s1 - xh[c][j] * s2));
        o[j] = dg * yv[c][j];
      }
      *reinterpret_cast<uint4*>(dst + 256 * c + 8 * lane) = pack8(o);
    }
  }
  // per-CTA partial of dgamma: warps in index order
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[warp][256 * c + 8 * lane + j] = acc[c][j];
  __syncthreads();
  for (int col = static_cast<int>(threadIdx.x); col < H; col += PRO_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < PRO_THREADS / 32; ++w) s += s_part[w][col];
    dgamma_partial[static_cast<size_t>(blockIdx.x) * H + col] = s;
  }
}

// dgamma[col] = sum over the per-CTA partials in a FIXED order: a block owns 32 columns, its eight warps each sum
// every eighth partial (coalesced 128-byte rows), then the eight slice sums are added in slice order.  (One thread
// per column walking all ~600 partials was a 24 us chain of dependent loads for 2 KB of output.)
__global__ void __launch_bounds__(256)
prologue_dgamma_reduce_kernel(const float* __restrict__ partial, int n_part, int H, float* __restrict__ dgamma) {
  __shared__ float s_sum[8][32];
  const int lane = static_cast<int>(threadIdx.x & 31u), slice = static_cast<int>(threadIdx.x >> 5);
  const int col = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (col < H)
    for (int p = slice; p < n_part; p += 8) s += partial[static_cast<size_t>(p) * H + col];
  s_sum[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && col < H) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_sum[w][lane];
    dgamma[col] = t;
  }
}

int prologue_grid(int T, int ctas_per_sm = 4) {
  const int want = (T + PRO_THREADS / 32 - 1) / (PRO_THREADS / 32);
  const int cap = ctas_per_sm * sm_count();  // one wave
  return want < cap ? (want < 1 ? 1 : want) : cap;
}

int check_shape(const char* who, int T, int H) {
  if (T < 1) return set_error(kErrInvalidArg, "%s: T=%d", who, T);
  if (H % 256 != 0 || H < 256 || H > 1024)
    return set_error(kErrUnsupportedShape, "%s: hidden size %d must be 256, 512, 768 or 1024", who, H);
  return kOk;
}

}  // namespace

size_t head_prologue_workspace_bytes(int T, int H) {
  (void)T;
  return 256 + static_cast<size_t>(4 * sm_count()) * H * sizeof(float);
}

int head_prologue_fwd_launch(const void* x, int rows_cap, const void* dense_w, const float* norm_w, float eps, int T,
                             int H, void* y, void* hidden, float* stats, void* workspace, size_t ws_bytes,
                             cudaStream_t stream) {
  int rc = check_shape("splade_head_prologue_fwd", T, H);
  if (rc != kOk) return rc;
  if (!x || !dense_w || !norm_w || !y || !hidden || !stats)
    return set_error(kErrInvalidArg, "splade_head_prologue_fwd: null pointer argument");
  if (rows_cap < T || rows_cap < 128)
    return set_error(kErrInvalidArg, "splade_head_prologue_fwd: buffers must hold max(T, 128) = %d rows, got %d",
                     T < 128 ? 128 : T, rows_cap);
  if (!workspace || ws_bytes < 256) return set_error(kErrWorkspace, "splade_head_prologue_fwd: workspace too small");
  int32_t* t_rows = static_cast<int32_t*>(workspace);
  SPLADE_CUDA_OK(cudaMemcpyAsync(t_rows, &T, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  rc = trunk_gemm_launch(EPI_BF16, x, rows_cap, dense_w, H, H, t_rows, y, H, nullptr, 0.f, 0, stream);
  if (rc != kOk) return rc;
  const int grid = prologue_grid(T);
  const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(y);
  __nv_bfloat16* hb = static_cast<__nv_bfloat16*>(hidden);
  float2* st = reinterpret_cast<float2*>(stats);
  switch (H / 256) {
    case 1: prologue_act_ln_kernel<1><<<grid, PRO_THREADS, 0, stream>>>(yb, norm_w, eps, T, hb, st); break;
    case 2: prologue_act_ln_kernel<2><<<grid, PRO_THREADS, 0, stream>>>(yb, norm_w, eps, T, hb, st); break;
    case 3: prologue_act_ln_kernel<3><<<grid, PRO_THREADS, 0, stream>>>(yb, norm_w, eps, T, hb, st); break;
    default: prologue_act_ln_kernel<4><<<grid, PRO_THREADS, 0, stream>>>(yb, norm_w, eps, T, hb, st); break;
  }
  return check_cuda(cudaGetLastError(), "prologue_act_ln_kernel");
}

int head_prologue_bwd_launch(const float* d_hidden, const void* y, const float* stats, const float* norm_w,
                             const void* dense_w_t, int rows_cap, int T, int H, void* dy, void* dx, float* d_norm_w,
                             void* workspace, size_t ws_bytes, cudaStream_t stream) {
  int rc = check_shape("splade_head_prologue_bwd", T, H);
  if (rc != kOk) return rc;
  if (!d_hidden || !y || !stats || !norm_w || !dy || !d_norm_w)
    return set_error(kErrInvalidArg, "splade_head_prologue_bwd: null pointer argument");
  if (dx != nullptr && !dense_w_t)
    return set_error(kErrInvalidArg, "splade_head_prologue_bwd: dx needs the transposed dense weight");
  if (rows_cap < T || rows_cap < 128)
    return set_error(kErrInvalidArg, "splade_head_prologue_bwd: buffers must hold max(T, 128) = %d rows, got %d",
                     T < 128 ? 128 : T, rows_cap);
  const int grid = prologue_grid(T, 2);  // the backward kernel holds two CTAs per SM
  const size_t need = 256 + static_cast<size_t>(grid) * H * sizeof(float);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "splade_head_prologue_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
  int32_t* t_rows = static_cast<int32_t*>(workspace);
  float* partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 256);
  const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(y);
  __nv_bfloat16* dyb = static_cast<__nv_bfloat16*>(dy);
  const float2* st = reinterpret_cast<const float2*>(stats);
  switch (H / 256) {
    case 1: prologue_bwd_kernel<1><<<grid, PRO_THREADS, 0, stream>>>(d_hidden, yb, st, norm_w, T, dyb, partial); break;
    case 2: prologue_bwd_kernel<2><<<grid, PRO_THREADS, 0, stream>>>(d_hidden, yb, st, norm_w, T, dyb, partial); break;
    case 3: prologue_bwd_kernel<3><<<grid, PRO_THREADS, 0, stream>>>(d_hidden, yb, st, norm_w, T, dyb, partial); break;
    default: prologue_bwd_kernel<4><<<grid, PRO_THREADS, 0, stream>>>(d_hidden, yb, st, norm_w, T, dyb, partial); break;
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  prologue_dgamma_reduce_kernel<<<(H + 31) / 32, 256, 0, stream>>>(partial, grid, H, d_norm_w);
  SPLADE_CUDA_OK(cudaGetLastError());
  if (dx == nullptr) return kOk;
  SPLADE_CUDA_OK(cudaMemcpyAsync(t_rows, &T, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  return trunk_gemm_launch(EPI_BF16, dy, rows_cap, dense_w_t, H, H, t_rows, dx, H, nullptr, 0.f, 0, stream);
}

}  // namespace splade
