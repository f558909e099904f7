// Row-wise backward kernels of the native trunk's TRAINING path (training side): LayerNorm backward
// on the fp32 residual stream and the GeGLU backward.  Everything else in the trunk's backward is a GEMM (dgrad on
// the tcgen05 kernel of trunk_gemm.cu with the transposed weights, wgrad on the caller's BLAS) or the attention
// backward; the orchestration is trunk_train.py.  Arithmetic = what torch.autograd does under autocast(bf16):
// LayerNorm and its backward in fp32, GELU / multiply backward on bf16 tensors with fp32 math and one rounding each.
//
// Both kernels run over ALL rows of the packed buffers: rows behind the last token hold zeros in every buffer (the
// forward is given zero-filled save buffers and only ever writes real rows or products of zero rows), and a zero
// gradient row stays a zero row through both, so no row count has to come back from the device.
#include "../../include/splade_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "rowwise.cuh"

namespace splade {
namespace {

using namespace rowwise;

constexpr int TB_THREADS = 256;

// dx[row] += LayerNorm_backward(dy[row]; x[row], gamma), dgamma partials per CTA (reduced in fixed order afterwards).
// transformers modeling_modernbert.py: attn_norm / mlp_norm / final_norm / embeddings.norm are nn.LayerNorm without
// bias; under autocast they run in fp32 on the fp32 residual stream.
template <int NCH, bool kDyF32>
__global__ void __launch_bounds__(TB_THREADS, 2)  // two CTAs per SM also at H = 768 (128 registers): twice the loads in flight
trunk_ln_bwd_kernel(const void* __restrict__ dy_raw, const float* __restrict__ x, const float* __restrict__ gamma,
                    float eps, int rows, float* __restrict__ dx, int accumulate, __nv_bfloat16* __restrict__ dx16,
                    float* __restrict__ dgamma_partial) {
  constexpr int H = 256 * NCH;
  __shared__ float s_part[TB_THREADS / 32][H];
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
  for (int row = blockIdx.x * (TB_THREADS / 32) + warp; row < rows; row += gridDim.x * (TB_THREADS / 32)) {
    const float* xs = x + static_cast<size_t>(row) * H;
    float xv[NCH][8], dv[NCH][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float4 x0 = *reinterpret_cast<const float4*>(xs + 256 * c + 8 * lane);
      const float4 x1 = *(reinterpret_cast<const float4*>(xs + 256 * c + 8 * lane) + 1);
      xv[c][0] = x0.x, xv[c][1] = x0.y, xv[c][2] = x0.z, xv[c][3] = x0.w;
      xv[c][4] = x1.x, xv[c][5] = x1.y, xv[c][6] = x1.z, xv[c][7] = x1.w;
      if (kDyF32) {
        const float* ds = static_cast<const float*>(dy_raw) + static_cast<size_t>(row) * H + 256 * c + 8 * lane;
        const float4 d0 = *reinterpret_cast<const float4*>(ds);
        const float4 d1 = *(reinterpret_cast<const float4*>(ds) + 1);
        dv[c][0] = d0.x, dv[c][1] = d0.y, dv[c][2] = d0.z, dv[c][3] = d0.w;
        dv[c][4] = d1.x, dv[c][5] = d1.y, dv[c][6] = d1.z, dv[c][7] = d1.w;
      } else {
        const __nv_bfloat16* ds = static_cast<const __nv_bfloat16*>(dy_raw) + static_cast<size_t>(row) * H;
        unpack8(*reinterpret_cast<const uint4*>(ds + 256 * c + 8 * lane), dv[c]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += xv[c][j];
    }
    const float mean = warp_sum(sum) * (1.0f / H);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xv[c][j] -= mean;
        var += xv[c][j] * xv[c][j];
      }
    const float rstd = rsqrtf(warp_sum(var) * (1.0f / H) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xv[c][j] *= rstd;                       // xhat
        acc[c][j] += dv[c][j] * xv[c][j];       // dgamma
        dv[c][j] *= gam[c][j];                  // a = dy * gamma
        s1 += dv[c][j];
        s2 += dv[c][j] * xv[c][j];
      }
    s1 = warp_sum(s1) * (1.0f / H);
    s2 = warp_sum(s2) * (1.0f / H);
    float* dst = dx + static_cast<size_t>(row) * H;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float4 o0, o1;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (dv[c][j] - s1 - xv[c][j] * s2);
      float4* p0 = reinterpret_cast<float4*>(dst + 256 * c + 8 * lane);
      if (accumulate) {
        o0 = p0[0], o1 = p0[1];
        o0.x += o[0], o0.y += o[1], o0.z += o[2], o0.w += o[3];
        o1.x += o[4], o1.y += o[5], o1.z += o[6], o1.w += o[7];
      } else {
        o0 = make_float4(o[0], o[1], o[2], o[3]);
        o1 = make_float4(o[4], o[5], o[6], o[7]);
      }
      p0[0] = o0;
      p0[1] = o1;
      if (dx16 != nullptr) {  // the bf16 copy the next dgrad GEMM reads (saves a separate cast pass over dx)
        const float ob[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        *reinterpret_cast<uint4*>(dx16 + static_cast<size_t>(row) * H + 256 * c + 8 * lane) = pack8(ob);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[warp][256 * c + 8 * lane + j] = acc[c][j];
  __syncthreads();
  for (int col = static_cast<int>(threadIdx.x); col < H; col += TB_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < TB_THREADS / 32; ++w) s += s_part[w][col];
    dgamma_partial[static_cast<size_t>(blockIdx.x) * H + col] = s;
  }
}

// dgamma[col] = sum over the per-CTA partials in a FIXED order: a block owns 32 columns, its eight warps each sum
// every eighth partial (coalesced 128-byte rows), then the eight slice sums are added in slice order.  (One thread
// per column walking all ~600 partials was a 24 us chain of dependent loads for 2 KB of output.)
__global__ void __launch_bounds__(256)
trunk_dgamma_reduce_kernel(const float* __restrict__ partial, int n_part, int H, float* __restrict__ dgamma) {
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

// GeGLU backward (modeling_modernbert.py ModernBertMLP: Wo(act(input) * gate), input | gate = Wi(x).chunk(2)):
// h = [a | g] [rows, 2I] bf16 (the Wi output, recomputed by the caller), dgg [rows, I] bf16 ->
// dh = [da | dg]:  act = bf16(gelu(a)),  dg = bf16(dgg * act),  da = bf16(bf16(dgg * g) * gelu'(a)).
__global__ void __launch_bounds__(256)
trunk_geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dgg, long long n_vec,
                       int I, __nv_bfloat16* __restrict__ dh) {
  const int vec_per_row = I / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vec_per_row;
    const int col = static_cast<int>(i - row * vec_per_row) * 8;
    const __nv_bfloat16* hr = h + row * 2 * I;
    float a[8], g[8], d[8], da[8], dg[8];
    unpack8(*reinterpret_cast<const uint4*>(hr + col), a);
    unpack8(*reinterpret_cast<const uint4*>(hr + I + col), g);
    unpack8(*reinterpret_cast<const uint4*>(dgg + row * I + col), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float cdf, pdf;
      gelu_parts(a[j], &cdf, &pdf);
      const float act = round_bf16(a[j] * cdf);
      dg[j] = d[j] * act;
      da[j] = round_bf16(d[j] * g[j]) * fmaf(a[j], pdf, cdf);
    }
    __nv_bfloat16* out = dh + row * 2 * I;
    *reinterpret_cast<uint4*>(out + col) = pack8(da);
    *reinterpret_cast<uint4*>(out + I + col) = pack8(dg);
  }
}

// Attention backward, last step: the gradients of the ROTATED q and k and of v, as the attention backward returns
// them ([B, heads, S, 64] bf16 with arbitrary batch / head / position strides), become the gradient of the Wqkv output
// row of every token: the transpose of the RoPE rotation (x1' = x1 cos + x2 sin, x2' = x2 cos - x1 sin on the pairs
// (j, j + 32); fp32 math, one bf16 rounding - apply_rotary_pos_emb's autograd, modeling_modernbert.py:205-228) and the
// interleave into [q heads | k heads | v heads].  row_of [B,S] gives the packed row of a token (>= rows: padding, skipped).
__global__ void __launch_bounds__(256)
trunk_dqkv_pack_kernel(const __nv_bfloat16* __restrict__ dq, const __nv_bfloat16* __restrict__ dk,
                       const __nv_bfloat16* __restrict__ dv, long long sb, long long sh, long long ss,
                       const long long* __restrict__ row_of, const float* __restrict__ cos_t,
                       const float* __restrict__ sin_t, int B, int S, int nh, long long rows,
                       __nv_bfloat16* __restrict__ dqkv) {
  const int H = nh * 64;
  const long long per_tok = static_cast<long long>(nh) * 16;  // q: 4 x (8+8 dims), k: 4 x (8+8), v: 8 x 8 dims
  const long long total = static_cast<long long>(B) * S * per_tok;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long tok = i / per_tok;
    const int r = static_cast<int>(i - tok * per_tok);
    const int hd = r >> 4, part = r & 15;
    const int b = static_cast<int>(tok / S), sp = static_cast<int>(tok - static_cast<long long>(b) * S);
    const long long row = row_of ? row_of[tok] : tok;
    if (row >= rows) continue;
    const long long src = b * sb + hd * sh + sp * ss;
    __nv_bfloat16* dst = dqkv + row * 3 * H + hd * 64;
    if (part < 8) {  // q (part 0..3) or k (4..7): dims [8 j, 8 j + 8) and their partners 32 above
      const int j0 = (part & 3) * 8;
      const __nv_bfloat16* g = (part < 4 ? dq : dk) + src;
      float x1[8], x2[8], o1[8], o2[8];
      unpack8(*reinterpret_cast<const uint4*>(g + j0), x1);
      unpack8(*reinterpret_cast<const uint4*>(g + 32 + j0), x2);
      const float* c = cos_t + static_cast<long long>(sp) * 32 + j0;
      const float* sn = sin_t + static_cast<long long>(sp) * 32 + j0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o1[j] = x1[j] * c[j] + x2[j] * sn[j];
        o2[j] = x2[j] * c[j] - x1[j] * sn[j];
      }
      __nv_bfloat16* d = dst + (part < 4 ? 0 : H);
      *reinterpret_cast<uint4*>(d + j0) = pack8(o1);
      *reinterpret_cast<uint4*>(d + 32 + j0) = pack8(o2);
    } else {
      const int j0 = (part - 8) * 8;
      *reinterpret_cast<uint4*>(dst + 2 * H + j0) = *reinterpret_cast<const uint4*>(dv + src + j0);
    }
  }
}

// padded[t, :] = row_of[t] < rows ? packed[row_of[t], :] : 0   (16-byte vectors; the packed -> [B,S] gather of q / k / v
// and of the attention output gradient in front of the library attention backward, ragged batches only)
__global__ void __launch_bounds__(256)
trunk_gather_rows_kernel(const uint4* __restrict__ packed, const long long* __restrict__ row_of, long long n_tok,
                         int vec_per_row, long long rows, uint4* __restrict__ padded) {
  const long long total = n_tok * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / vec_per_row;
    const int v = static_cast<int>(i - t * vec_per_row);
    const long long r = row_of[t];
    padded[i] = r < rows ? packed[r * vec_per_row + v] : make_uint4(0u, 0u, 0u, 0u);
  }
}

int ln_grid(int rows) {
  const int want = (rows + TB_THREADS / 32 - 1) / (TB_THREADS / 32);
  const int cap = 2 * sm_count();  // one wave of the two CTAs an SM holds
  return want < cap ? (want < 1 ? 1 : want) : cap;
}

}  // namespace

size_t trunk_ln_bwd_workspace_bytes(int H) { return static_cast<size_t>(4 * sm_count()) * H * sizeof(float); }

int trunk_ln_bwd_launch(const void* dy, int dy_is_f32, const float* x, const float* gamma, float eps, int rows, int H,
                        float* dx, int accumulate, void* dx_bf16, float* dgamma, void* workspace, size_t ws_bytes,
                        cudaStream_t stream) {
  if (!dy || !x || !gamma || !dx || !dgamma || rows < 1)
    return set_error(kErrInvalidArg, "splade_trunk_ln_bwd: bad argument");
  if (H % 256 != 0 || H < 256 || H > 1024)
    return set_error(kErrUnsupportedShape, "splade_trunk_ln_bwd: hidden size %d must be 256, 512, 768 or 1024", H);
  const int grid = ln_grid(rows);
  if (!workspace || ws_bytes < static_cast<size_t>(grid) * H * sizeof(float))
    return set_error(kErrWorkspace, "splade_trunk_ln_bwd: workspace too small");
  float* partial = static_cast<float*>(workspace);
  __nv_bfloat16* d16 = static_cast<__nv_bfloat16*>(dx_bf16);
#define SPLADE_LN_BWD(NCH)                                                                                          \
  do {                                                                                                              \
    if (dy_is_f32)                                                                                                  \
      trunk_ln_bwd_kernel<NCH, true><<<grid, TB_THREADS, 0, stream>>>(dy, x, gamma, eps, rows, dx, accumulate, d16, partial);  \
    else                                                                                                            \
      trunk_ln_bwd_kernel<NCH, false><<<grid, TB_THREADS, 0, stream>>>(dy, x, gamma, eps, rows, dx, accumulate, d16, partial); \
  } while (0)
  switch (H / 256) {
    case 1: SPLADE_LN_BWD(1); break;
    case 2: SPLADE_LN_BWD(2); break;
    case 3: SPLADE_LN_BWD(3); break;
    default: SPLADE_LN_BWD(4); break;
  }
#undef SPLADE_LN_BWD
  SPLADE_CUDA_OK(cudaGetLastError());
  trunk_dgamma_reduce_kernel<<<(H + 31) / 32, 256, 0, stream>>>(partial, grid, H, dgamma);
  return check_cuda(cudaGetLastError(), "trunk_dgamma_reduce_kernel");
}

int trunk_dqkv_pack_launch(const void* dq, const void* dk, const void* dv, long long sb, long long sh, long long ss,
                           const long long* row_of, const float* cos_t, const float* sin_t, int B, int S, int nh,
                           long long rows, void* dqkv, cudaStream_t stream) {
  if (!dq || !dk || !dv || !cos_t || !sin_t || !dqkv || B < 1 || S < 1 || nh < 1)
    return set_error(kErrInvalidArg, "splade_trunk_dqkv_pack: bad argument");
  if ((sb | sh | ss) % 8 != 0)
    return set_error(kErrInvalidArg, "splade_trunk_dqkv_pack: strides must be multiples of 8 elements (16-byte loads)");
  const long long total = static_cast<long long>(B) * S * nh * 16;
  const long long want = (total + 255) / 256;
  const int grid = static_cast<int>(want < 16LL * sm_count() ? want : 16LL * sm_count());
  trunk_dqkv_pack_kernel<<<grid, 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dq), static_cast<const __nv_bfloat16*>(dk),
      static_cast<const __nv_bfloat16*>(dv), sb, sh, ss, row_of, cos_t, sin_t, B, S, nh, rows,
      static_cast<__nv_bfloat16*>(dqkv));
  return check_cuda(cudaGetLastError(), "trunk_dqkv_pack_kernel");
}

int trunk_gather_rows_launch(const void* packed, const long long* row_of, long long n_tok, int row_bytes, long long rows,
                             void* padded, cudaStream_t stream) {
  if (!packed || !row_of || !padded || n_tok < 1 || row_bytes < 16 || row_bytes % 16 != 0)
    return set_error(kErrInvalidArg, "splade_trunk_gather_rows: bad argument (rows of a multiple of 16 bytes)");
  const int vec = row_bytes / 16;
  const long long want = (n_tok * vec + 255) / 256;
  const int grid = static_cast<int>(want < 16LL * sm_count() ? want : 16LL * sm_count());
  trunk_gather_rows_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(packed), row_of, n_tok, vec, rows,
                                                     static_cast<uint4*>(padded));
  return check_cuda(cudaGetLastError(), "trunk_gather_rows_kernel");
}

int trunk_geglu_bwd_launch(const void* h, const void* dgg, int rows, int I, void* dh, cudaStream_t stream) {
  if (!h || !dgg || !dh || rows < 1) return set_error(kErrInvalidArg, "splade_trunk_geglu_bwd: bad argument");
  if (I % 8 != 0) return set_error(kErrUnsupportedShape, "splade_trunk_geglu_bwd: intermediate size %d %% 8 != 0", I);
  const long long n_vec = static_cast<long long>(rows) * (I / 8);
  const long long want = (n_vec + 255) / 256;
  const int grid = static_cast<int>(want < 16LL * sm_count() ? want : 16LL * sm_count());
  trunk_geglu_bwd_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(h),
                                                   static_cast<const __nv_bfloat16*>(dgg), n_vec, I,
                                                   static_cast<__nv_bfloat16*>(dh));
  return check_cuda(cudaGetLastError(), "trunk_geglu_bwd_kernel");
}

}  // namespace splade
