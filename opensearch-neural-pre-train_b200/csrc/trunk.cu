// Native variable-length ModernBERT trunk for inference: the 22-layer encoder that
// feeds the SPLADE head, on PACKED tokens only (no padding anywhere), bf16 GEMMs on tcgen05 with fused epilogues
// (trunk_gemm.cu), and the small memory-bound kernels below.  Mirrors transformers' modeling_modernbert.py as
// the reference runs it under autocast(bf16): Linear outputs are bf16, the residual stream and every LayerNorm
// are fp32, RoPE is computed in fp32 on the bf16 q/k, softmax in fp32.
//
//   embeddings (:52-71)   tok_embeddings -> LayerNorm                                  trunk_embed_ln_kernel
//   layer (:327-343)      x += Wo(attn(rope(Wqkv(norm(x)))));  x += Wo(gelu(a) * g)    trunk_ln_kernel, GEMMs,
//                                                                                      trunk_attn_kernel
//   final_norm (:488), head.dense -> GELU -> head.norm (:501-502)                      trunk_ln_kernel, GEMM
//
// Attention: bidirectional, 64-wide heads, global every third layer and a +-64-token sliding window otherwise
// (config.local_attention = 128, :259-264); one CTA per (64 queries, head, document), flash-style online softmax
// on mma.sync m16n8k16 tensor-core tiles - it is ~7 % of the trunk's FLOPs at L=256, the GEMMs are the rest.
#include "../../include/splade_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <math.h>

namespace splade {
namespace {

size_t al256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// ------------------------------------------------------------------------------------------------ RoPE table
// (cos, sin)(pos * theta^(-2i/64)), fp32 as ModernBertRotaryEmbedding.forward (:158-172) computes them
__global__ void rope_table_kernel(float2* __restrict__ table, int max_pos, float theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * 32) return;
  const int pos = i >> 5, f = i & 31;
  const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * f) / 64.0f);
  const float ang = static_cast<float>(pos) * inv_freq;
  table[i] = make_float2(cosf(ang), sinf(ang));
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row; fp32 statistics (two passes over a row that sits in L1), weight only (norm_bias = false).
template <typename InT>
__device__ __forceinline__ float ld_as_float(const InT* p, int i);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p, int i) { return p[i]; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int i) {
  return __bfloat162float(p[i]);
}

template <typename InT>
__device__ __forceinline__ void warp_layernorm_row(const InT* __restrict__ src, const float* __restrict__ gamma,
                                                   float eps, int H, int lane, float* __restrict__ out_f32,
                                                   __nv_bfloat16* __restrict__ out_bf16) {
  float sum = 0.f;
  for (int i = lane; i < H; i += 32) sum += ld_as_float<InT>(src, i);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  const float mean = sum / static_cast<float>(H);
  float var = 0.f;
  for (int i = lane; i < H; i += 32) {
    const float d = ld_as_float<InT>(src, i) - mean;
    var += d * d;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
  const float rstd = rsqrtf(var / static_cast<float>(H) + eps);
  for (int i = lane; i < H; i += 32) {
    const float y = (ld_as_float<InT>(src, i) - mean) * rstd * gamma[i];
    if (out_f32 != nullptr) out_f32[i] = y;
    if (out_bf16 != nullptr) out_bf16[i] = __float2bfloat16_rn(y);
  }
}

template <typename InT>
__global__ void __launch_bounds__(256)
trunk_ln_kernel(const InT* __restrict__ in, const float* __restrict__ gamma, float eps,
                const int32_t* __restrict__ t_rows, int H, __nv_bfloat16* __restrict__ out) {
  const int T = __ldg(t_rows);
  const int lane = static_cast<int>(threadIdx.x & 31u);
  for (int row = blockIdx.x * 8 + static_cast<int>(threadIdx.x >> 5); row < T; row += gridDim.x * 8) {
    warp_layernorm_row<InT>(in + static_cast<size_t>(row) * H, gamma, eps, H, lane, nullptr,
                            out + static_cast<size_t>(row) * H);
  }
}

// x = LayerNorm(tok_embeddings[input_ids[doc, pos]]) for every packed column (replicas included: they carry the
// descriptor of the document's last token, so they stay exact duplicates of it through the whole trunk)
__global__ void __launch_bounds__(256)
trunk_embed_ln_kernel(const int64_t* __restrict__ input_ids, const float* __restrict__ tok_emb,
                      const float* __restrict__ gamma, float eps, const int32_t* __restrict__ t_rows,
                      const uint32_t* __restrict__ meta, int S, int H, float* __restrict__ x,
                      __nv_bfloat16* __restrict__ xb) {
  const int T = __ldg(t_rows);
  const int lane = static_cast<int>(threadIdx.x & 31u);
  for (int row = blockIdx.x * 8 + static_cast<int>(threadIdx.x >> 5); row < T; row += gridDim.x * 8) {
    const uint32_t m = __ldg(&meta[row]);
    const int doc = static_cast<int>((m >> 16) & 0x7FFFu), pos = static_cast<int>(m & 0xFFFFu);
    const long long id = input_ids[static_cast<size_t>(doc) * S + pos];
    warp_layernorm_row<float>(tok_emb + static_cast<size_t>(id) * H, gamma, eps, H, lane,
                              x + static_cast<size_t>(row) * H, xb + static_cast<size_t>(row) * H);
  }
}

// ------------------------------------------------------------------------------------------------ attention
constexpr int AT_BQ = 64;   // queries per CTA (4 warps x 16 rows)
constexpr int AT_BK = 64;   // keys per step
constexpr int AT_D = 64;    // head dim
constexpr int AT_LD = 72;   // smem row stride in bf16 (144 B: conflict-free 32-bit fragment loads)

struct AttnParams {
  const __nv_bfloat16* qkv;  // [T, 3H]: q | k | v, RoPE already applied to q and k
  __nv_bfloat16* out;        // [T, H]
  const int32_t* len;
  const int32_t* off;
  const uint32_t* meta;
  int H;
  int window;  // < 0: global attention; otherwise |pos_q - pos_k| <= window
  float scale;
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128)
trunk_attn_kernel(const AttnParams p) {
  const int b = static_cast<int>(blockIdx.z);
  const int h = static_cast<int>(blockIdx.y);
  const int q0 = static_cast<int>(blockIdx.x) * AT_BQ;
  const int L = __ldg(&p.len[b]);
  if (L == 0) return;
  const int L8 = (L + 7) & ~7;  // real tokens + replicas of the last one (they are queries, never keys)
  if (q0 >= L8) return;
  const int base = __ldg(&p.off[b]);
  const int H3 = 3 * p.H;

  __shared__ __align__(16) __nv_bfloat16 Qs[AT_BQ][AT_LD];
  __shared__ __align__(16) __nv_bfloat16 Ks[AT_BK][AT_LD];
  __shared__ __align__(16) __nv_bfloat16 Vt[AT_D][AT_LD];  // transposed: [dim][key]
  __shared__ int kpos[AT_BK];

  const int tid = static_cast<int>(threadIdx.x);
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  for (int i = 0; i < 4; ++i) {
    const int c = tid + 128 * i;
    const int r = c >> 3, ch = c & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + r < L8)
      v = __ldg(reinterpret_cast<const uint4*>(p.qkv + static_cast<size_t>(base + q0 + r) * H3 + h * AT_D + ch * 8));
    *reinterpret_cast<uint4*>(&Qs[r][ch * 8]) = v;
  }
  __syncthreads();
  uint32_t qa[4][4];
  {
    const int r0 = warp * 16 + g;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(&Qs[r0][ks * 16 + 2 * t]);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(&Qs[r0 + 8][ks * 16 + 2 * t]);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(&Qs[r0][ks * 16 + 8 + 2 * t]);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(&Qs[r0 + 8][ks * 16 + 8 + 2 * t]);
    }
  }
  int qpos[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int row = min(q0 + warp * 16 + g + 8 * e, L8 - 1);
    qpos[e] = static_cast<int>(__ldg(&p.meta[base + row]) & 0xFFFFu);
  }
  int qmin = 0, qmax = 0;
  if (p.window >= 0) {
    qmin = static_cast<int>(__ldg(&p.meta[base + q0]) & 0xFFFFu);
    qmax = static_cast<int>(__ldg(&p.meta[base + min(q0 + AT_BQ - 1, L8 - 1)]) & 0xFFFFu);
  }

  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;

  const int n_kv = (L + AT_BK - 1) / AT_BK;
  for (int kb = 0; kb < n_kv; ++kb) {
    const int kv0 = kb * AT_BK;
    if (p.window >= 0) {  // CTA-uniform: the whole key block is outside every query's window
      const int kfirst = static_cast<int>(__ldg(&p.meta[base + kv0]) & 0xFFFFu);
      const int klast = static_cast<int>(__ldg(&p.meta[base + min(kv0 + AT_BK - 1, L - 1)]) & 0xFFFFu);
      if (klast < qmin - p.window || kfirst > qmax + p.window) continue;
    }
    __syncthreads();  // everyone is done with the previous K / V tiles
    for (int i = 0; i < 4; ++i) {
      const int c = tid + 128 * i;
      {  // K rows: 16-byte chunks, chunk index fastest (coalesced)
        const int r = c >> 3, ch = c & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (kv0 + r < L)
          v = __ldg(reinterpret_cast<const uint4*>(p.qkv + static_cast<size_t>(base + kv0 + r) * H3 + p.H +
                                                   h * AT_D + ch * 8));
        *reinterpret_cast<uint4*>(&Ks[r][ch * 8]) = v;
      }
      {  // V rows, stored transposed: key index fastest so that the 2-byte stores of a warp are contiguous
        const int r = c & 63, ch = c >> 6;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (kv0 + r < L)
          v = __ldg(reinterpret_cast<const uint4*>(p.qkv + static_cast<size_t>(base + kv0 + r) * H3 + 2 * p.H +
                                                   h * AT_D + ch * 8));
        const __nv_bfloat16* e8 = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) Vt[ch * 8 + j][r] = e8[j];
      }
    }
    if (tid < AT_BK) {
      const int key = kv0 + tid;
      kpos[tid] = key < L ? static_cast<int>(__ldg(&p.meta[base + key]) & 0xFFFFu) : 0x3FFFFFFF;
    }
    __syncthreads();

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + 8 + 2 * t]);
        mma_bf16_16816(s[nt], qa[ks], b0, b1);
      }
    }
    // scale, mask (document end, sliding window), running maximum
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * t + (e & 1);
        const int r = e >> 1;
        bool ok = kv0 + col < L;
        if (p.window >= 0) {
          const int d = qpos[r] - kpos[col];
          ok = ok && d <= p.window && d >= -p.window;
        }
        s[nt][e] = ok ? s[nt][e] * p.scale : -INFINITY;
        mx[r] = fmaxf(mx[r], s[nt][e]);
      }
    }
    float corr[2], moff[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(m_run[r], mx[r]);
      moff[r] = mnew == -INFINITY ? 0.f : mnew;  // nothing visible yet: keep exp() finite
      corr[r] = __expf(m_run[r] - moff[r]);      // exp(-inf) = 0 on the first visible block
      m_run[r] = mnew;
      l_run[r] *= corr[r];
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = __expf(s[nt][0] - moff[0]);
      const float p1 = __expf(s[nt][1] - moff[0]);
      const float p2 = __expf(s[nt][2] - moff[1]);
      const float p3 = __expf(s[nt][3] - moff[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      const int ks = nt >> 1;
      if ((nt & 1) == 0) {
        pa[ks][0] = cvt_bf16x2(p1, p0);
        pa[ks][1] = cvt_bf16x2(p3, p2);
      } else {
        pa[ks][2] = cvt_bf16x2(p1, p0);
        pa[ks][3] = cvt_bf16x2(p3, p2);
      }
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= corr[0];
      o[dt][1] *= corr[0];
      o[dt][2] *= corr[1];
      o[dt][3] *= corr[1];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Vt[dt * 8 + g][ks * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Vt[dt * 8 + g][ks * 16 + 8 + 2 * t]);
        mma_bf16_16816(o[dt], pa[ks], b0, b1);
      }
    }
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.0f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.0f / l_run[1] : 0.f;
  const int row0 = q0 + warp * 16 + g;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = h * AT_D + dt * 8 + 2 * t;
    if (row0 < L8)
      *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(base + row0) * p.H + col) =
          cvt_bf16x2(o[dt][1] * inv0, o[dt][0] * inv0);
    if (row0 + 8 < L8)
      *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(base + row0 + 8) * p.H + col) =
          cvt_bf16x2(o[dt][3] * inv1, o[dt][2] * inv1);
  }
}

}  // namespace

int trunk_rope_table_launch(float2* table, int max_pos, float theta, cudaStream_t stream) {
  const int n = max_pos * 32;
  rope_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(table, max_pos, theta);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int trunk_embed_ln_launch(const int64_t* input_ids, const float* tok_emb, const float* gamma, float eps,
                          const VarlenLayout& lay, int S, int H, float* x, void* xb, cudaStream_t stream) {
  trunk_embed_ln_kernel<<<4 * sm_count(), 256, 0, stream>>>(input_ids, tok_emb, gamma, eps, &lay.plan->T, lay.meta, S,
                                                            H, x, static_cast<__nv_bfloat16*>(xb));
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int trunk_ln_launch(const void* in, int in_is_bf16, const float* gamma, float eps, const int32_t* t_rows,
                    int rows_cap, int H, void* out_bf16, cudaStream_t stream) {
  (void)rows_cap;
  if (in_is_bf16) {
    trunk_ln_kernel<__nv_bfloat16><<<4 * sm_count(), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), gamma, eps, t_rows, H, static_cast<__nv_bfloat16*>(out_bf16));
  } else {
    trunk_ln_kernel<float><<<4 * sm_count(), 256, 0, stream>>>(static_cast<const float*>(in), gamma, eps, t_rows, H,
                                                               static_cast<__nv_bfloat16*>(out_bf16));
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int trunk_attn_launch(const void* qkv, void* out, const VarlenLayout& lay, int B, int S, int H, int n_heads,
                      int window, cudaStream_t stream) {
  if (H != n_heads * AT_D) return set_error(kErrUnsupportedShape, "trunk attention: head dim must be 64");
  AttnParams p{};
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.len = lay.len;
  p.off = lay.off;
  p.meta = lay.meta;
  p.H = H;
  p.window = window;
  p.scale = 0.125f;  // head_dim ** -0.5
  const int s8 = (S + 7) & ~7;
  trunk_attn_kernel<<<dim3((s8 + AT_BQ - 1) / AT_BQ, n_heads, B), 128, 0, stream>>>(p);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------ layer loop
size_t trunk_workspace_bytes(int B, int S, int H, int I, int V) {
  const size_t rows = (static_cast<size_t>(B) * ((static_cast<size_t>(S) + 7) / 8 * 8) + 255) / 256 * 256;
  size_t bytes = al256(varlen_layout_bytes(B, S, V));
  bytes += al256(rows * H * 4);      // x   fp32 residual stream
  bytes += al256(rows * H * 2);      // xb  bf16 GEMM input / final hidden
  bytes += al256(rows * 3 * H * 2);  // qkv
  bytes += al256(rows * H * 2);      // attention output / head.dense output
  bytes += al256(rows * I * 2);      // GeGLU output
  bytes += 2 * al256(static_cast<size_t>(S) * 32 * sizeof(float2));  // RoPE tables (global, local)
  return bytes;
}

int trunk_encode_launch(const splade_trunk_args* a, cudaStream_t stream) {
  if (!a || !a->input_ids || !a->mask || !a->tok_emb || !a->layers || !a->dec_w || !a->dec_bias || !a->rep ||
      !a->argmax)
    return set_error(kErrInvalidArg, "splade_trunk_encode: null pointer argument");
  const int B = a->B, S = a->S, H = a->H, I = a->I, V = a->V;
  if (B < 1 || S < 1 || H % 64 != 0 || H < 256 || I % 128 != 0 || a->n_heads * 64 != H || a->n_layers < 1)
    return set_error(kErrUnsupportedShape,
                     "splade_trunk_encode: needs hidden %% 64 == 0 and >= 256, 64-wide heads, intermediate %% 128 == 0 "
                     "(B=%d S=%d H=%d I=%d heads=%d)", B, S, H, I, a->n_heads);
  int rc = require_sm100("splade_trunk_encode");
  if (rc != kOk) return rc;
  const size_t need = trunk_workspace_bytes(B, S, H, I, V);
  if (!a->workspace || a->ws_bytes < need)
    return set_error(kErrWorkspace, "splade_trunk_encode: workspace too small (%zu < %zu bytes)", a->ws_bytes, need);

  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  VarlenLayout lay{};
  rc = varlen_layout_build(a->mask, B, S, V, ws, varlen_layout_bytes(B, S, V), &lay, stream);
  if (rc != kOk) return rc;
  ws += al256(varlen_layout_bytes(B, S, V));
  const size_t rows = lay.rows;
  float* x = reinterpret_cast<float*>(ws);
  ws += al256(rows * H * 4);
  void* xb = ws;
  ws += al256(rows * H * 2);
  void* qkv = ws;
  ws += al256(rows * 3 * H * 2);
  void* ab = ws;
  ws += al256(rows * H * 2);
  void* gg = ws;
  ws += al256(rows * I * 2);
  float2* rope_g = reinterpret_cast<float2*>(ws);
  ws += al256(static_cast<size_t>(S) * 32 * sizeof(float2));
  float2* rope_l = reinterpret_cast<float2*>(ws);
  const int32_t* t_rows = &lay.plan->T;
  const int rows_i = static_cast<int>(rows);

  if ((rc = trunk_rope_table_launch(rope_g, S, a->global_theta, stream)) != kOk) return rc;
  if ((rc = trunk_rope_table_launch(rope_l, S, a->local_theta, stream)) != kOk) return rc;
  if ((rc = trunk_embed_ln_launch(a->input_ids, a->tok_emb, a->emb_norm, a->norm_eps, lay, S, H, x, xb, stream)) != kOk)
    return rc;

  for (int l = 0; l < a->n_layers; ++l) {
    const splade_trunk_layer& ly = a->layers[l];
    if (ly.attn_norm != nullptr) {
      if ((rc = trunk_ln_launch(x, 0, ly.attn_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
    }
    if ((rc = trunk_gemm_launch(EPI_QKV_ROPE, xb, rows_i, ly.wqkv, 3 * H, H, t_rows, qkv, 3 * H, lay.meta,
                                ly.is_global ? rope_g : rope_l, 2 * H, S, stream)) != kOk)
      return rc;
    if ((rc = trunk_attn_launch(qkv, ab, lay, B, S, H, a->n_heads, ly.is_global ? -1 : a->local_window, stream)) != kOk)
      return rc;
    if ((rc = trunk_gemm_launch(EPI_RESID, ab, rows_i, ly.wo, H, H, t_rows, x, H, nullptr, nullptr, 0, 0, stream)) != kOk)
      return rc;
    if ((rc = trunk_ln_launch(x, 0, ly.mlp_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
    if ((rc = trunk_gemm_launch(EPI_GEGLU, xb, rows_i, ly.wi, I, H, t_rows, gg, I, nullptr, nullptr, 0, 0, stream)) != kOk)
      return rc;
    if ((rc = trunk_gemm_launch(EPI_RESID, gg, rows_i, ly.wo2, H, I, t_rows, x, H, nullptr, nullptr, 0, 0, stream)) != kOk)
      return rc;
  }
  // final_norm -> head.dense -> GELU -> head.norm  (ModernBertPredictionHead)
  if ((rc = trunk_ln_launch(x, 0, a->final_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
  if ((rc = trunk_gemm_launch(EPI_GELU, xb, rows_i, a->head_dense, H, H, t_rows, ab, H, nullptr, nullptr, 0, 0,
                              stream)) != kOk)
    return rc;
  if ((rc = trunk_ln_launch(ab, 1, a->head_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
  if (a->hidden_out != nullptr) {
    // optional: the packed head input [rows, H] bf16 (tests compare it with the reference trunk)
    SPLADE_CUDA_OK(cudaMemcpyAsync(a->hidden_out, xb, rows * H * 2, cudaMemcpyDeviceToDevice, stream));
  }
  if (a->meta_out != nullptr)
    SPLADE_CUDA_OK(cudaMemcpyAsync(a->meta_out, lay.meta, rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  if (a->plan_out != nullptr)
    SPLADE_CUDA_OK(cudaMemcpyAsync(a->plan_out, lay.plan, 5 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  // SPLADE head on the packed rows
  return head_fwd_packed_launch(xb, lay, a->dec_w, a->dec_bias, B, S, H, V, a->rep, a->argmax, stream);
}

}  // namespace splade
