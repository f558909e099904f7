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
#include <cstdlib>
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

// Register-resident row (H = 128 * kV4 floats: each lane holds kV4 float4s, every global access of the warp is a
// contiguous 512 B): one read of the row, statistics and the normalised output all from registers.
template <int kV4>
__device__ __forceinline__ void warp_layernorm_regs(const float4 (&v)[kV4], const float* __restrict__ gamma, float eps,
                                                    int lane, float* __restrict__ out_f32,
                                                    __nv_bfloat16* __restrict__ out_bf16) {
  constexpr int H = 128 * kV4;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kV4; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  const float mean = sum * (1.0f / H);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    var += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
  const float rstd = rsqrtf(var * (1.0f / H) + eps);
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int c4 = lane + 32 * i;  // float4 index inside the row
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x;
    y.y = (v[i].y - mean) * rstd * g.y;
    y.z = (v[i].z - mean) * rstd * g.z;
    y.w = (v[i].w - mean) * rstd * g.w;
    if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32)[c4] = y;
    if (out_bf16 != nullptr) {
      uint2 o;
      o.x = cvt_bf16x2(y.y, y.x);
      o.y = cvt_bf16x2(y.w, y.z);
      reinterpret_cast<uint2*>(out_bf16)[c4] = o;
    }
  }
}

template <int kV4>
__device__ __forceinline__ void load_row_regs(const float* __restrict__ src, int lane, float4 (&v)[kV4]) {
#pragma unroll
  for (int i = 0; i < kV4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(src) + lane + 32 * i);
}
template <int kV4>
__device__ __forceinline__ void load_row_regs(const __nv_bfloat16* __restrict__ src, int lane, float4 (&v)[kV4]) {
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(src) + lane + 32 * i);
    v[i].x = __uint_as_float(u.x << 16);
    v[i].y = __uint_as_float(u.x & 0xFFFF0000u);
    v[i].z = __uint_as_float(u.y << 16);
    v[i].w = __uint_as_float(u.y & 0xFFFF0000u);
  }
}

// kV4 = 0: generic three-pass version for hidden sizes that are not 256 / 768 / 1024
template <typename InT, int kV4>
__global__ void __launch_bounds__(256)
trunk_ln_kernel(const InT* __restrict__ in, const float* __restrict__ gamma, float eps,
                const int32_t* __restrict__ t_rows, int H, __nv_bfloat16* __restrict__ out) {
  const int T = __ldg(t_rows);
  const int lane = static_cast<int>(threadIdx.x & 31u);
  for (int row = blockIdx.x * 8 + static_cast<int>(threadIdx.x >> 5); row < T; row += gridDim.x * 8) {
    if (kV4 > 0) {
      float4 v[kV4 > 0 ? kV4 : 1];
      load_row_regs<(kV4 > 0 ? kV4 : 1)>(in + static_cast<size_t>(row) * H, lane, v);
      warp_layernorm_regs<(kV4 > 0 ? kV4 : 1)>(v, gamma, eps, lane, nullptr, out + static_cast<size_t>(row) * H);
    } else {
      warp_layernorm_row<InT>(in + static_cast<size_t>(row) * H, gamma, eps, H, lane, nullptr,
                              out + static_cast<size_t>(row) * H);
    }
  }
}

// x = LayerNorm(tok_embeddings[input_ids[doc, pos]]) for every packed column (replicas included: they carry the
// descriptor of the document's last token, so they stay exact duplicates of it through the whole trunk)
template <int kV4>
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
    const float* src = tok_emb + static_cast<size_t>(id) * H;
    if (kV4 > 0) {
      float4 v[kV4 > 0 ? kV4 : 1];
      load_row_regs<(kV4 > 0 ? kV4 : 1)>(src, lane, v);
      warp_layernorm_regs<(kV4 > 0 ? kV4 : 1)>(v, gamma, eps, lane, x + static_cast<size_t>(row) * H,
                                                xb + static_cast<size_t>(row) * H);
    } else {
      warp_layernorm_row<float>(src, gamma, eps, H, lane, x + static_cast<size_t>(row) * H,
                                xb + static_cast<size_t>(row) * H);
    }
  }
}

// ------------------------------------------------------------------------------------------------ RoPE
// In place on the q and k thirds of qkv [T, 3H]: x*cos + rotate_half(x)*sin in fp32 on the bf16 Linear output,
// rounded back to bf16 (apply_rotary_pos_emb, modeling_modernbert.py:205-228).  One thread per (row, head of q or
// k, 2 dims): a warp covers one 64-wide head = 128 contiguous bytes; qkv of a layer is L2-resident.
__global__ void __launch_bounds__(256)
trunk_rope_kernel(__nv_bfloat16* __restrict__ qkv, const float2* __restrict__ rope, const uint32_t* __restrict__ meta,
                  const int32_t* __restrict__ t_rows, int H) {
  const int T = __ldg(t_rows);
  const int heads2 = 2 * H / 64;  // q heads then k heads
  const long long total = static_cast<long long>(T) * heads2 * 16;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d2 = static_cast<int>(i & 15);  // dims 2*d2, 2*d2+1 and their partners 32 above
    const long long rh = i >> 4;
    const int hd = static_cast<int>(rh % heads2);
    const int row = static_cast<int>(rh / heads2);
    const int pos = static_cast<int>(__ldg(&meta[row]) & 0xFFFFu);
    const float2 c0 = __ldg(&rope[pos * 32 + 2 * d2]);
    const float2 c1 = __ldg(&rope[pos * 32 + 2 * d2 + 1]);
    uint32_t* plo = reinterpret_cast<uint32_t*>(qkv + static_cast<size_t>(row) * 3 * H + hd * 64 + 2 * d2);
    uint32_t* phi = plo + 16;
    const uint32_t lo = *plo, hi = *phi;
    const float a0 = __uint_as_float(lo << 16), a1 = __uint_as_float(lo & 0xFFFF0000u);
    const float b0 = __uint_as_float(hi << 16), b1 = __uint_as_float(hi & 0xFFFF0000u);
    *plo = cvt_bf16x2(a1 * c1.x - b1 * c1.y, a0 * c0.x - b0 * c0.y);
    *phi = cvt_bf16x2(b1 * c1.x + a1 * c1.y, b0 * c0.x + a0 * c0.y);
  }
}

// ------------------------------------------------------------------------------------------------ attention
constexpr int AT_BK = 64;   // keys per step
constexpr int AT_D = 64;    // head dim
constexpr int AT_LD = 72;   // smem row stride in bf16 (144 B: conflict-free 32-bit fragment loads)

constexpr int TA_TILE_Q = 128, TA_TILE_KV = 256;
__host__ __device__ inline bool attn_tc_eligible(int L, int window);

struct AttnParams {
  const __nv_bfloat16* qkv;  // [T, 3H]: q | k | v (heads of 64), RoPE already applied to q and k
  __nv_bfloat16* out;        // [T, H]
  const int32_t* len;
  const int32_t* off;
  const uint32_t* meta;
  int H;
  int window;  // < 0: global attention; otherwise |pos_q - pos_k| <= window
  float scale;
  int skip_tc_tiles;  // mma.sync kernel: 1 = tiles the tcgen05 kernel handles are skipped
};

// 2^x on the SFU in one instruction (fast_exp2() adds a denormal-range fix-up around it that softmax does not need)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 16-byte asynchronous global -> shared copy; `valid == false` writes zeros (source size 0)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}

// kWarps x 16 query rows per CTA (8 warps = 128 queries for documents, 4 warps for query-length batches).  K and V
// tiles of 64 keys are double-buffered in shared memory with cp.async (the next tile streams in while the tensor
// cores work on the current one), stay row-major, and every fragment comes from one ldmatrix.x4 (V through .trans,
// so there is no transposition pass).
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
trunk_attn_kernel(const AttnParams p) {
  constexpr int BQ = kWarps * 16;
  constexpr int NT = kWarps * 32;
  const int b = static_cast<int>(blockIdx.z);
  const int h = static_cast<int>(blockIdx.y);
  const int q0 = static_cast<int>(blockIdx.x) * BQ;
  const int L = __ldg(&p.len[b]);
  if (L == 0) return;
  const int L8 = (L + 7) & ~7;  // real tokens + replicas of the last one (they are queries, never keys)
  if (q0 >= L8) return;
  if (p.skip_tc_tiles && attn_tc_eligible(L, p.window)) return;
  const int base = __ldg(&p.off[b]);
  const int H3 = 3 * p.H;

  extern __shared__ __align__(16) uint8_t attn_smem[];
  typedef __nv_bfloat16 (*Tile)[AT_LD];
  Tile Qs = reinterpret_cast<Tile>(attn_smem);
  Tile Kbuf = reinterpret_cast<Tile>(attn_smem + BQ * AT_LD * 2);                     // [2][AT_BK] rows
  Tile Vbuf = reinterpret_cast<Tile>(attn_smem + (BQ + 2 * AT_BK) * AT_LD * 2);       // [2][AT_BK] rows
  int* kpos_buf = reinterpret_cast<int*>(attn_smem + (BQ + 4 * AT_BK) * AT_LD * 2);  // [2][AT_BK]

  const int tid = static_cast<int>(threadIdx.x);
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // visible key blocks: all of them (global) or the contiguous range that intersects the queries' windows
  const int n_kv = (L + AT_BK - 1) / AT_BK;
  int kb_lo = 0, kb_hi = n_kv;
  if (p.window >= 0) {
    const int qmin = static_cast<int>(__ldg(&p.meta[base + q0]) & 0xFFFFu);
    const int qmax = static_cast<int>(__ldg(&p.meta[base + min(q0 + BQ - 1, L8 - 1)]) & 0xFFFFu);
    while (kb_lo < kb_hi &&
           static_cast<int>(__ldg(&p.meta[base + min(kb_lo * AT_BK + AT_BK - 1, L - 1)]) & 0xFFFFu) < qmin - p.window)
      ++kb_lo;
    while (kb_hi > kb_lo &&
           static_cast<int>(__ldg(&p.meta[base + (kb_hi - 1) * AT_BK]) & 0xFFFFu) > qmax + p.window)
      --kb_hi;
  }

  auto issue_kv = [&](int kb, int buf) {
    const int kv0 = kb * AT_BK;
    Tile Ks = Kbuf + buf * AT_BK;
    Tile Vs = Vbuf + buf * AT_BK;
    for (int c = tid; c < AT_BK * 8; c += NT) {
      const int r = c >> 3, ch = c & 7;
      const bool ok = kv0 + r < L;
      const __nv_bfloat16* src = p.qkv + static_cast<size_t>(base + (ok ? kv0 + r : 0)) * H3 + h * AT_D + ch * 8;
      cp_async16(&Ks[r][ch * 8], src + p.H, ok);
      cp_async16(&Vs[r][ch * 8], src + 2 * p.H, ok);
    }
    if (tid < AT_BK) {
      const int key = kv0 + tid;
      kpos_buf[buf * AT_BK + tid] = key < L ? static_cast<int>(__ldg(&p.meta[base + key]) & 0xFFFFu) : 0x3FFFFFFF;
    }
  };

  for (int c = tid; c < BQ * 8; c += NT) {
    const int r = c >> 3, ch = c & 7;
    const bool ok = q0 + r < L8;
    cp_async16(&Qs[r][ch * 8], p.qkv + static_cast<size_t>(base + (ok ? q0 + r : 0)) * H3 + h * AT_D + ch * 8, ok);
  }
  if (kb_lo < kb_hi) issue_kv(kb_lo, 0);
  cp_async_commit();

  const bool warp_active = q0 + warp * 16 < L8;  // warps whose 16 rows are all past the document only help loading
  uint32_t qa[4][4];
  int qpos[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int row = min(q0 + warp * 16 + g + 8 * e, L8 - 1);
    qpos[e] = static_cast<int>(__ldg(&p.meta[base + row]) & 0xFFFFu);
  }

  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;

  const float sl2 = p.scale * 1.4426950408889634f;  // softmax in base 2: exp(x) = exp2(x * log2 e)
  for (int kb = kb_lo; kb < kb_hi; ++kb) {
    const int buf = (kb - kb_lo) & 1;
    const int kv0 = kb * AT_BK;
    cp_async_wait_all();
    __syncthreads();  // tile kb has landed for everyone, and everyone is done with the other buffer
    if (kb + 1 < kb_hi) issue_kv(kb + 1, buf ^ 1);
    cp_async_commit();
    if (kb == kb_lo) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(qa[ks], &Qs[warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + (lane >> 4) * 8]);
    }
    if (!warp_active) continue;
    Tile Ks = Kbuf + buf * AT_BK;
    Tile Vs = Vbuf + buf * AT_BK;
    const int* kpos = kpos_buf + buf * AT_BK;

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {  // two 16-dim steps per ldmatrix.x4
        uint32_t kf[4];
        ldmatrix_x4(kf, &Ks[nt * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
        mma_bf16_16816(s[nt], qa[2 * kp], kf[0], kf[1]);
        mma_bf16_16816(s[nt], qa[2 * kp + 1], kf[2], kf[3]);
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
        s[nt][e] = ok ? s[nt][e] * sl2 : -INFINITY;
        mx[r] = fmaxf(mx[r], s[nt][e]);
      }
    }
    float corr[2], moff[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(m_run[r], mx[r]);
      moff[r] = mnew == -INFINITY ? 0.f : mnew;  // nothing visible yet: keep exp2() finite
      corr[r] = fast_exp2(m_run[r] - moff[r]);       // exp2(-inf) = 0 on the first visible block
      m_run[r] = mnew;
      l_run[r] *= corr[r];
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = fast_exp2(s[nt][0] - moff[0]);
      const float p1 = fast_exp2(s[nt][1] - moff[0]);
      const float p2 = fast_exp2(s[nt][2] - moff[1]);
      const float p3 = fast_exp2(s[nt][3] - moff[1]);
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
    }
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {  // two 8-dim output tiles per ldmatrix.x4.trans
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, &Vs[ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)][dp * 16 + (lane >> 4) * 8]);
        mma_bf16_16816(o[2 * dp], pa[ks], vf[0], vf[1]);
        mma_bf16_16816(o[2 * dp + 1], pa[ks], vf[2], vf[3]);
      }
    }
  }
  if (!warp_active) return;

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

// ------------------------------------------------------------------------------------------------ attention, tcgen05
// The same attention on the 5th-generation tensor cores, for every (128 queries, head, document) tile whose visible
// keys fit one 256-key chunk: all of a document of <= 256 tokens (global layers), or the +-window neighbourhood of
// the tile (sliding-window layers, 128 + 2*64 = 256 keys for ANY document length).
//   S[128 q, N keys] = Q K^T      tcgen05.mma, Q and K tiles TMA-loaded straight out of qkv (K-major, 128B swizzle),
//                                 fp32 scores in TMEM (256 columns)
//   softmax                       thread r owns query row r = TMEM lane r: two sweeps over its scores (max, then
//                                 exp2 / sum), P written as bf16 into shared memory in the K-major swizzled layout
//                                 the next MMA reads (over the dead Q / K tiles)
//   O[128 q, 64] = P V            tcgen05.mma with V as the MN-major B operand (the TMA-loaded [key][dim] tile as
//                                 it is: dims contiguous), accumulator in the TMEM columns S no longer needs
// 96 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM, so one tile's loads / softmax overlap the
// other's MMAs; 256 threads, two per query row (each sweeps half of the row's 32-column chunks).  Tiles that do not qualify (global attention over > 256 tokens) are left to the mma.sync kernel.
constexpr int TA_BQ = 128;
constexpr int TA_BKV = 256;
constexpr int TA_Q_OFF = 0;                    // 128 rows x 128 B
constexpr int TA_K_OFF = 16384;                // 256 rows x 128 B
constexpr int TA_V_OFF = 65536;                // 256 rows x 128 B (16 KB of padding before it: P needs 64 KB)
constexpr int TA_BAR_OFF = 98304;
constexpr int TA_KPOS_OFF = TA_BAR_OFF + 64;
constexpr int TA_XCHG_OFF = TA_KPOS_OFF + TA_BKV * 4;
constexpr int TA_SMEM_BYTES = TA_XCHG_OFF + 2 * 128 * 4 + 1024;
constexpr int TA_TMEM_COLS = 256;

__host__ __device__ inline bool attn_tc_eligible(int L, int window) {
  return window < 0 ? L <= TA_BKV : TA_BQ + 2 * window <= TA_BKV;
}

// MN-major operand tile with 128-byte swizzle: 64 contiguous MN elements (128 B) per K row, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;  // LBO: next 64-wide MN atom (unused: N = 64 is one atom)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;  // SBO: next group of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(256, 2)
trunk_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const AttnParams p) {
  const int b = static_cast<int>(blockIdx.z);
  const int h = static_cast<int>(blockIdx.y);
  const int q0 = static_cast<int>(blockIdx.x) * TA_BQ;
  const int L = __ldg(&p.len[b]);
  if (L == 0) return;
  const int L8 = (L + 7) & ~7;
  if (q0 >= L8 || !attn_tc_eligible(L, p.window)) return;  // CTA-uniform, before anything is allocated
  const int base = __ldg(&p.off[b]);
  const int kstart = p.window < 0 ? 0 : max(0, q0 - p.window);
  const int kend = p.window < 0 ? L : min(L, q0 + TA_BQ + p.window);
  const int nkeys = kend - kstart;  // 1..256

  extern __shared__ uint8_t attn_tc_smem[];
  const uint32_t raw_addr = smem_u32(attn_tc_smem);
  uint8_t* smem = attn_tc_smem + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + TA_BAR_OFF);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_load + 2);
  int* kpos = reinterpret_cast<int*>(smem + TA_KPOS_OFF);
  float* xchg = reinterpret_cast<float*>(smem + TA_XCHG_OFF);  // [2][128]: row maxima, then row sums, per half

  // 256 threads: thread t and thread t + 128 share query row (t & 127) = TMEM lane; `half` 0 takes the low 32-column
  // chunks of the row, half 1 the high ones (two warps per SM sub-partition instead of one: the softmax is
  // latency-bound, not issue-bound)
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const uint32_t row_t = tid & 127u;
  const uint32_t half = tid >> 7;
  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    // the loads do not depend on anything the other threads set up: start them before the block-wide sync
    mbar_arrive_expect_tx(bar_load, TA_BQ * 128 + 2 * TA_BKV * 128);
    tma_load_2d(smem + TA_Q_OFF, &tmap_q, bar_load, h * AT_D, base + q0);
    tma_load_2d(smem + TA_K_OFF, &tmap_kv, bar_load, p.H + h * AT_D, base + kstart);
    tma_load_2d(smem + TA_V_OFF, &tmap_kv, bar_load, 2 * p.H + h * AT_D, base + kstart);
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TA_TMEM_COLS);
  {
    const int key = kstart + static_cast<int>(tid);
    kpos[tid] = key < kend ? static_cast<int>(__ldg(&p.meta[base + key]) & 0xFFFFu) : 0x3FFFFFFF;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  if (tid == 0) {
    mbar_wait(bar_load, 0);
    tcgen05_fence_after();
    // S = Q K^T: M = 128 queries, N = keys rounded up to 16, K = 64 dims in four 16-wide steps
    const uint32_t n16 = static_cast<uint32_t>(max(16, (nkeys + 15) & ~15));
    const uint32_t idesc = make_idesc_bf16_f32(TA_BQ, n16);
    const uint32_t q_addr = smem_u32(smem + TA_Q_OFF), k_addr = smem_u32(smem + TA_K_OFF);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem, make_smem_desc_sw128(q_addr + k * 32), make_smem_desc_sw128(k_addr + k * 32), idesc,
                k != 0 ? 1u : 0u);
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tcgen05_fence_after();

  // ---- softmax: one query row per thread pair.  Positions grow with the column index, so the keys a row may see
  // are a contiguous column range [lo, hi): found once per row (two binary searches over the staged key positions
  // for the sliding window), and every 32-column chunk is then, for the whole warp, fully visible (no mask
  // arithmetic at all: the common case), invisible (skipped; its P columns are zero) or mixed (two integer compares
  // per score).
  const int qrow = min(q0 + static_cast<int>(row_t), L8 - 1);
  const int qpos = static_cast<int>(__ldg(&p.meta[base + qrow]) & 0xFFFFu);
  int lo = 0, hi = nkeys;
  if (p.window >= 0) {
    int l = 0, r = nkeys;  // first key with pos >= qpos - window
    while (l < r) {
      const int m = (l + r) >> 1;
      if (kpos[m] < qpos - p.window) l = m + 1; else r = m;
    }
    lo = l;
    r = nkeys;             // first key with pos > qpos + window
    while (l < r) {
      const int m = (l + r) >> 1;
      if (kpos[m] <= qpos + p.window) l = m + 1; else r = m;
    }
    hi = l;
  }
  const uint32_t t_lane = tmem + (((warp & 3u) * 32u) << 16);
  const int nchunks = (nkeys + 31) >> 5;
  const int c_mid = (nchunks + 1) >> 1;
  const int c_beg = half == 0 ? 0 : c_mid, c_end = half == 0 ? c_mid : nchunks;
  const float sl2 = p.scale * 1.4426950408889634f;
  float mx = -INFINITY;
  for (int c = c_beg; c < c_end; ++c) {
    const int c0 = 32 * c;
    if (__all_sync(0xffffffffu, hi <= c0 || lo >= c0 + 32)) continue;
    const bool full = __all_sync(0xffffffffu, lo <= c0 && hi >= c0 + 32);
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(c0), r);
    tmem_ld_wait();
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = (c0 + j >= lo && c0 + j < hi) ? fmaxf(mx, __uint_as_float(r[j])) : mx;
    }
  }
  xchg[half * 128u + row_t] = mx;
  __syncthreads();
  mx = fmaxf(mx, xchg[(half ^ 1u) * 128u + row_t]);
  if (mx == -INFINITY) mx = 0.f;  // rows past the document (never stored)
  const float moff = mx * sl2;
  float sum = 0.f;
  // P overwrites the Q / K tiles (the first MMA has completed): [64-key block][row][128 B], chunk c of row r at
  // chunk position c ^ (r & 7)
  for (int c = c_beg; c < c_end; ++c) {
    const int c0 = 32 * c;
    uint8_t* prow = smem + (c >> 1) * 16384 + row_t * 128u;
    if (__all_sync(0xffffffffu, hi <= c0 || lo >= c0 + 32)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + i);
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (row_t & 7u)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
      continue;
    }
    const bool full = __all_sync(0xffffffffu, lo <= c0 && hi >= c0 + 32);
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(c0), r);
    tmem_ld_wait();
    float pv[32];
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        pv[j] = fast_exp2(fmaf(__uint_as_float(r[j]), sl2, -moff));
        sum += pv[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        pv[j] = (c0 + j >= lo && c0 + j < hi) ? fast_exp2(fmaf(__uint_as_float(r[j]), sl2, -moff)) : 0.f;
        sum += pv[j];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 v;
      v.x = cvt_bf16x2(pv[8 * i + 1], pv[8 * i + 0]);
      v.y = cvt_bf16x2(pv[8 * i + 3], pv[8 * i + 2]);
      v.z = cvt_bf16x2(pv[8 * i + 5], pv[8 * i + 4]);
      v.w = cvt_bf16x2(pv[8 * i + 7], pv[8 * i + 6]);
      const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + i);
      *reinterpret_cast<uint4*>(prow + ((chunk ^ (row_t & 7u)) << 4)) = v;
    }
  }
  // The second MMA reads ceil(nkeys / 16) 16-key steps: all inside written P chunks (their tail columns are zero).
  // The V rows of that tail belong to other documents or to never-written workspace: 0 * NaN would poison O, so
  // they are cleared (whole 128-byte rows: the swizzle only permutes chunks inside a row).
  {
    const int vrow = nkeys + static_cast<int>(tid >> 3);
    if (tid < 128 && vrow < ((nkeys + 15) & ~15))
      *reinterpret_cast<uint4*>(smem + TA_V_OFF + vrow * 128 + (tid & 7u) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();                      // both halves have read the exchanged maxima
  xchg[half * 128u + row_t] = sum;      // row sums, exchanged after the next sync
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();

  if (tid == 0) {
    tcgen05_fence_after();
    // O = P V: M = 128, N = 64 dims (V MN-major), K = keys in 16-wide steps; accumulator reuses TMEM columns 0..63
    const uint32_t idesc = make_idesc_bf16_f32(TA_BQ, AT_D) | (1u << 16);
    const uint32_t p_addr = smem_u32(smem), v_addr = smem_u32(smem + TA_V_OFF);
    const int ksteps = (nkeys + 15) >> 4;
    for (int ks = 0; ks < ksteps; ++ks)
      umma_bf16(tmem, make_smem_desc_sw128(p_addr + (ks >> 2) * 16384 + (ks & 3) * 32),
                make_smem_desc_sw128_mn(v_addr + ks * 2048), idesc, ks != 0 ? 1u : 0u);
    umma_commit(bar_mma);
  }
  sum += xchg[(half ^ 1u) * 128u + row_t];
  mbar_wait(bar_mma, 1);
  tcgen05_fence_after();

  // each half stores 32 of the 64 output dims of its row
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  const int row = q0 + static_cast<int>(row_t);
  {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_lane + 32u * half, r);
    tmem_ld_wait();
    if (row < L8) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(base + row) * p.H + h * AT_D + 32 * half);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v;
        v.x = cvt_bf16x2(__uint_as_float(r[8 * i + 1]) * inv, __uint_as_float(r[8 * i + 0]) * inv);
        v.y = cvt_bf16x2(__uint_as_float(r[8 * i + 3]) * inv, __uint_as_float(r[8 * i + 2]) * inv);
        v.z = cvt_bf16x2(__uint_as_float(r[8 * i + 5]) * inv, __uint_as_float(r[8 * i + 4]) * inv);
        v.w = cvt_bf16x2(__uint_as_float(r[8 * i + 7]) * inv, __uint_as_float(r[8 * i + 6]) * inv);
        dst[i] = v;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, TA_TMEM_COLS);
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
  const int grid = 8 * sm_count();
  __nv_bfloat16* xb16 = static_cast<__nv_bfloat16*>(xb);
  const int32_t* t_rows = &lay.plan->T;
  if (H == 768) {
    trunk_embed_ln_kernel<6><<<grid, 256, 0, stream>>>(input_ids, tok_emb, gamma, eps, t_rows, lay.meta, S, H, x, xb16);
  } else if (H == 256) {
    trunk_embed_ln_kernel<2><<<grid, 256, 0, stream>>>(input_ids, tok_emb, gamma, eps, t_rows, lay.meta, S, H, x, xb16);
  } else if (H == 1024) {
    trunk_embed_ln_kernel<8><<<grid, 256, 0, stream>>>(input_ids, tok_emb, gamma, eps, t_rows, lay.meta, S, H, x, xb16);
  } else {
    trunk_embed_ln_kernel<0><<<grid, 256, 0, stream>>>(input_ids, tok_emb, gamma, eps, t_rows, lay.meta, S, H, x, xb16);
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

template <typename InT>
int ln_launch_typed(const InT* in, const float* gamma, float eps, const int32_t* t_rows, int H, __nv_bfloat16* out,
                    cudaStream_t stream) {
  const int grid = 8 * sm_count();
  if (H == 768) {
    trunk_ln_kernel<InT, 6><<<grid, 256, 0, stream>>>(in, gamma, eps, t_rows, H, out);
  } else if (H == 256) {
    trunk_ln_kernel<InT, 2><<<grid, 256, 0, stream>>>(in, gamma, eps, t_rows, H, out);
  } else if (H == 1024) {
    trunk_ln_kernel<InT, 8><<<grid, 256, 0, stream>>>(in, gamma, eps, t_rows, H, out);
  } else {
    trunk_ln_kernel<InT, 0><<<grid, 256, 0, stream>>>(in, gamma, eps, t_rows, H, out);
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int trunk_ln_launch(const void* in, int in_is_bf16, const float* gamma, float eps, const int32_t* t_rows,
                    int rows_cap, int H, void* out_bf16, cudaStream_t stream) {
  (void)rows_cap;
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_bf16);
  if (in_is_bf16) return ln_launch_typed(static_cast<const __nv_bfloat16*>(in), gamma, eps, t_rows, H, out, stream);
  return ln_launch_typed(static_cast<const float*>(in), gamma, eps, t_rows, H, out, stream);
}

int trunk_rope_launch(void* qkv, const float2* rope, const VarlenLayout& lay, int H, cudaStream_t stream) {
  trunk_rope_kernel<<<8 * sm_count(), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(qkv), rope, lay.meta,
                                                        &lay.plan->T, H);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

// SPLADE_B200_ATTN_MMA_SYNC=1 pins the mma.sync kernel (A/B timing, fallback while debugging)
static bool attn_tc_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("SPLADE_B200_ATTN_MMA_SYNC");
    return !(e != nullptr && e[0] == '1');
  }();
  return on;
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
  // tcgen05 kernel: every tile whose visible keys fit 256 (decided per document on the device)
  // (query-length batches, S <= 64, stay on the 64-row mma.sync kernel: a 128-row tcgen05 tile per 10-token query
  // is all fixed latency - measured 5.9 vs 4.5 ms for the whole encode of 512 queries)
  const bool tc = attn_tc_enabled() && s8 > 64 && lay.rows >= TA_BKV && (window < 0 || TA_BQ + 2 * window <= TA_BKV);
  const bool tc_covers_all = tc && attn_tc_eligible(s8, window);  // no document can be longer than S
  if (tc) {
    CUtensorMap tmap_q, tmap_kv;
    const uint64_t dims[2] = {static_cast<uint64_t>(3 * H), static_cast<uint64_t>(lay.rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(3 * H) * 2};
    const uint32_t box_q[2] = {AT_D, TA_BQ};
    const uint32_t box_kv[2] = {AT_D, TA_BKV};
    int rc = encode_tmap_bf16(&tmap_q, qkv, 2, dims, strides, box_q);
    if (rc != kOk) return rc;
    rc = encode_tmap_bf16(&tmap_kv, qkv, 2, dims, strides, box_kv);
    if (rc != kOk) return rc;
    SPLADE_CUDA_OK(cudaFuncSetAttribute(trunk_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TA_SMEM_BYTES));
    trunk_attn_tc_kernel<<<dim3((s8 + TA_BQ - 1) / TA_BQ, n_heads, B), 256, TA_SMEM_BYTES, stream>>>(tmap_q, tmap_kv, p);
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  if (!tc_covers_all) {
    p.skip_tc_tiles = tc ? 1 : 0;
    if (s8 <= 64) {
      constexpr int smem = (64 + 4 * AT_BK) * AT_LD * 2 + 2 * AT_BK * 4;
      SPLADE_CUDA_OK(cudaFuncSetAttribute(trunk_attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      trunk_attn_kernel<4><<<dim3(1, n_heads, B), 128, smem, stream>>>(p);
    } else {
      constexpr int smem = (128 + 4 * AT_BK) * AT_LD * 2 + 2 * AT_BK * 4;
      SPLADE_CUDA_OK(cudaFuncSetAttribute(trunk_attn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      trunk_attn_kernel<8><<<dim3((s8 + 127) / 128, n_heads, B), 256, smem, stream>>>(p);
    }
    SPLADE_CUDA_OK(cudaGetLastError());
  }
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

  (void)rope_g;
  (void)rope_l;
  if ((rc = trunk_embed_ln_launch(a->input_ids, a->tok_emb, a->emb_norm, a->norm_eps, lay, S, H, x, xb, stream)) != kOk)
    return rc;

  for (int l = 0; l < a->n_layers; ++l) {
    const splade_trunk_layer& ly = a->layers[l];
    if (ly.attn_norm != nullptr) {
      if ((rc = trunk_ln_launch(x, 0, ly.attn_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
    }
    // Wqkv with RoPE on the q / k heads fused into the epilogue (the separate trunk_rope_kernel cost 25 us a layer)
    if ((rc = trunk_gemm_launch(EPI_QKV_ROPE, xb, rows_i, ly.wqkv, 3 * H, H, t_rows, qkv, 3 * H, lay.meta,
                                ly.is_global ? a->global_theta : a->local_theta, 2 * H, stream)) != kOk)
      return rc;
    if ((rc = trunk_attn_launch(qkv, ab, lay, B, S, H, a->n_heads, ly.is_global ? -1 : a->local_window, stream)) != kOk)
      return rc;
    if ((rc = trunk_gemm_launch(EPI_RESID, ab, rows_i, ly.wo, H, H, t_rows, x, H, nullptr, 0.f, 0, stream)) != kOk)
      return rc;
    if ((rc = trunk_ln_launch(x, 0, ly.mlp_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
    if ((rc = trunk_gemm_launch(EPI_GEGLU, xb, rows_i, ly.wi, I, H, t_rows, gg, I, nullptr, 0.f, 0, stream)) != kOk)
      return rc;
    if ((rc = trunk_gemm_launch(EPI_RESID, gg, rows_i, ly.wo2, H, I, t_rows, x, H, nullptr, 0.f, 0, stream)) != kOk)
      return rc;
  }
  // final_norm -> head.dense -> GELU -> head.norm  (ModernBertPredictionHead)
  if ((rc = trunk_ln_launch(x, 0, a->final_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
  if ((rc = trunk_gemm_launch(EPI_GELU, xb, rows_i, a->head_dense, H, H, t_rows, ab, H, nullptr, 0.f, 0, stream)) != kOk)
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
