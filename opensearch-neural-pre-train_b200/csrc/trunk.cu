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
  int debug;          // timing ablations of the pipelined tcgen05 kernel (SPLADE_B200_ATTN_DEBUG; results are WRONG)
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


// ------------------------------------------------------------------------------------------------ attention, tcgen05,
// persistent and software-pipelined (the default).  Same tiles and arithmetic as trunk_attn_tc_kernel above, but
//   * ONE CTA per SM loops over its tiles: barriers, the TMEM allocation (all 512 columns) and the tensor maps are set
//     up once; a producer warp TMA-loads the Q / K / V tiles of tile n+1 (and n+2) into a two-stage ring while tile n
//     is in its softmax, and an MMA warp issues S(n+1) = Q K^T into the other half of TMEM as soon as they land:
//     loads, both MMAs and the per-tile metadata reads are off the softmax warps' critical path;
//   * two softmax groups of 256 threads (two threads per query row) alternate tiles, each with its own S buffer
//     (256 TMEM columns), so one group's TMEM-load / barrier latencies are filled by the other's arithmetic;
//   * P never touches shared memory: each thread packs its 32 probabilities of a chunk to bf16 pairs and stores
//     them with tcgen05.st over S columns it has ALREADY consumed; O = P V runs as tcgen05.mma with the A operand
//     in tensor memory (lane = query row, column j = keys 2j, 2j+1 - checked by tools/probes/tmem_a_operand.cu)
//     and V as the MN-major B operand.  This is what frees the 64 KB of P staging for the second Q/K/V stage.
// TMEM map of a slot (256 columns): S chunk c (32 keys) at [32c, 32c+32); the bf16 P of chunk c (16 columns) at 32c
// for c < 6, chunk 6 at 16, chunk 7 at 48 (the consumed upper halves of chunks 0 and 1 - same thread: a thread owns
// the even or the odd chunks of its row); O (64 columns) at 192, over chunks 6 / 7 once every thread is through.
constexpr int TP_STAGE_BYTES = 16384 + 32768 + 32768;  // Q | K | V
constexpr int TP_Q_OFF = 0;
constexpr int TP_K_OFF = 16384;
constexpr int TP_V_OFF = 49152;
constexpr int TP_QK_BYTES = 16384 + 32768;
constexpr int TP_V_BYTES = 32768;
constexpr int TP_BAR_OFF = 2 * TP_STAGE_BYTES;
constexpr int TP_KPOS_OFF = TP_BAR_OFF + 128;               // [4 sets][256] key positions
constexpr int TP_XCHG_OFF = TP_KPOS_OFF + 4 * TA_BKV * 4;    // [2 groups][max | sum][2 halves][128]
constexpr int TP_TILES_OFF = TP_XCHG_OFF + 2 * 2 * 2 * 128 * 4;
constexpr int TP_MAX_TILES = 256;                            // tiles per CTA the table holds
constexpr int TP_OST_OFF = TP_TILES_OFF + TP_MAX_TILES * 16 + 16;   // [4 sets][32 rows][128 B] output staging
constexpr int TP_META_OFF = TP_OST_OFF + 4 * 4096;                  // [512 threads][4] prefetched token descriptors
constexpr int TP_SMEM_BYTES = TP_META_OFF + 512 * 16 + 1024;
constexpr int TP_THREADS = 512 + 64;

// One (128 queries, head, document) tile of this CTA's list, decoded once at kernel start
struct __align__(16) AttnTile {
  int32_t q_row;     // first packed row of the query tile (base + q0)
  int32_t k_row;     // first packed row of the visible keys (base + kstart)
  int32_t rows_h;    // (valid query rows in the tile << 8) | head
  int32_t nkeys;     // visible keys, 1..256
};

__device__ __forceinline__ uint32_t attn_p_col(int chunk) {
  return chunk < 6 ? 32u * static_cast<uint32_t>(chunk) : (chunk == 6 ? 16u : 48u);
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(TP_THREADS, 1)
trunk_attn_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                      const AttnParams p, int n_tiles, int nqt, int n_heads) {
  extern __shared__ uint8_t attn_tp_smem[];
  const uint32_t raw_addr = smem_u32(attn_tp_smem);
  uint8_t* smem = attn_tp_smem + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TP_BAR_OFF);
  uint64_t* bar_full_qk = bars;      // [2] Q and K of the stage have landed                   (producer -> MMA)
  uint64_t* bar_full_v = bars + 2;   // [2] V of the stage has landed                          (producer -> MMA, group)
  uint64_t* bar_s_ready = bars + 4;  // [2] S = Q K^T is in the slot's TMEM; Q / K are dead    (MMA -> group, producer)
  uint64_t* bar_p_ready = bars + 6;  // [2] P is in TMEM, the V tail rows are cleared          (256 threads -> MMA)
  uint64_t* bar_o_ready = bars + 8;  // [2] O = P V done: O readable, V dead                   (MMA -> group, producer)
  uint64_t* bar_s_free = bars + 10;  // [2] the group has read O: the slot's TMEM is reusable  (256 threads -> MMA)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);
  AttnTile* tiles = reinterpret_cast<AttnTile*>(smem + TP_TILES_OFF);
  int* n_valid_smem = reinterpret_cast<int*>(smem + TP_TILES_OFF + TP_MAX_TILES * 16);

  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full_qk[s], 1);
      mbar_init(&bar_full_v[s], 1);
      mbar_init(&bar_s_ready[s], 1);
      mbar_init(&bar_p_ready[s], 16);   // one arrival per softmax warp
      mbar_init(&bar_o_ready[s], 1);
      mbar_init(&bar_s_free[s], 16);
    }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_ptr_smem, 512);
  if (warp == 16) {
    // this CTA's tile list (t = blockIdx.x, + gridDim.x, ...), invalid tiles (past a document's end, documents the
    // mma.sync kernel handles) dropped: one warp, 32 candidates per round, order preserved
    const uint32_t lane = tid & 31u;
    int count = 0;
    for (int k0 = 0; k0 < TP_MAX_TILES; k0 += 32) {
      const int t = static_cast<int>(blockIdx.x) + (k0 + static_cast<int>(lane)) * static_cast<int>(gridDim.x);
      bool valid = false;
      AttnTile ti{};
      if (t < n_tiles) {
        const int qt = t % nqt, bh = t / nqt;
        const int b = bh / n_heads, h = bh - b * n_heads;
        const int q0 = qt * TA_BQ;
        const int L = __ldg(&p.len[b]);
        const int L8 = (L + 7) & ~7;
        valid = L > 0 && q0 < L8 && attn_tc_eligible(L, p.window);
        if (valid) {
          const int base = __ldg(&p.off[b]);
          const int kstart = p.window < 0 ? 0 : max(0, q0 - p.window);
          const int kend = p.window < 0 ? L : min(L, q0 + TA_BQ + p.window);
          ti.q_row = base + q0;
          ti.k_row = base + kstart;
          ti.rows_h = (min(TA_BQ, L8 - q0) << 8) | h;
          ti.nkeys = kend - kstart;
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, valid);
      if (valid) tiles[count + __popc(m & ((1u << lane) - 1u))] = ti;
      count += __popc(m);
      if (__all_sync(0xffffffffu, t >= n_tiles)) break;
    }
    if (lane == 0) *n_valid_smem = count;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  const int n_valid = *n_valid_smem;

  // 18 warps: 0..15 softmax (a thread keeps its 64 scores in registers: 112 registers per thread at 576 threads),
  // 16 producer, 17 MMA issuer
  if (warp == 16) {
    // ---------------------------------------------------------------- producer: Q/K two tiles ahead, V one behind them
    if ((tid & 31u) == 0) {
      auto load_v = [&](int m) {
        const int s = m & 1;
        if (m >= 2) mbar_wait_quiet(&bar_o_ready[s], static_cast<uint32_t>(((m >> 1) - 1) & 1));  // P V of tile m-2 done
        const AttnTile ti = tiles[m];
        if (p.debug & 16) {
          mbar_arrive(&bar_full_v[s]);
          return;
        }
        mbar_arrive_expect_tx(&bar_full_v[s], TP_V_BYTES);
        tma_load_2d(smem + s * TP_STAGE_BYTES + TP_V_OFF, &tmap_kv, &bar_full_v[s], 2 * p.H + (ti.rows_h & 0xFF) * AT_D,
                    ti.k_row);
      };
      for (int n = 0; n < n_valid; ++n) {
        const int s = n & 1;
        if (n >= 2) mbar_wait_quiet(&bar_s_ready[s], static_cast<uint32_t>(((n >> 1) - 1) & 1));  // Q K^T of tile n-2 done
        const AttnTile ti = tiles[n];
        const int hcol = (ti.rows_h & 0xFF) * AT_D;
        uint8_t* st = smem + s * TP_STAGE_BYTES;
        if (p.debug & 32) {
          mbar_arrive(&bar_full_qk[s]);
        } else {
          mbar_arrive_expect_tx(&bar_full_qk[s], TP_QK_BYTES);
          tma_load_2d(st + TP_Q_OFF, &tmap_q, &bar_full_qk[s], hcol, ti.q_row);
          tma_load_2d(st + TP_K_OFF, &tmap_kv, &bar_full_qk[s], p.H + hcol, ti.k_row);
        }
        if (n >= 1) load_v(n - 1);
      }
      if (n_valid >= 1) load_v(n_valid - 1);
    }
  } else if (warp == 17) {
    // ---------------------------------------------------------------- MMA issuer: whichever of S(n1) = Q K^T and
    // O(n2) = P V (n2 < n1) has its inputs first
    if ((tid & 31u) == 0) {
      const uint32_t idesc_pv = make_idesc_bf16_f32(TA_BQ, AT_D) | (1u << 16);  // B (= V) is MN-major
      int n1 = 0, n2 = 0;
      long long t0 = clock64();
      while (n2 < n_valid) {
        bool progress = false;
        if (n2 < n1) {
          const int s = n2 & 1;
          const uint32_t ph = static_cast<uint32_t>((n2 >> 1) & 1);
          if (mbar_test_wait(smem_u32(&bar_p_ready[s]), ph) && mbar_test_wait(smem_u32(&bar_full_v[s]), ph)) {
            tcgen05_fence_after();
            const uint32_t slot = tmem + 256u * static_cast<uint32_t>(s);
            const uint32_t v_addr = smem_u32(smem + s * TP_STAGE_BYTES + TP_V_OFF);
            const int ksteps = (p.debug & 4) ? 1 : (tiles[n2].nkeys + 15) >> 4;
            for (int ks = 0; ks < ksteps; ++ks)
              umma_bf16_ts(slot + 192u, slot + 8u * static_cast<uint32_t>(ks),
                           make_smem_desc_sw128_mn(v_addr + ks * 2048), idesc_pv, ks != 0 ? 1u : 0u);
            umma_commit(&bar_o_ready[s]);
            ++n2;
            progress = true;
          }
        }
        if (n1 < n_valid && n1 - n2 < 2) {
          const int s = n1 & 1;
          const uint32_t ph = static_cast<uint32_t>((n1 >> 1) & 1);
          if (mbar_test_wait(smem_u32(&bar_full_qk[s]), ph) &&
              (n1 < 2 || mbar_test_wait(smem_u32(&bar_s_free[s]), ph ^ 1u))) {
            tcgen05_fence_after();
            const uint32_t n16 = static_cast<uint32_t>(max(16, (tiles[n1].nkeys + 15) & ~15));
            const uint32_t idesc = make_idesc_bf16_f32(TA_BQ, n16);
            const uint32_t q_addr = smem_u32(smem + s * TP_STAGE_BYTES + TP_Q_OFF);
            const uint32_t k_addr = smem_u32(smem + s * TP_STAGE_BYTES + TP_K_OFF);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem + 256u * static_cast<uint32_t>(s), make_smem_desc_sw128(q_addr + k * 32),
                        make_smem_desc_sw128(k_addr + k * 32), idesc, k != 0 ? 1u : 0u);
            umma_commit(&bar_s_ready[s]);
            ++n1;
            progress = true;
          }
        }
        if (progress) t0 = clock64();
        else if (clock64() - t0 > SPLADE_MBAR_TIMEOUT_CYCLES) __trap();
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: 512 threads, FOUR per query row
    // (TMEM lane = row; warp wg reads lane quarter wg & 3 and owns the 32-key chunks cq = wg >> 2 and cq + 4 of its
    // rows).  A thread's 64 scores stay in registers from the TMEM load to the bf16 P store: S is read ONCE (TMEM
    // reads, 64 B/clk/SM, were the floor of the two-pass version).
    const uint32_t wg = warp;
    const uint32_t gt = tid;
    const uint32_t row_t = (wg & 3u) * 32u + (tid & 31u);
    const int cq = static_cast<int>(wg >> 2);
    // The four warps that share a lane quarter (32 query rows) form a SET: everything they exchange (row maxima and
    // sums, the staged key positions) is per set, with a 128-thread named barrier - so the four sets drift apart and
    // one set's TMEM loads overlap another's exponentials instead of all sixteen warps moving in lock step
    const uint32_t set = wg & 3u;
    const uint32_t st_t = (wg >> 2) * 32u + (tid & 31u);  // 0..127 inside the set
    const uint32_t set_bar = 1u + set;
    int* kpos = reinterpret_cast<int*>(smem + TP_KPOS_OFF) + set * TA_BKV;
    uint8_t* ostage = smem + TP_OST_OFF + set * 4096u;  // the set's 32 output rows x 128 B
    float* xmax = reinterpret_cast<float*>(smem + TP_XCHG_OFF);  // [4][128]
    float* xsum = xmax + 512;                                    // [4][128]
    const uint32_t lane_off = ((wg & 3u) * 32u) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    // positions of this thread's query row and of "its" key (thread gt < 256 stages key gt), fetched one tile ahead
    // They are fetched one tile ahead with cp.async straight into a per-thread shared-memory slot: a prefetch into
    // REGISTERS was spilled by the compiler at once, and the spill store waited for the global load (10 % of all
    // stall samples, at the start of every tile).
    uint32_t* meta_slot = reinterpret_cast<uint32_t*>(smem + TP_META_OFF) + tid * 4u;
    auto cp_async4 = [](uint32_t* dst, const uint32_t* src) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    };
    auto fetch_meta = [&](int n) {
      if (n < n_valid) {
        const AttnTile ti = tiles[n];
        cp_async4(meta_slot, &p.meta[ti.q_row + min(static_cast<int>(row_t), (ti.rows_h >> 8) - 1)]);
        if (p.window >= 0) {  // thread st_t of the set stages keys st_t and st_t + 128
          if (static_cast<int>(st_t) < ti.nkeys) cp_async4(meta_slot + 1, &p.meta[ti.k_row + static_cast<int>(st_t)]);
          if (static_cast<int>(st_t) + 128 < ti.nkeys)
            cp_async4(meta_slot + 2, &p.meta[ti.k_row + static_cast<int>(st_t) + 128]);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // epilogue of tile m (its O is in TMEM once the P V MMAs have completed): 16 of the 64 output dims per thread
    auto epilogue = [&](int m, float inv) {
      const int s = m & 1;
      mbar_wait_quiet(&bar_o_ready[s], static_cast<uint32_t>((m >> 1) & 1));
      tcgen05_fence_after();
      uint32_t o[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]),
            "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
          : "r"(tmem + 256u * static_cast<uint32_t>(s) + lane_off + 192u + 16u * static_cast<uint32_t>(cq))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;"
                   : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]),
                     "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11]), "+r"(o[12]), "+r"(o[13]), "+r"(o[14]), "+r"(o[15])
                   :
                   : "memory");
      tcgen05_fence_before();
      __syncwarp();
      if ((tid & 31u) == 0) mbar_arrive(&bar_s_free[s]);  // O is in registers: the slot may take the next S
      // through shared memory so that a warp's store instruction covers four whole 128-byte lines (a thread's own 32
      // bytes of a row would make it 32 different lines: the output stores were 10 us of the 48 us kernel)
      const uint32_t lane = tid & 31u;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint4 v;
        v.x = cvt_bf16x2(__uint_as_float(o[8 * i + 1]) * inv, __uint_as_float(o[8 * i + 0]) * inv);
        v.y = cvt_bf16x2(__uint_as_float(o[8 * i + 3]) * inv, __uint_as_float(o[8 * i + 2]) * inv);
        v.z = cvt_bf16x2(__uint_as_float(o[8 * i + 5]) * inv, __uint_as_float(o[8 * i + 4]) * inv);
        v.w = cvt_bf16x2(__uint_as_float(o[8 * i + 7]) * inv, __uint_as_float(o[8 * i + 6]) * inv);
        const uint32_t chunk = static_cast<uint32_t>(2 * cq + i);
        *reinterpret_cast<uint4*>(ostage + lane * 128u + ((chunk ^ (lane & 7u)) << 4)) = v;
      }
      named_bar_sync(set_bar, 128);
      const AttnTile ti = tiles[m];
      const int rows = ti.rows_h >> 8;
      if (!(p.debug & 64)) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t q = st_t + 128u * static_cast<uint32_t>(i);
          const uint32_t r = q >> 3, chunk = q & 7u;
          const int row = static_cast<int>(set * 32u + r);
          if (row < rows) {
            const uint4 v = *reinterpret_cast<const uint4*>(ostage + r * 128u + ((chunk ^ (r & 7u)) << 4));
            *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(ti.q_row + row) * p.H + (ti.rows_h & 0xFF) * AT_D +
                                      chunk * 8u) = v;
          }
        }
      }
    };
    fetch_meta(0);
    float inv_prev = 0.f;
    for (int n = 0; n < n_valid; ++n) {
      const int s = n & 1;
      const uint32_t ph = static_cast<uint32_t>((n >> 1) & 1);
      const int nkeys = tiles[n].nkeys;
      const uint32_t t_lane = tmem + 256u * static_cast<uint32_t>(s) + lane_off;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      const uint32_t qmeta = meta_slot[0], kmeta0 = meta_slot[1], kmeta1 = meta_slot[2];
      const int qpos = static_cast<int>(qmeta & 0xFFFFu);
      int lo = 0, hi = nkeys;
      if (p.window >= 0) {
        kpos[st_t] = static_cast<int>(st_t) < nkeys ? static_cast<int>(kmeta0 & 0xFFFFu) : 0x3FFFFFFF;
        kpos[st_t + 128u] = static_cast<int>(st_t) + 128 < nkeys ? static_cast<int>(kmeta1 & 0xFFFFu) : 0x3FFFFFFF;
        named_bar_sync(set_bar, 128);
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
      fetch_meta(n + 1);  // in flight during this tile's softmax
      const int nchunks = (nkeys + 31) >> 5;
      const int ca = cq, cb = cq + 4;
      const bool va = ca < nchunks && !__all_sync(0xffffffffu, hi <= 32 * ca || lo >= 32 * ca + 32);
      const bool vb = cb < nchunks && !__all_sync(0xffffffffu, hi <= 32 * cb || lo >= 32 * cb + 32);
      const bool fa = __all_sync(0xffffffffu, lo <= 32 * ca && hi >= 32 * ca + 32);
      const bool fb = __all_sync(0xffffffffu, lo <= 32 * cb && hi >= 32 * cb + 32);
      mbar_wait_quiet(&bar_s_ready[s], ph);
      tcgen05_fence_after();

      // ---- the thread's scores: two 32-key chunks, both loads in flight; masked entries become -inf
      uint32_t ra[32], rb[32];
      if (p.debug & 8) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ra[j] = rb[j] = __float_as_uint(static_cast<float>(j + n));
      }
      if (va && !(p.debug & 8)) tmem_ld_32x32b_x32(t_lane + 32u * static_cast<uint32_t>(ca), ra);
      if (vb && !(p.debug & 8)) tmem_ld_32x32b_x32(t_lane + 32u * static_cast<uint32_t>(cb), rb);
      float mx = -INFINITY;
      if (va) {
        tmem_ld_wait_for(ra);
        if (!fa) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!(32 * ca + j >= lo && 32 * ca + j < hi)) ra[j] = 0xFF800000u;
        }
        float m0 = __uint_as_float(ra[0]), m1 = __uint_as_float(ra[1]), m2 = __uint_as_float(ra[2]),
              m3 = __uint_as_float(ra[3]);
#pragma unroll
        for (int j = 4; j < 32; j += 4) {
          m0 = fmaxf(m0, __uint_as_float(ra[j]));
          m1 = fmaxf(m1, __uint_as_float(ra[j + 1]));
          m2 = fmaxf(m2, __uint_as_float(ra[j + 2]));
          m3 = fmaxf(m3, __uint_as_float(ra[j + 3]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      if (vb) {
        tmem_ld_wait_for(rb);
        if (!fb) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!(32 * cb + j >= lo && 32 * cb + j < hi)) rb[j] = 0xFF800000u;
        }
        float m0 = __uint_as_float(rb[0]), m1 = __uint_as_float(rb[1]), m2 = __uint_as_float(rb[2]),
              m3 = __uint_as_float(rb[3]);
#pragma unroll
        for (int j = 4; j < 32; j += 4) {
          m0 = fmaxf(m0, __uint_as_float(rb[j]));
          m1 = fmaxf(m1, __uint_as_float(rb[j + 1]));
          m2 = fmaxf(m2, __uint_as_float(rb[j + 2]));
          m3 = fmaxf(m3, __uint_as_float(rb[j + 3]));
        }
        mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      }
      xmax[cq * 128 + static_cast<int>(row_t)] = mx;
      named_bar_sync(set_bar, 128);  // the set's threads hold their scores in registers: P may overwrite their S lanes

      // ---- the previous tile's output, while this tile's P V inputs are still being produced
      if (n >= 1) epilogue(n - 1, inv_prev);

      mx = fmaxf(fmaxf(xmax[row_t], xmax[128 + row_t]), fmaxf(xmax[256 + row_t], xmax[384 + row_t]));
      if (mx == -INFINITY) mx = 0.f;  // rows past the document (never stored)
      const float moff = mx * sl2;

      // ---- p = 2^(s * scale * log2 e - max) (2^-inf = 0 for masked keys), row sum, bf16 P into TMEM columns
      // [16 c, 16 c + 16) of the slot
      float sum0 = 0.f, sum1 = 0.f;
      {
        uint32_t pk[16];
        if (va) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p0 = fmaf(__uint_as_float(ra[2 * j]), sl2, -moff);
            float p1 = fmaf(__uint_as_float(ra[2 * j + 1]), sl2, -moff);
            if (!(p.debug & 1)) p0 = fast_exp2(p0), p1 = fast_exp2(p1);
            sum0 += p0;
            sum1 += p1;
            pk[j] = (p.debug & 2) ? ((__float_as_uint(p0) >> 16) | (__float_as_uint(p1) & 0xFFFF0000u)) : cvt_bf16x2(p1, p0);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
        if (ca < nchunks) tmem_st_32x32b_x16(t_lane + 16u * static_cast<uint32_t>(ca), pk);
        if (vb) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p0 = fmaf(__uint_as_float(rb[2 * j]), sl2, -moff);
            float p1 = fmaf(__uint_as_float(rb[2 * j + 1]), sl2, -moff);
            if (!(p.debug & 1)) p0 = fast_exp2(p0), p1 = fast_exp2(p1);
            sum0 += p0;
            sum1 += p1;
            pk[j] = (p.debug & 2) ? ((__float_as_uint(p0) >> 16) | (__float_as_uint(p1) & 0xFFFF0000u)) : cvt_bf16x2(p1, p0);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
        if (cb < nchunks) tmem_st_32x32b_x16(t_lane + 16u * static_cast<uint32_t>(cb), pk);
      }
      // The second MMA reads ceil(nkeys / 16) 16-key steps.  The V rows behind the last key belong to other
      // documents or to never-written workspace: 0 * NaN would poison O, so they are cleared (generic-proxy
      // stores into the TMA-written stage - after V has landed - made visible to the tensor core by the proxy fence).
      // Only the writing warps fence, and only when there is a tail (nkeys % 16 != 0): the proxy fence is a full
      // memory barrier for the thread - it would wait for the previous tile's output stores and the next tile's
      // metadata loads on every tile.
      if ((nkeys & 15) != 0 && gt < 128u) {
        const int vrow = nkeys + static_cast<int>(gt >> 3);
        if (vrow < ((nkeys + 15) & ~15)) {
          mbar_wait_quiet(&bar_full_v[s], ph);
          *reinterpret_cast<uint4*>(smem + s * TP_STAGE_BYTES + TP_V_OFF + vrow * 128 + (gt & 7u) * 16) =
              make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if ((tid & 31u) == 0) mbar_arrive(&bar_p_ready[s]);  // (512 arrivals on one mbarrier word serialise)
      xsum[cq * 128 + static_cast<int>(row_t)] = sum0 + sum1;
      named_bar_sync(set_bar, 128);
      const float sum = (xsum[row_t] + xsum[128 + row_t]) + (xsum[256 + row_t] + xsum[384 + row_t]);
      inv_prev = sum > 0.f ? 1.0f / sum : 0.f;
    }
    if (n_valid >= 1) epilogue(n_valid - 1, inv_prev);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 17) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 512);
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

// SPLADE_B200_ATTN_TC_V1=1 pins the one-tile-per-CTA tcgen05 kernel (A/B timing)
static bool attn_tc_pipelined() {
  static const bool on = [] {
    const char* e = std::getenv("SPLADE_B200_ATTN_TC_V1");
    return !(e != nullptr && e[0] == '1');
  }();
  return on;
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
  {
    const char* e = std::getenv("SPLADE_B200_ATTN_DEBUG");
    p.debug = e ? std::atoi(e) : 0;
  }
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
    const long long tiles_all = static_cast<long long>((s8 + TA_BQ - 1) / TA_BQ) * n_heads * B;
    if (attn_tc_pipelined() && tiles_all <= static_cast<long long>(TP_MAX_TILES) * sm_count()) {
      // persistent, software-pipelined kernel: one CTA per SM over the (q tile, head, document) list
      const int nqt = (s8 + TA_BQ - 1) / TA_BQ;
      const int n_tiles = nqt * n_heads * B;
      SPLADE_CUDA_OK(cudaFuncSetAttribute(trunk_attn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          TP_SMEM_BYTES));
      const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
      trunk_attn_tc2_kernel<<<grid, TP_THREADS, TP_SMEM_BYTES, stream>>>(tmap_q, tmap_kv, p, n_tiles, nqt, n_heads);
    } else {
      SPLADE_CUDA_OK(cudaFuncSetAttribute(trunk_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          TA_SMEM_BYTES));
      trunk_attn_tc_kernel<<<dim3((s8 + TA_BQ - 1) / TA_BQ, n_heads, B), 256, TA_SMEM_BYTES, stream>>>(tmap_q, tmap_kv,
                                                                                                     p);
    }
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

// ------------------------------------------------------------------------------------------------ training forward
// The same layer loop as trunk_encode_launch with every tensor the backward needs written straight into the caller's
// save buffers instead of scratch (the residual stream is copied at the two points a LayerNorm reads it), stopping at
// final_norm: the prediction head and the SPLADE head run as autograd nodes of their own (ops.py).  The save buffers
// must arrive zero-filled: rows behind the last token then stay zero everywhere (the kernels write real rows, or
// products of zero rows), which is what lets the backward run over whole buffers without knowing the row count.
size_t trunk_train_workspace_bytes(int B, int S, int H, int V) {
  const size_t rows = (static_cast<size_t>(B) * ((static_cast<size_t>(S) + 7) / 8 * 8) + 255) / 256 * 256;
  return al256(varlen_layout_bytes(B, S, V)) + al256(rows * H * 4) + 256;
}

int trunk_train_fwd_launch(const splade_trunk_args* a, const splade_trunk_train_saves* sv, cudaStream_t stream) {
  if (!a || !sv || !a->input_ids || !a->mask || !a->tok_emb || !a->layers || !a->emb_norm || !a->final_norm)
    return set_error(kErrInvalidArg, "splade_trunk_train_fwd: null pointer argument");
  if (!sv->x_attn_in || !sv->x_mlp_in || !sv->xb || !sv->qkv || !sv->ab || !sv->xb2 || !sv->gg || !sv->x_final ||
      !sv->out || !sv->meta || !sv->plan)
    return set_error(kErrInvalidArg, "splade_trunk_train_fwd: null save buffer");
  const int B = a->B, S = a->S, H = a->H, I = a->I;
  if (B < 1 || S < 1 || H % 64 != 0 || H < 256 || I % 128 != 0 || a->n_heads * 64 != H || a->n_layers < 1)
    return set_error(kErrUnsupportedShape,
                     "splade_trunk_train_fwd: needs hidden %% 64 == 0 and >= 256, 64-wide heads, intermediate %% 128 == 0 "
                     "(B=%d S=%d H=%d I=%d heads=%d)", B, S, H, I, a->n_heads);
  int rc = require_sm100("splade_trunk_train_fwd");
  if (rc != kOk) return rc;
  const size_t need = trunk_train_workspace_bytes(B, S, H, 1);
  if (!a->workspace || a->ws_bytes < need)
    return set_error(kErrWorkspace, "splade_trunk_train_fwd: workspace too small (%zu < %zu bytes)", a->ws_bytes, need);

  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  VarlenLayout lay{};
  rc = varlen_layout_build(a->mask, B, S, 1, ws, varlen_layout_bytes(B, S, 1), &lay, stream);
  if (rc != kOk) return rc;
  ws += al256(varlen_layout_bytes(B, S, 1));
  const size_t rows = lay.rows;
  float* x = reinterpret_cast<float*>(ws);
  const int32_t* t_rows = &lay.plan->T;
  const int rows_i = static_cast<int>(rows);
  const size_t rh = rows * static_cast<size_t>(H);
  SPLADE_CUDA_OK(cudaMemsetAsync(x, 0, rh * sizeof(float), stream));

  auto bf = [](void* base, size_t elems) { return static_cast<void*>(static_cast<__nv_bfloat16*>(base) + elems); };
  // embeddings -> x (fp32) and, for layer 0 whose attn_norm is Identity, its bf16 copy
  if ((rc = trunk_embed_ln_launch(a->input_ids, a->tok_emb, a->emb_norm, a->norm_eps, lay, S, H, x, sv->xb, stream)) != kOk)
    return rc;
  for (int l = 0; l < a->n_layers; ++l) {
    const splade_trunk_layer& ly = a->layers[l];
    void* xb = bf(sv->xb, l * rh);
    void* qkv = bf(sv->qkv, l * rh * 3);
    void* ab = bf(sv->ab, l * rh);
    void* xb2 = bf(sv->xb2, l * rh);
    void* gg = bf(sv->gg, l * rows * static_cast<size_t>(I));
    SPLADE_CUDA_OK(cudaMemcpyAsync(sv->x_attn_in + l * rh, x, rh * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if (ly.attn_norm != nullptr) {
      if ((rc = trunk_ln_launch(x, 0, ly.attn_norm, a->norm_eps, t_rows, rows_i, H, xb, stream)) != kOk) return rc;
    } else if (l != 0) {
      return set_error(kErrUnsupportedShape, "splade_trunk_train_fwd: only layer 0 may lack attn_norm");
    }
    if ((rc = trunk_gemm_launch(EPI_QKV_ROPE, xb, rows_i, ly.wqkv, 3 * H, H, t_rows, qkv, 3 * H, lay.meta,
                                ly.is_global ? a->global_theta : a->local_theta, 2 * H, stream)) != kOk)
      return rc;
    if ((rc = trunk_attn_launch(qkv, ab, lay, B, S, H, a->n_heads, ly.is_global ? -1 : a->local_window, stream)) != kOk)
      return rc;
    if ((rc = trunk_gemm_launch(EPI_RESID, ab, rows_i, ly.wo, H, H, t_rows, x, H, nullptr, 0.f, 0, stream)) != kOk)
      return rc;
    SPLADE_CUDA_OK(cudaMemcpyAsync(sv->x_mlp_in + l * rh, x, rh * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if ((rc = trunk_ln_launch(x, 0, ly.mlp_norm, a->norm_eps, t_rows, rows_i, H, xb2, stream)) != kOk) return rc;
    if ((rc = trunk_gemm_launch(EPI_GEGLU, xb2, rows_i, ly.wi, I, H, t_rows, gg, I, nullptr, 0.f, 0, stream)) != kOk)
      return rc;
    if ((rc = trunk_gemm_launch(EPI_RESID, gg, rows_i, ly.wo2, H, I, t_rows, x, H, nullptr, 0.f, 0, stream)) != kOk)
      return rc;
  }
  SPLADE_CUDA_OK(cudaMemcpyAsync(sv->x_final, x, rh * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if ((rc = trunk_ln_launch(x, 0, a->final_norm, a->norm_eps, t_rows, rows_i, H, sv->out, stream)) != kOk) return rc;
  SPLADE_CUDA_OK(cudaMemcpyAsync(sv->meta, lay.meta, rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  SPLADE_CUDA_OK(cudaMemcpyAsync(sv->plan, lay.plan, 5 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  return kOk;
}

}  // namespace splade
