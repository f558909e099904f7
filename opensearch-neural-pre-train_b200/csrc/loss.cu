// SPLADELossV33 on the pooled sparse vectors (sm_100a): FLOPS regulariser + nonzero monitoring, q.d score
// matrix, per-row hard-negative dots, the scalar stage (InfoNCE / KL-KD / MarginMSE / total) and the
// backward onto the sparse representations.  Reference: src/model/losses.py:57-73 (FLOPS), :136-181 (InfoNCE),
// :92-134 (MarginMSE), :239-253 (KL-KD), :271-295 (monitoring, loss_dict).
//
// All of these are HBM/L2-bandwidth or launch bound ([N,V] fp32 streams, V = 50 000): they are written as
// coalesced, vectorised CUDA-core kernels with warp-shuffle reductions; every cross-CTA reduction goes
// through a fixed-order second stage so results are run-to-run deterministic.
#include "../../include/splade_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace splade {
namespace {

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

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

// ---------------------------------------------------------------------------------------------
// FLOPS + nnz:  colmean[j] = mean_i rep[i,j];  flops = sum_j colmean_j^2;  nnz = (#rep>0) / N
// ---------------------------------------------------------------------------------------------
constexpr int kFlopsThreads = 256;

__global__ void __launch_bounds__(kFlopsThreads)
flops_cols_kernel(const float* __restrict__ rep, int N, int V, float* __restrict__ colmean,
                  float2* __restrict__ partials) {
  const int j = blockIdx.x * kFlopsThreads + threadIdx.x;
  float sum = 0.f, cnt = 0.f;
  if (j < V) {
    const float* p = rep + j;
    int i = 0;
    for (; i + 16 <= N; i += 16) {  // 16 independent loads in flight per thread (latency-bound otherwise)
      float x[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) x[u] = p[static_cast<size_t>(i + u) * V];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        sum += x[u];
        cnt += (x[u] > 0.f);
      }
    }
    for (; i < N; ++i) {
      const float a = p[static_cast<size_t>(i) * V];
      sum += a;
      cnt += (a > 0.f);
    }
  }
  const float mean = (N > 0) ? sum / static_cast<float>(N) : 0.f;
  if (j < V) colmean[j] = mean;
  float sq = warp_sum(mean * mean);
  float ct = warp_sum(cnt);
  __shared__ float s_sq[kFlopsThreads / 32], s_ct[kFlopsThreads / 32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    s_sq[w] = sq;
    s_ct[w] = ct;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < kFlopsThreads / 32; ++i) {
      a += s_sq[i];
      b += s_ct[i];
    }
    partials[blockIdx.x] = make_float2(a, b);
  }
}

__global__ void __launch_bounds__(256)
flops_final_kernel(const float2* __restrict__ partials, int nblocks, int N, float* __restrict__ out2) {
  __shared__ float s_a[256], s_b[256];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    a += partials[i].x;
    b += partials[i].y;
  }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_b[threadIdx.x] += s_b[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out2[0] = s_a[0];
    out2[1] = (N > 0) ? s_b[0] / static_cast<float>(N) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// raw[i,j] = sum_v q[i,v] docs[j,v]   (split over v, fixed-order reduction)
// ---------------------------------------------------------------------------------------------
constexpr int TI = 64, TJ = 64, TK = 32, TP = 65;

__global__ void __launch_bounds__(256)
scores_partial_kernel(const float* __restrict__ q, int Bq, const float* __restrict__ docs, int N, int V,
                      int round_bf16, int chunk, float* __restrict__ out /* [ksplit][Bq][N] */) {
  __shared__ float As[TK][TP];
  __shared__ float Bs[TK][TP];
  const int j0 = blockIdx.x * TJ, i0 = blockIdx.y * TI;
  const int v_beg = blockIdx.z * chunk;
  const int v_end = min(V, v_beg + chunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const int lrow = threadIdx.x >> 3;       // 0..31
  const int lk = (threadIdx.x & 7) * 4;    // 0,4,..,28
  for (int v0 = v_beg; v0 < v_end; v0 += TK) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = lrow + 32 * half;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int v = v0 + lk + e;
        float a = 0.f, b = 0.f;
        if (v < v_end) {
          if (i0 + r < Bq) a = q[static_cast<size_t>(i0 + r) * V + v];
          if (j0 + r < N) b = docs[static_cast<size_t>(j0 + r) * V + v];
          if (round_bf16) {
            a = bf16_round(a);
            b = bf16_round(b);
          }
        }
        As[lk + e][r] = a;
        Bs[lk + e][r] = b;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* o = out + static_cast<size_t>(blockIdx.z) * Bq * N;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i >= Bq) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + tx * 4 + b;
      if (j < N) o[static_cast<size_t>(i) * N + j] = acc[a][b];
    }
  }
}

// One warp per output element: lanes stride over the k-splits, then a fixed-order shuffle tree (deterministic).
__global__ void __launch_bounds__(256)
scores_reduce_kernel(const float* __restrict__ partial, int ksplit, size_t n, float* __restrict__ raw) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < ksplit; k += 32) s += partial[static_cast<size_t>(k) * n + i];
  s = warp_sum(s);
  if (lane == 0) raw[i] = s;
}

void scores_plan(int Bq, int N, int V, int* ksplit, int* chunk) {
  const int tiles = ((Bq + TI - 1) / TI) * ((N + TJ - 1) / TJ);
  int ks = std::max(1, (2 * sm_count()) / std::max(1, tiles));
  const int kblocks = (V + TK - 1) / TK;
  ks = std::min(ks, kblocks);
  const int per = (kblocks + ks - 1) / ks;
  *chunk = per * TK;
  *ksplit = (kblocks + per - 1) / per;
}

// ---------------------------------------------------------------------------------------------
// out[i,j] = sum_v q[i,v] * docs[i,j,v]    one CTA per (i,j)
// ---------------------------------------------------------------------------------------------
constexpr int kRowdotThreads = 1024;

__global__ void __launch_bounds__(kRowdotThreads)
rowdot_kernel(const float* __restrict__ q, const float* __restrict__ docs, int k, int V,
              float* __restrict__ out) {
  const int i = blockIdx.x / k;
  const float* a = q + static_cast<size_t>(i) * V;
  const float* b = docs + static_cast<size_t>(blockIdx.x) * V;
  float s = 0.f;
  if ((V & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const int n4 = V >> 2;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int v0 = threadIdx.x; v0 < n4; v0 += 4 * kRowdotThreads) {  // four independent 16-byte load pairs in flight
      float4 x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * kRowdotThreads;
        x[u] = (v < n4) ? __ldg(&a4[v]) : make_float4(0.f, 0.f, 0.f, 0.f);
        y[u] = (v < n4) ? __ldg(&b4[v]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      s0 = fmaf(x[0].x, y[0].x, fmaf(x[0].y, y[0].y, fmaf(x[0].z, y[0].z, fmaf(x[0].w, y[0].w, s0))));
      s1 = fmaf(x[1].x, y[1].x, fmaf(x[1].y, y[1].y, fmaf(x[1].z, y[1].z, fmaf(x[1].w, y[1].w, s1))));
      s2 = fmaf(x[2].x, y[2].x, fmaf(x[2].y, y[2].y, fmaf(x[2].z, y[2].z, fmaf(x[2].w, y[2].w, s2))));
      s3 = fmaf(x[3].x, y[3].x, fmaf(x[3].y, y[3].y, fmaf(x[3].z, y[3].z, fmaf(x[3].w, y[3].w, s3))));
    }
    s = (s0 + s1) + (s2 + s3);
  } else {
    for (int v = threadIdx.x; v < V; v += kRowdotThreads) s = fmaf(a[v], b[v], s);
  }
  s = warp_sum(s);
  __shared__ float sm[kRowdotThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kRowdotThreads / 32; ++w) t += sm[w];
    out[blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// Scalar stage: one warp per query row (8 rows per CTA) writes that row's loss terms and gradient rows;
// v33_finalize_kernel then adds the rows up in index order (deterministic) and forms the total.
// ---------------------------------------------------------------------------------------------
// One query row of the scalar stage, by one warp: that row's loss terms (returned on every lane) and its gradient
// rows g_inbatch[i,:] / g_hard[i,:].
__device__ __forceinline__ void v33_scalars_row(const splade_v33_scalar_args& a, int i, int lane, float& acc_nce,
                                                float& acc_kd, float& acc_mm) {
  const int Bq = a.Bq, N = a.N, k = a.k;
  const float invB = 1.f / static_cast<float>(Bq);
  const float* rin = a.raw_inbatch + static_cast<size_t>(i) * N;
  const float* rh = a.raw_hard + static_cast<size_t>(i) * k;
  float* gin = a.g_inbatch + static_cast<size_t>(i) * N;
  float* gh = a.g_hard + static_cast<size_t>(i) * k;
  const int label = a.label_offset + i;

  // ---- InfoNCE: CE over [in-batch | hard] scores (losses.py:155-181)
  float m = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    float s = rin[j];
    s = a.round_bf16 ? bf16_round(bf16_round(s) / a.temperature) : s / a.temperature;
    m = fmaxf(m, s);
  }
  for (int j = lane; j < k; j += 32) m = fmaxf(m, rh[j] / a.temperature);
  m = warp_max(m);
  float se = 0.f, s_lab = 0.f;
  for (int j = lane; j < N; j += 32) {
    float s = rin[j];
    s = a.round_bf16 ? bf16_round(bf16_round(s) / a.temperature) : s / a.temperature;
    se += expf(s - m);
    if (j == label) s_lab = s;
  }
  for (int j = lane; j < k; j += 32) se += expf(rh[j] / a.temperature - m);
  se = warp_sum(se);
  s_lab = warp_sum(s_lab);
  const float lse = m + logf(se);
  acc_nce += lse - s_lab;
  const float gscale = invB / a.temperature;
  for (int j = lane; j < N; j += 32) {
    float s = rin[j];
    s = a.round_bf16 ? bf16_round(bf16_round(s) / a.temperature) : s / a.temperature;
    gin[j] = (expf(s - lse) - (j == label ? 1.f : 0.f)) * gscale;
  }
  for (int j = lane; j < k; j += 32) gh[j] = expf(rh[j] / a.temperature - lse) * gscale;
  __syncwarp();

  // ---- optional KL-KD on the in-batch matrix (losses.py:239-253); requires N == Bq
  if (a.has_kd) {
    const float* tr = a.teacher_kd + static_cast<size_t>(i) * N;
    float ms = -INFINITY, mt = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = rin[j];
      s = a.round_bf16 ? bf16_round(bf16_round(s) / a.kd_temperature) : s / a.kd_temperature;
      ms = fmaxf(ms, s);
      mt = fmaxf(mt, tr[j] / a.kd_temperature);
    }
    ms = warp_max(ms);
    mt = warp_max(mt);
    float ss = 0.f, st = 0.f;
    for (int j = lane; j < N; j += 32) {
      float s = rin[j];
      s = a.round_bf16 ? bf16_round(bf16_round(s) / a.kd_temperature) : s / a.kd_temperature;
      ss += expf(s - ms);
      st += expf(tr[j] / a.kd_temperature - mt);
    }
    ss = warp_sum(ss);
    st = warp_sum(st);
    const float lss = ms + logf(ss), lst = mt + logf(st);
    float kl = 0.f;
    for (int j = lane; j < N; j += 32) {
      float s = rin[j];
      s = a.round_bf16 ? bf16_round(bf16_round(s) / a.kd_temperature) : s / a.kd_temperature;
      const float logp = s - lss;
      const float logt = tr[j] / a.kd_temperature - lst;
      const float t = expf(logt);
      if (t > 0.f) kl += t * (logt - logp);
      gin[j] += a.lambda_kd * (expf(logp) - t) * invB / a.kd_temperature;
    }
    acc_kd += warp_sum(kl);
    __syncwarp();
  }

  // ---- optional MarginMSE (losses.py:92-134): raw (un-tempered) fp32 dots
  if (a.has_mmse) {
    const float sp = a.raw_pos[i];
    const float tp = a.teacher_pos[i];
    const float c = 2.f * invB / static_cast<float>(k);
    float sq = 0.f, gpos = 0.f;
    for (int j = lane; j < k; j += 32) {
      const float diff = (sp - rh[j]) - (tp - a.teacher_neg[static_cast<size_t>(i) * k + j]);
      sq += diff * diff;
      gpos += c * diff;
      gh[j] += a.lambda_mmse * (-c * diff);
    }
    acc_mm += warp_sum(sq);
    gpos = warp_sum(gpos);
    if (lane == 0) gin[label] += a.lambda_mmse * gpos;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) v33_scalars_kernel(const splade_v33_scalar_args a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Bq = a.Bq;
  float acc_nce = 0.f, acc_kd = 0.f, acc_mm = 0.f;
  const int i = blockIdx.x * 8 + warp;
  if (i < Bq) v33_scalars_row(a, i, lane, acc_nce, acc_kd, acc_mm);
  if (i < Bq && lane == 0) {
    a.row_scratch[i] = acc_nce;
    a.row_scratch[Bq + i] = acc_kd;
    a.row_scratch[2 * Bq + i] = acc_mm;
  }
}

// Adds the per-row terms up in index order (fixed assignment + fixed tree => deterministic) and forms the total.
// Called by the first 256 threads of a CTA; `flops` = {flops_q, nnz_q, flops_d, nnz_d, flops_neg, nnz_neg}.
__device__ __forceinline__ void v33_finalize_block(const splade_v33_scalar_args& a, const float* flops,
                                                   float (*s_part)[256]) {
  const int Bq = a.Bq, k = a.k;
  float nce = 0.f, kd = 0.f, mm = 0.f;
  for (int i = threadIdx.x; i < Bq; i += 256) {
    nce += a.row_scratch[i];
    kd += a.row_scratch[Bq + i];
    mm += a.row_scratch[2 * Bq + i];
  }
  s_part[0][threadIdx.x] = nce;
  s_part[1][threadIdx.x] = kd;
  s_part[2][threadIdx.x] = mm;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int c = 0; c < 3; ++c) s_part[c][threadIdx.x] += s_part[c][threadIdx.x + o];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    const float invB = 1.f / static_cast<float>(Bq);
    nce = s_part[0][0] * invB;
    kd = s_part[1][0] * invB;                            // kl_div(reduction="batchmean")
    mm = s_part[2][0] * invB / static_cast<float>(k);    // mse_loss mean over [Bq,k]
    const float fq = flops[0], fd = flops[2], fn = flops[4];
    float loss = nce + a.lambda_q * fq + a.lambda_d * fd + a.lambda_neg * fn;
    if (a.has_kd) loss += a.lambda_kd * kd;
    if (a.has_mmse) loss += a.lambda_mmse * mm;
    a.scalars[0] = loss;
    a.scalars[1] = nce;
    a.scalars[2] = fq;
    a.scalars[3] = fd;
    a.scalars[4] = fn;
    a.scalars[5] = a.has_kd ? kd : 0.f;
    a.scalars[6] = a.has_mmse ? mm : 0.f;
    a.scalars[7] = flops[1];
    a.scalars[8] = flops[3];
  }
}

__global__ void __launch_bounds__(256) v33_finalize_kernel(const splade_v33_scalar_args a) {
  __shared__ float s_part[3][256];
  v33_finalize_block(a, a.flops3, s_part);
}

// ---------------------------------------------------------------------------------------------
// Fused forward (splade_v33_loss_fwd): three launches + the score GEMM instead of fourteen.
//
// (F1) v33_prep_kernel - ONE pass over q, the local positives and the negatives, one thread per vocabulary column
//      (coalesced across the CTA, 8 query rows = 8 (2 + k) loads in flight per thread):
//        * FLOPS column means of the three tensors + per-CTA partials of sum mean^2 and of the >0 counts
//          (same summation order as splade_flops_nnz => bit-identical),
//        * the bf16 operands of the score GEMM (the cast autocast performs),
//        * per-CTA partial dot products q_i . neg_{i,kk} and (MarginMSE) q_i . pos_i.
// (F2) the score GEMM (tcgen05 split-K under autocast, CUDA cores in fp32 mode) -> partial tiles
// (F3) v33_reduce2_kernel - fixed-order reductions of the split-K score tiles and of the dot-product partials
// (F4) v33_scalars_fused_kernel - ONE CTA: a warp per query row (CE / KL-KD / MarginMSE + gradient rows), then
//      the fixed-order totals, the FLOPS partials and the scalars in the same kernel (Bq <= 256).
// ---------------------------------------------------------------------------------------------
constexpr int kPrepRows = 8;

struct PrepParams {
  const float* q;       // [Bq, V]
  const float* docs;    // [N, V]   (local positives = rows label_offset .. label_offset + Bq)
  const float* negs;    // [Bq * k, V]
  int Bq, N, k, V, Vp, label_offset;
  int pack_q, pack_docs, want_pos;
  float* colmean;       // [3][V]
  float2* fpart;        // [3][nblk]
  __nv_bfloat16* q16;   // [Bq, Vp]
  __nv_bfloat16* d16;   // [N, Vp]
  float* rd_part;       // [nblk][Bq * (k + 1)]   (slot k of a row = q . pos)
};

template <int KMAX>  // k <= KMAX in {1, 2, 4, 8}
__global__ void __launch_bounds__(256) v33_prep_kernel(const PrepParams p) {
  __shared__ float s_rd[8][kPrepRows * (KMAX + 1)];  // [warp][row-in-chunk * (k+1) + kk]
  __shared__ float s_red[3][2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x * 256 + tid;
  const bool live = j < p.V;
  const int Bq = p.Bq, k = p.k, V = p.V;
  const int kk1 = k + 1;
  const float* qc = p.q + j;
  const float* dc = p.docs + static_cast<size_t>(p.label_offset) * V + j;
  const float* nc = p.negs + j;
  float sum_q = 0.f, cnt_q = 0.f, sum_d = 0.f, cnt_d = 0.f, sum_n = 0.f, cnt_n = 0.f;

  for (int i0 = 0; i0 < Bq; i0 += kPrepRows) {
    float xq[kPrepRows], xd[kPrepRows], xn[kPrepRows][KMAX];
#pragma unroll
    for (int u = 0; u < kPrepRows; ++u) {  // 8 (2 + k) independent loads in flight
      const int i = i0 + u;
      const bool ok = live && i < Bq;
      xq[u] = ok ? qc[static_cast<size_t>(i) * V] : 0.f;
      xd[u] = ok ? dc[static_cast<size_t>(i) * V] : 0.f;
#pragma unroll
      for (int kk = 0; kk < KMAX; ++kk)
        xn[u][kk] = (ok && kk < k) ? nc[(static_cast<size_t>(i) * k + kk) * V] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kPrepRows; ++u) {
      const int i = i0 + u;
      if (i < Bq) {
        sum_q += xq[u];
        cnt_q += (xq[u] > 0.f);
        sum_d += xd[u];
        cnt_d += (xd[u] > 0.f);
        if (j < p.Vp) {
          if (p.pack_q) p.q16[static_cast<size_t>(i) * p.Vp + j] = __float2bfloat16_rn(xq[u]);
          if (p.pack_docs)
            p.d16[static_cast<size_t>(p.label_offset + i) * p.Vp + j] = __float2bfloat16_rn(xd[u]);
        }
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) {  // in row order: the FLOPS sums run over rows r = i*k + kk ascending
          if (kk < k) {
            sum_n += xn[u][kk];
            cnt_n += (xn[u][kk] > 0.f);
            const float dot = warp_sum(xq[u] * xn[u][kk]);
            if (lane == 0) s_rd[warp][u * (KMAX + 1) + kk] = dot;
          }
        }
        if (p.want_pos) {
          const float dot = warp_sum(xq[u] * xd[u]);
          if (lane == 0) s_rd[warp][u * (KMAX + 1) + KMAX] = dot;
        }
      }
    }
    __syncthreads();  // the eight warps' partials of this chunk of rows, added in warp order
    if (tid < kPrepRows * (KMAX + 1)) {
      const int u = tid / (KMAX + 1), kk = tid - u * (KMAX + 1);
      const int i = i0 + u;
      if (i < Bq && (kk < k || (kk == KMAX && p.want_pos))) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_rd[w][tid];
        p.rd_part[static_cast<size_t>(blockIdx.x) * Bq * kk1 + static_cast<size_t>(i) * kk1 + (kk == KMAX ? k : kk)] = t;
      }
    }
    __syncthreads();
  }
  // documents that are not this rank's positives (gathered negatives, NCCL statement): only the bf16 operand
  if (p.pack_docs && j < p.Vp) {
    for (int r0 = 0; r0 < p.N; r0 += 8) {  // eight independent loads in flight
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + u;
        const bool other = r < p.N && !(r >= p.label_offset && r < p.label_offset + Bq);
        x[u] = (other && live) ? p.docs[static_cast<size_t>(r) * V + j] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + u;
        if (r < p.N && !(r >= p.label_offset && r < p.label_offset + Bq))
          p.d16[static_cast<size_t>(r) * p.Vp + j] = __float2bfloat16_rn(x[u]);
      }
    }
  }
  const float mq = sum_q / static_cast<float>(Bq), md = sum_d / static_cast<float>(Bq);
  const float mn = sum_n / static_cast<float>(Bq * k);
  if (live) {
    p.colmean[j] = mq;
    p.colmean[static_cast<size_t>(V) + j] = md;
    p.colmean[2 * static_cast<size_t>(V) + j] = mn;
  }
  // per-CTA partials, exactly as flops_cols_kernel forms them
  const float sq[3] = {warp_sum(mq * mq), warp_sum(md * md), warp_sum(mn * mn)};
  const float ct[3] = {warp_sum(cnt_q), warp_sum(cnt_d), warp_sum(cnt_n)};
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      s_red[t][0][warp] = sq[t];
      s_red[t][1][warp] = ct[t];
    }
  }
  __syncthreads();
  if (tid < 3) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += s_red[tid][0][w];
      b += s_red[tid][1][w];
    }
    p.fpart[static_cast<size_t>(tid) * gridDim.x + blockIdx.x] = make_float2(a, b);
  }
}

// One warp per output element: lanes stride over the partials, then a fixed-order shuffle tree (deterministic).
// Elements [0, nA) reduce partialA (the split-K score tiles), elements [nA, nA + nB) reduce partialB (the dots).
__global__ void __launch_bounds__(256)
v33_reduce2_kernel(const float* __restrict__ partialA, int ksA, size_t nA, float* __restrict__ outA,
                   const float* __restrict__ partialB, int ksB, size_t nB, float* __restrict__ outB, int k,
                   float* __restrict__ out_pos) {
  size_t i = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i < nA) {
    float s = 0.f;
    for (int t = lane; t < ksA; t += 32) s += partialA[static_cast<size_t>(t) * nA + i];
    s = warp_sum(s);
    if (lane == 0) outA[i] = s;
    return;
  }
  i -= nA;
  if (i >= nB) return;
  float s = 0.f;
  for (int t = lane; t < ksB; t += 32) s += partialB[static_cast<size_t>(t) * nB + i];
  s = warp_sum(s);
  if (lane == 0) {  // element i = row * (k+1) + kk; slot k is q . pos
    const int row = static_cast<int>(i / (k + 1)), kk = static_cast<int>(i % (k + 1));
    if (kk < k) {
      outB[static_cast<size_t>(row) * k + kk] = s;
    } else if (out_pos) {
      out_pos[row] = s;
    }
  }
}

__global__ void __launch_bounds__(1024)
v33_scalars_fused_kernel(const splade_v33_scalar_args a, const float2* __restrict__ fpart, int nblk, int rows_neg) {
  __shared__ float s_part[3][256];
  __shared__ float s_fa[3][256], s_fb[3][256];
  __shared__ float s_flops[6];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Bq = a.Bq;
  for (int i = warp; i < Bq; i += 32) {
    float acc_nce = 0.f, acc_kd = 0.f, acc_mm = 0.f;
    v33_scalars_row(a, i, lane, acc_nce, acc_kd, acc_mm);
    if (lane == 0) {
      a.row_scratch[i] = acc_nce;
      a.row_scratch[Bq + i] = acc_kd;
      a.row_scratch[2 * Bq + i] = acc_mm;
    }
  }
  // FLOPS partials of the three tensors: threads 0-255 / 256-511 / 512-767, the tree of flops_final_kernel
  const int t3 = threadIdx.x >> 8, tl = threadIdx.x & 255;
  if (t3 < 3) {
    float x = 0.f, y = 0.f;
    for (int i = tl; i < nblk; i += 256) {
      const float2 v = fpart[static_cast<size_t>(t3) * nblk + i];
      x += v.x;
      y += v.y;
    }
    s_fa[t3][tl] = x;
    s_fb[t3][tl] = y;
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t3 < 3 && tl < o) {
      s_fa[t3][tl] += s_fa[t3][tl + o];
      s_fb[t3][tl] += s_fb[t3][tl + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) {
    const float n_rows = static_cast<float>(threadIdx.x == 2 ? rows_neg : Bq);
    s_flops[2 * threadIdx.x] = s_fa[threadIdx.x][0];
    s_flops[2 * threadIdx.x + 1] = s_fb[threadIdx.x][0] / n_rows;
  }
  __syncthreads();  // row_scratch (global, written by this CTA) and s_flops are visible to the finalising threads
  if (threadIdx.x < 256) v33_finalize_block(a, s_flops, s_part);
}

// ---------------------------------------------------------------------------------------------
// Backward onto q / docs / negs.  One thread per vocabulary column, 128 columns per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int BT = 32;  // tile edge over queries / documents

__global__ void __launch_bounds__(128)
loss_bwd_kernel(const float* __restrict__ q, const float* __restrict__ docs, const float* __restrict__ negs,
                int Bq, int N, int k, int V, int label_offset, const float* __restrict__ g_in,
                const float* __restrict__ g_hard, const float* __restrict__ cm_q, const float* __restrict__ cm_d,
                const float* __restrict__ cm_n, float cq, float cd, float cn, const float* __restrict__ g_loss,
                float* __restrict__ dq, float* __restrict__ ddocs, float* __restrict__ dnegs) {
  __shared__ __align__(16) float Gs[BT][BT];
  const int col = blockIdx.x * 128 + threadIdx.x;
  const bool live = col < V;
  const float gl = *g_loss;
  const float fq = live ? cq * cm_q[col] : 0.f;
  const float fd = live ? cd * cm_d[col] : 0.f;
  const float fn = live ? cn * cm_n[col] : 0.f;

  // ---- dq (and dnegs)
  for (int i0 = 0; i0 < Bq; i0 += BT) {
    float acc[BT];
#pragma unroll
    for (int ii = 0; ii < BT; ++ii) acc[ii] = 0.f;
    for (int j0 = 0; j0 < N; j0 += BT) {
      __syncthreads();
      for (int e = threadIdx.x; e < BT * BT; e += 128) {
        const int ii = e / BT, jj = e % BT;
        Gs[ii][jj] = (i0 + ii < Bq && j0 + jj < N) ? g_in[static_cast<size_t>(i0 + ii) * N + j0 + jj] : 0.f;
      }
      __syncthreads();
      float d[BT];
#pragma unroll
      for (int jj = 0; jj < BT; ++jj)
        d[jj] = (live && j0 + jj < N) ? docs[static_cast<size_t>(j0 + jj) * V + col] : 0.f;
#pragma unroll
      for (int ii = 0; ii < BT; ++ii) {
#pragma unroll
        for (int jj = 0; jj < BT; jj += 4) {
          const float4 g = *reinterpret_cast<const float4*>(&Gs[ii][jj]);
          acc[ii] = fmaf(g.x, d[jj], acc[ii]);
          acc[ii] = fmaf(g.y, d[jj + 1], acc[ii]);
          acc[ii] = fmaf(g.z, d[jj + 2], acc[ii]);
          acc[ii] = fmaf(g.w, d[jj + 3], acc[ii]);
        }
      }
    }
    if (live) {
#pragma unroll
      for (int ii = 0; ii < BT; ++ii) {
        const int i = i0 + ii;
        if (i < Bq) {
          const float qv = q[static_cast<size_t>(i) * V + col];
          float extra = 0.f;
          for (int kk = 0; kk < k; ++kk) {
            const float gh = __ldg(&g_hard[static_cast<size_t>(i) * k + kk]);
            const size_t o = (static_cast<size_t>(i) * k + kk) * V + col;
            extra = fmaf(gh, negs[o], extra);
            dnegs[o] = gl * (gh * qv + fn);
          }
          dq[static_cast<size_t>(i) * V + col] = gl * (acc[ii] + extra + fq);
        }
      }
    }
  }

  // ---- ddocs
  for (int j0 = 0; j0 < N; j0 += BT) {
    float acc[BT];
#pragma unroll
    for (int jj = 0; jj < BT; ++jj) acc[jj] = 0.f;
    for (int i0 = 0; i0 < Bq; i0 += BT) {
      __syncthreads();
      for (int e = threadIdx.x; e < BT * BT; e += 128) {
        const int ii = e / BT, jj = e % BT;
        Gs[ii][jj] = (i0 + ii < Bq && j0 + jj < N) ? g_in[static_cast<size_t>(i0 + ii) * N + j0 + jj] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int ii = 0; ii < BT; ++ii) {
        const float qv = (live && i0 + ii < Bq) ? q[static_cast<size_t>(i0 + ii) * V + col] : 0.f;
#pragma unroll
        for (int jj = 0; jj < BT; jj += 4) {
          const float4 g = *reinterpret_cast<const float4*>(&Gs[ii][jj]);
          acc[jj] = fmaf(g.x, qv, acc[jj]);
          acc[jj + 1] = fmaf(g.y, qv, acc[jj + 1]);
          acc[jj + 2] = fmaf(g.z, qv, acc[jj + 2]);
          acc[jj + 3] = fmaf(g.w, qv, acc[jj + 3]);
        }
      }
    }
    if (live) {
#pragma unroll
      for (int jj = 0; jj < BT; ++jj) {
        const int j = j0 + jj;
        if (j < N) {
          const bool local = (j >= label_offset) && (j < label_offset + Bq);
          ddocs[static_cast<size_t>(j) * V + col] = gl * (acc[jj] + (local ? fd : 0.f));
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Vectorised backward (V % 4 == 0): each thread owns FOUR vocabulary columns (16-byte loads/stores), a CTA of
// 128 threads 512 columns and a tile of 16 queries (dq) or 16 documents (ddocs); the 16 x 32 block of G is
// broadcast from shared memory with one LDS.128 per 16 FMAs.
// ---------------------------------------------------------------------------------------------
constexpr int VT = 16;   // rows (queries or documents) per CTA
constexpr int VK = 32;   // reduction tile

__device__ __forceinline__ void fma4(float4& a, float g, const float4& x) {
  a.x = fmaf(g, x.x, a.x);
  a.y = fmaf(g, x.y, a.y);
  a.z = fmaf(g, x.z, a.z);
  a.w = fmaf(g, x.w, a.w);
}

__global__ void __launch_bounds__(128)
loss_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ docs, const float* __restrict__ negs,
                   int Bq, int N, int k, int V, const float* __restrict__ g_in, const float* __restrict__ g_hard,
                   const float* __restrict__ cm_q, const float* __restrict__ cm_n, float cq, float cn,
                   const float* __restrict__ g_loss, float* __restrict__ dq, float* __restrict__ dnegs) {
  __shared__ __align__(16) float Gs[VT][VK];
  const int c4 = blockIdx.x * 128 + threadIdx.x;  // float4 column index
  const int n4 = V >> 2;
  const bool live = c4 < n4;
  const int i0 = blockIdx.y * VT;
  const float gl = *g_loss;
  float4 acc[VT];
#pragma unroll
  for (int ii = 0; ii < VT; ++ii) acc[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* docs4 = reinterpret_cast<const float4*>(docs);
  for (int j0 = 0; j0 < N; j0 += VK) {
    __syncthreads();
    for (int e = threadIdx.x; e < VT * VK; e += 128) {
      const int ii = e / VK, jj = e % VK;
      Gs[ii][jj] = (i0 + ii < Bq && j0 + jj < N) ? g_in[static_cast<size_t>(i0 + ii) * N + j0 + jj] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int jj = 0; jj < VK; jj += 4) {
      float4 d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        d[u] = (live && j0 + jj + u < N) ? __ldg(&docs4[static_cast<size_t>(j0 + jj + u) * n4 + c4])
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ii = 0; ii < VT; ++ii) {
        const float4 g = *reinterpret_cast<const float4*>(&Gs[ii][jj]);
        fma4(acc[ii], g.x, d[0]);
        fma4(acc[ii], g.y, d[1]);
        fma4(acc[ii], g.z, d[2]);
        fma4(acc[ii], g.w, d[3]);
      }
    }
  }
  if (!live) return;
  const float4 mq = __ldg(reinterpret_cast<const float4*>(cm_q) + c4);
  const float4 mn = __ldg(reinterpret_cast<const float4*>(cm_n) + c4);
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const float4* negs4 = reinterpret_cast<const float4*>(negs);
  float4* dq4 = reinterpret_cast<float4*>(dq);
  float4* dnegs4 = reinterpret_cast<float4*>(dnegs);
#pragma unroll
  for (int ii = 0; ii < VT; ++ii) {
    const int i = i0 + ii;
    if (i >= Bq) break;
    const float4 qv = __ldg(&q4[static_cast<size_t>(i) * n4 + c4]);
    float4 a = acc[ii];
    for (int kk = 0; kk < k; ++kk) {
      const float gh = __ldg(&g_hard[static_cast<size_t>(i) * k + kk]);
      const size_t o = (static_cast<size_t>(i) * k + kk) * n4 + c4;
      const float4 nv = __ldg(&negs4[o]);
      fma4(a, gh, nv);
      dnegs4[o] = make_float4(gl * (gh * qv.x + cn * mn.x), gl * (gh * qv.y + cn * mn.y),
                              gl * (gh * qv.z + cn * mn.z), gl * (gh * qv.w + cn * mn.w));
    }
    dq4[static_cast<size_t>(i) * n4 + c4] = make_float4(gl * (a.x + cq * mq.x), gl * (a.y + cq * mq.y),
                                                        gl * (a.z + cq * mq.z), gl * (a.w + cq * mq.w));
  }
}

__global__ void __launch_bounds__(128)
loss_bwd_ddocs_kernel(const float* __restrict__ q, int Bq, int N, int V, int label_offset,
                      const float* __restrict__ g_in, const float* __restrict__ cm_d, float cd,
                      const float* __restrict__ g_loss, float* __restrict__ ddocs) {
  __shared__ __align__(16) float Gt[VT][VK];  // Gt[jj][ii] = g_in[i0+ii][j0+jj]
  const int c4 = blockIdx.x * 128 + threadIdx.x;
  const int n4 = V >> 2;
  const bool live = c4 < n4;
  const int j0 = blockIdx.y * VT;
  const float gl = *g_loss;
  float4 acc[VT];
#pragma unroll
  for (int jj = 0; jj < VT; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* q4 = reinterpret_cast<const float4*>(q);
  for (int i0 = 0; i0 < Bq; i0 += VK) {
    __syncthreads();
    for (int e = threadIdx.x; e < VT * VK; e += 128) {
      const int ii = e / VT, jj = e % VT;  // consecutive threads read consecutive j of one query row
      Gt[jj][ii] = (i0 + ii < Bq && j0 + jj < N) ? g_in[static_cast<size_t>(i0 + ii) * N + j0 + jj] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ii = 0; ii < VK; ii += 4) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        x[u] = (live && i0 + ii + u < Bq) ? __ldg(&q4[static_cast<size_t>(i0 + ii + u) * n4 + c4])
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int jj = 0; jj < VT; ++jj) {
        const float4 g = *reinterpret_cast<const float4*>(&Gt[jj][ii]);
        fma4(acc[jj], g.x, x[0]);
        fma4(acc[jj], g.y, x[1]);
        fma4(acc[jj], g.z, x[2]);
        fma4(acc[jj], g.w, x[3]);
      }
    }
  }
  if (!live) return;
  const float4 md = __ldg(reinterpret_cast<const float4*>(cm_d) + c4);
  float4* dd4 = reinterpret_cast<float4*>(ddocs);
#pragma unroll
  for (int jj = 0; jj < VT; ++jj) {
    const int j = j0 + jj;
    if (j >= N) break;
    const float f = (j >= label_offset && j < label_offset + Bq) ? cd : 0.f;
    dd4[static_cast<size_t>(j) * n4 + c4] = make_float4(gl * (acc[jj].x + f * md.x), gl * (acc[jj].y + f * md.y),
                                                        gl * (acc[jj].z + f * md.z), gl * (acc[jj].w + f * md.w));
  }
}

}  // namespace

size_t flops_workspace_bytes(int N, int V) {
  (void)N;
  return align256(static_cast<size_t>((V + kFlopsThreads - 1) / kFlopsThreads) * sizeof(float2));
}

int flops_nnz_launch(const float* rep, int N, int V, float* colmean, float* out2, void* workspace,
                     size_t ws_bytes, cudaStream_t stream) {
  if (!rep || !colmean || !out2) return set_error(kErrInvalidArg, "splade_flops_nnz: null pointer argument");
  if (N < 1 || V < 1) return set_error(kErrInvalidArg, "splade_flops_nnz: bad sizes N=%d V=%d", N, V);
  const size_t need = flops_workspace_bytes(N, V);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "splade_flops_nnz: workspace too small (%zu < %zu bytes)", ws_bytes, need);
  const int nblocks = (V + kFlopsThreads - 1) / kFlopsThreads;
  float2* partials = static_cast<float2*>(workspace);
  flops_cols_kernel<<<nblocks, kFlopsThreads, 0, stream>>>(rep, N, V, colmean, partials);
  SPLADE_CUDA_OK(cudaGetLastError());
  flops_final_kernel<<<1, 256, 0, stream>>>(partials, nblocks, N, out2);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

size_t scores_workspace_bytes(int Bq, int N, int V) {
  if (Bq < 1 || N < 1 || V < 1) return 0;
  int ksplit, chunk;
  scores_plan(Bq, N, V, &ksplit, &chunk);
  const size_t cuda_core = align256(static_cast<size_t>(ksplit) * Bq * N * sizeof(float));
  return std::max(cuda_core, scores_tc_workspace_bytes(Bq, N, V));
}

int scores_launch(const float* q, int Bq, const float* docs, int N, int V, int round_bf16, float* raw,
                  void* workspace, size_t ws_bytes, cudaStream_t stream, bool docs_prepacked) {
  if (!q || !docs || !raw) return set_error(kErrInvalidArg, "splade_scores: null pointer argument");
  if (Bq < 1 || N < 1 || V < 1) return set_error(kErrInvalidArg, "splade_scores: bad sizes Bq=%d N=%d V=%d", Bq, N, V);
  const size_t n_out = static_cast<size_t>(Bq) * N;
  if (round_bf16) {
    // autocast mode: bf16 operands, fp32 accumulate -> tcgen05 tensor-core path (scores_tc.cu)
    float* tc_partial = nullptr;
    int tc_ksplit = 0;
    const int rc = scores_tc_launch(q, Bq, docs, N, V, workspace, ws_bytes, &tc_partial, &tc_ksplit, docs_prepacked, stream);
    if (rc != kOk) return rc;
    scores_reduce_kernel<<<static_cast<unsigned>((n_out + 7) / 8), 256, 0, stream>>>(tc_partial, tc_ksplit, n_out,
                                                                                   raw);
    SPLADE_CUDA_OK(cudaGetLastError());
    return kOk;
  }
  // fp32 mode (no autocast): full-precision CUDA-core path
  int ksplit, chunk;
  scores_plan(Bq, N, V, &ksplit, &chunk);
  const size_t need = align256(static_cast<size_t>(ksplit) * Bq * N * sizeof(float));
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "splade_scores: workspace too small (%zu < %zu bytes)", ws_bytes, need);
  float* partial = static_cast<float*>(workspace);
  dim3 grid((N + TJ - 1) / TJ, (Bq + TI - 1) / TI, ksplit);
  scores_partial_kernel<<<grid, 256, 0, stream>>>(q, Bq, docs, N, V, round_bf16, chunk, partial);
  SPLADE_CUDA_OK(cudaGetLastError());
  const size_t n = static_cast<size_t>(Bq) * N;
  scores_reduce_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, stream>>>(partial, ksplit, n, raw);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int rowdot_launch(const float* q, const float* docs, int Bq, int k, int V, float* out, cudaStream_t stream) {
  if (!q || !docs || !out) return set_error(kErrInvalidArg, "splade_rowdot: null pointer argument");
  if (Bq < 1 || k < 1 || V < 1) return set_error(kErrInvalidArg, "splade_rowdot: bad sizes Bq=%d k=%d V=%d", Bq, k, V);
  rowdot_kernel<<<Bq * k, kRowdotThreads, 0, stream>>>(q, docs, k, V, out);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int v33_scalars_launch(const splade_v33_scalar_args* a, cudaStream_t stream) {
  if (!a) return set_error(kErrInvalidArg, "splade_v33_scalars: null args");
  if (a->Bq < 1 || a->N < 1 || a->k < 1)
    return set_error(kErrInvalidArg, "splade_v33_scalars: bad sizes Bq=%d N=%d k=%d", a->Bq, a->N, a->k);
  if (!a->raw_inbatch || !a->raw_hard || !a->flops3 || !a->scalars || !a->g_inbatch || !a->g_hard ||
      !a->row_scratch)
    return set_error(kErrInvalidArg, "splade_v33_scalars: null pointer argument");
  if (a->label_offset < 0 || a->label_offset + a->Bq > a->N)
    return set_error(kErrInvalidArg, "splade_v33_scalars: labels [%d,%d) outside the %d in-batch documents",
                     a->label_offset, a->label_offset + a->Bq, a->N);
  if (a->has_kd && (!a->teacher_kd || a->N != a->Bq))
    return set_error(kErrInvalidArg, "splade_v33_scalars: KL-KD needs teacher_scores [Bq,Bq] and N == Bq");
  if (a->has_mmse && (!a->raw_pos || !a->teacher_pos || !a->teacher_neg))
    return set_error(kErrInvalidArg, "splade_v33_scalars: MarginMSE needs raw_pos, teacher_pos and teacher_neg");
  if (!(a->temperature > 0.f) || (a->has_kd && !(a->kd_temperature > 0.f)))
    return set_error(kErrInvalidArg, "splade_v33_scalars: temperatures must be positive");
  v33_scalars_kernel<<<(a->Bq + 7) / 8, 256, 0, stream>>>(*a);
  v33_finalize_kernel<<<1, 256, 0, stream>>>(*a);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

// Workspace of splade_v33_loss_fwd: [FLOPS partials 3 x nblk float2 | score GEMM workspace | raw_in [Bq,N] |
// raw_hard [Bq,k] | raw_pos [Bq] | flops3 [6] | row_scratch [3 Bq] | dot-product partials [nblk][Bq (k+1)]].
static size_t v33_flops_area(int Bq, int k, int V) {
  return align256(3 * std::max(flops_workspace_bytes(Bq, V), flops_workspace_bytes(Bq * k, V)));
}

size_t v33_loss_fwd_workspace_bytes(int Bq, int N, int V, int k) {
  if (Bq < 1 || N < 1 || V < 1 || k < 1) return 0;
  const size_t nblk = (static_cast<size_t>((V + 7) / 8 * 8) + 255) / 256;
  return v33_flops_area(Bq, k, V) + align256(scores_workspace_bytes(Bq, N, V)) +
         align256(static_cast<size_t>(Bq) * N * 4) + align256(static_cast<size_t>(Bq) * k * 4) +
         align256(static_cast<size_t>(Bq) * 4) + align256(6 * 4) + align256(static_cast<size_t>(3) * Bq * 4) +
         align256(nblk * Bq * (k + 1) * 4);
}

// The scores workspace inside the forward workspace (layout of v33_loss_fwd_launch below).
void* v33_loss_fwd_scores_ws(const splade_v33_fwd_args* a) {
  return static_cast<uint8_t*>(a->workspace) + v33_flops_area(a->Bq, a->k, a->V);
}

int v33_loss_fwd_launch(const splade_v33_fwd_args* a, cudaStream_t stream, bool docs_prepacked) {
  if (!a) return set_error(kErrInvalidArg, "splade_v33_loss_fwd: null args");
  if (a->Bq < 1 || a->N < 1 || a->k < 1 || a->V < 1)
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd: bad sizes Bq=%d N=%d k=%d V=%d", a->Bq, a->N, a->k, a->V);
  if (!a->q || !a->docs || !a->negs || !a->colmean || !a->scalars || !a->g_inbatch || !a->g_hard)
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd: null pointer argument");
  if (a->label_offset < 0 || a->label_offset + a->Bq > a->N)
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd: labels outside the in-batch documents");
  const int Bq = a->Bq, N = a->N, k = a->k, V = a->V;
  const size_t need = v33_loss_fwd_workspace_bytes(Bq, N, V, k);
  if (!a->workspace || a->ws_bytes < need)
    return set_error(kErrWorkspace, "splade_v33_loss_fwd: workspace too small (%zu < %zu bytes)", a->ws_bytes, need);
  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  const size_t flops_area = v33_flops_area(Bq, k, V);
  void* w_flops = ws;
  ws += flops_area;
  const size_t scores_ws = align256(scores_workspace_bytes(Bq, N, V));
  void* w_scores = ws;
  ws += scores_ws;
  float* raw_in = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(Bq) * N * 4);
  float* raw_hard = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(Bq) * k * 4);
  float* raw_pos = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(Bq) * 4);
  float* flops3 = reinterpret_cast<float*>(ws);
  ws += align256(6 * 4);
  float* row_scratch = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(3) * Bq * 4);
  float* rd_part = reinterpret_cast<float*>(ws);

  splade_v33_scalar_args sa{};
  sa.Bq = Bq;
  sa.N = N;
  sa.k = k;
  sa.label_offset = a->label_offset;
  sa.round_bf16 = a->round_bf16;
  sa.has_kd = a->has_kd;
  sa.has_mmse = a->has_mmse;
  sa.temperature = a->temperature;
  sa.kd_temperature = a->kd_temperature;
  sa.lambda_q = a->lambda_q;
  sa.lambda_d = a->lambda_d;
  sa.lambda_neg = a->lambda_neg;
  sa.lambda_kd = a->lambda_kd;
  sa.lambda_mmse = a->lambda_mmse;
  sa.raw_inbatch = raw_in;
  sa.raw_hard = raw_hard;
  sa.raw_pos = a->has_mmse ? raw_pos : nullptr;
  sa.teacher_kd = a->teacher_kd;
  sa.teacher_pos = a->teacher_pos;
  sa.teacher_neg = a->teacher_neg;
  sa.flops3 = flops3;
  sa.scalars = a->scalars;
  sa.g_inbatch = a->g_inbatch;
  sa.g_hard = a->g_hard;
  sa.row_scratch = row_scratch;
  if (a->has_kd && (!a->teacher_kd || N != Bq))
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd: KL-KD needs teacher_scores [Bq,Bq] and N == Bq");
  if (a->has_mmse && (!a->teacher_pos || !a->teacher_neg))
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd: MarginMSE needs teacher_pos and teacher_neg");
  if (!(a->temperature > 0.f) || (a->has_kd && !(a->kd_temperature > 0.f)))
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd: temperatures must be positive");

  const float* local_docs = a->docs + static_cast<size_t>(a->label_offset) * V;
  if (k > 8 || Bq > 256) {
    // ---- general shapes: one kernel per term (the launch sequence of round 1)
    const size_t one = flops_area / 3;
    int rc = flops_nnz_launch(a->q, Bq, V, a->colmean, flops3, w_flops, one, stream);
    if (rc != kOk) return rc;
    rc = flops_nnz_launch(local_docs, Bq, V, a->colmean + V, flops3 + 2, w_flops, one, stream);
    if (rc != kOk) return rc;
    rc = flops_nnz_launch(a->negs, Bq * k, V, a->colmean + 2 * static_cast<size_t>(V), flops3 + 4, w_flops, one, stream);
    if (rc != kOk) return rc;
    rc = scores_launch(a->q, Bq, a->docs, N, V, a->round_bf16, raw_in, w_scores, scores_ws, stream,
                       docs_prepacked && a->round_bf16);
    if (rc != kOk) return rc;
    rc = rowdot_launch(a->q, a->negs, Bq, k, V, raw_hard, stream);
    if (rc != kOk) return rc;
    if (a->has_mmse) {
      rc = rowdot_launch(a->q, local_docs, Bq, 1, V, raw_pos, stream);
      if (rc != kOk) return rc;
    }
    return v33_scalars_launch(&sa, stream);
  }

  // ---- fused forward: prep -> score GEMM -> reduce2 -> scalars (one CTA)
  const int Vp = (V + 7) / 8 * 8;
  const int nblk = (Vp + 255) / 256;
  PrepParams pp{};
  pp.q = a->q, pp.docs = a->docs, pp.negs = a->negs;
  pp.Bq = Bq, pp.N = N, pp.k = k, pp.V = V, pp.Vp = Vp, pp.label_offset = a->label_offset;
  pp.pack_q = a->round_bf16 ? 1 : 0;
  pp.pack_docs = (a->round_bf16 && !docs_prepacked) ? 1 : 0;
  pp.want_pos = a->has_mmse ? 1 : 0;
  pp.colmean = a->colmean;
  pp.fpart = static_cast<float2*>(w_flops);
  pp.rd_part = rd_part;
  if (a->round_bf16) {
    int Vp2 = 0;
    pp.q16 = static_cast<__nv_bfloat16*>(w_scores);
    pp.d16 = static_cast<__nv_bfloat16*>(scores_tc_docs16(w_scores, Bq, N, V, &Vp2));
    if (Vp2 != Vp) return set_error(kErrInvalidArg, "splade_v33_loss_fwd: internal layout mismatch");
  }
  if (k <= 1) {
    v33_prep_kernel<1><<<nblk, 256, 0, stream>>>(pp);
  } else if (k <= 2) {
    v33_prep_kernel<2><<<nblk, 256, 0, stream>>>(pp);
  } else if (k <= 4) {
    v33_prep_kernel<4><<<nblk, 256, 0, stream>>>(pp);
  } else {
    v33_prep_kernel<8><<<nblk, 256, 0, stream>>>(pp);
  }
  SPLADE_CUDA_OK(cudaGetLastError());

  float* partial = nullptr;
  int ksplit = 0;
  if (a->round_bf16) {
    const int rc = scores_tc_launch(a->q, Bq, a->docs, N, V, w_scores, scores_ws, &partial, &ksplit,
                                    /*docs_prepacked=*/true, stream, /*q_prepacked=*/true);
    if (rc != kOk) return rc;
  } else {
    int chunk;
    scores_plan(Bq, N, V, &ksplit, &chunk);
    partial = static_cast<float*>(w_scores);
    dim3 grid((N + TJ - 1) / TJ, (Bq + TI - 1) / TI, ksplit);
    scores_partial_kernel<<<grid, 256, 0, stream>>>(a->q, Bq, a->docs, N, V, 0, chunk, partial);
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  const size_t nA = static_cast<size_t>(Bq) * N, nB = static_cast<size_t>(Bq) * (k + 1);
  v33_reduce2_kernel<<<static_cast<unsigned>((nA + nB + 7) / 8), 256, 0, stream>>>(
      partial, ksplit, nA, raw_in, rd_part, nblk, nB, raw_hard, k, a->has_mmse ? raw_pos : nullptr);
  SPLADE_CUDA_OK(cudaGetLastError());
  v33_scalars_fused_kernel<<<1, 1024, 0, stream>>>(sa, static_cast<const float2*>(w_flops), nblk, Bq * k);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int loss_bwd_launch(const float* q, const float* docs, const float* negs, int Bq, int N, int k, int V,
                    int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                    const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                    const float* g_loss, float* dq, float* ddocs, float* dnegs, cudaStream_t stream) {
  if (!q || !docs || !negs || !g_inbatch || !g_hard || !colmean_q || !colmean_d || !colmean_n || !g_loss || !dq ||
      !ddocs || !dnegs)
    return set_error(kErrInvalidArg, "splade_loss_bwd: null pointer argument");
  if (Bq < 1 || N < 1 || k < 1 || V < 1)
    return set_error(kErrInvalidArg, "splade_loss_bwd: bad sizes Bq=%d N=%d k=%d V=%d", Bq, N, k, V);
  auto aligned16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  const bool vec_ok = (V % 4 == 0) && aligned16(q) && aligned16(docs) && aligned16(negs) && aligned16(dq) &&
                      aligned16(ddocs) && aligned16(dnegs) && aligned16(colmean_q) && aligned16(colmean_d) &&
                      aligned16(colmean_n);
  if (vec_ok) {
    const unsigned cx = static_cast<unsigned>((V / 4 + 127) / 128);
    loss_bwd_dq_kernel<<<dim3(cx, (Bq + VT - 1) / VT), 128, 0, stream>>>(
        q, docs, negs, Bq, N, k, V, g_inbatch, g_hard, colmean_q, colmean_n, cq, cn, g_loss, dq, dnegs);
    loss_bwd_ddocs_kernel<<<dim3(cx, (N + VT - 1) / VT), 128, 0, stream>>>(q, Bq, N, V, label_offset, g_inbatch,
                                                                         colmean_d, cd, g_loss, ddocs);
  } else {
    loss_bwd_kernel<<<(V + 127) / 128, 128, 0, stream>>>(q, docs, negs, Bq, N, k, V, label_offset, g_inbatch,
                                                         g_hard, colmean_q, colmean_d, colmean_n, cq, cd, cn,
                                                         g_loss, dq, ddocs, dnegs);
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

// dq and dnegs only (vectorised kernel): the peer exchange computes ddocs itself (peer.cu).
int loss_bwd_dq_launch(const float* q, const float* docs, const float* negs, int Bq, int N, int k, int V,
                       const float* g_inbatch, const float* g_hard, const float* colmean_q, const float* colmean_n,
                       float cq, float cn, const float* g_loss, float* dq, float* dnegs, cudaStream_t stream) {
  if (!q || !docs || !negs || !g_inbatch || !g_hard || !colmean_q || !colmean_n || !g_loss || !dq || !dnegs)
    return set_error(kErrInvalidArg, "splade_peer_loss_bwd: null pointer argument");
  auto aligned16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  if (V % 4 != 0 || !aligned16(q) || !aligned16(docs) || !aligned16(negs) || !aligned16(dq) || !aligned16(dnegs) ||
      !aligned16(colmean_q) || !aligned16(colmean_n))
    return set_error(kErrUnsupportedShape, "splade_peer_loss_bwd: V %% 4 == 0 and 16-byte aligned rows are required");
  const unsigned cx = static_cast<unsigned>((V / 4 + 127) / 128);
  loss_bwd_dq_kernel<<<dim3(cx, (Bq + VT - 1) / VT), 128, 0, stream>>>(q, docs, negs, Bq, N, k, V, g_inbatch, g_hard,
                                                                      colmean_q, colmean_n, cq, cn, g_loss, dq, dnegs);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // namespace splade
