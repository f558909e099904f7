// Gather-driven sparse backward of the fused SPLADE-max head (sm_100a).
//
// The reference back-propagates through dense [B,S,V] tensors (autograd of src/model/splade_modern.py:69-83
// run from src/train/cli/train_v33_ddp.py:364): value_selecting_reduction_backward scatters g to the argmax
// position, the element-wise chain multiplies by mask / (1+x) / relu', and two dense GEMMs produce dX and dW.
// Because max-pooling routes the whole gradient of (doc b, term v) to ONE sequence position, dlogits has
// exactly one non-zero per (b,v):
//     coef[b,v] = g_rep[b,v] * d/dx log1p(x) = g_rep[b,v] * exp(-rep[b,v])      (rep > 0, else 0)
// so
//     dbias[v]      = sum_b coef[b,v]
//     dW[v,:]       = sum_b coef[b,v] * hidden[b, argmax[b,v], :]                 (gather of B rows per term)
//     dHidden[b,s,:]= sum_{v : argmax[b,v]=s} coef[b,v] * W[v,:]                  (gather after bucketing by s)
// Both gathers read 16-bit rows that live in the 126 MB L2 (hidden 25 MB, W 77 MB at the headline shape);
// the kernels are L2-gather-bound, not tensor-core work.
//
// Kernels: (1) bucket_count / bucket_scan / bucket_scatter: deterministic per-document counting sort of the active
//              terms by argmax position, one warp per (document, 1024-term chunk);
//          (2) dhidden_kernel: one warp per (b,s) row sums coef * W[v,:] over its bucket;
//          (3) dw_kernel: one warp per vocabulary row sums coef * hidden[b,argmax,:] over the batch.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace splade {
namespace {

constexpr int kBucketThreads = 1024;

__device__ __forceinline__ float coef_of(float g, float rep) { return rep > 0.f ? g * expf(-rep) : 0.f; }

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// ---------------------------------------------------------------------------------------------
// (1) DETERMINISTIC per-document counting sort of the active (b,v) pairs by position.
//     The vocabulary of every document is cut into chunks of kChunk terms; ONE WARP walks a chunk in term order:
//       count   : per-(doc, chunk, position) counts            (warp-private shared-memory counters)
//       scan    : row_ptr[b][s] and, per chunk, the first slot of each position (prefix over chunks)
//       scatter : the same walk again; slot = chunk base + running count + rank among the lanes of the same
//                 position (lanes are in term order)  => every bucket lists its terms in ascending term id,
//                 so dHidden sums in a fixed order: run-to-run bit-identical.
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 1024;  // vocabulary entries per warp

__global__ void __launch_bounds__(256)
bucket_count_kernel(const float* __restrict__ g_rep, const float* __restrict__ rep,
                    const int32_t* __restrict__ argmax, int B, int S, int V, int n_chunks, int warps_per_cta,
                    int32_t* __restrict__ chunk_counts /*[B][n_chunks][S]*/) {
  extern __shared__ int32_t sm[];  // [warps_per_cta][S]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps_per_cta) return;
  const long long job = static_cast<long long>(blockIdx.x) * warps_per_cta + warp;  // (b, chunk)
  if (job >= static_cast<long long>(B) * n_chunks) return;
  const int b = static_cast<int>(job / n_chunks), ch = static_cast<int>(job % n_chunks);
  int32_t* cnt = sm + static_cast<size_t>(warp) * S;
  for (int s = lane; s < S; s += 32) cnt[s] = 0;
  __syncwarp();
  const int v0 = ch * kChunk, v1 = min(V, v0 + kChunk);
  const size_t base = static_cast<size_t>(b) * V;
  for (int vb = v0; vb < v1; vb += 128) {  // four 32-term steps per trip: their loads are issued together
    float gg[4], rr[4];
    int pp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = vb + 32 * u + lane;
      gg[u] = rr[u] = 0.f;
      pp[u] = 0;
      if (v < v1) {
        gg[u] = g_rep[base + v];
        rr[u] = rep[base + v];
        pp[u] = argmax[base + v];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool act = coef_of(gg[u], rr[u]) != 0.f;
      const uint32_t am = __ballot_sync(0xffffffffu, act);
      if (act) {
        const uint32_t peers = __match_any_sync(am, pp[u]);
        if (lane == __ffs(peers) - 1) cnt[pp[u]] += __popc(peers);  // one lane per distinct position: no race
      }
      __syncwarp();
    }
  }
  int32_t* out = chunk_counts + (static_cast<size_t>(b) * n_chunks + ch) * S;
  for (int s = lane; s < S; s += 32) out[s] = cnt[s];
}

// chunk_counts -> row_ptr [B][S+1] (exclusive scan over positions of the per-position totals) and, in place,
// chunk_counts[b][ch][s] := row_ptr[b][s] + sum_{ch' < ch} counts[b][ch'][s].  One CTA per document.
__global__ void __launch_bounds__(kBucketThreads)
bucket_scan_kernel(int S, int n_chunks, int32_t* __restrict__ chunk_counts, int32_t* __restrict__ row_ptr) {
  __shared__ int32_t warp_tot[32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  int32_t* cc = chunk_counts + static_cast<size_t>(b) * n_chunks * S;
  int32_t* rp = row_ptr + static_cast<size_t>(b) * (S + 1);
  const int per = (S + kBucketThreads - 1) / kBucketThreads;
  const int lo = min(S, tid * per), hi = min(S, lo + per);
  int tsum = 0;
  for (int s = lo; s < hi; ++s) {
    int tot = 0;
    for (int ch = 0; ch < n_chunks; ++ch) tot += cc[static_cast<size_t>(ch) * S + s];
    rp[s] = tot;  // per-position total, turned into the exclusive offset below
    tsum += tot;
  }
  int incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += n;
  }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    const int w = warp_tot[tid];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, wi, o);
      if (tid >= o) wi += n;
    }
    warp_tot[tid] = wi - w;  // exclusive
  }
  __syncthreads();
  int run = warp_tot[tid >> 5] + incl - tsum;
  for (int s = lo; s < hi; ++s) {
    const int tot = rp[s];
    rp[s] = run;
    int acc = run;
    for (int ch = 0; ch < n_chunks; ++ch) {  // first slot of position s for each chunk
      const size_t i = static_cast<size_t>(ch) * S + s;
      const int c = cc[i];
      cc[i] = acc;
      acc += c;
    }
    run += tot;
  }
  if (tid == kBucketThreads - 1) rp[S] = run;  // total number of active terms of this document
}

__global__ void __launch_bounds__(256)
bucket_scatter_kernel(const float* __restrict__ g_rep, const float* __restrict__ rep,
                      const int32_t* __restrict__ argmax, int B, int S, int V, int n_chunks, int warps_per_cta,
                      const int32_t* __restrict__ chunk_base, int32_t* __restrict__ sorted_v,
                      float* __restrict__ sorted_c) {
  extern __shared__ int32_t sm[];  // [warps_per_cta][S] running slot per position
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps_per_cta) return;
  const long long job = static_cast<long long>(blockIdx.x) * warps_per_cta + warp;
  if (job >= static_cast<long long>(B) * n_chunks) return;
  const int b = static_cast<int>(job / n_chunks), ch = static_cast<int>(job % n_chunks);
  int32_t* cur = sm + static_cast<size_t>(warp) * S;
  const int32_t* cb = chunk_base + (static_cast<size_t>(b) * n_chunks + ch) * S;
  for (int s = lane; s < S; s += 32) cur[s] = cb[s];
  __syncwarp();
  const int v0 = ch * kChunk, v1 = min(V, v0 + kChunk);
  const size_t base = static_cast<size_t>(b) * V;
  for (int vb = v0; vb < v1; vb += 128) {
    float cc4[4];
    int pp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = vb + 32 * u + lane;
      cc4[u] = 0.f;
      pp[u] = 0;
      if (v < v1) {
        cc4[u] = coef_of(g_rep[base + v], rep[base + v]);
        pp[u] = argmax[base + v];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = vb + 32 * u + lane;
      const float c = cc4[u];
      const int pos = pp[u];
      const bool act = c != 0.f;
      const uint32_t am = __ballot_sync(0xffffffffu, act);
      uint32_t peers = 0;
      int slot = 0;
      if (act) {
        peers = __match_any_sync(am, pos);
        slot = cur[pos] + __popc(peers & ((1u << lane) - 1u));  // lanes are in term order
      }
      __syncwarp();
      if (act) {
        if (lane == __ffs(peers) - 1) cur[pos] += __popc(peers);
        sorted_v[base + slot] = v;
        sorted_c[base + slot] = c;
      }
      __syncwarp();
    }
  }
}

// acc[8] += c * (8 bf16 packed in a uint4)
__device__ __forceinline__ void fma_bf16x8(float (&acc)[8], float c, const uint4& u) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[2 * i] = fmaf(c, __uint_as_float(w[i] << 16), acc[2 * i]);
    acc[2 * i + 1] = fmaf(c, __uint_as_float(w[i] & 0xFFFF0000u), acc[2 * i + 1]);
  }
}

__device__ __forceinline__ void store_row8(float* dst, const float (&acc)[8], float scale) {
  float4 a = make_float4(acc[0] * scale, acc[1] * scale, acc[2] * scale, acc[3] * scale);
  float4 b = make_float4(acc[4] * scale, acc[5] * scale, acc[6] * scale, acc[7] * scale);
  reinterpret_cast<float4*>(dst)[0] = a;
  reinterpret_cast<float4*>(dst)[1] = b;
}

// ---------------------------------------------------------------------------------------------
// (2) dHidden: one warp per (b,s) row.  kChunks = ceil(H / 256); lane l owns h = i*256 + l*8 .. +8.
// ---------------------------------------------------------------------------------------------
template <int kChunks>
__global__ void __launch_bounds__(256)
dhidden_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ sorted_v,
               const float* __restrict__ sorted_c, const uint16_t* __restrict__ W, int B, int S, int H, int V,
               float* __restrict__ d_hidden) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * S) return;
  const int b = static_cast<int>(row / S);
  const int s = static_cast<int>(row - static_cast<long long>(b) * S);
  const int32_t* rp = row_ptr + static_cast<size_t>(b) * (S + 1);
  const int beg = rp[s], end = rp[s + 1];
  const size_t lbase = static_cast<size_t>(b) * V;

  float acc[kChunks][8];
#pragma unroll
  for (int i = 0; i < kChunks; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;

  for (int e0 = beg; e0 < end; e0 += 32) {
    const int n = min(32, end - e0);
    int my_v = 0;
    float my_c = 0.f;
    if (lane < n) {
      my_v = sorted_v[lbase + e0 + lane];
      my_c = sorted_c[lbase + e0 + lane];
    }
#pragma unroll 4
    for (int l = 0; l < n; ++l) {
      const int v = __shfl_sync(0xffffffffu, my_v, l);
      const float c = __shfl_sync(0xffffffffu, my_c, l);
      const uint16_t* wrow = W + static_cast<size_t>(v) * H;
#pragma unroll
      for (int i = 0; i < kChunks; ++i) {
        const int h = i * 256 + lane * 8;
        if (h < H) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(wrow + h));
          fma_bf16x8(acc[i], c, u);
        }
      }
    }
  }
  float* out = d_hidden + static_cast<size_t>(row) * H;
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const int h = i * 256 + lane * 8;
    if (h < H) store_row8(out + h, acc[i], 1.f);
  }
}

// ---------------------------------------------------------------------------------------------
// (3) dW + dbias.  A CTA owns 32 consecutive vocabulary rows: it first stages coef[b][32] and argmax[b][32]
//     for the whole batch in shared memory with fully coalesced 128-byte reads (the three [B,V] arrays are read
//     exactly once), then each of its 8 warps produces 4 rows, skipping inactive (b,v) pairs.
// ---------------------------------------------------------------------------------------------
template <int kChunks>
__global__ void __launch_bounds__(256)
dw_kernel(const float* __restrict__ g_rep, const float* __restrict__ rep, const int32_t* __restrict__ argmax,
          const uint16_t* __restrict__ hidden, int B, int S, int H, int V, float* __restrict__ dW,
          float* __restrict__ dbias) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  float* s_c = reinterpret_cast<float*>(dw_smem);                  // [B][32]
  int32_t* s_p = reinterpret_cast<int32_t*>(s_c + static_cast<size_t>(B) * 32);  // [B][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v0 = blockIdx.x * 32;

  for (int idx = threadIdx.x; idx < B * 32; idx += 256) {
    const int b = idx >> 5, j = idx & 31;
    const int v = v0 + j;
    float c = 0.f;
    int pos = 0;
    if (v < V) {
      const size_t o = static_cast<size_t>(b) * V + v;
      c = coef_of(g_rep[o], rep[o]);
      pos = argmax[o];
    }
    s_c[idx] = c;
    s_p[idx] = pos;
  }
  __syncthreads();

  for (int r = 0; r < 4; ++r) {
    const int j = warp * 4 + r;
    const int v = v0 + j;
    if (v >= V) break;  // warp-uniform
    float acc[kChunks][8];
#pragma unroll
    for (int i = 0; i < kChunks; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
    float bsum = 0.f;
#pragma unroll 4
    for (int b = 0; b < B; ++b) {
      const float c = s_c[b * 32 + j];  // broadcast
      if (c == 0.f) continue;           // warp-uniform
      bsum += c;
      const uint16_t* hrow = hidden + (static_cast<size_t>(b) * S + s_p[b * 32 + j]) * H;
#pragma unroll
      for (int i = 0; i < kChunks; ++i) {
        const int h = i * 256 + lane * 8;
        if (h < H) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(hrow + h));
          fma_bf16x8(acc[i], c, u);
        }
      }
    }
    if (lane == 0) dbias[v] = bsum;
    float* out = dW + static_cast<size_t>(v) * H;
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const int h = i * 256 + lane * 8;
      if (h < H) store_row8(out + h, acc[i], 1.f);
    }
  }
}

}  // namespace

size_t head_bwd_workspace_bytes(int B, int S, int V) {
  const size_t n_chunks = (static_cast<size_t>(V) + kChunk - 1) / kChunk;
  return align256(static_cast<size_t>(B) * (S + 1) * sizeof(int32_t)) +
         align256(static_cast<size_t>(B) * n_chunks * S * sizeof(int32_t)) +
         2 * align256(static_cast<size_t>(B) * V * sizeof(int32_t));
}

int head_bwd_launch(const float* g_rep, const float* rep, const int32_t* argmax, const void* hidden,
                    const void* W, int B, int S, int H, int V, float* d_hidden, float* dW, float* dbias,
                    void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (!g_rep || !rep || !argmax || !hidden || !W || !d_hidden || !dW || !dbias)
    return set_error(kErrInvalidArg, "splade_head_bwd: null pointer argument");
  if (B < 0 || S < 1 || V < 1 || H < 1)
    return set_error(kErrInvalidArg, "splade_head_bwd: bad sizes B=%d S=%d H=%d V=%d", B, S, H, V);
  if (H % 8 != 0 || H > 1024)
    return set_error(kErrUnsupportedShape, "splade_head_bwd: hidden size %d must be a multiple of 8 and <= 1024",
                     H);
  const size_t bucket_smem = static_cast<size_t>(S) * sizeof(int32_t);
  if (bucket_smem > 200 * 1024)
    return set_error(kErrUnsupportedShape, "splade_head_bwd: sequence length %d too long for the bucket pass", S);
  if (static_cast<size_t>(B) * 32 * 8 > 200 * 1024)
    return set_error(kErrUnsupportedShape, "splade_head_bwd: batch %d too large for the dW staging buffer (<= 800)", B);
  int rc = require_sm100("splade_head_bwd");
  if (rc != kOk) return rc;
  const size_t need = head_bwd_workspace_bytes(B, S, V);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "splade_head_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);

  if (B == 0) {
    SPLADE_CUDA_OK(cudaMemsetAsync(dW, 0, static_cast<size_t>(V) * H * sizeof(float), stream));
    SPLADE_CUDA_OK(cudaMemsetAsync(dbias, 0, static_cast<size_t>(V) * sizeof(float), stream));
    return kOk;
  }

  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int32_t* row_ptr = reinterpret_cast<int32_t*>(ws);
  ws += align256(static_cast<size_t>(B) * (S + 1) * sizeof(int32_t));
  const int n_chunks = (V + kChunk - 1) / kChunk;
  int32_t* chunk_tab = reinterpret_cast<int32_t*>(ws);
  ws += align256(static_cast<size_t>(B) * n_chunks * S * sizeof(int32_t));
  int32_t* sorted_v = reinterpret_cast<int32_t*>(ws);
  ws += align256(static_cast<size_t>(B) * V * sizeof(int32_t));
  float* sorted_c = reinterpret_cast<float*>(ws);

  // warp-private position counters live in shared memory: as many warps per CTA as fit
  const int warps_per_cta = static_cast<int>(std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / bucket_smem)));
  const size_t smem_cs = bucket_smem * warps_per_cta;
  if (smem_cs > 48 * 1024) {
    SPLADE_CUDA_OK(cudaFuncSetAttribute(bucket_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem_cs)));
    SPLADE_CUDA_OK(cudaFuncSetAttribute(bucket_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem_cs)));
  }
  const long long jobs = static_cast<long long>(B) * n_chunks;
  const unsigned grid_cs = static_cast<unsigned>((jobs + warps_per_cta - 1) / warps_per_cta);
  bucket_count_kernel<<<grid_cs, 256, smem_cs, stream>>>(g_rep, rep, argmax, B, S, V, n_chunks, warps_per_cta,
                                                        chunk_tab);
  bucket_scan_kernel<<<B, kBucketThreads, 0, stream>>>(S, n_chunks, chunk_tab, row_ptr);
  bucket_scatter_kernel<<<grid_cs, 256, smem_cs, stream>>>(g_rep, rep, argmax, B, S, V, n_chunks, warps_per_cta,
                                                          chunk_tab, sorted_v, sorted_c);
  SPLADE_CUDA_OK(cudaGetLastError());

  const uint16_t* Wb = static_cast<const uint16_t*>(W);
  const uint16_t* Hb = static_cast<const uint16_t*>(hidden);
  const long long rows = static_cast<long long>(B) * S;
  const unsigned grid_h = static_cast<unsigned>((rows + 7) / 8);
  const unsigned grid_w = static_cast<unsigned>((V + 31) / 32);
  const size_t dw_smem_bytes = static_cast<size_t>(B) * 32 * 8;
  const int chunks = (H + 255) / 256;
  switch (chunks) {
    case 1:
      dhidden_kernel<1><<<grid_h, 256, 0, stream>>>(row_ptr, sorted_v, sorted_c, Wb, B, S, H, V, d_hidden);
      SPLADE_CUDA_OK(cudaFuncSetAttribute(dw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(dw_smem_bytes)));
      dw_kernel<1><<<grid_w, 256, dw_smem_bytes, stream>>>(g_rep, rep, argmax, Hb, B, S, H, V, dW, dbias);
      break;
    case 2:
      dhidden_kernel<2><<<grid_h, 256, 0, stream>>>(row_ptr, sorted_v, sorted_c, Wb, B, S, H, V, d_hidden);
      SPLADE_CUDA_OK(cudaFuncSetAttribute(dw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(dw_smem_bytes)));
      dw_kernel<2><<<grid_w, 256, dw_smem_bytes, stream>>>(g_rep, rep, argmax, Hb, B, S, H, V, dW, dbias);
      break;
    case 3:
      dhidden_kernel<3><<<grid_h, 256, 0, stream>>>(row_ptr, sorted_v, sorted_c, Wb, B, S, H, V, d_hidden);
      SPLADE_CUDA_OK(cudaFuncSetAttribute(dw_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(dw_smem_bytes)));
      dw_kernel<3><<<grid_w, 256, dw_smem_bytes, stream>>>(g_rep, rep, argmax, Hb, B, S, H, V, dW, dbias);
      break;
    default:
      dhidden_kernel<4><<<grid_h, 256, 0, stream>>>(row_ptr, sorted_v, sorted_c, Wb, B, S, H, V, d_hidden);
      SPLADE_CUDA_OK(cudaFuncSetAttribute(dw_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(dw_smem_bytes)));
      dw_kernel<4><<<grid_w, 256, dw_smem_bytes, stream>>>(g_rep, rep, argmax, Hb, B, S, H, V, dW, dbias);
      break;
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // namespace splade
