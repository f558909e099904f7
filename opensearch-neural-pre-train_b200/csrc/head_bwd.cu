// Gather-driven sparse backward of the fused SPLADE-max head (sm_100a), for ONE or SEVERAL head calls at once.
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
//
// The reference trainer calls the model three times per micro-step (query, positive, negatives:
// train_v33_ddp.py:339-343) and back-propagates once (:364); all three head calls share ONE decoder matrix, so
// their dW / dbias are one sum.  A "segment" is one head call; the kernels take up to kMaxSegs of them:
//   (1) bwd_bucket_kernel   one CTA per document (all segments in one launch): deterministic counting sort of the
//                           document's active terms by argmax position (count in shared memory, scan, scatter) AND
//                           the activity bitmap [doc][V/32] the dW sweep is driven by;
//   (2) dhidden_kernel      one warp per (doc, position) row of any segment sums coef * W[v,:] over its bucket;
//   (3) dw_kernel           ONE sweep over the vocabulary for all segments: a CTA owns 32 consecutive terms, reads
//                           one bitmap word per document, stages the coefficients of the active (doc, term) pairs
//                           only, and every warp sums coef * hidden[doc, argmax, :] for 4 terms in a fixed order
//                           (segments in call order, documents ascending) => dW / dbias written once, bit-identical
//                           run to run.  `accumulate` adds into an existing gradient buffer instead.
// Gathers read 16-bit rows that live in the 126 MB L2 (createpolicy evict_last); the streams that pass through once
// (sorted lists, fp32 outputs) are marked evict_first so that they do not push W / hidden out of it.
#include "../../include/splade_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace splade {
namespace {

constexpr int kMaxSegs = SPLADE_HEAD_BWD_MAX_SEGMENTS;
constexpr int kBucketThreads = 512;
constexpr int kBucketWarps = kBucketThreads / 32;
constexpr int kDwDocs = 256;  // documents staged per pass of the dW sweep
constexpr int kRowsPerWarp = 8;  // dHidden: consecutive positions of one document per warp

struct Seg {
  const float* g;
  const float* rep;
  const int32_t* arg;
  const uint16_t* hidden;
  float* dh;
  int32_t* row_ptr;     // [B][S+1]
  int32_t* sorted_v;    // [B][V]
  float* sorted_c;      // [B][V]
  uint32_t* bitmap;     // [B][words]
  int B, S;
  int doc0;             // first global document index of this segment
  long long row0;       // first global row GROUP (kRowsPerWarp consecutive positions of one document) of this segment
};

struct Segs {
  Seg s[kMaxSegs];
  int n;
  int total_docs;
  long long total_rows;
};

__device__ __forceinline__ float coef_of(float g, float rep) { return rep > 0.f ? g * expf(-rep) : 0.f; }

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

__device__ __forceinline__ int seg_of_doc(const Segs& p, int doc) {
  int sel = 0;
#pragma unroll
  for (int i = 1; i < kMaxSegs; ++i)
    if (i < p.n && doc >= p.s[i].doc0) sel = i;
  return sel;
}

// ---------------------------------------------------------------------------------------------
// (0) activity bitmap: bit j of word w of document b = (rep[b, 32w + j] > 0).  The autograd path computes it right
//     after the forward (rep is L2-hot then); every backward kernel below is DRIVEN by it, so in the trained
//     (sparse) regime g_rep / rep / argmax are read at the ~300 active terms of a document only, not 50 000.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rep_bitmap_kernel(const float* __restrict__ rep, int B, int V, uint32_t* __restrict__ bitmap) {
  const int words = (V + 31) >> 5;
  const long long total = static_cast<long long>(B) * words;
  const int lane = threadIdx.x & 31;
  for (long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); w < total;
       w += static_cast<long long>(gridDim.x) * 8) {
    const int b = static_cast<int>(w / words), wi = static_cast<int>(w - static_cast<long long>(b) * words);
    const int v = wi * 32 + lane;
    const bool act = v < V && rep[static_cast<size_t>(b) * V + v] > 0.f;
    const uint32_t m = __ballot_sync(0xffffffffu, act);
    if (lane == 0) bitmap[w] = m;
  }
}

// ---------------------------------------------------------------------------------------------
// (1) per-document counting sort by position.  A CLUSTER of C CTAs owns one document (C is chosen so that the grid
//     fills the machine in one wave: 192 documents -> 2 CTAs each); every warp of the cluster owns a contiguous
//     range of 32-term bitmap words and walks the NON-ZERO ones in term order, four words per trip (their loads are
//     issued together):
//       pass 1 : per-(warp, position) counts in the CTA's shared memory
//       scan   : every CTA reads the counters of the whole cluster (its own directly, the others through
//                distributed shared memory) -> row_ptr[s] (exclusive over positions) and the first slot of each
//                of its (warp, position) pairs
//       pass 2 : the same walk (operands L2-hot); slot = base + running count + rank among the lanes with the same
//                position (lanes are in term order)
//     => every bucket lists its terms in ascending term id, so dHidden sums in a fixed order.
//     Shared memory: cnt [nw][S] | base [nw][S] | tot [S] int32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_nctaid() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ int32_t ld_dsmem_s32(uint32_t cluster_addr) {
  int32_t v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(cluster_addr));
  return v;
}

// One walk over the warp's words.  kScatter = false: count into cnt[]; true: cnt[] holds the running slot of every
// position, entries are written to the sorted lists.
template <bool kScatter>
__device__ __forceinline__ void bucket_walk(const float* __restrict__ g_rep, const float* __restrict__ rep,
                                            const int32_t* __restrict__ argmax, const uint32_t* __restrict__ bm,
                                            size_t base, int V, int w0, int w1, int lane, int32_t* cnt,
                                            int32_t* __restrict__ sorted_v, float* __restrict__ sorted_c) {
  for (int wc = w0; wc < w1; wc += 32) {
    const uint32_t myw = (wc + lane < w1) ? bm[wc + lane] : 0u;
    uint32_t nz = __ballot_sync(0xffffffffu, myw != 0u);
    while (nz) {
      uint32_t bits[4];
      int widx[4];
      float cc[4];
      int pp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = nz ? __ffs(nz) - 1 : 0;
        bits[u] = nz ? __shfl_sync(0xffffffffu, myw, k) : 0u;
        widx[u] = wc + k;
        nz &= nz - 1;  // (0 stays 0)
        cc[u] = 0.f;
        pp[u] = 0;
        if ((bits[u] >> lane) & 1u) {
          const size_t o = base + static_cast<size_t>(widx[u]) * 32 + lane;
          cc[u] = coef_of(g_rep[o], rep[o]);
          pp[u] = argmax[o];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool act = cc[u] != 0.f;
        const uint32_t am = __ballot_sync(0xffffffffu, act);
        if (am == 0u) continue;  // warp-uniform
        uint32_t peers = 0;
        int slot = 0;
        if (act) {
          peers = __match_any_sync(am, pp[u]);
          if (kScatter) slot = cnt[pp[u]] + __popc(peers & ((1u << lane) - 1u));  // lanes are in term order
        }
        __syncwarp();
        if (act) {
          if (lane == __ffs(peers) - 1) cnt[pp[u]] += __popc(peers);  // one lane per distinct position: no race
          if (kScatter) {
            sorted_v[base + slot] = widx[u] * 32 + lane;
            sorted_c[base + slot] = cc[u];
          }
        }
        __syncwarp();
      }
    }
  }
}

__global__ void __launch_bounds__(kBucketThreads)
bwd_bucket_kernel(const __grid_constant__ Segs p, int V, int nw) {
  extern __shared__ int32_t bk_smem[];
  __shared__ int32_t warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = static_cast<int>(cluster_nctaid());
  const int rank = static_cast<int>(cluster_ctarank());
  const int doc = blockIdx.x / C;
  const int si = seg_of_doc(p, doc);
  const Seg& sg = p.s[si];
  const int b = doc - sg.doc0;
  const int S = sg.S;
  int32_t* cnt_all = bk_smem;                                     // [nw][S]
  int32_t* base_all = bk_smem + static_cast<size_t>(nw) * S;      // [nw][S]
  int32_t* tot_s = bk_smem + static_cast<size_t>(2) * nw * S;     // [S]
  const int words = (V + 31) >> 5;
  const int gwarps = C * nw;
  const int wpw = (words + gwarps - 1) / gwarps;
  const int gw = rank * nw + warp;
  const int w0 = min(words, gw * wpw), w1 = (warp < nw) ? min(words, w0 + wpw) : w0;
  const size_t base = static_cast<size_t>(b) * V;
  const uint32_t* __restrict__ bm = sg.bitmap + static_cast<size_t>(b) * words;
  int32_t* cnt = cnt_all + static_cast<size_t>(min(warp, nw - 1)) * S;

  if (warp < nw)
    for (int s = lane; s < S; s += 32) cnt[s] = 0;
  __syncwarp();
  bucket_walk<false>(sg.g, sg.rep, sg.arg, bm, base, V, w0, w1, lane, cnt, nullptr, nullptr);
  cluster_sync_all();  // every counter of the cluster is final (also a CTA-wide barrier)

  // ---- scan over positions (thread t owns a contiguous range of positions)
  const int per = (S + kBucketThreads - 1) / kBucketThreads;
  const int lo = min(S, tid * per), hi = min(S, lo + per);
  const uint32_t cnt_addr = smem_u32(cnt_all);
  int tsum = 0;
  for (int s = lo; s < hi; ++s) {
    int tot = 0, before = 0;
    for (int r = 0; r < C; ++r) {
      int part = 0;
      if (r == rank) {
        for (int w = 0; w < nw; ++w) part += cnt_all[static_cast<size_t>(w) * S + s];
      } else {
        const uint32_t remote = map_to_cta(cnt_addr, static_cast<uint32_t>(r));
        for (int w = 0; w < nw; ++w) part += ld_dsmem_s32(remote + (static_cast<uint32_t>(w) * S + s) * 4u);
      }
      tot += part;
      if (r < rank) before += part;
    }
    tot_s[s] = tot;
    base_all[s] = before;  // row 0 of base_all, finalised below
    tsum += tot;
  }
  cluster_sync_all();  // nobody reads a peer's counters after this point (a CTA may finish before its peers)
  int incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (tid < 32) {
    const int w = tid < kBucketWarps ? warp_tot[tid] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, wi, o);
      if (tid >= o) wi += n;
    }
    warp_tot[tid] = wi - w;  // exclusive
  }
  __syncthreads();
  int32_t* rp = sg.row_ptr + static_cast<size_t>(b) * (S + 1);
  int run = warp_tot[warp] + incl - tsum;
  for (int s = lo; s < hi; ++s) {
    if (rank == 0) rp[s] = run;
    int slot = run + base_all[s];  // behind the slots of the lower-ranked CTAs
    for (int w = 0; w < nw; ++w) {  // first slot of position s for each of this CTA's warps
      const size_t i = static_cast<size_t>(w) * S + s;
      const int c = cnt_all[i];
      base_all[i] = slot;
      slot += c;
    }
    run += tot_s[s];
  }
  if (rank == 0 && tid == kBucketThreads - 1) rp[S] = run;  // total number of active terms of this document
  __syncthreads();

  int32_t* cur = base_all + static_cast<size_t>(min(warp, nw - 1)) * S;
  bucket_walk<true>(sg.g, sg.rep, sg.arg, bm, base, V, w0, w1, lane, cur, sg.sorted_v, sg.sorted_c);
}

// ---------------------------------------------------------------------------------------------
// Row fragments.  A lane owns 8 consecutive elements of every 256-element chunk of a row (h = c*256 + lane*8): one
// 128-bit load of bf16 per chunk (64-bit loads measured 1.5x slower in the L2-gather-bound dense regime), two
// 128-bit fp32 stores.  kC = chunks = ceil(H / 256).  (L2 eviction-priority hints on the gathers / output stores
// and st.global.cs were measured: no difference in either regime, so plain accesses.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fma_bf16x8(float (&acc)[8], float c, const uint4& u) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[2 * i] = fmaf(c, __uint_as_float(w[i] << 16), acc[2 * i]);
    acc[2 * i + 1] = fmaf(c, __uint_as_float(w[i] & 0xFFFF0000u), acc[2 * i + 1]);
  }
}
template <int kC>
__device__ __forceinline__ void load_row(uint4 (&u)[kC], const uint16_t* row, int H, int lane) {
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    const int h = c * 256 + lane * 8;
    if (h < H) u[c] = __ldg(reinterpret_cast<const uint4*>(row + h));
  }
}
template <int kC>
__device__ __forceinline__ void fma_row(float (&acc)[kC][8], float coef, const uint4 (&u)[kC], int H, int lane) {
#pragma unroll
  for (int c = 0; c < kC; ++c)
    if (c * 256 + lane * 8 < H) fma_bf16x8(acc[c], coef, u[c]);
}
template <int kC>
__device__ __forceinline__ void zero_acc(float (&acc)[kC][8]) {
#pragma unroll
  for (int c = 0; c < kC; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
}
template <int kC>
__device__ __forceinline__ void store_row(float* out, const float (&acc)[kC][8], int H, int lane) {
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    const int h = c * 256 + lane * 8;
    if (h < H) {
      *reinterpret_cast<float4*>(out + h) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      *reinterpret_cast<float4*>(out + h + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
    }
  }
}
template <int kC>
__device__ __forceinline__ void store_zero_row(float* out, int H, int lane) {
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    const int h = c * 256 + lane * 8;
    if (h < H) {
      *reinterpret_cast<float4*>(out + h) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(out + h + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// A document is summed by the dense-regime kernel when its buckets average at least this many terms per position
// (random-init models: ~195; trained models: ~1).  Decided per document from row_ptr[S], on the device.
constexpr int kDenseTermsPerPosition = 8;

// ---------------------------------------------------------------------------------------------
// (2a) dHidden, trained (sparse) regime: one warp per GROUP of kRowsPerWarp consecutive positions of one document.
//      row_ptr is monotone over the positions of a document, so the buckets of the group are ONE contiguous range
//      of the sorted lists: the warp reads the group's row_ptr entries (one load), stores the EMPTY positions' zero
//      rows at once (nearly all of them: the write stream does not wait for the gather chain), then reads the
//      range (coalesced) and walks it as a segmented sum, kDepth W rows in flight.  Three dependent global
//      latencies per GROUP instead of per position.
// ---------------------------------------------------------------------------------------------
template <int kC, int kDepth>
__global__ void __launch_bounds__(256)
dhidden_group_kernel(const __grid_constant__ Segs p, const uint16_t* __restrict__ W, int H, int V) {
  const int lane = threadIdx.x & 31;
  const long long grp = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (grp >= p.total_rows) return;
  int si = 0;
#pragma unroll
  for (int i = 1; i < kMaxSegs; ++i)
    if (i < p.n && grp >= p.s[i].row0) si = i;
  const Seg& sg = p.s[si];
  if (sg.dh == nullptr) return;
  const int S = sg.S;
  const int gps = (S + kRowsPerWarp - 1) / kRowsPerWarp;
  const long long lg = grp - sg.row0;
  const int b = static_cast<int>(lg / gps);
  const int s0 = static_cast<int>(lg - static_cast<long long>(b) * gps) * kRowsPerWarp;
  const int nrows = min(kRowsPerWarp, S - s0);
  const int32_t* rp_doc = sg.row_ptr + static_cast<size_t>(b) * (S + 1);
  const int32_t* rp = rp_doc + s0;
  int my_rp = lane <= nrows ? rp[lane] : 0;
  if (lane == 31) my_rp = rp_doc[S];
  if (__shfl_sync(0xffffffffu, my_rp, 31) >= kDenseTermsPerPosition * S) return;  // dense document: kernel (2b)
  const int beg = __shfl_sync(0xffffffffu, my_rp, 0), end = __shfl_sync(0xffffffffu, my_rp, nrows);
  float* out = sg.dh + (static_cast<size_t>(b) * S + s0) * H;

  // empty positions first
  const int my_next = __shfl_down_sync(0xffffffffu, my_rp, 1);
  const uint32_t empty = __ballot_sync(0xffffffffu, lane < nrows && my_next == my_rp);
  for (uint32_t m = empty; m; m &= m - 1) store_zero_row<kC>(out + static_cast<size_t>(__ffs(m) - 1) * H, H, lane);
  if (beg == end) return;

  const size_t lbase = static_cast<size_t>(b) * V;
  const int32_t* __restrict__ sorted_v = sg.sorted_v;
  const float* __restrict__ sorted_c = sg.sorted_c;
  float acc[kC][8];
  zero_acc<kC>(acc);
  int row = 0;                                        // position s0 + row is being summed
  while ((empty >> row) & 1u) ++row;                  // (beg < end: some position is not empty)
  int row_end = __shfl_sync(0xffffffffu, my_rp, row + 1);

  for (int e0 = beg; e0 < end; e0 += 32) {
    const int n = min(32, end - e0);
    int my_v = 0;
    float my_c = 0.f;
    if (lane < n) {
      my_v = __ldcs(sorted_v + lbase + e0 + lane);
      my_c = __ldcs(sorted_c + lbase + e0 + lane);
    }
    for (int l = 0; l < n; l += kDepth) {
      uint4 u[kDepth][kC];
      float c[kDepth];
#pragma unroll
      for (int k = 0; k < kDepth; ++k) {
        const int src = min(l + k, n - 1);  // past the end: a valid row, never accumulated
        const int v = __shfl_sync(0xffffffffu, my_v, src);
        c[k] = __shfl_sync(0xffffffffu, my_c, src);
        load_row<kC>(u[k], W + static_cast<size_t>(v) * H, H, lane);
      }
#pragma unroll
      for (int k = 0; k < kDepth; ++k) {
        if (l + k < n) {
          while (e0 + l + k >= row_end) {  // warp-uniform: this position's bucket is complete
            store_row<kC>(out + static_cast<size_t>(row) * H, acc, H, lane);
            zero_acc<kC>(acc);
            ++row;
            while ((empty >> row) & 1u) ++row;
            row_end = __shfl_sync(0xffffffffu, my_rp, row + 1);
          }
          fma_row<kC>(acc, c[k], u[k], H, lane);
        }
      }
    }
  }
  store_row<kC>(out + static_cast<size_t>(row) * H, acc, H, lane);  // the last non-empty position
}

// ---------------------------------------------------------------------------------------------
// (2b) dHidden, dense regime (random init: every term of every document is active, ~V/S terms per position): one
//      warp per (document, position) row, 64 registers so that four CTAs of 8 warps stay resident per SM - this
//      regime is bound by the L2 -> SM gather of W rows (B*V rows of 2H bytes per head call), i.e. by how many loads
//      are in flight per SM.
// ---------------------------------------------------------------------------------------------
template <int kC>
__global__ void __launch_bounds__(256, 4)
dhidden_row_kernel(const __grid_constant__ Segs p, const uint16_t* __restrict__ W, int H, int V) {
  const int lane = threadIdx.x & 31;
  const long long grp = static_cast<long long>(blockIdx.x);  // one CTA per group of 8 positions, one warp per row
  if (grp >= p.total_rows) return;
  int si = 0;
#pragma unroll
  for (int i = 1; i < kMaxSegs; ++i)
    if (i < p.n && grp >= p.s[i].row0) si = i;
  const Seg& sg = p.s[si];
  if (sg.dh == nullptr) return;
  const int S = sg.S;
  const int gps = (S + kRowsPerWarp - 1) / kRowsPerWarp;
  const long long lg = grp - sg.row0;
  const int b = static_cast<int>(lg / gps);
  const int s = static_cast<int>(lg - static_cast<long long>(b) * gps) * kRowsPerWarp + (threadIdx.x >> 5);
  const int32_t* rp = sg.row_ptr + static_cast<size_t>(b) * (S + 1);
  if (rp[S] < kDenseTermsPerPosition * S) return;  // sparse document: kernel (2a)
  if (s >= S) return;
  const int beg = rp[s], end = rp[s + 1];
  const size_t lbase = static_cast<size_t>(b) * V;
  const int32_t* __restrict__ sorted_v = sg.sorted_v;
  const float* __restrict__ sorted_c = sg.sorted_c;
  float acc[kC][8];
  zero_acc<kC>(acc);
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int n = min(32, end - e0);
    int my_v = 0;
    float my_c = 0.f;
    if (lane < n) {
      my_v = __ldcs(sorted_v + lbase + e0 + lane);
      my_c = __ldcs(sorted_c + lbase + e0 + lane);
    }
#pragma unroll 4
    for (int l = 0; l < n; ++l) {
      const int v = __shfl_sync(0xffffffffu, my_v, l);
      const float c = __shfl_sync(0xffffffffu, my_c, l);
      uint4 u[kC];
      load_row<kC>(u, W + static_cast<size_t>(v) * H, H, lane);
      fma_row<kC>(acc, c, u, H, lane);
    }
  }
  store_row<kC>(sg.dh + (static_cast<size_t>(b) * S + s) * H, acc, H, lane);
}

// ---------------------------------------------------------------------------------------------
// (3) dW + dbias for ALL segments in one sweep.  A CTA owns the 32 terms of one bitmap word.  Per pass of
//     <= kDwDocs documents (one pass unless a micro-step has more head rows than that):
//       a. one bitmap word per document -> shared memory (one load per thread, all in flight together); terms that
//          no document activates get their zero row stored right away (the write stream does not wait for b-d);
//       b. the set bits are compacted into a list of (document, term) pairs (shared-memory slot counter; the order
//          of the list is irrelevant, every pair has its own staging slot);
//       c. g / rep / argmax of the listed pairs only - one global latency for the whole CTA in the trained regime,
//          coalesced 128-byte rows in the dense one - give coef and the hidden row index, staged as [term][doc];
//       d. warp w owns terms 4w..4w+3: per term it walks the active documents (ballot over the bitmap words, 32
//          documents at a time) in ascending document order, kDepth hidden rows in flight.
//     The sum order is fixed (segments in call order, documents ascending) => run-to-run bit-identical.
// ---------------------------------------------------------------------------------------------
template <int kC, int kDepth>
__global__ void __launch_bounds__(256)
dw_kernel(const __grid_constant__ Segs p, int H, int V, int docs_per_pass, int stride /* odd, >= docs_per_pass */,
          float* dW, float* dbias, int accumulate) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  float* s_c = reinterpret_cast<float*>(dw_smem);                                       // [32][stride]
  uint32_t* s_p = reinterpret_cast<uint32_t*>(s_c + static_cast<size_t>(32) * stride);  // [32][stride]
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_p + static_cast<size_t>(32) * stride);  // [32 * docs_per_pass]
  __shared__ const uint16_t* s_hid[kMaxSegs];
  __shared__ uint32_t s_bits[kDwDocs];
  __shared__ int s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int word = blockIdx.x, v0 = word * 32;
  const int words = (V + 31) >> 5;
  if (threadIdx.x < kMaxSegs) s_hid[threadIdx.x] = threadIdx.x < p.n ? p.s[threadIdx.x].hidden : nullptr;
  const bool single_pass = p.total_docs <= docs_per_pass;

  for (int d0 = 0; d0 < p.total_docs; d0 += docs_per_pass) {
    const int nb = min(docs_per_pass, p.total_docs - d0);
    __syncthreads();  // the previous pass is done with the staging buffers (and s_hid is visible)
    if (threadIdx.x == 0) s_n = 0;
    if (static_cast<int>(threadIdx.x) < nb) {  // a.
      const int doc = d0 + threadIdx.x;
      const Seg& sg = p.s[seg_of_doc(p, doc)];
      s_bits[threadIdx.x] = sg.bitmap[static_cast<size_t>(doc - sg.doc0) * words + word];
    }
    __syncthreads();
    // which of this warp's four terms does some document of the pass activate?
    uint32_t any = 0;
    for (int c0 = 0; c0 < nb; c0 += 32) any |= (c0 + lane < nb) ? s_bits[c0 + lane] : 0u;
#pragma unroll
    for (int o = 16; o; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
    if (single_pass && !accumulate) {
      for (int r = 0; r < 4; ++r) {
        const int j = warp * 4 + r;
        if (v0 + j < V && !((any >> j) & 1u)) {
          store_zero_row<kC>(dW + static_cast<size_t>(v0 + j) * H, H, lane);
          if (lane == 0) dbias[v0 + j] = 0.f;
        }
      }
    }
    {  // b.  thread t still knows document t's word: slots by a warp scan + one shared-memory atomic per warp
      const uint32_t bits = static_cast<int>(threadIdx.x) < nb ? s_bits[threadIdx.x] : 0u;
      const int n = __popc(bits);
      int incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      int wbase = 0;
      const int wtot = __shfl_sync(0xffffffffu, incl, 31);
      if (lane == 31 && wtot) wbase = atomicAdd(&s_n, wtot);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      int slot = wbase + incl - n;
      for (uint32_t m = bits; m; m &= m - 1)
        s_list[slot++] = static_cast<uint16_t>((threadIdx.x << 5) | (__ffs(m) - 1));
    }
    __syncthreads();
    const int n_pairs = s_n;
#pragma unroll 4
    for (int i = threadIdx.x; i < n_pairs; i += 256) {  // c.
      const int e = s_list[i];
      const int dl = e >> 5, j = e & 31;
      const int doc = d0 + dl;
      const int si = seg_of_doc(p, doc);
      const Seg& sg = p.s[si];
      const int b = doc - sg.doc0;
      const size_t o = static_cast<size_t>(b) * V + v0 + j;
      s_c[j * stride + dl] = coef_of(sg.g[o], sg.rep[o]);
      s_p[j * stride + dl] = (static_cast<uint32_t>(si) << 28) | static_cast<uint32_t>(b * sg.S + sg.arg[o]);
    }
    __syncthreads();

    const bool first = (d0 == 0) && !accumulate;
    for (int r = 0; r < 4; ++r) {  // d.
      const int j = warp * 4 + r;
      const int v = v0 + j;
      if (v >= V) break;  // warp-uniform
      const bool touched = (any >> j) & 1u;
      if (!touched && (single_pass ? !accumulate : !first)) continue;  // zero row already stored / nothing to add
      float* out = dW + static_cast<size_t>(v) * H;
      float acc[kC][8];
      float bsum = 0.f;
      if (first) {
        zero_acc<kC>(acc);
      } else {  // continue this term's sum (an earlier pass of this warp, or the caller's gradient buffer)
        if (!touched) continue;
#pragma unroll
        for (int c = 0; c < kC; ++c) {
          const int h = c * 256 + lane * 8;
          if (h < H) {
            const float4 a = *reinterpret_cast<const float4*>(out + h);
            const float4 b4 = *reinterpret_cast<const float4*>(out + h + 4);
            acc[c][0] = a.x, acc[c][1] = a.y, acc[c][2] = a.z, acc[c][3] = a.w;
            acc[c][4] = b4.x, acc[c][5] = b4.y, acc[c][6] = b4.z, acc[c][7] = b4.w;
          }
        }
        bsum = dbias[v];
      }
      const float* cj = s_c + j * stride;
      const uint32_t* pj = s_p + j * stride;
      for (int c0 = 0; c0 < nb; c0 += 32) {
        const uint32_t bits = (c0 + lane < nb) ? s_bits[c0 + lane] : 0u;
        uint32_t m = __ballot_sync(0xffffffffu, (bits >> j) & 1u);
        while (m) {  // up to kDepth active documents per trip: their hidden rows are independent loads
          int d[kDepth];
          float c[kDepth];
          uint4 u[kDepth][kC];
#pragma unroll
          for (int k = 0; k < kDepth; ++k) {
            d[k] = m ? (c0 + __ffs(m) - 1) : -1;
            m &= m - 1;  // (0 stays 0)
            const int dk = d[k] < 0 ? d[0] : d[k];
            c[k] = cj[dk];  // broadcast reads
            const uint32_t pw = pj[dk];
            load_row<kC>(u[k], s_hid[pw >> 28] + static_cast<size_t>(pw & 0x0FFFFFFFu) * H, H, lane);
          }
#pragma unroll
          for (int k = 0; k < kDepth; ++k) {
            if (d[k] >= 0) {
              bsum += c[k];
              fma_row<kC>(acc, c[k], u[k], H, lane);
            }
          }
        }
      }
      if (lane == 0) dbias[v] = bsum;
      store_row<kC>(out, acc, H, lane);  // (a later pass reads it back: same lane, same addresses)
    }
  }
}

size_t bitmap_bytes(int B, int V) { return align256(static_cast<size_t>(B) * ((V + 31) / 32) * sizeof(uint32_t)); }

// transient workspace of one segment's bucket + dHidden phase (row_ptr | sorted_v | sorted_c)
size_t seg_sort_bytes(int B, int S, int V) {
  return align256(static_cast<size_t>(B) * (S + 1) * sizeof(int32_t)) +
         2 * align256(static_cast<size_t>(B) * V * sizeof(int32_t));
}

// The dW sweep does not depend on the bucket sort / dHidden kernels (it needs the bitmaps, g / rep / argmax and the
// hidden rows only) and all three are chains of dependent latencies that leave most issue slots idle: they run
// CONCURRENTLY, the sweep on a side stream forked from and joined back into the caller's stream with events.  One
// non-blocking side stream + two events per device, created on first use (the only device-side state of the library).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  std::mutex enqueue;  // one fork / launch / join sequence at a time per device (the two events are shared)
};

int side_stream(SideStream** out) {
  static SideStream table[64];
  static std::mutex mu;
  int dev = 0;
  SPLADE_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return set_error(kErrInvalidArg, "splade_head_bwd: device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  SideStream& s = table[dev];
  if (!s.stream) {
    SPLADE_CUDA_OK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    SPLADE_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    SPLADE_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return kOk;
}

bool concurrent_sweep_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = std::getenv("SPLADE_B200_BWD_CONCURRENT");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// dW sweep over `p_dw` on the side stream while `p_sort` is bucketed / dHidden-summed on the caller's stream.
int launch_sort_dhidden(const Segs& p, const void* W, int H, int V, int s_max, bool any_dh, cudaStream_t stream);
int launch_dw(const Segs& p, int H, int V, float* dW, float* dbias, int accumulate, cudaStream_t stream);

int launch_concurrent(const Segs& p_sort, int s_max, bool any_dh, const Segs& p_dw, const void* W, int H, int V,
                      float* dW, float* dbias, int accumulate, cudaStream_t stream) {
  if (!concurrent_sweep_enabled()) {
    int rc = launch_sort_dhidden(p_sort, W, H, V, s_max, any_dh, stream);
    if (rc != kOk) return rc;
    return launch_dw(p_dw, H, V, dW, dbias, accumulate, stream);
  }
  SideStream* side = nullptr;
  int rc = side_stream(&side);
  if (rc != kOk) return rc;
  std::lock_guard<std::mutex> lock(side->enqueue);
  SPLADE_CUDA_OK(cudaEventRecord(side->fork, stream));
  SPLADE_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
  rc = launch_dw(p_dw, H, V, dW, dbias, accumulate, side->stream);
  if (rc == kOk) rc = launch_sort_dhidden(p_sort, W, H, V, s_max, any_dh, stream);
  // always join: the caller's stream must not run ahead of work already forked
  const cudaError_t e1 = cudaEventRecord(side->join, side->stream);
  const cudaError_t e2 = cudaStreamWaitEvent(stream, side->join, 0);
  if (rc != kOk) return rc;
  SPLADE_CUDA_OK(e1);
  SPLADE_CUDA_OK(e2);
  return kOk;
}

int validate_common(const char* who, int H, int V) {
  if (V < 1 || H < 1) return set_error(kErrInvalidArg, "%s: bad sizes H=%d V=%d", who, H, V);
  if (H % 8 != 0 || H > 1024)
    return set_error(kErrUnsupportedShape, "%s: hidden size %d must be a multiple of 8 and <= 1024", who, H);
  return kOk;
}

int validate_seg(const char* who, const splade_head_bwd_segment& s, int i, bool need_dh) {
  if (s.B < 0 || s.S < 1) return set_error(kErrInvalidArg, "%s: segment %d has bad sizes B=%d S=%d", who, i, s.B, s.S);
  if (s.B > 0 && (!s.g_rep || !s.rep || !s.argmax || !s.hidden || (need_dh && !s.d_hidden)))
    return set_error(kErrInvalidArg, "%s: segment %d has a null pointer argument", who, i);
  if (static_cast<size_t>(s.S) * sizeof(int32_t) * 3 > 220 * 1024)
    return set_error(kErrUnsupportedShape, "%s: sequence length %d too long for the bucket pass", who, s.S);
  if (static_cast<long long>(s.B) * s.S >= (1ll << 28))
    return set_error(kErrUnsupportedShape, "%s: segment %d has B*S >= 2^28 rows", who, i);
  return kOk;
}

// bucket + dHidden of the segments in `p` (their row_ptr / sorted_* / bitmap pointers are set)
int launch_sort_dhidden(const Segs& p, const void* W, int H, int V, int s_max, bool any_dh, cudaStream_t stream) {
  if (p.total_docs == 0) return kOk;
  // shared memory: (2 nw + 1) S counters; as many warps as fit
  const size_t per_pos = static_cast<size_t>(s_max) * sizeof(int32_t);
  int nw = kBucketWarps;
  while (nw > 1 && (2 * static_cast<size_t>(nw) + 1) * per_pos > 200 * 1024) nw >>= 1;
  const size_t smem = (2 * static_cast<size_t>(nw) + 1) * per_pos;
  if (smem > 220 * 1024)
    return set_error(kErrUnsupportedShape, "splade_head_bwd: sequence length %d too long for the bucket pass", s_max);
  if (smem > 44 * 1024)
    SPLADE_CUDA_OK(cudaFuncSetAttribute(bwd_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
  // one wave over the machine: C CTAs per document (a cluster), C a power of two <= 8
  int C = 1;
  while (C < 8 && static_cast<long long>(p.total_docs) * C < 2ll * sm_count()) C <<= 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(p.total_docs) * C);
  cfg.blockDim = dim3(kBucketThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(C);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPLADE_CUDA_OK(cudaLaunchKernelEx(&cfg, bwd_bucket_kernel, p, V, nw));
  SPLADE_CUDA_OK(cudaGetLastError());
  if (!any_dh) return kOk;
  const uint16_t* Wb = static_cast<const uint16_t*>(W);
  const unsigned grid_h = static_cast<unsigned>((p.total_rows + 7) / 8);  // 8 warps = 8 row groups per CTA
  // both regimes are enqueued; every warp / CTA looks at its document's row_ptr[S] and the kernel whose regime it is
  // not exits at once (no host synchronisation, mixed batches are handled per document)
  const unsigned grid_rows = static_cast<unsigned>(p.total_rows);
  switch ((H + 255) / 256) {
    case 1:
      dhidden_group_kernel<1, 4><<<grid_h, 256, 0, stream>>>(p, Wb, H, V);
      dhidden_row_kernel<1><<<grid_rows, 256, 0, stream>>>(p, Wb, H, V);
      break;
    case 2:
      dhidden_group_kernel<2, 4><<<grid_h, 256, 0, stream>>>(p, Wb, H, V);
      dhidden_row_kernel<2><<<grid_rows, 256, 0, stream>>>(p, Wb, H, V);
      break;
    case 3:
      dhidden_group_kernel<3, 4><<<grid_h, 256, 0, stream>>>(p, Wb, H, V);
      dhidden_row_kernel<3><<<grid_rows, 256, 0, stream>>>(p, Wb, H, V);
      break;
    default:
      dhidden_group_kernel<4, 4><<<grid_h, 256, 0, stream>>>(p, Wb, H, V);
      dhidden_row_kernel<4><<<grid_rows, 256, 0, stream>>>(p, Wb, H, V);
      break;
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

template <int kC>
int launch_dw_t(const Segs& p, int H, int V, float* dW, float* dbias, int accumulate, cudaStream_t stream) {
  const int docs = std::max(1, std::min(kDwDocs, p.total_docs));
  const int stride = docs | 1;  // odd: the transposed staging writes are bank-conflict free
  const size_t smem = static_cast<size_t>(32) * stride * 8 + static_cast<size_t>(32) * docs * sizeof(uint16_t);
  if (smem > 44 * 1024)  // (static shared memory counts against the 48 KB default too)
    SPLADE_CUDA_OK(cudaFuncSetAttribute(dw_kernel<kC, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
  dw_kernel<kC, 4><<<static_cast<unsigned>((V + 31) / 32), 256, smem, stream>>>(p, H, V, docs, stride, dW, dbias,
                                                                                accumulate);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

int launch_dw(const Segs& p, int H, int V, float* dW, float* dbias, int accumulate, cudaStream_t stream) {
  switch ((H + 255) / 256) {
    case 1: return launch_dw_t<1>(p, H, V, dW, dbias, accumulate, stream);
    case 2: return launch_dw_t<2>(p, H, V, dW, dbias, accumulate, stream);
    case 3: return launch_dw_t<3>(p, H, V, dW, dbias, accumulate, stream);
    default: return launch_dw_t<4>(p, H, V, dW, dbias, accumulate, stream);
  }
}

int launch_rep_bitmap(const float* rep, int B, int V, uint32_t* bitmap, cudaStream_t stream) {
  if (B == 0) return kOk;
  const long long total = static_cast<long long>(B) * ((V + 31) / 32);
  const unsigned grid = static_cast<unsigned>(std::min<long long>((total + 7) / 8, 16ll * sm_count()));
  rep_bitmap_kernel<<<grid, 256, 0, stream>>>(rep, B, V, bitmap);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // namespace

size_t head_bwd_bitmap_bytes(int B, int V) { return bitmap_bytes(B, V); }

int rep_bitmap_launch(const float* rep, int B, int V, uint32_t* bitmap, cudaStream_t stream) {
  if (B < 0 || V < 1) return set_error(kErrInvalidArg, "splade_rep_bitmap: bad sizes B=%d V=%d", B, V);
  if (B > 0 && (!rep || !bitmap)) return set_error(kErrInvalidArg, "splade_rep_bitmap: null pointer argument");
  int rc = require_sm100("splade_rep_bitmap");
  if (rc != kOk) return rc;
  return launch_rep_bitmap(rep, B, V, bitmap, stream);
}

size_t head_bwd_dhidden_workspace_bytes(int B, int S, int V) { return seg_sort_bytes(B, S, V); }

size_t head_bwd_workspace_bytes(int B, int S, int V) { return seg_sort_bytes(B, S, V) + bitmap_bytes(B, V); }

size_t head_bwd_multi_workspace_bytes(const splade_head_bwd_segment* segs, int n_segs, int V) {
  size_t total = 0;
  for (int i = 0; segs && i < n_segs; ++i)
    total += seg_sort_bytes(segs[i].B, segs[i].S, V) + bitmap_bytes(segs[i].B, V);
  return total;
}

// Phase 1 for ONE segment: bucket + dHidden; leaves the activity bitmap for phase 2.
int head_bwd_dhidden_launch(const splade_head_bwd_segment* seg, const void* W, int H, int V, uint32_t* bitmap,
                            void* workspace, size_t ws_bytes, cudaStream_t stream) {
  const char* who = "splade_head_bwd_dhidden";
  if (!seg || !W || (!bitmap && !seg->active)) return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  int rc = validate_common(who, H, V);
  if (rc != kOk) return rc;
  rc = validate_seg(who, *seg, 0, /*need_dh=*/false);  // d_hidden == NULL: only the bitmap is produced
  if (rc != kOk) return rc;
  rc = require_sm100(who);
  if (rc != kOk) return rc;
  const size_t need = seg_sort_bytes(seg->B, seg->S, V);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need);
  if (seg->B == 0) return kOk;
  Segs p{};
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  Seg& s = p.s[0];
  s.g = seg->g_rep, s.rep = seg->rep, s.arg = seg->argmax, s.hidden = static_cast<const uint16_t*>(seg->hidden);
  s.dh = seg->d_hidden, s.B = seg->B, s.S = seg->S, s.doc0 = 0, s.row0 = 0;
  if (seg->active) {
    s.bitmap = const_cast<uint32_t*>(seg->active);
  } else {  // the caller has no bitmap from the forward: one pass over rep
    rc = launch_rep_bitmap(seg->rep, seg->B, V, bitmap, stream);
    if (rc != kOk) return rc;
    s.bitmap = bitmap;
  }
  s.row_ptr = reinterpret_cast<int32_t*>(ws);
  ws += align256(static_cast<size_t>(s.B) * (s.S + 1) * sizeof(int32_t));
  s.sorted_v = reinterpret_cast<int32_t*>(ws);
  ws += align256(static_cast<size_t>(s.B) * V * sizeof(int32_t));
  s.sorted_c = reinterpret_cast<float*>(ws);
  p.n = 1, p.total_docs = s.B, p.total_rows = static_cast<long long>(s.B) * ((s.S + kRowsPerWarp - 1) / kRowsPerWarp);
  return launch_sort_dhidden(p, W, H, V, s.S, s.dh != nullptr, stream);
}

// Phase 2: ONE dW / dbias sweep over every segment whose bitmap phase 1 left behind.
int head_bwd_dw_launch(const splade_head_bwd_segment* segs, const uint32_t* const* bitmaps, int n_segs, int H, int V,
                       float* dW, float* dbias, int accumulate, cudaStream_t stream) {
  const char* who = "splade_head_bwd_dw";
  if (n_segs < 0 || n_segs > kMaxSegs)
    return set_error(kErrInvalidArg, "%s: %d segments (at most %d)", who, n_segs, kMaxSegs);
  if ((n_segs > 0 && (!segs || !bitmaps)) || !dW || !dbias)
    return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  int rc = validate_common(who, H, V);
  if (rc != kOk) return rc;
  Segs p{};
  for (int i = 0; i < n_segs; ++i) {
    rc = validate_seg(who, segs[i], i, /*need_dh=*/false);
    if (rc != kOk) return rc;
    if (segs[i].B > 0 && !bitmaps[i]) return set_error(kErrInvalidArg, "%s: segment %d has no bitmap", who, i);
    Seg& s = p.s[i];
    s.g = segs[i].g_rep, s.rep = segs[i].rep, s.arg = segs[i].argmax;
    s.hidden = static_cast<const uint16_t*>(segs[i].hidden);
    s.B = segs[i].B, s.S = segs[i].S, s.bitmap = const_cast<uint32_t*>(bitmaps[i]);
    s.doc0 = p.total_docs, s.row0 = p.total_rows;
    p.total_docs += s.B;
    p.total_rows += static_cast<long long>(s.B) * ((s.S + kRowsPerWarp - 1) / kRowsPerWarp);
  }
  p.n = n_segs;
  rc = require_sm100(who);
  if (rc != kOk) return rc;
  if (p.total_docs == 0 && !accumulate) {
    SPLADE_CUDA_OK(cudaMemsetAsync(dW, 0, static_cast<size_t>(V) * H * sizeof(float), stream));
    SPLADE_CUDA_OK(cudaMemsetAsync(dbias, 0, static_cast<size_t>(V) * sizeof(float), stream));
    return kOk;
  }
  if (p.total_docs == 0) return kOk;
  return launch_dw(p, H, V, dW, dbias, accumulate, stream);
}

// Everything for n segments in three launches: one bucket pass over all documents, one dHidden launch over all
// rows, one dW sweep.
int head_bwd_multi_launch(const splade_head_bwd_segment* segs, int n_segs, const void* W, int H, int V, float* dW,
                          float* dbias, int accumulate, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  const char* who = "splade_head_bwd_multi";
  if (n_segs < 0 || n_segs > kMaxSegs)
    return set_error(kErrInvalidArg, "%s: %d segments (at most %d)", who, n_segs, kMaxSegs);
  if ((n_segs > 0 && !segs) || !W || !dW || !dbias) return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  int rc = validate_common(who, H, V);
  if (rc != kOk) return rc;
  for (int i = 0; i < n_segs; ++i) {
    rc = validate_seg(who, segs[i], i, /*need_dh=*/false);
    if (rc != kOk) return rc;
  }
  rc = require_sm100(who);
  if (rc != kOk) return rc;
  const size_t need = head_bwd_multi_workspace_bytes(segs, n_segs, V);
  if (need > 0 && (!workspace || ws_bytes < need))
    return set_error(kErrWorkspace, "%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need);

  Segs p{};
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int s_max = 1;
  bool any_dh = false;
  for (int i = 0; i < n_segs; ++i) {
    Seg& s = p.s[i];
    s.g = segs[i].g_rep, s.rep = segs[i].rep, s.arg = segs[i].argmax;
    s.hidden = static_cast<const uint16_t*>(segs[i].hidden);
    s.dh = segs[i].d_hidden, s.B = segs[i].B, s.S = segs[i].S;
    s.doc0 = p.total_docs, s.row0 = p.total_rows;
    s.row_ptr = reinterpret_cast<int32_t*>(ws);
    ws += align256(static_cast<size_t>(s.B) * (s.S + 1) * sizeof(int32_t));
    s.sorted_v = reinterpret_cast<int32_t*>(ws);
    ws += align256(static_cast<size_t>(s.B) * V * sizeof(int32_t));
    s.sorted_c = reinterpret_cast<float*>(ws);
    ws += align256(static_cast<size_t>(s.B) * V * sizeof(int32_t));
    uint32_t* own_bitmap = reinterpret_cast<uint32_t*>(ws);
    ws += bitmap_bytes(s.B, V);
    if (segs[i].active) {
      s.bitmap = const_cast<uint32_t*>(segs[i].active);
    } else {
      rc = launch_rep_bitmap(s.rep, s.B, V, own_bitmap, stream);
      if (rc != kOk) return rc;
      s.bitmap = own_bitmap;
    }
    p.total_docs += s.B;
    p.total_rows += static_cast<long long>(s.B) * ((s.S + kRowsPerWarp - 1) / kRowsPerWarp);
    if (s.B > 0) s_max = std::max(s_max, s.S);
    any_dh = any_dh || (s.B > 0 && s.dh != nullptr);
  }
  p.n = n_segs;
  if (p.total_docs == 0) {
    if (!accumulate) {
      SPLADE_CUDA_OK(cudaMemsetAsync(dW, 0, static_cast<size_t>(V) * H * sizeof(float), stream));
      SPLADE_CUDA_OK(cudaMemsetAsync(dbias, 0, static_cast<size_t>(V) * sizeof(float), stream));
    }
    return kOk;
  }
  return launch_concurrent(p, s_max, any_dh, p, W, H, V, dW, dbias, accumulate, stream);
}

// Autograd's last head node of a backward pass: bucket + dHidden of ITS segment (workspace / bitmap as for
// splade_head_bwd_dhidden) concurrently with the dW sweep over ALL segments of the pass.
int head_bwd_last_launch(const splade_head_bwd_segment* last, const splade_head_bwd_segment* segs,
                         const uint32_t* const* bitmaps, int n_segs, const void* W, int H, int V, uint32_t* bitmap,
                         float* dW, float* dbias, int accumulate, void* workspace, size_t ws_bytes,
                         cudaStream_t stream) {
  const char* who = "splade_head_bwd_last";
  if (!last || !W || (!bitmap && !last->active) || !dW || !dbias || n_segs < 1 || n_segs > kMaxSegs || !segs ||
      !bitmaps)
    return set_error(kErrInvalidArg, "%s: bad argument", who);
  int rc = validate_common(who, H, V);
  if (rc != kOk) return rc;
  rc = validate_seg(who, *last, 0, /*need_dh=*/false);
  if (rc != kOk) return rc;
  rc = require_sm100(who);
  if (rc != kOk) return rc;
  const size_t need = seg_sort_bytes(last->B, last->S, V);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need);
  Segs pd{};
  for (int i = 0; i < n_segs; ++i) {
    rc = validate_seg(who, segs[i], i, /*need_dh=*/false);
    if (rc != kOk) return rc;
    if (segs[i].B > 0 && !bitmaps[i]) return set_error(kErrInvalidArg, "%s: segment %d has no bitmap", who, i);
    Seg& s = pd.s[i];
    s.g = segs[i].g_rep, s.rep = segs[i].rep, s.arg = segs[i].argmax;
    s.hidden = static_cast<const uint16_t*>(segs[i].hidden);
    s.B = segs[i].B, s.S = segs[i].S, s.bitmap = const_cast<uint32_t*>(bitmaps[i]);
    s.doc0 = pd.total_docs, s.row0 = pd.total_rows;
    pd.total_docs += s.B;
    pd.total_rows += static_cast<long long>(s.B) * ((s.S + kRowsPerWarp - 1) / kRowsPerWarp);
  }
  pd.n = n_segs;
  Segs ps{};
  {
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    Seg& s = ps.s[0];
    s.g = last->g_rep, s.rep = last->rep, s.arg = last->argmax, s.hidden = static_cast<const uint16_t*>(last->hidden);
    s.dh = last->d_hidden, s.B = last->B, s.S = last->S, s.doc0 = 0, s.row0 = 0;
    if (last->active) {
      s.bitmap = const_cast<uint32_t*>(last->active);
    } else {
      rc = launch_rep_bitmap(last->rep, last->B, V, bitmap, stream);
      if (rc != kOk) return rc;
      s.bitmap = bitmap;
    }
    s.row_ptr = reinterpret_cast<int32_t*>(ws);
    ws += align256(static_cast<size_t>(s.B) * (s.S + 1) * sizeof(int32_t));
    s.sorted_v = reinterpret_cast<int32_t*>(ws);
    ws += align256(static_cast<size_t>(s.B) * V * sizeof(int32_t));
    s.sorted_c = reinterpret_cast<float*>(ws);
    ps.n = 1, ps.total_docs = s.B;
    ps.total_rows = static_cast<long long>(s.B) * ((s.S + kRowsPerWarp - 1) / kRowsPerWarp);
  }
  if (pd.total_docs == 0) {
    if (!accumulate) {
      SPLADE_CUDA_OK(cudaMemsetAsync(dW, 0, static_cast<size_t>(V) * H * sizeof(float), stream));
      SPLADE_CUDA_OK(cudaMemsetAsync(dbias, 0, static_cast<size_t>(V) * sizeof(float), stream));
    }
    return launch_sort_dhidden(ps, W, H, V, last->S, last->d_hidden != nullptr, stream);
  }
  return launch_concurrent(ps, last->S, last->d_hidden != nullptr, pd, W, H, V, dW, dbias, accumulate, stream);
}

int head_bwd_launch(const float* g_rep, const float* rep, const int32_t* argmax, const void* hidden,
                    const void* W, int B, int S, int H, int V, float* d_hidden, float* dW, float* dbias,
                    void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (!g_rep || !rep || !argmax || !hidden || !W || !d_hidden || !dW || !dbias)
    return set_error(kErrInvalidArg, "splade_head_bwd: null pointer argument");
  if (B < 0 || S < 1 || V < 1 || H < 1)
    return set_error(kErrInvalidArg, "splade_head_bwd: bad sizes B=%d S=%d H=%d V=%d", B, S, H, V);
  splade_head_bwd_segment seg{};
  seg.g_rep = g_rep, seg.rep = rep, seg.argmax = argmax, seg.hidden = hidden, seg.d_hidden = d_hidden;
  seg.B = B, seg.S = S;
  return head_bwd_multi_launch(&seg, 1, W, H, V, dW, dbias, 0, workspace, ws_bytes, stream);
}

}  // namespace splade
