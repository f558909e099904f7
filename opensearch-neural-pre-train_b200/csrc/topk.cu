// Batched top-k pruning of sparse representations with the reference's production emit semantics
// (benchmark/encoders.py:320-343): keep weight > 0, drop block-listed ids, stable sort by weight descending
// (=> ties by ascending term id, because the reference builds its dict in ascending id order), first k.
// The reference does this per row on the host (.cpu(), python loops, sorted()); here one CTA handles one row:
//   1. 3-pass MSB radix select (11/11/10 bits) on the fp32 bit patterns (positive floats order like uints)
//      finds the k-th largest weight T;
//   2. a coalesced, id-ordered compaction takes every weight > T plus the lowest-id ties == T;
//   3. a bitonic sort of the <= k survivors on (weight desc, id asc).
// HBM/L2-bound: the row (V * 4 B) is streamed 5 times out of L2.
#include "common.cuh"
#include "kernels.h"

namespace splade {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 1024;

__device__ __forceinline__ uint32_t cand_key(const float* __restrict__ row, const uint8_t* __restrict__ block,
                                             int v) {
  const float x = row[v];
  if (!(x > 0.f)) return 0u;
  if (block != nullptr && block[v] != 0) return 0u;
  return __float_as_uint(x);  // > 0 for every candidate
}

__global__ void __launch_bounds__(kThreads)
topk_kernel(const float* __restrict__ rep, int V, const uint8_t* __restrict__ block, int k, int kpad,
            int32_t* __restrict__ ids, float* __restrict__ vals, int32_t* __restrict__ counts) {
  extern __shared__ unsigned long long s_keys[];  // [kpad] (weight bits << 32 | ~id), sorted descending
  __shared__ uint32_t hist[2048];
  __shared__ uint32_t s_prefix, s_need, s_total;
  __shared__ uint32_t w_gt[kWarps], w_eq[kWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = rep + static_cast<size_t>(blockIdx.x) * V;

  // ---- 1. radix select the k-th largest candidate key
  uint32_t prefix = 0, prefix_mask = 0;
  uint32_t need = static_cast<uint32_t>(k);  // rank (1-based, from the top) still to locate inside `prefix`
  bool take_all = false;
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = tid; i < 2048; i += kThreads) hist[i] = 0;
    __syncthreads();
    const uint32_t dmask = (1u << bits[pass]) - 1u;
    for (int v = tid; v < V; v += kThreads) {
      const uint32_t key = cand_key(row, block, v);
      if (key != 0u && (key & prefix_mask) == prefix) atomicAdd(&hist[(key >> shifts[pass]) & dmask], 1u);
    }
    __syncthreads();
    {
      // Parallel suffix scan over the 2048 bins: thread t owns bins [8t, 8t+8).  The wanted digit d* is the
      // largest d with (number of keys whose digit >= d) >= need.
      uint32_t local[8], tsum = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        local[e] = hist[tid * 8 + e];
        tsum += local[e];
      }
      uint32_t incl = tsum;  // inclusive suffix sum across lanes (towards higher lanes)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += n;
      }
      if (lane == 0) w_gt[warp] = incl;  // warp totals (w_gt reused as scratch before the compaction)
      __syncthreads();
      uint32_t above = 0, total = 0;
      for (int w = 0; w < kWarps; ++w) {
        if (w > warp) above += w_gt[w];
        total += w_gt[w];
      }
      uint32_t running = above + incl - tsum;  // keys with digit above this thread's bins
      if (tid == 0) s_total = total;
#pragma unroll
      for (int e = 7; e >= 0; --e) {
        if (running < need && running + local[e] >= need) {
          s_prefix = static_cast<uint32_t>(tid * 8 + e);
          s_need = need - running;
        }
        running += local[e];
      }
    }
    __syncthreads();
    if (pass == 0 && s_total <= static_cast<uint32_t>(k)) {
      take_all = true;  // fewer candidates than k: keep them all (block-uniform)
      break;
    }
    prefix |= s_prefix << shifts[pass];
    prefix_mask |= dmask << shifts[pass];
    need = s_need;
    __syncthreads();
  }
  // Now: take_all, or `prefix` = T (the k-th largest key) and `need` = how many keys == T to keep.
  const uint32_t T = take_all ? 0u : prefix;

  // ---- 2. id-ordered compaction: warp w owns the contiguous id range [w*seg, (w+1)*seg)
  const int seg = ((V + kWarps - 1) / kWarps + 31) / 32 * 32;
  const int vbeg = warp * seg, vend = min(V, vbeg + seg);
  uint32_t n_gt = 0, n_eq = 0;
  for (int v0 = vbeg; v0 < vend; v0 += 32) {
    const int v = v0 + lane;
    const uint32_t key = (v < vend) ? cand_key(row, block, v) : 0u;
    n_gt += __popc(__ballot_sync(0xffffffffu, key > T));
    n_eq += __popc(__ballot_sync(0xffffffffu, !take_all && key == T && key != 0u));
  }
  if (lane == 0) {
    w_gt[warp] = n_gt;
    w_eq[warp] = n_eq;
  }
  __syncthreads();
  uint32_t base_gt = 0, base_eq = 0, tot_gt = 0;
  for (int w = 0; w < kWarps; ++w) {
    if (w < warp) {
      base_gt += w_gt[w];
      base_eq += w_eq[w];
    }
    tot_gt += w_gt[w];
  }
  const uint32_t keep_eq = take_all ? 0u : need;
  const uint32_t n_out = min(static_cast<uint32_t>(k), tot_gt + keep_eq);

  for (int i = tid; i < kpad; i += kThreads) s_keys[i] = 0ull;
  __syncthreads();
  for (int v0 = vbeg; v0 < vend; v0 += 32) {
    const int v = v0 + lane;
    const uint32_t key = (v < vend) ? cand_key(row, block, v) : 0u;
    const bool gt = key > T;
    const bool eq = !take_all && key == T && key != 0u;
    const uint32_t m_gt = __ballot_sync(0xffffffffu, gt);
    const uint32_t m_eq = __ballot_sync(0xffffffffu, eq);
    const uint32_t lt = (1u << lane) - 1u;
    const unsigned long long packed =
        (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(v));
    if (gt) {
      const uint32_t slot = base_gt + __popc(m_gt & lt);
      if (slot < static_cast<uint32_t>(kpad)) s_keys[slot] = packed;
    }
    if (eq) {
      const uint32_t rank = base_eq + __popc(m_eq & lt);
      if (rank < keep_eq) s_keys[tot_gt + rank] = packed;
    }
    base_gt += __popc(m_gt);
    base_eq += __popc(m_eq);
  }
  __syncthreads();

  // ---- 3. bitonic sort, descending, on kpad (power of two) 64-bit keys
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kpad / 2; i += kThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = s_keys[lo], b = s_keys[hi];
        if ((a < b) == desc) {
          s_keys[lo] = b;
          s_keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }

  int32_t* oid = ids + static_cast<size_t>(blockIdx.x) * k;
  float* oval = vals + static_cast<size_t>(blockIdx.x) * k;
  for (int i = tid; i < k; i += kThreads) {
    if (static_cast<uint32_t>(i) < n_out) {
      const unsigned long long p = s_keys[i];
      oid[i] = static_cast<int32_t>(~static_cast<uint32_t>(p & 0xffffffffull));
      oval[i] = __uint_as_float(static_cast<uint32_t>(p >> 32));
    } else {
      oid[i] = -1;
      oval[i] = 0.f;
    }
  }
  if (tid == 0) counts[blockIdx.x] = static_cast<int32_t>(n_out);
}

}  // namespace

int topk_launch(const float* rep, int B, int V, const uint8_t* blocklist, int k, int32_t* ids, float* vals,
                int32_t* counts, cudaStream_t stream) {
  if (B == 0) return kOk;
  if (!rep || !ids || !vals || !counts) return set_error(kErrInvalidArg, "splade_topk: null pointer argument");
  if (B < 0 || V < 1) return set_error(kErrInvalidArg, "splade_topk: bad sizes B=%d V=%d", B, V);
  if (k < 1 || k > kMaxK) return set_error(kErrUnsupportedShape, "splade_topk: k=%d outside [1,%d]", k, kMaxK);
  int kpad = 2;
  while (kpad < k) kpad <<= 1;
  topk_kernel<<<B, kThreads, static_cast<size_t>(kpad) * sizeof(unsigned long long), stream>>>(
      rep, V, blocklist, k, kpad, ids, vals, counts);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // namespace splade
