// Batched top-k pruning of sparse representations with the reference's production emit semantics
// (benchmark/encoders.py:320-343): keep weight > 0, drop block-listed ids, stable sort by weight descending
// (=> ties by ascending term id, because the reference builds its dict in ascending id order), first k.
// The reference does this per row on the host (.cpu(), python loops, sorted()); here one CTA handles one row:
//   0. the row is read from HBM ONCE: filtered fp32 bit patterns ("keys", 0 = not a candidate) are staged in
//      shared memory (200 KB for V = 50 000, one CTA of 1024 threads per SM); rows that do not fit fall back to
//      re-reading the keys from global memory / L2 in every pass;
//   1. 3-pass MSB radix select (11/11/10 bits) on the keys (positive floats order like uints) finds the k-th
//      largest weight T; histogram updates are warp-aggregated (match.any) because pooled SPLADE weights take
//      only ~200 distinct values per row;
//   2. an id-ordered warp-ballot compaction takes every weight > T plus the lowest-id ties == T;
//   3. a bitonic sort of the <= k survivors on (weight desc, id asc).
// HBM-bound: algorithmic bytes = B * V * 4.
#include "common.cuh"
#include "kernels.h"

namespace splade {
namespace {

constexpr int kMaxK = 1024;
constexpr int kSmemBudget = 227 * 1024;
constexpr int32_t kNeedsFallback = -1;  // marker in counts[]: the fast kernel hands this row to the exact kernel
constexpr int kListCap = 4096;           // fast path: candidates >= lower bound
constexpr int kSortCap = 2048;           // fast path: survivors (> T plus all ties == T)

__device__ __forceinline__ uint32_t cand_key(const float* __restrict__ row, const uint8_t* __restrict__ block,
                                             int v) {
  const float x = row[v];
  if (!(x > 0.f)) return 0u;
  if (block != nullptr && block[v] != 0) return 0u;
  return __float_as_uint(x);  // > 0 for every candidate
}

// 3-pass MSB radix select (11/11/10 bits) over n keys delivered by key_fn (0 = not a candidate; keys below
// `floor_key` are ignored).  Finds the key T of rank `k_rank` from the top.  Returns true when there are at most
// k_rank candidates (then everything is kept); otherwise *T_out = T and *need_out = how many keys == T belong
// to the top k_rank.  Block-uniform result; all threads of the CTA must call it.
template <int kThreads, class KeyFn>
__device__ __forceinline__ bool radix_select(KeyFn key_fn, int n, uint32_t k_rank, uint32_t floor_key,
                                             uint32_t* hist, uint32_t* w_scratch, uint32_t* s_misc,
                                             uint32_t* T_out, uint32_t* need_out) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kBinsPerThread = 2048 / kThreads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t prefix = 0, prefix_mask = 0;
  uint32_t need = k_rank;  // rank (1-based, from the top) still to locate inside `prefix`
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = tid; i < 2048; i += kThreads) hist[i] = 0;
    __syncthreads();
    const uint32_t dmask = (1u << bits[pass]) - 1u;
    for (int v0 = warp * 32; v0 < n; v0 += kThreads) {
      const int v = v0 + lane;
      const uint32_t key = (v < n) ? key_fn(v) : 0u;
      const bool active = key != 0u && key >= floor_key && (key & prefix_mask) == prefix;
      const uint32_t am = __ballot_sync(0xffffffffu, active);
      if (active) {  // warp-aggregated histogram update: one shared-memory atomic per distinct digit
        const uint32_t digit = (key >> shifts[pass]) & dmask;
        const uint32_t peers = __match_any_sync(am, digit);
        if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], static_cast<uint32_t>(__popc(peers)));
      }
    }
    __syncthreads();
    {
      // Parallel suffix scan over the 2048 bins: thread t owns kBinsPerThread consecutive bins.  The wanted
      // digit d* is the largest d with (number of keys whose digit >= d) >= need.
      uint32_t local[kBinsPerThread], tsum = 0;
#pragma unroll
      for (int e = 0; e < kBinsPerThread; ++e) {
        local[e] = hist[tid * kBinsPerThread + e];
        tsum += local[e];
      }
      uint32_t incl = tsum;  // inclusive suffix sum across lanes (towards higher lanes)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t nn = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += nn;
      }
      if (lane == 0) w_scratch[warp] = incl;
      __syncthreads();
      uint32_t above = 0, total = 0;
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t t = w_scratch[w];
        if (w > warp) above += t;
        total += t;
      }
      uint32_t running = above + incl - tsum;  // keys with digit above this thread's bins
      if (tid == 0) s_misc[2] = total;
#pragma unroll
      for (int e = kBinsPerThread - 1; e >= 0; --e) {
        if (running < need && running + local[e] >= need) {
          s_misc[0] = static_cast<uint32_t>(tid * kBinsPerThread + e);
          s_misc[1] = need - running;
        }
        running += local[e];
      }
    }
    __syncthreads();
    const uint32_t total = s_misc[2], digit = s_misc[0], new_need = s_misc[1];
    __syncthreads();  // s_misc / w_scratch may be rewritten by the next pass
    if (pass == 0 && total <= k_rank) return true;  // fewer candidates than k_rank: keep them all
    prefix |= digit << shifts[pass];
    prefix_mask |= dmask << shifts[pass];
    need = new_need;
  }
  *T_out = prefix;
  *need_out = need;
  return false;
}

// kStaged: keys live in shared memory; otherwise they are recomputed from global memory on every access.
template <int kThreads, bool kStaged>
__global__ void __launch_bounds__(kThreads)
topk_kernel(const float* __restrict__ rep, int V, const uint8_t* __restrict__ block, int k, int kpad,
            int32_t* __restrict__ ids, float* __restrict__ vals, int32_t* __restrict__ counts, int only_flagged) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_dyn);     // [kpad] sorted descending
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_dyn + static_cast<size_t>(kpad) * 8);  // [2048]
  uint32_t* w_gt = hist + 2048;        // [kWarps]
  uint32_t* w_eq = w_gt + kWarps;      // [kWarps]
  uint32_t* s_misc = w_eq + kWarps;    // [4]: digit, need, total
  uint32_t* tmax = s_misc + 4;         // [kThreads] per-thread maxima
  uint32_t* staged = tmax + kThreads;  // [V] when kStaged

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = rep + static_cast<size_t>(blockIdx.x) * V;
  if (only_flagged && counts[blockIdx.x] != kNeedsFallback) return;  // block-uniform

  // ---- 0. read the row once; remember each thread's largest key.  16-byte loads, four in flight per thread:
  //          a single CTA has to pull its 200 KB row out of HBM with enough memory-level parallelism.
  uint32_t my_max = 0;
  const bool vec_ok = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
  if (vec_ok) {
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int n4 = V >> 2;
    for (int i0 = tid; i0 < n4; i0 += 4 * kThreads) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kThreads;
        x[u] = (i < n4) ? __ldg(&row4[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kThreads;
        if (i < n4) {
          const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
          uint32_t ks[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int v = 4 * i + e;
            uint32_t key = (xs[e] > 0.f) ? __float_as_uint(xs[e]) : 0u;
            if (key != 0u && block != nullptr && block[v] != 0) key = 0u;
            ks[e] = key;
            my_max = max(my_max, key);
          }
          if (kStaged) *reinterpret_cast<uint4*>(&staged[4 * i]) = make_uint4(ks[0], ks[1], ks[2], ks[3]);
        }
      }
    }
  } else {
    for (int v = tid; v < V; v += kThreads) {
      const uint32_t key = cand_key(row, block, v);
      if (kStaged) staged[v] = key;
      my_max = max(my_max, key);
    }
  }
  tmax[tid] = my_max;
  __syncthreads();
  auto key_at = [&](int v) -> uint32_t { return kStaged ? staged[v] : cand_key(row, block, v); };

  // ---- 1a. lower bound: the k largest per-thread maxima are k distinct elements of the row, so the k-th
  //          largest key of the row is >= their k-th largest.  Everything below that bound (all but ~2k
  //          elements of a dense row) is skipped by the histogram passes, whose cost is the number of
  //          shared-memory atomics.
  uint32_t floor_key = 0, T = 0, need = 0;
  if (k <= kThreads) {
    uint32_t dummy = 0;
    const bool few = radix_select<kThreads>([&](int i) -> uint32_t { return tmax[i]; }, kThreads,
                                            static_cast<uint32_t>(k), 0u, hist, w_gt, s_misc, &floor_key, &dummy);
    if (few) floor_key = 0;
  }
  // ---- 1b. exact k-th largest key of the row
  const bool take_all = radix_select<kThreads>(key_at, V, static_cast<uint32_t>(k), floor_key, hist, w_gt, s_misc,
                                               &T, &need);
  // "keep every candidate": with a lower bound in force the candidates are the keys >= floor_key
  if (take_all) T = floor_key ? floor_key - 1u : 0u;
  // Now: take_all, or T = the k-th largest key and `need` = how many keys == T to keep.

  // ---- 2. id-ordered compaction: warp w owns the contiguous id range [w*seg, (w+1)*seg)
  const int seg = ((V + kWarps - 1) / kWarps + 31) / 32 * 32;
  const int vbeg = warp * seg, vend = min(V, vbeg + seg);
  uint32_t n_gt = 0, n_eq = 0;
  for (int v0 = vbeg; v0 < vend; v0 += 32) {
    const int v = v0 + lane;
    const uint32_t key = (v < vend) ? key_at(v) : 0u;
    n_gt += __popc(__ballot_sync(0xffffffffu, key > T));
    n_eq += __popc(__ballot_sync(0xffffffffu, !take_all && key == T && key != 0u));
  }
  __syncthreads();  // every thread is done reading w_gt as scan scratch
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
    const uint32_t key = (v < vend) ? key_at(v) : 0u;
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


// ---------------------------------------------------------------------------------------------
// Fast path (the common case).  Two coalesced sweeps over the row instead of keeping it on chip:
//   A. per-thread maxima -> lower bound L on the k-th largest key (see topk_kernel, step 1a);
//   B. second sweep (L2-hot): every candidate >= L is appended to a small shared-memory list (warp-aggregated
//      slot allocation).  A dense random-init row leaves ~2k of 50 000 elements, a trained one ~60;
//   C. radix select on the LIST finds the exact k-th largest key T; everything > T and all ties == T move to
//      the sort buffer;
//   D. bitonic sort on (weight desc, id asc); the first k entries are the result.
// Rows whose candidate list or tie set overflows the buffers (pathological ties, e.g. a constant row) are
// flagged in counts[] and redone by the exact multi-pass kernel.
// ---------------------------------------------------------------------------------------------
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
topk_fast_kernel(const float* __restrict__ rep, int V, const uint8_t* __restrict__ block, int k,
                 int32_t* __restrict__ ids, float* __restrict__ vals, int32_t* __restrict__ counts) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  unsigned long long* list = reinterpret_cast<unsigned long long*>(smem_dyn);  // [kListCap]
  unsigned long long* sortbuf = list + kListCap;                               // [kSortCap]
  uint32_t* hist = reinterpret_cast<uint32_t*>(sortbuf + kSortCap);            // [2048]
  uint32_t* tmax = hist + 2048;                                                // [kThreads]
  uint32_t* w_scratch = tmax + kThreads;                                       // [kWarps]
  uint32_t* s_misc = w_scratch + kWarps;                                       // [4]
  uint32_t* s_cnt = s_misc + 4;                                                // [2]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = rep + static_cast<size_t>(blockIdx.x) * V;
  const bool vec_ok = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
  const float4* row4 = reinterpret_cast<const float4*>(row);
  const int n4 = V >> 2;
  if (tid < 2) s_cnt[tid] = 0;

  // ---- A. thread maxima over unblocked candidates
  uint32_t my_max = 0;
  auto consider = [&](float x, int v) {
    const uint32_t key = (x > 0.f) ? __float_as_uint(x) : 0u;
    if (key > my_max && !(block != nullptr && block[v] != 0)) my_max = key;
  };
  if (vec_ok) {
    for (int i0 = tid; i0 < n4; i0 += 4 * kThreads) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kThreads;
        x[u] = (i < n4) ? __ldg(&row4[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = 4 * (i0 + u * kThreads);
        consider(x[u].x, v);
        consider(x[u].y, v + 1);
        consider(x[u].z, v + 2);
        consider(x[u].w, v + 3);
      }
    }
  } else {
    for (int v = tid; v < V; v += kThreads) consider(row[v], v);
  }
  tmax[tid] = my_max;
  __syncthreads();
  uint32_t floor_key = 0, dummy = 0;
  if (radix_select<kThreads>([&](int i) -> uint32_t { return tmax[i]; }, kThreads, static_cast<uint32_t>(k), 0u,
                             hist, w_scratch, s_misc, &floor_key, &dummy))
    floor_key = 0;
  const uint32_t lo_key = max(floor_key, 1u);

  // ---- B. collect every candidate >= lower bound (unordered)
  auto append = [&](bool pred, uint32_t key, int v) {  // warp-collective
    const uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return;
    uint32_t base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(&s_cnt[0], static_cast<uint32_t>(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pred) {
      const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
      if (slot < static_cast<uint32_t>(kListCap))
        list[slot] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(v));
    }
  };
  auto test = [&](float x, int v, bool in) -> bool {
    if (!in || !(x > 0.f) || __float_as_uint(x) < lo_key) return false;
    return !(block != nullptr && block[v] != 0);
  };
  if (vec_ok) {
    for (int b0 = warp * 32; b0 < n4; b0 += 2 * kThreads) {  // warp-uniform trip count
      float4 x[2];
      bool in[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = b0 + u * kThreads + lane;
        in[u] = i < n4;
        x[u] = in[u] ? __ldg(&row4[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int v = 4 * (b0 + u * kThreads + lane);
        append(test(x[u].x, v, in[u]), __float_as_uint(x[u].x), v);
        append(test(x[u].y, v + 1, in[u]), __float_as_uint(x[u].y), v + 1);
        append(test(x[u].z, v + 2, in[u]), __float_as_uint(x[u].z), v + 2);
        append(test(x[u].w, v + 3, in[u]), __float_as_uint(x[u].w), v + 3);
      }
    }
  } else {
    for (int b0 = warp * 32; b0 < V; b0 += kThreads) {
      const int v = b0 + lane;
      const bool in = v < V;
      const float x = in ? row[v] : 0.f;
      append(test(x, v, in), __float_as_uint(x), v);
    }
  }
  __syncthreads();
  const uint32_t n_list = s_cnt[0];
  if (n_list > static_cast<uint32_t>(kListCap)) {  // block-uniform
    if (tid == 0) counts[blockIdx.x] = kNeedsFallback;
    return;
  }

  // ---- C. exact k-th largest inside the list; move survivors to the sort buffer
  uint32_t T = 0, need = 0;
  const bool take_all = radix_select<kThreads>([&](int i) -> uint32_t { return static_cast<uint32_t>(list[i] >> 32); },
                                               static_cast<int>(n_list), static_cast<uint32_t>(k), 0u, hist,
                                               w_scratch, s_misc, &T, &need);
  if (take_all) T = 0;  // every list entry survives (there are at most k of them)
  for (int b0 = warp * 32; b0 < static_cast<int>(n_list); b0 += kThreads) {
    const int i = b0 + lane;
    const unsigned long long e = (i < static_cast<int>(n_list)) ? list[i] : 0ull;
    const bool pred = (i < static_cast<int>(n_list)) && static_cast<uint32_t>(e >> 32) >= max(T, 1u);
    const uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) continue;
    uint32_t base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(&s_cnt[1], static_cast<uint32_t>(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pred) {
      const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
      if (slot < static_cast<uint32_t>(kSortCap)) sortbuf[slot] = e;
    }
  }
  __syncthreads();
  const uint32_t n_sort = s_cnt[1];
  if (n_sort > static_cast<uint32_t>(kSortCap)) {  // too many ties at the k-th value
    if (tid == 0) counts[blockIdx.x] = kNeedsFallback;
    return;
  }
  int npad = 2;
  while (npad < static_cast<int>(n_sort)) npad <<= 1;
  for (int i = static_cast<int>(n_sort) + tid; i < npad; i += kThreads) sortbuf[i] = 0ull;
  __syncthreads();

  // ---- D. bitonic sort, descending, on (weight bits, ~id)
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < npad / 2; i += kThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = sortbuf[lo], b = sortbuf[hi];
        if ((a < b) == desc) {
          sortbuf[lo] = b;
          sortbuf[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  const uint32_t n_out = min(static_cast<uint32_t>(k), n_sort);
  int32_t* oid = ids + static_cast<size_t>(blockIdx.x) * k;
  float* oval = vals + static_cast<size_t>(blockIdx.x) * k;
  for (int i = tid; i < k; i += kThreads) {
    if (static_cast<uint32_t>(i) < n_out) {
      const unsigned long long p = sortbuf[i];
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
  // 1. fast path for every row; 2. the exact kernel re-does only the rows the fast path flagged (it exits at once
  //    for all others), so there is no host round trip and no data-dependent launch.
  const size_t fast_smem = static_cast<size_t>(kListCap + kSortCap) * 8 + (2048 + 1024 + 32 + 4 + 2) * 4;
  SPLADE_CUDA_OK(cudaFuncSetAttribute(topk_fast_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(fast_smem)));
  topk_fast_kernel<1024><<<B, 1024, fast_smem, stream>>>(rep, V, blocklist, k, ids, vals, counts);
  SPLADE_CUDA_OK(cudaGetLastError());
  const size_t fixed_1024 = static_cast<size_t>(kpad) * 8 + 2048 * 4 + (2 * 32 + 4 + 1024) * 4;
  const size_t fixed_256 = static_cast<size_t>(kpad) * 8 + 2048 * 4 + (2 * 8 + 4 + 256) * 4;
  const size_t staged_bytes = fixed_1024 + static_cast<size_t>(V) * 4;
  if (staged_bytes <= static_cast<size_t>(kSmemBudget)) {
    auto kernel = topk_kernel<1024, true>;
    SPLADE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(staged_bytes)));
    kernel<<<B, 1024, staged_bytes, stream>>>(rep, V, blocklist, k, kpad, ids, vals, counts, 1);
  } else {
    topk_kernel<256, false><<<B, 256, fixed_256, stream>>>(rep, V, blocklist, k, kpad, ids, vals, counts, 1);
  }
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // namespace splade
