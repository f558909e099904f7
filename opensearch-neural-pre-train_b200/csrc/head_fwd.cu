// Fused SPLADE-max head forward for sm_100a.
//
// Replaces, in one kernel, the reference chain
//   logits = decoder(hidden)                       (transformers modeling_modernbert.py:562, K1)
//   w = log1p(relu(logits))                        (src/model/splade_modern.py:76, K2/K3)
//   w = w * attention_mask[..., None]              (src/model/splade_modern.py:79-80, K4)
//   sparse_repr, argmax = w.max(dim=1)             (src/model/splade_modern.py:83, K5)
//   token_weights = w.max(dim=-1).values           (src/model/splade_modern.py:86, K6)
// without ever writing the [B,S,V] logits to HBM.
//
// Orientation: the GEMM is computed TRANSPOSED, D[v, t] = sum_h W[v,h] * hidden[t,h]:
//   * UMMA M (= 128 TMEM lanes)  <- vocabulary rows of decoder.weight  (A operand, K-major, TMA 2D)
//   * UMMA N (<= 256 TMEM cols)  <- tokens of whole documents          (B operand, K-major, TMA 3D box
//                                                                       (64 h, s_box tokens, b_box docs))
// so every epilogue thread owns ONE vocabulary entry and walks the sequence positions of a document
// left-to-right in registers: max-pool over the sequence and its first-occurrence argmax need no
// cross-thread traffic and no atomics (for S <= 256).  The per-position max over the vocabulary
// (token_weights) is one REDUX per column plus one 32-lane RED per 32 positions.
//
// Numerics follow the reference's autocast(bf16) path exactly: fp32 accumulate, + bias in fp32, ONE
// RNE rounding to bf16, relu, compare on the rounded value with ties to the lowest position, padded
// positions contribute 0, argmax 0 when the pooled value is 0; log1p is applied once per (doc, term)
// after the max (log1p∘relu is monotone).
//
// Warp roles (256 threads, 1 CTA/SM, persistent over (vocab tile, token tile) items):
//   warp 0 lane 0 : TMA producer         (4-stage ring of {A 16 KB, B <=32 KB}, 128B swizzle)
//   warp 1 lane 0 : tcgen05.mma issuer   (128 x N x 16 per instruction, fp32 accumulators in TMEM)
//   warp 2        : TMEM allocator
//   warps 4..7    : epilogue             (TMEM lanes 32*(warp%4) .. +31), double-buffered accumulators;
//                   the next tile's mask word and bias are prefetched one tile ahead
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdlib>

namespace splade {
namespace {

constexpr int BLOCK_V = 128;  // UMMA M
constexpr int BLOCK_T = 256;  // max UMMA N
constexpr int BLOCK_K = 64;   // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BLOCK_V * BLOCK_K * 2;
constexpr int B_BYTES = BLOCK_T * BLOCK_K * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 256;
constexpr int TMEM_COLS = 512;  // two 128 x 256 fp32 accumulators
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;

struct HeadFwdParams {
  const float* bias;
  const int32_t* mask;  // [B,S], nonzero = real token
  const uint32_t* tile_masks;  // [n_tt][8]: 256-bit column validity bitmap per token tile (prepass)
  const int32_t* tile_n16;     // [n_tt]: columns up to the last valid one, rounded up to 16 (UMMA N of the tile)
  int trim_rows;               // 1: one document per tile -> the producer also loads only ceil64(n16) token rows
  float* rep;           // [B,V]
  int32_t* argmax;      // [B,V]
  uint32_t* packed;     // [B,V] (bf16 bits << 16 | 0xFFFF - pos), only when a doc spans several tiles
  uint32_t* tokw_bits;  // [B,S] bf16 bits of the per-position max logit, or nullptr
  int B, S, H, V;
  int s_box, b_box;     // tokens per doc / docs per token tile (TMA box of hidden)
  int n_tt, n_vt;       // token tiles, vocab tiles
  int chunks_per_doc;   // 1 when S <= 256
  int umma_n;
  uint32_t stage_tx_bytes;
  // CTA-pair kernel only: how the token rows of a tile are split between the two CTAs
  int pair_split_docs;   // 1: each CTA takes b_box/2 documents; 0: each CTA takes half of one document's rows
  int pair_half_rows;    // token rows per CTA (= umma_n / 2)
  const int32_t* plan_mode;  // device flag written by varlen_plan_kernel (nullptr: padded path unconditionally);
                             // != 0 means the packed kernel owns this call and the padded kernels exit at once
};

struct TileCoord {
  int vt, b0, nb, s0, ns;
};

// item -> (vocab tile, token tile).  vt_mul / vt_add map the item's vocabulary index to this CTA's 128-row tile:
// (1, 0) for the single-CTA kernel, (2, cluster rank) for the CTA-pair kernel (256 rows per item).
__device__ __forceinline__ TileCoord tile_coord(const HeadFwdParams& p, int item, int vt_mul = 1, int vt_add = 0) {
  TileCoord t;
  const int vi = item / p.n_tt;
  t.vt = vi * vt_mul + vt_add;
  const int tt = item - vi * p.n_tt;
  if (p.chunks_per_doc == 1) {
    t.b0 = tt * p.b_box;
    t.nb = min(p.b_box, p.B - t.b0);
    t.s0 = 0;
    t.ns = p.S;
  } else {
    t.b0 = tt / p.chunks_per_doc;
    t.nb = 1;
    t.s0 = (tt - t.b0 * p.chunks_per_doc) * BLOCK_T;
    t.ns = min(BLOCK_T, p.S - t.s0);
  }
  return t;
}

// Prepass: the 256-bit column-validity bitmap of every token tile (bit c = column c of the tile is a real,
// unmasked token).  One CTA of 256 threads (8 warps = 8 words) per token tile; a few microseconds, and it takes
// eight dependent global loads + ballots per tile out of the epilogue warps' critical path.
__global__ void __launch_bounds__(256)
build_tile_masks_kernel(const HeadFwdParams p, uint32_t* __restrict__ out, int32_t* __restrict__ n16_out) {
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = -1;
  __syncthreads();
  TileCoord t = tile_coord(p, blockIdx.x);  // vt = 0 since blockIdx.x < n_tt
  const int col = threadIdx.x;
  const int d = col / p.s_box;
  const int s = col - d * p.s_box;
  bool ok = false;
  if (d < t.nb && s < t.ns) ok = p.mask[static_cast<size_t>(t.b0 + d) * p.S + t.s0 + s] != 0;
  const uint32_t word = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0) {
    out[blockIdx.x * 8 + (threadIdx.x >> 5)] = word;
    if (word) atomicMax(&s_last, static_cast<int>(threadIdx.x) + 31 - __clz(word));
  }
  __syncthreads();
  // Padding is a suffix in practice (right-padded batches): everything after the last valid column of the tile
  // needs no MMA work at all.  N must be a multiple of 16 and at least 16.
  if (threadIdx.x == 0) n16_out[blockIdx.x] = min(p.umma_n, max(16, (s_last + 1 + 15) / 16 * 16));
}

// One 32-column window of one document, for one vocabulary entry per thread.
// Returns max over the window of key_j = (bf16(acc_j + bias) << 16) | (31 - j) as a signed int:
// negative logits (sign bit set) lose against 0, positive bf16 bit patterns order like the floats,
// and equal values keep the smallest j.
template <bool kAllValid, bool kTokW>
__device__ __forceinline__ int window_max(const uint32_t (&r)[32], float bias_v, uint32_t vw, uint32_t lane,
                                          int& tok_mine) {
  int wkey = 0;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float x0 = __uint_as_float(r[j]) + bias_v;
    const float x1 = __uint_as_float(r[j + 1]) + bias_v;
    const uint32_t pk = cvt_bf16x2(x1, x0);
    int k0 = static_cast<int>((pk << 16) | static_cast<uint32_t>(31 - j));
    int k1 = static_cast<int>((pk & 0xFFFF0000u) | static_cast<uint32_t>(30 - j));
    if (!kAllValid) {
      k0 = ((vw >> j) & 1u) ? k0 : 0;
      k1 = ((vw >> (j + 1)) & 1u) ? k1 : 0;
    }
    wkey = max(wkey, max(k0, k1));
    if (kTokW) {
      const int t0 = warp_max_s32(k0);
      const int t1 = warp_max_s32(k1);
      tok_mine = (lane == static_cast<uint32_t>(j)) ? t0 : tok_mine;
      tok_mine = (lane == static_cast<uint32_t>(j + 1)) ? t1 : tok_mine;
    }
  }
  return wkey;
}

// Everything the epilogue warps need to know about the kernel they run in.
struct EpilogueCtx {
  uint32_t tmem_base;
  uint64_t* tfull_bar;    // [2] accumulator-ready barriers of THIS CTA
  uint32_t tempty_addr;   // shared::cluster address of tempty_bar[0] of the CTA that issues the MMAs
  int item0, item_step, n_items;
  int vt_mul, vt_add;
};

// Epilogue of one CTA: drains 128 TMEM lanes (vocabulary rows) x up to 256 columns (tokens) per item.
// Called by the four epilogue warps (warp % 4 = TMEM lane quarter).
template <bool kTokW>
__device__ __forceinline__ void epilogue_loop(const HeadFwdParams& p, const EpilogueCtx& c, uint32_t warp,
                                              uint32_t lane) {
  const uint32_t q = warp & 3u;  // TMEM lane quarter this warp may read
  uint32_t it = 0;
  // Operands of the NEXT tile (its mask word and this thread's bias) are fetched one tile ahead so that their
  // global-load latency overlaps the window processing of the current tile (one warp per SM sub-partition has
  // nothing else to hide it with).
  uint32_t next_word = 0;
  float next_bias = 0.f;
  if (c.item0 < c.n_items) {
    const int vi0 = c.item0 / p.n_tt;
    const int vt0 = vi0 * c.vt_mul + c.vt_add;
    const int tt0 = c.item0 - vi0 * p.n_tt;
    if (lane < 8) next_word = __ldg(&p.tile_masks[tt0 * 8 + lane]);
    const int v0 = vt0 * BLOCK_V + static_cast<int>(q * 32u + lane);
    if (v0 < p.V) next_bias = __ldg(&p.bias[v0]);
  }
  for (int item = c.item0; item < c.n_items; item += c.item_step, ++it) {
    const uint32_t acc = it & 1u;
    const uint32_t acc_phase = (it >> 1) & 1u;
    const TileCoord t = tile_coord(p, item, c.vt_mul, c.vt_add);
    const uint32_t my_word = next_word;  // lane w < 8 keeps word w of this tile's validity bitmap
    const float bias_v = next_bias;
    const int v = t.vt * BLOCK_V + static_cast<int>(q * 32u + lane);
    {
      const int nitem = item + c.item_step;
      next_word = 0;
      next_bias = 0.f;
      if (nitem < c.n_items) {
        const int nvi = nitem / p.n_tt;
        const int nvt = nvi * c.vt_mul + c.vt_add;
        const int ntt = nitem - nvi * p.n_tt;
        if (lane < 8) next_word = __ldg(&p.tile_masks[ntt * 8 + lane]);
        const int nv = nvt * BLOCK_V + static_cast<int>(q * 32u + lane);
        if (nv < p.V) next_bias = __ldg(&p.bias[nv]);
      }
    }
    // Is every column of every document of this tile valid?  (the common case: no padding inside the tile)
    const int cols_valid = (t.ns == p.s_box) ? t.nb * p.s_box : -1;
    uint32_t expect = 0;
    if (lane < 8) {
      const int rem = cols_valid - 32 * static_cast<int>(lane);
      expect = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    }
    const bool tile_full = (cols_valid >= 0) && (__ballot_sync(0xffffffffu, my_word != expect) == 0u);

    mbar_wait(&c.tfull_bar[acc], acc_phase);
    tcgen05_fence_after();
    const uint32_t t_lane = c.tmem_base + ((q * 32u) << 16) + acc * BLOCK_T;

    for (int d = 0; d < t.nb; ++d) {
      const int b = t.b0 + d;
      const int col_base = d * p.s_box;
      int best_val = 0, best_pos = 0;
      const int nwin = (t.ns + 31) >> 5;
      for (int cw = 0; cw < nwin; ++cw) {
        const int w0 = min(col_base + 32 * cw, BLOCK_T - 32);
        const int jlo = max(0, col_base - w0);
        const int jhi = min(32, col_base + t.ns - w0);
        uint32_t range = (jhi >= 32) ? 0xffffffffu : ((1u << jhi) - 1u);
        range &= 0xffffffffu << jlo;
        uint32_t vw = range;
        if (!tile_full) {  // warp-uniform
          const uint32_t lo = __shfl_sync(0xffffffffu, my_word, w0 >> 5);
          const uint32_t hi = __shfl_sync(0xffffffffu, my_word, min(7, (w0 >> 5) + 1));
          vw &= __funnelshift_r(lo, hi, w0 & 31);
        }
        if (vw == 0u) continue;  // warp-uniform: fully padded window

        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(w0), r);
        tmem_ld_wait();

        int tok_mine = 0;
        int wkey;
        if (vw == 0xffffffffu) {
          wkey = window_max<true, kTokW>(r, bias_v, vw, lane, tok_mine);
        } else {
          wkey = window_max<false, kTokW>(r, bias_v, vw, lane, tok_mine);
        }
        const int pos0 = t.s0 + (w0 - col_base);  // sequence position of window column 0
        const int wval = wkey >> 16;
        if (wval > best_val) {  // strict: earlier windows win ties (first occurrence)
          best_val = wval;
          best_pos = pos0 + 31 - (wkey & 31);
        }
        if (kTokW) {
          const int tv = tok_mine >> 16;
          if (tv > 0 && ((vw >> lane) & 1u)) {
            atomicMax(&p.tokw_bits[static_cast<size_t>(b) * p.S + pos0 + static_cast<int>(lane)],
                      static_cast<uint32_t>(tv));
          }
        }
      }
      if (v < p.V) {
        const size_t o = static_cast<size_t>(b) * p.V + v;
        if (p.chunks_per_doc == 1) {
          p.rep[o] = log1pf(bf16_bits_to_float(static_cast<uint32_t>(best_val)));
          p.argmax[o] = best_pos;
        } else {
          atomicMax(&p.packed[o], (static_cast<uint32_t>(best_val) << 16) |
                                      (0xFFFFu - static_cast<uint32_t>(best_pos)));
        }
      }
    }
    tcgen05_fence_before();
    mbar_arrive_cluster(c.tempty_addr + acc * 8u);
  }
}

template <bool kTokW>
__global__ void __launch_bounds__(NUM_THREADS, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                const __grid_constant__ CUtensorMap tmap_h64, const __grid_constant__ CUtensorMap tmap_h128,
                const __grid_constant__ CUtensorMap tmap_h192, const HeadFwdParams p) {
  if (p.plan_mode != nullptr && __ldg(p.plan_mode) != 0) return;  // grid-uniform: the packed kernel runs instead
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024 B alignment

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_h);
    if (p.trim_rows) {
      tma_prefetch_desc(&tmap_h64);
      tma_prefetch_desc(&tmap_h128);
      tma_prefetch_desc(&tmap_h192);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_items = p.n_vt * p.n_tt;
  const int num_kb = p.H / BLOCK_K;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const TileCoord t = tile_coord(p, item);
        // Tail trimming: with one document per tile only ceil64(n16) token rows are fetched.
        const CUtensorMap* hmap = &tmap_h;
        uint32_t tx_bytes = p.stage_tx_bytes;
        if (p.trim_rows) {
          const int n16 = __ldg(&p.tile_n16[item % p.n_tt]);
          const int sel = (n16 + 63) / 64;  // 1..4
          const int rows = min(p.s_box, 64 * sel);
          hmap = sel == 1 ? &tmap_h64 : (sel == 2 ? &tmap_h128 : (sel == 3 ? &tmap_h192 : &tmap_h));
          tx_bytes = static_cast<uint32_t>(A_BYTES + rows * BLOCK_K * 2);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* a_dst = smem + stage * STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_2d(a_dst, &tmap_w, &full_bar[stage], kb * BLOCK_K, t.vt * BLOCK_V);
          tma_load_3d(b_dst, hmap, &full_bar[stage], kb * BLOCK_K, t.s0, t.b0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        const uint32_t acc_phase = (it >> 1) & 1u;
        // UMMA N of this tile: columns after the last valid token are never computed (MMA time ~ N)
        const uint32_t idesc =
            make_idesc_bf16_f32(BLOCK_V, static_cast<uint32_t>(__ldg(&p.tile_n16[item % p.n_tt])));
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_T;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t a_desc = make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t b_desc = make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps)
    EpilogueCtx c;
    c.tmem_base = tmem_base;
    c.tfull_bar = tfull_bar;
    c.tempty_addr = smem_u32(&tempty_bar[0]);
    c.item0 = static_cast<int>(blockIdx.x);
    c.item_step = static_cast<int>(gridDim.x);
    c.n_items = total_items;
    c.vt_mul = 1;
    c.vt_add = 0;
    epilogue_loop<kTokW>(p, c, warp, lane);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2).  A cluster of two CTAs (one TPC) works on a 256-row vocabulary tile
// x one token tile: each CTA stages ITS 128 vocabulary rows (A) and HALF of the token rows (B); the leader's
// single MMA thread issues 256 x N x 16 instructions that read both CTAs' shared memory and write both CTAs'
// TMEM.  Per CTA and k-block that is 16 KB + 16 KB of TMA traffic instead of 16 KB + 32 KB, and the B operand is
// read from shared memory once per pair instead of once per CTA - less L2->SM and SMEM->tensor-core traffic for
// the same math.  Each CTA runs the same epilogue as the single-CTA kernel on its own 128 TMEM lanes.
// ---------------------------------------------------------------------------------------------------------
constexpr int P_STAGES = 6;
constexpr int P_A_BYTES = BLOCK_V * BLOCK_K * 2;         // 16 KB: this CTA's 128 vocabulary rows
constexpr int P_B_BYTES = (BLOCK_T / 2) * BLOCK_K * 2;   // 16 KB: this CTA's half of the token rows
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 256 + 1024;

template <bool kTokW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
head_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                     const HeadFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);  // used in the leader
  uint64_t* empty_bar = full_bar + P_STAGES;                                         // per CTA
  uint64_t* tfull_bar = empty_bar + P_STAGES;                                        // per CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                              // used in the leader
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_h);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's arrive.expect_tx; bytes come from both CTAs' TMA loads
      mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);     // one multicast tcgen05.commit
      mbar_init(&tempty_bar[a], 256);  // 128 epilogue threads of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / TMA credit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_vt2 = (p.n_vt + 1) / 2;
  const int total_items = n_vt2 * p.n_tt;
  const int num_kb = p.H / BLOCK_K;
  const int item0 = static_cast<int>(blockIdx.x >> 1);
  const int item_step = static_cast<int>(gridDim.x >> 1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t full0 = map_to_cta(smem_u32(&full_bar[0]), 0);  // leader's full barriers
      for (int item = item0; item < total_items; item += item_step) {
        const TileCoord t = tile_coord(p, item, 2, static_cast<int>(rank));
        // this CTA's half of the token rows: docs [b0 + rank*b_half, ..) or positions [s0 + rank*s_half, ..)
        const int my_s = t.s0 + (p.pair_split_docs ? 0 : static_cast<int>(rank) * p.pair_half_rows);
        const int my_b = t.b0 + (p.pair_split_docs ? static_cast<int>(rank) * (p.b_box / 2) : 0);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* a_dst = smem + stage * P_STAGE_BYTES;
          uint8_t* b_dst = a_dst + P_A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], p.stage_tx_bytes);
          const uint32_t bar = full0 + stage * 8u;
          tma_load_2d_pair(a_dst, &tmap_w, bar, kb * BLOCK_K, t.vt * BLOCK_V);
          tma_load_3d_pair(b_dst, &tmap_h, bar, kb * BLOCK_K, my_s, my_b);
          if (++stage == P_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(2 * BLOCK_V, static_cast<uint32_t>(p.umma_n));
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = item0; item < total_items; item += item_step, ++it) {
        const uint32_t acc = it & 1u;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);  // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_T;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * P_STAGE_BYTES);
          const uint32_t b_addr = a_addr + P_A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t a_desc = make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t b_desc = make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
            umma_bf16_pair(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], 0x3);  // frees this stage in BOTH CTAs
          if (++stage == P_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_pair(&tfull_bar[acc], 0x3);  // accumulator halves ready in both CTAs
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps, both CTAs)
    EpilogueCtx c;
    c.tmem_base = tmem_base;
    c.tfull_bar = tfull_bar;
    c.tempty_addr = map_to_cta(smem_u32(&tempty_bar[0]), 0);
    c.item0 = item0;
    c.item_step = item_step;
    c.n_items = total_items;
    c.vt_mul = 2;
    c.vt_add = static_cast<int>(rank);
    epilogue_loop<kTokW>(p, c, warp, lane);
  }

  __syncwarp();  // reconverge the single-lane role warps before the aligned cluster barrier
  tcgen05_fence_before();
  cluster_sync_all();  // the peer may still be reading our shared memory / arriving on our barriers
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// packed (bf16 bits << 16 | 0xFFFF - pos) -> fp32 log1p weight + int32 argmax (documents longer than one tile)
__global__ void finalize_packed_kernel(const uint32_t* __restrict__ packed, float* __restrict__ rep,
                                       int32_t* __restrict__ argmax, size_t n, const int32_t* plan_mode) {
  if (plan_mode != nullptr && __ldg(plan_mode) != 0) return;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = packed[i];
  rep[i] = log1pf(bf16_bits_to_float(k >> 16));
  argmax[i] = static_cast<int32_t>(0xFFFFu - (k & 0xFFFFu));
}

// token_weights[b,s] = log1p(max_v relu(bf16 logit)) ; 0 for padded positions (buffer was zero-initialised)
__global__ void finalize_tokw_kernel(const uint32_t* __restrict__ bits, float* __restrict__ tokw, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tokw[i] = log1pf(bf16_bits_to_float(bits[i]));
}

// ---------------------------------------------------------------------------------------------------------
// Packed (variable-length) path.  Right-padded batches waste tensor-core time on padding: a batch of queries
// padded to S=64 with ~10 real tokens each is 85 % zeros.  Here the REAL tokens of the whole batch are packed
// back to back into [T, H], the token operand is a plain 2-D TMA tile of 256 consecutive packed tokens, and the
// epilogue follows the document boundaries inside the tile: every thread still owns one vocabulary entry and
// walks its columns left to right, but it flushes (doc, term) whenever the column's document changes.
//
//   * Every document starts at a multiple of 8 packed columns; the <= 7 filler columns behind its last token are
//     REPLICAS of that token (same hidden row, same (doc, position) descriptor).  A replica can never change a
//     maximum, loses every tie to the original (it sits to its right) and sends token_weights the value the
//     original already sent - so the epilogue needs no per-column masks, and document boundaries fall on the
//     static 8-column groups of a 32-column window.
//   * Eight epilogue warps: warps 4-7 drain columns 0-127 of the accumulator, warps 8-11 columns 128-255 (two
//     warps per SM sub-partition hide each other's latencies).  A document that lies inside one 128-column half
//     is stored directly; one that straddles halves is merged through the same 32-bit atomicMax key as the
//     S > 256 padded path and decoded by a finalize pass.
//
// Everything is decided ON THE DEVICE (no host synchronisation): a plan kernel compares the MMA/L2 cost of the
// padded tiling with the packed one and sets plan.mode; both main kernels are launched and the one that is not
// selected exits at its first instruction.
// ---------------------------------------------------------------------------------------------------------
constexpr int VL_ALIGN = 8;                 // packed columns per document are rounded up to this
constexpr int VL_HALF = BLOCK_T / 2;        // columns drained by one epilogue warp group
constexpr int VL_THREADS = 384;             // producer, MMA, TMEM-alloc, idle + 2 x 4 epilogue warps
constexpr uint32_t VL_SPAN_BIT = 0x80000000u;

struct VarlenParams {
  const float* bias;
  const VarlenPlan* plan;
  const uint32_t* meta;   // [rows]: (straddles << 31) | (doc << 16) | position of every packed column
  const uint8_t* spans;   // [B]: 1 = the document's columns straddle a 128-column half-tile boundary
  const int32_t* len;     // [B]: real tokens per document
  float* rep;             // [B,V]
  int32_t* argmax;        // [B,V]
  uint32_t* keys;         // [B,V] (bf16 bits << 16 | 0xFFFF - pos) merge buffer of straddling documents
  uint32_t* tokw_bits;    // [B,S] or nullptr
  int B, S, H, V, n_vt;
};

// len[b] = number of unmasked positions of document b (one warp per document)
__global__ void __launch_bounds__(256)
varlen_count_kernel(const int32_t* __restrict__ mask, int B, int S, int32_t* __restrict__ len) {
  const int b = blockIdx.x * 8 + static_cast<int>(threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = static_cast<int>(threadIdx.x & 31u);
  int c = 0;
  for (int s = lane; s < S; s += 32) c += mask[static_cast<size_t>(b) * S + s] != 0 ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) len[b] = c;
}

// One CTA: exclusive scan of the 8-aligned lengths (packed offset of every document), straddle flags, and the
// decision.  Cost model in "operand rows streamed from L2 per (vocab tile, token tile)": 128 vocabulary rows +
// the token rows the tile loads (the kernel is within a few percent of both its tensor and its L2->SM limit, so
// a tile with fewer columns still pays for its vocabulary rows).
__global__ void __launch_bounds__(1024)
varlen_plan_kernel(int32_t* __restrict__ len, int B, const int32_t* __restrict__ tile_n16, int n_tt_pad,
                   int trim_rows, int pad_cols, int force_mode, int32_t* __restrict__ off,
                   uint8_t* __restrict__ spans, VarlenPlan* __restrict__ plan, int32_t* __restrict__ plan_out,
                   const int32_t* __restrict__ mask, int S) {
  // mask != nullptr: the lengths are counted here as well (small batches: saves the varlen_count_kernel launch)
  __shared__ int s_len[1024];
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  __shared__ int s_real;
  __shared__ unsigned long long s_cost;
  const int tid = static_cast<int>(threadIdx.x);
  const int lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_carry = 0;
    s_real = 0;
    s_cost = 0ull;
  }
  __syncthreads();
  int real = 0;
  for (int base = 0; base < B; base += 1024) {
    const int i = base + tid;
    if (mask != nullptr) {  // one warp per document, 32 documents per warp and chunk
      for (int d = warp; d < 1024 && base + d < B; d += 32) {
        const int32_t* mrow = mask + static_cast<size_t>(base + d) * S;
        int c = 0;
        for (int s = lane; s < S; s += 32) c += mrow[s] != 0 ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) {
          s_len[d] = c;
          len[base + d] = c;
        }
      }
      __syncthreads();
    }
    const int l = i < B ? (mask != nullptr ? s_len[tid] : len[i]) : 0;
    real += l;
    const int x = (l + VL_ALIGN - 1) / VL_ALIGN * VL_ALIGN;
    int incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      int wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += y;
      }
      s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int carry = s_carry;
    const int excl = carry + s_warp[warp] + incl - x;
    if (i < B) {
      off[i] = excl;
      spans[i] = (x > 0 && (excl / VL_HALF) != ((excl + x - 1) / VL_HALF)) ? 1 : 0;
    }
    __syncthreads();
    if (tid == 1023) s_carry = excl + x;
    __syncthreads();
  }
  real = __reduce_add_sync(0xffffffffu, real);
  if (lane == 0 && real) atomicAdd(&s_real, real);
  unsigned long long cost = 0;
  for (int t = tid; t < n_tt_pad; t += 1024) {
    const int rows = trim_rows ? min(pad_cols, (tile_n16[t] + 63) / 64 * 64) : pad_cols;
    cost += static_cast<unsigned long long>(BLOCK_V + rows);
  }
  for (int d = 16; d > 0; d >>= 1) cost += __shfl_down_sync(0xffffffffu, cost, d);
  if (lane == 0 && cost) atomicAdd(&s_cost, cost);
  __syncthreads();
  if (tid == 0) {
    const int T = s_carry;
    off[B] = T;
    const int n_tt = (T + BLOCK_T - 1) / BLOCK_T;
    const int tail = T - BLOCK_T * (n_tt - 1);
    const int n16_last = T > 0 ? max(16, (tail + 15) / 16 * 16) : 16;
    const unsigned long long cost_packed = static_cast<unsigned long long>(n_tt) * (BLOCK_V + BLOCK_T);
    // packing, the merge keys and the finalize pass cost a few percent of a full batch: ask for a 7 % win
    int mode = (T > 0 && cost_packed * 107ull < s_cost * 100ull) ? 1 : 0;
    if (force_mode >= 0) mode = force_mode;
    plan->mode = mode;
    plan->T = T;
    plan->n_tt = n_tt;
    plan->n16_last = n16_last;
    plan->T_real = s_real;
    if (plan_out != nullptr) {
      plan_out[0] = mode;
      plan_out[1] = s_real;
      plan_out[2] = n_tt;
      plan_out[3] = n16_last;
    }
  }
}

// grid (ceil(S/256), B): copy the unmasked rows of 256 positions of one document to their packed place and record
// the descriptor of every packed column.  The CTA of chunk 0 also writes the replicas of the document's last token
// behind it (up to the next multiple of 8 columns) and clears the merge keys if the document straddles half-tiles.
__global__ void __launch_bounds__(256)
varlen_pack_kernel(const uint4* __restrict__ hidden, const int32_t* __restrict__ mask,
                   const int32_t* __restrict__ len, const int32_t* __restrict__ off,
                   const uint8_t* __restrict__ spans, const VarlenPlan* __restrict__ plan, int S, int H, int V,
                   uint4* __restrict__ packed, uint32_t* __restrict__ meta, uint32_t* __restrict__ keys) {
  if (plan->mode == 0) return;
  __shared__ int s_cnt[8];
  __shared__ int s_red[8];
  __shared__ int s_src[256];
  __shared__ int s_last;
  const int b = static_cast<int>(blockIdx.y);
  const int c0 = static_cast<int>(blockIdx.x) * 256;
  const int tid = static_cast<int>(threadIdx.x);
  const int lane = tid & 31, warp = tid >> 5;
  const int32_t* mrow = mask + static_cast<size_t>(b) * S;
  const int L = len[b];
  if (L == 0) return;
  const uint32_t tag = (spans[b] ? VL_SPAN_BIT : 0u) | (static_cast<uint32_t>(b) << 16);
  const int row_u4 = H / 8;
  if (blockIdx.x == 0) {
    if (spans[b]) {
      uint32_t* krow = keys + static_cast<size_t>(b) * V;
      for (int v = tid; v < V; v += 256) krow[v] = 0u;
    }
    const int fill = (L + VL_ALIGN - 1) / VL_ALIGN * VL_ALIGN - L;  // replicas of the last real token
    if (fill > 0) {
      if (tid == 0) s_last = -1;
      __syncthreads();
      int last = -1;
      for (int s = tid; s < S; s += 256) last = mrow[s] != 0 ? s : last;
      last = __reduce_max_sync(0xffffffffu, last);
      if (lane == 0 && last >= 0) atomicMax(&s_last, last);
      __syncthreads();
      last = s_last;
      for (int r = warp; r < fill; r += 8) {
        const size_t row = static_cast<size_t>(off[b] + L + r);
        if (hidden != nullptr) {
          const uint4* src = hidden + (static_cast<size_t>(b) * S + last) * row_u4;
          uint4* dst = packed + row * row_u4;
          for (int i = lane; i < row_u4; i += 32) dst[i] = __ldg(&src[i]);
        }
        if (lane == 0) meta[row] = tag | static_cast<uint32_t>(last);
      }
    }
  }
  // unmasked positions of this document before this chunk
  int before = 0;
  for (int s = tid; s < c0; s += 256) before += mrow[s] != 0 ? 1 : 0;
  before = __reduce_add_sync(0xffffffffu, before);
  if (lane == 0) s_red[warp] = before;
  const int s = c0 + tid;
  const bool ok = s < S && mrow[s] != 0;
  const uint32_t bal = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) s_cnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0, wpre = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    base += s_red[w];
    wpre += w < warp ? s_cnt[w] : 0;
    total += s_cnt[w];
  }
  const int dst0 = off[b] + base;  // packed row of this chunk's first unmasked position
  if (ok) {
    const int r = wpre + __popc(bal & ((1u << lane) - 1u));
    s_src[r] = s;
    meta[dst0 + r] = tag | static_cast<uint32_t>(s);
  }
  __syncthreads();
  if (hidden == nullptr) return;  // descriptors only: the caller (the native trunk) produces the packed rows itself
  for (int r = warp; r < total; r += 8) {
    const uint4* src = hidden + (static_cast<size_t>(b) * S + s_src[r]) * row_u4;
    uint4* dst = packed + static_cast<size_t>(dst0 + r) * row_u4;
    for (int i = lane; i < row_u4; i += 32) dst[i] = __ldg(&src[i]);
  }
}

template <bool kTokW>
__global__ void __launch_bounds__(VL_THREADS, 1)
head_fwd_varlen_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                       const VarlenParams p) {
  if (__ldg(&p.plan->mode) == 0) return;  // grid-uniform: the padded kernel owns this call
  const int T = __ldg(&p.plan->T);
  const int n_tt = __ldg(&p.plan->n_tt);
  const int n16_last = __ldg(&p.plan->n16_last);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint32_t* meta_smem = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + 256);  // [8 warps][4][32]

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_items = p.n_vt * n_tt;
  const int num_kb = p.H / BLOCK_K;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int vt = item / n_tt;
        const int tt = item - vt * n_tt;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* a_dst = smem + stage * STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(a_dst, &tmap_w, &full_bar[stage], kb * BLOCK_K, vt * BLOCK_V);
          tma_load_2d(b_dst, &tmap_x, &full_bar[stage], kb * BLOCK_K, tt * BLOCK_T);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        const uint32_t acc_phase = (it >> 1) & 1u;
        const int tt = item % n_tt;
        const uint32_t idesc =
            make_idesc_bf16_f32(BLOCK_V, static_cast<uint32_t>(tt == n_tt - 1 ? n16_last : BLOCK_T));
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_T;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t a_desc = make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t b_desc = make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (2 groups x 4 warps)
    const uint32_t q = warp & 3u;           // TMEM lane quarter
    const uint32_t grp = (warp - 4u) >> 2;  // column half of the accumulator
    uint32_t* my_meta = meta_smem + (warp - 4u) * 128u;  // this warp's 128 column descriptors
    uint32_t it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int vt = item / n_tt;
      const int tt = item - vt * n_tt;
      const int v = vt * BLOCK_V + static_cast<int>(q * 32u + lane);
      const float bias_v = v < p.V ? __ldg(&p.bias[v]) : 0.f;
      const int col0 = static_cast<int>(grp) * VL_HALF;                  // first column of this group
      const int ncols = min(VL_HALF, T - tt * BLOCK_T - col0);           // multiple of 8, may be <= 0
      // column descriptors of the half-tile: fetched before the accumulator wait, whose latency hides them
      __syncwarp();
#pragma unroll
      for (int w = 0; w < 4; ++w)
        my_meta[w * 32 + lane] = __ldg(&p.meta[static_cast<size_t>(tt) * BLOCK_T + col0 + w * 32 + lane]);
      __syncwarp();

      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_lane = tmem_base + ((q * 32u) << 16) + acc * BLOCK_T + static_cast<uint32_t>(col0);

      uint32_t cur_tag = 0xFFFFFFFFu;  // (straddle bit | doc) of the running document, none yet
      int best_val = 0, best_pos = 0;
      auto flush = [&]() {
        if (cur_tag != 0xFFFFFFFFu && v < p.V) {
          const size_t o = static_cast<size_t>(cur_tag & 0x7FFFu) * p.V + v;
          if (cur_tag & 0x8000u) {
            atomicMax(&p.keys[o], (static_cast<uint32_t>(best_val) << 16) | (0xFFFFu - static_cast<uint32_t>(best_pos)));
          } else {
            p.rep[o] = log1pf(bf16_bits_to_float(static_cast<uint32_t>(best_val)));
            p.argmax[o] = best_pos;
          }
        }
      };
      const int nwin = (ncols + 31) >> 5;
      for (int w = 0; w < nwin; ++w) {
        const int ngrp = min(4, (ncols - 32 * w) >> 3);  // valid 8-column groups of this window
        const uint32_t m = my_meta[w * 32 + lane];
        const uint32_t tag_j = m >> 16;
        // group g (columns 8g..8g+7) opens a new document when its tag differs from the running one; inside a
        // group the tag is constant by construction (documents start at multiples of 8 columns)
        const uint32_t tag0 = __shfl_sync(0xffffffffu, tag_j, 0);
        const uint32_t tag1 = __shfl_sync(0xffffffffu, tag_j, 8);
        const uint32_t tag2 = __shfl_sync(0xffffffffu, tag_j, 16);
        const uint32_t tag3 = __shfl_sync(0xffffffffu, tag_j, 24);

        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(32 * w), r);
        tmem_ld_wait();
        // signed keys (bf16(acc + bias) << 16) | (31 - j), in place
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float x0 = __uint_as_float(r[j]) + bias_v;
          const float x1 = __uint_as_float(r[j + 1]) + bias_v;
          const uint32_t pk = cvt_bf16x2(x1, x0);
          r[j] = (pk << 16) | static_cast<uint32_t>(31 - j);
          r[j + 1] = (pk & 0xFFFF0000u) | static_cast<uint32_t>(30 - j);
        }
        if (kTokW) {
          int tok_mine = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int t = warp_max_s32(static_cast<int>(r[j]));
            tok_mine = (lane == static_cast<uint32_t>(j)) ? t : tok_mine;
          }
          const int tv = tok_mine >> 16;
          if (tv > 0 && static_cast<int>(lane) < 8 * ngrp) {
            atomicMax(&p.tokw_bits[static_cast<size_t>(tag_j & 0x7FFFu) * p.S + (m & 0xFFFFu)],
                      static_cast<uint32_t>(tv));
          }
        }
        int k[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          int kk = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) kk = max(kk, static_cast<int>(r[8 * g + j]));
          k[g] = kk;
        }
        if (ngrp == 4 && tag0 == cur_tag && tag1 == tag0 && tag2 == tag0 && tag3 == tag0) {
          // the common case for long documents: the whole window continues the running document
          const int wkey = max(max(k[0], k[1]), max(k[2], k[3]));
          const uint32_t mm = __shfl_sync(0xffffffffu, m, 31 - (wkey & 31));
          if ((wkey >> 16) > best_val) {
            best_val = wkey >> 16;
            best_pos = static_cast<int>(mm & 0xFFFFu);
          }
        } else {
          const uint32_t tags[4] = {tag0, tag1, tag2, tag3};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ngrp) {  // warp-uniform
              if (tags[g] != cur_tag) {
                flush();
                cur_tag = tags[g];
                best_val = 0;
                best_pos = 0;
              }
              const uint32_t mm = __shfl_sync(0xffffffffu, m, 31 - (k[g] & 31));
              if ((k[g] >> 16) > best_val) {  // strict: the earliest position wins ties
                best_val = k[g] >> 16;
                best_pos = static_cast<int>(mm & 0xFFFFu);
              }
            }
          }
        }
      }
      flush();
      tcgen05_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Rows of straddling documents are decoded from the merge keys, rows of documents with no real token are zero;
// every other row was stored by the epilogue.  Fixed-size grid striding over (document, 1024-term slab) items, so
// that the launch costs nothing when the padded tiling was chosen (a grid proportional to B*V of immediately
// exiting CTAs costs ~0.1 ms at B=512).
__global__ void __launch_bounds__(256)
finalize_varlen_kernel(const VarlenPlan* __restrict__ plan, const int32_t* __restrict__ len,
                       const uint8_t* __restrict__ spans, const uint32_t* __restrict__ keys,
                       float* __restrict__ rep, int32_t* __restrict__ argmax, int B, int V) {
  if (plan->mode == 0) return;
  const int slabs = (V + 1023) / 1024;
  const long long items = static_cast<long long>(B) * slabs;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = static_cast<int>(item / slabs);
    const int L = len[b];
    if (L != 0 && spans[b] == 0) continue;
    const int v0 = static_cast<int>(item - static_cast<long long>(b) * slabs) * 1024;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = v0 + u * 256 + static_cast<int>(threadIdx.x);
      if (v >= V) continue;
      const size_t o = static_cast<size_t>(b) * V + v;
      if (L == 0) {
        rep[o] = 0.f;
        argmax[o] = 0;
      } else {
        const uint32_t k = keys[o];
        rep[o] = log1pf(bf16_bits_to_float(k >> 16));
        argmax[o] = static_cast<int32_t>(0xFFFFu - (k & 0xFFFFu));
      }
    }
  }
}

constexpr int VARLEN_SMEM_BYTES = SMEM_BYTES + 8 * 128 * 4;

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// SPLADE_B200_CTA_PAIR=1 selects the cta_group::2 kernel where the shape allows it (default: single-CTA kernel).
bool pair_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("SPLADE_B200_CTA_PAIR");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

}  // namespace

void head_fwd_plan(int B, int S, int* s_box, int* b_box, int* chunks_per_doc, int* n_tt, int* umma_n) {
  if (S <= BLOCK_T) {
    *s_box = S;
    *b_box = std::max(1, std::min(BLOCK_T / S, B));
    *chunks_per_doc = 1;
    *n_tt = (B + *b_box - 1) / *b_box;
  } else {
    *s_box = BLOCK_T;
    *b_box = 1;
    *chunks_per_doc = (S + BLOCK_T - 1) / BLOCK_T;
    *n_tt = B * *chunks_per_doc;
  }
  const int cols = *s_box * *b_box;
  *umma_n = std::max(16, (cols + 15) / 16 * 16);
}

// Shapes the packed path can take: (doc, position) of a column fit 16 bits each, offsets fit int32.
bool varlen_shape_ok(int B, int S) {
  return B >= 1 && B <= 32767 && S >= 1 && S <= 65535 && static_cast<size_t>(B) * S <= (static_cast<size_t>(1) << 30);
}

// rows of the packed token buffer: every document rounded up to 8 columns, a whole number of 256-token tiles
size_t varlen_rows(int B, int S) {
  const size_t rows = static_cast<size_t>(B) * ((static_cast<size_t>(S) + VL_ALIGN - 1) / VL_ALIGN * VL_ALIGN);
  return (rows + BLOCK_T - 1) / BLOCK_T * BLOCK_T;
}

size_t head_fwd_workspace_bytes(int B, int S, int H, int V, bool want_tokw) {
  if (B <= 0 || S <= 0) return 0;
  int s_box, b_box, cpd, n_tt, umma_n;
  head_fwd_plan(B, S, &s_box, &b_box, &cpd, &n_tt, &umma_n);
  size_t bytes = align256(static_cast<size_t>(n_tt) * 8 * sizeof(uint32_t)) +
                 align256(static_cast<size_t>(n_tt) * sizeof(int32_t));
  const bool varlen = varlen_shape_ok(B, S);
  if (S > BLOCK_T || varlen) bytes += align256(static_cast<size_t>(B) * V * sizeof(uint32_t));
  if (want_tokw) bytes += align256(static_cast<size_t>(B) * S * sizeof(uint32_t));
  if (varlen) {
    bytes += 256;                                                          // VarlenPlan
    bytes += align256(static_cast<size_t>(B) * sizeof(int32_t));           // len
    bytes += align256((static_cast<size_t>(B) + 1) * sizeof(int32_t));     // off
    bytes += align256(static_cast<size_t>(B));                             // spans
    bytes += align256(varlen_rows(B, S) * sizeof(uint32_t));               // meta
    bytes += align256(varlen_rows(B, S) * static_cast<size_t>(H) * 2);     // packed tokens (bf16)
  }
  return bytes;
}

// SPLADE_B200_PACK = auto (default) | 0 | 1: token layout policy of splade_head_fwd (see splade_head_fwd_mode).
int pack_mode_from_env() {
  static const int mode = [] {
    const char* e = std::getenv("SPLADE_B200_PACK");
    if (e == nullptr) return -1;
    if (e[0] == '0') return 0;
    if (e[0] == '1') return 1;
    return -1;
  }();
  return mode;
}

size_t varlen_rows_public(int B, int S) { return varlen_rows(B, S); }

size_t varlen_layout_bytes(int B, int S, int V) {
  if (!varlen_shape_ok(B, S)) return 0;
  return 256 + align256(static_cast<size_t>(B) * sizeof(int32_t)) +
         align256((static_cast<size_t>(B) + 1) * sizeof(int32_t)) + align256(static_cast<size_t>(B)) +
         align256(varlen_rows(B, S) * sizeof(uint32_t)) + align256(static_cast<size_t>(B) * V * sizeof(uint32_t));
}

// Packed layout of a batch from its attention mask alone (count -> plan -> column descriptors), pinned to the
// packed tiling.  Used by the native trunk, whose kernels all work on the packed rows.
int varlen_layout_build(const int32_t* mask, int B, int S, int V, void* workspace, size_t ws_bytes,
                        VarlenLayout* out, cudaStream_t stream) {
  if (!varlen_shape_ok(B, S))
    return set_error(kErrUnsupportedShape, "packed layout: needs B <= 32767, S <= 65535, B*S <= 2^30 (B=%d S=%d)", B, S);
  if (!workspace || ws_bytes < varlen_layout_bytes(B, S, V))
    return set_error(kErrWorkspace, "packed layout: workspace too small");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  out->plan = reinterpret_cast<VarlenPlan*>(ws);
  ws += 256;
  out->len = reinterpret_cast<int32_t*>(ws);
  ws += align256(static_cast<size_t>(B) * sizeof(int32_t));
  out->off = reinterpret_cast<int32_t*>(ws);
  ws += align256((static_cast<size_t>(B) + 1) * sizeof(int32_t));
  out->spans = ws;
  ws += align256(static_cast<size_t>(B));
  out->meta = reinterpret_cast<uint32_t*>(ws);
  ws += align256(varlen_rows(B, S) * sizeof(uint32_t));
  out->keys = reinterpret_cast<uint32_t*>(ws);
  out->rows = varlen_rows(B, S);
  const bool count_inline = static_cast<size_t>(B) * S <= 131072;
  if (!count_inline) {
    varlen_count_kernel<<<(B + 7) / 8, 256, 0, stream>>>(mask, B, S, out->len);
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  varlen_plan_kernel<<<1, 1024, 0, stream>>>(out->len, B, nullptr, 0, 0, 0, /*force_mode=*/1, out->off, out->spans,
                                             out->plan, nullptr, count_inline ? mask : nullptr, S);
  SPLADE_CUDA_OK(cudaGetLastError());
  varlen_pack_kernel<<<dim3((S + 255) / 256, B), 256, 0, stream>>>(nullptr, mask, out->len, out->off, out->spans,
                                                                    out->plan, S, 64, V, nullptr, out->meta,
                                                                    out->keys);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

// Fused head on rows that are ALREADY in the packed layout `lay` (hidden_packed [lay.rows, H] bf16).
int head_fwd_packed_launch(const void* hidden_packed, const VarlenLayout& lay, const void* W, const float* bias,
                           int B, int S, int H, int V, float* rep, int32_t* argmax, cudaStream_t stream) {
  if (H % BLOCK_K != 0)
    return set_error(kErrUnsupportedShape, "splade_head_fwd: hidden size %d is not a multiple of %d", H, BLOCK_K);
  CUtensorMap tmap_w, tmap_x;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(H), static_cast<uint64_t>(V)};
    const uint64_t strides[1] = {static_cast<uint64_t>(H) * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_V};
    const int rc = encode_tmap_bf16(&tmap_w, W, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(H), static_cast<uint64_t>(lay.rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(H) * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_T};
    const int rc = encode_tmap_bf16(&tmap_x, hidden_packed, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  VarlenParams q{};
  q.bias = bias;
  q.plan = lay.plan;
  q.meta = lay.meta;
  q.spans = lay.spans;
  q.len = lay.len;
  q.rep = rep;
  q.argmax = argmax;
  q.keys = lay.keys;
  q.tokw_bits = nullptr;
  q.B = B;
  q.S = S;
  q.H = H;
  q.V = V;
  q.n_vt = (V + BLOCK_V - 1) / BLOCK_V;
  auto kernel = head_fwd_varlen_kernel<false>;
  SPLADE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VARLEN_SMEM_BYTES));
  kernel<<<sm_count(), VL_THREADS, VARLEN_SMEM_BYTES, stream>>>(tmap_w, tmap_x, q);
  SPLADE_CUDA_OK(cudaGetLastError());
  const long long items = static_cast<long long>(B) * ((V + 1023) / 1024);
  const int grid = static_cast<int>(std::min<long long>(items, 8ll * sm_count()));
  finalize_varlen_kernel<<<grid, 256, 0, stream>>>(lay.plan, lay.len, lay.spans, lay.keys, rep, argmax, B, V);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

// The same decision as varlen_plan_kernel, for callers whose attention mask is still in HOST memory
// (splade_head_fwd_host): the tiling is pinned before anything is enqueued, so no gated no-op kernels are launched.
int head_fwd_pack_mode_host(const int32_t* mask_host, int B, int S) {
  const int env = pack_mode_from_env();
  if (env >= 0) return env;
  if (!varlen_shape_ok(B, S) || pair_enabled()) return 0;
  int s_box, b_box, cpd, n_tt, umma_n;
  head_fwd_plan(B, S, &s_box, &b_box, &cpd, &n_tt, &umma_n);
  const bool trim = b_box == 1 && s_box > 64;
  unsigned long long cost = 0, T = 0;
  for (int tt = 0; tt < n_tt; ++tt) {
    int rows = s_box * b_box;
    if (trim) {  // one document (chunk) per tile: rows up to the last unmasked position, rounded up to 64
      const int b = tt / cpd;
      const int s0 = (tt - b * cpd) * BLOCK_T;
      const int ns = std::min(BLOCK_T, S - s0);
      int last = -1;
      for (int s = ns - 1; s >= 0; --s) {
        if (mask_host[static_cast<size_t>(b) * S + s0 + s] != 0) {
          last = s;
          break;
        }
      }
      const int n16 = std::min(umma_n, std::max(16, (last + 1 + 15) / 16 * 16));
      rows = std::min(s_box, (n16 + 63) / 64 * 64);
    }
    cost += static_cast<unsigned long long>(BLOCK_V + rows);
  }
  for (int b = 0; b < B; ++b) {
    unsigned long long l = 0;
    for (int s = 0; s < S; ++s) l += mask_host[static_cast<size_t>(b) * S + s] != 0 ? 1 : 0;
    T += (l + VL_ALIGN - 1) / VL_ALIGN * VL_ALIGN;
  }
  const unsigned long long n_tp = (T + BLOCK_T - 1) / BLOCK_T;
  return (T > 0 && n_tp * (BLOCK_V + BLOCK_T) * 107ull < cost * 100ull) ? 1 : 0;
}

int head_fwd_launch(const void* hidden, const void* W, const float* bias, const int32_t* mask, int B, int S,
                    int H, int V, float* rep, int32_t* argmax, float* tokw, void* workspace, size_t ws_bytes,
                    int pack_mode, int32_t* plan_out, cudaStream_t stream) {
  if (B == 0) return kOk;
  if (!hidden || !W || !bias || !mask || !rep || !argmax)
    return set_error(kErrInvalidArg, "splade_head_fwd: null pointer argument");
  if (B < 0 || S < 1 || V < 1 || H < 1)
    return set_error(kErrInvalidArg, "splade_head_fwd: bad sizes B=%d S=%d H=%d V=%d", B, S, H, V);
  if (H % BLOCK_K != 0)
    return set_error(kErrUnsupportedShape, "splade_head_fwd: hidden size %d is not a multiple of %d", H, BLOCK_K);
  if (S > 65535)
    return set_error(kErrUnsupportedShape, "splade_head_fwd: sequence length %d exceeds 65535", S);
  if ((reinterpret_cast<uintptr_t>(hidden) | reinterpret_cast<uintptr_t>(W)) & 15u)
    return set_error(kErrInvalidArg, "splade_head_fwd: hidden and W must be 16-byte aligned");
  int rc = require_sm100("splade_head_fwd");
  if (rc != kOk) return rc;

  HeadFwdParams p{};
  p.bias = bias;
  p.mask = mask;
  p.rep = rep;
  p.argmax = argmax;
  p.B = B;
  p.S = S;
  p.H = H;
  p.V = V;
  head_fwd_plan(B, S, &p.s_box, &p.b_box, &p.chunks_per_doc, &p.n_tt, &p.umma_n);
  p.n_vt = (V + BLOCK_V - 1) / BLOCK_V;
  p.stage_tx_bytes = static_cast<uint32_t>(A_BYTES + p.s_box * p.b_box * BLOCK_K * 2);

  if (pack_mode < -1 || pack_mode > 1)
    return set_error(kErrInvalidArg, "splade_head_fwd: pack_mode must be -1 (auto), 0 (padded) or 1 (packed)");
  const size_t need = head_fwd_workspace_bytes(B, S, H, V, tokw != nullptr);
  if (need > 0 && (!workspace || ws_bytes < need))
    return set_error(kErrWorkspace, "splade_head_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint32_t* tile_masks = reinterpret_cast<uint32_t*>(ws);
  p.tile_masks = tile_masks;
  ws += align256(static_cast<size_t>(p.n_tt) * 8 * sizeof(uint32_t));
  int32_t* tile_n16 = reinterpret_cast<int32_t*>(ws);
  p.tile_n16 = tile_n16;
  ws += align256(static_cast<size_t>(p.n_tt) * sizeof(int32_t));
  const bool varlen_shape = varlen_shape_ok(B, S);
  uint32_t* keys = nullptr;
  if (p.chunks_per_doc > 1 || varlen_shape) {
    keys = reinterpret_cast<uint32_t*>(ws);
    const size_t n = static_cast<size_t>(B) * V * sizeof(uint32_t);
    if (p.chunks_per_doc > 1) {  // (the packed path clears only the rows of straddling documents, on the device)
      p.packed = keys;
      SPLADE_CUDA_OK(cudaMemsetAsync(keys, 0, n, stream));
    }
    ws += align256(n);
  }
  if (tokw) {
    p.tokw_bits = reinterpret_cast<uint32_t*>(ws);
    SPLADE_CUDA_OK(cudaMemsetAsync(p.tokw_bits, 0, static_cast<size_t>(B) * S * sizeof(uint32_t), stream));
    ws += align256(static_cast<size_t>(B) * S * sizeof(uint32_t));
  }
  VarlenPlan* plan = nullptr;
  int32_t* vl_len = nullptr;
  int32_t* vl_off = nullptr;
  uint8_t* vl_spans = nullptr;
  uint32_t* vl_meta = nullptr;
  void* vl_packed = nullptr;
  if (varlen_shape) {
    plan = reinterpret_cast<VarlenPlan*>(ws);
    ws += 256;
    vl_len = reinterpret_cast<int32_t*>(ws);
    ws += align256(static_cast<size_t>(B) * sizeof(int32_t));
    vl_off = reinterpret_cast<int32_t*>(ws);
    ws += align256((static_cast<size_t>(B) + 1) * sizeof(int32_t));
    vl_spans = reinterpret_cast<uint8_t*>(ws);
    ws += align256(static_cast<size_t>(B));
    vl_meta = reinterpret_cast<uint32_t*>(ws);
    ws += align256(varlen_rows(B, S) * sizeof(uint32_t));
    vl_packed = ws;
  }

  // CTA-pair (cta_group::2) path: needs the tile's token rows to split evenly between two CTAs.
  const int cols = p.s_box * p.b_box;
  bool use_pair = pair_enabled() && (sm_count() % 2 == 0) && (cols % 16 == 0) &&
                  ((p.b_box % 2 == 0) || (p.b_box == 1 && p.s_box % 16 == 0));
  // The packed path is considered unless the caller (or the CTA-pair experiment) pins the padded one.
  const bool varlen = varlen_shape && !use_pair && pack_mode != 0;
  if (plan_out != nullptr && !varlen) {
    SPLADE_CUDA_OK(cudaMemsetAsync(plan_out, 0, 4 * sizeof(int32_t), stream));
  }
  CUtensorMap tmap_w, tmap_h, tmap_trim[3];
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(H), static_cast<uint64_t>(V)};
    const uint64_t strides[1] = {static_cast<uint64_t>(H) * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_V};
    rc = encode_tmap_bf16(&tmap_w, W, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(H), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(H) * 2, static_cast<uint64_t>(S) * H * 2};
    uint32_t box[3] = {BLOCK_K, static_cast<uint32_t>(p.s_box), static_cast<uint32_t>(p.b_box)};
    if (use_pair) {
      p.pair_split_docs = (p.b_box % 2 == 0) ? 1 : 0;
      p.pair_half_rows = cols / 2;
      if (p.pair_split_docs) box[2] = static_cast<uint32_t>(p.b_box / 2);
      else box[1] = static_cast<uint32_t>(p.s_box / 2);
      p.umma_n = cols;
      p.stage_tx_bytes = static_cast<uint32_t>(2 * (P_A_BYTES + p.pair_half_rows * BLOCK_K * 2));
    }
    rc = encode_tmap_bf16(&tmap_h, hidden, 3, dims, strides, box);
    if (rc != kOk) return rc;
    // Trimmed boxes (64 / 128 / 192 token rows) for tiles whose document ends early; single-CTA kernel only.
    p.trim_rows = (!use_pair && p.b_box == 1 && p.s_box > 64) ? 1 : 0;
    for (int i = 0; i < 3; ++i) {
      uint32_t tbox[3] = {BLOCK_K, static_cast<uint32_t>(p.trim_rows ? std::min(p.s_box, 64 * (i + 1)) : p.s_box),
                          static_cast<uint32_t>(p.b_box)};
      rc = encode_tmap_bf16(&tmap_trim[i], hidden, 3, dims, strides, tbox);
      if (rc != kOk) return rc;
    }
  }

  build_tile_masks_kernel<<<p.n_tt, 256, 0, stream>>>(p, tile_masks, tile_n16);
  SPLADE_CUDA_OK(cudaGetLastError());

  CUtensorMap tmap_x;
  if (varlen) {
    const bool count_inline = static_cast<size_t>(B) * S <= 131072;  // small batch: the plan kernel counts itself
    if (!count_inline) {
      varlen_count_kernel<<<(B + 7) / 8, 256, 0, stream>>>(mask, B, S, vl_len);
      SPLADE_CUDA_OK(cudaGetLastError());
    }
    varlen_plan_kernel<<<1, 1024, 0, stream>>>(vl_len, B, tile_n16, p.n_tt, p.trim_rows,
                                               p.trim_rows ? p.s_box : cols, pack_mode, vl_off, vl_spans, plan,
                                               plan_out, count_inline ? mask : nullptr, S);
    SPLADE_CUDA_OK(cudaGetLastError());
    varlen_pack_kernel<<<dim3((S + 255) / 256, B), 256, 0, stream>>>(
        static_cast<const uint4*>(hidden), mask, vl_len, vl_off, vl_spans, plan, S, H, V,
        static_cast<uint4*>(vl_packed), vl_meta, keys);
    SPLADE_CUDA_OK(cudaGetLastError());
    const uint64_t dims[2] = {static_cast<uint64_t>(H), static_cast<uint64_t>(varlen_rows(B, S))};
    const uint64_t strides[1] = {static_cast<uint64_t>(H) * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_T};
    rc = encode_tmap_bf16(&tmap_x, vl_packed, 2, dims, strides, box);
    if (rc != kOk) return rc;
    p.plan_mode = &plan->mode;
  }

  if (varlen && pack_mode == 1) {
    // the caller pinned the packed path: the padded kernel is not even launched
  } else if (use_pair) {
    const int total_pairs = ((p.n_vt + 1) / 2) * p.n_tt;
    const int grid = 2 * std::min(sm_count() / 2, total_pairs);
    auto kernel = tokw ? head_fwd_pair_kernel<true> : head_fwd_pair_kernel<false>;
    SPLADE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    kernel<<<grid, NUM_THREADS, P_SMEM_BYTES, stream>>>(tmap_w, tmap_h, p);
  } else {
    const int total_items = p.n_vt * p.n_tt;
    const int grid = std::min(sm_count(), total_items);
    auto kernel = tokw ? head_fwd_kernel<true> : head_fwd_kernel<false>;
    SPLADE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmap_w, tmap_h, tmap_trim[0], tmap_trim[1], tmap_trim[2], p);
  }
  SPLADE_CUDA_OK(cudaGetLastError());

  if (p.chunks_per_doc > 1 && !(varlen && pack_mode == 1)) {
    const size_t n = static_cast<size_t>(B) * V;
    finalize_packed_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(p.packed, rep, argmax, n,
                                                                                       p.plan_mode);
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  if (varlen) {
    VarlenParams q{};
    q.bias = bias;
    q.plan = plan;
    q.meta = vl_meta;
    q.spans = vl_spans;
    q.len = vl_len;
    q.rep = rep;
    q.argmax = argmax;
    q.keys = keys;
    q.tokw_bits = p.tokw_bits;
    q.B = B;
    q.S = S;
    q.H = H;
    q.V = V;
    q.n_vt = p.n_vt;
    // persistent over (vocab tile, packed token tile); the tile count is only known on the device
    auto kernel = tokw ? head_fwd_varlen_kernel<true> : head_fwd_varlen_kernel<false>;
    SPLADE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VARLEN_SMEM_BYTES));
    kernel<<<sm_count(), VL_THREADS, VARLEN_SMEM_BYTES, stream>>>(tmap_w, tmap_x, q);
    SPLADE_CUDA_OK(cudaGetLastError());
    {
      const long long items = static_cast<long long>(B) * ((V + 1023) / 1024);
      const int grid = static_cast<int>(std::min<long long>(items, 8ll * sm_count()));
      finalize_varlen_kernel<<<grid, 256, 0, stream>>>(plan, vl_len, vl_spans, keys, rep, argmax, B, V);
    }
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  if (tokw) {
    const size_t n = static_cast<size_t>(B) * S;
    finalize_tokw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(p.tokw_bits, tokw, n);
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  return kOk;
}

}  // namespace splade
