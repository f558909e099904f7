// tcgen05 GEMM of the variable-length ModernBERT trunk:
//   C[t, n] = sum_k A[t, k] * W[n, k]      A = packed token activations [T, K] bf16, W = nn.Linear weight [N, K] bf16
// with the layer's element-wise tail fused into the epilogue, so that no intermediate ever makes an extra HBM
// round trip.  Reference arithmetic (transformers modeling_modernbert.py, under torch.autocast(bf16) as the
// trainer runs it, src/train/cli/train_v33_ddp.py:339): every nn.Linear returns bf16, the residual stream is fp32.
//
//   EPI_BF16       plain bf16 output                                                        -> bf16 [T, N]
//   EPI_QKV_ROPE   attn.Wqkv (:278) + apply_rotary_pos_emb (:205-228) on the q and k heads -> bf16 [T, 3H]
//   EPI_RESID      attn.Wo (:309) / mlp.Wo (:91) + residual add (:341-342)                -> fp32 x [T, H] (in place)
//   EPI_GEGLU      mlp.Wi (:90) + act(input) * gate (:91)                                  -> bf16 [T, I]
//   EPI_GELU       head.dense + GELU (:502)                                                -> bf16 [T, H]
//
// Orientation: UMMA M = 128 tokens (TMEM lanes), UMMA N = 256 output features (TMEM columns), K in 64-element
// (128-byte swizzle) blocks through a 4-stage TMA ring, two 128x256 fp32 accumulators so the epilogue of one tile
// overlaps the MMAs of the next.  An epilogue thread owns ONE token and walks its output features, so RoPE pairs
// (d, d+32 of a head) and GeGLU pairs (input n, gate n) are in the same thread's registers; results leave through
// per-warp shared-memory tiles and TMA bulk stores / reduce-adds.  The number of
// token rows T is only known on the device (it comes from the attention mask): the kernel is persistent and reads
// it from the packed-layout plan.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace splade {
namespace {

constexpr int G_BM = 128;
constexpr int G_BN = 256;
constexpr int G_BK = 64;
constexpr int G_UK = 16;
constexpr int G_STAGES = 4;
constexpr int G_A_BYTES = G_BM * G_BK * 2;
constexpr int G_B_BYTES = G_BN * G_BK * 2;
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_THREADS = 384;  // producer, MMA, TMEM alloc, idle, 2 x 4 epilogue warps
constexpr int G_TMEM_COLS = 512;
constexpr int G_EPI_WARPS = 8;
constexpr int G_EPI_STAGE_BYTES = 4096;  // per epilogue warp: 32 rows x 128 B (64 bf16 or 32 fp32), 128B-swizzled
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE_BYTES + G_EPI_WARPS * G_EPI_STAGE_BYTES + 256 + 1024;

struct GemmParams {
  const int32_t* t_rows;  // device: number of valid token rows (packed columns) T
  int N;                  // output features (GeGLU: the intermediate size I; W has 2*I rows)
  int K;
  void* out;              // bf16 [T, ldo] or fp32 x [T, ldo] (EPI_RESID, read-modify-write)
  int ldo;
  const uint32_t* meta;   // EPI_QKV_ROPE: packed column descriptors (position in the low 16 bits)
  int rope_cols;          // EPI_QKV_ROPE: output features below this are q / k heads (2*H)
  float inv_freq[32];     // EPI_QKV_ROPE: theta^(-2i/64)
  int debug;              // timing experiments only (SPLADE_B200_GEMM_DEBUG): 1 = no global stores, 2 = no epilogue
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// An epilogue thread owns one token ROW of the accumulator (a TMEM lane); written straight to global memory that is
// 32 different rows per store instruction (measured: qkv 60 us with such stores, 39.7 us without any).  So every
// 32-row block leaves through a per-warp shared-memory tile and ONE TMA bulk tensor store (bf16 outputs) or TMA
// bulk REDUCE-ADD (the fp32 residual stream: x += tile, performed at the L2 - no read-modify-write through the
// SM, and exactly one add per element, i.e. the same single rounding as a load-add-store).  The tile is written in
// the 128-byte swizzle TMA expects: 16-byte chunk c of row r sits at chunk position c ^ (r & 7), which also makes
// the thread-per-row writes bank-conflict free.  Measured at [16384,768]: qkv 60 -> 50 us, attn.Wo 39 -> 27 us,
// mlp.Wo 42 -> 32 us (per-lane REDG / float4 read-modify-write variants: 39-58 us).
__device__ __forceinline__ void stage_chunk16(uint8_t* stage, uint32_t lane, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(stage + lane * 128u + ((static_cast<uint32_t>(chunk) ^ (lane & 7u)) << 4)) = v;
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 32 fp32 values of this thread's row -> chunks [chunk0, chunk0 + 4) of its staged bf16 row
__device__ __forceinline__ void stage_bf16x32(uint8_t* stage, uint32_t lane, int chunk0, const float (&y)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v;
    v.x = cvt_bf16x2(y[8 * i + 1], y[8 * i + 0]);
    v.y = cvt_bf16x2(y[8 * i + 3], y[8 * i + 2]);
    v.z = cvt_bf16x2(y[8 * i + 5], y[8 * i + 4]);
    v.w = cvt_bf16x2(y[8 * i + 7], y[8 * i + 6]);
    stage_chunk16(stage, lane, chunk0 + i, v);
  }
}

template <int kEpi>
__global__ void __launch_bounds__(G_THREADS, 1)
trunk_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ GemmParams p) {
  const int T = __ldg(p.t_rows);
  if (T <= 0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* epi_stage = smem + G_STAGES * G_STAGE_BYTES;  // 1024-byte aligned: G_EPI_WARPS x 4 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + G_EPI_WARPS * G_EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G_STAGES;
  uint64_t* tfull_bar = empty_bar + G_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, G_TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tiles: (token tile, feature tile), feature tile fastest so that consecutive tiles share the activation rows
  constexpr int kOutPerTile = (kEpi == EPI_GEGLU) ? G_BN / 2 : G_BN;  // output features produced per tile
  const int m_tiles = (T + G_BM - 1) / G_BM;
  const int n_tiles = (p.N + kOutPerTile - 1) / kOutPerTile;
  const int total = m_tiles * n_tiles;
  const int num_kb = p.K / G_BK;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int mt = tile / n_tiles;
        const int nt = tile - mt * n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* a_dst = smem + stage * G_STAGE_BYTES;
          uint8_t* b_dst = a_dst + G_A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], G_STAGE_BYTES);
          tma_load_2d(a_dst, &tmap_a, &full_bar[stage], kb * G_BK, mt * G_BM);
          if (kEpi == EPI_GEGLU) {
            // 128 "input" rows and the 128 matching "gate" rows of Wi (nn.Linear(H, 2I): chunk(2, dim=-1))
            tma_load_2d(b_dst, &tmap_w, &full_bar[stage], kb * G_BK, nt * (G_BN / 2));
            tma_load_2d(b_dst + G_B_BYTES / 2, &tmap_w, &full_bar[stage], kb * G_BK, p.N + nt * (G_BN / 2));
          } else {
            tma_load_2d(b_dst, &tmap_w, &full_bar[stage], kb * G_BK, nt * G_BN);
          }
          if (++stage == G_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(G_BM, G_BN);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * G_STAGE_BYTES);
          const uint32_t b_addr = a_addr + G_A_BYTES;
#pragma unroll
          for (int k = 0; k < G_BK / G_UK; ++k) {
            const uint64_t a_desc = make_smem_desc_sw128(a_addr + k * G_UK * 2);
            const uint64_t b_desc = make_smem_desc_sw128(b_addr + k * G_UK * 2);
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == G_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    const uint32_t q = warp & 3u;           // TMEM lane quarter = 32 token rows
    const uint32_t grp = (warp - 4u) >> 2;  // which half of the tile's output features
    uint8_t* stage = epi_stage + (warp - 4u) * G_EPI_STAGE_BYTES;
    uint32_t it = 0;
    // hand the staged 32-row block to the TMA engine (one elected lane), after making the generic-proxy writes
    // visible to the async proxy; the previous block's read of the stage buffer is awaited before it is rewritten
    auto stage_begin = [&]() {
      if (lane == 0) bulk_wait_read();
      __syncwarp();
    };
    auto stage_commit = [&](int col, int row, bool reduce) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && p.debug != 1) {
        if (reduce) tma_reduce_add_2d(&tmap_o, stage, col, row);
        else tma_store_2d(&tmap_o, stage, col, row);
        bulk_commit();
      }
    };
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int mt = tile / n_tiles;
      const int nt = tile - mt * n_tiles;
      const int row0 = mt * G_BM + static_cast<int>(q * 32u);  // first token row of this warp
      // EPI_QKV_ROPE: (cos, sin) of this thread's token for the 32 frequencies, once per tile of q / k heads, while
      // the MMAs of the tile still run (fp32 sincosf as ModernBertRotaryEmbedding.forward :158-172 computes them)
      float rc[kEpi == EPI_QKV_ROPE ? 32 : 1], rs[kEpi == EPI_QKV_ROPE ? 32 : 1];
      if (kEpi == EPI_QKV_ROPE) {
        if (nt * G_BN + static_cast<int>(grp) * 128 < p.rope_cols) {  // warp-uniform
          const int row = min(row0 + static_cast<int>(lane), T - 1);
          const float pos = static_cast<float>(__ldg(&p.meta[row]) & 0xFFFFu);
#pragma unroll
          for (int j = 0; j < 32; ++j) sincosf(pos * p.inv_freq[j], &rs[j], &rc[j]);
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_lane = tmem_base + ((q * 32u) << 16) + acc * G_BN;
      if (p.debug == 2) {
        tcgen05_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        continue;
      }

      if (kEpi == EPI_GEGLU) {
        // this group: input columns [64 grp, 64 grp + 64) and the gate columns 128 above them -> one 64-wide block
        const int n0 = nt * (G_BN / 2) + static_cast<int>(grp) * 64;
        if (n0 < p.N) {  // warp-uniform
          stage_begin();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const int col = static_cast<int>(grp) * 64 + c * 32;
            uint32_t a[32], g[32];
            tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(col), a);
            tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(G_BN / 2 + col), g);
            tmem_ld_wait();
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float in = bf16_round(__uint_as_float(a[j]));
              const float gate = bf16_round(__uint_as_float(g[j]));
              y[j] = bf16_round(gelu_erf(in)) * gate;
            }
            stage_bf16x32(stage, lane, 4 * c, y);
          }
          stage_commit(n0, row0, false);
        }
      } else if (kEpi == EPI_RESID) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {  // fp32: 32 features = one 128-byte row per block
          const int col = static_cast<int>(grp) * 128 + c * 32;
          const int n0 = nt * G_BN + col;
          uint32_t a[32];
          tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(col), a);
          tmem_ld_wait();
          if (n0 < p.N) {  // warp-uniform
            stage_begin();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 v;  // the Linear's bf16 output, as fp32 addends
              v.x = __float_as_uint(bf16_round(__uint_as_float(a[4 * i + 0])));
              v.y = __float_as_uint(bf16_round(__uint_as_float(a[4 * i + 1])));
              v.z = __float_as_uint(bf16_round(__uint_as_float(a[4 * i + 2])));
              v.w = __float_as_uint(bf16_round(__uint_as_float(a[4 * i + 3])));
              stage_chunk16(stage, lane, i, v);
            }
            stage_commit(n0, row0, true);
          }
        }
      } else if (kEpi == EPI_QKV_ROPE) {
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {  // two 64-wide heads per group
          const int col = static_cast<int>(grp) * 128 + hb * 64;
          const int n0 = nt * G_BN + col;
          uint32_t lo[32], hi[32];
          tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(col), lo);
          tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(col + 32), hi);
          tmem_ld_wait();
          if (n0 < p.N) {  // warp-uniform
            float y1[32], y2[32];
            if (n0 < p.rope_cols) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x1 = bf16_round(__uint_as_float(lo[j]));  // the Linear's bf16 output, then .float()
                const float x2 = bf16_round(__uint_as_float(hi[j]));
                y1[j] = x1 * rc[j] - x2 * rs[j];                      // q*cos + rotate_half(q)*sin
                y2[j] = x2 * rc[j] + x1 * rs[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                y1[j] = __uint_as_float(lo[j]);
                y2[j] = __uint_as_float(hi[j]);
              }
            }
            stage_begin();
            stage_bf16x32(stage, lane, 0, y1);
            stage_bf16x32(stage, lane, 4, y2);
            stage_commit(n0, row0, false);
          }
        }
      } else {
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {  // two 64-feature blocks per group
          const int n0 = nt * G_BN + static_cast<int>(grp) * 128 + hb * 64;
          if (n0 < p.N) stage_begin();  // warp-uniform
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const int col = static_cast<int>(grp) * 128 + hb * 64 + c * 32;
            uint32_t a[32];
            tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(col), a);
            tmem_ld_wait();
            if (n0 < p.N) {
              float y[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(a[j]);
                y[j] = (kEpi == EPI_GELU) ? gelu_erf(bf16_round(v)) : v;
              }
              stage_bf16x32(stage, lane, 4 * c, y);
            }
          }
          if (n0 < p.N) stage_commit(n0, row0, false);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
    if (lane == 0) bulk_wait_all();  // all bulk stores have completed before the CTA (and its shared memory) goes away
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, G_TMEM_COLS);
  }
}

template <int kEpi>
int launch_one(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w, const CUtensorMap& tmap_o, const GemmParams& p,
               int grid, cudaStream_t stream) {
  auto kernel = trunk_gemm_kernel<kEpi>;
  SPLADE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
  kernel<<<grid, G_THREADS, G_SMEM_BYTES, stream>>>(tmap_a, tmap_w, tmap_o, p);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // namespace

// A [rows_cap, K] bf16 (rows_cap >= 128), W [w_rows, K] bf16, out [rows_cap, ldo] (bf16, or fp32 for EPI_RESID).
// EPI_QKV_ROPE: meta[rows] (position in the low 16 bits), rope_theta, rope_cols = 2*H.
int trunk_gemm_launch(int epi, const void* A, int rows_cap, const void* W, int N, int K, const int32_t* t_rows,
                      void* out, int ldo, const uint32_t* meta, float rope_theta, int rope_cols,
                      cudaStream_t stream) {
  if (K % G_BK != 0 || K < G_BK)
    return set_error(kErrUnsupportedShape, "trunk gemm: K=%d is not a multiple of %d", K, G_BK);
  if (N % 128 != 0 || (epi != EPI_GEGLU && N < G_BN))
    return set_error(kErrUnsupportedShape, "trunk gemm: N=%d must be a multiple of 128 (and >= 256)", N);
  if (epi == EPI_QKV_ROPE && (rope_cols % 128 != 0 || !meta || !(rope_theta > 0.f)))
    return set_error(kErrInvalidArg, "trunk gemm: RoPE needs positions, theta > 0 and 64-wide heads");
  if (rows_cap < G_BM) return set_error(kErrUnsupportedShape, "trunk gemm: fewer than %d activation rows", G_BM);
  if (ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15u))
    return set_error(kErrInvalidArg, "trunk gemm: output rows must be 16-byte aligned");
  const int w_rows = epi == EPI_GEGLU ? 2 * N : N;
  CUtensorMap tmap_a, tmap_w, tmap_o;
  {  // output: 32-row blocks of 128 bytes (64 bf16 features, or 32 fp32 features of the residual stream)
    const bool f32 = epi == EPI_RESID;
    const uint64_t dims[2] = {static_cast<uint64_t>(ldo), static_cast<uint64_t>(rows_cap)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldo) * (f32 ? 4 : 2)};
    const uint32_t box[2] = {f32 ? 32u : 64u, 32u};
    const int rc = f32 ? encode_tmap_f32(&tmap_o, out, 2, dims, strides, box)
                       : encode_tmap_bf16(&tmap_o, out, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows_cap)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {G_BK, G_BM};
    const int rc = encode_tmap_bf16(&tmap_a, A, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(w_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {G_BK, static_cast<uint32_t>(epi == EPI_GEGLU ? G_BN / 2 : G_BN)};
    const int rc = encode_tmap_bf16(&tmap_w, W, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  GemmParams p{};
  p.t_rows = t_rows;
  p.N = N;
  p.K = K;
  p.out = out;
  p.ldo = ldo;
  p.meta = meta;
  p.rope_cols = rope_cols;
  if (epi == EPI_QKV_ROPE)
    for (int i = 0; i < 32; ++i) p.inv_freq[i] = 1.0f / powf(rope_theta, static_cast<float>(2 * i) / 64.0f);
  {
    static const int dbg = [] {
      const char* e = std::getenv("SPLADE_B200_GEMM_DEBUG");
      return e ? std::atoi(e) : 0;
    }();
    p.debug = dbg;
  }
  const int grid = sm_count();
  switch (epi) {
    case EPI_BF16:
      return launch_one<EPI_BF16>(tmap_a, tmap_w, tmap_o, p, grid, stream);
    case EPI_GELU:
      return launch_one<EPI_GELU>(tmap_a, tmap_w, tmap_o, p, grid, stream);
    case EPI_RESID:
      return launch_one<EPI_RESID>(tmap_a, tmap_w, tmap_o, p, grid, stream);
    case EPI_GEGLU:
      return launch_one<EPI_GEGLU>(tmap_a, tmap_w, tmap_o, p, grid, stream);
    case EPI_QKV_ROPE:
      return launch_one<EPI_QKV_ROPE>(tmap_a, tmap_w, tmap_o, p, grid, stream);
    default:
      return set_error(kErrInvalidArg, "trunk gemm: unknown epilogue %d", epi);
  }
}

}  // namespace splade
