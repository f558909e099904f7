// tcgen05 GEMM of the variable-length ModernBERT trunk:
//   C[t, n] = sum_k A[t, k] * W[n, k]      A = packed token activations [T, K] bf16, W = nn.Linear weight [N, K] bf16
// with the layer's element-wise tail fused into the epilogue, so that no intermediate ever makes an extra HBM
// round trip.  Reference arithmetic (transformers modeling_modernbert.py, under torch.autocast(bf16) as the
// trainer runs it, src/train/cli/train_v33_ddp.py:339): every nn.Linear returns bf16, the residual stream is fp32.
//
//   EPI_BF16       attn.Wqkv (:278)  (RoPE is applied by the attention kernel's loader)      -> bf16 [T, 3H]
//   EPI_RESID      attn.Wo (:309) / mlp.Wo (:91) + residual add (:341-342)                -> fp32 x [T, H] (in place)
//   EPI_GEGLU      mlp.Wi (:90) + act(input) * gate (:91)                                  -> bf16 [T, I]
//   EPI_GELU       head.dense + GELU (:502)                                                -> bf16 [T, H]
//
// Orientation: UMMA M = 128 tokens (TMEM lanes), UMMA N = 256 output features (TMEM columns), K in 64-element
// (128-byte swizzle) blocks through a 4-stage TMA ring, two 128x256 fp32 accumulators so the epilogue of one tile
// overlaps the MMAs of the next.  An epilogue thread owns ONE token and walks its output features, so the GeGLU
// pair (input n, gate n) is in the same thread's registers; results leave through a per-warp shared-memory
// transpose so that global accesses are row-coalesced.  The number of
// token rows T is only known on the device (it comes from the attention mask): the kernel is persistent and reads
// it from the packed-layout plan.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

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
constexpr int G_STAGE_LD = 33;  // fp32 words per staged row (conflict-free row writes and column reads)
constexpr int G_EPI_STAGE_BYTES = 32 * G_STAGE_LD * 4;
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE_BYTES + 256 + G_EPI_WARPS * G_EPI_STAGE_BYTES + 1024;

struct GemmParams {
  const int32_t* t_rows;  // device: number of valid token rows (packed columns) T
  int N;                  // output features (GeGLU: the intermediate size I; W has 2*I rows)
  int K;
  void* out;              // bf16 [T, ldo] or fp32 x [T, ldo] (EPI_RESID, read-modify-write)
  int ldo;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// An epilogue thread owns one token ROW of the accumulator, but global memory wants a warp to touch one row at a
// time.  Every 32 x 32 block therefore goes through a per-warp shared-memory stage: thread `lane` writes its 32
// values as row `lane` (bank = lane + j: conflict-free), then the warp reads the block back one row per
// instruction (bank = row + lane: conflict-free) and issues fully coalesced global accesses.
__device__ __forceinline__ void stage_rows(float* stage, uint32_t lane, const float (&y)[32]) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) stage[lane * G_STAGE_LD + j] = y[j];
  __syncwarp();
}

// x[row][n0 + lane] += bf16(acc): the Linear's bf16 output added to the fp32 residual stream, as fire-and-forget
// reductions (REDG.ADD.F32, one coalesced 128-byte line per instruction).  Every element receives exactly ONE add
// per GEMM, so the result is the same single rounding as a load-add-store.  Measured alternatives at
// [16384,768]x[768,768]: thread-per-row float4 load/add/store 58 us, this 39 us, coalesced load/add/store with the
// loads prefetched one / two blocks ahead 52 / 47 us (the load latency stays exposed at 4 blocks per tile).
__device__ __forceinline__ void emit_resid_block(const float* stage, uint32_t lane, float* x, int ldo, int row0,
                                                 int n0, int T) {
  float* px = x + static_cast<size_t>(row0) * ldo + n0 + lane;
#pragma unroll 8
  for (int rr = 0; rr < 32; ++rr) {
    if (row0 + rr < T) atomicAdd(px + static_cast<size_t>(rr) * ldo, bf16_round(stage[rr * G_STAGE_LD + lane]));
  }
}

// direct form for bf16 outputs: the thread's 32 features of its own row are 64 contiguous bytes (two full sectors)
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&y)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v;
    v.x = cvt_bf16x2(y[8 * i + 1], y[8 * i + 0]);
    v.y = cvt_bf16x2(y[8 * i + 3], y[8 * i + 2]);
    v.z = cvt_bf16x2(y[8 * i + 5], y[8 * i + 4]);
    v.w = cvt_bf16x2(y[8 * i + 7], y[8 * i + 6]);
    d4[i] = v;
  }
}

template <int kEpi>
__global__ void __launch_bounds__(G_THREADS, 1)
trunk_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const GemmParams p) {
  const int T = __ldg(p.t_rows);
  if (T <= 0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G_STAGES;
  uint64_t* tfull_bar = empty_bar + G_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
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
    float* stage = reinterpret_cast<float*>(smem + G_STAGES * G_STAGE_BYTES + 256) +
                   (warp - 4u) * (G_EPI_STAGE_BYTES / 4);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int mt = tile / n_tiles;
      const int nt = tile - mt * n_tiles;
      const int row0 = mt * G_BM + static_cast<int>(q * 32u);  // first token row of this warp
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_lane = tmem_base + ((q * 32u) << 16) + acc * G_BN;

      if (kEpi == EPI_GEGLU) {
        // this group: input columns [64 grp, 64 grp + 64) and the gate columns 128 above them
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int col = static_cast<int>(grp) * 64 + c * 32;
          uint32_t a[32], g[32];
          tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(col), a);
          tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(G_BN / 2 + col), g);
          tmem_ld_wait();
          const int n0 = nt * (G_BN / 2) + col;
          if (n0 < p.N) {  // warp-uniform
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float in = bf16_round(__uint_as_float(a[j]));
              const float gate = bf16_round(__uint_as_float(g[j]));
              y[j] = bf16_round(gelu_erf(in)) * gate;
            }
            if (row0 + static_cast<int>(lane) < T)
              store_bf16x32(static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row0 + lane) * p.ldo + n0, y);
          }
        }
      } else if (kEpi == EPI_RESID) {
        float* x = static_cast<float*>(p.out);
        const int nbase = nt * G_BN + static_cast<int>(grp) * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int n0 = nbase + c * 32;
          uint32_t a[32];
          tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(grp * 128u + c * 32), a);
          tmem_ld_wait();
          if (n0 < p.N) {  // warp-uniform
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(a[j]);
            stage_rows(stage, lane, y);
            emit_resid_block(stage, lane, x, p.ldo, row0, n0, T);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = static_cast<int>(grp) * 128 + c * 32;
          uint32_t a[32];
          tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(col), a);
          tmem_ld_wait();
          const int n0 = nt * G_BN + col;
          if (n0 < p.N && row0 + static_cast<int>(lane) < T) {
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(a[j]);
              y[j] = (kEpi == EPI_GELU) ? gelu_erf(bf16_round(v)) : v;
            }
            store_bf16x32(static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row0 + lane) * p.ldo + n0, y);
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, G_TMEM_COLS);
  }
}

template <int kEpi>
int launch_one(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w, const GemmParams& p, int grid,
               cudaStream_t stream) {
  auto kernel = trunk_gemm_kernel<kEpi>;
  SPLADE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
  kernel<<<grid, G_THREADS, G_SMEM_BYTES, stream>>>(tmap_a, tmap_w, p);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // namespace

// A [rows_cap, K] bf16 (rows_cap >= 128, a whole number of 256-row tiles in the trunk), W [w_rows, K] bf16.
int trunk_gemm_launch(int epi, const void* A, int rows_cap, const void* W, int N, int K, const int32_t* t_rows,
                      void* out, int ldo, cudaStream_t stream) {
  if (K % G_BK != 0 || K < G_BK)
    return set_error(kErrUnsupportedShape, "trunk gemm: K=%d is not a multiple of %d", K, G_BK);
  if (N % 128 != 0 || (epi != EPI_GEGLU && N < G_BN))
    return set_error(kErrUnsupportedShape, "trunk gemm: N=%d must be a multiple of 128 (and >= 256)", N);
  if (rows_cap < G_BM) return set_error(kErrUnsupportedShape, "trunk gemm: fewer than %d activation rows", G_BM);
  const int w_rows = epi == EPI_GEGLU ? 2 * N : N;
  CUtensorMap tmap_a, tmap_w;
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
  const int grid = sm_count();
  switch (epi) {
    case EPI_BF16:
      return launch_one<EPI_BF16>(tmap_a, tmap_w, p, grid, stream);
    case EPI_GELU:
      return launch_one<EPI_GELU>(tmap_a, tmap_w, p, grid, stream);
    case EPI_RESID:
      return launch_one<EPI_RESID>(tmap_a, tmap_w, p, grid, stream);
    case EPI_GEGLU:
      return launch_one<EPI_GEGLU>(tmap_a, tmap_w, p, grid, stream);
    default:
      return set_error(kErrInvalidArg, "trunk gemm: unknown epilogue %d", epi);
  }
}

}  // namespace splade
