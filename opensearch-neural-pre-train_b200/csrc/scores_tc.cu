// Tensor-core q.d score GEMM for InfoNCE (autocast mode): raw[i,j] = sum_v bf16(q[i,v]) * bf16(docs[j,v]),
// fp32 accumulate — what torch.mm(anchor, positive.t()) computes under torch.autocast(bf16)
// (src/model/losses.py:155, :243).
//
// Shape: M = Bq (64) queries, N = in-batch documents (64 ... W*64 with gathered negatives), K = V = 50 000:
// a short-and-fat GEMM whose cost is reading the operands once.  It is split over K so that every SM
// streams a different slice of the vocabulary:
//   1. pack_bf16_kernel: fp32 -> bf16 copies (the cast autocast performs), rows padded to a multiple of 8;
//   2. scores_tc_kernel: grid (k-splits, m-tiles, n-tiles); each CTA accumulates its K slice with
//      tcgen05.mma (128 x N x 16, operands TMA-staged with 128 B swizzle, 4-stage mbarrier ring,
//      accumulator in TMEM) and writes an fp32 partial tile;
//   3. the fixed-order reduction over k-splits (scores_reduce in loss.cu) -> deterministic result.
// The bf16 rounding of the OUTPUT (and the division by the temperature) is applied by the scalar stage.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace splade {
namespace {

constexpr int TM = 128;       // UMMA M
constexpr int TN_MAX = 256;   // max UMMA N
constexpr int BK = 64;
constexpr int UK = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * BK * 2;
constexpr int B_BYTES = TN_MAX * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
constexpr int TMEM_COLS = 256;

// fp32 [rows, V] -> bf16 [rows, Vp] (Vp = V rounded up to 8, zero padded); grid (column chunks, rows).
__global__ void __launch_bounds__(256)
pack_bf16_kernel(const float* __restrict__ src, int V, int Vp, __nv_bfloat16* __restrict__ dst) {
  const float* s = src + static_cast<size_t>(blockIdx.y) * V;
  __nv_bfloat16* d = dst + static_cast<size_t>(blockIdx.y) * Vp;
  const int v0 = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (v0 >= Vp) return;
  float x[8];
  if (v0 + 8 <= V && (reinterpret_cast<uintptr_t>(s + v0) & 15u) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(s + v0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(s + v0 + 4));
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = (v0 + e < V) ? s[v0 + e] : 0.f;
  }
  uint4 o;
  o.x = cvt_bf16x2(x[1], x[0]);
  o.y = cvt_bf16x2(x[3], x[2]);
  o.z = cvt_bf16x2(x[5], x[4]);
  o.w = cvt_bf16x2(x[7], x[6]);
  *reinterpret_cast<uint4*>(d + v0) = o;  // Vp % 8 == 0 and the workspace is 256-byte aligned
}

struct ScoresTcParams {
  float* partial;  // [ksplit][Bq][N]
  int Bq, N;
  int n_tile;      // UMMA N of this launch (multiple of 16)
  int kb_total;    // number of 64-wide k blocks
  int kb_per_cta;
  uint32_t stage_tx_bytes;
};

__global__ void __launch_bounds__(128, 1)
scores_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                 const ScoresTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done_bar + 1);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int kb0 = blockIdx.x * p.kb_per_cta;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_cta);
  const int m0 = blockIdx.y * TM;
  const int n0 = blockIdx.z * p.n_tile;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0 && lane == 0) {
    uint32_t stage = 0, phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* a_dst = smem + stage * STAGE_BYTES;
      mbar_arrive_expect_tx(&full_bar[stage], p.stage_tx_bytes);
      tma_load_2d(a_dst, &tmap_q, &full_bar[stage], kb * BK, m0);
      tma_load_2d(a_dst + A_BYTES, &tmap_d, &full_bar[stage], kb * BK, n0);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(TM, static_cast<uint32_t>(p.n_tile));
    uint32_t stage = 0, phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
      const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
      for (int k = 0; k < BK / UK; ++k) {
        umma_bf16(tmem_base, make_smem_desc_sw128(a_addr + k * UK * 2), make_smem_desc_sw128(b_addr + k * UK * 2),
                  idesc, (kb > kb0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
    umma_commit(done_bar);
  }
  __syncwarp();

  // ---- epilogue: all four warps drain their TMEM lane quarter into the fp32 partial tile
  if (kb1 > kb0) {
    mbar_wait(done_bar, 0);
    tcgen05_fence_after();
  }
  __syncwarp();
  const int i = m0 + static_cast<int>(warp * 32u + lane);
  float* out_row = p.partial + (static_cast<size_t>(blockIdx.x) * p.Bq + (i < p.Bq ? i : 0)) * p.N;
  for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
    uint32_t r[32];
    if (kb1 > kb0) {
      tmem_ld_32x32b_x32(tmem_base + ((warp * 32u) << 16) + static_cast<uint32_t>(c0), r);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
    if (i < p.Bq) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = n0 + c0 + j;
        if (c0 + j < p.n_tile && col < p.N) out_row[col] = __uint_as_float(r[j]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

void plan(int Bq, int N, int V, int* Vp, int* n_tile, int* m_tiles, int* n_tiles, int* kb_total, int* kb_per_cta,
          int* ksplit) {
  *Vp = (V + 7) / 8 * 8;
  const int n16 = (N + 15) / 16 * 16;
  *n_tile = std::min(TN_MAX, n16);
  *m_tiles = (Bq + TM - 1) / TM;
  *n_tiles = (N + *n_tile - 1) / *n_tile;
  *kb_total = (*Vp + BK - 1) / BK;
  const int want = std::max(1, sm_count() / std::max(1, *m_tiles * *n_tiles));
  *kb_per_cta = std::max(1, (*kb_total + want - 1) / want);
  *ksplit = (*kb_total + *kb_per_cta - 1) / *kb_per_cta;
}

}  // namespace

size_t scores_tc_workspace_bytes(int Bq, int N, int V) {
  int Vp, n_tile, m_tiles, n_tiles, kb_total, kb_per_cta, ksplit;
  plan(Bq, N, V, &Vp, &n_tile, &m_tiles, &n_tiles, &kb_total, &kb_per_cta, &ksplit);
  return align256(static_cast<size_t>(Bq) * Vp * 2) + align256(static_cast<size_t>(N) * Vp * 2) +
         align256(static_cast<size_t>(ksplit) * Bq * N * sizeof(float));
}

// Where scores_tc_launch keeps the bf16 document operand inside its workspace ([N, *Vp] bf16): the peer gather
// kernel (peer.cu) writes it there directly and then launches with docs_prepacked = true.
void* scores_tc_docs16(void* workspace, int Bq, int N, int V, int* Vp_out) {
  int Vp, n_tile, m_tiles, n_tiles, kb_total, kb_per_cta, ksplit;
  plan(Bq, N, V, &Vp, &n_tile, &m_tiles, &n_tiles, &kb_total, &kb_per_cta, &ksplit);
  *Vp_out = Vp;
  return static_cast<uint8_t*>(workspace) + align256(static_cast<size_t>(Bq) * Vp * 2);
}

// Returns the number of k-splits written to `*partial_out` ([ksplit][Bq][N]); the caller reduces them.
int scores_tc_launch(const float* q, int Bq, const float* docs, int N, int V, void* workspace, size_t ws_bytes,
                     float** partial_out, int* ksplit_out, bool docs_prepacked, cudaStream_t stream,
                     bool q_prepacked) {
  int rc = require_sm100("splade_scores");
  if (rc != kOk) return rc;
  int Vp, n_tile, m_tiles, n_tiles, kb_total, kb_per_cta, ksplit;
  plan(Bq, N, V, &Vp, &n_tile, &m_tiles, &n_tiles, &kb_total, &kb_per_cta, &ksplit);
  const size_t need = scores_tc_workspace_bytes(Bq, N, V);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "splade_scores: workspace too small (%zu < %zu bytes)", ws_bytes, need);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* q16 = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align256(static_cast<size_t>(Bq) * Vp * 2);
  __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align256(static_cast<size_t>(N) * Vp * 2);
  float* partial = reinterpret_cast<float*>(ws);

  const unsigned chunks = static_cast<unsigned>((Vp / 8 + 255) / 256);
  if (!q_prepacked) pack_bf16_kernel<<<dim3(chunks, Bq), 256, 0, stream>>>(q, V, Vp, q16);
  if (!docs_prepacked) pack_bf16_kernel<<<dim3(chunks, N), 256, 0, stream>>>(docs, V, Vp, d16);
  SPLADE_CUDA_OK(cudaGetLastError());

  CUtensorMap tmap_q, tmap_d;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Vp), static_cast<uint64_t>(Bq)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Vp) * 2};
    const uint32_t box[2] = {BK, TM};
    rc = encode_tmap_bf16(&tmap_q, q16, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Vp), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Vp) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(n_tile)};
    rc = encode_tmap_bf16(&tmap_d, d16, 2, dims, strides, box);
    if (rc != kOk) return rc;
  }
  ScoresTcParams p{};
  p.partial = partial;
  p.Bq = Bq;
  p.N = N;
  p.n_tile = n_tile;
  p.kb_total = kb_total;
  p.kb_per_cta = kb_per_cta;
  p.stage_tx_bytes = static_cast<uint32_t>(A_BYTES + n_tile * BK * 2);
  SPLADE_CUDA_OK(cudaFuncSetAttribute(scores_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(ksplit, m_tiles, n_tiles);
  scores_tc_kernel<<<grid, 128, SMEM_BYTES, stream>>>(tmap_q, tmap_d, p);
  SPLADE_CUDA_OK(cudaGetLastError());
  *partial_out = partial;
  *ksplit_out = ksplit;
  return kOk;
}

}  // namespace splade
