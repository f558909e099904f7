// Tensor-core backward of the InfoNCE score GEMM (autocast mode):
//     dq[i,v]    = gl * ( sum_j G[i,j] * D[j,v]  +  sum_kk g_hard[i,kk] * negs[i,kk,v]  +  cq * colmean_q[v] )
//     ddocs[j,v] = gl * ( sum_i G[i,j] * Q[i,v]  +  [j is a local positive] cd * colmean_d[v] )
//     dnegs[i,kk,v] = gl * ( g_hard[i,kk] * Q[i,v] + cn * colmean_n[v] )
// G = d loss / d scores [Bq, N] (written by the scalar stage), D / Q = the pooled vectors.  Under torch.autocast the
// reference's backward of torch.mm(anchor, positive.t()) (src/model/losses.py:155, :243) is itself a bf16 GEMM on the
// bf16 casts of its operands with a bf16-rounded G and a bf16-rounded result; here the operands are the SAME bf16 casts
// the forward packed for the score GEMM (kept alive for the backward), G is rounded to bf16, accumulation and output
// stay fp32 (strictly closer to the exact gradient than the reference's).  Round 1 ran these products on CUDA cores
// in fp32: 47 + 20 us at N = 64 documents and 0.2 ms each at N = 512 (eight ranks' positives gathered), where they
// were the largest term of the gathered-negatives loss.
//
// Both products are C[m, v] = sum_k A[m,k] * X[k,v] with X = a [K, Vp] bf16 matrix whose ROWS are the contraction
// index (D: rows = documents, Q: rows = queries):
//   A  K-major  [rows, Kp] bf16   dq: G16 [Bq, N]      ddocs: G16^T [N, Bq]      TMA box (64 k, 128 rows), SW128
//   B  MN-major                   the TMA-loaded [64 k rows][64 v] tile as it lies in X (128 B swizzle rows): two
//                                 64-column atoms per CTA = 128 vocabulary columns, two 128x64 TMEM accumulators
//   tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 64, K = 16 per instruction; 4-stage TMA ring over K blocks of 64.
// Grid = (ceil(V / 128), ceil(rows / 128)), ~68 KB of shared memory => three CTAs per SM, one wave at V = 50 000.
// Epilogue: the accumulator tile goes TMEM -> shared memory, then all threads walk (row, 4 columns) items so that every
// warp instruction touches 512 contiguous bytes of one fp32 row (the first version kept a thread = one row: half-used
// sectors and a serial chain of dependent loads per thread, 101 us at N = 64 against 61 us for the CUDA-core kernels).
// Memory-bound by construction: X is read once per row tile (51 MB at N = 512), the outputs written once.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace splade {
namespace {

constexpr int TM = 128;
constexpr int BK = 64;
constexpr int UK = 16;
constexpr int NATOM = 2;                   // 64-column atoms per CTA
constexpr int TV = 64 * NATOM;             // vocabulary columns per CTA
constexpr int STAGES = 2;                  // the contraction is at most 8 blocks long
constexpr int A_BYTES = TM * BK * 2;       // 16 KB
constexpr int B_ATOM_BYTES = BK * 64 * 2;  // 8 KB
constexpr int STAGE_BYTES = A_BYTES + NATOM * B_ATOM_BYTES;
constexpr int TILE_LD = TV + 4;            // padded row of the staged fp32 accumulator tile (conflict-free float4)
constexpr int TILE_BYTES = TM * TILE_LD * 4;
constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = (PIPE_BYTES > TILE_BYTES ? PIPE_BYTES : TILE_BYTES) + 256 + 1024;  // ~68 KB: 3 CTAs/SM
constexpr int TMEM_COLS = 128;

// MN-major operand tile with 128-byte swizzle: 64 contiguous MN elements (128 B) per K row, 8-row groups 1024 B apart
// (the layout the attention kernel of trunk.cu uses for V).
__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct BwdTcParams {
  int rows;          // valid output rows (dq: Bq, ddocs: N)
  int kb_total;      // 64-wide blocks of the contraction
  int V;
  int mode;          // 0: dq + dnegs, 1: ddocs
  int Bq, k, label_offset;
  const float* g_loss;
  // mode 0
  const float* q;        // [Bq, V]
  const float* negs;     // [Bq*k, V]
  const float* g_hard;   // [Bq, k]
  const float* cm_q;
  const float* cm_n;
  float cq, cn;
  float* dq;
  float* dnegs;
  // mode 1
  const float* cm_d;
  float cd;
  float* ddocs;
};

__global__ void __launch_bounds__(128, 3)
loss_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_x,
                   const BwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (PIPE_BYTES > TILE_BYTES ? PIPE_BYTES : TILE_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done_bar + 1);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int v0 = blockIdx.x * TV;
  const int m0 = blockIdx.y * TM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_x);
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
    for (int kb = 0; kb < p.kb_total; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* a_dst = smem + stage * STAGE_BYTES;
      mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
      tma_load_2d(a_dst, &tmap_a, &full_bar[stage], kb * BK, m0);
#pragma unroll
      for (int a = 0; a < NATOM; ++a)
        tma_load_2d(a_dst + A_BYTES + a * B_ATOM_BYTES, &tmap_x, &full_bar[stage], v0 + a * 64, kb * BK);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(TM, 64) | (1u << 16);  // B operand MN-major
    uint32_t stage = 0, phase = 0;
    for (int kb = 0; kb < p.kb_total; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
      const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
      for (int ks = 0; ks < BK / UK; ++ks) {
#pragma unroll
        for (int a = 0; a < NATOM; ++a)
          umma_bf16(tmem_base + static_cast<uint32_t>(a * 64), make_smem_desc_sw128(a_addr + ks * UK * 2),
                    make_desc_sw128_mn(b_addr + a * B_ATOM_BYTES + ks * UK * 128), idesc,
                    (kb > 0 || ks > 0) ? 1u : 0u);
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

  // ---- epilogue.  Phase 1: TMEM -> shared memory (the pipeline stages are free: every MMA has completed), a thread
  //      = one accumulator row.  Phase 2: all 128 threads walk (row, 4 columns) items so that a warp touches 512
  //      contiguous bytes of ONE row of q / negs / dq / dnegs / ddocs per instruction, four items in flight.
  mbar_wait(done_bar, 0);
  tcgen05_fence_after();
  __syncwarp();
  float* tile = reinterpret_cast<float*>(smem);
  {
    const uint32_t my_row = warp * 32u + lane;
    for (int c0 = 0; c0 < TV; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((warp * 32u) << 16) + static_cast<uint32_t>(c0), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(tile + my_row * TILE_LD + c0 + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3]));
    }
  }
  __syncthreads();
  const int rows_valid = min(TM, p.rows - m0);
  const float gl = __ldg(p.g_loss);
  const int V = p.V;
  constexpr int C4 = TV / 4;
  const int n_items = rows_valid * C4;
  for (int it0 = static_cast<int>(threadIdx.x); it0 < n_items; it0 += 128 * 4) {
    if (p.mode == 0) {
      float4 acc[4], qv[4], mq[4], mn[4];
      int rr[4], vv[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * 128;
        rr[u] = it / C4;
        vv[u] = v0 + (it - rr[u] * C4) * 4;
        ok[u] = it < n_items && vv[u] < V;
        if (ok[u]) {
          acc[u] = *reinterpret_cast<const float4*>(tile + rr[u] * TILE_LD + (vv[u] - v0));
          qv[u] = __ldg(reinterpret_cast<const float4*>(p.q + static_cast<size_t>(m0 + rr[u]) * V + vv[u]));
          mq[u] = __ldg(reinterpret_cast<const float4*>(p.cm_q + vv[u]));
          mn[u] = __ldg(reinterpret_cast<const float4*>(p.cm_n + vv[u]));
        }
      }
      for (int kk = 0; kk < p.k; ++kk) {
        float4 nv[4];
        float gh[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (ok[u]) {
            const size_t rk = static_cast<size_t>(m0 + rr[u]) * p.k + kk;
            gh[u] = __ldg(&p.g_hard[rk]);
            nv[u] = __ldg(reinterpret_cast<const float4*>(p.negs + rk * V + vv[u]));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (ok[u]) {
            const size_t rk = static_cast<size_t>(m0 + rr[u]) * p.k + kk;
            acc[u].x = fmaf(gh[u], nv[u].x, acc[u].x);
            acc[u].y = fmaf(gh[u], nv[u].y, acc[u].y);
            acc[u].z = fmaf(gh[u], nv[u].z, acc[u].z);
            acc[u].w = fmaf(gh[u], nv[u].w, acc[u].w);
            *reinterpret_cast<float4*>(p.dnegs + rk * V + vv[u]) =
                make_float4(gl * (gh[u] * qv[u].x + p.cn * mn[u].x), gl * (gh[u] * qv[u].y + p.cn * mn[u].y),
                            gl * (gh[u] * qv[u].z + p.cn * mn[u].z), gl * (gh[u] * qv[u].w + p.cn * mn[u].w));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ok[u])
          *reinterpret_cast<float4*>(p.dq + static_cast<size_t>(m0 + rr[u]) * V + vv[u]) =
              make_float4(gl * (acc[u].x + p.cq * mq[u].x), gl * (acc[u].y + p.cq * mq[u].y),
                          gl * (acc[u].z + p.cq * mq[u].z), gl * (acc[u].w + p.cq * mq[u].w));
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * 128;
        const int r_ = it / C4;
        const int v = v0 + (it - r_ * C4) * 4;
        if (it < n_items && v < V) {
          const int row = m0 + r_;
          const float f = (row >= p.label_offset && row < p.label_offset + p.Bq) ? p.cd : 0.f;
          const float4 a = *reinterpret_cast<const float4*>(tile + r_ * TILE_LD + (v - v0));
          const float4 md = __ldg(reinterpret_cast<const float4*>(p.cm_d + v));
          *reinterpret_cast<float4*>(p.ddocs + static_cast<size_t>(row) * V + v) =
              make_float4(gl * (a.x + f * md.x), gl * (a.y + f * md.y), gl * (a.z + f * md.z), gl * (a.w + f * md.w));
        }
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

// G [Bq, N] fp32 -> G16 [Bq, Np] and G16^T [N, Bp] bf16 (Np / Bp = N / Bq rounded up to 8; pads zero)
__global__ void __launch_bounds__(256)
pack_g_kernel(const float* __restrict__ g, int Bq, int N, int Np, int Bp, __nv_bfloat16* __restrict__ g16,
              __nv_bfloat16* __restrict__ gt16) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < Bq * Np) {
    const int i = idx / Np, j = idx - i * Np;
    g16[idx] = __float2bfloat16_rn(j < N ? g[static_cast<size_t>(i) * N + j] : 0.f);
  }
  if (gt16 && idx < N * Bp) {
    const int j = idx / Bp, i = idx - j * Bp;
    gt16[idx] = __float2bfloat16_rn(i < Bq ? g[static_cast<size_t>(i) * N + j] : 0.f);
  }
}

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

}  // namespace

size_t loss_bwd_tc_workspace_bytes(int Bq, int N) {
  const size_t Np = (static_cast<size_t>(N) + 7) / 8 * 8, Bp = (static_cast<size_t>(Bq) + 7) / 8 * 8;
  return align256(static_cast<size_t>(Bq) * Np * 2) + align256(static_cast<size_t>(N) * Bp * 2);
}

// q16 [Bq, Vp] / d16 [N, Vp]: the bf16 operands the forward packed (Vp = V rounded up to 8).  ddocs == nullptr: dq and
// dnegs only (the peer exchange computes ddocs itself).
int loss_bwd_tc_launch(const float* q, const float* negs, const void* q16, const void* d16, int Bq, int N, int k, int V,
                       int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                       const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                       const float* g_loss, float* dq, float* ddocs, float* dnegs, void* workspace, size_t ws_bytes,
                       cudaStream_t stream) {
  const char* who = "splade_loss_bwd_tc";
  if (!q || !negs || !q16 || !d16 || !g_inbatch || !g_hard || !colmean_q || !colmean_n || !g_loss || !dq || !dnegs)
    return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  if (ddocs && !colmean_d) return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  if (Bq < 1 || N < 1 || k < 1 || V < 1) return set_error(kErrInvalidArg, "%s: bad sizes", who);
  auto aligned16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  if (V % 4 != 0 || !aligned16(q) || !aligned16(negs) || !aligned16(dq) || !aligned16(dnegs) ||
      !aligned16(colmean_q) || !aligned16(colmean_n) || (ddocs && (!aligned16(ddocs) || !aligned16(colmean_d))) ||
      !aligned16(q16) || !aligned16(d16))
    return set_error(kErrUnsupportedShape, "%s: V %% 4 == 0 and 16-byte aligned rows are required", who);
  int rc = require_sm100(who);
  if (rc != kOk) return rc;
  const size_t need = loss_bwd_tc_workspace_bytes(Bq, N);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need);
  const int Vp = (V + 7) / 8 * 8;
  const int Np = (N + 7) / 8 * 8, Bp = (Bq + 7) / 8 * 8;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* g16 = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align256(static_cast<size_t>(Bq) * Np * 2);
  __nv_bfloat16* gt16 = reinterpret_cast<__nv_bfloat16*>(ws);
  const int n_pack = std::max(Bq * Np, ddocs ? N * Bp : 0);
  pack_g_kernel<<<(n_pack + 255) / 256, 256, 0, stream>>>(g_inbatch, Bq, N, Np, Bp, g16, ddocs ? gt16 : nullptr);
  SPLADE_CUDA_OK(cudaGetLastError());
  SPLADE_CUDA_OK(cudaFuncSetAttribute(loss_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));

  auto encode_a = [&](CUtensorMap* m, const void* base, int rows, int K, int Kp) {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Kp) * 2};
    const uint32_t box[2] = {BK, TM};
    return encode_tmap_bf16(m, base, 2, dims, strides, box);
  };
  auto encode_x = [&](CUtensorMap* m, const void* base, int rows) {
    const uint64_t dims[2] = {static_cast<uint64_t>(Vp), static_cast<uint64_t>(rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Vp) * 2};
    const uint32_t box[2] = {64, BK};
    return encode_tmap_bf16(m, base, 2, dims, strides, box);
  };
  const unsigned gx = static_cast<unsigned>((V + TV - 1) / TV);

  BwdTcParams p{};
  p.V = V, p.Bq = Bq, p.k = k, p.label_offset = label_offset, p.g_loss = g_loss;
  {  // dq (+ dnegs): A = G16 [Bq, N], X = D16 [N, Vp]
    CUtensorMap ta, tx;
    rc = encode_a(&ta, g16, Bq, N, Np);
    if (rc != kOk) return rc;
    rc = encode_x(&tx, d16, N);
    if (rc != kOk) return rc;
    p.mode = 0, p.rows = Bq, p.kb_total = (N + BK - 1) / BK;
    p.q = q, p.negs = negs, p.g_hard = g_hard, p.cm_q = colmean_q, p.cm_n = colmean_n, p.cq = cq, p.cn = cn;
    p.dq = dq, p.dnegs = dnegs;
    loss_bwd_tc_kernel<<<dim3(gx, (Bq + TM - 1) / TM), 128, SMEM_BYTES, stream>>>(ta, tx, p);
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  if (ddocs) {  // A = G16^T [N, Bq], X = Q16 [Bq, Vp]
    CUtensorMap ta, tx;
    rc = encode_a(&ta, gt16, N, Bq, Bp);
    if (rc != kOk) return rc;
    rc = encode_x(&tx, q16, Bq);
    if (rc != kOk) return rc;
    p.mode = 1, p.rows = N, p.kb_total = (Bq + BK - 1) / BK;
    p.cm_d = colmean_d, p.cd = cd, p.ddocs = ddocs;
    loss_bwd_tc_kernel<<<dim3(gx, (N + TM - 1) / TM), 128, SMEM_BYTES, stream>>>(ta, tx, p);
    SPLADE_CUDA_OK(cudaGetLastError());
  }
  return kOk;
}

}  // namespace splade
