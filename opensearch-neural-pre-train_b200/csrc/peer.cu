// Cross-GPU negatives for SPLADELossV33 over NVLink peer memory (include/splade_b200.h, splade_peer_*).
//
// The reference scores a rank's queries against that rank's documents only (src/model/losses.py:136-181); the
// north-star extension scores them against every rank's positives.  The NCCL statement of
// that step is all-gather -> score GEMM forward and GEMM -> reduce-scatter backward (distributed.py).  Here the
// transfers are folded into the kernels: every rank exports one "exchange window" with CUDA IPC, peers map it,
// and the kernels read each other's windows directly (NVSwitch gives every pair the full NVLink bandwidth).
//
//   window (per rank):   [flags 2x16 u32 | status u32 | pad to 512 B]
//                        parity 0/1: Q [Bq,V] f32 | P [Bq,V] f32 | G [Bq, W*Bq] f32 | gl f32
//
//   forward  epoch e, parity e&1:
//     D2D   q -> Q, local positives -> P                                   (publish)
//     K1    peer_barrier_kernel      flag e into every peer's slot 0, wait for theirs
//     K2    peer_gather_kernel       pull P of every rank -> docs_all fp32 (kept for backward) and, under
//                                    autocast, the bf16 operand of the tcgen05 score GEMM in the same pass
//     ...   the unchanged single-GPU loss pipeline on docs_all, with g_inbatch written into G
//   backward:
//     D2D   g_loss -> gl                                                   (publish)
//     K3    loss_bwd_dq_kernel       dq, dnegs from the local docs_all copy (loss.cu)
//     K4    peer_bwd_sync_kernel     CTA s: flag e into peer s's slot 1, wait for its flag, pull gl_s * G_s[:, mine]
//     K5    peer_ddocs_kernel        ddocs_local = sum_s (gl_s G_s[:, mine])^T q_s, q_s TMA-bulk-copied from peer
//                                    s's window into shared memory (3-stage mbarrier pipeline)
//
// Double buffering makes one barrier per phase sufficient: a window half is rewritten at epoch e+2, after its
// owner passed the epoch e+1 barrier, which every peer only signals after its epoch e kernels (stream order).
#include "../../include/splade_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>
#include <cstring>

namespace splade {
namespace {

constexpr int kMaxPeers = SPLADE_MAX_PEERS;
constexpr size_t kHeaderBytes = 512;   // flags [2][kMaxPeers] u32 at 0, status u32 at 256
constexpr size_t kStatusOffset = 256;

__host__ __device__ inline size_t al256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct PeerLayout {
  size_t q[2], p[2], g[2], gl[2], total;
};

__host__ __device__ inline PeerLayout peer_layout(int Bq, int V, int W) {
  PeerLayout L;
  size_t o = kHeaderBytes;
  const size_t rep = al256(static_cast<size_t>(Bq) * V * sizeof(float));
  const size_t g = al256(static_cast<size_t>(Bq) * W * Bq * sizeof(float));
  for (int b = 0; b < 2; ++b) {
    L.q[b] = o;
    o += rep;
    L.p[b] = o;
    o += rep;
    L.g[b] = o;
    o += g;
    L.gl[b] = o;
    o += 256;
  }
  L.total = o;
  return L;
}

struct PeerPtrs {
  uint8_t* win[kMaxPeers];
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Loads of peer memory: plain (coherent) global loads, never the non-coherent path.
__device__ __forceinline__ float4 ld_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_f2(const float* p) {
  float2 v;
  asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_f1(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Tell `peer` that this rank reached `value` in phase `slot`, then wait until the peer says the same.
// Flags only grow; a peer that does not answer within timeout_ns marks the status word and the wait gives up
// (the step's numbers are then garbage, but the GPU is not hung; splade_peer_status reports it).
__device__ bool peer_signal_wait(const PeerPtrs& g, int peer, int slot, uint32_t value, unsigned long long timeout_ns) {
  uint32_t* remote = reinterpret_cast<uint32_t*>(g.win[peer]) + slot * kMaxPeers + g.rank;
  __threadfence_system();
  st_release_sys(remote, value);
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(g.win[g.rank]) + slot * kMaxPeers + peer;
  const uint32_t* status = reinterpret_cast<const uint32_t*>(g.win[g.rank] + kStatusOffset);
  const unsigned long long t0 = global_ns();
  while (static_cast<int32_t>(ld_acquire_sys(mine) - value) < 0) {
    // a timeout is sticky: once one barrier of this rank gave up, later ones do not wait another full period
    if (ld_acquire_sys(status) != 0u) return false;
    if (global_ns() - t0 > timeout_ns) {
      atomicExch(reinterpret_cast<unsigned int*>(g.win[g.rank] + kStatusOffset),
                 0x100u * static_cast<unsigned>(slot + 1) + static_cast<unsigned>(peer) + 1u);
      return false;
    }
    __nanosleep(128);
  }
  return true;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerPtrs g, int slot, uint32_t value,
                                                          unsigned long long timeout_ns) {
  const int w = threadIdx.x;
  if (w < g.world && w != g.rank) peer_signal_wait(g, w, slot, value, timeout_ns);
}

// ---------------------------------------------------------------------------------------------
// K2: docs_all[s*rows + r, :] = P_s[r, :]  (+ the bf16 GEMM operand, rows padded to Vp) — 8 columns per thread
// ---------------------------------------------------------------------------------------------
struct GatherParams {
  const float* seg[kMaxPeers];
  int W, rank;
  int rows, V, Vp;
  float* out32;            // [W*rows, V]
  __nv_bfloat16* out16;    // [W*rows, Vp] or nullptr
};

__global__ void __launch_bounds__(256) peer_gather_kernel(const GatherParams p) {
  // Consecutive CTAs read from DIFFERENT ranks, and rank r starts its rotation at r+1: at any moment every
  // rank's requests are spread over all peers, so no peer's NVLink egress serves W-1 readers while the rest idle.
  const int r = blockIdx.y / p.W;
  const int s = (p.rank + 1 + static_cast<int>(blockIdx.y % p.W)) % p.W;
  const int row = s * p.rows + r;
  const int v0 = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (v0 >= p.Vp) return;
  const float* src = p.seg[s] + static_cast<size_t>(r) * p.V;
  float* dst = p.out32 + static_cast<size_t>(row) * p.V;
  float x[8];
  if (v0 + 8 <= p.V) {  // V % 4 == 0 and 256-byte aligned windows: rows are 16-byte aligned
    const float4 a = ld_f4(src + v0);
    const float4 b = ld_f4(src + v0 + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    *reinterpret_cast<float4*>(dst + v0) = a;
    *reinterpret_cast<float4*>(dst + v0 + 4) = b;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      x[e] = (v0 + e < p.V) ? ld_f1(src + v0 + e) : 0.f;
      if (v0 + e < p.V) dst[v0 + e] = x[e];
    }
  }
  if (p.out16) {
    uint4 o;
    o.x = cvt_bf16x2(x[1], x[0]);
    o.y = cvt_bf16x2(x[3], x[2]);
    o.z = cvt_bf16x2(x[5], x[4]);
    o.w = cvt_bf16x2(x[7], x[6]);
    *reinterpret_cast<uint4*>(p.out16 + static_cast<size_t>(row) * p.Vp + v0) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// K4: CTA s synchronises with peer s (phase 1) and pulls Gloc[s][i][j] = gl_s * G_s[i][rank*Bq + j]
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  PeerPtrs g;
  size_t q_off, g_off, gl_off;
  int Bq, V;
  uint32_t epoch;
  unsigned long long timeout_ns;
  float* gloc;            // [W][Bq][Bq]
  const float* cm_d;      // [V]
  float cd;
  float* ddocs;           // [Bq,V]
  int round_bf16;         // autocast mode: G and q enter the product rounded to bf16 (what splade_loss_bwd_tc multiplies)
};

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__global__ void __launch_bounds__(256) peer_bwd_sync_kernel(const BwdParams p) {
  const int s = blockIdx.x;
  if (threadIdx.x == 0 && s != p.g.rank) peer_signal_wait(p.g, s, 1, p.epoch, p.timeout_ns);
  __syncthreads();
  const int Bq = p.Bq, N = p.g.world * Bq;
  const float* Gs = reinterpret_cast<const float*>(p.g.win[s] + p.g_off);
  const float gls = ld_f1(reinterpret_cast<const float*>(p.g.win[s] + p.gl_off));
  float* out = p.gloc + static_cast<size_t>(s) * Bq * Bq;
  for (int base = 0; base < Bq * Bq; base += 256 * 8) {  // eight remote loads in flight per thread
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * 256 + static_cast<int>(threadIdx.x);
      const int i = e / Bq, j = e - i * Bq;
      v[u] = (e < Bq * Bq) ? ld_f1(Gs + static_cast<size_t>(i) * N + p.g.rank * Bq + j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * 256 + static_cast<int>(threadIdx.x);
      if (e < Bq * Bq) out[e] = gls * (p.round_bf16 ? bf16r(v[u]) : v[u]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K5: ddocs_local[j, :] = sum_s sum_i Gloc[s][i][j] * q_s[i, :] + gl_rank * cd * colmean_d
// A CTA owns 256 vocabulary columns and PD local documents; each thread two columns, accumulators in
// registers.  The q rows come straight out of the owning rank's window: one elected thread issues TMA bulk
// copies (cp.async.bulk, 1 KB per row, 16 rows per stage, 3 stages) that land in shared memory and complete
// on an mbarrier, so ~32 KB per CTA are in flight over NVLink while the previous tile is being multiplied.
// Every q element crosses the link exactly once.  CTAs start their rotation over the ranks at different
// peers (rank + 1 + blockIdx.x), which spreads the reads over all links at all times.
// ---------------------------------------------------------------------------------------------
constexpr int PD = 64;        // local documents per CTA
constexpr int PK = 16;        // query rows per stage
constexpr int PKP = 20;       // padded row of the staged G tile (keeps 16-byte alignment, spreads banks)
constexpr int PCOLS = 256;    // vocabulary columns per CTA
constexpr int PSTAGES = 3;
constexpr int PSTAGE_BYTES = PK * PCOLS * 4;
constexpr int PDDOCS_SMEM = PSTAGES * PSTAGE_BYTES + PD * PKP * 4 + PSTAGES * 8 + 128;

__device__ __forceinline__ void bulk_load_row(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// mbar_wait of common.cuh reports a timeout with printf, which gives the kernel a stack frame; the first launch of
// a kernel with a larger local-memory footprint than any before it makes the driver synchronise the device, and
// that must not happen while a peer's barrier kernel is waiting for this rank.  Same bound, no message.
__device__ __forceinline__ void mbar_wait_quiet(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(addr, parity)) {
    if (clock64() - t0 > SPLADE_MBAR_TIMEOUT_CYCLES) __trap();
  }
}

__global__ void __launch_bounds__(128) peer_ddocs_kernel(const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float* stage_base = reinterpret_cast<float*>(smem);                                   // [PSTAGES][PK][PCOLS]
  float(*Gt)[PKP] = reinterpret_cast<float(*)[PKP]>(smem + PSTAGES * PSTAGE_BYTES);    // Gt[jj][ii]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PSTAGES * PSTAGE_BYTES + PD * PKP * 4);

  const int Bq = p.Bq, V = p.V, W = p.g.world;
  const int c0 = blockIdx.x * PCOLS;
  const int cols = min(PCOLS, V - c0);            // multiple of 4 (V % 4 == 0)
  const uint32_t row_bytes = static_cast<uint32_t>(cols) * 4u;
  const int c = c0 + 2 * static_cast<int>(threadIdx.x);
  const bool live = 2 * static_cast<int>(threadIdx.x) < cols;
  const int j0 = blockIdx.y * PD;
  const int tiles_per_rank = (Bq + PK - 1) / PK;
  const int T = W * tiles_per_rank;
  const int s_first = (p.g.rank + 1 + static_cast<int>(blockIdx.x)) % W;

  if (threadIdx.x == 0) {
    for (int st = 0; st < PSTAGES; ++st) mbar_init(&full_bar[st], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // thread 0 only; a macro-like helper without a closure object (the kernel must not need a stack frame: the
  // first launch of a kernel with a larger local-memory footprint synchronises the device)
#define SPLADE_PEER_ISSUE(t_)                                                                                     \
  do {                                                                                                            \
    const int t__ = (t_);                                                                                         \
    const int s__ = (s_first + t__ / tiles_per_rank) % W;                                                         \
    const int i__ = (t__ % tiles_per_rank) * PK;                                                                  \
    const int rows__ = min(PK, Bq - i__);                                                                         \
    const int st__ = t__ % PSTAGES;                                                                               \
    const float* src__ =                                                                                          \
        reinterpret_cast<const float*>(p.g.win[s__] + p.q_off) + static_cast<size_t>(i__) * V + c0;               \
    float* dst__ = stage_base + static_cast<size_t>(st__) * PK * PCOLS;                                           \
    mbar_arrive_expect_tx(&full_bar[st__], row_bytes * static_cast<uint32_t>(rows__));                            \
    for (int r__ = 0; r__ < rows__; ++r__)                                                                        \
      bulk_load_row(dst__ + r__ * PCOLS, src__ + static_cast<size_t>(r__) * V, row_bytes, &full_bar[st__]);       \
  } while (0)

  if (threadIdx.x == 0)
    for (int t = 0; t < PSTAGES - 1 && t < T; ++t) SPLADE_PEER_ISSUE(t);

  float2 acc[PD];
#pragma unroll
  for (int jj = 0; jj < PD; ++jj) acc[jj] = make_float2(0.f, 0.f);

  for (int t = 0; t < T; ++t) {
    const int s = (s_first + t / tiles_per_rank) % W;
    const int i0 = (t % tiles_per_rank) * PK;
    const int st = t % PSTAGES;
    __syncthreads();  // everyone is done with tile t-1: its stage and Gt may be overwritten
    if (threadIdx.x == 0 && t + PSTAGES - 1 < T) SPLADE_PEER_ISSUE(t + PSTAGES - 1);
    const float* gl = p.gloc + static_cast<size_t>(s) * Bq * Bq;
    for (int e = threadIdx.x; e < PD * PK; e += 128) {
      const int ii = e / PD, jj = e - ii * PD;  // consecutive threads -> consecutive local documents
      Gt[jj][ii] = (i0 + ii < Bq && j0 + jj < Bq) ? gl[static_cast<size_t>(i0 + ii) * Bq + j0 + jj] : 0.f;
    }
    __syncthreads();
    mbar_wait_quiet(&full_bar[st], static_cast<uint32_t>((t / PSTAGES) & 1));
    const float* xs = stage_base + static_cast<size_t>(st) * PK * PCOLS + 2 * threadIdx.x;
#pragma unroll
    for (int h = 0; h < PK; h += 4) {
      float2 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)  // rows past the batch were not loaded: never touch their (stale) bytes
        x[u] = (live && i0 + h + u < Bq) ? *reinterpret_cast<const float2*>(xs + (h + u) * PCOLS) : make_float2(0.f, 0.f);
      if (p.round_bf16) {
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = make_float2(bf16r(x[u].x), bf16r(x[u].y));
      }
#pragma unroll
      for (int jj = 0; jj < PD; ++jj) {
        const float4 g = *reinterpret_cast<const float4*>(&Gt[jj][h]);
        float2 a = acc[jj];
        a.x = fmaf(g.x, x[0].x, a.x); a.y = fmaf(g.x, x[0].y, a.y);
        a.x = fmaf(g.y, x[1].x, a.x); a.y = fmaf(g.y, x[1].y, a.y);
        a.x = fmaf(g.z, x[2].x, a.x); a.y = fmaf(g.z, x[2].y, a.y);
        a.x = fmaf(g.w, x[3].x, a.x); a.y = fmaf(g.w, x[3].y, a.y);
        acc[jj] = a;
      }
    }
  }
  if (!live) return;
  const float glr = ld_f1(reinterpret_cast<const float*>(p.g.win[p.g.rank] + p.gl_off));
  const float2 md = *reinterpret_cast<const float2*>(p.cm_d + c);
  const float fx = glr * p.cd * md.x, fy = glr * p.cd * md.y;
#pragma unroll
  for (int jj = 0; jj < PD; ++jj) {
    const int j = j0 + jj;
    if (j < Bq)
      *reinterpret_cast<float2*>(p.ddocs + static_cast<size_t>(j) * V + c) = make_float2(acc[jj].x + fx, acc[jj].y + fy);
  }
#undef SPLADE_PEER_ISSUE
}

unsigned long long timeout_ns_from_env() {
  static unsigned long long cached = 0;
  if (cached == 0) {
    unsigned long long ms = 600000;  // NCCL-like patience: ranks legitimately skew by minutes (rank-0 evaluation, checkpoints)
    if (const char* e = std::getenv("SPLADE_B200_PEER_TIMEOUT_MS")) {
      const long long v = std::atoll(e);
      if (v > 0) ms = static_cast<unsigned long long>(v);
    }
    cached = ms * 1000000ull;
  }
  return cached;
}

int check_group(const splade_peer_group* g, const char* who) {
  if (!g) return set_error(kErrInvalidArg, "%s: null peer group", who);
  if (g->world < 1 || g->world > kMaxPeers || g->rank < 0 || g->rank >= g->world)
    return set_error(kErrInvalidArg, "%s: bad peer group (rank %d of %d, at most %d peers)", who, g->rank, g->world,
                     kMaxPeers);
  for (int s = 0; s < g->world; ++s)
    if (!g->window[s]) return set_error(kErrInvalidArg, "%s: window of rank %d is not mapped", who, s);
  return kOk;
}

PeerPtrs make_ptrs(const splade_peer_group* g) {
  PeerPtrs p{};
  for (int s = 0; s < g->world; ++s) p.win[s] = static_cast<uint8_t*>(g->window[s]);
  p.rank = g->rank;
  p.world = g->world;
  return p;
}

}  // namespace

size_t peer_window_bytes(int Bq, int V, int world) {
  if (Bq < 1 || V < 1 || world < 1 || world > kMaxPeers) return 0;
  return peer_layout(Bq, V, world).total;
}

size_t peer_loss_bwd_workspace_bytes(int Bq, int world) {
  if (Bq < 1 || world < 1) return 0;
  // [gloc: world * Bq * Bq fp32 | scratch of the tensor-core dq (bf16 G)]
  return al256(static_cast<size_t>(world) * Bq * Bq * sizeof(float)) + loss_bwd_tc_workspace_bytes(Bq, world * Bq);
}

}  // namespace splade

using namespace splade;

extern "C" {

size_t splade_peer_window_bytes(int Bq, int V, int world) { return peer_window_bytes(Bq, V, world); }

int splade_peer_alloc(size_t bytes, void** dptr) {
  if (!dptr || bytes == 0) return set_error(kErrInvalidArg, "splade_peer_alloc: bad argument");
  // Load the exchange kernels now: with lazy module loading the first launch of a kernel may synchronise the
  // context, which must not happen while another rank's barrier kernel is spinning on this device.
  cudaFuncAttributes fa;
  SPLADE_CUDA_OK(cudaFuncGetAttributes(&fa, peer_barrier_kernel));
  SPLADE_CUDA_OK(cudaFuncGetAttributes(&fa, peer_gather_kernel));
  SPLADE_CUDA_OK(cudaFuncGetAttributes(&fa, peer_bwd_sync_kernel));
  SPLADE_CUDA_OK(cudaFuncGetAttributes(&fa, peer_ddocs_kernel));
  // ... and configure them here for the same reason (the first cudaFuncSetAttribute on a function may block
  // behind running kernels); windows are allocated on every device that will launch these kernels.
  SPLADE_CUDA_OK(cudaFuncSetAttribute(peer_ddocs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PDDOCS_SMEM));
  void* p = nullptr;
  SPLADE_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return check_cuda(e, "zeroing the exchange window");
  }
  *dptr = p;
  return kOk;
}

int splade_peer_free(void* dptr) {
  if (!dptr) return kOk;
  SPLADE_CUDA_OK(cudaFree(dptr));
  return kOk;
}

int splade_peer_export(const void* dptr, void* handle_out) {
  if (!dptr || !handle_out) return set_error(kErrInvalidArg, "splade_peer_export: null pointer argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is documented as 64 bytes");
  cudaIpcMemHandle_t h;
  SPLADE_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(dptr)));
  std::memcpy(handle_out, &h, sizeof(h));
  return kOk;
}

int splade_peer_open(const void* handle, void** dptr_out) {
  if (!handle || !dptr_out) return set_error(kErrInvalidArg, "splade_peer_open: null pointer argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  SPLADE_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr_out = p;
  return kOk;
}

int splade_peer_close(void* dptr) {
  if (!dptr) return kOk;
  SPLADE_CUDA_OK(cudaIpcCloseMemHandle(dptr));
  return kOk;
}

int splade_peer_status(const splade_peer_group* group, int32_t* status_host) {
  int rc = check_group(group, "splade_peer_status");
  if (rc != kOk) return rc;
  if (!status_host) return set_error(kErrInvalidArg, "splade_peer_status: null pointer argument");
  // the caller's streams may be non-blocking ones the legacy-stream memcpy would not wait for
  SPLADE_CUDA_OK(cudaDeviceSynchronize());
  SPLADE_CUDA_OK(cudaMemcpy(status_host, static_cast<uint8_t*>(group->window[group->rank]) + kStatusOffset,
                            sizeof(int32_t), cudaMemcpyDeviceToHost));
  return kOk;
}

int splade_peer_v33_loss_fwd(const splade_v33_fwd_args* args, const splade_peer_group* group, uint32_t epoch,
                             float* docs_all, void* stream_) {
  const char* who = "splade_peer_v33_loss_fwd";
  int rc = check_group(group, who);
  if (rc != kOk) return rc;
  if (!args || !docs_all) return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  if (!args->q || !args->docs || !args->negs || !args->workspace)
    return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  const int Bq = args->Bq, V = args->V, W = group->world;
  if (Bq != group->Bq || V != group->V || args->N != Bq)
    return set_error(kErrInvalidArg, "%s: shapes (Bq=%d, N=%d, V=%d) do not match the exchange windows (Bq=%d, V=%d)",
                     who, Bq, args->N, V, group->Bq, group->V);
  if (V % 4 != 0) return set_error(kErrUnsupportedShape, "%s: V %% 4 == 0 is required (V=%d)", who, V);
  if (args->has_kd) return set_error(kErrInvalidArg, "%s: KL-KD teacher scores are [Bq,Bq]; not available with gathered positives", who);
  if (epoch == 0) return set_error(kErrInvalidArg, "%s: epochs start at 1", who);
  rc = require_sm100(who);
  if (rc != kOk) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const PeerLayout L = peer_layout(Bq, V, W);
  const int b = static_cast<int>(epoch & 1u);
  uint8_t* mine = static_cast<uint8_t*>(group->window[group->rank]);
  const size_t rep_bytes = static_cast<size_t>(Bq) * V * sizeof(float);

  splade_v33_fwd_args a = *args;
  a.N = W * Bq;
  a.docs = docs_all;
  a.label_offset = group->rank * Bq;
  a.g_inbatch = reinterpret_cast<float*>(mine + L.g[b]);
  const size_t need = v33_loss_fwd_workspace_bytes(Bq, a.N, V, a.k);
  if (a.k < 1 || a.ws_bytes < need)  // checked before anything is enqueued: a rank must not signal and then bail out
    return set_error(kErrWorkspace, "%s: workspace too small (%zu < %zu bytes) or k < 1", who, a.ws_bytes, need);

  SPLADE_CUDA_OK(cudaMemcpyAsync(mine + L.q[b], args->q, rep_bytes, cudaMemcpyDeviceToDevice, stream));
  SPLADE_CUDA_OK(cudaMemcpyAsync(mine + L.p[b], args->docs, rep_bytes, cudaMemcpyDeviceToDevice, stream));

  const PeerPtrs ptrs = make_ptrs(group);
  if (W > 1) peer_barrier_kernel<<<1, 32, 0, stream>>>(ptrs, 0, epoch, timeout_ns_from_env());

  GatherParams gp{};
  for (int s = 0; s < W; ++s) gp.seg[s] = reinterpret_cast<const float*>(ptrs.win[s] + L.p[b]);
  gp.W = W;
  gp.rank = group->rank;
  gp.rows = Bq;
  gp.V = V;
  gp.out32 = docs_all;
  gp.out16 = nullptr;
  gp.Vp = (V + 7) / 8 * 8;
  if (a.round_bf16) {
    int Vp = 0;
    gp.out16 = static_cast<__nv_bfloat16*>(scores_tc_docs16(v33_loss_fwd_scores_ws(&a), Bq, a.N, V, &Vp));
    gp.Vp = Vp;
  }
  const unsigned chunks = static_cast<unsigned>((gp.Vp / 8 + 255) / 256);
  peer_gather_kernel<<<dim3(chunks, static_cast<unsigned>(W * Bq)), 256, 0, stream>>>(gp);
  SPLADE_CUDA_OK(cudaGetLastError());
  rc = v33_loss_fwd_launch(&a, stream, /*docs_prepacked=*/a.round_bf16 != 0);
  if (rc != kOk) return rc;
  // The window's status word rides along with the monitoring scalars (raw int32 bits in scalars[15]): the one
  // device->host copy SPLADELossV33.forward makes anyway tells the caller when a barrier of this rank gave up.
  if (a.scalars)
    SPLADE_CUDA_OK(cudaMemcpyAsync(a.scalars + 15, mine + kStatusOffset, sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                   stream));
  return kOk;
}

int splade_peer_loss_bwd(const splade_peer_group* group, uint32_t epoch, const float* q, const float* docs_all,
                         const float* negs, int Bq, int k, int V, const float* g_hard, const float* colmean_q,
                         const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                         const float* g_loss, float* dq, float* ddocs_local, float* dnegs, const void* q16,
                         const void* d16_all, void* workspace, size_t ws_bytes, void* stream_) {
  const char* who = "splade_peer_loss_bwd";
  int rc = check_group(group, who);
  if (rc != kOk) return rc;
  if (!colmean_d || !ddocs_local || !g_loss) return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  const int W = group->world;
  if (Bq != group->Bq || V != group->V || k < 1)
    return set_error(kErrInvalidArg, "%s: shapes (Bq=%d, V=%d, k=%d) do not match the exchange windows (Bq=%d, V=%d)",
                     who, Bq, V, k, group->Bq, group->V);
  if (V % 4 != 0) return set_error(kErrUnsupportedShape, "%s: V %% 4 == 0 is required (V=%d)", who, V);
  if (epoch == 0) return set_error(kErrInvalidArg, "%s: epochs start at 1", who);
  const size_t need = peer_loss_bwd_workspace_bytes(Bq, W);
  if (!workspace || ws_bytes < need)
    return set_error(kErrWorkspace, "%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need);
  // everything loss_bwd_dq_launch would refuse is refused HERE, before the first enqueue: a rank must never
  // publish its gl / signal its peers and then bail out (they would be left to time out)
  if (!q || !docs_all || !negs || !g_hard || !colmean_q || !colmean_n || !dq || !dnegs)
    return set_error(kErrInvalidArg, "%s: null pointer argument", who);
  {
    auto aligned16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
    if (!aligned16(q) || !aligned16(docs_all) || !aligned16(negs) || !aligned16(dq) || !aligned16(dnegs) ||
        !aligned16(colmean_q) || !aligned16(colmean_n) || !aligned16(colmean_d) || !aligned16(ddocs_local) ||
        !aligned16(q16) || !aligned16(d16_all))
      return set_error(kErrUnsupportedShape, "%s: 16-byte aligned rows are required", who);
  }
  rc = require_sm100(who);
  if (rc != kOk) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const PeerLayout L = peer_layout(Bq, V, W);
  const int b = static_cast<int>(epoch & 1u);
  uint8_t* mine = static_cast<uint8_t*>(group->window[group->rank]);

  SPLADE_CUDA_OK(cudaMemcpyAsync(mine + L.gl[b], g_loss, sizeof(float), cudaMemcpyDeviceToDevice, stream));
  const float* g_win = reinterpret_cast<const float*>(mine + L.g[b]);
  if (q16 && d16_all) {  // autocast mode: dq = G D on tcgen05, on the bf16 operands the forward packed
    uint8_t* tc_ws = static_cast<uint8_t*>(workspace) + al256(static_cast<size_t>(W) * Bq * Bq * sizeof(float));
    rc = loss_bwd_tc_launch(q, negs, q16, d16_all, Bq, W * Bq, k, V, group->rank * Bq, g_win, g_hard, colmean_q,
                            nullptr, colmean_n, cq, 0.f, cn, g_loss, dq, /*ddocs=*/nullptr, dnegs, tc_ws,
                            loss_bwd_tc_workspace_bytes(Bq, W * Bq), stream);
  } else {
    rc = loss_bwd_dq_launch(q, docs_all, negs, Bq, W * Bq, k, V, g_win, g_hard, colmean_q, colmean_n, cq, cn, g_loss,
                            dq, dnegs, stream);
  }
  if (rc != kOk) return rc;

  BwdParams p{};
  p.g = make_ptrs(group);
  p.q_off = L.q[b];
  p.g_off = L.g[b];
  p.gl_off = L.gl[b];
  p.Bq = Bq;
  p.V = V;
  p.epoch = epoch;
  p.timeout_ns = timeout_ns_from_env();
  p.gloc = static_cast<float*>(workspace);
  p.cm_d = colmean_d;
  p.cd = cd;
  p.ddocs = ddocs_local;
  p.round_bf16 = (q16 && d16_all) ? 1 : 0;
  peer_bwd_sync_kernel<<<W, 256, 0, stream>>>(p);
  // (the kernel's dynamic shared memory limit is raised once per device in splade_peer_alloc)
  const unsigned cx = static_cast<unsigned>((V + PCOLS - 1) / PCOLS);
  peer_ddocs_kernel<<<dim3(cx, static_cast<unsigned>((Bq + PD - 1) / PD)), 128, PDDOCS_SMEM, stream>>>(p);
  SPLADE_CUDA_OK(cudaGetLastError());
  return kOk;
}

}  // extern "C"
