// C-ABI of libsplade_b200.so (declared in include/splade_b200.h) + error / device plumbing.
//
// Contract: plain pointers and sizes, every call returns 0 or an error code and never
// throws or aborts; the message is available through splade_last_error() (thread-local, because DDP runs
// backward on a different host thread than forward); all work is enqueued on the caller's stream; the
// library owns no device memory.
#include "../../include/splade_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>

namespace splade {

namespace {
thread_local char g_last_error[512] = "";
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return kOk;
  return set_error(kErrCuda, "CUDA error %s (%s) at: %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
}

// Device properties are immutable: look them up once per device (these sit on every launch path).
namespace {
constexpr int kMaxDevices = 64;
std::atomic<int> g_cc_major[kMaxDevices];  // 0 = not queried yet
std::atomic<int> g_sm_count[kMaxDevices];
}  // namespace

int require_sm100(const char* who) {
  int dev = 0;
  SPLADE_CUDA_OK(cudaGetDevice(&dev));
  int major = (dev >= 0 && dev < kMaxDevices) ? g_cc_major[dev].load(std::memory_order_relaxed) : 0;
  if (major == 0) {
    SPLADE_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (dev >= 0 && dev < kMaxDevices) g_cc_major[dev].store(major, std::memory_order_relaxed);
  }
  if (major != 10) {
    int minor = 0;
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return set_error(kErrUnsupportedDevice,
                     "%s: device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", who, dev,
                     major, minor);
  }
  return kOk;
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < kMaxDevices) {
    n = g_sm_count[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
  }
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev >= 0 && dev < kMaxDevices) g_sm_count[dev].store(n, std::memory_order_relaxed);
  return n;
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}
}  // namespace

// tensor with a contiguous innermost dimension, 128-byte swizzle, out-of-bounds elements read as zero / not written.
static int encode_tmap_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* gptr, int rank,
                             const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(kErrDriver, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  const CUresult r = fn(out, dtype, static_cast<cuuint32_t>(rank),
                        const_cast<void*>(gptr), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(kErrDriver, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", static_cast<int>(r),
                     rank);
  return kOk;
}

int encode_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, gptr, rank, dims, strides_bytes, box);
}

int encode_tmap_f32(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, gptr, rank, dims, strides_bytes, box);
}

}  // namespace splade

using namespace splade;

extern "C" {

int splade_abi_version(void) { return SPLADE_B200_ABI_VERSION; }

const char* splade_last_error(void) { return g_last_error; }

size_t splade_workspace_bytes(int op, int B, int S, int H, int V, int k) {
  switch (op) {
    case SPLADE_OP_HEAD_FWD:
      return head_fwd_workspace_bytes(B, S, H, V, /*want_tokw=*/false);
    case SPLADE_OP_HEAD_FWD_TOKW:
      return head_fwd_workspace_bytes(B, S, H, V, /*want_tokw=*/true);
    case SPLADE_OP_HEAD_BWD:
      return head_bwd_workspace_bytes(B, S, V);
    case SPLADE_OP_FLOPS_NNZ:
      return flops_workspace_bytes(B, V);
    case SPLADE_OP_SCORES:
      return scores_workspace_bytes(B, k, V);
    case SPLADE_OP_V33_LOSS_FWD:
      return v33_loss_fwd_workspace_bytes(B, S, V, k);
    case SPLADE_OP_TRUNK_ENCODE:
      return trunk_workspace_bytes(B, S, H, k, V);
    case SPLADE_OP_TRUNK_GEMM:
      return 256;
    case SPLADE_OP_TRUNK_ATTN:
      return ((varlen_layout_bytes(B, S, 1) + 255) & ~static_cast<size_t>(255)) +
             static_cast<size_t>(S > 0 ? S : 1) * 32 * sizeof(float2);
    case SPLADE_OP_PEER_LOSS_BWD:
      return peer_loss_bwd_workspace_bytes(B, k);
    case SPLADE_OP_HEAD_BWD_DHIDDEN:
      return head_bwd_dhidden_workspace_bytes(B, S, V);
    case SPLADE_OP_HEAD_BWD_BITMAP:
      return head_bwd_bitmap_bytes(B, V);
    case SPLADE_OP_LOSS_BWD_TC:
      return loss_bwd_tc_workspace_bytes(B, k);
    case SPLADE_OP_HEAD_PROLOGUE:
      return head_prologue_workspace_bytes(B, H);
    case SPLADE_OP_TRUNK_TRAIN_FWD:
      return trunk_train_workspace_bytes(B, S, H, 1);
    case SPLADE_OP_TRUNK_LN_BWD:
      return trunk_ln_bwd_workspace_bytes(H);
    default:
      return 0;
  }
}

int splade_head_fwd(const void* hidden, const void* W, const float* bias, const int32_t* mask, int B, int S, int H,
                    int V, float* rep, int32_t* argmax, float* tokw, void* workspace, size_t ws_bytes,
                    void* stream) {
  return head_fwd_launch(hidden, W, bias, mask, B, S, H, V, rep, argmax, tokw, workspace, ws_bytes,
                         pack_mode_from_env(), nullptr, static_cast<cudaStream_t>(stream));
}

int splade_head_fwd_mode(const void* hidden, const void* W, const float* bias, const int32_t* mask, int B, int S,
                         int H, int V, float* rep, int32_t* argmax, float* tokw, void* workspace, size_t ws_bytes,
                         int pack_mode, int32_t* plan_out, void* stream) {
  return head_fwd_launch(hidden, W, bias, mask, B, S, H, V, rep, argmax, tokw, workspace, ws_bytes, pack_mode,
                         plan_out, static_cast<cudaStream_t>(stream));
}

int splade_head_fwd_host(const void* hidden_host, const int32_t* mask_host, const void* W, const float* bias, int B,
                         int S, int H, int V, void* hidden_dev, int32_t* mask_dev, float* rep_dev,
                         int32_t* argmax_dev, float* rep_host, int32_t* argmax_host, void* workspace,
                         size_t ws_bytes, void* stream) {
  if (!hidden_host || !mask_host || !hidden_dev || !mask_dev || !rep_dev || !argmax_dev)
    return set_error(kErrInvalidArg, "splade_head_fwd_host: null pointer argument");
  if (B < 0 || S < 1 || H < 1 || V < 1)
    return set_error(kErrInvalidArg, "splade_head_fwd_host: bad sizes B=%d S=%d H=%d V=%d", B, S, H, V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t tokens = static_cast<size_t>(B) * S;
  SPLADE_CUDA_OK(cudaMemcpyAsync(hidden_dev, hidden_host, tokens * H * 2, cudaMemcpyHostToDevice, st));
  SPLADE_CUDA_OK(cudaMemcpyAsync(mask_dev, mask_host, tokens * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const int rc = head_fwd_launch(hidden_dev, W, bias, mask_dev, B, S, H, V, rep_dev, argmax_dev, nullptr, workspace,
                                 ws_bytes, head_fwd_pack_mode_host(mask_host, B, S), nullptr, st);
  if (rc != kOk) return rc;
  const size_t out = static_cast<size_t>(B) * V;
  if (rep_host) SPLADE_CUDA_OK(cudaMemcpyAsync(rep_host, rep_dev, out * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (argmax_host)
    SPLADE_CUDA_OK(cudaMemcpyAsync(argmax_host, argmax_dev, out * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return kOk;
}

int splade_head_bwd(const float* g_rep, const float* rep, const int32_t* argmax, const void* hidden, const void* W,
                    int B, int S, int H, int V, float* d_hidden, float* dW, float* dbias, void* workspace,
                    size_t ws_bytes, void* stream) {
  return head_bwd_launch(g_rep, rep, argmax, hidden, W, B, S, H, V, d_hidden, dW, dbias, workspace, ws_bytes,
                         static_cast<cudaStream_t>(stream));
}

int splade_rep_bitmap(const float* rep, int B, int V, uint32_t* bitmap, void* stream) {
  return rep_bitmap_launch(rep, B, V, bitmap, static_cast<cudaStream_t>(stream));
}

size_t splade_head_bwd_multi_workspace_bytes(const splade_head_bwd_segment* segs, int n_segs, int V) {
  return head_bwd_multi_workspace_bytes(segs, n_segs, V);
}

int splade_head_bwd_multi(const splade_head_bwd_segment* segs, int n_segs, const void* W, int H, int V, float* dW,
                          float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream) {
  return head_bwd_multi_launch(segs, n_segs, W, H, V, dW, dbias, accumulate, workspace, ws_bytes,
                               static_cast<cudaStream_t>(stream));
}

int splade_head_bwd_dhidden(const splade_head_bwd_segment* seg, const void* W, int H, int V, uint32_t* bitmap,
                            void* workspace, size_t ws_bytes, void* stream) {
  return head_bwd_dhidden_launch(seg, W, H, V, bitmap, workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}

int splade_head_bwd_last(const splade_head_bwd_segment* last, const splade_head_bwd_segment* segs,
                         const uint32_t* const* bitmaps, int n_segs, const void* W, int H, int V, uint32_t* bitmap,
                         float* dW, float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream) {
  return head_bwd_last_launch(last, segs, bitmaps, n_segs, W, H, V, bitmap, dW, dbias, accumulate, workspace, ws_bytes,
                              static_cast<cudaStream_t>(stream));
}

int splade_head_bwd_dw(const splade_head_bwd_segment* segs, const uint32_t* const* bitmaps, int n_segs, int H, int V,
                       float* dW, float* dbias, int accumulate, void* stream) {
  return head_bwd_dw_launch(segs, bitmaps, n_segs, H, V, dW, dbias, accumulate, static_cast<cudaStream_t>(stream));
}

int splade_flops_nnz(const float* rep, int N, int V, float* colmean, float* out2, void* workspace, size_t ws_bytes,
                     void* stream) {
  return flops_nnz_launch(rep, N, V, colmean, out2, workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}

int splade_scores(const float* q, int Bq, const float* docs, int N, int V, int round_bf16, float* raw,
                  void* workspace, size_t ws_bytes, void* stream) {
  return scores_launch(q, Bq, docs, N, V, round_bf16, raw, workspace, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

int splade_rowdot(const float* q, const float* docs, int Bq, int k, int V, float* out, void* stream) {
  return rowdot_launch(q, docs, Bq, k, V, out, static_cast<cudaStream_t>(stream));
}

int splade_v33_scalars(const splade_v33_scalar_args* args, void* stream) {
  return v33_scalars_launch(args, static_cast<cudaStream_t>(stream));
}

int splade_v33_loss_fwd(const splade_v33_fwd_args* args, void* stream) {
  return v33_loss_fwd_launch(args, static_cast<cudaStream_t>(stream));
}

int splade_loss_bwd(const float* q, const float* docs, const float* negs, int Bq, int N, int k, int V,
                    int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                    const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                    const float* g_loss, float* dq, float* ddocs, float* dnegs, void* stream) {
  return loss_bwd_launch(q, docs, negs, Bq, N, k, V, label_offset, g_inbatch, g_hard, colmean_q, colmean_d,
                         colmean_n, cq, cd, cn, g_loss, dq, ddocs, dnegs, static_cast<cudaStream_t>(stream));
}

int splade_v33_loss_fwd_operands(const splade_v33_fwd_args* args, const void** q16, const void** d16, int* Vp) {
  if (!args || !q16 || !d16 || !Vp || !args->workspace)
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd_operands: null pointer argument");
  if (!args->round_bf16)
    return set_error(kErrInvalidArg, "splade_v33_loss_fwd_operands: the forward packs bf16 operands in autocast mode only");
  void* sws = v33_loss_fwd_scores_ws(args);
  *q16 = sws;
  *d16 = scores_tc_docs16(sws, args->Bq, args->N, args->V, Vp);
  return kOk;
}

int splade_loss_bwd_tc(const float* q, const float* negs, const void* q16, const void* d16, int Bq, int N, int k, int V,
                       int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                       const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                       const float* g_loss, float* dq, float* ddocs, float* dnegs, void* workspace, size_t ws_bytes,
                       void* stream) {
  return loss_bwd_tc_launch(q, negs, q16, d16, Bq, N, k, V, label_offset, g_inbatch, g_hard, colmean_q, colmean_d,
                            colmean_n, cq, cd, cn, g_loss, dq, ddocs, dnegs, workspace, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}

int splade_trunk_encode(const splade_trunk_args* args, void* stream) {
  return trunk_encode_launch(args, static_cast<cudaStream_t>(stream));
}

size_t splade_packed_rows(int B, int S) { return varlen_rows_public(B, S); }

int splade_trunk_train_fwd(const splade_trunk_args* args, const splade_trunk_train_saves* saves, void* stream) {
  return trunk_train_fwd_launch(args, saves, static_cast<cudaStream_t>(stream));
}

int splade_trunk_ln_bwd(const void* dy, int dy_is_f32, const float* x, const float* gamma, float eps, int rows, int H,
                        float* dx, int accumulate, void* dx_bf16, float* dgamma, void* workspace, size_t ws_bytes,
                        void* stream) {
  int rc = require_sm100("splade_trunk_ln_bwd");
  if (rc != kOk) return rc;
  return trunk_ln_bwd_launch(dy, dy_is_f32, x, gamma, eps, rows, H, dx, accumulate, dx_bf16, dgamma, workspace, ws_bytes,
                             static_cast<cudaStream_t>(stream));
}

int splade_trunk_dqkv_pack(const void* dq, const void* dk, const void* dv, int64_t stride_b, int64_t stride_h,
                           int64_t stride_s, const int64_t* row_of, const float* cos_t, const float* sin_t, int B, int S,
                           int n_heads, int64_t rows, void* dqkv, void* stream) {
  return trunk_dqkv_pack_launch(dq, dk, dv, stride_b, stride_h, stride_s, reinterpret_cast<const long long*>(row_of),
                                cos_t, sin_t, B, S, n_heads, rows, dqkv, static_cast<cudaStream_t>(stream));
}

int splade_trunk_gather_rows(const void* packed, const int64_t* row_of, int64_t n_tokens, int row_bytes, int64_t rows,
                             void* padded, void* stream) {
  return trunk_gather_rows_launch(packed, reinterpret_cast<const long long*>(row_of), n_tokens, row_bytes, rows, padded,
                                  static_cast<cudaStream_t>(stream));
}

int splade_trunk_geglu_bwd(const void* h, const void* dgg, int rows, int I, void* dh, void* stream) {
  int rc = require_sm100("splade_trunk_geglu_bwd");
  if (rc != kOk) return rc;
  return trunk_geglu_bwd_launch(h, dgg, rows, I, dh, static_cast<cudaStream_t>(stream));
}

int splade_trunk_gemm(int epi, const void* A, int rows, const void* W, int N, int K, void* out, int ldo,
                      void* workspace, size_t ws_bytes, void* stream) {
  if (!A || !W || !out || rows < 1) return set_error(kErrInvalidArg, "splade_trunk_gemm: bad argument");
  if (!workspace || ws_bytes < 256) return set_error(kErrWorkspace, "splade_trunk_gemm: workspace too small");
  int rc = require_sm100("splade_trunk_gemm");
  if (rc != kOk) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* t_rows = static_cast<int32_t*>(workspace);
  SPLADE_CUDA_OK(cudaMemcpyAsync(t_rows, &rows, sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (epi == EPI_QKV_ROPE) return set_error(kErrInvalidArg, "splade_trunk_gemm: epilogue 4 (RoPE) is internal to splade_trunk_encode");
  return trunk_gemm_launch(epi, A, rows, W, N, K, t_rows, out, ldo, nullptr, 0.f, 0, st);
}

int splade_head_prologue_fwd(const void* x, const void* dense_w, const float* norm_w, float eps, int T, int rows_cap,
                             int H, void* y, void* hidden, float* stats, void* workspace, size_t ws_bytes,
                             void* stream) {
  int rc = require_sm100("splade_head_prologue_fwd");
  if (rc != kOk) return rc;
  return head_prologue_fwd_launch(x, rows_cap, dense_w, norm_w, eps, T, H, y, hidden, stats, workspace, ws_bytes,
                                  static_cast<cudaStream_t>(stream));
}

int splade_head_prologue_bwd(const float* d_hidden, const void* y, const float* stats, const float* norm_w,
                             const void* dense_w_t, int T, int rows_cap, int H, void* dy, void* dx, float* d_norm_w,
                             void* workspace, size_t ws_bytes, void* stream) {
  int rc = require_sm100("splade_head_prologue_bwd");
  if (rc != kOk) return rc;
  return head_prologue_bwd_launch(d_hidden, y, stats, norm_w, dense_w_t, rows_cap, T, H, dy, dx, d_norm_w, workspace,
                                  ws_bytes, static_cast<cudaStream_t>(stream));
}

int splade_trunk_attn(void* qkv_packed, const int32_t* mask, int B, int S, int H, int n_heads, int window,
                      float rope_theta, void* out_packed, uint32_t* meta_out, void* workspace, size_t ws_bytes,
                      void* stream) {
  if (!qkv_packed || !mask || !out_packed) return set_error(kErrInvalidArg, "splade_trunk_attn: null pointer");
  int rc = require_sm100("splade_trunk_attn");
  if (rc != kOk) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t lay_bytes = (varlen_layout_bytes(B, S, 1) + 255) & ~static_cast<size_t>(255);
  const size_t need = lay_bytes + static_cast<size_t>(S) * 32 * sizeof(float2);
  if (!workspace || ws_bytes < need) return set_error(kErrWorkspace, "splade_trunk_attn: workspace too small");
  VarlenLayout lay{};
  rc = varlen_layout_build(mask, B, S, 1, workspace, lay_bytes, &lay, st);
  if (rc != kOk) return rc;
  if (rope_theta > 0.f) {  // the trunk's own RoPE kernel, in place on the q and k thirds
    float2* rope = reinterpret_cast<float2*>(static_cast<uint8_t*>(workspace) + lay_bytes);
    rc = trunk_rope_table_launch(rope, S, rope_theta, st);
    if (rc != kOk) return rc;
    rc = trunk_rope_launch(qkv_packed, rope, lay, H, st);
    if (rc != kOk) return rc;
  }
  if (meta_out)
    SPLADE_CUDA_OK(cudaMemcpyAsync(meta_out, lay.meta, lay.rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  return trunk_attn_launch(qkv_packed, out_packed, lay, B, S, H, n_heads, window, st);
}

int splade_topk(const float* rep, int B, int V, const uint8_t* blocklist, int k, int32_t* ids, float* vals,
                int32_t* counts, void* stream) {
  return topk_launch(rep, B, V, blocklist, k, ids, vals, counts, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
