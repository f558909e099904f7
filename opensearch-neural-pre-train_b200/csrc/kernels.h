// Internal launcher declarations (host side). The public C-ABI is include/splade_b200.h.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct splade_v33_scalar_args;
struct splade_v33_fwd_args;
struct splade_trunk_args;
struct splade_head_bwd_segment;
struct splade_trunk_train_saves;

namespace splade {

// api.cu
int require_sm100(const char* who);
int sm_count();
int encode_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);
int encode_tmap_f32(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);

// head_fwd.cu
void head_fwd_plan(int B, int S, int* s_box, int* b_box, int* chunks_per_doc, int* n_tt, int* umma_n);
size_t head_fwd_workspace_bytes(int B, int S, int H, int V, bool want_tokw);
int pack_mode_from_env();
int head_fwd_pack_mode_host(const int32_t* mask_host, int B, int S);
// pack_mode: -1 decide on the device, 0 padded tiling, 1 packed (variable-length) tiling; plan_out: optional
// device int32[4] = {mode chosen, real tokens, packed tiles, UMMA N of the last packed tile}
int head_fwd_launch(const void* hidden, const void* W, const float* bias, const int32_t* mask, int B, int S,
                    int H, int V, float* rep, int32_t* argmax, float* tokw, void* workspace, size_t ws_bytes,
                    int pack_mode, int32_t* plan_out, cudaStream_t stream);

// Packed (variable-length) token layout shared by the head and the native trunk (head_fwd.cu)
struct VarlenPlan {
  int32_t mode;      // 0: padded kernel runs, 1: packed kernel runs
  int32_t T;         // packed columns in the batch (real tokens + replicas)
  int32_t n_tt;      // packed token tiles = ceil(T / 256)
  int32_t n16_last;  // UMMA N of the last packed tile
  int32_t T_real;    // real tokens in the batch
};
struct VarlenLayout {
  VarlenPlan* plan;   // device
  int32_t* len;       // [B] real tokens per document
  int32_t* off;       // [B+1] first packed row of every document (multiples of 8)
  uint8_t* spans;     // [B]
  uint32_t* meta;     // [rows] (straddles << 31) | (doc << 16) | position
  uint32_t* keys;     // [B,V] merge buffer of straddling documents
  size_t rows;        // rows of the packed buffers (a whole number of 256-row tiles)
};
size_t varlen_layout_bytes(int B, int S, int V);
int varlen_layout_build(const int32_t* mask, int B, int S, int V, void* workspace, size_t ws_bytes,
                        VarlenLayout* out, cudaStream_t stream);
int head_fwd_packed_launch(const void* hidden_packed, const VarlenLayout& lay, const void* W, const float* bias,
                           int B, int S, int H, int V, float* rep, int32_t* argmax, cudaStream_t stream);

// trunk_gemm.cu: tcgen05 GEMM with the trunk's fused epilogues
enum TrunkEpilogue { EPI_BF16 = 0, EPI_GELU = 1, EPI_RESID = 2, EPI_GEGLU = 3, EPI_QKV_ROPE = 4 };
int trunk_gemm_launch(int epi, const void* A, int rows_cap, const void* W, int N, int K, const int32_t* t_rows,
                      void* out, int ldo, const uint32_t* meta, float rope_theta, int rope_cols,
                      cudaStream_t stream);

// trunk.cu: embedding / LayerNorm / attention kernels and the layer loop
struct TrunkAttnArgs;
int trunk_rope_table_launch(float2* table, int max_pos, float theta, cudaStream_t stream);
int trunk_embed_ln_launch(const int64_t* input_ids, const float* tok_emb, const float* gamma, float eps,
                          const VarlenLayout& lay, int S, int H, float* x, void* xb, cudaStream_t stream);
int trunk_ln_launch(const void* in, int in_is_bf16, const float* gamma, float eps, const int32_t* t_rows,
                    int rows_cap, int H, void* out_bf16, cudaStream_t stream);
int trunk_rope_launch(void* qkv, const float2* rope, const VarlenLayout& lay, int H, cudaStream_t stream);
int trunk_attn_launch(const void* qkv, void* out, const VarlenLayout& lay, int B, int S, int H, int n_heads,
                      int window, cudaStream_t stream);

size_t trunk_workspace_bytes(int B, int S, int H, int I, int V);
int trunk_encode_launch(const struct splade_trunk_args* a, cudaStream_t stream);
size_t varlen_rows_public(int B, int S);
size_t trunk_train_workspace_bytes(int B, int S, int H, int V);
int trunk_train_fwd_launch(const struct splade_trunk_args* a, const struct splade_trunk_train_saves* saves,
                           cudaStream_t stream);

// trunk_bwd.cu: row-wise backward kernels of the trunk's training path
size_t trunk_ln_bwd_workspace_bytes(int H);
int trunk_ln_bwd_launch(const void* dy, int dy_is_f32, const float* x, const float* gamma, float eps, int rows, int H,
                        float* dx, int accumulate, void* dx_bf16, float* dgamma, void* workspace, size_t ws_bytes,
                        cudaStream_t stream);
int trunk_geglu_bwd_launch(const void* h, const void* dgg, int rows, int I, void* dh, cudaStream_t stream);
int trunk_gather_rows_launch(const void* packed, const long long* row_of, long long n_tok, int row_bytes, long long rows,
                             void* padded, cudaStream_t stream);
int trunk_dqkv_pack_launch(const void* dq, const void* dk, const void* dv, long long sb, long long sh, long long ss,
                           const long long* row_of, const float* cos_t, const float* sin_t, int B, int S, int nh,
                           long long rows, void* dqkv, cudaStream_t stream);

// head_prologue.cu: ModernBertPredictionHead (dense -> GELU -> LayerNorm) forward / backward on the training path
size_t head_prologue_workspace_bytes(int T, int H);
int head_prologue_fwd_launch(const void* x, int rows_cap, const void* dense_w, const float* norm_w, float eps, int T,
                             int H, void* y, void* hidden, float* stats, void* workspace, size_t ws_bytes,
                             cudaStream_t stream);
int head_prologue_bwd_launch(const float* d_hidden, const void* y, const float* stats, const float* norm_w,
                             const void* dense_w_t, int rows_cap, int T, int H, void* dy, void* dx, float* d_norm_w,
                             void* workspace, size_t ws_bytes, cudaStream_t stream);

// head_bwd.cu
size_t head_bwd_workspace_bytes(int B, int S, int V);
size_t head_bwd_bitmap_bytes(int B, int V);
int rep_bitmap_launch(const float* rep, int B, int V, uint32_t* bitmap, cudaStream_t stream);
size_t head_bwd_dhidden_workspace_bytes(int B, int S, int V);
size_t head_bwd_multi_workspace_bytes(const struct splade_head_bwd_segment* segs, int n_segs, int V);
int head_bwd_multi_launch(const struct splade_head_bwd_segment* segs, int n_segs, const void* W, int H, int V,
                          float* dW, float* dbias, int accumulate, void* workspace, size_t ws_bytes,
                          cudaStream_t stream);
int head_bwd_dhidden_launch(const struct splade_head_bwd_segment* seg, const void* W, int H, int V, uint32_t* bitmap,
                            void* workspace, size_t ws_bytes, cudaStream_t stream);
int head_bwd_last_launch(const struct splade_head_bwd_segment* last, const struct splade_head_bwd_segment* segs,
                         const uint32_t* const* bitmaps, int n_segs, const void* W, int H, int V, uint32_t* bitmap,
                         float* dW, float* dbias, int accumulate, void* workspace, size_t ws_bytes,
                         cudaStream_t stream);
int head_bwd_dw_launch(const struct splade_head_bwd_segment* segs, const uint32_t* const* bitmaps, int n_segs, int H,
                       int V, float* dW, float* dbias, int accumulate, cudaStream_t stream);
int head_bwd_launch(const float* g_rep, const float* rep, const int32_t* argmax, const void* hidden,
                    const void* W, int B, int S, int H, int V, float* d_hidden, float* dW, float* dbias,
                    void* workspace, size_t ws_bytes, cudaStream_t stream);

// loss.cu
size_t flops_workspace_bytes(int N, int V);
int flops_nnz_launch(const float* rep, int N, int V, float* colmean, float* out2, void* workspace,
                     size_t ws_bytes, cudaStream_t stream);
size_t scores_workspace_bytes(int Bq, int N, int V);
int scores_launch(const float* q, int Bq, const float* docs, int N, int V, int round_bf16, float* raw,
                  void* workspace, size_t ws_bytes, cudaStream_t stream, bool docs_prepacked = false);
int rowdot_launch(const float* q, const float* docs, int Bq, int k, int V, float* out, cudaStream_t stream);
int v33_scalars_launch(const struct splade_v33_scalar_args* args, cudaStream_t stream);
size_t v33_loss_fwd_workspace_bytes(int Bq, int N, int V, int k);
int v33_loss_fwd_launch(const struct splade_v33_fwd_args* args, cudaStream_t stream, bool docs_prepacked = false);
void* v33_loss_fwd_scores_ws(const struct splade_v33_fwd_args* args);
int loss_bwd_dq_launch(const float* q, const float* docs, const float* negs, int Bq, int N, int k, int V,
                       const float* g_inbatch, const float* g_hard, const float* colmean_q, const float* colmean_n,
                       float cq, float cn, const float* g_loss, float* dq, float* dnegs, cudaStream_t stream);
int loss_bwd_launch(const float* q, const float* docs, const float* negs, int Bq, int N, int k, int V,
                    int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                    const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                    const float* g_loss, float* dq, float* ddocs, float* dnegs, cudaStream_t stream);

// loss_bwd_tc.cu: tcgen05 backward of the score GEMM on the bf16 operands the forward packed
size_t loss_bwd_tc_workspace_bytes(int Bq, int N);
int loss_bwd_tc_launch(const float* q, const float* negs, const void* q16, const void* d16, int Bq, int N, int k, int V,
                       int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                       const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                       const float* g_loss, float* dq, float* ddocs, float* dnegs, void* workspace, size_t ws_bytes,
                       cudaStream_t stream);

// scores_tc.cu (tcgen05 path used when round_bf16 != 0)
size_t scores_tc_workspace_bytes(int Bq, int N, int V);
int scores_tc_launch(const float* q, int Bq, const float* docs, int N, int V, void* workspace, size_t ws_bytes,
                     float** partial_out, int* ksplit_out, bool docs_prepacked, cudaStream_t stream,
                     bool q_prepacked = false);
void* scores_tc_docs16(void* workspace, int Bq, int N, int V, int* Vp_out);

// peer.cu: cross-GPU negatives over NVLink peer memory (include/splade_b200.h, splade_peer_*)
struct PeerGroupView;
size_t peer_window_bytes(int Bq, int V, int world);
size_t peer_loss_bwd_workspace_bytes(int Bq, int world);

// topk.cu
int topk_launch(const float* rep, int B, int V, const uint8_t* blocklist, int k, int32_t* ids, float* vals,
                int32_t* counts, cudaStream_t stream);

}  // namespace splade
