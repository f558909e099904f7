/*
 * splade_b200.h — C-ABI of libsplade_b200.so: the B200-native (sm_100a) SPLADE-max head + SPLADELossV33
 * hot path of opensearch-neural-pre-train.
 *
 * The reference has no FFI: its "operator interface" for this path is two Python nn.Modules
 * (src/model/splade_modern.py:19-114, src/model/losses.py:14-301).  Every entry point below replaces one
 * slice of those modules' forward/backward and is what a ctypes / cffi / pybind stub on the reference
 * side binds.  Conventions for all calls:
 *
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *   - row-major, innermost dimension contiguous; "bf16" is passed as const void* (raw 16-bit words);
 *   - the caller (PyTorch) owns every buffer, including workspaces sized by splade_workspace_bytes();
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation inside;
 *   - return value 0 = success, otherwise an error code; the message is splade_last_error()
 *     (thread-local).  Nothing throws or aborts across this boundary;
 *   - unsupported device (not sm_100) or shape is an ERROR, never a fallback.
 */
#ifndef SPLADE_B200_H_
#define SPLADE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPLADE_B200_ABI_VERSION 5

/* error codes */
#define SPLADE_OK 0
#define SPLADE_ERR_INVALID_ARG 1
#define SPLADE_ERR_UNSUPPORTED_SHAPE 2
#define SPLADE_ERR_UNSUPPORTED_DEVICE 3
#define SPLADE_ERR_WORKSPACE 4
#define SPLADE_ERR_CUDA 5
#define SPLADE_ERR_DRIVER 6

/* op ids for splade_workspace_bytes */
#define SPLADE_OP_HEAD_FWD 0      /* (B,S,H,V)            head forward, token_weights not requested   */
#define SPLADE_OP_HEAD_FWD_TOKW 1 /* (B,S,H,V)            head forward with token_weights             */
#define SPLADE_OP_HEAD_BWD 2      /* (B,S,H,V)            head backward                               */
#define SPLADE_OP_FLOPS_NNZ 3     /* (B=N rows, V)        FLOPS regulariser + nonzero count           */
#define SPLADE_OP_SCORES 4        /* (B=Bq, k=N docs, V)  q.d score matrix                            */
#define SPLADE_OP_V33_LOSS_FWD 5  /* (B=Bq, S=N docs, V, k) whole SPLADELossV33 forward               */
#define SPLADE_OP_TRUNK_ENCODE 6  /* (B,S,H,V, k = intermediate size) native trunk + head encode      */
#define SPLADE_OP_TRUNK_GEMM 7    /* (-)                      splade_trunk_gemm scratch               */
#define SPLADE_OP_TRUNK_ATTN 8    /* (B,S,-,V=1)              splade_trunk_attn scratch               */
#define SPLADE_OP_PEER_LOSS_BWD 9 /* (B=Bq, k=world)          splade_peer_loss_bwd scratch            */
#define SPLADE_OP_HEAD_BWD_DHIDDEN 10 /* (B,S,-,V)            splade_head_bwd_dhidden scratch (one segment)   */
#define SPLADE_OP_HEAD_BWD_BITMAP 11  /* (B,-,-,V)            bytes of one segment's activity bitmap          */
#define SPLADE_OP_LOSS_BWD_TC 12      /* (B=Bq, k=N docs)     splade_loss_bwd_tc scratch                      */
#define SPLADE_OP_HEAD_PROLOGUE 13    /* (B=T rows, H)        splade_head_prologue_fwd / _bwd scratch         */
#define SPLADE_OP_TRUNK_TRAIN_FWD 14  /* (B,S,H)              splade_trunk_train_fwd scratch                  */
#define SPLADE_OP_TRUNK_LN_BWD 15     /* (-,-,H)              splade_trunk_ln_bwd scratch                     */

int splade_abi_version(void);
const char* splade_last_error(void);
size_t splade_workspace_bytes(int op, int B, int S, int H, int V, int k);

/*
 * Fused vocabulary projection + log1p(relu) + padding mask + max-pool over the sequence (+ argmax).
 * Replaces src/model/splade_modern.py:69-86 (decoder Linear of transformers' ModernBertForMaskedLM,
 * modeling_modernbert.py:562, included).  Numerics = the reference under torch.autocast(bf16):
 * fp32 accumulate, + bias, one RNE rounding to bf16, relu, max with ties to the lowest position.
 *
 *   hidden [B,S,H] bf16   output of model.head (dense->GELU->LayerNorm), cast to bf16
 *   W      [V,H]   bf16   decoder.weight (tied to the token embeddings)
 *   bias   [V]     fp32   decoder.bias (pass it pre-rounded to bf16 to mirror autocast)
 *   mask   [B,S]   int32  attention mask, nonzero = real token
 *   rep    [B,V]   fp32   sparse_repr                       (out)
 *   argmax [B,V]   int32  position of the max, 0 if rep==0  (out; what autograd keeps in the reference)
 *   tokw   [B,S]   fp32   token_weights, or NULL to skip    (out)
 * H must be a multiple of 64; S <= 65535.
 */
int splade_head_fwd(const void* hidden, const void* W, const float* bias, const int32_t* mask, int B, int S, int H,
                    int V, float* rep, int32_t* argmax, float* tokw, void* workspace, size_t ws_bytes,
                    void* stream);

/*
 * splade_head_fwd with an explicit token-layout policy.  Right-padded batches (the reference's collator pads every
 * batch to its longest member, src/train/data/dataloader.py:46-164; queries are ~10 real tokens in S=64) spend
 * tensor-core time on padding.  The packed layout gathers the REAL tokens of the batch back to back ([T,H],
 * T = sum of lengths) and tiles those; results are bit-identical to the padded layout.
 *   pack_mode  SPLADE_PACK_AUTO: a device-side plan compares the cost of both tilings and picks one - no host
 *              synchronisation, both kernels are enqueued and the one not chosen exits immediately;
 *              SPLADE_PACK_OFF / SPLADE_PACK_ON pin the padded / packed tiling.
 *   plan_out   optional DEVICE int32[4]: {tiling used (0 padded, 1 packed), real tokens T, packed 256-token
 *              tiles, UMMA N of the last packed tile}; all zero when the packed path was not considered.
 * splade_head_fwd itself uses the policy in the environment variable SPLADE_B200_PACK (auto | 0 | 1, default
 * auto).  The packed path needs B <= 32767 and B*S <= 2^30; other shapes silently use the padded tiling.
 */
#define SPLADE_PACK_AUTO (-1)
#define SPLADE_PACK_OFF 0
#define SPLADE_PACK_ON 1
int splade_head_fwd_mode(const void* hidden, const void* W, const float* bias, const int32_t* mask, int B, int S,
                         int H, int V, float* rep, int32_t* argmax, float* tokw, void* workspace, size_t ws_bytes,
                         int pack_mode, int32_t* plan_out, void* stream);

/*
 * Host-buffer form of splade_head_fwd for callers whose activations live in host memory (and for the end-to-end
 * benchmark): enqueues  H2D(hidden_host, mask_host) -> splade_head_fwd -> D2H(rep, argmax)  on `stream` and
 * returns without synchronising.  W and bias are model weights and stay resident on the device.  The caller owns
 * the device staging buffers (hidden_dev [B,S,H] bf16, mask_dev [B,S] int32, rep_dev [B,V] fp32, argmax_dev [B,V]
 * int32, workspace) and must keep the host buffers alive (ideally pinned) until the stream has drained.
 * rep_host / argmax_host may be NULL to skip that copy.  Use two slots on two streams to overlap the copies of
 * one batch with the compute of the other.
 */
int splade_head_fwd_host(const void* hidden_host, const int32_t* mask_host, const void* W, const float* bias, int B,
                         int S, int H, int V, void* hidden_dev, int32_t* mask_dev, float* rep_dev,
                         int32_t* argmax_dev, float* rep_host, int32_t* argmax_host, void* workspace,
                         size_t ws_bytes, void* stream);

/*
 * Gather-driven sparse backward of splade_head_fwd (replaces autograd of splade_modern.py:69-83, invoked at
 * src/train/cli/train_v33_ddp.py:364).  dlogit[b,argmax[b,v],v] = g_rep[b,v] * exp(-rep[b,v]) for rep>0.
 *   d_hidden [B,S,H] fp32, dW [V,H] fp32, dbias [V] fp32 are fully overwritten (dense, DDP-ready).
 */
int splade_head_bwd(const float* g_rep, const float* rep, const int32_t* argmax, const void* hidden, const void* W,
                    int B, int S, int H, int V, float* d_hidden, float* dW, float* dbias, void* workspace,
                    size_t ws_bytes, void* stream);

/*
 * The same backward for SEVERAL head calls that share one decoder matrix.  The reference trainer runs the model
 * three times per micro-step (query, positive, negatives: src/train/cli/train_v33_ddp.py:339-343) and
 * back-propagates once (:364); autograd then sums three dense [V,H] gradients of the tied decoder / embedding
 * matrix.  Here each head call is a "segment" and dW / dbias are produced by ONE sweep over the vocabulary that
 * walks all segments (fixed order: segments as given, documents ascending => run-to-run bit-identical).
 *   accumulate != 0: dW / dbias are ADDED to the buffers' contents (e.g. straight into tok_embeddings.weight.grad
 *                    on the later micro-steps of a gradient-accumulation window) instead of overwriting them.
 *   d_hidden may be NULL for a segment whose input needs no gradient.
 * splade_head_bwd_multi does everything in three launches.  The two-phase form serves autograd, which delivers the
 * three g_rep one after the other: splade_head_bwd_dhidden (per segment, as its g_rep arrives; uses segment.active
 * or, when that is NULL, writes the activity bitmap uint32 [B, ceil(V/32)] into `bitmap`) and, after the last one,
 * splade_head_bwd_dw over all segments and their bitmaps.  The backward kernels are DRIVEN by the bitmap: in the
 * trained (sparse) regime g_rep / rep / argmax are read at the active terms only.
 * Workspaces: splade_head_bwd_multi_workspace_bytes; SPLADE_OP_HEAD_BWD_DHIDDEN via splade_workspace_bytes.
 */
#define SPLADE_HEAD_BWD_MAX_SEGMENTS 8
typedef struct splade_head_bwd_segment {
  const float* g_rep;     /* [B,V] fp32  gradient of sparse_repr                 */
  const float* rep;       /* [B,V] fp32  sparse_repr saved from the forward      */
  const int32_t* argmax;  /* [B,V] int32 saved from the forward                  */
  const void* hidden;     /* [B,S,H] bf16 the forward's input                    */
  float* d_hidden;        /* [B,S,H] fp32 out, or NULL                           */
  int32_t B, S;
  const uint32_t* active; /* optional [B, ceil(V/32)]: bit j of word w = (rep[b, 32w+j] > 0), as splade_rep_bitmap
                             writes it (cheapest right after the forward, rep is L2-hot); NULL = computed here */
} splade_head_bwd_segment;
int splade_rep_bitmap(const float* rep, int B, int V, uint32_t* bitmap, void* stream);
size_t splade_head_bwd_multi_workspace_bytes(const splade_head_bwd_segment* segs, int n_segs, int V);
int splade_head_bwd_multi(const splade_head_bwd_segment* segs, int n_segs, const void* W, int H, int V, float* dW,
                          float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream);
int splade_head_bwd_dhidden(const splade_head_bwd_segment* seg, const void* W, int H, int V, uint32_t* bitmap,
                            void* workspace, size_t ws_bytes, void* stream);
int splade_head_bwd_dw(const splade_head_bwd_segment* segs, const uint32_t* const* bitmaps, int n_segs, int H, int V,
                       float* dW, float* dbias, int accumulate, void* stream);
/* The last head node of a backward pass: splade_head_bwd_dhidden for `last` and splade_head_bwd_dw over all `segs`
 * (which include `last`) in one call; the sweep runs CONCURRENTLY with the bucket sort / dHidden kernels on a side
 * stream forked from and joined back into `stream` (as splade_head_bwd_multi does; SPLADE_B200_BWD_CONCURRENT=0
 * serialises them). */
int splade_head_bwd_last(const splade_head_bwd_segment* last, const splade_head_bwd_segment* segs,
                         const uint32_t* const* bitmaps, int n_segs, const void* W, int H, int V, uint32_t* bitmap,
                         float* dW, float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream);

/*
 * FLOPS regulariser (src/model/losses.py:57-73) and nonzero monitoring (losses.py:272-273) in one pass.
 *   rep [N,V] fp32 -> colmean [V] (kept for backward), out2[0] = sum_j colmean_j^2, out2[1] = mean_i #(rep_i > 0)
 */
int splade_flops_nnz(const float* rep, int N, int V, float* colmean, float* out2, void* workspace, size_t ws_bytes,
                     void* stream);

/*
 * raw[i,j] = sum_v q[i,v] * docs[j,v]  (torch.mm(anchor, positive.t()), losses.py:155,243).
 * round_bf16 != 0 mirrors autocast: operands are rounded to bf16 first (fp32 accumulate); the output
 * rounding is applied by splade_v33_scalars.  Deterministic split-K (fixed-order reduction).
 */
int splade_scores(const float* q, int Bq, const float* docs, int N, int V, int round_bf16, float* raw,
                  void* workspace, size_t ws_bytes, void* stream);

/* out[i,j] = sum_v q[i,v] * docs[i,j,v], fp32 (hard-negative scores losses.py:161-169, margins :115-130). */
int splade_rowdot(const float* q, const float* docs, int Bq, int k, int V, float* out, void* stream);

/* Scalar stage of SPLADELossV33.forward (losses.py:136-181, 209-269): InfoNCE cross-entropy, optional KL-KD and
 * MarginMSE on the tiny score matrices, total loss, and the gradients w.r.t. the raw dot products. */
typedef struct splade_v33_scalar_args {
  /* sizes */
  int32_t Bq;           /* local queries                                                    */
  int32_t N;            /* in-batch documents (Bq, or world*Bq with gathered negatives)      */
  int32_t k;            /* hard negatives per query (>=1)                                   */
  int32_t label_offset; /* positive of query i is document label_offset + i                 */
  int32_t round_bf16;   /* mirror autocast: raw_inbatch -> bf16 -> /T -> bf16               */
  int32_t has_kd;       /* teacher_scores present and lambda_kd > 0 (requires N == Bq)      */
  int32_t has_mmse;     /* teacher pos/neg scores present and lambda_margin_mse > 0         */
  int32_t reserved;
  /* hyper-parameters (already scheduled on the host, losses.py:75-90) */
  float temperature, kd_temperature;
  float lambda_q, lambda_d, lambda_neg, lambda_kd, lambda_mmse;
  float reserved_f;
  /* inputs */
  const float* raw_inbatch; /* [Bq,N]  from splade_scores                                  */
  const float* raw_hard;    /* [Bq,k]  from splade_rowdot                                  */
  const float* raw_pos;     /* [Bq]    fp32 q_i.p_i (MarginMSE only) or NULL               */
  const float* teacher_kd;  /* [Bq,Bq] or NULL                                             */
  const float* teacher_pos; /* [Bq]    or NULL                                             */
  const float* teacher_neg; /* [Bq,k]  or NULL                                             */
  const float* flops3;      /* [6]: {flops_q,nnz_q, flops_d,nnz_d, flops_neg,nnz_neg}      */
  /* outputs */
  float* scalars;   /* [16]: 0 loss, 1 infonce, 2 flops_q, 3 flops_d, 4 flops_neg, 5 kd, 6 margin_mse,
                             7 nonzero_q, 8 nonzero_d                                       */
  float* g_inbatch; /* [Bq,N] d loss / d raw_inbatch                                       */
  float* g_hard;    /* [Bq,k] d loss / d raw_hard                                          */
  float* row_scratch; /* [3*Bq] per-row partial sums (workspace)                           */
} splade_v33_scalar_args;

int splade_v33_scalars(const splade_v33_scalar_args* args, void* stream);

/*
 * The whole SPLADELossV33 forward (losses.py:209-295) in ONE call: the three FLOPS/nonzero passes, the q.d score
 * GEMM, the hard-negative (and MarginMSE positive) row dots and the scalar stage are enqueued back to back on
 * `stream` - one host call and no intermediate allocation instead of ~10.  Same arithmetic as the fine-grained
 * entry points above (it calls the same launchers).
 */
typedef struct splade_v33_fwd_args {
  int32_t Bq, N, k, V;  /* queries, in-batch documents, hard negatives per query, vocabulary */
  int32_t label_offset, round_bf16, has_kd, has_mmse;
  float temperature, kd_temperature;
  float lambda_q, lambda_d, lambda_neg, lambda_kd, lambda_mmse;
  float reserved_f;
  const float* q;           /* [Bq,V]   */
  const float* docs;        /* [N,V]    positives; rows [label_offset, label_offset+Bq) are the local ones */
  const float* negs;        /* [Bq,k,V] */
  const float* teacher_kd;  /* [Bq,Bq] or NULL */
  const float* teacher_pos; /* [Bq]    or NULL */
  const float* teacher_neg; /* [Bq,k]  or NULL */
  float* colmean;           /* [3,V] out: column means of q, local docs, negatives (kept for backward) */
  float* scalars;           /* [16]  out, layout as in splade_v33_scalar_args */
  float* g_inbatch;         /* [Bq,N] out */
  float* g_hard;            /* [Bq,k] out */
  void* workspace;          /* splade_workspace_bytes(SPLADE_OP_V33_LOSS_FWD, Bq, N, 0, V, k) */
  size_t ws_bytes;
} splade_v33_fwd_args;

int splade_v33_loss_fwd(const splade_v33_fwd_args* args, void* stream);

/*
 * Backward of the scoring + FLOPS terms onto the sparse representations (autograd of losses.py:155-181,
 * 72-73).  With gl = *g_loss:
 *   dq[i,:]     = gl * ( sum_j g_inbatch[i,j] docs[j,:] + sum_k g_hard[i,k] negs[i,k,:] + cq * colmean_q )
 *   ddocs[j,:]  = gl * ( sum_i g_inbatch[i,j] q[i,:] + [label_offset <= j < label_offset+Bq] cd * colmean_d )
 *   dnegs[i,k,:]= gl * ( g_hard[i,k] q[i,:] + cn * colmean_n )
 * cq = 2*lambda_q/Bq, cd = 2*lambda_d/Bq, cn = 2*lambda_neg/(Bq*k).
 */
int splade_loss_bwd(const float* q, const float* docs, const float* negs, int Bq, int N, int k, int V,
                    int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                    const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                    const float* g_loss, float* dq, float* ddocs, float* dnegs, void* stream);

/*
 * The same backward in autocast mode on the 5th-generation tensor cores.  Under torch.autocast the reference's
 * backward of torch.mm(anchor, positive.t()) (src/model/losses.py:155) is a bf16 GEMM on the bf16 casts of its
 * operands; splade_v33_loss_fwd (round_bf16 != 0) already packed those casts for the score GEMM and leaves them in
 * its workspace: splade_v33_loss_fwd_operands returns where (q16 [Bq,Vp], d16 [N,Vp] bf16, Vp = V rounded up to 8) -
 * keep the forward workspace alive until the backward ran.  G is rounded to bf16, accumulation and outputs are
 * fp32.  ddocs may be NULL (dq and dnegs only).  Needs V % 4 == 0.
 */
int splade_v33_loss_fwd_operands(const splade_v33_fwd_args* args, const void** q16, const void** d16, int* Vp);
int splade_loss_bwd_tc(const float* q, const float* negs, const void* q16, const void* d16, int Bq, int N, int k, int V,
                       int label_offset, const float* g_inbatch, const float* g_hard, const float* colmean_q,
                       const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                       const float* g_loss, float* dq, float* ddocs, float* dnegs, void* workspace, size_t ws_bytes,
                       void* stream);

/*
 * Top-k pruning with the production emit semantics (benchmark/encoders.py:320-343): keep weight > 0, drop
 * block-listed term ids (special tokens and tokens starting with '[' or '<'), order by weight descending
 * with ties by ascending term id, keep the first k.
 *   rep [B,V] fp32, blocklist [V] uint8 (nonzero = drop) or NULL, k <= 1024
 *   ids [B,k] int32 (-1 padded), vals [B,k] fp32 (0 padded), counts [B] int32
 */
int splade_topk(const float* rep, int B, int V, const uint8_t* blocklist, int k, int32_t* ids, float* vals,
                int32_t* counts, void* stream);

/*
 * HOST-side emit (no device work): serialise compact top-k buffers (host copies of splade_topk's outputs) into
 * one JSON object per document, the value of the `sparse_embedding` field the reference indexes
 * (benchmark/encoders.py:320-343 -> benchmark/indexer.py:112-138).
 *   token_bytes/token_offsets [n_tokens+1]: concatenated UTF-8 token strings -> {"token": w} (rank_features);
 *   NULL -> {"<term id>": w} (sparse_vector, scripts/neural_sparse_search_aws.py:745-770);
 *   round_decimals >= 0 -> fixed decimals (the AWS script uses 4), < 0 -> shortest fp32 round-trip (%.9g).
 *   out [out_cap] receives the objects back to back; out_offsets [B+1] their byte ranges; *bytes_needed the total
 *   (call once with out_cap = 0 to size the buffer: returns SPLADE_ERR_WORKSPACE and the size).
 */
int splade_emit_json(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B, int k,
                     const char* token_bytes, const int64_t* token_offsets, int n_tokens, int round_decimals,
                     char* out, size_t out_cap, int64_t* out_offsets, size_t* bytes_needed);

/*
 * The emitters split large batches (>= 1024 documents, output buffer given) into contiguous document ranges
 * formatted on worker threads and concatenated in order - byte-identical to one thread.  n > 0 fixes the thread
 * count, 0 = automatic (env SPLADE_B200_EMIT_THREADS, else the host's cores, at most 16).  Returns the old value.
 */
int splade_emit_set_threads(int n);

/*
 * HOST-side: the body of an OpenSearch `_bulk` request for B documents - per document the action line and the
 * source line that opensearch-py's helpers.bulk serialises from the reference's actions
 *   benchmark/indexer.py:112-138            {"_index", "_id", "_source": {"doc_id", "content", "sparse_embedding"}}
 *   scripts/neural_sparse_search_aws.py:993-1013  "_source": {"title", "content", "category", "sparse_embedding"},
 *                                                 documents whose vector is empty are skipped (skip_empty != 0)
 * namely
 *   {"index":{"_index":"<index_name>","_id":"<id>"}}\n
 *   {<head>,"<field_name>":{"term":w,...}}\n
 * Document ids are UTF-8 strings (id_bytes / id_offsets [B+1]); `head` (head_bytes / head_offsets [B+1], or both
 * NULL) is the caller-serialised leading members of _source WITHOUT the braces (the per-document strings are the
 * caller's; the per-term work is done here).  The sparse object and the sizing protocol are splade_emit_json's;
 * out_offsets [B+1] delimit each document's two lines (an empty range for a skipped document).
 */
int splade_emit_bulk(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B, int k,
                     const char* token_bytes, const int64_t* token_offsets, int n_tokens, int round_decimals,
                     const char* index_name, const char* field_name, const char* id_bytes,
                     const int64_t* id_offsets, const char* head_bytes, const int64_t* head_offsets, int skip_empty,
                     char* out, size_t out_cap, int64_t* out_offsets, size_t* bytes_needed);

/*
 * HOST-side, query leg (BASELINE config 3): the search request bodies NeuralSparseSearcher.search builds from a
 * pruned query vector (benchmark/searchers.py:158-185) - one `rank_feature` should-clause per kept term, highest
 * weight first, at most max_clauses (the reference keeps 64) -
 *   {"size":<size>,"query":{"bool":{"should":[{"rank_feature":{"field":"<field_name>.<token>","boost":w}},...]}}}
 * for B queries from splade_topk's compact buffers (already ordered by weight desc, ties by term id asc = the
 * reference's stable sort).  msearch_index != NULL emits the `_msearch` form instead: a {"index":"<name>"} header
 * line before and a newline after every body.  A query with no surviving term produces no bytes (the reference
 * returns an empty response without calling OpenSearch).  Sizing protocol and out_offsets [B+1] as above.
 */
int splade_emit_search(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B, int k,
                       const char* token_bytes, const int64_t* token_offsets, int n_tokens, int max_clauses,
                       int size, const char* field_name, const char* msearch_index, char* out, size_t out_cap,
                       int64_t* out_offsets, size_t* bytes_needed);

/*
 * ---------------------------------------------------------------------------------------------------------------
 * Cross-GPU negatives without NCCL on the data path (north_star multi-GPU step).
 * The reference has local negatives only (src/model/losses.py:136-181 sees one rank's batch); the extension
 * scores every rank's queries against ALL ranks' positives.  Instead of all-gather -> GEMM and GEMM ->
 * reduce-scatter, the exchange is folded into the kernels themselves over NVLink peer memory:
 *
 *   forward : every rank publishes q and its positives in an "exchange window" (device memory it exports with
 *             CUDA IPC); after a flag barrier written straight into the peers' windows, ONE kernel pulls the
 *             other ranks' positives over NVLink, casts them to the bf16 operand of the tcgen05 score GEMM and
 *             keeps the fp32 copy for backward - the gathered fp32 tensor never takes a separate pass;
 *   backward: ddocs of the LOCAL positives only, = sum over ranks s of gl_s * G_s[:, my columns]^T q_s, computed
 *             by a kernel that reads q_s directly from the peers' windows (G_s, gl_s are pulled by the barrier
 *             kernel).  The [W*Bq, V] gradient of the gathered tensor and its reduce-scatter never exist.
 *
 * Windows are double-buffered by the parity of `epoch` (a counter the ranks advance together, first call = 1),
 * so one flag barrier per phase is enough.  A barrier that is not answered within the timeout
 * (env SPLADE_B200_PEER_TIMEOUT_MS, default 600000 - NCCL-like patience) sets the window's status word instead of hanging the GPU.
 * ---------------------------------------------------------------------------------------------------------------
 */
#define SPLADE_MAX_PEERS 16

typedef struct splade_peer_group {
  int32_t rank, world; /* world <= SPLADE_MAX_PEERS */
  int32_t Bq, V;       /* the shapes every window was sized for (splade_peer_window_bytes) */
  void* window[SPLADE_MAX_PEERS]; /* window[s] = base of rank s's window as mapped in THIS process */
} splade_peer_group;

size_t splade_peer_window_bytes(int Bq, int V, int world);
/* window memory: cudaMalloc'ed and zeroed by the library on the current device because CUDA IPC exports whole
 * allocations (the one exception to "PyTorch owns every buffer"); the caller frees it with splade_peer_free. */
int splade_peer_alloc(size_t bytes, void** dptr);
int splade_peer_free(void* dptr);
int splade_peer_export(const void* dptr, void* handle_out /* 64 bytes, cudaIpcMemHandle_t */);
int splade_peer_open(const void* handle /* 64 bytes from another process */, void** dptr_out);
int splade_peer_close(void* dptr);
/* synchronous read of the local window's status word: 0 = every barrier so far was answered */
int splade_peer_status(const splade_peer_group* group, int32_t* status_host);

/*
 * SPLADELossV33 forward with gathered positives.  `args` as for splade_v33_loss_fwd with args->docs = the
 * LOCAL positives [Bq,V] and args->N = Bq; labels are rank*Bq + i; args->g_inbatch is ignored (the [Bq, W*Bq]
 * gradient lives in the window, where the peers read it); args->workspace is sized by
 * splade_workspace_bytes(SPLADE_OP_V33_LOSS_FWD, Bq, world*Bq, 0, V, k).  KL-KD is not available (teacher
 * scores are [Bq,Bq]).  docs_all [world*Bq, V] fp32 (out) must be kept for splade_peer_loss_bwd.  V % 4 == 0.
 */
int splade_peer_v33_loss_fwd(const splade_v33_fwd_args* args, const splade_peer_group* group, uint32_t epoch,
                             float* docs_all, void* stream);

/*
 * Backward of the same (formulas of splade_loss_bwd); ddocs_local [Bq,V] is the gradient of the LOCAL positives
 * summed over every rank's queries.  workspace: splade_workspace_bytes(SPLADE_OP_PEER_LOSS_BWD, Bq,0,0,0, world).
 * q16 / d16_all: the bf16 operands the forward left in its workspace (splade_v33_loss_fwd_operands on the forward's
 * args; autocast mode) - dq then runs on the tensor cores (splade_loss_bwd_tc); both NULL = fp32 CUDA-core dq.
 */
int splade_peer_loss_bwd(const splade_peer_group* group, uint32_t epoch, const float* q, const float* docs_all,
                         const float* negs, int Bq, int k, int V, const float* g_hard, const float* colmean_q,
                         const float* colmean_d, const float* colmean_n, float cq, float cd, float cn,
                         const float* g_loss, float* dq, float* ddocs_local, float* dnegs, const void* q16,
                         const void* d16_all, void* workspace, size_t ws_bytes, void* stream);

/*
 * ---------------------------------------------------------------------------------------------------------------
 * Native variable-length ModernBERT trunk for inference.  The reference runs
 * transformers' ModernBertForMaskedLM (src/model/splade_modern.py:38,69-73) on right-padded batches; this entry
 * runs the same network - embeddings, n_layers x (LayerNorm, Wqkv + RoPE, global / +-local_window attention, Wo,
 * LayerNorm, GeGLU MLP), final_norm, head.dense + GELU + head.norm - on the REAL tokens only, packed back to
 * back, followed by the fused SPLADE head, in one call and with no host synchronisation.  Arithmetic = the
 * reference under torch.autocast(bf16): bf16 Linear outputs, fp32 residual stream / LayerNorm / softmax / RoPE.
 * Weights: the caller passes bf16 copies of the nn.Linear matrices and the fp32 LayerNorm weights / embeddings.
 * Needs hidden % 64 == 0 and >= 256, 64-wide heads, intermediate % 128 == 0, B <= 32767, S <= 65535.
 */
typedef struct splade_trunk_layer {
  const float* attn_norm; /* [H] fp32, NULL = Identity (layer 0, modeling_modernbert.py:318-319) */
  const void* wqkv;       /* [3H,H] bf16 attn.Wqkv.weight                                        */
  const void* wo;         /* [H,H]  bf16 attn.Wo.weight                                          */
  const float* mlp_norm;  /* [H] fp32                                                            */
  const void* wi;         /* [2I,H] bf16 mlp.Wi.weight (rows [0,I) input, [I,2I) gate)           */
  const void* wo2;        /* [H,I]  bf16 mlp.Wo.weight                                           */
  int32_t is_global;      /* 1: full attention + global rope theta, 0: sliding window + local    */
  int32_t reserved;
} splade_trunk_layer;

typedef struct splade_trunk_args {
  int32_t B, S, H, I, V, n_heads, n_layers;
  int32_t local_window;             /* keys with |pos_q - pos_k| <= local_window (config.local_attention / 2) */
  float norm_eps, global_theta, local_theta, reserved_f;
  const int64_t* input_ids;         /* [B,S] int64 (device) */
  const int32_t* mask;              /* [B,S] int32 (device), nonzero = real token */
  const float* tok_emb;             /* [V,H] fp32 embeddings.tok_embeddings.weight (tied to decoder.weight) */
  const float* emb_norm;            /* [H] */
  const splade_trunk_layer* layers; /* HOST array of n_layers entries (device pointers inside) */
  const float* final_norm;          /* [H] */
  const void* head_dense;           /* [H,H] bf16 head.dense.weight */
  const float* head_norm;           /* [H] */
  const void* dec_w;                /* [V,H] bf16 decoder.weight */
  const float* dec_bias;            /* [V] fp32 (pre-rounded to bf16 to mirror autocast) */
  float* rep;                       /* [B,V] out */
  int32_t* argmax;                  /* [B,V] out */
  void* hidden_out;                 /* optional [rows,H] bf16: the packed head input (rows from the workspace op) */
  uint32_t* meta_out;               /* optional [rows]: (straddles<<31 | doc<<16 | position) of every packed row */
  int32_t* plan_out;                /* optional int32[5]: {1, packed rows T, tiles, last-tile N, real tokens} */
  void* workspace;                  /* splade_workspace_bytes(SPLADE_OP_TRUNK_ENCODE, B, S, H, V, I) */
  size_t ws_bytes;
} splade_trunk_args;

int splade_trunk_encode(const splade_trunk_args* args, void* stream);

/* rows of the packed buffers for a (B,S) batch: every document rounded up to 8 tokens, whole 256-row tiles */
size_t splade_packed_rows(int B, int S);

/*
 * Building blocks of the trunk, exported for unit tests and reuse.
 * splade_trunk_gemm: out = epilogue(A [rows,K] bf16 x W[N,K]^T bf16), rows >= 128, K % 64 == 0, N % 128 == 0.
 *   epi 0 bf16 [rows,N]; 1 GELU -> bf16; 2 out (fp32 [rows,N]) += bf16(acc); 3 GeGLU: W is [2N,K], out bf16 [rows,N].
 * splade_trunk_attn: qkv_packed [rows,3H] bf16 -> out_packed [rows,H] bf16 in the packed layout of `mask`;
 *   window < 0 = global attention, otherwise keys with |pos_q - pos_k| <= window; rope_theta > 0 first rotates
 *   q and k IN PLACE in qkv_packed (the trunk's RoPE kernel, that base), <= 0 means they are already rotated.
 */
int splade_trunk_gemm(int epi, const void* A, int rows, const void* W, int N, int K, void* out, int ldo,
                      void* workspace, size_t ws_bytes, void* stream);
int splade_trunk_attn(void* qkv_packed, const int32_t* mask, int B, int S, int H, int n_heads, int window,
                      float rope_theta, void* out_packed, uint32_t* meta_out, void* workspace, size_t ws_bytes,
                      void* stream);

/*
 * Native trunk, TRAINING side (opt-in, NativeTrunkTrain in trunk_train.py).  splade_trunk_train_fwd runs
 * the layer loop of splade_trunk_encode up to final_norm (same kernels, same autocast arithmetic as transformers'
 * ModernBertModel.forward, modeling_modernbert.py:424-490) and writes everything the backward needs into the caller's
 * save buffers, which must arrive ZERO-FILLED (rows behind the last token then stay zero in every buffer, so the
 * backward can run over whole buffers).  rows = splade_packed_rows(B, S); the head fields of `args` are ignored.
 * The backward is orchestrated by the caller: dgrad GEMMs = splade_trunk_gemm on the transposed weights, wgrad GEMMs =
 * the caller's BLAS, attention backward = the caller's (library) attention on the saved q/k/v, and the two row-wise
 * kernels below:
 *   splade_trunk_ln_bwd    dx (+)= LayerNorm_backward(dy bf16 or fp32; x fp32, gamma) in fp32, dgamma [H] (deterministic);
 *                          dx_bf16 (optional) also receives the updated dx rounded to bf16 (the next dgrad's operand)
 *   splade_trunk_geglu_bwd h = [a | g] [rows,2I] bf16 (Wi output), dgg [rows,I] bf16 -> dh = [da | dg] [rows,2I] bf16
 *   splade_trunk_dqkv_pack the attention backward's (dq, dk, dv) -> the gradient rows of the Wqkv output (RoPE transposed)
 */
typedef struct splade_trunk_train_saves {
  float* x_attn_in; /* [n_layers][rows][H] fp32  residual stream entering each layer                      */
  float* x_mlp_in;  /* [n_layers][rows][H] fp32  residual stream after the attention block                */
  void* xb;         /* [n_layers][rows][H] bf16  attention-norm output (layer 0: bf16 embedding LayerNorm) */
  void* qkv;        /* [n_layers][rows][3H] bf16 Wqkv output with RoPE applied to q and k                 */
  void* ab;         /* [n_layers][rows][H] bf16  attention output                                         */
  void* xb2;        /* [n_layers][rows][H] bf16  mlp-norm output                                          */
  void* gg;         /* [n_layers][rows][I] bf16  GeGLU output                                             */
  float* x_final;   /* [rows][H] fp32            residual stream entering final_norm                      */
  void* out;        /* [rows][H] bf16            final_norm output: the trunk's output, packed            */
  uint32_t* meta;   /* [rows]                    (doc << 16 | position) of every packed row               */
  int32_t* plan;    /* int32[5]                  {1, packed rows T, tiles, last-tile N, real tokens}      */
} splade_trunk_train_saves;
int splade_trunk_train_fwd(const splade_trunk_args* args, const splade_trunk_train_saves* saves, void* stream);
int splade_trunk_ln_bwd(const void* dy, int dy_is_f32, const float* x, const float* gamma, float eps, int rows, int H,
                        float* dx, int accumulate, void* dx_bf16, float* dgamma, void* workspace, size_t ws_bytes,
                        void* stream);
int splade_trunk_geglu_bwd(const void* h, const void* dgg, int rows, int I, void* dh, void* stream);
/* padded[t,:] = row_of[t] < rows ? packed[row_of[t],:] : 0 for n_tokens = B*S positions (rows of row_bytes, a multiple of 16):
 * the packed -> padded gather in front of the library attention backward (ragged batches) */
int splade_trunk_gather_rows(const void* packed, const int64_t* row_of, int64_t n_tokens, int row_bytes, int64_t rows,
                             void* padded, void* stream);
/* gradients of the rotated q / k and of v ([B,heads,S,64] bf16, element strides given) -> dqkv [rows,3H] bf16: transpose of
 * the RoPE rotation (cos / sin [S,32] fp32) + interleave; row_of [B,S] int64 = packed row of a token (>= rows: padding),
 * NULL = row b*S+s */
int splade_trunk_dqkv_pack(const void* dq, const void* dk, const void* dv, int64_t stride_b, int64_t stride_h,
                           int64_t stride_s, const int64_t* row_of, const float* cos_t, const float* sin_t, int B, int S,
                           int n_heads, int64_t rows, void* dqkv, void* stream);

/*
 * Head prologue on the training / default path: transformers' ModernBertPredictionHead
 * `norm(act(dense(x)))` (modeling_modernbert.py:493-502), which the reference runs between the trunk and the decoder
 * (src/model/splade_modern.py:69-73), as the reference computes it under torch.autocast(bf16):
 *   y = bf16(x @ dense_w^T)   g = bf16(gelu_erf(y))   hidden = bf16(LayerNorm_fp32(g) * norm_w)
 * classifier_bias = norm_bias = false (the ModernBERT / A.X-Encoder configuration); H in {256, 512, 768, 1024}.
 *
 *   x       [rows_cap,H] bf16  trunk output cast to bf16 (what autocast hands F.linear); T valid rows
 *   dense_w [H,H] bf16         head.dense.weight;   dense_w_t = its transpose (row i = input feature i)
 *   norm_w  [H] fp32           head.norm.weight
 *   y       [rows_cap,H] bf16  out, kept for the backward (the dense output autograd would keep)
 *   hidden  [rows_cap,H] bf16  out, the input of splade_head_fwd
 *   stats   [T,2] fp32         out, per-row (mean, rstd) of the LayerNorm
 * rows_cap = rows allocated in every [rows_cap,H] buffer, >= max(T, 128) (the GEMM tiles 128 rows; rows >= T are
 * scratch).  Backward: d_hidden [T,H] fp32 (splade_head_bwd's output) -> dy [rows_cap,H] bf16 (gradient of y),
 * dx [rows_cap,H] bf16 (gradient of x; NULL to skip), d_norm_w [H] fp32 (deterministic two-stage sum).  The dense
 * weight gradient dy^T x is a plain GEMM left to the caller's BLAS.
 * Workspace: SPLADE_OP_HEAD_PROLOGUE.
 */
int splade_head_prologue_fwd(const void* x, const void* dense_w, const float* norm_w, float eps, int T, int rows_cap,
                             int H, void* y, void* hidden, float* stats, void* workspace, size_t ws_bytes,
                             void* stream);
int splade_head_prologue_bwd(const float* d_hidden, const void* y, const float* stats, const float* norm_w,
                             const void* dense_w_t, int T, int rows_cap, int H, void* dy, void* dx, float* d_norm_w,
                             void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* SPLADE_B200_H_ */
