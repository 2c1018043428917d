/* mvptr_b200.h -- C-ABI of the B200-native MVPTR hot path (libmvptr_b200.so).
 *
 * The reference (Junction4Nako/mvp_pytorch) has no FFI: its seam is the Python
 * class layer in oscar/modeling/modeling_vlbert.py.  Each entry point below is
 * the device-side replacement of one group of ATen calls made by those classes;
 * the comment above each function names the reference lines it replaces.
 * Python (mvp_pytorch_b200/_lib.py, ctypes) is the only caller.
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name ends in _host; `stream` is a cudaStream_t passed as void*;
 *   - return 0 on success, a negative MVPTR_ERR_* otherwise; never throw, never
 *     allocate device memory, never synchronise the device;
 *   - mvptr_last_error() returns a thread-local message for the last failure;
 *   - bf16 tensors are row-major with explicit leading dimensions in ELEMENTS.
 */
#ifndef MVPTR_B200_H
#define MVPTR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVPTR_OK 0
#define MVPTR_ERR_ARG (-1)      /* bad shape / alignment / unsupported option */
#define MVPTR_ERR_CUDA (-2)     /* CUDA runtime or driver error (see mvptr_last_error) */
#define MVPTR_ERR_UNSUPPORTED (-3)

#define MVPTR_ABI_VERSION 1

int mvptr_abi_version(void);
/* Number of kernels this library has launched in this process (every launch is counted at its
 * launch site): what bench.py reports as gpu_launches. */
unsigned long long mvptr_launch_count(void);
const char* mvptr_last_error(void);
/* Optional per-launch profiler: when enabled, every entry point brackets its kernel launches with
 * CUDA events on the launching stream; collect() synchronises and returns (name, work, ms). */
int mvptr_profile_enable(int on);
int mvptr_profile_collect(const char** names_host, double* work_host, float* ms_host, int cap);

/* ---- dense contraction (tcgen05 + TMEM + TMA) --------------------------------
 * D[M,N] (+)= epilogue( alpha * A[M,K] . B[N,K]^T )
 * Replaces every nn.Linear / matmul on the path: modeling_bert.py:293-295 (Q/K/V),
 * :349 (attention output dense), :395 (intermediate), :408 (output dense), :471
 * (pooler), :488 (head transform), :514/:531 (vocab / answer decoder);
 * modeling_vlbert.py:498 (region projection), :525-527 (txt/vis proj, sim_mat);
 * and their autograd dgrad / wgrad products.
 *
 * A and B are bf16.  a_mn / b_mn select the storage of the operand:
 *   0: "K-major"  -- A is [M rows][K contiguous] with pitch lda  (an activation, an nn.Linear weight)
 *   1: "MN-major" -- A is [K rows][M contiguous] with pitch lda  (the transposed view, used by wgrad / dgrad)
 * Pitches must be multiples of 8 elements and base pointers 16-byte aligned.
 *
 * Epilogue, applied per element in this order (null pointer = skipped):
 *   v = alpha*acc;  v += bias[n] (fp32 or bf16 per bias_is_bf16);
 *   pre_act[m,n] = bf16(v)                       (saved for backward; bf16(gelu'(v)) with aux_is_gelu_grad)
 *   v = act(v)            act: 0 none, 1 erf-GELU, 2 tanh
 *   v *= gelu'(gelu_grad_of[m,n])                (backward of act=1; v *= gelu_grad_of[m,n] with aux_is_gelu_grad)
 *   colsum[n] += sum_m bf16(v)                   (fp32 atomics; only with gelu_grad_of: the bias gradient of
 *                                                 the GELU layer, i.e. the column sums of exactly what is stored)
 *   v = keep(seed, m*N+n) ? v/keep_prob : 0      (dropout, p_drop > 0)
 *   v += residual[m,n]
 *   D[m,n] = v   or   D[m,n] += v  (accumulate=1: TMA reduce-add; required when split_k > 1)
 * D is bf16 (d_is_f32=0) or fp32 (d_is_f32=1).
 */
typedef struct {
  const void* A;
  const void* B;
  void* D;
  int M, N, K;
  int lda, ldb, ldd;
  int a_mn, b_mn;
  int d_is_f32;
  int accumulate;
  int split_k; /* <=1: no split */
  float alpha;
  const void* bias;
  int bias_is_bf16;
  void* pre_act; /* bf16 [M,N], pitch ld_aux */
  int act;
  const void* gelu_grad_of; /* bf16 [M,N], pitch ld_aux */
  const void* residual;     /* bf16 [M,N], pitch ld_aux */
  int ld_aux;
  float p_drop;
  uint32_t seed;
  int block_n;  /* 0 = auto, else 128 or 256 */
  int cta_pair; /* 0 = auto, 1 = single-CTA tiles (128 x block_n), 2 = CTA-pair tiles (256 x 256, cta_group::2) */
  float* colsum; /* fp32 [N] (+=), nullable; requires gelu_grad_of and the lean GELU' epilogue (see mvptr_gemm) */
  int aux_is_gelu_grad; /* 1: the auxiliary tensor carries gelu'(pre-activation) instead of the pre-activation:
                           pre_act[m,n] = bf16(gelu'(v)) in a forward call (act = 1), and a backward call multiplies
                           by gelu_grad_of[m,n] itself.  Moves the erf of GELU' out of the dgrad epilogue (it shares
                           the forward's erf evaluation); the [M,N] auxiliary tensor keeps its size. */
} mvptr_gemm_args;

int mvptr_gemm(const mvptr_gemm_args* args, void* stream);
/* Persistent CTAs a GEMM launch may occupy (0 = all 148 SMs, the default).  Data-parallel runs leave a few SMs
 * to the NCCL kernels that overlap with backward (mvp_pytorch_b200/parallel.py). */
int mvptr_gemm_set_max_ctas(int n);


/* ---- embeddings + LayerNorm ---------------------------------------------------
 * y = dropout(LN(word[ids] + pos[pos_ids or arange(L)] + type[type_ids]))
 * Replaces BertEmbeddings.forward, modeling_bert.py:262-277 (called twice per forward,
 * modeling_vlbert.py:479-482).  ids/type_ids/pos_ids int64 [B,L]; tables bf16.
 * y row (b,t) is written at y + b*y_batch_stride + t*H when y_rows_per_batch > 0
 * (lets the tag embeddings land directly inside the [B, Lt+R, H] visual sequence,
 * replacing the torch.cat of modeling_vlbert.py:506); otherwise densely.
 * pre/mean/rstd (nullable) are saved for backward. */
int mvptr_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const void* word,
                       const void* pos, const void* type, const void* gamma, const void* beta, void* y,
                       int y_rows_per_batch, long long y_batch_stride, void* pre, float* mean, float* rstd, int B,
                       int L, int H, float eps, int vocab, int max_pos, int n_types, float p_drop, uint32_t seed,
                       void* stream);
/* Backward of the gather: word rows by fp32 atomics (row padding_idx gets none, as
 * nn.Embedding(padding_idx=0), modeling_bert.py:253), position/type rows by per-position reduction. */
int mvptr_embed_bwd(const void* dpre, const int64_t* ids, const int64_t* type_ids, float* dword, float* dpos,
                    float* dtype, int B, int L, int H, int vocab, int n_types, int padding_idx, void* stream);

/* y = dropout(LN(x)), TF-style eps inside the sqrt.  Replaces BertLayerNorm.forward,
 * modeling_bert.py:242-246 (9 ATen kernels) at :351, :410, :490 and modeling_vlbert.py:499-503. */
int mvptr_ln_fwd(const void* x, const void* gamma, const void* beta, void* y, int y_rows_per_batch,
                 long long y_batch_stride, float* mean, float* rstd, int rows, int H, float eps, float p_drop,
                 uint32_t seed, void* stream);
/* out = LN(dropout(x) + residual): the tail of BertSelfOutput / BertOutput (modeling_bert.py:350-351,
 * 409-410) as one coalesced pass; pre_out (nullable) saves the pre-LN sum for backward. */
int mvptr_add_ln_fwd(const void* x, const void* residual, float in_p_drop, uint32_t in_seed, void* pre_out,
                     const void* gamma, const void* beta, void* y, float* mean, float* rstd, int rows, int H,
                     float eps, void* stream);
/* erf-GELU of BertIntermediate (modeling_bert.py:396) and its backward fused with the dense bias gradient */
int mvptr_gelu_fwd(const void* x, void* y, size_t n, void* stream);
int mvptr_gelu_bwd_colsum(const void* dy, const void* pre, void* dx, float* dbias, int M, int N, void* stream);
/* LayerNorm backward; also emits dx with the dense-output dropout mask re-applied (dx_drop)
 * and accumulates dgamma / dbeta / dbias (fp32, atomics). */
int mvptr_ln_bwd(const void* dy, int dy_rows_per_batch, long long dy_batch_stride, const void* x, const float* mean,
                 const float* rstd, const void* gamma, void* dx, void* dx_drop, float* dgamma, float* dbeta,
                 float* dbias, int rows, int H, float out_p_drop, uint32_t out_seed, float in_p_drop,
                 uint32_t in_seed, void* stream);
/* out[n] += sum_m x[m,n]  (bias gradient of a dense layer) */
int mvptr_colsum(const void* x, int ldx, float* out, int M, int N, void* stream);

/* region features [rows,K] bf16|fp32|fp16 (src_kind 0|1|2; any pitch) -> bf16 [rows, ld_dst] zero padded so that the
 * K=2054 projection (modeling_vlbert.py:498) has a 16-byte aligned pitch for TMA.  fp16 is what the reference's
 * --half_evaluation (run_retrieval.py:1047-1048 + prepare_inputs) and DeepSpeed fp16 loaders (run_pretrain_ml.py:504)
 * deliver. */
int mvptr_pad_cast(const void* src, int src_kind, long long ld_src, void* dst, int ld_dst, int rows, int K,
                   void* stream);
/* additive mask (1-mask)*-10000 of modeling_vlbert.py:430-460 for the sequence
 * [a(row_a[r], 0..La) | b(row_b[r], b_col0..Lb)] -- also assembles the joint and
 * hard-negative masks of :542-566, :587 (row_a/row_b nullable = identity). */
int mvptr_mask_prepare(const int64_t* mask_a, int La, const int64_t* mask_b, int Lb, int b_col0, const int64_t* row_a,
                       const int64_t* row_b, float* out, int rows, void* stream);
/* out[r] = cat(a[row_a[r]], b[row_b[r], b_col0:])  -- torch.cat + index_select of
 * modeling_vlbert.py:542-566 and :586 as one gather; and its backward (fp32 scatter-add). */
int mvptr_concat_rows(const void* a, int La, const void* b, int Lb, int b_col0, const int64_t* row_a,
                      const int64_t* row_b, void* out, int rows, int H, void* stream);
int mvptr_concat_rows_bwd(const void* dout, int La, int Lb, int b_col0, const int64_t* row_a, const int64_t* row_b,
                          float* da, float* db, int rows, int H, void* stream);
/* masked_select of MLM rows (modeling_vlbert.py:1232, 1246) and its backward */
int mvptr_gather_rows(const void* src, const int64_t* idx, void* out, int n, int H, void* stream);
int mvptr_scatter_rows_add(const void* src, const int64_t* idx, float* dst, int n, int H, void* stream);
int mvptr_cast_f32_bf16(const float* src, void* dst, size_t n, void* stream);
/* widen a bf16 buffer into fp32 (data-parallel gradient path: bf16 all-reduce over NVLink, fp32 arena);
 * max_ctas > 0 caps the grid so the cast shares the SMs politely with the backward GEMMs it overlaps */
int mvptr_cast_bf16_f32(const void* src, float* dst, size_t n, int max_ctas, void* stream);
int mvptr_add_cast(const float* a, const void* b, void* d, size_t n, void* stream);

/* Which attention kernels run for L <= 128: the tcgen05 / TMA / TMEM kernels of csrc/attention_tc.cu or the mma.sync
 * kernels of csrc/attention.cu (the only ones for 128 < L <= 256).  Per direction: -1 follow MVPTR_ATTN_FWD_TC /
 * MVPTR_ATTN_BWD_TC, 0 auto (the faster one per shape, as measured), 1 tcgen05, 2 mma.sync.  Both families implement the
 * same entry points below bit-compatibly in their dropout masks, so forward and backward may use different families. */
int mvptr_attn_set_path(int fwd_mode, int bwd_mode);
/* ---- fused masked attention, head_dim 64, L <= 256 ---------------------------------
 * ctx = dropout(softmax(q k^T / 8 + maskadd)) v per head, reading the fused QKV projection
 * [B*L, 3H] and writing head-merged context [B*L, H].  Replaces CaptionBertSelfAttention.forward
 * modeling_vlbert.py:79-100 (2 bmm + div + add + softmax + dropout + permute) and
 * transpose_for_scores modeling_bert.py:299-303.  lse [B,nh,L] (nullable) is saved for backward. */
int mvptr_attn_fwd(const void* qkv, int ld_qkv, const float* maskadd, void* ctx, int ld_ctx, float* lse, int B, int L,
                   int nh, int H, float p_drop, uint32_t seed, void* stream);
/* dbias (nullable, fp32 [3H], +=): column sums of dqkv, i.e. the gradient of the fused QKV bias */
int mvptr_attn_bwd(const void* qkv, int ld_qkv, const float* maskadd, const void* ctx, const void* dctx, int ld_ctx,
                   const float* lse, void* dqkv, float* dbias, int B, int L, int nh, int H, float p_drop,
                   uint32_t seed, void* stream);

/* ---- losses -------------------------------------------------------------------------
 * CrossEntropyLoss(ignore_index=-1) over fp32 logits [n,V]: modeling_vlbert.py:1229,1235,1249.
 * fwd accumulates sum of row losses and the valid-row count; bwd writes bf16 dlogits. */
int mvptr_ce_fwd(const float* logits, int ld, const int64_t* labels, int n, int V, int ignore_index, float* row_lse,
                 float* loss_sum, float* n_valid, void* stream);
int mvptr_ce_bwd(const float* logits, int ld, const int64_t* labels, int n, int V, int ignore_index,
                 const float* row_lse, const float* n_valid, const float* gscale, void* dlogits, int ld_d,
                 void* stream);
/* F.normalize(p=2) of modeling_vlbert.py:525-526 */
int mvptr_l2norm_fwd(const float* x, float* y, void* y16, float* norm, int n, int H, void* stream);
int mvptr_l2norm_bwd(const float* dy, const float* y, const float* norm, void* dx16, int n, int H, void* stream);
/* VSC loss (modeling_vlbert.py:1238-1241) fused with the in-batch hardest-negative argmax (:530-534) */
int mvptr_vsc_fwd(const float* sim, int B, const float* logit_scale, float* row_lse, float* col_lse, float* loss,
                  int64_t* hard_img, int64_t* hard_txt, void* stream);
int mvptr_vsc_bwd(const float* sim, int B, const float* logit_scale, const float* row_lse, const float* col_lse,
                  const float* gscale, float* dsim, float* dlogit_scale, void* stream);
/* ITM / retrieval classifier [n,H]x[C,H]^T, C small (modeling_vlbert.py:1247, 1680, 1708) + CE (:1251, :1682) */
int mvptr_small_head_fwd(const void* x, int ldx, const void* W, const void* bias, float* logits, int n, int H, int C,
                         void* stream);
int mvptr_small_head_bwd(const float* dlogits, const void* x, int ldx, const void* W, void* dx, int ld_dx, float* dW,
                         float* db, int n, int H, int C, void* stream);
/* acc = float[2] {sum of row losses, valid rows}: accumulated by the forward call (dlogits == NULL, acc pre-zeroed),
 * read by the backward call (dlogits != NULL).  Labels outside [0, C) (e.g. ignore_index -1 of the QA loss,
 * modeling_vlbert.py:1262-1264) are ignored: no loss, zero gradient, mean over the valid rows. */
int mvptr_small_ce(const float* logits, const int64_t* labels, int n, int C, float* acc, float* dlogits,
                   const float* gscale, void* stream);

/* ---- optimizer ------------------------------------------------------------------------
 * AdamW.step of transformers/pytorch_transformers/optimization.py:130-189 over one flat fp32
 * arena (decay-first layout), also refreshing the bf16 compute copy; grad clipping folded in. */
/* dyn_lr_step (nullable, device float[2] = {lr, step}): read at execution time instead of the by-value
 * lr / step, so a captured CUDA graph follows the LR schedule and the bias correction. */
int mvptr_adamw(float* p, const float* g, float* m, float* v, void* p16, size_t n, size_t decay_end, float lr,
                float beta1, float beta2, float eps, float weight_decay, int step, int correct_bias,
                const float* grad_sumsq, float max_norm, const float* dyn_lr_step, void* stream);
/* Copies *src (device or pinned-host word, read when the copy executes) into the dropout epoch mixed
 * into every dropout hash: lets CUDA-graph replays draw fresh masks with baked-in seeds. */
int mvptr_set_dropout_epoch(const uint32_t* src, void* stream);
/* Per-replay parameters of a CUDA-graph training step (run_pretrain_ml.py:528-545: scheduler.step() and the
 * optimizer's step count advance once per iteration).  ring: `slots` 16-byte records {float lr, float step,
 * uint32 dropout epoch, pad} in PINNED host (or device) memory, filled by the host for replay n at slot
 * n % slots before it launches the replay; counter: device word = replays executed so far.  One 1-thread
 * kernel reads record counter % slots at EXECUTION time, publishes {lr, step} to dyn_lr_step (nullable, the
 * mvptr_adamw argument) and the epoch to every dropout site, then increments the counter -- so a host that
 * runs several replays ahead still gives every replay its own values. */
int mvptr_step_params(const void* ring, int slots, uint32_t* counter, float* dyn_lr_step, void* stream);
int mvptr_sumsq(const float* g, size_t n, float* out, void* stream);

/* ---- weakly-supervised phrase grounding (WRA), batched ----------------------------------
 * Replaces the per-sample Python loops of modeling_vlbert.py:1288-1300 and helpers
 * mask_slice_and_stack :1502-1508, t2i_sim :1543-1550, get_pos_neg_sims :1553-1596.
 * seq bf16 [B,Ltot,H]; phrase_index/img_index int64 [B,2]; neg_img int64 [B] (the
 * random.choice of :1573); rand_pos/rand_neg int64 [B,P] in [0,3) (the torch.randint of
 * :1548).  Outputs pos/neg mean cosine similarity per sample and the selected region
 * token of every phrase (sel_*, int32 [B, mvptr_wra_max_phrases()], -1 = none). */
int mvptr_wra_max_phrases(void);
int mvptr_wra_fwd(const void* seq, int B, int Ltot, int H, const int64_t* phrase_index, const int64_t* img_index,
                  const int64_t* neg_img, const int64_t* rand_pos, const int64_t* rand_neg, int P, float* pos_out,
                  float* neg_out, int* sel_pos, int* sel_neg, void* stream);
int mvptr_wra_bwd(const void* seq, int B, int Ltot, int H, const int64_t* phrase_index, const int64_t* neg_img,
                  const int* sel_pos, const int* sel_neg, const float* dpos, const float* dneg, float* dseq,
                  void* stream);

/* ---- referring-expression head ------------------------------------------------------------
 * BiImageBertForRE.forward mod 1 / mod 2 (modeling_vlbert.py:1936-1956): logits[b, j] = score of region token
 * first + j against the [CLS] token of sequence b -- cosine similarity (normalize=1: F.normalize + bmm) or raw
 * dot product (normalize=0).  p_drop > 0 applies self.dropout(sequence_output) (:1932) on the fly.
 * inv_norm [B*R, 2] fp32 is saved for backward; dseq must be zeroed by the caller. */
int mvptr_cls_region_score_fwd(const void* seq, int B, int Ltot, int H, int first, int R, int normalize, float* logits,
                               float* inv_norm, float p_drop, uint32_t seed, void* stream);
int mvptr_cls_region_score_bwd(const void* seq, int B, int Ltot, int H, int first, int R, int normalize,
                               const float* logits, const float* inv_norm, const float* dlogits, void* dseq,
                               float p_drop, uint32_t seed, void* stream);
/* dx = dy * gelu'(pre): backward of the head-transform activation, modeling_bert.py:489 */
int mvptr_gelu_bwd(const void* dy, const void* pre, void* dx, size_t n, void* stream);
/* instance_bce_with_logits, modeling_vlbert.py:878-883 (VQA loss): loss += sum BCE / n */
int mvptr_bce_fwd(const float* logits, int ld, const float* labels, int n, int C, float* loss, void* stream);
int mvptr_bce_bwd(const float* logits, int ld, const float* labels, int n, int C, const float* gscale, void* dlogits,
                  int ld_d, void* stream);

/* ---- one encoder layer per call --------------------------------------------------------
 * CaptionBertLayer.forward (modeling_vlbert.py:191-199): attention (:63-103) -> BertSelfOutput
 * (modeling_bert.py:348-352) -> BertIntermediate (:394-397) -> BertOutput (:407-411), and its
 * backward.  The caller owns every buffer; [M = B*L rows]:
 *   x,att,pre1,a1,pre2,out [M,H] bf16; qkv [M,3H]; pre_g,inter [M,I]; lse [B,nh,L] f32;
 *   st1,st2 [2,M] f32 (mean | rstd).  save=0 (inference) skips lse/st/pre1/pre2.
 * Backward reads the saved activations + dout and writes dx; d* are scratch of the same
 * shapes as their forward counterparts; g_* are fp32 gradient accumulators (+=). */
typedef struct {
  int B, L, H, I, nh, save;
  float eps, p_hidden, p_attn;
  uint32_t seed_attn, seed1, seed2;
  const float* maskadd;
  const void *w_qkv, *b_qkv, *w_o, *b_o, *ln1_g, *ln1_b, *w_i, *b_i, *w_o2, *b_o2, *ln2_g, *ln2_b;
  const void* x;
  void *qkv, *att;
  float* lse;
  void* pre1;
  float* st1;
  void *a1, *pre_g, *inter, *pre2;
  float* st2;
  void* out;
  void* tmp; /* [M,H] bf16 forward scratch (dense outputs before dropout+residual+LN) */
  const void* dout;
  void *dx, *dpre2, *dpre2d, *dpre_g, *da1, *dpre1, *dpre1d, *datt, *dqkv;
  float *g_w_qkv, *g_b_qkv, *g_w_o, *g_b_o, *g_ln1_g, *g_ln1_b, *g_w_i, *g_b_i, *g_w_o2, *g_b_o2, *g_ln2_g, *g_ln2_b;
} mvptr_layer_args;
int mvptr_layer_fwd(const mvptr_layer_args* args, void* stream);
int mvptr_layer_bwd(const mvptr_layer_args* args, void* stream);

/* ---- retrieval scoring (run_retrieval.py) ------------------------------------------------
 * Per-row top-k of fp32 scores in the reference's ranking order: np.argsort(sim)[::-1][:k] of
 * compute_ranks_coarse (run_retrieval.py:481-522) / compute_ranks (:429-478), made deterministic
 * as descending score then descending index.  idx_out int64 [rows,k]; val_out (nullable) [rows,k]. */
int mvptr_topk_rows(const float* x, long long ld, int rows, int n, int k, int64_t* idx_out, float* val_out,
                    void* stream);
/* softmax(logits)[:,1] of the 2-way ITM classifier (run_retrieval.py:776-777, 818-820) */
int mvptr_match_prob(const float* logits, float* prob, int n, void* stream);

/* ---- input pipeline (SURVEY f-2; csrc/data_prep.cu) -------------------------------------------------
 * Region features as the reference's TSV files hold them: base64 text of float32[num_boxes, K] per sample
 * (oscar_datasets_ml/oscar_tsv4.py:696-727: np.frombuffer(base64.b64decode(..), float32).reshape(num_boxes, K) ->
 * torch.tensor(dtype)), zero-padded to R rows.  src = the samples' texts back to back, offsets[B+1] their byte
 * offsets; dst [B, R, ld_dst] bf16 or fp32 (rows >= num_boxes and columns >= K zeroed).  max_boxes >= max(num_boxes)
 * sizes the grid.  error_flag (device int, caller-zeroed): bit 0 length mismatch, bit 1 invalid character, bit 2
 * misplaced padding. */
int mvptr_b64_decode_features(const void* src, const int64_t* offsets, const int32_t* num_boxes, void* dst,
                              int dst_is_f32, int B, int R, int K, int ld_dst, int max_boxes, int* error_flag,
                              void* stream);
/* BERT token masking (random_word, oscar_tsv4.py:782-820) and phrase masking (random_phrases, :822-850, labels
 * dropped by :960) in place on ids [B, L]; labels [B, L] receives the original id at selected token positions, -1
 * elsewhere.  u [B, L] uniforms / r [B, L] non-negative integer draws replay recorded random numbers (nullable:
 * hash of seed).  links [B, L, max_links] int32 (nullable): phrase indexes tied to caption token i, -1 padded. */
int mvptr_mlm_mask(int64_t* ids, int64_t* labels, const int32_t* tok_first, const int32_t* tok_count,
                   const int32_t* phr_first, const int32_t* phr_count, const int32_t* links, int max_links,
                   const float* u, const int64_t* r, int B, int L, long long mask_id, long long word_vocab,
                   long long phrase_vocab, long long vocab_size, uint32_t seed, void* stream);

/* ---- fp32 verification tier (csrc/fp32_tier.cu) ---------------------------------------------
 * north_star: "bit-exact for token/region indexing, masking and top-k ranking order under fp32", "1e-4 in fp32".
 * The reference computes in fp32 by default (oscar/tmp_config_FP32.json; run_retrieval.py:1047 halves only on a flag).
 * fp32 storage everywhere; contractions run on the tcgen05 GEMM as six bf16 products of the 3-way split
 * x = hi + mid + lo produced by mvptr_f32_split3 (mvptr_gemm with fp32 D and accumulate = 1).  Plain kernels,
 * not a performance path.  Same reference lines as the bf16 entry points of the same name. */
int mvptr_f32_split3(const float* src, long long ld_src, int rows, int K, void* hi, void* mid, void* lo, int ld_dst,
                     void* stream);
int mvptr_f32_ln_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* y,
                     int y_rows_per_batch, long long y_batch_stride, float* pre_out, float* mean_out, float* rstd_out,
                     int rows, int H, float eps, void* stream);
int mvptr_f32_ln_bwd(const float* dy, int dy_rows_per_batch, long long dy_batch_stride, const float* pre,
                     const float* mean, const float* rstd, const float* gamma, float* dx, float* dgamma, float* dbeta,
                     int rows, int H, void* stream);
int mvptr_f32_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const float* word,
                           const float* pos, const float* type, const float* gamma, const float* beta, float* y,
                           int y_rows_per_batch, long long y_batch_stride, float* pre_out, float* mean_out,
                           float* rstd_out, int B, int L, int H, float eps, int vocab, int max_pos, int n_types,
                           void* stream);
int mvptr_f32_embed_bwd(const float* dpre, const int64_t* ids, const int64_t* type_ids, float* dword, float* dpos,
                        float* dtype, int B, int L, int H, void* stream);
/* act: 1 erf-GELU (modeling_bert.py:142-148), 2 tanh, 3 relu; in place.  _bwd: saved = pre-activation (GELU) or output */
int mvptr_f32_act(float* x, size_t n, int act, void* stream);
int mvptr_f32_act_bwd(const float* dy, const float* saved, float* dx, size_t n, int act, void* stream);
/* probs (nullable): softmax probabilities [B, nh, L, L] saved for the backward */
int mvptr_f32_attn_fwd(const float* qkv, int ld_qkv, const float* maskadd, float* ctx, int ld_ctx, float* probs, int B,
                       int L, int nh, int H, void* stream);
int mvptr_f32_attn_bwd(const float* qkv, int ld_qkv, const float* probs, const float* dctx, int ld_ctx, float* dqkv,
                       int B, int L, int nh, int H, void* stream);
int mvptr_f32_colsum(const float* x, int ldx, float* out, int M, int N, void* stream);
int mvptr_f32_small_head_fwd(const float* x, long long ldx, const float* W, const float* bias, float* logits, int n,
                             int H, int C, void* stream);
int mvptr_f32_small_head_bwd(const float* dlogits, const float* x, long long ldx, const float* W, float* dx, float* dW,
                             float* db, int n, int H, int C, void* stream);
int mvptr_f32_concat_rows_bwd(const float* dout, int La, int Lb, int b_col0, const int64_t* row_a, const int64_t* row_b,
                              float* da, float* db, int rows, int H, void* stream);
int mvptr_f32_scatter_rows_add(const float* src, const int64_t* idx, float* dst, int n, int H, void* stream);
int mvptr_f32_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx, int n, int H, void* stream);
int mvptr_f32_ce_bwd(const float* logits, int ld, const int64_t* labels, int n, int V, int ignore_index,
                     const float* row_lse, const float* n_valid, const float* gscale, float* dlogits, int ld_d,
                     void* stream);
int mvptr_f32_bce_bwd(const float* logits, int ld, const float* labels, int n, int C, const float* gscale,
                      float* dlogits, int ld_d, void* stream);
int mvptr_f32_wra_fwd(const float* seq, int B, int Ltot, int H, const int64_t* phrase_index, const int64_t* img_index,
                      const int64_t* neg_img, const int64_t* rand_pos, const int64_t* rand_neg, int P, float* pos_out,
                      float* neg_out, int* sel_pos, int* sel_neg, void* stream);
int mvptr_f32_wra_bwd(const float* seq, int B, int Ltot, int H, const int64_t* phrase_index, const int64_t* neg_img,
                      const int* sel_pos, const int* sel_neg, const float* dpos, const float* dneg, float* dseq,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVPTR_B200_H */
