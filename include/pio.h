/*
 * pio.h -- C ABI of libpio_sm100.so: the B200 (sm_100a) patch -> region -> caption hot path of
 * Patch-ioner, as a drop-in below the reference's Python facade.
 *
 * The reference has no FFI / plugin interface: its seam is Python.  Every entry
 * point below therefore cites the reference *call site* it replaces (file:line under
 * /root/reference/Patch-ioner) -- this is what a maintainer would bind with ctypes (see
 * INTEGRATION.md).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer into caller-owned memory unless the
 *     parameter name starts with h_ (host).  Row-major, innermost dimension contiguous.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing
 *     on the hot path (workspaces are caller-provided; sizes from the *_workspace_bytes calls) and
 *     returns 0 on success, a negative PIO_E* code otherwise; pio_last_error() gives the text.
 *   - a handle belongs to one (process, device); calls on it are stream-ordered, not re-entrant.
 *   - there is no CPU fallback and no backend dispatch: without a CUDA device every compute
 *     entry point fails with PIO_ECUDA.
 */
#ifndef PIO_H_
#define PIO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIO_OK 0
#define PIO_EINVAL (-1)
#define PIO_ECUDA (-2)
#define PIO_EUNSUPPORTED (-3)

/* arithmetic mode of the dense layers */
#define PIO_FP32 0 /* fp32 operands, fp32 FFMA accumulate: the 'fp32 parity' mode            */
#define PIO_BF16 1 /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM          */

/* element types of loose buffers */
#define PIO_DT_F32 0
#define PIO_DT_BF16 1

/* region weighting (bbox_utils.py:46-92) */
#define PIO_POOL_MEAN 0
#define PIO_POOL_GAUSS 1
#define PIO_POOL_ATTN 2

/* activations of the generic linear layer */
#define PIO_ACT_NONE 0
#define PIO_ACT_GELU_ERF 1 /* DINOv2 mlp (exact erf GELU)                 */
#define PIO_ACT_GELU_NEW 2 /* GPT-2 'gelu_new' (tanh form)                */
#define PIO_ACT_TANH 3     /* Talk2DINO projection MLP (talk2dino.py:73-83); fp32 mode only */
#define PIO_ACT_RELU 4     /* fp32 mode only                              */

const char* pio_last_error(void);
int pio_version(void);
/* number of kernels this library launched since the last pio_reset_launch_count() (bench evidence) */
long long pio_launch_count(void);
void pio_reset_launch_count(void);
/* frees the library's lazily allocated per-(device, stream) scratch (deterministic split-K partials and counters); */
/* call with those streams idle, e.g. at interpreter exit.  The scratch is re-created on demand.                    */
void pio_release_scratch(void);

/* ------------------------------------------------------------------------------------------ */
/* generic dense layer:  C[map(m), n] = res + gamma[n] * act( alpha * colscale[n] * (A W^T)[m,n] + bias[n] ) */
/* Replaces every nn.Linear / Conv1D on the path (DINOv2 qkv/proj/fc1/fc2, GPT-2 c_attn/c_proj/c_fc,   */
/* lm_head, clip_project decap.py:71, the two GEMMs of im2txtprojection.py:370,377).                   */
typedef struct {
  const void* A;   /* [M, lda] K-contiguous, dtype a_dt                                         */
  const void* W;   /* [N, ldw] K-contiguous (torch Linear layout), dtype a_dt                    */
  void* C;         /* [*, ldc], dtype c_dt                                                        */
  int M, N, K;
  int lda, ldw, ldc;
  int a_dt, c_dt;
  const float* bias;      /* [N] or NULL                                                          */
  const float* colscale;  /* [N] or NULL                                                          */
  const float* gamma;     /* [N] or NULL (LayerScale)                                             */
  const float* residual;  /* fp32 [*, ldres] or NULL; may alias C                                 */
  const float* res_rowscale; /* [M] or NULL: residual row m is multiplied by res_rowscale[m]      */
  int ldres;
  float alpha;
  int act;
  /* optional row remap: output row = (m / rows_per_group) * group_stride + group_offset + m % rows_per_group */
  int rows_per_group, group_stride, group_offset; /* rows_per_group == 0 -> identity             */
  /* fused row arg-max (PIO_BF16 mode only): when argmax_val != NULL nothing is written to C; instead every  */
  /* (row, column-slab) pair s writes its maximum, the FIRST column index reaching it and sum exp(v - max)    */
  /* to argmax_*[row * argmax_ld + s]; pio_argmax_slabs() tells how many slabs a call produces and          */
  /* pio_argmax_finish() reduces them.  Replaces logits -> softmax -> argmax of decap.py:133-141.           */
  float* argmax_val;
  int* argmax_idx;
  float* argmax_sumexp; /* may be NULL: no sum of exponentials (only the log-prob score needs it) */
  int argmax_ld;
  /* fused exponential (PIO_BF16 mode only, C must be bf16): when exp_ref != NULL the epilogue stores           */
  /*   C[m,n] = exp2( alpha*colscale[n]*acc - exp_ref[m] )   and writes, per (row, column-slab) pair s,           */
  /*   exp_psum[m*exp_ld+s] = sum of those values, exp_pmax[m*exp_ld+s] = max of alpha*colscale[n]*acc.           */
  /* This is the streaming softmax numerator of im2txtprojection.py:376 with a lagging reference maximum.        */
  const float* exp_ref;
  float* exp_psum;
  float* exp_pmax;
  int exp_ld;
  /* w_static != 0: W is not written by any kernel still in flight on the stream (weights, caption bank, wte); the  */
  /* tcgen05 kernels then fetch their first W tiles AHEAD of the programmatic-dependent-launch wait.  Leave it 0   */
  /* when W may be the output of the previous call on the same stream (then every load waits for that kernel).     */
  int w_static;
} PioLinear;
int pio_argmax_slabs(int M, int N);
/* ids[row*ids_ld + t] = arg-max over the slabs (first index wins ties); logprob_sum[row] += log softmax at it (or NULL) */
int pio_argmax_finish(const float* val, const int* idx, const float* sumexp, int ld, int slabs, int M, int* ids, int ids_ld,
                      int t, float* logprob_sum, void* stream);
int pio_linear(const PioLinear* p, int mode, void* stream);

/* LayerNorm over the last dim (DINOv2 eps 1e-6, GPT-2 eps 1e-5); x fp32 [rows, dim] with row stride ldx. */
int pio_layernorm(const float* x, int ldx, const float* w, const float* b, void* out, int out_dt, int ldo,
                  int rows, int dim, float eps, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* DINOv2 ViT-B/14-reg4 forward:  replaces  self.dino(imgs, is_training=True)  (src/model.py:783) and */
/* the forward hook on blocks[-1].attn.qkv (src/model.py:589-590, src/dino_extraction.py:8-9).         */
typedef struct {
  const float *ln1_w, *ln1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *ls1;
  const float *ln2_w, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b, *ls2;
} PioVitBlock;
typedef struct {
  const float* cls_token;       /* [768]                                                         */
  const float* register_tokens; /* [4,768]                                                       */
  const float* patch_w;         /* [768, 3*14*14] (Conv2d weight flattened c,ky,kx)              */
  const float* patch_b;         /* [768]                                                         */
  PioVitBlock blk[12];
  const float *norm_w, *norm_b;
} PioVitWeights;
typedef struct PioVit PioVit;
/* Copies / repacks the weights into library-owned device memory (bf16 copies in PIO_BF16 mode). */
int pio_vit_create(PioVit** out, const PioVitWeights* w, int mode, void* stream);
void pio_vit_destroy(PioVit* h);
size_t pio_vit_workspace_bytes(const PioVit* h, int B, int S);
/* imgs fp32 [B,3,S,S]; pos_embed fp32 [1+g*g,768] already resized to this grid (host, once);       */
/* out_tokens fp32 [B,N,768] = final LayerNorm of all tokens (cls | 4 reg | patches), N = 5+g*g;    */
/* out_attn   fp32 [B,g*g]   = softmax_j(<q_cls,k_j>/128) of the LAST block (process_self_attention,  */
/*                             dino_extraction.py:24-34) or NULL;                                   */
/* out_qkv    fp32 [B,N,2304] raw hooked qkv of the last block or NULL (debug / parity only).       */
int pio_vit_forward(PioVit* h, const float* imgs, int B, int S, const float* pos_embed, float* out_tokens,
                    float* out_attn, float* out_qkv, void* workspace, size_t workspace_bytes, void* stream);

/* Multi-head self-attention of one ViT block (head_dim 64): qkv [B,N,3*H*64] = [q|k|v] -> out [B,N,H*64], same dtype  */
/* (PIO_DT_F32: fp32 FFMA flash kernel; PIO_DT_BF16: tcgen05 kernel, needs a workspace for the transposed V copy).      */
/* Replaces dinov2 Attention.forward inside self.dino(...) (src/model.py:783).                                          */
size_t pio_attention_workspace_bytes(int dt, int B, int N, int H);
int pio_vit_attention(const void* qkv, void* out, int dt, int B, int N, int H, void* workspace, size_t workspace_bytes,
                      void* stream);

/* CLS attention map alone, from a hooked qkv tensor [B,N,3*D] (dino_extraction.py:24-34). */
int pio_cls_attention(const void* qkv, int qkv_dt, int B, int N, int D, int num_global, float* out_attn,
                      void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Region aggregation:  replaces extract_bboxes_feats (src/bbox_utils.py:8-109).                  */
/* boxes: [B,R,4] xywh in crop pixels, float32 (boxes_dt = PIO_DT_F32) or int32 (boxes_dt = 2);    */
/* tokens: fp32 patch tokens, image b at tokens + b*img_stride, patch p at + p*row_stride, D floats. */
/* out_bounds int32 [B,R,4] = (y_lo,y_hi,x_lo,x_hi) slice actually pooled (hi exclusive) or NULL.   */
/* attn_map fp32 [B,P] is only read for PIO_POOL_ATTN; it is NOT modified (the reference mutates     */
/* its CPU copy; the sequential rescaling is reproduced on a private copy in `workspace`).          */
/* set_mode = 0: out [B,R,D];  set_mode = 1 (get_single_embedding_per_image): out [B,D].            */
#define PIO_DT_I32 2
size_t pio_pool_workspace_bytes(int B, int R, int grid);
int pio_pool_boxes(const float* tokens, long long img_stride, long long row_stride, int B, int grid, int D,
                   const void* boxes, int boxes_dt, int R, int patch_size, int mode, float variance,
                   const float* attn_map, int set_mode, float* out, int* out_bounds, void* workspace,
                   size_t workspace_bytes, void* stream);
/* Weighted pooling with explicit weights [B,R,grid*grid]: out[b,r,:] = scale * sum_p w[b,r,p] x[b,p,:].  */
/* Replaces the trace branch src/model.py:1052-1054 (scale = 1/g^2), avg_self_attn_token :869 (1/P),   */
/* compute_region_means :45-94 and the masks= generalisation.                                        */
int pio_pool_grid(const float* tokens, long long img_stride, long long row_stride, int B, int grid, int D,
                  const float* weights, int R, float scale, float* out, void* stream);
/* map_traces_to_grid (src/bbox_utils.py:158-168): points double (x,y) pairs, trace t owns points      */
/* [offsets[t], offsets[t+1]); counts fp32 [T,grid,grid] (zeroed here); optional attn multiply        */
/* (model.py:1052-1053): counts *= attn[t].                                                          */
int pio_trace_bins(const double* points_xy, const int* offsets, int T, int grid, const float* attn, float* counts,
                   void* stream);
/* Per-"head" CLS attention maps (process_self_attention(..., ret_self_attn_maps=True), dino_extraction.py:24-34, */
/* softmaxed over the patches as at model.py:871): the hooked qkv re-cut into heads = 16 groups of D/16 channels,  */
/* out_maps fp32 [B, heads, P].  Feeds get_attn_heads_capt through pio_pool_grid (model.py:872, 950-960).           */
int pio_cls_head_attention(const void* qkv, int qkv_dt, int B, int N, int D, int num_global, int heads, float scale,
                           float* out_maps, void* stream);

/* ctx_cleaner (model.py:1425-1436): rows fp32 [B,P,D] (strided view allowed) cleaned against one context vector per image */
/* (ctx fp32 [B,D]) -> out fp32 [B,P,D] contiguous.  mode 0 = orthogonal_projection (alpha), 1 = contrastive_mask (eps).    */
/* prenorm != 0 L2-normalises rows and context first (clean_after_projection=False, model.py:905-913).                      */
int pio_ctx_clean(const float* rows, long long img_stride, long long row_stride, const float* ctx, long long ctx_stride, int B, int P,
                  int D, int mode, float alpha, float eps, int prenorm, float* out, void* stream);

/* Gaussian / uniform whole-image weights of compute_region_means (model.py:45-94) -> weights [grid*grid] */
int pio_region_mean_weights(int grid, float variance, float* weights, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* double-DINO (extract_bboxes_feats_double_dino, src/bbox_utils.py:300-403, called at model.py:983-992): the last      */
/* block of the backbone is run again on [cls | registers | the patches of one box] for every box.                      */
/* pio_vit_block_rows runs block `layer` (negative counts from the end) on PACKED sequences: x fp32 [T,768] in/out,      */
/* rows grouped in buckets of sequences of equal length (HOST arrays bucket_nseq / bucket_len; bucket rows contiguous).  */
size_t pio_vit_block_workspace_bytes(const PioVit* h, int T);
int pio_vit_block_rows(PioVit* h, int layer, float* x, int T, const int* bucket_nseq, const int* bucket_len, int nbuckets,
                       void* workspace, size_t workspace_bytes, void* stream);
/* out[t,:] = src[idx[t],:] (src rows src_ld floats apart); out[s,:] = mean of x rows [seg_start[s], +seg_len[s])        */
int pio_gather_rows(const float* src, long long src_ld, const int* idx, int T, int D, float* out, void* stream);
int pio_segment_mean(const float* x, const int* seg_start, const int* seg_len, int nseg, int D, float* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Image preprocessing on the device: replaces the PIL / torchvision transforms of src/model.py:347-357                 */
/*   Resize(resize_dim, BICUBIC) [+ CenterCrop(crop_dim)] + ToTensor + Normalize                                        */
/* for B images of one size.  imgs u8 [B,H,W,3]; kx/bx, ky/by: Pillow's fixed-point coefficient tables of the           */
/* horizontal / vertical pass (k [out, ksize] int32, bounds [out, 2] = first input index, tap count), built on the      */
/* host (patch-ioner_b200/preprocess.py); only the crop window [crop_top, +crop_h) x [crop_left, +crop_w) of the        */
/* resized image is computed; [row_first, +rows) are the input rows the vertical pass of that window needs.             */
/* mean3 / std3 are HOST pointers.  out fp32 [B,3,crop_h,crop_w].  Resized bytes are bit-identical to Pillow's.         */
size_t pio_preprocess_workspace_bytes(int B, int rows, int crop_w);
int pio_preprocess(const unsigned char* imgs, int B, int H, int W, const int* kx, const int* bx, int ksize_x, const int* ky,
                   const int* by, int ksize_y, int crop_left, int crop_top, int crop_w, int crop_h, int row_first, int rows,
                   const float* mean3, const float* std3, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* DeCap caption-memory projection: replaces Im2TxtProjector.project (im2txtprojection.py:353-385).  */
typedef struct PioBank PioBank;
/* bank fp32 [M,D] (zero rows already dropped, :345).  Builds library-owned copies: the bank, its     */
/* transpose, and 1/|row| (the reference re-normalises the whole bank on every call, :367).           */
int pio_bank_create(PioBank** out, const float* bank, long long M, int D, int mode, void* stream);
void pio_bank_destroy(PioBank* h);
long long pio_bank_rows(const PioBank* h);
size_t pio_project_workspace_bytes(const PioBank* h, int R);
/* q fp32 [R,D] (not modified; the reference normalises it in place, :368); out fp32 [R,D].          */
/* normalize != 0 -> out /= |out| (:379-380).  Partial form for a row-sharded bank:    */
/* if part_m/part_l are non-NULL the un-normalised (m[R], l[R], O[R,D]) of THIS shard are written     */
/* (O into out) and pio_project_finish() completes after the all-reduces.                             */
int pio_project(PioBank* h, const float* q, int R, float temperature, int normalize, float* out, float* part_m,
                float* part_l, void* workspace, size_t workspace_bytes, void* stream);
/* rescale a shard's partial by exp(m_local - m_global): O *= f, l *= f (before the SUM all-reduce) */
int pio_project_rescale(float* O, float* l, const float* m_local, const float* m_global, int R, int D, void* stream);
/* O / l and optional L2 normalisation (after the SUM all-reduce) */
int pio_project_finish(float* O, const float* l, int R, int D, int normalize, void* stream);
/* return_n_best_sims (im2txtprojection.py:382-383): the n largest cosine similarities <q^, bank_j^> per query,   */
/* descending -> out_sims fp32 [R,n], out_rows int32 [R,n] (bank row of each; may be NULL).  1 <= n <= 32.        */
/* The bank is walked in the same chunks as pio_project (same workspace size).  A row-sharded bank merges the      */
/* per-shard lists on the host side (n values per rank).                                                           */
int pio_best_sims(PioBank* h, const float* q, int R, int n, float* out_sims, int* out_rows, void* workspace,
                  size_t workspace_bytes, void* stream);
/* revert_transformation (embedding_utils.py:17-24) is pio_linear with W = A_pinv, bias = -A_pinv b. */

/* ------------------------------------------------------------------------------------------ */
/* DeCap prefix decoder: replaces decoding_batched (src/decap/decap.py:116-160) up to token ids.     */
typedef struct {
  const float *ln1_w, *ln1_b, *attn_w /*[768,2304] Conv1D in,out*/, *attn_b, *proj_w /*[768,768]*/, *proj_b;
  const float *ln2_w, *ln2_b, *fc_w /*[768,3072]*/, *fc_b, *fc2_w /*[3072,768]*/, *fc2_b;
} PioGptBlock;
typedef struct {
  const float* wte; /* [50257,768] (tied lm_head)                                                 */
  const float* wpe; /* [1024,768]                                                                 */
  PioGptBlock blk[4];
  const float *lnf_w, *lnf_b;
  const float* prefix_w; /* clip_project.model.0.weight [768, prefix_size]                        */
  const float* prefix_b; /* [768]                                                                 */
  int prefix_size;
} PioDecoderWeights;
typedef struct PioDecoder PioDecoder;
int pio_decoder_create(PioDecoder** out, const PioDecoderWeights* w, int mode, void* stream);
void pio_decoder_destroy(PioDecoder* h);
size_t pio_decode_workspace_bytes(const PioDecoder* h, int R, int steps);
/* Small batches (bf16 mode, no log-prob sum) are decoded by ONE persistent cooperative kernel (csrc/decode_fused_sm100.cu)   */
/* instead of ~32 launches per position.  max_rows: largest batch routed to it (default 32, at most 64; 0 turns it off);  */
/* ctas: CTAs of its grid (0 = one per SM; a cooperative grid of every SM cannot overlap with kernels of another stream -- */
/* a serving loop that keeps two batches in flight uses half the SMs per decode).  A negative argument keeps the setting.  */
/* Process-wide; the environment (PIO_DECODE_FUSED, PIO_DECODE_FUSED_MAX_ROWS, PIO_DECODE_FUSED_CTAS) overrides it.        */
int pio_set_decode_fused(int max_rows, int ctas);
/* test / debug aid: byte offsets of the decode workspace regions for R rows (pio_decode_greedy layout):                */
/* [0] x fp32 [R,768], [1] LayerNorm rows, [2] qkv rows, [3] gelu rows, [4] attention rows (fused decode), [5] key      */
/* cache, [6] value cache, [7] per-CTA arg-max partials (fused decode)                                                  */
int pio_decode_debug_layout(const PioDecoder* h, int R, long long* offsets, int n);
/* prefix fp32 [R,prefix_size]; out_ids int32 [R,steps]; out_logprob_sum fp32 [R] or NULL            */
/* (compute_scores: sum_t log softmax(logits_t)[tok_t], decap.py:157-160).  Fixed `steps` (30) greedy  */
/* steps with a KV cache, argmax first-index tie-break, no EOS early exit.                            */
int pio_decode_greedy(PioDecoder* h, const float* prefix, int R, int steps, int* out_ids, float* out_logprob_sum,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* ViECap captioner on the same region embeddings (src/viecap/entrypoint.py:98-162)                         */
/* GPT-2 with n_layer blocks of 12 heads x 64 (pretrained 'gpt2': 12 layers; viecap/ClipCap.py:157), no prefix            */
/* projection: the prompt arrives as input embeddings.  KV cache of up to 128 positions.                                 */
typedef struct {
  const float* wte; /* [50257,768] (tied lm_head) */
  const float* wpe; /* [1024,768]                 */
  const float *lnf_w, *lnf_b;
  const PioGptBlock* blk; /* n_layer entries       */
  int n_layer, n_head;
} PioGpt2Weights;
int pio_decoder_create_gpt2(PioDecoder** out, const PioGpt2Weights* w, int mode, void* stream);
size_t pio_decode_prompt_workspace_bytes(const PioDecoder* h, int R, int prompt_len, int steps);
/* greedy_search (viecap/search.py:108-191): prompt fp32 [R,prompt_len,768] input embeddings (position embeddings are     */
/* added here), then `steps` arg-max tokens (on the logits, first-index tie-break), out_ids int32 [R,steps].  No early    */
/* exit (the reference only exits early at batch 1, search.py:173-176); the sentence is cut at '.' on the host.           */
int pio_decode_greedy_prompt(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int* out_ids,
                             float* out_logprob_sum, void* workspace, size_t workspace_bytes, void* stream);
/* compute_scores of the ViECap path (entrypoint.py:164-177): mean negative log-likelihood of every right-padded token row  */
/* ids int32 [R,n] (lens int32 [R] tokens each) under the language model = GPT2LMHeadModel(input_ids, labels=input_ids).loss  */
/* of that sentence; the perplexity is exp() of it.  NaN for rows with fewer than two tokens.  n <= 128.                     */
/* The same search, stopped once EVERY row has emitted one of the two end-of-sentence ids: the reference runs all `steps`         */
/* positions for a batch and then keeps each row up to its first '.' (search.py:184-190), so the kept tokens are identical.     */
/* Columns from *out_steps_run (host int, may be NULL) on are filled with eos0.  Synchronises the stream every eighth step.      */
int pio_decode_greedy_prompt_eos(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int eos0, int eos1, int* out_ids,
                                 float* out_logprob_sum, int* out_steps_run, void* workspace, size_t workspace_bytes, void* stream);
/* beam_search (viecap/search.py:193-285; the reference's default, entrypoint.py:77,139-143 -- one call per region there, all   */
/* R regions x beam_width beams as one batch here): prompt fp32 [R,prompt_len,768]; at most `steps` new tokens (max_len);      */
/* eos0 / eos1 = the two end-of-sentence token ids (search.py:218); temperature divides the logits (:232; <= 0 means 1).       */
/* out_ids int32 [R,beam_width,steps], out_len int32 [R,beam_width] (tokens of each beam that count, search.py:280),            */
/* out_score fp32 [R,beam_width] (length-normalised log-probability), beams best first (:281-283).  *out_steps_run (host int,   */
/* may be NULL) = steps executed before every beam had ended (:275-276); the call synchronises the stream every fourth step.    */
size_t pio_decode_beam_workspace_bytes(const PioDecoder* h, int R, int prompt_len, int steps, int beam_width);
int pio_decode_beam_prompt(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int beam_width, int eos0, int eos1,
                           float temperature, int* out_ids, int* out_len, float* out_score, int* out_steps_run, void* workspace,
                           size_t workspace_bytes, void* stream);
size_t pio_gpt2_score_workspace_bytes(const PioDecoder* h, int R, int n);
int pio_gpt2_score_tokens(PioDecoder* h, const int* ids, const int* lens, int R, int n, float* out_nll_mean, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Mapping network (viecap/ClipCap.py:122-153): Linear weights in torch layout [out,in]                                  */
typedef struct {
  const float *norm1_w, *norm1_b, *q_w /*[768,768]*/, *kv_w /*[1536,768]*/, *proj_w /*[768,768]*/, *proj_b;
  const float *norm2_w, *norm2_b, *fc1_w /*[hidden,768]*/, *fc1_b, *fc2_w /*[768,hidden]*/, *fc2_b;
} PioMapperLayer;
typedef struct {
  int clip_size, project_len, prefix_len, n_layer, n_head, hidden;
  const float* linear_w;     /* [project_len*768, clip_size] */
  const float* linear_b;     /* [project_len*768]            */
  const float* prefix_const; /* [prefix_len,768]             */
  const PioMapperLayer* layers;
} PioMapperWeights;
typedef struct PioMapper PioMapper;
int pio_mapper_create(PioMapper** out, const PioMapperWeights* w, int mode, void* stream);
void pio_mapper_destroy(PioMapper* h);
size_t pio_mapper_workspace_bytes(const PioMapper* h, int R);
/* feats fp32 [R,clip_size], L2-normalised by the caller (entrypoint.py:108) -> out fp32 [R,prefix_len,768]              */
int pio_mapper_forward(PioMapper* h, const float* feats, int R, float* out, void* workspace, size_t workspace_bytes,
                       void* stream);
/* Entity retrieval (retrieval_categories.py:87-115): p = softmax(q . E^T / temperature) over n_ent unit-norm entity      */
/* embeddings [n_ent,D] (q [R,D] unit-norm), then the k largest probabilities in descending order (lowest index first     */
/* on ties) -> out_prob fp32 [R,k], out_idx int32 [R,k].  fp32 arithmetic in both modes.  k <= 32, n_ent <= 8192.         */
int pio_entity_topk(const float* q, const float* ent, int R, int n_ent, int D, float temperature, int k, float* out_prob,
                    int* out_idx, void* stream);

/* L2-normalise rows in place */
int pio_l2_normalize(float* x, int rows, int dim, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Batched detokenisation (HOST buffers, no kernel): replaces the per-row loop at the end of decoding_batched            */
/* (src/decap/decap.py:162-181) and SimpleTokenizer.decode (src/clip/simple_tokenizer.py:129-131).                        */
/* ids int32 [R, ld] (T used columns); table = concatenated byte strings of the vocabulary, token i = table[offsets[i] :  */
/* offsets[i+1]] (offsets has vocab + 1 entries).  Row r's bytes = concatenation of its tokens up to (excluding) the first */
/* eot_id -> out[row_offsets[r] : row_offsets[r+1]]; row_status[r] = 0 whole row, 2 cut at eot_id, 1 an id outside the     */
/* table (row left empty: the reference raises KeyError there and returns None for the whole call, decap.py:180-181).      */
/* strip_trailing_sep: drop one trailing byte of rows that were not cut (tables whose entries end in a separator);         */
/* replace_eow: substitute the bytes "</w>" by " " (simple_tokenizer.py:130).  UTF-8 decoding is left to the caller;       */
/* *all_ascii (may be NULL) = 1 when every output byte is < 128 (the caller may then decode the buffer in one piece).      */
int pio_detok_rows(const int* ids, int R, int T, int ld, const unsigned char* table, const long long* offsets, int vocab,
                   int eot_id, int strip_trailing_sep, int replace_eow, unsigned char* out, long long out_cap,
                   long long* row_offsets, int* row_status, int* all_ascii);

#ifdef __cplusplus
}
#endif
#endif /* PIO_H_ */
