/* libdvst.so - C ABI of the B200-native scoring path (TimeSformer forward + DINO head / loss).
 *
 * The reference (Luke-537/Dino-Video-Summarization-Transformer) is pure Python: its boundary for this path is the
 * nn.Module level.  This header is the FFI the Python host side (dvst_b200/*.py, ctypes) binds;
 * each entry point cites the reference interface it stands in for.
 *
 * Conventions: every function returns 0 on success, DVST_ERR otherwise (message via dvst_last_error(), thread
 * local); nothing throws.  All data pointers are caller-owned DEVICE pointers (tensor.data_ptr()) unless stated
 * otherwise; `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).  No call
 * synchronises the device or allocates device memory.  A model handle is not re-entrant (one per process / GPU).
 */
#ifndef DVST_H_
#define DVST_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DVST_ABI_VERSION 1
#define DVST_ERR 1

/* attention_type: cfg.TIMESFORMER.ATTENTION_TYPE (models/timesformer.py:97,120-124) */
#define DVST_ATTN_DIVIDED 0
#define DVST_ATTN_JOINT 1
#define DVST_ATTN_SPACE_ONLY 2
/* precision: 0 = bf16 operands / fp32 accumulate; 1 = 2-term bf16 split (hi*hi + hi*lo + lo*hi), ~16 mantissa bits */
#define DVST_PRECISION_BF16 0
#define DVST_PRECISION_SPLIT 1

typedef struct dvst_config {
  int depth;          /* number of blocks (12 for ViT-B, models/timesformer.py:595) */
  int num_frames;     /* cfg.DATA.NUM_FRAMES = rows of time_embed (:234) */
  int pos_grid;       /* sqrt(#patch positions of pos_embed) = cfg.DATA.TRAIN_CROP_SIZE / 16 (:231) */
  int attention_type; /* DVST_ATTN_* */
  int precision;      /* DVST_PRECISION_* */
} dvst_config;

typedef struct dvst_model dvst_model; /* opaque */

const char* dvst_last_error(void);
/* Number of kernel launches issued by the library so far (bench.py's `gpu_launches`). */
long long dvst_launch_count(void);
int dvst_abi_version(void);

/* Tuning / test hook: 2 (default) lets large GEMMs use tcgen05 cta_group::2 CTA pairs (256x256 tiles); 1 forces the
 * single-CTA 128x256 kernel. */
void dvst_set_gemm_cta_group(int cg);

/* Tuning / test hook: residual boxes per epilogue warp of the K <= 1024 deferred-LayerNorm producer GEMMs: 5 (default,
 * residual loads four chunks ahead, four operand stages) or 3 (two ahead, five operand stages).  Same results. */
void dvst_set_producer_ring(int boxes);

/* Tuning / test hook: 1 (default) runs spatial / space-only attention (modes 1 and 3, up to 256 tokens, bf16) on the
 * tcgen05 + TMEM kernel (one-pass softmax where scratch is available); 2 the same kernel with the two-pass softmax;
 * 0 forces the mma.sync kernels for every mode. */
void dvst_set_attention_impl(int impl);

/* Tuning / test hook: 1 (default) lets models whose folded tensors are registered run the deferred
 * LayerNorm path (see dvst_gemm_bf16_ln); 0 forces the stand-alone LayerNorm kernel.  Call before dvst_workspace_bytes. */
void dvst_set_ln_fusion(int on);

/* Tuning / test hook: 1 (default) runs the temporal half of a divided block (temporal_norm1 -> qkv -> attention) as ONE
 * kernel when the head-permuted folded tensors are registered (see dvst_gemm_ln_temporal_attention); 0 runs the qkv GEMM
 * and the temporal attention kernel separately.  For T in {16, 32, 64} both forms give the same bits. */
void dvst_set_temporal_fusion(int on);

/* Tuning / test hook: 1 launches the forward's hot-path kernels with programmatic dependent launch (a kernel's prologue
 * overlaps the tail of its predecessor; results are identical), 0 (default) uses ordinary stream-ordered launches. */
void dvst_set_pdl(int on);

/* L2-residency tuning: walk a batch in chunks of about `rows` token rows (0 = whole batch per pass).  Results are
 * bit-identical for any value (rows of every kernel are independent).  Call before dvst_workspace_bytes. */
void dvst_set_chunk_rows(long long rows);

/* Live per-kernel timing for bench.py's roofline: while enabled, every launch is bracketed by CUDA events on its own
 * stream.  dvst_profile_collect waits for the recorded events, sums elapsed ms / work (FLOPs for kind 0 GEMM and
 * 1 attention, bytes for 2 layernorm, 3 elementwise, 4 head-loss, 5 the HBM-bound short-K deferred-LayerNorm producer
 * GEMMs, FLOPs for 6 the fused qkv-GEMM + temporal-attention kernel) / launch counts per kind, and clears the log. */
void dvst_profile_enable(int on);
int dvst_profile_collect(double* ms, double* work, long long* count, int nkinds);

/* ---- model handle: replaces VisionTransformer.__init__/load_state_dict (models/timesformer.py:212-263) ----------- */
int dvst_create(const dvst_config* cfg, dvst_model** out);
void dvst_destroy(dvst_model* m);
/* Registers one PACKED tensor (device pointer, must stay alive while the handle is used).  Names:
 *   patch_embed.w  bf16 [768, 768*k] (conv weight flattened (c,ky,kx); k = 2 in split precision: [hi | lo])
 *   patch_embed.b, cls_pos0 (= cls_token + pos_embed[0]), norm.w, norm.b          fp32 [768]
 *   pos_embed fp32 [1 + pos_grid^2, 768], time_embed fp32 [num_frames, 768]
 *   blocks.<i>.{temporal_norm1,norm1,norm2}.{w,b}                                   fp32 [768]
 *   blocks.<i>.{temporal_attn.qkv,attn.qkv}.w bf16 [2304, 768*k], .b fp32 [2304]
 *   blocks.<i>.temporal_projfc.w bf16 [768, 768*k] = temporal_fc.weight @ temporal_attn.proj.weight, .b fp32 [768]
 *   blocks.<i>.attn.proj.{w,b}, blocks.<i>.mlp.fc1.{w [3072,768*k], b}, blocks.<i>.mlp.fc2.{w [768,3072*k], b}
 * Optional (bf16 precision; all of them or none): the deferred-LayerNorm form of the three Linears that follow a norm,
 *   blocks.<i>.{temporal_attn.qkv,attn.qkv,mlp.fc1}.wf bf16 [N, 768] = bf16(gamma * W), .sf fp32 [N] = row sums of wf,
 *   .bf fp32 [N] = b + W beta   (gamma / beta of temporal_norm1 / norm1 / norm2 respectively)
 * Optional on top of those: blocks.<i>.temporal_attn.qkv.{wfp,sfp,bfp} = the same three tensors with their 2304 rows
 *   permuted head-major (row 192 h + 64 part + d <- row 768 part + 64 h + d, part = q, k, v): enables the fused
 *   qkv-GEMM + temporal-attention kernel.                                                                               */
int dvst_set_weight(dvst_model* m, const char* name, const void* dev_ptr);
int dvst_finalize_weights(dvst_model* m);

/* Scratch bytes dvst_forward_features needs for a [B,3,T,H,W] batch (-1 on bad arguments). */
long long dvst_workspace_bytes(const dvst_model* m, int B, int T, int H, int W);

/* VisionTransformer.forward(x, use_head=False) / forward_features(x, get_all) - models/timesformer.py:285-351.
 *   x   fp32 [B,3,T,H,W] (NCTHW, H and W multiples of 16)
 *   out fp32 [B,768] (get_all = 0: LayerNorm'ed CLS feature) or [B, 1+N*T, 768] (get_all = 1, reference token order;
 *       [B, 1+N, 768] for space_only) */
int dvst_forward_features(dvst_model* m, const float* x, int B, int T, int H, int W, float* out, int get_all,
                          void* workspace, long long workspace_bytes, void* stream);

/* forward_features(x, get_attn=True) - models/timesformer.py:325-331,155-157: the softmax(q k^T / 8) map of the LAST
 * block's spatial attention, attn fp32 [B*T, 12, 1+N, 1+N] (CLS token first), divided_space_time models only. */
int dvst_forward_attention_map(dvst_model* m, const float* x, int B, int T, int H, int W, float* attn,
                               void* workspace, long long workspace_bytes, void* stream);

/* Device-side view construction + forward: `frames` is a frame-major video fp32 [F,3,H,W]; clip b is the window of T
 * consecutive frames starting at starts[b] (int32 device array, n_clips entries).  Replaces the materialised
 * [2F,3,G,224,224] view tensor of get_views_of_video_same_size (datasets_custom/dino_loss_loader.py:126-171) followed by
 * model(views) (dino_similarity.py:78-84).  out fp32 [n_clips,768].  Workspace as for B = n_clips. */
int dvst_forward_windows(dvst_model* m, const float* frames, int F, const int* starts, int n_clips, int T, int H,
                         int W, float* out, void* workspace, long long workspace_bytes, void* stream);

/* The same straight from DECODED frames: `video` uint8 [n_frames, Hf, Wf, 3] (device); the frame axis of `starts` is the
 * SAMPLED one (every sampling_rate-th decoded frame, datasets_custom/dino_loss_loader.py:80-92); normalisation (/ 255, - mean,
 * / std: the loader's three fp32 operations, data_utils.py:308-325), centre crop to crop x crop (transform.py:228-229), window
 * gather and patch extraction happen in ONE pass - bit-identical to dvst_preprocess_u8 followed by dvst_forward_windows, without
 * the normalised fp32 frames in between.  For clips cut from a stream of frames (every frame used once); whole-video scoring,
 * where each frame is reused by ~33 windows, normalises once with dvst_preprocess_u8.  mean3 / std3 are HOST pointers. */
int dvst_forward_windows_u8(dvst_model* m, const unsigned char* video, int n_frames, int Hf, int Wf, int sampling_rate,
                            const float* mean3, const float* std3, int crop, const int* starts, int n_clips, int T,
                            float* out, void* workspace, long long workspace_bytes, void* stream);

/* Scoring DINOLoss.forward for `pairs` rows at once - dino_similarity.py:129-135:
 *   loss[p] = sum_j -softmax((t[p]-center)/teacher_temp)_j * log_softmax(s[p]/student_temp)_j
 * student / teacher fp32 [pairs, dim] (row strides lds / ldt), center fp32 [dim] or NULL (zeros), loss fp32 [pairs]. */
int dvst_pair_loss(const float* student, long long lds, const float* teacher, long long ldt, const float* center,
                   float teacher_temp, float student_temp, int dim, int pairs, float* loss, void* stream);

/* "Mode H": DINOHead last layer (weight-normalised 256 -> out_dim, vision_transformer.py:266-282) fused with the
 * centred / sharpened softmax cross-entropy above; the out_dim-wide logits are never written to memory.
 *   zs, zt : bf16 [pairs, ldz] L2-normalised bottleneck vectors (256 wide) of the student / teacher views
 *   w      : bf16 [out_dim, 256] = weight_g * weight_v / ||weight_v||_row     (out_dim multiple of 64)
 *   center : fp32 [out_dim] or NULL;  loss fp32 [pairs];  scratch >= dvst_head_loss_scratch_bytes(pairs) bytes. */
long long dvst_head_loss_scratch_bytes(int pairs);
int dvst_head_loss(const void* zs, const void* zt, long long ldz, const void* w, const float* center,
                   float teacher_temp, float student_temp, int pairs, int out_dim, float* loss, void* scratch,
                   long long scratch_bytes, void* stream);

/* Same with row gathers: pair p uses student row s_idx[p] and teacher row t_idx[p] (int32 DEVICE arrays, both given or
 * both NULL; every index must lie inside its matrix - the kernel does not range-check, the Python layer validates dtype /
 * device always and the range under DVST_CHECK_INDICES=1).
 * Covers the per-frame pairing of dino_similarity.py:87-88 over de-duplicated window features and the (iq, v != iq)
 * terms of the training DINOLoss (train_ssl.py:660-667). */
int dvst_pair_loss_indexed(const float* student, long long lds, const int* s_idx, const float* teacher, long long ldt,
                           const int* t_idx, const float* center, float teacher_temp, float student_temp, int dim,
                           int pairs, float* loss, void* stream);

/* DINOLoss.update_center pieces (train_ssl.py:686-691): out[c] = sum_r in[r,c];  center = center*momentum +
 * (batch_sum/denom)*(1-momentum).  The all_reduce between them stays in torch.distributed (NCCL). */
int dvst_colsum(const float* in, long long ldi, float* out, int rows, int cols, void* stream);
int dvst_center_ema(float* center, const float* batch_sum, float denom, float momentum, int cols, void* stream);

/* fp32 [M,cols] -> bf16 GEMM operand (split: also the rounding residual at column + lo_off). */
int dvst_cast_bf16(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off, long long M,
                   int cols, void* stream);
/* F.normalize(x, dim=-1, p=2) (vision_transformer.py:280): bf16 (+lo) rows for the last-layer GEMM, optional fp32 copy. */
int dvst_l2norm_rows(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off,
                     float* out_f32, long long M, int cols, void* stream);

/* ---- building blocks (exported for unit tests and the DINO head) -------------------------------------------------- */
/* out[M,N] = epilogue(A[M,K] x W[N,K]^T).  bf16 operands, fp32 accumulate (tcgen05 / TMEM).
 * Replaces torch.nn.Linear as used by models/timesformer.py:41-52,63-64,71,83,108,133 and
 * vision_transformer.py:254-266.
 *   split    : 0 plain; 1 = operands hold [hi | lo] bf16 halves (lo at column K), 3-term product.
 *   bias     : fp32 [N] or NULL.   residual : NULL, or the fp32 output itself (residual == out, ldr == ldo, out_kind 1):
 *              the fused residual add is IN PLACE (out += A W^T + bias through a TMA reduce-add); any other pointer is
 *              rejected.
 *   out_kind : 0 bf16, 1 fp32, 2 bf16 hi/lo pair (lo written at column + lo_off).   act : 0 none, 1 exact GELU. */
int dvst_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias, int M, int N, int K,
                   int split, const float* residual, long long ldr, void* out, long long ldo, int out_kind,
                   long long lo_off, int act, void* stream);
/* Deferred LayerNorm (bf16 precision): the stand-alone LayerNorm pass between a residual add and the next Linear
 * (models/timesformer.py:127-128,140-141,184) is split between the two GEMMs around it.
 *   producer form (x_inout != NULL, N = 768): x += A W^T + bias in place (fp32), xb = bf16(x) [M, ldxb],
 *       stats_out[row][N / 128][2] = (sum, sum of squares) of the new row per 128-column block;
 *   consumer form (ln_stats != NULL, K = 768): out (bf16) = act(rstd (A W^T) - rstd mean colsum + bias) where A = xb,
 *       W = bf16(gamma * weight) row-major [N, 768], ln_colsum[c] = sum_k W[c, k], bias = b + weight beta  and
 *       (mean, rstd) come from ln_stats (eps inside the rsqrt as in nn.LayerNorm).
 * Row offsets (producer form only; xoff fp32 [M] or NULL = 0): xb = bf16(x - xoff[r]) and the statistics are those of
 * x - xoff[r]; LayerNorm is shift invariant, so the consumer needs nothing else, and both the bf16 rounding and the
 * variance act on centred values (rows with |mean| >> std lose nothing).  dvst_row_offsets768 moves xoff[r] to the current
 * row mean (xoff[r] += mean of the centred row) - call it right before a producer that rewrites the same rows.
 * dvst_rowstats768 produces (xb, stats, and the exact mean in xoff if given) for rows that no residual GEMM has written
 * (the CLS rows after embedding). */
int dvst_rowstats768(const float* x, long long ldx, void* xb, long long ldxb, float* stats, float* xoff, long long M,
                     void* stream);
int dvst_row_offsets768(const float* stats, float* xoff, long long M, void* stream);
int dvst_gemm_bf16_ln(const void* A, long long lda, const void* W, long long ldw, const float* bias, int M, int N, int K,
                      const float* ln_stats, const float* ln_colsum, float ln_eps, float* x_inout, long long ldx,
                      void* xb, long long ldxb, float* stats_out, const float* xoff, void* out, long long ldo, int act,
                      void* stream);
/* temporal_norm1 -> temporal_attn.qkv -> softmax(q k^T / 8) v of temporal_attn, heads concatenated
 * (models/timesformer.py:127-131 with Attention.forward :68-81, up to but not including temporal_attn.proj) in one kernel:
 * the deferred-LayerNorm consumer GEMM above with N = 2304, whose epilogue keeps q | k | v of a (128-row, one-head) tile in
 * shared memory and runs the temporal attention from there - the packed qkv tensor is never written.
 *   xb, ln_stats : as for the consumer form; M rows = whole sequences of T consecutive rows ((b, n, t) order), 1 <= T <= 64
 *   w_perm bf16 [2304, 768], bias_perm / colsum_perm fp32 [2304]: head-permuted (row 192 h + 64 part + d)
 *   clip_rows: 0, or the rows per clip (N T; a multiple of T that divides M): rows are then tiled clip by clip, so that the
 *       place of a sequence inside its MMA tile - and with it the rounding of its output - does not depend on the batch
 *   out bf16 [M, ldo]: attention output, head h in columns [64 h, 64 h + 64) - the A operand of the proj GEMM. */
int dvst_gemm_ln_temporal_attention(const void* xb, long long ldxb, const void* w_perm, const float* bias_perm,
                                    const float* colsum_perm, const float* ln_stats, float ln_eps, long long M, int T,
                                    long long clip_rows, void* out, long long ldo, void* stream);
/* Host-side geometry of that kernel (no device work): rows per CTA tile R = whole sequences (floor(128 / T) T, or T where a 16-row query tile would otherwise touch more than 64 rows)
 * and the key-block size (16, 32 or 64) that holds every sequence a 16-row query tile touches; DVST_ERR for T outside 1..64. */
int dvst_temporal_tile_geometry(int T, int* rows_per_tile, int* key_block);
/* nn.LayerNorm(768, eps) over M rows of fp32 (models/timesformer.py:99,105,112,245).  out_kind as above. */
int dvst_layernorm768(const float* x, long long ldx, const float* gamma, const float* beta, void* out, long long ldo,
                      int out_kind, long long lo_off, long long M, float eps, void* stream);
/* Attention.forward minus the linears (models/timesformer.py:77-81) on the packed qkv rows [*, ld] bf16.
 *   mode 0 temporal, 1 spatial (shared CLS row per clip, CLS outputs -> cls_out fp32 [B,T,768]), 2 joint,
 *   3 space-only; cls_base = index of the first CLS row; out bf16 [*, ldo]. */
int dvst_attention(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                   int N, long long cls_base, void* stream);

/* Same with scratch (>= dvst_attention_scratch_bytes(B, T, N) bytes, 16-byte aligned): modes 1 / 3 then run the tcgen05
 * kernel's one-pass softmax (reference point = maximum of the first 32 keys; rows it cannot normalise are recomputed
 * exactly in fp32 by a fix-up kernel, so results are exact for every input). */
long long dvst_attention_scratch_bytes(int B, int T, int N);
int dvst_attention_ws(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                      int N, long long cls_base, void* scratch, long long scratch_bytes, void* stream);

/* Split-precision form (precision = "high"): every operand is a bf16 [hi | lo] pair - qkv rows [q k v hi (2304) | q k v lo
 * (2304)] (ld >= 4608), out rows [hi (768) | lo (768)] (ldo >= 1536), lo = bf16(x - float(hi)) - and every product is
 * evaluated as hi*hi + hi*lo + lo*hi (~16 mantissa bits end to end; what reference-exact mode-R scoring needs).  Modes 1 / 3
 * with scratch and at most 208 key columns run the tcgen05 kernel of csrc/attention_hp_sm100.cu, everything else the
 * mma.sync split kernels.  cls_out (mode 1) stays fp32. */
int dvst_attention_hp(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                      int N, long long cls_base, void* scratch, long long scratch_bytes, void* stream);

/* Adaptive frame selection for a batch of videos (datasets_custom/frame_selection_loader.py:115-172) on the device, in
 * the reference's fp64 arithmetic (numpy pairwise sum, sequential cumsum, first-minimum argmin): identical index lists.
 *   loss        fp64 [n_videos, ld] device: per-sampled-frame losses of video v in row v (the JSON values; fp32 scores
 *               widen exactly)
 *   n_loss      int32 [n_videos] device: losses used per video = min(len(loss list), sampled frames)   (:125-126)
 *   n_unsampled int32 [n_videos] device: frames of the undecimated video (frames_unsampled.size(0))
 *   indices     int32 [n_videos, num_frames] device: selected frame indices in unsampled-video units; num_frames <= 64 */
int dvst_select_frames(const double* loss, long long ld, const int* n_loss, const int* n_unsampled, int n_videos,
                       int pre_sampling_rate, int num_frames, int* indices, void* stream);

/* Window plan of a batch of videos: get_views_of_video_same_size (datasets_custom/dino_loss_loader.py:126-171) as index
 * arithmetic on the device, for videos concatenated along the frame axis.  Replaces the [2F,3,G,224,224] view tensors and
 * the per-video Python loop: identical windows are listed once (F - local + 1 distinct local and F - G + 1 distinct global
 * windows per video, G = min(global, F)).
 *   lengths   HOST int [n_videos]: sampled frames per video, each >= local_clip_size (they travel as kernel parameters:
 *             no host-to-device copy, no synchronisation, capturable in a CUDA graph)
 *   s_starts  DEVICE int32 [sum(F - local + 1)]: first frame (concatenated axis) of every distinct local window, video-major
 *   t_starts  DEVICE int32 [sum(F - G + 1)]: the same for the global windows, grouped by G ascending, video-major inside
 *   s_idx, t_idx  DEVICE int32 [sum F]: per frame, the row of its local / global window in those lists
 *   n_s, n_t, n_groups, groups  HOST outputs: list lengths, number of distinct G, and (G, windows, first row) per group
 *             (groups holds 3 * n_videos ints at most) - what the caller needs to issue one dvst_forward_windows per G. */
int dvst_plan_windows(const int* lengths, int n_videos, int local_clip_size, int global_clip_size, int* s_starts,
                      int* t_starts, int* s_idx, int* t_idx, int* n_s, int* n_t, int* n_groups, int* groups, void* stream);

/* Loader front half (datasets_custom/dino_loss_loader.py:80-92) in one pass: decoded video uint8 [n_frames, H, W, 3]
 * (device) -> every sampling_rate-th frame, / 255, - mean, / std (data_utils.py:308-325), T H W C -> T C H W, centre crop
 * (transform.py:228-229): frames fp32 [ceil(n_frames / sampling_rate), 3, crop, crop].  mean3 / std3 are HOST pointers. */
int dvst_preprocess_u8(const unsigned char* video, int n_frames, int H, int W, int sampling_rate, const float* mean3,
                       const float* std3, int crop, float* frames, void* stream);

/* ---- weighted k-NN evaluation of frozen features: knn_classifier, eval_knn.py:138-178 -------------------------------
 * sim[i,j] = <test[i,:], train[j,:]> in fp32 FMA arithmetic (torch.mm(features, train_features.t()), :153); neighbour
 * order is the result, so no reduced-precision tensor-core product here.  dim and row strides multiples of 4. */
int dvst_knn_similarity(const float* test, long long ldt, const float* train, long long ldr, int n_test, int n_train,
                        int dim, float* sim, long long lds, void* stream);
/* similarity.topk(k, largest=True, sorted=True) per row (:154): dist fp32 / idx int32 [n_test, ldk], values descending,
 * equal values by ascending index; k <= 512.  The first k' columns are the top-k' for any k' <= k. */
int dvst_knn_topk(const float* sim, long long lds, int n_test, int n_train, int k, float* dist, int* idx, int ldk,
                  void* stream);
/* Votes of the k nearest neighbours weighted by exp(dist / temperature), summed per class in neighbour order (:155-166),
 * then rank[i] = position of test_labels[i] in probs.sort(descending) (:167; ties by ascending class index), so that
 * top1 = #(rank == 0), top5 = #(rank < 5) (:170-172).  pred int32 [n_test] (optional) = predictions[:, 0].
 * Labels are int64 device arrays (torch.long). */
int dvst_knn_vote(const float* dist, const int* idx, int ldk, const long long* train_labels,
                  const long long* test_labels, int n_test, int k, float temperature, int num_classes, int* rank,
                  int* pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVST_H_ */
