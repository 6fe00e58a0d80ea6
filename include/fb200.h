/* fb200.h - C ABI of libfb200.so, the B200 (sm_100a) fusion-head library.
 *
 * Drop-in boundary for ONE hot path of life-ufes/multimodal-model-skin-lesion-classifier:
 * the multimodal fusion head + MLP classifier + class-weighted cross-entropy, forward and
 * backward.  The reference has no FFI layer: its boundary is the Python class
 *   MultimodalModel(...)            src/scripts/benchmark/models/multimodalIntraInterModal.py:13-28
 *   MultimodalModel.forward         ... :162-416
 *   nn.CrossEntropyLoss(weight=w)   src/scripts/benchmark/train_pad_20.py:52,111
 * The Python host (fusion_b200.MultimodalModel) keeps that class' constructor, parameter
 * names and fusion strings and calls the entry points below through ctypes with raw device
 * pointers.  No torch types cross this boundary.  Every function returns an int status
 * (0 = OK, negative = FB200_E*), never throws, never exits; there is NO CPU fallback - a
 * host pointer or a missing GPU is FB200_EUNSUPPORTED / FB200_ECUDA.
 *
 * Threading: the library keeps no mutable global state besides a per-device cache of
 * immutable kernel attributes and, per host thread and device, one internal side stream with
 * a pool of timing-disabled events.  Calls are re-entrant (forward is called on the Python
 * main thread, backward on autograd's device thread).  All work is ordered on the stream
 * handed in: the head entry points may fork part of it onto the side stream, but join it back
 * before they return (FB200_FLAG_ONE_STREAM disables the fork), so they are CUDA-graph
 * capturable and callers see one stream.  Every kernel is launched with the
 * programmatic-stream-serialization attribute and waits (griddepcontrol.wait) before its first
 * global-memory access.  Environment switches for A/B measurements, read once: FB200_PDL=0
 * (plain launches), FB200_LANES=0 (one stream), FB200_TC_DBG=<bits> (timing-only GEMM
 * ablations; results are wrong by construction).
 */
#ifndef FB200_H
#define FB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_VERSION 100

/* status codes */
#define FB200_OK            0
#define FB200_EBADARG      -1   /* null pointer / inconsistent descriptor                 */
#define FB200_EUNSUPPORTED -2   /* shape / dtype / device the kernels do not cover        */
#define FB200_EALIGN       -3   /* pointer or leading dimension breaks a 16-byte contract */
#define FB200_ECUDA        -4   /* a CUDA runtime / driver call failed                    */
#define FB200_EABSENT      -5   /* parameter slot does not exist in this configuration    */

/* compute dtypes (fb200_desc.dtype).  Parameters and the API tensors are always fp32. */
#define FB200_F32   0   /* fp32 storage; GEMMs either FFMA (exact fp32) or 3xTF32 on tcgen05 */
#define FB200_BF16  1   /* bf16 operands, fp32 accumulation, fp32 LayerNorm / loss          */

/* fusion strings of MultimodalModel.forward (multimodalIntraInterModal.py:205-416), in the
 * order they are tested there.  fb200_mechanism_from_string() maps the literal strings. */
enum fb200_mechanism {
  FB200_NO_METADATA = 0,              /* :205 */
  FB200_NO_METADATA_WITHOUT_MLP,      /* :208 */
  FB200_CONCATENATION,                /* :211 */
  FB200_CROSSATTENTION,               /* :215 */
  FB200_WEIGHTED,                     /* :219 */
  FB200_GFCAM,                        /* :225 */
  FB200_CROSS_WEIGHTS_AFTER_CROSSATT, /* :231 */
  FB200_METABLOCK,                    /* :237 */
  FB200_RGATT2FUSEFEATURES,           /* :247 */
  FB200_RG_ATT,                       /* :253 */
  FB200_ATT_INTRAMODAL,               /* :265 */
  FB200_ATT_INTRAMODAL_RESIDUAL,      /* :273 */
  FB200_CROSS_ATTENTION_ONLY,         /* :285 */
  FB200_RESIDUAL_CROSSATT,            /* :301 */
  FB200_RGATT_FULL,                   /* :322 "att-intramodal+residual+cross-attention-metadados" */
  FB200_RGATT_FULL_RGATT2FUSE,        /* :343 */
  FB200_RGATT_FULL_METABLOCK,         /* :364 */
  FB200_RGATT_FULL_INTRAMODAL_RES,    /* :388 */
  FB200_NUM_MECHANISMS
};

/* fb200_desc.flags */
#define FB200_FLAG_NEED_DIMG   1   /* backward also produces d(img_feat) (backbone unfrozen) */
#define FB200_FLAG_NEED_DTEXT  2   /* backward also produces d(text_in) (trainable encoder)  */
#define FB200_FLAG_FORCE_SIMT  4   /* never use the tcgen05 GEMM (exact-fp32 FFMA everywhere) */
#define FB200_FLAG_FORCE_TC    8   /* use the tcgen05 GEMM wherever its shape rules allow     */
#define FB200_FLAG_NO_MEGA    32   /* small fp32 batches: keep the per-op kernels instead of the persistent step kernel */
#define FB200_FLAG_FORCE_MEGA 64   /* fp32: use the persistent step kernel up to its capacity of 64 rows (default: up to 32 rows;
                                      from 33 rows the tcgen05 GEMMs with cluster split-K are faster - measured) */
#define FB200_FLAG_ONE_STREAM 16   /* launch everything on the caller's stream (default: the metadata chain of large
                                      batches runs on an internal side stream, forked from and joined back into the
                                      caller's stream inside the call - graph-capturable, invisible to the caller) */

/* dropout sites, reference call order (nn.Dropout modules reached by forward) */
enum fb200_dropout_site {
  FB200_DROP_IMG_RES = 0,   /* image_residual.dropout  p=0.1  [B,D]   gatedResidualBlock.py:9,14 */
  FB200_DROP_TXT_RES,       /* text_residual.dropout   p=0.1  [B,D]                              */
  FB200_DROP_IMG_RES2,      /* second call of image_residual (strings :343, :388)                */
  FB200_DROP_TXT_RES2,      /* second call of text_residual  (string :388)                       */
  FB200_DROP_FC1,           /* fc_fusion[3] p=0.5 / after-metablock MLP[3] p=0.3  [B,D]   :139,153 */
  FB200_DROP_FC2,           /* fc_fusion[7] p=0.5 / after-metablock MLP[7] p=0.3  [B,D/2] :143,157 */
  FB200_NUM_DROPOUT_SITES
};

/* One head instance = the reference constructor arguments that shape the path
 * (multimodalIntraInterModal.py:14-28) plus the batch and the execution mode. */
typedef struct fb200_desc {
  int32_t mechanism;   /* enum fb200_mechanism  <- attention_mecanism                     */
  int32_t B;           /* rows in this call (any B >= 1)                                  */
  int32_t F;           /* cnn_dim_output: width of img_feat                               */
  int32_t V;           /* vocab_size: width of the one-hot metadata (text_mode 0)         */
  int32_t T;           /* text_encoder_dim_output (512 one-hot, 85 tab-transformer)       */
  int32_t D;           /* common_dim (multiple of 8; the residual blocks hard-code 8 heads) */
  int32_t H;           /* num_heads (only validated: D % H == 0, as nn.MultiheadAttention) */
  int32_t C;           /* num_classes                                                     */
  int32_t n;           /* ctor argument n (fc_fusion input = n*D; the six strings need 2) */
  int32_t text_mode;   /* 0: text_in = [B,V] one-hot -> text_fc; 1: text_in = [B,T] encoder output */
  int32_t dtype;       /* FB200_F32 | FB200_BF16                                          */
  int32_t train;       /* 1: dropout active (masks or Philox), 0: eval                    */
  int32_t flags;       /* FB200_FLAG_*                                                    */
  int32_t reserved;
} fb200_desc;

/* ---- introspection (host only, no GPU needed) ------------------------------------- */
int          fb200_version(void);
const char*  fb200_strerror(int status);
int          fb200_mechanism_from_string(const char* attention_mecanism); /* -1: unknown */
const char*  fb200_mechanism_string(int mechanism);
int          fb200_num_params(void);                 /* 78 slots, reference state_dict order */
const char*  fb200_param_name(int slot);             /* e.g. "image_projector.weight"        */
/* rows/cols of a slot under desc (cols = 0 for 1-D tensors); FB200_EABSENT if the slot does
 * not exist (text_fc.* when text_mode = 1). */
int          fb200_param_shape(const fb200_desc* d, int slot, int64_t* rows, int64_t* cols);
/* Element offset of the slot's gradient inside the flat gradient buffer, or -1 when the
 * reference leaves .grad = None for it under this mechanism. */
int64_t      fb200_grad_offset(const fb200_desc* d, int slot);
int64_t      fb200_grad_elems(const fb200_desc* d);  /* size of the flat gradient buffer (fp32 elements) */
/* Element ranges [begin, end) of the flat gradient buffer that can be non-zero (the W_q / W_k rows of S=1 attention are
 * structural zeros): writes up to cap pairs into out, returns the number of ranges.  The DP all-reduce moves only these. */
int          fb200_grad_live_ranges(const fb200_desc* d, int64_t* out, int cap);
int          fb200_workspace_bytes(const fb200_desc* d, size_t* bytes);
float        fb200_dropout_p(const fb200_desc* d, int site);   /* 0 when the site is unused */
int          fb200_dropout_shape(const fb200_desc* d, int site, int64_t* rows, int64_t* cols);
/* Algorithmic work of one train step (forward + backward) under desc: FLOPs = 6*B*MAC_live
 * minus the dX GEMMs nobody needs, bytes = 3*P_live*4 + input/grad traffic. */
int          fb200_algorithmic_work(const fb200_desc* d, double* flops, double* bytes, int64_t* live_params);
/* Number of kernel launches one forward / one backward issues (for bench.py's gpu_launches). */
int          fb200_launch_count(const fb200_desc* d, int* forward, int* backward);

/* GEMM launches of one train step: writes up to cap entries of 5 int32 {layout, engine, M, N, K}
 * (layouts / engines as in fb200_gemm) and returns how many were written. */
int          fb200_list_gemms(const fb200_desc* d, int32_t* out, int cap);

/* ---- the hot path ----------------------------------------------------------------- */
/* All pointers are DEVICE pointers on the current device; `stream` is a cudaStream_t.
 * params[slot]  : fp32 parameter tensors, contiguous, reference shapes (NULL for absent slots)
 * img_feat      : [B,F] fp32 row-major (output of image_encoder, :167-170)
 * text_in       : [B,V] (text_mode 0) or [B,T] (text_mode 1) fp32 row-major
 * masks[site]   : optional uint8 keep-masks ({0,1}, row-major, fb200_dropout_shape); when
 *                 train = 1 and masks == NULL (or masks[site] == NULL) the kernels draw the
 *                 mask from Philox4x32-10 keyed by (seed, offset, site, element)
 * rng_state     : optional device uint64[2] = {seed, offset}; when non-NULL the kernels read the
 *                 Philox key from it at run time instead of the immediate seed/offset, so a captured
 *                 CUDA graph draws fresh masks on every replay (advance it with fb200_rng_advance)
 * logits        : [B,C] fp32 out
 * ws            : fb200_workspace_bytes() bytes, 256-byte aligned; must stay untouched
 *                 between forward and the matching backward (it holds the saved activations) */
int fb200_head_forward(const fb200_desc* d, const void* const* params,
                       const void* img_feat, const void* text_in,
                       const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                       void* logits, void* ws, void* stream);

/* Debug / measurement aid: launches only the GEMM kernels of one train step (same kernels, grids and operands as
 * fb200_head_train_step) so that bench.py can time the dominant kernel family with CUDA events on the launch stream. */
int fb200_debug_gemm_replay(const fb200_desc* d, const void* const* params, const void* img_feat, const void* text_in,
                            void* logits, void* grads, void* ws, void* stream);

/* Debug aid (not part of the drop-in surface): record clock64 stamps of the TMA / MMA pipeline of
 * CTA (0,0,0) of each later tcgen05 GEMM into device_buf (int64[8 * k_blocks]); NULL switches it off. */
/* debug / measurement aid: programmatic dependent launch on (default; FB200_PDL=0 in the environment turns it off) or
 * off for every kernel launched afterwards; returns the previous setting. */
int fb200_debug_set_pdl(int on);
int fb200_debug_tc_trace(void* device_buf);
/* debug: grid-wide timeline of the tcgen05 GEMM launches that follow.  Launch l (in host launch order, grouped weight-gradient
 * launches excluded) writes, for each of its CTAs, int64[8] = {entry, dependency wait passed, tile stored (globaltimer ns), SM id, accumulator in
 * registers, tensor memory released, accumulator read out by every worker, MMA warp left the k-loop} at device_buf + (l * 1024 + cta) * 8; launches with more than 1024 CTAs are skipped.  max_launches = slots in device_buf; NULL
 * switches it off.  Returns the number of launches recorded since the previous call. */
int fb200_debug_tc_timeline(void* device_buf, int max_launches);
/* debug: clock64 stamps of CTA 0 of the persistent step kernel (>= 2 + 2 * stages int64; NULL disables), and that kernel
 * with `nstages` empty stages (launch + grid-barrier cost alone; ws256: 256 bytes of device memory) */
int fb200_debug_mega_trace(void* device_buf);
/* debug: clock64 stamps of CTA 0 of the TabTransformer kernels after every phase of its first sample (int64[32]; NULL disables) */
int fb200_debug_tabt_trace(void* device_buf);
/* host only: shape of the program the persistent step kernel (fp32, B <= 32; B <= 64 with FB200_FLAG_FORCE_MEGA) runs for `d`; pass 0 forward, 1 backward, 2 fused
 * train step; out[4] = stages, GEMM ops, row ops, tile tasks.  FB200_EUNSUPPORTED when `d` takes the per-op kernels. */
int fb200_mega_program_info(const fb200_desc* d, int pass, int* out);
int fb200_debug_mega_barriers(int nstages, void* ws256, void* stream);

/* rng_state[1] += increment, on `stream` (one tiny kernel; graph-capturable). */
int fb200_rng_advance(void* rng_state, uint64_t increment, void* stream);

/* dlogits : [B,C] fp32.  grads : flat fp32 buffer of fb200_grad_elems() elements; the call
 * overwrites it (rows of in_proj_weight/bias that belong to W_q/W_k are written as zeros,
 * exactly like autograd does for S=1 attention).  d_img_feat / d_text_in : [B,F] / [B,V|T]
 * fp32 out, required iff the matching FB200_FLAG_NEED_* bit is set, else may be NULL. */
int fb200_head_backward(const fb200_desc* d, const void* const* params,
                        const void* img_feat, const void* text_in,
                        const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                        const void* dlogits, void* grads, void* d_img_feat, void* d_text_in,
                        void* ws, void* stream);

/* Fused log-softmax + class-weighted NLL (nn.CrossEntropyLoss(weight, reduction='mean')),
 * forward and dlogits in one launch pair.
 * class_w  : [C] fp32 or NULL (all ones).
 * denom    : device scalar holding the global sum_i w[y_i] (data-parallel runs), or NULL to
 *            use this batch's own sum.
 * loss_out : device float[3] = { loss, numerator, local weight sum }.
 * dlogits  : [B,C] fp32 out, or NULL for loss only. */
int fb200_cross_entropy(const void* logits, const int64_t* labels, const float* class_w,
                        const float* denom, int B, int C, float* loss_out, void* dlogits,
                        void* stream);

/* forward + cross-entropy + backward in one call (what model.forward_loss / bench.py use). */
int fb200_head_train_step(const fb200_desc* d, const void* const* params,
                          const void* img_feat, const void* text_in,
                          const int64_t* labels, const float* class_w, const float* denom,
                          const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                          void* logits, float* loss_out, void* grads,
                          void* d_img_feat, void* d_text_in, void* ws, void* stream);

/* Data-parallel gradient SUM all-reduce over NVLink / NVSwitch in ONE kernel, in place, over the live ranges of the flat
 * gradient buffer only (the reference has no collective - single process, train_pad_20.py:509).  The buffer
 * of every rank lives in symmetric memory: `multicast_ptr` = the NVSwitch multicast address aliasing all ranks' buffers
 * (multimem.ld_reduce / multimem.st: the switch adds and broadcasts), or NULL with `peer_ptrs[world]` = every rank's
 * buffer mapped into this process (plain peer loads / stores).  ranges: nranges pairs [begin, end) of element offsets
 * (multiples of 4, ascending; fb200_grad_live_ranges).  The caller brackets the call with two cross-rank barriers on the
 * stream: every rank's gradients complete before, every rank's stores complete after.  max_ctas: 0 = full grid; a small
 * value (e.g. 32) keeps the kernel beside a GEMM it overlaps with (256-thread CTAs, no shared memory).
 * signal_pads[world] != NULL: the two barriers run INSIDE the kernel on uint32 slots [slot, slot + 2 * world) of every rank's
 * zero-initialised, peer-mapped signal pad (`state`: 16 bytes of local device memory for the kernel's own bookkeeping); NULL:
 * the caller brackets the call with its own barriers. */
int fb200_dp_allreduce(void* multicast_ptr, void* const* peer_ptrs, const int64_t* ranges, int nranges, int rank, int world, int max_ctas,
                       void* const* signal_pads, void* state, int slot, void* stream);

/* Data-parallel variant of fb200_head_train_step (the head's gradients are all-reduced after every step).
 * Every gradient whose offset in the flat buffer is below fb200_dp_bucket_split(d) is final when `mid_event`
 * (a cudaEvent_t; recorded on `stream`, as an external event-record node when the stream is being captured) fires;
 * the tcgen05 weight gradients above the split are launched after it.  The caller waits for the event on its
 * communication stream and all-reduces the first bucket while the second half of the weight gradients is computed.
 * mid_event = NULL: identical to fb200_head_train_step.  fb200_dp_bucket_split: 0 when the configuration has nothing to
 * split (FFMA path, weights applied twice) - the event then fires at the end of the step. */
int fb200_head_train_step_dp(const fb200_desc* d, const void* const* params, const void* img_feat, const void* text_in,
                             const int64_t* labels, const float* class_w, const float* denom,
                             const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                             void* logits, float* loss_out, void* grads, void* d_img_feat, void* d_text_in, void* ws, void* stream,
                             void* mid_event);
int64_t fb200_dp_bucket_split(const fb200_desc* d);

/* The other losses the reference's loops use, forward + dlogits in one launch, mean reduction folded in:
 * kind 1 = FocalLoss(alpha, gamma) (models/focalLoss.py:6-26; targets = int64 labels [B], weight = alpha [C] or NULL)
 * kind 2 = SoftTargetCrossEntropy(weight) (models/softtargetsCrossEntropy.py:5-22; targets = fp32 soft labels [B,C]).
 * loss_out: device float; dlogits: [B,C] fp32 out or NULL. */
int fb200_aux_loss(int kind, const void* logits, const void* targets, const float* weight, float gamma, int B, int C,
                   float* loss_out, void* dlogits, void* stream);
/* Metadata one-hot + StandardScaler on the device: replaces np.hstack((ohe.transform(categorical), scaler.transform(
 * numerical))) of SkinLesionDataset.one_hot_encoding (models/skinLesionDatasets.py:133-176; same construction in
 * skinLesionDatasetsISIC2019.py / ...ISIC2020.py).  codes [B, n_cat] int32: index of each categorical value inside its
 * column's fitted category list, < 0 for an unknown value (handle_unknown='ignore' -> all-zero group); col_of [cat_total]:
 * the categorical column an output position belongs to; col_base [n_cat]: first output position of a column; numeric
 * [B, n_num] float64 raw values (NaN already replaced by -1 as at :152); mean / scale [n_num] float64 = StandardScaler.mean_
 * / .scale_.  out [B, cat_total + n_num] fp32, bit-identical to the float64 scikit-learn result rounded to fp32. */
int fb200_metadata_encode(const int32_t* codes, const int32_t* col_of, const int32_t* col_base, const double* numeric,
                          const double* mean, const double* scale, int B, int n_cat, int cat_total, int n_num, float* out, void* stream);

/* Evaluation tail: probs = softmax(logits) and pred = argmax (utils/model_metrics.py:57-58); either output may be NULL. */
int fb200_softmax_argmax(const void* logits, int B, int C, void* probs, int64_t* pred, void* stream);

/* Fused multi-tensor Adam step with torch.optim.Adam semantics (coupled L2 weight decay, bias correction, eps
 * outside the square root) - the optimizer the reference builds right after the path: optim.Adam(model.parameters(),
 * lr=5e-5, weight_decay=1e-4) (train_pad_20.py:54) and steps at :113.  Host arrays of ntensors device pointers;
 * `step` is the 1-based step count; grads are multiplied by grad_scale first (1/world_size style rescaling). */
int fb200_adam_step(int ntensors, void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                    const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                    float grad_scale, void* stream);

/* ---- primitives (exported for unit parity tests and micro-benchmarks) --------------- */
/* C[M,N] (+)= op(A) * op(B) (+ bias) with fp32 tensors.
 * layout: 0 = NT  C = A[M,K] * B[N,K]^T   (nn.Linear forward)
 *         1 = NN  C = A[M,K] * B[K,N]     (dX = dY * W)
 *         2 = TN  C = A[K,M]^T * B[K,N]   (dW = dY^T * X)
 * engine: 0 = FFMA SIMT kernel, 1 = tcgen05 3xTF32, 2 = tcgen05 bf16 (operands rounded to bf16)
 * bias: [N] or NULL; relu: apply max(.,0); accumulate: C += result. */
int fb200_gemm(int layout, int engine, int M, int N, int K,
               const float* A, int lda, const float* B, int ldb, float* C, int ldc,
               const float* bias, int relu, int accumulate, void* ws, size_t ws_bytes, void* stream);
int fb200_gemm_workspace_bytes(int layout, int engine, int M, int N, int K, size_t* bytes);

/* ---- multi-head attention on token sequences ------------------------------------------------------------------
 * Replaces torch.nn.MultiheadAttention(embed_dim = D, num_heads = H, batch_first = False, dropout = 0) as the
 * reference's sequence models call it (models/multimodalGated.py:118-206: image tokens attending to metadata tokens;
 * models/multimodalIntraInterModal.py:78-100 builds the same module and calls it with S = 1, which the head entry
 * points above lower to two GEMMs).  query [Sq,B,D], key / value [Skv,B,D], out [Sq,B,D], all contiguous fp32;
 * in_proj_weight [3D,D] = [W_q; W_k; W_v], in_proj_bias [3D], out_proj_weight [D,D], out_proj_bias [D].
 * The head-averaged attention weights the module can also return are not produced (every caller discards them).
 * ws: fb200_mha_workspace_bytes() bytes, 256-byte aligned; forward leaves Q, K, V, the head outputs and the row
 * log-sum-exps there for backward (the probabilities are recomputed, never stored).  dquery / dkey / dvalue may be
 * NULL (not needed).  When query, key and value are one tensor the caller adds the three input gradients. */
#define FB200_MHA_POOL_MEAN 1   /* out / dout are [B, D]: the mean over the S_q query tokens (models/multimodalGated.py:200-205:
                                 * cross_att.permute(1,0,2).mean(dim=1)), folded in front of the output projection it commutes with */
typedef struct { int32_t Sq, Skv, B, D, H, flags; } fb200_mha_desc;
int fb200_mha_workspace_bytes(const fb200_mha_desc* d, size_t* bytes);
int fb200_mha_forward(const fb200_mha_desc* d, const float* query, const float* key, const float* value,
                      const float* in_proj_weight, const float* in_proj_bias, const float* out_proj_weight, const float* out_proj_bias,
                      float* out, void* ws, void* stream);
int fb200_mha_backward(const fb200_mha_desc* d, const float* query, const float* key, const float* value,
                       const float* in_proj_weight, const float* out_proj_weight, const float* dout,
                       float* dquery, float* dkey, float* dvalue, float* d_in_proj_weight, float* d_in_proj_bias,
                       float* d_out_proj_weight, float* d_out_proj_bias, void* ws, void* stream);

/* ---- TabTransformer encoder (models/tab_transformer.py:6-60) ------------------------------------------
 * Replaces, for the whole batch in ONE launch per pass, the embedding lookups + torch.stack (:42-43), the
 * nn.TransformerEncoder stack (:19-27, :46: post-norm layers, ReLU feed-forward, dropout p on the attention
 * probabilities, after the attention, inside the feed-forward and after it) and the flatten (:47).
 *   B samples, T categorical columns (tokens), D = embed_dim, H heads, F = dim_feedforward, L layers.
 *   codes    [B, T] int64 category codes (x_categorical), emb_base[t] = first row of column t's table inside the
 *            stacked embedding table (n_emb_rows rows = sum of the cardinalities);
 *   params   one flat fp32 buffer: L layer blocks, each in nn.TransformerEncoderLayer's state_dict order
 *            (self_attn.in_proj_weight [3D,D], in_proj_bias [3D], out_proj.weight [D,D], out_proj.bias [D],
 *            linear1.weight [F,D], linear1.bias [F], linear2.weight [D,F], linear2.bias [D], norm1.weight, norm1.bias,
 *            norm2.weight, norm2.bias [D each]), then embeddings.0.weight ... embeddings.T-1.weight stacked
 *            [n_emb_rows, D]; fb200_tabt_param_elems gives the block and total sizes;
 *   out      [B, T*D] with row stride ldo >= T*D (the caller may hand in the left part of the concatenated feature matrix);
 *   saved    fb200_tabt_workspace_bytes(saved_bytes): the inputs of layers 1..L-1, written by forward, read by backward
 *            (everything else is recomputed on chip);
 *   masks    NULL, or 4 optional uint8 keep-masks {attention [L,B,H,T,T], after-attention [L,B*T,D], feed-forward
 *            [L,B*T,F], after-feed-forward [L,B*T,D]}; a NULL entry draws from Philox (seed, offset | rng_state) when train = 1;
 *   dparams  gradient of `params`, same layout, OVERWRITTEN (bit-reproducible: no atomics); ws = bwd_ws_bytes scratch.
 * 16-byte aligned params / out / saved / dout / dparams / ws.  FB200_EUNSUPPORTED when a sample's working set does not
 * fit the 227 KB of one SM, D or F is not a multiple of 4, or D / H is not 2, 4, 8, 16 or 32. */
#define FB200_TABT_MAX_CTAS 160
typedef struct { int32_t B, T, D, H, F, L, n_emb_rows, train; float p; int32_t flags; } fb200_tabt_desc;
int fb200_tabt_param_elems(const fb200_tabt_desc* d, int64_t* layer_elems, int64_t* total_elems);
int fb200_tabt_workspace_bytes(const fb200_tabt_desc* d, size_t* saved_bytes, size_t* bwd_ws_bytes);
int fb200_tabt_forward(const fb200_tabt_desc* d, const int64_t* codes, const int32_t* emb_base, const float* params,
                       const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                       float* out, int ldo, void* saved, void* stream);
int fb200_tabt_backward(const fb200_tabt_desc* d, const int64_t* codes, const int32_t* emb_base, const float* params,
                        const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                        const void* saved, const float* dout, int lddo, float* dparams, void* ws, void* stream);

/* nn.Linear on row-major fp32 tensors with explicit row strides (tab_transformer.py:30 numeric_projection, :33-38 fc):
 * y[M,N] = x[M,K] W[N,K]^T + bias, optionally followed by ReLU; the engine (tcgen05 3xTF32 / FFMA) is chosen per shape.
 * backward: dx = dy W (dx may be NULL), dW = dy^T x and db = column sums of dy (both OVERWRITTEN). */
int fb200_linear_forward(int M, int N, int K, const float* x, int ldx, const float* W, const float* bias, int relu,
                         float* y, int ldy, void* stream);
int fb200_linear_backward(int M, int N, int K, const float* x, int ldx, const float* W, const float* dy, int lddy,
                          float* dx, int lddx, float* dW, float* db, void* stream);

/* y = dropout(relu(LayerNorm(x))), rows of width N (fc_fusion[1:4], [5:8]); stats = [B,2] (mean, rstd) */
int fb200_ln_relu_dropout_fwd(const float* x, const float* gamma, const float* beta,
                              const uint8_t* mask, float p, int train, uint64_t seed, uint64_t offset, int site,
                              int B, int N, float* y, float* stats, void* stream);
int fb200_ln_relu_dropout_bwd(const float* x, const float* y, const float* gamma, const float* stats,
                              const float* dy, float p, int train, int B, int N,
                              float* dx, float* dgamma, float* dbeta, void* stream);
/* MetaBlock modulation: y = sigmoid(tanh(v * LN(f)) + LN(g))  (metablock.py:22-32) */
int fb200_metablock_fwd(const float* v, const float* f, const float* g,
                        const float* gamma_f, const float* beta_f, const float* gamma_g, const float* beta_g,
                        int B, int N, float* y, float* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FB200_H */
