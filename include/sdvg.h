/* sdvg.h - C ABI of libsdvg.so: the B200-native (sm_100a) implementation of the sd-video-gen rollout hot path.
 *
 * This is the drop-in boundary.  The reference (jeremy-collins/sd-video-gen) exposes this path only as a
 * Python nn.Module; every entry point below names the reference interface it stands in for.  Plain C types,
 * device pointers and sizes only - no torch types.  The Python mirror of the reference API
 * (sd-video-gen_b200/transformer.py, predict.py) binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every call returns 0 on success or a negative sdvg_status; sdvg_last_error() gives the message.
 *     No C++ exception crosses the ABI.  There is NO CPU fallback: without a CUDA device every compute
 *     entry point fails with SDVG_ERR_CUDA.
 *   - a handle is bound to one device and is not thread-safe.  All work is enqueued on the caller's
 *     cudaStream_t (passed as void*); no call synchronises the device except sdvg_create/sdvg_destroy,
 *     sdvg_finalize_weights and sdvg_timing_read.
 *   - the caller owns every input/output buffer; the library owns its packed weights and workspace, all
 *     allocated in sdvg_create (nothing is allocated on the hot path).
 *   - tensors are fp32, row-major, clip-major ("batch first") unless stated otherwise.
 */
#ifndef SDVG_H_
#define SDVG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDVG_VERSION 100 /* 0.1.0 */

typedef enum sdvg_status {
  SDVG_OK = 0,
  SDVG_ERR_INVALID = -1,   /* bad argument / shape outside the handle's limits                          */
  SDVG_ERR_CUDA = -2,      /* CUDA runtime error, no device, or wrong architecture (needs sm_100)        */
  SDVG_ERR_STATE = -3,     /* e.g. forward before all weights were set                                   */
  SDVG_ERR_BATCH = -4,     /* B > 64 without pe_index: the reference raises here too (see sdvg_forward)  */
  SDVG_ERR_UNSUPPORTED = -5
} sdvg_status;

/* GEMM operand precision.  All accumulate in fp32; LayerNorm, softmax, residual stream are fp32 in every mode. */
typedef enum sdvg_precision {
  SDVG_FP32_SIMT = 0, /* CUDA-core fp32 FMA GEMMs: bring-up / cross-check mode                                   */
  SDVG_FP32 = 1,      /* tensor cores, split operands (fp16 hi + 2^-11 fp16 lo, 3 MMAs): <=1e-4 parity mode       */
  SDVG_FP16 = 2,      /* tensor cores, single fp16 operands: the full-rate mode that meets the 5e-3 bar          */
  SDVG_BF16 = 3,      /* tensor cores, single bf16 operands (same rate; misses 5e-3 on this model) */
  SDVG_MIXED = 4      /* SDVG_FP16 with the embedding and layer-0 Q/K/V projections in split precision            */
} sdvg_precision;

/* Architecture of the model: the arguments of Transformer.__init__ (models/transformer.py:12-45) plus the
 * latent width E = 4*(FRAME_SIZE/8)^2 that the reference reads from its yaml (models/transformer.py:28-29,37). */
typedef struct sdvg_config {
  int32_t dim_model;          /* d          (models/transformer.py:15)                                 */
  int32_t num_heads;          /* H          (:16)                                                      */
  int32_t num_encoder_layers; /* Le         (:17)                                                      */
  int32_t num_decoder_layers; /* Ld         (:18)                                                      */
  int32_t latent_dim;         /* E          (:37,45)                                                   */
  int32_t dim_feedforward;    /* 2048: nn.Transformer default, never overridden by the reference (:38-44) */
  float layer_norm_eps;       /* 1e-5                                                                 */
  int32_t max_clips;          /* largest B of any call                                                */
  int32_t max_tokens;         /* largest S (window incl. SOS) of any call, <= 32                      */
  int32_t max_history;        /* rollout: context + predicted frames per clip (0: forward only)        */
  int32_t precision;          /* sdvg_precision                                                       */
  int32_t device;             /* CUDA device ordinal                                                  */
} sdvg_config;

typedef struct sdvg_handle sdvg_handle;

/* Library version (SDVG_VERSION). */
int sdvg_version(void);

/* Message for the last error on this handle (or of the last failed sdvg_create when h == NULL). */
const char* sdvg_last_error(const sdvg_handle* h);

/* Replaces Transformer.__init__ (models/transformer.py:12-45): allocates packed-weight storage and workspace. */
int sdvg_create(const sdvg_config* cfg, sdvg_handle** out);
void sdvg_destroy(sdvg_handle* h);

/* Bytes of device memory owned by the handle (weights + workspace). */
int sdvg_workspace_bytes(const sdvg_handle* h, size_t* bytes);

/* Replaces nn.Module.load_state_dict (prediction/predict.py:51, trainers/trainer.py:362-363): one call per
 * state_dict entry, `key` being the reference's own key ("embedding.weight",
 * "transformer.encoder.layers.0.self_attn.in_proj_weight", ..., "positional_encoder.pos_encoding", "out.bias").
 * `data` is fp32, contiguous, host or device memory; it is copied (caller keeps ownership).  `shape`/`ndim`
 * are checked against the architecture. */
int sdvg_set_weight(sdvg_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim);

/* Number of state_dict entries the architecture expects / the i-th key (for completeness checks). */
int sdvg_num_weights(const sdvg_handle* h);
const char* sdvg_weight_key(const sdvg_handle* h, int32_t i);

/* Packs the operand planes of all weights for the configured precision (called implicitly by the first
 * forward/rollout; explicit call lets a caller keep it out of a timed region).  Synchronises `stream`. */
int sdvg_finalize_weights(sdvg_handle* h, void* stream);

/* Replaces Transformer.forward(src, tgt, tgt_mask) (models/transformer.py:47-68) in eval mode.
 *   src (B, S_src, E), tgt (B, S_tgt, E) device fp32.  out (S_tgt, B, E) device fp32 - sequence-first like
 *   the reference's return value.
 *   mask_kind 0: none; 1: causal (== get_tgt_mask(S_tgt), models/transformer.py:70-89, never materialised);
 *             2: additive fp32 device matrix `mask` (S_tgt x S_tgt), 0 / -inf.
 *   pe_index (B) device int32 or NULL.  The reference adds pos_encoding[b] - indexed by BATCH position - to
 *   every token of clip b (models/positional_encoding.py:35) and therefore raises for B > 64; NULL reproduces
 *   exactly that (SDVG_ERR_BATCH for B > 64).  A caller that shards or chunks clips passes the row each clip
 *   would have had in the reference call (global_index mod 64).
 *   Padding masks (src_pad_mask / tgt_pad_mask) are None at every reference call site and are not supported. */
int sdvg_forward(sdvg_handle* h, const float* src, const float* tgt, int32_t B, int32_t S_src, int32_t S_tgt,
                 int32_t mask_kind, const float* mask, const int32_t* pe_index, float* out, void* stream);

/* Replaces predict() and the rollout hot loop (prediction/predict.py:16-42 and :143-197), for B clips at once.
 *   ctx (B, C, E) device fp32 context latents; out (B, n_pred, E) device fp32 predictions.
 *   Every step runs the model on src = tgt = window with the causal mask and keeps the last position
 *   (predict.py:24-26,42), then slides the window over [context, predictions].
 *   window   : tokens per step (the reference hard-codes 5, predict.py:196); clipped to the history length.
 *   flags    : bit 0 (SDVG_ROLLOUT_FAITHFUL) = the literal predict.py sequence instead of a plain sliding window:
 *              requires C == 5; first window is [SOS, f1..f5] with SOS = 2.0 (predict.py:124-130,
 *              utils/sd_utils.py:31,151-153), later windows are the last 5 of [f1..f4, p1..pk] (predict.py:193-196
 *              drops the last real frame).
 *              bit 1 (SDVG_ROLLOUT_RESIDUAL) = the predict_diff.py variant: every prediction is the model output
 *              plus the second-to-last frame of its window (prediction/predict_diff.py:33).
 *   teacher  : NULL, or (B, n_pred, E) device fp32 frames fed back instead of the model's own predictions
 *              (teacher forcing, used by the parity tests for reduced-precision modes).
 *   pe_index : (B) device int32 or NULL (NULL -> b mod 64, i.e. the reference run in chunks of 64 clips).
 *   scale_in / scale_out : multiply the context on ingest / the predictions on egress (0.18215 and 1/0.18215
 *              are the VAE latent scale of utils/sd_utils.py:143,159; 1.0 when the caller's latents are
 *              already scaled).  The fed-back frames are never rescaled. */
#define SDVG_ROLLOUT_FAITHFUL 1
#define SDVG_ROLLOUT_RESIDUAL 2
int sdvg_rollout(sdvg_handle* h, const float* ctx, int32_t B, int32_t C, int32_t n_pred, int32_t window,
                 int32_t flags, const float* teacher, const int32_t* pe_index, float scale_in, float scale_out,
                 float* out, void* stream);

/* Per-kernel-class device timing (CUDA events around every launch of the library's kernels).  Off by default.
 * sdvg_timing_read synchronises, accumulates and clears; classes: 0 GEMM (tensor core), 1 GEMM (SIMT),
 * 2 attention, 3 LayerNorm, 4 pack/export, 5 persistent small-batch kernel (one launch = a whole forward pass or
 * rollout when clips x tokens <= 128; bytes = weight-plane bytes it streams, flops = 2*M*N*K of its GEMM ops).
 * flops = algorithmic 2*M*N*K of the launches (0 for non-GEMM), bytes = algorithmic bytes moved by the launches. */
#define SDVG_NUM_KERNEL_CLASSES 6
int sdvg_timing_enable(sdvg_handle* h, int32_t on);
int sdvg_timing_read(sdvg_handle* h, double* ms, int64_t* launches, double* flops, double* bytes);

/* Kernels launched by this handle since creation (every launch of the library's own kernels is counted). */
int64_t sdvg_launch_count(const sdvg_handle* h);

/* Stand-alone GEMM entry point used by the kernel unit tests and micro-benchmarks:
 *   C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ ReLU), fp32 in / fp32 out on device, computed with the same
 *   kernels as the model (precision as in sdvg_precision; SDVG_MIXED is treated as SDVG_FP16).
 *   block_n: 0 = automatic tile choice, or one of 32/64/128/256 to force a tile width (tensor-core modes).
 *   iters > 1 repeats the GEMM launch back to back (operand packing excluded) and, when ms != NULL, returns
 *   the mean device time per launch in *ms (synchronises). */
int sdvg_gemm(int32_t device, int32_t precision, const float* A, const float* W, const float* bias, int32_t relu,
              float* C, int32_t M, int32_t N, int32_t K, int32_t block_n, int32_t iters, float* ms, void* stream);

/* Replaces the loss evaluation of the reference's validation / training loops - Trainer.criterion
 * (trainers/trainer.py:88-109) = use_mse*MSE + use_l1*L1 + use_gdl*lambda_gdl*GDL(alpha) (trainers/trainer.py:65-83)
 * + use_contrastive*lambda_contrastive*BiPatchNCE(temperature) (models/contrastive_loss.py:28-60) - forward values.
 *   pred, target: device fp32 (P, B, E) sequence-first, E = 4*h*w - the slices pred[-P:], y_expected[-P:] of
 *   trainers/trainer.py:145,224.  out: device fp32 [5] = total, MSE, L1, GDL, contrastive (all five are always
 *   computed except the contrastive term when use_contrastive == 0).  Like the reference, use_mse && use_l1 is
 *   rejected (SDVG_ERR_INVALID).  Asynchronous on `stream`; scratch is cached per device. */
int sdvg_criterion(int32_t device, const float* pred, const float* target, int32_t P, int32_t B, int32_t h, int32_t w,
                   int32_t use_mse, int32_t use_l1, int32_t use_gdl, float lambda_gdl, float alpha,
                   int32_t use_contrastive, float temperature, float lambda_contrastive, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training step - replaces the body of Trainer.train_loop (trainers/trainer.py:123-165): teacher-forced forward
 * model(new_batch, y_input, tgt_mask) (:141), criterion on the last frames_to_predict positions (:145),
 * opt.zero_grad(); loss.backward(); opt.step() (:160-162) with opt = Adam(model.parameters(), lr) (:365).
 * The handle must use a tensor-core precision (SDVG_FP32 for <=1e-4 gradient parity).  Dropout is off unless
 * sdvg_train_set_dropout is called. */
typedef struct sdvg_loss_config {   /* arguments of Trainer.criterion, trainers/trainer.py:88 */
  int32_t frames_to_predict;        /* P: loss over pred[-P:], y_expected[-P:] (:145) */
  int32_t use_mse, use_l1, use_gdl;
  float lambda_gdl, alpha;
  int32_t use_contrastive;
  float temperature, lambda_contrastive;
} sdvg_loss_config;

/* Forward with saved activations, criterion and backward pass.
 *   src (B, S_src, E) = new_batch, tgt (B, S_tgt, E) = y_input, device fp32, clip-major, contiguous (:123-128);
 *   expected (S_tgt, B, E) = y_expected after its permute(1, 0, 2) (:131-132), device fp32;
 *   the causal target mask of get_tgt_mask (:136-137) is implied.  losses: device fp32 [5] = total, MSE, L1, GDL,
 *   contrastive, or NULL.  Gradients of all parameters land unscaled in the flat fp32 vector of
 *   sdvg_train_gradients (layout = the parameter arena, see sdvg_param_range); every call overwrites them
 *   (opt.zero_grad() is implied).
 *   part 0: the whole step.  part 1: forward, criterion and the decoder-side backward - afterwards the gradients in
 *   [decoder_offset, count) are final, so a data-parallel caller can start their all-reduce while part 2 (target /
 *   source embedding and encoder backward, gradients in [0, decoder_offset)) runs. */
int sdvg_train_backward(sdvg_handle* h, const float* src, const float* tgt, const float* expected, int32_t B, int32_t S_src,
                        int32_t S_tgt, const sdvg_loss_config* loss, const int32_t* pe_index, float* losses, int32_t part,
                        void* stream);

/* The two halves of sdvg_train_backward around a loss that lives in the caller - the torch.autograd bridge that lets
 * the reference's loop body run unmodified: `pred = model(new_batch, y_input, tgt_mask)` (trainers/trainer.py:141),
 * `loss = loss_fn(pred[-P:], ...)` (:145, the reference's own torch criterion), `opt.zero_grad(); loss.backward();
 * opt.step()` (:163-165) with `optim.Adam(model.parameters())` (:365).
 *   sdvg_train_forward: Transformer.forward in train() mode - activations saved, dropout as set by
 *     sdvg_train_set_dropout, causal target mask; pred (S_tgt, B, E) device fp32 is written.  src / tgt must stay
 *     alive until the backward call (the embedding weight gradient reads them).
 *   sdvg_train_backward_from: dpred (S_tgt, B, E) device fp32 = dL/dpred from the caller's autograd graph; the
 *     parameter gradients land in the flat vector of sdvg_train_gradients exactly as with sdvg_train_backward. */
int sdvg_train_forward(sdvg_handle* h, const float* src, const float* tgt, int32_t B, int32_t S_src, int32_t S_tgt,
                       const int32_t* pe_index, float* pred, void* stream);
int sdvg_train_backward_from(sdvg_handle* h, const float* dpred, void* stream);

/* The flat gradient vector (device fp32, `count` elements) and the offset of the decoder-side bucket.  This is the
 * buffer a data-parallel trainer hands to ncclAllReduce (torch.distributed.all_reduce on a tensor view of it). */
int sdvg_train_gradients(sdvg_handle* h, float** grads, int64_t* count, int64_t* decoder_offset);

/* Training-mode dropout with probability p (DROPOUT_P of the reference's configs) at nn.Transformer's sites: the
 * embedding + positional sum (models/positional_encoding.py:35), the attention probabilities, every sub-layer output
 * before its residual add and the FFN hidden activation.  Masks are a counter-based hash of (seed, training step,
 * site, element index) - csrc/common.cuh drop_hash, restated in oracle/dropout.py - regenerated in the backward pass.
 * torch's own Philox stream is NOT reproduced: with p > 0 a run is statistically, not bitwise, comparable to the
 * reference; parity tests feed the same masks to the oracle.  p = 0 (default) disables. */
int sdvg_train_set_dropout(sdvg_handle* h, float p, uint64_t seed);

/* Finer-grained overlap: while sdvg_train_backward enqueues the backward pass it calls `fn(user, offset, count)` (on
 * the calling host thread) each time a range of the flat gradient vector is final given the work enqueued so far on
 * `stream` - from the end of the vector to its start, every `layers_per_bucket` layers; the ranges tile the vector.
 * The caller records an event on `stream` there and starts that range's all-reduce on its own stream.  NULL clears. */
typedef void (*sdvg_grad_ready_fn)(void* user, long long offset, long long count);
int sdvg_train_set_ready_callback(sdvg_handle* h, sdvg_grad_ready_fn fn, void* user, int32_t layers_per_bucket);

/* Offset / element count of a state_dict entry inside the flat parameter (and gradient) vector. */
int sdvg_param_range(const sdvg_handle* h, const char* key, int64_t* offset, int64_t* count);

/* Device pointer to the (S_tgt, B, E) prediction of the last training forward (what train_loop logs from, :164-172). */
int sdvg_train_prediction(sdvg_handle* h, const float** pred);

/* torch.optim.Adam.step() (betas, eps as given; no weight decay / amsgrad - the reference's defaults) on every
 * parameter, using grad_mul * gradient (1 / world_size after a sum all-reduce), then rebuilds the operand planes. */
int sdvg_train_adam_step(sdvg_handle* h, float lr, float beta1, float beta2, float eps, float grad_mul, void* stream);

/* The same update for the parameters [offset, offset + count) only - the ranges announced by the gradient-ready
 * callback - so the update (and plane rebuild) of finished layers overlaps the rest of the backward pass.  The first
 * range of a training step passes begin_step = 1 (advances the bias-correction step), the others 0; ranges must
 * arrive from the end of the vector to its start and tile it (the last one starts at offset 0). */
int sdvg_train_adam_step_range(sdvg_handle* h, float lr, float beta1, float beta2, float eps, float grad_mul, int64_t offset,
                               int64_t count, int32_t begin_step, void* stream);

/* Reads a parameter back (state_dict()[key], for torch.save at trainers/trainer.py:294): `out` is host or device
 * fp32 with room for the entry.  Synchronises `stream`. */
int sdvg_get_weight(sdvg_handle* h, const char* key, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDVG_H_ */
