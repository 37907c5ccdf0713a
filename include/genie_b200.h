/* libgenie_b200 — C ABI of the B200-native GENIE ST-transformer + MaskGIT path.
 *
 * The reference (1x-technologies/1xgpt) has no FFI layer: its seams for this path are nn.Module
 * methods.  Each entry point below states the reference interface it sits behind (file:line under
 * /root/reference).  INTEGRATION.md shows the ctypes stub a maintainer adds to bind them.
 *
 * Conventions
 *   - plain C types; every data pointer is a DEVICE pointer on the handle's device unless the name
 *     ends in `_host`; the caller owns all buffers; contiguous layouts as documented.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  All work is enqueued on it;
 *     entry points that return scalars to host memory synchronise that stream before returning.
 *   - return 0 on success, negative on error; message via gn_last_error() (thread-local).
 *   - token ids are int32 (the reference uses int64 tensors, disk format is uint32: data.py:46);
 *     the mask token is cfg.image_vocab_size (st_mask_git.py:51).
 *   - a handle is bound to one device, is not re-entrant, one handle per GPU / rank.
 */
#ifndef GENIE_B200_H
#define GENIE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GN_ABI_VERSION 4

/* precision modes of the linear layers / activations */
#define GN_PREC_BF16 0 /* tcgen05 kind::f16, bf16 operands+activations, fp32 accumulate/residual/LN/softmax */
#define GN_PREC_TF32 1 /* tcgen05 kind::tf32, fp32 activations (parity mode, <= 1e-3 rel) */
#define GN_PREC_FP32 2 /* CUDA-core fp32 FMA everywhere (exactness checks; slow) */
#define GN_PREC_FP16 3 /* tcgen05 kind::f16 with IEEE fp16 operands+activations (11-bit mantissa = tf32's), fp32
                          accumulate/residual/LN/softmax: the bf16 kernels at the bf16 speed, logits <= 1e-3 rel
                          (parity mode of the tensor-core path; conversions saturate at +-65504) */

#define GN_UNMASK_RANDOM 0 /* st_mask_git.py:204-206 (default of the reference) */
#define GN_UNMASK_GREEDY 1 /* st_mask_git.py:201-203 */

typedef struct gn_model gn_model;

/* mirrors genie/config.py:7-55 (GenieConfig) + runtime knobs */
typedef struct gn_config {
  int32_t num_layers;
  int32_t num_heads;
  int32_t d_model;
  int32_t T;
  int32_t S;
  int32_t image_vocab_size;
  int32_t num_factored_vocabs;
  int32_t factored_vocab_size;
  int32_t use_mup;
  int32_t qkv_bias;
  int32_t proj_bias;
  int32_t qk_norm;
  int32_t mlp_bias;
  float mlp_ratio;
  /* runtime */
  int32_t precision;     /* GN_PREC_* */
  int32_t chunk_tokens;  /* tokens per L2-resident work chunk (0 = library default) */
  int32_t kv_cache;      /* 1: temporal K/V cache + causal frame trimming in maskgit/generate/eval (exact same
                            results, fewer FLOPs); 0: recompute the full T-frame window every step like the
                            reference (st_mask_git.py:163,169) */
  int32_t generic_attention; /* 1: force the CUDA-core attention kernels (debug / cross-check) */
  int32_t fold_ln;       /* bf16 mode, pre-LN configs (qk_norm = 0): 1 = apply norm1 / norm2 inside the epilogue of the
                            QKV / fc1 GEMM (W*diag(gamma) folded into the weights, row statistics produced by the
                            previous residual epilogue) instead of a separate LayerNorm pass; 0 = separate pass */
  int32_t cuda_graphs;   /* 1: the per-chunk layer stack (~330 kernel launches) is captured once per shape into a CUDA
                            graph and replayed (only when the caller's stream is not the legacy default stream);
                            removes the host launch bound at small batch */
  int32_t lanes;         /* clips are independent: the chunks of a MaskGIT step are dealt round-robin to this many
                            concurrent streams (lane 0 = the caller's stream, the others library-owned, fork/join by
                            events), each with its own workspace, so that one lane's HBM-bound kernels and GEMM tail
                            waves overlap the other lane's tensor-bound kernels.  Results are bit-identical for any
                            value.  0 = library default, 1 = single stream, max 4 */
} gn_config;

int gn_version(void);
const char* gn_last_error(void);

int gn_model_create(gn_model** out, const gn_config* cfg, int device);
void gn_model_destroy(gn_model* m);

/* Copy + repack one tensor of the reference's state_dict (keys as saved by STMaskGIT.save_pretrained,
 * SURVEY.md section 8b; e.g. "decoder.layers.3.mlp.fc1.weight").  `src` is fp32 on the device. */
int gn_model_set_weight(gn_model* m, const char* key, const float* src, const int64_t* shape, int ndim, void* stream);
/* All tensors required by cfg present?  (0 / error naming the first missing key) */
int gn_model_check_weights(gn_model* m);

/* STTransformerDecoder.forward (genie/st_transformer.py:115-120): x [B,T,S,d] fp32 -> y [B,T,S,d] fp32 */
int gn_decoder_forward(gn_model* m, const float* x, float* y, int B, void* stream);

/* SelfAttention.forward (genie/attention.py:36-61 / :67-83): x [n_seq, n_tok, d] fp32 -> y, using the weights
 * of layer `layer`, attention `which` (0 spatial, 1 temporal) of the loaded model. */
int gn_attention_forward(gn_model* m, int layer, int which, const float* x, float* y, int n_seq, int n_tok,
                         int causal, void* stream);

/* STMaskGIT.compute_logits (genie/st_mask_git.py:255-265): ids [B,T,S] -> logits [B, NV*V, T, S] fp32 */
int gn_compute_logits(gn_model* m, const int32_t* ids, int B, float* logits, void* stream);

/* STMaskGIT.maskgit_generate (genie/st_mask_git.py:123-229).
 *   prompt [B,T,S] in/out (frame out_t is overwritten, frames >= out_t must be mask on entry: checked on
 *   the device, reported as an error after the call's work is enqueued and the stream synchronised);
 *   noise [steps-1, B, S] fp32 replaces torch.rand_like for GN_UNMASK_RANDOM (required when steps > 1);
 *   samples [B,S] out; logits0 [B, NV*V, S] fp32 out (step-0 logits of frame out_t; nullable).
 *   temperature <= 1e-8: argmax per factored vocab.  temperature > 1e-8: Categorical(probs / temperature) per
 *   factored vocab (st_mask_git.py:182-187; Categorical renormalises, so the draw is from softmax(logits) whatever
 *   the temperature), drawn by inverse CDF from `uniform` [steps, B, S, NV] fp32 in [0,1) (required then; it
 *   replaces the reference's global torch RNG the way `noise` replaces torch.rand_like). */
int gn_maskgit_generate(gn_model* m, int32_t* prompt, int B, int out_t, int steps, float temperature, int unmask_mode,
                        const float* noise, const float* uniform, int32_t* samples, float* logits0, void* stream);

/* genie/generate.py:77-103 / STMaskGIT.generate (st_mask_git.py:65-113): autoregressively fill frames
 * t_prompt..T-1 of tokens [B,T,S] (frames >= t_prompt are overwritten with mask first).
 *   noise [T-t_prompt, steps-1, B, S];  uniform [T-t_prompt, steps, B, S, NV] (temperature > 1e-8 only, else
 *   nullable);  logits0 [B, NV*V, T-t_prompt, S] nullable. */
int gn_generate(gn_model* m, int32_t* tokens, int B, int t_prompt, int steps, float temperature, int unmask_mode,
                const float* noise, const float* uniform, float* logits0, void* stream);
/* same, host buffers (pageable or pinned): H2D of tokens(+noise, +uniform), D2H of tokens inside the call */
int gn_generate_host(gn_model* m, int32_t* tokens_host, int B, int t_prompt, int steps, float temperature,
                     int unmask_mode, const float* noise_host, const float* uniform_host, void* stream);

/* genie/evaluate.py:82-122,173-179 + eval_utils.py:44-77: temporally teacher-forced evaluation of gt [B,T,S].
 * For t in 1..T-1: mask frames >= t, MaskGIT `steps` steps, CE of the step-0 logits against gt[:,t], token
 * accuracy of the final samples.  acc (device, 4 doubles) is ACCUMULATED into:
 *   acc[0] += sum of per-token CE, acc[1] += tokens, acc[2] += argmax-correct tokens (step-0 logits),
 *   acc[3] += sample == gt tokens.   noise [T-1, steps-1, B, S];  uniform [T-1, steps, B, S, NV] (temperature >
 *   1e-8 only, evaluate.py --temperature).  samples_out [B, T-1, S] nullable. */
int gn_teacher_forced_eval(gn_model* m, const int32_t* gt, int B, int steps, float temperature, int unmask_mode,
                           const float* noise, const float* uniform, int32_t* samples_out, double* acc, void* stream);

/* STMaskGIT.forward (st_mask_git.py:267-279): masked-mean factored CE + accuracy over frames 1..T-1.
 * input_ids/labels [B,T,S]; acc (device, 4 doubles) accumulated as above over positions where
 * input_ids == mask; logits [B,NV*V,T,S] nullable. */
int gn_forward_loss(gn_model* m, const int32_t* input_ids, const int32_t* labels, int B, float* logits, double* acc,
                    void* stream);

/* Test hook for the linear layer kernel: out[M,N] = epi(A[M,K] . W[N,K]^T + bias) (+ resid).
 * a/w dtype: in_bf16 0 fp32 / 1 bf16 / 2 fp16; out dtype: out_bf16 0 fp32 / 1 bf16 / 2 fp16 (16-bit in and out share
 * one format); epi 0 store / 1 erf-GELU / 2 residual; out2 (nullable) 16-bit copy for epi 2; force_simt selects the
 * CUDA-core kernel. */
int gn_linear_forward(const void* a, const void* w, const float* bias, const float* resid, void* out, void* out2,
                      int M, int N, int K, int epi, int in_bf16, int out_bf16, int force_simt, void* stream);

/* Test hook for the spatial attention kernels (the attention core of SelfAttention.forward, attention.py:48-58,
 * non-causal): qkv [n_frames*S, 3*d] bf16 (column order (3, h, hd)) -> out [n_frames*S, d] bf16, d = n_heads*head_dim.
 * kernel: 0 = library default (tcgen05 kernel when supported), 1 = mma.sync kernel, 2 = CUDA-core generic kernel;
 * + 0x100: the buffers hold IEEE fp16 instead of bf16. */
int gn_spatial_attention(const void* qkv, void* out, int n_frames, int S, int n_heads, int head_dim, float scale,
                         int kernel, void* stream);

/* decode-step kernels exposed for isolated bit-exact tests against the oracle */
/* uniform: nullptr = argmax; [R, NV] = inverse-CDF categorical draw (see gn_maskgit_generate) */
int gn_sample_tokens(const float* logits_rows, int R, int V, int NV, const float* uniform, int32_t* samples,
                     float* conf, void* stream);
int gn_remask_step(int32_t* prompt_frame, int64_t clip_stride, const int32_t* samples, const float* conf_or_noise,
                   uint8_t* unmasked, int32_t* samples_out, int B, int S, int n_mask, int last_step, int mask_id,
                   void* stream);
int gn_cross_entropy(const float* logits_rows, const int32_t* targets, int R, int V, int NV, const uint8_t* weight,
                     double* acc, void* stream);

/* Live per-launch timing (bench.py roofline leg): between begin and end every hot-path kernel launch is bracketed
 * by CUDA events on its stream.  out[2c] = total ms, out[2c+1] = launches of category c, for the
 * GN_PROF_CATEGORIES categories {gemm store, gemm gelu, gemm residual, prep/LN, spatial attn, temporal attn,
 * other}; out[2*GN_PROF_CATEGORIES] = sum of 2*M*N*K over the tcgen05 GEMM launches. */
#define GN_PROF_CATEGORIES 7
int gn_profile_begin(void);
int gn_profile_end(double* out /* 2*GN_PROF_CATEGORIES + 1 doubles */);

/* ------------------------------------------------------------------ MAGVIT2 tokenizer (second kernel family)
 * Encoder -> LFQ -> tokens and tokens -> LFQ^-1 -> Decoder.
 * reference: magvit2/modules/diffusionmodules/improved_model.py:54-182, lookup_free_quantize.py:181-194,241-257,
 * magvit2/models/lfqgan.py:121-129 (VQModel.encode/decode), visualize.py:84-116 (decode wrapper). */
typedef struct gn_vq gn_vq;
typedef struct gn_vq_config { /* mirrors magvit2/config.py:12-18 (VQConfig) */
  int32_t in_channels;    /* 3 */
  int32_t z_channels;     /* 18 */
  int32_t out_channels;   /* 3 */
  int32_t base_channels;  /* 128 */
  int32_t num_blocks;     /* len(ch_mult) = 5 */
  int32_t ch_mult[8];     /* (1, 1, 2, 2, 4) */
  int32_t num_res_blocks; /* 2 */
  int32_t precision;      /* operand format of the convolutions: GN_PREC_FP16 (default of the Python mirror), GN_PREC_BF16
                             (visualize.py:97 decodes in bf16), or GN_PREC_FP32: every convolution on the CUDA-core
                             fp32 kernel - the exact mode in which the LFQ token ids equal the reference's */
} gn_vq_config;
int gn_vq_create(gn_vq** out, const gn_vq_config* cfg, int device);
void gn_vq_destroy(gn_vq* m);
/* keys of VQModel.state_dict(): "encoder.conv_in.weight", "encoder.down.0.block.0.norm1.weight", ...,
 * "decoder.up.4.upsample.conv1.weight", ...; src is fp32 on the device in PyTorch layout (OIHW for convs). */
int gn_vq_set_weight(gn_vq* m, const char* key, const float* src, const int64_t* shape, int ndim, void* stream);
int gn_vq_check_weights(gn_vq* m, int need_encoder, int need_decoder);
/* img [B,3,H,W] fp32 in [-1,1] -> ids [B, (H/16)*(W/16)] (big-endian LFQ index, lookup_free_quantize.py:257);
 * z_out [B, z_channels, H/16, W/16] (pre-quantisation latents, nullable). */
int gn_vq_encode(gn_vq* m, const float* img, int B, int H, int W, int32_t* ids, float* z_out, void* stream);
/* ids [B, h*w] -> img_f32 [B,3,16h,16w] (decoder output, nullable) and/or img_u8 ((v+1)*127.5 clamped, nullable).
 * little_endian=1 reproduces visualize.py:111-116 (get_codebook_entry(...).flip(1): dataset tokens store bit c in
 * latent channel c); little_endian=0 inverts gn_vq_encode's own indices. */
int gn_vq_decode(gn_vq* m, const int32_t* ids, int B, int h, int w, int little_endian, float* img_f32, uint8_t* img_u8,
                 void* stream);

/* counters for bench accounting */
uint64_t gn_kernel_launches(void);              /* kernels launched by this library since load */
uint64_t gn_fallback_launches(void);            /* of those: launches that left the intended Blackwell kernel for a slower
                                                   variant because of a shape / alignment cliff (CUDA-core GEMM on a
                                                   tensor-core handle, mma.sync / generic attention on a 16-bit handle);
                                                   0 on the production shapes (GENIE_138M, d=1024 h=16), asserted in tests */
double gn_model_flops_per_clip_forward(gn_model* m); /* dense reference-equivalent FLOPs */
double gn_model_flops_executed(gn_model* m);    /* FLOPs actually issued by linear+attention kernels since reset */
double gn_model_bytes_executed(gn_model* m);    /* algorithmic HBM bytes of those launches since reset: every operand and
                                                   output of a launch once (GEMM A + W + out + residual, attention
                                                   q/k/v/o + temporal K/V cache rows, LayerNorm rows in + out) */
void gn_model_reset_counters(gn_model* m);

#ifdef __cplusplus
}
#endif
#endif /* GENIE_B200_H */
