// Launcher declarations for every kernel of the hot path (all enqueue on the given stream and
// return GN_OK / negative code).
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace gn {

// ---- elementwise.cu
int launch_embed(const int32_t* ids, const float* E, const float* mask_embed, const float* pos, float* x, int B,
                 int T, int S, int t0, int Tact, int d, int V, int NV, int mask_id, cudaStream_t st);
// out_bf16: 0 = fp32 output, 1 = bf16, 2 = fp16
int launch_prep(const float* x, void* out, int out_bf16, const float* gamma, const float* beta, int n_rows, int d,
                float scale, int S, int Tact, int tsel, cudaStream_t st, int round_tf32 = 0);
// prep variant that also writes the row statistics {sum, sumsq} (1 partial per row) next to the bf16 cast
int launch_prep_stats(const float* x, bf16* out, float* stats, int n_rows, int d, cudaStream_t st);
// fold LayerNorm(gamma, beta) into a linear layer: Wf = bf16(W * diag(gamma)), colsum[n] = sum_k float(Wf[n,k]),
// bias_f[n] = sum_k beta[k] * W[n,k] + bias[n]
int launch_fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, bf16* Wf, float* colsum,
                   float* bias_f, int N, int K, cudaStream_t st);
int launch_cast_h16(const float* in, void* out, int fp16, int64_t n, cudaStream_t st);   // fp32 -> bf16 / fp16
int launch_round_tf32(const float* in, float* out, int64_t n, cudaStream_t st);
int launch_logits_transpose(const float* rows, float* out, int B, int Tl, int S, int C, int Tout, int tslot0,
                            cudaStream_t st);

// ---- attention.cu
struct AttnArgs {
  const void* qkv;        // [n_seq * n_q_tok(+..), 3*d] fused projection output, column order (3, h, hd)
  void* out;              // [rows, d]
  int act_bf16;           // activation dtype of qkv/out: 1 = 16-bit, 0 = f32
  int fp16;               // 16-bit activations are IEEE fp16 instead of bf16
  int n_heads, head_dim;
  float scale;
  const float* qk_gamma;  // [hd] shared q/k LayerNorm affine (attention.py:34,43-44) or nullptr
  const float* qk_beta;
  int round_tf32;         // fp32 activations only: round the output to tf32 (feeds a kind::tf32 GEMM)
};
// spatial: sequences = frames; tokens of a sequence are S consecutive rows.  non-causal.
int launch_spatial_attention(const AttnArgs& a, int n_frames, int S, int force_generic, cudaStream_t st);
// mma.sync spatial kernel (attention_fast.cu): bf16, head_dim 64 / 32, S in {128, 256}, optional qk-LayerNorm
bool fast_spatial_supported(const AttnArgs& a, int S);
int fast_spatial_attention(const AttnArgs& a, int n_frames, int S, cudaStream_t st);
// tcgen05 spatial attention (attention_tc.cu): bf16, head_dim 64, S in {128, 256}, no qk-LayerNorm
bool tc_spatial_supported(const AttnArgs& a, int S);
int tc_spatial_attention(const AttnArgs& a, int n_frames, int S, cudaStream_t st);
// temporal: sequences = (clip, spatial position); token (b, tl, s) is row (b*Tq + tl)*S + s of qkv (fresh
// frames t0..t0+Tq-1).  Keys/values of frames < t0 come from kcache/vcache [B, S, T, d] (nullptr when t0 == 0);
// fresh k/v are written back to the caches when they are non-null.  causal.
int launch_temporal_attention(const AttnArgs& a, int B, int S, int T, int t0, int Tq, void* kcache, void* vcache,
                              int force_generic, cudaStream_t st);
// temporal v2 (bf16, head_dim 64, no qk-LayerNorm): K/V of frames [0, t0+Tq) are read from head-major caches
// [clip*S + s][head][T][64] which the temporal QKV GEMM has already filled (LinearArgs::kv_*); `a.qkv` supplies Q only.
bool temporal_v2_supported(const AttnArgs& a, int S, int T);
int launch_temporal_attention_v2(const AttnArgs& a, int nb, int S, int T, int t0, int Tq, const void* kcache,
                                 const void* vcache, cudaStream_t st);
// generic standalone attention over [n_seq, n_tok] (SelfAttention.forward contract, any small shape)
int launch_generic_attention(const AttnArgs& a, int n_seq, int n_tok, int causal, cudaStream_t st);

// ---- readout_sample.cu: readout GEMM fused with the factored softmax / argmax / confidence (temperature 0).
// A [R, K] 16-bit rows of the decoded frame, W [NV*512, K], bias [NV*512] fp32 -> samples [R], conf [R]; logits
// [R, NV*512] fp32 only when `logits` != nullptr.
bool readout_sample_supported(int V, int K, int is_16bit);
int launch_readout_sample(const void* A, const void* W, const float* bias, float* logits, int32_t* samples, float* conf,
                          int R, int K, int NV, int fp16, cudaStream_t st);

// ---- decode.cu
// factored softmax / argmax-or-categorical / confidence of one frame's logits rows [R, NV*V]  (st_mask_git.py:171-190)
// uniform: nullptr = greedy argmax; else [R, NV] uniforms in [0,1) for the inverse-CDF categorical draw
int launch_sample(const float* logits, int R, int V, int NV, const float* uniform, int32_t* samples, float* conf,
                  cudaStream_t st);
// cosine re-mask + scatter for one MaskGIT step (st_mask_git.py:192-223), one CTA per clip
int launch_remask(int32_t* prompt_frame, int64_t clip_stride, const int32_t* samples, const float* conf_or_noise,
                  uint8_t* unmasked, int32_t* samples_out, int B, int S, int n_mask, int last_step, int mask_id,
                  cudaStream_t st);
// factored cross-entropy + argmax accuracy over logits rows (eval_utils.py:72-77, st_mask_git.py:236-250)
//   targets: int32 per row (unfactorized id); weight: optional uint8 per row (relevant mask)
//   acc[0] += sum loss, acc[1] += rows counted, acc[2] += rows whose per-vocab argmax all match
int launch_ce(const float* logits, const int32_t* targets, int64_t target_stride_b, int rows_per_b, int R, int V, int NV,
              const uint8_t* weight, double* acc, cudaStream_t st);
// acc[3] += #(a == b)
int launch_count_equal(const int32_t* a, int64_t a_stride_b, const int32_t* b, int64_t b_stride_b, int rows_per_b, int R,
                       double* acc, cudaStream_t st);
// all(prompt[:, t_from:] == mask_id) -> flag (0 = ok); device-side replacement of st_mask_git.py:155
int launch_check_masked(const int32_t* prompt, int B, int T, int S, int t_from, int mask_id, int* flag, cudaStream_t st);

}  // namespace gn
