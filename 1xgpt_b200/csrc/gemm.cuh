// Linear-layer GEMM  C[M,N] = epilogue(A[M,K] . W[N,K]^T)  for the ST-transformer
// (QKV / proj / fc1 / fc2 / readout: reference genie/attention.py:27,32,38,60,
//  genie/st_transformer.py:15-25, genie/st_mask_git.py:60-61,262).
//
// Two device paths, selected by SHAPE (not by backend):
//   * gemm_tcgen05_kernel : persistent, warp-specialised sm_100a kernel.  TMA (128B-swizzled K-major
//     tiles) -> smem ring -> tcgen05.mma (one issuing thread, fp32 accumulators double-buffered in
//     TMEM) -> epilogue warps (tcgen05.ld, bias / erf-GELU / residual in registers) -> swizzled smem
//     staging -> TMA store.
//   * gemm_simt_kernel    : plain CUDA-core fp32 tile kernel for shapes the tensor path does not take
//     (N % 64 != 0, K % 8 != 0, tiny test shapes) and for the fp32 "exact" precision mode.
#pragma once
#include "common.cuh"

namespace gn {

enum EpiKind : int { EPI_STORE = 0, EPI_GELU = 1, EPI_RESID = 2 };

// Implicit-GEMM 3x3 convolution (padding 1, stride 1 or 2) on the same kernel: A is the NHWC bf16 input
// [Nimg, Hi, Wi, Cin]; the K loop runs over (tap, 64-channel block) and the A tile of a k-block is the TMA box of
// the 128 output pixels of the tile shifted by the tap (out-of-bounds = zero padding).  W is [Cout, 9*Cin]
// (tap-major).  Output rows are flat NHWC output pixels.
struct ConvGeom {
  int Nimg, Hi, Wi, Cin;   // input
  int Ho, Wo;              // output spatial size
  int stride;              // 1 or 2
};

struct LinearArgs {
  const void* A;      // [M, lda]  bf16 or f32
  int64_t lda;
  const void* W;      // [N, ldw]  same dtype as A
  int64_t ldw;
  const float* bias;  // [N] or nullptr
  const float* resid; // [M, ldr] f32 (EPI_RESID) or nullptr
  int64_t ldr;
  void* out;          // [M, ldo]  bf16 or f32
  int64_t ldo;
  void* out2;         // optional second output, bf16 copy of `out` (only EPI_RESID with f32 out)
  int64_t ldo2;
  int M, N, K;
  int epi;            // EpiKind
  int in_bf16;        // 1: A/W are 16-bit (kind::f16), 0: f32 (kind::tf32 on the tensor path)
  int out_bf16;       // 1: out is 16-bit (same format as A/W), 0: f32
  int fp16;           // 16-bit tensors are IEEE fp16 instead of bf16 (the parity format, see common.cuh H16)
  int force_simt;     // 1: CUDA-core fp32 path regardless of shape
  int round_out_tf32; // 1: round the fp32 output to tf32 (it feeds a kind::tf32 GEMM next)
  // qk-LayerNorm (attention.py:42-47: LayerNorm(head_dim) with one shared affine on q and k) applied in the epilogue of
  // the QKV projection: output columns [0, qkn_cols) are normalised per group of 64 (= one head) in fp32 before the
  // bf16 rounding; bf16 tensor path with the store epilogue only (head_dim must be 64).  nullptr: off.
  const float* qkn_gamma;
  const float* qkn_beta;
  int qkn_cols;
  int qkn_hd;         // head_dim of the normalised groups: 64 (a pair of 32-column chunks) or 32 (one chunk)
  int a_evict_first;  // 1: A is not read again after this GEMM (L2 evict_first hint on its loads)
  int red_add;        // 1 (EPI_STORE, fp32 out, tensor path): out += A.W^T + bias through TMA reduce-add (the residual
                      // update x += f(x) without loading x into the SM; bit-identical to EPI_RESID in place)
  const ConvGeom* conv; // non-null: implicit-GEMM 3x3 convolution (A = NHWC input, K = 9*Cin, M = Nimg*Ho*Wo)
  // LayerNorm folded into the epilogue (A holds the RAW bf16 rows, W holds W*diag(gamma)):
  //   out[m,n] = rstd[m] * (acc[m,n] - mean[m] * colsum[n]) + bias[n]        (bias already contains beta . W^T)
  // row statistics come as `ln_np` partial {sum, sumsq} pairs per row (written by the previous residual epilogue)
  const float* ln_stats;  // [M, ln_np, 2] or nullptr
  int ln_np;
  int ln_d;               // normalised width (d_model)
  const float* ln_colsum; // [N]
  // residual epilogue: also emit per-row partial statistics of the new residual stream: [M, N/BLOCK_N, 2]
  float* stats_out;
  // temporal QKV projection (EPI_STORE, bf16 out, N = 3*kv_d): output columns [kv_d, 2*kv_d) / [2*kv_d, 3*kv_d) are
  // written straight into the temporal K / V caches (layout [clip*S + s][head][T][hd]) instead of `out`;
  // GEMM row (b, tl, s) = (b*kv_Tact + tl)*kv_S + s goes to frame kv_t0 + tl.  kv_k == nullptr: off.
  void* kv_k;
  void* kv_v;
  int kv_d, kv_hd, kv_T, kv_S, kv_Tact, kv_t0, kv_clips;
};

// Enqueue on `stream`.  Returns GN_OK or a negative code (message via last_error()).
int linear_forward(const LinearArgs& a, cudaStream_t stream);

// Live per-launch timing of the tcgen05 GEMM kernel (bench.py's roofline leg): while enabled, every tensor-path
// launch is bracketed by CUDA events on its own stream.  profile_end() synchronises those events and returns
// {sum of durations [ms], sum of 2*M*N*K, launches} for the launches since profile_begin().
int profile_begin();
int profile_end(double* out);   // see runtime.cu: per-category {ms, launches}
extern double g_gemm_flops_issued;   // 2*M*N*K summed over tensor-path launches (reset by the caller)

// BLOCK_N the tensor path uses for a residual-epilogue GEMM (= number of stats_out partials per row is N / this)
int resid_block_n(int N, int K, bool dual);

// number of kernels launched by linear_forward so far (bench's gpu_launches accounting)
extern unsigned long long g_launch_count;
// launches that fell off the intended Blackwell kernel onto a slower variant (see gemm.cu)
extern unsigned long long g_fallback_launches;

}  // namespace gn
