// Launchers of the MAGVIT2 tokenizer kernels (vq_kernels.cu).
#pragma once
#include "common.cuh"

namespace gn {
int launch_stem_conv(const float* img, const float* w, float* out, int B, int H, int W, int Cout, cudaStream_t st);
// o16 (here and below): operand format of the convolution inputs: 1 = bf16, 2 = fp16, 0 = fp32 (exact mode)
int launch_gn_swish(const float* x, double* stats, const float* gamma, const float* beta, void* out, int o16, int B,
                    int HW, int C, cudaStream_t st);
// single persistent kernel (statistics + apply, second read from L2) for the 16-bit operand modes; `ready` = 128 zero-
// initialised flags, `epoch` = a value that differs from every earlier launch on this buffer; falls back to the two-kernel
// form for o16 == 0 and unsupported shapes.  The stats buffer must hold kGnFusedMaxChunks (512) chunks per image.
int launch_gn_swish_fused(const float* x, double* stats, unsigned int* ready, unsigned int epoch, const float* gamma,
                          const float* beta, void* out, int o16, int B, int HW, int C, cudaStream_t st);
int launch_depth_to_space(const float* in, float* out, int B, int H, int W, int Cp, cudaStream_t st);
int launch_vq_head(const float* x, double* stats, const float* gamma, const float* beta, const float* w, const float* bias,
                   int32_t* ids, float* z_out, int B, int HW, int C, int Z, cudaStream_t st);
int launch_vq_tail(const int32_t* ids, const float* w, const float* bias, float* out, int B, int H, int W, int Z, int Cout,
                   int little_endian, cudaStream_t st);
int launch_out_conv(const void* a, int o16, const float* w, const float* bias, float* out_f32, uint8_t* out_u8, int B, int H,
                    int W, int C, cudaStream_t st);
// tensor-core output conv (16-bit operand modes): `wf` = fragment-ordered weights, 9 * C/16 * 32 uint2, built by
// launch_out_conv_pack whenever decoder.conv_out.weight changes
int launch_out_conv_pack(const float* w, void* wf, int o16, int C, cudaStream_t st);
int launch_out_conv_mma(const void* a, int o16, const void* wf, const float* bias, float* out_f32, uint8_t* out_u8, int B,
                        int H, int W, int C, cudaStream_t st);
int launch_repack_conv_w(const float* w, void* out, int o16, int Cout, int Cin, int taps, cudaStream_t st);
}  // namespace gn
