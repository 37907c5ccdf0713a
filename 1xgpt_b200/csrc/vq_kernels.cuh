// Launchers of the MAGVIT2 tokenizer kernels (vq_kernels.cu).
#pragma once
#include "common.cuh"

namespace gn {
int launch_stem_conv(const float* img, const float* w, float* out, int B, int H, int W, int Cout, cudaStream_t st);
int launch_gn_swish(const float* x, double* stats, const float* gamma, const float* beta, bf16* out, int B, int HW, int C,
                    cudaStream_t st);
int launch_depth_to_space(const float* in, float* out, int B, int H, int W, int Cp, cudaStream_t st);
int launch_vq_head(const float* x, double* stats, const float* gamma, const float* beta, const float* w, const float* bias,
                   int32_t* ids, float* z_out, int B, int HW, int C, int Z, cudaStream_t st);
int launch_vq_tail(const int32_t* ids, const float* w, const float* bias, float* out, int B, int H, int W, int Z, int Cout,
                   int little_endian, cudaStream_t st);
int launch_out_conv(const bf16* a, const float* w, const float* bias, float* out_f32, uint8_t* out_u8, int B, int H, int W,
                    int C, cudaStream_t st);
int launch_repack_conv_w(const float* w, bf16* out, int Cout, int Cin, int taps, cudaStream_t st);
}  // namespace gn
