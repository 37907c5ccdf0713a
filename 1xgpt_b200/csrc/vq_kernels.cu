// MAGVIT2 tokenizer kernels that are NOT the 3x3 implicit-GEMM convolution (that one is the tcgen05 GEMM kernel
// in conv mode, gemm.cu): GroupNorm(32)+swish, the 3->C stem, the 18-channel LFQ head / tail, depth-to-space.
// Activations are NHWC: fp32 trunk, bf16 GEMM operands.
// reference: magvit2/modules/diffusionmodules/improved_model.py, magvit2/modules/vqvae/lookup_free_quantize.py
#include "kernels.cuh"
#include "vq_kernels.cuh"

namespace gn {

// -------------------------------------------------------------------------------------
// stem: conv 3x3 (Cin = 3, padding 1, no bias) from the NCHW fp32 image to the NHWC fp32 trunk
// improved_model.py:67-73 (Encoder.conv_in).  w: [Cout, 3, 3, 3] (PyTorch OIHW).
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w, float* __restrict__ out, int H, int W,
                 int Cout) {
  extern __shared__ float sw[];  // [27][Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;  // k = ci*9 + ky*3 + kx
    sw[i] = w[(int64_t)co * 27 + k];
  }
  __syncthreads();
  const int n = blockIdx.y;
  const int groups = Cout / 32;                      // 32 output channels per thread
  const int pix = blockIdx.x * (blockDim.x / groups) + threadIdx.x / groups;
  const int cg = threadIdx.x % groups;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix % W;
  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        in[ci * 9 + ky * 3 + kx] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(((int64_t)n * 3 + ci) * H + yy) * W + xx] : 0.f;
      }
  float* o = out + ((int64_t)n * H * W + pix) * Cout + cg * 32;
#pragma unroll 4
  for (int c = 0; c < 32; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) acc = fmaf(in[k], sw[k * Cout + cg * 32 + c], acc);
    o[c] = acc;
  }
}

int launch_stem_conv(const float* img, const float* w, float* out, int B, int H, int W, int Cout, cudaStream_t st) {
  GN_REQUIRE(Cout % 32 == 0 && Cout <= 256, "stem conv: Cout %d unsupported", Cout);
  const int groups = Cout / 32, ppb = 256 / groups;
  dim3 grid(ceil_div(H * W, ppb), B);
  stem_conv_kernel<<<grid, 256, 27 * Cout * sizeof(float), st>>>(img, w, out, H, W, Cout);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// GroupNorm(32 groups, eps 1e-6) statistics over NHWC fp32: stats[n][g] = {sum, sumsq} (double)
// improved_model.py:25-26,116 (nn.GroupNorm(32, C, eps=1e-6))
// -------------------------------------------------------------------------------------
// Deterministic two-stage reduction (no atomics: LFQ takes the SIGN of the encoder output, so run-to-run
// summation-order noise would flip bits of near-zero latents):
//   stage 1: block (chunk of pixels, image) -> partial[n][chunk][g] = {sum, sumsq} in a fixed order
//   stage 2: stats[n][g] = sum over chunks in index order (double)
__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ x, double* __restrict__ partial, int HW, int C, int pix_per_block) {
  __shared__ float s_sum[256], s_sq[256];
  const int n = blockIdx.y;
  const int cg = C / 32;                 // channels per group (4, 8, 16)
  const int vpp = C / 4;                 // float4 per pixel; 256 % vpp == 0 -> each thread owns one channel quad
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, HW);
  const float4* base = reinterpret_cast<const float4*>(x + (int64_t)n * HW * C) + (int64_t)p0 * vpp;
  const int total = (p1 - p0) * vpp;
  float sum = 0.f, sq = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const float4 f = base[i];
    sum += (f.x + f.y) + (f.z + f.w);
    sq += (f.x * f.x + f.y * f.y) + (f.z * f.z + f.w * f.w);
  }
  s_sum[threadIdx.x] = sum;
  s_sq[threadIdx.x] = sq;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    const int q0 = g * cg / 4, q1 = (g + 1) * cg / 4;   // channel quads of this group
    double a = 0.0, b2 = 0.0;
    for (int t0 = 0; t0 < 256; t0 += vpp)
      for (int q = q0; q < q1; ++q) { a += (double)s_sum[t0 + q]; b2 += (double)s_sq[t0 + q]; }
    double* o = partial + (((int64_t)n * gridDim.x + blockIdx.x) * 32 + g) * 2;
    o[0] = a;
    o[1] = b2;
  }
}
__global__ void gn_finalize_kernel(const double* __restrict__ partial, double* __restrict__ stats, int chunks) {
  const int n = blockIdx.x, g = threadIdx.x;   // 32 threads
  double a = 0.0, b = 0.0;
  for (int c = 0; c < chunks; ++c) {
    const double* p = partial + (((int64_t)n * chunks + c) * 32 + g) * 2;
    a += p[0];
    b += p[1];
  }
  stats[((int64_t)n * 32 + g) * 2] = a;
  stats[((int64_t)n * 32 + g) * 2 + 1] = b;
}
// stats buffer layout: [B*64 doubles final stats][partials]
static int gn_stats(const float* x, double* stats, int B, int HW, int C, cudaStream_t st) {
  GN_REQUIRE(C % 128 == 0 && 1024 % C == 0, "GroupNorm: C %d unsupported (128, 256, 512, 1024)", C);
  const int ppb = HW >= 4096 ? 128 : (HW >= 1024 ? 32 : 8);
  const int chunks = ceil_div(HW, ppb);
  GN_REQUIRE(chunks <= 512, "GroupNorm: too many partial chunks");
  double* partial = stats + (int64_t)B * 64;
  gn_partial_kernel<<<dim3(chunks, B), 256, 0, st>>>(x, partial, HW, C, ppb);
  GN_CUDA_CHECK(cudaGetLastError());
  gn_finalize_kernel<<<B, 32, 0, st>>>(partial, stats, chunks);
  GN_CUDA_CHECK(cudaGetLastError());
  g_launch_count += 2;
  return GN_OK;
}

// y = swish(GN(x)) -> bf16 NHWC (the A operand of the next convolution)     improved_model.py:8-10,41-46
__global__ void __launch_bounds__(256)
gn_apply_swish_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                      const float* __restrict__ beta, bf16* __restrict__ out, int HW, int C, int64_t total_vec) {
  const int vec_per_pix = C / 4, cg = C / 32;
  const double cnt = (double)HW * cg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_pix);
    const int64_t pix = i / vec_per_pix;
    const int n = (int)(pix / HW);
    const int g = (v * 4) / cg;
    const double s = stats[((int64_t)n * 32 + g) * 2], q = stats[((int64_t)n * 32 + g) * 2 + 1];
    const double mean_d = s / cnt;
    const double var_d = fmax(q / cnt - mean_d * mean_d, 0.0);
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var_d + 1e-6));
    const float4 f = reinterpret_cast<const float4*>(x)[i];
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + v);
    float y[4] = {(f.x - mean) * rstd * gm.x + bt.x, (f.y - mean) * rstd * gm.y + bt.y, (f.z - mean) * rstd * gm.z + bt.z,
                  (f.w - mean) * rstd * gm.w + bt.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = y[k] / (1.f + __expf(-y[k]));
    uint2 p;
    p.x = pack_bf16x2(y[0], y[1]);
    p.y = pack_bf16x2(y[2], y[3]);
    reinterpret_cast<uint2*>(out)[i] = p;
  }
}

int launch_gn_swish(const float* x, double* stats, const float* gamma, const float* beta, bf16* out, int B, int HW, int C,
                    cudaStream_t st) {
  GN_PROPAGATE(gn_stats(x, stats, B, HW, C, st));
  const int64_t total_vec = (int64_t)B * HW * (C / 4);
  const int g2 = (int)std::min<int64_t>(ceil_div64(total_vec, 256), 148 * 16);
  gn_apply_swish_kernel<<<g2, 256, 0, st>>>(x, stats, gamma, beta, out, HW, C, total_vec);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// depth-to-space (DCR, block 2): [B,H,W,4C'] -> [B,2H,2W,C'], channel = (b1*2 + b2)*C' + c'
// improved_model.py:185-237
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
depth_to_space_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Cp, int64_t total_vec) {
  const int vec_c = Cp / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_c);
    int64_t r = i / vec_c;
    const int ox = (int)(r % (2 * W)); r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    const int b1 = oy & 1, b2 = ox & 1, h = oy >> 1, w = ox >> 1;
    const float4 f = reinterpret_cast<const float4*>(in)[(((int64_t)n * H + h) * W + w) * (4 * vec_c) +
                                                          (b1 * 2 + b2) * vec_c + v];
    reinterpret_cast<float4*>(out)[i] = f;
  }
}
int launch_depth_to_space(const float* in, float* out, int B, int H, int W, int Cp, cudaStream_t st) {
  const int64_t total_vec = (int64_t)B * 4 * H * W * (Cp / 4);
  const int grid = (int)std::min<int64_t>(ceil_div64(total_vec, 256), 148 * 16);
  depth_to_space_kernel<<<grid, 256, 0, st>>>(in, out, H, W, Cp, total_vec);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// encoder head: z = conv1x1(swish(GN(x))) + b  (C -> Z channels), LFQ: q = sign(z) (z > 0 -> +1 else -1),
// index = sum_c (z_c > 0) << (Z-1-c)   (big endian, lookup_free_quantize.py:152,248,257).  One warp per pixel.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_head_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
               const float* __restrict__ beta, const float* __restrict__ w /*[Z][C]*/, const float* __restrict__ bias,
               int32_t* __restrict__ ids, float* __restrict__ z_out /*[B,Z,HW] nullable*/, int HW, int C, int Z,
               int n_pix) {
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= n_pix) return;
  const int lane = threadIdx.x & 31;
  const int n = pix / HW;
  const int cg = C / 32;
  const double cnt = (double)HW * cg;
  float acc[32];
#pragma unroll
  for (int z = 0; z < 32; ++z) acc[z] = 0.f;
  for (int c = lane; c < C; c += 32) {
    const int g = c / cg;
    const double s = stats[((int64_t)n * 32 + g) * 2], q = stats[((int64_t)n * 32 + g) * 2 + 1];
    const double mean_d = s / cnt;
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(fmax(q / cnt - mean_d * mean_d, 0.0) + 1e-6));
    float y = (x[(int64_t)pix * C + c] - mean) * rstd * gamma[c] + beta[c];
    y = y / (1.f + __expf(-y));
    for (int z = 0; z < Z; ++z) acc[z] = fmaf(y, w[(int64_t)z * C + c], acc[z]);
  }
  int idx = 0;
  for (int z = 0; z < Z; ++z) {
    const float v = warp_sum(acc[z]) + bias[z];
    if (v > 0.f) idx |= 1 << (Z - 1 - z);
    if (z_out != nullptr && lane == 0) z_out[((int64_t)n * Z + z) * HW + (pix % HW)] = v;
  }
  if (lane == 0) ids[pix] = idx;
}
int launch_vq_head(const float* x, double* stats, const float* gamma, const float* beta, const float* w, const float* bias,
                   int32_t* ids, float* z_out, int B, int HW, int C, int Z, cudaStream_t st) {
  GN_REQUIRE(Z <= 31, "LFQ: at most 31 bits");
  GN_PROPAGATE(gn_stats(x, stats, B, HW, C, st));
  const int n_pix = B * HW;
  vq_head_kernel<<<ceil_div(n_pix, 8), 256, 0, st>>>(x, stats, gamma, beta, w, bias, ids, z_out, HW, C, Z, n_pix);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// decoder stem: ids -> +-1 latents (Z channels; little_endian: bit c <-> channel c, i.e. get_codebook_entry
// followed by visualize.py:115 `.flip(1)`; big endian: bit (Z-1-c) <-> channel c) -> conv 3x3 (Z -> Cout) + bias
// lookup_free_quantize.py:181-194, improved_model.py:135-137,164.   w: [Cout, Z, 3, 3].  One warp per pixel.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_tail_kernel(const int32_t* __restrict__ ids, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ out, int H, int W, int Z, int Cout, int little_endian, int n_pix) {
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= n_pix) return;
  const int lane = threadIdx.x & 31;
  const int n = pix / (H * W), rem = pix % (H * W), y = rem / W, x = rem % W;
  int nb[9];
  bool ok[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    ok[t] = yy >= 0 && yy < H && xx >= 0 && xx < W;
    nb[t] = ok[t] ? ids[((int64_t)n * H + yy) * W + xx] : 0;
  }
  for (int co = lane; co < Cout; co += 32) {
    float acc = bias[co];
    const float* wc = w + (int64_t)co * Z * 9;
    for (int c = 0; c < Z; ++c) {
      const int bit = little_endian ? c : (Z - 1 - c);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (ok[t]) acc += ((nb[t] >> bit) & 1) ? wc[c * 9 + t] : -wc[c * 9 + t];
      }
    }
    out[(int64_t)pix * Cout + co] = acc;
  }
}
int launch_vq_tail(const int32_t* ids, const float* w, const float* bias, float* out, int B, int H, int W, int Z, int Cout,
                   int little_endian, cudaStream_t st) {
  const int n_pix = B * H * W;
  vq_tail_kernel<<<ceil_div(n_pix, 8), 256, 0, st>>>(ids, w, bias, out, H, W, Z, Cout, little_endian, n_pix);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// decoder output conv: 3x3, C -> 3, + bias, input = swish(GN(x)) as bf16 NHWC; writes fp32 NCHW and/or the uint8
// image ((v + 1) * 127.5 clamped to [0, 255], truncated: visualize.py:84-92).  One warp per output pixel.
// w: [3, C, 3, 3]
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
out_conv_kernel(const bf16* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ out_f32, uint8_t* __restrict__ out_u8, int H, int W, int C, int n_pix) {
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= n_pix) return;
  const int lane = threadIdx.x & 31;
  const int n = pix / (H * W), rem = pix % (H * W), y = rem / W, x = rem % W;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const bf16* ap = a + (((int64_t)n * H + yy) * W + xx) * C;
    for (int c = lane * 2; c < C; c += 64) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ap + c));
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        acc[co] = fmaf(v.x, w[((int64_t)co * C + c) * 9 + t], acc[co]);
        acc[co] = fmaf(v.y, w[((int64_t)co * C + c + 1) * 9 + t], acc[co]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 3; ++co) {
    const float v = warp_sum(acc[co]) + bias[co];
    if (lane == 0) {
      const int64_t o = (((int64_t)n * 3 + co) * H + y) * W + x;
      if (out_f32) out_f32[o] = v;
      if (out_u8) out_u8[o] = (uint8_t)fminf(fmaxf((v + 1.f) * 127.5f, 0.f), 255.f);
    }
  }
}
int launch_out_conv(const bf16* a, const float* w, const float* bias, float* out_f32, uint8_t* out_u8, int B, int H, int W,
                    int C, cudaStream_t st) {
  const int n_pix = B * H * W;
  out_conv_kernel<<<ceil_div(n_pix, 8), 256, 0, st>>>(a, w, bias, out_f32, out_u8, H, W, C, n_pix);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// conv weight repack: PyTorch [Cout, Cin, kh, kw] fp32 -> [Cout, kh*kw, Cin] bf16 (tap-major K for the implicit GEMM)
__global__ void repack_conv_w_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int taps) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int t = (int)((i / Cin) % taps);
    const int co = (int)(i / ((int64_t)Cin * taps));
    out[i] = __float2bfloat16_rn(w[((int64_t)co * Cin + ci) * taps + t]);
  }
}
int launch_repack_conv_w(const float* w, bf16* out, int Cout, int Cin, int taps, cudaStream_t st) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), 4096);
  repack_conv_w_kernel<<<grid, 256, 0, st>>>(w, out, Cout, Cin, taps);
  GN_CUDA_CHECK(cudaGetLastError());
  return GN_OK;
}

}  // namespace gn
