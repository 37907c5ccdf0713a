// MAGVIT2 tokenizer kernels that are NOT the 3x3 implicit-GEMM convolution (that one is the tcgen05 GEMM kernel
// in conv mode, gemm.cu): GroupNorm(32)+swish, the 3->C stem, the 18-channel LFQ head / tail, depth-to-space.
// Activations are NHWC: fp32 trunk, bf16 GEMM operands.
// reference: magvit2/modules/diffusionmodules/improved_model.py, magvit2/modules/vqvae/lookup_free_quantize.py
#include "kernels.cuh"
#include "vq_kernels.cuh"
#include <cstdlib>

namespace gn {

// -------------------------------------------------------------------------------------
// stem: conv 3x3 (Cin = 3, padding 1, no bias) from the NCHW fp32 image to the NHWC fp32 trunk
// improved_model.py:67-73 (Encoder.conv_in).  w: [Cout, 3, 3, 3] (PyTorch OIHW).
// -------------------------------------------------------------------------------------
// One block = ROWS consecutive image rows (round 2b; the first TMA-free version took one row per block, so the 108 weight
// loads per lane, the staging of 3 input rows and the barrier were paid for every 32 pixels a warp produces: 214 us per 8
// images against ~45 us of HBM time for the 268 MB trunk it writes).  The ROWS + 2 input rows (3 channels, zero padded)
// are staged in shared memory once; a warp takes (row, 32-pixel segment) items and walks the segment with the 3x3x3
// window sliding through registers (9 broadcast LDS per pixel); each lane owns Cout/32 output channels with their 27 taps
// in registers, so a pixel's Cout floats leave the warp as ONE contiguous store (512 B for Cout = 128).  The tap order of
// the per-pixel sum is unchanged (k = ci*9 + ky*3 + kx), results are bit-identical for every ROWS.
template <int CPL, int ROWS>   // output channels per lane, image rows per block
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w, float* __restrict__ out, int H, int W) {
  constexpr int Cout = CPL * 32;
  constexpr int RS = ROWS + 2;           // staged input rows per channel
  extern __shared__ float srow[];        // [3 ci][ROWS + 2][W + 2]
  const int n = blockIdx.y, y0 = blockIdx.x * ROWS;
  const int Wp = W + 2;
  for (int i = threadIdx.x; i < 3 * RS * Wp; i += blockDim.x) {
    const int r = i / Wp, xx = i % Wp - 1;
    const int ci = r / RS, yy = y0 + r % RS - 1;
    srow[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(((int64_t)n * 3 + ci) * H + yy) * W + xx] : 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float wr[CPL][27];                     // w: [Cout, 3, 3, 3] (OIHW): k = ci*9 + ky*3 + kx
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int k = 0; k < 27; ++k) wr[c][k] = __ldg(w + (int64_t)(lane * CPL + c) * 27 + k);
  __syncthreads();
  const int segs = (W + 31) / 32;
  for (int item = warp; item < ROWS * segs; item += nwarps) {
    const int ry = item / segs, x0 = (item % segs) * 32;
    if (y0 + ry >= H) break;             // items are row-major: every later item of this warp is out of range too
    float* orow = out + ((int64_t)n * H + y0 + ry) * W * Cout;
    int so[9];                           // offset of the staged input row behind window row r = ci*3 + ky
#pragma unroll
    for (int r = 0; r < 9; ++r) so[r] = ((r / 3) * RS + ry + r % 3) * Wp;
    float win[9][3];                     // [ci*3 + ky][kx]
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      win[r][1] = srow[so[r] + x0];      // padded column x0 - 1 + 1
      win[r][2] = srow[so[r] + x0 + 1];
    }
    const int xe = min(x0 + 32, W);
    for (int x = x0; x < xe; ++x) {
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        win[r][0] = win[r][1];
        win[r][1] = win[r][2];
        win[r][2] = srow[so[r] + x + 2];
      }
      float acc[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 27; ++k) a = fmaf(win[k / 3][k % 3], wr[c][k], a);   // same tap order as before
        acc[c] = a;
      }
      float* o = orow + (int64_t)x * Cout + lane * CPL;
      if (CPL == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else {
#pragma unroll
        for (int c = 0; c < CPL; ++c) o[c] = acc[c];
      }
    }
  }
}

template <int ROWS>
static int launch_stem_rows(const float* img, const float* w, float* out, int B, int H, int W, int Cout, cudaStream_t st) {
  const size_t smem = (size_t)3 * (ROWS + 2) * (W + 2) * sizeof(float);
  dim3 grid(ceil_div(H, ROWS), B);
  switch (Cout / 32) {
    case 1: stem_conv_kernel<1, ROWS><<<grid, 256, smem, st>>>(img, w, out, H, W); break;
    case 2: stem_conv_kernel<2, ROWS><<<grid, 256, smem, st>>>(img, w, out, H, W); break;
    case 4: stem_conv_kernel<4, ROWS><<<grid, 256, smem, st>>>(img, w, out, H, W); break;
    case 8: stem_conv_kernel<8, ROWS><<<grid, 256, smem, st>>>(img, w, out, H, W); break;
    default: set_error("stem conv: Cout %d unsupported (32, 64, 128, 256)", Cout); return GN_ERR_UNSUPPORTED;
  }
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

int launch_stem_conv(const float* img, const float* w, float* out, int B, int H, int W, int Cout, cudaStream_t st) {
  GN_REQUIRE(Cout % 32 == 0 && Cout <= 256, "stem conv: Cout %d unsupported", Cout);
  GN_REQUIRE((size_t)9 * (W + 2) * sizeof(float) <= 48 * 1024, "stem conv: image width %d too large", W);
  // 8 rows per block when the 10 staged rows fit the default 48 KB of shared memory (W <= 407), else one row per block.
  // GENIE_B200_STEM_ROWS=1 forces the one-row variant (A/B; results are bit-identical)
  const char* sr = getenv("GENIE_B200_STEM_ROWS");   // once per encode pass: re-read so that tests can flip it
  const bool one_row = sr && sr[0] == '1';
  if (!one_row && (size_t)3 * 10 * (W + 2) * sizeof(float) <= 48 * 1024)
    return launch_stem_rows<8>(img, w, out, B, H, W, Cout, st);
  return launch_stem_rows<1>(img, w, out, B, H, W, Cout, st);
}

// -------------------------------------------------------------------------------------
// GroupNorm(32 groups, eps 1e-6) statistics over NHWC fp32: stats[n][g] = {sum, sumsq} (double)
// improved_model.py:25-26,116 (nn.GroupNorm(32, C, eps=1e-6))
// -------------------------------------------------------------------------------------
// Deterministic two-stage reduction (no atomics: LFQ takes the SIGN of the encoder output, so run-to-run
// summation-order noise would flip bits of near-zero latents):
//   stage 1: block (chunk of pixels, image) -> partial[n][chunk][g] = {sum, sumsq} in a fixed order
//   stage 2: stats[n][g] = sum over chunks in index order (double)
// The block that finishes LAST for an image (atomic ticket) folds that image's partials into stats[n][g] in chunk-index
// order: which block is last varies from run to run, the summation order does not.  This replaces a separate
// finalize launch (350 launches of 26 us per 64-image encode+decode in round 1).
__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ x, double* __restrict__ partial, double* __restrict__ stats,
                  unsigned int* __restrict__ tickets, int HW, int C, int pix_per_block) {
  __shared__ float s_sum[256], s_sq[256];
  __shared__ bool s_last;
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
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[n], 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) tickets[n] = 0;          // ready for the next launch on this stream
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // 8 threads per group: thread j of a group adds the chunks c = j, j + 8, ... in index order (<= 8 independent loads
  // each), then the 8 sums are folded by a fixed shuffle tree
  const int g = threadIdx.x >> 3, j = threadIdx.x & 7, chunks = gridDim.x;
  double a = 0.0, b = 0.0;
#pragma unroll 8
  for (int c = j; c < chunks; c += 8) {
    const double2 p = __ldcg(reinterpret_cast<const double2*>(partial + (((int64_t)n * chunks + c) * 32 + g) * 2));
    a += p.x;
    b += p.y;
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o, 8);
    b += __shfl_down_sync(0xffffffffu, b, o, 8);
  }
  if (j == 0) {
    // mean and 1 / sqrt(var + eps) of the group in double, handed to the apply / head kernels as ONE float2 (they used
    // to redo this double-precision division and square root per element: FP64 runs at 1/64 rate on this part)
    const double cnt = (double)HW * (C / 32);
    const double mean = a / cnt;
    const double var = fmax(b / cnt - mean * mean, 0.0);
    reinterpret_cast<float2*>(stats)[(int64_t)n * 32 + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-6)));
  }
}
// stats buffer layout: [64 doubles = 128 uint tickets at a FIXED place, zeroed at allocation][B*64 doubles final
// stats][partials]; kStatsOff = offset of the final stats that the apply / head kernels read
constexpr int kStatsOff = 64;
static int gn_stats(const float* x, double* stats, int B, int HW, int C, cudaStream_t st) {
  GN_REQUIRE(C % 128 == 0 && 1024 % C == 0, "GroupNorm: C %d unsupported (128, 256, 512, 1024)", C);
  // <= 64 chunks per image: the last block's fold (below) is a short, latency-bound tail
  int ppb = 8;
  while (ceil_div(HW, ppb) > 64) ppb *= 2;
  const int chunks = ceil_div(HW, ppb);
  GN_REQUIRE(chunks <= 512, "GroupNorm: too many partial chunks");
  GN_REQUIRE(B <= 128, "GroupNorm: at most 128 images per pass");
  unsigned int* tickets = reinterpret_cast<unsigned int*>(stats);
  double* partial = stats + kStatsOff + (int64_t)B * 64;
  gn_partial_kernel<<<dim3(chunks, B), 256, 0, st>>>(x, partial, stats + kStatsOff, tickets, HW, C, ppb);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// y = swish(GN(x)) -> NHWC in the operand format of the next convolution (bf16 / fp16 / fp32)  improved_model.py:8-10,41-46
template <typename OutT>
__global__ void __launch_bounds__(256)
gn_apply_swish_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                      const float* __restrict__ beta, OutT* __restrict__ out, int HW, int C, int64_t total_vec) {
  // 32-bit index arithmetic (total_vec < 2^31, checked by the launcher) and a per-thread-constant channel quad: the
  // block size and the grid stride are multiples of vec_per_pix, so v / g / gamma / beta are hoisted out of the loop
  // (the first version spent its time in 64-bit div / mod per element: 41 % of the copy peak)
  const uint32_t vec_per_pix = C / 4, cg = C / 32;
  const float2* mr = reinterpret_cast<const float2*>(stats);       // {mean, rstd} per (image, group)
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const uint32_t v = tid % vec_per_pix, g = (v * 4) / cg;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v);
  const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + v);
  const uint32_t vec_per_img = (uint32_t)HW * vec_per_pix;
  for (uint32_t i = tid; i < (uint32_t)total_vec; i += stride) {
    const uint32_t n = i / vec_per_img;
    const float2 ms = __ldg(mr + n * 32 + g);
    const float mean = ms.x, rstd = ms.y;
    const float4 f = __ldcs(reinterpret_cast<const float4*>(x) + i);
    float y[4] = {(f.x - mean) * rstd * gm.x + bt.x, (f.y - mean) * rstd * gm.y + bt.y, (f.z - mean) * rstd * gm.z + bt.z,
                  (f.w - mean) * rstd * gm.w + bt.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = y[k] / (1.f + __expf(-y[k]));
    if constexpr (sizeof(OutT) == 4) {
      reinterpret_cast<float4*>(out)[i] = make_float4(y[0], y[1], y[2], y[3]);
    } else {
      uint2 p;
      p.x = pack_h2<OutT>(y[0], y[1]);
      p.y = pack_h2<OutT>(y[2], y[3]);
      reinterpret_cast<uint2*>(out)[i] = p;
    }
  }
}

int launch_gn_swish(const float* x, double* stats, const float* gamma, const float* beta, void* out, int o16, int B,
                    int HW, int C, cudaStream_t st) {
  GN_PROPAGATE(gn_stats(x, stats, B, HW, C, st));
  const int64_t total_vec = (int64_t)B * HW * (C / 4);
  GN_REQUIRE(total_vec < (1ll << 31) && 256 % (C / 4) == 0, "GroupNorm apply: pass too large or C %d unsupported", C);
  const int g2 = (int)std::min<int64_t>(ceil_div64(total_vec, 256), 148 * 16);
  if (o16 == 2)
    gn_apply_swish_kernel<f16><<<g2, 256, 0, st>>>(x, stats + kStatsOff, gamma, beta, static_cast<f16*>(out), HW, C, total_vec);
  else if (o16 == 1)
    gn_apply_swish_kernel<bf16><<<g2, 256, 0, st>>>(x, stats + kStatsOff, gamma, beta, static_cast<bf16*>(out), HW, C, total_vec);
  else
    gn_apply_swish_kernel<float><<<g2, 256, 0, st>>>(x, stats + kStatsOff, gamma, beta, static_cast<float*>(out), HW, C, total_vec);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// GroupNorm statistics + apply + swish in ONE persistent kernel (16-bit operand modes).
// The two-kernel form above reads the fp32 trunk twice from HBM (statistics pass, apply pass) and writes the operand:
// 10 bytes per element, 30 % of an encode + decode pass.  Here a chunk of an image is summed (phase 1) and normalised
// (phase 2) by the SAME block a short time apart, so the second read is served by L2: 6 bytes per element of HBM traffic.
//   * work item = (image n, chunk c) of `ppb` pixels (64 KB of fp32); block b takes items b, b + grid, ... in order
//   * phase 1(item): the block's partial {sum, sumsq} per group -> partial[n][c][g]; the block whose ticket completes
//     image n folds the image's partials IN CHUNK-INDEX ORDER (deterministic, as above) into {mean, rstd}, then
//     publishes ready[n] = epoch (release)
//   * phase 2(item) = the apply pass on the item's pixels, after an acquire spin on ready[n].  A block runs
//     phase 1(item k) BEFORE phase 2(item k - 1): the fold of an image overlaps the next item's loads, and the in-flight
//     footprint is 2 items per block (2 x 592 x 64 KB = 76 MB < L2).
//   * progress: every block reaches phase 1 of its item of image n without waiting on image n itself (only on earlier
//     images), so by induction over n every image completes - provided all blocks are co-resident: the grid is sized by
//     the occupancy calculator and launched with cudaLaunchCooperativeKernel, which enforces exactly that.
// The per-element arithmetic equals gn_apply_swish_kernel's; the statistics differ from the two-kernel form only by the
// chunking of the fixed-order sum, which is why the fp32 exact mode (token ids equal to the reference's) keeps that form.
// -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(256)
gn_fused_kernel(const float* __restrict__ x, double* __restrict__ partial, double* __restrict__ stats,
                unsigned int* __restrict__ tickets, unsigned int* __restrict__ ready, unsigned int epoch,
                const float* __restrict__ gamma, const float* __restrict__ beta, OutT* __restrict__ out, int HW, int C,
                int ppb, int chunks, int n_items) {
  __shared__ float s_sum[256], s_sq[256];
  __shared__ bool s_last;
  const int cg = C / 32;                 // channels per group (4, 8, 16, 32)
  const int vpp = C / 4;                 // float4 per pixel; 256 % vpp == 0 -> each thread owns one channel quad
  const int tid = threadIdx.x;
  const int v = tid % vpp, g_of_thread = (v * 4) / cg;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v);
  const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + v);

  auto phase1 = [&](int item) {
    const int n = item / chunks, c = item % chunks;
    const int p0 = c * ppb, p1 = min(p0 + ppb, HW);
    const float4* base = reinterpret_cast<const float4*>(x + (int64_t)n * HW * C) + (int64_t)p0 * vpp;
    const int total = (p1 - p0) * vpp;
    float sum = 0.f, sq = 0.f;
    for (int i = tid; i < total; i += 256) {
      const float4 f = base[i];
      sum += (f.x + f.y) + (f.z + f.w);
      sq += (f.x * f.x + f.y * f.y) + (f.z * f.z + f.w * f.w);
    }
    __syncthreads();                     // s_sum / s_sq / s_last of the previous item are no longer read
    s_sum[tid] = sum;
    s_sq[tid] = sq;
    __syncthreads();
    if (tid < 32) {
      const int g = tid;
      const int q0 = g * cg / 4, q1 = (g + 1) * cg / 4;   // channel quads of this group
      double a = 0.0, b2 = 0.0;
      for (int t0 = 0; t0 < 256; t0 += vpp)
        for (int q = q0; q < q1; ++q) { a += (double)s_sum[t0 + q]; b2 += (double)s_sq[t0 + q]; }
      double* o = partial + (((int64_t)n * chunks + c) * 32 + g) * 2;
      o[0] = a;
      o[1] = b2;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int t = atomicAdd(&tickets[n], 1u);
      s_last = (t == (unsigned int)chunks - 1);
      if (s_last) tickets[n] = 0;        // ready for the next launch on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int g = tid >> 3, j = tid & 7;
    double a = 0.0, b = 0.0;
#pragma unroll 8
    for (int cc = j; cc < chunks; cc += 8) {
      const double2 p = __ldcg(reinterpret_cast<const double2*>(partial + (((int64_t)n * chunks + cc) * 32 + g) * 2));
      a += p.x;
      b += p.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o, 8);
      b += __shfl_down_sync(0xffffffffu, b, o, 8);
    }
    if (j == 0) {
      const double cnt = (double)HW * (C / 32);
      const double mean = a / cnt;
      const double var = fmax(b / cnt - mean * mean, 0.0);
      reinterpret_cast<float2*>(stats)[(int64_t)n * 32 + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-6)));
    }
    __syncthreads();                     // the 32 statistics are written ...
    if (tid == 0) {
      __threadfence();                   // ... and visible device-wide before the flag
      st_release_u32(&ready[n], epoch);
    }
  };

  auto phase2 = [&](int item) {
    const int n = item / chunks, c = item % chunks;
    if (tid == 0) {
      while (ld_acquire_u32(&ready[n]) != epoch) __nanosleep(64);
    }
    __syncthreads();
    const float2 ms = __ldcg(reinterpret_cast<const float2*>(stats) + (int64_t)n * 32 + g_of_thread);
    const float mean = ms.x, rstd = ms.y;
    const int p0 = c * ppb, p1 = min(p0 + ppb, HW);
    const int64_t off = ((int64_t)n * HW + p0) * vpp;            // in float4 / 4-element vectors
    const float4* base = reinterpret_cast<const float4*>(x) + off;
    const int total = (p1 - p0) * vpp;
    for (int i = tid; i < total; i += 256) {
      const float4 f = __ldcs(base + i);
      float y[4] = {(f.x - mean) * rstd * gm.x + bt.x, (f.y - mean) * rstd * gm.y + bt.y, (f.z - mean) * rstd * gm.z + bt.z,
                    (f.w - mean) * rstd * gm.w + bt.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = y[k] / (1.f + __expf(-y[k]));
      uint2 p;
      p.x = pack_h2<OutT>(y[0], y[1]);
      p.y = pack_h2<OutT>(y[2], y[3]);
      reinterpret_cast<uint2*>(out)[off + i] = p;
    }
  };

  int prev = -1;
  for (int item = blockIdx.x;; item += gridDim.x) {
    if (item < n_items) phase1(item);
    if (prev >= 0) phase2(prev);
    if (item >= n_items) break;
    prev = item;
  }
}

constexpr int kGnFusedMaxChunks = 512;
// -> GN_OK and *done = true when the fused kernel ran; *done = false when this shape is left to the two-kernel form
template <typename OutT>
static int launch_gn_fused_t(const float* x, double* stats, unsigned int* ready, unsigned int epoch, const float* gamma,
                             const float* beta, OutT* out, int B, int HW, int C, cudaStream_t st, bool* done) {
  *done = false;
  if (!(C % 128 == 0 && 1024 % C == 0) || B > 128) return GN_OK;
  int ppb = 65536 / (C * 4);             // 64 KB of fp32 per item
  if (ppb < 8) ppb = 8;
  while (ceil_div(HW, ppb) > kGnFusedMaxChunks) ppb *= 2;
  const int chunks = ceil_div(HW, ppb);
  const int n_items = B * chunks;
  static int blocks_per_sm_dev[kMaxDevices] = {};
  int& bps = blocks_per_sm_dev[current_device()];
  if (!bps) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gn_fused_kernel<OutT>, 256, 0) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return GN_OK;
    }
    bps = n < 4 ? n : 4;                 // 4 x 148 blocks x 2 items x 64 KB = 76 MB in flight: inside L2
  }
  int grid = bps * device_sm_count();
  if (grid > n_items) grid = n_items;
  // the progress argument needs at most two items of one block inside one image: chunks <= 2 * grid (B200: grid = 592
  // or one item per block); anything else takes the two-kernel form
  if (chunks > 2 * grid) return GN_OK;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(stats);
  double* fin = stats + kStatsOff;
  double* partial = stats + kStatsOff + (int64_t)B * 64;
  int ppb_ = ppb, chunks_ = chunks, items_ = n_items, HW_ = HW, C_ = C;
  void* args[] = {(void*)&x, (void*)&partial, (void*)&fin, (void*)&tickets, (void*)&ready, (void*)&epoch, (void*)&gamma,
                  (void*)&beta, (void*)&out, (void*)&HW_, (void*)&C_, (void*)&ppb_, (void*)&chunks_, (void*)&items_};
  GN_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)gn_fused_kernel<OutT>, dim3(grid), dim3(256), args, 0, st));
  ++g_launch_count;
  *done = true;
  return GN_OK;
}

int launch_gn_swish_fused(const float* x, double* stats, unsigned int* ready, unsigned int epoch, const float* gamma,
                          const float* beta, void* out, int o16, int B, int HW, int C, cudaStream_t st) {
  bool done = false;
  if (o16 == 2)
    GN_PROPAGATE(launch_gn_fused_t<f16>(x, stats, ready, epoch, gamma, beta, static_cast<f16*>(out), B, HW, C, st, &done));
  else if (o16 == 1)
    GN_PROPAGATE(launch_gn_fused_t<bf16>(x, stats, ready, epoch, gamma, beta, static_cast<bf16*>(out), B, HW, C, st, &done));
  if (done) return GN_OK;
  return launch_gn_swish(x, stats, gamma, beta, out, o16, B, HW, C, st);
}

// -------------------------------------------------------------------------------------
// depth-to-space (DCR, block 2): [B,H,W,4C'] -> [B,2H,2W,C'], channel = (b1*2 + b2)*C' + c'
// improved_model.py:185-237
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
depth_to_space_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Cp, int64_t total_vec) {
  const uint32_t vec_c = Cp / 4;                    // 32-bit index arithmetic (total_vec < 2^31, checked by the launcher)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)total_vec; i += gridDim.x * blockDim.x) {
    const uint32_t v = i % vec_c;
    uint32_t r = i / vec_c;
    const uint32_t ox = r % (2 * W); r /= (2 * W);
    const uint32_t oy = r % (2 * H);
    const uint32_t n = r / (2 * H);
    const uint32_t b1 = oy & 1, b2 = ox & 1, h = oy >> 1, w = ox >> 1;
    const float4 f = __ldcs(reinterpret_cast<const float4*>(in) + ((n * H + h) * W + w) * (4 * vec_c) + (b1 * 2 + b2) * vec_c + v);
    reinterpret_cast<float4*>(out)[i] = f;
  }
}
int launch_depth_to_space(const float* in, float* out, int B, int H, int W, int Cp, cudaStream_t st) {
  const int64_t total_vec = (int64_t)B * 4 * H * W * (Cp / 4);
  GN_REQUIRE(total_vec < (1ll << 31), "depth-to-space: pass too large");
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
  const float2* mr = reinterpret_cast<const float2*>(stats);       // {mean, rstd} per (image, group)
  float acc[32];
#pragma unroll
  for (int z = 0; z < 32; ++z) acc[z] = 0.f;
  for (int c = lane; c < C; c += 32) {
    const int g = c / cg;
    const float2 ms = __ldg(mr + (int64_t)n * 32 + g);
    const float mean = ms.x, rstd = ms.y;
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
  vq_head_kernel<<<ceil_div(n_pix, 8), 256, 0, st>>>(x, stats + kStatsOff, gamma, beta, w, bias, ids, z_out, HW, C, Z, n_pix);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// decoder stem: ids -> +-1 latents (Z channels; little_endian: bit c <-> channel c, i.e. get_codebook_entry
// followed by visualize.py:115 `.flip(1)`; big endian: bit (Z-1-c) <-> channel c) -> conv 3x3 (Z -> Cout) + bias
// lookup_free_quantize.py:181-194, improved_model.py:135-137,164.   w: [Cout, Z, 3, 3].  One warp per pixel.
// -------------------------------------------------------------------------------------
// Block = 64 pixels x 32 output channels: the channel tile's weights (32 x Z x 9 fp32, 20 KB for Z = 18) are staged in
// shared memory once and reused by the 64 pixels; thread (p = tid / 4, q = tid % 4) owns pixel p and 8 channels.  (The
// first version - one warp per pixel, every lane streaming its 16 channels' 162 taps from global memory - took 334 us
// per 8 images for 170 MFMA of work.)
constexpr int VT_PIX = 64, VT_CO = 32;
__global__ void __launch_bounds__(256)
vq_tail_kernel(const int32_t* __restrict__ ids, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ out, int H, int W, int Z, int Cout, int little_endian, int n_pix) {
  extern __shared__ float vt_w[];                 // [VT_CO][Z * 9]
  const int co0 = blockIdx.y * VT_CO;
  const int zk = Z * 9;
  for (int i = threadIdx.x; i < VT_CO * zk; i += blockDim.x) vt_w[i] = __ldg(w + (int64_t)co0 * zk + i);
  __syncthreads();
  const int pix = blockIdx.x * VT_PIX + (threadIdx.x >> 2), q = threadIdx.x & 3;
  if (pix >= n_pix) return;
  const int n = pix / (H * W), rem = pix % (H * W), y = rem / W, x = rem % W;
  int nb[9];
  bool ok[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    ok[t] = yy >= 0 && yy < H && xx >= 0 && xx < W;
    nb[t] = ok[t] ? ids[((int64_t)n * H + yy) * W + xx] : 0;
  }
  float acc[VT_CO / 4];
#pragma unroll
  for (int j = 0; j < VT_CO / 4; ++j) acc[j] = bias[co0 + q * (VT_CO / 4) + j];
  for (int c = 0; c < Z; ++c) {                   // same (channel, tap) summation order as the first version
    const int bit = little_endian ? c : (Z - 1 - c);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!ok[t]) continue;
      const bool pos = (nb[t] >> bit) & 1;
#pragma unroll
      for (int j = 0; j < VT_CO / 4; ++j) {
        const float wv = vt_w[(q * (VT_CO / 4) + j) * zk + c * 9 + t];
        acc[j] += pos ? wv : -wv;
      }
    }
  }
  float* o = out + (int64_t)pix * Cout + co0 + q * (VT_CO / 4);
  *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
}
int launch_vq_tail(const int32_t* ids, const float* w, const float* bias, float* out, int B, int H, int W, int Z, int Cout,
                   int little_endian, cudaStream_t st) {
  GN_REQUIRE(Cout % VT_CO == 0, "LFQ tail: Cout %d must be a multiple of %d", Cout, VT_CO);
  const int n_pix = B * H * W;
  const size_t smem = (size_t)VT_CO * Z * 9 * sizeof(float);
  GN_REQUIRE(smem <= 48 * 1024, "LFQ tail: z_channels %d too large", Z);
  vq_tail_kernel<<<dim3(ceil_div(n_pix, VT_PIX), Cout / VT_CO), 256, smem, st>>>(ids, w, bias, out, H, W, Z, Cout,
                                                                                  little_endian, n_pix);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// decoder output conv: 3x3, C -> 3, + bias, input = swish(GN(x)) as bf16 NHWC; writes fp32 NCHW and/or the uint8
// image ((v + 1) * 127.5 clamped to [0, 255], truncated: visualize.py:84-92).   w: [3, C, 3, 3]
// -------------------------------------------------------------------------------------
// Block = 64 consecutive pixels of one output row, 4 warps.  The 3 x 66 input pixels (C bf16 each) are staged in shared
// memory with a padded pixel pitch (C*2 + 16 B: conflict-free 16-byte reads at a one-pixel lane stride), the weights as
// [tap][c][4] fp32.  Warp s owns the channel slice [s*C/4, (s+1)*C/4), lane l the pixels l and l + 32: every weight read
// is a warp-uniform broadcast and is used for two pixels; the 4 slice partials are reduced through shared memory in a
// fixed order.  (The first version - one warp per pixel, weights gathered from global memory with a 9-float stride,
// three warp reductions per pixel - ran 1.02 ms per 8 images against ~25 us of HBM time for its input.)
constexpr int OC_PIX = 64;
template <typename InT>
__global__ void __launch_bounds__(128)
out_conv_kernel(const InT* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ out_f32, uint8_t* __restrict__ out_u8, int H, int W, int C) {
  extern __shared__ __align__(16) uint8_t oc_smem[];
  constexpr int ES = (int)sizeof(InT);                        // 2 (bf16 / fp16) or 4 (fp32 exact mode)
  constexpr int EPV = 16 / ES;                                // elements per 16-byte vector
  const int pitch = C * ES + 16;                              // bytes per staged pixel
  uint8_t* sin = oc_smem;                                     // [3][OC_PIX + 2][pitch]
  float* sw = reinterpret_cast<float*>(oc_smem + 3 * (OC_PIX + 2) * pitch);   // [9][C][4]
  float* sred = sw + 9 * C * 4;                               // [4 slices][OC_PIX][3]
  const int n = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * OC_PIX;
  const int tid = threadIdx.x, lane = tid & 31, slice = tid >> 5;
  // weights: w[co][c][t] -> sw[t][c][co]
  for (int i = tid; i < 9 * C; i += blockDim.x) {
    const int t = i / C, c = i % C;
    float4 v;
    v.x = __ldg(w + ((int64_t)0 * C + c) * 9 + t);
    v.y = __ldg(w + ((int64_t)1 * C + c) * 9 + t);
    v.z = __ldg(w + ((int64_t)2 * C + c) * 9 + t);
    v.w = 0.f;
    reinterpret_cast<float4*>(sw)[i] = v;
  }
  // input patch, 16-byte vectors, zero outside the image
  const int vec_per_pix = C / EPV;
  for (int i = tid; i < 3 * (OC_PIX + 2) * vec_per_pix; i += blockDim.x) {
    const int v = i % vec_per_pix, p = (i / vec_per_pix) % (OC_PIX + 2), r = i / (vec_per_pix * (OC_PIX + 2));
    const int yy = y + r - 1, xx = x0 + p - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      val = *reinterpret_cast<const uint4*>(a + (((int64_t)n * H + yy) * W + xx) * C + v * EPV);
    *reinterpret_cast<uint4*>(sin + (r * (OC_PIX + 2) + p) * pitch + v * 16) = val;
  }
  __syncthreads();
  const int cps = C / 4;                                      // channels per slice
  float acc[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  for (int t = 0; t < 9; ++t) {
    const int r = t / 3, dx = t % 3;
    const uint8_t* p0 = sin + (r * (OC_PIX + 2) + lane + dx) * pitch + slice * cps * ES;
    const uint8_t* p1 = p0 + 32 * pitch;
    const float4* wt = reinterpret_cast<const float4*>(sw) + t * C + slice * cps;
    for (int c8 = 0; c8 < cps; c8 += EPV) {
      const uint4 u0 = *reinterpret_cast<const uint4*>(p0 + c8 * ES);
      const uint4 u1 = *reinterpret_cast<const uint4*>(p1 + c8 * ES);
      const uint32_t* h0 = reinterpret_cast<const uint32_t*>(&u0);
      const uint32_t* h1 = reinterpret_cast<const uint32_t*>(&u1);
#pragma unroll
      for (int k = 0; k < EPV / 2; ++k) {
        float2 v0, v1;
        if constexpr (ES == 2) {
          v0 = unpack_h2<InT>(h0[k]);
          v1 = unpack_h2<InT>(h1[k]);
        } else {
          v0 = make_float2(__uint_as_float(h0[2 * k]), __uint_as_float(h0[2 * k + 1]));
          v1 = make_float2(__uint_as_float(h1[2 * k]), __uint_as_float(h1[2 * k + 1]));
        }
        const float4 wa = wt[c8 + 2 * k], wb = wt[c8 + 2 * k + 1];
        acc[0][0] = fmaf(v0.x, wa.x, acc[0][0]); acc[0][1] = fmaf(v0.x, wa.y, acc[0][1]); acc[0][2] = fmaf(v0.x, wa.z, acc[0][2]);
        acc[0][0] = fmaf(v0.y, wb.x, acc[0][0]); acc[0][1] = fmaf(v0.y, wb.y, acc[0][1]); acc[0][2] = fmaf(v0.y, wb.z, acc[0][2]);
        acc[1][0] = fmaf(v1.x, wa.x, acc[1][0]); acc[1][1] = fmaf(v1.x, wa.y, acc[1][1]); acc[1][2] = fmaf(v1.x, wa.z, acc[1][2]);
        acc[1][0] = fmaf(v1.y, wb.x, acc[1][0]); acc[1][1] = fmaf(v1.y, wb.y, acc[1][1]); acc[1][2] = fmaf(v1.y, wb.z, acc[1][2]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int co = 0; co < 3; ++co) sred[(slice * OC_PIX + lane + 32 * q) * 3 + co] = acc[q][co];
  __syncthreads();
  for (int i = tid; i < OC_PIX * 3; i += blockDim.x) {
    const int co = i / OC_PIX, p = i % OC_PIX;   // consecutive threads -> consecutive x of one output plane
    const int x = x0 + p;
    if (x >= W) continue;
    float v = bias[co];
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) v += sred[(sl * OC_PIX + p) * 3 + co];
    const int64_t o = (((int64_t)n * 3 + co) * H + y) * W + x;
    if (out_f32) out_f32[o] = v;
    if (out_u8) out_u8[o] = (uint8_t)fminf(fmaxf((v + 1.f) * 127.5f, 0.f), 255.f);
  }
}
template <typename InT>
static int launch_out_conv_t(const InT* a, const float* w, const float* bias, float* out_f32, uint8_t* out_u8, int B, int H,
                             int W, int C, cudaStream_t st) {
  const size_t smem = (size_t)3 * (OC_PIX + 2) * (C * sizeof(InT) + 16) + (size_t)9 * C * 16 + (size_t)4 * OC_PIX * 3 * 4;
  GN_REQUIRE(smem <= 227 * 1024, "output conv: C %d does not fit shared memory in this precision", C);
  static DevSmemOptIn optin;
  GN_CUDA_CHECK(ensure_smem_optin(optin, out_conv_kernel<InT>, (int)smem));
  dim3 grid(ceil_div(W, OC_PIX), H, B);
  out_conv_kernel<InT><<<grid, 128, smem, st>>>(a, w, bias, out_f32, out_u8, H, W, C);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}
int launch_out_conv(const void* a, int o16, const float* w, const float* bias, float* out_f32, uint8_t* out_u8, int B, int H,
                    int W, int C, cudaStream_t st) {
  GN_REQUIRE(C % 32 == 0 && C <= 512, "output conv: C %d unsupported (multiple of 32, <= 512)", C);
  if (o16 == 2) return launch_out_conv_t<f16>(static_cast<const f16*>(a), w, bias, out_f32, out_u8, B, H, W, C, st);
  if (o16 == 1) return launch_out_conv_t<bf16>(static_cast<const bf16*>(a), w, bias, out_f32, out_u8, B, H, W, C, st);
  return launch_out_conv_t<float>(static_cast<const float*>(a), w, bias, out_f32, out_u8, B, H, W, C, st);
}

// -------------------------------------------------------------------------------------
// decoder output conv on the warp-level tensor cores (16-bit operand modes): 3x3, C -> 3, + bias, as an implicit GEMM
// with mma.sync m16n8k16 (fp32 accumulate): M = 16 consecutive pixels of one image row, N = 8 (3 used), K = 9 taps x C.
// The CUDA-core kernel above spends 7x its FMA floor (1.55 ms per 32 images, 6.6 % of an encode + decode pass); here the
// arithmetic is 72 MMAs per 16 pixels and the kernel is bound by streaming the activation through L1.
//   * A fragments come straight from global memory (L1-cached; every line is reused by up to 9 taps): K is PERMUTED so
//     that lane (r = lane / 4, j = lane % 4) reads 32 contiguous bytes of pixel r (and of pixel r + 8) per 64-channel
//     block - the 4 lanes of a pixel cover one full 128-byte line - and serves 4 k-steps from them: in k-step s of block
//     ss, k-slots {2j, 2j+1} are channels 64 ss + 16 j + 4 s + {0, 1} and k-slots {2j+8, 2j+9} are ... + {2, 3}.
//   * B fragments (the weights in the operand format, same permutation, n >= 3 zero) are laid out once per weight upload
//     as wf[(tap * C/64 + ss) * 4 + s][lane] = {b0, b1} (out_conv_pack_kernel) and staged in shared memory per block.
//   * block = 8 warps = 8 consecutive rows x 16 pixels, so the 3 input rows of a warp are shared through L1 with its
//     neighbours; zero padding = predicated loads.
// The fp32 exact mode keeps the CUDA-core kernel (fp32 weights and activations).
// -------------------------------------------------------------------------------------
template <typename H>
__global__ void out_conv_pack_kernel(const float* __restrict__ w /*[3][C][3][3]*/, uint2* __restrict__ wf, int C) {
  const int CB = C / 64;
  const int total = 9 * CB * 4 * 32;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int lane = idx & 31, ks = idx >> 5;
  const int s = ks & 3, ss = (ks >> 2) % CB, t = ks / (4 * CB);
  const int j = lane & 3, n = lane >> 2;
  const int c0 = 64 * ss + 16 * j + 4 * s;
  uint2 v = make_uint2(0u, 0u);
  if (n < 3) {
    const float* wn = w + (int64_t)n * C * 9 + t;
    v.x = pack_h2<H>(__ldg(wn + (int64_t)(c0 + 0) * 9), __ldg(wn + (int64_t)(c0 + 1) * 9));
    v.y = pack_h2<H>(__ldg(wn + (int64_t)(c0 + 2) * 9), __ldg(wn + (int64_t)(c0 + 3) * 9));
  }
  wf[idx] = v;
}

template <typename H>
__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
  if constexpr (H16<H>::UMMA_FMT == 1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

// 32 contiguous bytes through the non-coherent path as ONE 256-bit load (LDG.E.256, sm_100): the 4 lanes of a pixel then
// cover its full 128-byte line with one instruction.  Motivation: with two 128-bit loads ncu shows l1tex throughput at
// 95 % (a warp-wide 128-bit load of this fragment layout touches 8 lines for 512 bytes).  MEASURED AND LEFT OFF
// (GENIE_B200_OUT_CONV_L256=1; frames bit-identical): decode 2441 / 2400 img/s vs 2517 / 2456 with 128-bit loads, same
// process, interleaved (scripts/out_conv_l256_check.py, profiles/r02b_out_conv_l256_ab.json).
__device__ __forceinline__ void ldg256(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

constexpr int OM_ROWS = 8, OM_PIX = 16;   // block tile: 8 rows x 16 pixels, one warp per row
template <typename H, bool L256>
__global__ void __launch_bounds__(32 * OM_ROWS)
out_conv_mma_kernel(const H* __restrict__ a, const uint2* __restrict__ wf, const float* __restrict__ bias,
                    float* __restrict__ out_f32, uint8_t* __restrict__ out_u8, int Hh, int W, int C) {
  extern __shared__ __align__(16) uint2 om_wf[];               // [9 * C/16][32]
  const int CB = C / 64;
  const int nfrag = 9 * CB * 4 * 32;
  for (int i = threadIdx.x; i < nfrag / 2; i += blockDim.x)
    reinterpret_cast<uint4*>(om_wf)[i] = __ldg(reinterpret_cast<const uint4*>(wf) + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.z, y = blockIdx.y * OM_ROWS + warp, x0 = blockIdx.x * OM_PIX;
  if (y >= Hh) return;
  const int r = lane >> 2, j = lane & 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1;
    const int xa = x0 + r + t % 3 - 1, xb = xa + 8;
    const bool rowok = yy >= 0 && yy < Hh;
    const bool va = rowok && xa >= 0 && xa < W, vb = rowok && xb >= 0 && xb < W;
    const H* pa = a + (((int64_t)n * Hh + yy) * W + xa) * C + 16 * j;
    const H* pb = pa + (int64_t)8 * C;
    for (int ss = 0; ss < CB; ++ss) {
      uint4 qa0 = make_uint4(0, 0, 0, 0), qa1 = qa0, qb0 = qa0, qb1 = qa0;
      if constexpr (L256) {
        if (va) ldg256(pa + 64 * ss, qa0, qa1);
        if (vb) ldg256(pb + 64 * ss, qb0, qb1);
      } else {
        if (va) {
          qa0 = __ldg(reinterpret_cast<const uint4*>(pa + 64 * ss));
          qa1 = __ldg(reinterpret_cast<const uint4*>(pa + 64 * ss + 8));
        }
        if (vb) {
          qb0 = __ldg(reinterpret_cast<const uint4*>(pb + 64 * ss));
          qb1 = __ldg(reinterpret_cast<const uint4*>(pb + 64 * ss + 8));
        }
      }
      const uint2* bf = om_wf + ((t * CB + ss) * 4) * 32 + lane;
      const uint2 b0 = bf[0], b1 = bf[32], b2 = bf[64], b3 = bf[96];
      mma_m16n8k16<H>(acc, qa0.x, qb0.x, qa0.y, qb0.y, b0.x, b0.y);   // s = 0: channels +0..3
      mma_m16n8k16<H>(acc, qa0.z, qb0.z, qa0.w, qb0.w, b1.x, b1.y);   // s = 1: +4..7
      mma_m16n8k16<H>(acc, qa1.x, qb1.x, qa1.y, qb1.y, b2.x, b2.y);   // s = 2: +8..11
      mma_m16n8k16<H>(acc, qa1.z, qb1.z, qa1.w, qb1.w, b3.x, b3.y);   // s = 3: +12..15
    }
  }
  // accumulator layout: acc[0], acc[1] = pixel r, outputs 2j, 2j+1; acc[2], acc[3] = pixel r + 8
  auto put = [&](int co, int x, float v) {
    if (x >= W) return;
    v += __ldg(bias + co);
    const int64_t o = (((int64_t)n * 3 + co) * Hh + y) * W + x;
    if (out_f32) out_f32[o] = v;
    if (out_u8) out_u8[o] = (uint8_t)fminf(fmaxf((v + 1.f) * 127.5f, 0.f), 255.f);
  };
  if (j == 0) {
    put(0, x0 + r, acc[0]); put(1, x0 + r, acc[1]);
    put(0, x0 + r + 8, acc[2]); put(1, x0 + r + 8, acc[3]);
  } else if (j == 1) {
    put(2, x0 + r, acc[0]);
    put(2, x0 + r + 8, acc[2]);
  }
}

int launch_out_conv_pack(const float* w, void* wf, int o16, int C, cudaStream_t st) {
  GN_REQUIRE(o16 != 0 && C % 64 == 0, "output conv (tensor path): 16-bit operands and C %% 64 == 0");
  const int total = 9 * (C / 64) * 4 * 32;
  if (o16 == 2) out_conv_pack_kernel<f16><<<ceil_div(total, 256), 256, 0, st>>>(w, static_cast<uint2*>(wf), C);
  else out_conv_pack_kernel<bf16><<<ceil_div(total, 256), 256, 0, st>>>(w, static_cast<uint2*>(wf), C);
  GN_CUDA_CHECK(cudaGetLastError());
  return GN_OK;
}

template <typename H, bool L256>
static int launch_om(const void* a, const void* wf, const float* bias, float* out_f32, uint8_t* out_u8, dim3 grid, size_t smem,
                     int Hh, int W, int C, cudaStream_t st) {
  static DevSmemOptIn optin;             // one per kernel instantiation
  GN_CUDA_CHECK(ensure_smem_optin(optin, out_conv_mma_kernel<H, L256>, (int)smem));
  out_conv_mma_kernel<H, L256><<<grid, 32 * OM_ROWS, smem, st>>>(static_cast<const H*>(a), static_cast<const uint2*>(wf), bias,
                                                                out_f32, out_u8, Hh, W, C);
  return GN_OK;
}

int launch_out_conv_mma(const void* a, int o16, const void* wf, const float* bias, float* out_f32, uint8_t* out_u8, int B,
                        int H, int W, int C, cudaStream_t st) {
  GN_REQUIRE(o16 != 0 && C % 64 == 0 && C <= 512, "output conv (tensor path): 16-bit operands, C %% 64 == 0, C <= 512");
  const size_t smem = (size_t)9 * (C / 16) * 32 * sizeof(uint2);      // 18 KB for C = 128
  GN_REQUIRE(B <= 65535, "output conv: too many images per pass");
  dim3 grid(ceil_div(W, OM_PIX), ceil_div(H, OM_ROWS), B);
  // GENIE_B200_OUT_CONV_L256=1: 256-bit activation loads (needs 32-byte aligned pixel rows: C % 16 == 0 holds, base checked)
  const char* e256 = getenv("GENIE_B200_OUT_CONV_L256");
  const bool l256 = e256 && e256[0] == '1' && reinterpret_cast<uintptr_t>(a) % 32 == 0;
  if (o16 == 2) {
    if (l256) GN_PROPAGATE((launch_om<f16, true>(a, wf, bias, out_f32, out_u8, grid, smem, H, W, C, st)));
    else GN_PROPAGATE((launch_om<f16, false>(a, wf, bias, out_f32, out_u8, grid, smem, H, W, C, st)));
  } else {
    if (l256) GN_PROPAGATE((launch_om<bf16, true>(a, wf, bias, out_f32, out_u8, grid, smem, H, W, C, st)));
    else GN_PROPAGATE((launch_om<bf16, false>(a, wf, bias, out_f32, out_u8, grid, smem, H, W, C, st)));
  }
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// conv weight repack: PyTorch [Cout, Cin, kh, kw] fp32 -> [Cout, kh*kw, Cin] bf16 (tap-major K for the implicit GEMM)
template <typename OutT>
__global__ void repack_conv_w_kernel(const float* __restrict__ w, OutT* __restrict__ out, int Cout, int Cin, int taps) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int t = (int)((i / Cin) % taps);
    const int co = (int)(i / ((int64_t)Cin * taps));
    out[i] = from_f32<OutT>(w[((int64_t)co * Cin + ci) * taps + t]);
  }
}
int launch_repack_conv_w(const float* w, void* out, int o16, int Cout, int Cin, int taps, cudaStream_t st) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), 4096);
  if (o16 == 2) repack_conv_w_kernel<f16><<<grid, 256, 0, st>>>(w, static_cast<f16*>(out), Cout, Cin, taps);
  else if (o16 == 1) repack_conv_w_kernel<bf16><<<grid, 256, 0, st>>>(w, static_cast<bf16*>(out), Cout, Cin, taps);
  else repack_conv_w_kernel<float><<<grid, 256, 0, st>>>(w, static_cast<float*>(out), Cout, Cin, taps);
  GN_CUDA_CHECK(cudaGetLastError());
  return GN_OK;
}

}  // namespace gn
